"""Command line front end: `vilma fit ...` and `vilma sim ...` (drop-in for vilma.frontend).

The reference's other sub-commands (make_ld_schema, check_ld_schema) are pre-processing /
diagnostics outside the fitting hot path and are not part of this package.
"""
import logging
import sys
from argparse import ArgumentParser

from . import REFERENCE_VERSION, VERSION
from .sim import args as sim_args
from .sim import main as sim
from .vi_options import args as fit_args
from .vi_options import main as fit

COMMANDS = {'fit': {'cmd': fit, 'parser': fit_args}, 'sim': {'cmd': sim, 'parser': sim_args}}
OUT_OF_SCOPE = ('make_ld_schema', 'check_ld_schema')


def main(argv=None):
    parser = ArgumentParser(
        description='vilma_b200 v%s: B200-native `vilma fit` (interface of vilma v%s): '
                    'variational inference of variant effect sizes from GWAS summary data.'
                    % (VERSION, REFERENCE_VERSION),
        usage='vilma <command> <options>')
    subparsers = parser.add_subparsers(title='Commands', dest='command')
    for cmd in COMMANDS:
        cmd_parser = COMMANDS[cmd]['parser'](subparsers)
        cmd_parser.add_argument('--logfile', required=False, type=str, default='',
                                help='File to store information about the run. To print to '
                                     'stdout use "-". Defaults to no logging.')
        cmd_parser.add_argument('--verbose', dest='verbose', action='store_true',
                                help='Log all information (as opposed to just warnings)')
    argv = sys.argv[1:] if argv is None else list(argv)
    if argv and argv[0] in OUT_OF_SCOPE:
        parser.error('%s is not part of vilma_b200 (only the `fit` path and `sim` are); use the '
                     'reference vilma for it' % argv[0])
    args = parser.parse_args(argv)
    try:
        func = COMMANDS[args.command]['cmd']
    except KeyError:
        parser.print_help()
        return 0
    level = 10 if args.verbose else 30
    if args.logfile == '-':
        logging.basicConfig(level=level)
    elif args.logfile:
        logging.basicConfig(filename=args.logfile, level=level)
    func(args)
    return 0


if __name__ == '__main__':
    sys.exit(main())
