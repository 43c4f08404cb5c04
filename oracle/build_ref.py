"""Recipe for `oracle/_ref/`: the UNMODIFIED reference package, installed for use as the CPU baseline.

Test / benchmark infrastructure only (see oracle/__init__.py).  The reference is pure Python
(+ numba JIT), so "building" it is installing its package: this script pip-installs `/root/reference`
(from a scratch copy -- the source tree is read-only and setuptools writes an egg-info next to
setup.py) into `oracle/_ref/` with `--no-deps` (its pins -- numba==0.55.2, plinkio, h5py -- are not
in this image; numba 0.65 / numpy 2.3 / pandas 3 are, and run it with the import-time shims of
`oracle/ref_loader.py`).  No reference source is edited or committed: `oracle/_ref/` is git-ignored
and only travels to the GPU box with the gpurun snapshot, like the built `.so`.

    python oracle/build_ref.py            # in the build container (needs /root/reference)

`__graft_entry__.build()` runs it when /root/reference is present; on the GPU box the
installed copy is used as is.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SRC = '/root/reference'


def build(force=False):
    """Install the reference into oracle/_ref.  Returns DEST, or None when /root/reference is absent."""
    marker = os.path.join(DEST, 'vilma', 'variational_inference.py')
    if os.path.exists(marker) and not force:
        return DEST
    if not os.path.isdir(SRC):
        return DEST if os.path.exists(marker) else None
    tmp = tempfile.mkdtemp(prefix='vilma_ref_')
    try:
        work = os.path.join(tmp, 'reference')
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns('.git', 'example', 'tests'))
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, '-m', 'pip', 'install', '--quiet', '--no-index', '--no-build-isolation',
               '--no-deps', '--find-links', '/opt/wheelhouse', '--target', DEST, work]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not os.path.exists(marker):
            # same files, without pip: the package is a plain directory of .py modules
            sys.stderr.write('pip install of the reference failed (%s); copying src/vilma instead\n'
                             % (res.stderr.strip().splitlines() or ['?'])[-1])
            os.makedirs(DEST, exist_ok=True)
            shutil.copytree(os.path.join(SRC, 'src', 'vilma'), os.path.join(DEST, 'vilma'),
                            dirs_exist_ok=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return DEST if os.path.exists(marker) else None


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
