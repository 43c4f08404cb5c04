"""CPU oracle for the `vilma fit` hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's algorithm (jeffspence/vilma v0.0.16:
``variational_inference.py``, ``numerics.py``, ``matrix_structures.py``), written
from the maths in SURVEY.md Appendix A, each function citing the reference
file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline -- never as part of the product path.  ``vilma_b200`` does
not import it and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this oracle against
(a) the reference's own committed goldens (``copy_vilma_run.*``,
``example/*estimates.tsv`` / ``.npz``, packed in ``tests/golden/cli_*.npz``) and
(b) trajectories produced by the unmodified reference run in the build container
(``tests/golden/make_golden.py`` -> ``vischeme_*.npz``, ``syn_*.npz``).
"""
from . import numerics_np, ld_np, vi_np  # noqa: F401
