"""CPU oracle for the Grad DFT per-SCF-iteration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``graddft_b200/`` (the product) may import this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker / CPU comparator, never
as the thing measured as the product.

What it is: a float64 torch-CPU restatement of the reference's in-tree arithmetic
(``/root/reference/grad_dft/molecule.py``, ``functional.py``, ``popular_functionals.py``,
``train.py``), function by function, each citing the file:line it follows.  torch-CPU autograd over
the same expressions gives the reference VJPs (the analogue of ``jax.value_and_grad``).

Pinning: the reference needs jax/flax/pyscf, none of which exist in the build container or on the
GPU box.  ``tests/golden/make_golden.py`` therefore executes the reference's OWN source files
(imported from /root/reference, unmodified) on top of a small torch-backed stand-in for the handful
of ``jax.numpy`` symbols they use, and commits the resulting input/output vectors under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this restatement against those
vectors, so the oracle is pinned to the reference's own formulas, evaluated in IEEE float64 by
torch instead of XLA (same einsum strings, same pointwise expressions, same where/clip guards).
"""

from .reference_math import *  # noqa: F401,F403
