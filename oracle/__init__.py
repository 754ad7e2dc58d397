"""CPU oracle for the NeuralQuantum.jl variational-Monte-Carlo hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy (float64 / complex128) restatement of
the reference's algorithm, function by function, each citing the reference file:line it
follows (paths relative to the reference checkout).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product (``neuralquantum.jl_b200/``) never does and has no CPU fallback.

Pinning status:
  * Julia is not installed in this image, so the reference itself cannot run here and it
    ships no golden vectors.  Machines, operators and the Liouvillian are pinned the way the
    reference's own active tests pin them (tests/test_oracle_*.py re-implement
    test/Machines/test_grad.jl, test/Operators/operators.jl, test/Operators/ising.jl with
    independent dense-matrix / finite-difference evaluation).
  * Local estimators, force, S, the solve and the sampler's accept rule have no active
    reference test ("parity unpinned" by the reference); they are pinned here against dense
    linear algebra ((H psi)/psi, (L rho)/rho, full-space sums, numpy.linalg.solve).
"""
from . import hilbert, operators, machines, estimators, sampler, sr, stats  # noqa: F401
