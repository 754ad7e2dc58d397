"""The C restatement (oracle/cref.c, the CPU baseline of bench.py) agrees with the NumPy oracle."""
import numpy as np
import pytest

from oracle import cref, estimators as E, machines as M
from oracle.models import lindblad_ising_1d, random_states


@pytest.mark.parametrize("act,fock", [(M.SOFTPLUS, True), (M.LOGCOSH, False)])
def test_cref_matches_numpy_oracle(act, fock):
    N = 5
    hilb, _, _, liouv = lindblad_ising_1d(N, fock=fock)
    net = M.random_machine("ndm", N, 2, act=act, seed=4, std=0.3)
    sr, sc = random_states(hilb, 9, 1), random_states(hilb, 9, 2)
    out, O = net.logpsi_grad(sr, sc)
    cout, cO = cref.ndm_logpsi_grad(net, sr, sc)
    assert np.allclose(cout, out, rtol=1e-13, atol=1e-13) and np.allclose(cO, O, rtol=1e-13, atol=1e-13)
    loc, g = E.local_grad_super(net, liouv, sr, sc)
    cloc, cg = cref.local_grad_super(net, liouv, sr, sc, nthreads=2)
    assert np.allclose(cloc, loc, rtol=1e-12, atol=1e-12) and np.allclose(cg, g, rtol=1e-12, atol=1e-12)
    assert cref.max_threads() >= 1
