"""NDMSymm on the device: tied parameters (set_bare_params!, NDMSymm.jl:79-128), symmetrised gradient
rows (NDMSymmBatched.jl:16-36) and a full SR iteration in the symmetric parameter space, against oracle.machines.NDMSymm."""
import numpy as np
import pytest

import helpers as H
from oracle import machines as OM, sr as OSR
from oracle.models import lindblad_ising_1d

pytestmark = pytest.mark.gpu


def translations(N):
    return [[(i + s) % N + 1 for i in range(N)] for s in range(N)]


def make(nq, ctx, N, ah, aa, dtype, perms, seed=21):
    om = OM.random_ndmsymm(N, ah, aa, perms, seed=seed, std=0.2)
    w = om.params().astype(dtype)
    om.set_params(w.astype(np.float64))
    pm = nq.NDMSymm(ctx, nq.HomogeneousFock(N), dtype, ah, aa, perms)
    assert pm.P == om.P and pm.Pb == om.bare.P
    pm.set_params(w)
    return om, pm


@pytest.mark.parametrize("N,ah,aa,dtype,perms", [(6, 2, 1, np.float64, "T"), (5, 1, 2, np.float64, "T"), (6, 2, 1, np.float32, "T"),
                                                 (4, 2, 2, np.float64, [[1, 2, 3, 4], [4, 3, 2, 1]])])
def test_ndmsymm_params_and_gradient(nq, ctx, N, ah, aa, dtype, perms):
    perms = translations(N) if perms == "T" else perms
    om, pm = make(nq, ctx, N, ah, aa, dtype, perms)
    tol = H.TOL[np.dtype(dtype)]
    # tied parameters: the local biases are averaged in the symmetric vector itself, the bare net follows the permutations
    H.assert_close(pm.params(), om.params(), tol, "symm params")
    H.assert_close(pm.bare.params(), om.bare.params(), tol, "bare params")
    B = 37
    sr, sc = H.rand_states("fock", N, B, 3), H.rand_states("fock", N, B, 4)
    out, O = pm.logpsi_and_grad((sr, sc))
    oout, oO = om.logpsi_grad(sr, sc)
    H.assert_close(out, oout, tol, "log rho")
    H.assert_close(O, oO, tol, "symmetrised O")
    # update!(opt, cnet::NDMSymm, dw): step in the symmetric space, re-tie the bare net
    dw = np.random.default_rng(0).standard_normal(pm.P).astype(dtype)
    pm.update(dw, 0.05)
    om.set_params(om.params() - 0.05 * dw.astype(np.float64))
    H.assert_close(pm.params(), om.params(), tol, "updated symm params")
    H.assert_close(pm.bare.params(), om.bare.params(), tol, "updated bare params")


def test_ndmsymm_sr_iteration(nq, ctx):
    N, B, Lc = 6, 16, 4
    perms = translations(N)
    om, pm = make(nq, ctx, N, 2, 1, np.float64, perms)
    _, _, _, ol = lindblad_ising_1d(N)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N)
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=5, seed=1), pl, nq.SR(np.float32, eps=0.001), batch_sz=B)
    R, Cc = H.rand_states("fock", N, B * Lc, 1), H.rand_states("fock", N, B * Lc, 2)
    bs.set_samples((R.reshape(N, B, Lc, order="F"), Cc.reshape(N, B, Lc, order="F")))
    bs.sample_(sample=False)
    ref = OSR.iteration_liouvillian(om, ol, R, Cc, OSR.eps_f32(0.001))
    H.assert_close(bs.loc.cpu().numpy(), ref["Lloc"], 1e-11, "L_loc")
    H.assert_close(bs.O.cpu().numpy().T + bs.avg.cpu().numpy()[:, None], ref["O"], 1e-11, "O (symm)")
    H.assert_close(bs.gloc.cpu().numpy().T, ref["gLloc"], 1e-11, "grad L_loc (symm)")
    H.assert_close(bs.S.cpu().numpy().T, ref["S"], 1e-11, "S")
    H.assert_close(bs.F.cpu().numpy(), ref["F"], 1e-11, "F")
    dw = bs.precondition_().cpu().numpy()
    assert np.linalg.norm(dw - ref["dw"]) <= 1e-7 * np.linalg.norm(ref["dw"])
    bs.update_(nq.Descent(0.01))
    om.set_params(om.params() - 0.01 * ref["dw"])
    H.assert_close(pm.params(), om.params(), 1e-9, "params after the step")
    # a sampled iteration runs end to end (the chain evaluates the bare net) and lowers nothing silently
    stat, _ = bs.sample_()
    assert np.isfinite(stat.mean.real) and stat.mean.real > 0
