"""GPU parity: the whole iteration (eval+grad -> local estimator -> centre -> force -> S -> solve ->
update) on identical sample batches vs the oracle, for the BASELINE cfg1/cfg2 shapes, plus a short
end-to-end optimisation whose energy must approach exact diagonalisation within Monte-Carlo error."""
import numpy as np
import pytest

import helpers as H
from oracle import machines as OM
from oracle import operators as OOPS
from oracle import sr as OSR
from oracle.models import lindblad_ising_1d, tfim_1d

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,algo", [(np.complex128, "sr_cholesky"), (np.float64, "sr_cholesky"),
                                        (np.complex128, "sr_cg"), (np.complex64, "sr_cholesky")])
def test_cfg1_ground_state_iteration(nq, ctx, dtype, algo):
    """cfg1: TFIM 1D N=10, RBM alpha=2 logcosh, B=8 x L=125."""
    N, B, Lc = 10, 8, 125
    oh, oH = tfim_1d(N)
    ph, pH = H.p_tfim_1d(nq, N)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, dtype, OM.LOGCOSH)
    tol = H.TOL[np.dtype(dtype)]
    eps = 0.1
    sr_algo = nq.SR(np.float32, eps=eps, algorithm=algo, precision=1e-13, full_matrix=False)
    smp = nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=10, seed=1)
    bs = nq.BatchedSampler(pm, smp, pH, sr_algo, batch_sz=B)
    S = H.rand_states("spin", N, B * Lc, 4321)
    bs.set_samples(S.reshape(N, B, Lc, order="F"))
    stat, _ = bs.sample_(sample=False)
    ref = OSR.iteration_ket(om, oH, S, OSR.eps_f32(eps))
    H.assert_close(bs.logpsi.cpu().numpy(), ref["logpsi"], tol, "logpsi")
    H.assert_close(bs.loc.cpu().numpy(), ref["Eloc"], tol, "E_loc")
    H.assert_close(bs.avg.cpu().numpy(), ref["O_avg"], tol, "<O>")
    gref = OSR.force_ket_ld(ref["Eloc"], ref["O"])              # long-double accumulation of the oracle's formula
    assert np.linalg.norm(gref - ref["gradC"]) <= 1e-12 * np.linalg.norm(gref)
    H.assert_close(bs.gradC.cpu().numpy(), gref, tol, "grad C")
    H.assert_close(bs.F.cpu().numpy(), np.real(gref) if bs.real_params else gref, tol, "F")
    if bs.S is not None:
        H.assert_close(bs.S.cpu().numpy().T, ref["S"], tol, "S")
    assert abs(stat.mean - ref["Eloc"].mean()) <= tol * abs(ref["Eloc"].mean()) * 10
    dw = bs.precondition_().cpu().numpy()
    cond = np.linalg.cond(ref["S"] + OSR.eps_f32(eps) * np.eye(pm.P))
    lim = max(1e-8 if algo == "sr_cg" else 1e-10, 50 * tol * cond if np.dtype(dtype).itemsize <= 8 else 0)
    assert np.linalg.norm(dw - ref["dw"]) <= lim * np.linalg.norm(ref["dw"]), "dw"
    w0 = pm.params()
    bs.update_(nq.Descent(0.1))
    H.assert_close(pm.params(), w0 - np.asarray(0.1 * dw).astype(w0.dtype), 10 * tol, "update")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cfg2_steady_state_iteration(nq, ctx, dtype):
    """cfg2: dissipative Ising N=8, NDM alpha=2, B=16 x L=125 (the real chain length; the C restatement
    oracle/cref.c cross-checks the NumPy oracle's L_loc / grad L_loc on the full batch in test_oracle_cref.py)."""
    N, B, Lc = 8, 16, 125
    _, _, _, ol = lindblad_ising_1d(N)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N)
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, dtype, OM.SOFTPLUS)
    tol = H.TOL[np.dtype(dtype)]
    eps = 0.001
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=10, seed=2), pl,
                           nq.SR(np.float32, eps=eps, algorithm="sr_cholesky"), batch_sz=B)
    R, Cc = H.rand_states("fock", N, B * Lc, 11), H.rand_states("fock", N, B * Lc, 12)
    bs.set_samples((R.reshape(N, B, Lc, order="F"), Cc.reshape(N, B, Lc, order="F")))
    stat, _ = bs.sample_(sample=False)
    ref = OSR.iteration_liouvillian(om, ol, R, Cc, OSR.eps_f32(eps))
    H.assert_close(bs.loc.cpu().numpy(), ref["Lloc"], tol, "L_loc")
    H.assert_close(bs.gloc.cpu().numpy().T, ref["gLloc"], tol, "grad L_loc")
    # centred rows are differences of O(1) numbers: bound relative to max |O| (both sides round O - <O> once)
    Oc_ref = ref["O"] - ref["O_avg"][:, None]
    assert np.max(np.abs(bs.O.cpu().numpy().T - Oc_ref)) <= tol * np.abs(ref["O"]).max(), "O centred"
    assert abs(bs.cost - ref["C"]) <= tol * ref["C"] and abs(stat.mean.real - ref["C"]) <= tol * ref["C"]
    gref = OSR.force_liouvillian_ld(ref["Lloc"], ref["gLloc"], ref["O"])
    assert np.linalg.norm(gref - ref["gradC"]) <= 1e-12 * np.linalg.norm(gref)
    H.assert_close(bs.gradC.cpu().numpy(), gref, tol, "grad C")
    H.assert_close(bs.S.cpu().numpy().T, ref["S"], tol, "S")
    H.assert_close(bs.F.cpu().numpy(), np.real(gref), tol, "F")
    if np.dtype(dtype) == np.float64:
        dw = bs.precondition_().cpu().numpy()
        cond = np.linalg.cond(ref["S"] + OSR.eps_f32(eps) * np.eye(pm.P))
        assert np.linalg.norm(dw - ref["dw"]) <= max(1e-10, 1e-14 * cond) * np.linalg.norm(ref["dw"])


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_cfg3_ground_state_iteration(nq, ctx, dtype):
    """cfg3 shape: TFIM 2D 6x6 (h=3, J=1), complex RBM alpha=4 logcosh (P = 5364), 1024 chains x L=1:
    log psi, E_loc, <O>, force, S (explicit, complex Hermitian) and the Cholesky update against the oracle."""
    from oracle.models import tfim_2d
    Lx, B, Lc = 6, 1024, 1
    N = Lx * Lx
    oh, oH = tfim_2d(Lx, 3.0, 1.0)
    ph, pH = H.p_tfim_2d(nq, Lx, 3.0, 1.0)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 4, dtype, OM.LOGCOSH)
    assert pm.P == 5364
    tol = H.TOL[np.dtype(dtype)]
    eps = 0.1
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=2, seed=1), pH,
                           nq.SR(np.float32, eps=eps, algorithm="sr_cholesky"), batch_sz=B)
    S = H.rand_states("spin", N, B * Lc, 77)
    bs.set_samples(S.reshape(N, B, Lc, order="F"))
    stat, _ = bs.sample_(sample=False)
    ref = OSR.iteration_ket(om, oH, S, OSR.eps_f32(eps))
    H.assert_close(bs.logpsi.cpu().numpy(), ref["logpsi"], tol, "logpsi")
    H.assert_close(bs.loc.cpu().numpy(), ref["Eloc"], tol, "E_loc")
    H.assert_close(bs.avg.cpu().numpy(), ref["O_avg"], tol, "<O>")
    gref = OSR.force_ket_ld(ref["Eloc"], ref["O"])
    H.assert_close(bs.gradC.cpu().numpy(), gref, tol, "grad C")
    H.assert_close(bs.S.cpu().numpy().T, ref["S"], tol, "S")
    if np.dtype(dtype) == np.complex128:
        dw = bs.precondition_().cpu().numpy()
        # Ns < P: S is singular, lambda_min(S + eps I) = eps, lambda_max <= tr S  ->  cond <= (tr S + eps) / eps
        cond = (np.trace(ref["S"]).real + OSR.eps_f32(eps)) / OSR.eps_f32(eps)
        assert np.linalg.norm(dw - ref["dw"]) <= max(1e-10, 1e-14 * cond) * np.linalg.norm(ref["dw"]), "dw"


def test_ground_state_optimisation_reaches_exact_energy(nq, ctx):
    """TFIM 1D N=6 (h=J=1): sampled SR descent approaches the exact ground energy."""
    N = 6
    oh, oH = tfim_1d(N)
    E0 = np.linalg.eigvalsh(OOPS.to_matrix(oH))[0]
    ph, pH = H.p_tfim_1d(nq, N)
    net = nq.RBM(ctx, ph, np.float64, 2, nq.af_logcosh)
    nq.init_random_pars_(net, sigma=0.01, seed=1234)
    bs = nq.BatchedSampler(net, nq.MetropolisSampler(nq.LocalRule(), 64, N, burn=30, seed=5), pH,
                           nq.SR(np.float32, eps=0.1, algorithm="sr_cholesky"), batch_sz=64)
    opt = nq.Descent(0.1)
    hist = []
    for it in range(120):
        stat, _ = bs.sample_()
        bs.precondition_(it + 1)
        bs.update_(opt)
        hist.append(stat)
    last = hist[-1]
    best = np.mean([h.mean.real for h in hist[-10:]])
    assert hist[0].mean.real > best
    assert abs(best - E0) < 0.02 * abs(E0) + 5 * last.error, (best, E0, last.error)


def test_steady_state_optimisation_reduces_cost(nq, ctx):
    N = 3
    ph, _, _, pl = H.p_lindblad_ising_1d(nq, N)
    net = nq.NDM(ctx, ph, np.float64, 1, 1, nq.af_softplus, seed=3)
    bs = nq.BatchedSampler(net, nq.MetropolisSampler(nq.LocalRule(), 64, N, burn=30, seed=6), pl,
                           nq.SR(np.float32, eps=0.001, algorithm="sr_cholesky"), batch_sz=32)
    costs = []
    for it in range(60):
        stat, _ = bs.sample_()
        bs.precondition_(it + 1)
        bs.update_(nq.Descent(0.02))
        costs.append(stat.mean.real)
    assert np.mean(costs[-5:]) < 0.5 * np.mean(costs[:5])


def test_ket_observables_on_stored_samples(nq, ctx):
    """compute_observables (BatchedObsKetSampler.jl:32-59): O_loc of the stored samples through the E_loc kernel +
    chain statistics, against the oracle's local estimator and stat_analysis."""
    from oracle import estimators as OE
    from oracle import stats as OST
    from oracle import operators as OO
    N, B, Lc = 8, 16, 20
    oh, oH = tfim_1d(N)
    ph, pH = H.p_tfim_1d(nq, N)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, np.complex128, OM.LOGCOSH)
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=10, seed=3), pH,
                           nq.SR(np.float32, eps=0.1, algorithm="sr_cholesky"), batch_sz=B)
    S = H.rand_states("spin", N, B * Lc, 77)
    bs.set_samples(S.reshape(N, B, Lc, order="F"))
    bs.evaluate()
    # magnetisation along x and a two-site correlator, built with the same operator algebra on both sides
    p_mx, o_mx = nq.LocalOperator(ph), None
    for i in range(1, N + 1):
        p_mx = p_mx + (1.0 / N) * nq.sigmax(ph, i)
        o_mx = OO.add(o_mx, OO.scale(1.0 / N, OO.sigmax(oh, i)))
    p_zz = nq.LocalOperator(ph) + nq.sigmaz(ph, 1) * nq.sigmaz(ph, 4)
    o_zz = OO.mul(OO.sigmaz(oh, 1), OO.sigmaz(oh, 4))
    bs.add_observable_("mx", p_mx)
    bs.add_observable_("zz14", p_zz)
    res = bs.compute_observables()
    assert set(res) == {"mx", "zz14"}
    for name, oop in (("mx", o_mx), ("zz14", o_zz)):
        ref_loc = OE.local_scalar_ket(om, oop, S)
        ref = OST.stat_analysis(ref_loc.reshape(B, Lc, order="F"))
        m = res[name]
        assert abs(m.mean - ref["mean"]) <= 1e-11 * max(1.0, abs(ref["mean"])), name
        assert abs(m.error - ref["error"]) <= 1e-9 * max(1e-3, ref["error"]), name


def test_nesterov_update_rule(nq, ctx):
    """Optimisers.Nesterov (rules.jl:36-55): d = mu^2 v - (1+mu) lr dw, v <- mu v - lr dw, w <- w + d."""
    import torch
    N = 6
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, np.complex128, OM.LOGCOSH)
    ph, pH = H.p_tfim_1d(nq, N)
    bs = nq.BatchedSampler(pm, nq.MetropolisSampler(nq.LocalRule(), 4, N, burn=2, seed=3), pH,
                           nq.SR(np.float32, eps=0.1, algorithm="sr_cholesky"), batch_sz=4)
    opt = nq.Nesterov(0.05, 0.9)
    rng = np.random.default_rng(2)
    w = pm.params().astype(np.complex128)
    v = np.zeros_like(w)
    for it in range(3):
        g = rng.standard_normal(pm.P) + 1j * rng.standard_normal(pm.P)
        bs.update_(opt, torch.from_numpy(g).cuda())
        d = 0.81 * v - 1.9 * 0.05 * g
        v = 0.9 * v - 0.05 * g
        w = w + d
        H.assert_close(pm.params(), w, 1e-13, "Nesterov step %d" % it)
