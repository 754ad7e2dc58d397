"""Pins the oracle's estimators, force, S and solve against dense linear algebra
(the reference's own tests for these are disabled)."""
import numpy as np
import pytest

from oracle import estimators as E
from oracle import machines as M
from oracle import operators as ops
from oracle import sr
from oracle.hilbert import super_state
from oracle.models import lindblad_ising_1d, tfim_1d


def full_super_states(h):
    D = h.spacedim()
    rows, cols = zip(*[super_state(h, i) for i in range(1, D * D + 1)])
    return np.stack(rows, 1), np.stack(cols, 1)


@pytest.mark.parametrize("cplx", [False, True])
def test_eloc_equals_Hpsi_over_psi(cplx):
    hilb, H = tfim_1d(4, h=0.9, J=1.1)
    net = M.random_machine("rbm", 4, 2, act=M.LOGCOSH, complex_weights=cplx, seed=5, std=0.3)
    S = hilb.all_states()
    psi = np.exp(net.logpsi(S))
    Hm = ops.to_matrix(H)
    Eloc = E.local_scalar_ket(net, H, S)
    assert np.allclose(Eloc, (Hm @ psi) / psi, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("kind", ["ndm", "rbmsplit"])
def test_Lloc_equals_Lrho_over_rho(kind):
    hilb, H, jumps, liouv = lindblad_ising_1d(3, 0.4, 2.0)
    net = M.random_machine(kind, 3, 2, seed=6, std=0.3, complex_weights=(kind == "rbmsplit"))
    sr_, sc_ = full_super_states(hilb)
    lr = net.logpsi(sr_, sc_)
    rho = np.exp(lr)                      # vec index (idx(col)-1) D + idx(row) by construction
    Lm = ops.to_matrix(liouv)
    Lloc, gL = E.local_grad_super(net, liouv, sr_, sc_)
    assert np.allclose(Lloc, (Lm @ rho) / rho, rtol=1e-11, atol=1e-12)
    assert np.allclose(E.local_scalar_super(net, liouv, sr_, sc_), Lloc, rtol=1e-12, atol=1e-13)
    # grad L_loc_k = sum_eta L[s,eta] rho(eta)/rho(s) O_k(eta)
    _, O = net.logpsi_grad(sr_, sc_)
    ref = ((Lm * rho[None, :]) @ O.T).T / rho[None, :]
    assert np.allclose(gL, ref, rtol=1e-10, atol=1e-11)


def test_liouvillian_force_is_cost_gradient_full_space():
    """2 Re F_k == dC/dtheta_k with C = ||L rho||^2/||rho||^2, probabilities |rho|^2/Z."""
    hilb, H, jumps, liouv = lindblad_ising_1d(3, 0.4, 2.0)
    net = M.random_machine("ndm", 3, 1, seed=11, std=0.3)
    sr_, sc_ = full_super_states(hilb)
    Lm = ops.to_matrix(liouv)

    def cost(p):
        n2 = M.random_machine("ndm", 3, 1, seed=11, std=0.3)
        n2.set_params(p)
        rho = np.exp(n2.logpsi(sr_, sc_))
        return np.linalg.norm(Lm @ rho) ** 2 / np.linalg.norm(rho) ** 2

    out, O = net.logpsi_grad(sr_, sc_)
    prob = np.abs(np.exp(out)) ** 2
    prob /= prob.sum()
    Lloc, gL = E.local_grad_super(net, liouv, sr_, sc_, out)
    C = np.sum(prob * np.abs(Lloc) ** 2)
    assert abs(C - cost(net.params())) < 1e-12
    Oavg = O @ prob
    F = (gL * np.conj(Lloc)[None, :]) @ prob - C * Oavg
    p = net.params()
    for k in range(0, len(p), 3):
        e = np.zeros(len(p)); e[k] = 1e-6
        fd = (cost(p + e) - cost(p - e)) / 2e-6
        assert abs(2 * F[k].real - fd) < 1e-7
    # the sampled-force formula with uniform weights reproduces its own definition
    g = sr.force_liouvillian(Lloc, gL, O.mean(1))
    ref = (gL * np.conj(Lloc)[None, :]).mean(1) - np.mean(np.abs(Lloc) ** 2) * O.mean(1)
    assert np.allclose(g, ref, atol=1e-14)


def test_ket_force_is_energy_gradient_full_space():
    hilb, H = tfim_1d(4)
    net = M.random_machine("rbm", 4, 1, act=M.LOGCOSH, seed=12, std=0.3)
    S = hilb.all_states()
    Hm = ops.to_matrix(H)

    def energy(p):
        n2 = M.random_machine("rbm", 4, 1, act=M.LOGCOSH, seed=12)
        n2.set_params(p)
        psi = np.exp(n2.logpsi(S))
        return np.real(psi.conj() @ Hm @ psi / (psi.conj() @ psi))

    out, O = net.logpsi_grad(S)
    prob = np.abs(np.exp(out)) ** 2; prob /= prob.sum()
    Eloc = E.local_scalar_ket(net, H, S, out)
    Oc = O - (O @ prob)[:, None]
    F = (np.conj(Oc) * Eloc[None, :]) @ prob
    p = net.params()
    for k in range(len(p)):
        e = np.zeros(len(p)); e[k] = 1e-6
        assert abs(2 * F[k].real - (energy(p + e) - energy(p - e)) / 2e-6) < 1e-7


@pytest.mark.parametrize("real_params", [True, False])
def test_sr_setup_and_solvers(real_params):
    rng = np.random.default_rng(3)
    P, Ns = 12, 200
    O = rng.standard_normal((P, Ns)) + 1j * rng.standard_normal((P, Ns))
    avg, Oc = sr.center(O)
    assert np.allclose(Oc.mean(1), 0, atol=1e-14)
    gradC = rng.standard_normal(P) + 1j * rng.standard_normal(P)
    S, F = sr.sr_setup(Oc, gradC, real_params)
    # S_kl = <conj(O_k) O_l>_c (complex nets) or its real part
    ref = (np.conj(Oc) @ Oc.T) / Ns
    assert np.allclose(S, ref.real if real_params else ref, atol=1e-13)
    eps = sr.eps_f32(0.001)
    assert eps == 0.0010000000474974513
    dw = sr.solve_cholesky(S, F, eps)
    assert np.allclose(dw, np.linalg.solve(S + eps * np.eye(P), F), rtol=1e-10)
    x, it, ok = sr.solve_cg(Oc, F, eps, 1e-12, real_params)
    assert ok and np.allclose(x, dw, rtol=1e-8, atol=1e-10)
    x2, it2, ok2 = sr.solve_cg_explicit(S, F, eps, 1e-12)
    assert ok2 and np.allclose(x2, dw, rtol=1e-8, atol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        sr.solve_cholesky(-np.eye(3), np.ones(3), 0.0)


def test_iteration_drivers_run():
    hilb, H = tfim_1d(4)
    net = M.random_machine("rbm", 4, 2, act=M.LOGCOSH, complex_weights=True, seed=1)
    S = hilb.all_states()
    r = sr.iteration_ket(net, H, S, 0.1)
    assert r["dw"].shape == (net.P,) and np.all(np.isfinite(r["dw"]))
    hilb, _, _, liouv = lindblad_ising_1d(3)
    ndm = M.random_machine("ndm", 3, 1, seed=2)
    a, b = full_super_states(hilb)
    r = sr.iteration_liouvillian(ndm, liouv, a, b, 1e-3)
    assert r["dw"].dtype == np.float64 and np.all(np.isfinite(r["dw"]))


def test_density_matrix_observable_estimator_is_trace_formula():
    """Observables of a density matrix (BatchedObsDMSampler.jl): with p(sigma) ~ rho(sigma, sigma) and the operator
    acting on the row index, sum_sigma p(sigma) O_loc(sigma) = Tr(O rho) / Tr(rho) on the full space."""
    from oracle.hilbert import HomogeneousFock
    N = 3
    hilb = HomogeneousFock(N)
    net = M.random_machine("ndm", N, 2, seed=9, std=0.4)
    allS = hilb.all_states()
    D = allS.shape[1]
    rho = np.exp(net.logpsi(np.repeat(allS, D, axis=1), np.tile(allS, (1, D)))).reshape(D, D)     # rho[row, col]
    assert np.allclose(rho, rho.conj().T, atol=1e-12)                  # Hermitian by construction of the NDM
    p = np.diag(rho).real
    assert np.all(p > 0)
    O = ops.add(ops.mul(ops.sigmay(hilb, 1), ops.sigmaz(hilb, 3)), ops.scale(0.5, ops.sigmax(hilb, 2)))
    left = ops.KLocalLiouvillian(hilb, ops._tensor_left(O), None, None)
    loc = E.local_scalar_super(net, left, allS, allS)
    exact = np.trace(ops.to_matrix(O) @ rho) / np.trace(rho)
    assert abs(np.sum(p * loc) / p.sum() - exact) <= 1e-12 * max(1.0, abs(exact))


@pytest.mark.parametrize("cplx", [False, True])
def test_minres_restatement_converges_to_the_direct_solution(cplx):
    rng = np.random.default_rng(3)
    P = 80
    X = rng.standard_normal((P, 3 * P)) + (1j * rng.standard_normal((P, 3 * P)) if cplx else 0)
    S = X @ X.conj().T / (3 * P)
    F = rng.standard_normal(P) + (1j * rng.standard_normal(P) if cplx else 0)
    x, it, ok = sr.solve_minres_explicit(S, F, 1e-3, 1e-12)
    ref = np.linalg.solve(S + 1e-3 * np.eye(P), F)
    assert ok and 0 < it <= 10 * P
    assert np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)
    x3, it3, ok3 = sr.solve_minres_explicit(S, F, 1e-3, 1e-30, maxiter=3)
    assert it3 == 3 and not ok3
