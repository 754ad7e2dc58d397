"""Pins the oracle's machines the way the reference's active tests pin the reference:
test/Machines/test_grad.jl:108-138 (hand gradient == AD of the uncached definition at every
basis state of N=4; here AD is replaced by central finite differences of an independently
written per-sample definition) and test/Machines/test_batched.jl:132-209 (batched == single)."""
import numpy as np
import pytest

from oracle import machines as M
from oracle.hilbert import HomogeneousFock, HomogeneousSpin

N = 4


def _sp(x):
    return np.log(1 + np.exp(x))


def _lc(x):
    return np.log(np.cosh(x))


def defn_rbm(p, sig, N, Mh, act):
    a, b, W = p[:N], p[N:N + Mh], p[N + Mh:].reshape((Mh, N), order="F")
    f = _sp if act == M.SOFTPLUS else _lc
    return a @ sig + np.sum(f(b + W @ sig))


def defn_rbmsplit(p, sr, sc, N, Mh):
    o = 0
    ar = p[o:o + N]; o += N
    ac = p[o:o + N]; o += N
    b = p[o:o + Mh]; o += Mh
    Wr = p[o:o + Mh * N].reshape((Mh, N), order="F"); o += Mh * N
    Wc = p[o:].reshape((Mh, N), order="F")
    return ar @ sr + ac @ sc + np.sum(_sp(b + Wr @ sr + Wc @ sc))


def defn_ndm(p, sr, sc, N, Mh, A, act):
    f = _sp if act == M.SOFTPLUS else _lc
    o = 0
    def take(n, shape=None):
        nonlocal o
        v = p[o:o + n]; o += n
        return v if shape is None else v.reshape(shape, order="F")
    b_mu, h_mu, w_mu, u_mu = take(N), take(Mh), take(Mh * N, (Mh, N)), take(A * N, (A, N))
    b_la, h_la, d_la, w_la, u_la = take(N), take(Mh), take(A), take(Mh * N, (Mh, N)), take(A * N, (A, N))
    ss, ds = sr + sc, sr - sc
    pi = 0.5 * u_la @ ss + 0.5j * (u_mu @ ds) + d_la
    gl = 0.5 * (np.sum(f(h_la + w_la @ sr)) + np.sum(f(h_la + w_la @ sc)) + b_la @ ss)
    gm = 0.5j * (np.sum(f(h_mu + w_mu @ sr)) - np.sum(f(h_mu + w_mu @ sc)) + b_mu @ ds)
    return gl + gm + np.sum(f(pi.astype(complex)))


def fd_grad(fun, p, h=1e-6):
    g = np.zeros(len(p), dtype=complex)
    for k in range(len(p)):
        e = np.zeros(len(p)); e[k] = h
        g[k] = (fun(p + e) - fun(p - e)) / (2 * h)
        if np.iscomplexobj(p):   # holomorphic: the real-direction derivative is the derivative
            pass
    return g


@pytest.mark.parametrize("act", [M.SOFTPLUS, M.LOGCOSH])
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("hk", ["spin", "fock"])
def test_rbm_value_and_grad_all_basis_states(act, cplx, hk):
    hilb = HomogeneousSpin(N) if hk == "spin" else HomogeneousFock(N)
    net = M.random_machine("rbm", N, 2, act=act, complex_weights=cplx, seed=7, std=0.3)
    S = hilb.all_states()
    out, O = net.logpsi_grad(S)
    assert np.allclose(out, net.logpsi(S), atol=1e-14)
    p = net.params()
    for s in range(S.shape[1]):
        ref = defn_rbm(p, S[:, s], N, net.M, act)
        assert abs(ref - out[s]) < 1e-12
        g = fd_grad(lambda q: defn_rbm(q, S[:, s], N, net.M, act), p)
        assert np.allclose(O[:, s], g, atol=1e-8)


@pytest.mark.parametrize("cplx", [False, True])
def test_rbmsplit_value_and_grad(cplx):
    hilb = HomogeneousFock(N)
    net = M.random_machine("rbmsplit", N, 2, complex_weights=cplx, seed=8, std=0.3)
    S = hilb.all_states()
    D = S.shape[1]
    sr = np.repeat(S, D, axis=1); sc = np.tile(S, (1, D))
    out, O = net.logpsi_grad(sr, sc)
    assert np.allclose(out, net.logpsi(sr, sc), atol=1e-14)
    p = net.params()
    for s in range(0, D * D, 3):
        ref = defn_rbmsplit(p, sr[:, s], sc[:, s], N, net.M)
        assert abs(ref - out[s]) < 1e-12
        g = fd_grad(lambda q: defn_rbmsplit(q, sr[:, s], sc[:, s], N, net.M), p)
        assert np.allclose(O[:, s], g, atol=1e-8)


@pytest.mark.parametrize("act", [M.SOFTPLUS, M.LOGCOSH])
@pytest.mark.parametrize("hk", ["spin", "fock"])
def test_ndm_value_and_grad(act, hk):
    hilb = HomogeneousSpin(N) if hk == "spin" else HomogeneousFock(N)
    net = M.random_machine("ndm", N, 2, act=act, seed=9, std=0.3, alpha_a=1)
    S = hilb.all_states()
    D = S.shape[1]
    sr = np.repeat(S, D, axis=1); sc = np.tile(S, (1, D))
    out, O = net.logpsi_grad(sr, sc)
    assert np.allclose(out, net.logpsi(sr, sc), atol=1e-14)
    p = net.params()
    for s in range(0, D * D, 5):
        ref = defn_ndm(p, sr[:, s], sc[:, s], N, net.M, net.A, act)
        assert abs(ref - out[s]) < 1e-12
        g = fd_grad(lambda q: defn_ndm(q, sr[:, s], sc[:, s], N, net.M, net.A, act), p)
        assert np.allclose(O[:, s], g, atol=1e-8)
    # on the diagonal rho is real positive: log rho(s, s) is real  (NDM is positive by construction)
    od = net.logpsi(S, S)
    assert np.allclose(od.imag, 0, atol=1e-13)
    # hermiticity: rho(s, s') = conj(rho(s', s))
    assert np.allclose(net.logpsi(sr, sc), np.conj(net.logpsi(sc, sr)), atol=1e-13)


def test_param_roundtrip_and_layout():
    for kind in ("rbm", "rbmsplit", "ndm"):
        net = M.random_machine(kind, 3, 2, seed=3)
        p = net.params()
        assert len(p) == net.P
        net2 = M.random_machine(kind, 3, 2, seed=99)
        net2.set_params(p)
        assert np.array_equal(net2.params(), p)
    r = M.random_machine("rbm", 3, 2, seed=3)
    # W[k, j] sits at N + M + k + M*j (column-major, tuple_logic.jl:104-118)
    assert r.params()[3 + 6 + 1 + 6 * 2] == r.W[1, 2]


def test_logcosh_large_argument_branch():
    x = np.array([-30.0, -12.0, 12.0, 13.0, 40.0])
    assert np.allclose(M.logcosh(x), np.abs(x) + np.log1p(np.exp(-2 * np.abs(x))) - np.log(2), atol=1e-9)


def test_ndmsymm_gradient_is_the_derivative_in_the_symmetric_space():
    """NDMSymm.jl:36-76: grad_symm = G grad_bare must be the derivative of log rho(bare(w_symm)) with respect to the
    symmetric parameters (the local biases enter through their mean)."""
    from oracle import machines as OM
    N = 4
    perms = [[(i + s) % N + 1 for i in range(N)] for s in range(N)]
    net = OM.random_ndmsymm(N, 2, 1, perms, seed=5, std=0.3)
    assert net.bare.M == 2 * N and net.bare.A == N
    rng = np.random.default_rng(1)
    sr = rng.integers(0, 2, size=(N, 5)).astype(float)
    sc = rng.integers(0, 2, size=(N, 5)).astype(float)
    out, O = net.logpsi_grad(sr, sc)
    w0 = net.params().copy()
    h = 1e-6
    for p in rng.choice(net.P, size=12, replace=False):
        wp, wm = w0.copy(), w0.copy()
        wp[p] += h
        wm[p] -= h
        net.set_params(wp)
        fp = net.logpsi(sr, sc)
        net.set_params(wm)
        fm = net.logpsi(sr, sc)
        assert np.allclose((fp - fm) / (2 * h), O[p], atol=1e-7), p
    net.set_params(w0)
    # translation invariance of the symmetrised density matrix: rho(T sigma, T sigma') = rho(sigma, sigma')
    shift = lambda a: np.roll(a, 1, axis=0)
    assert np.allclose(net.logpsi(shift(sr), shift(sc)), net.logpsi(sr, sc), atol=1e-12)
