"""Oracle: RBM / RBMSplit / NDM values and gradients (batched).  Test infrastructure only.

Follows
  src/Networks/activation.jl:5-29
  src/Networks/ClosedSystems/RBM.jl:50-51, :87-105;  RBMBatched.jl:37-91
  src/Networks/MixedDensityMatrix/RBMSplit.jl:49-50, :111-129; RBMSplitBatched.jl:35-101
  src/Networks/MixedDensityMatrix/NDM.jl:76-96, :241-337;  NDMBatched.jl:94-280
  src/tuple_logic.jl:82-118, src/base_batched_networks.jl:124-145 (flat gradient layout:
  one [P, B] buffer, parameter index fastest, fields in functor order, matrices column-major)
  src/base_batched_networks.jl:255-259 (log_prob = 2 Re log psi)
All arithmetic in float64 / complex128 regardless of the precision mode under test.
"""
import numpy as np

SOFTPLUS, LOGCOSH = 0, 1


# ----- activations: activation.jl ------------------------------------------------------
def softplus(x):            # logL(x) = log1p(exp(x))            :15
    return np.log1p(np.exp(x))


def d_softplus(x):          # dlogL(x) = 1/(1+exp(-x))           :13
    return 1.0 / (1.0 + np.exp(-x))


def logcosh(x):             # logL2                              :7-10
    x = np.asarray(x)
    if np.iscomplexobj(x):
        re, im = x.real, x.imag
        return _logcosh_real(re) + np.log(np.cos(im) + 1j * np.tanh(re) * np.sin(im))
    return _logcosh_real(x)


def _logcosh_real(x):
    ax = np.abs(x)
    small = ax <= 12.0
    out = ax - np.log(2.0)
    with np.errstate(over="ignore"):
        out = np.where(small, np.log(np.cosh(np.where(small, x, 0.0))), out)
    return out


def d_logcosh(x):           # tanh                               :6,26
    return np.tanh(x)


ACT = {SOFTPLUS: (softplus, d_softplus), LOGCOSH: (logcosh, d_logcosh)}


def _f64(sig):
    return np.asarray(sig, dtype=np.float64)


# ----- RBM ------------------------------------------------------------------------------
class RBM:
    """Parameters a[N], b[M], W[M,N]; real or complex.  Flat order [a, b, vec(W)] col-major."""
    kind = "rbm"
    doubled = False

    def __init__(self, a, b, W, act=SOFTPLUS):
        self.a, self.b, self.W, self.act = np.asarray(a), np.asarray(b), np.asarray(W), act
        self.N, self.M = len(self.a), len(self.b)

    @property
    def P(self):
        return self.N + self.M + self.M * self.N

    @property
    def is_complex(self):
        return np.iscomplexobj(self.W)

    def params(self):
        return np.concatenate([self.a, self.b, self.W.reshape(-1, order="F")])

    def set_params(self, w):
        N, M = self.N, self.M
        self.a, self.b = w[:N].copy(), w[N:N + M].copy()
        self.W = w[N + M:].reshape((M, N), order="F").copy()

    def logpsi(self, sig):                          # RBM.jl:50-51
        sig = _f64(sig)
        f, _ = ACT[self.act]
        theta = self.W @ sig + self.b[:, None]
        return self.a @ sig + f(theta).sum(axis=0)

    def logpsi_grad(self, sig):                     # RBMBatched.jl:58-91
        sig = _f64(sig)
        f, df = ACT[self.act]
        B = sig.shape[1]
        theta = self.W @ sig + self.b[:, None]
        out = self.a @ sig + f(theta).sum(axis=0)
        d = df(theta)
        dt = np.result_type(self.W.dtype, np.float64)
        O = np.empty((self.P, B), dtype=dt)
        N, M = self.N, self.M
        O[:N] = sig
        O[N:N + M] = d
        # grad.W[k, j, b] = d[k, b] * sig[j, b], flattened k + M*j   (utils/math.jl:23-32)
        O[N + M:] = (sig[:, None, :] * d[None, :, :]).reshape(N * M, B)
        return out, O


# ----- RBMSplit -------------------------------------------------------------------------
class RBMSplit:
    """ar, ac [N]; b [M]; Wr, Wc [M,N].  Activation hard-wired to softplus."""
    kind = "rbmsplit"
    doubled = True
    act = SOFTPLUS

    def __init__(self, ar, ac, b, Wr, Wc):
        self.ar, self.ac, self.b = np.asarray(ar), np.asarray(ac), np.asarray(b)
        self.Wr, self.Wc = np.asarray(Wr), np.asarray(Wc)
        self.N, self.M = len(self.ar), len(self.b)

    @property
    def P(self):
        return 2 * self.N + self.M + 2 * self.M * self.N

    @property
    def is_complex(self):
        return np.iscomplexobj(self.Wr)

    def params(self):
        return np.concatenate([self.ar, self.ac, self.b,
                               self.Wr.reshape(-1, order="F"), self.Wc.reshape(-1, order="F")])

    def set_params(self, w):
        N, M = self.N, self.M
        o = 0
        self.ar = w[o:o + N].copy(); o += N
        self.ac = w[o:o + N].copy(); o += N
        self.b = w[o:o + M].copy(); o += M
        self.Wr = w[o:o + M * N].reshape((M, N), order="F").copy(); o += M * N
        self.Wc = w[o:o + M * N].reshape((M, N), order="F").copy()

    def logpsi(self, sr, sc):                       # RBMSplit.jl:49-50
        sr, sc = _f64(sr), _f64(sc)
        theta = self.Wr @ sr + self.Wc @ sc + self.b[:, None]
        return self.ar @ sr + self.ac @ sc + softplus(theta).sum(axis=0)

    def logpsi_grad(self, sr, sc):                  # RBMSplitBatched.jl:64-101 (no conj: quirk Q1)
        sr, sc = _f64(sr), _f64(sc)
        B = sr.shape[1]
        N, M = self.N, self.M
        theta = self.Wr @ sr + self.Wc @ sc + self.b[:, None]
        out = self.ar @ sr + self.ac @ sc + softplus(theta).sum(axis=0)
        s = d_softplus(theta)
        dt = np.result_type(self.Wr.dtype, np.float64)
        O = np.empty((self.P, B), dtype=dt)
        o = 0
        O[o:o + N] = sr; o += N
        O[o:o + N] = sc; o += N
        O[o:o + M] = s; o += M
        O[o:o + M * N] = (sr[:, None, :] * s[None, :, :]).reshape(N * M, B); o += M * N
        O[o:o + M * N] = (sc[:, None, :] * s[None, :, :]).reshape(N * M, B)
        return out, O


# ----- NDM ------------------------------------------------------------------------------
class NDM:
    """Real parameters, complex output.  Functor order (NDM.jl:17-19):
    b_mu[N], h_mu[M], w_mu[M,N], u_mu[A,N], b_lam[N], h_lam[M], d_lam[A], w_lam[M,N], u_lam[A,N]."""
    kind = "ndm"
    doubled = True
    is_complex = False

    def __init__(self, b_mu, h_mu, w_mu, u_mu, b_lam, h_lam, d_lam, w_lam, u_lam, act=SOFTPLUS):
        self.b_mu, self.h_mu, self.w_mu, self.u_mu = map(np.asarray, (b_mu, h_mu, w_mu, u_mu))
        self.b_lam, self.h_lam, self.d_lam = map(np.asarray, (b_lam, h_lam, d_lam))
        self.w_lam, self.u_lam = np.asarray(w_lam), np.asarray(u_lam)
        self.act = act
        self.N, self.M, self.A = len(self.b_mu), len(self.h_mu), len(self.d_lam)

    @property
    def P(self):
        N, M, A = self.N, self.M, self.A
        return 2 * N + 2 * M + A + 2 * M * N + 2 * A * N

    def _fields(self):
        return ["b_mu", "h_mu", "w_mu", "u_mu", "b_lam", "h_lam", "d_lam", "w_lam", "u_lam"]

    def params(self):
        return np.concatenate([getattr(self, n).reshape(-1, order="F") for n in self._fields()])

    def set_params(self, w):
        o = 0
        for n in self._fields():
            cur = getattr(self, n)
            sz = cur.size
            setattr(self, n, np.asarray(w[o:o + sz]).real.reshape(cur.shape, order="F").copy())
            o += sz

    def _common(self, sr, sc):
        f, df = ACT[self.act]
        th_l = self.w_lam @ sr + self.h_lam[:, None]
        th_m = self.w_mu @ sr + self.h_mu[:, None]
        th_lp = self.w_lam @ sc + self.h_lam[:, None]
        th_mp = self.w_mu @ sc + self.h_mu[:, None]
        ssum, sdif = sr + sc, sr - sc
        pi = 0.5 * (self.u_lam @ ssum) + 0.5j * (self.u_mu @ sdif) + self.d_lam[:, None]
        g_l = 0.5 * (f(th_l).sum(0) + f(th_lp).sum(0) + self.b_lam @ ssum)
        g_m = 0.5 * (f(th_m).sum(0) - f(th_mp).sum(0) + self.b_mu @ sdif)
        out = g_l + 1j * g_m + f(pi).sum(0)
        return out, (th_l, th_m, th_lp, th_mp, pi, ssum, sdif)

    def logpsi(self, sr, sc):                       # NDM.jl:76-96 / NDMBatched.jl:94-175
        return self._common(_f64(sr), _f64(sc))[0]

    def logpsi_grad(self, sr, sc):                  # NDMBatched.jl:177-280
        sr, sc = _f64(sr), _f64(sc)
        _, df = ACT[self.act]
        out, (th_l, th_m, th_lp, th_mp, pi, ssum, sdif) = self._common(sr, sc)
        dl, dm, dlp, dmp, dpi = df(th_l), df(th_m), df(th_lp), df(th_mp), df(pi)
        B = sr.shape[1]
        N, M, A = self.N, self.M, self.A

        def outer(d, s):          # [K,B],[N,B] -> [K*N,B], index k + K*j
            return (s[:, None, :] * d[None, :, :]).reshape(-1, B)
        blocks = [
            0.5j * sdif,                                   # b_mu
            0.5j * (dm - dmp),                             # h_mu
            0.5j * (outer(dm, sr) - outer(dmp, sc)),       # w_mu
            0.5j * outer(dpi, sdif),                       # u_mu
            0.5 * ssum + 0j,                               # b_lam
            0.5 * (dl + dlp) + 0j,                         # h_lam
            dpi,                                           # d_lam
            0.5 * (outer(dl, sr) + outer(dlp, sc)) + 0j,   # w_lam
            0.5 * outer(dpi, ssum),                        # u_lam
        ]
        O = np.concatenate(blocks, axis=0).astype(np.complex128)
        assert O.shape == (self.P, B)
        return out, O


def log_prob(out):                                  # base_batched_networks.jl:255-259
    return 2.0 * np.real(out)


# ----- synthetic parameters (SURVEY 8d / utils/rng.jl:27-32) ------------------------------
def random_machine(kind, N, alpha, *, act=SOFTPLUS, complex_weights=False, seed=1234,
                   std=0.1, alpha_a=None):
    """N(0, std^2) parameters from numpy Philox(seed); complex: re, im each N(0, std^2/2)."""
    rng = np.random.Generator(np.random.Philox(seed))
    M = int(alpha * N)

    def rn(*shape):
        if complex_weights:
            return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * std / np.sqrt(2)
        return rng.standard_normal(shape) * std
    if kind == "rbm":
        return RBM(rn(N), rn(M), rn(M, N), act)
    if kind == "rbmsplit":
        return RBMSplit(rn(N), rn(N), rn(M), rn(M, N), rn(M, N))
    if kind == "ndm":
        assert not complex_weights
        A = int((alpha if alpha_a is None else alpha_a) * N)
        return NDM(rn(N), rn(M), rn(M, N), rn(A, N), rn(N), rn(M), rn(A), rn(M, N), rn(A, N), act)
    raise ValueError(kind)


# ----- NDMSymm --------------------------------------------------------------------------
class NDMSymm:
    """NDMSymm.jl:3-25: a symmetric NDM (alpha_h, alpha_a FEATURES) expanded into a bare NDM with alpha * n_symm units
    by the site permutations; gradients of the bare net are multiplied by the dense 0/1 matrices of
    construct_grad_matrices (NDMSymm.jl:130-181, ones/n for the local biases), field by field
    (symmetrize_grad_NDM!, NDMSymm.jl:65-76).  `permutations`: n_symm lists of N 1-based sites."""
    kind = "ndmsymm"
    doubled = True
    is_complex = False

    def __init__(self, symm_net, permutations):
        self.symm = symm_net
        self.perms = [list(p) for p in permutations]
        ns, N = len(self.perms), symm_net.N
        Ms, As = symm_net.M, symm_net.A
        z = np.zeros
        self.bare = NDM(z(N), z(Ms * ns), z((Ms * ns, N)), z((As * ns, N)), z(N), z(Ms * ns), z(As * ns), z((Ms * ns, N)),
                        z((As * ns, N)), symm_net.act)
        self.N, self.act = N, symm_net.act
        self._matrices()
        self.set_bare_params()

    @property
    def P(self):
        return self.symm.P

    def params(self):
        return self.symm.params()

    def set_params(self, w):
        self.symm.set_params(w)
        self.set_bare_params()

    def set_bare_params(self):                          # NDMSymm.jl:79-128
        s, b, ns = self.symm, self.bare, len(self.perms)
        s.b_mu = np.full_like(s.b_mu, s.b_mu.sum() / len(s.b_mu))
        s.b_lam = np.full_like(s.b_lam, s.b_lam.sum() / len(s.b_lam))
        b.b_mu, b.b_lam = s.b_mu.copy(), s.b_lam.copy()
        for i in range(s.M):
            b.h_mu[i * ns:(i + 1) * ns] = s.h_mu[i]
            b.h_lam[i * ns:(i + 1) * ns] = s.h_lam[i]
        for i in range(s.A):
            b.d_lam[i * ns:(i + 1) * ns] = s.d_lam[i]
        for f in range(s.M):
            for j, perm in enumerate(self.perms):
                for i, ip in enumerate(perm):
                    b.w_mu[j + f * ns, ip - 1] = s.w_mu[f, i]
                    b.w_lam[j + f * ns, ip - 1] = s.w_lam[f, i]
        for f in range(s.A):
            for j, perm in enumerate(self.perms):
                for i, ip in enumerate(perm):
                    b.u_mu[j + f * ns, ip - 1] = s.u_mu[f, i]
                    b.u_lam[j + f * ns, ip - 1] = s.u_lam[f, i]

    def _matrices(self):                                # NDMSymm.jl:130-181
        s, ns, N = self.symm, len(self.perms), self.N
        self.Gb = np.ones((N, N)) / N
        self.Gh = np.zeros((s.M, s.M * ns))
        for k in range(s.M):
            self.Gh[k, k * ns:(k + 1) * ns] = 1
        self.Gd = np.zeros((s.A, s.A * ns))
        for k in range(s.A):
            self.Gd[k, k * ns:(k + 1) * ns] = 1

        def wmat(K):
            G = np.zeros((K * N, K * ns * N))
            for f in range(K):
                for j, perm in enumerate(self.perms):
                    for i, ip in enumerate(perm):
                        G[f + K * i, (j + f * ns) + K * ns * (ip - 1)] = 1
            return G
        self.Gw, self.Gu = wmat(s.M), wmat(s.A)

    def logpsi(self, sr, sc):
        return self.bare.logpsi(sr, sc)

    def logpsi_grad(self, sr, sc):                      # NDMSymmBatched.jl:16-36
        out, Ob = self.bare.logpsi_grad(sr, sc)
        N, Mb, Ab = self.N, self.bare.M, self.bare.A
        sizes = [N, Mb, Mb * N, Ab * N, N, Mb, Ab, Mb * N, Ab * N]
        mats = [self.Gb, self.Gh, self.Gw, self.Gu, self.Gb, self.Gh, self.Gd, self.Gw, self.Gu]
        blocks, o = [], 0
        for sz, G in zip(sizes, mats):
            blocks.append(G @ Ob[o:o + sz])
            o += sz
        return out, np.concatenate(blocks, axis=0)


def random_ndmsymm(N, alpha_h, alpha_a, permutations, act=SOFTPLUS, seed=1234, std=0.1):
    rng = np.random.Generator(np.random.Philox(seed))
    rn = lambda *shape: rng.standard_normal(shape) * std
    Ms, As = alpha_h, alpha_a
    return NDMSymm(NDM(rn(N), rn(Ms), rn(Ms, N), rn(As, N), rn(N), rn(Ms), rn(As), rn(Ms, N), rn(As, N), act), permutations)
