"""Machines: host mirror of the reference's network types over libnqcuda handles.

ref: src/Networks/ClosedSystems/RBM.jl, src/Networks/MixedDensityMatrix/{RBMSplit,NDM}.jl,
     src/base_batched_networks.jl:20-145 (cached / logpsi! / logpsi_and_grad! / grad_cache),
     src/utils/rng.jl:20-32 (init_random_pars!), src/Networks/utils.jl:7-9 (rescaled_normal).
Parameters live on the device; `params` / `set_params` move the flat functor-ordered vector.
"""
import ctypes as C

import numpy as np

from . import _lib as L

af_softplus = L.NQ_SOFTPLUS
af_logcosh = L.NQ_LOGCOSH


def _rescaled_normal(rng, dtype, scale, *dims):
    """rescaled_normal(T, scale, dims...) = randn(T, dims) * scale * sqrt(24/sum(dims))."""
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        x = (rng.standard_normal(dims) + 1j * rng.standard_normal(dims)) / np.sqrt(2.0)
    else:
        x = rng.standard_normal(dims)
    return (x * scale * np.sqrt(24.0 / sum(dims))).astype(dtype)


class Machine:
    kind = None
    fields = ()

    def __init__(self, ctx, hilb, dtype, N, M, A, act, shapes, init):
        self.ctx, self.hilb = ctx, hilb
        self.dtype = np.dtype(dtype)
        self.N, self.M, self.A, self.act = int(N), int(M), int(A), act
        self.shapes = shapes
        h = C.c_void_p()
        L.check(L.lib.nq_machine_create(ctx.h, self.kind, hilb.code, self.N, self.M, self.A, act,
                                        L.nq_dtype(self.dtype), C.byref(h)), ctx.h)
        self.h = h
        n = C.c_int64()
        L.check(L.lib.nq_machine_nparams(h, C.byref(n)), ctx.h)
        self.P = n.value
        od = C.c_int()
        L.check(L.lib.nq_machine_out_dtype(h, C.byref(od)), ctx.h)
        self.out_code = od.value
        self.out_dtype = np.dtype(L.NP_OF[od.value])            # out_type(net)
        self.cdtype = np.dtype(L.NP_OF[L.complex_of(od.value)])  # complex of the same precision
        self.rdtype = np.dtype(L.NP_OF[L.real_of(od.value)])
        self.set_params(np.concatenate([np.asarray(a).reshape(-1, order="F") for a in init]))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                L.lib.nq_machine_destroy(self.h)
            self.h = None
        except Exception:
            pass

    # ---- plugin interface ----
    is_analytic = True
    doubled = True

    @property
    def real_params(self):
        """trainable parameters are real (S = Re(O O')/N, F = Re(grad C); SRDirect.jl:36-38)."""
        return self.dtype.kind != "c"

    def params(self):
        w = np.zeros(self.P, dtype=self.dtype)
        L.check(L.lib.nq_machine_get_params(self.h, L.ptr(w), self.P), self.ctx.h)
        return w

    def set_params(self, w):
        w = np.ascontiguousarray(w, dtype=self.dtype)
        L.check(L.lib.nq_machine_set_params(self.h, L.ptr(w), w.size), self.ctx.h)

    def named_params(self):
        w, out, o = self.params(), {}, 0
        for name, shp in zip(self.fields, self.shapes):
            n = int(np.prod(shp))
            out[name] = w[o:o + n].reshape(shp, order="F")
            o += n
        return out

    def _states(self, sigma):
        if self.doubled:
            sr, sc = sigma
            sr = np.asfortranarray(sr)
            sc = np.asfortranarray(sc, dtype=sr.dtype)
            return sr, sc, sr.shape[1]
        s = np.asfortranarray(sigma)
        return s, None, s.shape[1]

    def logpsi(self, sigma, out=None):
        """logpsi!(out, net, cache, sigma...): sigma [N,B] (ket) or (sigma, sigma') pair."""
        sr, sc, B = self._states(sigma)
        if out is None:
            out = np.zeros(B, dtype=self.out_dtype)
        L.check(L.lib.nq_logpsi(self.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype), B, L.ptr(out)), self.ctx.h)
        return out

    def log_prob(self, sigma):
        sr, sc, B = self._states(sigma)
        out = np.zeros(B, dtype=self.rdtype)
        L.check(L.lib.nq_log_prob(self.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype), B, L.ptr(out)), self.ctx.h)
        return out

    def logpsi_and_grad(self, sigma, out=None, grad=None):
        """logpsi_and_grad!(grad, out, net, cache, sigma...): grad is ONE [P, B] buffer, parameter fastest."""
        sr, sc, B = self._states(sigma)
        if out is None:
            out = np.zeros(B, dtype=self.out_dtype)
        if grad is None:
            grad = np.zeros((self.P, B), dtype=self.out_dtype, order="F")
        L.check(L.lib.nq_logpsi_grad(self.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype), B, L.ptr(out), L.ptr(grad),
                                     self.P), self.ctx.h)
        return out, grad

    def grad_views(self, grad):
        """Named views into the flat gradient buffer (RealDerivative, AD/RealDerivatives.jl:1-25)."""
        out, o = {}, 0
        for name, shp in zip(self.fields, self.shapes):
            n = int(np.prod(shp))
            out[name] = grad[o:o + n].reshape(tuple(shp) + grad.shape[1:], order="F")
            o += n
        return out

    def update(self, dw, eta):
        """Optimisers.update!(Descent(eta), net, dw): w <- w - eta dw on the device."""
        dw = np.ascontiguousarray(dw, dtype=self.dtype) if isinstance(dw, np.ndarray) else dw
        L.check(L.lib.nq_update(self.h, L.ptr(dw), float(eta)), self.ctx.h)


class RBM(Machine):
    kind = L.NQ_RBM
    fields = ("a", "b", "W")
    doubled = False

    def __init__(self, ctx, hilb, dtype, alpha, act=af_softplus, seed=0):
        N = hilb.n
        M = int(alpha * N)
        rng = np.random.Generator(np.random.Philox(seed))
        init = [_rescaled_normal(rng, dtype, 0.01, N), _rescaled_normal(rng, dtype, 0.01, M),
                _rescaled_normal(rng, dtype, 0.01, M, N)]
        super().__init__(ctx, hilb, dtype, N, M, 0, act, [(N,), (M,), (M, N)], init)


class RBMSplit(Machine):
    kind = L.NQ_RBMSPLIT
    fields = ("ar", "ac", "b", "Wr", "Wc")

    def __init__(self, ctx, hilb, dtype, alpha, seed=0):
        N = hilb.n
        M = int(alpha * N)
        rng = np.random.Generator(np.random.Philox(seed))
        init = [_rescaled_normal(rng, dtype, 0.01, N), _rescaled_normal(rng, dtype, 0.01, N),
                _rescaled_normal(rng, dtype, 0.005, M), _rescaled_normal(rng, dtype, 0.01, M, N),
                _rescaled_normal(rng, dtype, 0.01, M, N)]
        super().__init__(ctx, hilb, dtype, N, M, 0, af_softplus, [(N,), (N,), (M,), (M, N), (M, N)], init)


class NDM(Machine):
    kind = L.NQ_NDM
    fields = ("b_mu", "h_mu", "w_mu", "u_mu", "b_lam", "h_lam", "d_lam", "w_lam", "u_lam")

    def __init__(self, ctx, hilb, dtype, alpha_h, alpha_a, act=af_softplus, seed=0):
        dtype = np.dtype(dtype)
        if dtype.kind == "c":               # NDM(T<:Complex, ...) = NDM(real(T), ...)  NDM.jl:46-47
            dtype = np.dtype(np.float32 if dtype == np.complex64 else np.float64)
        N = hilb.n
        M, A = int(alpha_h * N), int(alpha_a * N)
        rng = np.random.Generator(np.random.Philox(seed))
        rn = lambda s, *d: _rescaled_normal(rng, dtype, s, *d)
        init = [rn(0.005, N), rn(0.005, M), rn(0.01, M, N), rn(0.01, A, N),
                rn(0.005, N), rn(0.005, M), rn(0.005, A), rn(0.01, M, N), rn(0.01, A, N)]
        shapes = [(N,), (M,), (M, N), (A, N), (N,), (M,), (A,), (M, N), (A, N)]
        super().__init__(ctx, hilb, dtype, N, M, A, act, shapes, init)


def _ndm_offsets(N, M, A):
    """Offsets of the NDM fields in the flat functor-ordered vector (NDM.jl:17-19)."""
    sizes = [N, M, M * N, A * N, N, M, A, M * N, A * N]
    return dict(zip(NDM.fields, np.concatenate([[0], np.cumsum(sizes)[:-1]]).tolist())), int(sum(sizes))


def symmetry_maps(N, alpha_h, alpha_a, permutations):
    """The two maps of an NDMSymm as index lists (0-based): gradient gather (ptr, idx, scale) = the rows of the 0/1
    matrices of construct_grad_matrices (NDMSymm.jl:130-181; 1/n for the local biases) and the parameter source `src` of
    set_bare_params! (NDMSymm.jl:79-128).  `permutations`: n_symm lists of N 1-based sites."""
    perms = [[int(x) - 1 for x in p] for p in permutations]
    ns = len(perms)
    assert all(sorted(p) == list(range(N)) for p in perms), "permutations of 1..N expected"
    Ms, As, Mb, Ab = alpha_h, alpha_a, alpha_h * ns, alpha_a * ns
    so, Ps = _ndm_offsets(N, Ms, As)
    bo, Pb = _ndm_offsets(N, Mb, Ab)
    lists = [[] for _ in range(Ps)]
    scale = np.ones(Ps)
    src = np.full(Pb, -1, dtype=np.int32)
    for f in ("b_mu", "b_lam"):
        for i in range(N):
            lists[so[f] + i] = [bo[f] + j for j in range(N)]
            scale[so[f] + i] = 1.0 / N
            src[bo[f] + i] = so[f] + i
    for f, K in (("h_mu", Ms), ("h_lam", Ms), ("d_lam", As)):
        for s_ in range(K):
            lists[so[f] + s_] = [bo[f] + s_ * ns + j for j in range(ns)]
            for j in range(ns):
                src[bo[f] + s_ * ns + j] = so[f] + s_
    for f, Ks, Kb in (("w_mu", Ms, Mb), ("w_lam", Ms, Mb), ("u_mu", As, Ab), ("u_lam", As, Ab)):
        for ft in range(Ks):
            for j, perm in enumerate(perms):
                for i, ip in enumerate(perm):
                    q = bo[f] + (j + ft * ns) + Kb * ip
                    p_ = so[f] + ft + Ks * i
                    lists[p_].append(q)
                    src[q] = p_
    assert (src >= 0).all()
    ptr = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    idx = np.array([q for l in lists for q in l], dtype=np.int32)
    avg = np.array([[so["b_mu"], so["b_mu"] + N], [so["b_lam"], so["b_lam"] + N]], dtype=np.int64)
    return dict(Ps=Ps, Pb=Pb, ptr=ptr, idx=idx, scale=scale, src=src, avg=avg, Ms=Ms, As=As, Mb=Mb, Ab=Ab)


class NDMSymm:
    """NDMSymm(T, hilb, alpha_h, alpha_a, permutations, act): an NDM whose weights are tied by the site permutations of a
    symmetry group (NDMSymm.jl:3-25).  alpha_h / alpha_a count FEATURES (hidden units of the symmetric net); the bare net
    that every kernel evaluates has alpha * n_symm units.  Trainable parameters, gradients, S and the update live in the
    symmetric space; `h` is the bare machine's handle (samplers and estimators take it as is)."""
    kind = L.NQ_NDM
    fields = NDM.fields
    doubled = True
    is_analytic = True
    real_params = True

    def __init__(self, ctx, hilb, dtype, alpha_h, alpha_a, permutations, act=af_softplus, seed=0):
        N = hilb.n
        mp = symmetry_maps(N, int(alpha_h), int(alpha_a), permutations)
        self.maps, self.permutations = mp, [list(p) for p in permutations]
        # NDM(T, n_in, alpha//n_in * n_symm, ...): Mb = alpha_h n_symm hidden units
        self.bare = NDM(ctx, hilb, dtype, mp["Mb"] / N, mp["Ab"] / N, act, seed=seed)
        assert self.bare.P == mp["Pb"] and self.bare.M == mp["Mb"] and self.bare.A == mp["Ab"]
        self.ctx, self.hilb, self.N, self.act = ctx, hilb, N, act
        self.M, self.A = self.bare.M, self.bare.A
        self.dtype, self.out_dtype, self.cdtype, self.rdtype = self.bare.dtype, self.bare.out_dtype, self.bare.cdtype, self.bare.rdtype
        self.out_code = self.bare.out_code
        self.P, self.Pb = mp["Ps"], mp["Pb"]
        self.shapes = [(N,), (mp["Ms"],), (mp["Ms"], N), (mp["As"], N), (N,), (mp["Ms"],), (mp["As"],), (mp["Ms"], N), (mp["As"], N)]
        g = C.c_void_p()
        L.check(L.lib.nq_symm_create(self.bare.h, self.P, L.ptr(mp["ptr"]), L.ptr(mp["idx"]), L.ptr(mp["scale"]),
                                     L.ptr(mp["src"]), 2, L.ptr(np.ascontiguousarray(mp["avg"])), C.byref(g)), ctx.h)
        self.g = g
        rng = np.random.Generator(np.random.Philox(seed))
        rn = lambda s_, *d: _rescaled_normal(rng, self.dtype, s_, *d)
        Ms, As = mp["Ms"], mp["As"]
        init = [rn(0.005, N), rn(0.005, Ms), rn(0.01, Ms, N), rn(0.01, As, N), rn(0.005, N), rn(0.005, Ms), rn(0.005, As),
                rn(0.01, Ms, N), rn(0.01, As, N)]
        self.set_params(np.concatenate([np.asarray(a).reshape(-1, order="F") for a in init]))

    @property
    def h(self):
        return self.bare.h

    def __del__(self):
        try:
            if self.g and self.ctx.h:
                L.lib.nq_symm_destroy(self.g)
            self.g = None
        except Exception:
            pass

    def params(self):
        w = np.zeros(self.P, dtype=self.dtype)
        L.check(L.lib.nq_symm_get_params(self.g, L.ptr(w), self.P), self.ctx.h)
        return w

    def set_params(self, w):
        w = np.ascontiguousarray(w, dtype=self.dtype)
        L.check(L.lib.nq_symm_set_params(self.g, L.ptr(w), w.size), self.ctx.h)

    named_params = Machine.named_params
    grad_views = Machine.grad_views
    _states = Machine._states

    def logpsi(self, sigma, out=None):
        return self.bare.logpsi(sigma, out)

    def log_prob(self, sigma):
        return self.bare.log_prob(sigma)

    def symmetrize(self, bare_ptr, ldb, Ns, out_ptr, lds):
        """symmetrize_grad_NDM_batched!: rows of a bare gradient buffer -> rows of the symmetric one (device or host)."""
        L.check(L.lib.nq_symm_gradient(self.g, bare_ptr, ldb, Ns, L.nq_dtype(self.out_dtype), out_ptr, lds), self.ctx.h)

    def logpsi_and_grad(self, sigma, out=None, grad=None):
        sr, sc, B = self._states(sigma)
        out, gb = self.bare.logpsi_and_grad(sigma, out)
        if grad is None:
            grad = np.zeros((self.P, B), dtype=self.out_dtype, order="F")
        self.symmetrize(L.ptr(gb), self.Pb, B, L.ptr(grad), self.P)
        return out, grad

    def update(self, dw, eta):
        dw = np.ascontiguousarray(dw, dtype=self.dtype) if isinstance(dw, np.ndarray) else dw
        L.check(L.lib.nq_symm_update(self.g, L.ptr(dw), float(eta)), self.ctx.h)


def fullspace_size(net):
    n = C.c_int64()
    L.check(L.lib.nq_fullspace_size(net.h, C.byref(n)), net.ctx.h)
    return n.value


def ket(net, hilb=None, norm=True):
    """ket(net, hilb, norm): exp(log psi) on all basis states (basis number = digits, site 1 least significant),
    2-normalised when `norm`.  A density-matrix machine gives vec(rho) like the reference's ket(::MatrixNet, ...),
    normalised as a vector.  ref: utils/densitymatrix.jl:40-68."""
    if net.doubled:
        v = densitymatrix(net, hilb, False).reshape(-1, order="F")
        return v / np.linalg.norm(v) if norm else v
    out = np.zeros(fullspace_size(net), dtype=net.out_dtype)
    L.check(L.lib.nq_fullspace_state(net.h, 1 if norm else 0, L.ptr(out)), net.ctx.h)
    return out


def densitymatrix(net, hilb=None, norm=True):
    """densitymatrix(net, hilb, norm): rho[i, j] = exp(log rho(row state i, column state j)), divided by its trace
    when `norm`.  ref: utils/densitymatrix.jl:9-38."""
    if not net.doubled:
        raise ValueError("densitymatrix needs a density-matrix machine (RBMSplit, NDM)")
    D = 1 << net.N
    out = np.zeros((D, D), dtype=net.out_dtype, order="F")
    L.check(L.lib.nq_fullspace_state(net.h, 1 if norm else 0, L.ptr(out)), net.ctx.h)
    return out


def init_random_pars_(net, sigma=0.01, seed=1234):
    """init_random_pars!(net, sigma): every parameter ~ N(0, sigma) (std sqrt(sigma)); utils/rng.jl:27-32."""
    rng = np.random.Generator(np.random.Philox(seed))
    if net.dtype.kind == "c":
        w = (rng.standard_normal(net.P) + 1j * rng.standard_normal(net.P)) * np.sqrt(sigma / 2.0)
    else:
        w = rng.standard_normal(net.P) * np.sqrt(sigma)
    net.set_params(w.astype(net.dtype))
    return net
