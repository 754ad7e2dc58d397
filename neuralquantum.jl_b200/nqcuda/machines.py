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


def init_random_pars_(net, sigma=0.01, seed=1234):
    """init_random_pars!(net, sigma): every parameter ~ N(0, sigma) (std sqrt(sigma)); utils/rng.jl:27-32."""
    rng = np.random.Generator(np.random.Philox(seed))
    if net.dtype.kind == "c":
        w = (rng.standard_normal(net.P) + 1j * rng.standard_normal(net.P)) * np.sqrt(sigma / 2.0)
    else:
        w = rng.standard_normal(net.P) * np.sqrt(sigma)
    net.set_params(w.astype(net.dtype))
    return net
