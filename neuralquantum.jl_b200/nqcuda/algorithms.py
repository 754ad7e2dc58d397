"""Local estimators, SR and optimisers over libnqcuda (host mirror of the reference interfaces).

ref: src/IterativeInterface/Accumulators/*.jl (local estimators), src/Algorithms/SR/{SR,SRDirect,SRIterative}.jl
     (SR, setup_algorithm!, precondition!), src/Optimisers/{rules,apply}.jl (Descent, update!),
     src/utils/stats.jl (Measurement, stat_analysis).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L

sr_cholesky, sr_cg, sr_minres, sr_qlp = "sr_cholesky", "sr_cg", "sr_minres", "sr_qlp"
sr_shift, sr_multiplicative, sr_none = "sr_shift", "sr_multiplicative", "sr_none"


def local_scalar(net, op, sigma):
    """E_loc / L_loc for every configuration (AccumulatorObsScalar semantics)."""
    sr, sc, B = net._states(sigma)
    out = np.zeros(B, dtype=net.cdtype)
    L.check(L.lib.nq_local_scalar(net.h, op.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype), B, None, L.ptr(out)), net.ctx.h)
    return out


def local_grad(net, op, sigma):
    """(L_loc [B], grad L_loc [P, B]) for a Liouvillian (AccumulatorObsGrad semantics)."""
    sr, sc, B = net._states(sigma)
    out = np.zeros(B, dtype=net.cdtype)
    Pk = getattr(net, "Pb", net.P)                   # symmetrised machines: the kernel writes rows of the bare net
    g = np.zeros((Pk, B), dtype=net.cdtype, order="F")
    L.check(L.lib.nq_local_grad(net.h, op.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype), B, None, L.ptr(out), L.ptr(g),
                                Pk), net.ctx.h)
    if Pk != net.P:
        gs = np.zeros((net.P, B), dtype=net.cdtype, order="F")
        net.symmetrize(L.ptr(g), Pk, B, L.ptr(gs), net.P)
        g = gs
    return out, g


@dataclass
class Measurement:
    mean: complex
    error: float
    variance: float
    tau: float
    R: float

    def __str__(self):
        return "(%.4f%+.4fim) ± %.4f [var=%.4f, tau=%.4f, R=%.4f]" % (
            self.mean.real, self.mean.imag, self.error, self.variance, self.tau, self.R)


def stat_analysis(ctx, vals):
    """vals [B chains, L] (host numpy or (device_ptr, B, L, np dtype))."""
    if isinstance(vals, tuple):
        p, B, Lc, dt = vals
    else:
        vals = np.asfortranarray(vals)
        B, Lc = vals.shape
        p, dt = L.ptr(vals), vals.dtype
    out = (C.c_double * 6)()
    L.check(L.lib.nq_stat_analysis(ctx.h, p, B, Lc, L.nq_dtype(dt), out), ctx.h)
    return Measurement(complex(out[0], out[1]), out[2], out[3], out[4], out[5])


class SR:
    """SR(T=Float32; eps=0.001, precision=1e-4, precondition_type=sr_shift, algorithm=sr_cholesky, full_matrix=false,
    lambda0=100, b=0.95, lambda_min=1e-4)  (SR/SR.jl:37-57).  eps, precision and the multiplicative-regulariser
    constants are stored in precision T exactly like the reference (quirk Q16).
    precondition_type: sr_shift (S + eps I), sr_multiplicative (S + lambda Diagonal(diag S), lambda = max(lambda0 b^iter,
    lambda_min); explicit S only) or sr_none.  algorithm: sr_cholesky | sr_cg | sr_minres | sr_qlp."""

    def __init__(self, T=np.float32, eps=0.001, precision=10e-5, algorithm=sr_cholesky, full_matrix=False,
                 precondition_type=sr_shift, lambda0=100.0, b=0.95, lambda_min=1e-4, maxiter=None):
        if precondition_type not in (sr_shift, sr_multiplicative, sr_none):
            raise ValueError("precondition_type: sr_shift, sr_multiplicative or sr_none")
        if algorithm not in (sr_cholesky, sr_cg, sr_minres, sr_qlp):
            raise NotImplementedError("sr_cholesky, sr_cg, sr_minres and sr_qlp are built (sr_diag / sr_div / sr_lsq: SURVEY 8f)")
        if precondition_type == sr_multiplicative and not (algorithm == sr_cholesky or full_matrix):
            raise ValueError("sr_multiplicative needs the explicit S (sr_cholesky or full_matrix=True): diag(S) of the "
                             "matrix-free operator is not defined in the reference either (SR_notfull.jl)")
        t = np.dtype(T).type
        self.sr_diag_shift = float(t(eps))
        self.sr_precision = float(t(precision))
        self.lambda0, self.b, self.lambda_min = float(t(lambda0)), float(t(b)), float(t(lambda_min))
        self.algorithm, self.full_matrix, self.precondition_type = algorithm, full_matrix, precondition_type
        self.maxiter = 0 if maxiter is None else int(maxiter)      # 0 = the reference's fixed 10 P (SRIterative.jl:95)

    def regulariser(self, iter_n):
        """(eps, lambda) applied at iteration iter_n: S + eps I + lambda Diagonal(diag S)."""
        if self.precondition_type == sr_shift:
            return self.sr_diag_shift, 0.0
        if self.precondition_type == sr_multiplicative:
            return 0.0, max(self.lambda0 * self.b ** iter_n, self.lambda_min)
        return 0.0, 0.0


class Descent:
    """Optimisers.Descent(eta): w <- w - eta dw."""

    def __init__(self, eta=0.1):
        self.eta = eta


    def step(self, dw, state=None, ctx=None):
        """(delta, eta) such that w <- w - eta * delta."""
        return dw, self.eta


class Nesterov:
    """Optimisers.Nesterov(lr, mu) (rules.jl:36-55): v is the velocity kept per parameter vector,
        d = mu^2 v - (1 + mu) lr dw;   v <- mu v - lr dw;   w <- w + d.
    The velocity lives with the optimiser object (the reference keys an IdDict by the parameter array) as a device vector;
    the step is one libnqcuda kernel (nq_nesterov)."""

    def __init__(self, lr=0.1, mu=0.9, gclip=0.0):
        self.lr, self.mu, self.gclip = float(lr), float(mu), float(gclip)
        self.eta = 1.0
        self.v = None
        self.delta = None

    def step(self, dw, state=None, ctx=None):
        """dw: device tensor.  Returns (delta, 1.0) with w <- w - delta."""
        import torch
        if self.v is None or self.v.shape != dw.shape or self.v.dtype != dw.dtype:
            self.v = torch.zeros_like(dw)
            self.delta = torch.zeros_like(dw)
        dw = dw.contiguous()
        code = {torch.float32: L.NQ_F32, torch.float64: L.NQ_F64, torch.complex64: L.NQ_C64, torch.complex128: L.NQ_C128}[dw.dtype]
        L.check(L.lib.nq_nesterov(ctx.h, self.v.data_ptr(), dw.data_ptr(), dw.numel(), code, self.lr, self.mu,
                                  self.delta.data_ptr()), ctx.h)
        return self.delta, 1.0


def update_(opt, net, dw):
    """Optimisers.update!(opt, net, dw) with a host or device dw."""
    if isinstance(opt, Nesterov) and isinstance(dw, np.ndarray):
        import torch
        with torch.cuda.stream(net.ctx.torch_stream()):
            dw = torch.as_tensor(np.ascontiguousarray(dw, dtype=net.dtype), device=torch.device("cuda", net.ctx.device))
    delta, eta = opt.step(dw, ctx=net.ctx)
    net.update(delta, eta)
