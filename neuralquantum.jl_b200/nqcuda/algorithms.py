"""Local estimators, SR and optimisers over libnqcuda (host mirror of the reference interfaces).

ref: src/IterativeInterface/Accumulators/*.jl (local estimators), src/Algorithms/SR/{SR,SRDirect,SRIterative}.jl
     (SR, setup_algorithm!, precondition!), src/Optimisers/{rules,apply}.jl (Descent, update!),
     src/utils/stats.jl (Measurement, stat_analysis).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L

sr_cholesky, sr_cg, sr_minres = "sr_cholesky", "sr_cg", "sr_minres"
sr_shift = "sr_shift"


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
    g = np.zeros((net.P, B), dtype=net.cdtype, order="F")
    L.check(L.lib.nq_local_grad(net.h, op.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype), B, None, L.ptr(out), L.ptr(g),
                                net.P), net.ctx.h)
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
    """SR(T=Float32; eps=0.001, precision=1e-4, algorithm=sr_cholesky, full_matrix=false).
    eps and precision are stored in precision T exactly like the reference (quirk Q16)."""

    def __init__(self, T=np.float32, eps=0.001, precision=10e-5, algorithm=sr_cholesky, full_matrix=False,
                 precondition_type=sr_shift):
        if precondition_type != sr_shift:
            raise NotImplementedError("only sr_shift is on the built path (SURVEY 8f)")
        if algorithm not in (sr_cholesky, sr_cg, sr_minres):
            raise NotImplementedError("sr_cholesky, sr_cg and sr_minres are on the built path (SURVEY 8f)")
        self.sr_diag_shift = float(np.dtype(T).type(eps))
        self.sr_precision = float(np.dtype(T).type(precision))
        self.algorithm, self.full_matrix = algorithm, full_matrix


class Descent:
    """Optimisers.Descent(eta): w <- w - eta dw."""

    def __init__(self, eta=0.1):
        self.eta = eta


    def step(self, dw, state=None):
        """(delta, eta) such that w <- w - eta * delta."""
        return dw, self.eta


class Nesterov:
    """Optimisers.Nesterov(lr, mu) (rules.jl:36-55): v is the velocity kept per parameter vector,
        d = mu^2 v - (1 + mu) lr dw;   v <- mu v - lr dw;   w <- w + d.
    The velocity lives with the optimiser object (the reference keys an IdDict by the parameter array)."""

    def __init__(self, lr=0.1, mu=0.9, gclip=0.0):
        self.lr, self.mu, self.gclip = float(lr), float(mu), float(gclip)
        self.eta = 1.0
        self.v = None

    def step(self, dw, state=None):
        if self.v is None or self.v.shape != dw.shape:
            self.v = dw * 0
        d = (self.mu ** 2) * self.v - (1.0 + self.mu) * self.lr * dw
        self.v = self.mu * self.v - self.lr * dw
        return -d, 1.0                      # w <- w - 1 * (-d)


def update_(opt, net, dw):
    delta, eta = opt.step(np.asarray(dw))
    net.update(delta, eta)
