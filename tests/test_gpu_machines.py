"""GPU parity: fused eval+grad kernels (K1-K3) vs the oracle through the C ABI.
Tolerances (north_star): 1e-11 relative in FP64 mode, 1e-5 in FP32 mode (vs the FP64 oracle)."""
import numpy as np
import pytest

import helpers as H
from oracle import machines as OM

pytestmark = pytest.mark.gpu

CASES = [
    # kind, hilbert, N, alpha, dtype, act, B
    ("rbm", "spin", 10, 2, np.complex128, OM.LOGCOSH, 37),
    ("rbm", "spin", 10, 2, np.float64, OM.LOGCOSH, 8),
    ("rbm", "fock", 7, 3, np.complex128, OM.SOFTPLUS, 5),
    ("rbm", "spin", 36, 4, np.complex128, OM.LOGCOSH, 64),
    ("rbm", "spin", 36, 4, np.complex64, OM.LOGCOSH, 64),
    ("rbm", "spin", 20, 1, np.float32, OM.LOGCOSH, 33),
    ("rbm", "spin", 70, 5, np.float64, OM.SOFTPLUS, 11),      # N > 64 (two words), M = 350 > 256 (two chunks)
    ("rbmsplit", "fock", 6, 2, np.complex128, OM.SOFTPLUS, 19),
    ("rbmsplit", "fock", 6, 2, np.float64, OM.SOFTPLUS, 19),
    ("rbmsplit", "spin", 5, 3, np.complex64, OM.SOFTPLUS, 9),
    ("ndm", "fock", 8, 2, np.float64, OM.SOFTPLUS, 21),
    ("ndm", "fock", 16, 2, np.float64, OM.SOFTPLUS, 50),
    ("ndm", "spin", 7, 1, np.float64, OM.LOGCOSH, 13),
    ("ndm", "fock", 16, 2, np.float32, OM.SOFTPLUS, 50),
    ("ndm", "spin", 66, 4, np.float64, OM.SOFTPLUS, 6),       # multi-word, M = 264 > 256
    # corners of the cfg5 throughput sweep (N = 64 ... 256, alpha = 1 ... 8)
    ("rbm", "spin", 256, 1, np.complex128, OM.LOGCOSH, 9),    # four packed words
    ("rbm", "spin", 128, 8, np.complex64, OM.LOGCOSH, 5),     # M = 1024
    ("rbm", "spin", 64, 8, np.float64, OM.SOFTPLUS, 7),
    ("ndm", "fock", 128, 2, np.float64, OM.SOFTPLUS, 4),      # P = 131 968
    ("ndm", "fock", 256, 1, np.float32, OM.SOFTPLUS, 3),
    ("rbmsplit", "fock", 200, 2, np.complex128, OM.SOFTPLUS, 3),
]


@pytest.mark.parametrize("kind,hk,N,alpha,dtype,act,B", CASES)
def test_logpsi_and_grad_match_oracle(nq, ctx, kind, hk, N, alpha, dtype, act, B):
    om, pm, hilb = H.make_pair(nq, ctx, kind, hk, N, alpha, dtype, act)
    tol = H.TOL[np.dtype(dtype)]
    sr = H.rand_states(hk, N, B, 4321)
    sc = H.rand_states(hk, N, B, 4322)
    sigma = (sr, sc) if om.doubled else sr
    ref_out, ref_O = om.logpsi_grad(*sigma) if om.doubled else om.logpsi_grad(sigma)
    out, O = pm.logpsi_and_grad(sigma)
    assert out.dtype == pm.out_dtype and O.shape == (pm.P, B)
    # cfg5 corners in FP32 mode: theta sums N >= 128 FP32 terms (rounding ~1e-6) and complex tanh amplifies it by
    # 1 + |tanh|^2 near its poles, so single entries of O move by ~1e-4 of the largest one: element-wise floor max|ref|
    floor = 1.0 if (H.TOL[np.dtype(dtype)] > 1e-6 and N >= 128) else None
    H.assert_close(out, ref_out, tol, "logpsi", floor=floor)
    H.assert_close(O, ref_O, tol, "O", floor=floor)
    # value-only entry point and log-probability
    H.assert_close(pm.logpsi(sigma), ref_out, tol, "logpsi!")
    H.assert_close(pm.log_prob(sigma), OM.log_prob(ref_out), tol, "log_prob")
    # float32 configuration arrays are accepted too
    s32 = tuple(s.astype(np.float32) for s in sigma) if om.doubled else sigma.astype(np.float32)
    H.assert_close(pm.logpsi(s32), ref_out, tol, "logpsi! (Float32 states)")


def test_edge_cases(nq, ctx):
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", 10, 2, np.complex128, OM.LOGCOSH)
    # empty batch
    out, O = pm.logpsi_and_grad(np.zeros((10, 0), order="F"))
    assert out.shape == (0,) and O.shape == (pm.P, 0)
    # single configuration, batch not a multiple of the CTA tile
    s = H.rand_states("spin", 10, 1, 1)
    H.assert_close(pm.logpsi(s), om.logpsi(s), 1e-11, "B=1")
    # parameters round-trip through the device in functor order
    w = pm.params()
    assert np.array_equal(w, om.params())
    # leading dimension larger than P
    import ctypes as C
    B, ld = 6, pm.P + 5
    s = H.rand_states("spin", 10, B, 2)
    out = np.zeros(B, np.complex128)
    Obuf = np.full((ld, B), np.nan + 0j, order="F")
    L = nq._lib
    L.check(L.lib.nq_logpsi_grad(pm.h, L.ptr(s), None, L.NQ_F64, B, L.ptr(out), L.ptr(Obuf), ld), ctx.h)
    H.assert_close(Obuf[:pm.P], om.logpsi_grad(s)[1], 1e-11, "ldO > P")
    assert np.all(np.isnan(Obuf[pm.P:].real))
    # wrong arity is an error, not a crash
    with pytest.raises(nq.NQError):
        L.check(L.lib.nq_logpsi(pm.h, L.ptr(s), L.ptr(s), L.NQ_F64, B, L.ptr(out)), ctx.h)


def test_device_pointers_and_packed_path(nq, ctx):
    import torch
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", 16, 2, np.float64, OM.SOFTPLUS)
    B = 300
    sr, sc = H.rand_states("fock", 16, B, 10), H.rand_states("fock", 16, B, 11)
    pr, pc = ctx.pack(hilb, sr), ctx.pack(hilb, sc)
    assert np.array_equal(ctx.unpack(hilb, pr), sr)
    dpr = torch.from_numpy(pr.view(np.int64)).cuda()
    dpc = torch.from_numpy(pc.view(np.int64)).cuda()
    out = torch.zeros(B, dtype=torch.complex128, device="cuda")
    O = torch.zeros((B, pm.P), dtype=torch.complex128, device="cuda")
    L = nq._lib
    L.check(L.lib.nq_logpsi_grad_packed(pm.h, dpr.data_ptr(), dpc.data_ptr(), B, out.data_ptr(), O.data_ptr(), pm.P), ctx.h)
    torch.cuda.synchronize()
    ro, rO = om.logpsi_grad(sr, sc)
    H.assert_close(out.cpu().numpy(), ro, 1e-11, "packed logpsi")
    H.assert_close(O.cpu().numpy().T, rO, 1e-11, "packed O")
    assert ctx.launches > 0


def test_update_descent(nq, ctx):
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", 6, 2, np.complex128, OM.LOGCOSH)
    rng = np.random.default_rng(0)
    dw = (rng.standard_normal(pm.P) + 1j * rng.standard_normal(pm.P))
    w0 = pm.params()
    nq.update_(nq.Descent(0.1), pm, dw)
    H.assert_close(pm.params(), w0 - 0.1 * dw, 1e-15, "Descent")
