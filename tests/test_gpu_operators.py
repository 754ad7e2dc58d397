"""GPU parity: connection enumeration (bit-exact) and local estimators (K5) vs the oracle."""
import numpy as np
import pytest

import helpers as H
from oracle import estimators as OE
from oracle import machines as OM
from oracle.models import lindblad_ising_1d, tfim_1d, tfim_2d

pytestmark = pytest.mark.gpu


def _mask(words):
    return sum(int(w) << (64 * i) for i, w in enumerate(words))


def test_connections_ket_bit_exact(nq, ctx):
    N = 12
    oh, oH = tfim_1d(N, 0.9, 1.1)
    ph, pH = H.p_tfim_1d(nq, N, 0.9, 1.1)
    dop = pH.to_device(ctx)
    S = H.rand_states("spin", N, 40, 3)
    counts, mels, fr, fc = dop.connections(S)
    assert fc is None
    for b in range(S.shape[1]):
        ref = OE.connection_list_ket(oH, S[:, b])
        assert counts[b] == len(ref)
        got = [(complex(mels[b, c]), _mask(fr[b, c])) for c in range(counts[b])]
        assert got == ref                      # mel bit patterns and flip masks, in reference order


@pytest.mark.parametrize("fock", [True, False])
def test_connections_liouvillian_bit_exact(nq, ctx, fock):
    N = 8
    _, _, _, ol = lindblad_ising_1d(N, 0.4, 2.0, fock=fock)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N, 0.4, 2.0, fock=fock)
    dop = pl.to_device(ctx)
    kind = "fock" if fock else "spin"
    R, Cc = H.rand_states(kind, N, 50, 1), H.rand_states(kind, N, 50, 2)
    counts, mels, fr, fc = dop.connections(R, Cc)
    for b in range(50):
        ref = OE.connection_list_super(ol, R[:, b], Cc[:, b])
        assert counts[b] == len(ref) <= dop.max_connections
        got = [(complex(mels[b, c]), _mask(fr[b, c]), _mask(fc[b, c])) for c in range(counts[b])]
        assert got == ref


def test_connections_multiword(nq, ctx):
    N = 70
    oh, oH = tfim_1d(N)
    ph, pH = H.p_tfim_1d(nq, N)
    dop = pH.to_device(ctx)
    S = H.rand_states("spin", N, 3, 9)
    counts, mels, fr, _ = dop.connections(S)
    for b in range(3):
        ref = OE.connection_list_ket(oH, S[:, b])
        assert [(complex(mels[b, c]), _mask(fr[b, c])) for c in range(counts[b])] == ref


@pytest.mark.parametrize("dtype,act,N,alpha", [
    (np.complex128, OM.LOGCOSH, 10, 2), (np.float64, OM.LOGCOSH, 10, 2), (np.complex128, OM.SOFTPLUS, 6, 3),
    (np.complex64, OM.LOGCOSH, 10, 2), (np.float32, OM.LOGCOSH, 12, 1), (np.complex128, OM.LOGCOSH, 20, 32)])
def test_local_energy_rbm(nq, ctx, dtype, act, N, alpha):
    oh, oH = tfim_1d(N)
    ph, pH = H.p_tfim_1d(nq, N)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, alpha, dtype, act)
    dop = pH.to_device(ctx)
    S = H.rand_states("spin", N, 33, 4321)
    ref = OE.local_scalar_ket(om, oH, S)
    got = nq.local_scalar(pm, dop, S)
    H.assert_close(got, ref, H.TOL[np.dtype(dtype)], "E_loc")


def test_local_energy_2d(nq, ctx):
    oh, oH = tfim_2d(4)
    ph, pH = H.p_tfim_2d(nq, 4)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", 16, 4, np.complex128, OM.LOGCOSH)
    S = H.rand_states("spin", 16, 21, 5)
    H.assert_close(nq.local_scalar(pm, pH.to_device(ctx), S), OE.local_scalar_ket(om, oH, S), 1e-11, "E_loc 2D")


@pytest.mark.parametrize("kind,dtype,act,N,alpha,hk", [
    ("ndm", np.float64, OM.SOFTPLUS, 8, 2, "fock"), ("ndm", np.float64, OM.LOGCOSH, 6, 1, "spin"),
    ("ndm", np.float32, OM.SOFTPLUS, 8, 2, "fock"), ("rbmsplit", np.complex128, OM.SOFTPLUS, 6, 2, "fock"),
    ("rbmsplit", np.float64, OM.SOFTPLUS, 5, 2, "fock"), ("ndm", np.float64, OM.SOFTPLUS, 16, 2, "fock")])
def test_local_liouvillian_value_and_gradient(nq, ctx, kind, dtype, act, N, alpha, hk):
    _, _, _, ol = lindblad_ising_1d(N, 0.4, 2.0, fock=(hk == "fock"))
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N, 0.4, 2.0, fock=(hk == "fock"))
    om, pm, hilb = H.make_pair(nq, ctx, kind, hk, N, alpha, dtype, act)
    dop = pl.to_device(ctx)
    B = 17
    R, Cc = H.rand_states(hk, N, B, 31), H.rand_states(hk, N, B, 32)
    Cc[:, :3] = R[:, :3]                                   # include diagonal configurations
    ref_l, ref_g = OE.local_grad_super(om, ol, R, Cc)
    tol = H.TOL[np.dtype(dtype)]
    loc, g = nq.local_grad(pm, dop, (R, Cc))
    H.assert_close(loc, ref_l, tol, "L_loc")
    H.assert_close(g, ref_g, tol, "grad L_loc")
    H.assert_close(nq.local_scalar(pm, dop, (R, Cc)), OE.local_scalar_super(om, ol, R, Cc), tol, "L_loc (scalar)")


@pytest.mark.parametrize("N,alpha,alpha_a,act,hk", [
    (44, 0.5, 0.5, OM.SOFTPLUS, "fock"),      # 3 N = 132 (site, pattern) weights > 128 threads per configuration
    (12, 3, 1, OM.SOFTPLUS, "fock"),          # M = 36 > 32 hidden units per layer (two passes per lane), A = 12 != M
    (10, 1, 4, OM.LOGCOSH, "spin"),           # A = 40 > M = 10
])
def test_local_liouvillian_odd_shapes(nq, ctx, N, alpha, alpha_a, act, hk):
    """Shapes that leave the one-hidden-unit-per-lane / one-thread-per-weight sweet spot of the fused NDM kernel."""
    _, _, _, ol = lindblad_ising_1d(N, 0.4, 2.0, fock=(hk == "fock"))
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N, 0.4, 2.0, fock=(hk == "fock"))
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", hk, N, alpha, np.float64, act, alpha_a=alpha_a)
    dop = pl.to_device(ctx)
    B = 6
    R, Cc = H.rand_states(hk, N, B, 41), H.rand_states(hk, N, B, 42)
    Cc[:, :2] = R[:, :2]
    ref_l, ref_g = OE.local_grad_super(om, ol, R, Cc)
    loc, g = nq.local_grad(pm, dop, (R, Cc))
    H.assert_close(loc, ref_l, 1e-11, "L_loc")
    H.assert_close(g, ref_g, 1e-11, "grad L_loc")
    ref_lp, ref_O = om.logpsi_grad(R, Cc)
    # fused entry point (eval + grad + estimator in one launch) on the same configurations
    L = nq._lib
    import torch
    W = L.lib.nq_states_words(N)
    prow = torch.zeros((B, W), dtype=torch.int64, device="cuda")
    pcol = torch.zeros_like(prow)
    for arr, buf in ((R, prow), (Cc, pcol)):
        a = np.asfortranarray(arr)
        L.check(L.lib.nq_pack_states(ctx.h, hilb.code, N, B, L.ptr(a), L.nq_dtype(a.dtype), buf.data_ptr()), ctx.h)
    lp = torch.zeros(B, dtype=torch.complex128, device="cuda")
    O = torch.zeros((B, pm.P), dtype=torch.complex128, device="cuda")
    lo = torch.zeros(B, dtype=torch.complex128, device="cuda")
    gl = torch.zeros((B, pm.P), dtype=torch.complex128, device="cuda")
    L.check(L.lib.nq_logpsi_grad_local_packed(pm.h, dop.h, prow.data_ptr(), pcol.data_ptr(), B, lp.data_ptr(), O.data_ptr(),
                                              pm.P, lo.data_ptr(), gl.data_ptr(), pm.P), ctx.h)
    H.assert_close(lp.cpu().numpy(), ref_lp, 1e-11, "log rho (fused)")
    H.assert_close(O.cpu().numpy().T, ref_O, 1e-11, "O (fused)")
    H.assert_close(lo.cpu().numpy(), ref_l, 1e-11, "L_loc (fused)")
    H.assert_close(gl.cpu().numpy().T, ref_g, 1e-11, "grad L_loc (fused)")


def test_estimator_argument_errors(nq, ctx):
    _, pH = H.p_tfim_1d(nq, 6)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, 6)
    om, rbm, _ = H.make_pair(nq, ctx, "rbm", "spin", 6, 1, np.float64, OM.LOGCOSH)
    S = H.rand_states("spin", 6, 4, 1)
    with pytest.raises(nq.NQError):            # Liouvillian with a ket machine
        nq.local_scalar(rbm, pl.to_device(ctx), S)
    _, pH5 = H.p_tfim_1d(nq, 5)
    with pytest.raises(nq.NQError):            # site-count mismatch
        nq.local_scalar(rbm, pH5.to_device(ctx), S)


def _xx_models(nq, N, fock):
    """XX + Z chain with sigma^- jumps: connections flip TWO sites (generic, non site-local path)."""
    from oracle import operators as OO
    from oracle.hilbert import HomogeneousFock, HomogeneousSpin
    oh = HomogeneousFock(N) if fock else HomogeneousSpin(N)
    ph = nq.HomogeneousFock(N) if fock else nq.HomogeneousSpin(N)
    oHm, pHm, oj, pj = None, nq.LocalOperator(ph), [], []
    for i in range(1, N):
        oHm = OO.add(oHm, OO.mul(OO.scale(0.7, OO.sigmax(oh, i)), OO.sigmax(oh, i + 1)))
        pHm = pHm + (0.7 * nq.sigmax(ph, i)) * nq.sigmax(ph, i + 1)
    for i in range(1, N + 1):
        oHm = OO.add(oHm, OO.scale(0.3, OO.sigmaz(oh, i)))
        pHm = pHm + 0.3 * nq.sigmaz(ph, i)
        oHm = OO.add(oHm, OO.scale(0.2, OO.sigmay(oh, i)))
        pHm = pHm + 0.2 * nq.sigmay(ph, i)
        oj.append(OO.sigmam(oh, i))
        pj.append(nq.sigmam(ph, i))
    return oh, oHm, OO.liouvillian(oHm, oj), ph, pHm, nq.liouvillian(pHm, pj)


def test_generic_two_site_flips(nq, ctx):
    N = 6
    oh, oHm, ol, ph, pHm, pl = _xx_models(nq, N, fock=False)
    # ket
    om, pm, _ = H.make_pair(nq, ctx, "rbm", "spin", N, 2, np.complex128, OM.LOGCOSH)
    S = H.rand_states("spin", N, 25, 3)
    dH = pHm.to_device(ctx)
    counts, mels, fr, _ = dH.connections(S)
    for b in range(5):
        ref = OE.connection_list_ket(oHm, S[:, b])
        assert [(complex(mels[b, c]), _mask(fr[b, c])) for c in range(counts[b])] == ref
    H.assert_close(nq.local_scalar(pm, dH, S), OE.local_scalar_ket(om, oHm, S), 1e-11, "E_loc XX")
    # Liouvillian with gradient (shared-memory atomics path)
    for kind, dtype in (("ndm", np.float64), ("rbmsplit", np.complex128)):
        om, pm, _ = H.make_pair(nq, ctx, kind, "spin", N, 2, dtype, OM.SOFTPLUS)
        R, Cc = H.rand_states("spin", N, 11, 41), H.rand_states("spin", N, 11, 42)
        ref_l, ref_g = OE.local_grad_super(om, ol, R, Cc)
        loc, g = nq.local_grad(pm, pl.to_device(ctx), (R, Cc))
        H.assert_close(loc, ref_l, 1e-11, "L_loc XX " + kind)
        H.assert_close(g, ref_g, 1e-11, "grad L_loc XX " + kind)
