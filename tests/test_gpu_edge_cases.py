"""Edge cases through the C ABI: empty and single-element batches, padded leading dimensions, the largest supported
configuration width, burn-only sampler runs, argument errors reported as statuses (no crash, no silent fallback)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import estimators as OE, machines as OM, sampler as OS, sr as OSR
from oracle.models import lindblad_ising_1d, tfim_1d

pytestmark = pytest.mark.gpu


def test_empty_and_single_batches(nq, ctx):
    N = 6
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, np.complex128, OM.LOGCOSH)
    _, oH = tfim_1d(N)
    _, pH = H.p_tfim_1d(nq, N)
    op = pH.to_device(ctx)
    e0 = nq.local_scalar(pm, op, np.zeros((N, 0), order="F"))
    assert e0.shape == (0,)
    s1 = H.rand_states("spin", N, 1, 3)
    H.assert_close(nq.local_scalar(pm, op, s1), OE.local_scalar_ket(om, oH, s1), 1e-11, "E_loc, B = 1")
    om2, pm2, hilb2 = H.make_pair(nq, ctx, "ndm", "fock", 4, 2, np.float64, OM.SOFTPLUS)
    _, _, _, ol = lindblad_ising_1d(4)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, 4)
    opl = pl.to_device(ctx)
    l0, g0 = nq.local_grad(pm2, opl, (np.zeros((4, 0), order="F"), np.zeros((4, 0), order="F")))
    assert l0.shape == (0,) and g0.shape == (pm2.P, 0)
    r1, c1 = H.rand_states("fock", 4, 1, 5), H.rand_states("fock", 4, 1, 6)
    l1, g1 = nq.local_grad(pm2, opl, (r1, c1))
    rl, rg = OE.local_grad_super(om2, ol, r1, c1)
    H.assert_close(l1, rl, 1e-11, "L_loc, B = 1")
    H.assert_close(g1, rg, 1e-11, "grad L_loc, B = 1")
    # the zero operator (KLocalZero): E_loc = 0 for every configuration
    zero = nq.LocalOperator(hilb).to_device(ctx)
    assert np.all(nq.local_scalar(pm, zero, H.rand_states("spin", N, 5, 1)) == 0)


def test_single_sample_and_padded_leading_dimension(nq, ctx):
    L = nq._lib
    rng = np.random.default_rng(3)
    P, Ns, ld = 37, 50, 64
    O = (rng.standard_normal((P, Ns)) + 1j * rng.standard_normal((P, Ns))) + 0.3
    pad = np.full((ld, Ns), 7.0 + 7.0j, dtype=np.complex128, order="F")          # rows P..ld-1 must never be read or written
    pad[:P] = O
    d = torch.from_numpy(np.ascontiguousarray(pad.T)).cuda()
    avg = np.zeros(P, np.complex128)
    L.check(L.lib.nq_center(ctx.h, d.data_ptr(), ld, P, Ns, L.NQ_C128, L.ptr(avg)), ctx.h)
    H.assert_close(avg, O.mean(axis=1), 1e-13, "<O> with ld > P")
    back = d.cpu().numpy().T
    assert np.all(back[P:] == 7.0 + 7.0j)
    S = np.zeros((P, P), np.complex128, order="F")
    F = np.zeros(P, np.complex128)
    g = np.ones(P, np.complex128)
    L.check(L.lib.nq_sr_setup(ctx.h, d.data_ptr(), ld, P, Ns, Ns, L.NQ_C128, L.ptr(g), 0, L.ptr(S), L.ptr(F)), ctx.h)
    rS, _ = OSR.sr_setup(O - O.mean(axis=1)[:, None], g, False)
    H.assert_close(S, rS, 1e-11, "S with ld > P")
    # one sample: the centred matrix is zero and so is S
    one = torch.from_numpy(np.ascontiguousarray(O[:, :1].T)).cuda()
    L.check(L.lib.nq_center(ctx.h, one.data_ptr(), P, P, 1, L.NQ_C128, L.ptr(avg)), ctx.h)
    assert np.allclose(avg, O[:, 0]) and np.all(one.cpu().numpy() == 0)


def test_widest_configuration_and_limits(nq, ctx):
    """N = 256 sites = four 64-bit words (the widest the sampler supports); N = 257 is refused with a status."""
    N = 256
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 1, np.float64, OM.LOGCOSH, std=0.05)
    smp = nq.MetropolisSampler(nq.LocalRule(), 4, 3)
    cache = nq.MetropolisSamplerCache(smp, pm, 6)
    st = H.rand_states("spin", N, 6, 1)
    cache.set_state(st)
    rng = np.random.default_rng(8)
    sites = rng.integers(1, N + 1, size=(3, 6))
    sites[0, :3] = [1, 64, 65]
    sites[1, :3] = [128, 129, 256]                       # word boundaries and the last site
    u = rng.random((3, 6))
    new, acc_ref, margin = OS.samplenext_replay(om, H.ohilb("spin", N), st, sites, u)
    acc = cache.replay(sites, u)
    assert np.abs(margin).min() > 1e-9 and np.array_equal(acc, acc_ref) and np.array_equal(cache.get_state(), new)
    om2, pm2, hilb2 = H.make_pair(nq, ctx, "rbm", "spin", 257, 1, np.float64, OM.LOGCOSH, std=0.05)
    c2 = nq.MetropolisSamplerCache(nq.MetropolisSampler(nq.LocalRule(), 4, 3), pm2, 2)
    with pytest.raises(nq.NQError) as e:
        c2.sample(burn=1, L_store=1)
    assert e.value.status == nq._lib.NQ_ERR_UNSUPPORTED
    # the machine kernels themselves have no such limit
    s = H.rand_states("spin", 257, 3, 2)
    H.assert_close(pm2.logpsi(s), om2.logpsi(s), 1e-11, "log psi, N = 257")


def test_burn_only_run_and_counters(nq, ctx):
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", 5, 2, np.float64, OM.SOFTPLUS)
    smp = nq.MetropolisSampler(nq.LocalRule(), 4, 5, burn=0, seed=3)
    cache = nq.MetropolisSamplerCache(smp, pm, 7)
    cache.randomize()
    before = cache.get_state()
    L = nq._lib
    L.check(L.lib.nq_sampler_sample(cache.h, 6, 0, None, None, None, None, L.NQ_F64), ctx.h)      # burn 6, store nothing
    done, acc = cache.counters()
    assert done == 6 * 5 * 7 and 0 < acc <= done
    after = cache.get_state()
    assert any(not np.array_equal(a, b) for a, b in zip(before, after))
    L.check(L.lib.nq_sampler_sample(cache.h, 0, 0, None, None, None, None, L.NQ_F64), ctx.h)      # nothing to do
    assert cache.counters() == (done, acc)


def test_argument_errors_are_statuses(nq, ctx):
    L = nq._lib
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", 6, 2, np.complex128, OM.LOGCOSH)
    with pytest.raises(nq.NQError) as e:                     # wrong parameter count
        pm.set_params(np.zeros(pm.P + 1, np.complex128))
    assert e.value.status == L.NQ_ERR_SHAPE
    d = torch.zeros((10, 8), dtype=torch.complex128, device="cuda")
    avg = np.zeros(8, np.complex128)
    assert L.lib.nq_center(ctx.h, d.data_ptr(), 4, 8, 10, L.NQ_C128, L.ptr(avg)) == L.NQ_ERR_ARG          # ld < P
    assert L.lib.nq_center(ctx.h, d.data_ptr(), 8, 8, 0, L.NQ_C128, L.ptr(avg)) == L.NQ_ERR_ARG           # no samples
    with pytest.raises(nq.NQError):                          # a ket machine with a doubled configuration
        L.check(L.lib.nq_logpsi(pm.h, L.ptr(np.zeros((6, 2), order="F")), L.ptr(np.zeros((6, 2), order="F")), L.NQ_F64, 2,
                                L.ptr(np.zeros(2, np.complex128))), ctx.h)
    S = torch.eye(4, dtype=torch.float64, device="cuda") * -1.0
    F = torch.ones(4, dtype=torch.float64, device="cuda")
    dw = torch.zeros(4, dtype=torch.float64, device="cuda")
    its = L.C.c_int64()
    st = L.lib.nq_sr_solve(ctx.h, S.data_ptr(), F.data_ptr(), 4, L.NQ_F64, 0.0, L.NQ_SOLVE_CHOLESKY, 0.0, 0, dw.data_ptr(), L.C.byref(its))
    assert st == L.NQ_ERR_NOT_POSDEF
    with pytest.raises(nq.PosDefException):
        L.check(st, ctx.h)


@pytest.mark.parametrize("dtype,P,Ns,ld,real_params", [(np.complex128, 37, 4200, 64, False), (np.complex64, 37, 4200, 64, True),
                                                       (np.float64, 130, 150000, 130, True)])
def test_tensor_core_s_assembly_shapes(nq, ctx, dtype, P, Ns, ld, real_params):
    """The tcgen05 S-assembly paths (FP64: Ozaki digits on kind::i8; FP32 mode: TMA-fed 3xTF32) on awkward shapes: a padded
    leading dimension (rows P..ld-1 hold garbage that must not enter S), sizes that are not multiples of the 128-row tiles or
    the 64-sample boxes, and a K so long that one CTA drains its int32 accumulators several times (150 000 samples in one
    split: three accumulation windows)."""
    L = nq._lib
    rng = np.random.default_rng(23)
    dtype = np.dtype(dtype)
    cplx = dtype.kind == "c"
    O = rng.standard_normal((P, Ns)) + (1j * rng.standard_normal((P, Ns)) if cplx else 0.0)
    O = (O - O.mean(axis=1, keepdims=True)).astype(dtype)
    pad = np.full((ld, Ns), 1e6, dtype=dtype, order="F")
    pad[:P] = O
    d = torch.from_numpy(np.ascontiguousarray(pad.T)).cuda()
    single = dtype in (np.dtype(np.float32), np.dtype(np.complex64))
    sdt = np.dtype(dtype if (cplx and not real_params) else (np.float32 if single else np.float64))
    S = np.zeros((P, P), sdt, order="F")
    F = np.zeros(P, sdt)
    g = np.ones(P, np.complex64 if single else np.complex128)
    L.check(L.lib.nq_sr_setup(ctx.h, d.data_ptr(), ld, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(g), int(real_params), L.ptr(S), L.ptr(F)), ctx.h)
    O64 = O.astype(np.complex128 if cplx else np.float64)
    rS, _ = OSR.sr_setup(O64, g.astype(np.complex128), real_params)
    H.assert_close(S, rS, H.TOL[dtype], "S")
