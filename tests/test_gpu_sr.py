"""GPU parity: centring, force vectors, S assembly (DMMA SYRK/HERK), Cholesky / CG solves, chain
statistics (K6-K9) vs the oracle."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import sr as OSR
from oracle import stats as OST

pytestmark = pytest.mark.gpu


def _dev(a):
    """numpy [P, Ns] (Fortran) -> torch device tensor [Ns, P] sharing the column-major layout."""
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).cuda()


def _rand(rng, shape, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


@pytest.mark.parametrize("dtype,P,Ns,real_params", [
    (np.complex128, 230, 1000, False), (np.complex128, 576, 2000, True), (np.float64, 130, 515, True),
    (np.complex64, 200, 777, False), (np.float32, 129, 600, True), (np.complex128, 300, 70, False),
    (np.complex128, 1, 9, False),
    # >= 4096 samples: the FP64 S assembly runs on the integer tensor cores (Ozaki scheme, nq_syrk_ozaki.cu): complex S (two
    # launches), real S of complex rows, real rows; odd sizes exercise the zero padding of rows and samples
    (np.complex128, 230, 4200, False), (np.complex128, 576, 4500, True), (np.float64, 130, 4100, True)])
def test_center_force_setup(nq, ctx, dtype, P, Ns, real_params):
    L = nq._lib
    rng = np.random.default_rng(5)
    dtype = np.dtype(dtype)
    cdt = np.dtype(np.complex64 if dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else np.complex128)
    tol = H.TOL[dtype]
    O = _rand(rng, (P, Ns), dtype) + dtype.type(0.5)
    O64 = O.astype(np.complex128 if dtype.kind == "c" else np.float64)
    E = _rand(rng, Ns, cdt)
    dO = _dev(O)
    avg = np.zeros(P, dtype)
    L.check(L.lib.nq_center(ctx.h, dO.data_ptr(), P, P, Ns, L.nq_dtype(dtype), L.ptr(avg)), ctx.h)
    ravg, rOc = OSR.center(O64)
    H.assert_close(avg, ravg, tol, "<O>")
    # centred values are differences of O(1) numbers: tolerance relative to |O|
    Oc = dO.cpu().numpy().T
    assert np.max(np.abs(Oc - rOc)) <= 4 * tol * np.abs(O64).max()
    Oc64 = Oc.astype(O64.dtype)
    # ket force
    g = np.zeros(P, cdt)
    L.check(L.lib.nq_force_ket(ctx.h, dO.data_ptr(), P, P, Ns, L.nq_dtype(dtype), L.ptr(E), L.ptr(g)), ctx.h)
    # the force cancels to a fraction of its terms: reference accumulated in long double (oracle/sr.py), device in
    # double-double (nq_sr.cu colsum) -> the stated tolerance holds without a multiplier
    rg = OSR.force_ket_ld(E.astype(np.complex128), O64)
    H.assert_close(g, rg, tol, "grad C (ket)")
    # S, F
    sdt = np.dtype(dtype if (dtype.kind == "c" and not real_params) else (np.float32 if cdt == np.complex64 else np.float64))
    S = np.zeros((P, P), sdt, order="F")
    F = np.zeros(P, sdt)
    L.check(L.lib.nq_sr_setup(ctx.h, dO.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(g), int(real_params),
                              L.ptr(S), L.ptr(F)), ctx.h)
    rS, rF = OSR.sr_setup(Oc64, g.astype(np.complex128), real_params)
    H.assert_close(S, rS, tol, "S")
    H.assert_close(F, rF, tol, "F")
    if sdt.kind == "c":
        assert np.allclose(S, S.conj().T, atol=0)          # Hermitian by construction (mirrored tiles)


@pytest.mark.parametrize("dtype,P,Ns,real_params,structured", [
    (np.complex64, 256, 5000, False, False), (np.complex64, 300, 4100, True, False),
    (np.float32, 260, 3000, True, False), (np.complex64, 516, 3001, True, True),
    (np.float32, 132, 70000, True, False), (np.complex64, 64, 40, False, False)])
def test_sr_setup_fp32_tensor_path(nq, ctx, dtype, P, Ns, real_params, structured):
    """FP32-mode S assembly runs on tcgen05 (3xTF32, nq_syrk_tf32.cu); tolerance 1e-5 vs the FP64 oracle."""
    L = nq._lib
    rng = np.random.default_rng(11)
    dtype = np.dtype(dtype)
    O = _rand(rng, (P, Ns), dtype)
    if structured:                       # real-parameter NDM: lambda rows real, mu rows imaginary
        O[: P // 3] = O[: P // 3].real
        O[P // 3:] = 1j * O[P // 3:].imag
    O -= O.mean(axis=1, keepdims=True)
    O64 = O.astype(np.complex128 if dtype.kind == "c" else np.float64)
    dO = _dev(O)
    g = _rand(rng, P, np.complex64)
    sdt = np.dtype(dtype if (dtype.kind == "c" and not real_params) else np.float32)
    S = np.zeros((P, P), sdt, order="F")
    F = np.zeros(P, sdt)
    L.check(L.lib.nq_sr_setup(ctx.h, dO.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(g), int(real_params),
                              L.ptr(S), L.ptr(F)), ctx.h)
    rS, rF = OSR.sr_setup(O64, g.astype(np.complex128), real_params)
    # entries of S are sums of Ns products of O(1) numbers / Ns: tolerance relative to the diagonal scale
    scale = float(np.abs(np.diag(rS)).max())
    assert np.max(np.abs(S - rS)) <= 1e-5 * scale, np.max(np.abs(S - rS)) / scale
    H.assert_close(F, rF, 1e-5, "F")
    if sdt.kind == "c":
        assert np.allclose(S, S.conj().T, atol=0)


def test_force_liouvillian(nq, ctx):
    L = nq._lib
    rng = np.random.default_rng(6)
    P, Ns = 200, 640
    Lloc = _rand(rng, Ns, np.complex128)
    gL = _rand(rng, (P, Ns), np.complex128)
    avg = _rand(rng, P, np.complex128)
    g = np.zeros(P, np.complex128)
    import ctypes as C
    cost = C.c_double()
    dg = _dev(gL)
    L.check(L.lib.nq_force_liouvillian(ctx.h, L.ptr(Lloc), dg.data_ptr(), P, P, Ns, L.NQ_C128, L.ptr(avg), L.ptr(g),
                                       C.byref(cost)), ctx.h)
    H.assert_close(g, OSR.force_liouvillian(Lloc, gL, avg), 1e-11, "grad C (Liouvillian)")
    assert abs(cost.value - np.mean(np.abs(Lloc) ** 2)) < 1e-12 * cost.value


def _spd(rng, P, cplx):
    Ns = 3 * P
    O = rng.standard_normal((P, Ns)) + (1j * rng.standard_normal((P, Ns)) if cplx else 0)
    S = (O.conj() @ O.T) / Ns
    return np.asfortranarray(S)


@pytest.mark.parametrize("P,cplx", [(230, True), (576, False), (33, False), (1, True), (100, True)])
def test_cholesky_and_cg_solve(nq, ctx, P, cplx):
    import ctypes as C
    L = nq._lib
    rng = np.random.default_rng(7)
    S = _spd(rng, P, cplx)
    F = rng.standard_normal(P) + (1j * rng.standard_normal(P) if cplx else 0)
    eps = OSR.eps_f32(0.001)
    ref = OSR.solve_cholesky(S, F, eps)
    sd = L.NQ_C128 if cplx else L.NQ_F64
    for algo, tol in ((L.NQ_SOLVE_CHOLESKY, 1e-10), (L.NQ_SOLVE_CG, 1e-8)):
        Sw = S.copy(order="F")
        dw = np.zeros_like(F)
        its = C.c_int64()
        L.check(L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), P, sd, eps, algo, 1e-12, 0, L.ptr(dw), C.byref(its)), ctx.h)
        cond = np.linalg.cond(S + eps * np.eye(P))
        assert np.linalg.norm(dw - ref) <= max(tol, 1e-15 * cond) * np.linalg.norm(ref), (algo, its.value)
        if algo == L.NQ_SOLVE_CG:
            assert 0 < its.value <= 10 * P
            x, it_ref, ok = OSR.solve_cg_explicit(S, F, eps, 1e-12)
            assert abs(its.value - it_ref) <= 2          # same stopping rule as the oracle (unpinned in the reference)


@pytest.mark.parametrize("P,cplx", [(1700, True), (3200, False)])
def test_cholesky_large_systems(nq, ctx, P, cplx):
    """Large systems take the DMMA trailing update and the multi-CTA per-block triangular solves."""
    import ctypes as C
    L = nq._lib
    rng = np.random.default_rng(17)
    X = rng.standard_normal((P, 2 * P)) + (1j * rng.standard_normal((P, 2 * P)) if cplx else 0)
    S = np.asfortranarray((X @ X.conj().T) / (2 * P))
    F = rng.standard_normal(P) + (1j * rng.standard_normal(P) if cplx else 0)
    eps = OSR.eps_f32(0.001)
    ref = np.linalg.solve(S + eps * np.eye(P), F)
    dw = np.zeros_like(F)
    its = C.c_int64()
    Sw = S.copy(order="F")               # keep the buffer alive across the call (L.ptr does not hold a reference)
    L.check(L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), P, L.NQ_C128 if cplx else L.NQ_F64, eps,
                              L.NQ_SOLVE_CHOLESKY, 1e-12, 0, L.ptr(dw), C.byref(its)), ctx.h)
    cond = np.linalg.cond(S + eps * np.eye(P))
    assert np.linalg.norm(dw - ref) <= max(1e-10, 1e-14 * cond) * np.linalg.norm(ref)


def test_cholesky_not_posdef_and_cg_not_converged(nq, ctx):
    import ctypes as C
    L = nq._lib
    S = -np.eye(40, order="F")
    S[0, 0] = 1.0
    F = np.ones(40)
    dw = np.zeros(40)
    its = C.c_int64()
    Sw = S.copy(order="F")
    st = L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), 40, L.NQ_F64, 0.0, L.NQ_SOLVE_CHOLESKY, 0.0, 0,
                           L.ptr(dw), C.byref(its))
    assert st == L.NQ_ERR_NOT_POSDEF
    with pytest.raises(nq.PosDefException):
        L.check(st, ctx.h)
    info = C.c_int64()
    L.lib.nq_ctx_last_info(ctx.h, C.byref(info))
    assert info.value == 1
    rng = np.random.default_rng(1)
    S = _spd(rng, 50, False)
    F50, dw50 = np.ones(50), np.zeros(50)
    st = L.lib.nq_sr_solve(ctx.h, L.ptr(S), L.ptr(F50), 50, L.NQ_F64, 0.0, L.NQ_SOLVE_CG, 1e-30, 3,
                           L.ptr(dw50), C.byref(its))
    assert st == L.NQ_ERR_NOT_CONVERGED and its.value == 3


@pytest.mark.parametrize("dtype,real_params", [(np.complex128, False), (np.complex128, True), (np.float64, True)])
def test_matrix_free_cg(nq, ctx, dtype, real_params):
    import ctypes as C
    L = nq._lib
    rng = np.random.default_rng(8)
    P, Ns = 150, 900
    O = _rand(rng, (P, Ns), dtype)
    O64 = O.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    _, Oc = OSR.center(O64)
    cplx_out = np.dtype(dtype).kind == "c" and not real_params
    F = rng.standard_normal(P) + (1j * rng.standard_normal(P) if cplx_out else 0)
    eps = 0.01
    ref, it_ref, ok = OSR.solve_cg(Oc, F, eps, 1e-10, real_params)
    assert ok
    dOc = _dev(Oc.astype(dtype))
    dw = np.zeros_like(F)
    its = C.c_int64()
    L.check(L.lib.nq_sr_solve_matfree(ctx.h, dOc.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(F), int(real_params),
                                      eps, 1e-10, 0, L.ptr(dw), C.byref(its)), ctx.h)
    assert np.linalg.norm(dw - ref) <= 1e-8 * np.linalg.norm(ref)
    assert abs(its.value - it_ref) <= 2
    S, _ = OSR.sr_setup(Oc, F, real_params)
    assert np.linalg.norm(dw - np.linalg.solve(S + eps * np.eye(P), F)) <= 1e-7 * np.linalg.norm(ref)


@pytest.mark.parametrize("dtype,real_params", [(np.complex128, False), (np.complex128, True), (np.float64, True)])
def test_minres_explicit_and_matrix_free(nq, ctx, dtype, real_params):
    """sr_minres (SRIterative.jl:101-125): MINRES on the explicit S and matrix-free, against the oracle's MINRES and a
    direct solve (the reference pins neither iterates nor counts: converged solutions are compared)."""
    import ctypes as C
    L = nq._lib
    rng = np.random.default_rng(18)
    P, Ns = 150, 900
    O = _rand(rng, (P, Ns), dtype)
    O64 = O.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    _, Oc = OSR.center(O64)
    cplx_out = np.dtype(dtype).kind == "c" and not real_params
    F = rng.standard_normal(P) + (1j * rng.standard_normal(P) if cplx_out else 0)
    eps = 0.01
    S, _ = OSR.sr_setup(Oc, F, real_params)
    direct = np.linalg.solve(S + eps * np.eye(P), F)
    ref, it_ref, ok = OSR.solve_minres_explicit(S, F, eps, 1e-10)
    assert ok and np.linalg.norm(ref - direct) <= 1e-8 * np.linalg.norm(direct)
    # explicit S
    Sw = np.asfortranarray(S.astype(np.complex128 if cplx_out else np.float64))
    dw = np.zeros_like(F)
    its = C.c_int64()
    L.check(L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), P, L.NQ_C128 if cplx_out else L.NQ_F64, eps, L.NQ_SOLVE_MINRES,
                              1e-10, 0, L.ptr(dw), C.byref(its)), ctx.h)
    assert np.linalg.norm(dw - direct) <= 1e-8 * np.linalg.norm(direct)
    assert abs(its.value - it_ref) <= 2
    # matrix-free
    dOc = _dev(Oc.astype(dtype))
    dw2 = np.zeros_like(F)
    L.check(L.lib.nq_sr_solve_matfree_algo(ctx.h, dOc.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(F), int(real_params),
                                           eps, L.NQ_SOLVE_MINRES, 1e-10, 0, L.ptr(dw2), C.byref(its)), ctx.h)
    assert np.linalg.norm(dw2 - direct) <= 1e-7 * np.linalg.norm(direct)
    assert abs(its.value - it_ref) <= 2
    # exhausted iterations are reported like CG's
    st = L.lib.nq_sr_solve_matfree_algo(ctx.h, dOc.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(F), int(real_params),
                                        eps, L.NQ_SOLVE_MINRES, 1e-30, 3, L.ptr(dw2), C.byref(its))
    assert st == L.NQ_ERR_NOT_CONVERGED and its.value == 3


@pytest.mark.parametrize("dtype,real_params", [(np.complex128, False), (np.complex128, True), (np.float64, True)])
def test_minresqlp_explicit_and_matrix_free(nq, ctx, dtype, real_params):
    """sr_qlp (SRIterative.jl:92-100 -> External/IterativeSolvers/minresqlp.jl): the device solver against the oracle's
    restatement (same exit flag, iteration count within 2, same solution) and a direct solve."""
    import ctypes as C
    from oracle import minresqlp as OQ
    L = nq._lib
    rng = np.random.default_rng(28)
    P, Ns = 150, 900
    O = _rand(rng, (P, Ns), dtype)
    O64 = O.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    _, Oc = OSR.center(O64)
    cplx_out = np.dtype(dtype).kind == "c" and not real_params
    F = rng.standard_normal(P) + (1j * rng.standard_normal(P) if cplx_out else 0)
    eps = 0.01
    S, _ = OSR.sr_setup(Oc, F, real_params)
    direct = np.linalg.solve(S + eps * np.eye(P), F)
    ref, info = OQ.solve_qlp_explicit(S, F, eps, 1e-10)
    assert info["flag"] == 1 and np.linalg.norm(ref - direct) <= 1e-7 * np.linalg.norm(direct)
    Sw = np.asfortranarray(S.astype(np.complex128 if cplx_out else np.float64))
    dw = np.zeros_like(F)
    its, flag = C.c_int64(), C.c_int64()
    sd = L.NQ_C128 if cplx_out else L.NQ_F64
    L.check(L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), P, sd, eps, L.NQ_SOLVE_QLP, 1e-10, 0, L.ptr(dw), C.byref(its)), ctx.h)
    L.check(L.lib.nq_ctx_last_info(ctx.h, C.byref(flag)), ctx.h)
    assert flag.value == info["flag"] and abs(its.value - info["iters"]) <= 2
    assert np.linalg.norm(dw - ref) <= 1e-8 * np.linalg.norm(ref)
    # matrix-free
    dOc = _dev(Oc.astype(dtype))
    dw2 = np.zeros_like(F)
    L.check(L.lib.nq_sr_solve_matfree_algo(ctx.h, dOc.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(F), int(real_params),
                                           eps, L.NQ_SOLVE_QLP, 1e-10, 0, L.ptr(dw2), C.byref(its)), ctx.h)
    assert np.linalg.norm(dw2 - direct) <= 1e-7 * np.linalg.norm(direct) and abs(its.value - info["iters"]) <= 2
    # warm start from a perturbed solution: same answer, far fewer iterations
    dw3 = (ref * (1 + 1e-6)).astype(F.dtype)
    L.check(L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), P, sd, eps, L.NQ_SOLVE_QLP_WARM, 1e-6, 0, L.ptr(dw3), C.byref(its)), ctx.h)
    assert np.linalg.norm(dw3 - direct) <= 1e-7 * np.linalg.norm(direct)
    # the iteration limit is the one exit reported as not converged (flag 8)
    st = L.lib.nq_sr_solve(ctx.h, L.ptr(Sw), L.ptr(F), P, sd, eps, L.NQ_SOLVE_QLP, 1e-30, 3, L.ptr(dw2), C.byref(its))
    L.check(L.lib.nq_ctx_last_info(ctx.h, C.byref(flag)), ctx.h)
    assert st == L.NQ_ERR_NOT_CONVERGED and its.value == 3 and flag.value == 8


def test_minresqlp_singular_system_where_minres_fails(nq, ctx):
    """Ns < P and no shift: S is singular and F has a component outside its range.  MINRES-QLP returns the
    minimum-length least-squares solution pinv(S) F (exit flag of the oracle), plain MINRES does not converge."""
    import ctypes as C
    from oracle import minresqlp as OQ
    L = nq._lib
    rng = np.random.default_rng(31)
    P, Ns = 60, 20
    X = rng.standard_normal((P, Ns))
    S = np.asfortranarray(X @ X.T / Ns)
    F = rng.standard_normal(P)
    want = np.linalg.pinv(S) @ F
    # the Krylov space is exhausted after rank + 1 = 21 steps: the last rotation is singular, the solver drops that
    # component and stops on the solution-norm guard (flag 6) with the minimum-length least-squares solution
    ref, info = OQ.minresqlp(lambda v: S @ v, F, tol=1e-12, maxiter=10 * P)
    assert info["flag"] == 6 and np.linalg.norm(ref - want) <= 1e-6 * np.linalg.norm(want)
    dw = np.zeros(P)
    its, flag = C.c_int64(), C.c_int64()
    Sa = S.copy(order="F")              # keep the arrays alive across the call (S is overwritten by some solvers)
    st = L.lib.nq_sr_solve(ctx.h, L.ptr(Sa), L.ptr(F), P, L.NQ_F64, 0.0, L.NQ_SOLVE_QLP, 1e-12, 0, L.ptr(dw), C.byref(its))
    L.check(st, ctx.h)
    L.check(L.lib.nq_ctx_last_info(ctx.h, C.byref(flag)), ctx.h)
    assert flag.value == info["flag"] and abs(its.value - info["iters"]) <= 1
    assert np.linalg.norm(dw - want) <= 1e-6 * np.linalg.norm(want)
    dwm = np.zeros(P)
    Sb = S.copy(order="F")
    stm = L.lib.nq_sr_solve(ctx.h, L.ptr(Sb), L.ptr(F), P, L.NQ_F64, 0.0, L.NQ_SOLVE_MINRES, 1e-12, 0, L.ptr(dwm), C.byref(its))
    assert stm == L.NQ_ERR_NOT_CONVERGED or np.linalg.norm(dwm) > 1e3 * np.linalg.norm(want)


def test_restart_ladder_and_multiplicative_regulariser(nq, ctx):
    """precondition!: an unconverged CG is followed by warm-started MINRES-QLP restarts (SRIterative.jl:133-150);
    sr_multiplicative solves (S + lambda Diagonal(diag S)) dw = F (SRDirect.jl:66-72, SRIterative.jl:84-90)."""
    import warnings
    from oracle.models import tfim_1d
    N, B, Lc = 6, 8, 40
    oh, oH = tfim_1d(N)
    ph, pH = H.p_tfim_1d(nq, N)
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, np.float64, 1)
    S_ = H.rand_states("spin", N, B * Lc, 4321)
    ref = OSR.iteration_ket(om, oH, S_, OSR.eps_f32(0.01))
    smp = nq.MetropolisSampler(nq.LocalRule(), Lc, N, burn=2, seed=1)
    # (a) CG cut short (maxiter is 10 P in the reference; the mirror lets tests lower it): the ladder's warm-started
    #     MINRES-QLP runs (tolerance sqrt(eps)) finish the job
    want = np.linalg.solve(ref["S"] + 0.01 * np.eye(pm.P), ref["F"])
    bs = nq.BatchedSampler(pm, smp, pH, nq.SR(np.float64, eps=0.01, algorithm=nq.sr_cg, precision=1e-12, maxiter=25), batch_sz=B)
    bs.set_samples(S_.reshape(N, B, Lc, order="F"))
    bs.sample_(sample=False)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        dw = bs.precondition_().cpu().numpy()
    assert 1 <= bs.restarts <= 5 and bs.converged and len(wlist) == bs.restarts and "not converged" in str(wlist[0].message)
    assert np.linalg.norm(dw - want) <= 1e-6 * np.linalg.norm(want)
    #     ... and with 2 iterations per run nothing converges: five restarts, then the zero update
    bs = nq.BatchedSampler(pm, smp, pH, nq.SR(np.float64, eps=0.01, algorithm=nq.sr_cg, precision=1e-12, maxiter=2), batch_sz=B)
    bs.set_samples(S_.reshape(N, B, Lc, order="F"))
    bs.sample_(sample=False)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        dw = bs.precondition_().cpu().numpy()
    assert bs.restarts == 5 and not bs.converged and len(wlist) == 5 and np.all(dw == 0)
    # (b) multiplicative regulariser, explicit S, iteration 7
    algo = nq.SR(np.float64, algorithm=nq.sr_cholesky, precondition_type=nq.sr_multiplicative, lambda0=100.0, b=0.95, lambda_min=1e-4)
    bs = nq.BatchedSampler(pm, smp, pH, algo, batch_sz=B)
    bs.set_samples(S_.reshape(N, B, Lc, order="F"))
    bs.sample_(sample=False)
    dw = bs.precondition_(7).cpu().numpy()
    lam = max(100.0 * 0.95 ** 7, 1e-4)
    want = np.linalg.solve(ref["S"] + lam * np.diag(np.diag(ref["S"])), ref["F"])
    assert np.linalg.norm(dw - want) <= 1e-9 * np.linalg.norm(want)
    # (c) sr_qlp through the mirror
    bs = nq.BatchedSampler(pm, smp, pH, nq.SR(np.float64, eps=0.01, algorithm=nq.sr_qlp, precision=1e-10), batch_sz=B)
    bs.set_samples(S_.reshape(N, B, Lc, order="F"))
    bs.sample_(sample=False)
    dw = bs.precondition_().cpu().numpy()
    want = np.linalg.solve(ref["S"] + 0.01 * np.eye(pm.P), ref["F"])
    assert bs.converged and np.linalg.norm(dw - want) <= 1e-7 * np.linalg.norm(want)


@pytest.mark.parametrize("dtype", [np.complex128, np.float64, np.complex64])
def test_stat_analysis(nq, ctx, dtype):
    rng = np.random.default_rng(9)
    v = _rand(rng, (16, 125), dtype)
    m = nq.stat_analysis(ctx, v)
    r = OST.stat_analysis(v.astype(np.complex128))
    tol = 1e-5 if np.dtype(dtype) == np.complex64 else 1e-12
    for a, b in ((m.mean, r["mean"]), (m.error, r["error"]), (m.variance, r["variance"]), (m.tau, r["tau"]), (m.R, r["R"])):
        assert abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize("dtype,P,Ns,real_params,chunks", [
    (np.complex128, 230, 3000, False, (1000, 1000, 1000)), (np.complex128, 576, 2100, True, (512, 1024, 564)),
    (np.float64, 130, 900, True, (900,)), (np.complex64, 256, 4096, False, (2048, 2048)), (np.complex128, 140, 5000, True, (4096, 904))])
def test_streaming_sr_assembly(nq, ctx, dtype, P, Ns, real_params, chunks):
    """nq_sr_accumulate / nq_sr_finish (O produced and consumed chunk by chunk, algebraic centring) reproduce the S of
    centre + setup on the whole batch: vs the oracle at the stated tolerance, with a mean of the order of the fluctuations
    (the cancellation case of the algebraic centring)."""
    L = nq._lib
    rng = np.random.default_rng(17)
    dtype = np.dtype(dtype)
    tol = H.TOL[dtype]
    O = _rand(rng, (P, Ns), dtype) + dtype.type(1.5)
    if dtype.kind == "c" and real_params:
        O[: P // 3] = O[: P // 3].real                       # purely real / purely imaginary row blocks (zero-plane skipping)
        O[P // 3: 2 * P // 3] = 1j * O[P // 3: 2 * P // 3].imag
    O64 = O.astype(np.complex128 if dtype.kind == "c" else np.float64)
    single = dtype in (np.dtype(np.float32), np.dtype(np.complex64))
    sdt = np.dtype(dtype if (dtype.kind == "c" and not real_params) else (np.float32 if single else np.float64))
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.complex64): torch.complex64,
           np.dtype(np.complex128): torch.complex128}
    S = torch.zeros((P, P), dtype=tdt[sdt], device="cuda")
    sumO = torch.zeros(2 * P, dtype=torch.complex128, device="cuda")
    assert sum(chunks) == Ns
    c0 = 0
    buf = torch.zeros((max(chunks), P), dtype=tdt[dtype], device="cuda")          # ONE reused chunk buffer
    for i, n in enumerate(chunks):
        buf[:n].copy_(_dev(O[:, c0:c0 + n]))
        L.check(L.lib.nq_sr_accumulate(ctx.h, buf.data_ptr(), P, P, n, Ns, L.nq_dtype(dtype), int(real_params), S.data_ptr(),
                                       sumO.data_ptr(), int(i == 0)), ctx.h)
        c0 += n
    L.check(L.lib.nq_sr_finish(ctx.h, S.data_ptr(), sumO.data_ptr(), P, Ns, L.nq_dtype(dtype), int(real_params)), ctx.h)
    _, rOc = OSR.center(O64)
    rS, _ = OSR.sr_setup(rOc, np.zeros(P, np.complex128), real_params)
    H.assert_close(S.cpu().numpy().T, rS, tol, "streamed S")
    H.assert_close(sumO.cpu().numpy()[:P], O64.mean(axis=1), 1e-13, "<O>")


@pytest.mark.parametrize("dtype,P,Ns,real_params", [(np.complex128, 230, 4200, False), (np.complex128, 300, 4500, True),
                                                    (np.float64, 130, 4100, True), (np.complex128, 90, 1000, False)])
def test_deferred_centring(nq, ctx, dtype, P, Ns, real_params):
    """nq_center_lazy returns <O> and may leave O uncentred; the S assembly then subtracts the means on the fly (same S as
    centre + setup), repeated setups keep working, and nq_center_finish produces the centred rows."""
    L = nq._lib
    rng = np.random.default_rng(31)
    dtype = np.dtype(dtype)
    O = (_rand(rng, (P, Ns), dtype) + dtype.type(0.7)).astype(dtype)
    O64 = O.astype(np.complex128 if dtype.kind == "c" else np.float64)
    d = _dev(O)
    avg = np.zeros(P, dtype)
    flag = L.C.c_int(-1)
    L.check(L.lib.nq_center_lazy(ctx.h, d.data_ptr(), P, P, Ns, L.nq_dtype(dtype), L.ptr(avg), L.C.byref(flag)), ctx.h)
    ravg, rOc = OSR.center(O64)
    H.assert_close(avg, ravg, 1e-13, "<O>")
    assert flag.value == (1 if Ns >= 4096 else 0)
    if flag.value:
        assert np.array_equal(d.cpu().numpy().T, O)              # untouched
    sdt = np.dtype(dtype if (dtype.kind == "c" and not real_params) else np.float64)
    g = np.ones(P, np.complex128)
    rS, _ = OSR.sr_setup(rOc, g, real_params)
    for rep in range(2):                                          # the second call finds no row maxima left behind
        S = np.zeros((P, P), sdt, order="F")
        F = np.zeros(P, sdt)
        L.check(L.lib.nq_sr_setup(ctx.h, d.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(g), int(real_params), L.ptr(S), L.ptr(F)), ctx.h)
        H.assert_close(S, rS, 1e-11, "S, deferred centring, call %d" % rep)
    L.check(L.lib.nq_center_finish(ctx.h, d.data_ptr(), P, P, Ns, L.nq_dtype(dtype)), ctx.h)
    assert np.max(np.abs(d.cpu().numpy().T - rOc)) <= 4e-11 * np.abs(O64).max()
    L.check(L.lib.nq_center_finish(ctx.h, d.data_ptr(), P, P, Ns, L.nq_dtype(dtype)), ctx.h)      # nothing pending: no-op
    assert np.max(np.abs(d.cpu().numpy().T - rOc)) <= 4e-11 * np.abs(O64).max()
    S = np.zeros((P, P), sdt, order="F")
    F = np.zeros(P, sdt)
    L.check(L.lib.nq_sr_setup(ctx.h, d.data_ptr(), P, P, Ns, Ns, L.nq_dtype(dtype), L.ptr(g), int(real_params), L.ptr(S), L.ptr(F)), ctx.h)
    H.assert_close(S, rS, 1e-11, "S after the explicit centring")
