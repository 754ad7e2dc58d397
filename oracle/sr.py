"""Oracle: gradient centring, force vector, SR S-matrix, solve, update.  Test infra only.

Follows
  src/IterativeInterface/Samplers/BaseIterativeSampler.jl:19-26      (centre O)
  src/IterativeInterface/Samplers/CostFun/BatchedValSampler.jl:97-115 (ket force)
  src/IterativeInterface/Samplers/CostFun/BatchedGradSampler.jl:99-118 (Liouvillian force)
  src/Algorithms/SR/SRDirect.jl:26-90   (explicit S, shift, Cholesky)
  src/Algorithms/SR/SRIterative.jl:71-153, SR_notfull.jl:47-148 (matrix-free S, CG)
  src/Optimisers/rules.jl:15-17, apply.jl:25-32 (Descent)
CG restates IterativeSolvers.jl v0.8.1 `cg` (Manifest.toml pin; src/cg.jl of that package --
not vendored in the reference): x0 = 0, stop when ||r|| <= tol*||b||, at most maxiter
iterations.  No reference test pins its iterates ("parity unpinned"): tests compare the
converged solution with numpy.linalg.solve.
"""
import numpy as np


def center(O):
    """<O> over samples and O - <O>.  O [P, Ns]."""
    avg = O.mean(axis=1)
    return avg, O - avg[:, None]


def force_ket(Eloc, Oc):
    """grad C = E_loc[1,Ns] * Oc' / Ns  ->  F_k = <E_loc conj(Oc_k)>."""
    Ns = Oc.shape[1]
    return (np.asarray(Eloc)[None, :] @ Oc.conj().T).reshape(-1) / Ns


def force_liouvillian(Lloc, gLloc, O_avg):
    """F = conj( L_loc * gradL_loc' / Ns - <|L_loc|^2> <O>' )."""
    Ns = gLloc.shape[1]
    C = np.mean(np.abs(Lloc) ** 2)
    g = (np.asarray(Lloc)[None, :] @ gLloc.conj().T).reshape(-1) / Ns
    g = g - C * O_avg.conj()
    return g.conj()


# ---- the same three formulas accumulated in extended precision (x87 long double) --------------------------------
# The force is an average of O(1) terms that cancels to a small fraction of their size; numpy's float64 reductions
# carry an error ~1e-16 * log2(Ns) * sum|terms| -- the same order as the 1e-11 element-wise parity bound once the
# cancellation is counted.  The parity tests therefore compare the device (double-double accumulation) against these.
_LD = np.longdouble


def _ld_wsum(w, X):
    """sum_s w_s X[k, s] for complex w [Ns], X [P, Ns] in long double; returns (re, im) long double arrays."""
    wr, wi = np.real(w).astype(_LD), np.imag(w).astype(_LD)
    Xr, Xi = np.real(X).astype(_LD), np.imag(X).astype(_LD)
    return Xr @ wr - Xi @ wi, Xr @ wi + Xi @ wr


def center_ld(O):
    Or, Oi = np.real(O).astype(_LD), np.imag(O).astype(_LD)
    ar, ai = Or.mean(axis=1), Oi.mean(axis=1)
    return (ar, ai), (Or - ar[:, None], Oi - ai[:, None])


def force_ket_ld(Eloc, O):
    """force_ket(Eloc, O - <O>) from the UNcentred O, accumulated in long double; complex128 result.
    F_k = (1/Ns) sum_s E_s conj(Oc_ks)."""
    Ns = O.shape[1]
    _, (cr, ci) = center_ld(O)
    er, ei = np.real(Eloc).astype(_LD), np.imag(Eloc).astype(_LD)
    re = (cr @ er + ci @ ei) / Ns                  # Re[E conj(Oc)] = Er Or + Ei Oi
    im = (cr @ ei - ci @ er) / Ns                  # Im[E conj(Oc)] = Ei Or - Er Oi
    return np.asarray(re, dtype=np.float64) + 1j * np.asarray(im, dtype=np.float64)


def force_liouvillian_ld(Lloc, gLloc, O):
    """force_liouvillian(Lloc, gLloc, <O>) with <O> from the uncentred O, accumulated in long double.
    F_k = conj( (1/Ns) sum_s L_s conj(gL_ks) - C conj(<O_k>) ),  C = <|L|^2>."""
    Ns = gLloc.shape[1]
    (ar, ai), _ = center_ld(O)
    lr, li = np.real(Lloc).astype(_LD), np.imag(Lloc).astype(_LD)
    gr, gi = np.real(gLloc).astype(_LD), np.imag(gLloc).astype(_LD)
    C = np.mean(lr * lr + li * li)
    re = (gr @ lr + gi @ li) / Ns - C * ar         # Re[L conj(gL)] - C Re[conj(avg)]
    im = (gr @ li - gi @ lr) / Ns + C * ai         # Im[L conj(gL)] - C Im[conj(avg)]
    return np.asarray(re, dtype=np.float64) - 1j * np.asarray(im, dtype=np.float64)


def sr_setup(Oc, gradC, real_params):
    """SRDirect.jl:26-49.  real_params: S = Re(Oc Oc^H)/Ns, F = Re(gradC);
    complex: S = conj(Oc Oc^H)/Ns, F = gradC."""
    Ns = Oc.shape[1]
    Sc = Oc @ Oc.conj().T
    if real_params:
        return np.real(Sc) / Ns, np.real(gradC)
    return Sc.conj() / Ns, np.asarray(gradC)


def eps_f32(eps):
    """Quirk Q16: SR() stores eps as Float32 (SR/SR.jl:50-57)."""
    return float(np.float32(eps))


def solve_cholesky(S, F, eps):
    """SRDirect.jl:62-64 (S_ii += eps), :78-81 cholesky!(Hermitian(S)) + ldiv!.
    Hermitian(S) reads the upper triangle."""
    S = np.array(S, copy=True)
    P = S.shape[0]
    S[np.arange(P), np.arange(P)] += eps
    U = np.triu(S)
    H = U + np.triu(S, 1).conj().T
    L = np.linalg.cholesky(H)          # raises LinAlgError if not PD (check=true)
    y = np.linalg.solve(L, F)
    return np.linalg.solve(L.conj().T, y)


def cg(matvec, b, tol, maxiter):
    """IterativeSolvers 0.8.1 cg (see module docstring).  Returns (x, iters, converged)."""
    b = np.asarray(b)
    x = np.zeros_like(b)
    u = np.zeros_like(b)
    r = b.copy()
    residual = np.linalg.norm(b)
    prev_residual = 1.0
    reltol = residual * tol
    it = 0
    while it < maxiter and not residual <= reltol:
        beta = residual ** 2 / prev_residual ** 2
        u = r + beta * u
        c = matvec(u)
        alpha = residual ** 2 / np.vdot(u, c)
        x = x + alpha * u
        r = r - alpha * c
        prev_residual = residual
        residual = np.linalg.norm(r)
        it += 1
    return x, it, bool(residual <= reltol)


def sr_matvec_factory(Oc, eps, real_params):
    """SR_notfull.jl: complex nets v -> eps v + conj(Oc)(conj(Oc)^H v)/Ns  (:47-80);
    real nets (Or Or^T + Oi Oi^T) v / Ns + eps v  (:123-148)."""
    Ns = Oc.shape[1]
    if real_params:
        Or, Oi = Oc.real, Oc.imag
        return lambda v: eps * v + (Or @ (Or.T @ v) + Oi @ (Oi.T @ v)) / Ns
    Ob = Oc.conj()
    return lambda v: eps * v + (Ob @ (Ob.conj().T @ v)) / Ns


def solve_cg(Oc, F, eps, tol, real_params, maxiter=None):
    P = Oc.shape[0]
    mv = sr_matvec_factory(Oc, eps, real_params)
    x, it, ok = cg(mv, F, tol, 10 * P if maxiter is None else maxiter)
    if real_params:
        x = np.real(x)
    return x, it, ok


def solve_cg_explicit(S, F, eps, tol, maxiter=None):
    """CG on an explicit S + eps I (SRIterative.jl with full_matrix=true, :88, :117-125)."""
    P = S.shape[0]
    return cg(lambda v: S @ v + eps * v, F, tol, 10 * P if maxiter is None else maxiter)


def descent_update(w, dw, eta):
    """Optimisers.Descent: w <- w - eta*dw."""
    return w - eta * dw


# ----- whole iteration drivers (sample batches supplied) --------------------------------
def iteration_ket(net, op, sigmas, eps, real_params=None):
    """BatchedValSampler.sample! (:117-139) without the sampling step + SRDirect."""
    from .estimators import local_scalar_ket
    out, O = net.logpsi_grad(sigmas)
    Eloc = local_scalar_ket(net, op, sigmas, out)
    avg, Oc = center(O)
    gradC = force_ket(Eloc, Oc)
    rp = (not net.is_complex) if real_params is None else real_params
    S, F = sr_setup(Oc, gradC, rp)
    dw = solve_cholesky(S, F, eps)
    return dict(logpsi=out, O=O, Eloc=Eloc, O_avg=avg, gradC=gradC, S=S, F=F, dw=dw)


def iteration_liouvillian(net, liouv, srow, scol, eps, real_params=None):
    """BatchedGradSampler.sample! (:77-123) without the sampling step + SRDirect."""
    from .estimators import local_grad_super
    out, O = net.logpsi_grad(srow, scol)
    Lloc, gL = local_grad_super(net, liouv, srow, scol, out)
    avg, Oc = center(O)
    gradC = force_liouvillian(Lloc, gL, avg)
    rp = (not net.is_complex) if real_params is None else real_params
    S, F = sr_setup(Oc, gradC, rp)
    dw = solve_cholesky(S, F, eps)
    return dict(logpsi=out, O=O, Lloc=Lloc, gLloc=gL, O_avg=avg, gradC=gradC, S=S, F=F, dw=dw,
                C=np.mean(np.abs(Lloc) ** 2))


def solve_minres_explicit(S, F, eps, tol, maxiter=None):
    """MINRES (Paige & Saunders 1975) on A = S + eps I, x0 = 0, no preconditioner, stop when the recurrence residual
    phibar <= tol ||F||.  The reference's sr_minres branch calls IterativeSolvers 0.8.1 `minres` (SRIterative.jl:101-125);
    neither its iterates nor its iteration count are pinned by any reference test, so this restatement is only used
    for the converged solution and an iteration-count sanity check.  Returns (x, iterations, converged)."""
    A = np.asarray(S, dtype=np.complex128 if np.iscomplexobj(S) or np.iscomplexobj(F) else np.float64)
    A = A + eps * np.eye(A.shape[0])
    b = np.asarray(F, dtype=A.dtype)
    P = b.size
    maxiter = 10 * P if maxiter is None else maxiter
    x = np.zeros(P, A.dtype)
    r1 = np.zeros(P, A.dtype)
    r2 = b.copy()
    beta1 = np.sqrt(np.vdot(r2, r2).real)
    if beta1 == 0.0:
        return x, 0, True
    oldb, beta, dbar, epsln, phibar, cs, sn = 0.0, beta1, 0.0, 0.0, beta1, -1.0, 0.0
    wa, wb = np.zeros(P, A.dtype), np.zeros(P, A.dtype)
    it = 0
    while it < maxiter and not (phibar <= tol * beta1):
        it += 1
        v = r2 / beta
        y = A @ v
        if it >= 2:
            y = y - (beta / oldb) * r1
        alfa = np.vdot(v, y).real
        y = y - (alfa / beta) * r2
        r1, r2 = r2, y
        oldb = beta
        beta = np.sqrt(np.vdot(r2, r2).real)
        oldeps, delta, gbar = epsln, cs * dbar + sn * alfa, sn * dbar - cs * alfa
        epsln, dbar = sn * beta, -cs * beta
        gamma = max(np.sqrt(gbar * gbar + beta * beta), 1e-300)
        cs, sn = gbar / gamma, beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        wn = (v - oldeps * wa - delta * wb) / gamma
        x = x + phi * wn
        wa, wb = wb, wn
        if beta == 0.0:
            break
    return x, it, bool(phibar <= tol * beta1 or beta == 0.0)
