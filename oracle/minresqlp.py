"""Oracle: MINRES-QLP as the reference runs it (src/External/IterativeSolvers/minresqlp.jl).  Test infrastructure only.

The reference file is a Julia port of Choi-Paige-Saunders' minresQLP (via its Python port).  This restatement follows the
reference's EFFECTIVE control flow, line by line, with its deviations from the published algorithm kept and named:

  Q19  the MINRES/QLP switch reads `Acond < TranCond && flag != flag0 && QLPiter == 0` (minresqlp.jl:285); `flag == flag0`
       holds for every iteration that is still running, so the MINRES branch is never taken: the solver performs QLP
       updates from the first iteration (the published algorithm starts in MINRES mode and transfers when Acond grows).
       The transfer block (:298-310) -- which references undefined locals -- is therefore only reached with iters == 1,
       where it is skipped.
  Q20  `gmin = min(gminl2, gamal, abs_gama)` (:410) assigns a local: the iterable's gmin stays the first iteration's gama,
       and Acond = Anorm / gama_1.
  Q21  the constructor passes TranCond twice (:150): Acondlim = TranCond = 1e7 (the published default is 1e15), and
       maxxnorm = 1e7: flag 7 / flag 6 stop the solver early on ill-conditioned or large-norm systems and roll the
       iteration counters back (:453-459).
  Q22  `converged(m) = resnorm <= tol || flag != flag0` (:154): every way of stopping counts as converged, including
       flag 8 (iteration limit), so SRIterative.jl's "not converged -> restart" ladder (:133-150) never fires after
       sr_qlp; it can fire after sr_cg / sr_minres, where the restart call `minresqlp(dw, S, F; ...)` has one positional
       argument too many (MethodError).  The host mirror implements the ladder the code describes: up to 5 warm-started
       QLP runs with the default tolerance sqrt(eps), then a zero update.
  shift = 0, no preconditioner, real Lanczos scalars (Hermitian A).

Returns (x, info) with info = dict(flag, iters, qlp_iters, relres, relAres, Anorm, Acond, xnorm).
"""
import numpy as np


def sym_givens(a, b):
    """SymGivens (minresqlp.jl:486-513)."""
    if b == 0:
        c = 1.0 if a == 0 else float(np.sign(a))
        return c, 0.0, abs(a)
    if a == 0:
        return 0.0, float(np.sign(b)), abs(b)
    if abs(b) > abs(a):
        t = a / b
        s = np.sign(b) / np.sqrt(1 + t * t)
        c = s * t
        return c, s, b / s
    t = b / a
    c = np.sign(a) / np.sqrt(1 + t * t)
    s = c * t
    return c, s, a / c


def minresqlp(matvec, b, tol=None, maxiter=None, x0=None, TranCond=10e6, maxxnorm=10e6):
    """minresqlp(A, b; tol, maxiter) -- reference defaults tol = sqrt(eps), maxiter = size(A, 2)."""
    b = np.asarray(b)
    T = np.complex128 if np.iscomplexobj(b) else np.float64
    n = b.size
    tol = np.sqrt(np.finfo(np.float64).eps) if tol is None else tol
    maxiter = n if maxiter is None else maxiter
    Acondlim = TranCond                                   # Q21
    realmin = np.finfo(np.float64).tiny
    x = np.zeros(n, T)
    # warm start (the ladder of SRIterative.jl:133-150): iterate on the residual of x0, add at the end
    if x0 is not None:
        x0 = np.asarray(x0, T)
        b = b - matvec(x0)
    r1 = np.zeros(n, T)
    r2 = b.astype(T).copy()
    r3 = r2.copy()
    beta1 = float(np.linalg.norm(r2))
    info = dict(flag=0, iters=0, qlp_iters=0, relres=0.0, relAres=0.0, Anorm=0.0, Acond=1.0, xnorm=0.0)
    if beta1 == 0.0:
        return (x if x0 is None else x + x0), info
    relres = beta1 / (beta1 + 1e-50)
    w = np.zeros(n, T); wl = np.zeros(n, T); xl2 = np.zeros(n, T); wl2 = np.zeros(n, T)
    flag0 = -2
    flag = flag0
    iters = QLPiter = 0
    beta = tau = taul = 0.0
    phi = betan = beta1
    gmin = 0.0
    cs, sn, cr1, sr1, cr2, sr2 = -1.0, 0.0, -1.0, 0.0, -1.0, 0.0
    dltan = eplnn = gama = gamal = gamal2 = 0.0
    eta = etal = etal2 = vepln = veplnl = veplnl2 = 0.0
    ul3 = ul2 = ul = u = 0.0
    rnorm = beta1
    xnorm = xl2norm = Axnorm = Anorm = 0.0
    Acond = 1.0
    gminl = 0.0
    relAres = 0.0
    resnorm = 123.0
    iteration = 1
    while not (iteration > maxiter or resnorm <= tol or flag != flag0):        # done(m, iteration)
        iteration += 1
        iters += 1
        betal = beta
        beta = betan
        v = r3 / beta
        r3 = matvec(v)
        if iters > 1:
            r3 = r3 - (beta / betal) * r1
        alfa = float(np.real(np.vdot(r3, v)))
        r3 = r3 - (alfa / beta) * r2
        r1 = r2
        r2 = r3
        betan = float(np.linalg.norm(r3))
        if iters == 1 and betan == 0.0:
            if alfa == 0.0:
                flag = 0
            else:
                flag = -1
                x = b / alfa
            break
        pnorm = np.sqrt(betal ** 2 + alfa ** 2 + betan ** 2)
        # previous left rotation Q_{k-1}
        dbar = dltan
        dlta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        eplnn = sn * betan
        dltan = -cs * betan
        # current left rotation Q_k
        gamal2 = gamal
        gamal = gama
        cs, sn, gama = sym_givens(gbar, betan)
        taul2 = taul
        taul = tau
        tau = cs * phi
        Axnorm = np.sqrt(Axnorm ** 2 + tau ** 2)
        phi = sn * phi
        # previous right rotation P_{k-2,k}
        if iters > 2:
            veplnl2 = veplnl
            etal2 = etal
            etal = eta
            dlta_tmp = sr2 * vepln - cr2 * dlta
            veplnl = cr2 * vepln + sr2 * dlta
            dlta = dlta_tmp
            eta = sr2 * gama
            gama = -cr2 * gama
        # current right rotation P_{k-1,k}
        if iters > 1:
            cr1, sr1, gamal = sym_givens(gamal, dlta)
            vepln = sr1 * gama
            gama = -cr1 * gama
        # update xnorm
        ul4 = ul3
        ul3 = ul2
        if iters > 2:
            ul2 = (taul2 - etal2 * ul4 - veplnl2 * ul3) / gamal2
        if iters > 1:
            ul = (taul - etal * ul3 - veplnl * ul2) / gamal
        xnorm_tmp = np.sqrt(xl2norm ** 2 + ul2 ** 2 + ul ** 2)
        if abs(gama) > realmin and xnorm_tmp < maxxnorm:
            u = (tau - eta * ul2 - vepln * ul) / gama
            if np.sqrt(xnorm_tmp ** 2 + u ** 2) > maxxnorm:
                u = 0.0
                flag = 6
        else:
            u = 0.0
            flag = 9
        xl2norm = np.sqrt(xl2norm ** 2 + ul2 ** 2)
        xnorm = np.sqrt(xl2norm ** 2 + ul ** 2 + u ** 2)
        # update w and x: always the QLP branch (Q19)
        QLPiter += 1
        if iters == 1:
            wl2 = wl
            wl = v * sr1
            w = -v * cr1
        elif iters == 2:
            wl2 = wl
            wl, w = w * cr1 + v * sr1, w * sr1 - v * cr1
        else:
            wl2 = wl
            wl = w
            w = wl2 * sr2 - v * cr2
            wl2 = wl2 * cr2 + v * sr2
            v = wl * cr1 + w * sr1
            w = wl * sr1 - w * cr1
            wl = v
        xl2 = xl2 + wl2 * ul2
        x = xl2 + wl * ul + w * u
        # next right rotation P_{k-1,k+1}
        cr2, sr2, gamal = sym_givens(gamal, eplnn)
        # norms
        abs_gama = abs(gama)
        Anorm = max(Anorm, pnorm, gamal, abs_gama)
        if iters == 1:
            gmin = gama
            gminl = gmin
        else:
            gminl = gmin                                  # Q20: gmin itself is never updated
        Acondl = Acond
        Acond = Anorm / gmin
        rnorml = rnorm
        relresl = relres
        if flag != 9:
            rnorm = phi
        relres = rnorm / (Anorm * xnorm + beta1)
        rootl = np.sqrt(gbar ** 2 + dltan ** 2)
        relAres = rootl / Anorm
        epsx = Anorm * xnorm * np.finfo(np.float64).eps
        if flag == flag0 or flag == 9:
            t1 = 1 + relres
            t2 = 1 + relAres
            if iters >= maxiter:
                flag = 8
            if Acond >= Acondlim:
                flag = 7
            if xnorm >= maxxnorm:
                flag = 6
            if epsx >= beta1:
                flag = 5
            if t2 <= 1:
                flag = 4
            if t1 <= 1:
                flag = 3
            if relAres <= tol:
                flag = 2
            if relres <= tol:
                flag = 1
        if flag in (2, 4, 6, 7):
            iters -= 1
            Acond = Acondl
            rnorm = rnorml
            relres = relresl
        resnorm = min(relres, relAres)
    info.update(flag=flag, iters=iters, qlp_iters=QLPiter, relres=relres, relAres=relAres, Anorm=Anorm, Acond=Acond,
                xnorm=xnorm)
    return (x if x0 is None else x + x0), info


def solve_qlp_explicit(S, F, eps, tol, maxiter=None):
    """sr_qlp on the explicit S (SRIterative.jl:84-100): A = S + eps I, maxiter = 10 P.  Real-parameter nets keep
    Re(x) (:128-132).  Returns (x, info)."""
    A = np.asarray(S) + eps * np.eye(np.asarray(S).shape[0])
    P = A.shape[0]
    return minresqlp(lambda v: A @ v, np.asarray(F), tol=tol, maxiter=10 * P if maxiter is None else maxiter)
