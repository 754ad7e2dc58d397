"""The MINRES-QLP restatement (oracle/minresqlp.py, following External/IterativeSolvers/minresqlp.jl with its quirks
Q19-Q22) against dense linear algebra: parity of the iterates is unpinned by the reference (no test calls the solver),
so the oracle is pinned on what the algorithm is FOR: x = A^{-1} b on definite systems and the minimum-length
least-squares solution pinv(A) b on singular ones -- where plain MINRES diverges."""
import numpy as np

from oracle import minresqlp as Q
from oracle import sr as OSR


def test_spd_real_and_hermitian_complex():
    rng = np.random.default_rng(1)
    P = 60
    X = rng.standard_normal((P, 200))
    S = X @ X.T / 200
    F = rng.standard_normal(P)
    x, info = Q.solve_qlp_explicit(S, F, 1e-3, 1e-12)
    ref = np.linalg.solve(S + 1e-3 * np.eye(P), F)
    assert info["flag"] == 1 and info["iters"] <= P and np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)
    Xc = rng.standard_normal((P, 200)) + 1j * rng.standard_normal((P, 200))
    Sc = (Xc @ Xc.conj().T / 200).conj()
    Fc = rng.standard_normal(P) + 1j * rng.standard_normal(P)
    x, info = Q.solve_qlp_explicit(Sc, Fc, 1e-3, 1e-12)
    ref = np.linalg.solve(Sc + 1e-3 * np.eye(P), Fc)
    assert info["flag"] == 1 and np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)


def test_singular_systems_minimum_length_solution():
    rng = np.random.default_rng(2)
    P = 60
    X = rng.standard_normal((P, 20))
    S = X @ X.T / 20
    pinv = np.linalg.pinv(S)
    F = S @ rng.standard_normal(P)                       # consistent
    x, info = Q.minresqlp(lambda v: S @ v, F, tol=1e-12, maxiter=10 * P)
    assert info["flag"] == 1 and np.linalg.norm(x - pinv @ F) <= 1e-10 * np.linalg.norm(pinv @ F)
    F = rng.standard_normal(P)                           # inconsistent: least squares
    x, info = Q.minresqlp(lambda v: S @ v, F, tol=1e-12, maxiter=10 * P)
    assert info["flag"] in (2, 4, 6) and info["iters"] <= 21
    assert np.linalg.norm(x - pinv @ F) <= 1e-6 * np.linalg.norm(pinv @ F)
    xm, it, ok = OSR.solve_minres_explicit(S, F, 0.0, 1e-12)
    assert not ok and np.linalg.norm(xm) > 1e6 * np.linalg.norm(pinv @ F)      # what QLP is there for


def test_warm_start_and_quirks():
    rng = np.random.default_rng(3)
    P = 40
    X = rng.standard_normal((P, 100))
    A = X @ X.T / 100 + 0.01 * np.eye(P)
    b = rng.standard_normal(P)
    ref = np.linalg.solve(A, b)
    x, info = Q.minresqlp(lambda v: A @ v, b, tol=1e-6, maxiter=400, x0=ref * (1 + 1e-5))
    assert np.linalg.norm(x - ref) <= 1e-8 * np.linalg.norm(ref) and info["iters"] < P
    # Q19: every iteration is a QLP update; Q20 / Q21: Acond is Anorm / gama_1 and the limit is 1e7
    x, info = Q.minresqlp(lambda v: A @ v, b, tol=1e-12, maxiter=400)
    assert info["qlp_iters"] >= info["iters"] > 0
    big = np.diag(np.concatenate([[1.0], np.linspace(1e8, 2e8, P - 1)]))
    e1 = np.zeros(P); e1[0] = 1.0
    x, info = Q.minresqlp(lambda v: big @ v, e1 + 1e-8 * rng.standard_normal(P), tol=1e-16, maxiter=400)
    assert info["flag"] == 7                              # stops on Anorm / gama_1 >= 1e7 (Q20, Q21) and rolls back
    # sym_givens
    for a_, b_ in ((3.0, 4.0), (-3.0, 4.0), (0.0, 2.0), (2.0, 0.0), (0.0, 0.0), (5.0, -1.0)):
        c, s, r = Q.sym_givens(a_, b_)
        assert abs(c * a_ + s * b_ - r) < 1e-12 and abs(-s * a_ + c * b_) < 1e-12 and (abs(c * c + s * s - 1) < 1e-12)
