"""Shared helpers for the parity tests: matched (oracle, product) objects and error metrics."""
import numpy as np

from oracle import machines as OM
from oracle import models as OMOD

TOL = {np.dtype(np.float64): 1e-11, np.dtype(np.complex128): 1e-11,
       np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5}


def assert_close(x, ref, tol, what="", floor=None):
    """norm-wise relative error <= tol AND element-wise
    |x - ref| <= tol * (|ref| + floor * max|ref|), floor = 1e-3 in FP64 mode.  In FP32 mode (tol >= 1e-6, compared with
    the FP64 oracle) every stored parameter, table entry and partial sum carries a rounding of 6e-8 of the LARGEST term
    it was built from: an element that cancels to << max|ref| cannot be held to 1e-5 of its own size.  Measured on the
    B200 (E_loc of cfg1 in complex64: N = 10 ratio terms of M = 20 FP32 factors each; log psi of N = 128, alpha = 8):
    absolute errors of 1-3e-7 max|ref|.  The floor there is 1e-1 max|ref|, i.e. every element is held to 1e-6 of the
    largest one (round 1 used floor = 1: 1e-5 of the largest)."""
    if floor is None:
        floor = 1e-3 if tol < 1e-6 else 1e-1
    x = np.asarray(x)
    ref = np.asarray(ref)
    assert x.shape == ref.shape, (what, x.shape, ref.shape)
    assert np.all(np.isfinite(x)), what + ": non-finite values"
    if ref.size == 0:
        return
    nrm = np.linalg.norm(ref.ravel())
    err = np.linalg.norm((x - ref).ravel())
    assert err <= tol * max(nrm, 1e-300), "%s: norm-wise rel err %.3e > %.1e" % (what, err / max(nrm, 1e-300), tol)
    bound = tol * (np.abs(ref) + floor * np.abs(ref).max())
    worst = np.max(np.abs(x - ref) - bound)
    assert worst <= 0, "%s: element-wise excess %.3e" % (what, worst)


def make_pair(nq, ctx, kind, hilb_kind, N, alpha, dtype, act=0, seed=1234, std=0.1, alpha_a=None):
    """Oracle machine (float64/complex128 copy of the parameters as stored in `dtype`) + product machine."""
    dtype = np.dtype(dtype)
    cplx = dtype.kind == "c"
    om = OM.random_machine(kind, N, alpha, act=act, complex_weights=cplx, seed=seed, std=std, alpha_a=alpha_a)
    w = om.params().astype(dtype)
    om.set_params(w.astype(np.complex128 if cplx else np.float64))     # oracle sees the rounded parameters
    hilb = nq.HomogeneousSpin(N) if hilb_kind == "spin" else nq.HomogeneousFock(N)
    if kind == "rbm":
        pm = nq.RBM(ctx, hilb, dtype, alpha, act)
    elif kind == "rbmsplit":
        pm = nq.RBMSplit(ctx, hilb, dtype, alpha)
    else:
        pm = nq.NDM(ctx, hilb, dtype, alpha, alpha if alpha_a is None else alpha_a, act)
    assert pm.P == om.P
    pm.set_params(w)
    return om, pm, hilb


def ohilb(hilb_kind, N):
    from oracle.hilbert import HomogeneousFock, HomogeneousSpin
    return HomogeneousSpin(N) if hilb_kind == "spin" else HomogeneousFock(N)


def rand_states(hilb_kind, N, B, seed):
    return np.asfortranarray(OMOD.random_states(ohilb(hilb_kind, N), B, seed))


# ---- the same physical models written with the PRODUCT's operator algebra (nqcuda/models.py) ----
def p_tfim_1d(nq, N, h=1.0, J=1.0):
    return nq.models.tfim_1d(N, h, J)


def p_tfim_2d(nq, Lx, h=3.0, J=1.0):
    return nq.models.tfim_2d(Lx, h, J)


def p_lindblad_ising_1d(nq, N, g=0.4, V=2.0, fock=True):
    return nq.models.lindblad_ising_1d(N, g, V, fock)


def enumerate_tables(tb, bits_row, bits_col=None):
    """Host enumeration of the flattened tables for ONE configuration given as digit arrays
    (what the device kernel does; used by CPU tests of the table builder)."""
    out = []
    site_ptr = np.concatenate([[0], np.cumsum(tb["part_nsites"])])
    row0 = np.concatenate([[0], np.cumsum(2 ** tb["part_nsites"].astype(np.int64))])

    def rows(p, bits):
        s = tb["part_sites"][site_ptr[p]:site_ptr[p + 1]]
        r = sum(int(bits[j]) << i for i, j in enumerate(s))
        e0, e1 = tb["row_ptr"][row0[p] + r], tb["row_ptr"][row0[p] + r + 1]
        res = []
        for e in range(e0, e1):
            m = 0
            for i, j in enumerate(s):
                if (int(tb["entry_flip"][e]) >> i) & 1:
                    m |= 1 << int(j)
            res.append((complex(tb["entry_mel"][e]), m))
        return res
    for t in range(tb["n_terms"]):
        Lp, Rp = int(tb["term_left"][t]), int(tb["term_right"][t])
        if Lp >= 0 and Rp < 0:
            out += [(m, f, 0) for m, f in rows(Lp, bits_row)]
        elif Lp < 0 and Rp >= 0:
            out += [(m, 0, f) for m, f in rows(Rp, bits_col)]
        else:
            for ml, fl in rows(Lp, bits_row):
                for mr, fr in rows(Rp, bits_col):
                    out.append((ml * mr, fl, fr))
    return out
