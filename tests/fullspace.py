"""Full-Hilbert-space action of a ket operator driven ONLY by its connection tables (test infrastructure).

`compile_terms` turns a list of k-local terms -- each (sites, rows) with rows[r] = [(mel, global flip mask), ...] as
the oracle's `op_conns` or the product's `LocalTerm.rows` hold them -- into a diagonal vector and a list of
(flip mask, broadcastable coefficient table) passes over the state vector viewed as an N-dimensional 2x2x...x2 array.
Basis index = sum_j digit_j 2^j with site 1 least significant (HomogeneousSpin.jl:156-162), local row
r = sum_i digit(sites[i]) 2^i (HomogeneousSpin.jl:171-179): the same two rules the device enumeration uses, so a
wrong sign, bond, row order or flip mask in either table builder moves the spectrum.
"""
import numpy as np
from scipy.linalg import eigh_tridiagonal


def terms_from_oracle(op):
    """oracle.operators KLocalOperator / KLocalOperatorSum -> [(sites0, rows)]"""
    from oracle import operators as OO
    out = []
    for t in OO.terms(op):
        rows = []
        for conns in t.op_conns:
            row = []
            for mel, (cng_sites, _new_values) in conns:
                mask = 0
                for s in cng_sites:
                    mask |= 1 << (int(s) - 1)
                row.append((complex(mel), mask))
            rows.append(row)
        out.append(([int(s) - 1 for s in t.sites], rows))
    return out


def terms_from_product(op):
    """nqcuda.operators.LocalOperator (host tables, what nq_operator_create uploads) -> [(sites0, rows)]"""
    out = []
    for t in op.terms:
        sites0 = [int(s) - 1 for s in t.sites]
        rows = []
        for r in t.rows:
            row = []
            for mel, flip in r:
                mask = 0
                for i, s in enumerate(sites0):
                    if (int(flip) >> i) & 1:
                        mask |= 1 << s
                row.append((complex(mel), mask))
            rows.append(row)
        out.append((sites0, rows))
    return out


class CompiledOperator:
    def __init__(self, N, terms):
        self.N = N
        self.diag_tabs = []       # (sites0, table[2^k]) of the zero-flip entries
        self.flips = {}           # global mask -> list of (sites0, table[2^k])
        for sites0, rows in terms:
            k = len(sites0)
            by_mask = {}
            for r, row in enumerate(rows):
                for mel, mask in row:
                    by_mask.setdefault(mask, np.zeros(1 << k, dtype=np.complex128))[r] += mel
            for mask, tab in by_mask.items():
                if not np.any(tab != 0):
                    continue
                if mask == 0:
                    self.diag_tabs.append((tuple(sites0), tab))
                else:
                    self.flips.setdefault(mask, []).append((tuple(sites0), tab))
        self._diag = None

    def canonical(self):
        """Order-independent description, for exact comparison of two table builders."""
        d = {}
        for sites0, tab in self.diag_tabs:
            key = (0, sites0)
            d[key] = d.get(key, 0) + tab
        for mask, lst in self.flips.items():
            for sites0, tab in lst:
                key = (mask, sites0)
                d[key] = d.get(key, 0) + tab
        return d

    def _bcast(self, sites0, tab, real):
        # table indexed by r = sum_i digit(sites[i]) << i  ->  array broadcastable against v.reshape([2] * N),
        # where axis N-1-j carries the digit of site j (C order: last axis = least significant bit)
        N, k = self.N, len(sites0)
        nd = tab.reshape([2] * k)                       # axes (i = k-1, ..., 0)
        axes_global = [N - 1 - sites0[i] for i in range(k - 1, -1, -1)]
        order = np.argsort(axes_global)
        nd = nd.transpose(order)
        shape = [1] * N
        for ax in axes_global:
            shape[ax] = 2
        nd = nd.reshape(shape)
        return np.ascontiguousarray(nd.real) if real else nd

    def prepare(self):
        self.real = all(np.all(t.imag == 0) for _, t in self.diag_tabs) and \
            all(np.all(t.imag == 0) for lst in self.flips.values() for _, t in lst)
        dt = np.float64 if self.real else np.complex128
        diag = np.zeros([2] * self.N, dtype=dt)
        for sites0, tab in self.diag_tabs:
            diag += self._bcast(sites0, tab, self.real)
        self._diag = diag
        self._passes, self._fast = [], []
        for mask, lst in self.flips.items():
            axes = tuple(self.N - 1 - j for j in range(self.N) if (mask >> j) & 1)
            coef = None
            for sites0, tab in lst:
                b = self._bcast(sites0, tab, self.real)
                coef = b if coef is None else coef + b
            if self.real and len(axes) == 1 and np.all(coef == coef.flat[0]):
                self._fast.append((mask.bit_length() - 1, float(coef.flat[0])))     # one site flips, same element in every row
            else:
                self._passes.append((axes, coef))
        return self

    def matvec(self, v):
        """(H v)[s] = sum_c mel_c(s) v[eta_c(s)]: the action the local estimator samples (E_loc = (H psi)/psi)."""
        if self._diag is None:
            self.prepare()
        vn = v.reshape([2] * self.N)
        out = self._diag * vn
        for axes, coef in self._passes:
            out += coef * np.flip(vn, axes)
        out = out.reshape(-1)
        if self._fast:
            # single-site flips with a constant element (sigma^x fields): in-place adds on the two halves of the flipped
            # bit, multi-threaded through torch when it is there (same arithmetic, shared memory)
            try:
                import torch
                tv, to = torch.from_numpy(np.ascontiguousarray(v)), torch.from_numpy(out)
            except Exception:
                tv, to = v, out
            for c in sorted(set(c for _, c in self._fast)):
                bits = [b for b, cc in self._fast if cc == c]
                acc = to * 0
                for bit in bits:
                    a3, v3 = acc.reshape(-1, 2, 1 << bit), tv.reshape(-1, 2, 1 << bit)
                    a3[:, 0, :] += v3[:, 1, :]
                    a3[:, 1, :] += v3[:, 0, :]
                acc *= c
                to += acc
        return out


def lanczos_ground_energy(matvec, dim, tol=1e-12, maxit=200, dtype=np.float64):
    """Lowest eigenvalue by plain Lanczos (three vectors, no re-orthogonalisation: fine for the extreme eigenvalue),
    started from the uniform vector; stops when the lowest Ritz value moves by < tol (relative) twice in a row."""
    v = np.full(dim, 1.0 / np.sqrt(dim), dtype=dtype)
    v_prev = np.zeros(dim, dtype=dtype)
    alphas, betas = [], []
    beta, last, calm = 0.0, None, 0
    for it in range(maxit):
        w = matvec(v)
        a = float(np.real(np.vdot(v, w)))
        w -= a * v
        if it:
            w -= beta * v_prev
        alphas.append(a)
        e0 = eigh_tridiagonal(np.array(alphas), np.array(betas), select="i", select_range=(0, 0))[0][0] if betas else a
        if last is not None and abs(e0 - last) <= tol * abs(e0):
            calm += 1
            if calm >= 2:
                return e0, it + 1
        else:
            calm = 0
        last = e0
        beta = float(np.linalg.norm(w))
        if beta < 1e-14:
            return e0, it + 1
        betas.append(beta)
        v_prev, v = v, w / beta
    return last, maxit
