"""k-local operators: host-side algebra and the flattened connection tables uploaded to the GPU.

The operator *algebra* (+, -, *, adjoint, kron products, liouvillian) stays on the host exactly as
in the reference; what goes to the device is the per-local-row table of every term
(mel, flip mask) and the term list, and the per-sample enumeration runs there (nq_operator.cu).

For local dimension 2 a "change list" (sites, new values) of the reference is fully described by
which sites flip, i.e. by the XOR of the local row and column indices: flip = r ^ c.

ref: src/Operators/Operators/KLocalOperator.jl:54-112, :226-250, :271-346;
     KLocalOperatorSum.jl:74-111; KLocalOperatorTensor.jl:19-43, :195-204;
     KLocalLiouvillian.jl:9-37; OpConnections/OpConnection.jl:115-125; SimpleOperators.jl:11-147.
"""
import ctypes as C

import numpy as np

from . import _lib as L

_CUT = 10e-6   # KLocalOperator.jl:86


class LocalTerm:
    """One KLocalOperator: `sites` (1-based, local digit i <-> sites[i]), dense `mat`, and per local
    row the ordered entries [mel, flip] with flip = r ^ c; entry 0 is always the diagonal."""

    __slots__ = ("hilb", "sites", "mat", "rows")

    def __init__(self, hilb, sites, mat, rows=None):
        self.hilb = hilb
        self.sites = tuple(int(s) for s in sites)
        self.mat = np.array(mat, dtype=np.complex128)
        D = 1 << len(self.sites)
        if self.mat.shape != (D, D):
            raise ValueError("matrix shape %r does not match %d sites" % (self.mat.shape, len(self.sites)))
        if rows is None:
            rows = []
            for r in range(D):
                keep = [c for c in range(D) if c != r and abs(self.mat[r, c]) >= _CUT]
                rows.append([[self.mat[r, r], 0]] + [[self.mat[r, c], r ^ c] for c in keep])
        self.rows = rows

    def copy(self):
        return LocalTerm(self.hilb, self.sites, self.mat, [[list(e) for e in row] for row in self.rows])

    def merge_(self, other):
        """_add_samesite!: matrices add; row entries merge by identical change list, new ones append."""
        self.mat = self.mat + other.mat
        for mine, theirs in zip(self.rows, other.rows):
            index = {e[1]: e for e in reversed(mine)}      # first occurrence wins
            for mel, flip in theirs:
                if flip in index:
                    index[flip][0] = index[flip][0] + mel
                else:
                    mine.append([mel, flip])
                    index[flip] = mine[-1]
        return self

    def with_matrix(self, mat):
        return LocalTerm(self.hilb, self.sites, mat)

    def conj(self):
        out = self.copy()
        out.mat = out.mat.conj()
        for row in out.rows:
            for e in row:
                e[0] = np.conj(e[0])
        return out

    def product(self, right):
        a, b = self, right
        if a.sites == b.sites:
            return a.with_matrix(a.mat @ b.mat)
        if not set(a.sites) & set(b.sites):
            if len(a.sites) != 1 or len(b.sites) != 1:
                raise NotImplementedError("tensor product of multi-site operators (not in the reference either)")
            if a.sites[0] > b.sites[0]:
                a, b = b, a
            return LocalTerm(a.hilb, (a.sites[0], b.sites[0]), np.kron(b.mat, a.mat))
        if len(a.sites) != 1 and len(b.sites) != 1:
            raise NotImplementedError("product of overlapping multi-site operators (not in the reference either)")
        flipped = len(a.sites) == 1
        big, one = (b, a) if flipped else (a, b)
        pos = big.sites.index(one.sites[0])
        full = np.array([[1.0 + 0j]])
        for i in reversed(range(len(big.sites))):
            full = np.kron(full, one.mat if i == pos else np.eye(2))
        return big.with_matrix(full @ big.mat if flipped else big.mat @ full)


class LocalOperator:
    """KLocalOperatorZero / KLocalOperator / KLocalOperatorSum in one type: an ordered list of terms
    with distinct site sets (first-appearance order, KLocalOperatorSum.jl:74-85)."""

    def __init__(self, hilb, terms=(), is_sum=None):
        self.hilb = hilb
        self.terms = list(terms)
        # the reference distinguishes KLocalOperator from KLocalOperatorSum (it decides which operand
        # of `+` keeps its term order): track it
        self.is_sum = (len(self.terms) > 1) if is_sum is None else is_sum

    # -- algebra -----------------------------------------------------------------------
    def copy(self):
        return LocalOperator(self.hilb, [t.copy() for t in self.terms], self.is_sum)

    def _absorb(self, other):
        for t in other.terms:
            for mine in self.terms:
                if mine.sites == t.sites:
                    mine.merge_(t)
                    break
            else:
                self.terms.append(t.copy())
        return self

    def __add__(self, other):
        if isinstance(other, LocalOperator):
            if other.is_sum and not self.is_sum and self.terms:   # op + ops = ops + op
                out = other.copy()._absorb(self)
            else:
                out = self.copy()._absorb(other)
            out.is_sum = self.is_sum or other.is_sum or len(out.terms) > 1
            return out
        return NotImplemented

    def __neg__(self):
        return LocalOperator(self.hilb, [t.with_matrix(-t.mat) for t in self.terms])

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, other):
        if isinstance(other, LocalOperator):
            if not self.terms or not other.terms:
                return LocalOperator(self.hilb)
            if len(other.terms) == 1 and not other.is_sum:
                return LocalOperator(self.hilb, [t.product(other.terms[0]) for t in self.terms], self.is_sum)
            if len(self.terms) == 1 and not self.is_sum:
                return LocalOperator(self.hilb, [self.terms[0].product(t) for t in other.terms], other.is_sum)
            raise NotImplementedError("product of two operator sums (not in the reference either)")
        return self.__rmul__(other)

    def __rmul__(self, a):
        out = LocalOperator(self.hilb)
        for t in self.terms:                 # sum([a*op ...]) re-merges term by term
            out._absorb(LocalOperator(self.hilb, [t.with_matrix(a * t.mat)]))
        return out

    def __truediv__(self, a):
        return (1.0 / a) * self

    def transpose(self):
        return LocalOperator(self.hilb, [t.with_matrix(t.mat.T.copy()) for t in self.terms])

    def conj(self):
        return LocalOperator(self.hilb, [t.conj() for t in self.terms])

    def adjoint(self):
        return self.transpose().conj()

    dagger = adjoint

    # -- device tables ------------------------------------------------------------------
    def tables(self):
        return _flatten([(t, None) for t in self.terms], L.NQ_KET, self.hilb)

    def to_device(self, ctx):
        return DeviceOperator(ctx, self.tables(), self.hilb)

    def to_device_left(self, ctx):
        """The operator acting on the row index of a density matrix, O (x) 1: what the observable accumulator of a
        density-matrix machine visits (connections change sigma, sigma' stays; BatchedObsDMSampler.jl:94-101)."""
        return DeviceOperator(ctx, _flatten([(t, None) for t in self.terms], L.NQ_SUPER, self.hilb), self.hilb)


def KLocalOperatorRow(hilb, sites, mat):
    return LocalOperator(hilb, [LocalTerm(hilb, sites, mat)])


def _site(hilb, i, mat):
    return KLocalOperatorRow(hilb, [i], np.array(mat, dtype=np.complex128))


def sigmax(h, i): return _site(h, i, [[0, 1], [1, 0]])
def sigmay(h, i): return _site(h, i, [[0, -1j], [1j, 0]])
def sigmaz(h, i): return _site(h, i, [[1, 0], [0, -1]])
def sigmam(h, i): return _site(h, i, [[0, 0], [1, 0]])
def sigmap(h, i): return _site(h, i, [[0, 1], [0, 0]])
def destroy(h, i): return _site(h, i, [[0, 1], [0, 0]])
def create(h, i): return _site(h, i, [[0, 0], [1, 0]])
def number(h, i): return _site(h, i, [[0, 0], [0, 1]])


class Liouvillian:
    """KLocalLiouvillian: -i (H_nh (x) 1) + i (1 (x) H_nh') + sum_j L_j (x) conj(L_j), visited in
    that order (KLocalLiouvillian.jl:9-16, 46-52).  Each tensor term is (left part | None, right part | None)."""

    def __init__(self, hilb, groups):
        self.hilb = hilb
        self.groups = groups            # three lists of (LocalTerm|None, LocalTerm|None)

    def tables(self):
        return _flatten([p for g in self.groups for p in g], L.NQ_SUPER, self.hilb)

    def to_device(self, ctx):
        return DeviceOperator(ctx, self.tables(), self.hilb)


def liouvillian(H, jump_ops):
    hilb = H.hilb if H is not None else jump_ops[0].hilb
    HnH = H.copy() if H is not None else LocalOperator(hilb)
    for Lj in jump_ops:
        HnH = HnH + (-0.5j * Lj.adjoint()) * Lj
    left = [(t, None) for t in (-1.0j * HnH).terms]
    right = [(None, t) for t in (1.0j * HnH.adjoint()).terms]      # HnH' (quirk Q7)
    # sum of L (x) conj(L): same (sites_l, sites_r) pairs merge part-wise (KLocalOperatorTensor.jl:171-177)
    jumps = []
    for Lj in jump_ops:
        if len(Lj.terms) != 1:
            raise NotImplementedError("jump operators must be single local terms")
        l, r = Lj.terms[0].copy(), Lj.terms[0].conj()
        for pl, pr in jumps:
            if pl.sites == l.sites and pr.sites == r.sites:
                pl.merge_(l)
                pr.merge_(r)
                break
        else:
            jumps.append((l, r))
    return Liouvillian(hilb, [left, right, jumps])


def _flatten(pairs, space, hilb):
    """[(left LocalTerm|None, right LocalTerm|None)] -> arrays of nq_operator_create."""
    parts, part_id = [], {}

    def pid(t):
        if t is None:
            return -1
        if id(t) not in part_id:
            part_id[id(t)] = len(parts)
            parts.append(t)
        return part_id[id(t)]
    term_left = np.array([pid(l) for l, _ in pairs], dtype=np.int32)
    term_right = np.array([pid(r) for _, r in pairs], dtype=np.int32)
    nsites = np.array([len(p.sites) for p in parts], dtype=np.int32)
    sites = np.array([s - 1 for p in parts for s in p.sites], dtype=np.int32)
    row_ptr, mel, flip = [0], [], []
    for p in parts:
        for row in p.rows:
            for m, f in row:
                mel.append(complex(m))
                flip.append(f)
            row_ptr.append(len(mel))
    return dict(space=space, N=hilb.n, n_parts=len(parts), part_nsites=nsites, part_sites=sites,
                row_ptr=np.array(row_ptr, dtype=np.int64), entry_mel=np.array(mel, dtype=np.complex128),
                entry_flip=np.array(flip, dtype=np.uint32), n_terms=len(pairs),
                term_left=term_left, term_right=term_right)


class DeviceOperator:
    """nq_operator_t: the tables resident on the GPU."""

    def __init__(self, ctx, tb, hilb):
        self.ctx, self.hilb, self.space, self.tb = ctx, hilb, tb["space"], tb
        h = C.c_void_p()
        keep = [np.ascontiguousarray(tb[k]) for k in
                ("part_nsites", "part_sites", "row_ptr", "entry_mel", "entry_flip", "term_left", "term_right")]
        L.check(L.lib.nq_operator_create(ctx.h, tb["space"], tb["N"], tb["n_parts"], L.ptr(keep[0]), L.ptr(keep[1]),
                                         L.ptr(keep[2]), L.ptr(keep[3]), L.ptr(keep[4]), tb["n_terms"],
                                         L.ptr(keep[5]), L.ptr(keep[6]), C.byref(h)), ctx.h)
        self.h = h
        n = C.c_int64()
        L.check(L.lib.nq_operator_max_connections(h, C.byref(n)), ctx.h)
        self.max_connections = n.value

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                L.lib.nq_operator_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def connections(self, srow, scol=None):
        """row_valdiff over a batch: returns (counts[B], mels[B, max_conn], flips_row[B, max_conn, W64],
        flips_col | None) in reference order, zero matrix elements included."""
        srow = np.asfortranarray(srow)
        N, B = srow.shape
        W = L.lib.nq_states_words(N)
        mc = self.max_connections
        counts = np.zeros(B, dtype=np.int32)
        mels = np.zeros((B, mc), dtype=np.complex128)
        fr = np.zeros((B, mc, W), dtype=np.uint64)
        fc = np.zeros((B, mc, W), dtype=np.uint64) if scol is not None else None
        scol_f = np.asfortranarray(scol, dtype=srow.dtype) if scol is not None else None
        L.check(L.lib.nq_connections(self.h, self.hilb.code, L.ptr(srow), L.ptr(scol_f), L.nq_dtype(srow.dtype), B, mc,
                                     L.ptr(counts), L.ptr(mels), L.ptr(fr), L.ptr(fc)), self.ctx.h)
        return counts, mels, fr, fc
