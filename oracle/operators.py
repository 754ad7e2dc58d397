"""Oracle: k-local operators as per-local-row connection tables.  Test infrastructure only.

Restates (structure and ordering included, because the connection order is compared
bit-exactly with the CUDA enumeration):
  src/Operators/Operators/KLocalOperator.jl:54-112   (row tables), :183-199 (enumeration),
      :226-250 (_add_samesite!), :271-346 (products)
  src/Operators/Operators/KLocalOperatorSum.jl:65-111
  src/Operators/Operators/KLocalOperatorTensor.jl:19-43, 129-157, 195-204
  src/Operators/Operators/KLocalLiouvillian.jl:9-52
  src/Operators/OpConnections/OpConnection.jl:115-125 (merge rule)
  src/Operators/SimpleOperators.jl:11-147
  src/Operators/OpConversion.jl:6-24 (to_matrix)
"""
import numpy as np
from .hilbert import Hilbert, super_toint, super_state

THRESH = 10e-6  # KLocalOperator.jl:86 (`abs(val) < 10e-6 && continue`)


class KLocalOperator:
    """KLocalOperatorRow(hilb, sites, mat)  --  KLocalOperator.jl:54-112."""

    def __init__(self, hilb, sites, mat):
        self.hilb = hilb
        self.sites = [int(s) for s in sites]
        self.mat = np.array(mat, dtype=np.complex128)
        k = len(self.sites)
        D = 2 ** k
        assert self.mat.shape == (D, D)
        # local basis states: local index r (1-based) -> digits, site i least significant
        def loc_vals(r):
            digs = np.array([((r - 1) >> i) & 1 for i in range(k)])
            return hilb.value(digs)
        self.op_conns = []
        for r in range(1, D + 1):
            row = self.mat[r - 1]
            conns = [[row[r - 1], ((), ())]]          # diagonal always first (:75-81)
            st = loc_vals(r)
            for c in range(1, D + 1):
                if c == r:
                    continue
                val = row[c - 1]
                if abs(val) < THRESH:
                    continue
                st1 = loc_vals(c)
                cng = tuple(self.sites[i] for i in range(k) if st[i] != st1[i])
                nwv = tuple(float(st1[i]) for i in range(k) if st[i] != st1[i])
                conns.append([val, (cng, nwv)])
            self.op_conns.append(conns)

    def rebuilt(self, mat):            # KLocalOperator(op, mat)  :114-115
        return KLocalOperator(self.hilb, list(self.sites), mat)

    def duplicate(self):
        o = KLocalOperator.__new__(KLocalOperator)
        o.hilb, o.sites, o.mat = self.hilb, list(self.sites), self.mat.copy()
        o.op_conns = [[[m, c] for m, c in row] for row in self.op_conns]
        return o

    def add_samesite_(self, other):    # _add_samesite!  :226-250 (op_conns part: OpConnection.jl:115-125)
        assert self.sites == other.sites
        self.mat = self.mat + other.mat
        for cl, cr in zip(self.op_conns, other.op_conns):
            for mel, cng in cr:
                for e in cl:
                    if e[1] == cng:
                        e[0] = e[0] + mel
                        break
                else:
                    cl.append([mel, cng])
        return self

    def __neg__(self):
        return self.rebuilt(-self.mat)

    def transpose(self):
        return self.rebuilt(self.mat.T.copy())

    def conj(self):                    # conj! :255-262 keeps table structure
        o = self.duplicate()
        o.mat = o.mat.conj()
        for row in o.op_conns:
            for e in row:
                e[0] = np.conj(e[0])
        return o

    def adjoint(self):                 # adjoint = conj(transpose(op)) :265
        return self.transpose().conj()

    def scale(self, a):                # _op_alpha_prod :269-270
        return self.rebuilt(a * self.mat)

    def mul(self, opr):                # Base.:* :272-346
        opl = self
        if opl.sites == opr.sites:
            return opl.rebuilt(opl.mat @ opr.mat)
        disjoint = not any(s in opl.sites for s in opr.sites)
        if disjoint:
            assert len(opl.sites) == 1 and len(opr.sites) == 1, "not implemented in the reference"
            if opl.sites[0] > opr.sites[0]:
                opl, opr = opr, opl
            mat = np.kron(opr.mat, opl.mat)          # :303-305
            return KLocalOperator(opl.hilb, [opl.sites[0], opr.sites[0]], mat)
        assert len(opl.sites) == 1 or len(opr.sites) == 1, "not implemented in the reference"
        rev = False
        if len(opl.sites) == 1:
            opl, opr, rev = opr, opl, True
        idx = opl.sites.index(opr.sites[0])
        mats = [np.eye(2, dtype=np.complex128) for _ in opl.sites]
        mats[idx] = opr.mat
        mat_r = np.array([[1.0 + 0j]])
        for m in reversed(mats):                     # kron(reverse(matrices)...) :331
            mat_r = np.kron(mat_r, m)
        prod = mat_r @ opl.mat if rev else opl.mat @ mat_r
        return opl.rebuilt(prod)


class KLocalOperatorSum:
    """KLocalOperatorSum.jl.  Terms keep first-appearance order (`_add!` :74-85)."""

    def __init__(self, hilb, sites=None, operators=None):
        self.hilb = hilb
        self.sites = sites or []
        self.operators = operators or []

    def duplicate(self):
        return KLocalOperatorSum(self.hilb, [_copy_sites(s) for s in self.sites],
                                 [o.duplicate() for o in self.operators])

    def add_(self, op):
        if isinstance(op, KLocalOperatorSum):
            for o in op.operators:
                self.add_(o)
            return self
        s = op.sites
        for i, si in enumerate(self.sites):
            if si == s:
                self.operators[i].add_samesite_(op)
                return self
        self.sites.append(_copy_sites(s))
        self.operators.append(op.duplicate())
        return self

    def __neg__(self):
        return KLocalOperatorSum(self.hilb, [_copy_sites(s) for s in self.sites],
                                 [-o for o in self.operators])

    def transpose(self):
        return KLocalOperatorSum(self.hilb, [_copy_sites(s) for s in self.sites],
                                 [o.transpose() for o in self.operators])

    def conj(self):
        return KLocalOperatorSum(self.hilb, [_copy_sites(s) for s in self.sites],
                                 [o.conj() for o in self.operators])

    def adjoint(self):
        return self.transpose().conj()

    def scale(self, a):                # _op_alpha_prod :180-183: sum([a*op ...])
        return _sum([o.scale(a) for o in self.operators])


class KLocalOperatorTensor:
    """op_l (x) op_r on (row, col).  None stands for KLocalIdentity.  KLocalOperatorTensor.jl"""

    def __init__(self, op_l, op_r):
        self.op_l, self.op_r = op_l, op_r
        self.hilb = (op_l or op_r).hilb
        self.sites = (tuple(op_l.sites) if op_l is not None else (),
                      tuple(op_r.sites) if op_r is not None else ())

    def duplicate(self):
        return KLocalOperatorTensor(self.op_l.duplicate() if self.op_l is not None else None,
                                    self.op_r.duplicate() if self.op_r is not None else None)

    def add_samesite_(self, other):
        if self.op_l is not None:
            self.op_l.add_samesite_(other.op_l)
        if self.op_r is not None:
            self.op_r.add_samesite_(other.op_r)
        return self

    def scale(self, a):                # :195-204 (sqrt split for a full tensor, quirk Q14)
        if self.op_l is None:
            return KLocalOperatorTensor(None, self.op_r.scale(a))
        if self.op_r is None:
            return KLocalOperatorTensor(self.op_l.scale(a), None)
        s = np.sqrt(complex(a))
        return KLocalOperatorTensor(self.op_l.scale(s), self.op_r.scale(s))


class KLocalLiouvillian:
    def __init__(self, hilb, HnH_l, HnH_r, LLdag):
        self.hilb, self.HnH_l, self.HnH_r, self.LLdag = hilb, HnH_l, HnH_r, LLdag


def _copy_sites(s):
    return list(s) if isinstance(s, list) else tuple(tuple(x) for x in s)


def _sum(ops):
    acc = ops[0]
    for o in ops[1:]:
        acc = add(acc, o)
    return acc


# ----- algebra front-end (host side in the reference too) -------------------------------
def add(a, b):
    """`+` dispatch: KLocalZero.jl, KLocalOperatorSum.jl:96-111, KLocalOperatorTensor.jl:189-192."""
    if a is None:
        return b.duplicate()
    if b is None:
        return a.duplicate()
    if isinstance(a, KLocalOperatorSum):
        return a.duplicate().add_(b)
    if isinstance(b, KLocalOperatorSum):
        return b.duplicate().add_(a)          # op + ops = ops + op  (:93)
    if a.sites == b.sites:
        return a.duplicate().add_samesite_(b)
    return KLocalOperatorSum(a.hilb, [_copy_sites(a.sites)], [a.duplicate()]).add_(b)


def neg(a):
    return None if a is None else -a


def sub(a, b):
    if b is None:
        return a.duplicate()
    return add(a, neg(b)) if a is not None else neg(b)


def scale(alpha, a):
    return None if a is None else a.scale(alpha)


def mul(a, b):
    if a is None or b is None:
        return None
    if isinstance(a, KLocalOperatorSum):       # KLocalOperatorSum.jl:146-154
        out = a.duplicate()
        for i, o in enumerate(out.operators):
            n = o.mul(b)
            out.operators[i], out.sites[i] = n, list(n.sites)
        return out
    if isinstance(b, KLocalOperatorSum):
        out = b.duplicate()
        for i, o in enumerate(out.operators):
            n = a.mul(o)
            out.operators[i], out.sites[i] = n, list(n.sites)
        return out
    return a.mul(b)


def adjoint(a):
    return None if a is None else a.adjoint()


# ----- SimpleOperators.jl:11-147 (local dimension 2: S = 1/2) ---------------------------
def _one_site(h, i, mat):
    return KLocalOperator(h, [i], np.array(mat, dtype=np.complex128))


def sigmax(h, i):
    return _one_site(h, i, [[0, 1], [1, 0]])


def sigmay(h, i):          # diagm(-1=>D, 1=>-D), D = [1im]
    return _one_site(h, i, [[0, -1j], [1j, 0]])


def sigmaz(h, i):          # diagm(0 => [2m for m=S:-1:-S]) = diag(1,-1)
    return _one_site(h, i, [[1, 0], [0, -1]])


def sigmam(h, i):          # diagm(-1 => D)
    return _one_site(h, i, [[0, 0], [1, 0]])


def sigmap(h, i):          # diagm(1 => D)
    return _one_site(h, i, [[0, 1], [0, 0]])


def destroy(h, i):
    return _one_site(h, i, [[0, 1], [0, 0]])


def create(h, i):
    return _one_site(h, i, [[0, 0], [1, 0]])


def number(h, i):
    return _one_site(h, i, [[0, 0], [0, 1]])


# ----- Liouvillian: KLocalLiouvillian.jl:9-37 -------------------------------------------
def _tensor_left(op):      # KLocalOperatorTensor(HnH, Identity)  (Sum version :36-39)
    if isinstance(op, KLocalOperatorSum):
        return _sum([KLocalOperatorTensor(o, None) for o in op.operators])
    return KLocalOperatorTensor(op, None)


def _tensor_right(op):
    if isinstance(op, KLocalOperatorSum):
        return _sum([KLocalOperatorTensor(None, o) for o in op.operators])
    return KLocalOperatorTensor(None, op)


def liouvillian(H, Lops):
    HnH = None if H is None else H.duplicate()
    for L in Lops:
        HnH = add(HnH, mul(scale(-0.5j, adjoint(L)), L))          # :32
    hilb = (HnH if HnH is not None else Lops[0]).hilb
    HnH_l = scale(-1.0j, _tensor_left(HnH)) if HnH is not None else None       # :12
    HnH_r = scale(1.0j, _tensor_right(adjoint(HnH))) if HnH is not None else None  # :13 (quirk Q7)
    LL = [KLocalOperatorTensor(L, L.conj()) for L in Lops]                    # :15
    LLdag = _sum(LL) if LL else None
    return KLocalLiouvillian(hilb, HnH_l, HnH_r, LLdag)


# ----- enumeration ---------------------------------------------------------------------
def terms(op):
    if op is None:
        return []
    if isinstance(op, KLocalOperatorSum):
        return list(op.operators)
    return [op]


def connections_ket(op, sigma):
    """accumulate_connections! for a ket operator: KLocalOperatorSum.jl:65-72 ->
    KLocalOperator.jl:183-199.  Returns [(mel, (sites, new_values))], zero mels included."""
    out = []
    h = (op.hilb)
    for t in terms(op):
        r = h.local_index(sigma, t.sites)
        for mel, cng in t.op_conns[r - 1]:
            out.append((mel, cng))
    return out


def _tensor_connections(t, row, col):
    """KLocalOperatorTensor.jl:129-157."""
    h = t.hilb
    out = []
    if t.op_r is None:
        r = h.local_index(row, t.op_l.sites)
        for mel, cng in t.op_l.op_conns[r - 1]:
            out.append((mel, cng, None))
    elif t.op_l is None:
        r = h.local_index(col, t.op_r.sites)
        for mel, cng in t.op_r.op_conns[r - 1]:
            out.append((mel, None, cng))
    else:
        rr = h.local_index(row, t.op_l.sites)
        rc = h.local_index(col, t.op_r.sites)
        for mel_r, cng_r in t.op_l.op_conns[rr - 1]:
            for mel_c, cng_c in t.op_r.op_conns[rc - 1]:
                out.append((mel_r * mel_c, cng_r, cng_c))
    return out


def connections_super(liouv, row, col):
    """KLocalLiouvillian.jl:46-52: HnH_l, then HnH_r, then LLdag.
    Returns [(mel, changes_row|None, changes_col|None)], zero mels included."""
    out = []
    for grp in (liouv.HnH_l, liouv.HnH_r, liouv.LLdag):
        for t in terms(grp):
            out.extend(_tensor_connections(t, row, col))
    return out


def apply_changes(sigma, cng):
    """States/ApplyStateChanges.jl:38-43."""
    s = np.array(sigma, dtype=np.float64, copy=True)
    if cng is not None:
        for site, val in zip(*cng):
            s[site - 1] = val
    return s


def n_changes(cng):
    return 0 if cng is None else len(cng[0])


# ----- dense matrices (OpConversion.jl:6-24) -------------------------------------------
def to_matrix(op):
    if isinstance(op, KLocalLiouvillian):
        h = op.hilb
        D = h.spacedim()
        mat = np.zeros((D * D, D * D), dtype=np.complex128)
        for i in range(1, D * D + 1):
            row, col = super_state(h, i)
            for mel, cl, cr in connections_super(op, row, col):
                j = super_toint(h, apply_changes(row, cl), apply_changes(col, cr))
                mat[i - 1, j - 1] += mel
        return mat
    h = op.hilb
    D = h.spacedim()
    mat = np.zeros((D, D), dtype=np.complex128)
    for i in range(1, D + 1):
        s = h.state(i)
        for mel, cng in connections_ket(op, s):
            mat[i - 1, h.toint(apply_changes(s, cng)) - 1] += mel
    return mat
