"""Oracle: Hilbert spaces (value encoding, index order, flips).  Test infrastructure only.

Follows src/Hilbert/HomogeneousSpin.jl and src/Hilbert/HomogeneousFock.jl (local dimension 2
only: the B200 path packs configurations as bit-vectors) and src/Hilbert/SuperHilbert.jl.
"""
import numpy as np


class Hilbert:
    """Homogeneous d=2 space.  kind = "spin" (values -1/+1) or "fock" (values 0/1)."""

    def __init__(self, n, kind):
        assert kind in ("spin", "fock")
        self.n = int(n)
        self.kind = kind
        self.d = 2

    # value <-> digit.  HomogeneousSpin.jl:116-125 (set!: value = digit*2-(N-1)),
    # HomogeneousFock.jl:93-101 (value = digit).
    def value(self, digit):
        digit = np.asarray(digit)
        return (2 * digit - 1) if self.kind == "spin" else digit

    def digit(self, value):
        value = np.asarray(value)
        if self.kind == "spin":
            return ((np.real(value) + 1) / 2).astype(np.int64)
        return np.real(value).astype(np.int64)

    def spacedim(self):
        return 2 ** self.n

    def state(self, index):
        """set!(sigma, h, index): 1-based index, site 1 least significant.
        HomogeneousSpin.jl:116-125 / HomogeneousFock.jl:93-101."""
        v = int(index) - 1
        digs = np.array([(v >> i) & 1 for i in range(self.n)], dtype=np.int64)
        return self.value(digs).astype(np.float64)

    def toint(self, sigma):
        """HomogeneousSpin.jl:156-162 / HomogeneousFock.jl:140-146 (1-based)."""
        d = self.digit(sigma)
        return int(sum(int(d[i]) << i for i in range(self.n))) + 1

    def local_index(self, sigma, sites):
        """HomogeneousSpin.jl:171-179 / HomogeneousFock.jl:155-163.  `sites` 1-based."""
        d = self.digit(sigma)
        return 1 + sum(int(d[s - 1]) * 2 ** i for i, s in enumerate(sites))

    def flip_value(self, v):
        """flipat! for local dim 2: HomogeneousSpin.jl:95-104, HomogeneousFock.jl:72-81."""
        if self.kind == "spin":
            return -1.0 if v == 1.0 else 1.0
        return 1.0 if v == 0.0 else 0.0

    def all_states(self):
        return np.stack([self.state(i + 1) for i in range(self.spacedim())], axis=1)

    def __eq__(self, o):
        return isinstance(o, Hilbert) and (self.n, self.kind) == (o.n, o.kind)


def HomogeneousSpin(n):
    return Hilbert(n, "spin")


def HomogeneousFock(n, d=2):
    assert d == 2
    return Hilbert(n, "fock")


def super_toint(h, row, col):
    """SuperHilbert.jl:50-56: index = (toint(col)-1)*D + toint(row)."""
    return (h.toint(col) - 1) * h.spacedim() + h.toint(row)


def super_state(h, index):
    """SuperHilbert.jl:23-36: i_r = div(i-1, D) -> col, remainder -> row."""
    i = int(index) - 1
    D = h.spacedim()
    i_r, i_c = divmod(i, D)
    return h.state(i_c + 1), h.state(i_r + 1)  # (row, col)
