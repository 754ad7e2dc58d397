"""Oracle: the BASELINE model builders, written with the oracle's operator algebra exactly as
the reference examples write them.  Test infrastructure only.

  examples/ising1d.jl:12-18           H = sum_i -h sx_i + J sz_i sz_{i+1}   (PBC)
  examples/ising2d.jl                 same on an LxL periodic square lattice
  examples/dissipative_ising1d.jl:10-27   H = sum_i g/2 sx_i + V/4 sz_i sz_{i+1}, L_i = sm_i
"""
import numpy as np
from .hilbert import HomogeneousSpin, HomogeneousFock
from . import operators as ops


def tfim_1d(N, h=1.0, J=1.0, hilb=None):
    hilb = hilb or HomogeneousSpin(N)
    H = None
    for i in range(1, N + 1):
        H = ops.sub(H, ops.scale(h, ops.sigmax(hilb, i)))
        H = ops.add(H, ops.mul(ops.scale(J, ops.sigmaz(hilb, i)), ops.sigmaz(hilb, i % N + 1)))
    return hilb, H


def tfim_2d(L, h=3.0, J=1.0):
    """Periodic LxL square lattice; site(x,y) = 1 + x + L*y; bonds to +x and +y neighbours."""
    N = L * L
    hilb = HomogeneousSpin(N)
    H = None
    for i in range(1, N + 1):
        H = ops.sub(H, ops.scale(h, ops.sigmax(hilb, i)))
    for y in range(L):
        for x in range(L):
            i = 1 + x + L * y
            for j in (1 + (x + 1) % L + L * y, 1 + x + L * ((y + 1) % L)):
                if i == j:
                    continue
                H = ops.add(H, ops.mul(ops.scale(J, ops.sigmaz(hilb, i)), ops.sigmaz(hilb, j)))
    return hilb, H


def lindblad_ising_1d(N, g=0.4, V=2.0, fock=True):
    hilb = HomogeneousFock(N, 2) if fock else HomogeneousSpin(N)
    H = None
    jumps = []
    for i in range(1, N + 1):
        H = ops.add(H, ops.scale(g / 2.0, ops.sigmax(hilb, i)))
        H = ops.add(H, ops.mul(ops.scale(V / 4.0, ops.sigmaz(hilb, i)), ops.sigmaz(hilb, i % N + 1)))
        jumps.append(ops.sigmam(hilb, i))
    return hilb, H, jumps, ops.liouvillian(H, jumps)


def random_states(hilb, B, seed=4321):
    """i.i.d. uniform configurations, [N, B] float64."""
    rng = np.random.Generator(np.random.Philox(seed))
    return hilb.value(rng.integers(0, 2, size=(hilb.n, B))).astype(np.float64)
