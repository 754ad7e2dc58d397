"""The models of the reference's example scripts, written with this package's operator algebra the way the scripts
write them (benchmarks, examples and the parity tests share these builders).

ref: examples/ising1d.jl:11-18, examples/ising2d.jl:17-28, examples/dissipative_ising1d.jl:10-27.
"""
from .core import HomogeneousFock, HomogeneousSpin
from .operators import LocalOperator, liouvillian, sigmam, sigmax, sigmaz


def tfim_1d(N, h=1.0, J=1.0):
    """H = sum_i -h sx_i + J sz_i sz_{i+1}, periodic chain (examples/ising1d.jl:14-18)."""
    hilb = HomogeneousSpin(N)
    H = LocalOperator(hilb)
    for i in range(1, N + 1):
        H = H - h * sigmax(hilb, i)
        H = H + (J * sigmaz(hilb, i)) * sigmaz(hilb, i % N + 1)
    return hilb, H


def tfim_2d(Lx, h=3.0, J=1.0):
    """Periodic Lx x Lx square lattice, site(x, y) = 1 + x + Lx y, bonds to the +x and +y neighbours."""
    N = Lx * Lx
    hilb = HomogeneousSpin(N)
    H = LocalOperator(hilb)
    for i in range(1, N + 1):
        H = H - h * sigmax(hilb, i)
    for y in range(Lx):
        for x in range(Lx):
            i = 1 + x + Lx * y
            for j in (1 + (x + 1) % Lx + Lx * y, 1 + x + Lx * ((y + 1) % Lx)):
                if i != j:
                    H = H + (J * sigmaz(hilb, i)) * sigmaz(hilb, j)
    return hilb, H


def lindblad_ising_1d(N, g=0.4, V=2.0, fock=True):
    """H = sum_i g/2 sx_i + V/4 sz_i sz_{i+1}, jump operators sigma^-_i (examples/dissipative_ising1d.jl:10-27).
    Returns (hilb, H, jumps, liouvillian)."""
    hilb = HomogeneousFock(N, 2) if fock else HomogeneousSpin(N)
    H = LocalOperator(hilb)
    jumps = []
    for i in range(1, N + 1):
        H = H + (g / 2.0) * sigmax(hilb, i)
        H = H + ((V / 4.0) * sigmaz(hilb, i)) * sigmaz(hilb, i % N + 1)
        jumps.append(sigmam(hilb, i))
    return hilb, H, jumps, liouvillian(H, jumps)
