"""Pin the operator tables (oracle AND product) to the two exact energies the reference holds
(examples/ising1d.jl:46-47, examples/ising2d.jl:55): matrix-free Lanczos on the operator action
defined by the connection tables.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import operators as OO
from oracle.hilbert import HomogeneousSpin
import fullspace as FS

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "energies.json")))
RTOL = 1e-9


# ---- the models, written the way the example scripts write them -------------------------------------------------
def o_ising1d(N, h, J):                     # examples/ising1d.jl:11-18 with the oracle's algebra
    hilb = HomogeneousSpin(N)
    H = None
    for i in range(1, N + 1):
        H = OO.sub(H, OO.scale(h, OO.sigmax(hilb, i)))
        H = OO.add(H, OO.mul(OO.scale(J, OO.sigmaz(hilb, i)), OO.sigmaz(hilb, i % N + 1)))
    return H


def p_ising1d(nq, N, h, J):                 # same with the product's algebra
    hilb = nq.HomogeneousSpin(N)
    H = nq.LocalOperator(hilb)
    for i in range(1, N + 1):
        H = H - h * nq.sigmax(hilb, i)
        H = H + J * nq.sigmaz(hilb, i) * nq.sigmaz(hilb, i % N + 1)
    return H


def _coord(i, j, dims):                     # examples/ising2d.jl:15
    return i + (j - 1) * dims[1]


def o_ising2d(dims, h, J):                  # examples/ising2d.jl:17-28
    hilb = HomogeneousSpin(dims[0] * dims[1])
    H = None
    for i in range(1, dims[0] + 1):
        for j in range(1, dims[1] + 1):
            p = _coord(i, j, dims)
            H = OO.sub(H, OO.scale(h, OO.sigmax(hilb, p)))
            hop = OO.add(OO.mul(OO.scale(J, OO.sigmaz(hilb, p)), OO.sigmaz(hilb, _coord(i % dims[0] + 1, j, dims))),
                         OO.mul(OO.scale(J, OO.sigmaz(hilb, p)), OO.sigmaz(hilb, _coord(i, j % dims[1] + 1, dims))))
            H = OO.add(H, hop)
    return H


def p_ising2d(nq, dims, h, J):
    hilb = nq.HomogeneousSpin(dims[0] * dims[1])
    H = nq.LocalOperator(hilb)
    for i in range(1, dims[0] + 1):
        for j in range(1, dims[1] + 1):
            p = _coord(i, j, dims)
            H = H - h * nq.sigmax(hilb, p)
            hop = (J * nq.sigmaz(hilb, p) * nq.sigmaz(hilb, _coord(i % dims[0] + 1, j, dims)) +
                   J * nq.sigmaz(hilb, p) * nq.sigmaz(hilb, _coord(i, j % dims[1] + 1, dims)))
            H = H + hop
    return H


def _same_compiled(a, b):
    ca, cb = a.canonical(), b.canonical()
    assert set(ca) == set(cb)
    for k in ca:
        assert np.array_equal(ca[k], cb[k]), k


def test_fullspace_matvec_is_the_dense_matrix():
    """The table-driven action equals oracle.to_matrix (which reproduces test/Operators/ising.jl:28) on N = 6."""
    H = o_ising1d(6, 0.7, 1.3)
    comp = FS.CompiledOperator(6, FS.terms_from_oracle(H))
    dense = OO.to_matrix(H)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(64)
    assert np.allclose(comp.matvec(v), dense @ v, atol=1e-13)


def test_tfim_1d_N20_energy_from_oracle_and_product_tables(nq):
    g = GOLD["tfim_1d_N20"]
    N = g["N"]
    assert abs(g["energy"] - g["per_site"] * N) < 1e-10
    e = {}
    for name, terms in (("oracle", FS.terms_from_oracle(o_ising1d(N, g["h"], g["J"]))),
                        ("product", FS.terms_from_product(p_ising1d(nq, N, g["h"], g["J"])))):
        comp = FS.CompiledOperator(N, terms).prepare()
        assert comp.real
        e[name], its = FS.lanczos_ground_energy(comp.matvec, 1 << N)
        assert abs(e[name] - g["energy"]) <= RTOL * abs(g["energy"]), (name, e[name], g["energy"], its)
    assert abs(e["oracle"] - e["product"]) <= 1e-12 * abs(g["energy"])


def test_tfim_2d_5x5_energy_from_oracle_and_product_tables(nq):
    """2^25 states: the Lanczos run (under a minute) is done once, on the oracle's tables; the product's tables are
    required to compile to EXACTLY the same passes (same diagonal tables, flip masks and coefficients) and to give
    the same action on a random vector, which makes their spectrum the same number.
    Quirk Q18: the literal of examples/ising2d.jl:55 is the energy of the model with J = -1 (see golden/energies.json);
    the script's J = +1 Hamiltonian is frustrated on the odd torus.  The tables are pinned with the sign the literal
    belongs to; the J = +1 tables differ from them only in the sign of the 50 bond tables, checked below."""
    g = GOLD["tfim_2d_5x5"]
    dims = g["dims"]
    N = dims[0] * dims[1]
    J = g["J_of_the_literal"]
    co = FS.CompiledOperator(N, FS.terms_from_oracle(o_ising2d(dims, g["h"], J)))
    cp = FS.CompiledOperator(N, FS.terms_from_product(p_ising2d(nq, dims, g["h"], J)))
    _same_compiled(co, cp)
    co.prepare()
    e0, its = FS.lanczos_ground_energy(co.matvec, 1 << N, tol=1e-12)
    assert abs(e0 - g["energy"]) <= RTOL * abs(g["energy"]), (e0, g["energy"], its)
    cp.prepare()
    v = np.random.default_rng(1).standard_normal(1 << N)
    assert np.array_equal(co.matvec(v), cp.matvec(v))
    # the script's own sign: same tables with the bond elements negated
    cs = FS.CompiledOperator(N, FS.terms_from_product(p_ising2d(nq, dims, g["h"], g["J"]))).canonical()
    cf = cp.canonical()
    assert set(cs) == set(cf) and len([k for k in cs if k[0] == 0]) == 2 * N
    for k in cs:
        assert np.array_equal(cs[k], -cf[k] if k[0] == 0 else cf[k]), k
