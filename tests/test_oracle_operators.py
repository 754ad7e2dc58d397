"""Pins the oracle's operator tables the way the reference does:
test/Operators/operators.jl:24-107 (row tables reconstruct the local matrix and its kron
embeddings) and test/Operators/ising.jl:28-42 (to_matrix(H), to_matrix(liouvillian(H, sm)) equal
QuantumOptics dense matrices).  QuantumOptics is replaced by an independent kron construction in
its convention: first subsystem fastest, spin index 1 = up, spre(A) = 1 (x) A,
spost(B) = B^T (x) 1 acting on the column-major vec(rho)."""
import numpy as np
import pytest

from oracle import operators as ops
from oracle.hilbert import HomogeneousFock, HomogeneousSpin, super_state, super_toint
from oracle.models import lindblad_ising_1d, tfim_1d, tfim_2d

SX = np.array([[0, 1], [1, 0]], complex)
SY = np.array([[0, -1j], [1j, 0]], complex)
SZ = np.array([[1, 0], [0, -1]], complex)
SM = np.array([[0, 0], [1, 0]], complex)
I2 = np.eye(2, dtype=complex)


def embed(N, i, m):
    """QuantumOptics embed: subsystem 1 is the fastest index -> kron(I.., m at i, ..I) reversed."""
    out = np.array([[1.0 + 0j]])
    for s in range(N, 0, -1):
        out = np.kron(out, m if s == i else I2)
    return out


def dense_ising(N, g, V):
    H = np.zeros((2 ** N, 2 ** N), complex)
    for i in range(1, N + 1):
        H += g / 2 * embed(N, i, SX)
        H += V / 4 * embed(N, i, SZ) @ embed(N, i % N + 1, SZ)
    return H


def dense_liouvillian(H, Js):
    D = H.shape[0]
    Id = np.eye(D)
    spre = lambda A: np.kron(Id, A)
    spost = lambda B: np.kron(B.T, Id)
    L = -1j * (spre(H) - spost(H))
    for J in Js:
        JdJ = J.conj().T @ J
        L += spre(J) @ spost(J.conj().T) - 0.5 * spre(JdJ) - 0.5 * spost(JdJ)
    return L


@pytest.mark.parametrize("hk", ["spin", "fock"])
def test_row_tables_reconstruct_local_matrix(hk):
    h = HomogeneousSpin(4) if hk == "spin" else HomogeneousFock(4)
    rng = np.random.default_rng(0)
    m1 = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    op = ops.KLocalOperator(h, [3], m1)
    assert np.allclose(ops.to_matrix(op), embed(4, 3, m1))
    m2 = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    op2 = ops.KLocalOperator(h, [2, 4], m2)
    # local index = 1 + d(site2) + 2 d(site4): site 2 is the fast local digit
    full = np.zeros((16, 16), complex)
    for i in range(16):
        for j in range(16):
            di = [(i >> s) & 1 for s in range(4)]
            dj = [(j >> s) & 1 for s in range(4)]
            if di[0] == dj[0] and di[2] == dj[2]:
                full[i, j] = m2[di[1] + 2 * di[3], dj[1] + 2 * dj[3]]
    assert np.allclose(ops.to_matrix(op2), full)
    # diagonal entry is always first and kept even when zero; small off-diagonals dropped
    z = ops.KLocalOperator(h, [1], np.array([[0, 1e-6], [1, 0]]))
    assert z.op_conns[0] == [[0j, ((), ())]]
    assert len(z.op_conns[1]) == 2 and z.op_conns[1][0][1] == ((), ())


def test_products_and_sums():
    h = HomogeneousSpin(4)
    zz = ops.mul(ops.sigmaz(h, 3), ops.sigmaz(h, 1))
    assert zz.sites == [1, 3]
    assert np.allclose(ops.to_matrix(zz), embed(4, 3, SZ) @ embed(4, 1, SZ))
    xy = ops.mul(ops.sigmax(h, 2), ops.sigmay(h, 2))
    assert np.allclose(ops.to_matrix(xy), embed(4, 2, SX @ SY))
    s = ops.add(ops.add(ops.sigmax(h, 1), zz), ops.scale(0.5, ops.sigmay(h, 1)))
    assert [t.sites for t in s.operators] == [[1], [1, 3]]
    assert np.allclose(ops.to_matrix(s), embed(4, 1, SX + 0.5 * SY) + ops.to_matrix(zz))
    # joint product of a 2-site and a 1-site operator
    j = ops.mul(zz, ops.sigmax(h, 3))
    assert np.allclose(ops.to_matrix(j), embed(4, 3, SZ) @ embed(4, 1, SZ) @ embed(4, 3, SX))
    j2 = ops.mul(ops.sigmax(h, 3), zz)
    assert np.allclose(ops.to_matrix(j2), embed(4, 3, SX) @ embed(4, 3, SZ) @ embed(4, 1, SZ))
    assert np.allclose(ops.to_matrix(ops.adjoint(ops.sigmam(h, 2))), embed(4, 2, SM.conj().T))


def test_term_order_ising1d():
    """examples/ising1d.jl:15-18 => [1],[1,2],[2],[2,3],...,[N],[1,N]."""
    _, H = tfim_1d(5)
    assert [t.sites for t in H.operators] == [[1], [1, 2], [2], [2, 3], [3], [3, 4], [4], [4, 5], [5], [1, 5]]


@pytest.mark.parametrize("fock", [False, True])
def test_ising_hamiltonian_and_liouvillian_match_dense(fock):
    """test/Operators/ising.jl:28-42, N=4, g=0.7, V=2."""
    N, g, V = 4, 0.7, 2.0
    hilb, H, jumps, liouv = lindblad_ising_1d(N, g, V, fock=fock)
    Hq = dense_ising(N, g, V)
    assert np.allclose(ops.to_matrix(H), Hq, atol=1e-14)
    Jq = [embed(N, i, SM) for i in range(1, N + 1)]
    assert np.allclose(ops.to_matrix(ops.liouvillian(H, [])), dense_liouvillian(Hq, []), atol=1e-14)
    assert np.allclose(ops.to_matrix(ops.liouvillian(None, jumps)),
                       dense_liouvillian(np.zeros_like(Hq), Jq), atol=1e-14)
    assert np.allclose(ops.to_matrix(liouv), dense_liouvillian(Hq, Jq), atol=1e-14)


def test_tfim_dense():
    hilb, H = tfim_1d(4, h=0.8, J=1.3)
    Hq = sum(-0.8 * embed(4, i, SX) + 1.3 * embed(4, i, SZ) @ embed(4, i % 4 + 1, SZ) for i in range(1, 5))
    assert np.allclose(ops.to_matrix(H), Hq)
    hilb2, H2 = tfim_2d(3, h=3.0, J=1.0)
    Hm = ops.to_matrix(H2)
    assert np.allclose(Hm, Hm.conj().T)
    # 9 sites x 2 bonds each, all distinct for L=3
    assert sum(len(t.sites) == 2 for t in H2.operators) == 18


def test_super_index_roundtrip():
    h = HomogeneousFock(3)
    for i in range(1, 65):
        r, c = super_state(h, i)
        assert super_toint(h, r, c) == i
