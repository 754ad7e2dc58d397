"""CPU checks of the oracle restatements added for SURVEY 8f rows 3-4: full-space reconstruction / ExactSampler
(oracle/fullspace.py) and the non-local transition rules (oracle/rules.py)."""
import numpy as np

from oracle import fullspace as OF, machines as OM, operators as OO, rules as OR
from oracle.hilbert import HomogeneousFock, HomogeneousSpin
from oracle.models import lindblad_ising_1d, tfim_1d


def test_ket_and_densitymatrix_definitions():
    N = 3
    h = HomogeneousSpin(N)
    net = OM.random_machine("rbm", N, 2, act=OM.LOGCOSH, complex_weights=True, seed=1, std=0.3)
    psi = OF.ket(net, h)
    assert abs(np.linalg.norm(psi) - 1) < 1e-14
    for i in range(1, 2 ** N + 1):                      # psi[i] belongs to set_index!(v, hilb, i)
        assert np.isclose(OF.ket(net, h, False)[i - 1], np.exp(net.logpsi(h.state(i)[:, None]))[0])
    hf = HomogeneousFock(N, 2)
    ndm = OM.random_machine("ndm", N, 2, seed=2, std=0.3)
    rho = OF.densitymatrix(ndm, hf)
    assert abs(np.trace(rho) - 1) < 1e-14 and np.abs(rho - rho.conj().T).max() < 1e-14
    assert np.linalg.eigvalsh(rho).min() > -1e-14         # NDM is positive by construction
    raw = OF.densitymatrix(ndm, hf, False)
    for i, j in [(1, 1), (2, 5), (8, 3)]:
        assert np.isclose(raw[i - 1, j - 1], np.exp(ndm.logpsi(hf.state(i)[:, None], hf.state(j)[:, None]))[0])


def test_exact_table_and_draws():
    N = 4
    h = HomogeneousSpin(N)
    net = OM.random_machine("rbm", N, 2, act=OM.LOGCOSH, complex_weights=True, seed=3, std=0.3)
    cdf = OF.exact_cdf(net, h)
    p = np.abs(OF.ket(net, h)) ** 2
    assert np.allclose(np.diff(np.concatenate([[0], cdf])), p, atol=1e-14) and abs(cdf[-1] - 1) < 1e-15
    # searchsortedfirst: first index whose cumulative value is >= r
    assert OF.exact_draw(cdf, np.array([0.0]))[0] == 1
    assert OF.exact_draw(cdf, np.array([cdf[4]]))[0] == 5
    assert OF.exact_draw(cdf, np.array([np.nextafter(cdf[4], 1)]))[0] == 6
    u = np.random.default_rng(0).random(200000)
    cnt = np.bincount(OF.exact_draw(cdf, u) - 1, minlength=2 ** N)
    assert np.abs(cnt / u.size - p).max() < 5e-3
    hf = HomogeneousFock(2, 2)
    ndm = OM.random_machine("ndm", 2, 2, seed=4, std=0.3)
    c2 = OF.exact_cdf(ndm, hf)
    p2 = np.abs(OF.densitymatrix(ndm, hf, False).reshape(-1, order="F")) ** 2      # super index = row + D (col - 1)
    assert np.allclose(np.diff(np.concatenate([[0], c2])), p2 / p2.sum(), atol=1e-14)


def test_rule_proposals():
    N = 4
    h, H = tfim_1d(N)
    coup = OR.couplings(H)
    assert sorted(tuple(sorted(c)) for c in coup) == [(1, 2), (1, 4), (2, 3), (3, 4)]
    s = np.array([[1.0, -1.0, -1.0, 1.0]]).T
    d = np.zeros((1, 4), dtype=np.int64)
    d[0, 0] = 1 + coup.index(next(c for c in coup if set(c) == {1, 2}))
    new, bias = OR.propose("exchange", h, s, d, coup=coup)
    assert np.array_equal(new[:, 0], [-1.0, 1.0, -1.0, 1.0]) and bias[0] == 0
    # operator rule on the TFIM: 3N connections everywhere, no bias; r selects the connection in term order
    conns = OO.connections_ket(H, s[:, 0])
    assert len(conns) == 3 * N
    for k in (0, 1, 5):
        d[0, 0] = ((k << 32) + len(conns) - 1) // len(conns) if k else 0
        new, bias = OR.propose("operator", h, s, d, operator=H)
        assert np.array_equal(new[:, 0], OO.apply_changes(s[:, 0], conns[(int(d[0, 0]) * len(conns)) >> 32][1])) and bias[0] == 0
    # Lindbladian: the number of connections depends on the configuration -> non-zero bias, antisymmetric under reversal
    hf, Hf, _, liouv = lindblad_ising_1d(3)
    row, col = np.array([[0.0, 1.0, 1.0]]).T, np.array([[1.0, 0.0, 1.0]]).T
    nf = len(OO.connections_super(liouv, row[:, 0], col[:, 0]))
    seen = False
    for k in range(nf):
        d[0, 0] = ((k << 32) + nf - 1) // nf
        (nr, nc), bias = OR.propose("operator", hf, (row, col), d, operator=liouv)
        nb = len(OO.connections_super(liouv, nr[:, 0], nc[:, 0]))
        assert np.isclose(bias[0], np.log(nf / nb))
        seen |= nb != nf
    assert seen
    # Nagy moves
    coup3 = [(1, 2), (2, 3), (3, 1)]
    for move, aux, aux2, exp_r, exp_c in [(1, 2, 0, [1, 0, 1], [1, 0, 1]),     # hop in sigma: site 1 and element 2 of couple 1
                                          (3, 1, 0, [0, 1, 1], [1, 0, 1]),     # hop in sigma': site 1 twice -> unchanged
                                          (5, 0, 0, [1, 1, 1], [1, 0, 1]), (6, 0, 0, [0, 1, 1], [0, 0, 1]),
                                          (7, 1, 5, [1, 1, 1], [0, 0, 1]),     # row empty + coin 1 -> excited; col occupied -> decays
                                          (7, 2, 1, [0, 1, 1], [0, 0, 1]),     # row empty, coin != 1 -> stays
                                          (8, 3, 0, [1, 1, 1], [1, 0, 0])]:
        d[0] = [move, 1, aux, aux2]
        (nr, nc), bias = OR.propose("nagy", hf, (row, col), d, coup=coup3)
        assert nr[:, 0].tolist() == exp_r and nc[:, 0].tolist() == exp_c and bias[0] == 0, move
