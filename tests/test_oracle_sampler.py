"""Pins the oracle's Metropolis replay the way test/Samplers/test_samplers.jl:26-85 pins the
reference: chi-square (p >= 0.01) of the sampled histogram vs the exact |psi|^2 / |rho|^2, N=4."""
import numpy as np
from scipy import stats as sst

from oracle import machines as M
from oracle import sampler as S
from oracle import stats as ST
from oracle.hilbert import HomogeneousFock, HomogeneousSpin
from oracle.models import random_states


def _run(net, hilb, doubled, n_store, passes, B=64, seed=99):
    rng = np.random.Generator(np.random.Philox(seed))
    nsite = hilb.n * (2 if doubled else 1)
    st = random_states(hilb, B, seed=1)
    if doubled:
        st = (st, random_states(hilb, B, seed=2))
    hist = {}
    acc_tot = 0
    for it in range(n_store + 20):
        sites = rng.integers(1, nsite + 1, size=(passes, B))
        u = rng.random((passes, B))
        st, acc, margin = S.samplenext_replay(net, hilb, st, sites, u)
        acc_tot += acc.sum()
        if it < 20:
            continue
        for c in range(B):
            key = (hilb.toint(st[0][:, c]), hilb.toint(st[1][:, c])) if doubled else hilb.toint(st[:, c])
            hist[key] = hist.get(key, 0) + 1
    return hist, acc_tot


def test_ket_chain_samples_psi_squared():
    hilb = HomogeneousSpin(4)
    net = M.random_machine("rbm", 4, 1, complex_weights=True, seed=123, std=0.2)
    hist, acc = _run(net, hilb, False, 400, 3)
    allS = hilb.all_states()
    p = np.abs(np.exp(net.logpsi(allS))) ** 2
    p /= p.sum()
    obs = np.array([hist.get(i + 1, 0) for i in range(16)])
    chi = sst.chisquare(obs, p * obs.sum())
    assert chi.pvalue >= 0.01
    assert acc > 0


def test_doubled_chain_samples_rho_squared():
    hilb = HomogeneousFock(2)
    net = M.random_machine("ndm", 2, 2, seed=123, std=0.3)
    hist, _ = _run(net, hilb, True, 400, 3)
    allS = hilb.all_states()
    keys = [(i + 1, j + 1) for i in range(4) for j in range(4)]
    sr = np.stack([allS[:, i] for i, j in [(a - 1, b - 1) for a, b in keys]], 1)
    sc = np.stack([allS[:, j] for i, j in [(a - 1, b - 1) for a, b in keys]], 1)
    p = np.abs(np.exp(net.logpsi(sr, sc))) ** 2
    p /= p.sum()
    obs = np.array([hist.get(k, 0) for k in keys])
    assert sst.chisquare(obs, p * obs.sum()).pvalue >= 0.01


def test_diagonal_chain_samples_the_diagonal_of_rho():
    """Density-matrix observables chain (BatchedObsDMSampler.jl): p(sigma) ~ rho(sigma, sigma), which is real positive
    for the NDM; sigma' follows sigma."""
    hilb = HomogeneousFock(3)
    net = M.random_machine("ndm", 3, 2, seed=5, std=0.4)
    rng = np.random.Generator(np.random.Philox(3))
    B, passes = 64, 3
    st = random_states(hilb, B, seed=1)
    hist = np.zeros(8)
    for it in range(170):
        sites = rng.integers(1, hilb.n + 1, size=(passes, B))
        u = rng.random((passes, B))
        st, acc, margin = S.samplenext_diagonal_replay(net, hilb, st, sites, u)
        if it >= 20 and it % 2 == 0:
            for c in range(B):
                hist[hilb.toint(st[:, c]) - 1] += 1
    allS = hilb.all_states()
    d = np.exp(net.logpsi(allS, allS))
    assert np.max(np.abs(d.imag)) <= 1e-12 * np.abs(d).max() and np.all(d.real > 0)
    p = d.real / d.real.sum()
    assert sst.chisquare(hist, p * hist.sum()).pvalue >= 0.01


def test_replay_decision_rule_and_site_mapping():
    hilb = HomogeneousFock(3)
    net = M.random_machine("rbmsplit", 3, 1, seed=4, std=0.5)
    r = np.array([[0.0], [1.0], [0.0]]); c = np.array([[1.0], [1.0], [0.0]])
    # site 5 -> col[2]
    new = S.propose(hilb, (r, c), [5])
    assert np.array_equal(new[0], r) and new[1][1, 0] == 0.0
    lp0 = M.log_prob(net.logpsi(r, c))[0]
    lp1 = M.log_prob(net.logpsi(*new))[0]
    ratio = np.exp(lp1 - lp0)
    for u, expect in ((ratio * 0.999, True), (min(ratio * 1.001, 0.9999999), ratio * 1.001 < ratio or ratio > 1)):
        st, acc, margin = S.samplenext_replay(net, hilb, (r, c), np.array([[5]]), np.array([[u]]))
        assert bool(acc[0, 0]) == bool(u - ratio < 0)
        assert np.array_equal(st[1], new[1] if acc[0, 0] else c)


def test_stat_analysis_matches_definitions():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((8, 50)) + 1j * rng.standard_normal((8, 50))
    m = ST.stat_analysis(v)
    assert np.isclose(m["mean"], v.mean())
    mu_ch = v.mean(1)
    assert np.isclose(m["error"], np.sqrt(np.var(mu_ch, ddof=1) / 8))
    assert np.isclose(m["variance"], np.mean(np.var(v, axis=1, ddof=1)))
    t = np.var(mu_ch, ddof=1) / np.var(v, ddof=1)
    assert np.isclose(m["tau"], max(0, 0.5 * (t * 50 - 1))) and np.isclose(m["R"], np.sqrt(49 / 50 + t))
