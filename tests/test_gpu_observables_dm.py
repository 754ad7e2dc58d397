"""GPU parity: density-matrix observables: diagonal Markov chain (nq_sampler_set_mode) and the
row-acting local estimator, vs the oracle.  Bit-exact accept/reject in replay mode; estimator to 1e-11; <O> exact
against the dense density matrix; chi-square of the production chain against rho(sigma, sigma)."""
import numpy as np
import pytest
import scipy.stats as sst

import helpers as H
from oracle import estimators as OE
from oracle import machines as OM
from oracle import operators as OO
from oracle import sampler as OS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hk,N,alpha,act", [("fock", 8, 2, OM.SOFTPLUS), ("spin", 6, 1, OM.LOGCOSH), ("fock", 16, 2, OM.SOFTPLUS)])
def test_diagonal_chain_replay_bit_exact(nq, ctx, hk, N, alpha, act):
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", hk, N, alpha, np.float64, act, std=0.3)
    oh = H.ohilb(hk, N)
    B, nsteps = 24, 6
    smp = nq.MetropolisSampler(nq.LocalRule(), 10, 5)
    cache = nq.MetropolisSamplerCache(smp, pm, B)
    cache.set_mode(True)
    passes = smp.passes
    rng = np.random.Generator(np.random.Philox(5))
    st = H.rand_states(hk, N, B, 1)
    cache.set_state((st, st))
    for it in range(nsteps):
        sites = rng.integers(1, N + 1, size=(passes, B))
        u = rng.random((passes, B))
        new, acc_ref, margin = OS.samplenext_diagonal_replay(om, oh, st, sites, u)
        acc = cache.replay(sites, u)
        edge = np.abs(margin) < 1e-9 * np.maximum(1.0, u - margin)
        assert not edge.any()
        assert np.array_equal(acc, acc_ref), "accept/reject decisions differ at step %d" % it
        got = cache.get_state()
        assert np.array_equal(got[0], new) and np.array_equal(got[1], new)       # sigma' follows sigma
        st = new
    done, accepted = cache.counters()
    assert done == nsteps * passes * B and 0 < accepted <= done


def _observables(nq, ph, oh, N):
    p_mx, o_mx = nq.LocalOperator(ph), None
    for i in range(1, N + 1):
        p_mx = p_mx + (1.0 / N) * nq.sigmax(ph, i)
        o_mx = OO.add(o_mx, OO.scale(1.0 / N, OO.sigmax(oh, i)))
    p_c = nq.LocalOperator(ph) + nq.sigmay(ph, 1) * nq.sigmaz(ph, 3) + 0.5 * nq.sigmaz(ph, 2)
    o_c = OO.add(OO.mul(OO.sigmay(oh, 1), OO.sigmaz(oh, 3)), OO.scale(0.5, OO.sigmaz(oh, 2)))
    return {"mx": (p_mx, o_mx), "y1z3+z2/2": (p_c, o_c)}


def test_dm_local_estimator_and_exact_expectation(nq, ctx):
    N = 4
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, OM.SOFTPLUS, seed=11, std=0.3)
    oh = H.ohilb("fock", N)
    allS = oh.all_states()                                   # [N, 16]
    D = allS.shape[1]
    obs = nq.BatchedObsDMSampler(pm, nq.MetropolisSampler(nq.LocalRule(), 1, N, burn=0, seed=1), batch_sz=D, chain_length=1)
    obs.set_samples(allS.reshape(N, D, 1))
    # dense rho(row, col) on the full space
    R = np.repeat(allS, D, axis=1)
    Cc = np.tile(allS, (1, D))
    rho = np.exp(om.logpsi(R, Cc)).reshape(D, D)             # rho[row, col]
    p = np.real(np.diag(rho))
    assert np.all(p > 0) and np.max(np.abs(np.imag(np.diag(rho)))) <= 1e-12 * p.max()
    for name, (pop, oop) in _observables(nq, hilb, oh, N).items():
        obs.add_observable_(name, pop)
        loc = obs.local_values(name).cpu().numpy()
        ref = OE.local_scalar_super(om, OO.KLocalLiouvillian(oh, OO._tensor_left(oop), None, None), allS, allS)
        H.assert_close(loc, ref, 1e-11, "O_loc " + name)
        dense = OO.to_matrix(oop)
        # basis index of all_states column k must match to_matrix ordering: <O> = Tr(O rho) / Tr(rho)
        exact = np.trace(dense @ rho) / np.trace(rho)
        est = np.sum(p * loc) / p.sum()
        assert abs(est - exact) <= 1e-10 * max(1.0, abs(exact)), name


def test_dm_production_chain_samples_the_diagonal(nq, ctx):
    N = 3
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, OM.SOFTPLUS, seed=5, std=0.4)
    oh = H.ohilb("fock", N)
    smp = nq.MetropolisSampler(nq.LocalRule(), 300, 3, burn=50, seed=9)
    obs = nq.BatchedObsDMSampler(pm, smp, batch_sz=64)
    obs.sample_states()
    import torch
    assert torch.equal(obs.prow, obs.pcol)
    codes = obs.prow.cpu().numpy().reshape(300, 64)[::4].ravel()
    allS = oh.all_states()
    p = np.real(np.exp(om.logpsi(allS, allS)))
    p /= p.sum()
    # packed word bit j = site j+1 occupied; all_states column k: check the same code convention
    code_of = (allS.astype(int) * (1 << np.arange(N))[:, None]).sum(0)
    pk = np.zeros(8)
    pk[code_of] = p
    cnt = np.bincount(codes, minlength=8)
    assert sst.chisquare(cnt, pk * cnt.sum()).pvalue >= 0.01
    # and the sampled expectation agrees with the exact one within a few standard errors
    ph = hilb
    pop, oop = _observables(nq, ph, oh, N)["mx"]
    obs.add_observable_("mx", pop)
    res = obs.compute_observables(sample=False)["mx"]
    R = np.repeat(allS, 8, axis=1)
    Cc = np.tile(allS, (1, 8))
    rho = np.exp(om.logpsi(R, Cc)).reshape(8, 8)
    exact = np.trace(OO.to_matrix(oop) @ rho) / np.trace(rho)
    assert abs(res.mean - exact) <= 6 * max(res.error, 1e-3)
