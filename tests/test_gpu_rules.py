"""GPU parity: the transition rules other than LocalRule (csrc/nq_sampler.cu, sampler_rule_kernel) against the oracle
restatement (oracle/rules.py) of MCMCRules/{ExchangeRule,Nagy,OperatorRule}.jl: accept/reject decisions and final states
bit-exact in replay mode (knife-edge rule of SURVEY Appendix D.3), production chains against exact distributions."""
import numpy as np
import pytest
from scipy import stats as sst

import helpers as H
from oracle import machines as OM
from oracle import rules as OR
from oracle.models import lindblad_ising_1d, tfim_1d

pytestmark = pytest.mark.gpu


def _draws(rule, rng, passes, B, N, ncoup):
    d = np.zeros((passes, B, 4), dtype=np.int64)
    if rule == "exchange":
        d[..., 0] = rng.integers(1, ncoup + 1, size=(passes, B))
    elif rule == "nagy":
        mv = rng.integers(1, 9, size=(passes, B))
        d[..., 0] = mv
        d[..., 1] = rng.integers(1, N + 1, size=(passes, B))
        d[..., 2] = np.where(mv <= 4, rng.integers(1, 3, size=(passes, B)),
                             np.where(mv == 7, rng.integers(1, 11, size=(passes, B)), rng.integers(1, N + 1, size=(passes, B))))
        d[..., 3] = rng.integers(1, 11, size=(passes, B))
    else:
        d[..., 0] = rng.integers(0, 1 << 32, size=(passes, B))
    return d


def _replay_rule_case(nq, ctx, rule, kind, hk, N, dtype, act, B=16, passes=5, nsteps=5, seed=7):
    om, pm, hilb = H.make_pair(nq, ctx, kind, hk, N, 2, dtype, act, std=0.3)
    oh = H.ohilb(hk, N)
    if kind == "rbm":
        _, oH = tfim_1d(N)
        _, pH = H.p_tfim_1d(nq, N)
        oop, pop = oH, pH
    else:
        _, oH, _, ol = lindblad_ising_1d(N, fock=(hk == "fock"))
        _, pH, _, pl = H.p_lindblad_ising_1d(nq, N, fock=(hk == "fock"))
        oop, pop = ol, pl
    coup = OR.couplings(oH)
    prule = {"exchange": lambda: nq.ExchangeRule(pH), "nagy": lambda: nq.NagyRule(pH), "operator": lambda: nq.OperatorRule(pop)}[rule]()
    if rule == "exchange":
        assert prule.distances == coup
    if rule == "nagy":
        assert prule.adjacency_list == coup
    smp = nq.MetropolisSampler(prule, 10, passes)
    cache = nq.MetropolisSamplerCache(smp, pm, B)
    passes = smp.passes
    rng = np.random.Generator(np.random.Philox(seed))
    st = H.rand_states(hk, N, B, 1)
    if om.doubled:
        st = (st, H.rand_states(hk, N, B, 2))
    cache.set_state(st)
    rdt = np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64
    knife = 1e-4 if rdt == np.float32 else 1e-9
    excluded = total = moved = 0
    for it in range(nsteps):
        d = _draws(rule, rng, passes, B, N, len(coup))
        u = rng.random((passes, B)).astype(rdt)
        new, acc_ref, margin = OR.samplenext_rule_replay(om, oh, st, rule, d, u.astype(np.float64), operator=oop, coup=coup, dtype=rdt)
        acc = cache.replay_rule(d, u)
        ratio = u.astype(np.float64) - margin
        edge = np.abs(margin) < knife * np.maximum(1.0, ratio)
        total += acc.size
        excluded += int(edge.sum())
        if edge.any():
            assert np.array_equal(acc[~edge], acc_ref[~edge])
            cache.set_state(new)
        else:
            assert np.array_equal(acc, acc_ref), "accept/reject decisions differ at step %d" % it
            got = cache.get_state()
            if om.doubled:
                assert np.array_equal(got[0], new[0]) and np.array_equal(got[1], new[1])
            else:
                assert np.array_equal(got, new)
        moved += int(np.any(np.asarray(new) != np.asarray(st)))
        st = new
    assert moved > 0
    return excluded, total


@pytest.mark.parametrize("rule,kind,hk,N,dtype,act", [
    ("exchange", "rbm", "spin", 8, np.complex128, OM.LOGCOSH),
    ("exchange", "rbm", "fock", 6, np.float64, OM.SOFTPLUS),
    ("operator", "rbm", "spin", 8, np.complex128, OM.LOGCOSH),
    ("operator", "ndm", "fock", 5, np.float64, OM.SOFTPLUS),
    ("operator", "rbmsplit", "fock", 4, np.complex128, OM.SOFTPLUS),
    ("nagy", "ndm", "fock", 6, np.float64, OM.SOFTPLUS),
    ("nagy", "ndm", "spin", 5, np.float64, OM.LOGCOSH),
    ("nagy", "rbmsplit", "fock", 5, np.complex128, OM.SOFTPLUS),
])
def test_rule_replay_bit_exact_fp64(nq, ctx, rule, kind, hk, N, dtype, act):
    excluded, total = _replay_rule_case(nq, ctx, rule, kind, hk, N, dtype, act)
    assert excluded == 0, "%d of %d decisions sat on the knife edge" % (excluded, total)


@pytest.mark.parametrize("rule,kind,hk,N,dtype,act", [("exchange", "rbm", "spin", 8, np.complex64, OM.LOGCOSH),
                                                       ("nagy", "ndm", "fock", 6, np.float32, OM.SOFTPLUS)])
def test_rule_replay_fp32(nq, ctx, rule, kind, hk, N, dtype, act):
    excluded, total = _replay_rule_case(nq, ctx, rule, kind, hk, N, dtype, act)
    assert excluded <= total // 50


def test_rule_argument_checks(nq, ctx):
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", 6, 1, np.float64, OM.LOGCOSH)
    om2, pm2, hilb2 = H.make_pair(nq, ctx, "ndm", "fock", 4, 1, np.float64, OM.SOFTPLUS)
    _, pH = H.p_tfim_1d(nq, 6)
    _, pH4, _, pl4 = H.p_lindblad_ising_1d(nq, 4)
    with pytest.raises(nq.NQError):          # NagyRule needs doubled states
        nq.MetropolisSamplerCache(nq.MetropolisSampler(nq.NagyRule(pH), 4, 3), pm, 4)
    with pytest.raises(nq.NQError):          # ExchangeRule is ket-only in the reference ("not implemented")
        nq.MetropolisSamplerCache(nq.MetropolisSampler(nq.ExchangeRule(pH4), 4, 3), pm2, 4)
    with pytest.raises(nq.NQError):          # operator space must match the machine
        nq.MetropolisSamplerCache(nq.MetropolisSampler(nq.OperatorRule(pH), 4, 3), pm2, 4)
    with pytest.raises(nq.NQError):          # the reference indexes the couple list by site: N-1 couples (open chain) < N sites
        nq.MetropolisSamplerCache(nq.MetropolisSampler(nq.NagyRule([(1, 2), (2, 3), (3, 4)]), 4, 3), pm2, 4)


def test_production_operator_rule_samples_psi_squared(nq, ctx):
    """OperatorRule with the TFIM Hamiltonian: every state has N+1 connections (N flips + the diagonal term), so the
    bias vanishes and the chain samples |psi|^2; chi-square like test/Samplers/test_samplers.jl:26-85."""
    N = 4
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 1, np.complex128, OM.SOFTPLUS, seed=123, std=0.2)
    _, pH = H.p_tfim_1d(nq, N)
    smp = nq.MetropolisSampler(nq.OperatorRule(pH), 400, 5, burn=50, seed=11)
    cache = nq.MetropolisSamplerCache(smp, pm, 64)
    cache.randomize()
    S = cache.sample()
    idx = ((S.reshape(N, -1, order="F") + 1) / 2).astype(int)
    codes = (idx * (1 << np.arange(N))[:, None]).sum(0)
    allS = H.ohilb("spin", N).all_states()
    p = np.abs(np.exp(om.logpsi(allS))) ** 2
    p /= p.sum()
    thin = codes.reshape(400, 64)[::4].ravel()
    obs = np.bincount(thin, minlength=16)
    assert sst.chisquare(obs, p * obs.sum()).pvalue >= 0.01
    done, accepted = cache.counters()
    assert done == 64 * 450 * 5 and 0 < accepted < done


def test_production_exchange_rule_conserves_magnetisation_and_samples_the_sector(nq, ctx):
    N = 6
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 1, np.complex128, OM.LOGCOSH, seed=5, std=0.3)
    _, pH = H.p_tfim_1d(nq, N)
    B = 64
    smp = nq.MetropolisSampler(nq.ExchangeRule(pH), 300, 7, burn=50, seed=3)
    cache = nq.MetropolisSamplerCache(smp, pm, B)
    start = np.tile(np.array([1, 1, 1, -1, -1, -1], dtype=np.float64)[:, None], (1, B))
    cache.set_state(np.asfortranarray(start))
    S = cache.sample()
    assert np.all(S.sum(0) == 0)                                     # exchanges never leave the sector
    idx = ((S.reshape(N, -1, order="F") + 1) / 2).astype(int)
    codes = (idx * (1 << np.arange(N))[:, None]).sum(0)
    allS = H.ohilb("spin", N).all_states()
    p = np.abs(np.exp(om.logpsi(allS))) ** 2
    p[allS.sum(0) != 0] = 0
    p /= p.sum()
    thin = codes.reshape(300, B)[::5].ravel()
    obs = np.bincount(thin, minlength=1 << N)
    keep = p > 0
    assert obs[~keep].sum() == 0
    assert sst.chisquare(obs[keep], p[keep] * obs.sum()).pvalue >= 0.01


def test_production_nagy_rule_sharding_invariance(nq, ctx):
    """The Nagy moves are not symmetric (the dissipator move), so the reference's chain has no closed-form law to
    test against; production mode is checked for reproducibility and independence of the sharding of the chains."""
    N = 4
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, OM.SOFTPLUS, seed=9, std=0.3)
    coup = [(1, 2), (2, 3), (3, 4), (4, 1)]
    smp = nq.MetropolisSampler(nq.NagyRule(coup), 20, 5, burn=10, seed=21)
    full = nq.MetropolisSamplerCache(smp, pm, 32)
    full.randomize()
    sr, sc = full.sample()
    assert set(np.unique(sr)) <= {0.0, 1.0} and len(np.unique(sr.reshape(N, -1, order="F"), axis=1).T) > 4
    for off in (0, 16):
        part = nq.MetropolisSamplerCache(smp, pm, 16, chain_offset=off)
        part.randomize()
        pr, pc = part.sample()
        assert np.array_equal(pr, sr[:, off:off + 16]) and np.array_equal(pc, sc[:, off:off + 16])
    done, accepted = full.counters()
    assert done == 32 * 30 * 5 and 0 < accepted <= done
