"""GPU parity: Metropolis kernel (K4).  Replay mode: accept/reject decisions bit-exact vs the oracle
under a shared (site, uniform) stream (knife-edge rule of SURVEY Appendix D.3); production mode:
chi-square of the sampled histogram vs the exact distribution (test/Samplers/test_samplers.jl:26-85)."""
import numpy as np
import pytest
from scipy import stats as sst

import helpers as H
from oracle import machines as OM
from oracle import sampler as OS

pytestmark = pytest.mark.gpu


def _replay_case(nq, ctx, kind, hk, N, alpha, dtype, act, B, passes, nsteps, seed=99):
    om, pm, hilb = H.make_pair(nq, ctx, kind, hk, N, alpha, dtype, act, std=0.3)
    oh = H.ohilb(hk, N)
    smp = nq.MetropolisSampler(nq.LocalRule(), 10, passes)
    cache = nq.MetropolisSamplerCache(smp, pm, B)
    passes = smp.passes
    rng = np.random.Generator(np.random.Philox(seed))
    st = H.rand_states(hk, N, B, 1)
    if om.doubled:
        st = (st, H.rand_states(hk, N, B, 2))
    cache.set_state(st)
    nsite = N * (2 if om.doubled else 1)
    rdt = np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64
    knife = 1e-4 if rdt == np.float32 else 1e-9
    excluded = total = 0
    for it in range(nsteps):
        sites = rng.integers(1, nsite + 1, size=(passes, B))
        u = rng.random((passes, B)).astype(rdt)
        # oracle pass by pass so that the chain can be re-synchronised after an excluded decision
        new, acc_ref, margin = OS.samplenext_replay(om, oh, st, sites, u.astype(np.float64), dtype=rdt)
        acc = cache.replay(sites, u)
        ratio = u.astype(np.float64) - margin
        edge = np.abs(margin) < knife * np.maximum(1.0, ratio)
        total += acc.size
        excluded += int(edge.sum())
        if edge.any():
            assert np.array_equal(acc[~edge], acc_ref[~edge])
            cache.set_state(new)               # re-synchronise to the oracle
        else:
            assert np.array_equal(acc, acc_ref), "accept/reject decisions differ at step %d" % it
            got = cache.get_state()
            if om.doubled:
                assert np.array_equal(got[0], new[0]) and np.array_equal(got[1], new[1])
            else:
                assert np.array_equal(got, new)
        st = new
    done, accepted = cache.counters()
    assert done == nsteps * passes * B and 0 < accepted <= done
    return excluded, total


@pytest.mark.parametrize("kind,hk,N,alpha,dtype,act", [
    ("rbm", "spin", 10, 2, np.complex128, OM.LOGCOSH),
    ("rbm", "spin", 10, 2, np.float64, OM.LOGCOSH),
    ("rbm", "fock", 8, 5, np.complex128, OM.SOFTPLUS),
    ("rbmsplit", "fock", 6, 2, np.complex128, OM.SOFTPLUS),
    ("ndm", "fock", 8, 2, np.float64, OM.SOFTPLUS),
    ("ndm", "spin", 6, 1, np.float64, OM.LOGCOSH),
    ("ndm", "fock", 16, 2, np.float64, OM.SOFTPLUS),
])
def test_replay_bit_exact_fp64(nq, ctx, kind, hk, N, alpha, dtype, act):
    excluded, total = _replay_case(nq, ctx, kind, hk, N, alpha, dtype, act, B=24, passes=5, nsteps=6)
    assert excluded == 0, "%d of %d decisions sat on the knife edge" % (excluded, total)


@pytest.mark.parametrize("kind,hk,N,alpha,dtype,act", [
    ("rbm", "spin", 10, 2, np.complex64, OM.LOGCOSH),
    ("ndm", "fock", 8, 2, np.float32, OM.SOFTPLUS),
])
def test_replay_fp32(nq, ctx, kind, hk, N, alpha, dtype, act):
    excluded, total = _replay_case(nq, ctx, kind, hk, N, alpha, dtype, act, B=24, passes=5, nsteps=6)
    assert excluded <= total // 100


def test_replay_multiword(nq, ctx):
    excluded, _ = _replay_case(nq, ctx, "rbm", "spin", 70, 1, np.float64, OM.LOGCOSH, B=5, passes=3, nsteps=3)
    assert excluded == 0


def test_production_chain_samples_psi_squared(nq, ctx):
    N = 4
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 1, np.complex128, OM.SOFTPLUS, seed=123, std=0.2)
    smp = nq.MetropolisSampler(nq.LocalRule(), 400, 3, burn=50, seed=123)
    cache = nq.MetropolisSamplerCache(smp, pm, 64)
    cache.randomize()
    S = cache.sample()                                   # [N, B, L]
    assert S.shape == (N, 64, 400) and set(np.unique(S)) <= {-1.0, 1.0}
    idx = ((S.reshape(N, -1, order="F") + 1) / 2).astype(int)
    codes = (idx * (1 << np.arange(N))[:, None]).sum(0)
    obs = np.bincount(codes, minlength=16)
    allS = H.ohilb("spin", N).all_states()
    p = np.abs(np.exp(om.logpsi(allS))) ** 2
    p /= p.sum()
    # thin the chain so that samples are close to independent
    thin = codes.reshape(400, 64)[::4].ravel()
    obs = np.bincount(thin, minlength=16)
    assert sst.chisquare(obs, p * obs.sum()).pvalue >= 0.01
    done, accepted = cache.counters()
    assert done == 64 * 450 * 3 and 0 < accepted < done


def test_production_density_matrix_and_sharding_invariance(nq, ctx):
    N = 2
    om, pm, hilb = H.make_pair(nq, ctx, "ndm", "fock", N, 2, np.float64, OM.SOFTPLUS, seed=123, std=0.3)
    smp = nq.MetropolisSampler(nq.LocalRule(), 300, 3, burn=50, seed=7)
    full = nq.MetropolisSamplerCache(smp, pm, 64)
    full.randomize()
    sr, sc = full.sample()
    code = lambda a: (a.reshape(N, -1, order="F").astype(int) * (1 << np.arange(N))[:, None]).sum(0)
    joint = code(sr) + 4 * code(sc)
    allS = H.ohilb("fock", N).all_states()
    R = np.stack([allS[:, k % 4] for k in range(16)], 1)
    Cc = np.stack([allS[:, k // 4] for k in range(16)], 1)
    p = np.abs(np.exp(om.logpsi(R, Cc))) ** 2
    p /= p.sum()
    thin = joint.reshape(300, 64)[::4].ravel()
    obs = np.bincount(thin, minlength=16)
    assert sst.chisquare(obs, p * obs.sum()).pvalue >= 0.01
    # chains are keyed by their GLOBAL id: two shards of 32 chains reproduce the 64-chain run exactly
    for off in (0, 32):
        part = nq.MetropolisSamplerCache(smp, pm, 32, chain_offset=off)
        part.randomize()
        pr, pc = part.sample()
        assert np.array_equal(pr, sr[:, off:off + 32]) and np.array_equal(pc, sc[:, off:off + 32])


def test_exact_sampler_ket_and_density_matrix(nq, ctx):
    """ExactSampler (Samplers/Exact.jl): inverse-CDF draws from the full probability table; chi-square against the
    oracle's |psi|^2 / |rho|^2 and a sampled energy within a few standard errors of exact diagonalisation."""
    import torch
    from oracle import operators as OO
    from oracle.models import tfim_1d
    N = 5
    om, pm, hilb = H.make_pair(nq, ctx, "rbm", "spin", N, 2, np.complex128, OM.LOGCOSH, seed=3, std=0.3)
    oh, oH = tfim_1d(N)
    ph, pH = H.p_tfim_1d(nq, N)
    bs = nq.BatchedSampler(pm, nq.ExactSampler(200, seed=5), pH, nq.SR(np.float32, eps=0.1), batch_sz=64)
    stat, _ = bs.sample_()
    codes = bs.prow.cpu().numpy().ravel()
    allS = H.ohilb("spin", N).all_states()
    code_of = (((allS + 1) / 2).astype(int) * (1 << np.arange(N))[:, None]).sum(0)
    psi = np.exp(om.logpsi(allS))
    p = np.abs(psi) ** 2
    p /= p.sum()
    pk = np.zeros(1 << N)
    pk[code_of] = p
    cnt = np.bincount(codes, minlength=1 << N)
    assert cnt.sum() == 200 * 64
    exp = pk * cnt.sum()
    big = exp >= 5                       # pool the bins whose expected count is too small for the chi-square statistic
    obs_p = np.append(cnt[big], cnt[~big].sum())
    exp_p = np.append(exp[big], exp[~big].sum())
    assert sst.chisquare(obs_p, exp_p).pvalue >= 0.01
    Hm = OO.to_matrix(oH)
    exact = (psi.conj() @ (Hm @ psi) / (psi.conj() @ psi)).real
    assert abs(stat.mean.real - exact) <= 6 * max(stat.error, 1e-3)
    # density matrix: joint (row, col) index
    N2 = 2
    om2, pm2, hilb2 = H.make_pair(nq, ctx, "ndm", "fock", N2, 2, np.float64, OM.SOFTPLUS, seed=123, std=0.3)
    _, _, _, pl = H.p_lindblad_ising_1d(nq, N2)
    bs2 = nq.BatchedSampler(pm2, nq.ExactSampler(300, seed=7), pl, nq.SR(np.float32, eps=0.001), batch_sz=32)
    bs2.sample_states()
    joint = bs2.prow.cpu().numpy().ravel() + 4 * bs2.pcol.cpu().numpy().ravel()
    allS2 = H.ohilb("fock", N2).all_states()
    R = np.stack([allS2[:, k % 4] for k in range(16)], 1)
    Cc = np.stack([allS2[:, k // 4] for k in range(16)], 1)
    p2 = np.abs(np.exp(om2.logpsi(R, Cc))) ** 2
    p2 /= p2.sum()
    cnt2 = np.bincount(joint, minlength=16)
    assert sst.chisquare(cnt2, p2 * cnt2.sum()).pvalue >= 0.01
