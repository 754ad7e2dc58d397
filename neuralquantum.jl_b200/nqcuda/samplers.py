"""Metropolis-Hastings sampler (LocalRule) over libnqcuda.

ref: src/Samplers/Metropolis.jl:4-40 (MetropolisSampler), :54-93 (cache), :101-167 (init_sampler!,
     samplenext!), src/Samplers/MCMCRules/LocalRule.jl.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class LocalRule:
    """Transition rule: flip one random site (LocalRule.jl:10-28)."""


class MetropolisSampler:
    """MetropolisSampler(rule, chain_length, passes; burn=0, seed).  Even `passes` is bumped to odd."""

    def __init__(self, rule, chain_length, passes, burn=0, seed=None):
        if not isinstance(rule, LocalRule):
            raise NotImplementedError("the device sampler implements LocalRule (north_star scope)")
        assert passes > 0 and chain_length > 0
        self.rule, self.chain_length, self.burn_length = rule, int(chain_length), int(burn)
        self.passes = passes + 1 if passes % 2 == 0 else passes
        self.seed = int(np.random.SeedSequence().entropy % (1 << 63)) if seed is None else int(seed)


class MetropolisSamplerCache:
    """Device state of B chains (nq_sampler_t).  chain_offset = global id of the first chain, so the
    Philox streams do not depend on how chains are sharded over GPUs (worker_local_seed replacement)."""

    def __init__(self, sampler, net, batch_sz, chain_offset=0, num_workers=1):
        self.s, self.net, self.B = sampler, net, int(batch_sz)
        self.loc_chain_length = -(-sampler.chain_length // num_workers)      # Metropolis.jl:75
        h = C.c_void_p()
        L.check(L.lib.nq_sampler_create(net.h, self.B, sampler.passes, sampler.seed, int(chain_offset), C.byref(h)),
                net.ctx.h)
        self.h = h

    def __del__(self):
        try:
            if self.h and self.net.ctx.h:
                L.lib.nq_sampler_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def set_state(self, sigma):
        sr, sc, _ = self.net._states(sigma)
        L.check(L.lib.nq_sampler_set_state(self.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(sr.dtype)), self.net.ctx.h)

    def get_state(self, dtype=np.float64):
        sr = np.zeros((self.net.N, self.B), dtype=dtype, order="F")
        sc = np.zeros_like(sr) if self.net.doubled else None
        L.check(L.lib.nq_sampler_get_state(self.h, L.ptr(sr), L.ptr(sc), L.nq_dtype(dtype)), self.net.ctx.h)
        return (sr, sc) if self.net.doubled else sr

    def set_mode(self, diagonal):
        """diagonal=True: chain over rho(sigma, sigma) (density-matrix observables); False: joint (sigma, sigma')."""
        L.check(L.lib.nq_sampler_set_mode(self.h, 1 if diagonal else 0), self.net.ctx.h)

    def randomize(self):
        L.check(L.lib.nq_sampler_randomize(self.h), self.net.ctx.h)

    def replay(self, sites, uniforms):
        """One samplenext! with supplied randomness.  sites [passes, B] int (1-based), uniforms [passes, B]."""
        sites = np.ascontiguousarray(sites, dtype=np.int32)
        uniforms = np.ascontiguousarray(uniforms, dtype=self.net.rdtype)
        assert sites.shape == (self.s.passes, self.B) == uniforms.shape
        acc = np.zeros((self.s.passes, self.B), dtype=np.uint8)
        L.check(L.lib.nq_sampler_replay(self.h, L.ptr(sites), L.ptr(uniforms), L.ptr(acc)), self.net.ctx.h)
        return acc.astype(bool)

    def sample(self, burn=None, L_store=None, dtype=np.float64, packed_out=None):
        """init_sampler!-style run: `burn` discarded + L stored samples per chain.
        Returns float arrays [N, B, L] (host) or fills device packed buffers (prow_ptr, pcol_ptr)."""
        burn = self.s.burn_length if burn is None else burn
        Ls = self.loc_chain_length if L_store is None else L_store
        if packed_out is not None:
            pr, pc = packed_out
            L.check(L.lib.nq_sampler_sample(self.h, burn, Ls, pr, pc, None, None, L.NQ_F64), self.net.ctx.h)
            return None
        sr = np.zeros((self.net.N, self.B, Ls), dtype=dtype, order="F")
        sc = np.zeros_like(sr) if self.net.doubled else None
        L.check(L.lib.nq_sampler_sample(self.h, burn, Ls, None, None, L.ptr(sr), L.ptr(sc), L.nq_dtype(dtype)),
                self.net.ctx.h)
        return (sr, sc) if self.net.doubled else sr

    def counters(self):
        a, b = C.c_int64(), C.c_int64()
        L.check(L.lib.nq_sampler_counters(self.h, C.byref(a), C.byref(b)), self.net.ctx.h)
        return a.value, b.value


class ExactSampler:
    """ExactSampler(n_samples; seed) (Samplers/Exact.jl:3-21): builds the full probability table of the state and
    samples it exactly.  Indexable spaces only -- 2^N (ket) or 4^N (density matrix) table entries."""

    def __init__(self, n_samples, seed=None):
        self.samples_length = int(n_samples)
        self.chain_length = self.samples_length
        self.burn_length = 0
        self.seed = int(np.random.SeedSequence().entropy % (1 << 63)) if seed is None else int(seed)


class ExactSamplerCache:
    """Exact.jl:135-181 on the device: log-probabilities of ALL basis states through the machine kernel (packed
    index = basis number: site 1 is the least significant digit, super-index = col * D + row), a cumulative table,
    and one independent inverse-CDF draw per (chain, slot).  The table and the draws use torch ops: this is the
    reference's validation sampler, not part of the hot path.  Julia's MersenneTwister stream is not reproduced."""
    MAX_TABLE = 1 << 24

    def __init__(self, sampler, net, batch_sz, chain_offset=0, num_workers=1):
        import torch
        if net.N > 62:
            raise ValueError("ExactSampler needs an indexable space")
        self.s, self.net, self.B = sampler, net, int(batch_sz)
        self.loc_chain_length = -(-sampler.samples_length // num_workers)            # Exact.jl:41
        self.D = 1 << net.N
        self.size = self.D * self.D if net.doubled else self.D
        if self.size > self.MAX_TABLE:
            raise ValueError("ExactSampler: %d table entries (Closed N < 24, Open N < 12)" % self.size)
        self.gen = torch.Generator(device=torch.device("cuda", net.ctx.device))
        self.gen.manual_seed((sampler.seed + 0x9E3779B97F4A7C15 * (chain_offset + 1)) % (1 << 63))
        self.cdf = None

    def init_sampler(self):
        """init_sampler!: the probability table of the CURRENT parameters."""
        import torch
        net, dev = self.net, torch.device("cuda", self.net.ctx.device)
        idx = torch.arange(self.size, dtype=torch.int64, device=dev)
        row = (idx % self.D).contiguous() if net.doubled else idx
        col = (idx // self.D).contiguous() if net.doubled else None
        from .iterative import _tdtype
        out = torch.zeros(self.size, dtype=_tdtype(net.out_dtype), device=dev)
        L.check(L.lib.nq_logpsi_packed(net.h, row.data_ptr(), col.data_ptr() if col is not None else None, self.size,
                                       out.data_ptr()), net.ctx.h)
        lp = 2.0 * out.real.to(torch.float64) if out.is_complex() else 2.0 * out.to(torch.float64)   # log_prob_psi
        p = torch.exp(lp - lp.max())
        self.cdf = torch.cumsum(p, 0)
        self.cdf /= self.cdf[-1].clone()
        return self.cdf

    def sample_into(self, L_store, prow, pcol):
        """samplenext! for every (slot, chain): basis number by inverse CDF, written as packed words [L, B, 1]."""
        import torch
        self.init_sampler()
        u = torch.rand((L_store, self.B), generator=self.gen, device=self.cdf.device, dtype=torch.float64)
        hi = torch.searchsorted(self.cdf, u.reshape(-1)).clamp_(max=self.size - 1).reshape(L_store, self.B)
        if self.net.doubled:
            prow[:, :, 0] = hi % self.D
            pcol[:, :, 0] = hi // self.D
        else:
            prow[:, :, 0] = hi
        return hi
