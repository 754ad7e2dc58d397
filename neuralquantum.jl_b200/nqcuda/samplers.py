"""Metropolis-Hastings sampler (LocalRule) over libnqcuda.

ref: src/Samplers/Metropolis.jl:4-40 (MetropolisSampler), :54-93 (cache), :101-167 (init_sampler!,
     samplenext!), src/Samplers/MCMCRules/LocalRule.jl.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class LocalRule:
    """Transition rule: flip one random site (LocalRule.jl:10-28)."""
    code = L.NQ_RULE_LOCAL


def _couplings(op):
    """The site pairs of the 2-site terms of an operator, in term order (ExchangeRule.jl:17-34); 3-site couplings are
    ignored with a warning like the reference."""
    import warnings
    out = []
    for t in op.terms:
        if len(t.sites) == 1:
            continue
        if len(t.sites) > 2:
            warnings.warn("Can't exchange between 3-site couplings. This coupling is ignored")
            continue
        out.append((int(t.sites[0]), int(t.sites[1])))
    return out


class ExchangeRule:
    """ExchangeRule(H): at every step a random couple of sites coupled by a 2-body term of H is switched
    (ExchangeRule.jl:3-34).  `distances` = list of 1-based site pairs.  Ket states only, like the reference."""
    code = L.NQ_RULE_EXCHANGE

    def __init__(self, graph):
        self.distances = list(graph) if isinstance(graph, (list, tuple)) else _couplings(graph)
        if not self.distances:
            raise ValueError("ExchangeRule: the operator has no 2-site couplings")


class NagyRule:
    """NagyRule(H): the 8 moves of Nagy.jl:40-115 on doubled states (hopping in sigma / sigma', single flips, the
    dissipator move, the jumper).  `adjacency_list` is the list of 2-site couplings; the reference indexes it by SITE
    number (Nagy.jl:51), so it needs at least nsites entries."""
    code = L.NQ_RULE_NAGY

    def __init__(self, graph):
        self.adjacency_list = list(graph) if isinstance(graph, (list, tuple)) else _couplings(graph)


class OperatorRule:
    """OperatorRule(O): a move is drawn uniformly from the connections of O; log_prob_bias = log(n_forward / n_back)
    (OperatorRule.jl:3-59)."""
    code = L.NQ_RULE_OPERATOR

    def __init__(self, operator):
        self.operator = operator


class MetropolisSampler:
    """MetropolisSampler(rule, chain_length, passes; burn=0, seed).  Even `passes` is bumped to odd."""

    def __init__(self, rule, chain_length, passes, burn=0, seed=None):
        if not isinstance(rule, (LocalRule, ExchangeRule, NagyRule, OperatorRule)):
            raise TypeError("unknown transition rule %r" % (rule,))
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
        rule = sampler.rule
        if not isinstance(rule, LocalRule):
            coup, op, n = None, None, 0
            if isinstance(rule, (ExchangeRule, NagyRule)):
                pairs = rule.distances if isinstance(rule, ExchangeRule) else rule.adjacency_list
                coup = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2) - 1)     # 0-based
                n = coup.shape[0]
            else:
                self._rule_op = rule.operator.to_device(net.ctx)       # kept alive with the chain
                op = self._rule_op.h
            L.check(L.lib.nq_sampler_set_rule(h, rule.code, n, L.ptr(coup), op), net.ctx.h)

    def replay_rule(self, draws, uniforms):
        """One samplenext! of a non-local rule with supplied randomness: draws [passes, B, 4] (the integers the reference
        draws, see nq_sampler_replay_rule), uniforms [passes, B]."""
        draws = np.ascontiguousarray(draws, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(self.s.passes, self.B, 4)
        uniforms = np.ascontiguousarray(uniforms, dtype=self.net.rdtype)
        acc = np.zeros((self.s.passes, self.B), dtype=np.uint8)
        L.check(L.lib.nq_sampler_replay_rule(self.h, L.ptr(draws), L.ptr(uniforms), L.ptr(acc)), self.net.ctx.h)
        return acc.astype(bool)

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
    """Exact.jl:135-181 on the device (csrc/nq_fullspace.cu): log-probabilities of ALL basis states through the machine
    kernel (basis number = packed digits, site 1 least significant; super-index = col * D + row), a cumulative table
    built by a deterministic device scan (nq_exact_table), and one inverse-CDF draw per (slot, chain) with Philox
    uniforms keyed by the global chain id (nq_exact_sample).  Julia's MersenneTwister stream is not reproduced; replay
    with supplied uniforms gives the reference's searchsortedfirst indices bit for bit."""
    MAX_TABLE = 1 << 24

    def __init__(self, sampler, net, batch_sz, chain_offset=0, num_workers=1):
        import torch
        self.s, self.net, self.B = sampler, net, int(batch_sz)
        self.loc_chain_length = -(-sampler.samples_length // num_workers)            # Exact.jl:41
        self.chain_offset = int(chain_offset)
        size = C.c_int64()
        L.check(L.lib.nq_fullspace_size(net.h, C.byref(size)), net.ctx.h)
        self.size = size.value
        self.D = 1 << net.N
        if self.size > self.MAX_TABLE:
            raise ValueError("ExactSampler: %d table entries (Closed N <= 24, Open N <= 12)" % self.size)
        self.cdf = torch.zeros(self.size, dtype=torch.float64, device=torch.device("cuda", net.ctx.device))
        self.draws = 0
        self.valid = False

    def init_sampler(self):
        """init_sampler!: the probability table of the CURRENT parameters."""
        L.check(L.lib.nq_exact_table(self.net.h, self.cdf.data_ptr()), self.net.ctx.h)
        self.valid = True
        return self.cdf

    def sample_into(self, L_store, prow, pcol, uniforms=None, indices=None):
        """samplenext! for every (slot, chain): packed words [L, B, 1] on the device; `uniforms` [L, B] replays."""
        self.init_sampler()
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        L.check(L.lib.nq_exact_sample(self.net.h, self.cdf.data_ptr(), self.s.seed, self.chain_offset, self.draws,
                                      self.B, int(L_store), L.ptr(u), L.ptr(prow), L.ptr(pcol), L.ptr(indices)),
                self.net.ctx.h)
        if uniforms is None:
            self.draws += int(L_store)
