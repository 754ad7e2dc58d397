"""Oracle: Metropolis-Hastings local-flip sampler in REPLAY mode.  Test infrastructure only.

Follows src/Samplers/Metropolis.jl:124-167 (samplenext!), src/Samplers/MCMCRules/LocalRule.jl:19-28,
flips in src/Hilbert/HomogeneousSpin.jl:95-104, HomogeneousFock.jl:72-81, DoubledHilbert.jl:19-27.
Julia's MersenneTwister stream cannot be reproduced without Julia, so the proposal sites and
the uniforms are INPUTS (replay): per pass all B sites (chain order) then all B uniforms --
exactly the reference's consumption order (Metropolis.jl:141-149).
"""
import numpy as np
from .machines import log_prob


def _logp(net, state):
    if net.doubled:
        return log_prob(net.logpsi(state[0], state[1]))
    return log_prob(net.logpsi(state))


def propose(hilb, state, sites):
    """LocalRule.propose_step! on a batch; `sites` 1-based in 1..N (ket) or 1..2N (doubled:
    j<=N flips row, j>N flips col[j-N] -- DoubledHilbert.jl:19-27)."""
    doubled = isinstance(state, tuple)
    new = tuple(np.array(s, copy=True) for s in state) if doubled else np.array(state, copy=True)
    N = hilb.n
    for c, j in enumerate(sites):
        j = int(j)
        if doubled:
            arr, jj = (new[1], j - N) if j > N else (new[0], j)
        else:
            arr, jj = new, j
        arr[jj - 1, c] = hilb.flip_value(arr[jj - 1, c])
    return new


def samplenext_replay(net, hilb, state, sites, uniforms, dtype=np.float64):
    """One stored sample per chain = `passes` MH steps.
    state: [N,B] array (ket) or (row, col) tuple.  sites, uniforms: [passes, B].
    Accept iff u - exp(lp' - lp) < 0, evaluated in precision `dtype` (Metropolis.jl:148-154).
    Returns (new_state, accept[passes, B] bool, margin[passes, B] = u - exp(dlp) in float64)."""
    doubled = isinstance(state, tuple)
    cur = tuple(np.array(s, dtype=np.float64) for s in state) if doubled \
        else np.array(state, dtype=np.float64)
    lp = _logp(net, cur)
    passes, B = np.shape(sites)
    acc = np.zeros((passes, B), dtype=bool)
    margin = np.zeros((passes, B), dtype=np.float64)
    for i in range(passes):
        prop = propose(hilb, cur, sites[i])
        lpp = _logp(net, prop)
        ratio = np.exp(lpp - lp)
        margin[i] = np.asarray(uniforms[i], np.float64) - ratio
        prob = np.asarray(uniforms[i], dtype) - ratio.astype(dtype)
        rejected = prob >= 0
        a = ~rejected
        acc[i] = a
        if doubled:
            cur = tuple(np.where(a[None, :], p, c) for p, c in zip(prop, cur))
        else:
            cur = np.where(a[None, :], prop, cur)
        lp = np.where(a, lpp, lp)
    return cur, acc, margin


def samplenext_diagonal_replay(net, hilb, state, sites, uniforms, dtype=np.float64):
    """Diagonal chain of the density-matrix observables sampler (BatchedObsDMSampler.jl:59-105; diagonal evaluation
    base_batched_networks.jl:244-246): sigma' = sigma, a proposal flips site j (1..N) of both, the log-probability is
    Re log rho(sigma, sigma) -- the mathematically intended reading of base_batched_networks.jl:249-253, which writes
    abs.(log rho) (not reproduced).  state [N, B]; sites, uniforms [passes, B]."""
    cur = np.array(state, dtype=np.float64)
    lp = np.real(net.logpsi(cur, cur))
    passes, B = np.shape(sites)
    acc = np.zeros((passes, B), dtype=bool)
    margin = np.zeros((passes, B), dtype=np.float64)
    for i in range(passes):
        prop = propose(hilb, cur, sites[i])
        lpp = np.real(net.logpsi(prop, prop))
        ratio = np.exp(lpp - lp)
        margin[i] = np.asarray(uniforms[i], np.float64) - ratio
        a = (np.asarray(uniforms[i], dtype) - ratio.astype(dtype)) < 0
        acc[i] = a
        cur = np.where(a[None, :], prop, cur)
        lp = np.where(a, lpp, lp)
    return cur, acc, margin
