"""Oracle: Metropolis-Hastings transition rules other than LocalRule, in REPLAY mode.  Test infrastructure only.

Follows src/Samplers/MCMCRules/ExchangeRule.jl:14-68, Nagy.jl:7-115, OperatorRule.jl:27-59 and the accept test of
src/Samplers/Metropolis.jl:141-149 (prob = rand - exp(lp' - lp + log_prob_bias); accepted iff prob < 0).
Julia's MersenneTwister stream cannot be reproduced, so the integers each rule draws are INPUTS: per proposal a
4-vector `draws` (1-based like Julia's rand(1:n)):
    Exchange: (couple, -, -, -)
    Nagy    : (move 1..8, site 1..N, aux, aux2): aux = element 1..2 of the couple adjacency_list[site] (moves 1-4; the
              reference indexes its list of COUPLES by the site number, Nagy.jl:51), rand(1:10) of the row side (move 7;
              aux2 = that of the column side; consumed only when the site is empty), column site (move 8)
    Operator: (r, -, -, -): r a raw 32-bit random; the connection drawn by rand(conns) is number floor(r n / 2^32).
"""
import numpy as np

from . import operators as OO
from .sampler import _logp


def couplings(op):
    """ExchangeRule(H) / NagyRule(H): the site pairs of the 2-site terms, in term order (ExchangeRule.jl:17-34)."""
    out = []
    for t in OO.terms(op):
        if len(t.sites) == 2:
            out.append((int(t.sites[0]), int(t.sites[1])))
    return out


def _flip(hilb, arr, site, c):
    arr[site - 1, c] = hilb.flip_value(arr[site - 1, c])


def propose(rule, hilb, state, draws, operator=None, coup=None):
    """propose_step! on a batch.  Returns (new_state, log_prob_bias[B])."""
    doubled = isinstance(state, tuple)
    new = tuple(np.array(s, copy=True) for s in state) if doubled else np.array(state, copy=True)
    B = draws.shape[0]
    bias = np.zeros(B)
    N = hilb.n
    for c in range(B):
        d = [int(x) for x in draws[c]]
        if rule == "exchange":                         # ExchangeRule.jl:55-66
            i, j = coup[d[0] - 1]
            new[i - 1, c], new[j - 1, c] = new[j - 1, c], new[i - 1, c]
        elif rule == "nagy":                           # Nagy.jl:40-95
            row, col = new
            move, s1 = d[0], d[1]
            if move in (1, 2, 3, 4):
                s2 = coup[s1 - 1][d[2] - 1]
                arr = row if move <= 2 else col
                _flip(hilb, arr, s1, c)
                _flip(hilb, arr, s2, c)
            elif move == 5:
                _flip(hilb, row, s1, c)
            elif move == 6:
                _flip(hilb, col, s1, c)
            elif move == 7:
                if row[s1 - 1, c] == 0:
                    if d[2] == 1:
                        _flip(hilb, row, s1, c)
                else:
                    _flip(hilb, row, s1, c)
                if col[s1 - 1, c] == 0:
                    if d[3] == 1:
                        _flip(hilb, col, s1, c)
                else:
                    _flip(hilb, col, s1, c)
            else:
                _flip(hilb, row, s1, c)
                _flip(hilb, col, d[2], c)
        else:                                          # OperatorRule.jl:42-58
            r = d[0] & 0xFFFFFFFF
            if doubled:
                conns = OO.connections_super(operator, new[0][:, c], new[1][:, c])
                k = (r * len(conns)) >> 32
                _, cl, cr = conns[k]
                new[0][:, c] = OO.apply_changes(new[0][:, c], cl)
                new[1][:, c] = OO.apply_changes(new[1][:, c], cr)
                nb = len(OO.connections_super(operator, new[0][:, c], new[1][:, c]))
            else:
                conns = OO.connections_ket(operator, new[:, c])
                k = (r * len(conns)) >> 32
                new[:, c] = OO.apply_changes(new[:, c], conns[k][1])
                nb = len(OO.connections_ket(operator, new[:, c]))
            bias[c] = np.log(len(conns) / nb)
    return new, bias


def samplenext_rule_replay(net, hilb, state, rule, draws, uniforms, operator=None, coup=None, dtype=np.float64):
    """One stored sample per chain.  draws [passes, B, 4], uniforms [passes, B].
    Returns (new_state, accept[passes, B], margin[passes, B] = u - exp(dlp + bias))."""
    doubled = isinstance(state, tuple)
    cur = tuple(np.array(s, dtype=np.float64) for s in state) if doubled else np.array(state, dtype=np.float64)
    lp = _logp(net, cur)
    passes, B = np.shape(uniforms)
    acc = np.zeros((passes, B), dtype=bool)
    margin = np.zeros((passes, B))
    for i in range(passes):
        prop, bias = propose(rule, hilb, cur, draws[i], operator, coup)
        lpp = _logp(net, prop)
        ratio = np.exp(lpp - lp + bias)
        margin[i] = np.asarray(uniforms[i], np.float64) - ratio
        a = (np.asarray(uniforms[i], dtype) - ratio.astype(dtype)) < 0
        acc[i] = a
        if doubled:
            cur = tuple(np.where(a[None, :], p, c) for p, c in zip(prop, cur))
        else:
            cur = np.where(a[None, :], prop, cur)
        lp = np.where(a, lpp, lp)
    return cur, acc, margin
