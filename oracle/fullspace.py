"""Oracle: full-space reconstruction and the ExactSampler.  Test infrastructure only.

Follows src/utils/densitymatrix.jl:9-62 (densitymatrix, ket) and src/Samplers/Exact.jl:135-181 (init_sampler!:
pdf[h] = exp(log_prob_psi(state h)), cumsum!, ./= total; samplenext!: searchsortedfirst(pdf, rand())).
Basis numbers follow set!(sigma, hilb, i): digits of i-1, site 1 least significant (oracle/hilbert.py); the
super-operator space orders (row, col) with the row index fastest.
"""
import numpy as np


def ket(net, hilb, norm=True):
    """densitymatrix.jl:46-62: psi[i] = exp(net(state i)); normalize!(psi)."""
    psi = np.exp(net.logpsi(hilb.all_states()))
    return psi / np.linalg.norm(psi) if norm else psi


def densitymatrix(net, hilb, norm=True):
    """densitymatrix.jl:9-33: rho[i, j] = exp(net(row = state i, col = state j)); rho ./= tr(rho)."""
    S = hilb.all_states()
    D = S.shape[1]
    R = np.repeat(S[:, :, None], D, axis=2).reshape(hilb.n, D * D, order="F")      # row index fastest
    C = np.repeat(S[:, None, :], D, axis=1).reshape(hilb.n, D * D, order="F")
    rho = np.exp(net.logpsi(R, C)).reshape(D, D, order="F")
    return rho / np.trace(rho) if norm else rho


def exact_cdf(net, hilb):
    """Exact.jl:140-157.  The reference exponentiates log_prob_psi directly; the table is invariant under the common
    factor exp(-max) the device path takes out for range safety."""
    if net.doubled:
        S = hilb.all_states()
        D = S.shape[1]
        R = np.repeat(S[:, :, None], D, axis=2).reshape(hilb.n, D * D, order="F")
        C = np.repeat(S[:, None, :], D, axis=1).reshape(hilb.n, D * D, order="F")
        lp = 2.0 * np.real(net.logpsi(R, C))
    else:
        lp = 2.0 * np.real(net.logpsi(hilb.all_states()))
    pdf = np.exp(lp - lp.max())
    tot = pdf.sum()
    return np.cumsum(pdf) / tot


def exact_draw(cdf, r):
    """Exact.jl:172-176: hi = searchsortedfirst(pdf, r) (1-based basis number), clamped to the table."""
    return np.minimum(np.searchsorted(cdf, r, side="left"), len(cdf) - 1) + 1
