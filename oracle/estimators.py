"""Oracle: local estimators built from operator connections.  Test infrastructure only.

Follows
  src/IterativeInterface/Accumulators/AccumulatorObsScalar.jl:70-132  (E_loc)
  src/IterativeInterface/Accumulators/AccumulatorObsGrad.jl:63-127    (L_loc, grad L_loc)
  src/IterativeInterface/Samplers/CostFun/BatchedValSampler.jl:70-95
  src/IterativeInterface/Samplers/CostFun/BatchedGradSampler.jl:77-97
The reference buffers `local_batch_sz` connected configurations per machine call; batching does
not change any value, so the oracle evaluates all connections of a sample in one call.
"""
import numpy as np
from .operators import connections_ket, connections_super, apply_changes, n_changes


def local_scalar_ket(net, op, sigmas, logpsi=None):
    """E_loc(s) = sum_{no-change} mel + sum_{others} mel*exp(logpsi(eta)-logpsi(s)); zero mels
    skipped (AccumulatorObsScalar.jl:94-107,123-132).  sigmas [N, Ns]."""
    sigmas = np.asarray(sigmas, dtype=np.float64)
    Ns = sigmas.shape[1]
    if logpsi is None:
        logpsi = net.logpsi(sigmas)
    out = np.zeros(Ns, dtype=np.complex128)
    for s in range(Ns):
        sig = sigmas[:, s]
        res = 0.0 + 0.0j
        mels, etas = [], []
        for mel, cng in connections_ket(op, sig):
            if mel == 0.0:
                continue
            if n_changes(cng) == 0:
                res += mel
            else:
                mels.append(mel)
                etas.append(apply_changes(sig, cng))
        if mels:
            lp = net.logpsi(np.stack(etas, axis=1))
            res += np.sum(np.array(mels) * np.exp(lp - logpsi[s]))
        out[s] = res
    return out


def local_scalar_super(net, liouv, srow, scol, logpsi=None):
    """Scalar accumulator on doubled states (AccumulatorObsScalar.jl:70-88): diagonal mels are
    summed directly, the others weighted by rho(eta)/rho(sigma)."""
    srow, scol = np.asarray(srow, np.float64), np.asarray(scol, np.float64)
    Ns = srow.shape[1]
    if logpsi is None:
        logpsi = net.logpsi(srow, scol)
    out = np.zeros(Ns, dtype=np.complex128)
    for s in range(Ns):
        r, c = srow[:, s], scol[:, s]
        res = 0.0 + 0.0j
        mels, er, ec = [], [], []
        for mel, cl, cr in connections_super(liouv, r, c):
            if mel == 0.0:
                continue
            if n_changes(cl) == 0 and n_changes(cr) == 0:
                res += mel
            else:
                mels.append(mel)
                er.append(apply_changes(r, cl))
                ec.append(apply_changes(c, cr))
        if mels:
            lp = net.logpsi(np.stack(er, 1), np.stack(ec, 1))
            res += np.sum(np.array(mels) * np.exp(lp - logpsi[s]))
        out[s] = res
    return out


def local_grad_super(net, liouv, srow, scol, logpsi=None):
    """L_loc = sum_c mel_c r_c ; grad L_loc = sum_c mel_c r_c grad logrho(eta_c), with
    r_c = exp(logrho(eta_c) - logrho(sigma)) for EVERY non-zero-mel connection including the
    diagonal ones (the shortcut is disabled by `&& false`, AccumulatorObsGrad.jl:72); the
    -grad logrho(sigma) term is absent (quirk Q10, :113).  Returns ([Ns], [P, Ns])."""
    srow, scol = np.asarray(srow, np.float64), np.asarray(scol, np.float64)
    Ns = srow.shape[1]
    if logpsi is None:
        logpsi = net.logpsi(srow, scol)
    loc = np.zeros(Ns, dtype=np.complex128)
    gloc = np.zeros((net.P, Ns), dtype=np.complex128)
    for s in range(Ns):
        r, c = srow[:, s], scol[:, s]
        mels, er, ec = [], [], []
        for mel, cl, cr in connections_super(liouv, r, c):
            if mel == 0.0:
                continue
            mels.append(mel)
            er.append(apply_changes(r, cl))
            ec.append(apply_changes(c, cr))
        if not mels:
            continue
        lp, O = net.logpsi_grad(np.stack(er, 1), np.stack(ec, 1))
        w = np.array(mels) * np.exp(lp - logpsi[s])
        loc[s] = w.sum()
        gloc[:, s] = O @ w
    return loc, gloc


def connection_list_ket(op, sigma):
    """Integer artefact compared bit-exactly with nq_connections: per connection
    (mel, flip mask over N sites as python int, bit j-1 <-> site j).  Zero mels included."""
    out = []
    for mel, cng in connections_ket(op, sigma):
        m = 0
        for site in (cng[0] if cng is not None else ()):
            m |= 1 << (site - 1)
        out.append((complex(mel), m))
    return out


def connection_list_super(liouv, row, col):
    out = []
    for mel, cl, cr in connections_super(liouv, row, col):
        ml = mr = 0
        for site in (cl[0] if cl is not None else ()):
            ml |= 1 << (site - 1)
        for site in (cr[0] if cr is not None else ()):
            mr |= 1 << (site - 1)
        out.append((complex(mel), ml, mr))
    return out
