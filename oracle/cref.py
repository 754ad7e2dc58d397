"""ctypes wrapper of oracle/cref.c (C restatement of the reference CPU algorithm; test/baseline
infrastructure only).  Tables are flattened from the ORACLE's operator objects."""
import ctypes as C
import os

import numpy as np

from . import build_cref
from .operators import KLocalLiouvillian, terms

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_cref.LIB if os.path.exists(build_cref.LIB) else build_cref.build()
        _lib = C.CDLL(path)
        _lib.cref_max_threads.restype = C.c_int
    return _lib


def flatten(liouv):
    """oracle KLocalLiouvillian -> the part/term arrays cref.c walks (flip mask = which of the part's
    sites change; valid for local dimension 2)."""
    assert isinstance(liouv, KLocalLiouvillian)
    parts, pid, pairs = [], {}, []

    def part(op):
        if op is None:
            return -1
        if id(op) not in pid:
            pid[id(op)] = len(parts)
            parts.append(op)
        return pid[id(op)]
    for grp in (liouv.HnH_l, liouv.HnH_r, liouv.LLdag):
        for t in terms(grp):
            pairs.append((part(t.op_l), part(t.op_r)))
    nsites = np.array([len(p.sites) for p in parts], np.int32)
    site_ptr = np.concatenate([[0], np.cumsum(nsites)]).astype(np.int32)
    sites = np.array([s - 1 for p in parts for s in p.sites], np.int32)
    row0 = np.concatenate([[0], np.cumsum(2 ** nsites.astype(np.int64))])[:-1].astype(np.int64)
    row_ptr, mel, flip = [0], [], []
    for p in parts:
        for row in p.op_conns:
            for m, (cs, _) in row:
                f = 0
                for s in cs:
                    f |= 1 << p.sites.index(s)
                mel.append(complex(m))
                flip.append(f)
            row_ptr.append(len(mel))
    return dict(n_parts=len(parts), nsites=nsites, site_ptr=site_ptr, sites=sites, row0=row0,
                row_ptr=np.array(row_ptr, np.int64), mel=np.array(mel, np.complex128), flip=np.array(flip, np.uint32),
                n_terms=len(pairs), left=np.array([a for a, _ in pairs], np.int32),
                right=np.array([b for _, b in pairs], np.int32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ndm_logpsi_grad(net, sr, sc, grad=True, nthreads=0):
    sr = np.ascontiguousarray(np.asarray(sr, np.float64).T)      # [B, N] row-major == [N, B] column-major
    sc = np.ascontiguousarray(np.asarray(sc, np.float64).T)
    B = sr.shape[0]
    par = np.ascontiguousarray(net.params(), np.float64)
    out = np.zeros(B, np.complex128)
    O = np.zeros((B, net.P), np.complex128) if grad else None
    lib().cref_ndm_logpsi_grad(_p(par), net.N, net.M, net.A, int(net.act), _p(sr), _p(sc), C.c_int64(B), _p(out),
                               _p(O) if grad else None, int(nthreads))
    return out, (O.T if grad else None)


def local_grad_super(net, liouv, sr, sc, grad=True, nthreads=0, tables=None):
    tb = tables or flatten(liouv)
    sr = np.ascontiguousarray(np.asarray(sr, np.float64).T)
    sc = np.ascontiguousarray(np.asarray(sc, np.float64).T)
    B = sr.shape[0]
    par = np.ascontiguousarray(net.params(), np.float64)
    loc = np.zeros(B, np.complex128)
    g = np.zeros((B, net.P), np.complex128) if grad else None
    lib().cref_local_grad(_p(par), net.N, net.M, net.A, int(net.act), int(liouv.hilb.kind == "fock"),
                          tb["n_parts"], _p(tb["nsites"]), _p(tb["site_ptr"]), _p(tb["sites"]), _p(tb["row0"]),
                          _p(tb["row_ptr"]), _p(tb["mel"]), _p(tb["flip"]), tb["n_terms"], _p(tb["left"]), _p(tb["right"]),
                          _p(sr), _p(sc), C.c_int64(B), _p(loc), _p(g) if grad else None, int(nthreads))
    return loc, (g.T if grad else None)


def max_threads():
    return lib().cref_max_threads()
