"""ctypes binding of libnqcuda (include/nqcuda.h).  No torch types cross this boundary: only
pointers (host numpy buffers or raw device addresses) and sizes.

The library is built in-tree by ``neuralquantum.jl_b200/build.py``; if it is missing the import
fails loudly -- there is no CPU fallback anywhere in the product path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NQCUDA_LIB", os.path.join(os.path.dirname(HERE), "libnqcuda.so"))

# enums (mirror include/nqcuda.h)
NQ_OK = 0
NQ_ERR_ARG, NQ_ERR_SHAPE, NQ_ERR_CUDA, NQ_ERR_NCCL = -1, -2, -3, -4
NQ_ERR_NOT_POSDEF, NQ_ERR_NOT_CONVERGED, NQ_ERR_UNSUPPORTED, NQ_ERR_ALLOC = -5, -6, -7, -8
NQ_RBM, NQ_RBMSPLIT, NQ_NDM = 0, 1, 2
NQ_SOFTPLUS, NQ_LOGCOSH = 0, 1
NQ_F32, NQ_F64, NQ_C64, NQ_C128 = 0, 1, 2, 3
NQ_SPIN, NQ_FOCK = 0, 1
NQ_KET, NQ_SUPER = 0, 1
NQ_SOLVE_CHOLESKY, NQ_SOLVE_CG, NQ_SOLVE_MINRES, NQ_SOLVE_QLP, NQ_SOLVE_QLP_WARM = 0, 1, 2, 3, 4
NQ_RULE_LOCAL, NQ_RULE_EXCHANGE, NQ_RULE_NAGY, NQ_RULE_OPERATOR = 0, 1, 2, 3
NQ_UNIQUE_ID_BYTES = 128

NP_OF = {NQ_F32: np.float32, NQ_F64: np.float64, NQ_C64: np.complex64, NQ_C128: np.complex128}
NQ_OF = {np.dtype(v): k for k, v in NP_OF.items()}


def nq_dtype(dt):
    return NQ_OF[np.dtype(dt)]


def complex_of(code):
    return NQ_C128 if code in (NQ_F64, NQ_C128) else NQ_C64


def real_of(code):
    return NQ_F64 if code in (NQ_F64, NQ_C128) else NQ_F32


class NQError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libnqcuda: %s (status %d)" % (msg, status))
        self.status = status


class PosDefException(NQError):
    """Mirror of LinearAlgebra.PosDefException raised by cholesky!(..., check=true) (SRDirect.jl:79)."""


class NotConvergedError(NQError):
    """CG exhausted maxiter (SRIterative.jl:133-150)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libnqcuda.so not found at %s -- run `python neuralquantum.jl_b200/build.py` "
        "(the product has no CPU fallback)" % LIB_PATH)

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

_vp, _i32, _i64, _u64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_PROTOS = {
    "nq_version": (C.c_int, []),
    "nq_status_string": (C.c_char_p, [_i32]),
    "nq_ctx_create": (_i32, [_i32, _vp, C.POINTER(_vp)]),
    "nq_ctx_destroy": (_i32, [_vp]),
    "nq_last_error": (C.c_char_p, [_vp]),
    "nq_ctx_sync": (_i32, [_vp]),
    "nq_ctx_launch_count": (_i32, [_vp, C.POINTER(_u64)]),
    "nq_ctx_last_info": (_i32, [_vp, C.POINTER(_i64)]),
    "nq_states_words": (_i32, [_i32]),
    "nq_pack_states": (_i32, [_vp, _i32, _i32, _i64, _vp, _i32, _vp]),
    "nq_unpack_states": (_i32, [_vp, _i32, _i32, _i64, _vp, _vp, _i32]),
    "nq_machine_create": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(_vp)]),
    "nq_machine_destroy": (_i32, [_vp]),
    "nq_machine_nparams": (_i32, [_vp, C.POINTER(_i64)]),
    "nq_machine_out_dtype": (_i32, [_vp, C.POINTER(_i32)]),
    "nq_machine_set_params": (_i32, [_vp, _vp, _i64]),
    "nq_machine_get_params": (_i32, [_vp, _vp, _i64]),
    "nq_logpsi": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp]),
    "nq_log_prob": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp]),
    "nq_logpsi_grad": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _i64]),
    "nq_logpsi_packed": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "nq_logpsi_grad_packed": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64]),
    "nq_operator_create": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, C.POINTER(_vp)]),
    "nq_operator_destroy": (_i32, [_vp]),
    "nq_operator_max_connections": (_i32, [_vp, C.POINTER(_i64)]),
    "nq_connections": (_i32, [_vp, _i32, _vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp]),
    "nq_local_scalar": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "nq_local_grad": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _i64]),
    "nq_logpsi_grad_local_packed": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64]),
    "nq_logpsi_grad_local_host": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64]),
    "nq_local_scalar_packed": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "nq_local_grad_packed": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64]),
    "nq_sampler_create": (_i32, [_vp, _i64, _i32, _u64, _i64, C.POINTER(_vp)]),
    "nq_sampler_destroy": (_i32, [_vp]),
    "nq_sampler_set_state": (_i32, [_vp, _vp, _vp, _i32]),
    "nq_sampler_get_state": (_i32, [_vp, _vp, _vp, _i32]),
    "nq_sampler_randomize": (_i32, [_vp]),
    "nq_sampler_set_mode": (_i32, [_vp, _i32]),
    "nq_sampler_replay": (_i32, [_vp, _vp, _vp, _vp]),
    "nq_sampler_sample": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32]),
    "nq_sampler_set_rule": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "nq_sampler_replay_rule": (_i32, [_vp, _vp, _vp, _vp]),
    "nq_sampler_counters": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "nq_symm_create": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, C.POINTER(_vp)]),
    "nq_symm_destroy": (_i32, [_vp]),
    "nq_symm_nparams": (_i32, [_vp, C.POINTER(_i64)]),
    "nq_symm_set_params": (_i32, [_vp, _vp, _i64]),
    "nq_symm_get_params": (_i32, [_vp, _vp, _i64]),
    "nq_symm_update": (_i32, [_vp, _vp, _dbl]),
    "nq_symm_gradient": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _i64]),
    "nq_fullspace_size": (_i32, [_vp, C.POINTER(_i64)]),
    "nq_fullspace_state": (_i32, [_vp, _i32, _vp]),
    "nq_exact_table": (_i32, [_vp, _vp]),
    "nq_exact_sample": (_i32, [_vp, _vp, _u64, _i64, _u64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "nq_center": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "nq_center_lazy": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, C.POINTER(_i32)]),
    "nq_center_finish": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32]),
    "nq_force_ket": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp]),
    "nq_force_liouvillian": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, C.POINTER(_dbl)]),
    "nq_sr_setup": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _i32, _vp, _vp]),
    "nq_sr_hint_row_planes": (_i32, [_vp, _vp, _i64]),
    "nq_sr_solve": (_i32, [_vp, _vp, _vp, _i64, _i32, _dbl, _i32, _dbl, _i64, _vp, C.POINTER(_i64)]),
    "nq_sr_scale_diagonal": (_i32, [_vp, _vp, _i64, _i32, _dbl]),
    "nq_sr_solve_matfree": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _i32, _dbl, _dbl, _i64, _vp,
                                   C.POINTER(_i64)]),
    "nq_sr_solve_matfree_algo": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _i32, _dbl, _i32, _dbl, _i64, _vp,
                                        C.POINTER(_i64)]),
    "nq_sr_accumulate": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _i32]),
    "nq_sr_finish": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32]),
    "nq_nesterov": (_i32, [_vp, _vp, _vp, _i64, _i32, _dbl, _dbl, _vp]),
    "nq_update": (_i32, [_vp, _vp, _dbl]),
    "nq_stat_analysis": (_i32, [_vp, _vp, _i64, _i64, _i32, C.POINTER(_dbl)]),
    "nq_abs2": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "nq_comm_unique_id": (_i32, [_vp]),
    "nq_comm_init": (_i32, [_vp, _i32, _i32, _vp]),
    "nq_comm_destroy": (_i32, [_vp]),
    "nq_comm_size": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32)]),
    "nq_allreduce_sum": (_i32, [_vp, _vp, _i64, _i32]),
    "nq_allreduce_mean": (_i32, [_vp, _vp, _i64, _i32]),
    "nq_comm_set_global_samples": (_i32, [_vp, _i64]),
}
for _name, (_res, _args) in _PROTOS.items():
    _f = getattr(lib, _name)          # AttributeError here = header/library mismatch: fail loudly
    _f.restype = _res
    _f.argtypes = _args

EXPORTS = sorted(_PROTOS)


def ptr(x):
    """Address of a numpy array (host), a torch tensor (host or device), an int address, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError("cannot take the address of %r" % type(x))


def check(status, ctx=None):
    if status == NQ_OK:
        return
    msg = lib.nq_status_string(status).decode()
    if ctx is not None:
        detail = lib.nq_last_error(ctx).decode()
        if detail:
            msg = "%s: %s" % (msg, detail)
    if status == NQ_ERR_NOT_POSDEF:
        raise PosDefException(status, msg)
    if status == NQ_ERR_NOT_CONVERGED:
        raise NotConvergedError(status, msg)
    raise NQError(status, msg)
