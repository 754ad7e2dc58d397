"""Context and Hilbert spaces (host side of the drop-in boundary).

ref: src/Hilbert/HomogeneousSpin.jl, src/Hilbert/HomogeneousFock.jl (local dimension 2 only: the
device path packs configurations as bit-vectors), src/Parallel/* (one context per worker).
"""
import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    """One device + one stream (nq_ctx_t).  `stream` is a raw cudaStream_t address (e.g.
    torch.cuda.current_stream().cuda_stream); None = the CUDA default stream."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        L.check(L.lib.nq_ctx_create(int(device), stream, C.byref(h)))
        self.h = h
        self.device = int(device)
        self.stream = int(stream) if stream else 0

    def torch_stream(self):
        """The context's stream as a torch stream: the host mirror's few torch ops (buffer fills, F copy, Nesterov
        velocity, ExactSampler table) run under it so that they are ordered with the library's kernels."""
        import torch
        if not self.stream:
            return torch.cuda.default_stream(torch.device("cuda", self.device))
        return torch.cuda.ExternalStream(self.stream, device=torch.device("cuda", self.device))

    def close(self):
        if getattr(self, "h", None):
            L.lib.nq_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(L.lib.nq_ctx_sync(self.h), self.h)

    @property
    def launches(self):
        n = C.c_uint64()
        L.check(L.lib.nq_ctx_launch_count(self.h, C.byref(n)), self.h)
        return n.value

    # ---- parallel backend (ref: Parallel/not_parallel.jl, Parallel/MPI/mpi.jl) ----
    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_uint8 * L.NQ_UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        L.check(L.lib.nq_comm_init(self.h, nranks, rank, buf), self.h)

    def comm_size(self):
        n, r = C.c_int(), C.c_int()
        L.check(L.lib.nq_comm_size(self.h, C.byref(n), C.byref(r)), self.h)
        return n.value, r.value

    def allreduce_sum(self, dev_ptr, n, dtype):
        L.check(L.lib.nq_allreduce_sum(self.h, dev_ptr, n, dtype), self.h)

    def set_global_samples(self, ns_total):
        """Global sample count nq_center / nq_force_* normalise by (0 = Ns * nranks)."""
        L.check(L.lib.nq_comm_set_global_samples(self.h, int(ns_total)), self.h)

    def allreduce_mean(self, dev_ptr, n, dtype):
        L.check(L.lib.nq_allreduce_mean(self.h, dev_ptr, n, dtype), self.h)

    def pack(self, hilb, sigma):
        """float configurations [N, B] (column-major / Fortran order) -> uint64 words [B, W64]."""
        sigma = np.asfortranarray(sigma)
        N, B = sigma.shape
        out = np.zeros((B, L.lib.nq_states_words(N)), dtype=np.uint64)
        L.check(L.lib.nq_pack_states(self.h, hilb.code, N, B, L.ptr(sigma), L.nq_dtype(sigma.dtype), L.ptr(out)), self.h)
        return out

    def unpack(self, hilb, packed, dtype=np.float64):
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        B = packed.shape[0]
        out = np.zeros((hilb.n, B), dtype=dtype, order="F")
        L.check(L.lib.nq_unpack_states(self.h, hilb.code, hilb.n, B, L.ptr(packed), L.ptr(out), L.nq_dtype(dtype)), self.h)
        return out


def unique_id():
    buf = (C.c_uint8 * L.NQ_UNIQUE_ID_BYTES)()
    L.check(L.lib.nq_comm_unique_id(buf))
    return bytes(buf)


class Hilbert:
    def __init__(self, n, kind):
        self.n = int(n)
        self.kind = kind
        self.code = L.NQ_SPIN if kind == "spin" else L.NQ_FOCK

    def values(self, digits):
        digits = np.asarray(digits)
        return 2 * digits - 1 if self.kind == "spin" else digits

    def spacedimension(self):
        return 2 ** self.n

    def state(self, index, dtype=np.float64):
        """set!(sigma, hilb, index): 1-based index, site 1 least significant."""
        v = int(index) - 1
        return self.values(np.array([(v >> i) & 1 for i in range(self.n)])).astype(dtype)

    def toint(self, sigma):
        d = (np.asarray(sigma).real > (0 if self.kind == "spin" else 0.5)).astype(np.int64)
        return int(sum(int(b) << i for i, b in enumerate(d))) + 1

    def random_states(self, B, rng, dtype=np.float64):
        return np.asfortranarray(self.values(rng.integers(0, 2, size=(self.n, B))).astype(dtype))

    def __eq__(self, o):
        return isinstance(o, Hilbert) and (self.n, self.kind) == (o.n, o.kind)

    def __repr__(self):
        return "Homogeneous%s(%d, 2)" % ("Spin" if self.kind == "spin" else "Fock", self.n)


def HomogeneousSpin(n, S=None):
    return Hilbert(n, "spin")


def HomogeneousFock(n, d=2):
    if d != 2:
        raise NotImplementedError("the device path packs configurations as bits: local dimension 2 only")
    return Hilbert(n, "fock")
