"""BatchedSampler: the per-iteration driver (sample! -> precondition! -> update!) with every
buffer resident on the GPU; PyTorch is used only to own device memory.

ref: src/IterativeInterface/build_Batched.jl:1-14, Samplers/CostFun/BatchedValSampler.jl:117-139,
     Samplers/CostFun/BatchedGradSampler.jl:77-123, Samplers/BaseIterativeSampler.jl:5-26,
     src/Algorithms/SR/SRDirect.jl, SRIterative.jl.
Differences kept on purpose: all L slices of the chain are filled (Q3), S is
normalised by the global sample count under sharding (Q5), statistics are global (Q15).
"""
import copy
import ctypes as C

import numpy as np

from . import _lib as L
import warnings

from .algorithms import Measurement, SR, sr_cg, sr_cholesky, sr_minres, sr_qlp, stat_analysis
from .operators import Liouvillian
from .samplers import ExactSampler, ExactSamplerCache, MetropolisSamplerCache


def _torch():
    import torch
    return torch


_TORCH_OF = None


def _tdtype(np_dtype):
    torch = _torch()
    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[np.dtype(np_dtype)]


class BatchedSampler:
    """BatchedSampler(net, sampler, problem, algo; batch_sz).  `problem` is a LocalOperator (ground
    state, BatchedValSampler) or a Liouvillian (steady state, BatchedGradSampler)."""

    def __init__(self, net, sampler, problem, algo=None, batch_sz=16, chain_length=None):
        torch = _torch()
        self.net, self.sampler, self.problem = net, sampler, problem
        self.algo = algo if algo is not None else SR()
        self.ctx = ctx = net.ctx
        self.is_liouvillian = isinstance(problem, Liouvillian)
        if self.is_liouvillian != net.doubled:
            raise ValueError("ket machines pair with Hamiltonians, density-matrix machines with Liouvillians")
        self.op = problem.to_device(ctx)
        self.nranks, self.rank = ctx.comm_size()
        self.B = int(batch_sz)
        dev = torch.device("cuda", ctx.device)
        # chains of this rank inside the global batch: ranks may own different numbers of chains (shard_chains hands
        # the remainder to the first ranks), so the offset and the global count come from one all-reduce of the counts
        self.chain_offset, self.B_total = self.rank * self.B, self.B * self.nranks
        if self.nranks > 1:
            counts = torch.zeros(self.nranks, dtype=torch.float64, device=dev)
            counts[self.rank] = self.B
            torch.cuda.current_stream(dev).synchronize()
            ctx.allreduce_sum(counts.data_ptr(), self.nranks, L.NQ_F64)
            counts = [int(round(c)) for c in counts.cpu().tolist()]
            self.chain_offset, self.B_total = sum(counts[:self.rank]), sum(counts)
        cache_type = ExactSamplerCache if isinstance(sampler, ExactSampler) else MetropolisSamplerCache
        self.cache = cache_type(sampler, net, self.B, chain_offset=self.chain_offset, num_workers=self.nranks)
        self.L = self.cache.loc_chain_length if chain_length is None else int(chain_length)
        self.Ns = self.B * self.L
        self.Ns_total = self.B_total * self.L
        W = L.lib.nq_states_words(net.N)
        P, Ns = net.P, self.Ns
        ot, ct = _tdtype(net.out_dtype), _tdtype(net.cdtype)
        self.prow = torch.zeros((self.L, self.B, W), dtype=torch.int64, device=dev)
        self.pcol = torch.zeros_like(self.prow) if net.doubled else None
        self.logpsi = torch.zeros(Ns, dtype=ot, device=dev)
        self._O = torch.zeros((Ns, P), dtype=ot, device=dev)           # = [P, Ns] column-major; see the `O` property
        self._center_pending = False
        self.loc = torch.zeros(Ns, dtype=ct, device=dev)
        self.gloc = torch.zeros((Ns, P), dtype=ct, device=dev) if self.is_liouvillian else None
        # symmetrised machines (NDMSymm): the kernels write rows of the BARE net, symmetrised into O / gloc afterwards
        self.symm = hasattr(net, "bare")
        if self.symm:
            self.O_bare = torch.zeros((Ns, net.Pb), dtype=ot, device=dev)
            self.gloc_bare = torch.zeros((Ns, net.Pb), dtype=ct, device=dev) if self.is_liouvillian else None
        self.avg = torch.zeros(P, dtype=ot, device=dev)
        self.gradC = torch.zeros(P, dtype=ct, device=dev)
        self.real_params = net.real_params
        st = _tdtype(net.rdtype) if self.real_params else ct
        self.sdtype = np.dtype(net.rdtype) if self.real_params else np.dtype(net.cdtype)
        explicit = self.algo.algorithm == sr_cholesky or self.algo.full_matrix
        self.S = torch.zeros((P, P), dtype=st, device=dev) if explicit else None
        self.F = torch.zeros(P, dtype=st, device=dev)
        self.dw = torch.zeros(P, dtype=st, device=dev)
        self.abs2 = torch.zeros(Ns, dtype=_tdtype(net.rdtype), device=dev) if self.is_liouvillian else None
        # structure of the gradient rows of the real-parameter NDM (NDMBatched.jl:262-277): b/h/w of mu purely imaginary,
        # b/h/w of lambda purely real, the ancilla fields complex -> the S assembly skips the zero planes without scanning O
        self.row_planes = None
        if net.kind == L.NQ_NDM:
            planes = {"b_mu": 2, "h_mu": 2, "w_mu": 2, "u_mu": 3, "b_lam": 1, "h_lam": 1, "d_lam": 3, "w_lam": 1, "u_lam": 3}
            self.row_planes = np.concatenate([np.full(int(np.prod(shp)), planes[f], dtype=np.uint8)
                                              for f, shp in zip(net.fields, net.shapes)])
            assert self.row_planes.size == P
        self.cost = None
        self.last_iters = 0

    @property
    def O(self):
        """The gradient rows [Ns, P] (= [P, Ns] column-major): as written by evaluate(), centred after assemble().  The
        Liouvillian + explicit-S iteration defers the subtraction of <O> (nq_center_lazy: the S assembly subtracts on the fly
        and nothing else reads the centred rows); reading this attribute applies a pending subtraction first."""
        if self._center_pending:
            L.check(L.lib.nq_center_finish(self.ctx.h, self._O.data_ptr(), self.net.P, self.net.P, self.Ns,
                                           L.nq_dtype(self.net.out_dtype)), self.ctx.h)
            self._center_pending = False
        return self._O

    # ---- the hot path -----------------------------------------------------------------
    def sample_states(self):
        """_sample_state!: re-randomise the chains, burn, fill the L slices."""
        if isinstance(self.cache, ExactSamplerCache):
            with _torch().cuda.stream(self.ctx.torch_stream()):
                self.cache.sample_into(self.L, self.prow, self.pcol)
            return
        self.cache.randomize()
        self.cache.sample(self.sampler.burn_length, self.L,
                          packed_out=(self.prow.data_ptr(), self.pcol.data_ptr() if self.pcol is not None else None))

    def set_samples(self, sigma):
        """Bypass the sampler with supplied configurations [N, B, L] (parity tests / benchmarks)."""
        net, ctx = self.net, self.ctx
        sr, sc = sigma if net.doubled else (sigma, None)
        for arr, buf in ((sr, self.prow), (sc, self.pcol)):
            if arr is None:
                continue
            a = np.asfortranarray(arr).reshape(net.N, self.Ns, order="F")
            L.check(L.lib.nq_pack_states(ctx.h, net.hilb.code, net.N, self.Ns, L.ptr(a), L.nq_dtype(a.dtype),
                                         buf.data_ptr()), ctx.h)

    def evaluate(self):
        """logpsi_and_grad! + local estimator on the stored samples (one fused pass)."""
        net, ctx, Ns = self.net, self.ctx, self.Ns
        self._center_pending = False          # new rows are about to be written
        pc = self.pcol.data_ptr() if self.pcol is not None else None
        gl = self.gloc.data_ptr() if self.is_liouvillian else None
        if self.symm:       # logpsi_and_grad!(::NDMSymm): bare rows, then symmetrize_grad_NDM_batched! (NDMSymmBatched.jl:16-36)
            glb = self.gloc_bare.data_ptr() if self.is_liouvillian else None
            L.check(L.lib.nq_logpsi_grad_local_packed(net.h, self.op.h, self.prow.data_ptr(), pc, Ns, self.logpsi.data_ptr(),
                                                      self.O_bare.data_ptr(), net.Pb, self.loc.data_ptr(), glb, net.Pb), ctx.h)
            net.symmetrize(self.O_bare.data_ptr(), net.Pb, Ns, self._O.data_ptr(), net.P)
            if self.is_liouvillian:
                net.symmetrize(glb, net.Pb, Ns, gl, net.P)
            return
        L.check(L.lib.nq_logpsi_grad_local_packed(net.h, self.op.h, self.prow.data_ptr(), pc, Ns, self.logpsi.data_ptr(),
                                                  self._O.data_ptr(), net.P, self.loc.data_ptr(), gl, net.P), ctx.h)

    def evaluate_host(self, sigma, chunks=None):
        """set_samples + evaluate for HOST configurations [N, B, L] (pinned memory makes the copies asynchronous), software
        pipelined: the batch is cut into pieces; the float arrays of piece c+1 are copied on a side stream while the
        fused kernel works on piece c (the kernel is persistent and fills every SM, so only the copy engine overlaps; the
        two small packing kernels of a piece run between two fused launches).  Same results as set_samples(sigma);
        evaluate() -- every copy still happens inside the call.

        chunks = None (default): ONE library call, nq_logpsi_grad_local_host -- the pipeline runs inside libnqcuda on a growing
        schedule in units of one ROUND of the persistent kernel (one configuration per resident warp): 2 rounds, 7 rounds,
        the rest.  Only the copy of the first piece is exposed, a piece copies ~3.6x faster than it computes so every later
        copy hides behind the piece before it, and whole rounds keep the total number of rounds of the launch.
        chunks = k: k equal pieces, pipelined from this mirror with torch streams (kept for comparison)."""
        from .core import Context
        torch = _torch()
        net, ctx, Ns, P = self.net, self.ctx, self.Ns, self.net.P
        if self.symm:
            self.set_samples(sigma)
            return self.evaluate()
        if chunks is None:
            sr, sc = sigma if net.doubled else (sigma, None)
            hs = [np.asfortranarray(a).reshape(net.N, Ns, order="F") for a in (sr, sc) if a is not None]
            if hs[0].dtype not in (np.float32, np.float64) or any(h.dtype != hs[0].dtype for h in hs):
                hs = [np.asfortranarray(h, dtype=np.float64) for h in hs]
            self._center_pending = False
            gl = self.gloc.data_ptr() if self.is_liouvillian else None
            L.check(L.lib.nq_logpsi_grad_local_host(net.h, self.op.h, hs[0].ctypes.data, hs[1].ctypes.data if len(hs) > 1 else None,
                                                    L.nq_dtype(hs[0].dtype), Ns, self.prow.data_ptr(),
                                                    self.pcol.data_ptr() if self.pcol is not None else None,
                                                    self.logpsi.data_ptr(), self._O.data_ptr(), P, self.loc.data_ptr(), gl, P), ctx.h)
            return
        bounds = [0, Ns] if (chunks <= 1 or Ns < 8192 * chunks) else [(Ns * c) // chunks for c in range(chunks + 1)]
        if len(bounds) == 2:                   # measured: 8192 samples per GPU are faster in one piece
            self.set_samples(sigma)
            return self.evaluate()
        dev = torch.device("cuda", ctx.device)
        sr, sc = sigma if net.doubled else (sigma, None)
        hosts = [torch.from_numpy(np.asfortranarray(a).reshape(net.N, Ns, order="F").T) for a in (sr, sc) if a is not None]
        if getattr(self, "_side", None) is None or self._side[2].dtype != hosts[0].dtype:
            st = torch.cuda.Stream(device=dev)
            self._side = (st, Context(ctx.device, st.cuda_stream), torch.empty((2, Ns, net.N), dtype=hosts[0].dtype, device=dev))
        side_stream, side_ctx, stage = self._side
        main_stream = ctx.torch_stream()
        bufs = [self.prow, self.pcol][:len(hosts)]
        W = self.prow.shape[-1]
        es, cs = self._O.element_size(), self.loc.element_size()
        self._center_pending = False
        fcode = L.nq_dtype(np.dtype(str(hosts[0].dtype).replace("torch.", "")))
        side_stream.wait_stream(main_stream)                       # the staging / packed buffers may still be read by earlier work
        for c in range(len(bounds) - 1):
            c0, n = bounds[c], bounds[c + 1] - bounds[c]
            with torch.cuda.stream(side_stream):
                for i, h in enumerate(hosts):
                    stage[i, c0:c0 + n].copy_(h[c0:c0 + n], non_blocking=True)
            for i, buf in enumerate(bufs):
                L.check(L.lib.nq_pack_states(side_ctx.h, net.hilb.code, net.N, n, stage[i, c0:c0 + n].data_ptr(), fcode,
                                             buf.data_ptr() + c0 * W * 8), side_ctx.h)
            ev = torch.cuda.Event()
            ev.record(side_stream)
            main_stream.wait_event(ev)
            pc = self.pcol.data_ptr() + c0 * W * 8 if self.pcol is not None else None
            gl = self.gloc.data_ptr() + c0 * P * cs if self.is_liouvillian else None
            L.check(L.lib.nq_logpsi_grad_local_packed(net.h, self.op.h, self.prow.data_ptr() + c0 * W * 8, pc, n,
                                                      self.logpsi.data_ptr() + c0 * es, self._O.data_ptr() + c0 * P * es, P,
                                                      self.loc.data_ptr() + c0 * cs, gl, P), ctx.h)

    def assemble(self):
        """centre O, force vector, SR setup (S, F)."""
        net, ctx, Ns, P = self.net, self.ctx, self.Ns, self.net.P
        oc = L.nq_dtype(net.out_dtype)
        cc = L.nq_dtype(net.cdtype)
        ctx.set_global_samples(self.Ns_total if self.nranks > 1 else 0)
        if self.is_liouvillian and self.S is not None:
            # nothing but the S assembly reads the centred rows in this iteration: the subtraction may be deferred
            flag = C.c_int(0)
            L.check(L.lib.nq_center_lazy(ctx.h, self._O.data_ptr(), P, P, Ns, oc, self.avg.data_ptr(), C.byref(flag)), ctx.h)
            self._center_pending = bool(flag.value)
        else:
            L.check(L.lib.nq_center(ctx.h, self._O.data_ptr(), P, P, Ns, oc, self.avg.data_ptr()), ctx.h)
        if self.is_liouvillian:
            cost = C.c_double()
            L.check(L.lib.nq_force_liouvillian(ctx.h, self.loc.data_ptr(), self.gloc.data_ptr(), P, P, Ns, cc,
                                               self._avg_complex(), self.gradC.data_ptr(), C.byref(cost)), ctx.h)
            self.cost = cost.value
        else:
            L.check(L.lib.nq_force_ket(ctx.h, self.O.data_ptr(), P, P, Ns, oc, self.loc.data_ptr(),
                                       self.gradC.data_ptr()), ctx.h)
        if self.S is not None:
            if self.row_planes is not None:
                L.check(L.lib.nq_sr_hint_row_planes(ctx.h, L.ptr(self.row_planes), P), ctx.h)
            L.check(L.lib.nq_sr_setup(ctx.h, self._O.data_ptr(), P, P, Ns, self.Ns_total, oc, self.gradC.data_ptr(),
                                      int(self.real_params), self.S.data_ptr(), self.F.data_ptr()), ctx.h)
            if self.nranks > 1:      # C4 (global mean, quirk Q5)
                ctx.allreduce_sum(self.S.data_ptr(), P * P, L.nq_dtype(self.sdtype))
        else:
            torch = _torch()
            with torch.cuda.stream(ctx.torch_stream()):
                self.F.copy_(self.gradC.real if self.real_params else self.gradC)

    def _avg_complex(self):
        if self.avg.dtype == self.gradC.dtype:
            return self.avg.data_ptr()
        with _torch().cuda.stream(self.ctx.torch_stream()):
            self._avgc = self.avg.to(self.gradC.dtype)
        return self._avgc.data_ptr()

    def statistics(self):
        net, ctx = self.net, self.ctx
        if self.is_liouvillian:
            L.check(L.lib.nq_abs2(ctx.h, self.loc.data_ptr(), self.Ns, L.nq_dtype(net.cdtype), self.abs2.data_ptr()),
                    ctx.h)
            return stat_analysis(ctx, (self.abs2.data_ptr(), self.B, self.L, net.rdtype))
        return stat_analysis(ctx, (self.loc.data_ptr(), self.B, self.L, net.cdtype))

    # ---- observables on the stored samples (BatchedObsKetSampler.jl:32-59) ----
    def add_observable_(self, name, obs):
        """add_observable!(is, name, obs): `obs` is a LocalOperator on the physical Hilbert space."""
        if self.is_liouvillian:
            # the gradient sampler of a Liouvillian owns a diagonal-chain observables sampler (BatchedGradSampler.jl)
            if getattr(self, "_obs_dm", None) is None:
                self._obs_dm = BatchedObsDMSampler(self.net, self.sampler, batch_sz=self.B, chain_length=self.L,
                                                   chain_offset=self.chain_offset)
            self._obs_dm.add_observable_(name, obs)
            return
        if not hasattr(self, "observables"):
            self.observables, self._obs_dev, self._obs_loc = {}, {}, None
        self.observables[name] = obs
        self._obs_dev[name] = obs.to_device(self.ctx)

    def compute_observables(self):
        """compute_observables(is): O_loc of every stored configuration (same kernel as E_loc, chain reuse) and its
        chain statistics, one Measurement per observable."""
        torch = _torch()
        if self.is_liouvillian:
            dm = getattr(self, "_obs_dm", None)
            return dm.compute_observables() if dm is not None else {}
        res = {}
        if not getattr(self, "observables", None):
            return res
        if self._obs_loc is None:
            self._obs_loc = torch.zeros(self.Ns, dtype=self.loc.dtype, device=self.loc.device)
        for name, dev in self._obs_dev.items():
            L.check(L.lib.nq_local_scalar_packed(self.net.h, dev.h, self.prow.data_ptr(), None, self.Ns,
                                                 self._obs_loc.data_ptr()), self.ctx.h)
            res[name] = stat_analysis(self.ctx, (self._obs_loc.data_ptr(), self.B, self.L, self.net.cdtype))
        return res

    def sample_(self, sample=True):
        """sample!(is) -> (Measurement of the cost, self as the preconditioner cache)."""
        if sample:
            self.sample_states()
        self.evaluate()
        stat = self.statistics()
        self.assemble()
        return stat, self

    def precondition_(self, iter_n=1):
        """precondition!(cache, algo, iter) -> dw (device tensor; .cpu().numpy() for the host copy).
        Iterative solvers follow SRIterative.jl:71-153: one run of the chosen solver (maxiter 10 P, tol = precision);
        if it did not converge, up to 5 MINRES-QLP runs warm-started from the current dw with the solver's default
        tolerance sqrt(eps(T)) (the restart call passes none), a warning each; still not converged -> dw = 0.
        (In the reference that ladder cannot run: its restart call has one positional argument too many, and sr_qlp
        itself reports every exit as converged -- quirk Q22 in oracle/minresqlp.py.  This is the ladder the code spells.)"""
        ctx, P, algo = self.ctx, self.net.P, self.algo
        its = C.c_int64()
        sd = L.nq_dtype(self.sdtype)
        solver = {sr_cholesky: L.NQ_SOLVE_CHOLESKY, sr_cg: L.NQ_SOLVE_CG, sr_minres: L.NQ_SOLVE_MINRES,
                  sr_qlp: L.NQ_SOLVE_QLP}[algo.algorithm]
        eps, lam = algo.regulariser(iter_n)
        if lam != 0.0:      # sr_multiplicative (explicit S only): S + lambda Diagonal(diag S)
            L.check(L.lib.nq_sr_scale_diagonal(ctx.h, self.S.data_ptr(), P, sd, lam), ctx.h)

        def run(solver_code, tol):
            if self.S is not None:
                return L.lib.nq_sr_solve(ctx.h, self.S.data_ptr(), self.F.data_ptr(), P, sd, eps, solver_code, tol, algo.maxiter,
                                         self.dw.data_ptr(), C.byref(its))
            return L.lib.nq_sr_solve_matfree_algo(ctx.h, self.O.data_ptr(), P, P, self.Ns, self.Ns_total,
                                                  L.nq_dtype(self.net.out_dtype), self.F.data_ptr(),
                                                  int(self.real_params), eps, solver_code, tol, algo.maxiter, self.dw.data_ptr(),
                                                  C.byref(its))
        st = run(solver, algo.sr_precision)
        self.last_iters, self.restarts, self.converged = its.value, 0, st == L.NQ_OK
        if st == L.NQ_ERR_NOT_CONVERGED:
            default_tol = float(np.sqrt(np.finfo(np.dtype(self.net.rdtype)).eps))
            while st == L.NQ_ERR_NOT_CONVERGED and self.restarts < 5:
                self.restarts += 1
                warnings.warn("minresqlp not converged. Additional %d iters for the %d time." % (10 * P, self.restarts))
                st = run(L.NQ_SOLVE_QLP_WARM, default_tol)
                self.last_iters += its.value
            self.converged = st == L.NQ_OK
            if st == L.NQ_ERR_NOT_CONVERGED:       # success = false; dw .= 0.0
                with _torch().cuda.stream(ctx.torch_stream()):
                    self.dw.zero_()
                return self.dw
        L.check(st, ctx.h)
        return self.dw

    def update_(self, opt, dw=None):
        dw = self.dw if dw is None else dw
        net = self.net
        with _torch().cuda.stream(self.ctx.torch_stream()):
            delta, eta = opt.step(dw, ctx=self.ctx)     # Descent: (dw, eta); Nesterov: velocity kernel on the device vectors
            if delta.dtype != _tdtype(net.dtype):
                delta = delta.to(_tdtype(net.dtype))
            delta = delta.contiguous()
        if self.symm:
            L.check(L.lib.nq_symm_update(net.g, delta.data_ptr(), float(eta)), self.ctx.h)
            return
        L.check(L.lib.nq_update(net.h, delta.data_ptr(), float(eta)), self.ctx.h)


class BatchedObsDMSampler:
    """Observables of a density-matrix machine (BatchedObsDMSampler.jl:3-105): an independent Markov chain over the
    diagonal, p(sigma) ~ rho(sigma, sigma), and per stored sigma the local estimator
        O_loc(sigma) = sum_eta O(sigma, eta) rho(eta, sigma) / rho(sigma, sigma)
    so that <O> = Tr(O rho) / Tr(rho) = mean O_loc.  Log-probability: Re log rho(sigma, sigma) (the reference writes
    abs.(log rho), SURVEY quirk Q4 -- not reproduced)."""

    def __init__(self, net, sampler, batch_sz=16, chain_length=None, chain_offset=0):
        torch = _torch()
        if not net.doubled or net.kind != L.NQ_NDM:
            raise ValueError("BatchedObsDMSampler needs an NDM density matrix")
        self.net, self.sampler, self.ctx = net, sampler, net.ctx
        self.B = int(batch_sz)
        # its own Philox key: with the main chain's seed both chains would consume identical (site, uniform) streams
        obs_sampler = copy.copy(sampler)
        obs_sampler.seed = (int(sampler.seed) ^ 0x6F62735F646D) % (1 << 63)          # "obs_dm"
        self.cache = MetropolisSamplerCache(obs_sampler, net, self.B, chain_offset=chain_offset)
        self.cache.set_mode(True)
        self.L = self.cache.loc_chain_length if chain_length is None else int(chain_length)
        self.Ns = self.B * self.L
        dev = torch.device("cuda", self.ctx.device)
        W = L.lib.nq_states_words(net.N)
        self.prow = torch.zeros((self.L, self.B, W), dtype=torch.int64, device=dev)
        self.pcol = torch.zeros_like(self.prow)
        self.loc = torch.zeros(self.Ns, dtype=_tdtype(net.cdtype), device=dev)
        self.observables, self._dev, self.results = {}, {}, {}

    def add_observable_(self, name, obs):
        self.observables[name] = obs
        self._dev[name] = obs.to_device_left(self.ctx)

    def sample_states(self):
        self.cache.randomize()
        self.cache.sample(self.sampler.burn_length, self.L, packed_out=(self.prow.data_ptr(), self.pcol.data_ptr()))

    def set_samples(self, sigma):
        """Bypass the chain with supplied diagonal configurations [N, B, L] (parity tests)."""
        a = np.asfortranarray(sigma).reshape(self.net.N, self.Ns, order="F")
        for buf in (self.prow, self.pcol):
            L.check(L.lib.nq_pack_states(self.ctx.h, self.net.hilb.code, self.net.N, self.Ns, L.ptr(a), L.nq_dtype(a.dtype),
                                         buf.data_ptr()), self.ctx.h)

    def local_values(self, name):
        L.check(L.lib.nq_local_scalar_packed(self.net.h, self._dev[name].h, self.prow.data_ptr(), self.pcol.data_ptr(),
                                             self.Ns, self.loc.data_ptr()), self.ctx.h)
        return self.loc

    def compute_observables(self, sample=True):
        if not self.observables:
            return None
        if sample:
            self.sample_states()
        for name in self.observables:
            self.local_values(name)
            self.results[name] = stat_analysis(self.ctx, (self.loc.data_ptr(), self.B, self.L, self.net.cdtype))
        return self.results
