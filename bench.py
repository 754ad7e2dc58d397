#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for libnqcuda.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
    dissipative Ising 1D, N=16, g=0.4, V=2, gamma=1 (L_i = sigma^-_i), NDM alpha_h=alpha_a=2 (softplus),
    Float64, 65 536 samples per iteration = 4096 chains x 16 stored samples, sharded over the GPUs
    (strong scaling: the global sample count is fixed), SR eps=0.001 Cholesky, Descent(0.01).

Metric: MC samples/s of {log rho + gradient rows O, local Liouvillian estimator L_loc + grad L_loc}
("eval+grad+local estimator"), inputs resident in HBM; `sr_iteration` reports seconds per full SR
iteration (sampling + eval/grad + estimator + centring/force + S assembly + all-reduce + solve + update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
For N>1 the driver launches it under torchrun; one rank per GPU, NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOAD = dict(N=16, alpha=2, g=0.4, V=2.0, chains=4096, L=16, passes=17, burn=100, eps=0.001, eta=0.01)
WORKLOAD_NAME = ("cfg4: dissipative Ising 1D N=16 steady state, NDM alpha=2 (softplus, Float64), "
                 "65536 samples/iter = 4096 chains x 16, SR eps=1e-3 Cholesky")
METRIC = "MC samples/s (eval+grad+local estimator)"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on the host cores (the one place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------
_CPU_STATE = {}


def _cpu_setup():
    """The reference CPU path = oracle/cref.c (C restatement of the Julia algorithm, OpenMP over samples):
    full forward pass + full gradient for every connected configuration, like AccumulatorObsGrad."""
    if not _CPU_STATE:
        from oracle import cref, machines as OM
        from oracle.models import lindblad_ising_1d
        w = WORKLOAD
        hilb, _, _, liouv = lindblad_ising_1d(w["N"], w["g"], w["V"])
        net = OM.random_machine("ndm", w["N"], w["alpha"], seed=1234, std=0.1)
        _CPU_STATE.update(cref=cref, hilb=hilb, liouv=liouv, net=net, tables=cref.flatten(liouv))
    return _CPU_STATE


def cpu_samples_per_s(n_samples, threads):
    """eval+grad+local estimator of n_samples configurations of the workload on `threads` host threads."""
    from oracle.models import random_states
    st = _cpu_setup()
    cref, net = st["cref"], st["net"]
    sr_, sc_ = random_states(st["hilb"], n_samples, 100), random_states(st["hilb"], n_samples, 101)
    t0 = time.perf_counter()
    cref.ndm_logpsi_grad(net, sr_, sc_, nthreads=threads)
    cref.local_grad_super(net, st["liouv"], sr_, sc_, nthreads=threads, tables=st["tables"])
    dt = time.perf_counter() - t0
    return n_samples / dt, n_samples


def _cpu_sample_size(threads, seconds):
    cpu_samples_per_s(16 * threads, threads)                 # load the library, start the OpenMP team
    v, _ = cpu_samples_per_s(128 * threads, threads)         # calibration
    n = int(v * seconds)
    return max(64 * threads, min(WORKLOAD["chains"] * WORKLOAD["L"], (n // 1024) * 1024 or 1024))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = _cpu_sample_size(threads, 4.0)
    for _ in range(args.warmup):
        cpu_samples_per_s(max(1024, n // 8), threads)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, used = cpu_samples_per_s(n, threads)
        vals.append(v)
    dt = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = ("%d configurations per step on %d OpenMP threads; C restatement of the reference algorithm "
              "(oracle/cref.c; Julia is not installed, so the Julia code itself cannot be timed)" % (used, threads))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def ncu_fused_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused kernel on the cfg4 batch, taken from the
    committed `ncu --set full` capture (profiles/ncu_traffic.py writes the file); None when the file is absent."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2m_fused_kernel_traffic.json")))
        return float(rec["dram_bytes"]), rec.get("report")
    except Exception:
        return None, None


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its version there on
    rank 0), so fd 1 is pointed at stderr for the whole run and the line goes to a private copy of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


FP64_TENSOR_PEAK = 37.2   # TFLOP/s, mma.sync m8n8k4 f64 at 8 warps/SM on this pool's B200 (profiles/fp64_peaks.json)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        # nvidia-smi needs a moment to start: wait for its first line so that short timed regions are covered
        t0 = time.time()
        while self.p is not None and time.time() - t0 < 3.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        out["sm_mhz"] = float(np.median(sm)) if sm else None
        out["sm_max_mhz"] = float(rows[0][2]) if rows[0][2].strip().replace(".", "").isdigit() else None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        out["reasons"] = sorted(reasons)
        out["power_w_max"] = max(float(r[3]) for r in rows if r[3].strip().replace(".", "").isdigit())
        return out


def multi_gpu_parity(nq, torch, dist, ctx, net, liouv, world, rank, local, n_sub=None):
    """Cross-rank check of the exchange steps on a subset of the workload with 4096 samples PER RANK (so that every rank runs
    the production path of the iteration: deferred centring + the integer-tensor-core S assembly): the sharded iteration
    (all-reduced <O>, F, S; replicated solve) against a SINGLE-RANK recomputation of the same samples on rank 0 (second
    context without a communicator), and bit-equality of the update across ranks.  ref: Parallel/MPI/mpi.jl:21-74."""
    w = WORKLOAD
    N, Lc = w["N"], w["L"]
    n_sub = 4096 * world if n_sub is None else n_sub
    Bt = n_sub // Lc
    Bt -= Bt % world
    rng = np.random.Generator(np.random.Philox(777))          # same subset on every rank
    R = rng.integers(0, 2, size=(N, Bt, Lc)).astype(np.float64)
    Cc = rng.integers(0, 2, size=(N, Bt, Lc)).astype(np.float64)
    B = Bt // world
    sl = slice(rank * B, (rank + 1) * B)
    algo = nq.SR(np.float32, eps=w["eps"], algorithm="sr_cholesky")
    smp = nq.MetropolisSampler(nq.LocalRule(), Lc * world, w["passes"] - 1, burn=1, seed=5)
    bs = nq.BatchedSampler(net, smp, liouv, algo, batch_sz=B, chain_length=Lc)
    bs.set_samples((np.asfortranarray(R[:, sl]), np.asfortranarray(Cc[:, sl])))
    stat, _ = bs.sample_(sample=False)
    S_sh, F_sh = bs.S.clone(), bs.F.clone()
    dw_sh = bs.precondition_().clone()
    lo, hi = dw_sh.clone(), dw_sh.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = {"samples": int(Bt * Lc), "dw_bit_identical_across_ranks": bool(torch.equal(lo, hi))}
    if rank == 0:
        ctx1 = nq.Context(local, torch.cuda.current_stream().cuda_stream)     # no communicator: NotParallel
        net1 = nq.NDM(ctx1, net.hilb, np.float64, w["alpha"], w["alpha"], nq.af_softplus, seed=1234)
        net1.set_params(net.params())
        b1 = nq.BatchedSampler(net1, smp, liouv, algo, batch_sz=Bt, chain_length=Lc)
        b1.set_samples((np.asfortranarray(R), np.asfortranarray(Cc)))
        stat1, _ = b1.sample_(sample=False)

        def rel(a, b):
            return float((torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1))).item())
        out["S_rel"], out["F_rel"] = rel(S_sh, b1.S), rel(F_sh, b1.F)
        out["dw_rel"] = rel(dw_sh, b1.precondition_())
        out["cost_rel"] = abs(stat.mean - stat1.mean) / abs(stat1.mean)
        out["stat_error_rel"] = abs(stat.error - stat1.error) / abs(stat1.error)
        out["max_rel"] = max(out["S_rel"], out["F_rel"], out["dw_rel"], out["cost_rel"], out["stat_error_rel"])
        out["tolerance"] = "1e-11 on S, F, cost; dw up to the conditioning of S + eps I"
    dist.barrier()
    return out


def run_gpu(args):
    # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    import nqcuda as nq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched under torchrun with %d ranks" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = nq.Context(local, torch.cuda.current_stream().cuda_stream)
    if world > 1:
        box = [nq.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(world, rank, box[0])

    w = WORKLOAD
    N, Lc = w["N"], w["L"]
    B = w["chains"] // world                       # chains of this rank (strong scaling)
    Ns = B * Lc
    Ns_global = Ns * world
    hilb, _, _, liouv = nq.models.lindblad_ising_1d(N, w["g"], w["V"])
    net = nq.NDM(ctx, hilb, np.float64, w["alpha"], w["alpha"], nq.af_softplus, seed=1234)
    nq.init_random_pars_(net, sigma=0.01, seed=1234)
    smp = nq.MetropolisSampler(nq.LocalRule(), Lc * world, w["passes"] - 1, burn=w["burn"], seed=99)
    bs = nq.BatchedSampler(net, smp, liouv, nq.SR(np.float32, eps=w["eps"], algorithm="sr_cholesky"), batch_sz=B,
                           chain_length=Lc)
    P = net.P

    # synthetic configurations (uniform, seed 4321 + rank), resident on the device before timing
    rng = np.random.Generator(np.random.Philox(4321 + rank))
    host_r = torch.empty((Ns, N), dtype=torch.float64).pin_memory()
    host_c = torch.empty((Ns, N), dtype=torch.float64).pin_memory()
    host_r.numpy()[:] = rng.integers(0, 2, size=(Ns, N))
    host_c.numpy()[:] = rng.integers(0, 2, size=(Ns, N))
    host_loc = torch.empty(Ns, dtype=torch.complex128).pin_memory()
    sig = (host_r.numpy().T.reshape(N, B, Lc, order="F"), host_c.numpy().T.reshape(N, B, Lc, order="F"))
    bs.set_samples(sig)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time (CUDA events), max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- headline: eval + grad + local estimator on resident configurations ----
    for _ in range(max(3, args.warmup)):
        bs.evaluate()
    clocks = ClockSampler(local)
    l0 = ctx.launches
    ms = timed(bs.evaluate, args.steps)
    launches = ctx.launches - l0
    value = Ns_global * args.steps / (ms * 1e-3)

    # ---- per-kernel durations inside the same loop (events around each launch) ----
    pc = bs.pcol.data_ptr()
    L_ = nq._lib

    def k_evalgrad():     # stand-alone fused eval+grad (the machine plugin call nq_logpsi_grad)
        L_.check(L_.lib.nq_logpsi_grad_packed(net.h, bs.prow.data_ptr(), pc, Ns, bs.logpsi.data_ptr(), bs.O.data_ptr(), P), ctx.h)

    def k_fused():        # the step's kernel: log rho + O + L_loc + grad L_loc in one launch
        L_.check(L_.lib.nq_logpsi_grad_local_packed(net.h, bs.op.h, bs.prow.data_ptr(), pc, Ns, bs.logpsi.data_ptr(),
                                                    bs.O.data_ptr(), P, bs.loc.data_ptr(), bs.gloc.data_ptr(), P), ctx.h)
    for _ in range(3):
        k_evalgrad(); k_fused()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        ev[s][0].record(); k_evalgrad(); ev[s][1].record(); k_fused(); ev[s][2].record()
    barrier()
    t_eval = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    t_loc = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    # ---- full SR iteration ----
    opt = nq.Descent(w["eta"])

    def sr_iter():
        bs.sample_()
        bs.precondition_()
        bs.update_(opt)
    phases = {}

    def phase(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        phases[name] = phases.get(name, 0.0) + a.elapsed_time(b)
    sr_iter()
    n_it = max(2, min(args.steps, 5))
    ms_sr = timed(sr_iter, n_it)
    for _ in range(2):
        phase("sampler", bs.sample_states)
        phase("evalgrad+estimator", bs.evaluate)
        phase("statistics", bs.statistics)
        phase("centre+force+S", bs.assemble)
        phase("solve", bs.precondition_)
        phase("update", lambda: bs.update_(opt))
    phases = {k: v / 2 for k, v in phases.items()}
    # ---- S assembly alone (nq_sr_setup on the centred rows of the last iteration): tensor-pipe roofline ----
    oc = L_.nq_dtype(net.out_dtype)

    def k_setup():
        L_.check(L_.lib.nq_sr_setup(ctx.h, bs.O.data_ptr(), P, P, Ns, bs.Ns_total, oc, bs.gradC.data_ptr(),
                                    int(bs.real_params), bs.S.data_ptr(), bs.F.data_ptr()), ctx.h)
    k_setup()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a_.record()
    for _ in range(3):
        k_setup()
    b_.record()
    torch.cuda.synchronize()
    t_setup = a_.elapsed_time(b_) / 3
    # DMMA work issued: (tile pair, component) units whose planes are not identically zero, 128 x 128 x Ns MACs each
    Ov = torch.view_as_real(bs.O) if bs.O.is_complex() else bs.O.unsqueeze(-1)
    nzp = (Ov != 0).any(dim=0)                                   # [P, NC]
    ntile = (P + 127) // 128
    fl = [[bool(nzp[t * 128:(t + 1) * 128, c].any()) for c in range(nzp.shape[1])] for t in range(ntile)]
    units = sum(sum(1 for c in range(len(fl[0])) if fl[i][c] and fl[j][c]) for i in range(ntile) for j in range(i + 1))
    flops_setup = units * 128.0 * 128.0 * Ns * 2.0
    # restore the synthetic configurations (the SR iterations above sampled new ones)
    bs.set_samples(sig)

    # ---- e2e: public API with HOST buffers (pinned), H2D of the configurations and D2H of L_loc inside ----
    def e2e_step():
        bs.evaluate_host(sig)          # set_samples + evaluate, copies of piece c+1 overlapped with the kernel on piece c
        host_loc.copy_(bs.loc, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = Ns_global * args.steps / (ms_e2e * 1e-3)
    clk = clocks.stop()      # sampled over every timed region above (headline, per-kernel, SR iteration, e2e)
    parity = multi_gpu_parity(nq, torch, dist, ctx, net, liouv, world, rank, local) if world > 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_kind = peaks()
    es = 16
    bytes_eval = Ns * (P * es + es + 2 * 8)               # O row + log rho + two packed words per configuration
    bytes_fused = Ns * (2 * P * es + 2 * es + 2 * 8)      # O row + grad L_loc row + log rho + L_loc + packed words
    ach = bytes_fused / (t_loc * 1e-3) / 1e9
    traffic, traffic_src = ncu_fused_traffic()
    roof = {"bound": "hbm", "kernel": "local_ndm5_kernel<double,softplus,grad,with_O> (fused eval+grad+estimator)",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this batch (ncu --set full)
            "traffic": traffic if (world == 1 and Ns == 65536) else None, "traffic_source": traffic_src,
            "peak_source": pk_kind, "ms_per_launch": t_loc, "algorithmic_bytes_per_launch": bytes_fused,
            "note": "write-bound: DRAM traffic = algorithmic bytes; remaining gap is instruction issue, DESIGN.md section 4",
            "all": {"ndm_evalgrad_kernel (stand-alone nq_logpsi_grad)": {
                        "ms": t_eval, "GB/s": bytes_eval / (t_eval * 1e-3) / 1e9,
                        "frac": bytes_eval / (t_eval * 1e-3) / 1e9 / pk["hbm_gbs"]},
                    "local_ndm5_kernel (fused)": {"ms": t_loc, "GB/s": ach, "frac": ach / pk["hbm_gbs"]},
                    "S assembly (nq_sr_setup: Ozaki scheme on tcgen05 kind::i8, digit pre-pass + syrk_ozaki2_kernel + finalize)": {
                        "bound": "tensor", "ms": t_setup, "achieved": flops_setup / (t_setup * 1e-3) / 1e12, "unit": "TFLOP/s (FP64-equivalent)",
                        "peak": FP64_TENSOR_PEAK, "frac": flops_setup / (t_setup * 1e-3) / 1e12 / FP64_TENSOR_PEAK,
                        "peak_source": "FP64 DMMA issue rate measured with profiles/probe/dmma_probe.cu (cuBLAS DGEMM: 35.5); "
                                       "frac > 1 = faster than the FP64 tensor instruction allows: the products run as 34 exact int8 "
                                       "MMAs (ncu: tensor pipe 54 % active, bound by L2 -> SM operand traffic)",
                        "flops": "non-zero (tile pair, component) units x 128 x 128 x Ns x 2"}}}
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "chains_per_gpu": B, "stored_per_chain": Lc, "P": P,
                       "parallelism": "dp%d (chains sharded, NCCL all-reduce of <O>, F, S)" % world,
                       "l2": "every step writes O and grad L_loc (%.2f GB per GPU) >> 126 MB L2" % (2 * Ns * P * es / 1e9)},
            "sr_iteration": {"value": ms_sr * 1e-3 / n_it, "unit": "s", "iterations": n_it, "phases_ms": phases,
                             "includes": "sampler(burn=100,passes=17)+eval/grad+estimator+centre+force+S(+allreduce)+Cholesky+update"},
            "roofline": roof, "clocks": clk, "gpu_launches": int(launches), "parity": parity,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(2 * Ns * N * 8),
                    "d2h_bytes_per_step": int(Ns * 16), "ms_per_step": ms_e2e / args.steps}}
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n = _cpu_sample_size(threads, 12.0)
        v, used = cpu_samples_per_s(n, threads)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                "sample": "%d configurations of the same workload on %d OpenMP threads; C restatement of "
                                          "the reference algorithm (oracle/cref.c; Julia unavailable)" % (used, threads)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nqcuda", choices=["nqcuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
