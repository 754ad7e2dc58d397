"""Per-phase timings of BASELINE cfg1 / cfg2 / cfg3 on one B200 -> profiles/r2_cfg{1,2,3}.json.

cfg1: TFIM 1D N=10, RBM alpha=2 logcosh complex128, B=8 chains x L=125, passes 11, SR eps=0.1 CG tol 1e-3 (examples/ising1d.jl)
cfg2: Lindblad Ising N=8, NDM alpha=2 softplus, B=16 x L=125, passes 9, SR eps=1e-3 Cholesky (examples/dissipative_ising1d.jl)
cfg3: TFIM 2D 6x6 h=3, RBM alpha=4 logcosh, 16384 chains x L=1, passes 37, FP64 (complex128) and FP32 (complex64) modes

usage: python profiles/run_cfgs.py [out_dir] [tag]     (CUDA events on the context stream, 5 repetitions after a warm-up)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import nqcuda as nq  # noqa: E402
from bench import ClockSampler, FP64_TENSOR_PEAK, peaks  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")
tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
pk, pk_kind = peaks()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(name, net, smp, problem, algo, B, L, passes_note):
    bs = nq.BatchedSampler(net, smp, problem, algo, batch_sz=B, chain_length=L)
    Ns, P = B * L, net.P
    es = net.out_dtype.itemsize
    opt = nq.Descent(0.01)

    def it():
        bs.sample_()
        bs.precondition_()
        bs.update_(opt)
    it()
    clocks = ClockSampler(0)
    ph = {"sampler": timed(bs.sample_states), "evalgrad+estimator": timed(bs.evaluate), "statistics": timed(bs.statistics)}

    def asm_solve():          # the factorisation overwrites S: time assembly + solve together, then the assembly alone
        bs.assemble()
        bs.precondition_()
    t_both = timed(asm_solve, 3)
    ph["centre+force+S"] = timed(bs.assemble, 3)
    ph["solve"] = t_both - ph["centre+force+S"]
    bs.sample_()
    t_iter = timed(it)
    clk = clocks.stop()
    rows = 2 if bs.is_liouvillian else 1          # O row (+ grad L_loc row) written per sample
    bytes_eval = Ns * (rows * P * es + 2 * es)
    cplx_o = np.dtype(net.out_dtype).kind == "c"
    # S assembly, dense count: real-parameter nets = real SYRK over K = 2 Ns (2 P^2 Ns flops), complex nets = HERK 4 P^2 Ns
    flops_S = (2.0 if bs.real_params else 4.0) * P * P * Ns * (1.0 if cplx_o else 0.5)
    rec = {"config": name, "P": P, "Ns": Ns, "chains": B, "stored_per_chain": L, "passes": passes_note,
           "dtype": str(np.dtype(net.dtype)), "phases_ms": ph, "sr_iteration_ms": t_iter,
           "samples_per_s": {"evalgrad+estimator": Ns / (ph["evalgrad+estimator"] * 1e-3), "full_iteration": Ns / (t_iter * 1e-3)},
           "roofline": {"evalgrad+estimator": {"bound": "hbm", "GB/s": bytes_eval / (ph["evalgrad+estimator"] * 1e-3) / 1e9,
                                                "frac": bytes_eval / (ph["evalgrad+estimator"] * 1e-3) / 1e9 / pk["hbm_gbs"], "peak": pk["hbm_gbs"],
                                                "note": "launch-latency bound when Ns P is small (cfg1, cfg2: a few MB per launch)"},
                        "S assembly (centre+force+S)": {"bound": "tensor (FP64 DMMA)" if net.rdtype == np.float64 else "tensor (tcgen05 3xTF32)",
                                                        "dense_TFLOP/s": flops_S / (ph["centre+force+S"] * 1e-3) / 1e12,
                                                        "frac_of_fp64_dmma_peak": flops_S / (ph["centre+force+S"] * 1e-3) / 1e12 / FP64_TENSOR_PEAK,
                                                        "peak": FP64_TENSOR_PEAK}},
           "clocks": clk, "peak_source": pk_kind, "solver_iterations": bs.last_iters}
    del bs
    return rec


def main():
    hilb, Hm = nq.models.tfim_1d(10, 1.0, 1.0)
    net = nq.RBM(ctx, hilb, np.complex128, 2, nq.af_logcosh)
    nq.init_random_pars_(net, 0.01, 1234)
    r1 = run("cfg1: TFIM 1D N=10, RBM alpha=2 logcosh c128, 8 chains x 125, SR eps=0.1 CG tol=1e-3", net,
             nq.MetropolisSampler(nq.LocalRule(), 125, 10, burn=100, seed=5), Hm,
             nq.SR(np.float32, eps=0.1, algorithm="sr_cg", precision=1e-3), 8, 125, 11)
    json.dump(r1, open(os.path.join(out_dir, tag + "_cfg1.json"), "w"), indent=1)
    hilb, _, _, liouv = nq.models.lindblad_ising_1d(8, 0.4, 2.0)
    net = nq.NDM(ctx, hilb, np.float64, 2, 2, nq.af_softplus, seed=1234)
    r2 = run("cfg2: Lindblad Ising N=8, NDM alpha=2 softplus f64, 16 chains x 125, SR eps=1e-3 Cholesky", net,
             nq.MetropolisSampler(nq.LocalRule(), 125, 8, burn=100, seed=5), liouv,
             nq.SR(np.float32, eps=0.001, algorithm="sr_cholesky"), 16, 125, 9)
    json.dump(r2, open(os.path.join(out_dir, tag + "_cfg2.json"), "w"), indent=1)
    hilb, Hm = nq.models.tfim_2d(6, 3.0, 1.0)
    r3 = {}
    for mode, dt in (("fp64 (complex128)", np.complex128), ("fp32 (complex64)", np.complex64)):
        net = nq.RBM(ctx, hilb, dt, 4, nq.af_logcosh)
        nq.init_random_pars_(net, 0.01, 1234)
        r3[mode] = run("cfg3: TFIM 2D 6x6 h=3, RBM alpha=4 logcosh, 16384 chains x 1, SR eps=0.1 Cholesky, " + mode, net,
                       nq.MetropolisSampler(nq.LocalRule(), 1, 36, burn=100, seed=5), Hm,
                       nq.SR(np.float32, eps=0.1, algorithm="sr_cholesky"), 16384, 1, 37)
        del net
    json.dump(r3, open(os.path.join(out_dir, tag + "_cfg3.json"), "w"), indent=1)
    print(json.dumps({"cfg1": r1["phases_ms"], "cfg2": r2["phases_ms"], "cfg3": {k: v["phases_ms"] for k, v in r3.items()}}, indent=1))


if __name__ == "__main__":
    main()
