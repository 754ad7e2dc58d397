"""BASELINE cfg5: throughput sweep -- RBM / NDM, N = 64 ... 256, alpha = 1 ... 8, batches 4k ... 256k samples,
eval+grad (O materialised) and SR assembly only, random states and N(0, 0.01) parameters.  Points whose O or S do
not fit the budget below are reported as infeasible with their size.  usage: python profiles/run_cfg5.py [quick]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import nqcuda as nq  # noqa: E402

L = nq._lib
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
quick = len(sys.argv) > 1
HBM = 6392.8
O_BUDGET, S_BUDGET = 60e9, 20e9
grid_N = (64, 256) if quick else (64, 128, 256)
grid_a = (1, 8) if quick else (1, 2, 4, 8)
grid_Ns = (4096, 65536) if quick else (4096, 16384, 65536, 262144)
rows = []


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for kind in ("rbm", "ndm"):
    for N in grid_N:
        hilb = nq.HomogeneousSpin(N) if kind == "rbm" else nq.HomogeneousFock(N, 2)
        for alpha in grid_a:
            if kind == "rbm":
                net = nq.RBM(ctx, hilb, np.complex128, alpha, nq.af_logcosh)
            else:
                net = nq.NDM(ctx, hilb, np.float64, alpha, alpha, nq.af_softplus)
            nq.init_random_pars_(net, sigma=0.01, seed=1234)
            P, es = net.P, 16
            W = L.lib.nq_states_words(N)
            for Ns in grid_Ns:
                row = {"machine": kind, "N": N, "alpha": alpha, "P": P, "Ns": Ns}
                ob = P * Ns * es
                if ob > O_BUDGET:
                    row["infeasible"] = "O = %.1f GB" % (ob / 1e9)
                    rows.append(row)
                    continue
                g = torch.Generator(device="cuda").manual_seed(4321)
                mask = (1 << min(64, N)) - 1 if N < 64 else -1
                prow = torch.randint(-2**63, 2**63 - 1, (Ns, W), dtype=torch.int64, device="cuda", generator=g)
                pcol = torch.randint(-2**63, 2**63 - 1, (Ns, W), dtype=torch.int64, device="cuda", generator=g)
                if N % 64:
                    prow[:, -1] &= (1 << (N % 64)) - 1
                    pcol[:, -1] &= (1 << (N % 64)) - 1
                out = torch.zeros(Ns, dtype=torch.complex128, device="cuda")
                O = torch.zeros((Ns, P), dtype=torch.complex128, device="cuda")
                pc = pcol.data_ptr() if kind == "ndm" else None
                ms = timeit(lambda: L.check(L.lib.nq_logpsi_grad_packed(net.h, prow.data_ptr(), pc, Ns, out.data_ptr(),
                                                                        O.data_ptr(), P), ctx.h))
                row["evalgrad_ms"] = ms
                row["evalgrad_GBs"] = (ob + Ns * (es + 8 * W * (2 if kind == "ndm" else 1))) / ms / 1e6
                row["evalgrad_frac_hbm"] = row["evalgrad_GBs"] / HBM
                row["samples_per_s"] = Ns / ms * 1e3
                real_params = kind == "ndm"
                sb = P * P * (8 if real_params else 16)
                if sb > S_BUDGET:
                    row["sr_assembly"] = "infeasible: S = %.1f GB (matrix-free solve only)" % (sb / 1e9)
                else:
                    avg = torch.zeros(P, dtype=torch.complex128, device="cuda")
                    L.check(L.lib.nq_center(ctx.h, O.data_ptr(), P, P, Ns, L.NQ_C128, avg.data_ptr()), ctx.h)
                    S = torch.zeros((P, P), dtype=torch.float64 if real_params else torch.complex128, device="cuda")
                    F = torch.zeros(P, dtype=S.dtype, device="cuda")
                    gc = np.ones(P, np.complex128)
                    ms2 = timeit(lambda: L.check(L.lib.nq_sr_setup(ctx.h, O.data_ptr(), P, P, Ns, Ns, L.NQ_C128, L.ptr(gc),
                                                                   int(real_params), S.data_ptr(), F.data_ptr()), ctx.h), reps=1)
                    flops = (2.0 if real_params else 4.0) * P * P * Ns
                    row["sr_assembly_ms"] = ms2
                    row["sr_assembly_TFLOPs_dense"] = flops / ms2 / 1e9
                    del S, F, avg
                rows.append(row)
                print(json.dumps(row), flush=True)
                del O, out, prow, pcol
                torch.cuda.empty_cache()
            del net
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "cfg5_sweep.json"), "w"), indent=1)
