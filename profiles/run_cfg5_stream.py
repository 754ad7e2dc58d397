"""BASELINE cfg5, the points the one-shot sweep (run_cfg5.py) reported infeasible because the gradient matrix O exceeds the
memory budget: O is produced chunk by chunk into ONE reused buffer (nq_logpsi_grad_packed) and consumed at once by the
streaming S assembly (nq_sr_accumulate / nq_sr_finish), so the batch size is no longer bounded by |O|.

Per point: eval+grad GB/s over all chunks (HBM write roofline), S-assembly TFLOP/s (dense count) for S <= S_BUDGET, and the
total time of eval+grad+S.  Points whose S assembly is estimated above T_CAP seconds are run on eval+grad only and say so.
usage: python profiles/run_cfg5_stream.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import nqcuda as nq  # noqa: E402
from bench import peaks  # noqa: E402

L = nq._lib
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "cfg5_stream.json")
HBM = peaks()[0]["hbm_gbs"]
O_ONE_SHOT, S_BUDGET, CHUNK_BYTES, T_CAP, T_TOTAL = 60e9, 40e9, 8e9, 45.0, 420.0
rows, t_start = [], time.time()


def ev():
    return torch.cuda.Event(enable_timing=True)


for kind in ("rbm", "ndm"):
    for N in (64, 128, 256):
        hilb = nq.HomogeneousSpin(N) if kind == "rbm" else nq.HomogeneousFock(N, 2)
        for alpha in (1, 2, 4, 8):
            Pn = N + alpha * N + alpha * N * N if kind == "rbm" else N * (2 + 3 * alpha) + 4 * alpha * N * N
            todo = [Ns for Ns in (4096, 16384, 65536, 262144) if Pn * Ns * 16 > O_ONE_SHOT]
            if not todo:
                continue
            net = (nq.RBM(ctx, hilb, np.complex128, alpha, nq.af_logcosh) if kind == "rbm"
                   else nq.NDM(ctx, hilb, np.float64, alpha, alpha, nq.af_softplus))
            nq.init_random_pars_(net, sigma=0.01, seed=1234)
            P, es, W = net.P, 16, L.lib.nq_states_words(N)
            real_params = kind == "ndm"
            sb = P * P * (8 if real_params else 16)
            for Ns in todo:
                row = {"machine": kind, "N": N, "alpha": alpha, "P": P, "Ns": Ns, "O_GB_never_materialised": P * Ns * es / 1e9}
                Nc = max(1024, min(Ns, int(CHUNK_BYTES // (P * es)) // 1024 * 1024))
                flops = (2.0 if real_params else 4.0) * P * P * Ns
                do_S = sb <= S_BUDGET
                est = flops / 28e12 * (0.65 if real_params else 1.0)
                if do_S and (est > T_CAP or time.time() - t_start + est > T_TOTAL):
                    do_S = False
                    row["sr_assembly"] = "feasible (S = %.1f GB) but estimated %.0f s: skipped to bound GPU minutes" % (sb / 1e9, est)
                elif not do_S:
                    row["sr_assembly"] = "S = %.1f GB > budget: matrix-free solve territory" % (sb / 1e9)
                g = torch.Generator(device="cuda").manual_seed(4321)
                prow = torch.randint(-2**63, 2**63 - 1, (Ns, W), dtype=torch.int64, device="cuda", generator=g)
                pcol = torch.randint(-2**63, 2**63 - 1, (Ns, W), dtype=torch.int64, device="cuda", generator=g)
                if N % 64:
                    prow[:, -1] &= (1 << (N % 64)) - 1
                    pcol[:, -1] &= (1 << (N % 64)) - 1
                out = torch.zeros(Ns, dtype=torch.complex128, device="cuda")
                O = torch.zeros((Nc, P), dtype=torch.complex128, device="cuda")
                S = torch.zeros((P, P), dtype=torch.float64 if real_params else torch.complex128, device="cuda") if do_S else None
                sumO = torch.zeros(2 * P, dtype=torch.complex128, device="cuda")
                t_eval = t_S = 0.0
                for rep in range(2 if not do_S else 1):            # eval+grad only: one warm-up pass
                    t_eval = 0.0
                    for c0 in range(0, Ns, Nc):
                        n = min(Nc, Ns - c0)
                        a, b, c = ev(), ev(), ev()
                        a.record()
                        pc = pcol[c0:].data_ptr() if kind == "ndm" else None
                        L.check(L.lib.nq_logpsi_grad_packed(net.h, prow[c0:].data_ptr(), pc, n, out[c0:].data_ptr(), O.data_ptr(), P), ctx.h)
                        b.record()
                        if do_S:
                            L.check(L.lib.nq_sr_accumulate(ctx.h, O.data_ptr(), P, P, n, Ns, L.NQ_C128, int(real_params), S.data_ptr(),
                                                           sumO.data_ptr(), int(c0 == 0)), ctx.h)
                        c.record()
                        torch.cuda.synchronize()
                        t_eval += a.elapsed_time(b)
                        t_S += b.elapsed_time(c)
                if do_S:
                    a, b = ev(), ev()
                    a.record()
                    L.check(L.lib.nq_sr_finish(ctx.h, S.data_ptr(), sumO.data_ptr(), P, Ns, L.NQ_C128, int(real_params)), ctx.h)
                    b.record()
                    torch.cuda.synchronize()
                    t_S += a.elapsed_time(b)
                    row["sr_assembly_ms"] = t_S
                    row["sr_assembly_TFLOPs_dense"] = flops / t_S / 1e9
                    row["S_GB"] = sb / 1e9
                    row["S_diag_min"] = float(torch.diagonal(S).real.min().item())      # sanity: a covariance has a non-negative diagonal
                row["chunks"] = -(-Ns // Nc)
                row["chunk_samples"] = Nc
                row["evalgrad_ms"] = t_eval
                row["evalgrad_GBs"] = (P * Ns * es + Ns * (es + 8 * W * (2 if kind == "ndm" else 1))) / t_eval / 1e6
                row["evalgrad_frac_hbm"] = row["evalgrad_GBs"] / HBM
                row["samples_per_s_evalgrad"] = Ns / t_eval * 1e3
                if do_S:
                    row["samples_per_s_evalgrad_plus_S"] = Ns / (t_eval + t_S) * 1e3
                rows.append(row)
                print(json.dumps(row), flush=True)
                del O, out, prow, pcol, S, sumO
                torch.cuda.empty_cache()
            del net
json.dump({"hbm_peak_GBs": HBM, "chunk_bytes": CHUNK_BYTES, "rows": rows}, open(out_path, "w"), indent=1)
