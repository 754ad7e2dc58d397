"""Per-phase cycle breakdown of the fused NDM kernel (instrumented build, -DNQ_PROFILE_PHASES).
usage: NQCUDA_LIB=neuralquantum.jl_b200/libnqcuda_prof.so python profiles/phase_cycles.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import nqcuda as nq  # noqa: E402
import helpers as H  # noqa: E402
import bench  # noqa: E402

w = bench.WORKLOAD
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
hilb, _, _, liouv = H.p_lindblad_ising_1d(nq, w["N"], w["g"], w["V"])
net = nq.NDM(ctx, hilb, np.float64, w["alpha"], w["alpha"], nq.af_softplus, seed=1234)
nq.init_random_pars_(net, sigma=0.01, seed=1234)
smp = nq.MetropolisSampler(nq.LocalRule(), w["L"], w["passes"] - 1, burn=w["burn"], seed=99)
bs = nq.BatchedSampler(net, smp, liouv, nq.SR(np.float32, eps=w["eps"]), batch_sz=w["chains"], chain_length=w["L"])
rng = np.random.default_rng(1)
sig = (rng.integers(0, 2, (w["N"], w["chains"], w["L"])).astype(float), rng.integers(0, 2, (w["N"], w["chains"], w["L"])).astype(float))
bs.set_samples(sig)
bs.evaluate()
out = (C.c_ulonglong * 8)()
lib = nq._lib.lib
lib.nq_debug_phase_cycles(out, 1)
bs.evaluate()
lib.nq_debug_phase_cycles(out, 0)
names = ["decode + operator pass", "base quantities", "S1 ratio steps (v3: phase AB)", "W weights (v3: phase C)", "(v3: totals)", "S2 + output"]
tot = sum(out[:6])
for n, v in zip(names, out):
    print("%-38s %6.1f%%  %8.0f cycles/sample" % (n, 100.0 * v / tot, v / (w["chains"] * w["L"])))
print("raw slots 3/6/7 per sample:", [round(out[i] / (w["chains"] * w["L"])) for i in (3, 6, 7)])
if False:
    print("phase C fast-path bodies of warp 0: %.0f cycles each, %.2f per sample" % (out[6] / out[7], out[7] / (w["chains"] * w["L"])))
