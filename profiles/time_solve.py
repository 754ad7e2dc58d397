"""Time the SR phases of cfg4 after one sampled iteration: assemble (centre + force + S) and the Cholesky solve.
usage: [NQ_CHOL=legacy] python profiles/time_solve.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralquantum.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import nqcuda as nq  # noqa: E402
import bench  # noqa: E402

w = bench.WORKLOAD
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
hilb, _, _, liouv = nq.models.lindblad_ising_1d(w["N"], w["g"], w["V"])
net = nq.NDM(ctx, hilb, np.float64, w["alpha"], w["alpha"], nq.af_softplus, seed=1234)
nq.init_random_pars_(net, sigma=0.01, seed=1234)
smp = nq.MetropolisSampler(nq.LocalRule(), w["L"], w["passes"] - 1, burn=w["burn"], seed=99)
bs = nq.BatchedSampler(net, smp, liouv, nq.SR(np.float32, eps=w["eps"], algorithm="sr_cholesky"), batch_sz=w["chains"],
                       chain_length=w["L"])
bs.sample_()
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("NQ_")}}


def timeit(fn, pre=None, reps=5):
    ts = []
    for _ in range(reps + 1):
        if pre:
            pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts[1:]))


out["sampler_ms"] = timeit(bs.sample_states)
out["evaluate_ms"] = timeit(bs.evaluate)
out["assemble_ms"] = timeit(bs.assemble, pre=bs.evaluate)
out["solve_ms"] = timeit(bs.precondition_, pre=lambda: (bs.evaluate(), bs.assemble()))
dw = bs.dw.clone()
out["dw_norm"] = float(torch.linalg.norm(dw).item())
print(json.dumps(out))
