"""Time the fused eval+grad+estimator call of cfg4 (and cfg2-like smaller shapes) -- a few reps, CUDA events.
usage: [NQ_NDM3_WARPS=8] [NQ_NDM3_KERNEL=list] python profiles/time_evaluate.py"""
import json
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
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("NQ_")}}
for dt in (np.float64, np.float32):
    net = nq.NDM(ctx, hilb, dt, w["alpha"], w["alpha"], nq.af_softplus, seed=1234)
    nq.init_random_pars_(net, sigma=0.01, seed=1234)
    smp = nq.MetropolisSampler(nq.LocalRule(), w["L"], w["passes"] - 1, burn=w["burn"], seed=99)
    bs = nq.BatchedSampler(net, smp, liouv, nq.SR(np.float32, eps=w["eps"], algorithm="sr_cholesky"), batch_sz=w["chains"],
                           chain_length=w["L"])
    bs.sample_states()
    for _ in range(3):
        bs.evaluate()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        bs.evaluate()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out[np.dtype(dt).name] = {"ms": ms, "Msamples/s": w["chains"] * w["L"] / ms / 1e3}
    del bs, net
print(json.dumps(out))
