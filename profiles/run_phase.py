"""Run one phase of the cfg4 iteration a few times (for ncu captures).
usage: python profiles/run_phase.py {sampler|evaluate|assemble|solve|all} [reps]"""
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

phase = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.WORKLOAD
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
hilb, _, _, liouv = H.p_lindblad_ising_1d(nq, w["N"], w["g"], w["V"])
net = nq.NDM(ctx, hilb, np.float64, w["alpha"], w["alpha"], nq.af_softplus, seed=1234)
nq.init_random_pars_(net, sigma=0.01, seed=1234)
smp = nq.MetropolisSampler(nq.LocalRule(), w["L"], w["passes"] - 1, burn=w["burn"], seed=99)
bs = nq.BatchedSampler(net, smp, liouv, nq.SR(np.float32, eps=w["eps"], algorithm="sr_cholesky"), batch_sz=w["chains"],
                       chain_length=w["L"])
bs.sample_()
bs.precondition_()
torch.cuda.synchronize()
for _ in range(reps):
    if phase in ("sampler", "all"):
        bs.sample_states()
    if phase in ("evaluate", "all"):
        bs.evaluate()
    if phase in ("assemble", "all"):
        bs.evaluate() if phase == "assemble" else None
        bs.assemble()
    if phase in ("solve", "all"):
        bs.precondition_()
torch.cuda.synchronize()
print("done", phase, ctx.launches)
