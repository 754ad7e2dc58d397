"""End-to-end check against the reference-held exact energy (examples/ising1d.jl:46-47): TFIM 1D N=20, h=J=1, RBM
(logcosh) + Metropolis local flips + SR, run to convergence on the GPU; the energy averaged over the last iterations is
recorded next to -1.274549484318 * 20 with its Monte Carlo error and the variational gap.
usage: python profiles/run_example_convergence.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralquantum.jl_b200"))
import torch  # noqa: E402
import nqcuda as nq  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "example_ising1d.json")
N, h, J = 20, 1.0, 1.0
exact = -1.274549484318 * 20
ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
hilb = nq.HomogeneousSpin(N)
H = nq.LocalOperator(hilb)
for i in range(1, N + 1):
    H = H - h * nq.sigmax(hilb, i)
    H = H + (J * nq.sigmaz(hilb, i)) * nq.sigmaz(hilb, i % N + 1)
res = {}
for name, dt, alpha, B, L, iters in (("example (Float32 real RBM alpha=1, 8 chains x 125)", np.float32, 1, 8, 125, 300),
                                     ("large batch (complex128 RBM alpha=2, 1024 chains x 16)", np.complex128, 2, 1024, 16, 400)):
    net = nq.RBM(ctx, hilb, dt, alpha, nq.af_logcosh)
    nq.init_random_pars_(net, sigma=0.01, seed=1234)
    it = nq.BatchedSampler(net, nq.MetropolisSampler(nq.LocalRule(), L, N, burn=100, seed=1234), H,
                           nq.SR(np.float32, eps=0.1, algorithm=nq.sr_cg, precision=1e-3), batch_sz=B)
    opt = nq.Descent(0.1 if dt == np.float32 else 0.05)
    hist = []
    t0 = time.time()
    for i in range(1, iters + 1):
        stat, _ = it.sample_()
        hist.append((stat.mean.real, stat.error))
        it.precondition_(i)
        it.update_(opt)
    torch.cuda.synchronize()
    tail = np.array(hist[-50:])
    e, err = float(tail[:, 0].mean()), float(np.sqrt((tail[:, 1] ** 2).mean() / len(tail)))
    res[name] = {"iterations": iters, "seconds": time.time() - t0, "energy_last50": e, "mc_error_of_the_mean": err,
                 "exact": exact, "rel_gap": (e - exact) / abs(exact), "gap_in_sigma": (e - exact) / err,
                 "energy_first": hist[0][0], "energy_every_50": [hist[k][0] for k in range(0, iters, 50)]}
    print(name, res[name], flush=True)
json.dump(res, open(out, "w"), indent=1)
