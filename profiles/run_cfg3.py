"""Phase timings of BASELINE cfg3 (TFIM 2D 6x6, RBM alpha=4 logcosh, 16384 chains, L=1, passes=37) in FP64
(complex128 weights) and FP32 (complex64) modes.  usage: python profiles/run_cfg3.py"""
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

ctx = nq.Context(0, torch.cuda.current_stream().cuda_stream)
hilb, Hm = H.p_tfim_2d(nq, 6, 3.0, 1.0)
res = {}
for name, dt in (("fp64(c128)", np.complex128), ("fp32(c64)", np.complex64)):
    net = nq.RBM(ctx, hilb, dt, 4, nq.af_logcosh)
    nq.init_random_pars_(net, 0.01, 1234)
    smp = nq.MetropolisSampler(nq.LocalRule(), 1, 36, burn=100, seed=5)
    bs = nq.BatchedSampler(net, smp, Hm, nq.SR(np.float32, eps=0.1, algorithm="sr_cholesky"), batch_sz=16384, chain_length=1)
    ph = {}

    def phase(k, fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ph[k] = a.elapsed_time(b) / reps
    phase("sampler(burn100+1, passes37)", bs.sample_states)
    phase("evalgrad+E_loc", bs.evaluate)
    phase("centre+force+S", bs.assemble)
    bs.assemble()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); bs.precondition_(); b.record(); torch.cuda.synchronize()
    ph["solve(cholesky)"] = a.elapsed_time(b)
    es = net.out_dtype.itemsize
    ph["P"] = net.P
    ph["O_GB"] = net.P * 16384 * es / 1e9
    res[name] = ph
    del bs, net
print(json.dumps(res, indent=1))
