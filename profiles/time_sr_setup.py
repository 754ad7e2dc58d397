"""Time S assembly (nq_sr_setup) alone on synthetic centred gradient rows of the cfg4 / cfg3 shapes.
usage: python profiles/time_sr_setup.py [fp32|fp64]   (NQ_SR_FP32_PATH=dmma selects the FP64-tensor fallback in FP32 mode)"""
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
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
rdt, cdt = (torch.float32, np.complex64) if mode == "fp32" else (torch.float64, np.complex128)
out = {"mode": mode, "fp32_path": os.environ.get("NQ_SR_FP32_PATH", "tcgen05")}
for name, P, Ns, real_params, structured in (("cfg4 NDM P=2244 Ns=65536", 2244, 65536, True, True),
                                             ("cfg3 RBM P=5364 Ns=16384", 5364, 16384, False, False)):
    g = torch.Generator(device="cuda").manual_seed(1)
    O = torch.randn((Ns, P, 2), device="cuda", dtype=rdt, generator=g)
    if structured:
        O[:, : P // 2, 1] = 0
        O[:, P // 2:, 0] = 0
    sdt = np.dtype(cdt if not real_params else (np.float32 if mode == "fp32" else np.float64))
    S = torch.zeros((P, P, 2 if not real_params else 1), device="cuda", dtype=rdt)
    F = torch.zeros((P, 2), device="cuda", dtype=rdt)
    gc = np.ones(P, cdt)

    def run():
        L.check(L.lib.nq_sr_setup(ctx.h, O.data_ptr(), P, P, Ns, Ns, L.nq_dtype(np.dtype(cdt)), L.ptr(gc), int(real_params),
                                  S.data_ptr(), F.data_ptr()), ctx.h)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    flops = 2.0 * P * P / 2 * Ns * (4 if not structured else 1) * (2 if not real_params else 1) / (2 if not real_params else 1)
    # useful real MACs: lower triangle, complex rows = 2 real components (structured: 1 active), HERK adds the imaginary part
    macs = P * (P + 1) / 2 * Ns * ((1 if structured else 2) * (2 if not real_params else 1))
    out[name] = {"ms": ms, "useful_TFLOPs": 2 * macs / ms / 1e9}
    del O, S
print(json.dumps(out, indent=1))
