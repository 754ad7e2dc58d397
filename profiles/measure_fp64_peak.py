"""Measure the FP64 GEMM peaks the SR-assembly roofline is quoted against (MEASURED_PEAKS.json has
HBM and bf16 only): cuBLAS DGEMM / ZGEMM through torch.matmul, best of 10, CUDA events.
Writes profiles/fp64_peaks.json.  usage (on the GPU box): python profiles/measure_fp64_peak.py"""
import json
import os

import torch

out = {}
for name, dt, n, flops_per in (("dgemm", torch.float64, 8192, 2), ("zgemm", torch.complex128, 4096, 8),
                               ("sgemm", torch.float32, 8192, 2)):
    a = torch.randn(n, n, device="cuda", dtype=dt)
    b = torch.randn(n, n, device="cuda", dtype=dt)
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_tflops"] = flops_per * n ** 3 / (best * 1e-3) / 1e12
out["gpu"] = torch.cuda.get_device_name(0)
out["how"] = "torch.matmul (cuBLAS) n=8192 (4096 complex), best of 10, CUDA events; real flops (complex mul-add = 8)"
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fp64_peaks.json")
if os.path.isdir(os.path.join(os.path.dirname(path), "..", "gpurun_out")):
    json.dump(out, open(os.path.join(os.path.dirname(path), "..", "gpurun_out", "fp64_peaks.json"), "w"), indent=1)
print(json.dumps(out))
