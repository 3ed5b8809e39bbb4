"""Measure the FP64 roofline denominators on the box's B200 and write gpurun_out/fp64_peak.json.

Runs tools/fp64_peak (DFMA / DMMA issue-rate microbenchmarks, built from tools/fp64_peak.cu) and a
cuBLAS DGEMM through torch.matmul (8192^3, best of 10 and back-to-back for ~4 s), the way
MEASURED_PEAKS.json's bf16 figure was taken.  BASELINE.md §2 asks for exactly this number.
"""
import json
import os
import subprocess
import sys
import time

import torch

here = os.path.dirname(os.path.abspath(__file__))
out = {"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
exe = os.path.join(here, "fp64_peak")
if not os.path.exists(exe):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe,
                           os.path.join(here, "fp64_peak.cu")])
out["issue_rate"] = json.loads(subprocess.check_output([exe]).decode())

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e30
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["cublas_dgemm_burst_tflops"] = 2.0 * n ** 3 / best * 1e-9
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); cnt = 0; t0 = time.time()
while time.time() - t0 < 4.0:
    for _ in range(5):
        torch.matmul(a, b, out=c)
    cnt += 5
    torch.cuda.synchronize()
e1.record(); e1.synchronize()
out["cublas_dgemm_sustained_tflops"] = 2.0 * n ** 3 * cnt / e0.elapsed_time(e1) * 1e-9
out["gpu"] = torch.cuda.get_device_name(0)
os.makedirs(os.path.join(here, "..", "gpurun_out"), exist_ok=True)
with open(os.path.join(here, "..", "gpurun_out", "fp64_peak.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
