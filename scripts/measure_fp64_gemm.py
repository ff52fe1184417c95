"""FP64 GEMM rate of the box (cuBLAS through torch.matmul, 8192^3): the independent denominator SURVEY 8d asks for next
to the DMMA rate derived from the ncu pipe-activity capture.  Prints one JSON line."""
import json

import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps(dict(what="cuBLAS DGEMM 8192^3 (torch.matmul float64), best of 5", ms=best, tflops=2 * n ** 3 / best / 1e9)))
