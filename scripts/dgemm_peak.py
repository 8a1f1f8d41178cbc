"""Yard-stick: cuBLAS FP64 GEMM rate on this GPU (the y-GEMM kernels are measured against it in DESIGN.md)."""
import torch, time
for n, m, k in ((8192, 8192, 8192), (129, 65536 * 4, 136), (4096, 4096, 4096)):
    a = torch.randn(n, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, m, dtype=torch.float64, device="cuda")
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        c = a @ b
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("dgemm %dx%dx%d: %.3f ms  %.2f TFLOP/s" % (n, m, k, ms, 2.0 * n * m * k / ms / 1e9))
