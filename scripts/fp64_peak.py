"""FP64 yard-stick for the roofline denominators MEASURED_PEAKS.json lacks: cuBLAS DGEMM (torch.matmul fp64) and a
device copy, CUDA-event timed. Test infrastructure only."""
import json
import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
best = 1e9
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = a @ b
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        best = min(best, e0.elapsed_time(e1))
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
cb = 1e9
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y.copy_(x)
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        cb = min(cb, e0.elapsed_time(e1))
print(json.dumps({"dgemm_tflops": 2 * n ** 3 / (best * 1e-3) / 1e12, "dgemm_ms": best,
                  "copy_gbs": 2 * x.numel() * 8 / (cb * 1e-3) / 1e9, "gpu": torch.cuda.get_device_name(0)}))
