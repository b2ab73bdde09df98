"""cuBLAS DGEMM throughput (torch.matmul float64) and device copy bandwidth: context numbers
for the FP64 roofline denominator.  Prints one JSON line."""
import json
import torch

def main():
    dev = torch.device("cuda:0")
    out = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"cublas_dgemm_{n}_tflops"] = 2.0 * n**3 / best * 1e-9
    x = torch.empty(1 << 28, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    for _ in range(2):
        y.copy_(x)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["copy_gbs"] = 2 * x.numel() * 8 / best * 1e-6
    print(json.dumps(out))

if __name__ == "__main__":
    main()
