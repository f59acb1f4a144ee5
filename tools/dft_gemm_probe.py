"""Measured cost of the DFT-as-GEMM formulation of the headline step, as a LIBRARY GEMM upper bound:
frames [N, 384] (N = 4096 clips x 834 frames; the 128 zero window taps dropped) times the windowed real-DFT matrix
[384, 514] (257 cos + 257 sin columns).  Framing, the |.| / log epilogue and the [F, T] transposition are NOT included, so
the GEMM formulation cannot be faster than these numbers.  Experiment only -- not a product path."""
import torch

dev = torch.device("cuda:0")
N, K, M = 4096 * 834, 384, 514


def timeit(fn, n=10, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = torch.arange(K, dtype=torch.float64)
w = 0.5 - 0.5 * torch.cos(2 * torch.pi * n / K)
k = torch.arange(257, dtype=torch.float64)
ang = 2 * torch.pi * (n[:, None] + 64) * k[None, :] / 512
D = torch.cat([w[:, None] * torch.cos(ang), -w[:, None] * torch.sin(ang)], 1).to(dev)        # [384, 514]
x32 = torch.randn(N, K, device=dev) * 0.1
ref = (x32[:4096].double() @ D).float()
for name, a, b, setup in (
    ("bf16 (1 pass)", x32.bfloat16(), D.bfloat16(), None),
    ("tf32", x32, D.float(), True),
    ("fp32 (no tf32)", x32, D.float(), False),
):
    if setup is not None:
        torch.backends.cuda.matmul.allow_tf32 = setup
    out = torch.empty(N, M, device=dev, dtype=a.dtype)
    ms = timeit(lambda: torch.matmul(a, b, out=out))
    err = float((out[:4096].float() - ref).abs().max() / ref.abs().max())
    print(f"{name:16s} {ms:8.3f} ms   {2.0 * N * K * M / ms / 1e9:8.1f} TFLOP/s   rel max-abs err vs float64 {err:.2e}", flush=True)
    del out
print("register FFT kernel, whole step incl. framing, epilogue and all HBM traffic: 1.59 ms (bench.py)")
