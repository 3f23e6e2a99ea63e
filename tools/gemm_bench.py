"""Time the tcgen05 GEMM engine on the C2 call-site shapes for each operand-major combination (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megacrn_b200 import _abi
lib = _abi.load(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream

def bench(M, N, K, ta, tb, reps=30):
    a = torch.randn((K, M) if ta else (M, K), device=dev); b = torch.randn((N, K) if tb else (K, N), device=dev)
    c = torch.empty(M, N, device=dev)
    call = lambda: lib.mcrn_gemm(M, N, K, a.data_ptr(), a.shape[1], ta, b.data_ptr(), b.shape[1], tb, c.data_ptr(), N, 2, st)
    s = call()
    if s != 0: return None
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12

shapes = [("propagate dec", 828, 8192, 208), ("propagate enc", 828, 4096, 208), ("propagate_T dec", 208, 8192, 828),
          ("propagate_T enc", 208, 4096, 828), ("gate dec (K=12x128)", 13248, 256, 1536), ("gate enc", 13248, 128, 768),
          ("dxp dec", 13248, 768, 256), ("acc_ds dec", 828, 208, 8192), ("big square", 4096, 4096, 4096)]
for name, M, N, K in shapes:
    row = []
    for ta, tb in [(0, 1), (0, 0), (1, 1), (1, 0)]:
        r = bench(M, N, K, ta, tb)
        row.append("   n/a   " if r is None else f"{r[0]:7.1f}us {r[1]:6.1f}TF")
    print(f"{name:22s} M={M:6d} N={N:5d} K={K:5d} | KK {row[0]} | K,MN {row[1]} | MN,K {row[2]} | MN,MN {row[3]}")
