"""Blind-debug aid for the tcgen05 GEMM: structured operands that expose layout / descriptor mistakes.
Prints, per case, rel-L2 error vs fp64 and (when wrong) which rows/cols/k-slices are affected."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megacrn_b200 import _abi

lib = _abi.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def run(M, N, K, ta, tb, a, b, eng):
    c = torch.full((M, N), float("nan"), device=dev)
    s = lib.mcrn_gemm(M, N, K, a.data_ptr(), a.shape[1], ta, b.data_ptr(), b.shape[1], tb, c.data_ptr(), N, eng, st)
    torch.cuda.synchronize()
    return s, c


def case(M, N, K, ta, tb, kind):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(K, N, generator=g)
    if kind == "b_eye":
        B = torch.zeros(K, N); B[torch.arange(min(K, N)), torch.arange(min(K, N))] = 1
    if kind == "a_eye":
        A = torch.zeros(M, K); A[torch.arange(min(M, K)), torch.arange(min(M, K))] = 1
    if kind == "ints":
        A = torch.randint(-4, 5, (M, K), generator=g).float(); B = torch.randint(-4, 5, (K, N), generator=g).float()
    a = (A.T.contiguous() if ta else A.contiguous()).to(dev)
    b = (B.T.contiguous() if tb else B.contiguous()).to(dev)
    ref = (A.double() @ B.double())
    s, c = run(M, N, K, ta, tb, a, b, 2)
    if s != 0:
        print(f"[{kind}] M={M} N={N} K={K} ta={ta} tb={tb}: status {s} {lib.mcrn_last_error().decode()}")
        return
    c = c.cpu().double()
    nan = torch.isnan(c).sum().item()
    err = ((c - ref).norm() / ref.norm()).item() if nan == 0 else float("nan")
    flag = "OK " if (nan == 0 and err < 3e-3) else "BAD"
    print(f"{flag} [{kind}] M={M} N={N} K={K} ta={ta} tb={tb}: rel-L2 {err:.3e} nan={nan}")
    if flag == "BAD":
        bad = ((c - ref).abs() > 1e-2 * ref.abs().max()) | torch.isnan(c)
        rows = bad.any(1).nonzero().flatten().tolist()
        cols = bad.any(0).nonzero().flatten().tolist()
        print(f"    bad rows {len(rows)}/{M}: {rows[:12]}...  bad cols {len(cols)}/{N}: {cols[:12]}...")
        if kind in ("b_eye", "a_eye", "ints"):
            print("    got[0,:8] ", [round(v, 3) for v in c[0, :8].tolist()])
            print("    ref[0,:8] ", [round(v, 3) for v in ref[0, :8].tolist()])
            print("    got[1,:8] ", [round(v, 3) for v in c[1, :8].tolist()])
            print("    ref[1,:8] ", [round(v, 3) for v in ref[1, :8].tolist()])
            # where did value ref[i,j] land?
            if kind == "b_eye":
                # C should equal A[:, :N]; find for c[0, j] which k it equals
                a0 = A[0].double()
                m = [(j, (a0 - c[0, j]).abs().argmin().item()) for j in range(min(N, 16))]
                print("    c[0,j] == A[0,k]:", m)


if __name__ == "__main__":
    print("engine check:", lib.mcrn_device_ok())
    for ta, tb in [(0, 1), (0, 0), (1, 1), (1, 0)]:
        for kind in ("ints", "b_eye", "rand"):
            case(128, 64, 32, ta, tb, kind)      # single k-iteration, single tile
        case(128, 128, 64, ta, tb, "rand")       # 2 k-iterations, BN=128
        case(256, 192, 208, ta, tb, "rand")      # multi tile, K tail
        case(828, 512, 208, ta, tb, "rand")
        case(100, 36, 72, ta, tb, "rand")        # M/N/K tails
