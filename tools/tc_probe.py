"""Stage-by-stage probe of the tcgen05 GEMM (debug build hooks): dumps shared-memory stage 0 after the TMA
load and pre-fills TMEM with a pattern, to tell apart TMA / descriptor / MMA / TMEM-read problems."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megacrn_b200 import _abi

lib = _abi.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def swz_k_major(tile):      # tile [rows][32] -> expected smem image with 128B swizzle (16B chunk ^= row%8)
    rows = tile.shape[0]
    out = torch.zeros(rows, 32)
    for r in range(rows):
        for c in range(8):
            out[r, 4 * (c ^ (r % 8)):4 * (c ^ (r % 8)) + 4] = tile[r, 4 * c:4 * c + 4]
    return out.reshape(-1)


def probe(M, N, K, ta, tb, BNexp):
    g = torch.Generator().manual_seed(1)
    A = torch.randint(-4, 5, (M, K), generator=g).float()
    B = torch.randint(-4, 5, (K, N), generator=g).float()
    a = (A.T.contiguous() if ta else A.contiguous()).to(dev)
    b = (B.T.contiguous() if tb else B.contiguous()).to(dev)
    c = torch.full((M, N), float("nan"), device=dev)
    dbg = torch.full((32768,), float("nan"), device=dev)
    s = lib.mcrn_debug_tc_gemm(M, N, K, a.data_ptr(), a.shape[1], ta, b.data_ptr(), b.shape[1], tb, c.data_ptr(), N,
                               dbg.data_ptr(), st)
    torch.cuda.synchronize()
    print(f"=== probe M={M} N={N} K={K} ta={ta} tb={tb}: status {s} {lib.mcrn_last_error().decode() if s else ''}")
    if s:
        return
    d = dbg.cpu()
    a_floats = 128 * 32
    b_floats = BNexp * 32
    sa, sb = d[:a_floats], d[a_floats:a_floats + b_floats]
    extra = d[a_floats + b_floats:a_floats + b_floats + 2].view(torch.int32)
    print(f"  tmem_base=0x{extra[0].item() & 0xffffffff:08x} smem_base=0x{extra[1].item() & 0xffffffff:08x}")
    print(f"  smem A: nan={torch.isnan(sa).sum().item()} zeros={(sa == 0).sum().item()}/{a_floats}  first16={sa[:16].tolist()}")
    print(f"  smem B: nan={torch.isnan(sb).sum().item()} zeros={(sb == 0).sum().item()}/{b_floats}  first16={sb[:16].tolist()}")
    Apad = torch.zeros(128, 32); Apad[:min(M, 128), :min(K, 32)] = A[:128, :32]
    Bpad = torch.zeros(BNexp, 32); Bpad[:min(N, BNexp), :min(K, 32)] = B[:32, :BNexp].T
    if not ta:
        print("  A K-major: matches swizzled image:", torch.equal(sa, swz_k_major(Apad)), " matches linear image:", torch.equal(sa, Apad.reshape(-1)))
    else:
        # slabs j: [32 k][32 m] with chunk swizzle by k%8
        exp = torch.cat([swz_k_major(Apad[32 * j:32 * j + 32].T.contiguous()) for j in range(4)])
        print("  A MN-major: matches swizzled slab image:", torch.equal(sa, exp))
    if tb:
        print("  B K-major: matches swizzled image:", torch.equal(sb, swz_k_major(Bpad)), " linear:", torch.equal(sb, Bpad.reshape(-1)))
    else:
        exp = torch.cat([swz_k_major(Bpad[32 * j:32 * j + 32].T.contiguous()) for j in range(BNexp // 32)])
        print("  B N-major: matches swizzled slab image:", torch.equal(sb, exp))
    cc = c.cpu()
    ref = A @ B
    print(f"  C: nan={torch.isnan(cc).sum().item()} zeros={(cc == 0).sum().item()} equal_ref={torch.equal(cc, ref)}"
          f" maxabs_err={(cc - ref).abs().max().item():.3f}")
    print("  C[0,:8]  ", cc[0, :8].tolist(), " ref ", ref[0, :8].tolist())
    print("  C[1,:8]  ", cc[1, :8].tolist())
    print("  C[33,:4] ", cc[33, :4].tolist(), " C[127,:4]", cc[127, :4].tolist() if M > 127 else None)


if __name__ == "__main__":
    for ta, tb in [(0, 1), (0, 0), (1, 1), (1, 0)]:
        probe(128, 64, 32, ta, tb, 64)
    probe(128, 128, 64, 0, 1, 128)
