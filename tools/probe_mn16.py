"""Round-2 probe: which shared-memory descriptor does tcgen05 kind::f16 take for an MN-major (N contiguous) fp16 B operand
loaded by TMA with SWIZZLE_128B?  Sweeps candidate (LBO, SBO, layout, K-step) encodings, each in its own subprocess (a bad
descriptor can fault the context), and prints the ones whose 128 x 128 x 64 product equals A @ B.
usage (GPU box): python tools/probe_mn16.py            |   python tools/probe_mn16.py one LBO SBO LAYOUT KSTEP"""
import itertools, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(lbo, sbo, layout, kstep):
    import torch
    from megacrn_b200 import _abi
    lib = _abi.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    A = torch.randint(-3, 4, (128, 64), generator=g).half()
    B = torch.randint(-3, 4, (64, 128), generator=g).half()
    C = torch.full((128, 128), float("nan"), device=dev)
    a, b = A.to(dev), B.to(dev)
    st = lib.mcrn_debug_probe_mn16(a.data_ptr(), b.data_ptr(), C.data_ptr(), lbo, sbo, layout, kstep, 1,
                                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = A.float() @ B.float()
    err = (C.cpu() - ref).abs().max().item()
    print(f"lbo={lbo} sbo={sbo} layout={layout} kstep={kstep}: status {st} max|err| {err}", flush=True)
    return 0 if (st == 0 and err == 0.0) else 1


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        sys.exit(one(*map(int, sys.argv[2:6])))
    good = []
    # B tile in smem: two boxes of [64 k rows][64 n = 128 B], 8192 B apart; swizzle atom = 8 rows x 128 B = 1024 B
    # most likely first: 8-row x 128-byte swizzle atoms stacked along K (SBO = 1024), 64-wide N blocks 8192 bytes apart (LBO),
    # K = 16 per MMA = 2048 bytes; ~8 s per candidate (fresh process), stop at the first match unless --all
    cands = [(8192, 1024, 2, 2048)] + [c for c in itertools.product((8192, 1024), (1024, 2048, 8192), (2,), (2048, 1024, 32))
                                       if c != (8192, 1024, 2, 2048)]
    for lbo, sbo, layout, kstep in cands:
        r = subprocess.run([sys.executable, __file__, "one", str(lbo), str(sbo), str(layout), str(kstep)], capture_output=True, text=True,
                           timeout=120)
        line = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
        if r.returncode == 0:
            good.append(line)
            print("MATCH", line, flush=True)
            if "--all" not in sys.argv:
                break
    print(f"{len(good)} matching encodings")
    for l in good:
        print("  ", l)
