import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import kernel_spec as K
from oracle import megacrn_oracle as O
from megacrn_b200 import _abi
lib = _abi.load(); dev = torch.device("cuda:0")
for N in (45, 207, 325):
    d = O.Dims(num_nodes=N)
    p = O.init_params(d, seed=0)
    s_ref, _ = K.supports_fwd({k: v.double() for k, v in p.items()}, 3)
    dims = _abi.Dims(batch=1, num_nodes=N, seq_len=1, horizon=1, input_dim=1, output_dim=1, ycov_dim=1, rnn_units=64,
                     num_layers=1, cheb_k=3, mem_num=20, mem_dim=64)
    ld = lib.mcrn_support_ld(N)
    nbytes = lib.mcrn_workspace_bytes(dims, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dp = {k: v.to(dev) for k, v in p.items()}
    for mask in (0, 0x40):
        lib.mcrn_set_debug_mask(mask)
        S = torch.zeros(4, N, ld, device=dev); Sr = torch.zeros(4, N, ld, device=dev)
        st = lib.mcrn_supports_fwd2(dims, dp["memory.Memory"].data_ptr(), dp["memory.We1"].data_ptr(), dp["memory.We2"].data_ptr(),
                                    S.data_ptr(), Sr.data_ptr(), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        S, Sr = S[:, :, :N].cpu().double(), Sr[:, :, :N].cpu().double()
        for k in range(4):
            e = (S[k] - s_ref[k]).abs(); er = (Sr[k] - s_ref[k]).abs()
            i = int(e.argmax()); 
            print(f"N={N} cheb={'simt' if mask else 'tc'} blk{k}: exact max|d|={e.max():.2e} at ({i//N},{i%N}) rel {float((S[k]-s_ref[k]).norm()/s_ref[k].norm()):.2e}; rounded rel {float((Sr[k]-s_ref[k]).norm()/s_ref[k].norm()):.2e}")
lib.mcrn_set_debug_mask(0)
