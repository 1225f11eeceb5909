import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from action_conditioned_gans_b200 import _lib
_lib.use_probe_library()
dev = torch.device("cuda:0")
torch.manual_seed(0)
R, N = 200, 64
A = torch.randn(R, 64, device=dev).to(torch.bfloat16)
B = torch.randn(N, 64, device=dev).to(torch.bfloat16)
for mode in (0, 1):
    for pitch in (8, 10, 12):
        res = []
        for shift in range(0, 12):
            if shift + 15 * pitch + 8 > R:
                continue
            out = torch.zeros(128, N, device=dev)
            _lib.call("acg_debug_umma_shift", A.data_ptr(), R, B.data_ptr(), N, shift, pitch, mode, out.data_ptr(), None)
            torch.cuda.synchronize()
            rows = torch.tensor([shift + (m // 8) * pitch + (m % 8) for m in range(128)], device=dev)
            ref = A[rows].float() @ B.float().t()
            err = (out - ref).abs().max().item()
            res.append("%d:%s" % (shift, "ok" if err < 1e-2 else "%.1f" % err))
        print("base_offset_mode", mode, "pitch", pitch, " ".join(res))
