"""Short ncu target: the stage-0 forward GEMMs (persistent kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
N.set_gemm_mode(1)
dev = "cuda"; M = 65536
x = torch.randn(M, 48, device=dev); w = torch.randn(192, 48, device=dev); b = torch.randn(192, device=dev)
pre = torch.empty(M, 192, device=dev)
x2 = torch.randn(M, 192, device=dev); w2 = torch.randn(48, 192, device=dev); b2 = torch.randn(48, device=dev); res = torch.randn(M, 48, device=dev)
for _ in range(3):
    ops.linear_fwd(x, 48, w, b, M, 192, 48, act=True, pre=pre)
    ops.linear_fwd(x2, 192, w2, b2, M, 48, 192, res=res)
torch.cuda.synchronize()
