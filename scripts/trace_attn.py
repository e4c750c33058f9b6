import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
lib = N.load(); lib.mic_debug_attn_trace.argtypes = [ctypes.c_void_p]
N.set_gemm_mode(1)
Bw, C, heads = 4096, 96, 3
qkv = torch.randn(Bw * 343, 3 * C, device="cuda")
for _ in range(2): ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.mic_debug_attn_trace(buf.data_ptr()); torch.cuda.synchronize()
ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7)); torch.cuda.synchronize()
lib.mic_debug_attn_trace(None)
t = buf.cpu().double(); t0 = t[0]
names = {0: "mma: item start", 1: "mma: rdy_qk", 20: "sm: item start", 21: "sm: rounding done"}
for mt in range(3):
    names.update({2 + mt*4: f"mma: QK{mt} issued", 3 + mt*4: f"mma: p_full{mt}", 4 + mt*4: f"mma: o_empty{mt}", 5 + mt*4: f"mma: PV{mt} issued",
                  22 + mt*5: f"sm: s_full{mt}", 23 + mt*5: f"sm: pass1 done{mt}", 24 + mt*5: f"sm: pass2 done{mt}", 25 + mt*5: f"sm: o_full{mt}", 26 + mt*5: f"sm: epilogue done{mt}"})
ev = sorted((float(t[i] - t0) / 1e3, names[i]) for i in names if t[i] > 0)
for us, n in ev: print(f"{us:8.2f} us  {n}")
