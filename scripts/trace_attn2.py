"""Phase trace of the two-issuer large-window attention forward (window_attn_tc2_fwd_kernel<2>), CTA 0, config 4.
issuer x, tile n of its pipeline: n*16 + x*8 + {0: S0 issue, 1: P chunk A seen, 2: chunk B seen, 3: chunk C seen, 4: chunk D
seen, 5: tile issued}; softmax group x: 256 + n*16 + x*8 + {0: waits S0, 1: S0 seen, 2: block 0 published / waits S1, 3: S1
seen, 4: block 1 published / waits O, 5: O seen, 6: stores issued}.  SM clocks at 1.965 GHz."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
lib = N.load(); lib.mic_debug_attn_trace.argtypes = [ctypes.c_void_p]
N.set_gemm_mode(1)
Bw, C, heads = 1024, 96, 3
qkv = torch.randn(Bw * 343, 3 * C, device="cuda")
for _ in range(2): ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
lib.mic_debug_attn_trace(buf.data_ptr()); torch.cuda.synchronize()
ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7)); torch.cuda.synchronize()
lib.mic_debug_attn_trace(None)
t = buf.cpu().tolist()
t0 = min(v for v in t if v)
us = lambda v: ("%7.2f" % ((v - t0) / 1965.0)) if v else "      -"
print("pipe tile | issuer: S0 issue | A seen | B seen | C seen | D seen | done || softmax: wait S0 | S0 seen | blk0 pub | S1 seen | blk1 pub | O seen | stored")
for n in range(2, 12):
    for x in range(2):
        a = t[n * 16 + x * 8: n * 16 + x * 8 + 6]; b = t[256 + n * 16 + x * 8: 256 + n * 16 + x * 8 + 7]
        print(f"  {x}   {n:2d}  | " + " | ".join(us(v) for v in a) + " || " + " | ".join(us(v) for v in b))
