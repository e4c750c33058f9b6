"""Phase trace of the tcgen05 window-attention backward (csrc/window_attn_tc_bwd.cu), CTA 0 of a config-4 launch.
Slots: g*8 + {0: issuer before S(g+1), 1: after, 2: p_ready(g) seen, 3: OUT(g) issued, 4: group waits s_full(g),
5: s_full(g) seen, 6: P/dS(g) published, 7: epilogue stored}; 1000.. item-level stamps.  SM clocks (1.965 GHz)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
lib = N.load(); lib.mic_debug_attn_bwd_trace.argtypes = [ctypes.c_void_p]
N.set_gemm_mode(1)
Bw, C, heads = 1024, 96, 3
qkv = torch.randn(Bw * 343, 3 * C, device="cuda"); do = torch.randn(Bw * 343, C, device="cuda")
o, lse = ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
for _ in range(2): ops.window_attn_bwd(qkv, o, do, lse, C, heads, Bw, (7, 7, 7), (7, 7, 7))
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.mic_debug_attn_bwd_trace(buf.data_ptr()); torch.cuda.synchronize()
ops.window_attn_bwd(qkv, o, do, lse, C, heads, Bw, (7, 7, 7), (7, 7, 7)); torch.cuda.synchronize()
lib.mic_debug_attn_bwd_trace(None)
t = buf.cpu().tolist()
t0 = t[1000]
us = lambda v: (v - t0) / 1965.0
print("item: issuer start 0.00 | rdy(Q,K) %.2f | stats start %.2f | stats done %.2f | conditioning+bar done %.2f us" %
      (us(t[1001]), us(t[1002]), us(t[1003]), us(t[1004])))
print(" g | S(g+1) issue  dur | p_ready(g) | OUT(g) done  dur || grp wait s_full | seen | published  simt | epilogue")
G = 36
for g in range(G):
    r = t[g * 8:g * 8 + 8]
    f = lambda i: ("%8.2f" % us(r[i])) if r[i] else "       -"
    d = lambda i, j: ("%5.2f" % ((r[j] - r[i]) / 1965.0)) if r[i] and r[j] else "    -"
    print(f"{g:2d} | {f(0)} {d(0,1)} | {f(2)} | {f(3)} {d(2,3)} || {f(4)} | {f(5)} | {f(6)} {d(5,6)} | {f(7)}")
last = max(v for v in t[:G * 8] if v)
print("item total %.2f us" % us(last))
