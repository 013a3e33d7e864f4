#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent attention forward kernel (build with DETR_B200_DEFINES=-DDETR_FWD_TIMELINE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch
from detr_b200 import _lib
from detr_b200.attention import attention_forward
B, nh, L, S = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (8, 8, 850, 850)
dev = torch.device("cuda:0"); C = nh * 32
q = torch.randn(B, L, C, device=dev).bfloat16(); k = torch.randn(B, S, C, device=dev).bfloat16(); v = torch.randn(B, S, C, device=dev).bfloat16()
for _ in range(3): attention_forward(q, k, v, dropout_p=0.1, seed=1)
dbg = torch.zeros(20 * 32 * 8, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.detr_attention_fwd_set_debug.argtypes = [ctypes.c_void_p]; lib.detr_attention_fwd_set_debug.restype = None
lib.detr_attention_fwd_set_debug(dbg.data_ptr())
attention_forward(q, k, v, dropout_p=0.1, seed=1)
torch.cuda.synchronize()
lib.detr_attention_fwd_set_debug(None)
d = dbg.view(20, 32, 8).cpu()
t0 = int(d[d > 0].min())
n = int((d[0, :, 0] > 0).sum())
print("softmax warp: [top, s_full seen, TMEM loaded, row max exchanged, P stored, arrived]")
for w in (0, 5, 10, 15):
    for j in range(min(n, 24)):
        r = [int(x) - t0 for x in d[w, j, :6]]
        print(f"warp {w:2d} pair {j:2d}: {r}  wait_s={r[1]-r[0]} ld={r[2]-r[1]} max+xch={r[3]-r[2]} math={r[4]-r[3]} fence={r[5]-r[4]}" + (f" period={int(d[w,j,0]-d[w,j-1,0])}" if j else ""))
print("MMA warp: [before p_full wait, after, PV + next scores issued]")
for j in range(min(n, 24)):
    r = [int(x) - t0 for x in d[17, j, :3]]
    print(f"pair {j:2d}: {r} waited={r[1]-r[0]} issue={r[2]-r[1]}")
