#!/usr/bin/env python
"""Timeline of CTA 0 of the fused attention backward kernel (clock64 stamps written when a debug buffer is set).
usage: python tools/attn_timeline.py [B nh L S]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch
from detr_b200 import _lib
from detr_b200.attention import attention_backward, attention_forward

B, nh, L, S = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (8, 8, 850, 850)
dev = torch.device("cuda:0")
C = nh * 32
q = torch.randn(B, L, C, device=dev).bfloat16(); k = torch.randn(B, S, C, device=dev).bfloat16()
v = torch.randn(B, S, C, device=dev).bfloat16(); do = torch.randn(B, L, C, device=dev).bfloat16()
o, lse = attention_forward(q, k, v, dropout_p=0.1, seed=1)
for _ in range(3):
    attention_backward(do, q, k, v, o, lse, dropout_p=0.1, seed=1)
dbg = torch.zeros(20 * 32 * 8 + 1024 * 4, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.detr_attention_bwd_set_debug.argtypes = [ctypes.c_void_p]; lib.detr_attention_bwd_set_debug.restype = None
lib.detr_attention_bwd_set_debug(dbg.data_ptr())
attention_backward(do, q, k, v, o, lse, dropout_p=0.1, seed=1)
torch.cuda.synchronize()
lib.detr_attention_bwd_set_debug(None)
cta = dbg[5120:].view(1024, 4).cpu()
d = dbg[:5120].view(20, 32, 8).cpu()
t0 = int(d[d > 0].min())
T = min(24, int((d[0, :, 0] > 0).sum()))   # pairs of CTA 0 with stamps (persistent kernel: several items)
print("cycles relative to the first stamp; compute warps: wait_sdp> <sdp_full | ld done> <ds_empty | math done | dq_readout done")
for w in (0, 5, 10, 15):
    for t in range(T):
        r = [int(x) - t0 if x > 0 else -1 for x in d[w, t, :6]]
        print(f"warp {w:2d} tile {t}: {r}   math={r[4]-r[3]} wait_sdp={r[1]-r[0]} wait_ds_empty={0} dq_readout={r[5]-r[4]}")
print("all math warps, pair 3: [start, sdp_full, ld done, math done, dq done]  math duration per tile")
for w in range(16):
    r = [int(x) - t0 for x in (d[w, 3, 0], d[w, 3, 1], d[w, 3, 3], d[w, 3, 4], d[w, 3, 5])]
    print(f"warp {w:2d} (smsp {w % 4}, kq {w // 4}): {r}  math/tile = {[int(d[w, t, 4] - d[w, t, 3]) for t in range(T)]}")
print("MMA warp: before ds_full wait | after | issued")
for t in range(T):
    r = [int(x) - t0 if x > 0 else -1 for x in d[17, t, :3]]
    print(f"tile {t}: {r}  waited={r[1]-r[0]}")

n = (cta[:, 0] > 0).sum().item()
if n:
    c = cta[:n]
    g0 = int(c[:, 0].min())
    dur = (c[:, 2] - c[:, 0]).float()
    setup = (c[:, 1] - c[:, 0]).float()
    print(f"CTAs {n}: kernel span {(int(c[:, 2].max()) - g0) / 1e3:.1f} us; CTA duration ns mean {dur.mean():.0f} min {dur.min():.0f} max {dur.max():.0f}; setup (to TMEM alloc) mean {setup.mean():.0f} ns")
    import collections
    per = collections.defaultdict(list)
    for i in range(n):
        per[int(c[i, 3])].append((int(c[i, 0]) - g0, int(c[i, 2]) - g0))
    for sm in sorted(per)[:4]:
        print("SM", sm, sorted(per[sm]))
    busy = sum(e - s_ for v in per.values() for s_, e in v)
    print(f"SMs used {len(per)}, mean CTAs/SM {n / len(per):.2f}, sum of CTA time / (SMs x span) = {busy / (len(per) * (int(c[:, 2].max()) - g0)):.2f}")
