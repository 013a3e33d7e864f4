#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the LayerNorm-prologue GEMM (needs a build with DETR_B200_DEFINES=-DDETR_GEMM_TIMELINE).
Stamps per warp: 0 after setup, 1 after griddepcontrol.wait, 2 prologue done (workers) / A tiles ready (MMA warp),
3.. per tile: epilogue done (workers) / MMAs committed (MMA warp), 9 all tiles done, 10 stores complete, 11 after the final sync."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch
from detr_b200 import _lib, gemm as G
dev = torch.device("cuda:0")
lib = _lib.load()
lib.detr_gemm_set_debug.argtypes = [ctypes.c_void_p]; lib.detr_gemm_set_debug.restype = None
for (M, N, npe, gelu, xdt) in ((6800, 768, 512, False, torch.bfloat16), (6800, 2048, 0, True, torch.bfloat16), (800, 768, 512, False, torch.float32)):
    x = torch.randn(M, 256, device=dev).to(xdt); gam = torch.ones(256, device=dev); bet = torch.zeros(256, device=dev)
    pos = torch.randn(M, 256, device=dev)
    w = (torch.randn(N, 256, device=dev) * 0.05).bfloat16(); b = torch.randn(N, device=dev)
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    def run():
        if gelu:
            G.gemm_ln(x, gam, bet, 1e-5, w, epilogue=G.EPI_GELU, bias=b, aux=aux, p=0.1, seed=1)
        else:
            G.gemm_ln(x, gam, bet, 1e-5, w, addend=pos, rows_per_batch=M, add_sb=0, add_sr=256, n_pos_end=npe, bias=b)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dbg = torch.zeros(10 * 16, dtype=torch.int64, device=dev)
    lib.detr_gemm_set_debug(dbg.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.detr_gemm_set_debug(None)
    d = dbg.cpu().view(10, 16)
    t0 = int(d[d > 0].min())
    print(f"--- M={M} N={N} n_pos_end={npe} gelu={gelu} x={xdt}: cycles since the first stamp (warps 0-7 workers, 9 = MMA issuer)")
    for wi in (0, 3, 7, 9):
        print(f"warp {wi}: " + " ".join(f"{k}:{int(v) - t0 if v > 0 else -1:6d}" for k, v in enumerate(d[wi].tolist()) if k < 12))
