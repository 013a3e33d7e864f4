#!/usr/bin/env python
"""One launch each of the LayerNorm-prologue GEMM at the encoder q|k|v and FFN1 shapes (the command ncu is pointed at)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch
from detr_b200 import gemm as G
dev = torch.device("cuda:0")
M = 6800
x = torch.randn(M, 256, device=dev).bfloat16(); gam = torch.ones(256, device=dev); bet = torch.zeros(256, device=dev)
pos = torch.randn(M, 256, device=dev)
w = (torch.randn(768, 256, device=dev) * 0.05).bfloat16(); b = torch.randn(768, device=dev)
w2 = (torch.randn(2048, 256, device=dev) * 0.05).bfloat16(); b2 = torch.randn(2048, device=dev)
aux = torch.empty(M, 2048, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    G.gemm_ln(x, gam, bet, 1e-5, w, addend=pos, rows_per_batch=M, add_sb=0, add_sr=256, n_pos_end=512, bias=b)
    G.gemm_ln(x, gam, bet, 1e-5, w2, epilogue=G.EPI_GELU, bias=b2, aux=aux, p=0.1, seed=1)
    G.gemm(x, w2, epilogue=G.EPI_GELU, bias=b2, aux=aux, p=0.1, seed=1)
torch.cuda.synchronize()
print("ok")
