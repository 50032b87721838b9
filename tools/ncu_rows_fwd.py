"""forward row-panel GEMMs alone (timing experiments under ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops
dev = "cuda:0"
P = 64 * 32 * 128
for K, N in ((64, 64), (64, 192)):
    x = torch.randn(P, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1; b = torch.randn(N, device=dev)
    for _ in range(2):
        ops.linear_fwd(x, w, b)
torch.cuda.synchronize()
