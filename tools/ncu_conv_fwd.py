"""conv3x3 64->64 forward alone at the bench shape (timing experiments under ncu); argv[1] = precision mode"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops
dev = "cuda:0"
if len(sys.argv) > 1:
    ops.set_precision(sys.argv[1])
B, h, w = 64, 32, 128
x = torch.randn(B, h, w, 64, device=dev); wt = torch.randn(64, 64, 3, 3, device=dev) * 0.05; b = torch.zeros(64, device=dev)
for _ in range(3):
    y = ops.conv2d_fwd(x, wt, b, 1)
torch.cuda.synchronize()
