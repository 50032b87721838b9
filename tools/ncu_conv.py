"""Run the dominant kernels alone (for `ncu --set full`): conv3x3 64->64 fwd at the bench shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops
dev = "cuda:0"
B, h, w = 64, 32, 128
x = torch.randn(B, h, w, 64, device=dev); wt = torch.randn(64, 64, 3, 3, device=dev) * 0.05; b = torch.zeros(64, device=dev)
dy = torch.randn(B, h, w, 64, device=dev)
for _ in range(3):
    y = ops.conv2d_fwd(x, wt, b, 1)
    ops.conv2d_bwd(x, wt, dy, 1)
torch.cuda.synchronize()
