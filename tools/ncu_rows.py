"""Run the row-panel kernels alone at the bench shape (for `ncu --set full -k regex:rows_`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops

dev = "cuda:0"
P = 64 * 32 * 128
for K, N in ((64, 64), (64, 192)):
    x = torch.randn(P, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.1
    b = torch.randn(N, device=dev)
    dy = torch.randn(P, N, device=dev)
    for _ in range(2):
        ops.linear_fwd(x, w, b)
        ops.linear_bwd_weight_rows(dy, x, True)
        ops.linear_bwd_data(dy, w)
torch.cuda.synchronize()
