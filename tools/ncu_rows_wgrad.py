"""One tatt_rows_wgrad call per NB = 1, 2, 3 at the benchmark's M (for `ncu -k regex:rows_wgrad -c 6 --set full`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tatt_b200 import ops

dev = torch.device("cuda:0")
M = 64 * 32 * 128
x = torch.randn(M, 64, device=dev)
for N in (64, 128, 192):
    dy = torch.randn(M, N, device=dev)
    ops.linear_bwd_weight_rows(dy, x, True)
torch.cuda.synchronize()
