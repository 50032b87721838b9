"""Run the SRB BiGRU scans alone at the bench shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops
dev = "cuda:0"
N, H, W = 64, 32, 128
P = N * H * W
gi = torch.randn(P, 192, device=dev); whh = torch.randn(2, 96, 32, device=dev) * 0.1; bhh = torch.zeros(2, 96, device=dev)
dout = torch.randn(P, 64, device=dev)
for geom in ((N * H, W, 1, W, 0, 1), (N * W, H, W, H * W, 1, W)):
    for _ in range(2):
        out, gates = ops.gru32_scan_fwd(gi, whh, bhh, *geom, save=True)
        ops.gru32_scan_bwd(dout, gates, whh, *geom)
torch.cuda.synchronize()
