"""Time the SRB BiGRU scans alone at the bench shape (CUDA events).  TATT_GRU_MMA=0 selects the scalar kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops
dev = "cuda:0"
N, H, W = 64, 32, 128
P = N * H * W
gi = torch.randn(P, 192, device=dev); whh = torch.randn(2, 96, 32, device=dev) * 0.1; bhh = torch.randn(2, 96, device=dev) * 0.1
dout = torch.randn(P, 64, device=dev)
for name, geom in (("horizontal T=128", (N * H, W, 1, W, 0, 1)), ("vertical T=32", (N * W, H, W, H * W, 1, W))):
    for _ in range(2):
        out, gates = ops.gru32_scan_fwd(gi, whh, bhh, *geom, save=True)
        ops.gru32_scan_bwd(dout, gates, whh, *geom)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(5):
        out, gates = ops.gru32_scan_fwd(gi, whh, bhh, *geom, save=True)
    ev[1].record()
    for _ in range(5):
        ops.gru32_scan_bwd(dout, gates, whh, *geom)
    ev[2].record()
    torch.cuda.synchronize()
    print("%s mma=%s: fwd %.1f us, bwd %.1f us" % (name, os.environ.get("TATT_GRU_MMA", "1"), ev[0].elapsed_time(ev[1]) * 200,
                                                 ev[1].elapsed_time(ev[2]) * 200))
