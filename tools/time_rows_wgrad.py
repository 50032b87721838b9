"""Times tatt_rows_wgrad (kernel + partial reduction) at the benchmark's M = 64 x 32 x 128 pixel rows for NB = 1, 2, 3
(8 back-to-back issues per event pair, inputs far larger than L2 in total).  TATT_WG_WS=0 selects the
one-barrier-per-slab kernels for comparison.  Prints the algorithmic HBM rate."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tatt_b200 import ops

dev = torch.device("cuda:0")
M = 64 * 32 * 128
torch.manual_seed(0)
x = torch.randn(M, 64, device=dev)
for N in (64, 128, 192):
    dy = torch.randn(M, N, device=dev)
    for _ in range(3):
        dW, db = ops.linear_bwd_weight_rows(dy, x, True)
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            ops.linear_bwd_weight_rows(dy, x, True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 8 * 1e3)
    t = min(ts)
    byt = 4.0 * M * (64 + N)
    ref = dy[:4096].double().t() @ x[:4096].double()
    got = ops.linear_bwd_weight_rows(dy[:4096].contiguous(), x[:4096].contiguous(), False)[0]
    full_ref = dy.double().t() @ x.double()
    err = (dW.double() - full_ref).abs().max().item() / full_ref.abs().max().item()
    print("TATT_WG_WS=%s NB=%d: %.1f us  %.0f GB/s algorithmic  (rel err vs fp64 %.2e)" % (
        os.environ.get("TATT_WG_WS", "1"), N // 64, t, byt / t / 1e3, err))
