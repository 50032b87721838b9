"""Times the data gradient of the tail 3x3 64 -> 256 convolution (the 256 -> 64 transposed problem) at the benchmark
geometry, 8 back-to-back issues per event pair.  TATT_ROLL_CIN=0 selects the im2col GEMM engine for comparison."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tatt_b200 import ops

dev = torch.device("cuda:0")
N, H, W = 64, 32, 128
torch.manual_seed(0)
x = torch.randn(N, H, W, 64, device=dev)
w = torch.randn(256, 64, 3, 3, device=dev) * 0.04
dy = torch.randn(N, H, W, 256, device=dev)
fill = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    return ops.conv2d_bwd(x, w, dy, 1, need_dx=True, need_dw=False, has_bias=False)[0]


ref = None
for _ in range(3):
    ref = run()
torch.cuda.synchronize()
ts = []
for rep in range(5):
    fill.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 8 * 1e3)
print("TATT_ROLL_CIN=%s dgrad 256->64 N=%d %dx%d: %s us (min %.1f)" % (os.environ.get("TATT_ROLL_CIN", "1"), N, H, W,
                                                                    " ".join("%.1f" % t for t in ts), min(ts)))
# fp64 check on a slice
xd = torch.nn.functional.conv_transpose2d(dy[:2].permute(0, 3, 1, 2).double(), w.double(), padding=1)
err = (ref[:2].permute(0, 3, 1, 2).double() - xd).abs().max().item() / xd.abs().max().item()
print("max rel err vs fp64:", err)
