import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from tatt_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
for (N, H, W) in [(2, 4, 128), (3, 32, 128), (2, 6, 256), (64, 32, 128)]:
    x = torch.randn(N, 64, H, W); w = torch.randn(64, 64, 3, 3) * 0.05; b = torch.randn(64)
    xd = ops.nchw_to_nhwc(x.to(dev), 64)
    y = ops.conv2d_fwd(xd, w.to(dev), b.to(dev), 1)
    torch.cuda.synchronize()
    if N <= 3:
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
        got = ops.nhwc_to_nchw(y, 64).cpu().double()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        print((N, H, W), "rel err", err, "mode", os.environ.get("TATT_TMA", "1"))
    else:
        flush = torch.empty(64 << 20, device=dev)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y = ops.conv2d_fwd(xd, w.to(dev), b.to(dev), 1); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("time ms (incl. pack+split)", min(ts), "TF/s", 2 * N * H * W * 64 * 576 / (min(ts) * 1e-3) / 1e12)
