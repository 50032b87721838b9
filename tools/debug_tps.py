import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from tatt_b200 import tsrn, ops
torch.manual_seed(0)
N, H, W = 3, 16, 64
dev = "cuda:0"
tps = tsrn.TPSSpatialTransformer(output_image_size=(H, W), num_control_points=20, margins=(0.05, 0.05))
x = torch.rand(N, 4, H, W)
gen = torch.Generator().manual_seed(5)
base = tsrn.STNHead(4, 20).stn_fc2.bias.detach().view(1, 20, 2)
for amp in (0.0, 0.08):
    ctrl = base + amp * torch.randn(N, 20, 2, generator=gen)
    Y = torch.cat([ctrl.double(), tps.padding_matrix.double().expand(N, 3, 2)], 1)
    M64 = torch.matmul(tps.inverse_kernel.double(), Y)
    src64 = torch.matmul(tps.target_coordinate_repr.double(), M64)
    out64 = F.grid_sample(x.double(), 2.0 * torch.clamp(src64.view(-1, H, W, 2), 0, 1) - 1.0, mode="bilinear",
                          padding_mode="zeros", align_corners=False)
    xd = ops.nchw_to_nhwc(x.to(dev), 4)
    od, sd = ops.tps_sample_fwd(xd, ctrl.to(dev).contiguous(), tps.inverse_kernel.to(dev), tps.target_coordinate_repr.to(dev), want_src=True)
    o = ops.nhwc_to_nchw(od, 4).cpu().double()
    es = (sd.cpu().double() - src64).abs()
    eo = (o - out64).abs()
    print("amp", amp, "src err max", es.max().item(), "out err max", eo.max().item())
    idx = eo.flatten().argmax().item()
    n, c, y, xx = idx // (4 * H * W), (idx // (H * W)) % 4, (idx // W) % H, idx % W
    p = y * W + xx
    print("  worst at", n, c, y, xx, "mine", o[n, c, y, xx].item(), "ref", out64[n, c, y, xx].item(),
          "src mine", sd[n, p].cpu().tolist(), "src64", src64[n, p].tolist())
