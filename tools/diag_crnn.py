"""Per-parameter gradient error of the CUDA CRNN against the fp64 oracle, in the fp32-parity tensor-core mode and on the
fp32 FFMA kernels (tells operand-rounding / ReLU-and-pooling decision flips from logic errors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import crnn_oracle as co
from tatt_b200 import ops
from tatt_b200.crnn import CRNN, parse_crnn_data

SEED = 1234
N = int(os.environ.get("N", 3))
torch.manual_seed(SEED)
net = CRNN(32, 1, 37, 256)
co.perturb_bn_(net.state_dict(), SEED + 1)
net.train()
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
for d in (sd, sd64):
    for k, v in d.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
x = torch.rand(N, 3, 16, 64, generator=torch.Generator().manual_seed(SEED))
net = net.cuda()
gray = parse_crnn_data(x.cuda())
wgt = torch.randn(26, N, 37, generator=torch.Generator().manual_seed(99))
l64 = co.crnn_forward(sd64, gray.cpu().double(), training=True)
(l64 * wgt.double()).sum().backward()
l32 = co.crnn_forward(sd, gray.cpu(), training=True)
(l32 * wgt).sum().backward()
res = {}
for mode in ("fp32", "ffma"):
    ops.set_precision(mode)
    net.zero_grad(set_to_none=True)
    lg = net(gray)
    (lg * wgt.cuda()).sum().backward()
    print(mode, "logits err %.3e" % ((lg.detach().cpu().double() - l64.detach()).abs().max().item() / l64.abs().max().item()))
    res[mode] = {n: p.grad.detach().double().cpu() for n, p in net.named_parameters()}
ops.set_precision("fp32")
print("%-40s %10s %10s %10s" % ("param", "tc fp32", "ffma", "cpu fp32"))
for n in res["fp32"]:
    g = sd64[n].grad
    den = max(g.norm().item(), 1e-30)
    print("%-40s %10.2e %10.2e %10.2e" % (n, (res["fp32"][n] - g).norm().item() / den, (res["ffma"][n] - g).norm().item() / den,
                                          (sd[n].grad.double() - g).norm().item() / den))
