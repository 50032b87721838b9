"""Per-parameter gradient error of the fused decoder-layer path and of the op-by-op path against the fp64 oracle
(G16, N = 3, dropout 0): tells whether a fused-vs-op-by-op difference is an error of one path or the noise of both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tatt_b200
from oracle import ref_harness as rh
from oracle import tatt_oracle as orc
from tatt_b200 import ops

kw = dict(scale_factor=2, width=128, height=32, STN=False, mask=True)
N = int(os.environ.get("N", 3))
torch.manual_seed(1234)
net = tatt_b200.TSRN_TL_TRANS(**kw)
rh.perturb_(net)
rh.zero_dropout(net)
net.train()
sd = orc.clone_sd(net.state_dict(), requires_grad=True)
sd64 = {}
for k, v in sd.items():
    t = v.detach().clone().double() if v.is_floating_point() else v.detach().clone()
    if v.requires_grad:
        t.requires_grad_(True)
    sd64[k] = t
x, tp = orc.synthetic_inputs(N, 16, 64, seed=5)
wgt = torch.randn(N, 4, 32, 128, generator=torch.Generator().manual_seed(9))
o64 = orc.tsrn_tl_trans_forward(sd64, x.double(), tp.double(), training=True, stn=False, dropout_p=0.0)[0]
(o64 * wgt.double()).sum().backward()
o32 = orc.tsrn_tl_trans_forward(sd, x, tp, training=True, stn=False, dropout_p=0.0)[0]
(o32 * wgt).sum().backward()
net = net.cuda()
res = {}
for fused in (True, False):
    ops._declayer_enabled = fused
    net.zero_grad(set_to_none=True)
    out, aux = net(x.cuda(), tp.cuda())
    (out * wgt.cuda()).sum().backward()
    res[fused] = {n: p.grad.detach().double().cpu() for n, p in net.named_parameters() if p.grad is not None}
    print("fused" if fused else "opbyop", "out err vs fp64 %.3e" % ((out.detach().double().cpu() - o64.detach()).abs().max().item() / o64.abs().max().item()))
print("%-70s %10s %10s %10s %10s" % ("param", "fused", "opbyop", "cpu fp32", "fus-vs-op"))
for n in res[True]:
    g = sd64[n].grad
    if g is None:
        continue
    den = max(g.norm().item(), 1e-30)
    e = [(res[f][n] - g).norm().item() / den for f in (True, False)]
    e32 = (sd[n].grad.double() - g).norm().item() / den
    d = (res[True][n] - res[False][n]).norm().item() / den
    if "infoGen" in n or max(e) > 3e-4:
        print("%-70s %10.2e %10.2e %10.2e %10.2e" % (n, e[0], e[1], e32, d))
