import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
for (M, N, K) in [(512, 64, 64), (512, 64, 576), (128, 3072, 1024), (4096, 192, 64)]:
    a = torch.randn(M, K); b = torch.randn(N, K)
    ref = a.double() @ b.double().t()
    absdot = a.abs().double() @ b.abs().double().t()
    ad, bd = a.to(dev), b.to(dev)
    ctc = ops.linear_fwd(ad, bd, None).cpu().double()
    with ops.full_fp32():
        cff = ops.linear_fwd(ad, bd, None).cpu().double()
    c32 = (a @ b.t()).double()
    # emulate bf16x3 on CPU in fp64 accumulate
    ah = a.bfloat16().float(); al = (a - ah).bfloat16().float(); bh = b.bfloat16().float(); bl = (b - bh).bfloat16().float()
    emu = (ah.double() @ bh.double().t()) + (ah.double() @ bl.double().t()) + (al.double() @ bh.double().t())
    f = lambda c: ((c - ref).abs() / absdot).max().item()
    g = lambda c: ((c - ref).norm() / ref.norm()).item()
    print((M, N, K), "err/sum|ab|: tc %.2e ffma %.2e cpu32 %.2e emu(bf16x3,exact acc) %.2e | relL2: tc %.2e ffma %.2e emu %.2e" % (
        f(ctc), f(cff), f(c32), f(emu), g(ctc), g(cff), g(emu)))
