"""Times tatt_tp_declayer_fwd alone at the bench shape (N = 64 samples x 4096 queries, 26 keys): train mode with the side
outputs and dropout 0.1, train mode without dropout, and eval mode.  CUDA events around 20 calls after warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops

N, Lq, Lk = int(os.environ.get("N", 64)), 4096, 26
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
P = N * Lq
r = lambda *s: torch.randn(*s, device=dev, generator=g)
ins = [r(P, 64), r(P, 64), r(N * Lk, 64), r(N * Lk, 64)] + [r(64, 64) * 0.1 for _ in range(4)] + [r(64) * 0.1 for _ in range(10)]
rng = ops.DeviceRNG.get(dev).snapshot()


def run(train, pd):
    new = lambda: torch.empty(P, 64, device=dev)
    outs = [new(), new(), torch.empty(N, Lq, Lk, device=dev)]
    if train:
        outs += [new() for _ in range(8)] + [torch.empty(2, P, device=dev) for _ in range(3)]
    f = lambda: ops.declayer_fwd(ins, outs, train, N, Lq, Lk, pd, rng if pd else None, (10, 11, 12, 13) if pd else None)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3


print("declayer fwd N=%d: train+dropout %.1f us, train p=0 %.1f us, eval %.1f us" % (
    N, run(True, (0.1, 0.1, 0.1, 0.1)), run(True, None), run(False, None)))

# the attention backward the decoder layers still use (csrc/attn.cu:mha_bwd_kernel)
q, k, v, do = ins[0], ins[2], ins[3], ins[1]


def time_mha_bwd(pd):
    f = lambda: ops.mha_bwd(q, k, v, do, N, Lq, Lk, pd, rng if pd else None, 10)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3


print("mha_bwd N=%d: dropout 0.1 %.1f us, p=0 %.1f us" % (N, time_mha_bwd(0.1), time_mha_bwd(0.0)))
