"""Time the per-pixel linear layers at the bench shape (P = 64*32*128 rows): fused-split row kernels
(csrc/tc4_rows.cu) vs the pre-split plane path (split pass + tc2 GEMM).  CUDA events, L2 flushed between launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import ops

dev = "cuda:0"
P = 64 * 32 * 128
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def timed(fn, n=8):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def planes_fwd(x, w, b):
    xP = ops.split_matrix(x)
    return ops.linear_fwd(x, w, b, xP=xP)


def planes_bwd(dy, x, w):
    db = ops.empty(dy.shape[1], like=dy)
    dyP = ops.split_matrix(dy, db)
    xP = ops.split_matrix(x)
    ops.linear_bwd_weight(dy, x, dyP=dyP, xP=xP)
    return ops.linear_bwd_data(dy, w, dyP=dyP)


def rows_bwd(dy, x, w):
    ops.linear_bwd_weight_rows(dy, x, True)
    return ops.linear_bwd_data(dy, w)


for K, N in ((64, 64), (64, 192), (128, 64)):
    x = torch.randn(P, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.1
    b = torch.randn(N, device=dev)
    dy = torch.randn(P, N, device=dev)
    gb = (P * (K + N) * 4) / 1e9
    ops.set_rows_kernels(True)
    t_rows = timed(lambda: ops.linear_fwd(x, w, b))
    t_rows_b = timed(lambda: rows_bwd(dy, x, w))
    t_wg = timed(lambda: ops.linear_bwd_weight_rows(dy, x, True))
    ops.set_rows_kernels(False)
    t_pl = timed(lambda: planes_fwd(x, w, b))
    t_pl_b = timed(lambda: planes_bwd(dy, x, w))
    ops.set_rows_kernels(True)
    print("K=%3d N=%3d  fwd: rows %.1f us (%.2f TB/s algorithmic) | split+planes %.1f us   bwd(dW+db+dx): rows %.1f us "
          "(wgrad alone %.1f us = %.2f TB/s) | planes %.1f us" % (K, N, t_rows, gb / t_rows * 1e3, t_pl, t_rows_b, t_wg,
                                                                gb / t_wg * 1e3, t_pl_b))
