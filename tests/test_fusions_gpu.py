"""Op-level checks of the round-2 fusions against their unfused twins (which are pinned to torch by tests/test_ops_gpu.py):
BatchNorm statistics from the convolution epilogue, BatchNorm backward / BatchNorm apply / PixelShuffle writing a
convolution's bf16 operand planes directly, the deterministic gradient-norm reduction, the packed-parameter copy."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def g(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def test_conv_epilogue_bn_statistics_match_the_separate_pass():
    from tatt_b200 import ops
    x = g(3, 8, 128, 64)                                   # NHWC, served by the persistent TMA kernel (W % 128 == 0)
    w, b = g(64, 64, 3, 3, seed=1, scale=1 / 24.0), g(64, seed=2)
    st = {}
    y = ops.conv2d_fwd(x, w, b, 1, stats=st)
    assert "acc" in st, "the statistics variant was not used"
    y0 = ops.conv2d_fwd(x, w, b, 1)
    assert torch.equal(y, y0)
    rm, rv = torch.zeros(64, device=DEV), torch.ones(64, device=DEV)
    rm0, rv0 = rm.clone(), rv.clone()
    P = y.numel() // 64
    m1, i1 = ops.bn_finalize(st["acc"], P, 64, 1e-5, 0.1, rm, rv)
    m0, i0 = ops.bn_stats(y0.view(-1, 64), 1e-5, 0.1, rm0, rv0)
    for a, c in ((m1, m0), (i1, i0), (rm, rm0), (rv, rv0)):
        assert rel(a, c) <= 1e-5
    yr = y.view(-1, 64).double()
    assert rel(m1.double(), yr.mean(0)) <= 1e-5 and rel(i1.double(), 1 / (yr.var(0, unbiased=False) + 1e-5).sqrt()) <= 1e-5


@pytest.mark.parametrize("shape", [(2, 8, 128), (3, 16, 64)])     # TMA data-gradient kernel / generic engine with valid planes
def test_bn_backward_writes_conv_dy_planes(shape):
    """tape-level: conv -> BatchNorm(mish) with the fused plane hand-over vs TATT_CONV_SHARE-less reference path"""
    from tatt_b200 import ops
    from tatt_b200.tape import Tape
    n, h, wd = shape
    x = g(n, h, wd, 64)
    conv = torch.nn.Conv2d(64, 64, 3, padding=1).to(DEV)
    bn = torch.nn.BatchNorm2d(64).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
    gy = g(n, h, wd, 64, seed=5)
    res = []
    for fused in (True, False):
        tape = Tape(True)
        y = tape.conv(x, conv.weight, conv.bias, 1, bn_next=fused)
        z = tape.batchnorm(y, bn, ops.ACT_MISH, True)
        if fused:
            assert tape._conv_keep == {}                                   # the BatchNorm picked the hand-over up
        tape.seed(z, gy)
        tape.backward()
        res.append([tape.grad(t).clone() for t in (x, conv.weight, bn.weight, bn.bias)] + [tape.grad(conv.bias).clone()])
    for a, c in zip(res[0][:4], res[1][:4]):
        assert rel(a, c) <= 2e-4
    assert res[0][4].abs().max().item() == 0.0                # exact zero: sum_p dX_bn == 0 behind a train-mode BatchNorm
    assert res[1][4].abs().max().item() <= 1e-3 * res[1][1].abs().max().item() * 64   # the unfused value is rounding noise


def test_producers_write_conv_operand_planes():
    from tatt_b200 import ops
    from tatt_b200.tape import Tape
    # BatchNorm + mish -> conv3x3
    x = g(2, 8, 128, 64)
    bn = torch.nn.BatchNorm2d(64).to(DEV).eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
    conv = torch.nn.Conv2d(64, 64, 3, padding=1).to(DEV)
    outs = []
    for fused in (True, False):
        tape = Tape(False)
        z = tape.batchnorm(x, bn, ops.ACT_MISH, False, planes_for=conv.weight if fused else None)
        assert (z.stride() == (0, 0, 0, 0)) == fused            # shape carrier only
        outs.append(tape.conv(z, conv.weight, conv.bias, 1))
    assert rel(outs[0], outs[1]) <= 1e-6
    # PixelShuffle + mish -> conv9x9 (kx-expansion path)
    u = g(2, 8, 32, 256, seed=3)
    fin = torch.nn.Conv2d(64, 4, 9, padding=4).to(DEV)
    outs = []
    for fused in (True, False):
        tape = Tape(False)
        s = tape.pixshuf_mish(u, planes_for=fin.weight if fused else None)
        outs.append(tape.conv(s, fin.weight, fin.bias, 4))
    assert rel(outs[0], outs[1]) <= 1e-6


def test_sqnorm_det_is_deterministic_and_right():
    from tatt_b200 import _cabi, ops
    x = g(7_570_638, seed=9)
    ws = torch.empty(1024, device=DEV)
    outs = []
    for _ in range(3):
        o = torch.empty(1, device=DEV)
        _cabi.call("tatt_sqnorm_det", x.data_ptr(), x.numel(), o.data_ptr(), ws.data_ptr(), 1024, ops._stream())
        outs.append(o.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert abs(outs[0].item() - x.double().pow(2).sum().item()) <= 1e-5 * outs[0].item()


def test_packed_parameter_copy():
    from tatt_b200 import ops
    srcs = [g(96, 64, seed=1), g(96, seed=2), g(2, 96, 32, seed=3), g(40000, seed=4)]
    views, flat = ops.packed(srcs, srcs[0], want_flat=True)
    for v, s in zip(views, srcs):
        assert v.shape == s.shape and torch.equal(v, s)
    assert flat.numel() == sum((s.numel() + 3) // 4 * 4 for s in srcs)
    again = ops.packed(srcs, srcs[0])                          # cached table, fresh buffer
    assert all(torch.equal(a, s) for a, s in zip(again, srcs)) and again[0].data_ptr() != views[0].data_ptr()
