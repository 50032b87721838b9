"""Parity of the fused-split row-panel kernels (csrc/tc4_rows.cu: tatt_rows_gemm / tatt_rows_wgrad) against
fp64 CPU matmuls.  fp32-parity mode (bf16 hi/lo split, 3 MMAs): max-abs error <= 2e-4 of the output scale for the
forward / data-gradient GEMMs, 5e-4 for the pixel-axis reductions; bf16 mode: 2e-2 (stated per test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def close(a, b, tol, name=""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item() / scale
    assert err <= tol, "%s: rel-max err %.3e > %.1e (scale %.3e)" % (name, err, tol, scale)


def g(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=gen) * scale


@pytest.fixture(autouse=True)
def _rows_on():
    from tatt_b200 import ops
    ops.set_rows_kernels(True)
    ops.set_precision("fp32")
    yield
    ops.set_precision("fp32")


@pytest.mark.parametrize("M", [128, 1000, 50000])            # 50000 rows: 391 tiles > 2 CTAs x 148 SMs
@pytest.mark.parametrize("K,N", [(64, 64), (64, 192), (64, 128), (128, 64), (192, 64)])
def test_rows_gemm(M, K, N):
    from tatt_b200 import ops
    x, w, b = g(M, K), g(N, K, seed=1, scale=0.2), g(N, seed=2)
    xd, wd, bd = x.to(dev()), w.to(dev()), b.to(dev())
    ref = x.double() @ w.double().t() + b.double()
    assert ops.rows_gemm_ok(M, K, N)
    y = ops.linear_fwd(xd, wd, bd)
    close(y, ref, 2e-4, "rows_gemm")
    close(ops.linear_fwd(xd, wd, bd, relu=True), ref.clamp_min(0), 2e-4, "rows_gemm relu")
    # accumulate into a strided view, no bias
    big = torch.zeros(M, 2 * N, device=dev())
    ops.linear_fwd(xd, wd, bd, out=big[:, N:])
    ops.linear_fwd(xd, wd, None, out=big[:, N:], accumulate=True)
    close(big[:, N:], 2 * ref - b.double(), 2e-4, "rows_gemm accumulate")
    assert big[:, :N].abs().max().item() == 0
    # data gradient: dy[M,N] @ w[N,K]   (the kernel reads w transposed)
    if ops.rows_gemm_ok(M, N, K):
        dy = g(M, N, seed=3)
        close(ops.linear_bwd_data(dy.to(dev()), wd), dy.double() @ w.double(), 2e-4, "rows dgrad")


def test_rows_gemm_two_sources():
    """1x1 conv over a channel concatenation that never exists in memory (tsrn.py:902 + 1075)"""
    from tatt_b200 import ops
    M = 3000
    a, b2, w, bias = g(M, 64), g(M, 64, seed=5), g(64, 128, seed=1, scale=0.2), g(64, seed=2)
    ad, bd = a.to(dev()), torch.zeros(M, 96, device=dev())
    bd[:, 16:80] = b2.to(dev())                               # second source is a strided view
    y = ops.rows_gemm([ad, bd[:, 16:80]], w.to(dev()), bias.to(dev()), torch.empty(M, 64, device=dev()))
    ref = torch.cat([a, b2], 1).double() @ w.double().t() + bias.double()
    close(y, ref, 2e-4, "rows_gemm cat")


@pytest.mark.parametrize("M", [128, 777, 50000, 100003])   # >= 4 x 148 row blocks: the warp-specialised kernel
@pytest.mark.parametrize("N,K", [(64, 64), (64, 128), (192, 64), (128, 64)])
def test_rows_wgrad(M, N, K):
    from tatt_b200 import ops
    dy, x = g(M, N, seed=3), g(M, K)
    assert ops.rows_wgrad_ok(M, N, K)
    dW, db = ops.linear_bwd_weight_rows(dy.to(dev()), x.to(dev()), True)
    close(dW, dy.double().t() @ x.double(), 5e-4, "rows_wgrad dW")
    close(db, dy.double().sum(0), 5e-4, "rows_wgrad db")
    dW2, none = ops.linear_bwd_weight_rows(dy.to(dev()), x.to(dev()), False)
    assert none is None
    close(dW2, dy.double().t() @ x.double(), 5e-4, "rows_wgrad dW (no bias)")


@pytest.mark.parametrize("M", [5000, 90001])
def test_rows_wgrad_parts_and_gates(M):
    from tatt_b200 import ops
    dy, p0, p1 = g(M, 64, seed=3), g(M, 64), g(M, 64, seed=7)
    dW, db = ops.linear_bwd_weight_rows_parts(dy.to(dev()), [p0.to(dev()), p1.to(dev())], True)
    close(dW, dy.double().t() @ torch.cat([p0, p1], 1).double(), 5e-4, "wgrad cat")
    close(db, dy.double().sum(0), 5e-4, "wgrad cat bias")
    # recurrent weights of both GRU directions from the gate tensor [M][320] and dGH [M][192]
    gates, dgh = g(M, 320, seed=9), g(M, 192, seed=11)
    dwhh = torch.empty(2, 96, 32, device=dev())
    dbhh = torch.empty(192, device=dev())
    ops.rows_wgrad(gates.to(dev()), (128, 288), ops._blocks64(dgh.to(dev())), dwhh, 2, 96, 32, True, rb=32, cb=96,
                   colsum_src=2, dbias=dbhh)
    for d in range(2):
        ref = dgh[:, 96 * d:96 * d + 96].double().t() @ gates[:, 160 * d + 128:160 * d + 160].double()
        close(dwhh[d], ref, 5e-4, "dW_hh[%d]" % d)
    close(dbhh, dgh.double().sum(0), 5e-4, "db_hh")


@pytest.mark.parametrize("M", [4096, 131072])
def test_rows_bf16_mode(M):
    from tatt_b200 import ops
    x, w, b, dy = g(M, 64), g(192, 64, seed=1, scale=0.2), g(192, seed=2), g(M, 192, seed=3)
    ops.set_precision("bf16")
    try:
        y = ops.linear_fwd(x.to(dev()), w.to(dev()), b.to(dev()))
        dW, db = ops.linear_bwd_weight_rows(dy.to(dev()), x.to(dev()), True)
    finally:
        ops.set_precision("fp32")
    close(y, x.double() @ w.double().t() + b.double(), 2e-2, "bf16 rows_gemm")
    close(dW, dy.double().t() @ x.double(), 2e-2, "bf16 rows_wgrad")
    close(db, dy.double().sum(0), 5e-4, "bf16 rows_wgrad bias (fp32 sums)")


def test_rows_toggle_matches_plane_path():
    """the tape picks the fused-split kernels by default; with TATT_ROWS=0 semantics the plane path gives the same
    layer gradients (both within the fp32-parity tolerance of each other)"""
    from tatt_b200 import ops
    from tatt_b200.tape import Tape
    M = 2048
    x, w, b, dy = (t.to(dev()) for t in (g(M, 64), g(64, 64, seed=1, scale=0.2), g(64, seed=2), g(M, 64, seed=3)))
    res = []
    for on in (True, False):
        ops.set_rows_kernels(on)
        t = Tape(True)
        y = t.linear(x, w, b, relu=True)
        t.seed(y, dy)
        t.backward()
        res.append((y, t.grad(w), t.grad(b), t.grad(x)))
    ops.set_rows_kernels(True)
    for a, c, n in zip(res[0], res[1], ("y", "dW", "db", "dx")):
        close(a, c, 3e-4, "toggle " + n)
