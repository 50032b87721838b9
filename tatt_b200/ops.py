"""Tensor-level wrappers over the C-ABI (allocation + pointer plumbing only; every FLOP runs in the
hand-written CUDA kernels of tatt_b200/csrc).  All tensors are fp32 CUDA tensors; feature maps are
channels-last [N,H,W,C]; "rows" tensors are [P, C] with unit inner stride."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi

Tensor = torch.Tensor
ACT_NONE, ACT_RELU, ACT_MISH = 0, 1, 2
F_ACCUM, F_RELU, F_SPLITK, F_ZEROC, F_FP32, F_APLANES, F_BPLANES, F_BF16, F_A_VALID = 1, 2, 4, 64, 128, 256, 512, 1024, 2048
_precision_flag = 0      # OR-ed into every GEMM / conv call; F_FP32 inside `full_fp32()`


def set_precision(mode: str) -> None:
    """'fp32' (default): fp32 parity on tensor cores via the bf16 hi/lo split (3 MMAs per k-step);
    'bf16': operands rounded to bf16, one MMA per k-step, fp32 accumulation (BASELINE configs 3/4);
    'ffma': every GEMM / convolution on the fp32 CUDA-core kernels."""
    global _precision_flag
    _precision_flag = {"fp32": 0, "bf16": F_BF16, "ffma": F_FP32}[mode]


def get_precision() -> str:
    return "ffma" if _precision_flag & F_FP32 else ("bf16" if _precision_flag & F_BF16 else "fp32")


class full_fp32:
    """Run the enclosed GEMMs / convolutions on the fp32 FFMA kernels instead of the tcgen05 bf16x3 path
    (relative error 2^-24 instead of 2^-16).  Used for the STN head, whose BatchNorm1d over a handful of
    samples amplifies operand rounding by orders of magnitude."""

    def __enter__(self):
        global _precision_flag
        self._old = _precision_flag
        _precision_flag = F_FP32

    def __exit__(self, *a):
        global _precision_flag
        _precision_flag = self._old


class precision_flag_scope:
    """Run the enclosed GEMMs / convolutions with the given precision flag (0 = fp32 parity on tensor cores,
    F_BF16, F_FP32) regardless of the global mode."""

    def __init__(self, flag: int):
        self.flag = flag

    def __enter__(self):
        global _precision_flag
        self._old = _precision_flag
        _precision_flag = self.flag

    def __exit__(self, *a):
        global _precision_flag
        _precision_flag = self._old


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---- side stream for work that is off the critical path of the backward pass (weight gradients): the data-gradient
# chain of a stage is a serial sequence of mostly latency- / HBM-latency-bound kernels, the weight-gradient kernels
# only feed the optimizer.  Inside `with side_stream():` launches go to a per-device second stream that first waits
# for everything enqueued on the caller's stream so far; `join_side()` makes the caller's stream wait for it.  Both
# are plain event edges, so a CUDA-graph capture turns them into a fork / join of the graph.  TATT_SIDE=0 disables it.
_side_enabled = os.environ.get("TATT_SIDE", "1") != "0"
_side_streams: dict = {}
_side_dirty: dict = {}


class side_stream:
    def __enter__(self):
        self.ctx = None
        if not _side_enabled:
            return self
        cur = torch.cuda.current_stream()
        dev = cur.device_index
        s = _side_streams.get(dev)
        if s is None:
            s = _side_streams[dev] = torch.cuda.Stream(device=dev)
        if cur == s:                                  # nested use: already on the side stream
            return self
        ev = torch.cuda.Event()
        ev.record(cur)
        s.wait_event(ev)
        _side_dirty[dev] = True
        self.ctx = torch.cuda.stream(s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        return False


_side_hold: dict = {}                 # device index -> tensors (one caller thread per device, like the streams)


def hold_until_join(*tensors) -> None:
    """keep main-stream tensors that side-stream kernels read alive until the next join_side() on their device"""
    for t in tensors:
        if t is not None:
            _side_hold.setdefault(t.device.index, []).append(t)


def join_side() -> None:
    """the current stream waits for everything launched on its device's side stream (no-op when called from code that
    itself runs on the side stream)"""
    cur = torch.cuda.current_stream()
    dev = cur.device_index
    if cur == _side_streams.get(dev):
        return
    if _side_dirty.get(dev):
        ev = torch.cuda.Event()
        ev.record(_side_streams[dev])
        cur.wait_event(ev)
        _side_dirty[dev] = False
    _side_hold.pop(dev, None)


# Bumped whenever parameters are updated through raw pointers (the fused clip+Adam kernel writes the flat parameter
# buffer without touching torch's version counters); caches derived from weights key on it (tsrn.TPInterpreter).
_weights_epoch = 0


def weights_epoch() -> int:
    return _weights_epoch


def bump_weights_epoch() -> None:
    global _weights_epoch
    _weights_epoch += 1


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: Tensor, name: str = "tensor") -> Tensor:
    if not t.is_cuda:
        raise RuntimeError("tatt_b200: %s must be a CUDA tensor (the hot path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("tatt_b200: %s must be float32, got %s" % (name, t.dtype))
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError("tatt_b200: %s lives on %s but the current device is cuda:%d; kernels launch on the current "
                           "device's stream -- wrap the call in torch.cuda.device(...)" % (
                               name, t.device, torch.cuda.current_device()))
    return t


def _rows(t: Tensor) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a 2-D tensor with unit inner stride."""
    assert t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1), (t.shape, t.stride())
    return t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


def _ws(like: Tensor, nelems: int):
    """Scratch for the bf16 hi/lo operand planes (2 planes x 2 bytes per element), or (None, 0)."""
    if _precision_flag & F_FP32:
        return None, 0
    nbytes = 4 * nelems + 1024
    return torch.empty(nbytes, dtype=torch.uint8, device=like.device), nbytes


def empty(*shape, like: Tensor) -> Tensor:
    return torch.empty(*shape, dtype=torch.float32, device=like.device)


# ------------------------------------------------------------------------------------------ GEMM
def gemm(amode: int, bmode: int, A: Tensor, lda: int, B: Tensor, ldb: int, C: Tensor, ldc: int,
         bias: Optional[Tensor], M: int, N: int, K: int, flags: int = 0, batch: int = 1, sA: int = 0, sB: int = 0,
         sC: int = 0, sBias: int = 0, loA: int = 0, loB: int = 0) -> None:
    """flags & F_APLANES / F_BPLANES: A / B are bf16 hi planes (see Planes), lo plane at +loA / +loB elements."""
    ws, wsb = (None, 0)
    if M >= 32 and K >= 32 and N > 4:
        ne = 0 if (flags & F_APLANES and flags & F_BPLANES) else batch * (_r8(M) * _r8(K) + _r8(N) * _r8(K))
        if ne:
            ws, wsb = _ws(C, ne)
    _cabi.call("tatt_gemm", amode, bmode, _p(A), lda, _p(B), ldb, _p(C), ldc, _p(bias), M, N, K, batch, sA, sB,
               sC, sBias, flags | _precision_flag, loA, loB, _p(ws), wsb, _stream())


class Planes:
    """bf16 hi/lo planes of an fp32 operand reused by many GEMMs: one bf16 tensor [2, *shape] (hi, lo)."""

    def __init__(self, shape, like: Tensor, zero: bool = False):
        self.t = torch.empty((2,) + tuple(shape), dtype=torch.bfloat16, device=like.device)
        self.lo_off = self.t[0].numel()
        if zero:
            _cabi.call("tatt_memset0", _p(self.t), self.t.numel() * 2, _stream())

    def split_from(self, src2d: Tensor, dst_index=None, transpose: bool = False) -> None:
        """src2d [rows, cols] fp32 -> planes[dst_index] ([rows, cols] or, transposed, [cols, rows]); cols % 8 == 0"""
        rows, cols, ld = _rows(src2d)
        hi = self.t[0] if dst_index is None else self.t[0][dst_index]
        lo = self.t[1] if dst_index is None else self.t[1][dst_index]
        assert (rows if transpose else cols) % 8 == 0
        _cabi.call("tatt_split_bf16", _p(src2d), ld, rows, cols, 1 if transpose else 0, _p(hi), _p(lo), None,
                   _stream())


class MatPlanes:
    """bf16 hi/lo planes of a [rows, cols] fp32 matrix (row stride `ld`, lo plane `lo_off` elements after hi).
    One split pass serves every GEMM that consumes the matrix (forward, data-gradient and weight-gradient)."""

    def __init__(self, hi: Tensor, lo_off: int, ld: int):
        self.hi, self.lo_off, self.ld = hi, lo_off, ld
        self.rows, self.cols = hi.shape

    def slice_cols(self, a: int, b: int) -> "MatPlanes":
        assert a % 8 == 0
        return MatPlanes(self.hi[:, a:b], self.lo_off, self.ld)


def split_matrix(x2: Tensor, colsum_out: Optional[Tensor] = None) -> Optional[MatPlanes]:
    """Split once for the tcgen05 engine; None when the matrix is not eligible (fp32 mode, odd shapes) -- callers
    then use the plain fp32 entry points (and compute the column sum separately)."""
    rows, cols, ld = _rows(x2)
    if (_precision_flag & F_FP32) or cols % 8 or cols < 32 or rows < 32:
        return None
    t = torch.empty(2, rows, cols, dtype=torch.bfloat16, device=x2.device)
    _cabi.call("tatt_split_bf16", _p(x2), ld, rows, cols, 0, _p(t[0]), _p(t[1]), _p(colsum_out), _stream())
    return MatPlanes(t[0], rows * cols, cols)


# ---- row-panel kernels with the operand split fused into the loader (csrc/tc4_rows.cu); TATT_ROWS=0 disables them
_rows_enabled = os.environ.get("TATT_ROWS", "1") != "0"


def set_rows_kernels(on: bool) -> None:
    global _rows_enabled
    _rows_enabled = bool(on)


def _al16(*ts) -> bool:
    return all(t is None or t.data_ptr() % 16 == 0 for t in ts)


def rows_gemm_ok(M: int, K: int, N: int) -> bool:
    """Y[M,N] = X[M,K] W^T is served by tatt_rows_gemm"""
    return (_rows_enabled and not (_precision_flag & F_FP32) and M >= 128 and K in (64, 128, 192)
            and N in (64, 128, 192) and (K // 64) * N <= 192)


def rows_wgrad_ok(M: int, N: int, K: int) -> bool:
    """dW[N,K] = dY[M,N]^T X[M,K] is served by tatt_rows_wgrad"""
    return (_rows_enabled and not (_precision_flag & F_FP32) and M >= 128
            and ((N == 64 and K in (64, 128, 192)) or (K == 64 and N in (128, 192))))


def _blocks64(x: Tensor) -> List[Tensor]:
    return [x[:, 64 * i:64 * (i + 1)] for i in range(x.shape[1] // 64)]


def rows_gemm(xs: Sequence[Tensor], w: Tensor, b: Optional[Tensor], out: Tensor, wtrans: bool = False,
              accumulate: bool = False, relu: bool = False) -> Tensor:
    """out[M,N] (=|+=) act(sum_i xs[i][M,64] @ Wblock_i^T + b); W is [N, 64*len(xs)], or its transpose (wtrans)."""
    M, KB = xs[0].shape[0], len(xs)
    N = out.shape[1]
    args = []
    for i in range(3):
        if i < KB:
            r, c, ld = _rows(xs[i])
            assert r == M and c == 64, (xs[i].shape,)
            args += [_p(xs[i]), ld]
        else:
            args += [None, 0]
    wr, wc, ldw = _rows(w)
    assert (wr, wc) == ((64 * KB, N) if wtrans else (N, 64 * KB)), (w.shape, N, KB, wtrans)
    _, _, ldo = _rows(out)
    fl = (F_ACCUM if accumulate else 0) | (F_RELU if relu else 0) | (_precision_flag & F_BF16)
    _cabi.call("tatt_rows_gemm", *args, _p(w), ldw, 1 if wtrans else 0, _p(b), _p(out), ldo, M, N, KB, fl, _stream())
    return out


def rows_wgrad(A: Tensor, a_offs: Tuple[int, int], Bs: Sequence[Tensor], out: Tensor, nb: int, ni: int, nj: int,
               transpose: bool, rb: int = 0, cb: int = 0, colsum_src: int = 0, dbias: Optional[Tensor] = None) -> None:
    """D[64, 64*len(Bs)] = A[:, cols]^T [B0|B1|B2]; out[b,i,j] = D[b*rb + (j if T else i), b*cb + (i if T else j)];
    dbias = column sums of A (colsum_src 1) or of the B blocks (2)."""
    M = A.shape[0]
    lda = A.stride(0)
    args = []
    for i in range(3):
        if i < len(Bs):
            r, c, ld = _rows(Bs[i])
            assert r == M and c == 64
            args += [_p(Bs[i]), ld]
        else:
            args += [None, 0]
    nbytes = _cabi.lib().tatt_rows_wgrad_ws_bytes()
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    assert out.is_contiguous() and out.numel() == nb * ni * nj
    _cabi.call("tatt_rows_wgrad", _p(A), lda, a_offs[0], a_offs[1], *args, len(Bs), M, colsum_src, _p(out), nb, ni,
               nj, 1 if transpose else 0, rb, cb, _p(dbias), 0 if dbias is None else dbias.numel(), _p(ws), nbytes,
               _precision_flag & F_BF16, _stream())


def linear_bwd_weight_rows(dy: Tensor, x: Tensor, want_bias: bool, out: Optional[Tensor] = None):
    """-> (dW[N,K] = dy^T x, db[N] | None) in one pass over dy and x (no pre-split planes)"""
    M, N, _ = _rows(dy)
    _, K, _ = _rows(x)
    if out is None:
        out = empty(N, K, like=dy)
    db = empty(N, like=dy) if want_bias else None
    if N == 64:
        rows_wgrad(dy, (0, 32), _blocks64(x), out, 1, 64, K, False, colsum_src=1 if want_bias else 0, dbias=db)
    else:
        assert K == 64
        rows_wgrad(x, (0, 32), _blocks64(dy), out, 1, N, 64, True, colsum_src=2 if want_bias else 0, dbias=db)
    return out, db


def linear_bwd_weight_rows_parts(dy: Tensor, parts: Sequence[Tensor], want_bias: bool):
    """dW[64, 64*len(parts)] = dy^T [part_0 | part_1 | ...] without materialising the concatenation"""
    out = empty(64, 64 * len(parts), like=dy)
    db = empty(64, like=dy) if want_bias else None
    rows_wgrad(dy, (0, 32), list(parts), out, 1, 64, 64 * len(parts), False, colsum_src=1 if want_bias else 0, dbias=db)
    return out, db


def linear_fwd(x: Tensor, w: Tensor, b: Optional[Tensor], out: Optional[Tensor] = None, accumulate: bool = False,
               relu: bool = False, xP: Optional[MatPlanes] = None) -> Tensor:
    """out[M,N] (=|+=) x[M,K] @ w[N,K]^T + b"""
    M, K, ldx = _rows(x)
    N, K2, ldw = _rows(w)
    assert K == K2, (x.shape, w.shape)
    if out is None:
        out = empty(M, N, like=x)
    _, _, ldo = _rows(out)
    if xP is None and rows_gemm_ok(M, K, N) and ldx % 4 == 0 and ldw % 4 == 0 and ldo % 4 == 0 and _al16(x, w, b, out):
        return rows_gemm(_blocks64(x), w, b, out, accumulate=accumulate, relu=relu)
    fl = (F_ACCUM if accumulate else 0) | (F_RELU if relu else 0)
    if xP is not None and N > 4:
        gemm(0, 1, xP.hi, xP.ld, w, ldw, out, ldo, b, M, N, K, fl | F_APLANES, loA=xP.lo_off)
    else:
        gemm(0, 1, x, ldx, w, ldw, out, ldo, b, M, N, K, fl)
    return out


def linear_bwd_data(dy: Tensor, w: Tensor, out: Optional[Tensor] = None, accumulate: bool = False,
                    dyP: Optional[MatPlanes] = None) -> Tensor:
    """out[M,K] (=|+=) dy[M,N] @ w[N,K]"""
    M, N, lddy = _rows(dy)
    N2, K, ldw = _rows(w)
    assert N == N2, (dy.shape, w.shape)
    if out is None:
        out = empty(M, K, like=dy)
    _, _, ldo = _rows(out)
    fl = F_ACCUM if accumulate else 0
    if dyP is None and rows_gemm_ok(M, N, K) and lddy % 4 == 0 and ldo % 4 == 0 and _al16(dy, w, out):
        return rows_gemm(_blocks64(dy), w, None, out, wtrans=True, accumulate=accumulate)
    if dyP is not None and K > 4:
        gemm(0, 0, dyP.hi, dyP.ld, w, ldw, out, ldo, None, M, K, N, fl | F_APLANES, loA=dyP.lo_off)
    else:
        gemm(0, 0, dy, lddy, w, ldw, out, ldo, None, M, K, N, fl)
    return out


def linear_bwd_weight(dy: Tensor, x: Tensor, out: Optional[Tensor] = None, dyP: Optional[MatPlanes] = None,
                      xP: Optional[MatPlanes] = None) -> Tensor:
    """out[N,K] += dy[M,N]^T @ x[M,K]  (split-K atomics; `out` must be zero-filled if given)"""
    M, N, lddy = _rows(dy)
    M2, K, ldx = _rows(x)
    assert M == M2, (dy.shape, x.shape)
    flags = F_SPLITK
    if out is None:
        out = empty(N, K, like=dy)
        flags |= F_ZEROC
    _, _, ldo = _rows(out)
    ok = N >= 32 and K > 4 and M >= 32
    A, lda, loA = (dyP.hi, dyP.ld, dyP.lo_off) if (dyP is not None and ok) else (dy, lddy, 0)
    B, ldb, loB = (xP.hi, xP.ld, xP.lo_off) if (xP is not None and ok) else (x, ldx, 0)
    if dyP is not None and ok:
        flags |= F_APLANES
    if xP is not None and ok:
        flags |= F_BPLANES
    gemm(1, 0, A, lda, B, ldb, out, ldo, None, N, K, M, flags, loA=loA, loB=loB)
    return out


def colsum(x: Tensor, out: Optional[Tensor] = None, zero_first: bool = True) -> Tensor:
    P, C, ldx = _rows(x)
    if out is None:
        out = empty(C, like=x)
    _cabi.call("tatt_colsum", _p(x), ldx, _p(out), P, C, 1 if zero_first else 0, _stream())
    return out


# ------------------------------------------------------------------------------------------ conv
def _pad4(c: int) -> int:
    return (c + 3) // 4 * 4


def conv_pack(w: Tensor, cin_p: int, cout_p: int, flip: bool) -> Tensor:
    co, ci, kh, kw = w.shape
    wt = empty(kh * kw * (cout_p if flip else cin_p), (cin_p if flip else cout_p), like=w)
    _cabi.call("tatt_conv_weight_pack", _p(w.contiguous()), _p(wt), co, ci, kh, kw, cin_p, cout_p, 1 if flip else 0,
               _stream())
    return wt


def _conv_ws(x: Tensor, nx: int, ndy: int, extra: int):
    """One scratch buffer per convolution layer, laid out so that the bf16 hi/lo planes written by one pass are
    reused by the next (F_A_VALID): [X planes: 2 x nx bf16][dY planes: 2 x ndy bf16][weights planes / split-K tiles].
    forward: workspace base = 0 (A = X); weight gradient: base = 0 (A = X valid, B = dY split here);
    data gradient: base = the dY planes (A = dY valid)."""
    if _precision_flag & F_FP32:
        return None
    nbytes = 4 * (_r8(nx) + _r8(ndy) + extra) + 1024
    return torch.empty(nbytes, dtype=torch.uint8, device=x.device)


def conv_fwd_ws_elems(cin_p: int, cout_p: int, kh: int, kw: int, P: int) -> Tuple[int, int]:
    """(dY-plane elements, extra elements) of a convolution layer's workspace (see _conv_ws)"""
    if cout_p == 4 and kw > 1 and cin_p >= 16:
        return P * _r8(kw * 4), kh * cin_p * _r8(kw * 4) + 148 * kh * cin_p * _r8(kw * 4)
    return P * _r8(cout_p), (max(kh * kw * cin_p * _r8(cout_p), kh * kw * cout_p * _r8(cin_p))
                             + 148 * kh * kw * cin_p * cout_p + 16)


def conv_x_planes_ok(shape, w: Tensor) -> bool:
    """a producer may hand this convolution its input as bf16 planes only: every engine that can serve the layer reads
    pre-split planes (flag 2048) and never the fp32 tensor"""
    n, h, wd, cin_p = shape
    return (_producer_planes and not (_precision_flag & F_FP32) and cin_p == 64 and n * h * wd >= 128
            and w.shape[1] <= 64 and w.shape[0] >= 3)


# off when a debugging switch routes convolutions to an engine without pre-split operand planes (or TATT_XPLANES=0)
_producer_planes = all(os.environ.get(k, "1") != "0" for k in ("TATT_XPLANES", "TATT_TC", "TATT_TC2"))


def conv_x_planes_alloc(shape, w: Tensor, like: Tensor):
    """workspace of the layer (same layout as conv2d_fwd allocates) + addresses of its X hi / lo planes"""
    n, h, wd, cin_p = shape
    co, ci, kh, kw = w.shape
    P = n * h * wd
    ndy, extra = conv_fwd_ws_elems(cin_p, _pad4(co), kh, kw, P)
    ws = _conv_ws(like, P * cin_p, ndy, extra)
    return ws, ws.data_ptr(), ws.data_ptr() + 2 * _r8(P * cin_p)


def conv2d_fwd(x: Tensor, w: Tensor, b: Optional[Tensor], pad: int, keep: Optional[dict] = None,
               relu: bool = False, stats: Optional[dict] = None, ws_pre: Optional[Tensor] = None) -> Tensor:
    """x [N,H,W,CinP] (CinP >= w.shape[1], multiple of 4) -> y [N,H,W,CoutP] (same H x W: out-of-image taps read 0).
    `keep` (a dict owned by the caller's tape entry) receives the workspace whose X planes the backward pass reuses;
    `relu` fuses max(., 0) into the epilogue (generic engine only); `stats` (a dict) receives under "acc" the
    per-channel {sum, sum of squares} of y (2*Cout doubles) when the layer is served by the kernel that produces them
    in its epilogue (tatt_conv3x3_stats) -- the BatchNorm that follows then skips its own statistics pass.
    `ws_pre`: the layer's workspace with the X planes ALREADY written by the producer of x (conv_x_planes_alloc); x itself
    is then only a shape carrier and is never read."""
    n, h, wd, cin_p = x.shape
    xv = F_A_VALID if ws_pre is not None else 0
    co, ci, kh, kw = w.shape
    cout_p = _pad4(co)
    if b is not None and cout_p != co:
        bp = torch.zeros(cout_p, dtype=torch.float32, device=x.device)
        bp[:co].copy_(b)
        b = bp
    y = empty(n, h, wd, cout_p, like=x)
    P = n * h * wd
    if cout_p == 4 and kw > 1 and cin_p >= 16 and not relu:
        # kx-expansion (see gemm.cu): vertical-tap GEMM with N = kw*4, then a horizontal shift-sum
        wte = empty(kh * cin_p, kw * 4, like=x)
        _cabi.call("tatt_conv_kxexp_pack", _p(w.contiguous()), _p(wte), co, ci, kh, kw, cin_p, 4, _stream())
        t = empty(P, kw * 4, like=x)
        ws = ws_pre if ws_pre is not None else _conv_ws(
            x, x.numel(), P * _r8(kw * 4), kh * cin_p * _r8(kw * 4) + 148 * kh * cin_p * _r8(kw * 4))
        _cabi.call("tatt_conv2d_igemm", _p(x), _p(wte), None, _p(t), n, h, wd, cin_p, kw * 4, kh, 1, pad, 0,
                   _precision_flag | xv, _p(ws), 0 if ws is None else ws.numel(), _stream())
        _cabi.call("tatt_conv_kxexp_reduce", _p(t), _p(b), _p(y), P, wd, kw, 4, pad, _stream())
        if keep is not None:
            keep["ws"] = ws
        return y
    wt = conv_pack(w, cin_p, cout_p, False)
    if (stats is not None and kh == 3 and kw == 3 and pad == 1 and not relu and not (_precision_flag & F_FP32)
            and _cabi.lib().tatt_conv3x3_stats_supported(h, wd, cin_p, cout_p)):
        ws = ws_pre if ws_pre is not None else _conv_ws(
            x, x.numel(), P * _r8(cout_p),
            max(kh * kw * cin_p * _r8(cout_p), kh * kw * cout_p * _r8(cin_p)) + 148 * kh * kw * cin_p * cout_p + 16)
        acc = torch.empty(CONV_STATS_ROWS, 2 * cout_p, dtype=torch.float32, device=x.device)   # one row per CTA
        _cabi.call("tatt_conv3x3_stats", _p(x), _p(wt), _p(b), _p(y), n, h, wd, _precision_flag | xv, _p(ws), ws.numel(),
                   _p(acc), _stream())
        stats["acc"] = acc
        if keep is not None:
            keep["ws"] = ws
        return y
    ws = ws_pre if ws_pre is not None else _conv_ws(
        x, x.numel(), P * _r8(cout_p),
        max(kh * kw * cin_p * _r8(cout_p), kh * kw * cout_p * _r8(cin_p)) + 148 * kh * kw * cin_p * cout_p + 16)
    _cabi.call("tatt_conv2d_igemm", _p(x), _p(wt), _p(b), _p(y), n, h, wd, cin_p, cout_p, kh, kw, pad, pad,
               _precision_flag | (F_RELU if relu else 0) | xv, _p(ws), 0 if ws is None else ws.numel(), _stream())
    if keep is not None:
        keep["ws"] = ws
    return y


F_B_VALID = 4096
CONV_STATS_ROWS = 160      # TATT_CONV_STATS_ROWS of include/tatt_b200.h


class _NoSide:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def conv2d_bwd(x: Tensor, w: Tensor, dy: Tensor, pad: int, need_dx: bool = True, need_dw: bool = True,
               has_bias: bool = True, keep: Optional[dict] = None, side=None, dy_planes: bool = False):
    """-> (dx [N,H,W,CinP] | None, dW [Cout,Cin,KH,KW] | None, db [Cout] | None).
    `side(*reads)` (optional, the tape's) returns a context manager that runs the enclosed launches on the side
    stream: the weight gradient (and the bias gradient) only feed the optimizer, so they run there, concurrently with
    the data gradient.  The two passes then never share scratch: the 3x3 / 64-channel layers split dY ONCE on the
    caller's stream (planes behind the X planes, bias gradient from the same pass) and both passes read them; every
    other layer gives the data-gradient pass its own workspace."""
    n, h, wd, cin_p = x.shape
    co, ci, kh, kw = w.shape
    cout_p = dy.shape[-1]
    P = n * h * wd
    dx = dw = db = None
    on_side = side if side is not None else (lambda *r: _NoSide())
    kx = cout_p == 4 and kw > 1 and cin_p >= 16
    ndy = P * _r8(kw * 4) if kx else P * _r8(cout_p)
    extra = (kh * cin_p * _r8(kw * 4) + 148 * kh * cin_p * _r8(kw * 4)) if kx else (
        max(kh * kw * cin_p * _r8(cout_p), kh * kw * cout_p * _r8(cin_p)) + 148 * kh * kw * cin_p * cout_p + 16)
    ws = keep.get("ws") if keep is not None else None
    # X planes written by the forward pass (only the shapes whose forward AND weight-gradient run on the plane engines)
    x_valid = ws is not None and not (_precision_flag & F_FP32) and cin_p % 64 == 0 and P >= 128
    if ws is None:
        ws = _conv_ws(x, x.numel(), ndy, extra)
    wsb = 0 if ws is None else ws.numel()
    # 3x3 convolutions with 64 input channels: both backward passes run on the plane engines and share the dY planes
    share = (ws is not None and _conv_share_dy and not kx and kh == 3 and kw == 3 and cin_p == 64
             and cout_p in (64, 128, 192, 256) and P >= 128)
    dy2 = dy.view(-1, cout_p)
    dy_valid = False
    if dy_planes:
        # the producer of dY (a train-mode BatchNorm backward) wrote the planes itself; dY's column sums -- the bias
        # gradient -- are exactly zero behind a train-mode BatchNorm (sum_p dX_bn = 0)
        assert share and wd % 64 == 0
        if has_bias:
            db = zeros(co, like=x)
        dy_valid = True
    elif share and need_dw and wd % 64 == 0:
        # one split of dY on this stream serves both passes; its column sums are the bias gradient
        off = 4 * _r8(x.numel())
        dbp = empty(cout_p, like=x) if has_bias else None
        _cabi.call("tatt_split_bf16", _p(dy2), cout_p, P, cout_p, 0, ws.data_ptr() + off,
                   ws.data_ptr() + off + 2 * _r8(P * cout_p), _p(dbp), _stream())
        if has_bias:
            db = dbp[:co]
        dy_valid = True
    if need_dw and kx:
        with on_side(x, dy, ws):
            dt = empty(P, kw * 4, like=x)
            _cabi.call("tatt_conv_kxexp_expand", _p(dy), _p(dt), P, wd, kw, 4, pad, _stream())
            dwte = empty(kh * cin_p, kw * 4, like=x)
            _cabi.call("tatt_conv2d_wgrad", _p(x), _p(dt), _p(dwte), n, h, wd, cin_p, kw * 4, kh, 1, pad, 0,
                       _precision_flag | (F_A_VALID if x_valid else 0), _p(ws), wsb, _stream())
            dw = empty(co, ci, kh, kw, like=x)
            _cabi.call("tatt_conv_kxexp_unpack_grad", _p(dwte), _p(dw), co, ci, kh, kw, cin_p, 4, _stream())
            if has_bias:
                db = colsum(dy2)[:co]
    elif need_dw:
        with on_side(x, dy, ws):
            dwt = empty(kh * kw * cin_p, cout_p, like=x)
            _cabi.call("tatt_conv2d_wgrad", _p(x), _p(dy), _p(dwt), n, h, wd, cin_p, cout_p, kh, kw, pad, pad,
                       _precision_flag | (F_A_VALID if x_valid else 0) | (F_B_VALID if dy_valid else 0), _p(ws), wsb,
                       _stream())
            dw = empty(co, ci, kh, kw, like=x)
            _cabi.call("tatt_conv_weight_unpack_grad", _p(dwt), _p(dw), co, ci, kh, kw, cin_p, cout_p, _stream())
            if has_bias and db is None:
                db = colsum(dy2)[:co]
        # without the shared split, the tcgen05 weight-gradient kernels still leave the dY planes behind the X planes;
        # they are only safe to reuse when that pass ran on THIS stream
        if share and not dy_valid and side is None:
            dy_valid = True
    if need_dx:
        wb = conv_pack(w, cin_p, cout_p, True)
        dx = empty(n, h, wd, cin_p, like=x)
        if dy_valid:
            wsd = ws[4 * _r8(x.numel()):]                       # bytes: skip the X planes
        elif need_dw and side is not None:
            # the weight-gradient pass owns `ws` on the side stream: separate scratch for this pass
            wsd = _conv_ws(x, dy.numel(), 0, kh * kw * cout_p * _r8(cin_p) + 16)
        else:
            wsd = ws
        _cabi.call("tatt_conv2d_igemm", _p(dy), _p(wb), None, _p(dx), n, h, wd, cout_p, cin_p, kh, kw,
                   kh - 1 - pad, kw - 1 - pad, _precision_flag | (F_A_VALID if dy_valid else 0), _p(wsd),
                   0 if wsd is None else wsd.numel(), _stream())
    return dx, dw, db


_conv_share_dy = os.environ.get("TATT_CONV_SHARE", "1") != "0"


# ------------------------------------------------------------------------------------------ norms
def bn_stats(x2: Tensor, eps: float, momentum: float, running_mean: Optional[Tensor],
             running_var: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    P, C = x2.shape
    st = empty(2, C, like=x2)
    ws = torch.empty(2 * C, dtype=torch.float64, device=x2.device)
    _cabi.call("tatt_bn_stats", _p(x2), P, C, eps, momentum, _p(st[0]), _p(st[1]), _p(running_mean),
               _p(running_var), _p(ws), _stream())
    return st[0], st[1]


def bn_finalize(acc: Tensor, P: int, C: int, eps: float, momentum: float, running_mean: Optional[Tensor],
                running_var: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """mean / invstd from {sum, sum of squares} accumulated by the producing convolution (conv2d_fwd(stats=...))"""
    st = torch.empty(2, C, dtype=torch.float32, device=acc.device)
    _cabi.call("tatt_bn_finalize", _p(acc), acc.shape[0], P, C, eps, momentum, _p(st[0]), _p(st[1]), _p(running_mean),
               _p(running_var), _stream())
    return st[0], st[1]


def bn_eval_stats(running_mean: Tensor, running_var: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    C = running_mean.numel()
    st = empty(2, C, like=running_mean)
    _cabi.call("tatt_bn_eval_stats", _p(running_mean), _p(running_var), eps, C, _p(st[0]), _p(st[1]), _stream())
    return st[0], st[1]


def bn_apply(x2: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor, act: int) -> Tensor:
    P, C = x2.shape
    y = torch.empty_like(x2)
    _cabi.call("tatt_bn_apply_fwd", _p(x2), _p(y), _p(mean), _p(invstd), _p(gamma), _p(beta), act, P, C, _stream())
    return y


def bn_apply_planes(x2: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor, act: int, hi_ptr: int,
                    lo_ptr: int) -> None:
    P, C = x2.shape
    _cabi.call("tatt_bn_apply_planes", _p(x2), hi_ptr, lo_ptr, _p(mean), _p(invstd), _p(gamma), _p(beta), act, P, C,
               _stream())


def bn_bwd(x2: Tensor, dy2: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor, act: int,
           training: bool, need_dx: bool = True):
    P, C = x2.shape
    dx = torch.empty_like(x2) if need_dx else None
    dg = empty(2, C, like=x2)
    ws = torch.empty(2 * C, dtype=torch.float64, device=x2.device)
    _cabi.call("tatt_bn_bwd", _p(x2), _p(dy2), _p(mean), _p(invstd), _p(gamma), _p(beta), act, 1 if training else 0,
               P, C, _p(dx), _p(dg[0]), _p(dg[1]), _p(ws), _stream())
    return dx, dg[0], dg[1]


def bn_bwd_planes(x2: Tensor, dy2: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor, act: int,
                  hi_ptr: int, lo_ptr: int):
    """train-mode BatchNorm backward whose dX goes straight into the bf16 operand planes at (hi_ptr, lo_ptr) -- the dY
    slot of the workspace of the convolution in front (see conv_dy_plane_ptrs); -> (dgamma, dbeta)"""
    P, C = x2.shape
    dg = empty(2, C, like=x2)
    ws = torch.empty(2 * C, dtype=torch.float64, device=x2.device)
    _cabi.call("tatt_bn_bwd_planes", _p(x2), _p(dy2), _p(mean), _p(invstd), _p(gamma), _p(beta), act, 1, P, C, hi_ptr,
               lo_ptr, _p(dg[0]), _p(dg[1]), _p(ws), _stream())
    return dg[0], dg[1]


def conv_dy_plane_ptrs(x: Tensor, cout_p: int, ws: Tensor) -> Tuple[int, int]:
    """addresses of the dY hi / lo planes inside a convolution workspace (layout of _conv_ws / conv2d_bwd)"""
    P = x.numel() // x.shape[-1]
    off = 4 * _r8(x.numel())
    return ws.data_ptr() + off, ws.data_ptr() + off + 2 * _r8(P * cout_p)


def conv_bwd_shares_dy(x: Tensor, w: Tensor, pad: int, ws: Optional[Tensor]) -> bool:
    """conv2d_bwd will run both backward passes of this layer on the dY planes behind the X planes of `ws`"""
    n, h, wd, cin_p = x.shape
    co, ci, kh, kw = w.shape
    cout_p = _pad4(co)
    return (ws is not None and _conv_share_dy and not (_precision_flag & F_FP32) and kh == 3 and kw == 3 and pad == 1
            and cin_p == 64 and cout_p in (64, 128, 192, 256) and n * h * wd >= 128 and wd % 64 == 0)


def layernorm_fwd(x2: Tensor, r2: Optional[Tensor], gamma: Tensor, beta: Tensor, save: bool):
    """y = LN(x + r); returns (y, S=x+r | None, stats[2,P] | None)"""
    P, C = x2.shape
    assert C == 64
    y = torch.empty_like(x2)
    S = torch.empty_like(x2) if (save and r2 is not None) else None
    st = empty(2, P, like=x2) if save else None
    _cabi.call("tatt_layernorm64_fwd", _p(x2), _p(r2), _p(gamma), _p(beta), _p(y), _p(S),
               _p(st[0]) if save else None, _p(st[1]) if save else None, P, 1e-5, _stream())
    if save and r2 is None:
        S = x2
    return y, S, st


def layernorm_bwd(dy2: Tensor, S: Tensor, st: Tensor, gamma: Tensor):
    P, C = dy2.shape
    dS = torch.empty_like(dy2)
    dgb = empty(2, 64, like=dy2)
    _cabi.call("tatt_layernorm64_bwd", _p(dy2), _p(S), _p(st[0]), _p(st[1]), _p(gamma), _p(dS), _p(dgb[0]),
               _p(dgb[1]), P, _stream())
    return dS, dgb[0], dgb[1]


# ------------------------------------------------------------------------------------------ GRU(32)
def gru32_scan_fwd(gi: Tensor, whh: Tensor, bhh: Tensor, nseq: int, T: int, s_inner: int, outer: int, inner: int,
                   tstride: int, save: bool):
    P = gi.shape[0]
    out = empty(P, 64, like=gi)
    gates = empty(P, 320, like=gi) if save else None
    _cabi.call("tatt_gru32_scan_fwd", _p(gi), _p(whh), _p(bhh), _p(out), _p(gates), nseq, T, s_inner, outer, inner,
               tstride, _stream())
    return out, gates


def gru32_scan_bwd(dout: Tensor, gates: Tensor, whh: Tensor, nseq: int, T: int, s_inner: int, outer: int,
                   inner: int, tstride: int):
    P = dout.shape[0]
    dgi = empty(P, 192, like=dout)
    dgh = empty(P, 192, like=dout)
    _cabi.call("tatt_gru32_scan_bwd", _p(dout), _p(gates), _p(whh), _p(dgi), _p(dgh), nseq, T, s_inner, outer,
               inner, tstride, _stream())
    return dgi, dgh


# ------------------------------------------------------------------------------------------ attention
def mha_fwd(q: Tensor, k: Tensor, v: Tensor, N: int, Lq: int, Lk: int, need_weights: bool, pdrop: float,
            rng: Optional[Tensor], site: int):
    o = torch.empty_like(q)
    aw = empty(N, Lq, Lk, like=q) if need_weights else None
    _cabi.call("tatt_mha64_fwd", _p(q), _p(k), _p(v), _p(o), _p(aw), N, Lq, Lk, pdrop, _p(rng), site, _stream())
    return o, aw


def declayer_ok(N: int, Lq: int, Lk: int) -> bool:
    """the fused tcgen05 decoder-layer kernel (csrc/tc6_declayer.cu) serves this shape / precision mode"""
    return (_declayer_enabled and not (_precision_flag & F_FP32) and Lq % 128 == 0 and 1 <= Lk <= 32)


_declayer_enabled = os.environ.get("TATT_DECLAYER", "1") != "0"


def declayer_fwd(ins: Sequence[Tensor], outs: Sequence[Optional[Tensor]], train: bool, N: int, Lq: int, Lk: int,
                 pdrop: Optional[Sequence[float]], rng: Optional[Tensor], sites: Optional[Sequence[int]]) -> None:
    """tatt_tp_declayer_fwd: ins = 18 tensors, outs = 3 (+ 11 training side outputs); see include/tatt_b200.h"""
    import ctypes
    assert len(ins) == 18 and len(outs) == (14 if train else 3)
    for t in ins:
        _chk(t, "declayer input")
        assert t.is_contiguous()
    pin = (ctypes.c_void_p * 18)(*[t.data_ptr() for t in ins])
    pout = (ctypes.c_void_p * 14)(*([None if t is None else t.data_ptr() for t in outs] + [None] * (14 - len(outs))))
    pd = (ctypes.c_float * 4)(*(pdrop if pdrop is not None else (0.0, 0.0, 0.0, 0.0)))
    st = (ctypes.c_ulonglong * 4)(*(sites if sites is not None else (0, 0, 0, 0)))
    _cabi.call("tatt_tp_declayer_fwd", pin, pout, 1 if train else 0, N, Lq, Lk, pd, _p(rng), st, _stream())


def mha_bwd(q: Tensor, k: Tensor, v: Tensor, do: Tensor, N: int, Lq: int, Lk: int, pdrop: float,
            rng: Optional[Tensor], site: int):
    dq = torch.empty_like(q)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    _cabi.call("tatt_mha64_bwd", _p(q), _p(k), _p(v), _p(do), _p(dq), _p(dk), _p(dv), N, Lq, Lk, pdrop, _p(rng),
               site, _stream())
    return dq, dk, dv


# ------------------------------------------------------------------------------------------ element-wise
def axpby(a: Tensor, b: Optional[Tensor], alpha: float = 1.0, beta: float = 1.0, out: Optional[Tensor] = None):
    if out is None:
        out = torch.empty_like(a)
    _cabi.call("tatt_axpby", _p(a), _p(b), alpha, beta, _p(out), a.numel(), _stream())
    return out


def add(a: Tensor, b: Tensor) -> Tensor:
    return axpby(a, b, 1.0, 1.0)


def add_bcast_rows(a: Optional[Tensor], b: Tensor, rows: int, period: int, cols: int) -> Tensor:
    out = empty(rows, cols, like=b)
    _cabi.call("tatt_add_bcast_rows", _p(a), _p(b), _p(out), rows, period, cols, _stream())
    return out


def prelu_fwd(x: Tensor, w: Tensor) -> Tensor:
    y = torch.empty_like(x)
    _cabi.call("tatt_prelu_fwd", _p(x), _p(w), _p(y), x.numel(), _stream())
    return y


def prelu_bwd(x: Tensor, w: Tensor, dy: Tensor, need_dx: bool = True):
    dx = torch.empty_like(x) if need_dx else None
    dw = empty(1, like=x)
    _cabi.call("tatt_prelu_bwd", _p(x), _p(w), _p(dy), _p(dx), _p(dw), x.numel(), _stream())
    return dx, dw


def dropout(x: Tensor, p: float, rng: Tensor, site: int) -> Tensor:
    y = torch.empty_like(x)
    _cabi.call("tatt_dropout", _p(x), _p(y), x.numel(), p, _p(rng), site, _stream())
    return y


def relu_bwd(y: Tensor, dy: Tensor) -> Tensor:
    dx = torch.empty_like(dy)
    _cabi.call("tatt_relu_bwd", _p(y), _p(dy), _p(dx), y.numel(), _stream())
    return dx


def nchw_to_nhwc(x: Tensor, cp: int) -> Tensor:
    n, c, h, w = x.shape
    out = empty(n, h, w, cp, like=x)
    _cabi.call("tatt_nchw_to_nhwc", _p(x), _p(out), n, c, h, w, cp, _stream())
    return out


def nhwc_to_nchw(x: Tensor, c: int, do_tanh: bool = False) -> Tensor:
    n, h, w, cp = x.shape
    out = empty(n, c, h, w, like=x)
    _cabi.call("tatt_nhwc_to_nchw", _p(x), _p(out), n, c, h, w, cp, 1 if do_tanh else 0, _stream())
    return out


def tanh_bwd_to_nhwc(dout: Tensor, out: Tensor, cp: int) -> Tensor:
    n, c, h, w = dout.shape
    dpre = empty(n, h, w, cp, like=dout)
    _cabi.call("tatt_tanh_bwd_nchw_to_nhwc", _p(dout), _p(out), _p(dpre), n, c, h, w, cp, _stream())
    return dpre


def pixshuf2_mish_fwd(x: Tensor) -> Tensor:
    n, h, w, c4 = x.shape
    out = empty(n, 2 * h, 2 * w, c4 // 4, like=x)
    _cabi.call("tatt_pixshuf2_mish_fwd", _p(x), _p(out), n, h, w, c4 // 4, _stream())
    return out


def pixshuf2_mish_planes(x: Tensor, hi_ptr: int, lo_ptr: int) -> None:
    n, h, w, c4 = x.shape
    _cabi.call("tatt_pixshuf2_mish_planes", _p(x), hi_ptr, lo_ptr, n, h, w, c4 // 4, _stream())


def pixshuf2_mish_bwd(x: Tensor, dout: Tensor) -> Tensor:
    n, h, w, c4 = x.shape
    din = torch.empty_like(x)
    _cabi.call("tatt_pixshuf2_mish_bwd", _p(x), _p(dout), _p(din), n, h, w, c4 // 4, _stream())
    return din


def maxpool_fwd(x: Tensor, kh: int, kw: int) -> Tensor:
    n, h, w, c = x.shape
    out = empty(n, h // kh, w // kw, c, like=x)
    _cabi.call("tatt_maxpool_fwd", _p(x), _p(out), n, h, w, c, kh, kw, _stream())
    return out


def maxpool_bwd(x: Tensor, dout: Tensor, kh: int, kw: int) -> Tensor:
    n, h, w, c = x.shape
    din = torch.empty_like(x)
    _cabi.call("tatt_maxpool_bwd", _p(x), _p(dout), _p(din), n, h, w, c, kh, kw, _stream())
    return din


def tps_sample_fwd(x: Tensor, ctrl: Tensor, invk: Tensor, repr_: Tensor, want_src: bool = False):
    n, h, w, c = x.shape
    assert c == 4
    invk, repr_, ctrl = invk.contiguous(), repr_.contiguous(), ctrl.contiguous()   # torch.inverse is column-major
    out = torch.empty_like(x)
    src = empty(n, h * w, 2, like=x) if want_src else None
    _cabi.call("tatt_tps_sample_fwd", _p(x), _p(ctrl), _p(invk), _p(repr_), _p(out), _p(src), n, h, w, _stream())
    return out, src


def tps_sample_bwd(x: Tensor, ctrl: Tensor, invk: Tensor, repr_: Tensor, dout: Tensor) -> Tensor:
    n, h, w, c = x.shape
    invk, repr_, ctrl = invk.contiguous(), repr_.contiguous(), ctrl.contiguous()
    dctrl = torch.empty_like(ctrl)
    _cabi.call("tatt_tps_sample_bwd", _p(x), _p(ctrl), _p(invk), _p(repr_), _p(dout), _p(dctrl), n, h, w, _stream())
    return dctrl


def memcpy(dst: Tensor, src: Tensor) -> None:
    assert dst.numel() * dst.element_size() == src.numel() * src.element_size()
    _cabi.call("tatt_memcpy_d2d", _p(dst), _p(src), src.numel() * src.element_size(), _stream())


_pack_tables: dict = {}


def packed(srcs: Sequence[Tensor], like: Tensor, want_flat: bool = False):
    """Copies of the (small, contiguous fp32) tensors `srcs` laid out back to back in ONE fresh buffer, made by ONE
    kernel launch (tatt_multi_copy) instead of one memcpy node per tensor; returns views shaped like the sources.
    Sources whose sizes are multiples of 4 end up densely concatenated (neighbouring views can be re-viewed as one
    tensor).  The {address, offset, count} table is cached per source-address tuple and staged through pinned host
    memory, so the first use is legal inside a CUDA-graph capture."""
    key = tuple(t.data_ptr() for t in srcs) + tuple(t.numel() for t in srcs) + (like.device.index,)
    ent = _pack_tables.get(key)
    if ent is None:
        rows, offs, off = [], [], 0
        for t in srcs:
            assert t.is_contiguous() and t.dtype == torch.float32
            n = t.numel()
            offs.append(off)
            for c in range(0, n, 16384):
                rows.append((t.data_ptr() + 4 * c, off + c, min(16384, n - c)))
            off += (n + 3) // 4 * 4
        host = torch.tensor(rows, dtype=torch.int64).pin_memory()
        dev = torch.empty(len(rows), 3, dtype=torch.int64, device=like.device)
        dev.copy_(host, non_blocking=True)
        if len(_pack_tables) > 4096:
            _pack_tables.clear()
        ent = _pack_tables[key] = (dev, host, offs, off, len(rows))
    dev, _, offs, total, nrows = ent
    flat = torch.empty(total, dtype=torch.float32, device=like.device)
    _cabi.call("tatt_multi_copy", _p(dev), nrows, _p(flat), None, _stream())
    views = [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, srcs)]
    return (views, flat) if want_flat else views


def zeros(*shape, like: Tensor) -> Tensor:
    t = torch.empty(*shape, dtype=torch.float32, device=like.device)
    _cabi.call("tatt_memset0", _p(t), t.numel() * 4, _stream())
    return t


# ------------------------------------------------------------------------------------------ RNG state
class DeviceRNG:
    """{seed, counter} in device memory so dropout is CUDA-graph safe.  `snapshot()` copies the state
    for one forward call (its backward re-derives the same masks) and advances the counter."""
    _per_device = {}
    _seed = None            # set by manual_seed(); None -> torch.initial_seed() at first use

    def __init__(self, device: torch.device, seed: int):
        self.state = torch.tensor([seed & ((1 << 63) - 1), 0], dtype=torch.int64, device=device)

    @classmethod
    def get(cls, device: torch.device) -> "DeviceRNG":
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in cls._per_device:
            seed = cls._seed
            if seed is None:        # default: torch's seed, decorrelated across data-parallel ranks
                import torch.distributed as dist
                rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
                seed = torch.initial_seed() + 0x9E3779B97F4A7C15 * rank
            cls._per_device[key] = DeviceRNG(device, seed)
        return cls._per_device[key]

    @classmethod
    def manual_seed(cls, seed: int) -> None:
        cls._seed = seed
        for r in cls._per_device.values():
            r.state.copy_(torch.tensor([seed & ((1 << 63) - 1), 0], dtype=torch.int64))

    def snapshot(self) -> Tensor:
        snap = torch.empty_like(self.state)
        _cabi.call("tatt_memcpy_d2d", _p(snap), _p(self.state), 16, _stream())
        _cabi.call("tatt_rng_advance", _p(self.state), _stream())
        return snap
