"""A minimal reverse-mode tape over raw CUDA tensors.

Each method runs the forward through the C-ABI kernels (tatt_b200.ops) and, when recording, appends a
closure that turns the gradient of its output into gradients of its inputs -- again only C-ABI calls.
Stages (tatt_b200/stages.py) build one Tape per autograd.Function call; torch.autograd only sees the
stage boundary.  Gradients are keyed by id(tensor); closures keep every keyed tensor alive."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops

Tensor = torch.Tensor


class Tape:
    def __init__(self, record: bool):
        self.record = record
        self._ops: List = []
        self._g: Dict[int, Tensor] = {}
        self._keep: List[Tensor] = []
        self._xplanes: Dict[int, tuple] = {}     # shape-carrier tensors whose values exist only as a conv's X planes
        self._conv_keep: Dict[int, tuple] = {}   # conv outputs whose BatchNorm may write the conv's dY planes directly
        self._stats: Dict[int, tuple] = {}   # conv outputs whose BatchNorm sums were produced by the conv kernel itself
        self._side_keep: List = []    # tensors read by side-stream kernels: kept alive until the join (see backward())
        self._pflag = None            # inside precision_scope(flag): forward ops AND their backward closures run in that mode

    # ------------------------------------------------------------------ gradient bookkeeping
    def grad(self, t: Tensor) -> Optional[Tensor]:
        return self._g.get(id(t))

    def add_grad(self, t: Tensor, g: Optional[Tensor]) -> None:
        if g is None:
            return
        k = id(t)
        if k in self._g:
            ops.join_side()           # either addend may have been produced on the side stream
            self._g[k] = ops.add(self._g[k], g.reshape(self._g[k].shape))
        else:
            self._g[k] = g
            self._keep.append(t)

    def seed(self, t: Tensor, g: Optional[Tensor]) -> None:
        if g is not None:
            self.add_grad(t, g.contiguous().reshape(t.shape))

    def backward(self) -> None:
        for fn in reversed(self._ops):
            fn()
        ops.join_side()               # weight-gradient kernels launched on the side stream (see _side)
        self._side_keep.clear()

    def _side(self, *reads: Optional[Tensor]):
        """`with self._side(t0, t1, ...):` -- run the enclosed launches on the side stream (ops.side_stream).  `reads`
        are the tensors those kernels read: main-stream temporaries among them could otherwise be freed -- and their
        memory reused by later main-stream work -- while the side stream still reads them, so they are kept alive
        until backward() has joined the streams.  Tensors ALLOCATED inside the block belong to the side stream; they
        are only consumed after the join (gradient packing / autograd accumulation)."""
        self._side_keep.extend(t for t in reads if t is not None)
        return ops.side_stream()

    def _push(self, fn) -> None:
        if self.record:
            if self._pflag is not None:
                inner, flag = fn, self._pflag

                def fn():
                    with ops.precision_flag_scope(flag):
                        inner()
            self._ops.append(fn)

    def precision_scope(self, flag: int):
        """Context manager: the ops recorded inside AND their backward closures run with precision flag `flag`
        (0 = fp32 parity on tensor cores, ops.F_FP32 = fp32 FFMA kernels, ops.F_BF16), whatever the global mode is."""
        tape = self

        class _Scope:
            def __enter__(self):
                self.prev, tape._pflag = tape._pflag, flag
                self.ctx = ops.precision_flag_scope(flag)
                self.ctx.__enter__()

            def __exit__(self, *a):
                self.ctx.__exit__(*a)
                tape._pflag = self.prev
        return _Scope()

    def fp32_scope(self):
        return self.precision_scope(ops.F_FP32)

    # ------------------------------------------------------------------ dense layers
    def _split_grad(self, dy: Tensor, bias: Optional[Tensor]):
        """bf16 planes of a gradient matrix (shared by its data- and weight-gradient GEMMs); the bias gradient
        (column sums) is produced by the same pass."""
        db = ops.empty(dy.shape[1], like=dy) if bias is not None else None
        dyP = ops.split_matrix(dy, db)
        if dyP is None and bias is not None:
            ops.colsum(dy, out=db)
        if bias is not None:
            self.add_grad(bias, db)
        return dyP

    def linear(self, x: Tensor, W: Tensor, b: Optional[Tensor], relu: bool = False) -> Tensor:
        """x [M,K] @ W[N,K]^T + b (optionally ReLU)"""
        M, K = x.shape
        N = W.shape[0]
        rows = ops.rows_gemm_ok(M, K, N) and ops.rows_gemm_ok(M, N, K) and ops.rows_wgrad_ok(M, N, K)
        xP = ops.split_matrix(x) if (self.record and not rows) else None
        y = ops.linear_fwd(x, W, b, relu=relu, xP=xP)

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            if relu:
                dy = ops.relu_bwd(y, dy)
            if rows:        # fused-split kernels: one pass over (dy, x) for dW + db, one over dy for dx
                with self._side(dy, x):
                    dW, db = ops.linear_bwd_weight_rows(dy, x, b is not None)
                self.add_grad(W, dW)
                if b is not None:
                    self.add_grad(b, db)
                self.add_grad(x, ops.linear_bwd_data(dy, W))
                return
            dyP = self._split_grad(dy, b)
            self.add_grad(W, ops.linear_bwd_weight(dy, x, dyP=dyP, xP=xP))
            self.add_grad(x, ops.linear_bwd_data(dy, W, dyP=dyP))
        self._push(bwd)
        return y

    def linear_cat(self, parts: Sequence[Tensor], W4: Tensor, b: Tensor) -> Tensor:
        """1x1 convolution over the channel-concatenation of `parts` ([P,Ci] each) without
        materialising the cat (tsrn.py:902 + 1075): y = sum_i parts_i @ W[:, off_i:off_i+Ci]^T + b."""
        W = W4.view(W4.shape[0], -1)
        M = parts[0].shape[0]
        rows = (len(parts) <= 3 and all(t.shape[1] == 64 for t in parts) and W.shape[0] == 64
                and ops.rows_gemm_ok(M, 64 * len(parts), 64) and ops.rows_wgrad_ok(M, 64, 64 * len(parts)))
        if rows:
            y = ops.rows_gemm(parts, W, b, ops.empty(M, 64, like=parts[0]))

            def bwd_rows():
                dy = self.grad(y)
                if dy is None:
                    return
                with self._side(dy, *parts):
                    dW, db = ops.linear_bwd_weight_rows_parts(dy, parts, True)
                self.add_grad(W4, dW.view_as(W4))
                self.add_grad(b, db)
                for i, t in enumerate(parts):
                    self.add_grad(t, ops.linear_bwd_data(dy, W[:, 64 * i:64 * (i + 1)]))
            self._push(bwd_rows)
            return y
        y = None
        off = 0
        pls = []
        for t in parts:
            ci = t.shape[1]
            tP = ops.split_matrix(t) if self.record else None
            pls.append(tP)
            y = ops.linear_fwd(t, W[:, off:off + ci], b if off == 0 else None, out=y, accumulate=off > 0, xP=tP)
            off += ci

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            dyP = self._split_grad(dy, b)
            dW = ops.zeros(W.shape[0], W.shape[1], like=dy)
            o = 0
            for t, tP in zip(parts, pls):
                ci = t.shape[1]
                ops.linear_bwd_weight(dy, t, out=dW[:, o:o + ci], dyP=dyP, xP=tP)
                self.add_grad(t, ops.linear_bwd_data(dy, W[:, o:o + ci], dyP=dyP))
                o += ci
            self.add_grad(W4, dW.view_as(W4))
        self._push(bwd)
        return y

    def conv(self, x4: Tensor, w: Tensor, b: Optional[Tensor], pad: int, need_dx: bool = True,
             relu: bool = False, bn_next: bool = False) -> Tensor:
        """bn_next: a train-mode BatchNorm consumes the output next -- let the convolution kernel accumulate its
        statistics in the epilogue when it can (picked up by batchnorm() through self._stats)."""
        keep = {} if self.record else None       # workspace whose X planes the backward pass reuses
        st = {} if bn_next else None
        pre = self._xplanes.pop(id(x4), None)    # the producer wrote the X planes of THIS layer (x4 carries the shape only)
        if pre is not None:
            assert pre[0] is x4 and pre[2] is w, "operand planes were prepared for another convolution"
        y = ops.conv2d_fwd(x4, w, b, pad, keep=keep, relu=relu, stats=st, ws_pre=None if pre is None else pre[1])
        if st:
            self._stats[id(y)] = (y, st["acc"])
        if bn_next and self.record and need_dx and ops.conv_bwd_shares_dy(x4, w, pad, keep.get("ws")):
            # the BatchNorm that consumes y may write its dX straight into this layer's dY planes (batchnorm())
            self._conv_keep[id(y)] = (y, keep, x4, ops._pad4(w.shape[0]))

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            if relu:
                dy = ops.relu_bwd(y, dy)
            dx, dw, db = ops.conv2d_bwd(x4, w, dy, pad, need_dx=need_dx, has_bias=b is not None, keep=keep,
                                        side=self._side, dy_planes=bool(keep.get("dy_planes")))
            self.add_grad(w, dw)
            if b is not None:
                self.add_grad(b, db)
            if need_dx:
                self.add_grad(x4, dx)
        self._push(bwd)
        return y

    # ------------------------------------------------------------------ normalisation
    def _planes_carrier(self, shape, w: Tensor, like: Tensor):
        """-> (shape-carrier tensor y, hi address, lo address): y has no values, they go to the X planes of the conv `w`"""
        ws, hi, lo = ops.conv_x_planes_alloc(shape, w, like)
        y = torch.empty(1, dtype=like.dtype, device=like.device).expand(shape)
        self._xplanes[id(y)] = (y, ws, w)
        return y, hi, lo

    def batchnorm(self, x: Tensor, bn: torch.nn.Module, act: int, training: bool,
                  planes_for: Optional[Tensor] = None) -> Tensor:
        """BatchNorm over all leading dims of a channels-last tensor, fused activation.  planes_for = weight of the
        convolution that is the ONLY consumer of the result: the result is then written as that layer's bf16 operand
        planes and the returned tensor only carries the shape (pass it to conv() and nothing else)."""
        C = x.shape[-1]
        x2 = x.view(-1, C)
        use_batch = training or bn.running_mean is None
        if use_batch:
            if bn.momentum is None and training and bn.running_mean is not None:
                # torch: cumulative moving average 1/num_batches_tracked (a host-side value; nothing on this path
                # constructs such a BatchNorm, so it is not a device-resident / graph-safe quantity)
                raise NotImplementedError("tatt_b200: BatchNorm(momentum=None) (cumulative average) is not supported")
            pre = self._stats.pop(id(x), None)
            mom = bn.momentum if bn.momentum is not None else 0.1
            rm, rv = (bn.running_mean, bn.running_var) if training else (None, None)
            if pre is not None and pre[0] is x and pre[1].shape[1] == 2 * C:
                mean, invstd = ops.bn_finalize(pre[1], x2.shape[0], C, bn.eps, mom, rm, rv)
            else:
                mean, invstd = ops.bn_stats(x2, bn.eps, mom, rm, rv)
            if training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
        else:
            mean, invstd = ops.bn_eval_stats(bn.running_mean, bn.running_var, bn.eps)
        gamma, beta = bn.weight, bn.bias
        if planes_for is not None and x.dim() == 4 and ops.conv_x_planes_ok(x.shape, planes_for):
            y, hi, lo = self._planes_carrier(x.shape, planes_for, x)
            ops.bn_apply_planes(x2, mean, invstd, gamma, beta, act, hi, lo)
        else:
            y = ops.bn_apply(x2, mean, invstd, gamma, beta, act).view(x.shape)
        ck = self._conv_keep.pop(id(x), None)
        if ck is not None and not (ck[0] is x and use_batch and training):
            ck = None

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            if ck is not None and self.grad(x) is None and ck[1].get("ws") is not None:
                # x is the output of a 3x3 convolution: dX is only ever read as that layer's dY operand planes
                hi, lo = ops.conv_dy_plane_ptrs(ck[2], ck[3], ck[1]["ws"])
                dg, db = ops.bn_bwd_planes(x2, dy.view(-1, C), mean, invstd, gamma, beta, act, hi, lo)
                ck[1]["dy_planes"] = True
                self.add_grad(gamma, dg)
                self.add_grad(beta, db)
                self.add_grad(x, torch.empty(1, dtype=x.dtype, device=x.device).expand(x.shape))   # marker, never read
                return
            dx, dg, db = ops.bn_bwd(x2, dy.view(-1, C), mean, invstd, gamma, beta, act, use_batch)
            self.add_grad(gamma, dg)
            self.add_grad(beta, db)
            self.add_grad(x, dx.view(x.shape))
        self._push(bwd)
        return y

    def add_layernorm(self, x: Tensor, r: Optional[Tensor], ln: torch.nn.Module) -> Tensor:
        """LN(x + r) over 64 channels."""
        y, S, st = ops.layernorm_fwd(x, r, ln.weight, ln.bias, save=self.record)

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            dS, dg, db = ops.layernorm_bwd(dy, S, st, ln.weight)
            self.add_grad(ln.weight, dg)
            self.add_grad(ln.bias, db)
            self.add_grad(x, dS)
            if r is not None:
                self.add_grad(r, dS)
        self._push(bwd)
        return y

    # ------------------------------------------------------------------ element-wise
    def view(self, x: Tensor, *shape) -> Tensor:
        """Reshape with gradient pass-through (grads are keyed by tensor identity)."""
        y = x.view(*shape)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(x, dy.view(x.shape))
        self._push(bwd)
        return y

    def add(self, a: Tensor, b: Tensor) -> Tensor:
        y = ops.add(a, b)

        def bwd():
            dy = self.grad(y)
            self.add_grad(a, dy)
            self.add_grad(b, dy)
        self._push(bwd)
        return y

    def scale(self, a: Tensor, alpha: float) -> Tensor:
        y = ops.axpby(a, None, alpha, 0.0)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(a, ops.axpby(dy, None, alpha, 0.0))
        self._push(bwd)
        return y

    def mean2(self, a: Tensor, b: Tensor) -> Tensor:
        y = ops.axpby(a, b, 0.5, 0.5)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                h = ops.axpby(dy, None, 0.5, 0.0)
                self.add_grad(a, h)
                self.add_grad(b, h)
        self._push(bwd)
        return y

    def prelu(self, x: Tensor, w: Tensor) -> Tensor:
        y = ops.prelu_fwd(x, w)

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            dx, dw = ops.prelu_bwd(x, w, dy)
            self.add_grad(w, dw)
            self.add_grad(x, dx)
        self._push(bwd)
        return y

    def dropout(self, x: Tensor, p: float, rng: Optional[Tensor], site: int) -> Tensor:
        if p <= 0.0 or rng is None:
            return x
        y = ops.dropout(x, p, rng, site)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(x, ops.dropout(dy, p, rng, site))
        self._push(bwd)
        return y

    def pixshuf_mish(self, x4: Tensor, planes_for: Optional[Tensor] = None) -> Tensor:
        """planes_for: as in batchnorm() -- the up-sampled map is consumed by one convolution only"""
        n, h, w_, c4 = x4.shape
        oshape = (n, 2 * h, 2 * w_, c4 // 4)
        if planes_for is not None and ops.conv_x_planes_ok(oshape, planes_for):
            y, hi, lo = self._planes_carrier(oshape, planes_for, x4)
            ops.pixshuf2_mish_planes(x4, hi, lo)
        else:
            y = ops.pixshuf2_mish_fwd(x4)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(x4, ops.pixshuf2_mish_bwd(x4, dy))
        self._push(bwd)
        return y

    def maxpool(self, x4: Tensor, kh: int, kw: int) -> Tensor:
        y = ops.maxpool_fwd(x4, kh, kw)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(x4, ops.maxpool_bwd(x4, dy, kh, kw))
        self._push(bwd)
        return y

    def to_nchw(self, x4: Tensor, c: int, tanh: bool = False) -> Tensor:
        """NHWC (padded channels) -> NCHW (first c channels), optional tanh."""
        y = ops.nhwc_to_nchw(x4, c, do_tanh=tanh)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(x4, ops.tanh_bwd_to_nhwc(dy, y if tanh else None, x4.shape[-1]))
        self._push(bwd)
        return y

    def to_nhwc(self, x: Tensor, cp: int) -> Tensor:
        y = ops.nchw_to_nhwc(x, cp)

        def bwd():
            dy = self.grad(y)
            if dy is not None:
                self.add_grad(x, ops.nhwc_to_nchw(dy, x.shape[1]))
        self._push(bwd)
        return y

    # ------------------------------------------------------------------ recurrent / attention composites
    def bigru32(self, c: Tensor, gru: torch.nn.GRU, nseq: int, T: int, s_inner: int, outer: int, inner: int,
                tstride: int) -> Tensor:
        """nn.GRU(64, 32, bidirectional) over strided sequences of the row tensor c [P,64]."""
        P = c.shape[0]
        w_ih = (gru.weight_ih_l0, gru.weight_ih_l0_reverse)
        w_hh = (gru.weight_hh_l0, gru.weight_hh_l0_reverse)
        b_ih = (gru.bias_ih_l0, gru.bias_ih_l0_reverse)
        b_hh = (gru.bias_hh_l0, gru.bias_hh_l0_reverse)
        # both directions' weights side by side (input projections as ONE N=192 GEMM): one pack launch, 8 sources
        _, flat = ops.packed([w_ih[0], w_ih[1], b_ih[0], b_ih[1], w_hh[0], w_hh[1], b_hh[0], b_hh[1]], c, want_flat=True)
        assert flat.numel() == 18816                                         # dense: every source size is a multiple of 4
        wih = flat[0:12288].view(192, 64)
        bih = flat[12288:12480]
        whh = flat[12480:18624].view(2, 96, 32)
        bhh = flat[18624:18816].view(2, 96)
        rows = ops.rows_gemm_ok(P, 64, 192) and ops.rows_gemm_ok(P, 192, 64) and ops.rows_wgrad_ok(P, 192, 64)
        cP = ops.split_matrix(c) if (self.record and not rows) else None
        gi = ops.linear_fwd(c, wih, bih, xP=cP)
        out, gates = ops.gru32_scan_fwd(gi, whh, bhh, nseq, T, s_inner, outer, inner, tstride, save=self.record)
        del gi

        def bwd():
            dout = self.grad(out)
            if dout is None:
                return
            dgi, dgh = ops.gru32_scan_bwd(dout, gates, whh, nseq, T, s_inner, outer, inner, tstride)
            if rows:
                with self._side(dgi, dgh, c, gates):
                    dwih, dbih = ops.linear_bwd_weight_rows(dgi, c, True)          # [192, 64], [192]
                    dwhh, dbhh = ops.empty(2, 96, 32, like=c), ops.empty(192, like=c)
                    # dW_hh[d] = dgh[:, 96d:96d+96]^T h_{t-1}[d]: the saved h_{t-1} of both directions (two 32-column
                    # segments of the gate tensor) form the 64 A channels; the diagonal blocks of the product are kept
                    ops.rows_wgrad(gates, (128, 288), ops._blocks64(dgh), dwhh, 2, 96, 32, True, rb=32, cb=96,
                                   colsum_src=2, dbias=dbhh)
                for d in range(2):
                    self.add_grad(w_hh[d], dwhh[d])
                    self.add_grad(b_hh[d], dbhh[d * 96:(d + 1) * 96])
                    self.add_grad(w_ih[d], dwih[d * 96:(d + 1) * 96])
                    self.add_grad(b_ih[d], dbih[d * 96:(d + 1) * 96])
                self.add_grad(c, ops.linear_bwd_data(dgi, wih))
                return
            dbih, dbhh = ops.empty(192, like=c), ops.empty(192, like=c)
            dgiP, dghP = ops.split_matrix(dgi, dbih), ops.split_matrix(dgh, dbhh)
            if dgiP is None:
                ops.colsum(dgi, out=dbih)
            if dghP is None:
                ops.colsum(dgh, out=dbhh)
            dwih = ops.linear_bwd_weight(dgi, c, dyP=dgiP, xP=cP)             # [192, 64]
            for d in range(2):
                gh_d = dgh[:, d * 96:(d + 1) * 96]
                hprev = gates[:, d * 160 + 128:d * 160 + 160]
                ghP = dghP.slice_cols(d * 96, (d + 1) * 96) if dghP is not None else None
                self.add_grad(w_hh[d], ops.linear_bwd_weight(gh_d, hprev, dyP=ghP))
                self.add_grad(b_hh[d], dbhh[d * 96:(d + 1) * 96])
                self.add_grad(w_ih[d], dwih[d * 96:(d + 1) * 96])
                self.add_grad(b_ih[d], dbih[d * 96:(d + 1) * 96])
            self.add_grad(c, ops.linear_bwd_data(dgi, wih, dyP=dgiP))
        self._push(bwd)
        return out

    def _lin_bwd(self, dy: Tensor, x: Tensor, W: Tensor, want_dx: bool = True):
        """(dW, db, dx) of y = x W^T + b for a gradient dy (row-panel kernels when the shape qualifies)"""
        M, K = x.shape
        N = W.shape[0]
        with self._side(dy, x):
            if ops.rows_wgrad_ok(M, N, K) and ops.rows_gemm_ok(M, N, K):
                dW, db = ops.linear_bwd_weight_rows(dy, x, True)
            else:
                dW, db = ops.linear_bwd_weight(dy, x), ops.colsum(dy)
        return dW, db, (ops.linear_bwd_data(dy, W) if want_dx else None)

    def dec_layer(self, tgt: Tensor, qp: Tensor, kin: Tensor, mem: Tensor, d: torch.nn.Module, lnf: torch.nn.Module,
                  N: int, Lq: int, Lk: int, need_weights: bool, pd: Sequence[float], rng: Optional[Tensor],
                  sites: Sequence[int]):
        """TransformerDecoderLayer_TP.forward_post + the decoder's final norm of its output as ONE tcgen05 kernel
        (csrc/tc6_declayer.cu; transformer_v2.py:806-833, 380-390).  -> (out, inter, attention weights | None).
        pd / sites = (attention, dropout2, dropout, dropout3).  The backward pass runs the round-1 kernels on the side
        outputs the fused forward writes (dropout masks are regenerated from the same Philox counters)."""
        at = d.multihead_attn
        Win, bin_ = at.in_proj_weight, at.in_proj_bias
        Wo, bo = at.out_proj.weight, at.out_proj.bias
        k = ops.linear_fwd(kin, Win[64:128], bin_[64:128])
        v = ops.linear_fwd(mem, Win[128:192], bin_[128:192])
        P = N * Lq
        train = self.record
        new = lambda: ops.empty(P, 64, like=tgt)
        out, inter = new(), new()
        aw = ops.empty(N, Lq, Lk, like=tgt) if need_weights else None
        outs = [out, inter, aw]
        if train:
            qin, q, a, S2, t1, h1, h1d, S3 = (new() for _ in range(8))
            st2, st3, stf = (ops.empty(2, P, like=tgt) for _ in range(3))
            outs += [qin, q, a, S2, t1, h1, h1d, S3, st2, st3, stf]
        if rng is None or not any(x > 0 for x in pd):
            pd, rng_ = (0.0, 0.0, 0.0, 0.0), None
        else:
            rng_ = rng
        ins = [tgt, qp, k, v, Win[0:64], Wo, d.linear1.weight, d.linear2.weight, bin_[0:64], bo, d.norm2.weight,
               d.norm2.bias, d.linear1.bias, d.linear2.bias, d.norm3.weight, d.norm3.bias, lnf.weight, lnf.bias]
        ops.declayer_fwd([t.contiguous() for t in ins], outs, train, N, Lq, Lk, pd, rng_, sites)

        def bwd():
            d_inter, d_out = self.grad(inter), self.grad(out)
            if d_inter is None and d_out is None:
                return
            drop = lambda g, i: g if (rng_ is None or pd[i] <= 0.0) else ops.dropout(g, pd[i], rng_, sites[i])
            if d_inter is not None:                               # final norm of this layer's output
                dS, dg, db = ops.layernorm_bwd(d_inter, out, stf, lnf.weight)
                self.add_grad(lnf.weight, dg)
                self.add_grad(lnf.bias, db)
                d_out = dS if d_out is None else ops.add(d_out, dS)
            dS3, dg, db = ops.layernorm_bwd(d_out, S3, st3, d.norm3.weight)
            self.add_grad(d.norm3.weight, dg)
            self.add_grad(d.norm3.bias, db)
            d_f = drop(dS3, 3)                                    # dropout3
            dW2, db2, d_h1d = self._lin_bwd(d_f, h1d, d.linear2.weight)
            self.add_grad(d.linear2.weight, dW2)
            self.add_grad(d.linear2.bias, db2)
            d_h1 = ops.relu_bwd(h1, drop(d_h1d, 2))
            dW1, db1, d_t1 = self._lin_bwd(d_h1, t1, d.linear1.weight)
            self.add_grad(d.linear1.weight, dW1)
            self.add_grad(d.linear1.bias, db1)
            d_t1 = ops.add(d_t1, dS3)
            dS2, dg, db = ops.layernorm_bwd(d_t1, S2, st2, d.norm2.weight)
            self.add_grad(d.norm2.weight, dg)
            self.add_grad(d.norm2.bias, db)
            d_y = drop(dS2, 1)                                    # dropout2
            dWo, dbo, da = self._lin_bwd(d_y, a, Wo)
            self.add_grad(Wo, dWo)
            self.add_grad(bo, dbo)
            dq, dk, dv = ops.mha_bwd(q, k, v, da, N, Lq, Lk, pd[0] if rng_ is not None else 0.0, rng_, sites[0])
            dWin = ops.empty(192, 64, like=tgt)
            dbin = ops.empty(192, like=tgt)
            for i, (g, src) in enumerate(((dq, qin), (dk, kin), (dv, mem))):
                dW_i, db_i, dx_i = self._lin_bwd(g, src, Win[i * 64:(i + 1) * 64])
                with self._side(dW_i, db_i):                      # same stream as their producers
                    ops.memcpy(dWin[i * 64:(i + 1) * 64], dW_i)
                    ops.memcpy(dbin[i * 64:(i + 1) * 64], db_i)
                if i == 0:
                    self.add_grad(tgt, ops.add(dx_i, dS2))
                    self.add_grad(qp, dx_i)
                else:
                    self.add_grad(src, dx_i)
            self.add_grad(Win, dWin)
            self.add_grad(bin_, dbin)
        self._push(bwd)
        return out, inter, aw

    def mha(self, q_in: Tensor, k_in: Tensor, v_in: Tensor, attn: torch.nn.MultiheadAttention, N: int, Lq: int,
            Lk: int, need_weights: bool, pdrop: float, rng: Optional[Tensor], site: int):
        """nn.MultiheadAttention(64, 4) on token-major [N*L, 64] inputs -> (out [N*Lq,64], weights)."""
        Win, bin_ = attn.in_proj_weight, attn.in_proj_bias
        Wo, bo = attn.out_proj.weight, attn.out_proj.bias
        srcs = (q_in, k_in, v_in)
        sP = [None, None, None]
        rows = all(ops.rows_gemm_ok(t.shape[0], 64, 64) and ops.rows_wgrad_ok(t.shape[0], 64, 64) for t in srcs)
        if self.record and not rows:
            for i in range(3):
                for j in range(i):
                    if srcs[i] is srcs[j]:
                        sP[i] = sP[j]
                        break
                else:
                    sP[i] = ops.split_matrix(srcs[i])
        q = ops.linear_fwd(q_in, Win[0:64], bin_[0:64], xP=sP[0])
        k = ops.linear_fwd(k_in, Win[64:128], bin_[64:128], xP=sP[1])
        v = ops.linear_fwd(v_in, Win[128:192], bin_[128:192], xP=sP[2])
        if pdrop <= 0.0 or rng is None:
            pdrop, rng_ = 0.0, None
        else:
            rng_ = rng
        a, aw = ops.mha_fwd(q, k, v, N, Lq, Lk, need_weights, pdrop, rng_, site)
        y = ops.linear_fwd(a, Wo, bo)

        def bwd():
            dy = self.grad(y)
            if dy is None:
                return
            if rows:
                dWo, dbo = ops.linear_bwd_weight_rows(dy, a, True)
                self.add_grad(Wo, dWo)
                self.add_grad(bo, dbo)
                da = ops.linear_bwd_data(dy, Wo)
                dq, dk, dv = ops.mha_bwd(q, k, v, da, N, Lq, Lk, pdrop, rng_, site)
                dWin = ops.empty(192, 64, like=dy)
                dbin = ops.empty(192, like=dy)
                for i, (g, src) in enumerate(((dq, q_in), (dk, k_in), (dv, v_in))):
                    _, db_i = ops.linear_bwd_weight_rows(g, src, True, out=dWin[i * 64:(i + 1) * 64])
                    ops.memcpy(dbin[i * 64:(i + 1) * 64], db_i)
                    self.add_grad(src, ops.linear_bwd_data(g, Win[i * 64:(i + 1) * 64]))
                self.add_grad(Win, dWin)
                self.add_grad(bin_, dbin)
                return
            dyP = self._split_grad(dy, bo)
            self.add_grad(Wo, ops.linear_bwd_weight(dy, a, dyP=dyP))
            da = ops.linear_bwd_data(dy, Wo, dyP=dyP)
            dq, dk, dv = ops.mha_bwd(q, k, v, da, N, Lq, Lk, pdrop, rng_, site)
            dWin = ops.zeros(192, 64, like=dy)
            dbin = ops.empty(192, like=dy)
            for i, (g, src) in enumerate(((dq, q_in), (dk, k_in), (dv, v_in))):
                gP = ops.split_matrix(g, dbin[i * 64:(i + 1) * 64])
                if gP is None:
                    ops.colsum(g, out=dbin[i * 64:(i + 1) * 64])
                ops.linear_bwd_weight(g, src, out=dWin[i * 64:(i + 1) * 64], dyP=gP, xP=sP[i])
                self.add_grad(src, ops.linear_bwd_data(g, Win[i * 64:(i + 1) * 64], dyP=gP))
            self.add_grad(Win, dWin)
            self.add_grad(bin_, dbin)
        self._push(bwd)
        return y, aw
