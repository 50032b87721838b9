"""Drop-in `CRNN` text-prior generator (the reference's `model/crnn/crnn.py` surface) plus the two helpers
that sit between it and the SR model: `parse_crnn_data` (`interfaces/base.py:797-815`) and the softmax / permute that
turns its logits into the `[N, 37, 1, 26]` text prior (`interfaces/super_resolution.py:794-799`).

Same constructor (`CRNN(imgH, nc, nclass, nh, n_rnn=2, leakyRelu=False)`), same `forward(input)` -> logits `[T, N, nclass]`,
same `state_dict` keys (`cnn.conv0.weight`, `cnn.batchnorm2.running_mean`, `rnn.0.rnn.weight_ih_l0`,
`rnn.0.embedding.weight`, ...) and -- the same torch.nn leaves created in the same order -- bit-identical fresh init.
The leaves are parameter containers; all arithmetic (forward and backward) runs through the C-ABI: convolutions /
BatchNorm / linear layers / recurrent GEMMs on the engines of the SR path, pooling / LSTM cell / bicubic / softmax in
csrc/crnn.cu.  One autograd node for the whole network (tatt_b200.stages.StageFn)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _cabi, ops
from .stages import In, Out, run_stage
from .tape import Tape

Tensor = torch.Tensor
__all__ = ["CRNN", "BidirectionalLSTM", "parse_crnn_data", "softmax_prior", "text_prior"]


def _no_forward(self, *a, **k):
    raise RuntimeError("%s is a parameter container in tatt_b200; its math runs inside the CUDA stage" % type(self).__name__)


class BidirectionalLSTM(nn.Module):
    """crnn.py:5-26 (container): nn.LSTM(nIn, nHidden, bidirectional=True) + nn.Linear(2 nHidden, nOut)"""

    def __init__(self, nIn, nHidden, nOut):
        super().__init__()
        self.rnn = nn.LSTM(nIn, nHidden, bidirectional=True)
        self.embedding = nn.Linear(nHidden * 2, nOut)
    forward = _no_forward


# ------------------------------------------------------------------------------------------------ tape ops (CRNN only)
def _maxpool2d(tape: Tape, x4: Tensor, k, s, p) -> Tensor:
    n, h, w, c = x4.shape
    oh, ow = (h + 2 * p[0] - k[0]) // s[0] + 1, (w + 2 * p[1] - k[1]) // s[1] + 1
    y = ops.empty(n, oh, ow, c, like=x4)
    args = (n, h, w, c, k[0], k[1], s[0], s[1], p[0], p[1])
    _cabi.call("tatt_maxpool2d_fwd", ops._p(x4), ops._p(y), *args, ops._stream())

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        dx = torch.empty_like(x4)
        _cabi.call("tatt_maxpool2d_bwd", ops._p(x4), ops._p(dy), ops._p(dx), *args, ops._stream())
        tape.add_grad(x4, dx)
    tape._push(bwd)
    return y


def _crop(tape: Tape, x4: Tensor, oh: int, ow: int) -> Tensor:
    n, h, w, c = x4.shape
    y = ops.empty(n, oh, ow, c, like=x4)
    _cabi.call("tatt_crop_nhwc", ops._p(x4), ops._p(y), n, h, w, oh, ow, c, 0, ops._stream())

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        dx = torch.empty_like(x4)
        _cabi.call("tatt_crop_nhwc", ops._p(dy), ops._p(dx), n, h, w, oh, ow, c, 1, ops._stream())
        tape.add_grad(x4, dx)
    tape._push(bwd)
    return y


def _permute_102(tape: Tape, x3: Tensor) -> Tensor:
    a, b, c = x3.shape
    y = ops.empty(b, a, c, like=x3)
    _cabi.call("tatt_permute_102", ops._p(x3), ops._p(y), a, b, c, ops._stream())

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        dx = torch.empty_like(x3)
        _cabi.call("tatt_permute_102", ops._p(dy), ops._p(dx), b, a, c, ops._stream())
        tape.add_grad(x3, dx)
    tape._push(bwd)
    return y


def _bilstm(tape: Tape, X: Tensor, lstm: nn.LSTM, T: int, Nb: int) -> Tensor:
    """nn.LSTM(nIn, H, bidirectional=True) over X [T*Nb, nIn] (row t*Nb + n) -> OUT [T*Nb, 2H] = [h_fwd | h_bwd].
    The input projections of all steps and both directions are ONE GEMM; the recurrence is T steps of a [Nb x H] x
    [H x 4H] GEMM per direction + the cell kernel (csrc/crnn.cu); backward mirrors it step by step."""
    H = lstm.hidden_size
    n_in = X.shape[1]
    w_ih = (lstm.weight_ih_l0, lstm.weight_ih_l0_reverse)
    w_hh = (lstm.weight_hh_l0, lstm.weight_hh_l0_reverse)
    b_ih = (lstm.bias_ih_l0, lstm.bias_ih_l0_reverse)
    b_hh = (lstm.bias_hh_l0, lstm.bias_hh_l0_reverse)
    st = ops._stream
    wih = ops.empty(8 * H, n_in, like=X)
    bih = ops.empty(8 * H, like=X)
    bhh = ops.empty(2, 4 * H, like=X)
    whh = [w.contiguous() for w in w_hh]
    for d in range(2):
        ops.memcpy(wih[d * 4 * H:(d + 1) * 4 * H], w_ih[d])
        ops.memcpy(bih[d * 4 * H:(d + 1) * 4 * H], b_ih[d])
        ops.memcpy(bhh[d], b_hh[d])
    G = ops.linear_fwd(X, wih, bih)                               # [T*Nb, 8H] -> activated gates in place
    CS = ops.empty(2, T, Nb, H, like=X)
    OUT = ops.empty(T * Nb, 2 * H, like=X)
    GH = ops.empty(2, Nb, 4 * H, like=X)
    for s in range(T):
        if s > 0:
            for d in range(2):
                tp = s - 1 if d == 0 else T - s
                hprev = OUT[tp * Nb:(tp + 1) * Nb, d * H:(d + 1) * H]
                ops.gemm(0, 1, hprev, 2 * H, whh[d], H, GH[d], 4 * H, bhh[d], Nb, 4 * H, H)
        _cabi.call("tatt_lstm_gate_fwd", ops._p(G), ops._p(GH) if s > 0 else None, ops._p(bhh), ops._p(CS), ops._p(OUT),
                   s, T, Nb, H, st())

    def bwd():
        dOUT = tape.grad(OUT)
        if dOUT is None:
            return
        dG = ops.empty(T * Nb, 8 * H, like=X)
        DH = ops.empty(2, Nb, H, like=X)
        DC = ops.empty(2, Nb, H, like=X)
        for s in range(T - 1, -1, -1):
            _cabi.call("tatt_lstm_gate_bwd", ops._p(G), ops._p(CS), ops._p(dOUT), ops._p(DH), ops._p(DC), ops._p(dG), s, T,
                       Nb, H, st())
            if s > 0:
                for d in range(2):
                    t = s if d == 0 else T - 1 - s
                    dg = dG[t * Nb:(t + 1) * Nb, d * 4 * H:(d + 1) * 4 * H]
                    ops.gemm(0, 0, dg, 8 * H, whh[d], H, DH[d], H, None, Nb, H, 4 * H)
        dwih = ops.linear_bwd_weight(dG, X)                       # [8H, nIn]
        db = ops.colsum(dG)                                       # b_ih and b_hh enter the gates as a sum
        for d in range(2):
            tape.add_grad(w_ih[d], dwih[d * 4 * H:(d + 1) * 4 * H])
            tape.add_grad(b_ih[d], db[d * 4 * H:(d + 1) * 4 * H])
            tape.add_grad(b_hh[d], db[d * 4 * H:(d + 1) * 4 * H])
            if T > 1:
                # dW_hh[d] = sum_t dgates_t^T h_prev(t): h_prev of time t is OUT[t-1] (forward) / OUT[t+1] (reverse)
                rows_g = slice(Nb, T * Nb) if d == 0 else slice(0, (T - 1) * Nb)
                rows_h = slice(0, (T - 1) * Nb) if d == 0 else slice(Nb, T * Nb)
                tape.add_grad(w_hh[d], ops.linear_bwd_weight(dG[rows_g, d * 4 * H:(d + 1) * 4 * H],
                                                             OUT[rows_h, d * H:(d + 1) * H]))
            else:
                tape.add_grad(w_hh[d], ops.zeros(4 * H, H, like=X))
        tape.add_grad(X, ops.linear_bwd_data(dG, wih))
    tape._push(bwd)
    return OUT


# ------------------------------------------------------------------------------------------------ the module
class CRNN(nn.Module):
    """crnn.py:29-93.  `forward(input)`: gray images [N, nc, 32, W] (NCHW, CUDA fp32) -> logits [W/4 + 1, N, nclass]."""

    _KS = [3, 3, 3, 3, 3, 3, 2]
    _PS = [1, 1, 1, 1, 1, 1, 0]
    _NM = [64, 128, 256, 256, 512, 512, 512]
    _BN = (2, 4, 6)
    _POOL = {0: ((2, 2), (2, 2), (0, 0)), 1: ((2, 2), (2, 2), (0, 0)), 3: ((2, 2), (2, 1), (0, 1)),
             5: ((2, 2), (2, 1), (0, 1))}

    def __init__(self, imgH, nc, nclass, nh, n_rnn=2, leakyRelu=False):
        super().__init__()
        assert imgH % 16 == 0, 'imgH has to be a multiple of 16'
        if leakyRelu:
            raise NotImplementedError("tatt_b200.CRNN implements the reference's default ReLU variant (leakyRelu=False)")
        cnn = nn.Sequential()
        for i in range(7):
            n_in = nc if i == 0 else self._NM[i - 1]
            cnn.add_module('conv{0}'.format(i), nn.Conv2d(n_in, self._NM[i], self._KS[i], 1, self._PS[i]))
            if i in self._BN:
                cnn.add_module('batchnorm{0}'.format(i), nn.BatchNorm2d(self._NM[i]))
            cnn.add_module('relu{0}'.format(i), nn.ReLU(True))
            if i in self._POOL:
                k, s, p = self._POOL[i]
                cnn.add_module('pooling{0}'.format({0: 0, 1: 1, 3: 2, 5: 3}[i]), nn.MaxPool2d(k, s, p))
        self.cnn = cnn
        self.rnn = nn.Sequential(BidirectionalLSTM(512, nh, nh), BidirectionalLSTM(nh, nh, nclass))
        self._nc = nc

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("tatt_b200 runs on sm_100a only: got a %s tensor (no CPU fallback exists)" % input.device)
        if input.dim() != 4 or input.shape[1] != self._nc:
            raise RuntimeError("expected input [N, %d, H, W], got %s" % (self._nc, tuple(input.shape)))
        ps = [p for p in self.parameters()]
        training = self.training
        cnn, rnn = self.cnn, self.rnn

        def build(tape: Tape, t):
            x = ops._chk(t[0].contiguous(), "input image")
            N = x.shape[0]
            x4 = ops.nchw_to_nhwc(x, ops._pad4(x.shape[1]))
            h = x4
            for i in range(7):
                conv = getattr(cnn, 'conv%d' % i)
                has_bn = i in self._BN
                if self._PS[i] == 0:
                    # k x k convolution without padding = same-size convolution (missing taps read 0) + top-left crop
                    full = tape.conv(h, conv.weight, conv.bias, 0, need_dx=True)
                    h = _crop(tape, full, full.shape[1] - self._KS[i] + 1, full.shape[2] - self._KS[i] + 1)
                    assert has_bn               # the only unpadded convolution (conv6) is followed by BatchNorm + ReLU
                else:
                    h = tape.conv(h, conv.weight, conv.bias, self._PS[i], need_dx=(i != 0), relu=not has_bn)
                if has_bn:
                    h = tape.batchnorm(h, getattr(cnn, 'batchnorm%d' % i), ops.ACT_RELU, training)
                if i in self._POOL:
                    h = _maxpool2d(tape, h, *self._POOL[i])
            n_, hh, ww, cc = h.shape
            if hh != 1:
                raise AssertionError("the height of conv must be 1")
            seq = _permute_102(tape, tape.view(h, N, ww, cc))               # [T, N, 512]
            T = ww
            cur = tape.view(seq, T * N, cc)
            for blk in rnn:
                rec = _bilstm(tape, cur, blk.rnn, T, N)
                cur = tape.linear(rec, blk.embedding.weight, blk.embedding.bias)
            logits = cur                                                     # [T*N, nclass]
            outs = [Out(logits, logits.view(T, N, -1), lambda g: g.contiguous().view(T * N, -1))]
            return outs, [In(None)] + [In(p) for p in ps]

        return run_stage(build, [input] + ps)


# ------------------------------------------------------------------------------------------------ glue around it
def parse_crnn_data(imgs_input_: Tensor, in_width: int = 100) -> Tensor:
    """`TextBase.parse_crnn_data` (interfaces/base.py:797-815, ratio_keep=False): bicubic resize of the NCHW batch to
    32 x in_width and RGB -> gray; returns [N, 1, 32, in_width].  Not differentiable (the reference feeds it detached
    images, super_resolution.py:786,793)."""
    x = ops._chk(imgs_input_.detach().contiguous(), "imgs_input_")
    n, c, h, w = x.shape
    out = ops.empty(n, 1, 32, in_width, like=x)
    _cabi.call("tatt_bicubic_gray", ops._p(x), ops._p(out), n, c, h, w, 32, in_width, ops._stream())
    return out


class _SoftmaxPriorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: Tensor):
        logits = ops._chk(logits.contiguous(), "logits")
        T, Nb, C = logits.shape
        probs = torch.empty_like(logits)
        prior = ops.empty(Nb, C, 1, T, like=logits)
        _cabi.call("tatt_softmax_prior_fwd", ops._p(logits), ops._p(probs), ops._p(prior), T, Nb, C, ops._stream())
        ctx.save_for_backward(probs)
        ctx.mark_non_differentiable(prior)
        return probs, prior

    @staticmethod
    def backward(ctx, dprobs: Tensor, _dprior: Optional[Tensor]):
        (probs,) = ctx.saved_tensors
        T, Nb, C = probs.shape
        dl = torch.empty_like(probs)
        _cabi.call("tatt_softmax_bwd", ops._p(probs), ops._p(dprobs.contiguous()), ops._p(dl), T * Nb, C, ops._stream())
        return dl


def softmax_prior(logits: Tensor):
    """super_resolution.py:796-799 in one launch: `label_vecs = softmax(logits, -1)` ([T, N, C], differentiable -- the
    distillation loss consumes it) and `label_vecs_final` = its permutation [N, C, 1, T] (the SR model's text prior; the
    reference passes it `.detach()`ed, so it is returned non-differentiable)."""
    return _SoftmaxPriorFn.apply(logits)


def text_prior(logits: Tensor) -> Tensor:
    return softmax_prior(logits)[1]
