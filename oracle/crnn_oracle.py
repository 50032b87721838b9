"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (tatt_b200/).

CPU restatement of the text-prior generator that feeds the hot path (the "next" row before the path):
`model/crnn/crnn.py:CRNN` (lines 29-93) with `BidirectionalLSTM` (:5-26), the pre-processing
`interfaces/base.py:parse_crnn_data` (:797-815: bicubic resize to 32 x 100, RGB -> gray) and the softmax / permute that
turns its logits into the `[N, 37, 1, 26]` text prior (`interfaces/super_resolution.py:794-799`).

State-dict functional form (same ATen ops the reference dispatches to), so it is bit-exact against the live module:
pinned by tests/test_oracle.py (live reference, build container only) and tests/golden/crnn_n3.pt (everywhere).
No CUDA twin exists yet -- this is the parity anchor for the round that builds it."""
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# (has_bn, pool) per conv of CRNN.__init__ (crnn.py:35-68); pool = None | (kernel, stride, padding)
_LAYERS = [(False, ((2, 2), (2, 2), (0, 0))), (False, ((2, 2), (2, 2), (0, 0))), (True, None),
           (False, ((2, 2), (2, 1), (0, 1))), (True, None), (False, ((2, 2), (2, 1), (0, 1))), (True, None)]
_PADS = [1, 1, 1, 1, 1, 1, 0]


def parse_crnn_data(imgs: Tensor, in_width: int = 100) -> Tensor:
    """interfaces/base.py:797-815 (ratio_keep=False): bicubic resize to 32 x in_width, then 0.299 R + 0.587 G + 0.114 B"""
    x = F.interpolate(imgs, (32, in_width), mode="bicubic")
    return 0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3]


def _bilstm(sd: Dict[str, Tensor], pfx: str, seq: Tensor) -> Tensor:
    """BidirectionalLSTM.forward (crnn.py:13-26): nn.LSTM(nIn, nHidden, bidirectional=True) over [T, b, nIn], then
    Linear(2 nHidden -> nOut) on every time step"""
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    flat = [sd[pfx + "rnn." + n] for n in names] + [sd[pfx + "rnn." + n + "_reverse"] for n in names]
    nh = flat[1].shape[1]
    h0 = seq.new_zeros(2, seq.shape[1], nh)
    out, _, _ = torch._VF.lstm(seq, (h0, h0.clone()), flat, True, 1, 0.0, False, True, False)
    T, b, h = out.shape
    return F.linear(out.reshape(T * b, h), sd[pfx + "embedding.weight"], sd[pfx + "embedding.bias"]).view(T, b, -1)


def crnn_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = False) -> Tensor:
    """CRNN.forward (crnn.py:74-93): gray image [N, 1, 32, W] -> logits [W/4 + 1, N, nclass].  In training mode the
    three BatchNorm2d layers use batch statistics and update `running_*` in `sd` (like the module)."""
    h = x
    for i, (bn, pool) in enumerate(_LAYERS):
        h = F.conv2d(h, sd["cnn.conv%d.weight" % i], sd["cnn.conv%d.bias" % i], 1, _PADS[i])
        if bn:
            p = "cnn.batchnorm%d." % i
            if training:
                sd[p + "num_batches_tracked"] += 1
            h = F.batch_norm(h, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                             training, 0.1, 1e-5)
        h = F.relu(h)
        if pool is not None:
            h = F.max_pool2d(h, pool[0], pool[1], pool[2])
    b, c, hh, w = h.shape
    assert hh == 1, "the height of conv must be 1"
    seq = h.squeeze(2).permute(2, 0, 1)                      # [w, b, c]
    return _bilstm(sd, "rnn.1.", _bilstm(sd, "rnn.0.", seq))


def text_prior(logits: Tensor) -> Tensor:
    """super_resolution.py:796-799: softmax over the classes, [T, N, C] -> [N, C, 1, T]"""
    return F.softmax(logits, -1).permute(1, 0, 2).unsqueeze(1).permute(0, 3, 1, 2)


def make_state_dict(seed: int, nc: int = 1, nclass: int = 37, nh: int = 256) -> Dict[str, Tensor]:
    """A fresh `CRNN(32, nc, nclass, nh)` state_dict WITHOUT the reference: the same torch.nn leaf modules created in the
    same order as crnn.py:35-72 under the same seed give bit-identical initial values (checked against the live class in
    tests/test_oracle.py).  Keys follow the reference (`cnn.conv0.weight`, `cnn.batchnorm2.running_mean`,
    `rnn.0.rnn.weight_ih_l0`, `rnn.0.embedding.weight`, ...)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    nm = [64, 128, 256, 256, 512, 512, 512]
    ks = [3, 3, 3, 3, 3, 3, 2]
    sd: Dict[str, Tensor] = {}
    for i in range(7):
        conv = nn.Conv2d(nc if i == 0 else nm[i - 1], nm[i], ks[i], 1, _PADS[i])
        for k, v in conv.state_dict().items():
            sd["cnn.conv%d.%s" % (i, k)] = v
        if _LAYERS[i][0]:
            bn = nn.BatchNorm2d(nm[i])
            for k, v in bn.state_dict().items():
                sd["cnn.batchnorm%d.%s" % (i, k)] = v
    for j, (n_in, n_out) in enumerate(((512, nh), (nh, nclass))):
        rnn = nn.LSTM(n_in, nh, bidirectional=True)
        emb = nn.Linear(nh * 2, n_out)
        for k, v in rnn.state_dict().items():
            sd["rnn.%d.rnn.%s" % (j, k)] = v
        for k, v in emb.state_dict().items():
            sd["rnn.%d.embedding.%s" % (j, k)] = v
    return sd


def perturb_bn_(sd: Dict[str, Tensor], seed: int) -> None:
    """non-trivial BatchNorm statistics / affine parameters (same recipe as tests/golden/make_golden_crnn.py:build)"""
    g = torch.Generator().manual_seed(seed)
    for i in (2, 4, 6):
        p = "cnn.batchnorm%d." % i
        n = sd[p + "weight"].numel()
        sd[p + "running_mean"].copy_(0.1 * torch.randn(n, generator=g))
        sd[p + "running_var"].copy_(0.5 + torch.rand(n, generator=g))
        sd[p + "weight"].copy_(1 + 0.2 * torch.randn(n, generator=g))
        sd[p + "bias"].copy_(0.1 * torch.randn(n, generator=g))
