"""TEST INFRASTRUCTURE ONLY (the oracle) -- CPU fp32 restatement of the TATT/TSRN hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this file.  The product path (`tatt_b200/`) never
does: it fails loudly when the CUDA C-ABI library is missing.

What it restates (all `file:line` under /root/reference):
  * `TSRN_TL_TRANS.forward`            model/tsrn.py:646-692
  * `TSRN.forward`                     model/tsrn.py:131-150
  * `TPInterpreter.forward`            model/tsrn.py:194-224
  * `RecurrentResidualBlock(TL)`       model/tsrn.py:862-871, 892-910
  * `GruBlock.forward`                 model/tsrn.py:1074-1084
  * `UpsampleBLock` / `mish`           model/tsrn.py:1049-1053, 1061-1064
  * `InfoTransformer.forward` (Q1)     model/transformer_v2.py:198-244
  * `TransformerEncoder` (Q2)          model/transformer_v2.py:256-280, 470-484
  * `TransformerDecoder(+Layer_TP)`    model/transformer_v2.py:355-392, 806-833
  * `PositionalEncoding`               model/transformer_v2.py:39-42
  * `STNHead.forward`                  model/stn_head.py:92-106
  * `TPSSpatialTransformer.forward`    model/tps_spatial_transformer.py:97-112

The arithmetic itself lives in PyTorch (third-party: conv2d / batch_norm / gru /
multi_head_attention_forward / layer_norm / grid_sample / pixel_shuffle); the oracle is
"those torch CPU ops in the reference's order" written as pure functions over a
state_dict, so it is also a fair CPU baseline (same kernels the reference dispatches to).

Pinning: `tests/test_oracle.py` checks this file against the live reference
module (in the build container, where /root/reference exists) and against the committed
fixtures in `tests/golden/` (everywhere).  The reference itself ships no golden vectors /
tests, so parity is anchored on reference outputs generated here by
`tests/golden/make_golden.py`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOM = 0.1


# ----------------------------------------------------------------------------- small pieces
def mish(x: Tensor) -> Tensor:  # tsrn.py:1061-1064
    return x * torch.tanh(F.softplus(x))


def _bn(sd, pfx: str, x: Tensor, training: bool) -> Tensor:
    # nn.BatchNorm{1,2}d defaults: eps 1e-5, momentum 0.1; updates running stats in-place
    if training and (pfx + "num_batches_tracked") in sd:
        sd[pfx + "num_batches_tracked"] += 1
    return F.batch_norm(x, sd[pfx + "running_mean"], sd[pfx + "running_var"],
                        sd[pfx + "weight"], sd[pfx + "bias"], training, BN_MOM, BN_EPS)


def _bigru(sd, pfx: str, seq: Tensor) -> Tensor:
    """nn.GRU(num_layers=1, bidirectional=True, batch_first=True), zero initial state."""
    hid = sd[pfx + "weight_hh_l0"].shape[1]
    flat = [sd[pfx + n + s] for s in ("", "_reverse")
            for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    h0 = seq.new_zeros(2, seq.shape[0], hid)
    out, _ = torch._VF.gru(seq, h0, flat, True, 1, 0.0, False, True, True)
    return out


def gru_block(sd, pfx: str, x: Tensor) -> Tensor:
    """GruBlock (tsrn.py:1074-1084): 1x1 conv, then a BiGRU along the LAST spatial axis of
    the NCHW tensor it receives (one sequence per (n, row))."""
    x = F.conv2d(x, sd[pfx + "conv1.weight"], sd[pfx + "conv1.bias"])
    n, c, h, w = x.shape
    seq = x.permute(0, 2, 3, 1).reshape(n * h, w, c)
    out = _bigru(sd, pfx + "gru.", seq)
    return out.reshape(n, h, w, c).permute(0, 3, 1, 2)


def srb(sd, pfx: str, x: Tensor, tp_map: Optional[Tensor], training: bool) -> Tensor:
    """RecurrentResidualBlock (tp_map None, tsrn.py:862-871) / ...TL (tsrn.py:892-910)."""
    r = F.conv2d(x, sd[pfx + "conv1.weight"], sd[pfx + "conv1.bias"], padding=1)
    r = mish(_bn(sd, pfx + "bn1.", r, training))
    r = F.conv2d(r, sd[pfx + "conv2.weight"], sd[pfx + "conv2.bias"], padding=1)
    r = _bn(sd, pfx + "bn2.", r, training)
    if tp_map is not None:
        r = torch.cat([r, tp_map], 1)
    r = gru_block(sd, pfx + "gru1.", r.transpose(-1, -2)).transpose(-1, -2)  # vertical
    return gru_block(sd, pfx + "gru2.", x + r)                               # horizontal


def _dropout(x: Tensor, p: float, training: bool) -> Tensor:
    return F.dropout(x, p, training) if (training and p > 0) else x


def _mha(sd, pfx: str, q, k, v, p_drop: float, training: bool):
    """nn.MultiheadAttention(64, 4) forward, seq-first, need_weights -> head-averaged."""
    return F.multi_head_attention_forward(
        q, k, v, q.shape[-1], 4,
        sd[pfx + "in_proj_weight"], sd[pfx + "in_proj_bias"], None, None, False,
        p_drop if training else 0.0,
        sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"],
        training=training, key_padding_mask=None, need_weights=True, attn_mask=None)


def _ln(sd, pfx: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[pfx + "weight"], sd[pfx + "bias"], 1e-5)


def _ffn(sd, pfx: str, x: Tensor, p: float, training: bool) -> Tensor:
    h = _dropout(F.relu(F.linear(x, sd[pfx + "linear1.weight"], sd[pfx + "linear1.bias"])), p, training)
    return F.linear(h, sd[pfx + "linear2.weight"], sd[pfx + "linear2.bias"])


# ----------------------------------------------------------------------------- TP interpreter
def recurrent_pos_encoding(sd, pfx: str, init_factor: Tensor, bs: int, H: int, W: int) -> Tensor:
    """Q1 (transformer_v2.py:201,215-221): the BiGRU is batch_first but is fed
    [W, bs, H*C] -> it recurs over the BATCH axis.  Returns query_pos [H*W, bs, C]."""
    C = init_factor.shape[1]
    qe = init_factor.unsqueeze(1).repeat(1, bs, 1)                         # [HW, bs, C]
    qe = qe.reshape(H, W, bs, C).permute(1, 2, 0, 3).reshape(W, bs, H * C)
    qe = _bigru(sd, pfx, qe)                                               # batch=W, T=bs
    return qe.reshape(W, bs, H, C).permute(2, 0, 1, 3).reshape(H * W, bs, C)


def tp_interpreter(sd, pfx: str, feat: Tensor, tp: Tensor, training: bool, p: float = 0.1):
    """TPInterpreter.forward (tsrn.py:194-224) -> (tp_map [N,C,H,W], pr_weights [N,HW,26])."""
    N, C, H, W = feat.shape
    tgt = feat.reshape(N, C, H * W).permute(2, 0, 1)                       # [HW, N, C]
    x = tp.permute(0, 3, 1, 2).squeeze(-1)                                 # [N, 26, 37]
    x = F.prelu(F.linear(x, sd[pfx + "fc_in.weight"], sd[pfx + "fc_in.bias"]), sd[pfx + "activation.weight"])
    Nt, L, _ = x.shape
    pos = _dropout(sd[pfx + "pe.pe"][:, :L].expand(Nt, L, C), p, training).permute(1, 0, 2)  # pe(zeros)
    src = x.permute(1, 0, 2)                                               # [26, N, C]
    t = pfx + "transformer."
    qpos = recurrent_pos_encoding(sd, t + "gru_encoding.", sd[pfx + "init_factor.weight"], Nt, H, W)

    # encoder, 1 layer, post-norm; Q2: layer input is output + src == 2*src (transformer_v2.py:261-275)
    e = t + "encoder.layers.0."
    s = src + src
    qk = s + pos
    a, _ = _mha(sd, e + "self_attn.", qk, qk, s, p, training)
    s = _ln(sd, e + "norm1.", s + _dropout(a, p, training))
    s = _ln(sd, e + "norm2.", s + _dropout(_ffn(sd, e, s, p, training), p, training))
    memory = s

    # decoder, 2 layers of TransformerDecoderLayer_TP.forward_post (no self-attention: Q3)
    outs, w = [], None
    out = tgt
    for li in range(2):
        d = t + "decoder.layers.%d." % li
        a, w = _mha(sd, d + "multihead_attn.", out + qpos, memory + pos, memory, p, training)
        out = _ln(sd, d + "norm2.", out + _dropout(a, p, training))
        out = _ln(sd, d + "norm3.", out + _dropout(_ffn(sd, d, out, p, training), p, training))
        outs.append(_ln(sd, t + "decoder.norm.", out))
    hs = torch.stack(outs).mean(0)                                         # [HW, N, C]
    return hs.permute(1, 2, 0).reshape(Nt, C, H, W), w


# ----------------------------------------------------------------------------- STN / TPS
def stn_head(sd, pfx: str, x: Tensor, training: bool):
    """STNHead.forward (stn_head.py:92-106)."""
    pools = {0: (2, 2), 2: (2, 2), 4: (2, 2), 6: (2, 2), 8: ((1, 2), (1, 2))}
    for i in (0, 2, 4, 6, 8, 10):
        c = pfx + "stn_convnet.%d." % i
        x = F.conv2d(x, sd[c + "0.weight"], sd[c + "0.bias"], padding=1)
        x = F.relu(_bn(sd, c + "1.", x, training))
        if i in pools:
            x = F.max_pool2d(x, pools[i][0], pools[i][1])
    x = x.reshape(x.shape[0], -1)
    f = F.linear(x, sd[pfx + "stn_fc1.0.weight"], sd[pfx + "stn_fc1.0.bias"])
    f = F.relu(_bn(sd, pfx + "stn_fc1.1.", f, training))
    c = F.linear(0.1 * f, sd[pfx + "stn_fc2.weight"], sd[pfx + "stn_fc2.bias"])
    return f, c.reshape(-1, sd[pfx + "stn_fc2.bias"].numel() // 2, 2)


def tps_warp(sd, pfx: str, x: Tensor, ctrl: Tensor, out_hw):
    """TPSSpatialTransformer.forward (tps_spatial_transformer.py:97-112); grid_sample is
    bilinear / zeros / align_corners=False (the torch default the reference relies on)."""
    n = ctrl.shape[0]
    Y = torch.cat([ctrl, sd[pfx + "padding_matrix"].expand(n, 3, 2)], 1)
    M = torch.matmul(sd[pfx + "inverse_kernel"], Y)
    src = torch.matmul(sd[pfx + "target_coordinate_repr"], M)
    grid = 2.0 * torch.clamp(src.view(-1, out_hw[0], out_hw[1], 2), 0, 1) - 1.0
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False), src


# ----------------------------------------------------------------------------- full models
def _tail(sd, x: Tensor, srb_nums: int) -> Tensor:
    p = "block%d." % (srb_nums + 3)
    x = F.conv2d(x, sd[p + "0.conv.weight"], sd[p + "0.conv.bias"], padding=1)
    x = mish(F.pixel_shuffle(x, 2))
    return F.conv2d(x, sd[p + "1.weight"], sd[p + "1.bias"], padding=4)


def tsrn_tl_trans_forward(sd: Dict[str, Tensor], x: Tensor, text_emb: Optional[Tensor] = None, *,
                          training: bool = False, stn: bool = False, srb_nums: int = 5,
                          dropout_p: float = 0.1):
    """TSRN_TL_TRANS.forward (tsrn.py:646-692), scale_factor 2.  Mutates BN running stats
    in `sd` when training (like the module).  Returns (output, aux, block) where aux is the
    `ret_mid` dict in training and `pr_weights` in eval."""
    if stn and training:
        _, ctrl = stn_head(sd, "stn_head.", x, training)
        x, _ = tps_warp(sd, "tps.", x, ctrl, x.shape[-2:])
    block = {"1": F.prelu(F.conv2d(x, sd["block1.0.weight"], sd["block1.0.bias"], padding=4), sd["block1.1.weight"])}
    if text_emb is None:
        text_emb = torch.zeros(1, 37, 1, 26)
    tp_map, pr_w = tp_interpreter(sd, "infoGen.", block["1"], text_emb, training, dropout_p)
    for i in range(srb_nums):
        block[str(i + 2)] = srb(sd, "block%d." % (i + 2), block[str(i + 1)], tp_map, training)
    k = srb_nums + 2
    y = F.conv2d(block[str(k - 1)], sd["block%d.0.weight" % k], sd["block%d.0.bias" % k], padding=1)
    block[str(k)] = _bn(sd, "block%d.1." % k, y, training)
    block[str(k + 1)] = _tail(sd, block["1"] + block[str(k)], srb_nums)
    out = torch.tanh(block[str(k + 1)])
    if training:
        aux = {"pr_weights": pr_w, "pr_weights_gt": None, "spatial_t_emb": tp_map,
               "spatial_t_emb_gt": None, "in_feat": block["1"], "trans_feat": tp_map}
    else:
        aux = pr_w
    return out, aux, block


def tsrn_forward(sd: Dict[str, Tensor], x: Tensor, *, training: bool = False, stn: bool = False,
                 srb_nums: int = 5):
    """TSRN.forward (tsrn.py:131-150)."""
    if stn and training:
        _, ctrl = stn_head(sd, "stn_head.", x, training)
        x, _ = tps_warp(sd, "tps.", x, ctrl, x.shape[-2:])
    block = {"1": F.prelu(F.conv2d(x, sd["block1.0.weight"], sd["block1.0.bias"], padding=4), sd["block1.1.weight"])}
    for i in range(srb_nums):
        block[str(i + 2)] = srb(sd, "block%d." % (i + 2), block[str(i + 1)], None, training)
    k = srb_nums + 2
    y = F.conv2d(block[str(k - 1)], sd["block%d.0.weight" % k], sd["block%d.0.bias" % k], padding=1)
    block[str(k)] = _bn(sd, "block%d.1." % k, y, training)
    block[str(k + 1)] = _tail(sd, block["1"] + block[str(k)], srb_nums)
    return torch.tanh(block[str(k + 1)]), block


# ----------------------------------------------------------------------------- inputs / helpers
def synthetic_inputs(N: int, h: int, w: int, seed: int = 1234, with_mask: bool = True):
    """SURVEY 8d synthetic inputs: rgb uniform, mask = gray <= mean (dataset.py:1313-1316),
    text prior = softmax(3*randn) over the 37 classes."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(N, 3, h, w, generator=g)
    tp = torch.softmax(3.0 * torch.randn(N, 37, 1, 26, generator=g), dim=1)
    if not with_mask:
        return rgb, tp
    gray = 0.299 * rgb[:, 0] + 0.587 * rgb[:, 1] + 0.114 * rgb[:, 2]
    mask = (gray <= gray.mean((1, 2), keepdim=True)).float()
    return torch.cat([rgb, mask.unsqueeze(1)], 1), tp


def clone_sd(sd: Dict[str, Tensor], requires_grad: bool = False) -> Dict[str, Tensor]:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and not any(
                s in k for s in ("running_", "pe.pe", "tps.")):
            t.requires_grad_(True)
        out[k] = t
    return out


def psnr(a: Tensor, b: Tensor) -> float:
    """utils/ssim_psnr.py:9-15: 20*log10(255/sqrt(mse)) on x255 of the first 3 channels."""
    mse = ((a[:, :3] * 255 - b[:, :3] * 255) ** 2).mean().item()
    return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))
