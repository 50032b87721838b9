"""Stage-granular autograd boundary of the hot path.

One `torch.autograd.Function` node per stage; inside a stage everything is C-ABI kernel calls
recorded on a `Tape`.  Feature maps cross stage boundaries as NCHW-*logical* tensors whose memory is
channels-last (a permuted view of the [N,H,W,C] buffer the kernels use) -- the reference itself hands
permuted views around (tsrn.py:1083), callers only consume values/shape/dtype.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _cabi, ops
from .tape import Tape

Tensor = torch.Tensor


class Out:
    """raw: tensor the tape keys gradients on; exposed: distinct view handed to autograd;
    to_raw: converts an incoming gradient (shaped like `exposed`) to raw's layout."""

    def __init__(self, raw: Tensor, exposed: Tensor, to_raw: Callable[[Tensor], Tensor], diff: bool = True):
        self.raw, self.exposed, self.to_raw, self.diff = raw, exposed, to_raw, diff


class In:
    """key: tensor the tape accumulated the gradient on (None -> no grad); from_raw: layout fix."""

    def __init__(self, key: Optional[Tensor], from_raw: Callable[[Tensor], Tensor] = lambda g: g):
        self.key, self.from_raw = key, from_raw


def _join_side_at_end_of_backward():
    ops.join_side()


class StageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, builder, *tensors):
        ctx.set_materialize_grads(False)
        record = any(ctx.needs_input_grad[1:])
        tape = Tape(record)
        outs, ins = builder(tape, tensors)
        ctx.tape = tape if record else None
        ctx.outs_raw = [(o.raw, o.to_raw) for o in outs] if record else None
        ctx.ins = ins if record else None
        nd = [o.exposed for o in outs if not o.diff]
        if nd:
            ctx.mark_non_differentiable(*nd)
        res = tuple(o.exposed for o in outs)
        return res if len(res) > 1 else res[0]

    @staticmethod
    def backward(ctx, *gouts):
        tape = ctx.tape
        for (raw, to_raw), g in zip(ctx.outs_raw, gouts):
            if g is not None:
                tape.seed(raw, to_raw(g))
        if getattr(tape, "params_only", False) and ops._side_enabled:
            # a stage whose backward feeds parameters only (the RPE): the whole chain runs on the side stream,
            # concurrently with whatever the engine runs next; joined when the backward pass ends (engine callback),
            # i.e. before any optimizer / packing kernel can read the gradients
            ops.hold_until_join(*tape._g.values())            # the seeded output gradients were made on this stream
            with ops.side_stream():
                tape.backward()
            torch.autograd.Variable._execution_engine.queue_callback(_join_side_at_end_of_backward)
        else:
            tape.backward()
        grads = []
        for i, spec in enumerate(ctx.ins):
            g = None
            if spec.key is not None and ctx.needs_input_grad[i + 1]:
                g = tape.grad(spec.key)
                if g is not None:
                    g = spec.from_raw(g)
            grads.append(g)
        tape._g.clear()
        return (None,) + tuple(grads)


def run_stage(builder, tensors: Sequence[Tensor]):
    return StageFn.apply(builder, *tensors)


# ----------------------------------------------------------------------------- layout helpers
def as_nhwc(x: Tensor) -> Tensor:
    """NCHW-logical -> contiguous [N,H,W,C] (free when the memory is already channels-last)."""
    return x.permute(0, 2, 3, 1).contiguous()


def expose_nchw(raw4: Tensor) -> Tensor:
    return raw4.permute(0, 3, 1, 2)


def fmap_out(raw4: Tensor, diff: bool = True) -> Out:
    return Out(raw4, expose_nchw(raw4), as_nhwc, diff)


def fmap_in(key: Optional[Tensor]) -> In:
    return In(key, expose_nchw)


def params_of(*mods) -> List[Tensor]:
    out = []
    for m in mods:
        out.extend(p for p in m.parameters())
    return out


def _geom_v(N: int, H: int, W: int):   # sequences along H (vertical), one per (n, x)
    return (N * W, H, W, H * W, 1, W)


def _geom_h(N: int, H: int, W: int):   # sequences along W (horizontal), one per (n, y)
    return (N * H, W, 1, W, 0, 1)


# ----------------------------------------------------------------------------- stem (a5)
def stem_stage(block1: torch.nn.Sequential, x: Tensor, x_is_nhwc4: bool):
    """conv9x9(in_planes->64) + PReLU (tsrn.py:596-599).  `x` is the user NCHW image, or -- after the
    STN -- the warped NHWC4 map (differentiable w.r.t. it)."""
    conv, act = block1[0], block1[1]

    def build(tape: Tape, t):
        xin = t[0]
        if x_is_nhwc4:
            x4 = as_nhwc(xin)
        else:
            x4 = ops.nchw_to_nhwc(ops._chk(xin.contiguous(), "input image"), 4)
        y = tape.conv(x4, conv.weight, conv.bias, conv.padding[0], need_dx=x_is_nhwc4)
        o = tape.prelu(y, act.weight)
        return [fmap_out(o)], [fmap_in(x4 if x_is_nhwc4 else None)] + [In(p) for p in params_of(block1)]

    return run_stage(build, [x] + params_of(block1))


# ----------------------------------------------------------------------------- SRB (a11, a12, a17)
def srb_stage(blk: torch.nn.Module, x: Tensor, tp_map: Optional[Tensor], training: bool):
    """RecurrentResidualBlock(TL).forward (tsrn.py:862-871 / 892-910)."""
    ps = params_of(blk)

    def build(tape: Tape, t):
        x4 = as_nhwc(t[0])
        N, H, W, C = x4.shape
        if C != 64:
            raise NotImplementedError("tatt_b200 kernels are specialised for hidden_units=32 (64 channels)")
        tp4 = as_nhwc(t[1]) if tp_map is not None else None
        r = tape.conv(x4, blk.conv1.weight, blk.conv1.bias, 1, bn_next=training)
        r = tape.batchnorm(r, blk.bn1, ops.ACT_MISH, training, planes_for=blk.conv2.weight)   # only consumer: conv2
        r = tape.conv(r, blk.conv2.weight, blk.conv2.bias, 1, bn_next=training)
        r = tape.batchnorm(r, blk.bn2, ops.ACT_NONE, training)
        parts = [tape.view(r, -1, C)] + ([tape.view(tp4, -1, tp4.shape[-1])] if tp4 is not None else [])
        c1 = tape.linear_cat(parts, blk.gru1.conv1.weight, blk.gru1.conv1.bias)
        g1 = tape.bigru32(c1, blk.gru1.gru, *_geom_v(N, H, W))
        s = tape.add(tape.view(x4, -1, C), g1)
        c2 = tape.linear_cat([s], blk.gru2.conv1.weight, blk.gru2.conv1.bias)
        o = tape.bigru32(c2, blk.gru2.gru, *_geom_h(N, H, W))
        in_specs = [fmap_in(x4)] + ([fmap_in(tp4)] if tp4 is not None else []) + [In(p) for p in ps]
        return [Out(o, expose_nchw(o.view(N, H, W, C)), lambda g: as_nhwc(g).view(-1, C))], in_specs

    ins = [x] + ([tp_map] if tp_map is not None else []) + ps
    return run_stage(build, ins)


# ----------------------------------------------------------------------------- RPE (a7, quirk Q1)
rpe_debug: dict = {}


def rpe_stage(init_factor: torch.nn.Embedding, gru: torch.nn.GRU, N: int, H: int, W: int):
    """query_pos [N, H*W, 64] of InfoTransformer.forward (transformer_v2.py:201,215-221): the BiGRU
    recurs over the BATCH axis (batch_first GRU fed [W, bs, H*C]).  Its input is identical at every
    step, so W_ih x is computed once (SURVEY 8d 'hoisted'); the recurrence is N sequential steps of a
    [W x Hd] x [Hd x 3Hd] GEMM per direction."""
    emb = init_factor.weight
    w_ih = (gru.weight_ih_l0, gru.weight_ih_l0_reverse)
    w_hh = (gru.weight_hh_l0, gru.weight_hh_l0_reverse)
    b_ih = (gru.bias_ih_l0, gru.bias_ih_l0_reverse)
    b_hh = (gru.bias_hh_l0, gru.bias_hh_l0_reverse)
    ps = [emb, w_ih[0], w_hh[0], b_ih[0], b_hh[0], w_ih[1], w_hh[1], b_ih[1], b_hh[1]]

    def build(tape: Tape, t):
        C = emb.shape[1]
        I = H * C
        Hd = I // 2
        Wd = W
        if emb.shape[0] != H * W or w_ih[0].shape != (3 * Hd, I):
            raise RuntimeError("TPInterpreter was built for a different feature size than the input "
                               "(init_factor %s vs H*W=%d)" % (tuple(emb.shape), H * W))
        st = ops._stream
        X = ops.empty(Wd, I, like=emb)
        _cabi.call("tatt_rpe_gather", emb.data_ptr(), X.data_ptr(), H, W, C, st())
        GI = ops.empty(2, Wd, 3 * Hd, like=emb)
        for d in range(2):
            ops.linear_fwd(X, w_ih[d], b_ih[d], out=GI[d])
        _, pk = ops.packed([w_hh[0], w_hh[1], b_hh[0], b_hh[1]], emb, want_flat=True)   # one launch, not 4 memcpy nodes
        WHH = pk[:2 * 3 * Hd * Hd].view(2, 3 * Hd, Hd)
        BHH = pk[2 * 3 * Hd * Hd:].view(2, 3 * Hd)
        HALL = ops.zeros(2, N + 1, Wd, Hd, like=emb)
        GATES = ops.empty(N, 2, Wd, 4, Hd, like=emb) if tape.record else None
        GH = ops.empty(2, Wd, 3 * Hd, like=emb)
        QPOS = ops.empty(N, H * W, C, like=emb)
        sH = (N + 1) * Wd * Hd
        # operands of the N sequential recurrent GEMMs as pre-split bf16 hi/lo planes: W_hh is split once, the
        # hidden state is written in plane form by the gate kernel itself
        fast = Hd % 8 == 0 and Wd >= 32 and not (ops._precision_flag & ops.F_FP32)
        if fast:
            WP = ops.Planes((2, 3 * Hd, Hd), emb)
            for d in range(2):
                WP.split_from(w_hh[d], d)
            HP = ops.Planes((2, N + 1, Wd, Hd), emb, zero=True)
        persist = fast and _cabi.lib().tatt_rpe_persist_supported(N, Wd, Hd, C) == 0
        if persist:
            # ONE cooperative launch for all N steps of both directions (csrc/tc5_rpe.cu)
            sync = torch.empty(_cabi.lib().tatt_rpe_sync_bytes(N) // 4, dtype=torch.int32, device=emb.device)
            _cabi.call("tatt_rpe_fwd", GI.data_ptr(), BHH.data_ptr(), WP.t.data_ptr(), WP.lo_off, HP.t.data_ptr(),
                       HP.lo_off, HALL.data_ptr(), None if GATES is None else GATES.data_ptr(), QPOS.data_ptr(),
                       sync.data_ptr(), N, Wd, Hd, C, H, st())
            rpe_debug["fwd_sync"] = sync              # sync[2N] != 0 after the run: a bounded barrier spin gave up
        for s in range(0 if persist else N):
            if fast:
                ops.gemm(0, 1, HP.t[0, 0, s], Hd, WP.t, Hd, GH, 3 * Hd, BHH, Wd, 3 * Hd, Hd,
                         ops.F_APLANES | ops.F_BPLANES, batch=2, sA=sH, sB=3 * Hd * Hd, sC=Wd * 3 * Hd, sBias=3 * Hd,
                         loA=HP.lo_off, loB=WP.lo_off)
            else:
                ops.gemm(0, 1, HALL[0, s], Hd, WHH, Hd, GH, 3 * Hd, BHH, Wd, 3 * Hd, Hd, 0, batch=2, sA=sH,
                         sB=3 * Hd * Hd, sC=Wd * 3 * Hd, sBias=3 * Hd)
            _cabi.call("tatt_rpe_gate_fwd", GI.data_ptr(), GH.data_ptr(), HALL.data_ptr(),
                       None if GATES is None else GATES.data_ptr(), QPOS.data_ptr(),
                       HP.t.data_ptr() if fast else None, HP.lo_off if fast else 0, s, N, Wd, Hd, C, H, st())
        if fast:
            del WP
            if not tape.record:
                del HP                     # training keeps the hidden-state planes: B operand of the dW_hh GEMM

        def bwd():
            dQ = tape.grad(QPOS)
            if dQ is None:
                return
            DH = ops.zeros(2, Wd, Hd, like=emb)
            DGI = ops.zeros(2, Wd, 3 * Hd, like=emb)
            DGH = ops.empty(2, N, Wd, 3 * Hd, like=emb)
            if fast:
                WTP = ops.Planes((2, Hd, 3 * Hd), emb)              # W_hh^T: B operand [N=Hd][K=3Hd] of dh += dgh W_hh
                for d in range(2):
                    WTP.split_from(w_hh[d], d, transpose=True)
                DP = ops.Planes((2, N, Wd, 3 * Hd), emb)
            for s in range(N - 1, -1, -1):
                _cabi.call("tatt_rpe_gate_bwd", dQ.data_ptr(), HALL.data_ptr(), GATES.data_ptr(), DH.data_ptr(),
                           DGI.data_ptr(), DGH.data_ptr(), DP.t.data_ptr() if fast else None,
                           DP.lo_off if fast else 0, s, N, Wd, Hd, C, H, st())
                if s > 0 and fast:
                    ops.gemm(0, 1, DP.t[0, 0, s], 3 * Hd, WTP.t, 3 * Hd, DH, Hd, None, Wd, Hd, 3 * Hd,
                             ops.F_ACCUM | ops.F_APLANES | ops.F_BPLANES, batch=2, sA=N * Wd * 3 * Hd, sB=3 * Hd * Hd,
                             sC=Wd * Hd, loA=DP.lo_off, loB=WTP.lo_off)
                elif s > 0:
                    ops.gemm(0, 0, DGH[0, s], 3 * Hd, WHH, Hd, DH, Hd, None, Wd, Hd, 3 * Hd, ops.F_ACCUM, batch=2,
                             sA=N * Wd * 3 * Hd, sB=3 * Hd * Hd, sC=Wd * Hd)
            dWHH = ops.empty(2, 3 * Hd, Hd, like=emb)
            if fast:
                # dW_hh = sum_s dgh_s^T h_{s-1}: both operands already exist as bf16 planes (dgh: written by the gate
                # kernel for the chain GEMMs; h: written by the forward pass) -- no split passes over 270 MB
                ops.gemm(1, 0, DP.t[0], 3 * Hd, HP.t[0], Hd, dWHH, Hd, None, 3 * Hd, Hd, N * Wd,
                         ops.F_SPLITK | ops.F_ZEROC | ops.F_APLANES | ops.F_BPLANES, batch=2, sA=N * Wd * 3 * Hd, sB=sH,
                         sC=3 * Hd * Hd, loA=DP.lo_off, loB=HP.lo_off)
            else:
                ops.gemm(1, 0, DGH, 3 * Hd, HALL, Hd, dWHH, Hd, None, 3 * Hd, Hd, N * Wd, ops.F_SPLITK | ops.F_ZEROC,
                         batch=2, sA=N * Wd * 3 * Hd, sB=sH, sC=3 * Hd * Hd)
            dX = None
            for d in range(2):
                tape.add_grad(w_hh[d], dWHH[d])
                tape.add_grad(b_hh[d], ops.colsum(DGH[d].view(N * Wd, 3 * Hd)))
                dWih = ops.empty(3 * Hd, I, like=emb)
                ops.gemm(1, 0, DGI[d], 3 * Hd, X, I, dWih, I, None, 3 * Hd, I, Wd, 0)
                tape.add_grad(w_ih[d], dWih)
                tape.add_grad(b_ih[d], ops.colsum(DGI[d]))
                dX = ops.linear_bwd_data(DGI[d], w_ih[d], out=dX, accumulate=d > 0)
            demb = torch.empty_like(emb)
            _cabi.call("tatt_rpe_scatter", dX.data_ptr(), demb.data_ptr(), H, W, C, st())
            tape.add_grad(emb, demb)
        tape._push(bwd)
        tape.params_only = True
        return [Out(QPOS, QPOS.view(N, H * W, C), lambda g: g.contiguous())], [In(p) for p in ps]

    return run_stage(build, ps)


# ----------------------------------------------------------------------------- TP interpreter (a6, a8-a10)
def tp_stage(ig: torch.nn.Module, feat: Tensor, text_emb: Tensor, qpos: Tensor, training: bool):
    """TPInterpreter.forward minus the RPE (tsrn.py:194-224; transformer_v2.py:240-243, 256-280,
    355-392, 470-484, 806-833).  Token tensors are [N*L, 64] (token-major == NHWC memory)."""
    tr = ig.transformer
    enc = tr.encoder.layers[0]
    decs = list(tr.decoder.layers)
    live = [ig.fc_in, ig.activation, enc] + [m for d in decs for m in
                                             (d.multihead_attn, d.linear1, d.linear2, d.norm2, d.norm3)]
    live.append(tr.decoder.norm)
    ps = params_of(*live)

    def build(tape: Tape, t):
        f4 = as_nhwc(t[0])
        N, H, W, C = f4.shape
        te = ops._chk(t[1].contiguous(), "text_emb")
        Nt, Ct, one, L = te.shape
        if Nt != N:
            raise RuntimeError("text prior batch (%d) != image batch (%d); the reference only broadcasts "
                               "the default zeros prior for N == 1" % (Nt, N))
        if not t[2].is_contiguous():
            ops.join_side()            # a copy kernel would read qpos before the join below
        qp_raw = t[2].contiguous()
        qp = tape.view(qp_raw, N * H * W, C)
        tgt = tape.view(f4, N * H * W, C)
        p = {}
        rng = None
        if training:
            for name, m in (("pe", ig.pe.dropout), ("e1", enc.dropout1), ("e", enc.dropout), ("e2", enc.dropout2)):
                p[name] = float(m.p)
            p["ea"] = float(enc.self_attn.dropout)
            for i, d in enumerate(decs):
                p["d%da" % i] = float(d.multihead_attn.dropout)
                p["d%d2" % i] = float(d.dropout2.p)
                p["d%d" % i] = float(d.dropout.p)
                p["d%d3" % i] = float(d.dropout3.p)
            if any(v > 0 for v in p.values()):
                rng = ops.DeviceRNG.get(f4.device).snapshot()
        drop = lambda x, key, site: tape.dropout(x, p.get(key, 0.0), rng, site)

        # text prior -> tokens [N*26, 37] -> fc_in + PReLU
        xt = tape.view(tape.to_nhwc(te, Ct), N * L, Ct)
        src = tape.prelu(tape.linear(xt, ig.fc_in.weight, ig.fc_in.bias), ig.activation.weight)
        pe = ig.pe.pe[0, :L].contiguous()
        pos = drop(ops.add_bcast_rows(None, pe, N * L, L, C), "pe", 1)

        # encoder layer on 2*src (Q2)
        s2 = tape.scale(src, 2.0)
        qk = tape.add(s2, pos)
        a, _ = tape.mha(qk, qk, s2, enc.self_attn, N, L, L, False, p.get("ea", 0.0), rng, 2)
        s3 = tape.add_layernorm(s2, drop(a, "e1", 3), enc.norm1)
        f = drop(tape.linear(s3, enc.linear1.weight, enc.linear1.bias, relu=True), "e", 4)
        f = drop(tape.linear(f, enc.linear2.weight, enc.linear2.bias), "e2", 5)
        mem = tape.add_layernorm(s3, f, enc.norm2)

        # decoder layers (cross attention only: Q3)
        kin = tape.add(mem, pos)
        out = tgt
        inter = []
        aw = None
        fused = ops.declayer_ok(N, H * W, L) and C == 64 and all(dd.linear1.weight.shape == (64, 64) for dd in decs)
        ops.join_side()                # qpos may still be in flight on the side stream (tsrn.TSRN_TL_TRANS.forward)
        for i, d in enumerate(decs):
            last = i == len(decs) - 1
            if fused:          # one tcgen05 kernel per decoder layer (csrc/tc6_declayer.cu)
                pd = (p.get("d%da" % i, 0.0), p.get("d%d2" % i, 0.0), p.get("d%d" % i, 0.0), p.get("d%d3" % i, 0.0))
                out, it_, w_ = tape.dec_layer(out, qp, kin, mem, d, tr.decoder.norm, N, H * W, L, last, pd, rng,
                                              (10 + 8 * i, 11 + 8 * i, 12 + 8 * i, 13 + 8 * i))
                inter.append(it_)
                if last:
                    aw = w_
                continue
            qin = tape.add(out, qp)
            a, w_ = tape.mha(qin, kin, mem, d.multihead_attn, N, H * W, L, last, p.get("d%da" % i, 0.0), rng,
                             10 + 8 * i)
            if last:
                aw = w_
            t1 = tape.add_layernorm(out, drop(a, "d%d2" % i, 11 + 8 * i), d.norm2)
            f = drop(tape.linear(t1, d.linear1.weight, d.linear1.bias, relu=True), "d%d" % i, 12 + 8 * i)
            f = drop(tape.linear(f, d.linear2.weight, d.linear2.bias), "d%d3" % i, 13 + 8 * i)
            out = tape.add_layernorm(t1, f, d.norm3)
            inter.append(tape.add_layernorm(out, None, tr.decoder.norm))
        if len(inter) != 2:
            raise NotImplementedError("TP interpreter is specialised for 2 decoder layers")
        tpm = tape.mean2(inter[0], inter[1])                      # text_prior.mean(0), tsrn.py:219
        tpm4 = tpm.view(N, H, W, C)
        outs = [Out(tpm, expose_nchw(tpm4), lambda g: as_nhwc(g).view(-1, C)),
                Out(aw, aw.view(N, H * W, L), lambda g: g, diff=False)]
        ins = [fmap_in(f4), In(te), In(qp_raw)] + [In(q) for q in ps]
        return outs, ins

    return run_stage(build, [feat, text_emb, qpos] + ps)


# ----------------------------------------------------------------------------- block7 (a14)
def conv_bn_stage(seq: torch.nn.Sequential, x: Tensor, training: bool):
    """conv3x3 + BatchNorm (tsrn.py:609-614)."""
    conv = seq[0]
    bn = seq[1] if len(seq) > 1 else None
    ps = params_of(seq)

    def build(tape: Tape, t):
        x4 = as_nhwc(t[0])
        y = tape.conv(x4, conv.weight, conv.bias, conv.padding[0], bn_next=(bn is not None and training))
        if bn is not None:
            y = tape.batchnorm(y, bn, ops.ACT_NONE, training)
        return [fmap_out(y)], [fmap_in(x4)] + [In(p) for p in ps]

    return run_stage(build, [x] + ps)


# ----------------------------------------------------------------------------- tail (a15, a16)
def tail_stage(seq: torch.nn.Sequential, skip_a: Tensor, skip_b: Tensor, out_planes: int):
    """(block1 + block7) -> UpsampleBLock(s) (conv3x3 64->256, PixelShuffle(2), mish) -> conv9x9 -> tanh
    (tsrn.py:672-675, 1049-1053).  Returns (tanh output NCHW, pre-tanh NCHW (non-differentiable view))."""
    ps = params_of(seq)
    ups = [m for m in seq if hasattr(m, "conv")]
    final = seq[len(seq) - 1]

    def build(tape: Tape, t):
        a4, b4 = as_nhwc(t[0]), as_nhwc(t[1])
        s = tape.add(a4, b4)
        for i, u in enumerate(ups):
            nxt = ups[i + 1].conv if i + 1 < len(ups) else final            # the only consumer of the up-sampled map
            s = tape.pixshuf_mish(tape.conv(s, u.conv.weight, u.conv.bias, u.conv.padding[0]), planes_for=nxt.weight)
        pre = tape.conv(s, final.weight, final.bias, final.padding[0])
        out = tape.to_nchw(pre, out_planes, tanh=True)
        pre_view = expose_nchw(pre)[:, :out_planes]
        outs = [Out(out, out.view(out.shape), lambda g: g.contiguous()),
                Out(pre, pre_view, lambda g: g, diff=False)]
        return outs, [fmap_in(a4), fmap_in(b4)] + [In(p) for p in ps]

    return run_stage(build, [skip_a, skip_b] + ps)


# ----------------------------------------------------------------------------- STN + TPS (a3, a4)
_stn_tc = __import__("os").environ.get("TATT_STN_TC", "1") != "0"


def stn_tps_stage(stn_head: torch.nn.Module, tps: torch.nn.Module, x: Tensor, training: bool):
    """STNHead.forward (stn_head.py:92-106) + TPSSpatialTransformer.forward
    (tps_spatial_transformer.py:97-112).  Returns the warped image as an NHWC4-backed NCHW view."""
    ps = params_of(stn_head)
    pools = {0: (2, 2), 2: (2, 2), 4: (2, 2), 6: (2, 2), 8: (1, 2)}

    def build(tape: Tape, t):
        xin = ops._chk(t[0].contiguous(), "input image")
        N, Cin, H, W = xin.shape
        x4 = ops.nchw_to_nhwc(xin, 4)
        h = x4
        # the localisation convnet runs on the tensor-core engines like the trunk (fp32 parity via the bf16 hi/lo split)
        # unless TATT_STN_TC=0; the head from fc1 on stays on the fp32 FFMA kernels: BatchNorm1d over the N samples
        # of the batch divides by a tiny batch deviation and amplifies operand rounding of fc1 by orders of magnitude
        conv_scope = tape.precision_scope(0) if _stn_tc else tape.fp32_scope()      # never single-plane bf16
        with conv_scope:
            for i in (0, 2, 4, 6, 8, 10):
                blk = stn_head.stn_convnet[i]
                h = tape.conv(h, blk[0].weight, blk[0].bias, 1, need_dx=(i != 0))
                h = tape.batchnorm(h, blk[1], ops.ACT_RELU, training)
                if i in pools:
                    h = tape.maxpool(h, *pools[i])
        with tape.fp32_scope():
            n_, hh, ww, cc = h.shape
            flat = tape.view(tape.to_nchw(h, cc), N, cc * hh * ww)     # NCHW flatten order (stn_head.py:95)
            fc1, bn1 = stn_head.stn_fc1[0], stn_head.stn_fc1[1]
            if flat.shape[1] != fc1.weight.shape[1]:
                raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (
                    N, flat.shape[1], fc1.weight.shape[1], fc1.weight.shape[0]))
            f = tape.linear(flat, fc1.weight, fc1.bias)
            if training and N < 2:
                raise ValueError("Expected more than 1 value per channel when training, got input size %s" % (
                    [N, f.shape[1]],))
            f = tape.batchnorm(f, bn1, ops.ACT_RELU, training)
            c = tape.linear(tape.scale(f, 0.1), stn_head.stn_fc2.weight, stn_head.stn_fc2.bias)   # [N, 40]
            if tuple(tps.target_coordinate_repr.shape) != (H * W, 23):
                raise RuntimeError("TPS grid was built for a different output size")
            invk, rep = tps.inverse_kernel.contiguous(), tps.target_coordinate_repr.contiguous()
            warped, _ = ops.tps_sample_fwd(x4, c, invk, rep)

            def bwd():
                dw = tape.grad(warped)
                if dw is not None:
                    tape.add_grad(c, ops.tps_sample_bwd(x4, c, invk, rep, dw))
            tape._push(bwd)
        return [fmap_out(warped)], [In(None)] + [In(p) for p in ps]

    return run_stage(build, [x] + ps)
