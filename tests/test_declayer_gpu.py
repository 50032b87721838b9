"""The fused tcgen05 decoder-layer kernel (csrc/tc6_declayer.cu: TransformerDecoderLayer_TP.forward_post,
/root/reference/model/transformer_v2.py:806-833) against the round-1 op-by-op path of the same layer, which is itself
pinned to the oracle / live-reference fixtures by tests/test_model_gpu.py.  With dropout ON the two paths must still
agree, because the fused kernel draws bit-identical masks (same Philox counters as tatt_dropout / tatt_mha64_*); this
is the only check of the dropout-on forward that is not statistical.  Tolerance 2e-4 of max-abs (both are fp32-parity
tensor-core paths; accumulation order differs).  Gradients: 5e-3 rel-L2 per parameter and 1e-3 over the whole gradient.
The per-parameter bound is what the FFN's ReLU allows, not what the GEMMs deliver: two forward passes that agree to 3e-5
still disagree on the sign of a handful of the N*Lq*64 pre-activations, and every flipped mask element moves the
gradients upstream of that ReLU by a full element -- measured against the fp64 oracle (tools/diag_declayer.py, G16, N = 3)
BOTH paths sit at 0.5-2e-3 on exactly those parameters (decoder layer 0 from linear1 upstream, encoder, fc_in, RPE) and
at 2-4e-5 everywhere else, while they agree with each other to 2e-5 downstream of the flips."""
import pytest
import torch

import re

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_ZERO_GRAD = re.compile(r"^(block[2-6]\.conv[12]\.bias|block7\.0\.bias)$")


def _run(fused, kw, N, p_drop, seed=77):
    import tatt_b200
    from oracle import ref_harness as rh
    from oracle import tatt_oracle as orc
    from tatt_b200 import ops
    torch.manual_seed(1234)
    net = tatt_b200.TSRN_TL_TRANS(**kw)
    rh.perturb_(net)
    if p_drop == 0.0:
        rh.zero_dropout(net)
    net = net.to(DEV).train()
    x, tp = orc.synthetic_inputs(N, kw["height"] // 2, kw["width"] // 2, seed=5, with_mask=kw["mask"])
    tatt_b200.manual_seed(seed)
    old = ops._declayer_enabled
    ops._declayer_enabled = fused
    try:
        out, aux = net(x.to(DEV), tp.to(DEV))
        wgt = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).to(DEV)
        (out * wgt).sum().backward()
    finally:
        ops._declayer_enabled = old
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
    return out.detach(), aux["pr_weights"].detach(), aux["spatial_t_emb"].detach(), grads


@pytest.mark.parametrize("geom,N,p_drop", [("g16", 3, 0.0), ("g16", 3, 0.1), ("g32", 2, 0.1), ("tiny", 2, 0.1)])
def test_fused_decoder_layer_equals_op_by_op_path(geom, N, p_drop):
    kw = {"g16": dict(scale_factor=2, width=128, height=32, STN=False, mask=True),
          "g32": dict(scale_factor=2, width=256, height=64, STN=False, mask=True),
          "tiny": dict(scale_factor=2, width=32, height=16, STN=False, mask=False)}[geom]
    a = _run(True, kw, N, p_drop)
    b = _run(False, kw, N, p_drop)
    for name, u, v in (("out", a[0], b[0]), ("pr_weights", a[1], b[1]), ("tp_map", a[2], b[2])):
        err = (u - v).abs().max().item() / max(v.abs().max().item(), 1e-12)
        assert err <= 2e-4, "%s: fused vs op-by-op %.3e" % (name, err)
    assert (a[1].sum(-1).mean() - 1).abs().item() < (1e-4 if p_drop == 0 else 2e-2)
    assert set(a[3]) == set(b[3])
    G = max(v.abs().max().item() for v in b[3].values())
    bad = []
    num2 = den2 = 0.0
    for n in b[3]:
        d = (a[3][n] - b[3][n]).double()
        num2 += d.pow(2).sum().item()
        den2 += b[3][n].double().pow(2).sum().item()
        if _ZERO_GRAD.match(n):
            # a conv bias in front of a train-mode BatchNorm has an exactly-zero gradient: what either path holds is
            # summation noise (1e-7 of the sibling weight gradient), not a quantity two paths can agree on
            assert a[3][n].abs().max().item() <= 1e-4 * G, n
            continue
        rel = d.norm().item() / max(b[3][n].double().norm().item(), 1e-6 * G * d.numel() ** 0.5)
        # a 1-element parameter (the PReLU slope) is one heavily cancelling sum over all tokens: 2e-2
        if rel > (5e-3 if d.numel() > 1 else 2e-2):
            bad.append("%s %.3e" % (n, rel))
    assert not bad, "fused vs op-by-op gradient rel-L2 > 5e-3: " + ", ".join(bad)
    assert (num2 / den2) ** 0.5 <= 1e-3


def test_fused_decoder_layer_eval_matches_op_by_op():
    import tatt_b200
    from oracle import ref_harness as rh
    from oracle import tatt_oracle as orc
    from tatt_b200 import ops
    torch.manual_seed(1234)
    net = tatt_b200.TSRN_TL_TRANS(scale_factor=2, width=128, height=32, STN=False, mask=True)
    rh.perturb_(net)
    net = net.to(DEV).eval()
    x, tp = orc.synthetic_inputs(4, 16, 64, seed=5)
    res = []
    for fused in (True, False):
        ops._declayer_enabled = fused
        try:
            with torch.no_grad():
                res.append(net(x.to(DEV), tp.to(DEV)))
        finally:
            ops._declayer_enabled = True
    for u, v in zip(res[0], res[1]):
        assert (u - v).abs().max().item() / v.abs().max().item() <= 2e-4
