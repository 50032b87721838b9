"""The oracle (oracle/tatt_oracle.py) is pinned two ways: against the committed fixtures generated from the
live reference (everywhere) and against the live reference itself (build container only)."""
import pytest
import torch

import golden_util as gu
from oracle import ref_harness as rh
from oracle import tatt_oracle as orc


def _case(case):
    import tatt_b200
    cls, kw, N, training = gu.CASES[case]
    torch.manual_seed(gu.SEED)
    net = getattr(tatt_b200, cls)(**kw)          # same fresh init as the reference (test_dropin_surface)
    rh.perturb_(net)
    sd = orc.clone_sd(net.state_dict(), requires_grad=training)
    x, tp = orc.synthetic_inputs(N, kw["height"] // 2, kw["width"] // 2, seed=gu.SEED, with_mask=kw.get("mask", True))
    return cls, kw, N, training, sd, x, tp


@pytest.mark.parametrize("case", ["tatt_g16_stn_train_n3", "tatt_g16_eval_n2", "tatt_tiny_rgb_train_n2",
                                  "tsrn_g16_stn_train_n3"])
def test_oracle_matches_reference_fixtures(case):
    cls, kw, N, training, sd, x, tp = _case(case)
    fx = gu.load(case)
    if cls == "TSRN":
        out, block = orc.tsrn_forward(sd, x, training=training, stn=kw["STN"])
        aux = None
    else:
        out, aux, block = orc.tsrn_tl_trans_forward(sd, x, tp, training=training, stn=kw["STN"], dropout_p=0.0)
    gu.check_summary("out", out, fx["out"], 1e-5)
    if aux is not None:
        gu.check_summary("pr_weights", aux["pr_weights"] if training else aux, fx["pr_weights"], 1e-5)
    if training:
        gen = torch.Generator().manual_seed(99)
        (out * torch.randn(out.shape, generator=gen)).sum().backward()
        for n, ref in fx["grads"].items():
            g = sd[n].grad
            if ref is None:
                assert g is None or g.abs().max().item() == 0, n
            else:
                gu.check_summary("grad " + n, g, ref, 1e-4)


@pytest.mark.skipif(not rh.available(), reason="live reference not present (build container only)")
def test_oracle_bit_exact_vs_live_reference():
    ref = rh.load()
    torch.manual_seed(7)
    net = ref.TSRN_TL_TRANS(scale_factor=2, width=128, height=32, STN=True, mask=True)
    rh.zero_dropout(net); rh.perturb_(net); net.train()
    x, tp = orc.synthetic_inputs(3, 16, 64, seed=3)
    sd = orc.clone_sd(net.state_dict(), requires_grad=True)
    out_r, aux_r = net(x, tp)
    out_o, aux_o, blk = orc.tsrn_tl_trans_forward(sd, x, tp, training=True, stn=True, dropout_p=0.0)
    assert torch.equal(out_r, out_o) and torch.equal(aux_r["pr_weights"], aux_o["pr_weights"])
    out_r.mean().backward(); out_o.mean().backward()
    for n, p in net.named_parameters():
        if p.grad is None:
            assert sd[n].grad is None or sd[n].grad.abs().max() == 0
        else:
            assert torch.equal(p.grad, sd[n].grad), n
    for n, b in net.named_buffers():
        assert torch.equal(b, sd[n]), n


def test_psnr_formula():
    a = torch.rand(1, 4, 8, 8)
    assert orc.psnr(a, a) == float("inf")
    assert abs(orc.psnr(a, a + 1.0 / 255) - 20 * torch.log10(torch.tensor(255.0)).item()) < 1e-3
