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


# ---------------------------------------------------------------------------------- loss block
def test_loss_oracle_matches_reference_fixture():
    """oracle/loss_oracle.py vs the fixture generated from the live `loss/image_loss.py:ImageLoss` (bit-exact)"""
    import os
    from oracle import loss_oracle as lo
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_loss_n3.pt"))
    o = fx["out"].clone().requires_grad_(True)
    loss = lo.image_loss(o, fx["target"])
    (loss.mean() * 100).backward()
    assert torch.equal(loss, fx["loss"]) and torch.equal(o.grad, fx["dout"])
    assert o.grad[0, :3, 4, 6].abs().max().item() < 1e-3            # flat patch: only the MSE term (zero there) acts
    o64 = fx["out"].double().requires_grad_(True)
    lo.training_scalar(o64, fx["target"].double()).backward()
    assert torch.allclose(o64.grad, fx["dout64"], rtol=0, atol=1e-12)


@pytest.mark.skipif(not rh.available(), reason="live reference not present (build container only)")
def test_loss_oracle_bit_exact_vs_live_reference():
    import importlib
    import warnings
    from oracle import loss_oracle as lo
    rh.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("loss.image_loss")
        crit = mod.ImageLoss(gradient=True, loss_weight=[20, 1e-4])
    g = torch.Generator().manual_seed(5)
    out = torch.tanh(torch.randn(2, 4, 9, 21, generator=g)).requires_grad_(True)
    tgt = torch.rand(2, 4, 9, 21, generator=g)
    l_ref = crit(out, tgt)
    (l_ref.mean() * 100).backward()
    g_ref = out.grad.clone()
    out.grad = None
    l_o = lo.image_loss(out, tgt)
    (l_o.mean() * 100).backward()
    assert torch.equal(l_ref, l_o) and torch.equal(g_ref, out.grad)


# ---------------------------------------------------------------------------------- CRNN text-prior generator
def _crnn_sd(training):
    from oracle import crnn_oracle as co
    sd = co.make_state_dict(1234)
    co.perturb_bn_(sd, 1235)
    sd = {k: v.detach().clone() for k, v in sd.items()}
    if training:
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    return sd


@pytest.mark.parametrize("training", [False, True])
def test_crnn_oracle_matches_reference_fixture(training):
    """oracle/crnn_oracle.py (state-dict functional CRNN + parse_crnn_data) vs the fixture generated from the live
    `model/crnn/crnn.py:CRNN(32, 1, 37, 256)`: logits, sampled gradients and BN buffers bit-exact"""
    import os
    from oracle import crnn_oracle as co
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crnn_n3.pt"))
    sd = _crnn_sd(training)
    x = torch.rand(3, 3, 16, 64, generator=torch.Generator().manual_seed(1234))
    gray = co.parse_crnn_data(x)
    assert torch.equal(gray, fx["gray"])
    logits = co.crnn_forward(sd, gray, training)
    assert torch.equal(logits, fx["train_logits" if training else "eval_logits"])
    tp = co.text_prior(logits)
    assert tuple(tp.shape) == (3, 37, 1, 26) and torch.allclose(tp.sum(1), torch.ones(3, 1, 26), atol=1e-5)
    if training:
        gen = torch.Generator().manual_seed(99)
        (logits * torch.randn(logits.shape, generator=gen)).sum().backward()
        for n, ref in fx["train_grads"].items():
            assert tuple(sd[n].grad.shape) == tuple(ref["shape"]), n
            assert torch.equal(sd[n].grad.reshape(-1)[ref["idx"]], ref["val"]), n
            assert abs(sd[n].grad.double().sum().item() - ref["sum"]) <= 1e-6 * max(1.0, abs(ref["sum"])), n
        for n, b in fx["train_buffers"].items():
            assert torch.equal(sd[n], b), n


@pytest.mark.skipif(not rh.available(), reason="live reference not present (build container only)")
def test_crnn_state_dict_matches_live_reference():
    import importlib
    import warnings
    from oracle import crnn_oracle as co
    rh.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("model.crnn.crnn")
    torch.manual_seed(77)
    ref_sd = mod.CRNN(32, 1, 37, 256).state_dict()
    sd = co.make_state_dict(77)
    assert list(ref_sd.keys()) == list(sd.keys())
    for k in ref_sd:
        assert torch.equal(ref_sd[k], sd[k]), k


@pytest.mark.skipif(not rh.available(), reason="live reference not present (build container only)")
def test_loss_block_oracle_bit_exact_vs_live_reference():
    """SemanticLoss, TRI_SSIM and torch_rotate_img restatements (oracle/loss_oracle.py) vs the reference's own code.
    `torch_rotate_img` is a method of a class whose module cannot be imported here (ten absent packages), so
    its FunctionDef is compiled straight from the reference file."""
    import ast
    import importlib
    import os
    import warnings
    import torch.nn.functional as F
    from oracle import loss_oracle as lo
    rh.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sl = importlib.import_module("loss.semantic_loss")
        sp = importlib.import_module("utils.ssim_psnr")
        g = torch.Generator().manual_seed(3)
        p = torch.softmax(torch.randn(26, 4, 37, generator=g), -1).requires_grad_(True)
        q = torch.softmax(torch.randn(26, 4, 37, generator=g), -1)
        l_ref = sl.SemanticLoss()(p, q)
        l_ref.backward()
        g_ref = p.grad.clone()
        p.grad = None
        l_o = lo.semantic_loss(p, q)
        l_o.backward()
        assert torch.equal(l_ref, l_o) and torch.equal(g_ref, p.grad)
        a, b, c = [torch.rand(3, 4, 32, 128, generator=g) for _ in range(3)]
        assert torch.equal(sp.TRI_SSIM()(a, b, c), lo.tri_ssim(a, b, c))
        assert torch.equal(sp.TRI_SSIM(size_average=False)(a, b, c), lo.tri_ssim(a, b, c, size_average=False))
        src = open(os.path.join(rh.REF_ROOT, "interfaces", "super_resolution.py")).read()
        fn = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "torch_rotate_img"][0]
        ns = {"torch": torch, "F": F}
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "torch_rotate_img", "exec"), ns)
        arcs = (torch.rand(3, generator=g) - 0.5) * 0.2
        offs = torch.rand(3, generator=g)
        assert torch.equal(ns["torch_rotate_img"](None, a, arcs, offs), lo.rotate_img(a, arcs, offs))


def test_loss_block_oracle_matches_reference_fixture():
    import os
    from oracle import loss_oracle as lo
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_block_n3.pt"))
    assert torch.equal(lo.semantic_loss(fx["pred"], fx["gt"]), fx["semantic"])
    assert torch.equal(lo.tri_ssim(fx["a"], fx["b"], fx["c"]), fx["tri_ssim"])
    assert torch.equal(lo.tri_ssim(fx["a"], fx["b"], fx["c"], size_average=False), fx["tri_ssim_per_sample"])
    assert torch.equal(lo.rotate_img(fx["a"], fx["arcs"], fx["offs"]), fx["rotated"])
