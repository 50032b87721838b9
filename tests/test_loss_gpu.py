"""Parity of the CUDA image loss (tatt_b200.losses.ImageLoss -> csrc/loss.cu) against the fixture generated from the
reference's `loss/image_loss.py:ImageLoss` and against the CPU oracle.  Tolerances: loss 1e-5 relative; gradient 1e-4 of
its max-abs against the fp64 reference gradient (the reference's own fp32 gradient deviates by the same order near the
|.| kink of the gradient-prior term)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_image_loss_matches_reference_fixture():
    from tatt_b200.losses import ImageLoss
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_loss_n3.pt"))
    dev = torch.device("cuda:0")
    out = fx["out"].to(dev).requires_grad_(True)
    crit = ImageLoss(gradient=True, loss_weight=[20, 1e-4])
    loss = crit(out, fx["target"].to(dev))
    assert loss.shape == fx["loss"].shape
    rel = ((loss.detach().cpu() - fx["loss"]).abs() / fx["loss"].abs()).max().item()
    assert rel <= 1e-5, "image loss: rel err %.3e" % rel
    (loss.mean() * 100).backward()
    ref = fx["dout64"].float()
    err = (out.grad.cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-4, "image loss gradient: rel-max err %.3e" % err
    with pytest.raises(UnboundLocalError):                       # the reference's gradient=False path is broken
        ImageLoss(gradient=False)(out, fx["target"].to(dev))


def test_image_loss_vs_oracle_large_and_odd():
    from oracle import loss_oracle as lo
    from tatt_b200.losses import ImageLoss
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    for shape in ((5, 4, 64, 256), (2, 3, 7, 13)):
        out = torch.tanh(torch.randn(*shape, generator=g))
        tgt = torch.rand(*shape, generator=g)
        o64 = out.double().requires_grad_(True)
        l64 = lo.image_loss(o64, tgt.double())
        w = torch.rand(shape[0], generator=g).double()
        (l64 * w).sum().backward()
        od = out.to(dev).requires_grad_(True)
        ld = ImageLoss()(od, tgt.to(dev))
        (ld * w.float().to(dev)).sum().backward()
        assert ((ld.detach().cpu().double() - l64.detach()).abs() / l64.detach().abs()).max().item() <= 1e-5
        ref = o64.grad
        assert (od.grad.cpu().double() - ref).abs().max().item() / ref.abs().max().item() <= 1e-4
    with pytest.raises(RuntimeError):
        ImageLoss()(torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4))      # CPU tensors: no fallback path
