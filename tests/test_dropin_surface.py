"""CPU checks of the drop-in nn.Module surface: constructor / forward signatures, state_dict
layout (against a manifest generated from the reference, and against the live reference when present),
bit-identical fresh init, strict load both ways, and that there is no CPU compute path."""
import inspect
import json
import os

import pytest
import torch

import golden_util as gu

MANIFEST = os.path.join(gu.GOLDEN_DIR, "state_dict_manifest.json")
CONFIGS = {
    "TSRN_TL_TRANS/g16_stn": ("TSRN_TL_TRANS", dict(scale_factor=2, width=128, height=32, STN=True, mask=True)),
    "TSRN_TL_TRANS/g32": ("TSRN_TL_TRANS", dict(scale_factor=2, width=256, height=64, STN=False, mask=True)),
    "TSRN_TL_TRANS/rgb": ("TSRN_TL_TRANS", dict(scale_factor=2, width=128, height=32, STN=False, mask=False)),
    "TSRN/g16_stn": ("TSRN", dict(scale_factor=2, width=128, height=32, STN=True, mask=True)),
    "TSRN/srb3": ("TSRN", dict(scale_factor=2, width=128, height=32, STN=False, mask=True, srb_nums=3)),
}


def _ref():
    from oracle import ref_harness as rh
    return rh.load() if rh.available() else None


def test_state_dict_manifest():
    import tatt_b200
    man = json.load(open(MANIFEST))
    for key, (cls, kw) in CONFIGS.items():
        sd = getattr(tatt_b200, cls)(**kw).state_dict()
        got = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
        assert got == man[key], key
    assert len(man["TSRN_TL_TRANS/g16_stn"]) == 304


def test_signatures():
    import tatt_b200
    sig = inspect.signature(tatt_b200.TSRN_TL_TRANS.__init__)
    assert list(sig.parameters)[1:] == ["scale_factor", "width", "height", "STN", "srb_nums", "mask", "hidden_units",
                                        "word_vec_d", "text_emb", "out_text_channels", "feature_rotate",
                                        "rotate_train"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [2, 128, 32, False, 5, True, 32, 300, 37, 64,
                                                                      False, 3.]
    assert list(inspect.signature(tatt_b200.TSRN_TL_TRANS.forward).parameters) == [
        "self", "x", "text_emb", "text_emb_gt", "feature_arcs", "rand_offs"]
    sig = inspect.signature(tatt_b200.TSRN.__init__)
    assert list(sig.parameters)[1:] == ["scale_factor", "width", "height", "STN", "srb_nums", "mask", "hidden_units"]
    assert list(inspect.signature(tatt_b200.TSRN.forward).parameters) == ["self", "x"]
    ref = _ref()
    if ref is not None:
        for cls in ("TSRN", "TSRN_TL_TRANS"):
            for fn in ("__init__", "forward"):
                a = inspect.signature(getattr(getattr(ref, cls), fn))
                b = inspect.signature(getattr(getattr(tatt_b200, cls), fn))
                assert [(p.name, p.default) for p in a.parameters.values()] == \
                       [(p.name, p.default) for p in b.parameters.values()], (cls, fn)


@pytest.mark.skipif(_ref() is None, reason="live reference not present (build container only)")
def test_fresh_init_bit_identical_and_strict_load_both_ways():
    import tatt_b200
    ref = _ref()
    for key, (cls, kw) in CONFIGS.items():
        torch.manual_seed(1234); a = getattr(ref, cls)(**kw)
        torch.manual_seed(1234); b = getattr(tatt_b200, cls)(**kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            assert torch.equal(sa[k], sb[k]), (key, k)
        assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
        b.load_state_dict(sa, strict=True); a.load_state_dict(sb, strict=True)


def test_dead_q3_parameters_exist_and_module_api_works():
    import tatt_b200
    net = tatt_b200.TSRN_TL_TRANS(STN=True)
    names = dict(net.named_parameters())
    for n in ("infoGen.fc_feature_in.weight", "infoGen.transformer.decoder.layers.0.self_attn.in_proj_weight",
              "infoGen.transformer.decoder.layers.1.norm1.weight"):
        assert n in names
    assert sum(p.numel() for p in net.parameters()) == 7608334
    assert sum(p.numel() for p in tatt_b200.TSRN(STN=True).parameters()) == 2681677
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.5, 0.999))   # base.py:557-558
    net.train(); net.eval()
    for p in net.parameters():
        p.requires_grad = False
    assert opt is not None


def test_no_cpu_fallback():
    import tatt_b200
    net = tatt_b200.TSRN().eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 4, 16, 64))
    with pytest.raises(RuntimeError, match="parameter container"):
        net.block2.gru1(torch.rand(1, 64, 4, 4))


def test_image_loss_surface_cpu():
    """`tatt_b200.losses.ImageLoss` mirrors `loss/image_loss.py:ImageLoss`: constructor, broken gradient=False path
    (UnboundLocalError in the reference, image_loss.py:32), and no CPU fallback."""
    import inspect

    import pytest
    import torch
    from tatt_b200.losses import ImageLoss
    sig = inspect.signature(ImageLoss.__init__)
    assert list(sig.parameters)[1:] == ["gradient", "loss_weight"]
    assert sig.parameters["gradient"].default is True and list(sig.parameters["loss_weight"].default) == [20, 1e-4]
    assert list(inspect.signature(ImageLoss.forward).parameters)[1:] == ["out_images", "target_images", "grad_mask"]
    x = torch.zeros(2, 4, 8, 8)
    with pytest.raises(UnboundLocalError):
        ImageLoss(gradient=False)(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ImageLoss()(x, x)


def test_crnn_surface_and_fresh_init():
    """`tatt_b200.crnn.CRNN` mirrors `model/crnn/crnn.py:CRNN`: constructor, state_dict keys / shapes and bit-identical
    fresh initialisation (same torch.nn leaves in the same order) -- against the reference-free factory of the oracle and,
    in the build container, against the live class; leaves are containers; CPU tensors raise (no fallback)."""
    import pytest
    import torch
    from oracle import crnn_oracle as co
    from oracle import ref_harness as rh
    from tatt_b200.crnn import CRNN
    torch.manual_seed(77)
    net = CRNN(32, 1, 37, 256)
    sd = net.state_dict()
    want = co.make_state_dict(77)
    assert list(sd.keys()) == list(want.keys())
    for k in want:
        assert torch.equal(sd[k], want[k]), k
    if rh.available():
        import importlib
        import warnings
        rh.load()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("model.crnn.crnn")
        torch.manual_seed(77)
        ref = mod.CRNN(32, 1, 37, 256)
        rsd = ref.state_dict()
        assert list(rsd.keys()) == list(sd.keys())
        assert all(torch.equal(rsd[k], sd[k]) for k in sd)
        ref.load_state_dict(sd, strict=True)
        net.load_state_dict(rsd, strict=True)
        assert [n for n, _ in ref.named_modules()] == [n for n, _ in net.named_modules()]
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 1, 32, 100))
    with pytest.raises(RuntimeError):
        net.rnn[0](torch.zeros(26, 2, 512))
    with pytest.raises(NotImplementedError):
        CRNN(32, 1, 37, 256, leakyRelu=True)
