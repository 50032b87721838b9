"""End-to-end parity of the drop-in modules (CUDA, through the C-ABI) against the oracle restatement and
the committed reference fixtures.  Stated tolerances (fp32): outputs / intermediates <= 1e-3 relative to
the tensor's max-abs (north_star bound) against the fp32 oracle.  Parameter gradients: <= 2e-3 relative to
the gradient's max-abs (plus an absolute floor of 1e-5 x the largest gradient in the model, for the conv biases
in front of a train-mode BatchNorm whose true gradient is exactly 0) against the oracle evaluated in FLOAT64.
Float64 because the reference's own fp32 backward is noisy on this path: with STN on, its fp32 gradients deviate
from its fp64 gradients by up to 1.2e-2 (block1.0.bias; measured), so fp32-vs-fp32 would compare
two rounding noises; per parameter: rel-L2 <= 1e-2 + 5x the reference's own fp32 rel-L2 deviation (the tcgen05 path computes
every GEMM as a bf16 hi/lo split with 3 MMAs: measured 4.5e-6 rel-L2 per GEMM vs fp64, ~10-30x fp32 rounding --
tools/probe_tc_precision.py -- which ReLU kinks amplify on a few small tensors to ~4e-3), max-abs <= 1e-2 of the
gradient's max-abs + 5x the reference's max deviation; whole gradient vector: rel-L2 <= 2e-3 + 5x reference noise (ReLU / max-pool flips
and the ill-conditioned TPS solve make the fp32 backward itself non-smooth).  The committed fixtures carry the
reference's float64 gradients and its fp32 noise, and are checked with the same bound."""
import re

import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def make(case, zero_drop=True):
    import tatt_b200
    from oracle import ref_harness as rh
    from oracle import tatt_oracle as orc
    cls, kw, N, training = gu.CASES[case]
    torch.manual_seed(gu.SEED)
    net = getattr(tatt_b200, cls)(**kw)
    if zero_drop:
        rh.zero_dropout(net)
    rh.perturb_(net)
    net.train(training)
    sd = orc.clone_sd(net.state_dict(), requires_grad=training)
    x, tp = orc.synthetic_inputs(N, kw["height"] // 2, kw["width"] // 2, seed=gu.SEED, with_mask=kw.get("mask", True))
    return net.to(DEV), sd, x, tp, cls, kw, N, training


def to_dtype(sd, dt):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().to(dt) if v.is_floating_point() else v.detach().clone()
        if v.requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out


def run_oracle(cls, kw, sd, x, tp, training):
    from oracle import tatt_oracle as orc
    dt = next(v.dtype for v in sd.values() if v.is_floating_point())
    x, tp = x.to(dt), tp.to(dt)
    if cls == "TSRN":
        out, block = orc.tsrn_forward(sd, x, training=training, stn=kw["STN"])
        return out, None, block
    return orc.tsrn_tl_trans_forward(sd, x, tp, training=training, stn=kw["STN"], dropout_p=0.0)


@pytest.mark.parametrize("case", list(gu.CASES))
def test_forward_backward_vs_oracle_and_golden(case):
    net, sd, x, tp, cls, kw, N, training = make(case)
    fx = gu.load(case)
    xo = x.to(DEV)
    if cls == "TSRN":
        out, aux = net(xo), None
    else:
        out, aux = net(xo, tp.to(DEV))
    sd64 = to_dtype(sd, torch.float64)
    o_out, o_aux, o_block = run_oracle(cls, kw, sd, x, tp, training)
    assert out.shape == o_out.shape and out.dtype == torch.float32
    assert relerr(out, o_out) <= 1e-3, "output vs oracle: %.3e" % relerr(out, o_out)
    gu.check_summary("out", out, fx["out"], 1e-3)
    for k in o_block:
        e = relerr(net.block[k], o_block[k])
        assert e <= 1e-3, "block %s vs oracle: %.3e" % (k, e)
        if "block" + k in fx:
            gu.check_summary("block" + k, net.block[k], fx["block" + k], 1e-3)
    if aux is not None:
        pw = aux["pr_weights"] if training else aux
        opw = o_aux["pr_weights"] if training else o_aux
        assert relerr(pw, opw) <= 1e-3
        assert (pw.sum(-1) - 1).abs().max().item() < 1e-4          # rows of the averaged attention sum to 1
        gu.check_summary("pr_weights", pw, fx["pr_weights"], 1e-3)
        if training:
            assert set(aux.keys()) == {"pr_weights", "pr_weights_gt", "spatial_t_emb", "spatial_t_emb_gt",
                                       "in_feat", "trans_feat"}
            assert relerr(aux["spatial_t_emb"], o_aux["spatial_t_emb"]) <= 1e-3
            gu.check_summary("tp_map", aux["spatial_t_emb"], fx["tp_map"], 1e-3)
    if not training:
        return
    gen = torch.Generator().manual_seed(99)
    wgt = torch.randn(out.shape, generator=gen)
    (out * wgt.to(DEV)).sum().backward()
    (o_out * wgt).sum().backward()
    o64 = run_oracle(cls, kw, sd64, x, tp, training)[0]
    (o64 * wgt.double()).sum().backward()
    G = max(v.grad.abs().max().item() for v in sd64.values() if v.requires_grad and v.grad is not None)
    worst = ("", 0.0)
    num2 = den2 = nnum2 = 0.0
    for n, p in net.named_parameters():
        og = sd64[n].grad
        ref = fx["grads"][n]
        if ref is None:                      # Q3 dead parameters never receive a gradient
            assert p.grad is None, n
            assert og is None or og.abs().max().item() == 0
            continue
        assert p.grad is not None, n
        d = p.grad.detach().double().cpu() - og
        nd = sd[n].grad.double() - og                                  # the fp32 reference's own deviation
        err, noise = d.abs().max().item(), nd.abs().max().item()
        bound = 1e-2 * og.abs().max().item() + 5.0 * noise + 1e-5 * G
        l2ref = max(og.norm().item(), 1e-6 * G * og.numel() ** 0.5)
        rel, nrel = d.norm().item() / l2ref, nd.norm().item() / l2ref
        if err / bound > worst[1]:
            worst = (n, err / bound)
        assert err <= bound, "grad %s vs fp64 oracle: max err %.3e > bound %.3e (max %.3e)" % (
            n, err, bound, og.abs().max().item())
        assert rel <= 1e-2 + 5.0 * nrel, "grad %s vs fp64 oracle: rel-L2 %.3e (reference fp32 noise %.3e)" % (
            n, rel, nrel)
        num2 += d.pow(2).sum().item()
        den2 += og.pow(2).sum().item()
        nnum2 += nd.pow(2).sum().item()
        gu.check_summary("grad " + n, p.grad, fx["grads64"][n], 1e-2, atol=5.0 * fx["noise"][n] + 1e-5 * fx["gmax"])
    rel_l2, noise_l2 = (num2 / den2) ** 0.5, (nnum2 / den2) ** 0.5
    print("worst grad (err/bound)", worst, "whole-gradient rel-L2 vs fp64 oracle %.3e (fp32 reference: %.3e)" % (
        rel_l2, noise_l2))
    assert rel_l2 <= 2e-3 + 5.0 * noise_l2, "whole-model gradient rel-L2 error %.3e" % rel_l2
    # BatchNorm running statistics were updated exactly like torch does
    for n, b in net.named_buffers():
        if "running" in n or "num_batches" in n:
            assert relerr(b.float(), sd[n].float()) <= 1e-4, n
            gu.check_summary("buffer " + n, b.float(), fx["buffers"][n], 1e-4)


def make_big(case):
    import tatt_b200
    from oracle import ref_harness as rh
    from oracle import tatt_oracle as orc
    cls, kw, N, training = gu.BIG_CASES[case]
    torch.manual_seed(gu.SEED)
    net = getattr(tatt_b200, cls)(**kw)
    rh.zero_dropout(net)
    rh.perturb_(net)
    net.train(training)
    x, tp = orc.synthetic_inputs(N, kw["height"] // 2, kw["width"] // 2, seed=gu.SEED, with_mask=kw.get("mask", True))
    return net.to(DEV), x, tp


def test_benchmarked_config_vs_golden():
    """BASELINE configs[1] itself: G32, N = 64, train mode (dropout 0), forward + backward -- the batch bench.py
    times, which selects kernel variants the small cases never reach (T = 64 batch-axis recurrence on bf16 hi/lo
    plane hidden state, one-wave split-K GEMMs, multi-tile rolling conv, > 296 row tiles).  Checked against the
    fixture generated from the LIVE reference (tests/golden/make_golden.py tatt_g32_train_n64; the oracle itself is
    too slow to re-run here).  The fixture holds the reference's fp32 gradients (its float64 pass does not fit the
    build container); at G32 without STN the reference's own fp32-vs-fp64 gradient noise is 3e-5 rel-L2.
    Bounds: outputs / intermediates 1e-3 of max-abs; BN buffers 1e-4; per-parameter gradient norm and 64 strided
    samples 1e-2 (+1e-5 of the model's largest gradient)."""
    net, x, tp = make_big("tatt_g32_train_n64")
    fx = gu.load("tatt_g32_train_n64")
    out, aux = net(x.to(DEV), tp.to(DEV))
    gu.check_summary("out", out, fx["out"], 1e-3)
    for k in ("1", "4", "7"):
        gu.check_summary("block" + k, net.block[k], fx["block" + k], 1e-3)
    gu.check_summary("pr_weights", aux["pr_weights"], fx["pr_weights"], 1e-3)
    gu.check_summary("tp_map", aux["spatial_t_emb"], fx["tp_map"], 1e-3)
    errs = {}
    for k, t in (("out", out), ("tp_map", aux["spatial_t_emb"]), ("pr_weights", aux["pr_weights"]),
                 ("block7", net.block["7"])):
        r = fx[k]
        got = t.detach().double().cpu().reshape(-1)[r["idx"].long()].float()
        errs[k] = ((got - r["val"]).abs().max() / r["val"].abs().max()).item()
    print("N=64 G32 forward errors vs the live-reference fixture (max-abs / max-abs):", errs)
    gen = torch.Generator().manual_seed(99)
    wgt = torch.randn(out.shape, generator=gen)
    (out * wgt.to(DEV)).sum().backward()
    worst = ("", 0.0)
    for n, p in net.named_parameters():
        ref = fx["grads"][n]
        if ref is None:
            assert p.grad is None, n
            continue
        assert p.grad is not None, n
        if re.fullmatch(r"block\d\.conv[12]\.bias|block7\.0\.bias", n):
            # conv bias in front of a train-mode BatchNorm: the true gradient is exactly 0, both sides hold rounding
            # noise only (the reference's own fp32 value is ~3e-5 of the model's largest gradient here)
            assert p.grad.abs().max().item() <= 1e-4 * fx["gmax"] and ref["l2"] <= 1e-4 * fx["gmax"] * ref["n"] ** 0.5, n
            continue
        gu.check_summary("grad " + n, p.grad, ref, 1e-2, atol=1e-5 * fx["gmax"])
        rel = abs(p.grad.double().norm().item() - ref["l2"]) / max(ref["l2"], 1e-5 * fx["gmax"])
        if rel > worst[1]:
            worst = (n, rel)
    print("N=64 G32 worst per-parameter gradient-norm deviation vs the reference:", worst)
    for n, b in net.named_buffers():
        if "running" in n or "num_batches" in n:
            gu.check_summary("buffer " + n, b.float(), fx["buffers"][n], 1e-4)


def test_eval_repeatable_and_qpos_cache():
    net, sd, x, tp, cls, kw, N, training = make("tatt_g16_eval_n2")
    with torch.no_grad():
        a, wa = net(x.to(DEV), tp.to(DEV))
        cache = net.infoGen._qpos_cache
        b, wb = net(x.to(DEV), tp.to(DEV))
    assert torch.equal(a, b) and torch.equal(wa, wb)
    assert net.infoGen._qpos_cache is cache                      # RPE reused: weights unchanged
    with torch.no_grad():
        net.infoGen.init_factor.weight.mul_(1.5)
        c, _ = net(x.to(DEV), tp.to(DEV))
    assert net.infoGen._qpos_cache is not cache and not torch.equal(a, c)


def test_eval_after_trainer_step_sees_updated_weights():
    """The fused clip+Adam kernel updates parameters through raw pointers (no torch version bump): the cached eval
    positional encoding must still be refreshed -- eval after Trainer.step == a fresh model loaded from state_dict."""
    import tatt_b200
    from tatt_b200.train import GraphedForward, Trainer
    net, sd, x, tp, cls, kw, N, training = make("tatt_g16_stn_train_n3")
    xd, td = x.to(DEV), tp.to(DEV)
    net.eval()
    with torch.no_grad():
        o0, _ = net(xd, td)                                       # fills the qpos cache with the initial weights
    gf = GraphedForward(net, tuple(x.shape), tuple(tp.shape))
    gf.x.copy_(xd); gf.text.copy_(td)
    gf.capture()
    net.train()
    # the reference's learning rate: with a huge step the scrambled network amplifies the fp32-atomics ordering noise of
    # the split-K kernels (~1e-6) to ~1e-3, which is not what this test is about; a stale positional encoding or stale
    # weights would still show up at ~1e-3 against the 1e-5 bound
    tr = Trainer(net, lr=1e-3)
    g = torch.randn(N, 4, 32, 128, generator=torch.Generator().manual_seed(3)).to(DEV)
    tr.step(xd, td, g)
    net.eval()
    with torch.no_grad():
        o1, w1 = net(xd, td)
    o1g, w1g = gf()
    fresh = getattr(tatt_b200, cls)(**kw).to(DEV).eval()
    fresh.load_state_dict(net.state_dict())
    with torch.no_grad():
        o2, w2 = fresh(xd, td)
    assert not torch.equal(o0, o1)
    assert relerr(o0, o1) > 1e-4                                  # the step did change the function
    assert relerr(o1, o2) <= 1e-5 and relerr(w1, w2) <= 1e-5
    assert relerr(o1g, o2) <= 1e-5 and relerr(w1g, w2) <= 1e-5    # graph replay reads the refreshed encoding / weights
    # optimizer state round trip
    st = tr.state_dict()
    assert st["step"] == 1 and len(st["exp_avg"]) == len(tr.bucket.params)
    tr.load_state_dict(st)


def test_batch_position_dependence_q1():
    """Quirk Q1: the same sample replicated in a batch gets different outputs (batch-axis recurrence)."""
    net, sd, x, tp, *_ = make("tatt_g16_eval_n2")
    xx, tt = x[:1].repeat(3, 1, 1, 1).to(DEV), tp[:1].repeat(3, 1, 1, 1).to(DEV)
    with torch.no_grad():
        out, _ = net(xx, tt)
    assert (out[0] - out[1]).abs().max().item() > 0


def test_default_text_prior_only_for_single_image():
    net, sd, x, tp, *_ = make("tatt_g16_eval_n2")
    from oracle import tatt_oracle as orc
    with torch.no_grad():
        out, pw = net(x[:1].to(DEV))                              # ptflops-style probe: no text prior
    o_out, o_pw, _ = orc.tsrn_tl_trans_forward(sd, x[:1], None, training=False)
    assert relerr(out, o_out) <= 1e-3
    with pytest.raises(RuntimeError):
        net(x.to(DEV))                                            # reference also fails for N > 1


def test_train_mode_with_dropout_runs_and_is_seeded():
    import tatt_b200
    tatt_b200.manual_seed(5)
    net, sd, x, tp, *_ = make("tatt_g16_stn_train_n3", zero_drop=False)
    o1, _ = net(x.to(DEV), tp.to(DEV))
    o1.mean().backward()
    g1 = net.block2.conv1.weight.grad.clone()
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all()
    o2, _ = net(x.to(DEV), tp.to(DEV))                            # counter advanced -> different masks
    assert not torch.equal(o1, o2)
    # BN running stats moved, so check the dropout stream restarts identically on a fresh copy
    tatt_b200.manual_seed(5)
    net2, *_ = make("tatt_g16_stn_train_n3", zero_drop=False)
    o3, _ = net2(x.to(DEV), tp.to(DEV))
    assert (o1 - o3).abs().max().item() <= 1e-4          # same masks (split-K atomics reorder fp32 sums)


def test_cpu_tensor_raises_no_fallback():
    import tatt_b200
    net = tatt_b200.TSRN_TL_TRANS(width=128, height=32)
    with pytest.raises(RuntimeError):
        net.eval()(torch.rand(1, 4, 16, 64))


def test_stn_at_g32_fails_like_reference():
    import tatt_b200
    net = tatt_b200.TSRN_TL_TRANS(width=256, height=64, STN=True).to(DEV).train()
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        net(torch.rand(2, 4, 32, 128, device=DEV), torch.rand(2, 37, 1, 26, device=DEV))


def test_graphed_trainer_matches_eager_trainer():
    """CUDA-graph replay of forward+backward+pack reproduces the eager launches' gradient (same weights, dropout
    off), and the replayed optimizer graph advances the device-side step counter."""
    from tatt_b200.train import GraphedTrainer, Trainer
    net_g, sd, x, tp, cls, kw, N, training = make("tatt_g16_stn_train_n3")
    net_e, *_ = make("tatt_g16_stn_train_n3")
    g = torch.randn(N, 4, 32, 128, generator=torch.Generator().manual_seed(3)).to(DEV) * 1e-3
    xe, te = x.to(DEV), tp.to(DEV)
    graphed = GraphedTrainer(net_g, tuple(x.shape), tuple(tp.shape), tuple(g.shape))
    graphed.x.copy_(xe); graphed.text.copy_(te); graphed.grad_out.copy_(g)
    sd_before = {k: v.clone() for k, v in net_g.state_dict().items()}
    graphed.capture(warmup=2)
    # capture is side-effect free: warm-up steps are rolled back (weights, BN buffers, Adam state, step counter)
    assert int(graphed.step_state[1].item()) == 0
    for k, v in net_g.state_dict().items():
        assert torch.equal(v, sd_before[k]), k
    assert float(graphed.m.abs().max()) == 0.0 and float(graphed.v.abs().max()) == 0.0
    graphed.step(x.pin_memory(), tp.pin_memory())
    torch.cuda.synchronize()
    assert int(graphed.step_state[1].item()) == 1
    # same weights + BN buffers in an eager model -> gradient of the next step must match the graph replay
    net_e.load_state_dict(net_g.state_dict())
    eager = Trainer(net_e)
    eager.forward_backward(xe, te, g)
    ge = eager._ensure_bucket().pack().clone()
    graphed.graph_fb.replay()
    torch.cuda.synchronize()
    gg = graphed.bucket.flat_grad
    assert ge.shape == gg.shape
    rel = ((ge - gg).norm() / ge.norm()).item()
    assert rel <= 1e-3, "graph replay vs eager gradient rel-L2 %.3e" % rel


def test_config1_single_crop_psnr_vs_reference_math():
    """BASELINE configs[0]: one 32x128 -> 64x256 crop, eval forward; PSNR (utils/ssim_psnr.py:9-15 formula) of our SR
    output against the reference's CPU output.  Identical images give +inf; fp32-level agreement is > 100 dB."""
    import tatt_b200
    from oracle import ref_harness as rh
    from oracle import tatt_oracle as orc
    torch.manual_seed(gu.SEED)
    net = tatt_b200.TSRN_TL_TRANS(scale_factor=2, width=256, height=64, STN=False, mask=True)
    rh.perturb_(net)
    net.eval()
    sd = orc.clone_sd(net.state_dict())
    x, tp = orc.synthetic_inputs(1, 32, 128, seed=gu.SEED)
    ref, ref_w, _ = orc.tsrn_tl_trans_forward(sd, x, tp, training=False)
    with torch.no_grad():
        out, w = net.to(DEV)(x.to(DEV), tp.to(DEV))
    assert out.shape == (1, 4, 64, 256) and w.shape == (1, 32 * 128, 26)
    psnr = orc.psnr(out.cpu(), ref)
    print("config-1 PSNR new-vs-reference: %.1f dB" % psnr)
    assert psnr > 100.0


def test_config5_two_geometry_buckets_in_one_process():
    """BASELINE configs[4] (the reference has no mixed-width batching): one G16 bucket with STN/TPS on and
    one G32 bucket with STN off, each trained for a step in the same process; both must match the oracle."""
    for case in ("tatt_g16_stn_train_n3", "tatt_g32_train_n2"):
        net, sd, x, tp, cls, kw, N, training = make(case)
        out, aux = net(x.to(DEV), tp.to(DEV))
        o_out, _, _ = run_oracle(cls, kw, sd, x, tp, training)
        assert relerr(out, o_out) <= 1e-3
        out.mean().backward()
        assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


# bf16 mode (BASELINE.json configs 3/4: the reference under autocast).  Operands of every GEMM / convolution are
# rounded to bf16 (unit roundoff 2^-9 = 2e-3) and accumulated in fp32; everything else stays fp32.  The stated
# tolerance for this mode: outputs within 5e-2 of the fp32 oracle (max-abs / max; measured 1.9e-2 / 2.3e-2), whole-gradient rel-L2 within 1e-1
# (measured 1.9e-2 / 3.7e-2)
# of the fp64 oracle.  (The default mode's 1e-3 bars are unaffected.)
BF16_OUT_TOL, BF16_GRAD_TOL = 5e-2, 1e-1


@pytest.mark.parametrize("case", ["tatt_g32_train_n2", "tatt_g16_stn_train_n3"])
def test_bf16_mode_within_stated_tolerance(case):
    import tatt_b200
    net, sd, x, tp, cls, kw, N, training = make(case)
    sd64 = to_dtype(sd, torch.float64)
    tatt_b200.set_precision("bf16")
    try:
        out, aux = net(x.to(DEV), tp.to(DEV))
        gen = torch.Generator().manual_seed(99)
        wgt = torch.randn(out.shape, generator=gen)
        (out * wgt.to(DEV)).sum().backward()
        torch.cuda.synchronize()
    finally:
        tatt_b200.set_precision("fp32")
    o_out, o_aux, _ = run_oracle(cls, kw, sd, x, tp, training)
    e_out = relerr(out, o_out)
    e_pw = relerr(aux["pr_weights"], o_aux["pr_weights"])
    o64 = run_oracle(cls, kw, sd64, x, tp, training)[0]
    (o64 * wgt.double()).sum().backward()
    num2 = den2 = 0.0
    for n, p in net.named_parameters():
        og = sd64[n].grad
        if og is None or p.grad is None:
            continue
        num2 += (p.grad.detach().double().cpu() - og).pow(2).sum().item()
        den2 += og.pow(2).sum().item()
    rel = (num2 / den2) ** 0.5
    print("bf16 mode %s: out %.3e, pr_weights %.3e, whole-gradient rel-L2 %.3e" % (case, e_out, e_pw, rel))
    assert e_out <= BF16_OUT_TOL and e_pw <= BF16_OUT_TOL
    assert rel <= BF16_GRAD_TOL
    assert e_out > 1e-6            # the mode really changed the arithmetic
