"""Op-level parity of the recurrent positional encoding (quirk Q1: InfoTransformer.forward,
/root/reference/model/transformer_v2.py:177,201,215-221) at LONG recurrences: the BiGRU recurs over the batch axis, so
the benchmarked batch (N = 64) is a T = 64 recurrence and BASELINE configs[3] (128 / GPU) a T = 128 one, on hidden
state that travels as bf16 hi/lo planes.  Checked against torch's own CPU nn.GRU (what the reference dispatches to,
torch._VF.gru) in fp32 (forward) and float64 (gradients): rounding error must not build up over the chain.
Tolerances: query_pos <= 1e-3 of its max-abs (north_star bound); gradients rel-L2 <= 2e-3 vs the float64 torch GRU."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def Wd_ok(W, H):
    return W >= 32


def _mk(H, W, C=64, seed=7):
    torch.manual_seed(seed)
    I = H * C
    emb = torch.nn.Embedding(H * W, C)
    gru = torch.nn.GRU(I, I // 2, bidirectional=True, batch_first=True)
    for p in gru.parameters():                       # InfoTransformer._reset_parameters (transformer_v2.py:193-196)
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
    with torch.no_grad():                            # non-trivial biases / larger recurrent gain than the fresh init
        for n, p in gru.named_parameters():
            if "bias" in n:
                p.uniform_(-0.5, 0.5)
            if "weight_hh" in n:
                p.mul_(3.0)
    return emb, gru


def _torch_rpe(emb, gru, N, H, W, dt):
    C = emb.weight.shape[1]
    e = emb.weight.to(dt)
    g = torch.nn.GRU(gru.input_size, gru.hidden_size, bidirectional=True, batch_first=True).to(dt)
    g.load_state_dict({k: v.to(dt) for k, v in gru.state_dict().items()})
    e = e.detach().clone().requires_grad_(True)
    qe = e.unsqueeze(1).repeat(1, N, 1).reshape(H, W, N, C).permute(1, 2, 0, 3).reshape(W, N, H * C)
    out, _ = g(qe)
    q = out.reshape(W, N, H, C).permute(2, 0, 1, 3).reshape(H * W, N, C)      # [HW, N, C] as the reference holds it
    return q.permute(1, 0, 2), e, g                                           # -> [N, HW, C]


@pytest.mark.parametrize("N,H,W", [(64, 32, 128), (128, 16, 64), (64, 16, 64), (5, 8, 32)])
def test_rpe_long_recurrence_vs_torch_gru(N, H, W):
    from tatt_b200 import stages
    emb, gru = _mk(H, W)
    q32, _, _ = _torch_rpe(emb, gru, N, H, W, torch.float32)
    q64, e64, g64 = _torch_rpe(emb, gru, N, H, W, torch.float64)
    wgt = torch.randn(q64.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    (q64 * wgt).sum().backward()

    embd, grud = emb.to(DEV), gru.to(DEV)
    q = stages.rpe_stage(embd, grud, N, H, W)
    assert q.shape == (N, H * W, 64)
    if "fwd_sync" in stages.rpe_debug and Wd_ok(W, H):
        assert int(stages.rpe_debug["fwd_sync"][2 * N].item()) == 0, "persistent RPE kernel: a step barrier timed out"
    scale = q32.abs().max().item()
    err = (q.detach().cpu() - q32.detach()).abs().max().item() / scale
    # error of the LAST step separately: rounding must not accumulate along the chain
    err_last = (q[-1].detach().cpu() - q32[-1].detach()).abs().max().item() / scale
    print("RPE N=%d H=%d W=%d: fwd err %.2e (last step %.2e)" % (N, H, W, err, err_last))
    assert err <= 1e-3 and err_last <= 1e-3
    (q * wgt.float().to(DEV)).sum().backward()
    refs = {"emb": (embd.weight.grad, e64.grad)}
    for n, p in grud.named_parameters():
        refs[n] = (p.grad, dict(g64.named_parameters())[n].grad)
    num = den = 0.0
    for n, (a, b) in refs.items():
        d = a.detach().double().cpu() - b
        rel = d.norm().item() / max(b.norm().item(), 1e-30)
        assert rel <= 2e-3, "grad %s rel-L2 %.3e" % (n, rel)
        num += d.pow(2).sum().item()
        den += b.pow(2).sum().item()
    print("RPE N=%d: whole-gradient rel-L2 vs float64 torch GRU %.2e" % (N, (num / den) ** 0.5))
