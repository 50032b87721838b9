"""Times the recurrent positional encoding (forward, and forward+backward) alone at the bench shape (G32, N=64):
CUDA events around 20 calls after warm-up.  TATT_RPE_PERSIST=0 selects the round-1 per-step path for comparison."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tatt_b200 import stages

N, H, W = int(os.environ.get("N", 64)), 32, 128
torch.manual_seed(0)
emb = torch.nn.Embedding(H * W, 64).cuda()
gru = torch.nn.GRU(H * 64, H * 32, bidirectional=True, batch_first=True).cuda()


def run(bwd):
    q = stages.rpe_stage(emb, gru, N, H, W)
    if bwd:
        q.backward(torch.ones_like(q))


for bwd in (False, True):
    for _ in range(3):
        run(bwd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run(bwd)
    e1.record()
    torch.cuda.synchronize()
    print("RPE N=%d %s: %.3f ms per call" % (N, "fwd+bwd" if bwd else "fwd (train mode, gates saved)", e0.elapsed_time(e1) / 20))
with torch.no_grad():
    for _ in range(3):
        stages.rpe_stage(emb, gru, N, H, W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        stages.rpe_stage(emb, gru, N, H, W)
    e1.record()
    torch.cuda.synchronize()
    print("RPE N=%d fwd (no_grad): %.3f ms per call" % (N, e0.elapsed_time(e1) / 20))
