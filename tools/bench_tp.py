"""BASELINE.json configs[2]: TP Interpreter alone (text-prior -> image-feature cross-attention with the recurrent
positional encoding), batch 256, bf16 operands (fp32 accumulate), 26-character prior (L=26, C=37), one B200.

  python tools/bench_tp.py [--geometry g16|g32] [--batch 256] [--dtype bf16|f32] [--steps 10] [--once]

Prints one JSON line: images/s of `infoGen(block1_feature, text_prior)` forward and forward+backward, CUDA events,
inputs resident in HBM (each step streams > 126 MB of activations, no explicit L2 flush).  `--once` runs a single
forward+backward (for `ncu --set full -k regex:"mha|rows_gemm"`; numbers printed under a profiler are not bench values)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tatt_b200
from tatt_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--geometry", default="g16", choices=["g16", "g32"])
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--once", action="store_true")
ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
a = ap.parse_args()

dev = torch.device("cuda:0")
h, w = (16, 64) if a.geometry == "g16" else (32, 128)
torch.manual_seed(1234)
net = tatt_b200.TSRN_TL_TRANS(scale_factor=2, width=2 * w, height=2 * h, STN=False, mask=True).to(dev).train()
tatt_b200.manual_seed(1234)
tatt_b200.set_precision("bf16" if a.dtype == "bf16" else "fp32")
ig = net.infoGen
g = torch.Generator().manual_seed(1234)
feat = torch.randn(a.batch, 64, h, w, generator=g).to(dev).requires_grad_(True)
tp = torch.softmax(3 * torch.randn(a.batch, 37, 1, 26, generator=g), dim=1).to(dev)


def fwd():
    return ig(feat, tp)


def fwd_bwd():
    for p in ig.parameters():
        p.grad = None
    feat.grad = None
    tp_map, _ = ig(feat, tp)
    tp_map.backward(go)


go = torch.randn(a.batch, 64, h, w, device=dev) / (a.batch * 64 * h * w)
if a.once:
    fwd_bwd()
    torch.cuda.synchronize()
    sys.exit(0)


def timed(fn, n):
    """CUDA-graph replay by default: the eager path is bound by Python launch overhead (~2200 C-ABI calls per
    forward+backward at N = 256), not by the GPU.  Nothing runs on the legacy default stream before the capture
    (autograd's AccumulateGrad nodes remember the stream they were created on)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()                                  # first call also fills the positional-encoding cache in eval mode
        l0 = _cabi.launch_count
        fn()
        per = _cabi.launch_count - l0
        if a.eager:
            for _ in range(2):
                fn()
            side.synchronize()
            return _events(fn, n), per
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    return _events(gr.replay, n), per


def _events(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ig.eval()
    ms_f, lf = timed(fwd, a.steps)       # eval forward: the positional encoding is cached (it depends on weights + N only)
    ig.train()
ms_fb, lfb = timed(fwd_bwd, a.steps)
print(json.dumps({"metric": "TP Interpreter images/sec", "config": {"workload": "TPInterpreter(37->64, %dx%d queries, 26 keys) "
                  "batch %d" % (h, w, a.batch), "launch": "eager (python-launched kernels)" if a.eager else "cuda graph"}, "dtype": a.dtype,
                  "forward_eval": {"value": a.batch / ms_f * 1e3, "unit": "images/s", "ms": ms_f, "gpu_launches": lf},
                  "forward_backward_train": {"value": a.batch / ms_fb * 1e3, "unit": "images/s", "ms": ms_fb,
                                             "gpu_launches": lfb, "note": "includes the N sequential steps of the "
                                             "batch-axis positional BiGRU (quirk Q1) forward and backward"}}))
