"""One `ncu --set full` sample of every time-relevant C-ABI entry of a training step, at the benchmarked shapes.

Runs bench.py's workload eagerly (G32, N = 64, fp32 parity mode, ImageLoss step), then ONE more step in which the first
call of every distinct (entry point, shape) is bracketed by cudaProfilerStart / cudaProfilerStop -- so, under

    ncu --set full --clock-control none --profile-from-start off -o /tmp/r2_step python tools/ncu_step_sampler.py
    ncu -i /tmp/r2_step.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_step_sample_raw.csv

exactly the kernels of those calls are captured (operand-split passes included), in step order and in a warm, realistic
memory state.  tools/ncu_traffic.py turns the raw CSV into the per-kernel summary + profiles/traffic_by_entry.json."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import tatt_b200
from oracle import tatt_oracle as orc
from tatt_b200 import _cabi
from tatt_b200.train import Trainer

SKIP = {"tatt_memcpy_d2d", "tatt_memset0", "tatt_rng_advance", "tatt_conv_weight_pack", "tatt_conv_weight_unpack_grad",
        "tatt_multi_copy", "tatt_bn_eval_stats"}
B = int(os.environ.get("N", 64))
kw, h, w = bench.geometry("g32")
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = tatt_b200.TSRN_TL_TRANS(**kw).to(dev).train()
tatt_b200.manual_seed(1234)
trainer = Trainer(model, image_loss=(1.0, 1e-4))
x, tp = orc.synthetic_inputs(B, h, w, seed=1234)
hr = torch.rand(B, 4, 2 * h, 2 * w, generator=torch.Generator().manual_seed(7))
x, tp, hr = x.to(dev), tp.to(dev), hr.to(dev)
for _ in range(2):
    trainer.step(x, tp, hr)
torch.cuda.synchronize()

seen = set()
orig = _cabi.call
log = []


def sampled(name, *args):
    key = (name, bench.entry_cost(name, args)[0])
    if name in SKIP or key in seen:
        return orig(name, *args)
    seen.add(key)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    try:
        return orig(name, *args)
    finally:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        log.append("%s %s" % key)


_cabi.call = sampled
for mod in list(sys.modules.values()):          # modules that did `from . import _cabi` call through the attribute: fine
    pass
trainer.step(x, tp, hr)
torch.cuda.synchronize()
_cabi.call = orig
print("sampled %d calls:" % len(log))
for l in log:
    print("  ", l)
