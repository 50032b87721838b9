"""Launch list of ONE eager training step at the benchmarked shapes (G32, N = 64, fp32 parity mode, ImageLoss step):

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches_one_step.csv python tools/ncu_step_launches.py

Two warm steps run unprofiled, then cudaProfilerStart / Stop bracket exactly one Trainer.step (every kernel of the step,
in launch order; serialised by ncu, so the side stream's overlap is not in these times)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import tatt_b200
from oracle import tatt_oracle as orc
from tatt_b200.train import Trainer

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
torch.cuda.profiler.start()
trainer.step(x, tp, hr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
