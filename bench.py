#!/usr/bin/env python
"""bench.py -- SR images/sec of the TATT hot path (BASELINE.json metric), one JSON line on rank 0.

Workload (BASELINE.json configs[1]): TATT-TSRN = TSRN_TL_TRANS, LR 32x128 -> SR 64x256 ("G32": width=256,
height=64, STN off -- the reference itself cannot run STN at this geometry), batch 64 per GPU, fp32,
train mode (dropout 0.1 active), synthetic inputs, seed-1234 random-init weights.
A "step" = forward + backward (+ one NCCL all-reduce of the flat gradient when N>1) + fused global-norm
clip + Adam over the flat parameter buffer.  Weak scaling: per-GPU batch is fixed.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                      the reference's CPU path (oracle port), rank 0 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "SR images/sec (32x128 LR, fwd+bwd)"
UNIT = "images/s"
NCU_CONV_TRAFFIC_BYTES = 182.5e6         # split pass 91.1 MB + conv3x3_roll_kernel 91.4 MB (profiles/r1d_ncu_full_conv3x3_roll.csv)
ALG_FLOPS_FWD_BWD_G32 = 3.0 * 9.33e9     # ~9.33 GFLOP/img forward (RPE input-proj hoisted) x3 for fwd+bwd


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--geometry", default="g32", choices=["g32", "g16"])
    ap.add_argument("--cpu-sample", type=int, default=4, help="images per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python (no CUDA graphs)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="f32 (default, the configuration BASELINE.json's metric is quoted on: fp32 parity via "
                         "bf16 hi/lo split MMAs) or bf16 (configs 3/4: single-plane bf16 operands, fp32 accumulate)")
    return ap.parse_args()


def geometry(name):
    if name == "g32":
        return dict(scale_factor=2, width=256, height=64, STN=False, mask=True), 32, 128
    return dict(scale_factor=2, width=128, height=32, STN=True, mask=True), 16, 64


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(ctor_kw, h, w, sample_n, steps, warmup):
    """The reference's own CPU path (oracle port: the same torch CPU ops in the same order), fwd+bwd,
    dropout 0.1, all host threads."""
    import tatt_b200
    from oracle import tatt_oracle as orc
    torch.manual_seed(1234)
    net = tatt_b200.TSRN_TL_TRANS(**ctor_kw)          # parameter container only (CPU); never run
    sd = orc.clone_sd(net.state_dict(), requires_grad=True)
    x, tp = orc.synthetic_inputs(sample_n, h, w, seed=1234)
    cores = torch.get_num_threads()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for v in sd.values():
            v.grad = None
        out, _, _ = orc.tsrn_tl_trans_forward(sd, x, tp, training=True, stn=ctor_kw["STN"], dropout_p=0.1)
        out.mean().backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tot = sum(times)
    return {"value": sample_n * len(times) / tot, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d-image batch per step x %d steps of the same geometry (fwd+bwd, dropout 0.1), "
                      "oracle port of the reference on torch CPU fp32" % (sample_n, len(times)),
            "ms_per_step": 1e3 * tot / len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kw, h, w = geometry(args.geometry)
    cb = cpu_baseline(kw, h, w, args.cpu_sample, args.steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TSRN_TL_TRANS %s fwd+bwd, CPU sample batch %d" % (args.geometry, args.cpu_sample)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def conv_roofline(dev, batch, h, w, peaks):
    """Dominant kernel family = the 3x3 64->64 implicit-GEMM convolution on tcgen05 (11 forward instances per image
    plus their data / weight gradients).  One "launch" = the operand split pass + conv3x3_roll_kernel of one conv, as
    the model's forward issues it.  Algorithmic FLOPs per launch = 2 * pixels * 64 * 576 (SURVEY 8d conv figure x
    pixels); timed alone with CUDA events on the launching stream, L2 flushed (256 MB write) between launches.
    `kernel_only_*`: the same call with the bf16 planes already in the workspace (flag 2048: what the data-gradient
    pass does), i.e. conv3x3_roll_kernel by itself."""
    from tatt_b200 import _cabi, ops
    x = torch.randn(batch, h, w, 64, device=dev)
    wt = torch.randn(64, 64, 3, 3, device=dev) * 0.05
    b = torch.zeros(64, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)          # 256 MiB > L2
    ws, wsb = ops._ws(x, x.numel() + 576 * 64)
    for _ in range(3):
        ops.conv2d_fwd(x, wt, b, 1)
    wtp = ops.conv_pack(wt, 64, 64, False)
    y = torch.empty_like(x)

    def timed(flags):
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.call("tatt_conv2d_igemm", x.data_ptr(), wtp.data_ptr(), b.data_ptr(), y.data_ptr(), batch, h, w, 64,
                       64, 3, 3, 1, 1, ops._precision_flag | flags, ws.data_ptr(), wsb, ops._stream())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return sum(ts) / len(ts)

    t = timed(0)
    tk = timed(ops.F_A_VALID)
    flops = 2.0 * batch * h * w * 64 * 576
    peak = peaks.get("bf16_tflops", 1590.0)
    # dram__bytes_read.sum + dram__bytes_write.sum of the split pass + conv kernel from the committed ncu --set full
    # capture (profiles/), same shape; algorithmic bytes = 67 MB in + 67 MB out
    traffic = NCU_CONV_TRAFFIC_BYTES if (batch, h, w) == (64, 32, 128) else None
    bf16 = bool(ops._precision_flag & ops.F_BF16)
    return {"bound": "tensor", "kernel": "split_dense_kernel + conv3x3_roll_kernel (conv3x3 64->64, NHWC; persistent "
            "rolling-halo TMA tiles + TMA weight ring, tcgen05 kind::f16 with a bf16 hi/lo operand split: A_hi x [B_hi|B_lo] "
            "(N=128) + A_lo x B_hi (N=64) per k-step, ping-pong fp32 TMEM accumulators)",
            "achieved": flops / t / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flops / t / 1e12 / peak,
            "traffic": traffic, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst), of measured" if "bf16_tflops" in peaks
            else "fallback 1.59 PFLOP/s, of fallback", "launch_ms": t * 1e3, "algorithmic_flops_per_launch": flops,
            "kernel_only_ms": tk * 1e3, "kernel_only_tflops": flops / tk / 1e12, "kernel_only_frac": flops / tk / 1e12 / peak,
            "note": ("bf16 mode: one MMA per product" if bf16 else
                     "fp32-parity mode costs 3 bf16 products per MAC (ceiling = 1/3 of the bf16 peak); with C_out = 64 "
                     "the MMAs are bound by shared-memory operand fetch, not by the tensor-pipe math rate")}


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import tatt_b200
    from tatt_b200 import _cabi, ops
    from tatt_b200.train import GraphedTrainer, Trainer
    from oracle import tatt_oracle as orc   # only for the synthetic-input recipe and the cpu_baseline leg

    kw, h, w = geometry(args.geometry)
    torch.manual_seed(1234)
    model = tatt_b200.TSRN_TL_TRANS(**kw).to(dev).train()
    tatt_b200.manual_seed(1234 + rank)
    tatt_b200.set_precision("bf16" if args.dtype == "bf16" else "fp32")
    B = args.batch
    if args.eager:
        trainer = Trainer(model)
    else:
        trainer = GraphedTrainer(model, (B, 4, h, w), (B, 37, 1, 26), (B, 4, 2 * h, 2 * w))
    x_h, tp_h = orc.synthetic_inputs(B, h, w, seed=1234 + rank)
    x_h, tp_h = x_h.pin_memory(), tp_h.pin_memory()
    g_h = torch.randn(B, 4, 2 * h, 2 * w, generator=torch.Generator().manual_seed(7)) / (B * 4 * 4 * h * w)
    x_d, tp_d, g_d = x_h.to(dev), tp_h.to(dev), g_h.to(dev)
    metric_d = torch.zeros(1, device=dev)
    metric_h = torch.zeros(1).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _cabi.launch_count
        e0.record()
        for _ in range(nsteps):
            if e2e and args.eager:
                xs = x_h.to(dev, non_blocking=True)
                ts_ = tp_h.to(dev, non_blocking=True)
            elif e2e:
                xs, ts_ = x_h, tp_h                                # pinned host -> static device buffers
            elif args.eager:
                xs, ts_ = x_d, tp_d
            else:
                xs, ts_ = None, None                               # inputs already resident in HBM
            out = trainer.step(xs, ts_, g_d if args.eager else None)
            if e2e:
                _cabi.call("tatt_sqnorm", out.data_ptr(), out.numel(), metric_d.data_ptr(), 1, ops._stream())
                metric_h.copy_(metric_d, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, _cabi.launch_count - l0

    if not args.eager:
        trainer.x.copy_(x_d); trainer.text.copy_(tp_d); trainer.grad_out.copy_(g_d)
        l0 = _cabi.launch_count
        trainer.capture()
        per_step_launches = (_cabi.launch_count - l0) // 3        # 2 eager warm-up steps + 1 captured step
    timed(max(args.warmup, 3), False)
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(args.steps, False)
    clocks = sampler.stop()
    timed(1, True)
    ms_e2e, _ = timed(args.steps, True)

    # forward-only (inference) throughput of the same model / batch, eval mode, CUDA graph, inputs resident
    from tatt_b200.train import GraphedForward
    fwd = GraphedForward(model, (B, 4, h, w), (B, 37, 1, 26))
    fwd.x.copy_(x_d); fwd.text.copy_(tp_d)
    fwd.capture()
    for _ in range(3):
        fwd()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        fwd()
    f1.record()
    barrier()
    ms_fwd = f0.elapsed_time(f1)
    if world > 1:
        tf = torch.tensor([ms_fwd], device=dev)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        ms_fwd = tf.item()
    model.train()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    if not args.eager:
        launches = per_step_launches * args.steps               # kernels replayed from the captured graphs
    value = B * world * args.steps / (ms * 1e-3)
    e2e_v = B * world * args.steps / (ms_e2e * 1e-3)
    roof = conv_roofline(dev, B, h, w, peaks)
    roof["model_algorithmic_tflops"] = value * ALG_FLOPS_FWD_BWD_G32 / 1e12 if args.geometry == "g32" else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "TSRN_TL_TRANS(width=%d,height=%d,STN=%s) fwd+bwd+clip+Adam, per-GPU batch %d, "
                                   "train mode dropout 0.1" % (kw["width"], kw["height"], kw["STN"], B),
                       "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "launch": "eager" if args.eager else "cuda-graphs (fwd+bwd+pack | allreduce | clip+Adam)",
                       "l2": "activations per step (>2 GB) exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_v, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (x_h.numel() + tp_h.numel()) * 4, "d2h_bytes_per_step": 4},
            "forward_only": {"value": B * world * args.steps / (ms_fwd * 1e-3), "unit": UNIT,
                             "ms_per_step": ms_fwd / args.steps,
                             "note": "eval-mode forward of the same batch, CUDA graph, cached positional encoding"},
            "roofline": roof}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(kw, h, w, args.cpu_sample, 2, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
