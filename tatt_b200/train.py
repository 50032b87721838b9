"""Data-parallel training step around the hot path (SURVEY 8e + 8f-3).

One process per GPU; the batch is sharded on N; the only exchange is ONE all-reduce of the flat fp32
gradient bucket (NCCL over NVLink on GPUs, gloo in the CPU tests).  Parameters that never receive a
gradient (the 14 dead Q3 tensors) are not part of the bucket.  After the all-reduce the global-norm clip
(clip_grad_norm_(., 0.25), interfaces/super_resolution.py:1083-1085) and Adam(lr 1e-3, betas (0.5, 0.999))
(interfaces/base.py:557-558) run as one fused CUDA kernel over the flat parameter buffer.

BatchNorm statistics stay per-rank, like each replica of the reference's nn.DataParallel (base.py:390):
an N-way run reproduces N reference replicas at batch N_local, not one reference run at N_local*world.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def shard_batch(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of rank's slice of a global batch (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(n_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat fp32 views of the parameters that receive gradients.  Host-side logic only (layout,
    packing, the collective); device-agnostic so it is covered by gloo tests on CPU."""

    def __init__(self, params: Sequence[torch.nn.Parameter], flatten_params: bool = True):
        self.params: List[torch.nn.Parameter] = list(params)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4            # keep every slice 16-byte aligned
        self.numel = off
        dev = self.params[0].device
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_param: Optional[Tensor] = None
        if flatten_params:
            self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
            for p, o in zip(self.params, self.offsets):
                self.flat_param[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + p.numel()].view(p.shape)

    @classmethod
    def from_model_after_backward(cls, model: torch.nn.Module, flatten_params: bool = True) -> "GradBucket":
        return cls([p for p in model.parameters() if p.grad is not None], flatten_params)

    def pack(self) -> Tensor:
        for p, o in zip(self.params, self.offsets):
            if p.grad is None:
                raise RuntimeError("parameter of the gradient bucket has no gradient this step")
            self.flat_grad[o:o + p.numel()].copy_(p.grad.reshape(-1))
        return self.flat_grad

    CHUNK = 16384

    def pack_cuda(self, sq: Optional[Tensor] = None) -> Tensor:
        """The same packing as ONE kernel launch (tatt_multi_copy) instead of one memcpy per parameter: a table of
        {source address, destination offset, count} chunks is staged through pinned host memory (graph replays re-read
        it; addresses inside a captured graph's private pool are stable).  `sq` (1 float) receives sum(g^2)."""
        from . import _cabi, ops
        rows = []
        for p, o in zip(self.params, self.offsets):
            g = p.grad
            if g is None:
                raise RuntimeError("parameter of the gradient bucket has no gradient this step")
            if not g.is_contiguous():
                g = g.contiguous()
                p.grad = g
            base, n = g.data_ptr(), p.numel()
            for c in range(0, n, self.CHUNK):
                rows.append((base + 4 * c, o + c, min(self.CHUNK, n - c)))
        if getattr(self, "_tab_host", None) is None or self._tab_host.shape[0] != len(rows):
            self._tab_host = torch.empty(len(rows), 3, dtype=torch.int64).pin_memory()
            self._tab_dev = torch.empty(len(rows), 3, dtype=torch.int64, device=self.flat_grad.device)
        self._tab_host.copy_(torch.tensor(rows, dtype=torch.int64))
        self._tab_dev.copy_(self._tab_host, non_blocking=True)
        _cabi.call("tatt_multi_copy", self._tab_dev.data_ptr(), len(rows), self.flat_grad.data_ptr(),
                   None if sq is None else sq.data_ptr(), ops._stream())
        return self.flat_grad

    def allreduce_mean(self, group=None) -> Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            self.flat_grad.div_(dist.get_world_size(group))
        return self.flat_grad

    def unpack_into_grads(self) -> None:
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)


class Trainer:
    """fwd + bwd + (all-reduce) + fused clip/Adam for TSRN_TL_TRANS / TSRN on CUDA (eager launches)."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.5, 0.999), eps: float = 1e-8,
                 max_norm: float = 0.25, group=None, image_loss: Optional[Sequence[float]] = None):
        """image_loss = (w0, w1): the third argument of step() / forward_backward() is then the HR TARGET image and the
        step backpropagates `ImageLoss(gradient=True, loss_weight=[w0, w1])(out, target).mean() * 100`, the reference's
        SR loss (interfaces/super_resolution.py:666; loss/image_loss.py:10-34; base.py builds it with [1, 1e-4]), computed
        by csrc/loss.cu (no torch ops).  `self.loss` then holds the per-sample loss vector of the last step.
        image_loss = None: the third argument is an explicit upstream gradient d(loss)/d(out)."""
        self.model, self.lr, self.betas, self.eps, self.max_norm, self.group = model, lr, betas, eps, max_norm, group
        self.image_loss = None if image_loss is None else (float(image_loss[0]), float(image_loss[1]))
        self.loss: Optional[Tensor] = None
        self._loss_ws = None
        self.bucket: Optional[GradBucket] = None
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    def forward_backward(self, x: Tensor, text_emb: Optional[Tensor], grad_out: Tensor):
        for p in self.model.parameters():
            p.grad = None
        res = self.model(x, text_emb) if text_emb is not None else self.model(x)
        out = res[0] if isinstance(res, tuple) else res
        if self.image_loss is not None:
            grad_out = self._image_loss_grad(out, grad_out)
        torch.autograd.backward([out], [grad_out])
        return out

    def _image_loss_grad(self, out: Tensor, target: Tensor) -> Tensor:
        """d/d(out) of ImageLoss(out, target).mean() * 100 through the C-ABI; keeps the per-sample losses in self.loss."""
        from . import _cabi, ops
        o = out.detach()
        if not o.is_contiguous():
            o = o.contiguous()
        n, c, h, w = o.shape
        if target.shape != o.shape:
            raise RuntimeError("Trainer(image_loss=...): target %s != output %s" % (tuple(target.shape), tuple(o.shape)))
        if self._loss_ws is None or self._loss_ws[0].shape[0] != n:
            self._loss_ws = (torch.empty(n, dtype=torch.float32, device=o.device),
                             torch.empty(n, 3, h, w, 2, dtype=torch.float32, device=o.device),
                             torch.empty(2 * n, dtype=torch.float64, device=o.device),
                             torch.full((n,), 100.0 / n, dtype=torch.float32, device=o.device))
        loss, G, ws, gl = self._loss_ws
        w0, w1 = self.image_loss
        st = ops._stream()
        _cabi.call("tatt_image_loss_fwd", o.data_ptr(), target.data_ptr(), loss.data_ptr(), G.data_ptr(), n, c, h, w,
                   w0, w1, ws.data_ptr(), st)
        dout = torch.empty_like(o)
        _cabi.call("tatt_image_loss_bwd", o.data_ptr(), target.data_ptr(), G.data_ptr(), gl.data_ptr(), dout.data_ptr(),
                   n, c, h, w, w0, w1, st)
        self.loss = loss
        return dout

    def _ensure_bucket(self) -> GradBucket:
        if self.bucket is None:
            self.bucket = GradBucket.from_model_after_backward(self.model)
            b = self.bucket
            self.m = torch.zeros_like(b.flat_grad)
            self.v = torch.zeros_like(b.flat_grad)
            self.sq = torch.zeros(1, dtype=torch.float32, device=b.flat_grad.device)
            self.step_state = torch.zeros(2, dtype=torch.int64, device=b.flat_grad.device)   # {unused, step}
        return self.bucket

    def _allreduce(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.bucket.flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def _pack(self) -> None:
        """Gradients -> flat bucket in one launch; on a single rank the same pass also yields sum(g^2)."""
        self._ensure_bucket().pack_cuda(self.sq if self.world == 1 else None)

    def _update_kernels(self) -> None:
        """sum(g^2) of the SUMMED gradient (already produced by the packing pass on one rank), then clip + Adam;
        the 1/world averaging is folded in."""
        from . import _cabi, ops
        b, st = self.bucket, ops._stream()
        g = b.flat_grad
        for p, o in zip(b.params[:4], b.offsets[:4]):          # cheap guard against re-assigned p.data (model.to(), ...)
            if p.data_ptr() != b.flat_param.data_ptr() + 4 * o:
                raise RuntimeError("a parameter was re-assigned after the Trainer flattened it (model.to()/.float()/"
                                   "load with assign=True?); rebuild the Trainer")
        if self.world > 1:
            # deterministic reduction: every replica must clip with the bit-identical norm of the all-reduced gradient
            if getattr(self, "_sq_ws", None) is None:
                self._sq_ws = torch.empty(1024, dtype=torch.float32, device=g.device)
            _cabi.call("tatt_sqnorm_det", g.data_ptr(), b.numel, self.sq.data_ptr(), self._sq_ws.data_ptr(), 1024, st)
        _cabi.call("tatt_rng_advance", self.step_state.data_ptr(), st)
        _cabi.call("tatt_adam_clip_step", b.flat_param.data_ptr(), g.data_ptr(), self.m.data_ptr(),
                   self.v.data_ptr(), b.numel, self.sq.data_ptr(), self.max_norm, self.lr, self.betas[0],
                   self.betas[1], self.eps, self.step_state.data_ptr(), 1.0 / self.world, st)
        ops.bump_weights_epoch()          # raw-pointer update: invalidate weight-derived caches (eval qpos)

    def optimizer_step(self) -> Tensor:
        self._pack()
        self._allreduce()
        self._update_kernels()
        return self.sq

    # ---- resume support: Adam moments + step counter (parameters / BN buffers are in model.state_dict())
    def state_dict(self) -> dict:
        if self.bucket is None:
            return {"step": 0, "exp_avg": {}, "exp_avg_sq": {}}
        b = self.bucket
        names = {id(p): n for n, p in self.model.named_parameters()}
        m = {names[id(p)]: self.m[o:o + p.numel()].view(p.shape).clone() for p, o in zip(b.params, b.offsets)}
        v = {names[id(p)]: self.v[o:o + p.numel()].view(p.shape).clone() for p, o in zip(b.params, b.offsets)}
        return {"step": int(self.step_state[1].item()), "exp_avg": m, "exp_avg_sq": v}

    def load_state_dict(self, sd: dict) -> None:
        """Needs the bucket layout, i.e. call after at least one forward_backward (or pass a model whose
        parameters all have .grad); restores Adam's exp_avg / exp_avg_sq / step."""
        b = self._ensure_bucket()
        names = {id(p): n for n, p in self.model.named_parameters()}
        for p, o in zip(b.params, b.offsets):
            n = names[id(p)]
            if n in sd["exp_avg"]:
                self.m[o:o + p.numel()].copy_(sd["exp_avg"][n].reshape(-1))
                self.v[o:o + p.numel()].copy_(sd["exp_avg_sq"][n].reshape(-1))
        self.step_state[1] = int(sd["step"])

    def step(self, x: Tensor, text_emb: Optional[Tensor], grad_out: Tensor):
        out = self.forward_backward(x, text_emb, grad_out)
        self.optimizer_step()
        return out


def graph_node_counts(g: "torch.cuda.CUDAGraph"):
    """[kernel, memcpy, memset, other] node counts of a captured (keep_graph=True, not yet instantiated) CUDA graph."""
    import ctypes
    from . import _cabi
    c = (ctypes.c_int * 4)()
    _cabi.call_host("tatt_graph_node_counts", ctypes.c_void_p(int(g.raw_cuda_graph())), c)
    return list(c)


class GraphedTrainer(Trainer):
    """The same step captured into CUDA graphs for fixed shapes: graph 1 = forward + backward + gradient
    packing, (eager NCCL all-reduce when world > 1), graph 2 = grad-norm + clip + Adam.  Three host launches
    per step instead of ~800; dropout and the Adam step counter read device-resident state, so replays
    advance them."""

    def __init__(self, model: torch.nn.Module, x_shape, text_shape, out_shape, **kw):
        super().__init__(model, **kw)
        dev = next(model.parameters()).device
        self.x = torch.zeros(x_shape, dtype=torch.float32, device=dev)
        self.text = torch.zeros(text_shape, dtype=torch.float32, device=dev) if text_shape is not None else None
        self.grad_out = torch.zeros(out_shape, dtype=torch.float32, device=dev)
        self.graph_fb = self.graph_opt = None
        self.out: Optional[Tensor] = None

    def _snapshot(self):
        """Everything a training step mutates: parameters, BN buffers, Adam state, step counter, dropout RNG."""
        from .ops import DeviceRNG
        dev = self.x.device
        snap = {"params": [p.detach().clone() for p in self.model.parameters()],
                "buffers": [b.detach().clone() for b in self.model.buffers()],
                "rng": DeviceRNG.get(dev).state.clone()}
        if self.bucket is not None:
            snap.update(m=self.m.clone(), v=self.v.clone(), step=self.step_state.clone())
        return snap

    def _restore(self, snap) -> None:
        from .ops import DeviceRNG
        with torch.no_grad():
            for p, s in zip(self.model.parameters(), snap["params"]):
                p.copy_(s)
            for b, s in zip(self.model.buffers(), snap["buffers"]):
                b.copy_(s)
            DeviceRNG.get(self.x.device).state.copy_(snap["rng"])
            if "m" in snap:
                self.m.copy_(snap["m"]); self.v.copy_(snap["v"]); self.step_state.copy_(snap["step"])
            else:
                self.m.zero_(); self.v.zero_(); self.step_state.zero_()

    def capture(self, warmup: int = 2) -> None:
        """Warm-up steps run on whatever is in the static input buffers and are ROLLED BACK afterwards (weights, BN
        running statistics, Adam moments, step counter, dropout stream), so capturing has no training side effect."""
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                Trainer.step(self, self.x, self.text, self.grad_out)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph_fb = torch.cuda.CUDAGraph(keep_graph=True)
        with torch.cuda.graph(self.graph_fb):
            self.out = self.forward_backward(self.x, self.text, self.grad_out)
            self._pack()
            if self.world == 1:                   # no collective in between: one graph for the whole step
                self._update_kernels()
        if self.world > 1:
            self.graph_opt = torch.cuda.CUDAGraph(keep_graph=True)
            with torch.cuda.graph(self.graph_opt, pool=self.graph_fb.pool()):
                self._update_kernels()
        self.node_counts = graph_node_counts(self.graph_fb)        # [kernel, memcpy, memset, other] per replayed step
        self.graph_fb.instantiate()
        if self.graph_opt is not None:
            self.node_counts = [a + b for a, b in zip(self.node_counts, graph_node_counts(self.graph_opt))]
            self.graph_opt.instantiate()
        torch.cuda.synchronize()
        self._restore(snap)                        # capture itself does not execute, but the warm-up did

    def step(self, x: Optional[Tensor] = None, text_emb: Optional[Tensor] = None, grad_out: Optional[Tensor] = None):
        """Inputs may live on the host (pinned) or the device; None -> reuse the static buffers."""
        if x is not None and x is not self.x:
            self.x.copy_(x, non_blocking=True)
        if text_emb is not None and text_emb is not self.text:
            self.text.copy_(text_emb, non_blocking=True)
        if grad_out is not None and grad_out is not self.grad_out:
            self.grad_out.copy_(grad_out, non_blocking=True)
        if self.graph_fb is None:
            self.capture()
        self.graph_fb.replay()
        if self.world > 1:
            self._allreduce()
            self.graph_opt.replay()
        from . import ops
        ops.bump_weights_epoch()
        return self.out


class GraphedForward:
    """Eval-mode forward (`model(x, text_emb)` under no_grad) captured into one CUDA graph for fixed shapes.  The
    recurrent positional encoding depends on the weights only, so it is computed once (cached by the module)."""

    def __init__(self, model: torch.nn.Module, x_shape, text_shape):
        self.model = model.eval()
        dev = next(model.parameters()).device
        self.x = torch.zeros(x_shape, dtype=torch.float32, device=dev)
        self.text = torch.zeros(text_shape, dtype=torch.float32, device=dev) if text_shape is not None else None
        self.graph = None
        self.out = None

    def _run(self):
        with torch.no_grad():
            res = self.model(self.x, self.text) if self.text is not None else self.model(self.x)
        return res

    def capture(self, warmup: int = 2) -> None:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph(keep_graph=True)
        with torch.cuda.graph(self.graph):
            self.out = self._run()
        self._ptrs = self._param_ptrs()
        self.node_counts = graph_node_counts(self.graph)
        self.graph.instantiate()

    def _param_ptrs(self):
        return [t.data_ptr() for t in list(self.model.parameters()) + list(self.model.buffers())]

    def _refresh_qpos(self) -> None:
        """The captured graph reads the module's cached positional encoding; after a weight update (tracked by
        ops.weights_epoch / torch version counters) recompute it eagerly INTO THE SAME STORAGE before replaying."""
        ig = getattr(self.model, "infoGen", None)
        if ig is not None and self.text is not None:
            with torch.no_grad():
                ig.query_pos(self.text.shape[0], self.x.shape[2], self.x.shape[3])

    def __call__(self, x: Optional[Tensor] = None, text_emb: Optional[Tensor] = None):
        if self.graph is not None and self._ptrs != self._param_ptrs():
            # the graph holds raw parameter addresses: a Trainer that flattened the parameters into its bucket (or
            # model.to() / load with assign=True) moved them -> the captured graph would read dead storage
            self.graph = None
        if self.graph is None:
            self.capture()
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if text_emb is not None:
            self.text.copy_(text_emb, non_blocking=True)
        self._refresh_qpos()
        self.graph.replay()
        return self.out
