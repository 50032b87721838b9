"""Loss block that follows the hot path in the reference's training step (first "next" row built):
`loss/image_loss.py:ImageLoss` -- same constructor, `forward(out_images, target_images, grad_mask=None)` and
per-sample return value; the arithmetic (forward and backward) runs in csrc/loss.cu through the C-ABI."""
from __future__ import annotations

import torch

from . import _cabi, ops

Tensor = torch.Tensor


class _ImageLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out: Tensor, tgt: Tensor, w0: float, w1: float):
        out = ops._chk(out.contiguous(), "out_images")
        tgt = ops._chk(tgt.contiguous(), "target_images")
        if out.shape != tgt.shape or out.dim() != 4 or out.shape[1] < 3:
            raise RuntimeError("ImageLoss: out/target must be equal-shaped [N, C>=3, H, W] tensors, got %s and %s" % (
                tuple(out.shape), tuple(tgt.shape)))
        n, c, h, w = out.shape
        need = ctx.needs_input_grad[0]
        loss = ops.empty(n, like=out)
        G = ops.empty(n, 3, h, w, 2, like=out) if need else None
        ws = torch.empty(2 * n, dtype=torch.float64, device=out.device)
        _cabi.call("tatt_image_loss_fwd", ops._p(out), ops._p(tgt), ops._p(loss), ops._p(G), n, c, h, w, w0, w1,
                   ops._p(ws), ops._stream())
        if need:
            ctx.save_for_backward(out, tgt, G)
        ctx.w = (w0, w1)
        return loss

    @staticmethod
    def backward(ctx, gloss: Tensor):
        out, tgt, G = ctx.saved_tensors
        n, c, h, w = out.shape
        dout = torch.empty_like(out)
        _cabi.call("tatt_image_loss_bwd", ops._p(out), ops._p(tgt), ops._p(G), ops._p(gloss.contiguous()), ops._p(dout),
                   n, c, h, w, ctx.w[0], ctx.w[1], ops._stream())
        return dout, None, None, None


class ImageLoss(torch.nn.Module):
    """Drop-in for `loss/image_loss.py:ImageLoss` (reference lines 10-34).  Only `gradient=True` is defined behaviour in
    the reference (`gradient=False` reads an unbound local, image_loss.py:32) and the same error is raised here.
    `grad_mask` is accepted and ignored, like the reference (its use is commented out, image_loss.py:21-23).
    Gradients flow into `out_images` only (the target is data)."""

    def __init__(self, gradient=True, loss_weight=[20, 1e-4]):
        super().__init__()
        self.gradient = gradient
        self.loss_weight = loss_weight

    def forward(self, out_images, target_images, grad_mask=None):
        if not self.gradient:
            raise UnboundLocalError("cannot access local variable 'mse_loss' where it is not associated with a value")
        return _ImageLossFn.apply(out_images, target_images.detach(), float(self.loss_weight[0]),
                                  float(self.loss_weight[1]))


# ----------------------------------------------------------------------------------------------- SemanticLoss
class _SemanticLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: Tensor, gt: Tensor):
        pred = ops._chk(pred.contiguous(), "pred_vec")
        gt = ops._chk(gt.contiguous(), "gt_vec")
        if pred.shape != gt.shape:
            raise RuntimeError("The size of tensor a %s must match the size of tensor b %s" % (
                tuple(gt.shape), tuple(pred.shape)))
        loss = ops.empty(1, like=pred)
        ws = torch.empty(2, dtype=torch.float64, device=pred.device)
        _cabi.call("tatt_semantic_loss_fwd", ops._p(pred), ops._p(gt), pred.numel(), ops._p(loss), ops._p(ws),
                   ops._stream())
        ctx.save_for_backward(pred, gt)
        return loss.view(())

    @staticmethod
    def backward(ctx, gloss: Tensor):
        pred, gt = ctx.saved_tensors
        dp = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        dg = torch.empty_like(gt) if ctx.needs_input_grad[1] else None
        _cabi.call("tatt_semantic_loss_bwd", ops._p(pred), ops._p(gt), ops._p(gloss.contiguous().view(1)), ops._p(dp),
                   ops._p(dg), pred.numel(), ops._stream())
        return dp, dg


class SemanticLoss(torch.nn.Module):
    """Drop-in for `loss/semantic_loss.py:SemanticLoss` (reference lines 10-37): same constructor, `forward(pred_vec,
    gt_vec)` -> scalar `mean|gt - pred| + KLDivLoss()(log(pred + 1e-20), gt + 1e-20)` (lambda1 = lambda2 = 1; the
    reference's cos_sim / margin members are unused there too).  Forward + backward in csrc/loss2.cu."""

    def __init__(self, margin=0.1):
        super().__init__()
        self.margin = margin
        self.lambda1 = 1.0
        self.lambda2 = 1.0

    def forward(self, pred_vec, gt_vec):
        return _SemanticLossFn.apply(pred_vec, gt_vec)


# ----------------------------------------------------------------------------------------------- TRI_SSIM
class _TriSSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1: Tensor, x2: Tensor, x3: Tensor, per_sample: bool):
        xs = [ops._chk(t.contiguous(), "TRI_SSIM image") for t in (x1, x2, x3)]
        if not (xs[0].shape == xs[1].shape == xs[2].shape) or xs[0].dim() != 4:
            raise RuntimeError("TRI_SSIM: the three images must be equal-shaped [N, C, H, W] tensors")
        n, c, h, w = xs[0].shape
        need = any(ctx.needs_input_grad[:3])
        out = ops.empty(n if per_sample else 1, like=xs[0])
        G = ops.empty(5, n, c, h, w, like=xs[0]) if need else None
        ws = torch.empty(n, dtype=torch.float64, device=xs[0].device)
        _cabi.call("tatt_tri_ssim_fwd", ops._p(xs[0]), ops._p(xs[1]), ops._p(xs[2]), ops._p(out), ops._p(G), n, c, h, w,
                   1 if per_sample else 0, ops._p(ws), ops._stream())
        if need:
            ctx.save_for_backward(xs[0], xs[1], xs[2], G)
        ctx.per_sample = per_sample
        return out if per_sample else out.view(())

    @staticmethod
    def backward(ctx, gout: Tensor):
        x1, x2, x3, G = ctx.saved_tensors
        n, c, h, w = x1.shape
        ds = [torch.empty_like(x1) if ctx.needs_input_grad[i] else None for i in range(3)]
        _cabi.call("tatt_tri_ssim_bwd", ops._p(x1), ops._p(x2), ops._p(x3), ops._p(G), ops._p(gout.contiguous().view(-1)),
                   ops._p(ds[0]), ops._p(ds[1]), ops._p(ds[2]), n, c, h, w, 1 if ctx.per_sample else 0, ops._stream())
        return ds[0], ds[1], ds[2], None


class TRI_SSIM(torch.nn.Module):
    """Drop-in for `utils/ssim_psnr.py:TRI_SSIM` (reference lines 231-256 -> `_tri_ssim` 99-128): same constructor and
    `forward(img1, img2, img3)`; only the reference's default 11-tap window is built (its sigma is hard-wired to 1.5,
    `create_window` :34-37).  Separable Gaussian in shared memory, forward + backward in csrc/loss2.cu."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        if window_size != 11:
            raise NotImplementedError("tatt_b200.TRI_SSIM is specialised for the reference's window_size=11")
        self.window_size = window_size
        self.size_average = size_average

    def forward(self, img1, img2, img3):
        return _TriSSIMFn.apply(img1, img2, img3, not self.size_average)


# ----------------------------------------------------------------------------------------------- torch_rotate_img
class _RotateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img: Tensor, arcs: Tensor, offs: Tensor, off_range: float):
        img = ops._chk(img.contiguous(), "image batch")
        n, c, h, w = img.shape
        arcs = ops._chk(arcs.to(device=img.device, dtype=torch.float32).contiguous().view(-1), "arc_batches")
        offs = ops._chk(offs.to(device=img.device, dtype=torch.float32).contiguous().view(-1), "rand_offs")
        if arcs.numel() != n or offs.numel() != n:
            raise RuntimeError("torch_rotate_img: arcs / rand_offs must hold one value per image")
        out = torch.empty_like(img)
        _cabi.call("tatt_rotate_img_fwd", ops._p(img), ops._p(arcs), ops._p(offs), off_range, ops._p(out), n, c, h, w,
                   ops._stream())
        ctx.save_for_backward(arcs, offs)
        ctx.off_range = off_range
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        arcs, offs = ctx.saved_tensors
        dout = dout.contiguous()
        n, c, h, w = dout.shape
        dimg = torch.empty_like(dout)
        _cabi.call("tatt_rotate_img_bwd", ops._p(dout), ops._p(arcs), ops._p(offs), ctx.off_range, ops._p(dimg), n, c, h,
                   w, ops._stream())
        return dimg, None, None, None


def torch_rotate_img(torch_image_batches, arc_batches, rand_offs, off_range=0.2):
    """`TextSR.torch_rotate_img` (interfaces/super_resolution.py:126-157) without the `self`: rotate every image of the
    NCHW batch by its own angle (radians) with the jittered aspect term; differentiable w.r.t. the images."""
    return _RotateFn.apply(torch_image_batches, torch.as_tensor(arc_batches), torch.as_tensor(rand_offs),
                           float(off_range))
