"""Generate tests/golden/image_loss_n3.pt from the LIVE reference `loss/image_loss.py:ImageLoss` (imported unmodified;
only the absent IPython module is stubbed).  Run in the build container only:  python tests/golden/make_golden_loss.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402


def load_ref_loss():
    rh.load()                                   # installs the IPython stub and the sys.path entry
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("loss.image_loss")


def inputs(n=3, h=16, w=40, seed=1234):
    g = torch.Generator().manual_seed(seed)
    out = torch.tanh(torch.randn(n, 4, h, w, generator=g))
    tgt = torch.rand(n, 4, h, w, generator=g)
    out[0, :3, 3:6, 5:9] = tgt[0, :3, 3:6, 5:9]            # a flat patch: |m_out - m_tgt| == 0 exactly there
    return out, tgt


def main():
    mod = load_ref_loss()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        crit = mod.ImageLoss(gradient=True, loss_weight=[20, 1e-4])
    out, tgt = inputs()
    fx = {"out": out, "target": tgt, "torch": str(torch.__version__)}
    o32 = out.clone().requires_grad_(True)
    loss = crit(o32, tgt)
    (loss.mean() * 100).backward()
    fx["loss"] = loss.detach().clone()
    fx["dout"] = o32.grad.clone()
    o64 = out.double().requires_grad_(True)
    l64 = crit(o64, tgt.double())
    (l64.mean() * 100).backward()
    fx["loss64"] = l64.detach().clone()
    fx["dout64"] = o64.grad.clone()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image_loss_n3.pt")
    torch.save(fx, path)
    print("wrote", path, "loss", loss.tolist())
    loss_block()


def loss_block():
    """SemanticLoss / TRI_SSIM / torch_rotate_img outputs of the live reference -> loss_block_n3.pt"""
    import ast
    import importlib
    import warnings
    import torch.nn.functional as F
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sl = importlib.import_module("loss.semantic_loss")
        sp = importlib.import_module("utils.ssim_psnr")
        g = torch.Generator().manual_seed(1234)
        p = torch.softmax(torch.randn(26, 3, 37, generator=g), -1)
        q = torch.softmax(torch.randn(26, 3, 37, generator=g), -1)
        a, b, c = [torch.rand(3, 4, 16, 40, generator=g) for _ in range(3)]
        arcs = (torch.rand(3, generator=g) - 0.5) * 0.2
        offs = torch.rand(3, generator=g)
        src = open(os.path.join(rh.REF_ROOT, "interfaces", "super_resolution.py")).read()
        fn = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "torch_rotate_img"][0]
        ns = {"torch": torch, "F": F}
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "torch_rotate_img", "exec"), ns)
        fx = {"pred": p, "gt": q, "semantic": sl.SemanticLoss()(p, q).clone(), "a": a, "b": b, "c": c,
              "tri_ssim": sp.TRI_SSIM()(a, b, c).clone(), "tri_ssim_per_sample": sp.TRI_SSIM(size_average=False)(a, b, c).clone(),
              "arcs": arcs, "offs": offs, "rotated": ns["torch_rotate_img"](None, a, arcs, offs).clone(),
              "torch": str(torch.__version__)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_block_n3.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
