"""Generate tests/golden/*.pt from the LIVE reference (/root/reference, imported unmodified through
oracle/ref_harness.py).  Run in the build container only:  python tests/golden/make_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import golden_util as gu  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle import tatt_oracle as orc  # noqa: E402


def main():
    ref = rh.load()
    cases = dict(gu.CASES)
    cases.update(gu.BIG_CASES)
    only = sys.argv[1:]
    for name, (cls, kw, N, training) in cases.items():
        if only and name not in only:
            continue
        torch.manual_seed(gu.SEED)
        net = getattr(ref, cls)(**kw)
        rh.zero_dropout(net)
        rh.perturb_(net)
        net.train(training)
        h, w = kw["height"] // 2, kw["width"] // 2
        x, tp = orc.synthetic_inputs(N, h, w, seed=gu.SEED, with_mask=kw.get("mask", True))
        fx = {"case": name, "torch": torch.__version__}
        if cls == "TSRN":
            out = net(x)
            aux = None
        else:
            out, aux = net(x, tp)
        fx["out"] = gu.summarize(out)
        if aux is not None:
            pw = aux["pr_weights"] if training else aux
            fx["pr_weights"] = gu.summarize(pw)
            if training:
                fx["tp_map"] = gu.summarize(aux["spatial_t_emb"])
        for k in ("1", "4", str(kw.get("srb_nums", 5) + 2)):
            fx["block" + k] = gu.summarize(net.block[k], nsamp=256)
        if training:
            # loss exercising every output element with non-uniform weights
            gen = torch.Generator().manual_seed(99)
            wgt = torch.randn(out.shape, generator=gen)
            (out * wgt).sum().backward()
            fx["grads"] = {n: (None if p.grad is None else gu.summarize(p.grad, nsamp=64 if name in gu.BIG_CASES else 8))
                           for n, p in net.named_parameters()}
            # the same backward with the reference evaluated in float64, and the reference's own
            # fp32-vs-fp64 deviation per parameter (its rounding noise: ReLU / max-pool flips, TPS conditioning)
            fx["buffers"] = {n: gu.summarize(b.float(), nsamp=4) for n, b in net.named_buffers()
                             if "running" in n or "num_batches" in n}
            del out, aux
            net.block = {}                                        # drop the fp32 autograd graph before the fp64 pass
            if name in gu.BIG_CASES:
                # the float64 reference pass at N = 64, G32 does not fit this container's 62 GB: the big case keeps
                # the reference's fp32 gradients only (its fp32-vs-fp64 noise at G32 without STN is 3e-5 rel-L2,
                # measured on tatt_g32_train_n2)
                fx["gmax"] = max(p.grad.abs().max().item() for p in net.parameters() if p.grad is not None)
                path = os.path.join(gu.GOLDEN_DIR, name + ".pt")
                torch.save(fx, path)
                print(name, os.path.getsize(path) // 1024, "KiB")
                continue
            torch.manual_seed(gu.SEED)
            n64 = getattr(ref, cls)(**kw)
            rh.zero_dropout(n64)
            rh.perturb_(n64)
            n64 = n64.double().train(True)
            o64 = n64(x.double()) if cls == "TSRN" else n64(x.double(), tp.double())[0]
            (o64 * wgt.double()).sum().backward()
            g32 = dict(net.named_parameters())
            fx["grads64"], fx["noise"] = {}, {}
            for n, p in n64.named_parameters():
                if p.grad is None:
                    fx["grads64"][n] = None
                    continue
                fx["grads64"][n] = gu.summarize(p.grad, nsamp=8)
                fx["noise"][n] = (g32[n].grad.double() - p.grad).abs().max().item()
            fx["gmax"] = max(p.grad.abs().max().item() for p in n64.parameters() if p.grad is not None)
        path = os.path.join(gu.GOLDEN_DIR, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
