"""Generate tests/golden/crnn_n3.pt from the LIVE reference `model/crnn/crnn.py:CRNN(32, 1, 37, 256)` (the student
text-prior generator, interfaces/base.py:712-726).  Build container only:  python tests/golden/make_golden_crnn.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402

SEED = 1234


def load_ref_crnn():
    rh.load()
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("model.crnn.crnn")


def build(mod):
    torch.manual_seed(SEED)
    net = mod.CRNN(32, 1, 37, 256)
    g = torch.Generator().manual_seed(SEED + 1)
    for m in net.modules():                                   # non-trivial BN statistics / affine
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(1 + 0.2 * torch.randn(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return net


def inputs(n=3, seed=SEED):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, 16, 64, generator=g)              # LR crops as the training loop feeds them (RGB part)


def sample(t, k=512, seed=3):
    idx = torch.randperm(t.numel(), generator=torch.Generator().manual_seed(seed))[:k].clone()
    return {"shape": tuple(t.shape), "idx": idx, "val": t.detach().reshape(-1)[idx].clone(),
            "absmax": t.detach().abs().max().item(), "sum": t.detach().double().sum().item()}


def main():
    sys.path.insert(0, ROOT)
    from oracle import crnn_oracle as co
    mod = load_ref_crnn()
    fx = {"torch": str(torch.__version__)}
    for training in (False, True):
        net = build(mod).train(training)
        gray = co.parse_crnn_data(inputs())
        logits = net(gray)
        key = "train" if training else "eval"
        fx[key + "_logits"] = logits.detach().clone()
        if training:
            gen = torch.Generator().manual_seed(99)
            (logits * torch.randn(logits.shape, generator=gen)).sum().backward()
            fx["train_grads"] = {n: sample(p.grad) for n, p in net.named_parameters()}
            fx["train_buffers"] = {n: b.detach().clone() for n, b in net.named_buffers()}
    fx["gray"] = gray.detach().clone()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crnn_n3.pt")
    torch.save(fx, path)
    print("wrote", path, tuple(fx["eval_logits"].shape), os.path.getsize(path))


if __name__ == "__main__":
    main()
