"""Generate tests/golden/data_collate_n5.pt from the LIVE reference collate code: the classes
`resizeNormalize` (dataset/dataset.py:1266-1319) and `alignCollate_realWTLAMask` (:1965-2076) are cut out of the
reference source with `ast` and executed unmodified (the module itself cannot be imported here: lmdb / imgaug / ... are
absent).  Inputs are synthetic PIL images (odd sizes, so the bicubic resize is exercised) and label strings covering
the padding rules (empty, 1 char, < 26, > 26, characters outside the alphabet).
Build container only:  python tests/golden/make_golden_data.py"""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"
LABELS = ["hello", "a", "", "supercalifragilisticexpialidocious", "B2b!", "it"]


def images(seed=1234):
    g = np.random.RandomState(seed)
    sizes = [(121, 37), (64, 16), (200, 50), (33, 9), (128, 32), (97, 41)]         # (W, H) of the HR crops
    hr = [Image.fromarray(g.randint(0, 256, (h, w, 3), dtype=np.uint8)) for w, h in sizes]
    lr = [Image.fromarray(g.randint(0, 256, (max(h // 2, 3), max(w // 2, 5), 3), dtype=np.uint8)) for w, h in sizes]
    flat = Image.fromarray(np.full((20, 70, 3), 128, dtype=np.uint8))             # constant image: L == mean everywhere
    hr[4], lr[4] = flat, flat.resize((35, 10))
    return hr, lr


def load_reference_classes():
    import cv2
    from torchvision import transforms
    src = open(os.path.join(rh.REF_ROOT, "dataset", "dataset.py")).read()
    tree = ast.parse(src)
    want = {"resizeNormalize", "alignCollate_realWTLAMask"}
    ns = {"np": np, "torch": torch, "Image": Image, "transforms": transforms, "cv2": cv2,
          "alignCollate_syn": type("alignCollate_syn", (), {})}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in want:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "dataset.py:" + node.name, "exec"), ns)
    return ns["resizeNormalize"], ns["alignCollate_realWTLAMask"]


def main():
    RN, Collate = load_reference_classes()
    c = object.__new__(Collate)                       # alignCollate_syn.__init__ needs imgaug + al_chinese.txt: set its fields
    c.imgH, c.imgW, c.down_sample_scale, c.mask = 32, 128, 2, True
    c.d2a = "-" + ALPHABET
    c.alsize = len(c.d2a)
    c.a2d = {ch: i for i, ch in enumerate(c.d2a)}
    c.transform = RN((c.imgW, c.imgH), c.mask)
    c.transform2 = RN((c.imgW // 2, c.imgH // 2), c.mask, blur=True)
    hr, lr = images()
    batch = [(hr[i], lr[i], hr[i], lr[i], LABELS[i]) for i in range(len(LABELS))]
    out = c(batch)
    names = ["images_HR", "images_pseudoLR", "images_lr", "images_HRy", "images_lry", "label_strs", "label_rebatches",
             "weighted_masks", "weighted_tics"]
    fx = dict(zip(names, out))
    assert torch.equal(fx["images_HRy"], fx["images_HR"]) and torch.equal(fx["images_lry"], fx["images_lr"])
    del fx["images_HRy"], fx["images_lry"]            # the same PIL images were passed for the Y-domain slots
    fx["nomask_HR0"] = RN((128, 32), False)(hr[0])
    fx["torch"] = str(torch.__version__)
    import PIL
    fx["pil"] = PIL.__version__
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_collate_n5.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), {k: (tuple(v.shape) if hasattr(v, "shape") else v) for k, v in fx.items()})


if __name__ == "__main__":
    main()
