"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (tatt_b200/).

CPU restatement of the TextZoom collate: `resizeNormalize.__call__` (dataset/dataset.py:1276-1319) and the
label encoding of `alignCollate_realWTLAMask.__call__` (:2012-2076), in numpy / plain Python.
Pinned: tests/golden/data_collate_n5.pt holds the outputs of the reference's own classes (cut out of its source with
ast and executed unmodified by tests/golden/make_golden_data.py); tests/test_data.py checks this file against it
bit-exactly."""
import numpy as np
import torch
from PIL import Image


def resize_normalize(img, size, mask: bool):
    """dataset.py:1276-1319 (ratio_keep=False, aug=None): PIL bicubic resize to size=(W, H); ToTensor (HWC uint8 -> CHW
    float / 255); with mask: L = img.convert('L'), 0 where L > mean(L) else 255, appended as a 4th channel"""
    img = img.resize(size, Image.BICUBIC)
    a = np.asarray(img, dtype=np.uint8)
    t = torch.from_numpy(a.transpose(2, 0, 1).copy()).to(torch.float32).div(255)
    if mask:
        L = np.asarray(img.convert('L'))
        thres = L.mean()
        m = np.where(L > thres, 0, 255).astype(np.uint8)
        t = torch.cat((t, torch.from_numpy(m)[None].to(torch.float32).div(255)), 0)
    return t


def encode_labels(label_strs, alphabet: str = "0123456789abcdefghijklmnopqrstuvwxyz", max_len: int = 26):
    """dataset.py:2012-2076: lower-case, spread words shorter than 26 characters with '-' fillers (int((26 - len) /
    (len - 1)) per gap), truncate longer ones; one-hot rows over '-' + alphabet; an empty label is the blank class with
    tic 0.  -> (label_rebatches [N, alsize, 1, 26], weighted_masks [sum len], weighted_tics [N])"""
    d2a = "-" + alphabet
    a2d = {ch: i for i, ch in enumerate(d2a)}
    alsize = len(d2a)
    batches, masks, tics = [], [], []
    for word in label_strs:
        word = word.lower()
        if len(word) <= 1:
            pass
        elif len(word) < 26:
            padding = int((26 - len(word)) / (len(word) - 1))
            new_word = word[0]
            for i in range(len(word) - 1):
                new_word += "-" * padding + word[i + 1]
            word = new_word
        else:
            word = word[:26]
        label_list = [a2d[ch] for ch in word if ch in a2d]
        if len(label_list) <= 0:
            masks.append(0)
            vec = torch.zeros((1, alsize))
            vec[0, 0] = 1.
            tics.append(0)
        else:
            masks.extend(label_list)
            vec = torch.zeros((len(label_list), alsize))
            vec[torch.arange(len(label_list)), torch.tensor(label_list)] = 1.
            tics.append(1)
        batches.append(vec)
    re = torch.zeros((len(label_strs), max_len, alsize))
    for i, v in enumerate(batches):
        re[i][:v.shape[0]] = v
    return re.unsqueeze(1).float().permute(0, 3, 1, 2), torch.tensor(masks).long(), torch.tensor(tics)
