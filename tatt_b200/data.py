"""Wire formats either side of the hot path, host logic in Python like the reference's:

* checkpoints -- `TextBase.save_checkpoint` / the resume branch of `generator_init` (interfaces/base.py:616-660, 400-422):
  a `torch.save`d dict whose `'state_dict_G'` entry is the model's `state_dict()`, loaded with `strict=False`;
* TextZoom lmdb -- `lmdbDataset_real` (dataset/dataset.py:565-687): keys `num-samples`, `image_hr-%09d`, `image_lr-%09d`,
  `label-%09d` (1-based), values = encoded image files / utf-8 strings.  The `lmdb` package is not a dependency here:
  `LmdbReader` is a read-only parser of LMDB's on-disk B+tree (`data.mdb`), `write_lmdb` emits a minimal valid
  environment (used by the tests; real TextZoom files were not available to check against -- format parity UNPINNED);
* collate -- `resizeNormalize` + `alignCollate_realWTLAMask` (dataset/dataset.py:1266-1319, 1965-2076): the PIL bicubic
  resize stays on the host (same library as the reference => same pixels), the uint8 batch goes to the GPU once and
  `ToTensor` + the mean-threshold mask channel run there (csrc/data.cu, bit-exact); label one-hots are host tensors.
"""
from __future__ import annotations

import io
import os
import string
import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi, ops

Tensor = torch.Tensor
__all__ = ["save_checkpoint", "load_checkpoint", "LmdbReader", "write_lmdb", "TextZoomLmdb", "str_filt", "encode_labels",
           "collate_images", "AlignCollate"]


# ------------------------------------------------------------------------------------------------ checkpoints
def save_checkpoint(netG: torch.nn.Module, path: str, arch: str = "tatt", iters: int = 0, epoch: int = 0,
                    batch_size: int = 0, voc_type: str = "all", scale_factor: int = 2, best_acc_dict=None,
                    best_model_info=None, converge_list=None) -> Dict:
    """The dict of interfaces/base.py:636-645, written with torch.save.  `netG` may be wrapped (`.module`)."""
    net = netG.module if hasattr(netG, "module") else netG
    save_dict = {
        'state_dict_G': net.state_dict(),
        'info': {'arch': arch, 'iters': iters, 'epochs': epoch, 'batch_size': batch_size, 'voc_type': voc_type,
                 'up_scale_factor': scale_factor},
        'best_history_res': best_acc_dict,
        'best_model_info': best_model_info,
        'param_num': sum([param.nelement() for param in net.parameters()]),
        'converge': converge_list,
    }
    torch.save(save_dict, path)
    return save_dict


def load_checkpoint(model: torch.nn.Module, path: str, map_location="cpu"):
    """interfaces/base.py:404-422: `model.load_state_dict(torch.load(path)['state_dict_G'], strict=False)`.  Returns
    torch's (missing_keys, unexpected_keys) result.  Weights loaded into a live Trainer's model: call
    `ops.bump_weights_epoch()` semantics are handled here (cached positional encodings are invalidated)."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    sd = ck['state_dict_G'] if isinstance(ck, dict) and 'state_dict_G' in ck else ck
    res = model.load_state_dict(sd, strict=False)
    ops.bump_weights_epoch()
    return res


# ------------------------------------------------------------------------------------------------ LMDB (read-only)
_P_BRANCH, _P_LEAF, _P_OVERFLOW, _P_META = 0x01, 0x02, 0x04, 0x08
_F_BIGDATA = 0x01
_MAGIC = 0xBEEFC0DE
_INVALID = 0xFFFFFFFFFFFFFFFF


class LmdbReader:
    """Read-only view of an LMDB environment's main database (what `lmdb.open(root, readonly=True).begin().get(key)`
    gives the reference, dataset/dataset.py:576-593,639-646).  `root` is the environment directory (holding data.mdb) or
    the data file itself.  Layout parsed (LMDB 0.9 data version 1, little endian): 16-byte page headers, meta pages 0/1
    (the one with the larger txnid wins), branch / leaf pages with 8-byte node headers, overflow pages for big values."""

    def __init__(self, root: str):
        path = os.path.join(root, "data.mdb") if os.path.isdir(root) else root
        self._f = open(path, "rb")
        self._buf = np.memmap(path, dtype=np.uint8, mode="r")
        metas = []
        for pg in (0, 1):
            off = pg * self._guess_psize()
            flags = struct.unpack_from("<H", self._buf, off + 10)[0]
            magic, version = struct.unpack_from("<II", self._buf, off + 16)
            if not (flags & _P_META) or magic != _MAGIC:
                continue
            if version != 1:
                raise RuntimeError("LmdbReader: unsupported LMDB data version %d" % version)
            # MDB_meta: magic, version, address, mapsize, dbs[2] x 48 bytes, last_pg, txnid
            psize = struct.unpack_from("<I", self._buf, off + 16 + 24)[0]
            main = struct.unpack_from("<IHHQQQQQ", self._buf, off + 16 + 24 + 48)
            last_pg, txnid = struct.unpack_from("<QQ", self._buf, off + 16 + 24 + 96)
            metas.append((txnid, psize, main, last_pg))
        if not metas:
            raise RuntimeError("LmdbReader: %s is not an LMDB data file (no valid meta page)" % path)
        txnid, self.psize, main, self.last_pg = max(metas, key=lambda m: m[0])
        self.depth, self.entries, self.root = main[2], main[6], main[7]

    def _guess_psize(self) -> int:
        # mm_psize lives inside meta page 0 itself
        return struct.unpack_from("<I", self._buf, 16 + 24)[0] or 4096

    def close(self) -> None:
        del self._buf
        self._f.close()

    def __len__(self) -> int:
        return int(self.entries)

    def _page(self, pgno: int):
        off = pgno * self.psize
        flags, lower = struct.unpack_from("<HH", self._buf, off + 10)
        return off, flags, (lower - 16) // 2

    def _node(self, off: int, i: int):
        p = off + struct.unpack_from("<H", self._buf, off + 16 + 2 * i)[0]
        lo, hi, fl, ks = struct.unpack_from("<HHHH", self._buf, p)
        return p, lo, hi, fl, ks, bytes(self._buf[p + 8:p + 8 + ks])

    def _value(self, p: int, lo: int, hi: int, fl: int, ks: int) -> bytes:
        size = lo | (hi << 16)
        d = p + 8 + ks
        if fl & _F_BIGDATA:
            ov = struct.unpack_from("<Q", self._buf, d)[0]
            start = ov * self.psize + 16
            return bytes(self._buf[start:start + size])
        return bytes(self._buf[d:d + size])

    def get(self, key: bytes) -> Optional[bytes]:
        if self.root == _INVALID:
            return None
        pg = self.root
        while True:
            off, flags, n = self._page(pg)
            if flags & _P_BRANCH:
                lo_i, hi_i = 0, n - 1                      # last node whose key <= target (node 0's key is -infinity)
                while lo_i < hi_i:
                    mid = (lo_i + hi_i + 1) // 2
                    if self._node(off, mid)[5] <= key:
                        lo_i = mid
                    else:
                        hi_i = mid - 1
                p, lo, hi, fl, ks, _ = self._node(off, lo_i)
                pg = lo | (hi << 16) | (fl << 32)
            elif flags & _P_LEAF:
                a, b = 0, n - 1
                while a <= b:
                    mid = (a + b) // 2
                    p, lo, hi, fl, ks, k = self._node(off, mid)
                    if k == key:
                        return self._value(p, lo, hi, fl, ks)
                    if k < key:
                        a = mid + 1
                    else:
                        b = mid - 1
                return None
            else:
                raise RuntimeError("LmdbReader: unexpected page flags 0x%x at page %d" % (flags, pg))

    def items(self) -> Iterator[Tuple[bytes, bytes]]:
        def walk(pg):
            off, flags, n = self._page(pg)
            for i in range(n):
                p, lo, hi, fl, ks, k = self._node(off, i)
                if flags & _P_BRANCH:
                    yield from walk(lo | (hi << 16) | (fl << 32))
                else:
                    yield k, self._value(p, lo, hi, fl, ks)
        if self.root != _INVALID:
            yield from walk(self.root)


def write_lmdb(root: str, items: Dict[bytes, bytes], psize: int = 4096) -> str:
    """Write `items` as a fresh single-transaction LMDB environment (`root/data.mdb`): sorted keys, leaf pages filled in
    order, values that do not fit a node (> ~psize/2) on overflow pages, branch levels up to one root.  Test fixture
    writer; the dataset itself is read-only in the reference."""
    os.makedirs(root, exist_ok=True)
    keys = sorted(items)
    nodemax = (((psize - 16) // 2) & ~1) - 2
    pages: List[bytearray] = [bytearray(psize), bytearray(psize)]
    n_branch = n_leaf = n_over = 0

    def new_page(flags: int) -> Tuple[int, bytearray]:
        pg = bytearray(psize)
        pgno = len(pages)
        struct.pack_into("<QHHHH", pg, 0, pgno, 0, flags, 16, psize)
        pages.append(pg)
        return pgno, pg

    def add_node(pg: bytearray, lo: int, hi: int, fl: int, key: bytes, data: bytes) -> bool:
        lower, upper = struct.unpack_from("<HH", pg, 12)
        size = 8 + len(key) + len(data)
        size += size & 1
        if upper - size < lower + 2:
            return False
        upper -= size
        struct.pack_into("<HHHH", pg, upper, lo, hi, fl, len(key))
        pg[upper + 8:upper + 8 + len(key)] = key
        pg[upper + 8 + len(key):upper + 8 + len(key) + len(data)] = data
        struct.pack_into("<H", pg, lower, upper)
        struct.pack_into("<HH", pg, 12, lower + 2, upper)
        return True

    level: List[Tuple[bytes, int]] = []              # (first key, pgno) of the pages of the current level
    cur = None
    for k in keys:
        v = items[k]
        if 8 + len(k) + len(v) > nodemax:
            npg = (16 + len(v) + psize - 1) // psize
            first = len(pages)
            blob = bytearray(npg * psize)
            struct.pack_into("<QHHI", blob, 0, first, 0, _P_OVERFLOW, npg)
            blob[16:16 + len(v)] = v
            for i in range(npg):
                pages.append(blob[i * psize:(i + 1) * psize])
            n_over += npg
            node = (len(v) & 0xFFFF, len(v) >> 16, _F_BIGDATA, k, struct.pack("<Q", first))
        else:
            node = (len(v) & 0xFFFF, len(v) >> 16, 0, k, v)
        if cur is None or not add_node(cur[1], *node):
            cur = new_page(_P_LEAF)
            n_leaf += 1
            level.append((k, cur[0]))
            assert add_node(cur[1], *node)
    depth = 1 if level else 0
    while len(level) > 1:
        nxt: List[Tuple[bytes, int]] = []
        cur = None
        for k, pgno in level:
            first_on_page = cur is None
            key = b"" if first_on_page else k
            node = (pgno & 0xFFFF, (pgno >> 16) & 0xFFFF, (pgno >> 32) & 0xFFFF, key, b"")
            if cur is None or not add_node(cur[1], *node):
                cur = new_page(_P_BRANCH)
                n_branch += 1
                nxt.append((k, cur[0]))
                assert add_node(cur[1], pgno & 0xFFFF, (pgno >> 16) & 0xFFFF, (pgno >> 32) & 0xFFFF, b"", b"")
        level = nxt
        depth += 1
    root_pg = level[0][1] if level else _INVALID
    last_pg = len(pages) - 1
    for i in (0, 1):
        pg = pages[i]
        struct.pack_into("<QHHHH", pg, 0, i, 0, _P_META, 0, 0)
        struct.pack_into("<IIQQ", pg, 16, _MAGIC, 1, 0, psize * max(len(pages), 16))
        struct.pack_into("<IHHQQQQQ", pg, 16 + 24, psize, 0, 0, 0, 0, 0, 0, _INVALID)           # free DB (empty)
        struct.pack_into("<IHHQQQQQ", pg, 16 + 24 + 48, 0, 0, depth, n_branch, n_leaf, n_over, len(keys), root_pg)
        struct.pack_into("<QQ", pg, 16 + 24 + 96, last_pg, i)                                   # txnid 0 / 1
    path = os.path.join(root, "data.mdb")
    with open(path, "wb") as f:
        for pg in pages:
            f.write(bytes(pg))
    return path


# ------------------------------------------------------------------------------------------------ dataset
def str_filt(str_: str, voc_type: str) -> str:
    """utils/util.py:12-32 (the non-Chinese vocabularies)"""
    alpha_dict = {'digit': string.digits, 'lower': string.digits + string.ascii_lowercase,
                  'upper': string.digits + string.ascii_letters,
                  'all': string.digits + string.ascii_letters + string.punctuation}
    if voc_type not in alpha_dict:
        raise NotImplementedError("str_filt: vocabulary %r needs the reference's al_chinese.txt" % voc_type)
    if voc_type == 'lower':
        str_ = str_.lower()
    for char in str_:
        if char not in alpha_dict[voc_type]:
            str_ = str_.replace(char, '')
    return str_


class TextZoomLmdb(torch.utils.data.Dataset):
    """`lmdbDataset_real` (dataset/dataset.py:565-687) without the optional augmentations: item i (0-based) ->
    (img_HR, img_lr, img_HRy, img_lry, label_str) with PIL images; the Y-domain pair is the RGB->YUV conversion of the
    same crops (cv2.COLOR_RGB2YUV like the reference; needs cv2)."""

    def __init__(self, root: str, voc_type: str = 'upper', max_len: int = 100, test: bool = False):
        self.env = LmdbReader(root)
        n = self.env.get(b'num-samples')
        if n is None:
            raise RuntimeError("TextZoomLmdb: %s has no 'num-samples' key" % root)
        self.nSamples = int(n)
        self.voc_type, self.max_len, self.test = voc_type, max_len, test

    def __len__(self):
        return self.nSamples

    def _img(self, key: bytes):
        from PIL import Image
        buf = self.env.get(key)
        if buf is None:
            raise IOError("missing key %r" % key)
        return Image.open(io.BytesIO(buf)).convert('RGB')

    def __getitem__(self, index):
        assert index <= len(self), 'index range error'
        index += 1
        from PIL import Image
        try:
            img_HR = self._img(b'image_hr-%09d' % index)
            img_lr = self._img(b'image_lr-%09d' % index)
            import cv2
            img_lry = Image.fromarray(cv2.cvtColor(np.array(img_lr).astype(np.uint8), cv2.COLOR_RGB2YUV))
            img_HRy = Image.fromarray(cv2.cvtColor(np.array(img_HR).astype(np.uint8), cv2.COLOR_RGB2YUV))
            word = self.env.get(b'label-%09d' % index)
            word = " " if word is None else str(word.decode())
        except IOError:
            return self[index + 1]
        return img_HR, img_lr, img_HRy, img_lry, str_filt(word, self.voc_type)


# ------------------------------------------------------------------------------------------------ collate
def encode_labels(label_strs: Sequence[str], alphabet: str = "0123456789abcdefghijklmnopqrstuvwxyz", max_len: int = 26):
    """Label tensors of `alignCollate_realWTLAMask.__call__` (dataset/dataset.py:2012-2076) ->
    (label_rebatches [N, 1 + len(alphabet), 1, 26] one-hot, weighted_masks [sum of label lengths] long, weighted_tics [N])"""
    d2a = "-" + alphabet
    a2d = {ch: i for i, ch in enumerate(d2a)}
    alsize = len(d2a)
    rebatch = torch.zeros((len(label_strs), max_len, alsize))
    masks: List[int] = []
    tics: List[int] = []
    for idx, word in enumerate(label_strs):
        word = word.lower()
        if 1 < len(word) < 26:
            gap = "-" * int((26 - len(word)) / (len(word) - 1))
            word = gap.join(word)
        elif len(word) >= 26:
            word = word[:26]
        ids = [a2d[ch] for ch in word if ch in a2d]
        if ids:
            masks.extend(ids)
            rebatch[idx, torch.arange(len(ids)), torch.tensor(ids)] = 1.
            tics.append(1)
        else:                                               # blank label
            masks.append(0)
            rebatch[idx, 0, 0] = 1.
            tics.append(0)
    return rebatch.unsqueeze(1).float().permute(0, 3, 1, 2), torch.tensor(masks).long(), torch.tensor(tics)


def collate_images(images, size: Tuple[int, int], mask: bool, device=None, pinned: Optional[Tensor] = None) -> Tensor:
    """`[resizeNormalize(size, mask)(img) for img in images]` stacked (dataset/dataset.py:1266-1319, 1993-1997), as a
    CUDA tensor [N, 3 + mask, H, W]: PIL bicubic resize on the host, ONE uint8 host->device copy, normalisation + mask
    channel on the device (tatt_collate_u8).  size = (W, H) like the reference."""
    from PIL import Image
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("tatt_b200: collate_images produces CUDA tensors only (no CPU fallback)")
    w, h = size
    n = len(images)
    host = pinned if pinned is not None else torch.empty((n, h, w, 3), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for i, img in enumerate(images):
        hv[i] = np.asarray(img.convert('RGB').resize(size, Image.BICUBIC), dtype=np.uint8)
    dev = host.to(device, non_blocking=True)
    out = torch.empty((n, 4 if mask else 3, h, w), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _cabi.call("tatt_collate_u8", dev.data_ptr(), out.data_ptr(), n, h, w, 1 if mask else 0, ops._stream())
    return out


class AlignCollate:
    """`alignCollate_realWTLAMask` (dataset/dataset.py:1965-2076; constructor fields of `alignCollate_syn` :1624-1675):
    `__call__(batch)` with batch items `(img_HR, img_lr, img_HRy, img_lry, label_str)` returns the reference's 9-tuple
    `(images_HR, None, images_lr, images_HRy, images_lry, label_strs, label_rebatches, weighted_masks, weighted_tics)`;
    the four image tensors are CUDA tensors.  `alphabet`: the reference reads `al_chinese.txt`; pass its first line to
    reproduce that, the default is the 36-character alphabet of its English variants (:1836)."""

    def __init__(self, imgH=64, imgW=256, down_sample_scale=4, keep_ratio=False, min_ratio=1, mask=False,
                 alphabet="0123456789abcdefghijklmnopqrstuvwxyz", train=True, device=None):
        self.imgH, self.imgW, self.down_sample_scale, self.mask = imgH, imgW, down_sample_scale, mask
        self.keep_ratio, self.min_ratio, self.train = keep_ratio, min_ratio, train
        self.alphabet = alphabet
        self.d2a = "-" + alphabet
        self.alsize = len(self.d2a)
        self.device = device

    def __call__(self, batch):
        images_HR, images_lr, images_HRy, images_lry, label_strs = zip(*batch)
        hr_size = (self.imgW, self.imgH)
        lr_size = (self.imgW // self.down_sample_scale, self.imgH // self.down_sample_scale)
        t_HR = collate_images(images_HR, hr_size, self.mask, self.device)
        t_lr = collate_images(images_lr, lr_size, self.mask, self.device)
        t_HRy = collate_images(images_HRy, hr_size, self.mask, self.device)
        t_lry = collate_images(images_lry, lr_size, self.mask, self.device)
        rebatch, wmask, wtics = encode_labels(label_strs, self.alphabet)
        return t_HR, None, t_lr, t_HRy, t_lry, label_strs, rebatch, wmask, wtics
