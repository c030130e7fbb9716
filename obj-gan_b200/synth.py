"""Seeded synthetic COCO-shaped inputs for the image_generation training step.

Shapes follow what the reference's data pipeline hands to ``condGANTrainer.train``
(reference: image_generation/trainDataset.py:79-128 ``prepare_data``,
image_generation/trainer.py:357-393) with the dataset replaced by random draws, as
SURVEY.md section 8(d) specifies.  Everything is generated on the CPU with an explicit
``torch.Generator`` so the same call gives the same tensors here and on the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch

from .config import cfg

IMG_SIZES = (64, 128, 256)


def _ellipse_masks(boxes: torch.Tensor, num_rois: torch.Tensor, size: int, scale: float) -> torch.Tensor:
    """Filled axis-aligned ellipse inside each box, 1.0 inside with a one-pixel soft edge.

    boxes: (B, R, 4) float64 [x, y, w, h] at 64-pixel scale; returns (B, R, size, size) float32.
    """
    B, R, _ = boxes.shape
    ys = torch.arange(size, dtype=torch.float64).view(1, 1, size, 1) + 0.5
    xs = torch.arange(size, dtype=torch.float64).view(1, 1, 1, size) + 0.5
    x = (boxes[..., 0] * scale).view(B, R, 1, 1)
    y = (boxes[..., 1] * scale).view(B, R, 1, 1)
    w = (boxes[..., 2] * scale).clamp(min=1.0).view(B, R, 1, 1)
    h = (boxes[..., 3] * scale).clamp(min=1.0).view(B, R, 1, 1)
    cx, cy, rx, ry = x + w / 2, y + h / 2, w / 2, h / 2
    # signed "distance" in pixels along the ellipse normal (approximate), soft edge of 1 px
    d = (((xs - cx) / rx) ** 2 + ((ys - cy) / ry) ** 2).sqrt()
    edge = (1.0 - d) * torch.minimum(rx, ry)
    m = edge.clamp(0.0, 1.0)
    valid = (torch.arange(R).view(1, R) < num_rois.view(B, 1)).view(B, R, 1, 1)
    return (m * valid).float()


def make_inputs(batch: int, seed: int = 1234, *, words: int | None = None, max_rois: int | None = None,
                parity: bool = False, num_classes: int = 80) -> dict:
    """Build one batch of step inputs (all CPU tensors).

    parity=False: throughput workload -- every caption has ``words`` tokens (mask all False) and
    every image has ``max_rois`` boxes.  parity=True: ragged caption lengths (sorted descending,
    as ``prepare_data`` does) and 1..max_rois boxes per image, so the mask quirk and the padded
    roi slots are exercised.
    """
    L = int(words if words is not None else cfg.TEXT.WORDS_NUM)
    RB = cfg.ROI.BOXES_NUM
    Rm = int(max_rois if max_rois is not None else RB)
    g = torch.Generator().manual_seed(seed)
    B = batch

    def randn(*s):
        return torch.randn(*s, generator=g)

    def rand(*s):
        return torch.rand(*s, generator=g)

    z = randn(B, cfg.GAN.Z_DIM)
    eps = randn(B, cfg.GAN.CONDITION_DIM)  # CA_NET reparametrisation noise, injected for parity
    sent_emb = rand(B, cfg.TEXT.EMBEDDING_DIM) * 2 - 1
    words_embs = rand(B, cfg.TEXT.EMBEDDING_DIM, L) * 2 - 1
    glove_words_embs = 0.5 * randn(B, cfg.TEXT.GLOVE_EMBEDDING_DIM, L)
    clabels_emb = 0.5 * randn(num_classes, cfg.TEXT.GLOVE_EMBEDDING_DIM)

    if parity:
        cap_lens = torch.randint(5, L + 1, (B,), generator=g)
        cap_lens[0] = L  # the text encoder emits max(cap_lens) words; keep that equal to L
        cap_lens = torch.sort(cap_lens, descending=True).values
        num_rois = torch.randint(1, Rm + 1, (B,), generator=g)
        num_rois[B // 2] = Rm
    else:
        cap_lens = torch.full((B,), L, dtype=torch.long)
        num_rois = torch.full((B,), Rm, dtype=torch.long)
    mask = torch.arange(L).view(1, L) >= cap_lens.view(B, 1)  # (B, L) bool, True = padding

    xy = rand(B, RB, 2).double() * 40.0
    wh = 6.0 + rand(B, RB, 2).double() * 18.0
    wh = torch.minimum(wh, 64.0 - xy)
    cls = torch.randint(0, num_classes, (B, RB), generator=g).double()
    rois0 = torch.cat([xy, wh, cls.unsqueeze(-1), torch.zeros(B, RB, 1, dtype=torch.float64)], dim=2)
    valid = (torch.arange(RB).view(1, RB) < num_rois.view(B, 1)).unsqueeze(-1)
    rois0 = rois0 * valid
    rois = []
    for i in range(3):
        r = rois0.clone()
        r[..., :4] *= 2.0 ** i
        rois.append(r)
    fm_rois = rois0.clone()
    fm_rois[..., :4] /= 2.0

    boxes = rois0[..., :4]
    bt_masks = [_ellipse_masks(boxes, num_rois, s, s / 64.0) for s in IMG_SIZES]
    fm_bt_masks = _ellipse_masks(boxes, num_rois, 32, 0.5)
    hmaps = []
    cls_idx = cls.long()
    for i, s in enumerate(IMG_SIZES):
        hm = torch.zeros(B, num_classes, s, s)
        hm.scatter_add_(1, cls_idx.view(B, RB, 1, 1).expand(B, RB, s, s), bt_masks[i])
        hmaps.append(hm.clamp_(max=1.0))

    # slabels_feat: class-label GloVe vectors per roi, (B, 50, Rmax, 1) -- reference
    # image_generation/miscc/utils.py:502-522 (form_clabels_feat)
    rmax = int(num_rois.max())
    slabels = torch.zeros(B, rmax, clabels_emb.shape[1])
    for b in range(B):
        n = int(num_rois[b])
        slabels[b, :n] = clabels_emb[cls_idx[b, :n]]
    slabels_feat = slabels.transpose(1, 2).unsqueeze(3).contiguous()

    imgs = [rand(B, 3, s, s) * 2 - 1 for s in IMG_SIZES]

    return dict(z=z, eps=eps, sent_emb=sent_emb, words_embs=words_embs, glove_words_embs=glove_words_embs,
                clabels_emb=clabels_emb, slabels_feat=slabels_feat, cap_lens=cap_lens, mask=mask,
                num_rois=num_rois, rois=rois, fm_rois=fm_rois, bt_masks=bt_masks, fm_bt_masks=fm_bt_masks,
                hmaps=hmaps, imgs=imgs, glb_max_num_roi=rmax)


HMAP_CLAMP = 1.0     # make_inputs clamps its synthetic class heat maps at 1 (the reference's loader does not clamp)


def compact(inp: dict) -> dict:
    """The batch as it needs to cross PCIe: without the two tensors the device rebuilds from the rest
    (``hmaps`` = per-class sums of ``bt_masks``, 86 % of the bytes; ``slabels_feat`` = label embeddings of the roi
    classes), plus the roi class ids as an int64 table (column 4 of ``rois[0]``)."""
    out = {k: v for k, v in inp.items() if k not in ("hmaps", "slabels_feat")}
    out["roi_cls"] = inp["rois"][0][..., 4].to(torch.int64).contiguous()
    return out


def input_bytes(inp: dict) -> int:
    """Bytes a step copies host->device (every tensor in the batch)."""
    n = 0
    for v in inp.values():
        if torch.is_tensor(v):
            n += v.numel() * v.element_size()
        elif isinstance(v, (list, tuple)):
            n += sum(t.numel() * t.element_size() for t in v if torch.is_tensor(t))
    return n


def shard(inp: dict, rank: int, world: int) -> dict:
    """Data-parallel shard: samples [rank*B/world, (rank+1)*B/world) of every per-sample tensor
    (what ``nn.DataParallel``'s scatter does on dim 0; reference trainer.py:136-152)."""
    B = inp["z"].shape[0]
    assert B % world == 0
    lo, hi = rank * B // world, (rank + 1) * B // world
    out = {}
    for k, v in inp.items():
        if k == "clabels_emb" or k == "glb_max_num_roi":
            out[k] = v
        elif torch.is_tensor(v):
            out[k] = v[lo:hi].contiguous()
        else:
            out[k] = [t[lo:hi].contiguous() for t in v]
    rmax = int(out["num_rois"].max())
    out["glb_max_num_roi"] = rmax
    out["slabels_feat"] = out["slabels_feat"][:, :, :rmax].contiguous()
    return out
