"""Loss functions of the hot path with the reference's names and argument order
(reference: image_generation/miscc/losses.py:163-211 ``patD_loss``, 364-399 patch-D part of ``G_loss``,
533-537 ``KL_loss``).  BCE / KL values and their gradients come from the fused kernels ``og_bce`` / ``og_kl``.
"""
from __future__ import annotations

import torch

from . import ops
from .config import cfg


def _heads(net):
    net = net.module if hasattr(net, "module") else net
    return net.COND_DNET, net.UNCOND_DNET


def patD_loss(netPatD, real_imgs, fake_imgs, conditions):
    """ref: miscc/losses.py:163-211 -- two body forwards (separate BatchNorm statistics), COND head on
    (real, fake, wrong = real[:B-1] x cond[1:]), UNCOND head on (real, fake)."""
    real_features = netPatD(real_imgs)
    fake_features = netPatD(fake_imgs.detach())
    cond, uncond = _heads(netPatD)
    ul, tl = cfg.TRAIN.SMOOTH.UNCOND_LAMBDA, cfg.TRAIN.SMOOTH.TXT_LAMBDA
    batch_size = real_features.size(0)
    cond_real = cond(real_features, conditions)
    cond_fake = cond(fake_features, conditions)
    cond_wrong = cond(real_features[:(batch_size - 1)], conditions[1:batch_size])
    if uncond is not None:
        real_logits = uncond(real_features)
        fake_logits = uncond(fake_features)
        # errD = (u_real*UL + c_real*TL)/2 + (u_fake*UL + (c_fake + c_wrong)*TL)/3, folded into the weights
        return (ops.bce(real_logits, 1.0, ul / 2.0) + ops.bce(cond_real, 1.0, tl / 2.0)
                + ops.bce(fake_logits, 0.0, ul / 3.0) + ops.bce(cond_fake, 0.0, tl / 3.0)
                + ops.bce(cond_wrong, 0.0, tl / 3.0))
    return ops.bce(cond_real, 1.0, tl) + ops.bce(cond_fake, 0.0, tl / 2.0) + ops.bce(cond_wrong, 0.0, tl / 2.0)


def G_loss_pat(netsPatD, fake_imgs, sent_emb, streams=None):
    """Patch-discriminator terms of ``G_loss`` (ref: miscc/losses.py:372-399).  ``streams``: optional CUDA streams, one
    per discriminator: the three branches are independent until their losses are summed, so they may run
    concurrently (autograd replays each branch's backward on the stream of its forward)."""
    import contextlib
    ul, tl = cfg.TRAIN.SMOOTH.UNCOND_LAMBDA, cfg.TRAIN.SMOOTH.TXT_LAMBDA
    per = []
    main = torch.cuda.current_stream() if streams else None
    for i in range(len(netsPatD)):
        if streams:
            streams[i].wait_stream(main)
        with (torch.cuda.stream(streams[i]) if streams else contextlib.nullcontext()):
            features = netsPatD[i](fake_imgs[i])
            cond, uncond = _heads(netsPatD[i])
            cond_err = ops.bce(cond(features, sent_emb), 1.0, tl if uncond is not None else 1.0)
            loss = cond_err
            if uncond is not None:
                loss = loss + ops.bce(uncond(features), 1.0, ul)
            per.append(loss)
    total = None
    for i, loss in enumerate(per):
        if streams:
            main.wait_stream(streams[i])
        total = loss if total is None else total + loss
    return total, per


def KL_loss(mu, logvar):
    """ref: miscc/losses.py:533-537.  ``mu`` / ``logvar`` are the column views G_NET returns; the fused
    kernel runs on their shared row buffer."""
    base = mu._base if mu._base is not None else None
    if base is not None and logvar._base is base and base.dim() == 2:
        return ops.kl_rows(base, mu.shape[1], 1.0)
    d = mu.shape[1]
    rows = torch.zeros(mu.shape[0], ops.cpad(2 * d), device=mu.device)
    rows = torch.cat((mu, logvar, rows[:, 2 * d:]), 1)
    return ops.kl_rows(rows, d, 1.0)


def permute_seg(seg_conditions, rois, num_rois):
    """ref: miscc/utils.py:445-462 -- per sample, shuffle the segmentation channels of the classes present (host
    `random.shuffle`, drawn in the reference's order); returns the permuted maps and the indices of the samples that
    changed.  The shuffles become one (B, C) channel-permutation table, applied by ONE kernel pass
    (``og_permute_channels``) instead of a clone plus two indexed copies per sample."""
    import random
    from copy import deepcopy

    import numpy as np
    rois_np = rois.detach().cpu().numpy() if torch.is_tensor(rois) else np.asarray(rois)
    nums = num_rois.detach().cpu().numpy().tolist() if torch.is_tensor(num_rois) else list(num_rois)
    bsz, nch = seg_conditions.size(0), seg_conditions.size(1)
    perm = np.tile(np.arange(nch, dtype=np.int64), (bsz, 1))
    valid = []
    for b in range(bsz):
        if nums[b] == 0:
            continue
        classes = list(np.unique(rois_np[b, :nums[b], 4]).astype(int))
        shuffled = deepcopy(classes)
        random.shuffle(shuffled)
        if classes != shuffled:
            valid.append(b)
            perm[b, classes] = shuffled                 # new_seg[b, classes] = seg[b, shuffled]
    if not valid:
        return seg_conditions, valid                    # nothing changed (the reference returns an equal clone)
    return ops.permute_channels(seg_conditions, ops.h2d(perm, seg_conditions.device)), valid


def shpD_loss(netShpD, real_imgs, fake_imgs, seg_conditions, rois, num_rois):
    """ref: miscc/losses.py:213-251."""
    real_features = netShpD(real_imgs, seg_conditions)
    fake_features = netShpD(fake_imgs.detach(), seg_conditions)
    fake_seg, valid = permute_seg(seg_conditions, rois, num_rois)
    net = netShpD.module if hasattr(netShpD, "module") else netShpD
    err = ops.bce(net.UNCOND_DNET(real_features), 1.0, 1.0)
    fake_err = ops.bce(net.UNCOND_DNET(fake_features), 0.0, 1.0)
    if len(valid) > 0:
        idx = ops.h2d(valid, real_imgs.device, torch.int64)
        wrong = netShpD(real_imgs[idx], fake_seg[idx])
        return err + (fake_err + ops.bce(net.UNCOND_DNET(wrong), 0.0, 1.0)) / 2.0
    return err + fake_err


def feat_select(pooled_feat, raw_bt_c_codes, fm_rois, num_rois, is_large_scale=False):
    """ref: miscc/utils.py:465-499 -- keep the rois of the requested scale class.  The box filter runs on the host
    copy of ``fm_rois`` like the reference; the kept rows are compacted on the device by ONE gather per tensor
    (the reference concatenates per-sample slices).  Returns (features (R', C, h, w), classes (host int array),
    bt_c_codes (R', D)); three empty lists when nothing is kept."""
    import numpy as np
    fm = fm_rois.detach().cpu().numpy() if torch.is_tensor(fm_rois) else np.asarray(fm_rois)
    nums = num_rois.detach().cpu().numpy().tolist() if torch.is_tensor(num_rois) else list(num_rois)
    boxes, code_slots = pooled_feat.shape[1], raw_bt_c_codes.shape[1]
    flat, flat_c, classes = [], [], []
    for b in range(len(nums)):
        for r in range(int(nums[b])):
            _, _, width, height = fm[b, r, :4]
            if width < 1.25 and height < 1.25:
                continue
            big = max(width, height) >= cfg.ROI.ROI_SIZE_THRS
            if big != bool(is_large_scale):
                continue
            flat.append(b * boxes + r)
            flat_c.append(b * code_slots + r)      # the generator's bt_c_code has max(num_rois) slots, not BOXES_NUM
            classes.append(int(fm[b, r, 4]))
    if not flat:
        return [], [], []
    idx = ops.h2d(flat, pooled_feat.device, torch.int64)
    idx_c = ops.h2d(flat_c, pooled_feat.device, torch.int64)
    feats = ops.gather_rows(pooled_feat.reshape((-1,) + tuple(pooled_feat.shape[2:])), idx)
    codes = ops.gather_rows(raw_bt_c_codes.reshape(-1, raw_bt_c_codes.shape[-1]), idx_c)
    return feats, np.asarray(classes, dtype=np.int64), codes


def objD_loss(netObjD, real_imgs, fake_imgs, seg_conditions, raw_conditions, raw_bt_c_codes, fm_rois, num_rois,
              is_large_scale=False):
    """ref: miscc/losses.py:254-361 -- object discriminator loss: real / fake / mismatched-condition / permuted-shape
    terms over the rois of one scale class.  Reproduces the reference's use of ``classes`` (not ``classes2``) for the
    permuted-shape conditions (losses.py:307-311)."""
    net = netObjD.module if hasattr(netObjD, "module") else netObjD
    dev = real_imgs.device

    def lookup(cls_idx, codes):
        idx = ops.h2d([int(c) for c in cls_idx], dev, torch.int64)
        return ops.cat_rows_const(ops.gather_rows(raw_conditions.detach(), idx), codes)

    # the shape branch of the net sees the same segmentation map in the real and the fake pass: evaluate it once
    shared = {"shape_features": net.shape_features(seg_conditions)} if hasattr(net, "shape_features") else {}
    real_pooled = netObjD(real_imgs, seg_conditions, fm_rois, num_rois, **shared)
    real_features, classes, bt_c_codes = feat_select(real_pooled, raw_bt_c_codes, fm_rois, num_rois, is_large_scale)
    fake_pooled = netObjD(fake_imgs.detach(), seg_conditions, fm_rois, num_rois, **shared)
    fake_features, _, _ = feat_select(fake_pooled, raw_bt_c_codes, fm_rois, num_rois, is_large_scale)
    fake_seg, valid = permute_seg(seg_conditions, fm_rois, num_rois)
    classes2 = []
    if len(valid) > 0:
        vi = ops.h2d(valid, dev, torch.int64)
        fm_t = fm_rois if torch.is_tensor(fm_rois) else torch.as_tensor(fm_rois)
        nr_t = num_rois if torch.is_tensor(num_rois) else torch.as_tensor(num_rois)
        # index the (host or device) box tables with an index on THEIR device: no device -> host read of `valid`
        sel = lambda t: t[ops.h2d(valid, t.device, torch.int64)]
        pooled2 = netObjD(real_imgs[vi], fake_seg[vi], sel(fm_t), sel(nr_t))
        fake_features2, classes2, bt_c_codes2 = feat_select(pooled2, raw_bt_c_codes, sel(fm_t), sel(nr_t),
                                                            is_large_scale)
    n = len(classes)
    if n == 0:
        return 0
    conditions = lookup(classes, bt_c_codes)
    cond_real = ops.bce(net.COND_DNET(real_features, conditions), 1.0, 1.0)
    cond_fake = ops.bce(net.COND_DNET(fake_features, conditions), 0.0, 1.0)
    cond_wrong = None
    if n > 1:
        cond_wrong = ops.bce(net.COND_DNET(real_features[:n - 1], conditions[1:n]), 0.0, 1.0)
    cond_wrong2 = None
    if len(valid) > 0 and len(classes2) > 0:
        conditions2 = lookup(classes[:len(classes2)], bt_c_codes2)
        cond_wrong2 = ops.bce(net.COND_DNET(fake_features2, conditions2), 0.0, 1.0)
    if net.UNCOND_DNET is not None:
        err = (ops.bce(net.UNCOND_DNET(real_features), 1.0, 1.0) + cond_real) / 2.0
        tmp = ops.bce(net.UNCOND_DNET(fake_features), 0.0, 1.0) + cond_fake
        denorm = 3.0
    else:
        err = cond_real
        tmp = cond_fake
        denorm = 2.0
    if cond_wrong is not None:
        tmp = tmp + cond_wrong
    if cond_wrong2 is not None:
        tmp = tmp + cond_wrong2
        denorm += 1.0
    return err + tmp / denorm


def _class_masks(class_ids, batch_size, device):
    """ref: miscc/losses.py:27-38 -- mask[i, j] = 1 where caption j comes from the same class as sample i (j != i)."""
    if class_ids is None:
        return None
    import numpy as np
    ids = np.asarray(class_ids)
    m = (ids.reshape(-1, 1) == ids.reshape(1, -1)).astype(np.uint8)
    np.fill_diagonal(m, 0)
    return ops.h2d(m[:batch_size, :batch_size].copy(), device)


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, top1=True, is_training=True):
    """ref: miscc/losses.py:22-71 (DAMSM sentence matching): scaled cosine matrix between the image codes and the
    sentence codes, two cross-entropy directions.  ``rnn_code`` is a constant here (trainer.py:369 detaches it).
    Returns (loss0, loss1, accuracy) with the accuracy a device scalar (percent)."""
    if cnn_code.dim() == 3:
        cnn_code, rnn_code = cnn_code[0], rnn_code[0]
    sim = ops.cosine_matrix(cnn_code, rnn_code, eps)
    mask = _class_masks(class_ids, batch_size, cnn_code.device)
    loss0, loss1, correct = ops.ce_pair(sim, mask, labels, cfg.TRAIN.SMOOTH.GAMMA3)
    accuracy = correct * (100.0 / (batch_size * 2.0)) if top1 else sim.detach() * cfg.TRAIN.SMOOTH.GAMMA3
    return loss0, loss1, accuracy


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size, top1=True, is_training=True,
               need_att_maps=True):
    """ref: miscc/losses.py:74-159 (DAMSM word-region matching).  img_features (B, nef, 17, 17) carries the gradient;
    words_emb (B, nef, T) is a constant.  All B x B (image, caption) pairs run in ONE launch (``ops.words_pairs``:
    func_attention of caption i's words against image b, cosine similarity word / attended context, Eq. (10) pooling)
    instead of the reference's per-caption Python loop; then the two cross-entropy directions over the B x B
    similarity matrix.  ``cap_lens``: int64 tensor on the device (no host synchronisation), or a host list / tensor.
    Returns (loss0, loss1, att_maps, accuracy); att_maps[i] = (1, n_i, 17, 17) like the reference when
    ``need_att_maps`` (slicing by n_i needs the lengths on the host), else the padded (B, T, 17, 17) tensor."""
    dev = img_features.device
    if torch.is_tensor(cap_lens) and cap_lens.device == dev:
        lens_dev = cap_lens.to(torch.int64).contiguous()
    else:
        lens_dev = ops.h2d([int(v) for v in cap_lens], dev, torch.int64)
    sim, attn = ops.words_pairs(img_features, words_emb.detach(), lens_dev[:batch_size], cfg.TRAIN.SMOOTH.GAMMA1,
                                cfg.TRAIN.SMOOTH.GAMMA2)                     # sim[image, caption]
    diag = attn[torch.arange(batch_size, device=dev), torch.arange(batch_size, device=dev)] if need_att_maps else None
    if need_att_maps:
        lens = cap_lens.detach().cpu().tolist() if torch.is_tensor(cap_lens) else list(cap_lens)
        att_maps = [diag[i:i + 1, :int(lens[i])].contiguous() for i in range(batch_size)]
    else:
        att_maps = None
    mask = _class_masks(class_ids, batch_size, dev)
    loss0, loss1, correct = ops.ce_pair(sim, mask, labels, cfg.TRAIN.SMOOTH.GAMMA3)
    accuracy = correct * (100.0 / (batch_size * 2.0)) if top1 else None
    return loss0, loss1, att_maps, accuracy


def _obj_g_term(netObjD, fake_img, seg, slabels_emb, raw_bt_c_codes, rois, num_rois, is_large_scale):
    """One object-discriminator term of G_loss (ref: miscc/losses.py:436-478 / 481-523); None when no roi of that
    scale class exists.  The generator's bt_c_code rows keep their gradient through the conditioning code."""
    net = netObjD.module if hasattr(netObjD, "module") else netObjD
    pooled = netObjD(fake_img, seg, rois, num_rois)
    feats, classes, codes = feat_select(pooled, raw_bt_c_codes, rois, num_rois, is_large_scale)
    if len(classes) == 0:
        return None
    idx = ops.h2d([int(c) for c in classes], fake_img.device, torch.int64)
    conditions = ops.cat_rows(ops.gather_rows(slabels_emb.detach(), idx), codes)
    err = ops.bce(net.COND_DNET(feats, conditions), 1.0, 1.0)
    if net.UNCOND_DNET is not None:
        err = err + ops.bce(net.UNCOND_DNET(feats), 1.0, 1.0)
    return err * cfg.TRAIN.SMOOTH.OBJ_LAMBDA


def G_loss(netsPatD, netsShpD, netObjSSD, netObjLSD, image_encoder, fake_imgs, seg_conditions, words_embs, sent_emb,
           slabels_emb, raw_bt_c_codes, match_labels, cap_lens, class_ids, rois, fm_rois, num_rois):
    """ref: miscc/losses.py:364-531 -- the generator's full adversarial + matching loss: per scale the patch-D terms
    (COND * TXT_LAMBDA + UNCOND * UNCOND_LAMBDA) and the shape-D term (SHP_LAMBDA); at the last scale the DAMSM
    word / sentence matching terms through ``image_encoder`` (the pretrained Inception CNN_ENCODER: a stock PyTorch
    module returning (region_features (B, nef, 17, 17), cnn_code (B, nef)); None skips the two terms); the small- and
    large-scale object-D terms (OBJ_LAMBDA).  Returns (errG_total, logs) with ``logs`` a dict of device scalars instead
    of the reference's formatted string (no per-step host synchronisation)."""
    batch_size = fake_imgs[0].size(0)
    ul, tl = cfg.TRAIN.SMOOTH.UNCOND_LAMBDA, cfg.TRAIN.SMOOTH.TXT_LAMBDA
    logs = {}
    total = None

    def add(x):
        nonlocal total
        total = x if total is None else total + x

    n_d = len(netsPatD)
    for i in range(n_d):
        features = netsPatD[i](fake_imgs[i])
        cond, uncond = _heads(netsPatD[i])
        pat = ops.bce(cond(features, sent_emb), 1.0, tl if uncond is not None else 1.0)
        if uncond is not None:
            pat = pat + ops.bce(uncond(features), 1.0, ul)
        add(pat)
        logs[f"pat_g_loss{i}"] = pat.detach()
        shp_net = netsShpD[i].module if hasattr(netsShpD[i], "module") else netsShpD[i]
        shp = ops.bce(shp_net.UNCOND_DNET(netsShpD[i](fake_imgs[i], seg_conditions[i])), 1.0, cfg.TRAIN.SMOOTH.SHP_LAMBDA)
        add(shp)
        logs[f"shp_g_loss{i}"] = shp.detach()
        if i == n_d - 1 and image_encoder is not None:
            region_features, cnn_code = image_encoder(fake_imgs[i])
            w0, w1, _, _ = words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size,
                                      need_att_maps=False)
            s0, s1, _ = sent_loss(cnn_code, sent_emb, match_labels, class_ids, batch_size)
            w_loss = (w0 + w1) * cfg.TRAIN.SMOOTH.DAMSM_LAMBDA
            s_loss = (s0 + s1) * cfg.TRAIN.SMOOTH.DAMSM_LAMBDA
            add(w_loss + s_loss)
            logs["w_loss"], logs["s_loss"] = w_loss.detach(), s_loss.detach()
    ss = _obj_g_term(netObjSSD, fake_imgs[-1], seg_conditions[-1], slabels_emb, raw_bt_c_codes, rois, num_rois, False)
    if ss is not None:
        add(ss)
        logs["objss_g_loss"] = ss.detach()
    ls = _obj_g_term(netObjLSD, fake_imgs[-1], seg_conditions[-1], slabels_emb, raw_bt_c_codes, fm_rois, num_rois, True)
    if ls is not None:
        add(ls)
        logs["objls_g_loss"] = ls.detach()
    return total, logs
