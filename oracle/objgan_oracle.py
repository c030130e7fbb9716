"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (torch-CPU fp32 / numpy, functional style over a plain ``state_dict``) of
the Obj-GAN ``image_generation`` G+D training step, used ONLY as the checker in ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs.
Nothing under ``obj-gan_b200/`` may import this file.

Parity pinning: the reference ships no tests / golden vectors, so
this restatement is pinned (a) live against the reference's own modules imported read-only
from /root/reference (tests/test_oracle_vs_reference.py, runs only where the reference is
mounted) and (b) against fixtures those modules produced here, committed under
tests/golden/ together with tests/golden/make_golden.py.

All "ref:" citations are paths under /root/reference/image_generation/.
The arithmetic primitives (conv2d, batch_norm, bmm, softmax ...) are torch-CPU fp32 ops --
the same third-party arithmetic the reference itself calls (README.md:15-17 pins
"Pytorch 0.4.1"; semantics of the ops used are unchanged in torch 2.x).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU = 0.2
EPS_NORM = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def glu(x):
    """ref: model.py:19-27 -- first half of the channels gated by sigmoid of the second half."""
    c = x.shape[1] // 2
    return x[:, :c] * torch.sigmoid(x[:, c:])


def batch_norm_train(x, sd, prefix, update=True):
    """nn.BatchNorm{1,2}d in train mode (the trainer never calls .eval(); ref: trainer.py:331-472).
    Batch statistics, biased variance for normalisation, running stats updated with momentum 0.1
    and the unbiased variance; num_batches_tracked += 1."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not update:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, BN_MOMENTUM, EPS_NORM)
    if update and (prefix + ".num_batches_tracked") in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return y


def instance_norm(x):
    """nn.InstanceNorm2d defaults: no affine, no running stats, eps 1e-5 (ref: model.py:70)."""
    return F.instance_norm(x, eps=EPS_NORM)


def reflect_conv3x3(x, w, b=None):
    """ReflectionPad2d(1) followed by a padding-0 3x3 conv (ref: model.py:67-69, 600-601)."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)


def up_block(x, sd, p, update=True):
    """ref: model.py:43-49 -- nearest x2, conv3x3 (no bias) C -> 2C', BatchNorm2d, GLU."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, sd[p + ".1.weight"], None, 1, 1)
    x = batch_norm_train(x, sd, p + ".2", update)
    return glu(x)


def hmap_res_block(x, sd, p):
    """ref: model.py:63-81 -- x + IN(conv(pad(GLU(IN(conv(pad(x)))))))."""
    y = reflect_conv3x3(x, sd[p + ".block.1.weight"])
    y = glu(instance_norm(y))
    y = reflect_conv3x3(y, sd[p + ".block.5.weight"])
    return x + instance_norm(y)


# --------------------------------------------------------------------------------------
# attention (GlobalAttention.py)
# --------------------------------------------------------------------------------------
def quirk_mask_rows(mask, B, Q):
    """The reference tiles the (B, L) caption mask with ``mask.repeat(queryL, 1)``
    (ref: GlobalAttention.py:108, 168) while the score rows are ordered (b, q) b-major, so row
    (b, q) is masked with the caption mask of sample (b*Q + q) mod B.  Returns (B, Q, L) bool."""
    idx = (torch.arange(B * Q) % B).view(B, Q)
    return mask[idx]


def global_attention_general(h, words, w_ctx, mask):
    """ATT_NET forward (ref: GlobalAttention.py:83-122).
    h (B, idf, ih, iw), words (B, cdf, L), w_ctx (idf, cdf, 1, 1), mask (B, L) bool or None
    -> weightedContext (B, idf, ih, iw), attn (B, L, ih, iw)."""
    B, idf, ih, iw = h.shape
    Q, L = ih * iw, words.shape[2]
    src = torch.matmul(w_ctx.view(idf, -1), words)                  # (B, idf, L)   conv1x1 on words
    s = torch.bmm(h.reshape(B, idf, Q).transpose(1, 2), src)        # (B, Q, L)
    if mask is not None:
        s = s.masked_fill(quirk_mask_rows(mask, B, Q), float("-inf"))
    a = torch.softmax(s, dim=-1)                                    # over the L words
    at = a.transpose(1, 2)                                          # (B, L, Q)
    wc = torch.bmm(src, at)                                         # (B, idf, Q)
    return wc.view(B, idf, ih, iw), at.reshape(B, L, ih, iw)


def global_bu_attention(labels, glove, words, w_ctx, mask, norm=True, eps=1e-8):
    """BT_ATT_NET forward (ref: GlobalAttention.py:136-181).
    labels (B, 50, R, 1), glove (B, 50, L), words (B, cdf, L) -> (B, idf, R, 1), (B, L, R, 1)."""
    B, _, R, _ = labels.shape
    L = words.shape[2]
    idf = w_ctx.shape[0]
    t = labels.reshape(B, -1, R).transpose(1, 2)                    # (B, R, 50)
    src = torch.matmul(w_ctx.view(idf, -1), words)                  # (B, idf, L)
    s = torch.bmm(t, glove)                                         # (B, R, L)
    if norm:
        nt = t.norm(2, dim=2, keepdim=True)
        ng = glove.norm(2, dim=1, keepdim=True)
        s = s / (nt * ng).clamp(min=eps)
    if mask is not None:
        s = s.masked_fill(quirk_mask_rows(mask, B, R), float("-inf"))
    a = torch.softmax(s, dim=-1)
    at = a.transpose(1, 2)                                          # (B, L, R)
    wc = torch.bmm(src, at)                                         # (B, idf, R)
    return wc.unsqueeze(3), at.unsqueeze(3)


def func_attention(query, context, gamma1):
    """DAMSM attention (ref: GlobalAttention.py:32-70).
    query (B, ndf, Lq), context (B, ndf, ih, iw) -> (B, ndf, Lq), (B, Lq, ih, iw)."""
    B, ndf, Lq = query.shape
    ih, iw = context.shape[2:]
    S = ih * iw
    ctx = context.reshape(B, ndf, S)
    s = torch.bmm(ctx.transpose(1, 2), query)                       # (B, S, Lq)
    p = torch.softmax(s, dim=-1)                                    # per region over words
    p2 = torch.softmax(p.transpose(1, 2) * gamma1, dim=-1)          # per word over regions (B, Lq, S)
    wc = torch.bmm(ctx, p2.transpose(1, 2))                         # (B, ndf, Lq)
    return wc, p2.reshape(B, Lq, ih, iw)


def pprocess_bt_attns(f, bt_mask):
    """ref: miscc/utils.py:401-413 -- out[b,k,y,x] = max_r f[b,k,r] * mask[b,r,y,x].
    f (B, num, R, 1), bt_mask (B, R, ih, iw) -> (B, num, ih, iw).  (The reference first expands
    the mask to (B, R, num, ih, iw); the product/max is the same.)"""
    prod = f[:, :, :, 0].transpose(1, 2)[:, :, :, None, None] * bt_mask[:, :, None]   # (B,R,num,ih,iw)
    return prod.max(dim=1)[0]


# --------------------------------------------------------------------------------------
# generator (model.py:455-795)
# --------------------------------------------------------------------------------------
def ca_net(sent_emb, sd, eps, p="ca_net"):
    """ref: model.py:455-483.  ``eps`` is the N(0,1) draw of ``reparametrize`` (injected)."""
    x = glu(F.linear(sent_emb, sd[p + ".fc.weight"], sd[p + ".fc.bias"]))
    c = x.shape[1] // 2
    mu, logvar = x[:, :c], x[:, c:]
    return eps * torch.exp(0.5 * logvar) + mu, mu, logvar


def init_stage_g(z, c, sd, p, update=True):
    """ref: model.py:486-518."""
    x = F.linear(torch.cat((c, z), 1), sd[p + ".fc.0.weight"])
    x = glu(batch_norm_train(x, sd, p + ".fc.1", update))
    ngf = sd[p + ".upsample1.1.weight"].shape[1]
    x = x.view(-1, ngf, 8, 8)
    x = up_block(x, sd, p + ".upsample1", update)
    return up_block(x, sd, p + ".upsample2", update)


def g_hmap(hmap, sd, p):
    """ref: model.py:589-617 -- pad, conv3x3 (+bias), IN, LReLU, conv3x3 s2 p1 (no bias), LReLU."""
    x = reflect_conv3x3(hmap, sd[p + ".conv3x3.1.weight"], sd[p + ".conv3x3.1.bias"])
    x = F.leaky_relu(instance_norm(x), LRELU)
    x = F.conv2d(x, sd[p + ".downsample1.0.weight"], None, 2, 1)
    return F.leaky_relu(x, LRELU)


def _bt_branch(sd, p, words, glove, slabels_feat, mask, bt_mask, rmax, buattn_norm):
    slabels = slabels_feat[:, :, :rmax]
    if rmax == 0:
        # ref: model.py:689-694 (NEXT stage `else`): zero code / attention / label maps and an empty raw code; the INIT
        # stage's `else` (571-576) reads an undefined `att` and crashes -- the same zeros are its evident intent
        b, (ih, iw) = slabels_feat.shape[0], bt_mask.shape[2:]
        z = lambda c: torch.zeros(b, c, ih, iw, dtype=words.dtype)
        idf = sd[p + ".bt_att.conv_context.weight"].shape[0]
        return torch.zeros(b, idf, 0, 1, dtype=words.dtype), z(idf), z(words.shape[2]), z(slabels_feat.shape[1])
    bt_c, bt_att = global_bu_attention(slabels, glove, words, sd[p + ".bt_att.conv_context.weight"], mask,
                                       buattn_norm)
    m = bt_mask[:, :rmax]
    return bt_c, pprocess_bt_attns(bt_c, m), pprocess_bt_attns(bt_att, m), pprocess_bt_attns(slabels, m)


def init_stage_g_main(h_hmap, h_sent, words, glove, slabels_feat, mask, num_rois, bt_mask, sd, p,
                      n_res, buattn_norm=True, update=True):
    """ref: model.py:521-586 (max_num_roi > 0 branch; the else branch is broken in the reference)."""
    rmax = int(num_rois.max())
    _, bt_c, _bt_att, bt_sl = _bt_branch(sd, p, words, glove, slabels_feat, mask, bt_mask, rmax, buattn_norm)
    x = torch.cat((h_hmap, h_sent, bt_c, bt_sl), 1)
    for i in range(n_res):
        x = hmap_res_block(x, sd, f"{p}.residual.{i}")
    return up_block(x, sd, p + ".upsample", update)


def next_stage_g_main(h, h_hmap, words, glove, slabels_feat, mask, num_rois, bt_mask, glb_max_num_roi, sd, p,
                      n_res, buattn_norm=True, update=True):
    """ref: model.py:620-705.  Returns out_code, raw_bt_c_code (B, glbR, idf), att, bt_att."""
    c_code, att = global_attention_general(h, words, sd[p + ".att.conv_context.weight"], mask)
    rmax = int(num_rois.max())
    raw, bt_c, bt_att, bt_sl = _bt_branch(sd, p, words, glove, slabels_feat, mask, bt_mask, rmax, buattn_norm)
    raw_full = torch.zeros(h.shape[0], h.shape[1], glb_max_num_roi, 1)
    raw_full[:, :, :rmax] = raw
    x = torch.cat((h + h_hmap, c_code, bt_c, bt_sl), 1)
    for i in range(n_res):
        x = hmap_res_block(x, sd, f"{p}.residual.{i}")
    out = up_block(x, sd, p + ".upsample", update)
    return out, raw_full.transpose(1, 2).squeeze(-1), att, bt_att


def get_image_g(h, sd, p):
    """ref: model.py:708-719."""
    return torch.tanh(F.conv2d(h, sd[p + ".img.0.weight"], None, 1, 1))


def g_net_forward(sd, inp, *, branch_num=3, glb_r_num=7, local_r_num=3, buattn_norm=True, update=True):
    """G_NET.forward (ref: model.py:747-795).  ``inp`` is a dict from objgan_b200.synth.make_inputs.
    Returns fake_imgs, bt_c_codes, att_maps, bt_att_maps, mu, logvar.  ``sd`` BN buffers are updated
    in place when ``update``."""
    words, glove, mask = inp["words_embs"], inp["glove_words_embs"], inp["mask"]
    sl, nr, glb = inp["slabels_feat"], inp["num_rois"], inp["glb_max_num_roi"]
    fake, btc, atts, btatts = [], [], [], []
    c_code, mu, logvar = ca_net(inp["sent_emb"], sd, inp["eps"])
    h1_hmap = g_hmap(inp["hmaps"][0], sd, "h_net1_hmap")
    h1_sent = init_stage_g(inp["z"], c_code, sd, "h_net1_sent", update)
    h = init_stage_g_main(h1_hmap, h1_sent, words, glove, sl, mask, nr, inp["fm_bt_masks"], sd, "h_net1_main",
                          glb_r_num, buattn_norm, update)
    fake.append(get_image_g(h, sd, "img_net1"))
    for k in range(2, branch_num + 1):
        hh = g_hmap(inp["hmaps"][k - 1], sd, f"h_net{k}_hmap")
        h, raw, att, bt_att = next_stage_g_main(h, hh, words, glove, sl, mask, nr, inp["bt_masks"][k - 2], glb,
                                                sd, f"h_net{k}_main", local_r_num, buattn_norm, update)
        fake.append(get_image_g(h, sd, f"img_net{k}"))
        btc.append(raw)
        atts.append(att)
        btatts.append(bt_att)
    return fake, btc, atts, btatts, mu, logvar


# --------------------------------------------------------------------------------------
# patch discriminators (model.py:989-1106)
# --------------------------------------------------------------------------------------
def pat_d_net(x, sd, update=True, n_layer=4):
    """PAT_D_NET{64,128,256}.forward = encode_image_by_ntimes (ref: model.py:999-1017, 1053-1106):
    conv4x4 s2 p1 -> LReLU, then (conv4x4 s2 p1 -> BN -> LReLU) x (n_layer-1); all bias-free."""
    x = F.leaky_relu(F.conv2d(x, sd["img_code.0.weight"], None, 2, 1), LRELU)
    for n in range(1, n_layer):
        i = 2 + 3 * (n - 1)
        x = F.conv2d(x, sd[f"img_code.{i}.weight"], None, 2, 1)
        x = F.leaky_relu(batch_norm_train(x, sd, f"img_code.{i + 1}", update), LRELU)
    return x


def d_get_logits(h, sd, p, c_code=None, update=True):
    """D_GET_LOGITS.forward (ref: model.py:1020-1048).  Conditional head when ``c_code`` is given:
    broadcast c_code over the grid, cat after h's channels, conv3x3 -> BN -> LReLU; then
    conv k4 s2 p0 (+bias) -> sigmoid."""
    if c_code is not None:
        cc = c_code.view(c_code.shape[0], -1, 1, 1).expand(-1, -1, h.shape[2], h.shape[3])
        h = F.conv2d(torch.cat((h, cc), 1), sd[p + ".jointConv.0.weight"], None, 1, 1)
        h = F.leaky_relu(batch_norm_train(h, sd, p + ".jointConv.1", update), LRELU)
    return torch.sigmoid(F.conv2d(h, sd[p + ".outlogits.0.weight"], sd[p + ".outlogits.0.bias"], 2, 0))


def shp_d_net(x, seg, sd, update=True):
    """SHP_D_NET{64,128,256}.forward (ref: model.py:1111-1179): cat(image, shp_code(seg)) -> conv encoder."""
    ns = reflect_conv3x3(seg, sd["shp_code.1.weight"], sd["shp_code.1.bias"])
    ns = F.leaky_relu(instance_norm(ns), LRELU)
    return pat_d_net(torch.cat((x, ns), 1), sd, update)


def bce(p, target):
    """nn.BCELoss() (mean) on probabilities; log clamped at -100 like PyTorch."""
    t = torch.full_like(p, float(target))
    return F.binary_cross_entropy(p, t)


def pat_d_loss(sd, real, fake, cond, *, uncond_lambda=1.0, txt_lambda=0.1, update=True):
    """ref: miscc/losses.py:163-211.  Two separate body forwards (real, fake.detach()), three COND
    head calls (real, fake, wrong = real[:B-1] with cond[1:]), two UNCOND head calls."""
    fr = pat_d_net(real, sd, update)
    ff = pat_d_net(fake.detach(), sd, update)
    B = fr.shape[0]
    c_real = bce(d_get_logits(fr, sd, "COND_DNET", cond, update), 1)
    c_fake = bce(d_get_logits(ff, sd, "COND_DNET", cond, update), 0)
    c_wrong = bce(d_get_logits(fr[: B - 1], sd, "COND_DNET", cond[1:B], update), 0)
    u_real = bce(d_get_logits(fr, sd, "UNCOND_DNET"), 1)
    u_fake = bce(d_get_logits(ff, sd, "UNCOND_DNET"), 0)
    return ((u_real * uncond_lambda + c_real * txt_lambda) / 2.0
            + (u_fake * uncond_lambda + (c_fake + c_wrong) * txt_lambda) / 3.0)


def g_loss_pat(sds, fakes, sent_emb, *, uncond_lambda=1.0, txt_lambda=0.1, update=True):
    """Patch-D part of G_loss (ref: miscc/losses.py:372-399)."""
    total = 0
    for sd, fake in zip(sds, fakes):
        f = pat_d_net(fake, sd, update)
        c = bce(d_get_logits(f, sd, "COND_DNET", sent_emb, update), 1)
        u = bce(d_get_logits(f, sd, "UNCOND_DNET"), 1)
        total = total + u * uncond_lambda + c * txt_lambda
    return total


def kl_loss(mu, logvar):
    """ref: miscc/losses.py:533-537."""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


# --------------------------------------------------------------------------------------
# ROIAlign (models/roi_align) -- numpy restatement; the C restatement is roi_align_oracle.c
# --------------------------------------------------------------------------------------
def roi_align_forward_np(feat, rois, ah, aw, scale):
    """ref: models/roi_align/src/roi_align.c:80-137 / roi_align_kernel.cu:15-70.
    feat (B,C,H,W) f32, rois (R,5) f32 [batch, x1, y1, x2, y2] -> (R, C, ah, aw) f32.
    Reproduces the C promotion rules: float products for the scaled corners, ``+ 1.`` and
    ``/ (ah - 1.)`` evaluated in double then narrowed to float, float sample coordinates,
    mixed float/double bilinear products (see the inline comment), double sum narrowed to float."""
    f32, f64 = np.float32, np.float64
    feat = np.ascontiguousarray(feat, dtype=f32)
    rois = np.ascontiguousarray(rois, dtype=f32)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, C, ah, aw), dtype=f32)
    sc = f32(scale)
    for n in range(R):
        b = int(rois[n, 0])
        x1, y1, x2, y2 = (f32(rois[n, 1] * sc), f32(rois[n, 2] * sc), f32(rois[n, 3] * sc), f32(rois[n, 4] * sc))
        rw = f32(max(f32(f64(f32(x2 - x1)) + 1.0), f32(0)))
        rh = f32(max(f32(f64(f32(y2 - y1)) + 1.0), f32(0)))
        bh = f32(f64(rh) / (ah - 1.0))
        bw = f32(f64(rw) / (aw - 1.0))
        for ph in range(ah):
            h = f32(f32(f32(ph) * bh) + y1)
            for pw in range(aw):
                w = f32(f32(f32(pw) * bw) + x1)
                if h < 0 or h >= H or w < 0 or w >= W:
                    continue
                hs = int(min(f32(math.floor(h)), f32(H - 2)))
                ws = int(min(f32(math.floor(w)), f32(W - 2)))
                hr32, wr32 = f32(h - f32(hs)), f32(w - f32(ws))
                hr, wr = f64(hr32), f64(wr32)
                ul, ur = feat[b, :, hs, ws].astype(f64), feat[b, :, hs, ws + 1].astype(f64)
                dl32, dr32 = feat[b, :, hs + 1, ws], feat[b, :, hs + 1, ws + 1]
                # C typing of "bottom[i] * h_ratio * ..." (roi_align.c:131-134): float * float stays float until a
                # double operand ("1. - w_ratio") appears; the four terms are then summed in double.
                t3 = (dl32 * hr32).astype(f32).astype(f64) * (1.0 - wr)
                t4 = ((dr32 * hr32).astype(f32) * wr32).astype(f32).astype(f64)
                v = ul * (1.0 - hr) * (1.0 - wr) + ur * (1.0 - hr) * wr + t3 + t4
                out[n, :, ph, pw] = v.astype(f32)
    return out


def roi_align_backward_np(top_diff, rois, feat_shape, ah, aw, scale):
    """Adjoint of roi_align_forward_np -- what roi_align_kernel.cu:94-143 computes with atomicAdd
    (the reference's CPU backward, roi_align.c:175, has an inverted bounds test and is unusable).
    Accumulated in float64 then narrowed, so it is the order-independent reference value."""
    f32, f64 = np.float32, np.float64
    B, C, H, W = feat_shape
    rois = np.ascontiguousarray(rois, dtype=f32)
    g = np.zeros(feat_shape, dtype=f64)
    sc = f32(scale)
    for n in range(rois.shape[0]):
        b = int(rois[n, 0])
        x1, y1, x2, y2 = (f32(rois[n, 1] * sc), f32(rois[n, 2] * sc), f32(rois[n, 3] * sc), f32(rois[n, 4] * sc))
        rw = f32(max(f32(f64(f32(x2 - x1)) + 1.0), f32(0)))
        rh = f32(max(f32(f64(f32(y2 - y1)) + 1.0), f32(0)))
        bh = f32(f64(rh) / (ah - 1.0))
        bw = f32(f64(rw) / (aw - 1.0))
        for ph in range(ah):
            h = f32(f32(f32(ph) * bh) + y1)
            for pw in range(aw):
                w = f32(f32(f32(pw) * bw) + x1)
                if h < 0 or h >= H or w < 0 or w >= W:
                    continue
                hs = int(min(f32(math.floor(h)), f32(H - 2)))
                ws = int(min(f32(math.floor(w)), f32(W - 2)))
                hr, wr = f64(f32(h - f32(hs))), f64(f32(w - f32(ws)))
                t = top_diff[n, :, ph, pw].astype(f64)
                g[b, :, hs, ws] += t * (1.0 - hr) * (1.0 - wr)
                g[b, :, hs, ws + 1] += t * (1.0 - hr) * wr
                g[b, :, hs + 1, ws] += t * hr * (1.0 - wr)
                g[b, :, hs + 1, ws + 1] += t * hr * wr
    return g.astype(f32)


def roi_align_avg_np(feat, rois, ah, aw, scale):
    """RoIAlignAvg.forward (ref: models/roi_align/modules/roi_align.py:18-29): align to
    (ah+1, aw+1) then avg_pool2d(kernel 2, stride 1)."""
    x = roi_align_forward_np(feat, rois, ah + 1, aw + 1, scale)
    return F.avg_pool2d(torch.from_numpy(x), kernel_size=2, stride=1).numpy()


def get_rois_blob_np(fm_rois_xywh, boxes_num=10):
    """ref: miscc/utils.py:365-399 + model.py:1213-1214, 1236-1240.  fm_rois (B, boxes_num, >=4)
    [x, y, w, h, ...] float64 -> (B*boxes_num, 5) float32 [image index, x1, y1, x2, y2]."""
    r = np.array(fm_rois_xywh, dtype=np.float64, copy=True)
    r[:, :, 2:4] = r[:, :, 0:2] + r[:, :, 2:4]
    B = r.shape[0]
    flat = r.reshape(B * boxes_num, -1)[:, :4]
    levels = np.repeat(np.arange(B), boxes_num).reshape(-1, 1).astype(np.float64)
    return np.hstack((levels, flat)).astype(np.float32)


class _RoIAlignAvgFn(torch.autograd.Function):
    """Differentiable wrapper of the numpy ROIAlign restatement (forward + adjoint), for oracle gradients."""

    @staticmethod
    def forward(ctx, feat, rois, ah, aw, scale):
        ctx.cfg = (tuple(feat.shape), ah, aw, scale)
        ctx.rois = rois
        x = roi_align_forward_np(feat.detach().numpy(), rois, ah + 1, aw + 1, scale)
        return F.avg_pool2d(torch.from_numpy(x), kernel_size=2, stride=1)

    @staticmethod
    def backward(ctx, g):
        shape, ah, aw, scale = ctx.cfg
        g6 = torch.zeros(g.shape[0], g.shape[1], ah + 1, aw + 1)
        for dh in (0, 1):
            for dw in (0, 1):
                g6[:, :, dh:dh + ah, dw:dw + aw] += g / 4
        return torch.from_numpy(roi_align_backward_np(g6.numpy(), ctx.rois, shape, ah + 1, aw + 1, scale)), None, None, \
            None, None


def obj_d_net_forward(sd, x, s, fm_rois, n_layer, img_size=512, update=True, boxes_num=10):
    """OBJ_SS_D_NET / OBJ_LS_D_NET .forward (ref: model.py:1212-1246 / 1278-1312; n_layer = 3 / 4).
    x (B,3,256,256), s (B,80,256,256), fm_rois (B,10,>=4) [x,y,w,h,...] -> (B, 10, 384, 4, 4).
    Lines 1227-1241 (the Variable/resize_ idiom that no longer runs) are restated from their intent: build the
    (B*10, 5) float32 roi blob with _get_rois_blob and pool every slot, valid or not."""
    x5 = F.interpolate(x, size=(img_size, img_size), mode="bilinear", align_corners=True)
    s5 = F.interpolate(s, size=(img_size, img_size), mode="bilinear", align_corners=True)
    ns = reflect_conv3x3(s5, sd["shp_code.1.weight"], sd["shp_code.1.bias"])
    ns = F.leaky_relu(instance_norm(ns), LRELU)
    code = pat_d_net(torch.cat((x5, ns), 1), sd, update, n_layer)
    rois = get_rois_blob_np(np.asarray(fm_rois, dtype=np.float64), boxes_num)
    pooled = _RoIAlignAvgFn.apply(code, rois, 5, 5, 1.0 / 16.0)
    out = F.leaky_relu(F.conv2d(pooled, sd["roi_code.0.weight"], sd["roi_code.0.bias"], 1, 1), LRELU)
    return out.view(x.shape[0], boxes_num, out.shape[1], out.shape[2], out.shape[3])


def permute_seg(seg, rois, num_rois):
    """ref: miscc/utils.py:445-462 -- per sample with rois, shuffle (host ``random.shuffle``) the segmentation channels
    of the classes present; returns the permuted maps and the indices of the samples whose order changed."""
    import random
    from copy import deepcopy
    new_seg = seg.clone()
    rois = np.asarray(rois)
    valid = []
    for b in range(seg.shape[0]):
        if int(num_rois[b]) == 0:
            continue
        classes = list(np.unique(rois[b, :int(num_rois[b]), 4]).astype(int))
        shuffled = deepcopy(classes)
        random.shuffle(shuffled)
        if classes != shuffled:
            valid.append(b)
            new_seg[b, classes] = seg[b, shuffled]
    return new_seg, valid


def feat_select(pooled, raw_bt_c_codes, fm_rois, num_rois, is_large_scale=False, size_thrs=16.0):
    """ref: miscc/utils.py:465-499 -- rois of one scale class: drop boxes smaller than 1.25 x 1.25 feature-map cells,
    keep max(w, h) >= ROI_SIZE_THRS for the large-scale net and < ROI_SIZE_THRS for the small-scale one."""
    fm = np.asarray(fm_rois)
    feats, classes, codes = [], [], []
    for b in range(len(num_rois)):
        keep = []
        for r in range(int(num_rois[b])):
            _, _, w, h = fm[b, r, :4]
            if w < 1.25 and h < 1.25:
                continue
            if (max(w, h) >= size_thrs) != bool(is_large_scale):
                continue
            keep.append(r)
        if not keep:
            continue
        feats.append(pooled[b, keep])
        classes.append(fm[b, keep, 4].astype(int))
        codes.append(raw_bt_c_codes[b, keep])
    if not classes:
        return [], [], []
    return torch.cat(feats, 0), np.concatenate(classes), torch.cat(codes, 0)


def obj_d_loss(sd, real, fake, seg, raw_conditions, raw_bt_c_codes, fm_rois, num_rois, n_layer, *,
               is_large_scale=False, update=True, size_thrs=16.0):
    """ref: miscc/losses.py:254-361 (objD_loss) on top of obj_d_net_forward / d_get_logits.  ``fm_rois``: host array
    (B, 10, >= 5) [x, y, w, h, class, ...]; ``num_rois``: host ints.  The permuted-shape conditions index ``classes``
    (not ``classes2``) exactly like the reference (losses.py:307-311)."""
    fm = np.asarray(fm_rois)
    nums = [int(v) for v in num_rois]
    real_pooled = obj_d_net_forward(sd, real, seg, fm, n_layer, update=update)
    rf, classes, codes = feat_select(real_pooled, raw_bt_c_codes, fm, nums, is_large_scale, size_thrs)
    fake_pooled = obj_d_net_forward(sd, fake.detach(), seg, fm, n_layer, update=update)
    ff, _, _ = feat_select(fake_pooled, raw_bt_c_codes, fm, nums, is_large_scale, size_thrs)
    fake_seg, valid = permute_seg(seg, fm, nums)
    classes2 = []
    if len(valid) > 0:
        pooled2 = obj_d_net_forward(sd, real[valid], fake_seg[valid], fm[valid], n_layer, update=update)
        ff2, classes2, codes2 = feat_select(pooled2, raw_bt_c_codes, fm[valid], [nums[v] for v in valid],
                                            is_large_scale, size_thrs)
    n = len(classes)
    if n == 0:
        return 0
    cond = torch.cat((raw_conditions[torch.as_tensor(classes)], codes), 1)
    c_real = bce(d_get_logits(rf, sd, "COND_DNET", cond, update), 1)
    c_fake = bce(d_get_logits(ff, sd, "COND_DNET", cond, update), 0)
    tmp = c_fake
    if n > 1:
        tmp = tmp + bce(d_get_logits(rf[:n - 1], sd, "COND_DNET", cond[1:n], update), 0)
    extra = 0.0
    if len(valid) > 0 and len(classes2) > 0:
        cond2 = torch.cat((raw_conditions[torch.as_tensor(classes[:len(classes2)])], codes2), 1)
        tmp = tmp + bce(d_get_logits(ff2, sd, "COND_DNET", cond2, update), 0)
        extra = 1.0
    if "UNCOND_DNET.outlogits.0.weight" in sd:
        err = (bce(d_get_logits(rf, sd, "UNCOND_DNET", None, update), 1) + c_real) / 2.0
        tmp = tmp + bce(d_get_logits(ff, sd, "UNCOND_DNET", None, update), 0)
        return err + tmp / (3.0 + extra)
    return c_real + tmp / (2.0 + extra)


# --------------------------------------------------------------------------------------
# DAMSM matching losses (miscc/losses.py:13-159)
# --------------------------------------------------------------------------------------
def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """ref: miscc/losses.py:13-19."""
    w12 = torch.sum(x1 * x2, dim)
    w1 = torch.norm(x1, 2, dim)
    w2 = torch.norm(x2, 2, dim)
    return (w12 / (w1 * w2).clamp(min=eps)).squeeze()


def _class_masks(class_ids, batch_size):
    """ref: miscc/losses.py:27-38 / 86-89."""
    if class_ids is None:
        return None
    ids = np.asarray(class_ids)
    masks = []
    for i in range(batch_size):
        m = (ids == ids[i]).astype(np.uint8)
        m[i] = 0
        masks.append(m.reshape(1, -1))
    return torch.from_numpy(np.concatenate(masks, 0)).bool()


def _two_ce(scores0, labels, batch_size):
    scores1 = scores0.transpose(0, 1)
    loss0 = F.cross_entropy(scores0, labels)
    loss1 = F.cross_entropy(scores1, labels)
    correct = (scores0.max(1)[1] == labels).sum().item() + (scores1.max(1)[1] == labels).sum().item()
    return loss0, loss1, 100.0 * correct / (batch_size * 2.0)


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, *, gamma3=10.0, eps=1e-8):
    """ref: miscc/losses.py:22-71."""
    masks = _class_masks(class_ids, batch_size)
    if cnn_code.dim() == 2:
        cnn_code, rnn_code = cnn_code.unsqueeze(0), rnn_code.unsqueeze(0)
    cn = torch.norm(cnn_code, 2, dim=2, keepdim=True)
    rn = torch.norm(rnn_code, 2, dim=2, keepdim=True)
    scores0 = torch.bmm(cnn_code, rnn_code.transpose(1, 2))
    norm0 = torch.bmm(cn, rn.transpose(1, 2))
    scores0 = (scores0 / norm0.clamp(min=eps) * gamma3).squeeze(0)
    if masks is not None:
        scores0 = scores0.masked_fill(masks, -float("inf"))
    return _two_ce(scores0, labels, batch_size)


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size, *, gamma1=4.0, gamma2=5.0,
               gamma3=10.0):
    """ref: miscc/losses.py:74-159.  Returns (loss0, loss1, att_maps, accuracy)."""
    masks = _class_masks(class_ids, batch_size)
    att_maps, sims = [], []
    for i in range(batch_size):
        n = int(cap_lens[i])
        word = words_emb[i, :, :n].unsqueeze(0).contiguous().repeat(batch_size, 1, 1)
        wei, attn = func_attention(word, img_features, gamma1)
        att_maps.append(attn[i].unsqueeze(0).contiguous())
        w = word.transpose(1, 2).contiguous().view(batch_size * n, -1)
        c = wei.transpose(1, 2).contiguous().view(batch_size * n, -1)
        row = cosine_similarity(w, c).view(batch_size, n)
        row = torch.log((row * gamma2).exp().sum(dim=1, keepdim=True))
        sims.append(row)
    sim = torch.cat(sims, 1) * gamma3
    if masks is not None:
        sim = sim.masked_fill(masks, -float("inf"))
    loss0, loss1, acc = _two_ce(sim, labels, batch_size)
    return loss0, loss1, att_maps, acc


def _obj_g_term(sd, fake, seg, slabels_emb, raw_bt_c_codes, rois, num_rois, n_layer, is_large_scale, obj_lambda,
                size_thrs, update):
    """ref: miscc/losses.py:436-478 / 481-523."""
    pooled = obj_d_net_forward(sd, fake, seg, np.asarray(rois), n_layer, update=update)
    feats, classes, codes = feat_select(pooled, raw_bt_c_codes, rois, num_rois, is_large_scale, size_thrs)
    if len(classes) == 0:
        return None
    cond = torch.cat((slabels_emb[torch.as_tensor(classes)], codes), 1)
    err = bce(d_get_logits(feats, sd, "COND_DNET", cond, update), 1)
    if "UNCOND_DNET.outlogits.0.weight" in sd:
        err = err + bce(d_get_logits(feats, sd, "UNCOND_DNET", None, update), 1)
    return err * obj_lambda


def g_loss(sds_pat, sds_shp, sd_ss, sd_ls, image_encoder, fake_imgs, seg_conditions, words_embs, sent_emb, slabels_emb,
           raw_bt_c_codes, match_labels, cap_lens, class_ids, rois, fm_rois, num_rois, *, uncond_lambda=1.0,
           txt_lambda=0.1, shp_lambda=1.0, obj_lambda=0.1, damsm_lambda=100.0, size_thrs=16.0, update=True):
    """G_loss (ref: miscc/losses.py:364-531) over state_dicts; ``image_encoder`` is any callable returning
    (region_features, cnn_code) or None (no DAMSM terms).  Returns the total and a dict of the terms."""
    B = fake_imgs[0].shape[0]
    nums = [int(v) for v in num_rois]
    terms = {}
    total = 0
    n_d = len(sds_pat)
    for i in range(n_d):
        f = pat_d_net(fake_imgs[i], sds_pat[i], update)
        pat = bce(d_get_logits(f, sds_pat[i], "COND_DNET", sent_emb, update), 1)
        if "UNCOND_DNET.outlogits.0.weight" in sds_pat[i]:
            pat = bce(d_get_logits(f, sds_pat[i], "UNCOND_DNET", None, update), 1) * uncond_lambda + pat * txt_lambda
        terms[f"pat_g_loss{i}"] = pat
        total = total + pat
        fs = shp_d_net(fake_imgs[i], seg_conditions[i], sds_shp[i], update)
        shp = bce(d_get_logits(fs, sds_shp[i], "UNCOND_DNET", None, update), 1) * shp_lambda
        terms[f"shp_g_loss{i}"] = shp
        total = total + shp
        if i == n_d - 1 and image_encoder is not None:
            region, code = image_encoder(fake_imgs[i])
            w0, w1, _, _ = words_loss(region, words_embs, match_labels, cap_lens, class_ids, B)
            s0, s1, _ = sent_loss(code, sent_emb, match_labels, class_ids, B)
            terms["w_loss"], terms["s_loss"] = (w0 + w1) * damsm_lambda, (s0 + s1) * damsm_lambda
            total = total + terms["w_loss"] + terms["s_loss"]
    ss = _obj_g_term(sd_ss, fake_imgs[-1], seg_conditions[-1], slabels_emb, raw_bt_c_codes, np.asarray(rois), nums, 3,
                     False, obj_lambda, size_thrs, update)
    if ss is not None:
        terms["objss_g_loss"] = ss
        total = total + ss
    ls = _obj_g_term(sd_ls, fake_imgs[-1], seg_conditions[-1], slabels_emb, raw_bt_c_codes, np.asarray(fm_rois), nums,
                     4, True, obj_lambda, size_thrs, update)
    if ls is not None:
        terms["objls_g_loss"] = ls
        total = total + ls
    return total, terms


# --------------------------------------------------------------------------------------
# optimiser (trainer.py:197-224, 461-462)
# --------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=2e-4, b1=0.5, b2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) single-tensor update, in place on p, m, v.
    ``step`` is the 1-based step count."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def ema_update(avg, p, decay=0.999):
    """ref: trainer.py:461-462 -- avg = 0.999*avg + 0.001*p."""
    avg.mul_(decay).add_(p, alpha=1 - decay)


# --------------------------------------------------------------------------------------
# Step-A: G fwd -> 3 PatD updates -> G update (PatD terms + KL) -> EMA
# --------------------------------------------------------------------------------------
PARAM_SUFFIXES = (".weight", ".bias")


def trainable_keys(sd):
    return [k for k, v in sd.items() if k.endswith(PARAM_SUFFIXES) and v.dtype.is_floating_point]


def _with_grad(sd, keys):
    out = dict(sd)
    leaves = {}
    for k in keys:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        out[k] = leaves[k]
    return out, leaves


class StepAState:
    """Holds G / 3 PatD state_dicts, Adam moments, EMA copy of G and the step counter."""

    def __init__(self, g_sd, d_sds):
        self.g = {k: v.clone() for k, v in g_sd.items()}
        self.ds = [{k: v.clone() for k, v in d.items()} for d in d_sds]
        self.step = 0
        self.g_keys = trainable_keys(self.g)
        self.d_keys = [trainable_keys(d) for d in self.ds]
        z = lambda sd, keys: {k: torch.zeros_like(sd[k]) for k in keys}
        self.g_m, self.g_v = z(self.g, self.g_keys), z(self.g, self.g_keys)
        self.d_m = [z(d, k) for d, k in zip(self.ds, self.d_keys)]
        self.d_v = [z(d, k) for d, k in zip(self.ds, self.d_keys)]
        self.g_avg = {k: self.g[k].clone() for k in self.g_keys}


def step_a(state: StepAState, inp: dict, *, lr=2e-4, keep=None):
    """One Step-A training step on the CPU (ref: trainer.py:388-462 restricted to G + PatD + KL).
    Returns a dict of losses; if ``keep`` is a dict it receives fake images and all gradients."""
    state.step += 1
    t = state.step
    # (2) generate fake images (graph kept for the G update)
    g_live, g_leaves = _with_grad(state.g, state.g_keys)
    fake, _btc, _att, _btatt, mu, logvar = g_net_forward(g_live, inp)
    for k, v in g_live.items():  # carry BN buffer updates back
        if k not in g_leaves:
            state.g[k] = v
    losses = {}
    sent = inp["sent_emb"]
    # (3-1) PatD updates
    d_grads = []
    for i, d in enumerate(state.ds):
        d_live, d_leaves = _with_grad(d, state.d_keys[i])
        err = pat_d_loss(d_live, inp["imgs"][i], fake[i], sent)
        grads = torch.autograd.grad(err, [d_leaves[k] for k in state.d_keys[i]])
        for k, v in d_live.items():
            if k not in d_leaves:
                d[k] = v
        for k, gk in zip(state.d_keys[i], grads):
            adam_step(d[k], gk, state.d_m[i][k], state.d_v[i][k], t, lr)
        losses[f"errPatD{i}"] = float(err.detach())
        d_grads.append(dict(zip(state.d_keys[i], grads)))
    # (4) G update through the (already updated) PatDs
    errg = g_loss_pat(state.ds, fake, sent)
    kl = kl_loss(mu, logvar)
    total = errg + kl
    ggrads = torch.autograd.grad(total, [g_leaves[k] for k in state.g_keys])
    for k, gk in zip(state.g_keys, ggrads):
        adam_step(state.g[k], gk, state.g_m[k], state.g_v[k], t, lr)
        ema_update(state.g_avg[k], state.g[k])
    losses["errG"] = float(errg)
    losses["kl"] = float(kl)
    if keep is not None:
        keep["fake"] = [f.detach() for f in fake]
        keep["g_grads"] = dict(zip(state.g_keys, ggrads))
        keep["d_grads"] = d_grads
    return losses


def step_a_dp(state: StepAState, shards: list, *, lr=2e-4, keep=None):
    """Step-A under data parallelism over len(shards) equal shards: every shard runs the reference
    step on its own samples (local BatchNorm statistics, local "wrong pair" shift, local mask quirk), the gradients of
    each network are AVERAGED over the shards (what one NCCL all-reduce per bucket + 1/N does) and one Adam step is
    taken from the averaged gradient; the generator update goes through the discriminators after THEIR averaged step.
    Returns per-shard losses; ``keep`` receives the averaged gradients."""
    state.step += 1
    t, n = state.step, len(shards)
    g_live, g_leaves = _with_grad(state.g, state.g_keys)
    outs = []
    for i, inp in enumerate(shards):
        live_i = g_live if i == 0 else dict(g_live)           # BatchNorm buffer updates of shard 0 are kept
        outs.append(g_net_forward(live_i, inp))
    for k, v in g_live.items():
        if k not in g_leaves:
            state.g[k] = v
    losses = [dict() for _ in shards]
    d_grads = []
    for j, d in enumerate(state.ds):
        d_live, d_leaves = _with_grad(d, state.d_keys[j])
        total = 0
        for i, inp in enumerate(shards):
            live_i = d_live if i == 0 else dict(d_live)
            err = pat_d_loss(live_i, inp["imgs"][j], outs[i][0][j], inp["sent_emb"])
            losses[i][f"errPatD{j}"] = float(err.detach())
            total = total + err / n
        grads = torch.autograd.grad(total, [d_leaves[k] for k in state.d_keys[j]])
        for k, v in d_live.items():
            if k not in d_leaves:
                d[k] = v
        for k, gk in zip(state.d_keys[j], grads):
            adam_step(d[k], gk, state.d_m[j][k], state.d_v[j][k], t, lr)
        d_grads.append(dict(zip(state.d_keys[j], grads)))
    total = 0
    for i, inp in enumerate(shards):
        ds_i = state.ds if i == 0 else [dict(d) for d in state.ds]
        errg = g_loss_pat(ds_i, outs[i][0], inp["sent_emb"])
        kl = kl_loss(outs[i][4], outs[i][5])
        losses[i]["errG"], losses[i]["kl"] = float(errg.detach()), float(kl.detach())
        total = total + (errg + kl) / n
    ggrads = torch.autograd.grad(total, [g_leaves[k] for k in state.g_keys])
    for k, gk in zip(state.g_keys, ggrads):
        adam_step(state.g[k], gk, state.g_m[k], state.g_v[k], t, lr)
        ema_update(state.g_avg[k], state.g[k])
    if keep is not None:
        keep["g_grads"] = dict(zip(state.g_keys, ggrads))
        keep["d_grads"] = d_grads
        keep["fake"] = [[f.detach() for f in o[0]] for o in outs]
    return losses


# --------------------------------------------------------------------------------------
# Step-B: the reference's complete step (trainer.py:385-462) -- SURVEY.md section 8, row a21 in full
# --------------------------------------------------------------------------------------
def shp_d_loss(sd, real, fake, seg, rois, num_rois, update=True):
    """ref: miscc/losses.py:213-251 (shpD_loss): real / fake / permuted-shape terms on the UNCOND head."""
    rf = shp_d_net(real, seg, sd, update)
    ff = shp_d_net(fake.detach(), seg, sd, update)
    fake_seg, valid = permute_seg(seg, rois, [int(v) for v in num_rois])
    err = bce(d_get_logits(rf, sd, "UNCOND_DNET", None, update), 1)
    fake_err = bce(d_get_logits(ff, sd, "UNCOND_DNET", None, update), 0)
    if len(valid) > 0:
        wrong = shp_d_net(real[valid], fake_seg[valid], sd, update)
        return err + (fake_err + bce(d_get_logits(wrong, sd, "UNCOND_DNET", None, update), 0)) / 2.0
    return err + fake_err


class StepBState(StepAState):
    """StepAState + the three shape discriminators and the two object discriminators with their Adam moments."""

    def __init__(self, g_sd, d_sds, shp_sds, ss_sd, ls_sd):
        super().__init__(g_sd, d_sds)
        self.extra = [{k: v.clone() for k, v in sd.items()} for sd in [*shp_sds, ss_sd, ls_sd]]
        self.x_keys = [trainable_keys(sd) for sd in self.extra]
        z = lambda sd, keys: {k: torch.zeros_like(sd[k]) for k in keys}
        self.x_m = [z(sd, k) for sd, k in zip(self.extra, self.x_keys)]
        self.x_v = [z(sd, k) for sd, k in zip(self.extra, self.x_keys)]
        self.x_step = [0] * len(self.extra)          # an object discriminator only steps when it has a loss


def step_b(state: StepBState, inp: dict, *, image_encoder=None, class_ids=None, lr=2e-4):
    """One complete step on the CPU: G forward; PatD x3, ShpD x3, ObjSS, ObjLS updates; G update through all of them
    (+ DAMSM terms when ``image_encoder`` is given) + KL; EMA.  Host randomness (permute_seg) is drawn from Python's
    ``random`` in exactly this order, which is also the product's order."""
    state.step += 1
    t = state.step
    g_live, g_leaves = _with_grad(state.g, state.g_keys)
    fake, btc, _att, _btatt, mu, logvar = g_net_forward(g_live, inp)
    for k, v in g_live.items():
        if k not in g_leaves:
            state.g[k] = v
    losses = {}
    sent, imgs, hmaps = inp["sent_emb"], inp["imgs"], inp["hmaps"]
    rois = [np.asarray(r) for r in inp["rois"]]
    fm, nums = np.asarray(inp["fm_rois"]), [int(v) for v in inp["num_rois"]]

    def d_update(sd, keys, m, v, loss_fn, step):
        live, leaves = _with_grad(sd, keys)
        err = loss_fn(live)
        if not torch.is_tensor(err):
            return None
        grads = torch.autograd.grad(err, [leaves[k] for k in keys], allow_unused=True)
        for k, val in live.items():
            if k not in leaves:
                sd[k] = val
        for k, gk in zip(keys, grads):
            adam_step(sd[k], gk if gk is not None else torch.zeros_like(sd[k]), m[k], v[k], step, lr)
        return float(err.detach())

    for i, d in enumerate(state.ds):
        losses[f"errPatD{i}"] = d_update(d, state.d_keys[i], state.d_m[i], state.d_v[i],
                                         lambda live, i=i: pat_d_loss(live, imgs[i], fake[i], sent), t)
    for i in range(3):
        losses[f"errShpD{i}"] = d_update(state.extra[i], state.x_keys[i], state.x_m[i], state.x_v[i],
                                         lambda live, i=i: shp_d_loss(live, imgs[i], fake[i], hmaps[i], rois[i], nums), t)
        state.x_step[i] += 1
    codes = btc[-1].detach()      # ref: trainer.py:393 detaches every bt_c_code right after the G forward
    for j, (name, boxes, n_layer, large) in enumerate((("errObjSSD", rois[0], 3, False), ("errObjLSD", fm, 4, True))):
        idx = 3 + j
        nxt = state.x_step[idx] + 1
        res = d_update(state.extra[idx], state.x_keys[idx], state.x_m[idx], state.x_v[idx],
                       lambda live: obj_d_loss(live, imgs[-1], fake[-1], hmaps[-1], inp["clabels_emb"], codes, boxes, nums,
                                               n_layer, is_large_scale=large), nxt)
        if res is not None:
            state.x_step[idx] = nxt
        losses[name] = res
    labels = torch.arange(fake[0].shape[0])
    errg, terms = g_loss(state.ds, state.extra[:3], state.extra[3], state.extra[4], image_encoder, fake, hmaps,
                         inp["words_embs"], sent, inp["clabels_emb"], codes, labels,
                         [int(v) for v in inp["cap_lens"]], class_ids, rois[0], fm, nums)
    kl = kl_loss(mu, logvar)
    ggrads = torch.autograd.grad(errg + kl, [g_leaves[k] for k in state.g_keys])
    for k, gk in zip(state.g_keys, ggrads):
        adam_step(state.g[k], gk, state.g_m[k], state.g_v[k], t, lr)
        ema_update(state.g_avg[k], state.g[k])
    losses["errG"], losses["kl"] = float(errg), float(kl)
    losses["terms"] = {k: float(v) for k, v in terms.items()}
    losses["fake"] = [f.detach() for f in fake]
    return losses
