"""Host-side mirror of the reference's module API for the hot path.

Same class names, constructor arguments, ``forward`` signatures and ``state_dict`` keys / shapes as
/root/reference/image_generation/model.py (G side: lines 19-81, 455-795; patch discriminators:
989-1106; object discriminators: 1184-1312) and GlobalAttention.py (73-181), so ``trainer.py`` /
``miscc/losses.py`` call sites stay drop-in and reference checkpoints load with ``strict=True``.
Every forward runs the sm_100a kernels of libobjgan_b200.so through ``ops`` -- no torch.nn compute.

Public tensors are NCHW fp32 exactly like the reference's; internally activations are NHWC with
channels padded to 8 (``*_nhwc`` methods exchange that layout directly and are what G_NET uses
between its stages).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .config import cfg
from .lib import (ACT_LRELU, ACT_NONE, ACT_SIGMOID, ACT_TANH, NA_GLU, NA_LRELU, NA_NONE, PAD_REFLECT, PAD_ZERO,
                  UPSAMPLE2X)
from .ops import cpad


# --------------------------------------------------------------------------------------------------
# parameter holders (leaf modules).  Their attribute names give the reference's state_dict keys.
# --------------------------------------------------------------------------------------------------
# Host-logic tests build dozens of networks on the CPU and never look at the values: they may swap the reference's
# orthogonal initialisation (a QR per weight, seconds for the large layers) for a plain normal draw.
FAST_INIT = False


def _init_weight(w):
    if FAST_INIT:
        w.normal_(0.0, 0.02)
    else:
        nn.init.orthogonal_(w, 1.0)


class Conv2dP(nn.Module):
    """Holds an OIHW ``weight`` (+ optional ``bias``) like nn.Conv2d; forward is an NHWC kernel call."""

    def __init__(self, cin, cout, k, stride=1, pad=1, bias=False, mode=PAD_ZERO, act=ACT_NONE, split=0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        self.stride, self.pad, self.mode, self.act, self.split = stride, pad, mode, act, split
        self._cache = ops.PackedWeights()
        _init_weight(self.weight.data)
        if bias:
            bound = 1.0 / np.sqrt(cin * k * k)
            nn.init.uniform_(self.bias.data, -bound, bound)

    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, self._cache, stride=self.stride, pad=self.pad, mode=self.mode,
                          act=self.act, split=self.split)


class LinearP(nn.Module):
    """nn.Linear parameters ((out, in) weight); runs as a 1x1 conv over a (B, 1, 1, in) tensor."""

    def __init__(self, cin, cout, bias=True, split=0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        self.split = split
        self._cache = ops.PackedWeights()
        _init_weight(self.weight.data)

    def forward(self, x2d):
        b = x2d.shape[0]
        w4 = self.weight.view(self.weight.shape[0], self.weight.shape[1], 1, 1)
        y = ops.conv2d(x2d.view(b, 1, 1, -1), w4, self.bias, self._cache, stride=1, pad=0, mode=PAD_ZERO,
                       act=ACT_NONE, split=self.split)
        return y.view(b, -1)


class BatchNormP(nn.Module):
    """nn.BatchNorm{1,2}d parameters/buffers; always train-mode statistics (the reference trainer never
    switches G or the Ds to eval(): trainer.py:331-472)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c).normal_(1.0, 0.02))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, y, act):
        return ops.batch_norm_act(y, self.weight, self.bias,
                                  (self.running_mean, self.running_var, self.num_batches_tracked), act)


class _Slots(nn.Module):
    """Children registered under the numeric names an nn.Sequential would give them, so that state_dict
    keys match the reference's Sequential-based blocks (parameter-free positions are simply absent)."""

    def __init__(self, **children):
        super().__init__()
        for name, m in children.items():
            self.add_module(name.lstrip("_"), m)

    def __getitem__(self, i):
        return getattr(self, str(i))


def weights_init(m):
    """Mirror of miscc/utils.py:309-319 for this package's parameter holders."""
    if isinstance(m, (Conv2dP, LinearP)):
        _init_weight(m.weight.data)
        if isinstance(m, LinearP) and m.bias is not None:
            m.bias.data.fill_(0.0)
    elif isinstance(m, BatchNormP):
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


class _Base(nn.Module):
    def _load_from_state_dict(self, *a, **k):
        ops.bump_param_epoch()
        return super()._load_from_state_dict(*a, **k)


# --------------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------------
class GLU(nn.Module):
    """ref: model.py:19-27 (NCHW in / out)."""

    def forward(self, x):
        nc = x.size(1)
        assert nc % 2 == 0, "channels dont divide 2!"
        if x.dim() == 2:
            return ops.glu(x)
        h = nc // 2
        y = ops.to_nhwc(x, nc) if nc % 4 == 0 else None
        assert y is not None
        return ops.to_nchw(ops.glu(y), h)


class UpBlock(_Slots):
    """upBlock (ref: model.py:43-49): keys ``1.weight`` (conv3x3, out*2 channels) and ``2.*`` (BatchNorm2d)."""

    def __init__(self, cin, cout):
        super().__init__(_1=Conv2dP(cin, cout * 2, 3, 1, 1, mode=UPSAMPLE2X, split=cout), _2=BatchNormP(cout * 2))
        assert cout % 8 == 0

    def forward(self, x):
        return self[2](self[1](x), NA_GLU)


def upBlock(in_planes, out_planes):
    return UpBlock(in_planes, out_planes)


class HmapResBlock(nn.Module):
    """ref: model.py:63-81.  keys ``block.1.weight`` (2C, C, 3, 3) and ``block.5.weight`` (C, C, 3, 3)."""

    def __init__(self, channel_num):
        super().__init__()
        c = channel_num
        self.block = _Slots(_1=Conv2dP(c, 2 * c, 3, 1, 1, mode=PAD_REFLECT, split=c),
                            _5=Conv2dP(c, c, 3, 1, 1, mode=PAD_REFLECT))

    next_pad = None      # set by the owner: halo (1 / 0) of the convolution that consumes this block's output

    def forward(self, x):  # NHWC
        # the GLU output is read by block.5 only: it is written directly as that convolution's fp16 operand copies
        # (reflection halo included) and never as fp32; the block output goes out as fp32 (the next residual add reads
        # it) plus, when the owner said what consumes it, the operand copies of that convolution
        y = ops.instance_norm_act(self.block[1](x), NA_GLU, split_pad=1, keep_f32=False)
        return ops.instance_norm_act(self.block[5](y), NA_NONE, res=x, split_pad=self.next_pad)


def _chain_res_blocks(blocks, last_pad):
    """Tell every block of a residual chain which operand layout its consumer wants: the next block's reflection-padded
    conv (halo 1), or ``last_pad`` for what follows the chain."""
    blocks = list(blocks)
    for b in blocks[:-1]:
        b.next_pad = 1
    if blocks:
        blocks[-1].next_pad = last_pad


class CA_NET(_Base):
    """ref: model.py:455-483.  ``eps`` can be injected (parity tests); otherwise drawn on the device."""

    def __init__(self):
        super().__init__()
        self.t_dim = cfg.TEXT.EMBEDDING_DIM
        self.c_dim = cfg.GAN.CONDITION_DIM
        self.fc = LinearP(self.t_dim, self.c_dim * 4, bias=True, split=self.c_dim * 2)
        self.eps_override = None

    def forward_rows(self, text_embedding):
        """Returns c_code rows (B, cpad(c_dim)) and the GLU output rows x = [mu | logvar | pad]."""
        emb = text_embedding.contiguous()
        assert emb.shape[1] % 8 == 0
        x = ops.glu(self.fc(emb))                      # (B, cpad(2*c_dim))
        eps = self.eps_override
        if eps is None:
            eps = torch.randn(emb.shape[0], self.c_dim, device=emb.device)
        c = ops.reparam(x, eps, self.c_dim)
        return c, x

    def forward(self, text_embedding):
        c, x = self.forward_rows(text_embedding)
        d = self.c_dim
        return c[:, :d], x[:, :d], x[:, d:2 * d]


class INIT_STAGE_G(_Base):
    """ref: model.py:486-518."""

    def __init__(self, ngf, ncf):
        super().__init__()
        self.gf_dim = ngf
        self.in_dim = cfg.GAN.Z_DIM + ncf
        nz = self.in_dim
        self.fc = _Slots(_0=LinearP(nz, ngf * 8 * 8 * 2, bias=False, split=ngf * 8 * 8), _1=BatchNormP(ngf * 8 * 8 * 2))
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)

    def forward_nhwc(self, z_code, c_rows):
        """z (B, Z) and c_code rows (B, cpad(ncf)) -> (B, 32, 32, ngf/4) NHWC."""
        b = z_code.shape[0]
        ncf = self.in_dim - cfg.GAN.Z_DIM
        cz = ops.cat_channels([c_rows.view(b, 1, 1, -1), z_code.contiguous().view(b, 1, 1, -1)],
                              [ncf, cfg.GAN.Z_DIM])
        y = self.fc[0](cz.view(b, -1))                                  # (B, 2*ngf*64)
        y = self.fc[1](y.view(b, 1, 1, -1), NA_GLU)                      # BatchNorm1d + GLU -> (B,1,1,ngf*64)
        # .view(-1, ngf, 8, 8) of the reference is an NCHW reinterpretation of the feature vector
        x = ops.to_nhwc(y.view(b, self.gf_dim, 8, 8), cpad(self.gf_dim))
        x = self.upsample1(x)
        return self.upsample2(x)

    def forward(self, z_code, c_code):
        b = z_code.shape[0]
        ncf = c_code.shape[1]
        c_rows = torch.zeros(b, cpad(ncf), device=z_code.device)
        c_rows[:, :ncf] = c_code
        return ops.to_nchw(self.forward_nhwc(z_code, c_rows), self.gf_dim // 4)


class G_HMAP(_Base):
    """ref: model.py:589-617.  keys ``conv3x3.1.{weight,bias}``, ``downsample1.0.weight``."""

    def __init__(self, ngf, ncf):
        super().__init__()
        self.gf_dim, self.in_dim = ngf, ncf
        self.conv3x3 = _Slots(_1=Conv2dP(ncf, ngf, 3, 1, 1, bias=True, mode=PAD_REFLECT))
        self.downsample1 = _Slots(_0=Conv2dP(ngf, ngf * 2, 3, 2, 1, act=ACT_LRELU))

    def forward_nhwc(self, hmap_nhwc):
        y = ops.instance_norm_act(self.conv3x3[1](hmap_nhwc), NA_LRELU)
        return self.downsample1[0](y)

    def forward(self, hmap):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(hmap)), self.gf_dim * 2)


class GlobalAttentionGeneral(_Base):
    """ATT_NET (ref: GlobalAttention.py:73-122)."""

    def __init__(self, idf, cdf):
        super().__init__()
        self.conv_context = Conv2dP(cdf, idf, 1, 1, 0)
        self.idf = idf
        self.mask = None

    def applyMask(self, mask):
        self.mask = mask

    def forward_nhwc(self, h_nhwc, context):
        src = ops.words_proj(context, self.conv_context.weight.view(self.idf, -1))
        return ops.att_general(h_nhwc, src, ops.mask_bytes(self.mask), self.idf)

    def forward(self, input, context):
        wc, attn = self.forward_nhwc(ops.to_nhwc(input), context)
        return ops.to_nchw(wc, self.idf), attn


class GlobalBUAttentionGeneral(_Base):
    """BT_ATT_NET (ref: GlobalAttention.py:125-181).  No gradient to the label / GloVe inputs (they are
    data in every caller)."""

    def __init__(self, idf, cdf):
        super().__init__()
        self.conv_context = Conv2dP(cdf, idf, 1, 1, 0)
        self.idf = idf
        self.mask = None
        self.eps = 1e-8

    def applyMask(self, mask):
        self.mask = mask

    def forward(self, input, context1, context2):
        b, e, r = input.shape[0], input.shape[1], input.shape[2] * input.shape[3]
        src = ops.words_proj(context2, self.conv_context.weight.view(self.idf, -1))
        wc, attn = ops.bu_att(input.reshape(b, e, r), context1, src, ops.mask_bytes(self.mask),
                              cfg.TRAIN.BUATTN_NORM)
        return wc.view(b, self.idf, input.shape[2], input.shape[3]), attn.view(b, -1, input.shape[2], input.shape[3])


ATT_NET = GlobalAttentionGeneral
BT_ATT_NET = GlobalBUAttentionGeneral


def pprocess_bt_attns(fmaps, ih, iw, bt_mask):
    """ref: miscc/utils.py:401-413.  Accepts the reference's expanded (B, R, num, ih, iw) mask or the compact
    (B, R, ih, iw) one; returns NCHW (B, num, ih, iw)."""
    if bt_mask.dim() == 5:
        bt_mask = bt_mask[:, :, 0]
    b, num, r = fmaps.shape[0], fmaps.shape[1], fmaps.shape[2]
    out = ops.paint_max(fmaps.reshape(b, num, r), bt_mask[:, :r].contiguous())
    return ops.to_nchw(out, num)


def _bt_branch(bt_att, slabels_feat, glove_word_embs, word_embs, mask, bt_mask, ef_dim2):
    """Shared by the INIT / NEXT stage mains (ref: model.py:552-569, 668-687)."""
    rmax = slabels_feat.shape[2]
    if rmax == 0:
        # ref: model.py:571-576 / 689-694 -- a batch without boxes contributes all-zero code / attention / label maps
        # (the reference's INIT-stage `else` reads an undefined `att`; its intent, zeros, is what the NEXT stage does)
        b, (ih, iw) = slabels_feat.shape[0], bt_mask.shape[2:]
        z = lambda c: torch.zeros((b, ih, iw, ops.cpad(c)), device=bt_mask.device, dtype=torch.float32)
        raw = torch.zeros((b, bt_att.idf, 0, 1), device=bt_mask.device, dtype=torch.float32)
        return raw, z(bt_att.idf), z(mask.shape[1]), z(ef_dim2)
    bt_att.applyMask(mask)
    bt_c_code, bt_att_map = bt_att(slabels_feat, glove_word_embs, word_embs)          # (B,idf,R,1), (B,L,R,1)
    b = bt_c_code.shape[0]
    m = bt_mask[:, :rmax].contiguous()
    code_nhwc = ops.paint_max(bt_c_code.reshape(b, -1, rmax), m)
    att_nhwc = ops.paint_max(bt_att_map.reshape(b, -1, rmax), m)
    slab_nhwc = ops.paint_max(slabels_feat.reshape(b, ef_dim2, rmax), m)
    return bt_c_code, code_nhwc, att_nhwc, slab_nhwc


class INIT_STAGE_G_MAIN(_Base):
    """ref: model.py:521-586."""

    def __init__(self, ngf, nef, nef2):
        super().__init__()
        self.gf_dim, self.ef_dim, self.ef_dim2 = ngf, nef, nef2
        self.bt_att = BT_ATT_NET(ngf, nef)
        c = ngf * 3 + nef2
        self.residual = nn.Sequential(*[HmapResBlock(c) for _ in range(cfg.GAN.GLB_R_NUM)])
        self.upsample = upBlock(c, ngf)
        _chain_res_blocks(self.residual, 0)          # the upsample conv reads the plain (halo-free) operand copies

    def forward_nhwc(self, h_code_hmap, h_code1_sent, word_embs, glove_word_embs, slabels_feat, mask, bt_mask):
        ngf = self.gf_dim
        # max_num_roi comes from slabels_feat's third dim (the trainer builds it with max(num_rois) slots,
        # miscc/utils.py:502-522), so no device->host sync on num_rois is needed here.
        _, code, _att, slab = _bt_branch(self.bt_att, slabels_feat, glove_word_embs, word_embs, mask, bt_mask,
                                         self.ef_dim2)
        x = ops.cat_channels([h_code_hmap, h_code1_sent, code, slab], [ngf, ngf, ngf, self.ef_dim2])
        x = self.residual(x)
        return self.upsample(x)

    def forward(self, h_code_hmap, h_code1_sent, c_code, word_embs, glove_word_embs, slabels_feat, mask, rois,
                num_rois, bt_mask, glb_max_num_roi):
        rmax = int(np.amax(num_rois.data.cpu().numpy()))
        out = self.forward_nhwc(ops.to_nhwc(h_code_hmap), ops.to_nhwc(h_code1_sent), word_embs, glove_word_embs,
                                slabels_feat[:, :, :rmax], mask, bt_mask)
        return ops.to_nchw(out, self.gf_dim)


class NEXT_STAGE_G_MAIN(_Base):
    """ref: model.py:620-705."""

    def __init__(self, ngf, nef, nef2):
        super().__init__()
        self.gf_dim, self.ef_dim, self.ef_dim2 = ngf, nef, nef2
        self.att = ATT_NET(ngf, nef)
        self.bt_att = BT_ATT_NET(ngf, nef)
        c = ngf * 3 + nef2
        self.residual = nn.Sequential(*[HmapResBlock(c) for _ in range(cfg.GAN.LOCAL_R_NUM)])
        self.upsample = upBlock(c, ngf)
        _chain_res_blocks(self.residual, 0)

    def forward_nhwc(self, h_code, h_code_hmap, word_embs, glove_word_embs, slabels_feat, mask, bt_mask,
                     glb_max_num_roi):
        ngf = self.gf_dim
        self.att.applyMask(mask)
        c_code, att = self.att.forward_nhwc(h_code, word_embs)
        raw, code, bt_att, slab = _bt_branch(self.bt_att, slabels_feat, glove_word_embs, word_embs, mask, bt_mask,
                                             self.ef_dim2)
        b, rmax = raw.shape[0], raw.shape[2]
        raw_full = raw.new_zeros(b, ngf, glb_max_num_roi, 1)
        raw_full[:, :, :rmax] = raw
        x = ops.cat_channels([ops.add(h_code, h_code_hmap), c_code, code, slab], [ngf, ngf, ngf, self.ef_dim2])
        x = self.residual(x)
        out = self.upsample(x)
        return out, raw_full.transpose(1, 2).squeeze(-1), att, bt_att

    def forward(self, h_code, h_code_hmap, c_code, word_embs, glove_word_embs, slabels_feat, mask, rois, num_rois,
                bt_mask, glb_max_num_roi):
        rmax = int(np.amax(num_rois.data.cpu().numpy()))
        out, raw, att, bt_att = self.forward_nhwc(ops.to_nhwc(h_code), ops.to_nhwc(h_code_hmap), word_embs,
                                                  glove_word_embs, slabels_feat[:, :, :rmax], mask, bt_mask,
                                                  glb_max_num_roi)
        L = att.shape[1]
        return ops.to_nchw(out, self.gf_dim), raw, att, ops.to_nchw(bt_att, L)


class GET_IMAGE_G(_Base):
    """ref: model.py:708-719.  key ``img.0.weight``."""

    def __init__(self, ngf):
        super().__init__()
        self.gf_dim = ngf
        self.img = _Slots(_0=Conv2dP(ngf, 3, 3, 1, 1, act=ACT_TANH))

    def forward_nhwc(self, h):
        return self.img[0](h)

    def forward(self, h_code):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(h_code)), 3)


class G_NET(_Base):
    """ref: model.py:722-795 -- same forward signature and return tuple."""

    def __init__(self, num_classes):
        super().__init__()
        ngf, nef, nef2, ncf = cfg.GAN.GF_DIM, cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.GLOVE_EMBEDDING_DIM, cfg.GAN.CONDITION_DIM
        self.ca_net = CA_NET()
        self.num_classes = num_classes
        self.branch_num = cfg.TREE.BRANCH_NUM
        if cfg.TREE.BRANCH_NUM > 0:
            self.h_net1_sent = INIT_STAGE_G(ngf * 4, ncf)
            self.h_net1_hmap = G_HMAP(ngf // 2, num_classes)
            self.h_net1_main = INIT_STAGE_G_MAIN(ngf, nef, nef2)
            self.img_net1 = GET_IMAGE_G(ngf)
        if cfg.TREE.BRANCH_NUM > 1:
            self.h_net2_hmap = G_HMAP(ngf // 2, num_classes)
            self.h_net2_main = NEXT_STAGE_G_MAIN(ngf, nef, nef2)
            self.img_net2 = GET_IMAGE_G(ngf)
        if cfg.TREE.BRANCH_NUM > 2:
            self.h_net3_hmap = G_HMAP(ngf // 2, num_classes)
            self.h_net3_main = NEXT_STAGE_G_MAIN(ngf, nef, nef2)
            self.img_net3 = GET_IMAGE_G(ngf)

    def forward(self, z_code, sent_emb, word_embs, glove_word_embs, slabels_feat, mask, hmaps, rois, fm_rois,
                num_rois, bt_masks, fm_bt_masks, glb_max_num_roi):
        fake_imgs, bt_c_codes, att_maps, bt_att_maps = [], [], [], []
        c_rows, x_rows = self.ca_net.forward_rows(sent_emb)
        d = self.ca_net.c_dim
        mu, logvar = x_rows[:, :d], x_rows[:, d:2 * d]
        self._ca_rows = x_rows            # [mu | logvar | pad] rows for the fused KL kernel (losses.KL_loss)
        words = word_embs.contiguous()
        glove = glove_word_embs.contiguous()
        L = words.shape[2]
        h = None
        if self.branch_num > 0:
            hh = self.h_net1_hmap.forward_nhwc(ops.to_nhwc(hmaps[0]))
            hs = self.h_net1_sent.forward_nhwc(z_code, c_rows)
            h = self.h_net1_main.forward_nhwc(hh, hs, words, glove, slabels_feat, mask, fm_bt_masks)
            fake_imgs.append(ops.to_nchw(self.img_net1.forward_nhwc(h), 3))
        for k in range(2, self.branch_num + 1):
            hmap_net = getattr(self, f"h_net{k}_hmap")
            main = getattr(self, f"h_net{k}_main")
            hh = hmap_net.forward_nhwc(ops.to_nhwc(hmaps[k - 1]))
            h, raw, att, bt_att = main.forward_nhwc(h, hh, words, glove, slabels_feat, mask, bt_masks[k - 2],
                                                    glb_max_num_roi)
            fake_imgs.append(ops.to_nchw(getattr(self, f"img_net{k}").forward_nhwc(h), 3))
            bt_c_codes.append(raw)
            att_maps.append(att)
            bt_att_maps.append(ops.to_nchw(bt_att, L))
        return fake_imgs, bt_c_codes, att_maps, bt_att_maps, mu, logvar


# --------------------------------------------------------------------------------------------------
# discriminators
# --------------------------------------------------------------------------------------------------
class _EncodeImage(nn.Module):
    """encode_image_by_ntimes (ref: model.py:999-1017): keys ``0.weight``, then ``{2,5,8}.weight`` convs and
    ``{3,6,9}.*`` BatchNorm2d."""

    def __init__(self, ngf, ndf, n_layer):
        super().__init__()
        self.add_module("0", Conv2dP(3 + ngf, ndf, 4, 2, 1, act=ACT_LRELU))
        self.n_layer = n_layer
        for n in range(1, n_layer):
            prev, cur = ndf * min(2 ** (n - 1), 8), ndf * min(2 ** n, 8)
            i = 2 + 3 * (n - 1)
            self.add_module(str(i), Conv2dP(prev, cur, 4, 2, 1))
            self.add_module(str(i + 1), BatchNormP(cur))

    def forward(self, x):  # NHWC
        x = getattr(self, "0")(x)
        for n in range(1, self.n_layer):
            i = 2 + 3 * (n - 1)
            x = getattr(self, str(i + 1))(getattr(self, str(i))(x), NA_LRELU)
        return x


class D_GET_LOGITS(_Base):
    """ref: model.py:1020-1048.  keys ``jointConv.0.weight``, ``jointConv.1.*``, ``outlogits.0.{weight,bias}``.
    ``forward`` takes / returns NCHW like the reference; ``forward_nhwc`` takes the NHWC feature map."""

    def __init__(self, ndf, nef, bcondition=False):
        super().__init__()
        self.df_dim, self.ef_dim, self.bcondition = ndf, nef, bcondition
        self.layer_num = cfg.GAN.LAYER_D_NUM
        c = ndf * pow(2, self.layer_num - 1)
        if bcondition:
            self.jointConv = _Slots(_0=Conv2dP(c + nef, c, 3, 1, 1), _1=BatchNormP(c))
        self.outlogits = _Slots(_0=Conv2dP(c, 1, 4, 2, 0, bias=True, act=ACT_SIGMOID))

    def forward_nhwc(self, h, c_code=None):
        if self.bcondition and c_code is not None:
            hc = ops.broadcast_cat(h, c_code.reshape(-1, self.ef_dim))
            h = self.jointConv[1](self.jointConv[0](hc), NA_LRELU)
        return ops.to_nchw(self.outlogits[0](h), 1)

    def forward(self, h_code, c_code=None):
        if isinstance(h_code, NHWCFeature):
            return self.forward_nhwc(h_code.t, c_code)
        return self.forward_nhwc(ops.to_nhwc(h_code), c_code)


class NHWCFeature:
    """Feature map handed from a D body to its logit heads without a layout round trip.  Behaves like the
    NCHW tensor the reference returns for the few things callers do with it (``size(0)``, batch slicing:
    miscc/losses.py:185-190); ``.nchw()`` materialises the reference layout."""

    def __init__(self, t, c):
        self.t, self.c = t, c

    def size(self, dim=None):
        n, h, w, _ = self.t.shape
        s = torch.Size((n, self.c, h, w))
        return s if dim is None else s[dim]

    @property
    def shape(self):
        return self.size()

    def __getitem__(self, idx):
        assert isinstance(idx, slice), "only batch slicing is supported on an NHWC feature handle"
        return NHWCFeature(self.t[idx], self.c)

    def nchw(self):
        return ops.to_nchw(self.t, self.c)


class _PatD(_Base):
    def __init__(self, b_jcu=True):
        super().__init__()
        ndf, nef = cfg.GAN.DF_DIM, cfg.TEXT.EMBEDDING_DIM
        self.img_code = _EncodeImage(0, ndf, cfg.GAN.LAYER_D_NUM)
        self.UNCOND_DNET = D_GET_LOGITS(ndf, nef, bcondition=False) if b_jcu else None
        self.COND_DNET = D_GET_LOGITS(ndf, nef, bcondition=True)
        self.out_channels = ndf * 8
        self.nhwc_features = True

    def forward(self, x_var):
        f = self.img_code(ops.to_nhwc(x_var))
        return NHWCFeature(f, self.out_channels) if self.nhwc_features else ops.to_nchw(f, self.out_channels)


class PAT_D_NET64(_PatD):
    """ref: model.py:1053-1068."""


class PAT_D_NET128(_PatD):
    """ref: model.py:1072-1087."""


class PAT_D_NET256(_PatD):
    """ref: model.py:1091-1106."""


# --------------------------------------------------------------------------------------------------
# shape discriminators (ref: model.py:1111-1179): image || shape-code(seg map) -> conv encoder; UNCOND head only
# --------------------------------------------------------------------------------------------------
class _ShpD(_Base):
    def __init__(self, num_classes):
        super().__init__()
        ndf, nef = cfg.GAN.DF_DIM, cfg.TEXT.EMBEDDING_DIM
        ngf = cfg.GAN.GF_DIM // 4
        self.img_code = _EncodeImage(ngf, ndf, cfg.GAN.LAYER_D_NUM)
        self.shp_code = _Slots(_1=Conv2dP(num_classes, ngf, 3, 1, 1, bias=True, mode=PAD_REFLECT))
        self.UNCOND_DNET = D_GET_LOGITS(ndf, nef, bcondition=False)
        self.ngf = ngf
        self.out_channels = ndf * 8
        self.nhwc_features = True

    def forward(self, x_var, s_var):
        new_s = ops.instance_norm_act(self.shp_code[1](ops.to_nhwc(s_var)), NA_LRELU)
        x_s = ops.cat_channels([ops.to_nhwc(x_var), new_s], [x_var.shape[1], self.ngf])
        f = self.img_code(x_s)
        return NHWCFeature(f, self.out_channels) if self.nhwc_features else ops.to_nchw(f, self.out_channels)


class SHP_D_NET64(_ShpD):
    """ref: model.py:1111-1132."""


class SHP_D_NET128(_ShpD):
    """ref: model.py:1135-1156."""


class SHP_D_NET256(_ShpD):
    """ref: model.py:1159-1179."""


# --------------------------------------------------------------------------------------------------
# object discriminators (ref: model.py:1184-1312): 512x512 bilinear front end, shape code, conv encoder,
# RoIAlignAvg over the 10 box slots of every image, roi code.  forward() returns (B, 10, 384, 4, 4) like the reference;
# the logit heads are D_GET_LOGITS(ndf // 2, nef) applied by objD_loss (loss glue: SURVEY 8f "next").
# --------------------------------------------------------------------------------------------------
class _ObjD(_Base):
    n_layer = 3

    def __init__(self, num_classes, b_jcu=True):
        super().__init__()
        ndf = cfg.GAN.DF_DIM
        nef = cfg.TEXT.GLOVE_EMBEDDING_DIM + cfg.GAN.GF_DIM
        ngf = cfg.GAN.GF_DIM // 4
        self.roi_size = cfg.ROI.ROI_BASE_SIZE
        self.im_scales = np.array([1])
        n_layer = self.n_layer
        self.img_code = _EncodeImage(ngf, ndf, n_layer)
        self.shp_code = _Slots(_1=Conv2dP(num_classes, ngf, 3, 1, 1, bias=True, mode=PAD_REFLECT))
        self.feat_dim = ndf * min(2 ** (n_layer - 1), 8)
        self.roi_code = _Slots(_0=Conv2dP(self.feat_dim, ndf * 4, 4, 1, 1, bias=True, act=ACT_LRELU))
        self.RoIAlignAvg = RoIAlignAvg(self.roi_size, self.roi_size, 1.0 / 16.0)
        self.UNCOND_DNET = D_GET_LOGITS(ndf // 2, nef, bcondition=False) if b_jcu else None
        self.COND_DNET = D_GET_LOGITS(ndf // 2, nef, bcondition=True)
        self.ngf = ngf

    def shape_features(self, s_var, img_size=512):
        """The shape branch (ref: model.py:1217-1219: bilinear -> ReflPad + conv3x3 80->12 + bias -> InstanceNorm ->
        LeakyReLU) of a segmentation map, NHWC.  It depends on the map and this net's weights only, so a loss that
        runs the net on several images with the SAME map (objD_loss: real and fake) computes it once and passes it
        to ``forward`` (autograd sums both uses' gradients): same values, one 80-channel 512^2 pass instead of two."""
        s = ops.cached_const(s_var, ("bilinear_nhwc", img_size),
                             lambda: ops.bilinear(ops.to_nhwc(s_var), img_size, img_size))   # same map for both object Ds
        return ops.instance_norm_act(self.shp_code[1](s), NA_LRELU)

    def forward(self, x_var, s_var, fm_rois, num_rois, img_size=512, shape_features=None):
        # (x, y, w, h) -> (x1, y1, x2, y2) on the host, on a COPY (the reference mutates a CPU caller's tensor in place,
        # ref: model.py:1213-1214; on CUDA it works on a copy too)
        fm = fm_rois.detach().cpu().numpy().astype(np.float64, copy=True)
        fm[:, :, [2, 3]] = fm[:, :, [0, 1]] + fm[:, :, [2, 3]]
        b = fm.shape[0]
        x = ops.bilinear(ops.to_nhwc(x_var), img_size, img_size)
        new_s = shape_features if shape_features is not None else self.shape_features(s_var, img_size)
        x_s = ops.cat_channels([x, new_s], [x_var.shape[1], self.ngf])
        code = self.img_code(x_s)                                           # NHWC (B, S/2^n, S/2^n, feat_dim)
        rois = _get_rois_blob(fm.reshape(b * fm.shape[1], fm.shape[2])[:, :4], np.array([1] * b * cfg.ROI.BOXES_NUM))
        rois_t = ops.h2d(rois, x_var.device)
        # the feature map is channels-last already: the channels-last RoIAlignAvg kernel (same values as the
        # reference-ABI NCHW op behind self.RoIAlignAvg) feeds roi_code without any layout round trip
        pooled = ops.roi_align_avg_nhwc(code, rois_t, self.RoIAlignAvg.aligned_height, self.RoIAlignAvg.aligned_width,
                                        self.RoIAlignAvg.spatial_scale)     # (B*10, 5, 5, C)
        pooled = self.roi_code[0](pooled)                                    # conv k4 s1 p1 + bias + LReLU
        out = ops.to_nchw(pooled, pooled.shape[3])
        return out.view(b, cfg.ROI.BOXES_NUM, out.shape[1], out.shape[2], out.shape[3])


class OBJ_SS_D_NET(_ObjD):
    """ref: model.py:1184-1246 (small-scale objects: 3 encoder layers, 64x64 feature map)."""
    n_layer = 3


class OBJ_LS_D_NET(_ObjD):
    """ref: model.py:1250-1312 (large-scale objects: 4 encoder layers, 32x32 feature map)."""
    n_layer = 4


# --------------------------------------------------------------------------------------------------
# ROIAlign modules (ref: models/roi_align/modules/roi_align.py) and the rois blob helper
# --------------------------------------------------------------------------------------------------
class RoIAlign(nn.Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super().__init__()
        self.aligned_width, self.aligned_height = int(aligned_width), int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return ops.roi_align(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)


class RoIAlignAvg(nn.Module):
    """ref: modules/roi_align.py:18-29 -- fused align (AH+1, AW+1) + avg_pool2d(2, 1) kernel."""

    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super().__init__()
        self.aligned_width, self.aligned_height = int(aligned_width), int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return ops.roi_align_avg(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)


def _get_rois_blob(im_rois, im_scale_factors):
    """ref: miscc/utils.py:365-399 (host numpy, float64 -> float32 [level, x1, y1, x2, y2])."""
    im_rois = np.asarray(im_rois, dtype=np.float64)
    n = cfg.ROI.BOXES_NUM
    levels = np.repeat(np.arange(im_rois.shape[0] // n), n).reshape(-1, 1)
    rois = im_rois * np.asarray(im_scale_factors, dtype=np.float64)[levels]
    return np.hstack((levels.astype(np.float64), rois)).astype(np.float32, copy=False)
