"""Generates tests/golden/*.npz by running the REFERENCE's own code (imported read-only from /root/reference,
so this script only works in the build container).  The fixtures pin oracle/objgan_oracle.py and the CUDA path
on machines where the reference is absent (the GPU box).

Weights are not stored: both sides rebuild them from a seed with objgan_b200.model's own constructors (CPU) and
load them into the reference modules with load_state_dict(strict=True).

    python tests/golden/make_golden.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import refimport  # noqa: E402
from objgan_b200 import model, synth  # noqa: E402

SEED_W, SEED_IN = 101, 202


def build_weights(seed=SEED_W):
    torch.manual_seed(seed)
    g = model.G_NET(80)
    g.apply(model.weights_init)
    ds = [model.PAT_D_NET64(), model.PAT_D_NET128(), model.PAT_D_NET256()]
    for d in ds:
        d.apply(model.weights_init)
    return g.state_dict(), [d.state_dict() for d in ds]


def sub(t, s):
    return t[..., ::s, ::s].contiguous().numpy()


class StubEncoder(torch.nn.Module):
    """Differentiable stand-in for the pretrained Inception CNN_ENCODER: (regions (B,256,17,17), code (B,256))."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(77)
        self.weight = torch.nn.Parameter(torch.randn(256, 3, 1, 1, generator=g) * 0.5)
        self.bias = torch.nn.Parameter(torch.randn(256, generator=g) * 0.1)

    def forward(self, x):
        r = torch.nn.functional.conv2d(torch.nn.functional.adaptive_avg_pool2d(x, 17), self.weight, self.bias)
        return r, r.mean((2, 3))


def g_loss_case(seed=404):
    """Seeded full-G_loss case: the eight discriminators (objgan_b200 constructors on the CPU) and the inputs."""
    torch.manual_seed(seed)
    nets = dict(pat=[model.PAT_D_NET64(), model.PAT_D_NET128(), model.PAT_D_NET256()],
                shp=[model.SHP_D_NET64(80), model.SHP_D_NET128(80), model.SHP_D_NET256(80)],
                ss=model.OBJ_SS_D_NET(80), ls=model.OBJ_LS_D_NET(80))
    for n in nets["pat"] + nets["shp"] + [nets["ss"], nets["ls"]]:
        for p_ in n.parameters():          # cheap, seeded, well-scaled init (weights_init's QR is slow at this size)
            if p_.dim() > 1:
                p_.data.normal_(0.0, 1.0 / np.sqrt(p_[0].numel()))
    inp = synth.make_inputs(2, seed=seed + 1, parity=True)
    g = torch.Generator().manual_seed(seed + 2)
    fakes = [torch.tanh(torch.randn(im.shape, generator=g)) for im in inp["imgs"]]
    raw_bt = torch.randn(2, 10, 48, generator=g)
    fm = inp["fm_rois"].clone()
    fm[..., 2:4] *= torch.tensor([3.0, 0.6]).view(2, 1, 1)
    return nets, inp, fakes, raw_bt, inp["rois"][0].clone(), fm, np.array([4, 9])


def obj_d_case(ref, cls, seed=303):
    """Seeded object-discriminator case shared by the fixture generator (ref given: returns the reference net with the
    weights loaded) and the golden test (ref None: returns the state_dict)."""
    torch.manual_seed(seed)
    mine = getattr(model, cls)(80)
    mine.apply(model.weights_init)
    sd = {k: v.clone() for k, v in mine.state_dict().items()}
    inp = synth.make_inputs(3, seed=seed + 1, parity=True)
    real, seg, fm, nr = inp["imgs"][2], inp["hmaps"][2], inp["fm_rois"].clone(), inp["num_rois"]
    fm[..., 2:4] *= torch.tensor([1.0, 3.0, 0.2]).view(3, 1, 1)       # boxes on both sides of the size threshold
    gen = torch.Generator().manual_seed(seed + 2)
    fake = torch.tanh(torch.randn(real.shape, generator=gen))
    raw_bt = torch.randn(3, 10, mine.COND_DNET.ef_dim - inp["clabels_emb"].shape[1], generator=gen)
    if ref is None:
        return sd, real, fake, seg, fm, nr, inp["clabels_emb"], raw_bt
    net = getattr(ref.model, cls)(80)
    net.load_state_dict(sd, strict=True)
    return net, real, fake, seg, fm, nr, inp["clabels_emb"], raw_bt


def main():
    ref = refimport.load()
    g_sd, d_sds = build_weights()
    # ---- G_NET forward, B=2, ragged captions / roi counts ------------------------------------
    g = ref.model.G_NET(80)
    g.load_state_dict(g_sd, strict=True)
    inp = synth.make_inputs(2, seed=SEED_IN, parity=True)
    torch.manual_seed(0)
    out = g(inp["z"], inp["sent_emb"], inp["words_embs"], inp["glove_words_embs"], inp["slabels_feat"], inp["mask"],
            inp["hmaps"], inp["rois"], inp["fm_rois"], inp["num_rois"], inp["bt_masks"], inp["fm_bt_masks"],
            inp["glb_max_num_roi"])
    torch.manual_seed(0)
    eps = torch.FloatTensor(2, 100).normal_()     # the draw CA_NET.reparametrize made (model.py:473-477)
    fake = [f.detach() for f in out[0]]
    np.savez_compressed(os.path.join(HERE, "g_forward.npz"), eps=eps.numpy(), fake64=fake[0].numpy(),
                        fake128=sub(fake[1], 4), fake256=sub(fake[2], 8), att1=sub(out[2][0].detach(), 4),
                        att2=sub(out[2][1].detach(), 8), bt_att1=sub(out[3][0].detach(), 4),
                        bt_c1=out[1][0].detach().numpy(), bt_c2=out[1][1].detach().numpy(), mu=out[4].detach().numpy(),
                        logvar=out[5].detach().numpy(),
                        bn_rm=g.state_dict()["h_net3_main.upsample.2.running_mean"].numpy())
    # ---- patD_loss value + gradient norms ------------------------------------------------------
    d = ref.model.PAT_D_NET64()
    d.load_state_dict(d_sds[0], strict=True)
    gen = torch.Generator().manual_seed(7)
    real = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    fk = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    cond = torch.rand(4, 256, generator=gen)
    err = ref.losses.patD_loss(d, real, fk, cond)
    err.backward()
    names = [n for n, _ in d.named_parameters()]
    np.savez_compressed(os.path.join(HERE, "pat_d_loss.npz"), err=float(err),
                        grad_norms=np.array([p.grad.norm().item() for p in d.parameters()]),
                        grad_first=np.stack([p.grad.reshape(-1)[:4].numpy() for p in d.parameters()
                                             if p.numel() >= 4]),
                        names=np.array(names))
    # ---- attention modules -----------------------------------------------------------------------
    torch.manual_seed(5)
    att = ref.GlobalAttention.GlobalAttentionGeneral(48, 256)
    W = torch.randn(48, 256, 1, 1, generator=torch.Generator().manual_seed(1)) * 0.1
    att.conv_context.weight.data.copy_(W)
    gen = torch.Generator().manual_seed(9)
    h, words = torch.randn(3, 48, 8, 8, generator=gen), torch.randn(3, 256, 18, generator=gen)
    mask = torch.arange(18).view(1, 18) >= torch.tensor([18, 11, 6]).view(3, 1)
    att.applyMask(mask)
    wc, a = att(h, words)
    q, ctx = torch.randn(3, 256, 14, generator=gen), torch.randn(3, 256, 17, 17, generator=gen)
    fw, fa = ref.GlobalAttention.func_attention(q, ctx, 4.0)
    np.savez_compressed(os.path.join(HERE, "attention.npz"), wc=wc.detach().numpy(), attn=a.detach().numpy(),
                        func_wc=fw.numpy(), func_attn=fa.numpy())
    # ---- ROIAlign from the reference C source compiled verbatim (oracle/_ref) -----------------------
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libroi_align_ref_cpu.so"))
    rng = np.random.RandomState(3)
    feat = rng.randn(2, 5, 16, 16).astype(np.float32)
    xy = rng.uniform(-4, 230, (12, 2))
    wh = rng.uniform(1, 128, (12, 2))
    rois = np.hstack((np.repeat(np.arange(2), 6).reshape(-1, 1), xy, xy + wh)).astype(np.float32)
    outp = np.zeros((12, 5, 6, 6), dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ROIAlignForwardCpu(feat.ctypes.data_as(fp), ctypes.c_float(1.0 / 16), 12, 16, 16, 5, 6, 6,
                           rois.ctypes.data_as(fp), outp.ctypes.data_as(fp))
    np.savez_compressed(os.path.join(HERE, "roi_align.npz"), feat=feat, rois=rois, out=outp)
    # ---- objD_loss: the reference's own loss / feat_select / permute_seg code (miscc/losses.py:254-361) on a net whose
    # body is built from the reference's sub-modules and its roi_align.c (model.py:1227-1241 uses a Variable/resize_
    # idiom that no longer runs; it is replaced by the three lines it stands for) ----------------------------------
    import random
    import torch.nn.functional as F
    res = {}
    for cls, n_layer, large in (("OBJ_SS_D_NET", 3, False), ("OBJ_LS_D_NET", 4, True)):
        net, real, fake, seg, fm, nr, raw_cond, raw_bt = obj_d_case(ref, cls)

        class Net:
            COND_DNET, UNCOND_DNET = net.COND_DNET, net.UNCOND_DNET

            def __call__(self, x, s, f, n):
                x5 = F.interpolate(x, size=(512, 512), mode="bilinear", align_corners=True)
                s5 = F.interpolate(s, size=(512, 512), mode="bilinear", align_corners=True)
                code = net.img_code(torch.cat([x5, net.shp_code(s5)], 1))
                fmn = f.numpy().copy()
                fmn[:, :, [2, 3]] = fmn[:, :, [0, 1]] + fmn[:, :, [2, 3]]
                nroi = fmn.shape[0] * fmn.shape[1]
                rois_ = ref.utils._get_rois_blob(fmn.reshape(nroi, fmn.shape[2])[:, :4], np.array([1] * nroi))
                featc = np.ascontiguousarray(code.detach().numpy())
                c, hw = featc.shape[1], featc.shape[2]
                o6 = np.zeros((nroi, c, 6, 6), dtype=np.float32)
                lib.ROIAlignForwardCpu(featc.ctypes.data_as(fp), ctypes.c_float(1 / 16), nroi, hw, hw, c, 6, 6,
                                       np.ascontiguousarray(rois_).ctypes.data_as(fp), o6.ctypes.data_as(fp))
                out_ = net.roi_code(F.avg_pool2d(torch.from_numpy(o6), 2, 1))
                return out_.view(fmn.shape[0], fmn.shape[1], out_.size(1), out_.size(2), out_.size(3))

        with torch.no_grad():
            random.seed(21)
            err = ref.losses.objD_loss(Net(), real, fake, seg, raw_cond, raw_bt, fm, nr, is_large_scale=large)
        res[cls] = float(err)
    np.savez_compressed(os.path.join(HERE, "obj_d_loss.npz"), err_ss=res["OBJ_SS_D_NET"], err_ls=res["OBJ_LS_D_NET"])
    # ---- full G_loss (miscc/losses.py:364-531): the reference's function, patch / shape discriminators and logit heads;
    # the object discriminators' bodies go through the (pinned, differentiable) oracle.obj_d_net_forward ------------
    from oracle import objgan_oracle as O
    torch.ByteTensor = lambda a: torch.from_numpy(np.asarray(a)).bool()      # torch 2: masked_fill_ wants bool masks
    nets, inp, fakes, raw_bt, rois0, fm, class_ids = g_loss_case()
    rnets = dict(pat=[ref.model.PAT_D_NET64(), ref.model.PAT_D_NET128(), ref.model.PAT_D_NET256()],
                 shp=[ref.model.SHP_D_NET64(80), ref.model.SHP_D_NET128(80), ref.model.SHP_D_NET256(80)],
                 ss=ref.model.OBJ_SS_D_NET(80), ls=ref.model.OBJ_LS_D_NET(80))
    for k in ("pat", "shp"):
        for a_, b_ in zip(rnets[k], nets[k]):
            a_.load_state_dict(b_.state_dict(), strict=True)
    rnets["ss"].load_state_dict(nets["ss"].state_dict(), strict=True)
    rnets["ls"].load_state_dict(nets["ls"].state_dict(), strict=True)

    def stand_in(net, n_layer):
        sdict = {k: v.clone() for k, v in net.state_dict().items()}

        class Net:
            COND_DNET, UNCOND_DNET = net.COND_DNET, net.UNCOND_DNET

            def __call__(self, x, s, f, n):
                return O.obj_d_net_forward(sdict, x, s, f.numpy(), n_layer, update=False)
        return Net()

    fk = [f.clone().requires_grad_(True) for f in fakes]
    bt = raw_bt.clone().requires_grad_(True)
    total, _ = ref.losses.G_loss(rnets["pat"], rnets["shp"], stand_in(rnets["ss"], 3), stand_in(rnets["ls"], 4),
                                 StubEncoder(), fk, inp["hmaps"], inp["words_embs"], inp["sent_emb"], inp["clabels_emb"],
                                 bt, torch.arange(2), inp["cap_lens"], class_ids, rois0, fm, inp["num_rois"])
    grads = torch.autograd.grad(total, fk + [bt])
    np.savez_compressed(os.path.join(HERE, "g_loss.npz"), total=float(total), g64=grads[0].numpy(),
                        g128=sub(grads[1], 2), g256=sub(grads[2], 4), gbt=grads[3].numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
