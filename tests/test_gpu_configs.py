"""GPU parity at the sizes BASELINE.json quotes its metric on (configs 2, 3 and 4), plus the measured basis for the
end-to-end gradient tolerance of the tensor-core engine (float64 oracle).

* config 2: full G_NET + 3 patch discriminators at batch 16, 256x256: forward images / attention maps and the
  three patD_loss values within 1e-3 of the oracle.
* config 3: grid attention (GlobalAttentionGeneral) at Q = 16384 regions, L = 18 words, batch 16 (the two-queries-
  per-thread path) and func_attention over 289 regions.
* config 4: bottom-up attention + mask paint at batch 32 with 10 boxes, RoIAlignAvg(5,5,1/16) over 320 rois x 384
  channels (bit-identical ROI index math is covered in test_gpu_parity.py; here the full config shape).
"""
import numpy as np
import pytest
import torch

from objgan_b200 import model, ops, synth, trainer
from oracle import objgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def close(a, b, tol=1e-3, what=""):
    r = rel(a, b)
    assert r <= tol, (what, r)


def _cpu_sd(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def test_config2_g_forward_and_patd_losses_b16():
    """BASELINE configs[1]: batch 16, shipped engine (f16x3).  Forward of G_NET (three images, attention maps, bottom-up
    codes) and the three discriminator losses on those images, each within 1e-3 of the CPU oracle."""
    from objgan_b200 import losses
    assert ops.CONV_ENGINE == "f16x3"          # the shipped default is what is tested
    B = 16
    t = trainer.StepATrainer(device=DEV, seed=41)
    gsd, dsds = _cpu_sd(t.netG), [_cpu_sd(d) for d in t.netsPatD]
    inp = synth.make_inputs(B, seed=42, parity=True)
    with torch.no_grad():
        ref = O.g_net_forward(gsd, inp)
        ref_err = [float(O.pat_d_loss(dsds[i], inp["imgs"][i], ref[0][i], inp["sent_emb"])) for i in range(3)]
    d = t.to_device(inp)
    with torch.no_grad():
        out = t.generate(d)
        for i in range(3):
            close(out[0][i], ref[0][i], what=f"fake{i}")
        for i in range(2):
            close(out[1][i], ref[1][i], what=f"bt_c{i}")
            close(out[2][i], ref[2][i], what=f"att{i}")
            close(out[3][i], ref[3][i], what=f"bt_att{i}")
        close(out[4], ref[4], what="mu")
        close(out[5], ref[5], what="logvar")
        for i, dnet in enumerate(t.netsPatD):
            err = float(losses.patD_loss(dnet, d["imgs"][i], out[0][i], d["sent_emb"]))
            assert abs(err - ref_err[i]) <= 1e-3 * max(1.0, abs(ref_err[i])), (i, err, ref_err[i])


@pytest.mark.parametrize("B,ih", [(16, 128), (16, 64), (32, 64)])
def test_config3_att_general_full_size(B, ih):
    """BASELINE configs[2]: Q = ih*ih regions (16384 = stage 3, 4096 = stage 2), 18 words, C = 48, ragged captions so
    the mask-row permutation quirk (GlobalAttention.py:108) is exercised at B | Q.  Forward and all three gradients."""
    L = 18
    g = torch.Generator().manual_seed(70 + B + ih)
    att = model.ATT_NET(48, 256).to(DEV)
    h = torch.randn(B, 48, ih, ih, generator=g, requires_grad=True)
    words = torch.randn(B, 256, L, generator=g, requires_grad=True)
    lens = torch.randint(3, L + 1, (B,), generator=g)
    lens[0] = L
    mask = torch.arange(L).view(1, L) >= lens.view(B, 1)
    w = att.conv_context.weight.detach().cpu().clone().requires_grad_(True)
    wc_r, a_r = O.global_attention_general(h, words, w, mask)
    gr = torch.randn(wc_r.shape, generator=g)
    wc_r.backward(gr)
    att.applyMask(mask.to(DEV))
    hg = h.detach().to(DEV).requires_grad_(True)
    wg = words.detach().to(DEV).requires_grad_(True)
    wc, a = att(hg, wg)
    close(wc, wc_r, what="wc")
    close(a, a_r, what="attn")
    wc.backward(gr.to(DEV))
    close(hg.grad, h.grad, what="g_h")
    close(att.conv_context.weight.grad, w.grad, what="g_W")
    close(wg.grad, words.grad, what="g_words")


@pytest.mark.parametrize("B", [16, 32])
def test_config3_func_attention_pairs(B):
    """BASELINE configs[2], DAMSM part: 289 regions, 256 channels, B images against one caption (what words_loss does B
    times); forward and both gradients."""
    g = torch.Generator().manual_seed(80 + B)
    q = torch.randn(B, 256, 18, generator=g, requires_grad=True)
    ctx = torch.randn(B, 256, 17, 17, generator=g, requires_grad=True)
    w_r, a_r = O.func_attention(q, ctx, 4.0)
    gw, ga = torch.randn(w_r.shape, generator=g), torch.randn(a_r.shape, generator=g)
    (w_r * gw).sum().add((a_r * ga).sum()).backward()
    qg, cg = q.detach().to(DEV).requires_grad_(True), ctx.detach().to(DEV).requires_grad_(True)
    w, a = ops.func_attention(qg, cg, 4.0)
    close(w, w_r)
    close(a, a_r)
    (w * gw.to(DEV)).sum().add((a * ga.to(DEV)).sum()).backward()
    close(qg.grad, q.grad, what="g_query")
    close(cg.grad, ctx.grad, what="g_context")


@pytest.mark.parametrize("B", [16, 32])
def test_config3_words_loss_all_pairs(B):
    """BASELINE configs[2], the B^2 = 256 / 1024 (image, caption) pairs of words_loss in ONE launch (the reference
    loops over the captions): both loss directions, the caption-wise attention maps and the gradient reaching the image
    features, ragged caption lengths passed as a DEVICE tensor (no host read), class-id mask on."""
    from objgan_b200 import losses
    g = torch.Generator().manual_seed(85 + B)
    img = torch.randn(B, 256, 17, 17, generator=g)
    words = torch.randn(B, 256, 18, generator=g)
    lens = torch.randint(4, 19, (B,), generator=g)
    lens[0] = 18
    class_ids = np.arange(B) // 2                         # pairs of captions share a class -> masked entries
    labels = torch.arange(B)
    ir = img.clone().requires_grad_(True)
    o0, o1, omaps, oacc = O.words_loss(ir, words, labels, lens.tolist(), class_ids, B)
    go = torch.autograd.grad(o0 + 2.0 * o1, ir)[0]
    ig = img.to(DEV).requires_grad_(True)
    w0, w1, maps, acc = losses.words_loss(ig, words.to(DEV), labels.to(DEV), lens.to(DEV), class_ids, B)
    close(w0, o0.detach(), what="w_loss0")
    close(w1, o1.detach(), what="w_loss1")
    assert abs(float(acc) - oacc) < 1e-4
    for a, b in zip(maps, omaps):
        close(a, b.detach(), what="att_map")
    (w0 + 2.0 * w1).backward()
    close(ig.grad, go, 2e-3, what="g_img_features")


@pytest.mark.parametrize("ih", [64, 128])
def test_config4_bu_attention_and_paint_b32(ih):
    """BASELINE configs[3]: batch 32, 10 boxes per image: bottom-up attention, then the three mask paints of a stage
    (48 code channels, 18 attention channels, 50 label channels) at 64^2 and 128^2, forward + gradient."""
    B, R, L = 32, 10, 18
    g = torch.Generator().manual_seed(90 + ih)
    bu = model.BT_ATT_NET(48, 256).to(DEV)
    lab = torch.randn(B, 50, R, 1, generator=g)
    glove, words = torch.randn(B, 50, L, generator=g), torch.randn(B, 256, L, generator=g)
    nr = torch.randint(1, R + 1, (B,), generator=g)
    for b in range(B):
        lab[b, :, int(nr[b]):] = 0
    lens = torch.randint(3, L + 1, (B,), generator=g)
    mask = torch.arange(L).view(1, L) >= lens.view(B, 1)
    w = bu.conv_context.weight.detach().cpu().clone().requires_grad_(True)
    wc_r, a_r = O.global_bu_attention(lab, glove, words, w, mask)
    boxes = torch.zeros(B, R, 4, dtype=torch.float64)
    boxes[..., :2] = torch.rand(B, R, 2, generator=g, dtype=torch.float64) * 40
    boxes[..., 2:] = 6 + torch.rand(B, R, 2, generator=g, dtype=torch.float64) * 18
    m = synth._ellipse_masks(boxes, nr, ih, ih / 64.0)
    p_r = [O.pprocess_bt_attns(x, m) for x in (wc_r, a_r, lab)]
    gs = [torch.randn(p.shape, generator=g) for p in p_r]
    (p_r[0] * gs[0]).sum().backward()
    bu.applyMask(mask.to(DEV))
    wc, a = bu(lab.to(DEV), glove.to(DEV), words.to(DEV))
    close(wc, wc_r, what="wc")
    close(a, a_r, what="attn")
    md = m.to(DEV)
    for x, want, name in zip((wc, a, lab.to(DEV)), p_r, ("code", "att", "labels")):
        close(model.pprocess_bt_attns(x, ih, ih, md), want.detach(), what=f"paint_{name}")
    model.pprocess_bt_attns(wc, ih, ih, md).backward(gs[0].to(DEV))
    close(bu.conv_context.weight.grad, w.grad, what="g_W")


@pytest.mark.parametrize("C,H", [(384, 64), (768, 32)])
def test_config4_roi_align_avg_320_rois(C, H):
    """BASELINE configs[3]: RoIAlignAvg(5, 5, 1/16) over 32 images x 10 boxes (320 rois) on the small-scale (384 ch,
    64^2) and large-scale (768 ch, 32^2) feature maps.  Index math: bit-identical to the reference's own
    roi_align_kernel.cu compiled for sm_100a (oracle/_ref) followed by avg_pool2d; against the CPU restatement the
    sample coordinate may differ by one ulp (nvcc contracts ph*bin+start to an FMA in the reference .cu as well,
    gcc does not), i.e. values by ~1e-5.  The adjoint against the float64-accumulated oracle."""
    import ctypes
    import os
    B, per = 32, 10
    g = torch.Generator().manual_seed(100 + C)
    feat = torch.randn(B, C, H, H, generator=g)
    rng = np.random.RandomState(C)
    xy = rng.uniform(0, 40 * H / 64 * 16, (B * per, 2))
    wh = rng.uniform(6 * 16 * H / 64, 24 * 16 * H / 64, (B * per, 2))
    rois = np.concatenate([np.repeat(np.arange(B), per)[:, None], xy, xy + wh], 1).astype(np.float32)
    want = O.roi_align_avg_np(feat.numpy(), rois, 5, 5, 1.0 / 16)
    fg = feat.to(DEV).requires_grad_(True)
    rd = torch.from_numpy(rois).to(DEV)
    got = model.RoIAlignAvg(5, 5, 1.0 / 16)(fg, rd)
    assert np.array_equal(got.detach().cpu().numpy() == 0, want == 0)
    close(got, torch.from_numpy(want), 2e-5, what="avg fwd vs CPU restatement")
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                      "libroi_align_ref_cuda.so")
    if os.path.exists(so):
        ref = ctypes.CDLL(so)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        ref.ROIAlignForwardLaucher.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, vp, vp, vp]
        o6 = torch.zeros(B * per, C, 6, 6, device=DEV)
        assert ref.ROIAlignForwardLaucher(fg.data_ptr(), 1.0 / 16, B * per, H, H, C, 6, 6, rd.data_ptr(), o6.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream) == 1
        torch.cuda.synchronize()
        pooled = (((o6[:, :, :-1, :-1] + o6[:, :, :-1, 1:]) + o6[:, :, 1:, :-1]) + o6[:, :, 1:, 1:]) / 4.0
        assert torch.equal(got.detach(), pooled), "fused RoIAlignAvg != reference .cu + 2x2 average, bitwise"
    gr = torch.randn(got.shape, generator=g)
    got.backward(gr.to(DEV))
    g6 = torch.zeros(rois.shape[0], C, 6, 6)
    for dh in (0, 1):
        for dw in (0, 1):
            g6[:, :, dh:dh + 5, dw:dw + 5] += gr / 4
    want_g = O.roi_align_backward_np(g6.numpy(), rois, feat.shape, 6, 6, 1.0 / 16)
    close(fg.grad, torch.from_numpy(want_g), 2e-5, what="avg bwd")


def _to64(x):
    if torch.is_tensor(x):
        return x.double() if x.dtype == torch.float32 else x
    if isinstance(x, (list, tuple)):
        return [_to64(t) for t in x]
    if isinstance(x, dict):
        return {k: _to64(v) for k, v in x.items()}
    return x


def _dist(a, b):
    return (a.detach().cpu().double() - b).norm().item()


def test_gradient_error_vs_float64_oracle(monkeypatch):
    """The measured basis of the end-to-end gradient tolerances (replaces the argument in DESIGN.md section 4).

    The generator gradient of (patch-D terms + KL) through the whole G -> 3 x D -> BCE stack, and the discriminator
    gradient of patD_loss, from the SAME weights (B = 2), computed four ways: the oracle in float64 (g64: exact to
    ~1e-12), the oracle in float32 on the CPU (what the reference computes), the CUDA path with exact-fp32 FMA
    contractions ("simt") and the shipped tensor-core path ("f16x3").  The G -> D -> BCE chain amplifies rounding by
    10^3..10^4 (LeakyReLU / max / tiny-batch BatchNorm), so even float32 on the CPU sits 3e-4 (relative L2, one
    thread count) .. 1e-2 (another) from g64; the GPU's fp32 FMA chains over K <= 9216 terms round ~10x more than
    oneDNN's blocked sums.  Asserted: both CUDA engines stay within 1e-2 of g64 overall (2e-2 per tensor), and the
    3xFP16 tensor-core engine is no further from g64 than 3x the exact-fp32 CUDA engine (+ 1e-3) -- i.e. the operand
    split adds nothing beyond ordinary fp32 summation-order noise.  Measured on a B200 (round 2): G gradient cpu-fp32
    2.7e-4, simt 2.8e-3, f16x3 6.3e-3; PatD128 gradient cpu-fp32 1.4e-6, simt 6.4e-5, f16x3 5.2e-4.  The measured numbers are printed (and quoted in DESIGN.md).

    Not covered here, and measured in test_step_a_parity: inside a full step the generator's gradient is taken through
    discriminators that have just taken their first Adam step, which is sign descent -- entries whose gradient is
    rounding noise move by +-lr depending on that noise, on the CPU as on the GPU."""
    from objgan_b200 import losses
    t = trainer.StepATrainer(device=DEV, seed=21)
    gsd, dsds = _cpu_sd(t.netG), [_cpu_sd(d) for d in t.netsPatD]
    inp = synth.make_inputs(2, seed=33, parity=True)

    def oracle(gsd_, dsds_, inp_):
        g_live, g_leaves = O._with_grad(gsd_, O.trainable_keys(gsd_))
        fake, _b, _a, _ba, mu, logvar = O.g_net_forward(g_live, inp_)
        total = O.g_loss_pat(dsds_, fake, inp_["sent_emb"], update=False) + O.kl_loss(mu, logvar)
        keys = list(g_leaves)
        gg = dict(zip(keys, torch.autograd.grad(total, [g_leaves[k] for k in keys])))
        d_live, d_leaves = O._with_grad(dsds_[1], O.trainable_keys(dsds_[1]))
        err = O.pat_d_loss(d_live, inp_["imgs"][1], fake[1].detach(), inp_["sent_emb"])
        dkeys = list(d_leaves)
        dg = dict(zip(dkeys, torch.autograd.grad(err, [d_leaves[k] for k in dkeys])))
        return gg, dg

    clone = lambda sd: {k: v.clone() for k, v in sd.items()}
    gg32, dg32 = oracle(clone(gsd), [clone(d) for d in dsds], inp)
    gg64, dg64 = oracle(_to64(gsd), _to64(dsds), _to64(inp))
    dev = t.to_device(inp)
    buf0 = [[b.clone() for b in m.buffers()] for m in [t.netG, *t.netsPatD]]
    res = {}
    for engine in ("simt", "f16x3"):
        monkeypatch.setattr(ops, "CONV_ENGINE", engine)
        for m, saved in zip([t.netG, *t.netsPatD], buf0):          # same BatchNorm buffers for both engines
            for b, s0 in zip(m.buffers(), saved):
                b.copy_(s0)
        t.bG.zero_grad()
        t.bG.requires_grad_(True)
        for b in t.bD:
            b.zero_grad()
            b.requires_grad_(False)
        fake, _b, _a, _ba, mu, logvar = t.generate(dev)
        (losses.G_loss_pat(t.netsPatD, fake, dev["sent_emb"])[0] + losses.KL_loss(mu, logvar)).backward()
        t.bD[1].zero_grad()
        t.bD[1].requires_grad_(True)
        losses.patD_loss(t.netsPatD[1], dev["imgs"][1], fake[1].detach(), dev["sent_emb"]).backward()
        for name, got, g32, g64 in (("G", dict(t.netG.named_parameters()), gg32, gg64),
                                    ("PatD128", dict(t.netsPatD[1].named_parameters()), dg32, dg64)):
            tot_gpu = tot_cpu = tot_ref = 0.0
            worst = (0.0, "")
            for k, ref in g64.items():
                if k.endswith("conv3x3.1.bias"):
                    continue                         # bias ahead of InstanceNorm: the exact gradient is zero
                n64, d_cpu, d_gpu = ref.norm().item(), _dist(g32[k], ref), _dist(got[k].grad, ref)
                tot_gpu, tot_cpu, tot_ref = tot_gpu + d_gpu ** 2, tot_cpu + d_cpu ** 2, tot_ref + n64 ** 2
                worst = max(worst, (d_gpu / n64, k))
            res[(engine, name)] = ((tot_gpu / tot_ref) ** 0.5, (tot_cpu / tot_ref) ** 0.5, worst)
            print("%-6s %-8s gradient, rel L2 distance to the float64 gradient: cuda %.3e  cpu-fp32 %.3e  worst tensor "
                  "%.3e (%s)" % (engine, name, *res[(engine, name)][:2], *worst))
    for name in ("G", "PatD128"):
        simt, f16 = res[("simt", name)], res[("f16x3", name)]
        assert simt[0] <= 1e-2 and f16[0] <= 1e-2, (name, simt, f16)
        assert simt[2][0] <= 2e-2 and f16[2][0] <= 2e-2, (name, simt, f16)
        assert f16[0] <= 3.0 * simt[0] + 1e-3, (name, simt, f16)


def _dist(a, b):
    return (a.detach().cpu().double() - b).norm().item()
