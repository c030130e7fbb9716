"""CPU tests: the oracle against the committed golden fixtures (produced by the reference's own code via
tests/golden/make_golden.py), the three ROIAlign restatements against each other, and the C ABI surface."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from objgan_b200 import lib, model, synth
from oracle import objgan_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402  (only its seed constants and build_weights are used here)


def _close(a, b, tol=1e-4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


def test_g_forward_golden():
    gold = np.load(os.path.join(GOLD, "g_forward.npz"))
    g_sd, _ = make_golden.build_weights()
    inp = synth.make_inputs(2, seed=make_golden.SEED_IN, parity=True)
    inp["eps"] = torch.from_numpy(gold["eps"])
    sd = {k: v.clone() for k, v in g_sd.items()}
    with torch.no_grad():
        fake, btc, att, btatt, mu, logvar = O.g_net_forward(sd, inp)
    _close(fake[0].numpy(), gold["fake64"])
    _close(fake[1][..., ::4, ::4].numpy(), gold["fake128"])
    _close(fake[2][..., ::8, ::8].numpy(), gold["fake256"])
    _close(att[0][..., ::4, ::4].numpy(), gold["att1"])
    _close(att[1][..., ::8, ::8].numpy(), gold["att2"])
    _close(btatt[0][..., ::4, ::4].numpy(), gold["bt_att1"])
    _close(btc[0].numpy(), gold["bt_c1"])
    _close(btc[1].numpy(), gold["bt_c2"])
    _close(mu.numpy(), gold["mu"])
    _close(logvar.numpy(), gold["logvar"])
    _close(sd["h_net3_main.upsample.2.running_mean"].numpy(), gold["bn_rm"])


def test_pat_d_loss_golden():
    gold = np.load(os.path.join(GOLD, "pat_d_loss.npz"))
    _, d_sds = make_golden.build_weights()
    sd = {k: v.clone() for k, v in d_sds[0].items()}
    gen = torch.Generator().manual_seed(7)
    real = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    fk = torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1
    cond = torch.rand(4, 256, generator=gen)
    keys = O.trainable_keys(sd)
    assert keys == list(gold["names"])
    live, leaves = O._with_grad(sd, keys)
    err = O.pat_d_loss(live, real, fk, cond)
    grads = torch.autograd.grad(err, [leaves[k] for k in keys])
    assert abs(float(err) - float(gold["err"])) < 1e-5
    _close(np.array([g.norm().item() for g in grads]), gold["grad_norms"], 1e-3)


@pytest.mark.parametrize("cls,n_layer,large,key", [("OBJ_SS_D_NET", 3, False, "err_ss"), ("OBJ_LS_D_NET", 4, True, "err_ls")])
def test_obj_d_loss_golden(cls, n_layer, large, key):
    """oracle.obj_d_loss (body + feat_select + permute_seg + heads + weights) against the value the reference's own
    objD_loss produced on its own sub-modules and roi_align.c (fixture made by make_golden.py)."""
    import random
    gold = np.load(os.path.join(GOLD, "obj_d_loss.npz"))
    sd, real, fake, seg, fm, nr, raw_cond, raw_bt = make_golden.obj_d_case(None, cls)
    with torch.no_grad():
        random.seed(21)
        got = O.obj_d_loss(sd, real, fake, seg, raw_cond, raw_bt, fm.numpy(), nr.tolist(), n_layer,
                           is_large_scale=large, update=False)
    want = float(gold[key])
    assert want > 0
    assert abs(float(got) - want) <= 2e-5 * max(1.0, abs(want)), (float(got), want)


def test_g_loss_golden():
    """oracle.g_loss (all patch / shape / object discriminator terms + DAMSM terms) against the total and the
    gradients the reference's own G_loss produced (fixture made by make_golden.py)."""
    gold = np.load(os.path.join(GOLD, "g_loss.npz"))
    nets, inp, fakes, raw_bt, rois0, fm, class_ids = make_golden.g_loss_case()
    sd = lambda n: {k: v.clone() for k, v in n.state_dict().items()}
    fk = [f.clone().requires_grad_(True) for f in fakes]
    bt = raw_bt.clone().requires_grad_(True)
    total, terms = O.g_loss([sd(n) for n in nets["pat"]], [sd(n) for n in nets["shp"]], sd(nets["ss"]), sd(nets["ls"]),
                            make_golden.StubEncoder(), fk, inp["hmaps"], inp["words_embs"], inp["sent_emb"],
                            inp["clabels_emb"], bt, torch.arange(2), inp["cap_lens"].tolist(), class_ids, rois0.numpy(),
                            fm.numpy(), inp["num_rois"].tolist(), update=False)
    assert {"objss_g_loss", "objls_g_loss", "w_loss", "s_loss"} <= set(terms)        # every branch is exercised
    want = float(gold["total"])
    assert abs(float(total) - want) <= 1e-5 * max(1.0, abs(want)), (float(total), want)
    grads = torch.autograd.grad(total, fk + [bt])
    _close(grads[0].numpy(), gold["g64"], 1e-4)
    _close(grads[1][..., ::2, ::2].numpy(), gold["g128"], 1e-4)
    _close(grads[2][..., ::4, ::4].numpy(), gold["g256"], 1e-4)
    _close(grads[3].numpy(), gold["gbt"], 1e-4)


def test_attention_golden():
    gold = np.load(os.path.join(GOLD, "attention.npz"))
    W = torch.randn(48, 256, 1, 1, generator=torch.Generator().manual_seed(1)) * 0.1
    gen = torch.Generator().manual_seed(9)
    h, words = torch.randn(3, 48, 8, 8, generator=gen), torch.randn(3, 256, 18, generator=gen)
    mask = torch.arange(18).view(1, 18) >= torch.tensor([18, 11, 6]).view(3, 1)
    wc, a = O.global_attention_general(h, words, W, mask)
    _close(wc.numpy(), gold["wc"])
    _close(a.numpy(), gold["attn"])
    q, ctx = torch.randn(3, 256, 14, generator=gen), torch.randn(3, 256, 17, 17, generator=gen)
    fw, fa = O.func_attention(q, ctx, 4.0)
    _close(fw.numpy(), gold["func_wc"])
    _close(fa.numpy(), gold["func_attn"])


def _c_oracle():
    so = os.path.join(ROOT, "oracle", "_ref", "libroi_align_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libroi_align_oracle.so"])
    return ctypes.CDLL(so)


def test_roi_align_restatements_bit_exact():
    """numpy oracle == C oracle == golden output of the reference's roi_align.c, bit for bit."""
    gold = np.load(os.path.join(GOLD, "roi_align.npz"))
    feat, rois, want = gold["feat"], gold["rois"], gold["out"]
    got_np = O.roi_align_forward_np(feat, rois, 6, 6, 1.0 / 16)
    assert np.array_equal(got_np, want)
    c = _c_oracle()
    fp = ctypes.POINTER(ctypes.c_float)
    out = np.zeros_like(want)
    c.og_oracle_roi_align_forward(feat.ctypes.data_as(fp), ctypes.c_float(1.0 / 16), rois.shape[0], 16, 16, 5, 6, 6,
                                  rois.ctypes.data_as(fp), out.ctypes.data_as(fp))
    assert np.array_equal(out, want)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libroi_align_ref_cpu.so")
    if os.path.exists(ref_so):  # the reference source compiled verbatim, when it was built here
        r = ctypes.CDLL(ref_so)
        rng = np.random.RandomState(11)
        feat2 = rng.randn(3, 4, 32, 32).astype(np.float32)
        xy = rng.uniform(-8, 480, (30, 2))
        rois2 = np.hstack((np.repeat(np.arange(3), 10).reshape(-1, 1), xy, xy + rng.uniform(0, 200, (30, 2))))
        rois2 = rois2.astype(np.float32)
        a = np.zeros((30, 4, 6, 6), dtype=np.float32)
        b = np.zeros_like(a)
        r.ROIAlignForwardCpu(feat2.ctypes.data_as(fp), ctypes.c_float(1.0 / 16), 30, 32, 32, 4, 6, 6,
                             rois2.ctypes.data_as(fp), a.ctypes.data_as(fp))
        c.og_oracle_roi_align_forward(feat2.ctypes.data_as(fp), ctypes.c_float(1.0 / 16), 30, 32, 32, 4, 6, 6,
                                      rois2.ctypes.data_as(fp), b.ctypes.data_as(fp))
        assert np.array_equal(a, b)
        assert np.array_equal(O.roi_align_forward_np(feat2, rois2, 6, 6, 1.0 / 16), a)
    # backward restatements agree (numpy vs C, both float64 accumulation)
    g = np.random.RandomState(2).randn(*want.shape).astype(np.float32)
    gb_np = O.roi_align_backward_np(g, rois, feat.shape, 6, 6, 1.0 / 16)
    gb_c = np.zeros(feat.shape, dtype=np.float64)
    c.og_oracle_roi_align_backward(g.ctypes.data_as(fp), ctypes.c_float(1.0 / 16), rois.shape[0], 16, 16, 5, 6, 6,
                                   rois.ctypes.data_as(fp), gb_c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    np.testing.assert_allclose(gb_c.astype(np.float32), gb_np, rtol=1e-6, atol=1e-7)


def test_get_rois_blob():
    fm = np.zeros((2, 10, 6))
    fm[:, :, :4] = np.random.RandomState(0).uniform(0, 30, (2, 10, 4))
    blob = O.get_rois_blob_np(fm)
    assert blob.shape == (20, 5) and blob.dtype == np.float32
    assert (blob[:10, 0] == 0).all() and (blob[10:, 0] == 1).all()
    np.testing.assert_allclose(blob[:, 3], (fm[:, :, 0] + fm[:, :, 2]).reshape(-1).astype(np.float32))
    # host helper of the product mirrors miscc/utils.py:365-399
    xyxy = fm.copy()
    xyxy[:, :, 2:4] += xyxy[:, :, 0:2]
    blob2 = model._get_rois_blob(xyxy.reshape(20, 6)[:, :4], np.array([1] * 20))
    assert np.array_equal(blob, blob2)


def test_cabi_exports_every_declared_symbol():
    """The shared library loads (no GPU needed) and exports everything include/objgan_b200.h declares."""
    L = lib.get()
    protos = lib.parse_header()
    assert len(protos) >= 37
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = sorted(set(protos) - exported)
    assert not missing, missing
    for name in protos:
        assert name in L.fn
    # the reference's FFI names are present verbatim (roi_align_kernel.h:13-27)
    assert {"ROIAlignForwardLaucher", "ROIAlignBackwardLaucher"} <= exported


def test_no_cpu_fallback():
    """CPU tensors are rejected (the product path has no CPU or library fallback)."""
    from objgan_b200 import ops
    with pytest.raises(RuntimeError):
        ops.to_nhwc(torch.zeros(1, 8, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "obj-gan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_step_b_oracle_composition():
    """oracle.step_b is the composition of the individually pinned pieces: its first discriminator losses equal
    pat_d_loss / shp_d_loss / obj_d_loss evaluated directly on the initial weights with the same shuffles, and its
    generator loss equals g_loss on the updated discriminators."""
    import random
    from objgan_b200 import model
    model.FAST_INIT = True
    try:
        torch.manual_seed(0)
        nets = [model.G_NET(80), model.PAT_D_NET64(), model.PAT_D_NET128(), model.PAT_D_NET256(), model.SHP_D_NET64(80),
                model.SHP_D_NET128(80), model.SHP_D_NET256(80), model.OBJ_SS_D_NET(80), model.OBJ_LS_D_NET(80)]
    finally:
        model.FAST_INIT = False
    for n in nets:
        for p in n.parameters():
            if p.dim() > 1:
                p.data.normal_(0.0, 1.0 / np.sqrt(p[0].numel()))
    sd = lambda n: {k: v.clone() for k, v in n.state_dict().items()}
    st = O.StepBState(sd(nets[0]), [sd(n) for n in nets[1:4]], [sd(n) for n in nets[4:7]], sd(nets[7]), sd(nets[8]))
    inp = make_golden.synth.make_inputs(2, seed=4, parity=True)
    with torch.no_grad():
        fake = O.g_net_forward(sd(nets[0]), inp)[0]
        random.seed(5)
        want_pat = float(O.pat_d_loss(sd(nets[1]), inp["imgs"][0], fake[0], inp["sent_emb"]))
        want_shp = float(O.shp_d_loss(sd(nets[4]), inp["imgs"][0], fake[0], inp["hmaps"][0], inp["rois"][0].numpy(),
                                      inp["num_rois"].tolist()))
    random.seed(5)
    out = O.step_b(st, inp)
    assert abs(out["errPatD0"] - want_pat) < 1e-5 and abs(out["errShpD0"] - want_shp) < 1e-5
    assert st.step == 1 and st.x_step[:3] == [1, 1, 1]
    assert all(np.isfinite(v) for k, v in out.items() if isinstance(v, float))
    total = sum(out["terms"].values())
    assert abs(total - out["errG"]) < 1e-4 * max(1.0, abs(out["errG"]))


def test_roi_align_edge_cases_bit_exact():
    """Edge cases against the reference's roi_align.c compiled verbatim (oracle/_ref, build container): boxes partly or
    entirely outside the map, inverted and zero-area boxes, sub-pixel boxes, a box covering the whole map, an empty roi
    list -- numpy and C restatements must agree with it bit for bit (ref: roi_align.c:80-150)."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libroi_align_ref_cpu.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (the reference sources are absent)")
    r, c = ctypes.CDLL(ref_so), _c_oracle()
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(5)
    C, H, W = 3, 12, 20
    feat = rng.standard_normal((2, C, H, W)).astype(np.float32)
    s = 16.0
    boxes = [
        [0, -40.0, -40.0, 60.0, 70.0],                 # hangs over the top-left corner
        [1, 250.0, 150.0, 400.0, 300.0],               # hangs over the bottom-right corner
        [0, 500.0, 500.0, 600.0, 600.0],               # entirely outside
        [1, -300.0, -300.0, -100.0, -100.0],           # entirely outside (negative)
        [0, 100.0, 80.0, 40.0, 20.0],                  # inverted (x2 < x1, y2 < y1)
        [1, 64.0, 64.0, 64.0, 64.0],                   # zero area
        [0, 33.3, 17.7, 33.9, 18.1],                   # sub-pixel
        [1, 0.0, 0.0, (W - 1) * s, (H - 1) * s],       # the whole map, corners exactly on the last samples
        [0, 0.0, 0.0, W * s, H * s],                   # one cell past the map
    ]
    rois = np.asarray(boxes, dtype=np.float32)
    for ah, aw in ((6, 6), (3, 5), (2, 2)):
        want = np.zeros((len(boxes), C, ah, aw), dtype=np.float32)
        r.ROIAlignForwardCpu(feat.ctypes.data_as(fp), ctypes.c_float(1.0 / s), len(boxes), H, W, C, ah, aw,
                             rois.ctypes.data_as(fp), want.ctypes.data_as(fp))
        assert np.array_equal(O.roi_align_forward_np(feat, rois, ah, aw, 1.0 / s), want), (ah, aw)
        out = np.zeros_like(want)
        c.og_oracle_roi_align_forward(feat.ctypes.data_as(fp), ctypes.c_float(1.0 / s), len(boxes), H, W, C, ah, aw,
                                      rois.ctypes.data_as(fp), out.ctypes.data_as(fp))
        assert np.array_equal(out, want), (ah, aw)
    empty = np.zeros((0, 5), dtype=np.float32)
    assert O.roi_align_forward_np(feat, empty, 6, 6, 1.0 / s).shape == (0, C, 6, 6)


def test_step_a_dp_one_shard_is_step_a():
    """oracle.step_a_dp (mean of the shards' gradients) with a single shard is oracle.step_a."""
    from objgan_b200 import lib, model, synth
    lib.DRY_RUN, model.FAST_INIT = True, False
    try:
        torch.manual_seed(5)
        g = model.G_NET(80)
        ds = [model.PAT_D_NET64(), model.PAT_D_NET128(), model.PAT_D_NET256()]
        for m in [g, *ds]:
            m.apply(model.weights_init)
    finally:
        lib.DRY_RUN = False
    sd = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
    inp = synth.make_inputs(2, seed=14, parity=True)
    a, b = O.StepAState(sd(g), [sd(d) for d in ds]), O.StepAState(sd(g), [sd(d) for d in ds])
    ka, kb = {}, {}
    la = O.step_a(a, inp, keep=ka)
    lb = O.step_a_dp(b, [inp], keep=kb)[0]
    for k in la:
        assert abs(la[k] - lb[k]) <= 1e-6 * max(1.0, abs(la[k])), k
    for k in a.g_keys:
        assert torch.allclose(ka["g_grads"][k], kb["g_grads"][k], rtol=1e-4, atol=1e-7), k
        assert torch.allclose(a.g[k], b.g[k], rtol=0, atol=4.1e-4), k


def test_torch_extension_registers_reference_entry_points():
    """The thin PyTorch C++ extension (north_star's boundary): torch.ops.objgan_b200 exposes roi_align_forward_cuda /
    roi_align_backward_cuda with the reference's argument order (roi_align_cuda.h:1-5) on the CUDA dispatch key; no
    compute here (no GPU) -- CPU tensors must be rejected, not silently handled."""
    import objgan_b200.torch_ext as ext
    ops_ = ext.load()
    for name in ("roi_align_forward_cuda", "roi_align_backward_cuda"):
        schema = getattr(ops_, name).default._schema
        assert [a.name for a in schema.arguments][:3] == ["aligned_height", "aligned_width", "spatial_scale"]
        assert len(schema.arguments) == 6
    with pytest.raises(Exception):
        ops_.roi_align_forward_cuda(6, 6, 1 / 16, torch.zeros(1, 2, 8, 8), torch.zeros(1, 5), torch.zeros(1, 2, 6, 6))


def test_hi_lo_operand_split_arithmetic():
    """The integer form of the fp16 hi / lo operand split used by attention_tc.cu (at_hilo: hi = (bits + 0x1000) &
    0xFFFFE000, lo = x - hi) restated in numpy: hi has 11 significant bits (exactly an fp16 once scaled into the fp16
    range), x - hi is exact in fp32, |lo| <= 2^-11 |x|, and hi + fp16(lo) reproduces x to 2^-21 relative -- the
    22-bit operand the three-product MMA scheme assumes.  Also the algebra of the stacked-weight form used for narrow
    layers: a_lo [w_hi | w_lo] + a_hi [w_hi | w_lo] = the three-product sum + a_lo w_lo."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * 2.0 ** rng.integers(-6, 13, 200000)).astype(np.float32)   # |x| < 2^14 after scaling
    bits = x.view(np.uint32)
    hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = x - hi
    normal = np.abs(x) >= 2.0 ** -14                                             # fp16 normal range
    assert np.array_equal(hi[normal].astype(np.float16).astype(np.float32), hi[normal])   # hi is an fp16 value there
    assert np.array_equal(hi.astype(np.float64) + lo.astype(np.float64), x.astype(np.float64))   # the remainder is exact
    assert np.all(np.abs(lo) <= np.abs(x) * 2.0 ** -11 * (1 + 1e-6))
    rec = hi.astype(np.float64) + lo.astype(np.float16).astype(np.float64)
    big = np.abs(x) >= 2.0 ** -3                                                 # lo stays a normal fp16 there
    assert np.max(np.abs(rec[big] - x[big]) / np.abs(x[big])) <= 2.0 ** -21
    a_hi, a_lo, w_hi, w_lo = (rng.standard_normal(64) for _ in range(4))
    three = a_lo * w_hi + a_hi * w_lo + a_hi * w_hi
    stacked = (a_lo * w_hi + a_hi * w_hi) + (a_lo * w_lo + a_hi * w_lo)
    assert np.allclose(stacked, three + a_lo * w_lo)
