"""GPU tests of the tcgen05 tensor-core convolution path (conv_tc.cu) against torch-CPU fp32 convolutions
(the oracle's arithmetic primitive) at sizes large enough for the dispatcher to pick the tensor cores.

Tolerances: 3xFP16 (the parity mode: 22-bit hi/lo operands) must land at fp32 rounding level (<= 1e-4 of max|ref|
asserted, ~1e-6 observed); a single fp16 product is only checked loosely (it is not a parity mode)."""
import pytest
import torch
import torch.nn.functional as F

from objgan_b200 import model, ops
from objgan_b200.lib import ACT_NONE, PAD_REFLECT, PAD_ZERO, UPSAMPLE2X

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # cin, cout, mode, split, N, H, W   (3x3 stride 1 pad 1 unless mode is the string "s2": 4x4 stride 2 pad 1)
    (194, 388, PAD_REFLECT, 194, 2, 64, 64),     # HmapResBlock conv1 (two N tiles of 208)
    (194, 194, PAD_REFLECT, 0, 4, 32, 64),       # HmapResBlock conv2
    (194, 96, UPSAMPLE2X, 48, 2, 64, 64),        # upBlock (four phases)
    (1024, 768, PAD_ZERO, 0, 36, 16, 16),        # jointConv @16x16 (three N tiles of 256)
    (64, 48, PAD_ZERO, 0, 2, 80, 72),            # ragged tiles: H, W not multiples of the patch
    (40, 24, PAD_REFLECT, 0, 5, 48, 48),         # C not a multiple of 32, small N tile
    (1024, 768, PAD_ZERO, 0, 8, 16, 16),         # jointConv on 16x16 maps (wgrad chunk = 16 x 2 pixels)
    (96, 192, "s2", 0, 4, 64, 64),               # discriminator conv4x4 s2 (space-to-depth phases)
    (192, 384, "s2", 0, 8, 32, 32),              # ... on 16x16 outputs
    (384, 768, "s2", 0, 32, 16, 16),             # ... on 8x8 outputs (tiles span 2 images)
    (48, 3, PAD_ZERO, 0, 2, 64, 64),             # GET_IMAGE_G: 3 output channels (dgrad reads an 8-channel source)
    (3, 96, "s2", 0, 4, 64, 64),                 # discriminator first layer: 3 (->8) input channels
    (80, 24, PAD_REFLECT, 0, 2, 64, 64),         # G_HMAP conv3x3
    (768, 1536, "s2", 0, 32, 8, 8),              # 4x4 outputs: wgrad pixel patch spans 2 images
    (40, 24, PAD_REFLECT, 0, 4, 16, 16),         # reflection halo on a 16-wide map (wgrad patch = 16 x 2)
    (24, 48, "s2k3", 0, 4, 64, 64),              # conv3x3 stride 2 (heat-map encoder) through the same phase blocks
    (384, 384, "k4s1", 0, 130, 5, 5),            # roi_code of the object discriminators: 4x4 stride 1 pad 1 on 5x5 pooled rois
    (384, 384, "k4s1", 0, 70, 5, 5),             # ... roi count that does not fill the last pixel tile / wgrad image group
    (192, 384, "s2", 0, 13, 16, 16),             # sub-batch of the permuted-shape pass: 13 images, tiles span 2 images
    (384, 768, "s2", 0, 19, 8, 8),               # ... 4x4 outputs, 8 images per tile
    (96, 48, UPSAMPLE2X, 24, 3, 8, 8),           # upsample phases with a batch that does not fill the 2-image tile
    (80, 12, PAD_REFLECT, 0, 2, 128, 128),       # shape code of the object discriminators: 12 (-> 16) output channels
    (15, 96, "s2", 0, 2, 128, 128),              # their first encoder layer: the input gradient has 15 (-> 16) channels
]


def _ref(x, w, mode):
    if mode in ("s2", "s2k3"):
        return F.conv2d(x, w, None, 2, 1)
    if mode == "k4s1":
        return F.conv2d(x, w, None, 1, 1)
    if mode == PAD_REFLECT:
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    if mode == UPSAMPLE2X:
        return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, 1, 1)
    return F.conv2d(x, w, None, 1, 1)


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert torch.isfinite(a).all()
    return (a - b).abs().max().item() / b.abs().max().item()


@pytest.mark.parametrize("engine,tol", [("f16x3", 1e-4), ("f16", 4e-3)])
@pytest.mark.parametrize("case", CASES)
def test_tc_conv_fwd_dgrad(case, engine, tol, monkeypatch):
    cin, cout, mode, split, N, H, W = case
    monkeypatch.setattr(ops, "CONV_ENGINE", engine)
    torch.manual_seed(cin + cout + H)
    if mode in ("s2", "s2k3"):
        m = model.Conv2dP(cin, cout, 4 if mode == "s2" else 3, 2, 1).to(DEV)
    elif mode == "k4s1":
        m = model.Conv2dP(cin, cout, 4, 1, 1).to(DEV)
    else:
        m = model.Conv2dP(cin, cout, 3, 1, 1, mode=mode, split=split).to(DEV)
    x = torch.randn(N, cin, H, W)
    w = m.weight.detach().cpu().clone()
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = _ref(xr, wr, mode)
    gy = torch.randn_like(yr)
    gxr, gwr = torch.autograd.grad(yr, (xr, wr), gy)
    xg = x.to(DEV).requires_grad_(True)
    k, st = (4, 2) if mode == "s2" else (3, 2) if mode == "s2k3" else (4, 1) if mode == "k4s1" else (3, 1)
    assert ops._tc_kind(N, H, W, ops.cpad(cin), k, k, st, 1, PAD_ZERO if (st == 2 or mode == "k4s1") else mode)
    y_nhwc = m(ops.to_nhwc(xg))
    if split:
        sp = ops.cpad(split)
        y = torch.cat((y_nhwc[..., :split], y_nhwc[..., sp:sp + split]), -1).permute(0, 3, 1, 2)
        assert (y_nhwc[..., split:sp] == 0).all() and (y_nhwc[..., sp + split:] == 0).all()
        gfull = torch.zeros_like(y_nhwc)
        g_nhwc = gy.permute(0, 2, 3, 1).to(DEV)
        gfull[..., :split] = g_nhwc[..., :split]
        gfull[..., sp:sp + split] = g_nhwc[..., split:]
        assert _rel(y, yr) <= tol, ("fwd", _rel(y, yr))
        y_nhwc.backward(gfull)
    else:
        y = ops.to_nchw(y_nhwc, cout)
        assert _rel(y, yr) <= tol, ("fwd", _rel(y, yr))
        y.backward(gy.to(DEV))
    assert _rel(xg.grad, gxr) <= tol, ("dgrad", _rel(xg.grad, gxr))
    assert _rel(m.weight.grad, gwr) <= tol, ("wgrad", _rel(m.weight.grad, gwr))


def test_prep_split_exact():
    """(hi + lo) / 2^k reconstructs x to 2^-21 (relative, or 2^-38 of max|x| for tiny elements), the scale puts
    max|x| in [2^13, 2^14), the reflection halo matches F.pad, the space-to-depth blocks are the four phases."""
    x = torch.randn(2, 8, 6, 8, device=DEV) * 3
    x[0, 0, 0, 0] = 1e-7          # far below the maximum: only absolute accuracy is promised
    xn = ops.to_nhwc(x)
    monkey = ops.CONV_ENGINE
    ops.CONV_ENGINE = "f16x3"
    try:
        hi, lo, am = ops._split(xn, 1)
        shi, slo, sam = ops._split(xn, s2d=True)
    finally:
        ops.CONV_ENGINE = monkey
    amax = x.abs().max().item()
    assert am.view(torch.float32).item() == amax and sam.view(torch.float32).item() == amax
    import math
    k = 13 - math.floor(math.log2(amax))
    assert 2.0 ** 13 <= hi.float().abs().max().item() < 2.0 ** 14
    want = F.pad(x, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)

    def close(got, ref):
        tol = torch.maximum(ref.abs() * 2.0 ** -21, torch.full_like(ref, amax * 2.0 ** -38))
        return ((got - ref).abs() <= tol).all()

    rec = (hi.double() + lo.double()) * 2.0 ** -k
    assert close(rec, want.double())
    srec = (shi.double() + slo.double()) * 2.0 ** -k
    for a in range(2):
        for b in range(2):
            blk = srec[(a * 2 + b) * 2:(a * 2 + b + 1) * 2]
            assert close(blk, x[:, :, a::2, b::2].permute(0, 2, 3, 1).double())


def test_two_tile_kernels_on_small_cases():
    """The two-tile-per-CTA kernels (conv_tc2 / wgrad2) are normally picked only for >= 592 tiles; force them on the
    small parity cases through OG_TC2_MIN in a fresh process (the library reads the variable once)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OG_TC2_MIN="4")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-k",
                        "test_tc_conv_fwd_dgrad and f16x3 and (case0 or case1 or case3 or case6)"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_stacked_hi_lo_weights_on_narrow_layers():
    """Layers with <= 16 output channels run the CTA-pair kernel with B = [w_hi | w_lo] stacked along N and two MMAs per
    k-step (each A copy read once) instead of three (the default; the normal run of these cases covers it).  Here the
    three-MMA form (OG_STACKED=0) is checked against the same parity bound in a fresh process (the library reads the
    variable once)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OG_STACKED="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-x", "-k",
                        "test_tc_conv_fwd_dgrad and f16x3 and (case10 or case12 or case21 or case22)"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_cta_pair_kernels_on_small_cases():
    """The cta_group::2 variant of the two-tile kernel (conv_tc2x: clusters of two CTAs issue one M = 256 MMA, each
    CTA holding half of every weight stage) on the parity cases that reach the two-tile path, forced through
    OG_TC2X=1 / OG_TC2_MIN in a fresh process (the library reads the variables once).  Odd tile-pair counts and
    ragged tiles included (case 4 has 5 x 5 tiles per image)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OG_TC2_MIN="4", OG_TC2X="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-x", "-k",
                        "test_tc_conv_fwd_dgrad and (case0 or case1 or case3 or case6)"],
                       env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_persistent_cta_pair_kernel_on_small_cases():
    """conv_tcp (persistent clusters of two CTAs, cta_group::2 MMAs, double-buffered TMEM accumulators) on the parity
    cases with 208- / 256-wide output tiles, forced through OG_TCP=1 / OG_TCP_MIN=1 in a fresh process; includes an
    odd tile count, more work units than clusters (several rounds through both accumulator stages) and fewer."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OG_TCP="1", OG_TCP_MIN="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-x", "-k",
                        "test_tc_conv_fwd_dgrad and (case0 or case1 or case3 or case6 or case13)"],
                       env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_persistent_cta_pair_kernel_all_tile_widths():
    """conv_tcp for EVERY output-tile width (32 / 64 / 112 / 208 / 256: OG_TCP=2) on all parity cases of this file --
    upsample phases (strided output pixels), space-to-depth stride-2 taps, ragged tiles, tiles spanning images."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OG_TCP="2", OG_TCP_MIN="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-x", "-k",
                        "test_tc_conv_fwd_dgrad and f16x3"],
                       env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("engine", ["f16x3", "f16"])
def test_norm_apply_emits_operand_copies(engine, monkeypatch):
    """A chain of residual blocks + upBlock with the producer-side operand split (og_norm_apply_split: the normalisation
    pass writes the next convolution's fp16 hi / lo copies, reflection halo included, scaled by an a-priori bound) against
    the same chain with separate og_norm_apply + og_prep_split passes: outputs and all gradients agree to fp32
    rounding level (the two differ only in the power of two the operands are scaled by)."""
    monkeypatch.setattr(ops, "CONV_ENGINE", engine)
    torch.manual_seed(3)
    c = 194
    blocks = torch.nn.Sequential(*[model.HmapResBlock(c) for _ in range(3)]).to(DEV)
    up = model.upBlock(c, 48).to(DEV)
    model._chain_res_blocks(blocks, 0)
    x0 = torch.randn(2, 32, 32, ops.cpad(c), device=DEV)
    x0[..., c:] = 0
    res = {}
    for fused in (False, True, False):     # the first pass also packs the weights: only the last two are compared / counted
        monkeypatch.setattr(ops, "FUSED_SPLIT", fused)
        calls0 = ops._lib.get().launches
        x = x0.clone().requires_grad_(True)
        for p_ in list(blocks.parameters()) + list(up.parameters()):
            p_.grad = None
        out = up(blocks(x))
        gout = torch.ones_like(out) * torch.linspace(-1, 1, out.shape[-1], device=DEV)
        out.backward(gout)
        res[fused] = (out.detach().clone(), x.grad.clone(), [p_.grad.clone() for p_ in blocks.parameters()],
                      ops._lib.get().launches - calls0)
    tol = 1e-5 if engine == "f16x3" else 4e-3
    assert _rel(res[True][0], res[False][0]) <= tol
    assert _rel(res[True][1], res[False][1]) <= 10 * tol
    for a, b in zip(res[True][2], res[False][2]):
        assert _rel(a, b) <= 10 * tol
