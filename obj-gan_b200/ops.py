"""Autograd-aware Python operators over the C ABI of libobjgan_b200.so.

PyTorch is plumbing here: device memory (``torch.empty``), the autograd tape and streams.  Every
arithmetic step is a kernel of this repo's library, reached through ``lib.call`` (ctypes); there is no
torch / cuDNN / cuBLAS compute path and no CPU fallback -- CPU tensors raise.

Internal activation layout: NHWC fp32, shape (N, H, W, Cp) with Cp = channels rounded up to 8 and the
pad lanes kept at exactly zero.
"""
from __future__ import annotations

import torch

from . import lib as _lib
from .lib import (ACT_LRELU, ACT_NONE, ACT_SIGMOID, ACT_TANH, NA_GLU, NA_LRELU, NA_NONE, PAD_REFLECT, PAD_ZERO,
                  TRANSPOSED, UPSAMPLE2X)

LRELU_SLOPE = 0.2
NORM_EPS = 1e-5
BN_MOMENTUM = 0.1


def cpad(c: int) -> int:
    return (c + 7) // 8 * 8


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda and not _lib.DRY_RUN:
            raise RuntimeError("objgan_b200 kernels need CUDA tensors (no CPU fallback)")
        if t.dtype not in (torch.float32, torch.uint8, torch.float64, torch.int64):
            raise RuntimeError(f"unsupported dtype {t.dtype}")


def _p(t):
    return 0 if t is None else t.data_ptr()


def _call(name, *args):
    if _lib.DRY_RUN:  # host-logic tracing only (tests): kernels are NOT executed, outputs stay uninitialised
        lib = _lib.get()
        want = len(lib.protos[name]) - 1          # every prototype ends with the stream
        if len(args) != want:
            raise TypeError(f"{name}: {len(args)} arguments passed, the header declares {want} (+ stream)")
        lib.launches += 1
        return
    _lib.get().call(name, *args, _lib.stream())


# weights change only through the optimiser / load_state_dict; both bump this epoch so cached packed copies
# are rebuilt.  In-place torch ops on a Parameter are caught through Tensor._version.
_param_epoch = [0]


def bump_param_epoch():
    _param_epoch[0] += 1


def _epoch_of(w):
    """(global epoch, epoch of the network that owns w): a network's optimiser step only invalidates ITS operand
    copies (a trainer attaches ``og_epoch`` = its bucket's counter to every parameter)."""
    own = getattr(w, "og_epoch", None)
    return (_param_epoch[0], own[0] if own is not None else 0)


class PackedWeights:
    """Cache of kernel-native copies (fprop and dgrad operand) of one OIHW parameter."""

    def __init__(self):
        self.key = None
        self.f = None
        self.t = None
        self.hl = {}
        self.amax = None

    def _refresh(self, w, cip, kp, split):
        key = (w.data_ptr(), w._version, _epoch_of(w), cip, kp, split)
        if key != self.key:
            self.key, self.f, self.t, self.hl, self.amax = key, None, None, {}, None

    def _amax(self, w):
        """Device word with the float bits of max|w| (the power-of-two scale of the fp16 operand copies)."""
        if self.amax is None:
            self.amax = torch.empty(1, device=w.device, dtype=torch.int32)
            _call("og_amax", _p(w), w.numel(), _p(self.amax))
        return self.amax

    def get(self, w, cip, kp, split, splitp, need_t):
        self._refresh(w, cip, kp, split)
        co, ci, kh, kw = w.shape
        if self.f is None:
            self.f = torch.empty(kh * kw * cip * kp, device=w.device, dtype=torch.float32)
            _call("og_pack_weights", _p(w), co, ci, kh, kw, cip, kp, split, splitp, 0, _p(self.f))
        if need_t and self.t is None:
            self.t = torch.empty(kh * kw * cip * kp, device=w.device, dtype=torch.float32)
            _call("og_pack_weights", _p(w), co, ci, kh, kw, cip, kp, split, splitp, 1, _p(self.t))
        return self.f, self.t

    def get_up_hilo(self, w, cip, kp, split, splitp, transposed):
        """Pre-summed 2x2 phase weights of an upsample+conv3x3 layer as (fp16 hi, fp16 lo, scale word), laid out
        [16][co][ci] (transposed) or [16][ci][co]."""
        self._refresh(w, cip, kp, split)
        key = ("up", transposed)
        if key not in self.hl:
            co, ci, _, _ = w.shape
            hi = torch.empty(16 * cip * kp, device=w.device, dtype=torch.float16)
            lo = torch.empty_like(hi) if _nsplit() == 3 else None
            amax_up = torch.empty(1, device=w.device, dtype=torch.int32)
            _call("og_pack_upsample_weights", _p(w), co, ci, cip, kp, split, splitp, transposed, _p(self._amax(w)),
                  _p(amax_up), _p(hi), _p(lo))
            self.hl[key] = (hi, lo, amax_up)
        return self.hl[key]

    def get_hilo(self, w, cip, kp, split, splitp, transposed):
        """(fp16 hi, fp16 lo, scale word) of the packed matrix ([tap][co][ci] when transposed else [tap][ci][co])."""
        self._refresh(w, cip, kp, split)
        if transposed not in self.hl:
            co, ci, kh, kw = w.shape
            # hi and lo live back to back in ONE buffer: the CTA-pair kernel can then address them as the two halves of
            # an [hi | lo] operand stacked along N (narrow layers, conv_tc.cu: TcParams::stacked)
            nel = kh * kw * cip * kp
            buf = torch.empty(nel * (2 if _nsplit() == 3 else 1), device=w.device, dtype=torch.float16)
            hi = buf[:nel]
            lo = buf[nel:] if _nsplit() == 3 else None
            _call("og_pack_weights_f16", _p(w), co, ci, kh, kw, cip, kp, split, splitp, transposed,
                  _p(self._amax(w)), _p(hi), _p(lo))
            self.hl[transposed] = (hi, lo, self._amax(w))
            plan = getattr(w, "og_pack_plan", None)
            if plan is not None:            # the owner re-packs all of its weights in one launch after its Adam step
                plan.record(self, w, cip, kp, split, splitp, transposed, hi, lo, self._amax(w))
        return self.hl[transposed]


class PackPlan:
    """All (weight, layout) fp16 operand copies of ONE network, re-derived by two launches (og_amax_multi +
    og_pack_weights_f16_multi) right after the network's optimiser step instead of lazily, layer by layer, at the next
    use (~300 tiny launches per training step).  Jobs are recorded the first time a layer asks for a copy; from then on
    ``run`` refreshes the SAME buffers and re-validates the layers' caches."""

    def __init__(self):
        self.jobs = []              # (cache, w, cip, kp, split, splitp, transposed, hi, lo, amax)
        self._tables = None
        self._n_built = 0

    def record(self, cache, w, cip, kp, split, splitp, transposed, hi, lo, amax):
        job = (cache, w, cip, kp, split, splitp, transposed, hi, lo, amax)
        for i, j in enumerate(self.jobs):
            if j[0] is cache and j[6] == transposed:      # same layer and layout packed again: new buffers
                self.jobs[i] = job
                self._tables = None
                return
        self.jobs.append(job)

    def _build(self, device):
        CH = 4096
        arows, prows, ab, pb, seen = [], [], 0, 0, {}
        for cache, w, cip, kp, split, splitp, tr, hi, lo, amax in self.jobs:
            if id(amax) not in seen:
                seen[id(amax)] = True
                arows.append([w.data_ptr(), w.numel(), amax.data_ptr(), ab])
                ab += (w.numel() + CH - 1) // CH
            co, ci, kh, kw = w.shape
            total = kh * kw * cip * kp
            prows.append([w.data_ptr(), hi.data_ptr(), 0 if lo is None else lo.data_ptr(), amax.data_ptr(), co, ci,
                          kh * kw, cip, kp, split, splitp, tr, total, pb])
            pb += (total + CH - 1) // CH
        mk = lambda rows: h2d(rows, device, torch.int64)
        self._tables = (mk(arows), len(arows), ab, mk(prows), len(prows), pb)
        self._n_built = len(self.jobs)

    def run(self):
        """Call after the weights changed (and after the parameter epoch was bumped)."""
        if not self.jobs or _lib.DRY_RUN:
            return
        if self._tables is None or self._n_built != len(self.jobs):
            self._build(self.jobs[0][1].device)
        at, an, ab, pt, pn, pb = self._tables
        _call("og_amax_multi", _p(at), an, ab)
        _call("og_pack_weights_f16_multi", _p(pt), pn, pb)
        touched = {}
        for cache, w, cip, kp, split, splitp, tr, hi, lo, amax in self.jobs:
            if id(cache) not in touched:
                touched[id(cache)] = cache
                cache.key = (w.data_ptr(), w._version, _epoch_of(w), cip, kp, split)
                cache.f = cache.t = None
                cache.hl = {}
                cache.amax = amax
            cache.hl[tr] = (hi, lo, amax)


# Contraction engine for the convolutions: "f16x3" = tcgen05 tensor cores with the error-compensated
# three-product fp16 hi/lo split (22-bit operands, fp32-level accuracy: the default and the parity mode), "f16" = one fp16 product
# (faster, outside the 1e-3 parity budget), "simt" = exact fp32 FMA on the CUDA cores.
import os as _os
CONV_ENGINE = _os.environ.get("OBJGAN_CONV", "f16x3")
TC_WGRAD = _os.environ.get("OBJGAN_TC_WGRAD", "1") == "1"
FUSED_AMAX = _os.environ.get("OBJGAN_FUSED_AMAX", "1") == "1"   # producers leave max|x| for the consuming conv
KEEP_SPLIT = _os.environ.get("OBJGAN_KEEP_SPLIT", "1") == "1"   # keep x's hi/lo copies from forward for the wgrad
FUSED_SPLIT = _os.environ.get("OBJGAN_FUSED_SPLIT", "1") == "1"  # norm-apply emits the consumer conv's hi/lo operand copies
TC_MIN_PIXELS = 256        # smaller problems go to the exact-fp32 SIMT kernels (the tc kernel splits K on small maps)


def _int_array(rows):
    import ctypes
    flat = [int(v) for r in rows for v in r]
    return (ctypes.c_int * len(flat))(*flat)


def _empty_slack(shape, device):
    """fp16 buffer with 512 bytes of readable slack after the last element."""
    n = 1
    for d in shape:
        n *= d
    flat = torch.empty(n + 256, device=device, dtype=torch.float16)
    return flat[:n].view(shape)


def _nsplit():
    return 3 if CONV_ENGINE == "f16x3" else 1


def _split(x, pad=0, s2d=False):
    """(fp16 hi, fp16 lo, scale word) of an NHWC tensor scaled by a power of two from its max|x|; pad=1 adds the
    reflection halo; s2d gives [4N, H/2, W/2, C]."""
    n, h, w, c = x.shape
    ready_made = getattr(x, "og_split", None)          # the producer already wrote them (og_norm_apply_split)
    if ready_made is not None and ready_made[3:] == (pad, bool(s2d), x._version, _nsplit()):
        return ready_made[:3]
    if getattr(x, "og_f32_invalid", False):
        raise RuntimeError("this activation exists only as tensor-core operand copies (pad=%s); a consumer asked for "
                           "another layout" % (ready_made[3] if ready_made else None))
    shape = (4 * n, h // 2, w // 2, c) if s2d else (n, h + 2 * pad, w + 2 * pad, c)
    xh = _empty_slack(shape, x.device)
    xl = _empty_slack(shape, x.device) if CONV_ENGINE == "f16x3" else None
    amax = _known_amax(x)
    ready = amax is not None
    if not ready:
        amax = torch.empty(1, device=x.device, dtype=torch.int32)
    _call("og_prep_split", _p(x), n, h, w, c, pad, 1 if s2d else 0, _p(amax), 1 if ready else 0, _p(xh), _p(xl))
    return xh, xl, amax


def _tag_amax(t):
    """Attach a fresh max|t| word to a tensor some kernel is about to produce (the kernel fills it); a consuming
    convolution then skips its own amax pass.  The tag is void once the tensor is modified in place."""
    word = torch.empty(1, device=t.device, dtype=torch.int32)
    t.og_amax = (word, t._version)
    return word


def _known_amax(t):
    tag = getattr(t, "og_amax", None)
    if tag is not None and tag[1] == t._version and FUSED_AMAX:
        return tag[0]
    return None


def _tc_launch(xs, n, ws, ntaps_w, kw_rows, y, oh, ow, k, osy, op, taps, bias=None, act=ACT_NONE, layout=0):
    """xs / ws: (hi, lo, scale word) operand triples; taps: (dh, dw, dn, widx) quadruples; the source is
    [SN, SH, SW, C] with SN = n or 4n (space-to-depth)."""
    xh, xl, ax = xs
    wh, wl, aw = ws
    sn, sh, sw, c = xh.shape
    _, ohf, owf, kc = y.shape
    arr = _int_array(taps)
    import ctypes
    _call("og_conv2d_tc", _p(xh), _p(xl), _p(ax), n, sn, sh, sw, c, _p(wh), _p(wl), _p(aw), ntaps_w, kw_rows, _p(y),
          oh, ow, k,
          ohf * owf * kc, owf * kc, kc, ohf, owf, osy, osy, op[0], op[1], ctypes.addressof(arr), len(taps), layout,
          _nsplit(), _p(bias), act, LRELU_SLOPE)


def _tile_n(oh, ow):
    tw = 1
    while tw * 2 <= ow and tw * 2 <= 16:
        tw *= 2
    th = 1
    while th * 2 <= oh and tw * th * 2 <= 128:
        th *= 2
    return 128 // (tw * th)


def _tc_kind(n, h, w, c, kh, kw, stride, pad, mode):
    """Which tensor-core formulation (if any) applies to this convolution.  The batch need not fill the last pixel
    tile: GEMM rows are independent and rows past the batch are never stored."""
    if CONV_ENGINE not in ("f16x3", "f16") or _lib.DRY_RUN:
        return None
    if kh == kw and stride == 1 and pad == 1 and (kh == 3 or (kh == 4 and mode == PAD_ZERO)):
        # 3x3 (any padding mode), and the 4x4 stride-1 zero-padded roi_code layer of the object discriminators
        oh, ow = _out_hw(h, w, kh, kw, stride, pad, mode)
        if n * oh * ow < TC_MIN_PIXELS:
            return None
        return "s1"
    if kh == kw and kh in (3, 4) and stride == 2 and pad == 1 and mode == PAD_ZERO and h % 2 == 0 and w % 2 == 0:
        if n * (h // 2) * (w // 2) < TC_MIN_PIXELS:
            return None
        return "s2"
    return None


_UP_OFF = ((-1, 0), (0, 1))   # low-res row/col offsets read by output phase 0 / 1 of upsample + conv3x3


def _tc_fprop(kind, x, cache, weight, kp, split, splitp, mode, bias_p, act):
    """Returns (y, xs): the output and the fp16 operand triple of x (reused by the weight gradient)."""
    n, h, w, c = x.shape
    dev = x.device
    if mode != UPSAMPLE2X:
        ws = cache.get_hilo(weight, c, kp, split, splitp, 1)      # [tap][co][ci]
    if kind == "s2":                                 # 4x4 / 3x3 stride 2 pad 1 on space-to-depth phases
        k = weight.shape[2]
        xs = _split(x, s2d=True)
        y = torch.empty((n, h // 2, w // 2, kp), device=dev, dtype=torch.float32)
        taps = []
        for kh in range(k):
            dh, a = divmod(kh - 1, 2)
            for kw in range(k):
                dw, b = divmod(kw - 1, 2)
                taps.append((dh, dw, (a * 2 + b) * n, kh * k + kw))
        _tc_launch(xs, n, ws, k * k, kp, y, h // 2, w // 2, kp, 1, (0, 0), taps, bias_p, act)
        return y, xs
    if mode == PAD_REFLECT:
        xs = _split(x, 1)
        y = torch.empty((n, h, w, kp), device=dev, dtype=torch.float32)
        taps = [(kh, kw, 0, kh * 3 + kw) for kw in range(3) for kh in range(3)]      # column by column (layout 1)
        _tc_launch(xs, n, ws, 9, kp, y, h, w, kp, 1, (0, 0), taps, bias_p, act, layout=1)
    elif mode == PAD_ZERO:
        k = weight.shape[2]
        xs = _split(x, 0)
        oh, ow = h + 3 - k, w + 3 - k
        y = torch.empty((n, oh, ow, kp), device=dev, dtype=torch.float32)
        if k == 3:
            taps = [(kh - 1, kw - 1, 0, kh * 3 + kw) for kw in range(3) for kh in range(3)]
        else:
            taps = [(kh - 1, kw - 1, 0, kh * k + kw) for kh in range(k) for kw in range(k)]
        _tc_launch(xs, n, ws, k * k, kp, y, oh, ow, kp, 1, (0, 0), taps, bias_p, act, layout=1 if k == 3 else 0)
    else:  # UPSAMPLE2X: four output phases of 2x2 taps at low resolution, weights pre-summed per phase
        ws = cache.get_up_hilo(weight, c, kp, split, splitp, 1)
        xs = _split(x, 0)
        y = torch.empty((n, 2 * h, 2 * w, kp), device=dev, dtype=torch.float32)
        for py in range(2):
            for px in range(2):
                taps = [(_UP_OFF[py][a], _UP_OFF[px][b], 0, ((py * 2 + px) * 2 + a) * 2 + b)
                        for a in range(2) for b in range(2)]
                _tc_launch(xs, n, ws, 16, kp, y, h, w, kp, 2, (py, px), taps, bias_p, act)
    return y, xs


def _split_x(kind, x, mode):
    """The operand copies _tc_fprop makes of x."""
    if kind == "s2":
        return _split(x, s2d=True)
    return _split(x, 1 if mode == PAD_REFLECT else 0)


def _split_g(kind, g, mode):
    """Hi/lo copies of the output gradient shared by the input- and weight-gradient kernels."""
    return _split(g, s2d=(kind == "s1" and mode == UPSAMPLE2X))


def _tc_dgrad(kind, gs, n, cache, weight, c, kp, split, splitp, mode, h, w, oh, ow):
    """Input gradient on the tensor cores.  gs: hi/lo of g (N, oh, ow, kp) from _split_g; returns (N, h, w, c)."""
    dev = gs[0].device
    if mode != UPSAMPLE2X:
        ws = cache.get_hilo(weight, c, kp, split, splitp, 0)      # [tap][ci][co]
    if kind == "s2":
        k = weight.shape[2]
        gx = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
        for a in range(2):                       # input pixel (2i + a, 2j + b) <- kernel rows with kh = a + 1 (mod 2)
            for b in range(2):
                taps = [((a + 1 - kh) // 2, (b + 1 - kw) // 2, 0, kh * k + kw)
                        for kh in range(k) if (a + 1 - kh) % 2 == 0 for kw in range(k) if (b + 1 - kw) % 2 == 0]
                _tc_launch(gs, n, ws, k * k, c, gx, oh, ow, c, 2, (a, b), taps)
        return gx
    if mode == PAD_REFLECT:
        gpad = torch.empty((n, h + 2, w + 2, c), device=dev, dtype=torch.float32)
        taps = [(-kh, -kw, 0, kh * 3 + kw) for kw in range(3) for kh in (2, 1, 0)]   # rows ascending: -2, -1, 0
        _tc_launch(gs, n, ws, 9, c, gpad, h + 2, w + 2, c, 1, (0, 0), taps, layout=1)
        gx = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
        _call("og_reflect_pad_bwd", _p(gpad), n, h, w, c, _p(gx))
        return gx
    if mode == PAD_ZERO:
        k = weight.shape[2]
        gx = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
        if k == 3:
            taps = [(1 - kh, 1 - kw, 0, kh * 3 + kw) for kw in range(3) for kh in (2, 1, 0)]
        else:       # gx[i, j] = sum g[i + 1 - kh, j + 1 - kw] W[kh, kw]^T over the (oh, ow) gradient grid
            taps = [(1 - kh, 1 - kw, 0, kh * k + kw) for kh in range(k) for kw in range(k)]
        _tc_launch(gs, n, ws, k * k, c, gx, h, w, c, 1, (0, 0), taps, layout=1 if k == 3 else 0)
        return gx
    # UPSAMPLE2X: adjoint of the four phase convolutions: gx[i,j] = sum_{p,q,a,b} G_pq[i - off(p,a), j - off(q,b)] Wp^T
    ws = cache.get_up_hilo(weight, c, kp, split, splitp, 0)
    gx = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
    taps = [(-_UP_OFF[p_][a], -_UP_OFF[q_][b], (p_ * 2 + q_) * n, ((p_ * 2 + q_) * 2 + a) * 2 + b)
            for p_ in range(2) for q_ in range(2) for a in range(2) for b in range(2)]
    _tc_launch(gs, n, ws, 16, c, gx, h, w, c, 1, (0, 0), taps)
    return gx


def _tc_wgrad_ok(kind, mode, n, oh, ow):
    """The wgrad kernel streams 64-pixel patches (cw x chh x cn, powers of two) of the gradient grid."""
    if kind == "s1" and mode == UPSAMPLE2X:
        oh, ow = oh // 2, ow // 2
    cw = 1
    while cw * 2 <= 64 and ow % (cw * 2) == 0:
        cw *= 2
    chh = 1
    while cw * chh * 2 <= 64 and oh % (chh * 2) == 0:
        chh *= 2
    # a last image group that the batch does not fill adds nothing (the unstacked operand is zero filled there)
    return True


def _grad_sink(param):
    """Persistent gradient buffer a trainer attached to a parameter (``param.og_grad_sink``): the backward kernels
    then accumulate into it directly and hand autograd no gradient (one kernel and one temporary less per use)."""
    sink = getattr(param, "og_grad_sink", None)
    if sink is not None and sink.shape == param.shape and sink.is_contiguous():
        return sink
    return None


def _tc_wgrad(kind, xs, gs, n, h, w, oh, ow, weight, c, kp, split, splitp, mode, sink=None):
    """Weight gradient on the tensor cores from the hi/lo copies of x (_split_x) and g (_split_g), both NHWC;
    (h, w) is the extent of x, (oh, ow) of g.  Returns the OIHW gradient."""
    import ctypes
    xh, xl, ax = xs
    gh, gl, ag = gs
    co, ci, kh_, kw_ = weight.shape
    if kind == "s2":
        ent = []
        for kh in range(kh_):
            dh, a = divmod(kh - 1, 2)
            for kw in range(kw_):
                dw, b = divmod(kw - 1, 2)
                ent.append((0, dh, dw, (a * 2 + b) * n, kh * kw_ + kw))
        nt = kh_ * kw_
    elif mode == UPSAMPLE2X:
        ent = [((p_ * 2 + q_) * n, _UP_OFF[p_][a], _UP_OFF[q_][b], 0, ((p_ * 2 + q_) * 2 + a) * 2 + b)
               for p_ in range(2) for q_ in range(2) for a in range(2) for b in range(2)]
        oh, ow, nt = h, w, 16
    else:
        off = 0 if mode == PAD_REFLECT else -1     # reflect: x copy carries the halo; zero pad: TMA fills
        ent = [(0, kh + off, kw + off, 0, kh * kw_ + kw) for kh in range(kh_) for kw in range(kw_)]
        nt = kh_ * kw_
    arr = _int_array(ent)
    dwp = torch.empty(nt * kp * c, device=xh.device, dtype=torch.float32)
    _call("og_conv2d_wgrad_tc", _p(gh), _p(gl), _p(ag), n, gh.shape[0], oh, ow, kp, _p(xh), _p(xl), _p(ax),
          xh.shape[0], xh.shape[1], xh.shape[2], c, _p(dwp), nt, ctypes.addressof(arr), len(ent), _nsplit())
    if mode == UPSAMPLE2X and kind == "s1":
        gw = torch.empty_like(weight)
        _call("og_unpack_upsample_wgrad", _p(dwp), co, ci, c, kp, split, splitp, _p(gw))
        return gw
    if sink is not None:
        _call("og_unpack_wgrad", _p(dwp), co, ci, kh_, kw_, c, kp, split, splitp, _p(sink), 1, 1)
        return None
    gw = torch.empty_like(weight)
    _call("og_unpack_wgrad", _p(dwp), co, ci, kh_, kw_, c, kp, split, splitp, _p(gw), 0, 1)
    return gw


def _out_hw(h, w, kh, kw, stride, pad, mode):
    if mode == UPSAMPLE2X:
        return 2 * h, 2 * w
    if mode == PAD_REFLECT:
        return (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    return (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1


def _conv_raw(x, wp, n, h, w, c, oh, ow, k, kh, kw, stride, pad, mode, bias, act, splitk=True):
    y = torch.empty((n, oh, ow, k), device=x.device, dtype=torch.float32)
    _call("og_conv2d_simt", _p(x), n, h, w, c, h * w * c, w * c, c, _p(wp), _p(y), oh, ow, k, oh * ow * k, ow * k, k,
          kh, kw, stride, pad, mode, _p(bias), act, LRELU_SLOPE, 1 if splitk else 0)
    return y


class _Conv2d(torch.autograd.Function):
    """y = act(conv(x, W) + b) on NHWC tensors; W, b are the reference's OIHW / (Co,) parameters."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache, stride, pad, mode, act, split):
        _chk(x, weight, bias)
        x = x.contiguous()
        n, h, w, c = x.shape
        co, ci, kh, kw = weight.shape
        assert c == cpad(ci), (c, ci)
        if split:
            assert co == 2 * split
            splitp = cpad(split)
            kp = 2 * splitp
        else:
            splitp, kp = 0, cpad(co)
        need_t = ctx.needs_input_grad[0]
        oh, ow = _out_hw(h, w, kh, kw, stride, pad, mode)
        kind = _tc_kind(n, h, w, c, kh, kw, stride, pad, mode)
        bias_p = None
        if bias is not None:
            if kp == co:
                bias_p = bias.detach()
            else:
                assert not split
                bias_p = torch.zeros(kp, device=x.device, dtype=torch.float32)
                bias_p[:co] = bias.detach()
        narrow = kind is None and kp == 8 and mode == PAD_ZERO and n * oh * ow <= 65536
        f32_missing = getattr(x, "og_f32_invalid", False)
        if f32_missing and not kind:
            raise RuntimeError("input exists only as tensor-core operand copies, but this convolution runs on the CUDA cores")
        xs = None
        if kind:
            y, xs = _tc_fprop(kind, x, cache, weight, kp, split, splitp, mode, bias_p, act)
            if not (KEEP_SPLIT and ctx.needs_input_grad[1] and TC_WGRAD and _tc_wgrad_ok(kind, mode, n, oh, ow)) \
                    and not f32_missing:
                xs = None
        elif narrow:
            wf, _ = cache.get(weight, c, kp, split, splitp, False)
            y = torch.empty((n, oh, ow, 8), device=x.device, dtype=torch.float32)
            _call("og_conv2d_narrow_fwd", _p(x), n, h, w, c, _p(wf), _p(y), oh, ow, kh, kw, stride, pad, _p(bias_p), act,
                  LRELU_SLOPE)
        else:
            wf, _ = cache.get(weight, c, kp, split, splitp, False)
            y = _conv_raw(x, wf, n, h, w, c, oh, ow, kp, kh, kw, stride, pad, mode, bias_p, act,
                          splitk=(bias is None and act == ACT_NONE))
        ctx.kind = kind
        ctx.wsink = _grad_sink(weight)
        ctx.bsink = _grad_sink(bias) if bias is not None else None
        ctx.xs = xs                      # hi/lo operand copies of x, reused by the weight gradient
        ctx.narrow = narrow
        ctx.cfg = (stride, pad, mode, act, split, splitp, kp, need_t)
        ctx.cache = cache
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, y = ctx.saved_tensors
        stride, pad, mode, act, split, splitp, kp, _ = ctx.cfg
        n, h, w, c = x.shape
        co, ci, kh, kw = weight.shape
        g = g.contiguous()
        oh, ow = g.shape[1], g.shape[2]
        if act != ACT_NONE:
            gp = torch.empty_like(g)
            _call("og_act_backward", _p(y), _p(g), g.numel(), act, LRELU_SLOPE, _p(gp))
            g = gp
        gx = gw = gb = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            scratch = torch.empty(kp, device=g.device, dtype=torch.float64)
            if ctx.bsink is not None:
                _call("og_channel_sum", _p(g), n * oh * ow, kp, _p(scratch), _p(ctx.bsink), co, 1)
            else:
                gb = torch.empty(co, device=g.device, dtype=torch.float32)
                _call("og_channel_sum", _p(g), n * oh * ow, kp, _p(scratch), _p(gb), co, 0)
        kind = ctx.kind
        if ctx.narrow:
            wf, _ = ctx.cache.get(weight, c, kp, split, splitp, False)
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                _call("og_conv2d_narrow_dgrad", _p(g), n, h, w, c, _p(wf), _p(gx), oh, ow, kh, kw, stride, pad)
            if ctx.needs_input_grad[1]:
                dwp = torch.empty(kh * kw * c * 8, device=g.device, dtype=torch.float32)
                _call("og_conv2d_narrow_wgrad", _p(x), n, h, w, c, _p(g), _p(dwp), oh, ow, kh, kw, stride, pad)
                if ctx.wsink is not None:
                    _call("og_unpack_wgrad", _p(dwp), co, ci, kh, kw, c, kp, split, splitp, _p(ctx.wsink), 1, 0)
                else:
                    gw = torch.empty_like(weight)
                    _call("og_unpack_wgrad", _p(dwp), co, ci, kh, kw, c, kp, split, splitp, _p(gw), 0, 0)
            return gx, gw, gb, None, None, None, None, None, None
        tc_w = bool(ctx.needs_input_grad[1] and kind and TC_WGRAD and _tc_wgrad_ok(kind, mode, n, oh, ow))
        gs = _split_g(kind, g, mode) if kind and (ctx.needs_input_grad[0] or tc_w) else None
        if ctx.needs_input_grad[0] and kind:
            gx = _tc_dgrad(kind, gs, n, ctx.cache, weight, c, kp, split, splitp, mode, h, w, oh, ow)
        elif ctx.needs_input_grad[0]:
            _, wt = ctx.cache.get(weight, c, kp, split, splitp, True)
            if mode == PAD_ZERO:
                gx = _conv_raw(g, wt, n, oh, ow, kp, h, w, c, kh, kw, stride, pad, TRANSPOSED, None, ACT_NONE)
            elif mode == PAD_REFLECT:
                assert stride == 1 and pad == 1
                gpad = _conv_raw(g, wt, n, oh, ow, kp, h + 2, w + 2, c, kh, kw, 1, 0, TRANSPOSED, None, ACT_NONE)
                gx = torch.empty_like(x)
                _call("og_reflect_pad_bwd", _p(gpad), n, h, w, c, _p(gx))
            else:  # UPSAMPLE2X
                gu = _conv_raw(g, wt, n, oh, ow, kp, 2 * h, 2 * w, c, kh, kw, 1, pad, TRANSPOSED, None, ACT_NONE)
                gx = torch.empty_like(x)
                _call("og_upsample2x_bwd", _p(gu), n, h, w, c, _p(gx))
        if tc_w:
            xs = ctx.xs if ctx.xs is not None else _split_x(kind, x, mode)
            ctx.xs = None
            gw = _tc_wgrad(kind, xs, gs, n, h, w, oh, ow, weight, c, kp, split, splitp, mode, ctx.wsink)
        elif ctx.needs_input_grad[1]:
            dwp = torch.empty(kh * kw * c * kp, device=g.device, dtype=torch.float32)
            _call("og_conv2d_wgrad_simt", _p(x), n, h, w, c, h * w * c, w * c, c, _p(g), oh, ow, kp, oh * ow * kp,
                  ow * kp, kp, _p(dwp), kh, kw, stride, pad, mode)
            if ctx.wsink is not None:
                _call("og_unpack_wgrad", _p(dwp), co, ci, kh, kw, c, kp, split, splitp, _p(ctx.wsink), 1, 0)
            else:
                gw = torch.empty_like(weight)
                _call("og_unpack_wgrad", _p(dwp), co, ci, kh, kw, c, kp, split, splitp, _p(gw), 0, 0)
        return gx, gw, gb, None, None, None, None, None, None


def conv2d(x, weight, bias, cache, *, stride=1, pad=1, mode=PAD_ZERO, act=ACT_NONE, split=0):
    return _Conv2d.apply(x, weight, bias, cache, stride, pad, mode, act, split)


# --------------------------------------------------------------------------------------------------
class _NormAct(torch.autograd.Function):
    """InstanceNorm2d (groups = N) or train-mode BatchNorm (groups = 1) + fused activation (+ residual)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, res, bn_buffers, instance, act, split_pad=None, keep_f32=True):
        _chk(y, gamma, beta, res)
        y = y.contiguous()
        n, h, w, cy = y.shape
        groups = n if instance else 1
        P = h * w if instance else n * h * w
        dev = y.device
        stats = torch.empty(groups * cy * 2, device=dev, dtype=torch.float64)
        mean = torch.empty(groups * cy, device=dev, dtype=torch.float32)
        rstd = torch.empty(groups * cy, device=dev, dtype=torch.float32)
        rm = rv = nbt = None
        if bn_buffers is not None:
            rm, rv, nbt = bn_buffers
            assert gamma.numel() == cy, "BatchNorm layers of this model never need channel padding"
        _call("og_norm_stats", _p(y), groups, P, cy, NORM_EPS, _p(stats), _p(mean), _p(rstd), _p(rm), _p(rv),
              BN_MOMENTUM, cy, _p(nbt))
        co = cy // 2 if act == NA_GLU else cy
        out = torch.empty((n, h, w, co), device=dev, dtype=torch.float32)
        if res is not None:
            res = res.contiguous()
            assert res.shape == out.shape
        fused = (split_pad is not None and FUSED_SPLIT and CONV_ENGINE in ("f16x3", "f16") and not _lib.DRY_RUN
                 and (res is None or _known_amax(res) is not None) and (split_pad == 0 or (h > 1 and w > 1)))
        if fused:
            # the consuming convolution's operand copies straight from this pass, scaled by an a-priori bound of max|out|
            word = torch.empty(1, device=dev, dtype=torch.int32)
            _call("og_norm_bound", _p(gamma), _p(beta), cy if gamma is not None else 0, P,
                  _p(_known_amax(res)) if res is not None else 0, _p(word))
            xh = _empty_slack((n, h + 2 * split_pad, w + 2 * split_pad, co), dev)
            xl = _empty_slack((n, h + 2 * split_pad, w + 2 * split_pad, co), dev) if CONV_ENGINE == "f16x3" else None
            _call("og_norm_apply_split", _p(y), n, h, w, cy, 1 if instance else 0, _p(mean), _p(rstd), _p(gamma),
                  _p(beta), _p(res), act, LRELU_SLOPE, _p(out) if keep_f32 else 0, _p(word), split_pad, _p(xh), _p(xl))
            out.og_amax = (word, out._version)          # an upper bound is a valid operand scale for other consumers
            out.og_split = (xh, xl, word, split_pad, False, out._version, _nsplit())
            if not keep_f32:
                out.og_f32_invalid = True               # nothing may read the fp32 tensor: it was never written
        else:
            _call("og_norm_apply", _p(y), groups, P, cy, _p(mean), _p(rstd), _p(gamma), _p(beta), _p(res), act,
                  LRELU_SLOPE, _p(out), _p(_tag_amax(out)))
        ctx.cfg = (groups, P, cy, act, res is not None)
        ctx.sinks = (_grad_sink(gamma), _grad_sink(beta)) if gamma is not None else (None, None)
        ctx.save_for_backward(y, mean, rstd, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, g):
        y, mean, rstd, gamma, beta = ctx.saved_tensors
        groups, P, cy, act, has_res = ctx.cfg
        g = g.contiguous()
        bstats = torch.empty(groups * cy * 2, device=g.device, dtype=torch.float64)
        dy = torch.empty_like(y)
        dgamma = dbeta = None
        want_pg = gamma is not None and ctx.needs_input_grad[1]     # frozen discriminators (G update): no param grads
        direct = want_pg and ctx.sinks[0] is not None and ctx.sinks[1] is not None
        if direct:
            dgamma, dbeta = ctx.sinks
        elif want_pg:
            dgamma = torch.empty_like(gamma)
            dbeta = torch.empty_like(beta)
        _call("og_norm_backward", _p(y), _p(g), groups, P, cy, _p(mean), _p(rstd), _p(gamma), _p(beta), act,
              LRELU_SLOPE, _p(bstats), _p(dy), _p(dgamma), _p(dbeta), 1 if direct else 0, _p(_tag_amax(dy)))
        if direct:
            dgamma = dbeta = None
        return dy, dgamma, dbeta, (g if has_res else None), None, None, None, None, None


def instance_norm_act(y, act, res=None, split_pad=None, keep_f32=True):
    """``split_pad`` (0 / 1): also emit the fp16 hi / lo operand copies the NEXT convolution needs (plain / with the
    reflection halo), so that convolution skips its split pass; ``keep_f32=False``: the fp32 tensor is not written at
    all (only valid when that convolution is the ONLY reader and runs on the tensor cores)."""
    return _NormAct.apply(y, None, None, res, None, True, act, split_pad, keep_f32)


def batch_norm_act(y, gamma, beta, buffers, act):
    return _NormAct.apply(y, gamma, beta, None, buffers, False, act)


# --------------------------------------------------------------------------------------------------
class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cp):
        _chk(x)
        x = x.contiguous()
        n, c, h, w = x.shape
        y = torch.empty((n, h, w, cp), device=x.device, dtype=torch.float32)
        _call("og_nchw_to_nhwc", _p(x), n, c, h, w, cp, _p(y))
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n, h, w, cp = g.shape
        gx = torch.empty((n, ctx.c, h, w), device=g.device, dtype=torch.float32)
        _call("og_nhwc_to_nchw", _p(g), n, ctx.c, h, w, cp, _p(gx))
        return gx, None


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, c):
        _chk(y)
        y = y.contiguous()
        n, h, w, cp = y.shape
        x = torch.empty((n, c, h, w), device=y.device, dtype=torch.float32)
        _call("og_nhwc_to_nchw", _p(y), n, c, h, w, cp, _p(x))
        ctx.cp = cp
        return x

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n, c, h, w = g.shape
        gy = torch.empty((n, h, w, ctx.cp), device=g.device, dtype=torch.float32)
        _call("og_nchw_to_nhwc", _p(g), n, c, h, w, ctx.cp, _p(gy))
        return gy, None


# Derived copies of CONSTANT inputs (no gradient: segmentation maps, real images): a step converts / resizes the same
# input tensor for several networks (the 80-channel 256^2 class maps go through NCHW -> NHWC eight times and through the
# 512^2 bilinear resize four times in one Step-B); the copy is kept while the source tensor is the same object at the
# same version.  Never used while a CUDA graph is being captured (a replay must recompute from the static buffers).
_CONST_CACHE = []          # [(weakref to the source, version, key, value)], newest last
_CONST_CACHE_SLOTS = 12


def cached_const(src, key, fn):
    import weakref
    if src.requires_grad or _lib.DRY_RUN or not src.is_cuda or torch.cuda.is_current_stream_capturing():
        return fn()
    for ref, ver, k, val in _CONST_CACHE:
        if k == key and ver == src._version and ref() is src:
            return val
    val = fn()
    _CONST_CACHE[:] = [e for e in _CONST_CACHE if e[0]() is not None][-(_CONST_CACHE_SLOTS - 1):]
    _CONST_CACHE.append((weakref.ref(src), src._version, key, val))
    return val


def to_nhwc(x, cp=None):
    cp = cpad(x.shape[1]) if cp is None else cp
    if not x.requires_grad:
        return cached_const(x, ("nhwc", cp), lambda: _ToNHWC.apply(x, cp))
    return _ToNHWC.apply(x, cp)


def to_nchw(y, c):
    return _ToNCHW.apply(y, c)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _chk(a, b)
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        _call("og_add", _p(a), _p(b), _p(out), a.numel())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a, b)


class _CatChannels(torch.autograd.Function):
    """torch.cat along channels for NHWC tensors with padded lanes: copies the first reals[i] channels of
    each input back to back and zero-fills the tail up to cpad(sum(reals))."""

    @staticmethod
    def forward(ctx, reals, *xs):
        _chk(*xs)
        xs = [x.contiguous() for x in xs]
        n, h, w, _ = xs[0].shape
        tot = sum(reals)
        cp = cpad(tot)
        out = torch.empty((n, h, w, cp), device=xs[0].device, dtype=torch.float32)
        if cp != tot:
            out.zero_()
        off = 0
        P = n * h * w
        for x, r in zip(xs, reals):
            _call("og_copy_channels", _p(x), x.shape[3], 0, _p(out), cp, off, r, P, 0)
            off += r
        ctx.reals = reals
        ctx.cps = [x.shape[3] for x in xs]
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n, h, w, cp = g.shape
        P = n * h * w
        outs, off = [], 0
        for i, (r, c) in enumerate(zip(ctx.reals, ctx.cps)):
            if ctx.needs_input_grad[1 + i]:
                gi = torch.empty((n, h, w, c), device=g.device, dtype=torch.float32)
                if c != r:
                    gi.zero_()
                _call("og_copy_channels", _p(g), cp, off, _p(gi), c, 0, r, P, 0)
                outs.append(gi)
            else:
                outs.append(None)
            off += r
        return (None, *outs)


def cat_channels(xs, reals):
    return _CatChannels.apply(tuple(reals), *xs)


class _GatherRows(torch.autograd.Function):
    """out[i] = x[idx[i]] over the leading dimension (idx: int64 device tensor)."""

    @staticmethod
    def forward(ctx, x, idx):
        _chk(x, idx)
        x = x.contiguous()
        n_in = x.shape[0]
        rowlen = x.numel() // max(n_in, 1)
        out = torch.empty((idx.numel(),) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
        _call("og_gather_rows", _p(x), _p(idx), idx.numel(), rowlen, _p(out))
        ctx.save_for_backward(idx)
        ctx.shape = tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        gx = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        n_in = ctx.shape[0]
        rowlen = gx.numel() // max(n_in, 1)
        _call("og_scatter_rows_add", _p(g), _p(idx), idx.numel(), n_in, rowlen, _p(gx))
        return gx, None


def gather_rows(x, idx):
    return _GatherRows.apply(x, idx)


def cat_rows_const(a, b):
    """torch.cat((a, b), dim=1) for two constant (no-grad) 2-D tensors: two strided row copies."""
    _chk(a, b)
    a, b = a.detach().contiguous(), b.detach().contiguous()
    r, ca, cb = a.shape[0], a.shape[1], b.shape[1]
    assert b.shape[0] == r
    out = torch.empty((r, ca + cb), device=a.device, dtype=torch.float32)
    _call("og_copy_channels", _p(a), ca, 0, _p(out), ca + cb, 0, ca, r, 0)
    _call("og_copy_channels", _p(b), cb, 0, _p(out), ca + cb, ca, cb, r, 0)
    return out


class _BroadcastCat(torch.autograd.Function):
    """D_GET_LOGITS conditioning: cat(h, c_code broadcast over the grid) along channels.  c_code is the detached
    sentence embedding in the patch-D callers (miscc/losses.py:169-190, 375-377); in the object-discriminator terms of
    G_loss it carries the generator's bt_c_code (losses.py:436-452), so its gradient is produced when asked for."""

    @staticmethod
    def forward(ctx, h, c_code):
        _chk(h, c_code)
        h, c_code = h.contiguous(), c_code.contiguous()
        n, hh, ww, ch = h.shape
        cc = c_code.shape[1]
        cp = cpad(ch + cc)
        out = torch.empty((n, hh, ww, cp), device=h.device, dtype=torch.float32)
        if cp != ch + cc:
            out.zero_()
        P = n * hh * ww
        _call("og_copy_channels", _p(h), ch, 0, _p(out), cp, 0, ch, P, 0)
        _call("og_broadcast_channels", _p(c_code), n, cc, _p(out), cp, ch, hh * ww)
        ctx.ch, ctx.cc = ch, cc
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n, hh, ww, cp = g.shape
        gh = torch.empty((n, hh, ww, ctx.ch), device=g.device, dtype=torch.float32)
        _call("og_copy_channels", _p(g), cp, 0, _p(gh), ctx.ch, 0, ctx.ch, n * hh * ww, 0)
        gc = None
        if ctx.needs_input_grad[1]:
            gc = torch.empty((n, ctx.cc), device=g.device, dtype=torch.float32)
            _call("og_broadcast_channels_bwd", _p(g), n, ctx.cc, cp, ctx.ch, hh * ww, _p(gc))
        return gh, gc


class _CatRows(torch.autograd.Function):
    """torch.cat((a, b), dim=1) of two 2-D tensors where only ``b`` carries a gradient."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[1]
        return cat_rows_const(a, b)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        r, c = g.shape
        cb = c - ctx.ca
        gb = torch.empty((r, cb), device=g.device, dtype=torch.float32)
        _call("og_copy_channels", _p(g), c, ctx.ca, _p(gb), cb, 0, cb, r, 0)
        return None, gb


def cat_rows(a_const, b):
    return _CatRows.apply(a_const, b)


def broadcast_cat(h, c_code):
    return _BroadcastCat.apply(h, c_code)


class _GLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _chk(x)
        x = x.contiguous()
        ch = x.shape[-1] // 2
        P = x.numel() // (2 * ch)
        out = torch.empty((*x.shape[:-1], ch), device=x.device, dtype=torch.float32)
        _call("og_glu_fwd", _p(x), P, ch, _p(out))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        ch = x.shape[-1] // 2
        gx = torch.empty_like(x)
        _call("og_glu_bwd", _p(x), _p(g.contiguous()), x.numel() // (2 * ch), ch, _p(gx))
        return gx


def glu(x):
    return _GLU.apply(x)


class _Reparam(torch.autograd.Function):
    """c = eps * exp(0.5 * logvar) + mu with x rows = [mu (D) | logvar (D) | pad]; c rows padded to cpad(D)."""

    @staticmethod
    def forward(ctx, x, eps, d):
        _chk(x, eps)
        x, eps = x.contiguous(), eps.contiguous()
        b = x.shape[0]
        cs = cpad(d)
        c = torch.zeros((b, cs), device=x.device, dtype=torch.float32)
        _call("og_reparam_fwd", _p(x), x.shape[1], _p(eps), b, d, _p(c), cs)
        ctx.d = d
        ctx.save_for_backward(x, eps)
        return c

    @staticmethod
    def backward(ctx, gc):
        x, eps = ctx.saved_tensors
        gc = gc.contiguous()
        gx = torch.zeros_like(x)
        _call("og_reparam_bwd", _p(x), x.shape[1], _p(eps), _p(gc), gc.shape[1], x.shape[0], ctx.d, _p(gx))
        return gx, None, None


def reparam(x, eps, d):
    return _Reparam.apply(x, eps, d)


# --------------------------------------------------------------------------------------------------
# losses: scalar = weight * mean(...)
# --------------------------------------------------------------------------------------------------
class _BCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, target, weight):
        _chk(p)
        p = p.contiguous()
        loss = torch.zeros((), device=p.device, dtype=torch.float32)
        gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        _call("og_bce", _p(p), p.numel(), float(target), float(weight), _p(loss), _p(gp))
        ctx.save_for_backward(gp)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return gp * g, None, None


def bce(p, target, weight=1.0):
    """weight * nn.BCELoss()(p, full_like(p, target))."""
    return _BCE.apply(p, target, weight)


class _KL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, d, weight):
        _chk(x)
        x = x.contiguous()
        loss = torch.zeros((), device=x.device, dtype=torch.float32)
        gx = torch.zeros_like(x)
        _call("og_kl", _p(x), x.shape[1], x.shape[0], d, float(weight), _p(loss), _p(gx))
        ctx.save_for_backward(gx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gx,) = ctx.saved_tensors
        return gx * g, None, None


def kl_rows(x, d, weight=1.0):
    return _KL.apply(x, d, weight)


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
def mask_bytes(mask):
    if mask is None:
        return None
    return mask.to(torch.uint8).contiguous()


class _WordsProj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, words, w):
        _chk(words, w)
        words = words.contiguous()
        b, cdf, l = words.shape
        idf = w.shape[0]
        src = torch.empty((b, idf, l), device=words.device, dtype=torch.float32)
        _call("og_words_proj", _p(words), _p(w), b, idf, cdf, l, _p(src))
        ctx.save_for_backward(words, w)
        return src

    @staticmethod
    def backward(ctx, g):
        words, w = ctx.saved_tensors
        g = g.contiguous()
        b, cdf, l = words.shape
        idf = w.shape[0]
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        gwords = torch.empty_like(words) if ctx.needs_input_grad[0] else None
        _call("og_words_proj_bwd", _p(words), _p(w), _p(g), b, idf, cdf, l, _p(gw), 0, _p(gwords))
        return gwords, gw


def words_proj(words, w):
    return _WordsProj.apply(words, w)


class _AttGeneral(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, src, mask_u8, idf):
        _chk(h, src, mask_u8)
        h, src = h.contiguous(), src.contiguous()
        b, ih, iw, cs = h.shape
        l = src.shape[2]
        q = ih * iw
        wc = torch.empty_like(h)
        attn = torch.empty((b, l, ih, iw), device=h.device, dtype=torch.float32)
        _call("og_att_general_fwd", _p(h), _p(src), _p(mask_u8), b, q, idf, cs, l, _p(wc), _p(attn))
        ctx.idf = idf
        ctx.save_for_backward(h, src, attn)
        ctx.mark_non_differentiable(attn)  # visualisation only (trainer.py:390 discards it)
        return wc, attn

    @staticmethod
    def backward(ctx, g_wc, _g_attn):
        h, src, attn = ctx.saved_tensors
        b, ih, iw, cs = h.shape
        l = src.shape[2]
        g_h = torch.empty_like(h)
        g_src = torch.empty_like(src)
        _call("og_att_general_bwd", _p(h), _p(src), _p(attn), _p(g_wc.contiguous()), 0, b, ih * iw, ctx.idf, cs, l,
              _p(g_h), _p(g_src))
        return g_h, g_src, None, None


def att_general(h, src, mask_u8, idf):
    return _AttGeneral.apply(h, src, mask_u8, idf)


class _BuAtt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, glove, src, mask_u8, norm):
        _chk(labels, glove, src, mask_u8)
        labels, glove, src = labels.contiguous(), glove.contiguous(), src.contiguous()
        b, e, r = labels.shape
        l = glove.shape[2]
        idf = src.shape[1]
        wc = torch.empty((b, idf, r), device=src.device, dtype=torch.float32)
        attn = torch.empty((b, l, r), device=src.device, dtype=torch.float32)
        _call("og_bu_att_fwd", _p(labels), _p(glove), _p(src), _p(mask_u8), b, e, r, l, idf, 1 if norm else 0, 1e-8,
              _p(wc), _p(attn))
        ctx.save_for_backward(attn)
        ctx.dims = (b, r, l, idf)
        ctx.mark_non_differentiable(attn)
        return wc, attn

    @staticmethod
    def backward(ctx, g_wc, _g):
        (attn,) = ctx.saved_tensors
        b, r, l, idf = ctx.dims
        g_src = torch.empty((b, idf, l), device=attn.device, dtype=torch.float32)
        _call("og_bu_att_bwd", _p(attn), _p(g_wc.contiguous()), b, r, l, idf, _p(g_src))
        return None, None, g_src, None, None


def bu_att(labels, glove, src, mask_u8, norm=True):
    return _BuAtt.apply(labels, glove, src, mask_u8, norm)


class _PaintMax(torch.autograd.Function):
    """pprocess_bt_attns: out[b, y, x, k] = max_r f[b, k, r] * m[b, r, y, x]; f (B, num, R), m (B, Rtot, ih, iw)."""

    @staticmethod
    def forward(ctx, f, m):
        _chk(f, m)
        f, m = f.contiguous(), m.contiguous()
        b, num, r = f.shape
        _, rtot, ih, iw = m.shape
        cp = cpad(num)
        out = torch.empty((b, ih, iw, cp), device=f.device, dtype=torch.float32)
        if cp != num:
            out.zero_()
        _call("og_paint_max_fwd", _p(f), _p(m), b, num, r, rtot, ih * iw, _p(out), cp, 0)
        ctx.save_for_backward(f, m)
        return out

    @staticmethod
    def backward(ctx, g):
        f, m = ctx.saved_tensors
        g = g.contiguous()
        b, num, r = f.shape
        _, rtot, ih, iw = m.shape
        g_f = torch.empty_like(f)
        _call("og_paint_max_bwd", _p(f), _p(m), _p(g), g.shape[3], 0, b, num, r, rtot, ih * iw, _p(g_f))
        return g_f, None


def paint_max(f, m):
    return _PaintMax.apply(f, m)


class _FuncAttention(torch.autograd.Function):
    """DAMSM attention (reference: GlobalAttention.py:32-70) on the reference's NCHW tensors."""

    @staticmethod
    def forward(ctx, query, context, gamma1):
        _chk(query, context)
        query, context = query.contiguous(), context.contiguous()
        b, ndf, lq = query.shape
        ih, iw = context.shape[2:]
        wc = torch.empty((b, ndf, lq), device=query.device, dtype=torch.float32)
        attn = torch.empty((b, lq, ih, iw), device=query.device, dtype=torch.float32)
        _call("og_func_attention_fwd", _p(query), _p(context), b, ndf, lq, ih * iw, float(gamma1), _p(wc), _p(attn))
        ctx.gamma1 = float(gamma1)
        ctx.save_for_backward(query, context, attn)
        return wc, attn

    @staticmethod
    def backward(ctx, g_wc, g_attn):
        query, context, attn = ctx.saved_tensors
        b, ndf, lq = query.shape
        s = context.shape[2] * context.shape[3]
        g_q = torch.empty_like(query)
        g_c = torch.empty_like(context)
        ga = g_attn.contiguous() if g_attn is not None else None
        _call("og_func_attention_bwd", _p(query), _p(context), _p(attn), _p(g_wc.contiguous()), _p(ga), b, ndf, lq, s,
              ctx.gamma1, _p(g_q), _p(g_c))
        return g_q, g_c, None


def func_attention(query, context, gamma1):
    return _FuncAttention.apply(query, context, gamma1)


class _CosineCL(torch.autograd.Function):
    """out[b, l] = cosine_similarity(word[:, l], wei[b, :, l]) (ref: miscc/losses.py:13-19 as used at 101-108).
    ``word`` (D, L) is one caption's embedding shared by all images and is treated as a constant."""

    @staticmethod
    def forward(ctx, word, wei, eps):
        _chk(word, wei)
        word, wei = word.detach().contiguous(), wei.contiguous()
        b, d, l = wei.shape
        assert word.shape == (d, l)
        out = torch.empty((b, l), device=wei.device, dtype=torch.float32)
        _call("og_cosine_cl_fwd", _p(word), _p(wei), b, d, l, float(eps), _p(out))
        ctx.eps = float(eps)
        ctx.save_for_backward(word, wei)
        return out

    @staticmethod
    def backward(ctx, g):
        word, wei = ctx.saved_tensors
        b, d, l = wei.shape
        gwei = torch.empty_like(wei)
        _call("og_cosine_cl_bwd", _p(word), _p(wei), _p(g.contiguous()), b, d, l, ctx.eps, _p(gwei))
        return None, gwei, None


def cosine_cl(word, wei, eps=1e-8):
    return _CosineCL.apply(word, wei, eps)


class _ExpSumLog(torch.autograd.Function):
    """out[b] = log(sum_l exp(gamma * s[b, l]))  (ref: miscc/losses.py:112-115)."""

    @staticmethod
    def forward(ctx, s, gamma):
        _chk(s)
        s = s.contiguous()
        b, l = s.shape
        out = torch.empty((b,), device=s.device, dtype=torch.float32)
        _call("og_expsumlog_fwd", _p(s), b, l, float(gamma), _p(out))
        ctx.gamma = float(gamma)
        ctx.save_for_backward(s)
        return out

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        b, l = s.shape
        gs = torch.empty_like(s)
        _call("og_expsumlog_bwd", _p(s), _p(g.contiguous()), b, l, ctx.gamma, _p(gs))
        return gs, None


def expsumlog(s, gamma):
    return _ExpSumLog.apply(s, gamma)


def h2d(data, device, dtype=None):
    """Small host table (list / numpy array / CPU tensor) -> device WITHOUT stalling the host: staged in pinned memory
    and copied asynchronously on the current stream (a pageable ``.to(device)`` blocks the calling thread until the GPU
    has drained the stream, which serialises an eager step with its own kernels).  The caching pinned allocator only
    reuses the staging block after the copy has executed."""
    import numpy as np
    if torch.is_tensor(data):
        t = data if dtype is None else data.to(dtype)
    elif isinstance(data, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(data))
        t = t if dtype is None else t.to(dtype)
    else:
        t = torch.tensor(data, dtype=dtype)
    dev = torch.device(device)
    if dev.type != "cuda" or t.device.type == "cuda":
        return t.to(dev)
    return t.contiguous().pin_memory().to(dev, non_blocking=True)


def permute_channels(x, perm):
    """out[b, c] = x[b, perm[b, c]] for an NCHW tensor and a (B, C) int64 table on the device."""
    _chk(x, perm)
    x = x.detach().contiguous()
    b, c, h, w = x.shape
    assert perm.shape == (b, c) and perm.dtype == torch.int64
    out = torch.empty_like(x)
    _call("og_permute_channels", _p(x), _p(perm.contiguous()), b, c, h * w, _p(out))
    return out


def zeros(shape, device):
    """Zero-filled fp32 tensor through cudaMemsetAsync (a memset node inside a captured graph, no library kernel)."""
    t = torch.empty(shape, device=device, dtype=torch.float32)
    _chk(t)
    _call("og_zero_bytes", _p(t), t.numel() * 4)
    return t


def form_hmaps(bt_masks, roi_cls, num_rois, num_classes, clamp_max=0.0, out=None):
    """Class heat maps (B, num_classes, S, S) from the per-roi masks (B, R, S, S) and the roi class ids (B, R) int64:
    what the reference's loader accumulates on the host (ref: miscc/load.py:160-176), rebuilt on the device."""
    _chk(bt_masks, roi_cls, num_rois)
    bt_masks = bt_masks.contiguous()
    b, r, h, w = bt_masks.shape
    assert roi_cls.dtype == torch.int64 and num_rois.dtype == torch.int64 and roi_cls.shape == (b, r)
    if out is None:
        out = torch.empty((b, num_classes, h, w), device=bt_masks.device, dtype=torch.float32)
    _call("og_form_hmaps", _p(bt_masks), _p(roi_cls.contiguous()), _p(num_rois.contiguous()), b, r, h * w, num_classes,
          float(clamp_max), _p(out))
    return out


def form_clabels_feat(clabels_emb, roi_cls, num_rois, rmax, out=None):
    """ref: miscc/utils.py:502-522 on the device: (B, E, rmax, 1) label embeddings of the first num_rois[b] boxes."""
    _chk(clabels_emb, roi_cls, num_rois)
    b, r = roi_cls.shape
    ncls, e = clabels_emb.shape
    if out is None:
        out = torch.empty((b, e, rmax, 1), device=clabels_emb.device, dtype=torch.float32)
    _call("og_form_clabels_feat", _p(clabels_emb.contiguous()), _p(roi_cls.contiguous()), _p(num_rois.contiguous()), b, r,
          rmax, e, ncls, _p(out))
    return out


class _WordsPairs(torch.autograd.Function):
    """All B x NC (image, caption) pairs of words_loss in one launch (ref: miscc/losses.py:87-127): func_attention,
    word / attended-context cosine and the Eq. (10) pooling.  ctx_feat (B, ndf, ih, iw) carries the gradient;
    words (NC, ndf, T) and lens (NC,) int64 on the device are constants.  Returns sim (B, NC), attn (B, NC, T, ih, iw)
    (rows beyond a caption's length are zero)."""

    @staticmethod
    def forward(ctx, feat, words, lens, gamma1, gamma2, eps):
        _chk(feat, words, lens)
        feat, words = feat.contiguous(), words.detach().contiguous()
        assert lens.dtype == torch.int64 and lens.is_contiguous()
        b, ndf, ih, iw = feat.shape
        nc, _, t = words.shape
        s = ih * iw
        wc = zeros((b * nc, ndf, t), feat.device)
        attn = zeros((b, nc, t, ih, iw), feat.device)
        sim = torch.empty((b, nc), device=feat.device, dtype=torch.float32)
        _call("og_words_pairs_fwd", _p(words), _p(feat), _p(lens), b, nc, ndf, t, s, float(gamma1), float(gamma2),
              float(eps), _p(wc), _p(attn), _p(sim))
        ctx.cfg = (b, nc, ndf, t, s, float(gamma1), float(gamma2), float(eps))
        ctx.save_for_backward(feat, words, lens, wc, attn)
        ctx.mark_non_differentiable(attn)
        return sim, attn

    @staticmethod
    def backward(ctx, g_sim, _g_attn):
        feat, words, lens, wc, attn = ctx.saved_tensors
        b, nc, ndf, t, s, g1, g2, eps = ctx.cfg
        g_feat = torch.empty_like(feat)
        _call("og_words_pairs_bwd", _p(words), _p(feat), _p(lens), _p(wc), _p(attn), _p(g_sim.contiguous()), b, nc, ndf,
              t, s, g1, g2, eps, _p(g_feat))
        return g_feat, None, None, None, None, None


def words_pairs(feat, words, lens, gamma1, gamma2, eps=1e-8):
    return _WordsPairs.apply(feat, words, lens, gamma1, gamma2, eps)


class _CosineMatrix(torch.autograd.Function):
    """out[i, j] = <a_i, b_j> / max(|a_i| |b_j|, eps) (ref: miscc/losses.py:43-50); ``b`` is a constant."""

    @staticmethod
    def forward(ctx, a, b, eps):
        _chk(a, b)
        a, b = a.contiguous(), b.detach().contiguous()
        out = torch.empty((a.shape[0], b.shape[0]), device=a.device, dtype=torch.float32)
        _call("og_cosine_matrix_fwd", _p(a), _p(b), a.shape[0], b.shape[0], a.shape[1], float(eps), _p(out))
        ctx.eps = float(eps)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = torch.empty_like(a)
        _call("og_cosine_matrix_bwd", _p(a), _p(b), _p(g.contiguous()), a.shape[0], b.shape[0], a.shape[1], ctx.eps,
              _p(ga))
        return ga, None, None


def cosine_matrix(a, b, eps=1e-8):
    return _CosineMatrix.apply(a, b, eps)


class _CEPair(torch.autograd.Function):
    """(CrossEntropyLoss(scores, labels), CrossEntropyLoss(scores^T, labels), top-1 hits) with scores = gamma3 * sim and
    the masked entries at -inf (ref: miscc/losses.py:50-68, 131-147)."""

    @staticmethod
    def forward(ctx, sim, mask, labels, gamma3):
        _chk(sim, mask, labels)
        sim = sim.contiguous()
        b = sim.shape[0]
        assert sim.shape == (b, b)
        dev = sim.device
        l0 = torch.empty((), device=dev, dtype=torch.float32)
        l1 = torch.empty((), device=dev, dtype=torch.float32)
        hits = torch.empty((), device=dev, dtype=torch.float32)
        g0 = torch.empty((b, b), device=dev, dtype=torch.float32)
        g1 = torch.empty((b, b), device=dev, dtype=torch.float32)
        m = mask.contiguous() if mask is not None else None
        _call("og_ce_pair", _p(sim), _p(m), _p(labels.contiguous()), b, float(gamma3), _p(l0), _p(l1), _p(g0), _p(g1),
              _p(hits))
        ctx.save_for_backward(g0, g1)
        ctx.mark_non_differentiable(hits)
        return l0, l1, hits

    @staticmethod
    def backward(ctx, gl0, gl1, _gc):
        g0, g1 = ctx.saved_tensors
        gsim = torch.empty_like(g0)
        a = gl0.contiguous() if gl0 is not None else None
        b = gl1.contiguous() if gl1 is not None else None
        _call("og_ce_pair_bwd", _p(g0), _p(g1), _p(a), _p(b), g0.numel(), _p(gsim))
        return gsim, None, None, None


def ce_pair(sim, mask, labels, gamma3):
    return _CEPair.apply(sim, mask, labels, gamma3)


class _Bilinear(torch.autograd.Function):
    """F.interpolate(size, mode='bilinear', align_corners=True) on NHWC tensors."""

    @staticmethod
    def forward(ctx, x, oh, ow):
        _chk(x)
        x = x.contiguous()
        n, h, w, c = x.shape
        y = torch.empty((n, oh, ow, c), device=x.device, dtype=torch.float32)
        _call("og_bilinear_fwd", _p(x), n, h, w, c, oh, ow, _p(y))
        ctx.dims = (n, h, w, c, oh, ow)
        return y

    @staticmethod
    def backward(ctx, g):
        n, h, w, c, oh, ow = ctx.dims
        gx = torch.empty((n, h, w, c), device=g.device, dtype=torch.float32)
        _call("og_bilinear_bwd", _p(g.contiguous()), n, h, w, c, oh, ow, _p(gx))
        return gx, None, None


def bilinear(x, oh, ow):
    return _Bilinear.apply(x, oh, ow)


# --------------------------------------------------------------------------------------------------
# ROIAlign (NCHW like the reference op)
# --------------------------------------------------------------------------------------------------
class _RoIAlign(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ah, aw, scale):
        _chk(features, rois)
        features, rois = features.contiguous(), rois.contiguous()
        assert rois.shape[1] == 5
        b, c, h, w = features.shape
        r = rois.shape[0]
        out = torch.zeros((r, c, ah, aw), device=features.device, dtype=torch.float32)
        _call("ROIAlignForwardLaucher", _p(features), float(scale), r, h, w, c, ah, aw, _p(rois), _p(out))
        ctx.cfg = (b, c, h, w, ah, aw, scale)
        ctx.save_for_backward(rois)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        b, c, h, w, ah, aw, scale = ctx.cfg
        gin = torch.zeros((b, c, h, w), device=g.device, dtype=torch.float32)
        _call("ROIAlignBackwardLaucher", _p(g.contiguous()), float(scale), b, rois.shape[0], h, w, c, ah, aw, _p(rois),
              _p(gin))
        return gin, None, None, None, None


def roi_align(features, rois, ah, aw, scale):
    return _RoIAlign.apply(features, rois, ah, aw, scale)


class _RoIAlignAvg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ah, aw, scale):
        _chk(features, rois)
        features, rois = features.contiguous(), rois.contiguous()
        assert rois.shape[1] == 5
        b, c, h, w = features.shape
        r = rois.shape[0]
        out = torch.empty((r, c, ah, aw), device=features.device, dtype=torch.float32)
        _call("og_roi_align_avg_fwd", _p(features), h, w, c, _p(rois), r, ah, aw, float(scale), _p(out))
        ctx.cfg = (b, c, h, w, ah, aw, scale)
        ctx.save_for_backward(rois)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        b, c, h, w, ah, aw, scale = ctx.cfg
        gin = torch.zeros((b, c, h, w), device=g.device, dtype=torch.float32)
        _call("og_roi_align_avg_bwd", _p(g.contiguous()), h, w, c, _p(rois), rois.shape[0], ah, aw, float(scale),
              _p(gin))
        return gin, None, None, None, None


def roi_align_avg(features, rois, ah, aw, scale):
    return _RoIAlignAvg.apply(features, rois, ah, aw, scale)


class _RoIAlignAvgNHWC(torch.autograd.Function):
    """Channels-last fused RoIAlignAvg: features (B, H, W, C) -> (R, ah, aw, C); same values as ``roi_align_avg``."""

    @staticmethod
    def forward(ctx, features, rois, ah, aw, scale):
        _chk(features, rois)
        features, rois = features.contiguous(), rois.contiguous()
        assert rois.shape[1] == 5 and features.shape[3] % 4 == 0
        b, h, w, c = features.shape
        r = rois.shape[0]
        out = torch.empty((r, ah, aw, c), device=features.device, dtype=torch.float32)
        _call("og_roi_align_avg_nhwc_fwd", _p(features), h, w, c, _p(rois), r, ah, aw, float(scale), _p(out))
        ctx.cfg = (b, h, w, c, ah, aw, scale)
        ctx.save_for_backward(rois)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        b, h, w, c, ah, aw, scale = ctx.cfg
        gin = zeros((b, h, w, c), g.device)
        _call("og_roi_align_avg_nhwc_bwd", _p(g.contiguous()), h, w, c, _p(rois), rois.shape[0], ah, aw, float(scale),
              _p(gin))
        return gin, None, None, None, None


def roi_align_avg_nhwc(features, rois, ah, aw, scale):
    return _RoIAlignAvgNHWC.apply(features, rois, ah, aw, scale)


# --------------------------------------------------------------------------------------------------
def adam_ema_(p, g, m, v, avg, step, *, lr=2e-4, b1=0.5, b2=0.999, eps=1e-8, gscale=1.0, decay=0.999, step_dev=None,
              bump=True):
    """Fused Adam (+EMA) over flat fp32 buffers, in place.  ``step_dev`` (int64 device scalar) replaces the host
    step count and is incremented first, which keeps the call replayable from a CUDA graph."""
    _chk(p, g, m, v, avg, step_dev)
    if step_dev is not None:
        _call("og_inc_i64", _p(step_dev))
    _call("og_adam_ema", _p(p), _p(g), _p(m), _p(v), _p(avg), p.numel(), float(lr), float(b1), float(b2), float(eps),
          int(step), _p(step_dev), float(gscale), float(decay))
    if bump:
        bump_param_epoch()
