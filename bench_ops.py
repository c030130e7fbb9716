#!/usr/bin/env python
"""Operator-level measurements for BASELINE.json configs[2] (attention isolation) and configs[3] (object path):
achieved HBM GB/s of the bandwidth kernels against the measured copy peak, and ROIAlign against the reference's own
roi_align_kernel.cu compiled for sm_100a (oracle/_ref) on the same GPU.  Prints one JSON object per line.

    python bench_ops.py > profiles/r01_ops_bench.jsonl
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from objgan_b200 import model, ops  # noqa: E402

DEV = "cuda"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0)
FLUSH = torch.empty(256 * 1024 * 1024 // 4, device=DEV)


def timeit(fn, iters=10, warm=3):
    ts = []
    for i in range(warm + iters):
        FLUSH.zero_()                                  # evict the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def attention_sweep():
    C, L = 48, 18
    for B in (16, 32, 64):
        for Q in (1024, 4096, 16384):
            ih = int(Q ** 0.5)
            h = torch.randn(B, ih, ih, C, device=DEV, requires_grad=True)
            src = torch.randn(B, C, L, device=DEV, requires_grad=True)
            ms = timeit(lambda: ops.att_general(h.detach(), src.detach(), None, C))
            byts = 4.0 * Q * (2 * C + L) * B
            wc, _ = ops.att_general(h, src, None, C)
            g = torch.randn_like(wc)

            def bwd():
                torch.autograd.grad(wc, (h, src), g, retain_graph=True)
            msb = timeit(bwd)
            bytb = 4.0 * Q * (3 * C + L) * B
            emit(config=3, op="GlobalAttentionGeneral", B=B, Q=Q, L=L, C=C, fwd_ms=round(ms, 4),
                 fwd_gbs=round(byts / ms / 1e6, 1), fwd_frac_of_hbm_peak=round(byts / ms / 1e6 / HBM, 3),
                 bwd_ms=round(msb, 4), bwd_gbs=round(bytb / msb / 1e6, 1), hbm_peak_gbs=HBM)
    for P in (256, 1024):                              # B^2 (image, caption) pairs of words_loss
        q = torch.randn(P, 256, 18, device=DEV)
        ctx = torch.randn(P, 256, 17, 17, device=DEV)
        ms = timeit(lambda: ops.func_attention(q, ctx, 4.0))
        byts = 4.0 * P * (256 * 289 + 256 * 18 * 2 + 18 * 289)
        emit(config=3, op="func_attention", pairs=P, regions=289, Lq=18, ndf=256, fwd_ms=round(ms, 4),
             fwd_gbs=round(byts / ms / 1e6, 1), fwd_frac_of_hbm_peak=round(byts / ms / 1e6 / HBM, 3))


def object_path():
    B, R, L = 32, 10, 18
    for ih in (64, 128):
        bu = model.BT_ATT_NET(48, 256).to(DEV)
        lab, glove, words = torch.randn(B, 50, R, 1, device=DEV), torch.randn(B, 50, L, device=DEV), torch.randn(B, 256, L, device=DEV)
        m = torch.rand(B, R, ih, ih, device=DEV)

        def run():
            wc, att = bu(lab, glove, words)
            ops.paint_max(wc.reshape(B, 48, R), m)
            ops.paint_max(att.reshape(B, L, R), m)
            ops.paint_max(lab.reshape(B, 50, R), m)
        with torch.no_grad():
            ms = timeit(run)
        byts = 3 * 4.0 * B * R * ih * ih + 4.0 * B * ih * ih * (48 + 24 + 56)
        emit(config=4, op="BT_ATT_NET + 3x pprocess_bt_attns", B=B, R=R, size=ih, ms=round(ms, 4), gbs=round(byts / ms / 1e6, 1),
             frac_of_hbm_peak=round(byts / ms / 1e6 / HBM, 3))
    so = os.path.join(ROOT, "oracle", "_ref", "libroi_align_ref_cuda.so")
    ref = ctypes.CDLL(so) if os.path.exists(so) else None
    if ref:
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        ref.ROIAlignForwardLaucher.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, vp, vp, vp]
        ref.ROIAlignBackwardLaucher.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    for (C, H) in ((384, 64), (768, 32)):
        feat = torch.randn(B, C, H, H, device=DEV)
        xy = torch.rand(B * R, 2) * 40
        wh = 6 + torch.rand(B * R, 2) * 18
        rois = torch.cat([torch.arange(B).repeat_interleave(R).float().unsqueeze(1), xy, xy + wh], 1).to(DEV)
        nr = B * R
        st = torch.cuda.current_stream().cuda_stream
        ms_f = timeit(lambda: ops.roi_align_avg(feat, rois, 5, 5, 1 / 16))
        g = torch.randn(nr, C, 5, 5, device=DEV)
        gin = torch.zeros_like(feat)

        def bwd():
            gin.zero_()
            ops._call("og_roi_align_avg_bwd", g.data_ptr(), H, H, C, rois.data_ptr(), nr, 5, 5, 1 / 16, gin.data_ptr())
        ms_b = timeit(bwd)
        rec = dict(config=4, op="RoIAlignAvg(5,5,1/16)", B=B, rois=nr, C=C, H=H, fused_fwd_ms=round(ms_f, 4),
                   fused_bwd_ms=round(ms_b, 4), fwd_out_gbs=round(4.0 * nr * C * 25 / ms_f / 1e6, 1))
        if ref:
            o6 = torch.zeros(nr, C, 6, 6, device=DEV)

            def ref_f():
                ref.ROIAlignForwardLaucher(feat.data_ptr(), 1 / 16, nr, H, H, C, 6, 6, rois.data_ptr(), o6.data_ptr(), st)
                torch.nn.functional.avg_pool2d(o6, 2, 1)
            g6 = torch.randn_like(o6)

            def ref_b():
                gin.zero_()
                ref.ROIAlignBackwardLaucher(g6.data_ptr(), 1 / 16, B, nr, H, H, C, 6, 6, rois.data_ptr(), gin.data_ptr(), st)
            rec["reference_cu_fwd_ms"] = round(timeit(ref_f), 4)
            rec["reference_cu_bwd_ms"] = round(timeit(ref_b), 4)
            rec["note"] = "reference = roi_align_kernel.cu compiled verbatim for sm_100a + torch avg_pool2d (fwd); its bwd excludes the pool adjoint"
        emit(**rec)


if __name__ == "__main__":
    attention_sweep()
    object_path()
