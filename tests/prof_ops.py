"""Launches every bandwidth-bound kernel of the hot path twice at its BASELINE-config size, for one
`ncu --set full` capture (see profiles/README.md for the command).  Not a bench: numbers under ncu are not reported."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from objgan_b200 import model, ops, synth, trainer
from objgan_b200.lib import NA_GLU, NA_NONE

DEV = "cuda"
B, C, L = 16, 48, 18
torch.manual_seed(0)
# config 3: grid attention, Q = 16384
h = torch.randn(B, 128, 128, C, device=DEV, requires_grad=True)
srcw = torch.randn(B, C, L, device=DEV, requires_grad=True)
for _ in range(2):
    wc, att = ops.att_general(h, srcw, None, C)
    torch.autograd.grad(wc, (h, srcw), torch.ones_like(wc))
# config 3: all B x B pairs of words_loss
feat = torch.randn(B, 256, 17, 17, device=DEV, requires_grad=True)
words = torch.randn(B, 256, 18, device=DEV)
lens = torch.full((B,), 18, dtype=torch.int64, device=DEV)
for _ in range(2):
    sim, _ = ops.words_pairs(feat, words, lens, 4.0, 5.0)
    torch.autograd.grad(sim, feat, torch.ones_like(sim))
# config 4: bottom-up attention + 3 paints at 128^2, B = 32, R = 10; RoIAlignAvg 320 rois
B4, R = 32, 10
bu = model.BT_ATT_NET(48, 256).to(DEV)
lab, glove, wrd = torch.randn(B4, 50, R, 1, device=DEV), torch.randn(B4, 50, L, device=DEV), torch.randn(B4, 256, L, device=DEV)
m = torch.rand(B4, R, 128, 128, device=DEV)
for _ in range(2):
    wc2, att2 = bu(lab, glove, wrd)
    p = ops.paint_max(wc2.reshape(B4, 48, R), m)
    ops.paint_max(att2.detach().reshape(B4, L, R), m)
    ops.paint_max(lab.reshape(B4, 50, R), m)
    p.sum().backward()
for (Cf, H) in ((384, 64), (768, 32)):
    f = torch.randn(B4, H, H, Cf, device=DEV, requires_grad=True)
    xy = torch.rand(B4 * R, 2) * 40 * 16 * H / 64
    wh = (6 + torch.rand(B4 * R, 2) * 18) * 16 * H / 64
    rois = torch.cat([torch.arange(B4).repeat_interleave(R).float().unsqueeze(1), xy, xy + wh], 1).to(DEV)
    for _ in range(2):
        o = ops.roi_align_avg_nhwc(f, rois, 5, 5, 1 / 16)
        o.backward(torch.ones_like(o))
# normalisation + GLU of the stage-3 residual block (400 -> 200 channels at 128^2, B = 16), Adam + EMA, operand split
y = torch.randn(B, 128, 128, 400, device=DEV, requires_grad=True)
for _ in range(2):
    a = ops.instance_norm_act(y, NA_GLU)
    a.backward(torch.ones_like(a))
y2 = torch.randn(B, 128, 128, 200, device=DEV, requires_grad=True)
res = torch.randn(B, 128, 128, 200, device=DEV)
for _ in range(2):
    a = ops.instance_norm_act(y2, NA_NONE, res)
    a.backward(torch.ones_like(a))
# residual block of stage 3 (194 channels at 128^2, B = 16): the producer-side operand split (norm_apply_split_kernel)
blk = torch.nn.Sequential(model.HmapResBlock(194), model.HmapResBlock(194)).to(DEV)
model._chain_res_blocks(blk, 1)
xb = torch.randn(B, 128, 128, 200, device=DEV)
xb[..., 194:] = 0
for _ in range(2):
    with torch.no_grad():
        blk(xb)
n = 19_340_000
pp, g, mm, v, avg = (torch.randn(n, device=DEV) for _ in range(5))
for t in (1, 2):
    ops.adam_ema_(pp, g, mm, v.abs_(), avg, t)
x = torch.randn(B, 128, 128, 200, device=DEV)
for _ in range(2):
    ops._split(x, 1)
inp = synth.compact(synth.make_inputs(B, seed=1, parity=False))
tr = trainer.StepATrainer(device=DEV, seed=1)
for _ in range(2):
    tr.to_device(inp)
torch.cuda.synchronize()
