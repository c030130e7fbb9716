"""Per-layer timing of every distinct convolution of one Step-A step (CUDA events; guidance only, not a bench value).
Records the (shape, mode) of each ops.conv2d call of an eager step together with how often it runs forward /
backward, then times forward, input gradient and weight gradient of each distinct layer alone."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objgan_b200 import synth, trainer, ops, model

calls = collections.OrderedDict()
_orig_fwd = ops._Conv2d.forward
_orig_bwd = ops._Conv2d.backward


def fwd(ctx, x, weight, bias, cache, stride, pad, mode, act, split):
    key = (tuple(x.shape), tuple(weight.shape), stride, pad, mode, split, bias is not None)
    ctx._key = key
    calls.setdefault(key, [0, 0, 0])[0] += 1
    return _orig_fwd(ctx, x, weight, bias, cache, stride, pad, mode, act, split)


def bwd(ctx, g):
    c = calls[ctx._key]
    c[1] += 1 if ctx.needs_input_grad[0] else 0
    c[2] += 1 if ctx.needs_input_grad[1] else 0
    return _orig_bwd(ctx, g)


ops._Conv2d.forward = staticmethod(fwd)
ops._Conv2d.backward = staticmethod(bwd)
tr = trainer.StepATrainer(device="cuda", seed=1234)
inp = synth.make_inputs(16, seed=1234, parity=False)
inp.pop("eps")
dev = tr.to_device(inp)
tr._eager_step(dev) if hasattr(tr, "_eager_step") else tr.step(dev)
torch.cuda.synchronize()
ops._Conv2d.forward = staticmethod(_orig_fwd)
ops._Conv2d.backward = staticmethod(_orig_bwd)


def timeit(fn, n=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for key, (nf, nd, nw) in calls.items():
    xs, ws, stride, pad, mode, split, has_bias = key
    co, ci, kh, kw = ws
    m = model.Conv2dP(ci, co, kh, stride, pad, bias=has_bias, mode=mode, split=split).cuda()
    x = torch.randn(xs, device="cuda")
    with torch.no_grad():
        tf = timeit(lambda: m(x))
    xg = x.clone().requires_grad_(True)
    y = m(xg)
    g = torch.randn_like(y)
    td = tw = 0.0
    if nd:
        m.weight.requires_grad_(False)
        xg2 = x.clone().requires_grad_(True)
        y2 = m(xg2)
        td = timeit(lambda: torch.autograd.grad(y2, xg2, g, retain_graph=True))
        m.weight.requires_grad_(True)
    if nw:
        x3 = x.clone()
        y3 = m(x3)
        tw = timeit(lambda: torch.autograd.grad(y3, m.weight, g, retain_graph=True))
    n, h, w, _ = xs
    oh, ow = y.shape[1], y.shape[2]
    macs = n * oh * ow * co * ci * kh * kw
    kind = ops._tc_kind(n, h, w, xs[3], kh, kw, stride, pad, mode)
    rows.append((nf * tf + nd * td + nw * tw, key, nf, nd, nw, tf, td, tw, macs, kind))
tot = sum(r[0] for r in rows)
print(f"total conv time per step (sum of isolated timings): {tot:.2f} ms")
print("share  calls(f/d/w)  fwd ms (TF/s)  dgrad ms (TF/s)  wgrad ms (TF/s)  kind  x-shape  w-shape  s p mode split")
for t, key, nf, nd, nw, tf, td, tw, macs, kind in sorted(rows, key=lambda r: -r[0]):
    fl = 2.0 * macs / 1e9
    s = lambda ms: f"{ms:7.3f} ({fl / ms:6.1f})" if ms else "      -        "
    print(f"{100 * t / tot:5.1f}%  {nf:2d}/{nd:2d}/{nw:2d}  {s(tf)}  {s(td)}  {s(tw)}  {kind}  {key[0]} {key[1]} {key[2]} {key[3]} {key[4]} {key[5]}")
