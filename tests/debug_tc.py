"""Manual diagnostics for conv_tc.cu (run on the GPU box: python tests/debug_tc.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objgan_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def run(x, w, taps, nsplit_engine="tf32", osy=1, op=(0, 0), oh=None, ow=None, yshape=None):
    """x (N,H,W,C) ; w (ntaps, K, C) ; returns y (N,OH,OW,K)"""
    ops.CONV_ENGINE = nsplit_engine
    n, h, wd, c = x.shape
    nt, k, _ = w.shape
    xh, xl = ops._split(x, 0)
    wh = (w.view(torch.int32) & ~0x1FFF).view(torch.float32).contiguous()
    wl = (w - wh).contiguous()
    oh = h if oh is None else oh
    ow = wd if ow is None else ow
    y = torch.full(yshape or (n, oh, ow, k), float("nan"), device=dev)
    ops._tc_launch(xh, xl, wh, wl, nt, k, y, oh, ow, k, osy, op, taps)
    torch.cuda.synchronize()
    return y


def ref(x, w, taps):
    n, h, wd, c = x.shape
    y = torch.zeros(n, h, wd, w.shape[1], device=dev, dtype=torch.float64)
    xp = torch.zeros(n, h + 8, wd + 8, c, device=dev, dtype=torch.float64)
    xp[:, 4:4 + h, 4:4 + wd] = x.double()
    for (dh, dw, wi) in taps:
        y += torch.einsum("nhwc,kc->nhwk", xp[:, 4 + dh:4 + dh + h, 4 + dw:4 + dw + wd], w[wi].double())
    return y.float()


def report(name, y, yr):
    bad = ~torch.isfinite(y)
    err = (y - yr).abs()
    err[bad] = float("inf")
    print(f"{name}: max|ref|={yr.abs().max().item():.4g} maxerr={err.max().item():.4g} nan={int(bad.sum())} "
          f"frac_bad={(err > 1e-3 * yr.abs().max()).float().mean().item():.4f}")
    return err


# A: identity weights, 1 tap, C=K=32, x[p,c]=row index within tile -> checks row mapping
N, H, W, C, K = 1, 8, 16, 32, 32
rowid = torch.arange(H * W, device=dev, dtype=torch.float32).view(1, H, W, 1).expand(1, H, W, C).contiguous()
eye = torch.eye(K, C, device=dev).view(1, K, C).contiguous()
y = run(rowid, eye, [(0, 0, 0)])
e = report("A rows (x=row id, W=I)", y, rowid)
print("   y[0,0,:4,0]=", y[0, 0, :4, 0].tolist(), " y[0,1,:4,0]=", y[0, 1, :4, 0].tolist(), " y[0,:4,0,5]=", y[0, :4, 0, 5].tolist())
colid = torch.arange(C, device=dev, dtype=torch.float32).view(1, 1, 1, C).expand(1, H, W, C).contiguous()
y = run(colid, eye, [(0, 0, 0)])
report("A cols (x=col id, W=I)", y, colid)
print("   y[0,0,0,:]=", y[0, 0, 0, :].tolist())
print("   y[0,3,5,:]=", y[0, 3, 5, :].tolist())
# B: random small-int data, 1 tap
xi = torch.randint(-4, 5, (1, 8, 16, 32), device=dev).float()
wi = torch.randint(-4, 5, (1, 32, 32), device=dev).float()
report("B int C=32 K=32", run(xi, wi, [(0, 0, 0)]), ref(xi, wi, [(0, 0, 0)]))
# C: two k-chunks
xi = torch.randint(-4, 5, (1, 8, 16, 64), device=dev).float()
wi = torch.randint(-4, 5, (1, 32, 64), device=dev).float()
report("C int C=64 K=32", run(xi, wi, [(0, 0, 0)]), ref(xi, wi, [(0, 0, 0)]))
# D: K=64, K=112, K=208, K=256
for K in (64, 112, 208, 256, 400):
    xi = torch.randint(-4, 5, (1, 8, 16, 32), device=dev).float()
    wi = torch.randint(-4, 5, (1, K, 32), device=dev).float()
    report(f"D int C=32 K={K}", run(xi, wi, [(0, 0, 0)]), ref(xi, wi, [(0, 0, 0)]))
# E: taps with shifts
xi = torch.randint(-4, 5, (2, 16, 16, 32), device=dev).float()
wi = torch.randint(-4, 5, (9, 32, 32), device=dev).float()
taps = [(kh - 1, kw - 1, kh * 3 + kw) for kh in range(3) for kw in range(3)]
report("E 3x3 zero-pad C=32 K=32 N=2 16x16", run(xi, wi, taps), ref(xi, wi, taps))
report("E single shifted tap (1,0)", run(xi, wi, [(1, 0, 3)]), ref(xi, wi, [(1, 0, 3)]))
report("E single shifted tap (0,-1)", run(xi, wi, [(0, -1, 5)]), ref(xi, wi, [(0, -1, 5)]))
# F: 3xTF32 on random floats
xf = torch.randn(2, 16, 16, 64, device=dev)
wf = torch.randn(9, 48, 64, device=dev)
report("F tf32x3 random", run(xf, wf, taps, "tf32x3"), ref(xf, wf, taps))
report("F tf32 random", run(xf, wf, taps, "tf32"), ref(xf, wf, taps))
# G: C=200 (ragged last chunk), K=200
xf = torch.randn(2, 16, 16, 200, device=dev)
wf = torch.randn(9, 200, 200, device=dev)
report("G tf32x3 C=200 K=200", run(xf, wf, taps, "tf32x3"), ref(xf, wf, taps))
