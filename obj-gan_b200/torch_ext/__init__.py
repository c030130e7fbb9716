"""The PyTorch C++ extension boundary of the reference's one native operator (north_star: "a thin PyTorch C++/CUDA
extension"): ``torch.ops.objgan_b200.roi_align_forward_cuda / roi_align_backward_cuda`` registered by
``roi_align_op.cpp`` (TORCH_LIBRARY) on top of the C-ABI launchers of ``libobjgan_b200.so``, plus ``RoIAlignFunction``
-- what ``models/roi_align/functions/roi_align.py`` becomes when its dead cffi import is replaced by this module
(same constructor arguments, same forward / backward contract; written as a modern static autograd.Function because
the legacy instance-style Function no longer runs on torch >= 1.3).

    from objgan_b200.torch_ext import roi_align            # plays the role of `from .._ext import roi_align`
    roi_align.roi_align_forward_cuda(ah, aw, scale, features, rois, output)
"""
from __future__ import annotations

import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SO = os.path.join(HERE, "libobjgan_b200_torch.so")
SRC = os.path.join(HERE, "roi_align_op.cpp")


def build(force: bool = False) -> str:
    """g++ -shared against the torch headers / libraries of the running interpreter and the in-tree libobjgan_b200.so
    (no nvcc needed: the file contains no kernel).  In-tree output, like the CUDA library."""
    core = os.path.join(PKG, "libobjgan_b200.so")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(SRC), os.path.getmtime(core)):
        return SO
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{os.path.join(os.path.dirname(PKG), 'include')}"]
    lib = ce.library_paths(device_type="cuda")
    cmd = (["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
            SRC, "-o", SO] + inc + [f"-L{p}" for p in lib] + [f"-L{PKG}", "-lobjgan_b200", "-ltorch", "-ltorch_cpu",
                                                              "-lc10", "-ltorch_cuda", "-lc10_cuda",
                                                              "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + lib[0]])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the torch extension failed:\n" + " ".join(cmd) + "\n" + r.stderr[-4000:])
    return SO


_loaded = False


def load():
    """Register the operators (idempotent).  Fails loudly if the extension is not built: there is no fallback."""
    global _loaded
    if not _loaded:
        if not os.path.exists(SO):
            raise ImportError(f"{SO} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        torch.ops.load_library(SO)
        _loaded = True
    return torch.ops.objgan_b200


class _Ext:
    """Namespace with the reference's `_ext.roi_align` entry points (roi_align_cuda.h:1-5)."""

    @staticmethod
    def roi_align_forward_cuda(aligned_height, aligned_width, spatial_scale, features, rois, output):
        return load().roi_align_forward_cuda(int(aligned_height), int(aligned_width), float(spatial_scale), features,
                                             rois, output)

    @staticmethod
    def roi_align_backward_cuda(aligned_height, aligned_width, spatial_scale, top_grad, rois, bottom_grad):
        return load().roi_align_backward_cuda(int(aligned_height), int(aligned_width), float(spatial_scale), top_grad,
                                              rois, bottom_grad)


roi_align = _Ext


class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale):
        ctx.cfg = (aligned_height, aligned_width, spatial_scale, tuple(features.size()))
        ctx.save_for_backward(rois)
        output = features.new_zeros(rois.size(0), features.size(1), aligned_height, aligned_width)
        roi_align.roi_align_forward_cuda(aligned_height, aligned_width, spatial_scale, features.contiguous(),
                                         rois.contiguous(), output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        ah, aw, scale, size = ctx.cfg
        assert grad_output.is_cuda                                    # functions/roi_align.py:38
        grad_input = rois.new_zeros(size)
        roi_align.roi_align_backward_cuda(ah, aw, scale, grad_output.contiguous(), rois.contiguous(), grad_input)
        return grad_input, None, None, None, None


class RoIAlignFunction:
    """Call-compatible with the reference's ``RoIAlignFunction(aligned_height, aligned_width, spatial_scale)
    (features, rois)`` (functions/roi_align.py:7-51)."""

    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width, self.aligned_height = int(aligned_width), int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _RoIAlignFn.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)
