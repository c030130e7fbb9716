// Thin PyTorch C++ operator layer over the C ABI of libobjgan_b200.so: the boundary BASELINE.json's north_star
// names ("exposed through a thin PyTorch C++/CUDA extension").  It registers, under torch.ops.objgan_b200, the two
// entry points the reference's Python side calls through its (long dead) cffi module:
//     roi_align_forward_cuda / roi_align_backward_cuda
//     (ref: image_generation/models/roi_align/src/roi_align_cuda.h:1-5; callers functions/roi_align.py:24-27, 44-47)
// with the same argument order and meaning.  Differences that the new op interface forces or that SURVEY.md 8b asks
// for: tensors are at::Tensor (no THCState global: the launch goes to at::cuda::getCurrentCUDAStream(), so the op is
// re-entrant across the per-device threads of nn.DataParallel), argument errors raise (TORCH_CHECK) instead of
// returning 0 / calling exit(), and the caller-allocated, zero-filled output convention is kept.
// No kernel lives here: the work is done by ROIAlignForwardLaucher / ROIAlignBackwardLaucher of include/objgan_b200.h.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "objgan_b200.h"

namespace {

void check(const at::Tensor& t, const char* name, int64_t dim) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(t.dim() == dim, name, " must have ", dim, " dimensions");
}

int64_t roi_align_forward_cuda(int64_t aligned_height, int64_t aligned_width, double spatial_scale,
                               const at::Tensor& features, const at::Tensor& rois, at::Tensor output) {
  check(features, "features", 4);
  check(rois, "rois", 2);
  check(output, "output", 4);
  TORCH_CHECK(rois.size(1) == 5, "rois must be (R, 5) [batch_index, x1, y1, x2, y2]");
  TORCH_CHECK(output.size(0) == rois.size(0) && output.size(1) == features.size(1) && output.size(2) == aligned_height &&
                  output.size(3) == aligned_width, "output must be (R, C, aligned_height, aligned_width)");
  c10::cuda::CUDAGuard guard(features.device());
  const int ok = ROIAlignForwardLaucher(features.data_ptr<float>(), (float)spatial_scale, (int)rois.size(0),
                                        (int)features.size(2), (int)features.size(3), (int)features.size(1),
                                        (int)aligned_height, (int)aligned_width, rois.data_ptr<float>(),
                                        output.data_ptr<float>(), at::cuda::getCurrentCUDAStream());
  TORCH_CHECK(ok == 1, "ROIAlignForwardLaucher failed");
  return 1;
}

int64_t roi_align_backward_cuda(int64_t aligned_height, int64_t aligned_width, double spatial_scale,
                                const at::Tensor& top_grad, const at::Tensor& rois, at::Tensor bottom_grad) {
  check(top_grad, "top_grad", 4);
  check(rois, "rois", 2);
  check(bottom_grad, "bottom_grad", 4);
  TORCH_CHECK(rois.size(1) == 5, "rois must be (R, 5) [batch_index, x1, y1, x2, y2]");
  c10::cuda::CUDAGuard guard(top_grad.device());
  const int ok = ROIAlignBackwardLaucher(top_grad.data_ptr<float>(), (float)spatial_scale, (int)bottom_grad.size(0),
                                         (int)rois.size(0), (int)bottom_grad.size(2), (int)bottom_grad.size(3),
                                         (int)bottom_grad.size(1), (int)aligned_height, (int)aligned_width,
                                         rois.data_ptr<float>(), bottom_grad.data_ptr<float>(),
                                         at::cuda::getCurrentCUDAStream());
  TORCH_CHECK(ok == 1, "ROIAlignBackwardLaucher failed");
  return 1;
}

}  // namespace

TORCH_LIBRARY(objgan_b200, m) {
  m.def("roi_align_forward_cuda(int aligned_height, int aligned_width, float spatial_scale, Tensor features, "
        "Tensor rois, Tensor(a!) output) -> int");
  m.def("roi_align_backward_cuda(int aligned_height, int aligned_width, float spatial_scale, Tensor top_grad, "
        "Tensor rois, Tensor(a!) bottom_grad) -> int");
}
TORCH_LIBRARY_IMPL(objgan_b200, CUDA, m) {
  m.impl("roi_align_forward_cuda", &roi_align_forward_cuda);
  m.impl("roi_align_backward_cuda", &roi_align_backward_cuda);
}
