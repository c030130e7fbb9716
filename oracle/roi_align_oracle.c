/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of the reference's ROIAlign arithmetic
 *   forward : /root/reference/image_generation/models/roi_align/src/roi_align.c:80-137
 *             (same expression types as roi_align_kernel.cu:15-70)
 *   backward: the adjoint that roi_align_kernel.cu:94-143 computes with atomicAdd.  The reference's own CPU
 *             backward (roi_align.c:139-190) tests the bounds the wrong way round (line 175) and cannot be
 *             used as a checker, so the adjoint is restated here from the forward.
 * Pinned by tests/test_oracle_golden.py against oracle/_ref/libroi_align_ref_cpu.so (the reference roi_align.c
 * compiled verbatim, see oracle/Makefile) and against tests/golden/roi_align.npz.
 * Build: gcc -O2 -std=c99 -ffp-contract=off -shared -fPIC roi_align_oracle.c -o _ref/libroi_align_oracle.so -lm
 */
#include <math.h>

typedef struct { int hstart, wstart, ok; float h_ratio, w_ratio; } og_sample;

static og_sample locate(const float* roi, float spatial_scale, int height, int width, int AH, int AW, int ph, int pw) {
  og_sample s;
  float x1 = roi[1] * spatial_scale, y1 = roi[2] * spatial_scale;
  float x2 = roi[3] * spatial_scale, y2 = roi[4] * spatial_scale;
  float roi_w = fmaxf(x2 - x1 + 1., 0.);          /* float difference, "+ 1." in double, narrowed by fmaxf */
  float roi_h = fmaxf(y2 - y1 + 1., 0.);
  float bin_h = roi_h / (AH - 1.);                /* double division, narrowed */
  float bin_w = roi_w / (AW - 1.);
  float h = (float)(ph)*bin_h + y1;               /* float; inclusive end points */
  float w = (float)(pw)*bin_w + x1;
  s.hstart = fminf(floor(h), height - 2);
  s.wstart = fminf(floor(w), width - 2);
  s.ok = !(h < 0 || h >= height || w < 0 || w >= width);
  s.h_ratio = h - (float)(s.hstart);
  s.w_ratio = w - (float)(s.wstart);
  return s;
}

void og_oracle_roi_align_forward(const float* bottom, float spatial_scale, int num_rois, int height, int width,
                                 int channels, int AH, int AW, const float* rois, float* top) {
  for (int n = 0; n < num_rois; ++n) {
    const float* roi = rois + n * 5;
    int img_start = roi[0] * channels * height * width;   /* float arithmetic, truncated */
    for (int c = 0; c < channels; ++c)
      for (int ph = 0; ph < AH; ++ph)
        for (int pw = 0; pw < AW; ++pw) {
          og_sample s = locate(roi, spatial_scale, height, width, AH, AW, ph, pw);
          float* out = top + ((n * channels + c) * AH + ph) * AW + pw;
          if (!s.ok) { *out = 0.; continue; }
          const float* p = bottom + img_start + (c * height + s.hstart) * width + s.wstart;
          *out = p[0] * (1. - s.h_ratio) * (1. - s.w_ratio) + p[1] * (1. - s.h_ratio) * s.w_ratio +
                 p[width] * s.h_ratio * (1. - s.w_ratio) + p[width + 1] * s.h_ratio * s.w_ratio;
        }
  }
}

/* bottom_diff must be zero-filled by the caller; accumulation in double, order-independent up to rounding */
void og_oracle_roi_align_backward(const float* top_diff, float spatial_scale, int num_rois, int height, int width,
                                  int channels, int AH, int AW, const float* rois, double* bottom_diff) {
  for (int n = 0; n < num_rois; ++n) {
    const float* roi = rois + n * 5;
    int img_start = roi[0] * channels * height * width;
    for (int c = 0; c < channels; ++c)
      for (int ph = 0; ph < AH; ++ph)
        for (int pw = 0; pw < AW; ++pw) {
          og_sample s = locate(roi, spatial_scale, height, width, AH, AW, ph, pw);
          if (!s.ok) continue;
          double g = top_diff[((n * channels + c) * AH + ph) * AW + pw];
          double* p = bottom_diff + img_start + (c * height + s.hstart) * width + s.wstart;
          p[0] += g * (1. - s.h_ratio) * (1. - s.w_ratio);
          p[1] += g * (1. - s.h_ratio) * s.w_ratio;
          p[width] += g * s.h_ratio * (1. - s.w_ratio);
          p[width + 1] += g * s.h_ratio * s.w_ratio;
        }
  }
}
