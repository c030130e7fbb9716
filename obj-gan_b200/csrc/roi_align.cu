// ROIAlign forward / backward for sm_100a, plus the fused RoIAlignAvg (align to (AH+1)x(AW+1), then
// 2x2 stride-1 average pool) used by OBJ_SS_D_NET / OBJ_LS_D_NET.
//
// Drop-in for the reference's only native operator (paths under /root/reference/image_generation/):
//   ROIAlignForwardLaucher / ROIAlignBackwardLaucher   models/roi_align/src/roi_align_kernel.h:13-27
//   (kernels roi_align_kernel.cu:15-70, 94-143; CPU twin roi_align.c:80-137)
//   RoIAlignAvg                                        models/roi_align/modules/roi_align.py:18-29
//
// Arithmetic contract (bit-exact sample coordinates): the reference mixes float and double --
// float products for the scaled roi corners, "+ 1." and "/ (AH - 1.)" in double narrowed to float,
// float sample coordinates, double bilinear weights and products, result narrowed to float.  The
// expressions below keep exactly those types and the same association, so the same compiler flags
// (nvcc default -fmad=true) give the same bits as the reference .cu built for sm_100a.
//
// Differences in structure (not arithmetic): one CTA per (roi, channel chunk); the roi geometry
// (sample coordinates, neighbour offsets, ratios) is computed once per CTA into shared memory
// instead of once per output element, output writes are fully coalesced, and the Avg variant never
// materialises the (AH+1)x(AW+1) intermediate in HBM.  Launch failures are returned, never exit().
#include "common.cuh"

constexpr int ROI_MAX_S = 64;  // max samples per axis (aligned size) supported by the smem tables

struct RoiAxis {  // per-sample data along one axis
  int start;      // floor index (clamped to size-2)
  float ratio;    // fractional offset (float, like the reference)
  int ok;         // in bounds
};

__device__ __forceinline__ void roi_axis(float start, float bin, int p, int size, RoiAxis& a) {
  float x = (float)(p)*bin + start;
  int s = fminf(floor(x), size - 2);
  a.ok = !(x < 0 || x >= size);
  a.start = s;
  a.ratio = x - (float)(s);
}

__device__ __forceinline__ void roi_setup(const float* __restrict__ roi, float spatial_scale, int height, int width,
                                          int AH, int AW, RoiAxis* hs, RoiAxis* ws, float& batch_ind) {
  // executed by threads 0 .. AH+AW-1
  batch_ind = roi[0];
  float roi_start_w = roi[1] * spatial_scale;
  float roi_start_h = roi[2] * spatial_scale;
  float roi_end_w = roi[3] * spatial_scale;
  float roi_end_h = roi[4] * spatial_scale;
  float roi_width = fmaxf(roi_end_w - roi_start_w + 1., 0.);
  float roi_height = fmaxf(roi_end_h - roi_start_h + 1., 0.);
  float bin_size_h = roi_height / (AH - 1.);
  float bin_size_w = roi_width / (AW - 1.);
  int t = threadIdx.x;
  if (t < AH)
    roi_axis(roi_start_h, bin_size_h, t, height, hs[t]);
  else if (t < AH + AW)
    roi_axis(roi_start_w, bin_size_w, t - AH, width, ws[t - AH]);
}

__device__ __forceinline__ float roi_sample(const float* __restrict__ plane, int width, const RoiAxis& h,
                                            const RoiAxis& w) {
  if (!(h.ok && w.ok)) return 0.f;
  float h_ratio = h.ratio, w_ratio = w.ratio;
  int upleft = h.start * width + w.start;
  int upright = upleft + 1, downleft = upleft + width, downright = downleft + 1;
  return plane[upleft] * (1. - h_ratio) * (1. - w_ratio) + plane[upright] * (1. - h_ratio) * w_ratio +
         plane[downleft] * h_ratio * (1. - w_ratio) + plane[downright] * h_ratio * w_ratio;
}

// grid: (num_rois, channel chunks)   block: 256
__global__ void __launch_bounds__(256) roi_align_fwd_kernel(const float* __restrict__ bottom, float spatial_scale,
                                                            int height, int width, int channels, int AH, int AW,
                                                            int ch_per_block, const float* __restrict__ rois,
                                                            float* __restrict__ top) {
  __shared__ RoiAxis hs[ROI_MAX_S], ws[ROI_MAX_S];
  __shared__ float s_batch;
  const int n = blockIdx.x;
  float bi;
  if (threadIdx.x < AH + AW) {
    roi_setup(rois + n * 5, spatial_scale, height, width, AH, AW, hs, ws, bi);
    if (threadIdx.x == 0) s_batch = bi;
  }
  __syncthreads();
  const int c0 = blockIdx.y * ch_per_block;
  const int c1 = min(channels, c0 + ch_per_block);
  const int S = AH * AW;
  int img_start = s_batch * channels * height * width;  // float arithmetic, like the reference
  float* out = top + ((long long)n * channels + c0) * S;
  const int total = (c1 - c0) * S;
  for (int i = threadIdx.x; i < total; i += 256) {
    int c = i / S, s = i - c * S;
    int ph = s / AW, pw = s - ph * AW;
    out[i] = roi_sample(bottom + img_start + (long long)(c0 + c) * height * width, width, hs[ph], ws[pw]);
  }
}

__global__ void __launch_bounds__(256) roi_align_bwd_kernel(const float* __restrict__ top_diff, float spatial_scale,
                                                            int height, int width, int channels, int AH, int AW,
                                                            int ch_per_block, const float* __restrict__ rois,
                                                            float* __restrict__ bottom_diff) {
  __shared__ RoiAxis hs[ROI_MAX_S], ws[ROI_MAX_S];
  __shared__ float s_batch;
  const int n = blockIdx.x;
  float bi;
  if (threadIdx.x < AH + AW) {
    roi_setup(rois + n * 5, spatial_scale, height, width, AH, AW, hs, ws, bi);
    if (threadIdx.x == 0) s_batch = bi;
  }
  __syncthreads();
  const int c0 = blockIdx.y * ch_per_block;
  const int c1 = min(channels, c0 + ch_per_block);
  const int S = AH * AW;
  int img_start = s_batch * channels * height * width;
  const float* g = top_diff + ((long long)n * channels + c0) * S;
  const int total = (c1 - c0) * S;
  for (int i = threadIdx.x; i < total; i += 256) {
    int c = i / S, s = i - c * S;
    int ph = s / AW, pw = s - ph * AW;
    const RoiAxis h = hs[ph], w = ws[pw];
    if (!(h.ok && w.ok)) continue;
    float h_ratio = h.ratio, w_ratio = w.ratio;
    float* plane = bottom_diff + img_start + (long long)(c0 + c) * height * width;
    int upleft = h.start * width + w.start;
    float gv = g[i];
    atomicAdd(plane + upleft, gv * (1. - h_ratio) * (1 - w_ratio));
    atomicAdd(plane + upleft + 1, gv * (1. - h_ratio) * w_ratio);
    atomicAdd(plane + upleft + width, gv * h_ratio * (1 - w_ratio));
    atomicAdd(plane + upleft + width + 1, gv * h_ratio * w_ratio);
  }
}

// fused RoIAlignAvg forward: samples (AH+1)x(AW+1) kept in shared memory, pooled AHxAW written.
// block handles CH_AVG channels of one roi.
constexpr int CH_AVG = 8;
__global__ void __launch_bounds__(256) roi_align_avg_fwd_kernel(const float* __restrict__ bottom,
                                                                float spatial_scale, int height, int width,
                                                                int channels, int AH, int AW,
                                                                const float* __restrict__ rois,
                                                                float* __restrict__ top) {
  extern __shared__ float samples[];  // [CH_AVG][(AH+1)*(AW+1)]
  __shared__ RoiAxis hs[ROI_MAX_S], ws[ROI_MAX_S];
  __shared__ float s_batch;
  const int n = blockIdx.x, SH = AH + 1, SW = AW + 1, S = SH * SW;
  float bi;
  if (threadIdx.x < SH + SW) {
    roi_setup(rois + n * 5, spatial_scale, height, width, SH, SW, hs, ws, bi);
    if (threadIdx.x == 0) s_batch = bi;
  }
  __syncthreads();
  const int c0 = blockIdx.y * CH_AVG;
  const int nc = min(CH_AVG, channels - c0);
  int img_start = s_batch * channels * height * width;
  for (int i = threadIdx.x; i < nc * S; i += 256) {
    int c = i / S, s = i - c * S;
    int ph = s / SW, pw = s - ph * SW;
    samples[i] = roi_sample(bottom + img_start + (long long)(c0 + c) * height * width, width, hs[ph], ws[pw]);
  }
  __syncthreads();
  const int O = AH * AW;
  float* out = top + ((long long)n * channels + c0) * O;
  for (int i = threadIdx.x; i < nc * O; i += 256) {
    int c = i / O, o = i - c * O;
    int oh = o / AW, ow = o - oh * AW;
    const float* sp = samples + c * S + oh * SW + ow;
    // avg_pool2d(kernel 2, stride 1): window summed row-major in fp32, divided by 4
    out[i] = (((sp[0] + sp[1]) + sp[SW]) + sp[SW + 1]) / 4.f;
  }
}

__global__ void __launch_bounds__(256) roi_align_avg_bwd_kernel(const float* __restrict__ top_diff,
                                                                float spatial_scale, int height, int width,
                                                                int channels, int AH, int AW,
                                                                const float* __restrict__ rois,
                                                                float* __restrict__ bottom_diff) {
  extern __shared__ float gp[];  // [CH_AVG][AH*AW] pooled gradients
  __shared__ RoiAxis hs[ROI_MAX_S], ws[ROI_MAX_S];
  __shared__ float s_batch;
  const int n = blockIdx.x, SH = AH + 1, SW = AW + 1, S = SH * SW, O = AH * AW;
  float bi;
  if (threadIdx.x < SH + SW) {
    roi_setup(rois + n * 5, spatial_scale, height, width, SH, SW, hs, ws, bi);
    if (threadIdx.x == 0) s_batch = bi;
  }
  const int c0 = blockIdx.y * CH_AVG;
  const int nc = min(CH_AVG, channels - c0);
  const float* g = top_diff + ((long long)n * channels + c0) * O;
  for (int i = threadIdx.x; i < nc * O; i += 256) gp[i] = g[i];
  __syncthreads();
  int img_start = s_batch * channels * height * width;
  for (int i = threadIdx.x; i < nc * S; i += 256) {
    int c = i / S, s = i - c * S;
    int ph = s / SW, pw = s - ph * SW;
    const RoiAxis h = hs[ph], w = ws[pw];
    if (!(h.ok && w.ok)) continue;
    // adjoint of the 2x2 stride-1 average: each sample feeds up to four pooled outputs
    float gv = 0.f;
    const float* gc = gp + c * O;
    for (int dh = -1; dh <= 0; ++dh)
      for (int dw = -1; dw <= 0; ++dw) {
        int oh = ph + dh, ow = pw + dw;
        if (oh >= 0 && oh < AH && ow >= 0 && ow < AW) gv += gc[oh * AW + ow] / 4.f;
      }
    float h_ratio = h.ratio, w_ratio = w.ratio;
    float* plane = bottom_diff + img_start + (long long)(c0 + c) * height * width;
    int upleft = h.start * width + w.start;
    atomicAdd(plane + upleft, gv * (1. - h_ratio) * (1 - w_ratio));
    atomicAdd(plane + upleft + 1, gv * (1. - h_ratio) * w_ratio);
    atomicAdd(plane + upleft + width, gv * h_ratio * (1 - w_ratio));
    atomicAdd(plane + upleft + width + 1, gv * h_ratio * w_ratio);
  }
}

// ------------------------------------------------------------------------------------------------
// Channels-last fused RoIAlignAvg: features NHWC [B][H][W][C] (C % 4 == 0), out NHWC [R][AH][AW][C].
// Same sample coordinates / bilinear arithmetic as above (bit-identical values); what changes is the memory side:
// a thread owns FOUR consecutive channels, so each of the four neighbour reads of a sample is one coalesced 16-byte
// load per thread (a warp reads 512 contiguous bytes of one pixel) instead of 4-byte gathers from 4 * C planes, and
// the pooled row is written with coalesced 16-byte stores.  One CTA = (roi, pooled output row): sample rows oh and
// oh + 1 (2 x (AW+1) samples per channel) are combined in registers; nothing but the result touches HBM.
// The discriminators' feature maps are produced channels-last by the conv kernels, so this variant also removes the
// NHWC -> NCHW -> NHWC round trip around the reference-ABI op (model.py:1241, 1307 call sites).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float roi_bilerp(float ul, float ur, float dl, float dr, float h_ratio, float w_ratio) {
  return ul * (1. - h_ratio) * (1. - w_ratio) + ur * (1. - h_ratio) * w_ratio + dl * h_ratio * (1. - w_ratio) +
         dr * h_ratio * w_ratio;
}
__global__ void __launch_bounds__(256) roi_align_avg_nhwc_fwd_kernel(const float* __restrict__ feat, float spatial_scale,
                                                                     int height, int width, int C, int AH, int AW,
                                                                     const float* __restrict__ rois,
                                                                     float* __restrict__ out) {
  __shared__ RoiAxis hs[ROI_MAX_S], ws[ROI_MAX_S];
  __shared__ float s_batch;
  const int n = blockIdx.x, oh = blockIdx.y, SH = AH + 1, SW = AW + 1;
  float bi;
  if (threadIdx.x < SH + SW) {
    roi_setup(rois + n * 5, spatial_scale, height, width, SH, SW, hs, ws, bi);
    if (threadIdx.x == 0) s_batch = bi;
  }
  __syncthreads();
  const RoiAxis h0 = hs[oh], h1 = hs[oh + 1];
  const float* img = feat + (long long)(int)s_batch * height * width * C;
  float* orow = out + ((long long)n * AH + oh) * AW * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 top_prev = make_float4(0.f, 0.f, 0.f, 0.f), bot_prev = top_prev;
    for (int pw = 0; pw < SW; ++pw) {
      const RoiAxis w = ws[pw];
      float4 top = make_float4(0.f, 0.f, 0.f, 0.f), bot = top;
      if (w.ok && h0.ok) {
        const float* p = img + ((long long)h0.start * width + w.start) * C + c;
        const float4 ul = ldg4(p), ur = ldg4(p + C), dl = ldg4(p + (long long)width * C), dr = ldg4(p + (long long)width * C + C);
        top.x = roi_bilerp(ul.x, ur.x, dl.x, dr.x, h0.ratio, w.ratio);
        top.y = roi_bilerp(ul.y, ur.y, dl.y, dr.y, h0.ratio, w.ratio);
        top.z = roi_bilerp(ul.z, ur.z, dl.z, dr.z, h0.ratio, w.ratio);
        top.w = roi_bilerp(ul.w, ur.w, dl.w, dr.w, h0.ratio, w.ratio);
      }
      if (w.ok && h1.ok) {
        const float* p = img + ((long long)h1.start * width + w.start) * C + c;
        const float4 ul = ldg4(p), ur = ldg4(p + C), dl = ldg4(p + (long long)width * C), dr = ldg4(p + (long long)width * C + C);
        bot.x = roi_bilerp(ul.x, ur.x, dl.x, dr.x, h1.ratio, w.ratio);
        bot.y = roi_bilerp(ul.y, ur.y, dl.y, dr.y, h1.ratio, w.ratio);
        bot.z = roi_bilerp(ul.z, ur.z, dl.z, dr.z, h1.ratio, w.ratio);
        bot.w = roi_bilerp(ul.w, ur.w, dl.w, dr.w, h1.ratio, w.ratio);
      }
      if (pw > 0) {
        // avg_pool2d(kernel 2, stride 1): window summed row-major in fp32, divided by 4 (same order as the NCHW kernel)
        float4 o;
        o.x = (((top_prev.x + top.x) + bot_prev.x) + bot.x) / 4.f;
        o.y = (((top_prev.y + top.y) + bot_prev.y) + bot.y) / 4.f;
        o.z = (((top_prev.z + top.z) + bot_prev.z) + bot.z) / 4.f;
        o.w = (((top_prev.w + top.w) + bot_prev.w) + bot.w) / 4.f;
        st4(orow + (long long)(pw - 1) * C + c, o);
      }
      top_prev = top;
      bot_prev = bot;
    }
  }
}

// adjoint: one CTA = (roi, sample row ph); grad_features NHWC, zero-filled by the caller, 16-byte vector atomics
__global__ void __launch_bounds__(256) roi_align_avg_nhwc_bwd_kernel(const float* __restrict__ gout, float spatial_scale,
                                                                     int height, int width, int C, int AH, int AW,
                                                                     const float* __restrict__ rois,
                                                                     float* __restrict__ gfeat) {
  __shared__ RoiAxis hs[ROI_MAX_S], ws[ROI_MAX_S];
  __shared__ float s_batch;
  const int n = blockIdx.x, ph = blockIdx.y, SH = AH + 1, SW = AW + 1;
  float bi;
  if (threadIdx.x < SH + SW) {
    roi_setup(rois + n * 5, spatial_scale, height, width, SH, SW, hs, ws, bi);
    if (threadIdx.x == 0) s_batch = bi;
  }
  __syncthreads();
  const RoiAxis h = hs[ph];
  if (!h.ok) return;
  float* img = gfeat + (long long)(int)s_batch * height * width * C;
  const float* g = gout + (long long)n * AH * AW * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    for (int pw = 0; pw < SW; ++pw) {
      const RoiAxis w = ws[pw];
      if (!w.ok) continue;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int dh = -1; dh <= 0; ++dh)
        for (int dw = -1; dw <= 0; ++dw) {
          const int oh = ph + dh, ow = pw + dw;
          if (oh >= 0 && oh < AH && ow >= 0 && ow < AW) {
            const float4 t = ldg4(g + ((long long)oh * AW + ow) * C + c);
            gv.x += t.x / 4.f; gv.y += t.y / 4.f; gv.z += t.z / 4.f; gv.w += t.w / 4.f;
          }
        }
      const float hr = h.ratio, wr = w.ratio;
      const float k00 = (1. - hr) * (1 - wr), k01 = (1. - hr) * wr, k10 = hr * (1 - wr), k11 = (double)hr * wr;
      float* p = img + ((long long)h.start * width + w.start) * C + c;
      atomicAdd(reinterpret_cast<float4*>(p), make_float4(gv.x * k00, gv.y * k00, gv.z * k00, gv.w * k00));
      atomicAdd(reinterpret_cast<float4*>(p + C), make_float4(gv.x * k01, gv.y * k01, gv.z * k01, gv.w * k01));
      atomicAdd(reinterpret_cast<float4*>(p + (long long)width * C),
                make_float4(gv.x * k10, gv.y * k10, gv.z * k10, gv.w * k10));
      atomicAdd(reinterpret_cast<float4*>(p + (long long)width * C + C),
                make_float4(gv.x * k11, gv.y * k11, gv.z * k11, gv.w * k11));
    }
  }
}

static int roi_ch_per_block(int num_rois, int channels, int S) {
  // enough CTAs to fill 148 SMs a few times over, at least ~2 passes of 256 threads per CTA
  int cpb = channels;
  while (cpb > 1 && (long long)num_rois * og_cdiv(channels, cpb) < 148 * 8 && cpb * S > 512) cpb = (cpb + 1) / 2;
  return cpb;
}

// ---- reference C ABI (same names, argument order and meaning as roi_align_kernel.h:13-27) ----
OG_API int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                  const int height, const int width, const int channels, const int aligned_height,
                                  const int aligned_width, const float* bottom_rois, float* top_data,
                                  cudaStream_t stream) {
  if (aligned_height > ROI_MAX_S || aligned_width > ROI_MAX_S || aligned_height + aligned_width > 256) return 0;
  if (num_rois == 0 || channels == 0) return 1;
  int cpb = roi_ch_per_block(num_rois, channels, aligned_height * aligned_width);
  dim3 grid(num_rois, og_cdiv(channels, cpb));
  roi_align_fwd_kernel<<<grid, 256, 0, stream>>>(bottom_data, spatial_scale, height, width, channels, aligned_height,
                                                 aligned_width, cpb, bottom_rois, top_data);
  return cudaGetLastError() == cudaSuccess ? 1 : 0;  // the reference returns 1 on success
}

OG_API int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                   const int num_rois, const int height, const int width, const int channels,
                                   const int aligned_height, const int aligned_width, const float* bottom_rois,
                                   float* bottom_diff, cudaStream_t stream) {
  (void)batch_size;
  if (aligned_height > ROI_MAX_S || aligned_width > ROI_MAX_S || aligned_height + aligned_width > 256) return 0;
  if (num_rois == 0 || channels == 0) return 1;
  int cpb = roi_ch_per_block(num_rois, channels, aligned_height * aligned_width);
  dim3 grid(num_rois, og_cdiv(channels, cpb));
  roi_align_bwd_kernel<<<grid, 256, 0, stream>>>(top_diff, spatial_scale, height, width, channels, aligned_height,
                                                 aligned_width, cpb, bottom_rois, bottom_diff);
  return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

// ---- fused RoIAlignAvg (0 on success like the rest of the og_* ABI) ----
OG_API int og_roi_align_avg_fwd(const float* features, int height, int width, int channels, const float* rois,
                                int num_rois, int AH, int AW, float spatial_scale, float* out, cudaStream_t stream) {
  if (AH + 1 > ROI_MAX_S || AW + 1 > ROI_MAX_S || AH + AW + 2 > 256) return (int)cudaErrorInvalidValue;
  if (num_rois == 0 || channels == 0) return 0;
  dim3 grid(num_rois, og_cdiv(channels, CH_AVG));
  size_t sm = sizeof(float) * CH_AVG * (AH + 1) * (AW + 1);
  roi_align_avg_fwd_kernel<<<grid, 256, sm, stream>>>(features, spatial_scale, height, width, channels, AH, AW, rois, out);
  OG_RETURN_LAST_ERROR();
}
// grad_features must be zero-filled by the caller (same ownership rule as the reference, functions/roi_align.py:42-43)
OG_API int og_roi_align_avg_bwd(const float* grad_out, int height, int width, int channels, const float* rois,
                                int num_rois, int AH, int AW, float spatial_scale, float* grad_features,
                                cudaStream_t stream) {
  if (AH + 1 > ROI_MAX_S || AW + 1 > ROI_MAX_S || AH + AW + 2 > 256) return (int)cudaErrorInvalidValue;
  if (num_rois == 0 || channels == 0) return 0;
  dim3 grid(num_rois, og_cdiv(channels, CH_AVG));
  size_t sm = sizeof(float) * CH_AVG * AH * AW;
  roi_align_avg_bwd_kernel<<<grid, 256, sm, stream>>>(grad_out, spatial_scale, height, width, channels, AH, AW, rois,
                                                      grad_features);
  OG_RETURN_LAST_ERROR();
}

// ---- channels-last fused RoIAlignAvg (features / out NHWC; C % 4 == 0) ----
OG_API int og_roi_align_avg_nhwc_fwd(const float* features, int height, int width, int channels, const float* rois,
                                     int num_rois, int AH, int AW, float spatial_scale, float* out, cudaStream_t stream) {
  if (AH + 1 > ROI_MAX_S || AW + 1 > ROI_MAX_S || AH + AW + 2 > 64 || channels % 4) return (int)cudaErrorInvalidValue;
  if (num_rois == 0 || channels == 0) return 0;
  int threads = (channels / 4 + 31) / 32 * 32;
  threads = threads < 64 ? 64 : (threads > 256 ? 256 : threads);
  roi_align_avg_nhwc_fwd_kernel<<<dim3(num_rois, AH), threads, 0, stream>>>(features, spatial_scale, height, width,
                                                                           channels, AH, AW, rois, out);
  OG_RETURN_LAST_ERROR();
}
// grad_features must be zero-filled by the caller
OG_API int og_roi_align_avg_nhwc_bwd(const float* grad_out, int height, int width, int channels, const float* rois,
                                     int num_rois, int AH, int AW, float spatial_scale, float* grad_features,
                                     cudaStream_t stream) {
  if (AH + 1 > ROI_MAX_S || AW + 1 > ROI_MAX_S || AH + AW + 2 > 64 || channels % 4) return (int)cudaErrorInvalidValue;
  if (num_rois == 0 || channels == 0) return 0;
  int threads = (channels / 4 + 31) / 32 * 32;
  threads = threads < 64 ? 64 : (threads > 256 ? 256 : threads);
  roi_align_avg_nhwc_bwd_kernel<<<dim3(num_rois, AH + 1), threads, 0, stream>>>(grad_out, spatial_scale, height, width,
                                                                               channels, AH, AW, rois, grad_features);
  OG_RETURN_LAST_ERROR();
}
