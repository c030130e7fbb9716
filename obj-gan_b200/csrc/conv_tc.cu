// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::tf32) with TMEM accumulators,
// operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier producer/consumer pipeline.
//
// Covers every stride-1 contraction of the hot path through a "tap list":
//     out[pix, :] = sum_t  src[pix + (dh_t, dw_t), :] * W[widx_t]          (src out-of-bounds = 0 via TMA fill)
// which is the forward of the reflection-padded 3x3 convs of HmapResBlock (model.py:63-81; the halo is
// materialised by og_prep_split), the zero-padded 3x3 convs (model.py:36-39, 1020-1048), their input
// gradients (taps mirrored, operand re-packed) and the four phases of nearest-2x-upsample + conv3x3
// (upBlock, model.py:43-49: tap offsets at low resolution, strided output pixels).
//
// Precision: fp32 parity (1e-3 end to end) is not reachable with one TF32 product (2^-11 operand rounding,
// ~35 stacked convs), so by default each product is error-compensated:  a*b ~= ah*bh + al*bh + ah*bl  with
// ah = a rounded down to tf32, al = a - ah (exact), three MMAs into the same fp32 TMEM accumulator ("3xTF32").
// nsplit = 1 runs the plain single-TF32 product (reported separately, never the parity mode).
//
// GEMM tiling: one CTA = 128 output pixels (a TN x TH x TW patch) x BN output channels (BN % 16 == 0,
// <= 256); K loop over taps x 32-channel chunks; STAGES-deep smem ring of {A_hi, A_lo, B_hi, B_lo}.
// Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> global, one TMEM lane quadrant each).
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

// round-to-nearest tf32 (unbiased; truncation would bias every product the same way and the bias adds up
// linearly over the reduction).  lo = v - hi is exact in fp32 and is itself rounded to tf32 so that the tensor
// core's own operand truncation is a no-op: |v - hi - lo| <= 2^-22 |v|.
__device__ __forceinline__ float tf32_rn(float v) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float tf32_hi(float v) { return tf32_rn(v); }
__device__ __forceinline__ float tf32_lo(float v, float hi) { return tf32_rn(v - hi); }

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;              // tf32 elements per k-chunk = one 128-byte swizzle row
constexpr int TC_MAX_TAPS = 36;
constexpr int TC_THREADS = 192;

struct TcTaps {
  int n;
  int dh[TC_MAX_TAPS], dw[TC_MAX_TAPS], dn[TC_MAX_TAPS], widx[TC_MAX_TAPS];
};

struct TcParams {
  int N, OH, OW;          // GEMM-row pixel grid (before the output stride/phase mapping)
  int TN, TH, TW;         // tile patch: TN*TH*TW == 128
  int tiles_h, tiles_w;   // tiles per image along h / w
  int cchunks;            // ceil(C / 32)
  int K;                  // output channels actually stored (<= gridDim.y * BN)
  int nsplit;             // 1 or 3
  float* y;
  long long ysn, ysh, ysw;  // output strides (elements) of the FULL-resolution output tensor
  int osy, osx, opy, opx;   // output pixel = (osy * h + opy, osx * w + opx)
  int OHfull, OWfull;       // bounds of the output tensor
  const float* bias;        // [K] or null
  int act;
  float slope;
  int ksplit;               // > 1: gridDim.z CTAs share one output tile, each reduces a slice of the (tap, chunk) loop
  int tall;                 // taps come in groups of 3 consecutive source rows: load one (TH+2)-row A patch per group
  TcTaps taps;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tc_act(float v, int act, float slope) {
  if (act == OG_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == OG_ACT_TANH) return tanhf(v);
  if (act == OG_ACT_SIGMOID) return og_sigmoid(v);
  return v;
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, dense
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;   // n_dim
  d |= (uint32_t)(M >> 4) << 24;   // m_dim
  return d;
}

// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const TcParams p) {
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
  constexpr uint32_t B_BYTES = BN * TC_BK * 4;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int t = blockIdx.x;
  const int tw_i = t % p.tiles_w;
  t /= p.tiles_w;
  const int th_i = t % p.tiles_h;
  const int tn_i = t / p.tiles_h;
  const int n0 = tn_i * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
  const int col0 = blockIdx.y * BN;
  const int nk_total = p.taps.n * p.cchunks;
  const int per_split = (nk_total + p.ksplit - 1) / p.ksplit;
  const int it_beg = blockIdx.z * per_split;
  const int nk = min(nk_total, it_beg + per_split) - it_beg;
  const uint32_t tx_bytes = (p.nsplit == 3) ? STAGE_BYTES : (A_BYTES + B_BYTES);
  if (nk <= 0) return;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_ah);
    prefetch_tmap(&map_bh);
    if (p.nsplit == 3) {
      prefetch_tmap(&map_al);
      prefetch_tmap(&map_bl);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nk; ++it) {
        const int tap = (it_beg + it) / p.cchunks, cc = (it_beg + it) - tap * p.cchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], tx_bytes);
        const int c0 = cc * TC_BK, ws = w0 + p.taps.dw[tap], hs = h0 + p.taps.dh[tap], widx = p.taps.widx[tap];
        const int ns = n0 + p.taps.dn[tap];   // image offset: selects a space-to-depth phase block of the source
        tma_load_4d(sa, &map_ah, &full_bar[stage], c0, ws, hs, ns);
        tma_load_3d(sa + 2 * A_BYTES, &map_bh, &full_bar[stage], c0, col0, widx);
        if (p.nsplit == 3) {
          tma_load_4d(sa + A_BYTES, &map_al, &full_bar[stage], c0, ws, hs, ns);
          tma_load_3d(sa + 2 * A_BYTES + B_BYTES, &map_bl, &full_bar[stage], c0, col0, widx);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nk; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t ah = make_desc_sw128(sa), al = make_desc_sw128(sa + A_BYTES);
        const uint64_t bh = make_desc_sw128(sa + 2 * A_BYTES), bl = make_desc_sw128(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);   // advance 32 bytes inside the swizzle row
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          if (p.nsplit == 3) {
            umma_tf32(tmem_base, al + koff, bh + koff, idesc, acc);   // small terms first
            umma_tf32(tmem_base, ah + koff, bl + koff, idesc, 1u);
            umma_tf32(tmem_base, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma_tf32(tmem_base, ah + koff, bh + koff, idesc, acc);
          }
        }
        umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tmem_full_bar);        // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;               // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;        // tile row = output pixel within the patch
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const int n = n0 + tn, h = h0 + th, w = w0 + tw;
    const int oy = p.osy * h + p.opy, ox = p.osx * w + p.opx;
    const bool row_ok = (n < p.N) && (h < p.OH) && (w < p.OW) && (oy < p.OHfull) && (ox < p.OWfull);
    float* yrow = p.y + (long long)n * p.ysn + (long long)oy * p.ysh + (long long)ox * p.ysw + col0;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (c + j < BN && col0 + c + j < p.K) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            if (p.bias) {
              float4 b = ldg4(p.bias + col0 + c + j);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (p.act != OG_ACT_NONE) {
              v.x = tc_act(v.x, p.act, p.slope); v.y = tc_act(v.y, p.act, p.slope);
              v.z = tc_act(v.z, p.act, p.slope); v.w = tc_act(v.w, p.act, p.slope);
            }
            if (p.ksplit > 1) {   // partial sum of this K slice (output zero-filled by the launcher)
              atomicAdd(yrow + c + j, v.x); atomicAdd(yrow + c + j + 1, v.y);
              atomicAdd(yrow + c + j + 2, v.z); atomicAdd(yrow + c + j + 3, v.w);
            } else {
              *reinterpret_cast<float4*>(yrow + c + j) = v;
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Two-tile variant for large problems: one CTA computes TWO 128-pixel tiles (adjacent tile indices) against
// the same weight tile, with two TMEM accumulators.  The single-tile kernel above is bound by L2->SM traffic
// (85 KB of operands per 12 MMAs at BN = 208, measured 9.3 TB/s = the LTS throughput cap, tensor pipe 41 %);
// sharing each B (weight) stage between two A (pixel) tiles cuts that to 58.5 KB per tile.  Separate smem rings:
// 3 A slots (hi+lo, 32 KB each) and 2 B slots.
// ------------------------------------------------------------------------------------------------
// TALL = true (3x3 stride-1 taps, 8x16 pixel patches): the three kernel rows of one kernel column read the same
// pixels shifted by whole patch rows, so ONE (8+2)-row A patch per (kernel column, channel chunk) serves three
// k-iterations through descriptor offsets of 16 rows (2 KB, swizzle-phase preserving): A traffic / 2.4.
template <int BN, bool TALL>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                const TcParams p) {
  constexpr int AS = 3, BS = 2;
  constexpr uint32_t A_BYTES = (TALL ? 160 : TC_BM) * TC_BK * 4;   // one of hi / lo: 128 rows, or 10 x 16 patch rows
  constexpr uint32_t B_BYTES = BN * TC_BK * 4;
  constexpr uint32_t A_SLOT = 2 * A_BYTES, B_SLOT = 2 * B_BYTES;
  constexpr uint32_t ACC_STRIDE = 256;               // TMEM column offset of the second accumulator
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + AS * A_SLOT;
  __shared__ __align__(8) uint64_t fullA[AS], emptyA[AS], fullB[BS], emptyB[BS], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * BN;
  const int nk = p.taps.n * p.cchunks;
  const bool split3 = p.nsplit == 3;
  // the two tiles of this CTA
  int tn0[2], th0[2], tw0[2];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    int t = blockIdx.x * 2 + hf;
    const int tw_i = t % p.tiles_w;
    t /= p.tiles_w;
    const int th_i = t % p.tiles_h;
    const int tn_i = t / p.tiles_h;     // may run past the last image for an odd tile count: TMA then zero-fills
    tn0[hf] = tn_i * p.TN; th0[hf] = th_i * p.TH; tw0[hf] = tw_i * p.TW;
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_ah);
    prefetch_tmap(&map_bh);
    for (int s = 0; s < AS; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto load_b = [&](int tap, int c0) {
        mbar_wait(&emptyB[bs], bph ^ 1);
        uint8_t* sb = smem_b + (size_t)bs * B_SLOT;
        mbar_expect_tx(&fullB[bs], split3 ? B_SLOT : B_BYTES);
        tma_load_3d(sb, &map_bh, &fullB[bs], c0, col0, p.taps.widx[tap]);
        if (split3) tma_load_3d(sb + B_BYTES, &map_bl, &fullB[bs], c0, col0, p.taps.widx[tap]);
        if (++bs == BS) { bs = 0; bph ^= 1; }
      };
      auto load_a = [&](int hf, int tap, int c0) {
        mbar_wait(&emptyA[as], aph ^ 1);
        uint8_t* sa = smem_a + (size_t)as * A_SLOT;
        mbar_expect_tx(&fullA[as], split3 ? A_SLOT : A_BYTES);
        const int ws = tw0[hf] + p.taps.dw[tap], hs = th0[hf] + p.taps.dh[tap], ns = tn0[hf] + p.taps.dn[tap];
        tma_load_4d(sa, &map_ah, &fullA[as], c0, ws, hs, ns);
        if (split3) tma_load_4d(sa + A_BYTES, &map_al, &fullA[as], c0, ws, hs, ns);
        if (++as == AS) { as = 0; aph ^= 1; }
      };
      if (TALL) {
        const int groups = (p.taps.n / 3) * p.cchunks;
        for (int g = 0; g < groups; ++g) {
          const int kwi = g / p.cchunks, cc = g - kwi * p.cchunks;
          const int c0 = cc * TC_BK, t0 = kwi * 3;       // taps t0, t0+1, t0+2: same column, consecutive rows
          load_a(0, t0, c0);
          load_b(t0, c0);                                // first weight stage before the second patch: the MMAs of
          load_a(1, t0, c0);                             // half 0 can start while half 1's patch is in flight
          load_b(t0 + 1, c0);
          load_b(t0 + 2, c0);
        }
      } else {
        for (int it = 0; it < nk; ++it) {
          const int tap = it / p.cchunks, cc = it - tap * p.cchunks;
          const int c0 = cc * TC_BK;
          load_b(tap, c0);
          load_a(0, tap, c0);
          load_a(1, tap, c0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto mma_half = [&](uint32_t sa, uint32_t rowoff, uint32_t sb, int hf, bool first) {
        const uint64_t ah = make_desc_sw128(sa + rowoff), al = make_desc_sw128(sa + A_BYTES + rowoff);
        const uint64_t bh = make_desc_sw128(sb), bl = make_desc_sw128(sb + B_BYTES);
        const uint32_t d = tmem_base + hf * ACC_STRIDE;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);
          const uint32_t acc = (!first || k > 0) ? 1u : 0u;
          if (split3) {
            umma_tf32(d, al + koff, bh + koff, idesc, acc);
            umma_tf32(d, ah + koff, bl + koff, idesc, 1u);
            umma_tf32(d, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma_tf32(d, ah + koff, bh + koff, idesc, acc);
          }
        }
      };
      if (TALL) {
        const int groups = (p.taps.n / 3) * p.cchunks;
        for (int g = 0; g < groups; ++g) {
          const int a0 = as, a1 = (as + 1) % AS;
          const uint32_t aph0 = aph, aph1 = (as + 1 == AS) ? (aph ^ 1) : aph;
          for (int j = 0; j < 3; ++j) {
            mbar_wait(&fullB[bs], bph);
            const uint32_t sb = smem_u32(smem_b + (size_t)bs * B_SLOT);
            if (j == 0) mbar_wait(&fullA[a0], aph0);
            tc_fence_after();
            mma_half(smem_u32(smem_a + (size_t)a0 * A_SLOT), j * 16 * 128, sb, 0, g == 0 && j == 0);
            if (j == 0) {
              mbar_wait(&fullA[a1], aph1);
              tc_fence_after();
            }
            mma_half(smem_u32(smem_a + (size_t)a1 * A_SLOT), j * 16 * 128, sb, 1, g == 0 && j == 0);
            umma_commit(&emptyB[bs]);
            if (++bs == BS) { bs = 0; bph ^= 1; }
          }
          umma_commit(&emptyA[a0]);
          umma_commit(&emptyA[a1]);
          for (int k = 0; k < 2; ++k)
            if (++as == AS) { as = 0; aph ^= 1; }
        }
      } else {
        for (int it = 0; it < nk; ++it) {
          mbar_wait(&fullB[bs], bph);
          const uint32_t sb = smem_u32(smem_b + (size_t)bs * B_SLOT);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&fullA[as], aph);
            tc_fence_after();
            mma_half(smem_u32(smem_a + (size_t)as * A_SLOT), 0, sb, hf, it == 0);
            umma_commit(&emptyA[as]);
            if (++as == AS) { as = 0; aph ^= 1; }
          }
          umma_commit(&emptyB[bs]);
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      const int n = tn0[hf] + tn, h = th0[hf] + th, w = tw0[hf] + tw;
      const int oy = p.osy * h + p.opy, ox = p.osx * w + p.opx;
      const bool row_ok = (n < p.N) && (h < p.OH) && (w < p.OW) && (oy < p.OHfull) && (ox < p.OWfull);
      float* yrow = p.y + (long long)n * p.ysn + (long long)oy * p.ysh + (long long)ox * p.ysw + col0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * ACC_STRIDE + c), r);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c + j < BN && col0 + c + j < p.K) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                     __uint_as_float(r[j + 3]));
              if (p.bias) {
                float4 b = ldg4(p.bias + col0 + c + j);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              if (p.act != OG_ACT_NONE) {
                v.x = tc_act(v.x, p.act, p.slope); v.y = tc_act(v.y, p.act, p.slope);
                v.z = tc_act(v.z, p.act, p.slope); v.w = tc_act(v.w, p.act, p.slope);
              }
              *reinterpret_cast<float4*>(yrow + c + j) = v;
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp32 tensor viewed as [d3][d2][d1][d0] (d0 contiguous); strides in elements for d1..d3
int make_map(CUtensorMap* m, const float* base, int rank, const unsigned long long* dims,
             const unsigned long long* strides_elems, const unsigned* box, bool atom32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_elems[i] * sizeof(float);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

template <int BN, int STAGES>
int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
              const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * TC_BM * TC_BK * 4 + 2 * BN * TC_BK * 4) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  conv_tc_kernel<BN, STAGES><<<grid, TC_THREADS, smem, stream>>>(ah, al, bh, bl, p);
  return (int)cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// weight gradient on the tensor cores.
//   dW[t][co][ci] += sum_{n,h,w} G[n,h,w,co] * X[n, h+dh_t, w+dw_t, ci]
// GEMM per tap: M = co, N = ci, reduction over pixels.  Both operands are read straight from the NHWC hi/lo
// tensors the forward / input-gradient kernels already use, i.e. they are "MN-major" for the MMA (channel index
// contiguous, reduction index = smem row).  For 32-bit operands tcgen05 accepts that only in the 128-byte
// swizzle with 32-byte atoms (descriptor layout type 1; the ordinary 128B swizzle silently yields zeros --
// tests/probes/probe_mnmajor.cu), which TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  A smem operand
// is a stack of "slabs" (32 channels x 32 pixels = 32 rows of 128 bytes, LBO = 4096 apart; 4-row groups SBO =
// 512 apart); one 5-D TMA box (32 channels-in-slab, w, h, n, slab) fills all slabs of an operand, the tap shift
// is a plain coordinate offset with out-of-bounds zero fill.  One MMA consumes 8 pixel rows (K = 8).
// Grid: (co tiles x ci tiles, taps, pixel splits); partial sums are reduced with fp32 atomics.
// ------------------------------------------------------------------------------------------------
struct TcWgradParams {
  int N, OH, OW;        // pixel grid of G
  int cw, chh, cn;      // pixel chunk of one stage: cw x chh x cn = 32 pixels (w fastest)
  int wchunks, hchunks; // chunks per row / per image column
  int total_chunks;
  int chunks_per_cta;
  int Kp, C;            // rows (co) and columns (ci) of each dW[t]
  int cotiles;          // blockIdx.x = citile * cotiles + cotile
  int nsplit;
  float* dw;            // [ntaps_out][Kp][C]
  // entry e: dW[out[e]] += G[n + gdn[e], h, w]^T * X[n + xdn[e], h + xdh[e], w + xdw[e]]
  int gdn[TC_MAX_TAPS], xdh[TC_MAX_TAPS], xdw[TC_MAX_TAPS], xdn[TC_MAX_TAPS], out[TC_MAX_TAPS];
};

constexpr int WG_PIX = 32;                     // pixels per stage (4 MMA k-steps)
constexpr uint32_t WG_SLAB = WG_PIX * 128;     // bytes of one slab: 32 pixels x 32 channels

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// MN-major slab stack, 128B swizzle with 32-byte atoms: LBO = slab pitch, SBO = 4 rows * 128 B
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WG_SLAB >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                          // layout type: SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32_mn(int M, int N) {
  return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);   // A and B both MN-major
}

template <int BNW, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_gh, const __grid_constant__ CUtensorMap map_gl,
                     const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const TcWgradParams p) {
  constexpr int NSB = (BNW + 31) / 32;                    // slabs of the X operand
  constexpr uint32_t A_BYTES = 4 * WG_SLAB;               // 128 co = 4 slabs = 16 KB
  constexpr uint32_t B_BYTES = NSB * WG_SLAB;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = BNW <= 32 ? 32 : BNW <= 64 ? 64 : BNW <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cotile = blockIdx.x % p.cotiles, citile = blockIdx.x / p.cotiles, tap = blockIdx.y;
  const int ch_beg = blockIdx.z * p.chunks_per_cta;
  const int ch_end = min(p.total_chunks, ch_beg + p.chunks_per_cta);
  const int nk = ch_end - ch_beg;
  const uint32_t tx_bytes = (p.nsplit == 3) ? STAGE_BYTES : (A_BYTES + B_BYTES);
  if (nk <= 0) return;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_gh);
    prefetch_tmap(&map_xh);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int gdn = p.gdn[tap], xdh = p.xdh[tap], xdw = p.xdw[tap], xdn = p.xdn[tap];
      const int aslab = cotile * 4, bslab = citile * NSB;
      for (int it = 0; it < nk; ++it) {
        const int ch = ch_beg + it;
        const int wc = ch % p.wchunks;
        const int t2 = ch / p.wchunks;
        const int hc = t2 % p.hchunks;
        const int w0 = wc * p.cw, h = hc * p.chh, n = (t2 / p.hchunks) * p.cn;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], tx_bytes);
        tma_load_5d(sa, &map_gh, &full_bar[stage], 0, w0, h, n + gdn, aslab);
        tma_load_5d(sa + 2 * A_BYTES, &map_xh, &full_bar[stage], 0, w0 + xdw, h + xdh, n + xdn, bslab);
        if (p.nsplit == 3) {
          tma_load_5d(sa + A_BYTES, &map_gl, &full_bar[stage], 0, w0, h, n + gdn, aslab);
          tma_load_5d(sa + 2 * A_BYTES + B_BYTES, &map_xl, &full_bar[stage], 0, w0 + xdw, h + xdh, n + xdn, bslab);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(TC_BM, BNW);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nk; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t ah = make_desc_mn(sa), al = make_desc_mn(sa + A_BYTES);
        const uint64_t bh = make_desc_mn(sa + 2 * A_BYTES), bl = make_desc_mn(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
        for (int k = 0; k < WG_PIX / 8; ++k) {
          const uint64_t koff = (uint64_t)((k * 8 * 128) >> 4);     // 8 pixel rows of 128 bytes
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          if (p.nsplit == 3) {
            umma_tf32(tmem_base, al + koff, bh + koff, idesc, acc);
            umma_tf32(tmem_base, ah + koff, bl + koff, idesc, 1u);
            umma_tf32(tmem_base, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma_tf32(tmem_base, ah + koff, bh + koff, idesc, acc);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int co = cotile * TC_BM + q * 32 + lane;
    const bool row_ok = co < p.Kp;
    float* drow = p.dw + ((long long)p.out[tap] * p.Kp + co) * p.C + citile * BNW;
    const int cleft = p.C - citile * BNW;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BNW; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < BNW && c + j < cleft) atomicAdd(drow + c + j, __uint_as_float(r[j]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// two co-tiles (256 output channels) per CTA sharing every X stage -- same idea as conv_tc2_kernel
template <int BNW>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad2_kernel(const __grid_constant__ CUtensorMap map_gh, const __grid_constant__ CUtensorMap map_gl,
                      const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                      const TcWgradParams p) {
  constexpr int AS = 3, BS = 2;
  constexpr int NSB = (BNW + 31) / 32;
  constexpr uint32_t A_BYTES = 4 * WG_SLAB;
  constexpr uint32_t B_BYTES = NSB * WG_SLAB;
  constexpr uint32_t A_SLOT = 2 * A_BYTES, B_SLOT = 2 * B_BYTES;
  constexpr uint32_t ACC_STRIDE = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + AS * A_SLOT;
  __shared__ __align__(8) uint64_t fullA[AS], emptyA[AS], fullB[BS], emptyB[BS], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int copairs = (p.cotiles + 1) / 2;
  const int copair = blockIdx.x % copairs, citile = blockIdx.x / copairs, tap = blockIdx.y;
  const int ch_beg = blockIdx.z * p.chunks_per_cta;
  const int ch_end = min(p.total_chunks, ch_beg + p.chunks_per_cta);
  const int nk = ch_end - ch_beg;
  const bool split3 = p.nsplit == 3;
  if (nk <= 0) return;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_gh);
    prefetch_tmap(&map_xh);
    for (int s = 0; s < AS; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const int gdn = p.gdn[tap], xdh = p.xdh[tap], xdw = p.xdw[tap], xdn = p.xdn[tap];
      const int bslab = citile * NSB;
      for (int it = 0; it < nk; ++it) {
        const int ch = ch_beg + it;
        const int wc = ch % p.wchunks;
        const int t2 = ch / p.wchunks;
        const int hc = t2 % p.hchunks;
        const int w0 = wc * p.cw, h = hc * p.chh, n = (t2 / p.hchunks) * p.cn;
        mbar_wait(&emptyB[bs], bph ^ 1);
        uint8_t* sb = smem_b + (size_t)bs * B_SLOT;
        mbar_expect_tx(&fullB[bs], split3 ? B_SLOT : B_BYTES);
        tma_load_5d(sb, &map_xh, &fullB[bs], 0, w0 + xdw, h + xdh, n + xdn, bslab);
        if (split3) tma_load_5d(sb + B_BYTES, &map_xl, &fullB[bs], 0, w0 + xdw, h + xdh, n + xdn, bslab);
        if (++bs == BS) { bs = 0; bph ^= 1; }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&emptyA[as], aph ^ 1);
          uint8_t* sa = smem_a + (size_t)as * A_SLOT;
          mbar_expect_tx(&fullA[as], split3 ? A_SLOT : A_BYTES);
          const int aslab = (copair * 2 + hf) * 4;      // past the last slab: out of bounds, zero filled
          tma_load_5d(sa, &map_gh, &fullA[as], 0, w0, h, n + gdn, aslab);
          if (split3) tma_load_5d(sa + A_BYTES, &map_gl, &fullA[as], 0, w0, h, n + gdn, aslab);
          if (++as == AS) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(TC_BM, BNW);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int it = 0; it < nk; ++it) {
        mbar_wait(&fullB[bs], bph);
        const uint32_t sb = smem_u32(smem_b + (size_t)bs * B_SLOT);
        const uint64_t bh = make_desc_mn(sb), bl = make_desc_mn(sb + B_BYTES);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&fullA[as], aph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + (size_t)as * A_SLOT);
          const uint64_t ah = make_desc_mn(sa), al = make_desc_mn(sa + A_BYTES);
          const uint32_t d = tmem_base + hf * ACC_STRIDE;
#pragma unroll
          for (int k = 0; k < WG_PIX / 8; ++k) {
            const uint64_t koff = (uint64_t)((k * 8 * 128) >> 4);
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            if (split3) {
              umma_tf32(d, al + koff, bh + koff, idesc, acc);
              umma_tf32(d, ah + koff, bl + koff, idesc, 1u);
              umma_tf32(d, ah + koff, bh + koff, idesc, 1u);
            } else {
              umma_tf32(d, ah + koff, bh + koff, idesc, acc);
            }
          }
          umma_commit(&emptyA[as]);
          if (++as == AS) { as = 0; aph ^= 1; }
        }
        umma_commit(&emptyB[bs]);
        if (++bs == BS) { bs = 0; bph ^= 1; }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int cleft = p.C - citile * BNW;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      const int co = (copair * 2 + hf) * TC_BM + q * 32 + lane;
      const bool row_ok = co < p.Kp;
      float* drow = p.dw + ((long long)p.out[tap] * p.Kp + co) * p.C + citile * BNW;
#pragma unroll 1
      for (int c = 0; c < BNW; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * ACC_STRIDE + c), r);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c + j < BNW && c + j < cleft) atomicAdd(drow + c + j, __uint_as_float(r[j]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BNW>
int launch_wgrad2(const CUtensorMap& gh, const CUtensorMap& gl, const CUtensorMap& xh, const CUtensorMap& xl,
                  const TcWgradParams& p, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)3 * (2 * 4 * WG_SLAB) + (size_t)2 * (2 * ((BNW + 31) / 32) * WG_SLAB) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad2_kernel<BNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  conv_tc_wgrad2_kernel<BNW><<<grid, TC_THREADS, smem, stream>>>(gh, gl, xh, xl, p);
  return (int)cudaGetLastError();
}

template <int BNW, int STAGES>
int launch_wgrad(const CUtensorMap& gh, const CUtensorMap& gl, const CUtensorMap& xh, const CUtensorMap& xl,
                 const TcWgradParams& p, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * 4 * WG_SLAB + 2 * ((BNW + 31) / 32) * WG_SLAB) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<BNW, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  conv_tc_wgrad_kernel<BNW, STAGES><<<grid, TC_THREADS, smem, stream>>>(gh, gl, xh, xl, p);
  return (int)cudaGetLastError();
}

template <int BN, bool TALL>
int launch_tc2(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
               const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)3 * (2 * (TALL ? 160 : TC_BM) * TC_BK * 4) + (size_t)2 * (2 * BN * TC_BK * 4) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, TALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  conv_tc2_kernel<BN, TALL><<<grid, TC_THREADS, smem, stream>>>(ah, al, bh, bl, p);
  return (int)cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI.
//   xh/xl : source activations, tf32 hi / lo parts (og_prep_split), NHWC [SN][SH][SW][C] contiguous, C % 4 == 0
//           (SN = N, or 4*N for a space-to-depth source: phase block (a*2+b) holds x[:, a::2, b::2])
//   wh/wl : weights hi / lo, [ntaps_w][Kw][C] (og_pack_weights with transposed=1 for fprop), Kw rows
//   y     : output NHWC, channel count K (K % 4 == 0), element strides ysn/ysh/ysw, spatial bounds OHf x OWf
//   rows of the GEMM are the pixel grid N x OH x OW; output pixel (osy*h + opy, osx*w + opx)
//   taps  : ntaps quadruples (dh, dw, dn, weight tap index): source pixel = image n + dn, (h + dh, w + dw), zero
//           outside [0,SH)x[0,SW);  bias/act: optional epilogue (bias[K], OG_ACT_*)
//   tap_layout = 1 promises that the taps come in groups of three with equal dw, dn and dh = d0, d0+1, d0+2 (3x3
//           kernels listed column by column), which lets the kernel reuse one tall pixel patch for three taps
// ------------------------------------------------------------------------------------------------
OG_API int og_conv2d_tc(const float* xh, const float* xl, int N, int SN, int SH, int SW, int C, const float* wh,
                        const float* wl, int ntaps_w, int Kw, float* y, int OH, int OW, int K, long long ysn,
                        long long ysh, long long ysw, int OHf, int OWf, int osy, int osx, int opy, int opx,
                        const int* taps_host, int ntaps, int tap_layout, int nsplit, const float* bias, int act,
                        float slope, cudaStream_t stream) {
  if (C % 4 || K % 4 || ntaps < 1 || ntaps > TC_MAX_TAPS || (nsplit != 1 && nsplit != 3)) return (int)cudaErrorInvalidValue;
  if ((long long)N * OH * OW == 0) return 0;
  TcParams p;
  p.N = N; p.OH = OH; p.OW = OW;
  // patch shape: TW = largest power of two <= min(OW, 16); TH fills up to 128 / TW within OH; TN the rest
  int TW = 1;
  while (TW * 2 <= OW && TW * 2 <= 16) TW *= 2;
  int TH = 1;
  while (TH * 2 <= OH && TW * TH * 2 <= TC_BM) TH *= 2;
  int TN = TC_BM / (TW * TH);
  // a tile must not straddle two space-to-depth phase blocks of the source (image index n + dn)
  if (SN != N && (N % TN) != 0) return (int)cudaErrorInvalidValue;
  p.TN = TN; p.TH = TH; p.TW = TW;
  p.bias = bias; p.act = act; p.slope = slope;
  p.tall = 0;
  p.tiles_w = og_cdiv(OW, TW);
  p.tiles_h = og_cdiv(OH, TH);
  const int tiles_n = og_cdiv(N, TN);
  p.cchunks = og_cdiv(C, TC_BK);
  p.K = K;
  p.nsplit = nsplit;
  p.y = y; p.ysn = ysn; p.ysh = ysh; p.ysw = ysw;
  p.osy = osy; p.osx = osx; p.opy = opy; p.opx = opx;
  p.OHfull = OHf; p.OWfull = OWf;
  p.taps.n = ntaps;
  for (int i = 0; i < ntaps; ++i) {
    p.taps.dh[i] = taps_host[4 * i];
    p.taps.dw[i] = taps_host[4 * i + 1];
    p.taps.dn[i] = taps_host[4 * i + 2];
    p.taps.widx[i] = taps_host[4 * i + 3];
  }
  // N tile: whole K if it fits one UMMA (<= 256), else the smallest number of equal tiles (multiple of 16)
  const int Kr = (K + 15) / 16 * 16;
  const int ntile = og_cdiv(Kr, 256);
  int BN = (og_cdiv(Kr, ntile) + 15) / 16 * 16;
  int BNsel = BN <= 64 ? 64 : BN <= 112 ? 112 : BN <= 208 ? 208 : 256;
  if (BN <= 32) BNsel = 32;
  dim3 grid(tiles_n * p.tiles_h * p.tiles_w, og_cdiv(K, BNsel), 1);
  // few output tiles (small feature maps): split the reduction over gridDim.z and combine with atomics
  p.ksplit = 1;
  {
    const long long tiles = (long long)grid.x * grid.y;
    const int nk_total = ntaps * p.cchunks;
    const bool dense = ysw == K && ysh == (long long)OWf * K && ysn == (long long)OHf * OWf * K && osy == 1 && osx == 1 &&
                       OH == OHf && OW == OWf;
    if (tiles < 100 && nk_total >= 16 && !bias && act == OG_ACT_NONE && dense) {
      int s = (int)((256 + tiles - 1) / tiles);
      if (s > nk_total / 8) s = nk_total / 8;
      if (s > 1) {
        p.ksplit = s;
        grid.z = s;
        OG_CHECK(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)N * OH * OW * K, stream));
      }
    }
  }

  CUtensorMap mah, mal, mbh, mbl;
  unsigned long long adims[4] = {(unsigned long long)C, (unsigned long long)SW, (unsigned long long)SH, (unsigned long long)SN};
  unsigned long long astr[3] = {(unsigned long long)C, (unsigned long long)SW * C, (unsigned long long)SH * SW * C};
  unsigned abox[4] = {(unsigned)TC_BK, (unsigned)TW, (unsigned)TH, (unsigned)TN};
  unsigned long long bdims[3] = {(unsigned long long)C, (unsigned long long)Kw, (unsigned long long)ntaps_w};
  unsigned long long bstr[2] = {(unsigned long long)C, (unsigned long long)Kw * C};
  unsigned bbox[3] = {(unsigned)TC_BK, (unsigned)BNsel, 1u};
  int rc;
  if ((rc = make_map(&mah, xh, 4, adims, astr, abox))) return rc;
  if ((rc = make_map(&mbh, wh, 3, bdims, bstr, bbox))) return rc;
  if (nsplit == 3) {
    if ((rc = make_map(&mal, xl, 4, adims, astr, abox))) return rc;
    if ((rc = make_map(&mbl, wl, 3, bdims, bstr, bbox))) return rc;
  } else {
    mal = mah;
    mbl = mbh;
  }
  // large problems: two pixel tiles per CTA share every weight stage (see conv_tc2_kernel)
  static const bool no_tc2 = getenv("OG_NO_TC2") != nullptr;
  static const long long tc2_min = getenv("OG_TC2_MIN") ? atoll(getenv("OG_TC2_MIN")) : 2 * 148 * 2;
  if (!no_tc2 && p.ksplit == 1 && (BNsel == 208 || BNsel == 256) && (long long)grid.x * grid.y >= tc2_min) {
    dim3 grid2((grid.x + 1) / 2, grid.y, 1);
    static const bool no_tall = getenv("OG_NO_TALL") != nullptr;
    if (!no_tall && tap_layout == 1 && ntaps % 3 == 0 && BNsel == 208 && TN == 1 && TH == 8 && TW == 16) {
      // taps are ordered (column-major) in groups of three consecutive source rows: one tall A patch per group
      unsigned tbox[4] = {(unsigned)TC_BK, 16u, 10u, 1u};
      CUtensorMap tah, tal;
      if ((rc = make_map(&tah, xh, 4, adims, astr, tbox))) return rc;
      if (nsplit == 3) {
        if ((rc = make_map(&tal, xl, 4, adims, astr, tbox))) return rc;
      } else {
        tal = tah;
      }
      p.tall = 1;
      return launch_tc2<208, true>(tah, tal, mbh, mbl, p, grid2, stream);
    }
    return BNsel == 208 ? launch_tc2<208, false>(mah, mal, mbh, mbl, p, grid2, stream)
                        : launch_tc2<256, false>(mah, mal, mbh, mbl, p, grid2, stream);
  }
  switch (BNsel) {
    case 32:  return launch_tc<32, 4>(mah, mal, mbh, mbl, p, grid, stream);
    case 64:  return launch_tc<64, 4>(mah, mal, mbh, mbl, p, grid, stream);
    case 112: return launch_tc<112, 3>(mah, mal, mbh, mbl, p, grid, stream);
    case 208: return launch_tc<208, 2>(mah, mal, mbh, mbl, p, grid, stream);
    default:  return launch_tc<256, 2>(mah, mal, mbh, mbl, p, grid, stream);
  }
}


// ------------------------------------------------------------------------------------------------
// weight gradient:  dw[t][co][ci] (fp32, [ntaps][Kp][C], zero-filled here) = sum_pixels G^T * X(shifted by tap t)
//   gh/gl : output gradient hi/lo, NHWC [GN][OH][OW][Kp]   (og_prep_split; GN = N, or 4N space-to-depth blocks)
//   xh/xl : source activations hi/lo, NHWC [XN][SH][SW][C] (og_prep_split: plain, reflection-padded or
//           space-to-depth); both need 512 readable bytes after the last element (partial last channel slab)
//   entries: nentries quintuples (g image offset, dh, dw, x image offset, output tap):
//           dw[tap] += sum_{n,h,w} G[n + gdn, h, w, :]^T  X[n + xdn, h + dh, w + dw, :]   (out of range = 0)
// ------------------------------------------------------------------------------------------------
OG_API int og_conv2d_wgrad_tc(const float* gh, const float* gl, int N, int GN, int OH, int OW, int Kp,
                              const float* xh, const float* xl, int XN, int SH, int SW, int C, float* dw,
                              int ntaps_out, const int* entries_host, int nentries, int nsplit, cudaStream_t stream) {
  if (nentries < 1 || nentries > TC_MAX_TAPS || (nsplit != 1 && nsplit != 3)) return (int)cudaErrorInvalidValue;
  OG_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)ntaps_out * Kp * C, stream));
  if ((long long)N * OH * OW == 0) return 0;
  TcWgradParams p;
  p.N = N; p.OH = OH; p.OW = OW;
  // one stage = 32 pixels: a cw x chh x cn patch of the gradient grid (w fastest)
  int cw = 1;
  while (cw * 2 <= WG_PIX && OW % (cw * 2) == 0) cw *= 2;
  int chh = 1;
  while (cw * chh * 2 <= WG_PIX && OH % (chh * 2) == 0) chh *= 2;
  const int cn = WG_PIX / (cw * chh);
  if (N % cn != 0) return (int)cudaErrorInvalidValue;
  p.cw = cw; p.chh = chh; p.cn = cn;
  p.wchunks = OW / cw;
  p.hchunks = OH / chh;
  p.total_chunks = p.wchunks * p.hchunks * (N / cn);
  p.Kp = Kp; p.C = C; p.nsplit = nsplit; p.dw = dw;
  for (int i = 0; i < nentries; ++i) {
    p.gdn[i] = entries_host[5 * i];
    p.xdh[i] = entries_host[5 * i + 1];
    p.xdw[i] = entries_host[5 * i + 2];
    p.xdn[i] = entries_host[5 * i + 3];
    p.out[i] = entries_host[5 * i + 4];
  }
  const int cotiles = og_cdiv(Kp, TC_BM);
  const int Cr = (C + 15) / 16 * 16;
  const int BNsel = Cr <= 32 ? 32 : Cr <= 64 ? 64 : Cr <= 112 ? 112 : Cr <= 208 ? 208 : 256;
  const int citiles = og_cdiv(C, BNsel);
  p.cotiles = cotiles;
  // pixel splits: ~2 waves of CTAs over 148 SMs, at least 8 chunks per CTA
  int splits = og_cdiv(296, cotiles * citiles * nentries);
  int maxs = p.total_chunks / 8;
  if (maxs < 1) maxs = 1;
  if (splits > maxs) splits = maxs;
  p.chunks_per_cta = og_cdiv(p.total_chunks, splits);
  splits = og_cdiv(p.total_chunks, p.chunks_per_cta);
  dim3 grid(cotiles * citiles, nentries, splits);
  CUtensorMap mgh, mgl, mxh, mxl;
  // [slab][n][h][w][32 channels of the slab]: the slab dimension strides by 32 floats inside a pixel's channel vector
  unsigned long long gd[5] = {32ull, (unsigned long long)OW, (unsigned long long)OH, (unsigned long long)GN,
                              (unsigned long long)og_cdiv(Kp, 32)};
  unsigned long long gs[4] = {(unsigned long long)Kp, (unsigned long long)OW * Kp, (unsigned long long)OH * OW * Kp, 32ull};
  unsigned gb[5] = {32u, (unsigned)cw, (unsigned)chh, (unsigned)cn, 4u};
  unsigned long long xd[5] = {32ull, (unsigned long long)SW, (unsigned long long)SH, (unsigned long long)XN,
                              (unsigned long long)og_cdiv(C, 32)};
  unsigned long long xs[4] = {(unsigned long long)C, (unsigned long long)SW * C, (unsigned long long)SH * SW * C, 32ull};
  unsigned xb[5] = {32u, (unsigned)cw, (unsigned)chh, (unsigned)cn, (unsigned)((BNsel + 31) / 32)};
  int rc;
  if ((rc = make_map(&mgh, gh, 5, gd, gs, gb, true))) return rc;
  if ((rc = make_map(&mxh, xh, 5, xd, xs, xb, true))) return rc;
  if (nsplit == 3) {
    if ((rc = make_map(&mgl, gl, 5, gd, gs, gb, true))) return rc;
    if ((rc = make_map(&mxl, xl, 5, xd, xs, xb, true))) return rc;
  } else {
    mgl = mgh;
    mxl = mxh;
  }
  static const bool no_tc2 = getenv("OG_NO_TC2") != nullptr;
  if (!no_tc2 && cotiles >= 2 && (BNsel == 208 || BNsel == 256)) {
    // two co-tiles per CTA share every X stage (L2 -> SM traffic is the bound of this kernel)
    const int copairs = (cotiles + 1) / 2;
    int sp = og_cdiv(296, copairs * citiles * nentries);
    if (sp > maxs) sp = maxs;
    if (sp < 1) sp = 1;
    p.chunks_per_cta = og_cdiv(p.total_chunks, sp);
    sp = og_cdiv(p.total_chunks, p.chunks_per_cta);
    dim3 grid2(copairs * citiles, nentries, sp);
    return BNsel == 208 ? launch_wgrad2<208>(mgh, mgl, mxh, mxl, p, grid2, stream)
                        : launch_wgrad2<256>(mgh, mgl, mxh, mxl, p, grid2, stream);
  }
  switch (BNsel) {
    case 32:  return launch_wgrad<32, 4>(mgh, mgl, mxh, mxl, p, grid, stream);
    case 64:  return launch_wgrad<64, 4>(mgh, mgl, mxh, mxl, p, grid, stream);
    case 112: return launch_wgrad<112, 3>(mgh, mgl, mxh, mxl, p, grid, stream);
    case 208: return launch_wgrad<208, 2>(mgh, mgl, mxh, mxl, p, grid, stream);
    default:  return launch_wgrad<256, 2>(mgh, mgl, mxh, mxl, p, grid, stream);
  }
}

__global__ void prep_split_kernel(const float* __restrict__ x, int N, int H, int W, int C4, int pad, int s2d,
                                  long long total, float* __restrict__ xh, float* __restrict__ xl) {
  // output: [N][H+2p][W+2p][C] or, for s2d, [4][N][H/2][W/2][C] with phase block (a*2+b) = x[:, a::2, b::2]
  const int Hp = s2d ? H / 2 : H + 2 * pad, Wp = s2d ? W / 2 : W + 2 * pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    long long t = i / C4;
    int w = (int)(t % Wp);
    t /= Wp;
    int h = (int)(t % Hp);
    long long n = t / Hp;
    int sh, sw;
    if (s2d) {
      int ph = (int)(n / N);
      n -= (long long)ph * N;
      sh = 2 * h + (ph >> 1);
      sw = 2 * w + (ph & 1);
    } else {
      sh = h - pad;
      sw = w - pad;
      if (sh < 0) sh = -sh;
      if (sh >= H) sh = 2 * H - 2 - sh;
      if (sw < 0) sw = -sw;
      if (sw >= W) sw = 2 * W - 2 - sw;
    }
    float4 v = ldg4(x + (((n * H + sh) * W + sw) * C4 + c) * 4);
    float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    st4(xh + i * 4, hi);
    if (xl) st4(xl + i * 4, make_float4(tf32_lo(v.x, hi.x), tf32_lo(v.y, hi.y), tf32_lo(v.z, hi.z), tf32_lo(v.w, hi.w)));
  }
}

// pad = 0: plain split; pad = 1: nn.ReflectionPad2d(1) halo (model.py:67) materialised while splitting
OG_API int og_prep_split(const float* x, int N, int H, int W, int C, int pad, int s2d, float* xh, float* xl,
                         cudaStream_t stream) {
  if (C % 4 || pad < 0 || pad > 1 || (s2d && (pad || (H & 1) || (W & 1)))) return (int)cudaErrorInvalidValue;
  long long total = s2d ? (long long)N * H * W * (C / 4) : (long long)N * (H + 2 * pad) * (W + 2 * pad) * (C / 4);
  if (total == 0) return 0;
  long long b = (total + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  prep_split_kernel<<<(int)b, 256, 0, stream>>>(x, N, H, W, C / 4, pad, s2d, total, xh, xl);
  OG_RETURN_LAST_ERROR();
}
