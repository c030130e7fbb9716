// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::f16, fp32 accumulate) with TMEM
// accumulators, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier producer/consumer pipeline.
//
// Covers every stride-1 contraction of the hot path through a "tap list":
//     out[pix, :] = sum_t  src[pix + (dh_t, dw_t), :] * W[widx_t]          (src out-of-bounds = 0 via TMA fill)
// which is the forward of the reflection-padded 3x3 convs of HmapResBlock (model.py:63-81; the halo is
// materialised by og_prep_split), the zero-padded 3x3 convs (model.py:36-39, 1020-1048), their input
// gradients (taps mirrored, operand re-packed) and the four phases of nearest-2x-upsample + conv3x3
// (upBlock, model.py:43-49: tap offsets at low resolution, strided output pixels).
//
// Precision: fp32 parity (1e-3 end to end) is not reachable with one 11-bit product (2^-11 operand rounding,
// ~35 stacked convs), so by default each product is error-compensated:  a*b ~= ah*bh + al*bh + ah*bl  with
// ah = fp16(a), al = fp16(a - ah): 22 mantissa bits per operand, three MMAs into the same fp32 TMEM accumulator
// ("3xFP16").  fp16 has the mantissa of tf32 at twice the MMA rate and half the operand bytes; its narrow exponent
// range is handled by scaling every operand tensor by a power of two taken from its max|x| (og_prep_split /
// og_pack_weights_f16 write x * 2^k; the epilogue multiplies the accumulator by 2^-(ka+kb)): elements within
// 2^-18 of the tensor maximum keep all 22 bits, smaller ones an absolute error <= 2^-39 of the maximum.
// nsplit = 1 runs a single fp16 product (reported separately, never the parity mode).
//
// GEMM tiling: one CTA = 128 output pixels (a TN x TH x TW patch) x BN output channels (BN % 16 == 0,
// <= 256); K loop over taps x 64-channel chunks; STAGES-deep smem ring of {A_hi, A_lo, B_hi, B_lo}.
// Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> global, one TMEM lane quadrant each).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;              // fp16 elements per k-chunk = one 128-byte swizzle row (4 MMA k-steps of 16)
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
  int cchunks;            // ceil(C / 64)
  int klast;              // MMA k-steps (16 channels each) that the last chunk really needs: 1..4
  const unsigned* amax_a; // float bits of max|x| of the two operand tensors (power-of-two operand scaling)
  const unsigned* amax_b;
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
  int stacked;              // CTA-pair kernel, <= 16 output channels: B = [w_hi (CTA 0) | w_lo (CTA 1)] along N, two MMAs per k-step
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
// bounded wait (CTA-pair kernel): a broken cross-CTA barrier protocol traps after ~2 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t done = 0;
  int spins = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if ((++spins & 255) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
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

// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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

// ---- CTA-pair (cta_group::2) forms: two CTAs of a (2,1,1) cluster issue ONE M = 256 MMA; PTX forms as in the
// vendored CUTLASS headers (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_*, mma_sm100_umma.hpp
// SM100_MMA_F16BF16_2x1SM_SS, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM, tmem_allocator_sm100.hpp) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: "the leader's"
// both CTAs of the pair load into their OWN shared memory; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs: 128 rows each] (+)= A[each CTA's own smem] * B[N/2 rows from each CTA's smem]; leader only
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ float tc_act(float v, int act, float slope) {
  if (act == OG_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == OG_ACT_TANH) return tanhf(v);
  if (act == OG_ACT_SIGMOID) return og_sigmoid(v);
  return v;
}

// accumulator -> fp32 result: undo the two power-of-two operand scales (two factors: their product may leave the
// float exponent range although each factor and the result do not)
struct TcScale { float fa, fb; };
__device__ __forceinline__ TcScale tc_scale(const unsigned* amax_a, const unsigned* amax_b) {
  TcScale s;
  s.fa = og_exp2i(-og_scale_exp(__ldg(amax_a)));
  s.fb = og_exp2i(-og_scale_exp(__ldg(amax_b)));
  return s;
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
// instruction descriptor: D = f32, A = B = f16 (format code 0), both K-major, dense
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
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
  constexpr uint32_t A_BYTES = TC_BM * 128;         // 128 rows of 128 bytes = 16 KB
  constexpr uint32_t B_BYTES = BN * 128;
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
      const uint32_t idesc = make_idesc_f16(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nk; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t ah = make_desc_sw128(sa), al = make_desc_sw128(sa + A_BYTES);
        const uint64_t bh = make_desc_sw128(sa + 2 * A_BYTES), bl = make_desc_sw128(sa + 2 * A_BYTES + B_BYTES);
        // the last channel chunk may hold fewer than 64 real channels: skip the k-steps that would multiply zeros
        const int ks = ((it_beg + it + 1) % p.cchunks == 0) ? p.klast : 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= ks) break;
          const uint64_t koff = (uint64_t)((k * 32) >> 4);      // advance 32 bytes (16 fp16) inside the swizzle row
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          if (p.nsplit == 3) {
            umma_f16(tmem_base, al + koff, bh + koff, idesc, acc);   // small terms first
            umma_f16(tmem_base, ah + koff, bl + koff, idesc, 1u);
            umma_f16(tmem_base, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma_f16(tmem_base, ah + koff, bh + koff, idesc, acc);
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
    const TcScale sc = tc_scale(p.amax_a, p.amax_b);
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
            float4 v = make_float4(__uint_as_float(r[j]) * sc.fa * sc.fb, __uint_as_float(r[j + 1]) * sc.fa * sc.fb,
                                   __uint_as_float(r[j + 2]) * sc.fa * sc.fb, __uint_as_float(r[j + 3]) * sc.fa * sc.fb);
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
  constexpr uint32_t A_BYTES = (TALL ? 160 : TC_BM) * 128;   // one of hi / lo: 128 rows, or 10 x 16 patch rows
  constexpr uint32_t B_BYTES = BN * 128;
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
      const uint32_t idesc = make_idesc_f16(TC_BM, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto mma_half = [&](uint32_t sa, uint32_t rowoff, uint32_t sb, int hf, bool first, int ks) {
        const uint64_t ah = make_desc_sw128(sa + rowoff), al = make_desc_sw128(sa + A_BYTES + rowoff);
        const uint64_t bh = make_desc_sw128(sb), bl = make_desc_sw128(sb + B_BYTES);
        const uint32_t d = tmem_base + hf * ACC_STRIDE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= ks) break;
          const uint64_t koff = (uint64_t)((k * 32) >> 4);
          const uint32_t acc = (!first || k > 0) ? 1u : 0u;
          if (split3) {
            umma_f16(d, al + koff, bh + koff, idesc, acc);
            umma_f16(d, ah + koff, bl + koff, idesc, 1u);
            umma_f16(d, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma_f16(d, ah + koff, bh + koff, idesc, acc);
          }
        }
      };
      if (TALL) {
        const int groups = (p.taps.n / 3) * p.cchunks;
        for (int g = 0; g < groups; ++g) {
          const int a0 = as, a1 = (as + 1) % AS;
          const uint32_t aph0 = aph, aph1 = (as + 1 == AS) ? (aph ^ 1) : aph;
          const int ks = ((g + 1) % p.cchunks == 0) ? p.klast : 4;
          for (int j = 0; j < 3; ++j) {
            mbar_wait(&fullB[bs], bph);
            const uint32_t sb = smem_u32(smem_b + (size_t)bs * B_SLOT);
            if (j == 0) mbar_wait(&fullA[a0], aph0);
            tc_fence_after();
            mma_half(smem_u32(smem_a + (size_t)a0 * A_SLOT), j * 16 * 128, sb, 0, g == 0 && j == 0, ks);
            if (j == 0) {
              mbar_wait(&fullA[a1], aph1);
              tc_fence_after();
            }
            mma_half(smem_u32(smem_a + (size_t)a1 * A_SLOT), j * 16 * 128, sb, 1, g == 0 && j == 0, ks);
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
          const int ks = ((it + 1) % p.cchunks == 0) ? p.klast : 4;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&fullA[as], aph);
            tc_fence_after();
            mma_half(smem_u32(smem_a + (size_t)as * A_SLOT), 0, sb, hf, it == 0, ks);
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
    const TcScale sc = tc_scale(p.amax_a, p.amax_b);
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
              float4 v = make_float4(__uint_as_float(r[j]) * sc.fa * sc.fb, __uint_as_float(r[j + 1]) * sc.fa * sc.fb,
                                     __uint_as_float(r[j + 2]) * sc.fa * sc.fb, __uint_as_float(r[j + 3]) * sc.fa * sc.fb);
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
// CTA-pair variant of the two-tile kernel (cta_group::2): the two CTAs of a (2,1,1) cluster each own TWO 128-pixel
// tiles (4 tiles per pair) and HALF of every weight stage.  One tcgen05.mma.cta_group::2 (M = 256 = 128 pixels of
// each CTA, N = BN) is issued by the leader per (tile half, product term, k-step); the tensor cores of both SMs read
// their own A rows and BN/2 weight rows from each CTA, so per SM the shared-memory read per MMA drops from
// 16*2*(128 + BN) bytes to 16*2*(128 + BN/2) and a weight stage from 2*BN*128 to BN*128 bytes (4 stages fit).
//   barriers: fullA / fullB live in the LEADER (both CTAs' TMA loads credit their bytes there: PEER_BIT_MASK);
//             emptyA / emptyB / tmem_full exist in both CTAs, released by the leader's multicast tcgen05.commit.
//   TMEM: allocated with cta_group::2 by warp 1 of each CTA; each CTA's epilogue reads its own 128 lanes.
// ------------------------------------------------------------------------------------------------
template <int BN, bool TALL>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2x_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                 const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                 const TcParams p) {
  constexpr int AS = 3, BS = 4;
  constexpr int BNH = BN / 2;                                 // weight rows held by each CTA
  constexpr uint32_t A_BYTES = (TALL ? 160 : TC_BM) * 128;
  constexpr uint32_t B_BYTES = BNH * 128;
  constexpr uint32_t A_SLOT = 2 * A_BYTES, B_SLOT = 2 * B_BYTES;
  constexpr uint32_t ACC_STRIDE = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + AS * A_SLOT;
  __shared__ __align__(8) uint64_t fullA[AS], emptyA[AS], fullB[BS], emptyB[BS], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int col0 = blockIdx.y * BN;
  const int nk = p.taps.n * p.cchunks;
  const bool split3 = p.nsplit == 3;
  int tn0[2], th0[2], tw0[2];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    int t = blockIdx.x * 2 + hf;      // blockIdx.x = 2 * pair + rank: tiles 4*pair + 2*rank + hf
    const int tw_i = t % p.tiles_w;
    t /= p.tiles_w;
    const int th_i = t % p.tiles_h;
    const int tn_i = t / p.tiles_h;   // past the last image for a padded tile count: TMA zero-fills, rows masked
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
  if (warp == 1) tmem_alloc2(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the peer's barriers are initialised before anything may arrive on them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto load_b = [&](int tap, int c0) {
        mbar_wait_bounded(&emptyB[bs], bph ^ 1);
        uint8_t* sb = smem_b + (size_t)bs * B_SLOT;
        if (leader) mbar_expect_tx(&fullB[bs], 2 * (split3 ? B_SLOT : B_BYTES));     // both CTAs' halves
        tma2_load_3d(sb, &map_bh, &fullB[bs], c0, col0 + (int)rank * BNH, p.taps.widx[tap]);
        if (split3) tma2_load_3d(sb + B_BYTES, &map_bl, &fullB[bs], c0, col0 + (int)rank * BNH, p.taps.widx[tap]);
        if (++bs == BS) { bs = 0; bph ^= 1; }
      };
      auto load_a = [&](int hf, int tap, int c0) {
        mbar_wait_bounded(&emptyA[as], aph ^ 1);
        uint8_t* sa = smem_a + (size_t)as * A_SLOT;
        if (leader) mbar_expect_tx(&fullA[as], 2 * (split3 ? A_SLOT : A_BYTES));
        const int ws = tw0[hf] + p.taps.dw[tap], hs = th0[hf] + p.taps.dh[tap], ns = tn0[hf] + p.taps.dn[tap];
        tma2_load_4d(sa, &map_ah, &fullA[as], c0, ws, hs, ns);
        if (split3) tma2_load_4d(sa + A_BYTES, &map_al, &fullA[as], c0, ws, hs, ns);
        if (++as == AS) { as = 0; aph ^= 1; }
      };
      if (TALL) {
        const int groups = (p.taps.n / 3) * p.cchunks;
        for (int g = 0; g < groups; ++g) {
          const int kwi = g / p.cchunks, cc = g - kwi * p.cchunks;
          const int c0 = cc * TC_BK, t0 = kwi * 3;
          load_a(0, t0, c0);
          load_b(t0, c0);
          load_a(1, t0, c0);
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
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_f16(2 * TC_BM, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto mma_half = [&](uint32_t sa, uint32_t rowoff, uint32_t sb, int hf, bool first, int ks) {
        const uint64_t ah = make_desc_sw128(sa + rowoff), al = make_desc_sw128(sa + A_BYTES + rowoff);
        const uint64_t bh = make_desc_sw128(sb), bl = make_desc_sw128(sb + B_BYTES);
        const uint32_t d = tmem_base + hf * ACC_STRIDE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= ks) break;
          const uint64_t koff = (uint64_t)((k * 32) >> 4);
          const uint32_t acc = (!first || k > 0) ? 1u : 0u;
          if (split3) {
            umma2_f16(d, al + koff, bh + koff, idesc, acc);
            umma2_f16(d, ah + koff, bl + koff, idesc, 1u);
            umma2_f16(d, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma2_f16(d, ah + koff, bh + koff, idesc, acc);
          }
        }
      };
      if (TALL) {
        const int groups = (p.taps.n / 3) * p.cchunks;
        for (int g = 0; g < groups; ++g) {
          const int a0 = as, a1 = (as + 1) % AS;
          const uint32_t aph0 = aph, aph1 = (as + 1 == AS) ? (aph ^ 1) : aph;
          const int ks = ((g + 1) % p.cchunks == 0) ? p.klast : 4;
          for (int j = 0; j < 3; ++j) {
            mbar_wait_bounded(&fullB[bs], bph);
            const uint32_t sb = smem_u32(smem_b + (size_t)bs * B_SLOT);
            if (j == 0) mbar_wait_bounded(&fullA[a0], aph0);
            tc_fence_after();
            mma_half(smem_u32(smem_a + (size_t)a0 * A_SLOT), j * 16 * 128, sb, 0, g == 0 && j == 0, ks);
            if (j == 0) {
              mbar_wait_bounded(&fullA[a1], aph1);
              tc_fence_after();
            }
            mma_half(smem_u32(smem_a + (size_t)a1 * A_SLOT), j * 16 * 128, sb, 1, g == 0 && j == 0, ks);
            umma2_commit(&emptyB[bs]);
            if (++bs == BS) { bs = 0; bph ^= 1; }
          }
          umma2_commit(&emptyA[a0]);
          umma2_commit(&emptyA[a1]);
          for (int k = 0; k < 2; ++k)
            if (++as == AS) { as = 0; aph ^= 1; }
        }
      } else {
        for (int it = 0; it < nk; ++it) {
          mbar_wait_bounded(&fullB[bs], bph);
          const uint32_t sb = smem_u32(smem_b + (size_t)bs * B_SLOT);
          const int ks = ((it + 1) % p.cchunks == 0) ? p.klast : 4;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait_bounded(&fullA[as], aph);
            tc_fence_after();
            mma_half(smem_u32(smem_a + (size_t)as * A_SLOT), 0, sb, hf, it == 0, ks);
            umma2_commit(&emptyA[as]);
            if (++as == AS) { as = 0; aph ^= 1; }
          }
          umma2_commit(&emptyB[bs]);
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
      }
      umma2_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const TcScale sc = tc_scale(p.amax_a, p.amax_b);
    mbar_wait_bounded(&tmem_full_bar, 0);
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
              float4 v = make_float4(__uint_as_float(r[j]) * sc.fa * sc.fb, __uint_as_float(r[j + 1]) * sc.fa * sc.fb,
                                     __uint_as_float(r[j + 2]) * sc.fa * sc.fb, __uint_as_float(r[j + 3]) * sc.fa * sc.fb);
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
  cluster_sync_all();                 // neither CTA frees TMEM / leaves while the pair's MMAs or arrivals are in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent CTA-pair kernel (cta_group::2 + static tile scheduler + double-buffered TMEM accumulators).
// grid = 2 x (number of SM pairs); cluster c processes work units c, c + nclusters, ...; a unit = (pair of adjacent
// 128-pixel tiles: one per CTA, N tile).  Per unit the leader issues M = 256 MMAs into accumulator stage (k & 1)
// (TMEM columns [0, BN) / [256, 256 + BN)); the epilogue warps of both CTAs drain stage k & 1 while the MMAs of unit
// k + 1 already run into the other stage, so TMEM allocation, barrier setup and the epilogue are off the critical
// path.  Weight stages are shared by the pair (each CTA loads BN/2 rows), which gives the same L2 -> SM weight
// traffic per pixel tile as two tiles per CTA, with one tile per CTA.
//   barriers: fullA/fullB (leader; both CTAs' TMA bytes), emptyA/emptyB/tmem_full[2] (both CTAs; multicast commit),
//             tmem_empty[2] (leader; 8 arrivals = 4 epilogue warps of each CTA).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

template <int BN, bool TALL>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tcp_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                const TcParams p, const int nunits, const int ntn) {
  constexpr int AS = 3, BS = 4;
  constexpr int BNH = BN / 2;
  constexpr uint32_t A_BYTES = (TALL ? 160 : TC_BM) * 128;
  constexpr uint32_t B_BYTES = BNH * 128;
  constexpr uint32_t A_SLOT = 2 * A_BYTES, B_SLOT = 2 * B_BYTES;
  constexpr uint32_t ACC_STRIDE = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + AS * A_SLOT;
  __shared__ __align__(8) uint64_t fullA[AS], emptyA[AS], fullB[BS], emptyB[BS], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nk = p.taps.n * p.cchunks;
  const bool split3 = p.nsplit == 3;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_ah);
    prefetch_tmap(&map_bh);
    for (int s = 0; s < AS; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // this CTA's tile of unit u
  auto tile_of = [&](int u, int& n0, int& h0, int& w0, int& col0) {
    const int pp = u / ntn;
    col0 = (u - pp * ntn) * BN;
    int t = pp * 2 + (int)rank;
    const int tw_i = t % p.tiles_w;
    t /= p.tiles_w;
    const int th_i = t % p.tiles_h;
    const int tn_i = t / p.tiles_h;   // past the last image for an odd tile count: TMA zero-fills, rows are masked
    n0 = tn_i * p.TN; h0 = th_i * p.TH; w0 = tw_i * p.TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int u = cluster; u < nunits; u += nclusters) {
        int n0, h0, w0, col0;
        tile_of(u, n0, h0, w0, col0);
        auto load_b = [&](int tap, int c0) {
          mbar_wait_bounded(&emptyB[bs], bph ^ 1);
          uint8_t* sb = smem_b + (size_t)bs * B_SLOT;
          if (p.stacked) {
            // one 4-D map over the {hi, lo} weight copies: CTA 0 takes the hi rows, CTA 1 the lo rows -- together the
            // N = 32 operand [w_hi | w_lo]
            if (leader) mbar_expect_tx(&fullB[bs], 2 * B_BYTES);
            tma2_load_4d(sb, &map_bh, &fullB[bs], c0, 0, (int)rank, p.taps.widx[tap]);
          } else {
            if (leader) mbar_expect_tx(&fullB[bs], 2 * (split3 ? B_SLOT : B_BYTES));
            tma2_load_3d(sb, &map_bh, &fullB[bs], c0, col0 + (int)rank * BNH, p.taps.widx[tap]);
            if (split3) tma2_load_3d(sb + B_BYTES, &map_bl, &fullB[bs], c0, col0 + (int)rank * BNH, p.taps.widx[tap]);
          }
          if (++bs == BS) { bs = 0; bph ^= 1; }
        };
        auto load_a = [&](int tap, int c0) {
          mbar_wait_bounded(&emptyA[as], aph ^ 1);
          uint8_t* sa = smem_a + (size_t)as * A_SLOT;
          if (leader) mbar_expect_tx(&fullA[as], 2 * (split3 ? A_SLOT : A_BYTES));
          const int ws = w0 + p.taps.dw[tap], hs = h0 + p.taps.dh[tap], ns = n0 + p.taps.dn[tap];
          tma2_load_4d(sa, &map_ah, &fullA[as], c0, ws, hs, ns);
          if (split3) tma2_load_4d(sa + A_BYTES, &map_al, &fullA[as], c0, ws, hs, ns);
          if (++as == AS) { as = 0; aph ^= 1; }
        };
        if (TALL) {
          const int groups = (p.taps.n / 3) * p.cchunks;
          for (int g = 0; g < groups; ++g) {
            const int kwi = g / p.cchunks, cc = g - kwi * p.cchunks;
            const int c0 = cc * TC_BK, t0 = kwi * 3;
            load_a(t0, c0);
            load_b(t0, c0);
            load_b(t0 + 1, c0);
            load_b(t0 + 2, c0);
          }
        } else {
          for (int it = 0; it < nk; ++it) {
            const int tap = it / p.cchunks, cc = it - tap * p.cchunks;
            load_b(tap, cc * TC_BK);
            load_a(tap, cc * TC_BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_f16(2 * TC_BM, BN);
      int as = 0, bs = 0, k = 0;
      uint32_t aph = 0, bph = 0;
      auto mma_step = [&](uint32_t sa, uint32_t rowoff, uint32_t sb, uint32_t d, bool first, int ks) {
        const uint64_t ah = make_desc_sw128(sa + rowoff), al = make_desc_sw128(sa + A_BYTES + rowoff);
        const uint64_t bh = make_desc_sw128(sb), bl = make_desc_sw128(sb + B_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk >= ks) break;
          const uint64_t koff = (uint64_t)((kk * 32) >> 4);
          const uint32_t acc = (!first || kk > 0) ? 1u : 0u;
          if (split3 && p.stacked) {
            // columns [0, 16): a_lo w_hi + a_hi w_hi; columns [16, 32): a_lo w_lo (negligible) + a_hi w_lo: each A copy
            // is read from shared memory ONCE (these narrow layers are bound by the A reads, not by the MMA rate)
            umma2_f16(d, al + koff, bh + koff, idesc, acc);
            umma2_f16(d, ah + koff, bh + koff, idesc, 1u);
          } else if (split3) {
            umma2_f16(d, al + koff, bh + koff, idesc, acc);
            umma2_f16(d, ah + koff, bl + koff, idesc, 1u);
            umma2_f16(d, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma2_f16(d, ah + koff, bh + koff, idesc, acc);
          }
        }
      };
      for (int u = cluster; u < nunits; u += nclusters, ++k) {
        const int st = k & 1;
        mbar_wait_bounded(&tmem_empty[st], ((k >> 1) & 1) ^ 1);       // both CTAs' epilogues drained this stage
        tc_fence_after();
        const uint32_t d = tmem_base + st * ACC_STRIDE;
        if (TALL) {
          const int groups = (p.taps.n / 3) * p.cchunks;
          for (int g = 0; g < groups; ++g) {
            const int ks = ((g + 1) % p.cchunks == 0) ? p.klast : 4;
            mbar_wait_bounded(&fullA[as], aph);
            const uint32_t sa = smem_u32(smem_a + (size_t)as * A_SLOT);
            for (int j = 0; j < 3; ++j) {
              mbar_wait_bounded(&fullB[bs], bph);
              tc_fence_after();
              mma_step(sa, j * 16 * 128, smem_u32(smem_b + (size_t)bs * B_SLOT), d, g == 0 && j == 0, ks);
              umma2_commit(&emptyB[bs]);
              if (++bs == BS) { bs = 0; bph ^= 1; }
            }
            umma2_commit(&emptyA[as]);
            if (++as == AS) { as = 0; aph ^= 1; }
          }
        } else {
          for (int it = 0; it < nk; ++it) {
            const int ks = ((it + 1) % p.cchunks == 0) ? p.klast : 4;
            mbar_wait_bounded(&fullB[bs], bph);
            mbar_wait_bounded(&fullA[as], aph);
            tc_fence_after();
            mma_step(smem_u32(smem_a + (size_t)as * A_SLOT), 0, smem_u32(smem_b + (size_t)bs * B_SLOT), d, it == 0, ks);
            umma2_commit(&emptyA[as]);
            umma2_commit(&emptyB[bs]);
            if (++as == AS) { as = 0; aph ^= 1; }
            if (++bs == BS) { bs = 0; bph ^= 1; }
          }
        }
        umma2_commit(&tmem_full[st]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const TcScale sc = tc_scale(p.amax_a, p.amax_b);
    int k = 0;
    for (int u = cluster; u < nunits; u += nclusters, ++k) {
      const int st = k & 1;
      int n0, h0, w0, col0;
      tile_of(u, n0, h0, w0, col0);
      const int n = n0 + tn, h = h0 + th, w = w0 + tw;
      const int oy = p.osy * h + p.opy, ox = p.osx * w + p.opx;
      const bool row_ok = (n < p.N) && (h < p.OH) && (w < p.OW) && (oy < p.OHfull) && (ox < p.OWfull);
      float* yrow = p.y + (long long)n * p.ysn + (long long)oy * p.ysh + (long long)ox * p.ysw + col0;
      mbar_wait_bounded(&tmem_full[st], (k >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * ACC_STRIDE + c), r);
        if (p.stacked) {      // columns [0, BN/2): products with w_hi, [BN/2, BN): with w_lo, of the same output channels
          constexpr int HALF = BN / 2;
          if (c >= HALF) break;
          if (HALF >= 32) {
            uint32_t r2[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * ACC_STRIDE + c + HALF), r2);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r[j + 16]));
#pragma unroll
            for (int j = 16; j < 32; ++j) r[j] = 0u;
          }
        }
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c + j < BN && col0 + c + j < p.K) {
              float4 v = make_float4(__uint_as_float(r[j]) * sc.fa * sc.fb, __uint_as_float(r[j + 1]) * sc.fa * sc.fb,
                                     __uint_as_float(r[j + 2]) * sc.fa * sc.fb, __uint_as_float(r[j + 3]) * sc.fa * sc.fb);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[st]);       // this warp's quadrant of stage st is drained
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
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

// fp16 tensor viewed as [d4][d3][d2][d1][d0] (d0 contiguous); strides in elements for d1..
int make_map(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
             const unsigned long long* strides_elems, const unsigned* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_elems[i] * sizeof(__half);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, (void*)base, gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

template <int BN, int STAGES>
int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
              const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * TC_BM * 128 + 2 * BN * 128) + 1024;
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
// contiguous, reduction index = smem row; descriptor layout and strides verified by tests/probes/probe_mnmajor*.cu:
// 16-bit operands take the ordinary 128B swizzle, 32-bit ones would need the 32-byte-atom variant).  A smem
// operand is a stack of "slabs" (64 channels x 64 pixels = 64 rows of 128 bytes, LBO = 8192 apart; 8-row groups
// SBO = 1024 apart); one 5-D TMA box (64 channels-in-slab, w, h, n, slab) fills all slabs of an operand, the tap
// shift is a plain coordinate offset with out-of-bounds zero fill.  One MMA consumes 16 pixel rows (K = 16).
// Grid: (co tiles x ci tiles, taps, pixel splits); partial sums are reduced with fp32 atomics.
// ------------------------------------------------------------------------------------------------
struct TcWgradParams {
  int N, OH, OW;        // pixel grid of G
  int cw, chh, cn;      // pixel chunk of one stage: cw x chh x cn = 64 pixels (w fastest)
  int wchunks, hchunks; // chunks per row / per image column
  int total_chunks;
  int chunks_per_cta;
  int Kp, C;            // rows (co) and columns (ci) of each dW[t]
  int cotiles;          // blockIdx.x = citile * cotiles + cotile
  int nsplit;
  float* dw;            // [ntaps_out][Kp][C]
  const unsigned* amax_g; // operand scales (see TcParams)
  const unsigned* amax_x;
  // entry e: dW[out[e]] += G[n + gdn[e], h, w]^T * X[n + xdn[e], h + xdh[e], w + xdw[e]]
  int gdn[TC_MAX_TAPS], xdh[TC_MAX_TAPS], xdw[TC_MAX_TAPS], xdn[TC_MAX_TAPS], out[TC_MAX_TAPS];
};

constexpr int WG_PIX = 64;                     // pixels per stage (4 MMA k-steps of 16)
constexpr int WG_CH = 64;                      // channels per slab = one 128-byte row
constexpr uint32_t WG_SLAB = WG_PIX * 128;     // bytes of one slab: 64 pixels x 64 channels

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// MN-major slab stack, 128B swizzle: LBO = slab pitch, SBO = 8 rows * 128 B
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WG_SLAB >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16_mn(int M, int N) {
  return make_idesc_f16(M, N) | (1u << 15) | (1u << 16);   // A and B both MN-major
}

template <int BNW, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_gh, const __grid_constant__ CUtensorMap map_gl,
                     const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const TcWgradParams p) {
  constexpr int NSB = (BNW + WG_CH - 1) / WG_CH;          // slabs of the X operand
  constexpr uint32_t A_BYTES = 2 * WG_SLAB;               // 128 co = 2 slabs = 16 KB
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
      const int aslab = cotile * 2, bslab = citile * NSB;
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
      const uint32_t idesc = make_idesc_f16_mn(TC_BM, BNW);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nk; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t ah = make_desc_mn(sa), al = make_desc_mn(sa + A_BYTES);
        const uint64_t bh = make_desc_mn(sa + 2 * A_BYTES), bl = make_desc_mn(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
        for (int k = 0; k < WG_PIX / 16; ++k) {
          const uint64_t koff = (uint64_t)((k * 16 * 128) >> 4);    // 16 pixel rows of 128 bytes
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          if (p.nsplit == 3) {
            umma_f16(tmem_base, al + koff, bh + koff, idesc, acc);
            umma_f16(tmem_base, ah + koff, bl + koff, idesc, 1u);
            umma_f16(tmem_base, ah + koff, bh + koff, idesc, 1u);
          } else {
            umma_f16(tmem_base, ah + koff, bh + koff, idesc, acc);
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
    const TcScale sc = tc_scale(p.amax_g, p.amax_x);
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BNW; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < BNW && c + j < cleft) atomicAdd(drow + c + j, __uint_as_float(r[j]) * sc.fa * sc.fb);
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
  constexpr int NSB = (BNW + WG_CH - 1) / WG_CH;
  constexpr uint32_t A_BYTES = 2 * WG_SLAB;
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
          const int aslab = (copair * 2 + hf) * 2;      // past the last slab: out of bounds, zero filled
          tma_load_5d(sa, &map_gh, &fullA[as], 0, w0, h, n + gdn, aslab);
          if (split3) tma_load_5d(sa + A_BYTES, &map_gl, &fullA[as], 0, w0, h, n + gdn, aslab);
          if (++as == AS) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16_mn(TC_BM, BNW);
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
          for (int k = 0; k < WG_PIX / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 16 * 128) >> 4);
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            if (split3) {
              umma_f16(d, al + koff, bh + koff, idesc, acc);
              umma_f16(d, ah + koff, bl + koff, idesc, 1u);
              umma_f16(d, ah + koff, bh + koff, idesc, 1u);
            } else {
              umma_f16(d, ah + koff, bh + koff, idesc, acc);
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
    const TcScale sc = tc_scale(p.amax_g, p.amax_x);
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
            if (c + j < BNW && c + j < cleft) atomicAdd(drow + c + j, __uint_as_float(r[j]) * sc.fa * sc.fb);
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
  constexpr size_t smem = (size_t)3 * (2 * 2 * WG_SLAB) + (size_t)2 * (2 * ((BNW + WG_CH - 1) / WG_CH) * WG_SLAB) + 1024;
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
  constexpr size_t smem = (size_t)STAGES * (2 * 2 * WG_SLAB + 2 * ((BNW + WG_CH - 1) / WG_CH) * WG_SLAB) + 1024;
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
  constexpr size_t smem = (size_t)3 * (2 * (TALL ? 160 : TC_BM) * 128) + (size_t)2 * (2 * BN * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, TALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  conv_tc2_kernel<BN, TALL><<<grid, TC_THREADS, smem, stream>>>(ah, al, bh, bl, p);
  return (int)cudaGetLastError();
}

// CTA-pair launch: grid.x = tile pairs rounded up to an even count, clusters of two consecutive CTAs along x
template <int BN, bool TALL>
int launch_tc2x(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)3 * (2 * (TALL ? 160 : TC_BM) * 128) + (size_t)4 * (2 * (BN / 2) * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2x_kernel<BN, TALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((grid.x + 1) / 2 * 2, grid.y, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, conv_tc2x_kernel<BN, TALL>, ah, al, bh, bl, p);
}

// persistent CTA-pair launch: one cluster of two CTAs per SM pair
template <int BN, bool TALL>
int launch_tcp(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
               const TcParams& p, int ntiles, int ntn, cudaStream_t stream) {
  constexpr size_t smem = (size_t)3 * (2 * (TALL ? 160 : TC_BM) * 128) + (size_t)4 * (2 * (BN / 2) * 128) + 1024;
  static bool configured = false;
  static int sm_pairs = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tcp_kernel<BN, TALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sm_pairs = sms / 2;
    configured = true;
  }
  const int nunits = ((ntiles + 1) / 2) * ntn;
  const int nclusters = nunits < sm_pairs ? nunits : sm_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nclusters, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, conv_tcp_kernel<BN, TALL>, ah, al, bh, bl, p, nunits, ntn);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI.
//   xh/xl : source activations, fp16 hi / lo parts (og_prep_split), NHWC [SN][SH][SW][C] contiguous, C % 8 == 0
//           (SN = N, or 4*N for a space-to-depth source: phase block (a*2+b) holds x[:, a::2, b::2])
//   wh/wl : weights fp16 hi / lo, [ntaps_w][Kw][C] (og_pack_weights_f16 with transposed=1 for fprop), Kw rows
//   amax_x / amax_w : device words holding the float bits of max|.| of the two operand tensors (their scaling)
//   y     : output NHWC, channel count K (K % 4 == 0), element strides ysn/ysh/ysw, spatial bounds OHf x OWf
//   rows of the GEMM are the pixel grid N x OH x OW; output pixel (osy*h + opy, osx*w + opx)
//   taps  : ntaps quadruples (dh, dw, dn, weight tap index): source pixel = image n + dn, (h + dh, w + dw), zero
//           outside [0,SH)x[0,SW);  bias/act: optional epilogue (bias[K], OG_ACT_*)
//   tap_layout = 1 promises that the taps come in groups of three with equal dw, dn and dh = d0, d0+1, d0+2 (3x3
//           kernels listed column by column), which lets the kernel reuse one tall pixel patch for three taps
// ------------------------------------------------------------------------------------------------
OG_API int og_conv2d_tc(const void* xh, const void* xl, const unsigned* amax_x, int N, int SN, int SH, int SW, int C,
                        const void* wh, const void* wl, const unsigned* amax_w, int ntaps_w, int Kw, float* y, int OH,
                        int OW, int K, long long ysn,
                        long long ysh, long long ysw, int OHf, int OWf, int osy, int osx, int opy, int opx,
                        const int* taps_host, int ntaps, int tap_layout, int nsplit, const float* bias, int act,
                        float slope, cudaStream_t stream) {
  if (C % 8 || K % 4 || ntaps < 1 || ntaps > TC_MAX_TAPS || (nsplit != 1 && nsplit != 3)) return (int)cudaErrorInvalidValue;
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
  // N need not be a multiple of TN: GEMM rows are independent, the last tile's rows n >= N are never stored (row_ok),
  // whatever they read (zero fill past the tensor, or -- in a phase-stacked source -- images of the next phase block)
  p.TN = TN; p.TH = TH; p.TW = TW;
  p.bias = bias; p.act = act; p.slope = slope;
  p.tall = 0;
  p.stacked = 0;
  p.tiles_w = og_cdiv(OW, TW);
  p.tiles_h = og_cdiv(OH, TH);
  const int tiles_n = og_cdiv(N, TN);
  p.cchunks = og_cdiv(C, TC_BK);
  p.klast = og_cdiv(C - (p.cchunks - 1) * TC_BK, 16);
  p.amax_a = amax_x; p.amax_b = amax_w;
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
  // persistent CTA pairs (cta_group::2, double-buffered accumulators): every BN = 208 / 256 problem with enough tiles
  static const int tcp = getenv("OG_TCP") ? atoi(getenv("OG_TCP")) : 2;
  static const long long tcp_min = getenv("OG_TCP_MIN") ? atoll(getenv("OG_TCP_MIN")) : 64;
  if (tcp && p.ksplit == 1 && (tcp > 1 || BNsel == 208 || BNsel == 256) && (long long)grid.x * grid.y >= tcp_min) {
    CUtensorMap pbh, pbl;
    unsigned hbox[3] = {(unsigned)TC_BK, (unsigned)(BNsel / 2), 1u};
    // <= 16 output channels with the lo copy stored right behind the hi copy: stack them along N (two MMAs per k-step)
    static const int stacked_on = getenv("OG_STACKED") ? atoi(getenv("OG_STACKED")) : 1;
    // up to 16 output channels: N = 32 = [16 hi | 16 lo]; 17..32: N = 64 = [32 hi | 32 lo]
    const int stk_rows = (K <= 16 && Kw <= 16) ? 16 : (K <= 32 && Kw <= 32) ? 32 : 0;
    p.stacked = (stacked_on && nsplit == 3 && BNsel == 32 && stk_rows > 0 && wl != nullptr &&
                 (const char*)wl == (const char*)wh + sizeof(__half) * (size_t)ntaps_w * Kw * C) ? 1 : 0;
    if (p.stacked) {
      unsigned long long sdims[4] = {(unsigned long long)C, (unsigned long long)Kw, 2ull, (unsigned long long)ntaps_w};
      unsigned long long sstr[3] = {(unsigned long long)C, (unsigned long long)ntaps_w * Kw * C, (unsigned long long)Kw * C};
      unsigned sbox[4] = {(unsigned)TC_BK, (unsigned)stk_rows, 1u, 1u};
      if ((rc = make_map(&pbh, wh, 4, sdims, sstr, sbox))) return rc;
      pbl = pbh;
      if (stk_rows == 32) return launch_tcp<64, false>(mah, mal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
    } else {
      if ((rc = make_map(&pbh, wh, 3, bdims, bstr, hbox))) return rc;
      if (nsplit == 3) {
        if ((rc = make_map(&pbl, wl, 3, bdims, bstr, hbox))) return rc;
      } else {
        pbl = pbh;
      }
    }
    static const bool no_tall = getenv("OG_NO_TALL") != nullptr;
    if (!no_tall && tap_layout == 1 && ntaps % 3 == 0 && BNsel == 208 && TN == 1 && TH == 8 && TW == 16) {
      unsigned tbox[4] = {(unsigned)TC_BK, 16u, 10u, 1u};
      CUtensorMap tah, tal;
      if ((rc = make_map(&tah, xh, 4, adims, astr, tbox))) return rc;
      if (nsplit == 3) {
        if ((rc = make_map(&tal, xl, 4, adims, astr, tbox))) return rc;
      } else {
        tal = tah;
      }
      p.tall = 1;
      return launch_tcp<208, true>(tah, tal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
    }
    switch (BNsel) {
      case 32:  return launch_tcp<32, false>(mah, mal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
      case 64:  return launch_tcp<64, false>(mah, mal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
      case 112: return launch_tcp<112, false>(mah, mal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
      case 208: return launch_tcp<208, false>(mah, mal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
      default:  return launch_tcp<256, false>(mah, mal, pbh, pbl, p, (int)grid.x, (int)grid.y, stream);
    }
  }
  // large problems: two pixel tiles per CTA share every weight stage (see conv_tc2_kernel)
  static const bool no_tc2 = getenv("OG_NO_TC2") != nullptr;
  static const long long tc2_min = getenv("OG_TC2_MIN") ? atoll(getenv("OG_TC2_MIN")) : 2 * 148 * 2;
  if (!no_tc2 && p.ksplit == 1 && (BNsel == 208 || BNsel == 256) && (long long)grid.x * grid.y >= tc2_min) {
    dim3 grid2((grid.x + 1) / 2, grid.y, 1);
    static const bool no_tall = getenv("OG_NO_TALL") != nullptr;
    // CTA pairs (cta_group::2): each CTA loads half of every weight stage (box of BNsel / 2 rows)
    static const bool pair = getenv("OG_TC2X") != nullptr && atoi(getenv("OG_TC2X")) != 0;
    CUtensorMap pbh = mbh, pbl = mbl;
    if (pair) {
      unsigned hbox[3] = {(unsigned)TC_BK, (unsigned)(BNsel / 2), 1u};
      if ((rc = make_map(&pbh, wh, 3, bdims, bstr, hbox))) return rc;
      if (nsplit == 3) {
        if ((rc = make_map(&pbl, wl, 3, bdims, bstr, hbox))) return rc;
      } else {
        pbl = pbh;
      }
    }
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
      if (pair) return launch_tc2x<208, true>(tah, tal, pbh, pbl, p, grid2, stream);
      return launch_tc2<208, true>(tah, tal, mbh, mbl, p, grid2, stream);
    }
    if (pair)
      return BNsel == 208 ? launch_tc2x<208, false>(mah, mal, pbh, pbl, p, grid2, stream)
                          : launch_tc2x<256, false>(mah, mal, pbh, pbl, p, grid2, stream);
    return BNsel == 208 ? launch_tc2<208, false>(mah, mal, mbh, mbl, p, grid2, stream)
                        : launch_tc2<256, false>(mah, mal, mbh, mbl, p, grid2, stream);
  }
  switch (BNsel) {
    // narrow output tiles are latency bound (tiny MMAs): two shallow-pipelined CTAs per SM instead of one deep one
    case 32:  return launch_tc<32, 2>(mah, mal, mbh, mbl, p, grid, stream);
    case 64:  return launch_tc<64, 2>(mah, mal, mbh, mbl, p, grid, stream);
    case 112: return launch_tc<112, 3>(mah, mal, mbh, mbl, p, grid, stream);
    case 208: return launch_tc<208, 2>(mah, mal, mbh, mbl, p, grid, stream);
    default:  return launch_tc<256, 2>(mah, mal, mbh, mbl, p, grid, stream);
  }
}


// ------------------------------------------------------------------------------------------------
// weight gradient:  dw[t][co][ci] (fp32, [ntaps][Kp][C], zero-filled here) = sum_pixels G^T * X(shifted by tap t)
//   gh/gl : output gradient fp16 hi/lo, NHWC [GN][OH][OW][Kp] (og_prep_split; GN = N, or 4N space-to-depth blocks)
//   xh/xl : source activations fp16 hi/lo, NHWC [XN][SH][SW][C] (og_prep_split: plain, reflection-padded or
//           space-to-depth); both need 512 readable bytes after the last element (partial last channel slab)
//   amax_g / amax_x : operand scales as in og_conv2d_tc
//   entries: nentries quintuples (g image offset, dh, dw, x image offset, output tap):
//           dw[tap] += sum_{n,h,w} G[n + gdn, h, w, :]^T  X[n + xdn, h + dh, w + dw, :]   (out of range = 0)
// ------------------------------------------------------------------------------------------------
OG_API int og_conv2d_wgrad_tc(const void* gh, const void* gl, const unsigned* amax_g, int N, int GN, int OH, int OW,
                              int Kp, const void* xh, const void* xl, const unsigned* amax_x, int XN, int SH, int SW,
                              int C, float* dw,
                              int ntaps_out, const int* entries_host, int nentries, int nsplit, cudaStream_t stream) {
  if (nentries < 1 || nentries > TC_MAX_TAPS || (nsplit != 1 && nsplit != 3) || Kp % 8 || C % 8)
    return (int)cudaErrorInvalidValue;
  OG_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)ntaps_out * Kp * C, stream));
  if ((long long)N * OH * OW == 0) return 0;
  TcWgradParams p;
  p.N = N; p.OH = OH; p.OW = OW;
  // one stage = 64 pixels: a cw x chh x cn patch of the gradient grid (w fastest)
  int cw = 1;
  while (cw * 2 <= WG_PIX && OW % (cw * 2) == 0) cw *= 2;
  int chh = 1;
  while (cw * chh * 2 <= WG_PIX && OH % (chh * 2) == 0) chh *= 2;
  const int cn = WG_PIX / (cw * chh);
  // a last image group with n >= N contributes nothing as long as ONE operand is unstacked (TMA zero-fills it there)
  if (N % cn != 0 && GN != N && XN != N) return (int)cudaErrorInvalidValue;
  p.cw = cw; p.chh = chh; p.cn = cn;
  p.wchunks = OW / cw;
  p.hchunks = OH / chh;
  p.total_chunks = p.wchunks * p.hchunks * og_cdiv(N, cn);
  p.Kp = Kp; p.C = C; p.nsplit = nsplit; p.dw = dw;
  p.amax_g = amax_g; p.amax_x = amax_x;
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
  // pixel splits: the CTAs are equal-sized and run one per SM, so fill at most two full waves of 148 (one CTA more
  // would cost a whole extra round); at least 8 chunks per CTA
  int splits = 296 / (cotiles * citiles * nentries);
  if (splits < 1) splits = 1;
  int maxs = p.total_chunks / 8;
  if (maxs < 1) maxs = 1;
  if (splits > maxs) splits = maxs;
  p.chunks_per_cta = og_cdiv(p.total_chunks, splits);
  splits = og_cdiv(p.total_chunks, p.chunks_per_cta);
  dim3 grid(cotiles * citiles, nentries, splits);
  CUtensorMap mgh, mgl, mxh, mxl;
  // [slab][n][h][w][64 channels of the slab]: the slab dimension strides by 64 halves inside a pixel's channel vector
  unsigned long long gd[5] = {(unsigned long long)WG_CH, (unsigned long long)OW, (unsigned long long)OH,
                              (unsigned long long)GN, (unsigned long long)og_cdiv(Kp, WG_CH)};
  unsigned long long gs[4] = {(unsigned long long)Kp, (unsigned long long)OW * Kp, (unsigned long long)OH * OW * Kp,
                              (unsigned long long)WG_CH};
  unsigned gb[5] = {(unsigned)WG_CH, (unsigned)cw, (unsigned)chh, (unsigned)cn, 2u};
  unsigned long long xd[5] = {(unsigned long long)WG_CH, (unsigned long long)SW, (unsigned long long)SH,
                              (unsigned long long)XN, (unsigned long long)og_cdiv(C, WG_CH)};
  unsigned long long xs[4] = {(unsigned long long)C, (unsigned long long)SW * C, (unsigned long long)SH * SW * C,
                              (unsigned long long)WG_CH};
  unsigned xb[5] = {(unsigned)WG_CH, (unsigned)cw, (unsigned)chh, (unsigned)cn, (unsigned)((BNsel + WG_CH - 1) / WG_CH)};
  int rc;
  if ((rc = make_map(&mgh, gh, 5, gd, gs, gb))) return rc;
  if ((rc = make_map(&mxh, xh, 5, xd, xs, xb))) return rc;
  if (nsplit == 3) {
    if ((rc = make_map(&mgl, gl, 5, gd, gs, gb))) return rc;
    if ((rc = make_map(&mxl, xl, 5, xd, xs, xb))) return rc;
  } else {
    mgl = mgh;
    mxl = mxh;
  }
  static const bool no_tc2 = getenv("OG_NO_TC2") != nullptr;
  if (!no_tc2 && cotiles >= 2 && (BNsel == 208 || BNsel == 256)) {
    // two co-tiles per CTA share every X stage (L2 -> SM traffic is the bound of this kernel)
    const int copairs = (cotiles + 1) / 2;
    int sp = 296 / (copairs * citiles * nentries);
    if (sp > maxs) sp = maxs;
    if (sp < 1) sp = 1;
    p.chunks_per_cta = og_cdiv(p.total_chunks, sp);
    sp = og_cdiv(p.total_chunks, p.chunks_per_cta);
    dim3 grid2(copairs * citiles, nentries, sp);
    return BNsel == 208 ? launch_wgrad2<208>(mgh, mgl, mxh, mxl, p, grid2, stream)
                        : launch_wgrad2<256>(mgh, mgl, mxh, mxl, p, grid2, stream);
  }
  switch (BNsel) {
    case 32:  return launch_wgrad<32, 2>(mgh, mgl, mxh, mxl, p, grid, stream);
    case 64:  return launch_wgrad<64, 2>(mgh, mgl, mxh, mxl, p, grid, stream);
    case 112: return launch_wgrad<112, 3>(mgh, mgl, mxh, mxl, p, grid, stream);
    case 208: return launch_wgrad<208, 2>(mgh, mgl, mxh, mxl, p, grid, stream);
    default:  return launch_wgrad<256, 2>(mgh, mgl, mxh, mxl, p, grid, stream);
  }
}

// max|x| over a tensor as float bits (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long n4, int tail,
                                                   unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = ldg4(x + i * 4);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) m = fmaxf(m, fabsf(x[n4 * 4 + threadIdx.x]));
  m = warp_max(m);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) atomicMax(out, __float_as_uint(m));
  }
}

__device__ __forceinline__ void split_f16(float v, float scale, __half& hi, __half& lo) {
  v *= scale;
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__global__ void __launch_bounds__(256) prep_split_kernel(const float* __restrict__ x, int N, int H, int W, int C8, int pad,
                                                         int s2d, long long total, const unsigned* __restrict__ amax,
                                                         __half* __restrict__ xh, __half* __restrict__ xl) {
  // output: [N][H+2p][W+2p][C] or, for s2d, [4][N][H/2][W/2][C] with phase block (a*2+b) = x[:, a::2, b::2]
  const int Hp = s2d ? H / 2 : H + 2 * pad, Wp = s2d ? W / 2 : W + 2 * pad;
  const float scale = og_exp2i(og_scale_exp(__ldg(amax)));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    long long t = i / C8;
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
    const float* src = x + (((n * H + sh) * W + sw) * C8 + c) * 8;
    const float4 v0 = ldg4(src), v1 = ldg4(src + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f16(v[j], scale, hi[j], lo[j]);
    *reinterpret_cast<uint4*>(xh + i * 8) = *reinterpret_cast<const uint4*>(hi);
    if (xl) *reinterpret_cast<uint4*>(xl + i * 8) = *reinterpret_cast<const uint4*>(lo);
  }
}

// fp16 hi/lo operand copies of an NHWC fp32 tensor, scaled by the power of two that og_scale_exp derives from
// max|x| (*amax as float bits: computed here, or, with amax_ready, left there by the kernel that produced x --
// any upper bound of max|x| is valid, it only positions the 22-bit window).  pad = 0: plain; pad = 1: nn.ReflectionPad2d(1) halo (model.py:67)
// materialised while splitting; s2d = 1: four space-to-depth phase blocks.  xl may be null (single-product mode).
OG_API int og_prep_split(const float* x, int N, int H, int W, int C, int pad, int s2d, unsigned* amax, int amax_ready,
                         void* xh, void* xl, cudaStream_t stream) {
  if (C % 8 || pad < 0 || pad > 1 || (s2d && (pad || (H & 1) || (W & 1)))) return (int)cudaErrorInvalidValue;
  const long long n4 = (long long)N * H * W * C / 4;
  if (!amax_ready) OG_CHECK(cudaMemsetAsync(amax, 0, sizeof(unsigned), stream));
  if (n4 == 0) return 0;
  if (!amax_ready) {      // the producer of x did not leave max|x| behind: one extra read pass
    long long ab = (n4 + 255) / 256;
    if (ab > 148LL * 8) ab = 148LL * 8;
    amax_kernel<<<(int)ab, 256, 0, stream>>>(x, n4, 0, amax);
  }
  long long total = s2d ? (long long)N * H * W * (C / 8) : (long long)N * (H + 2 * pad) * (W + 2 * pad) * (C / 8);
  long long b = (total + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  prep_split_kernel<<<(int)b, 256, 0, stream>>>(x, N, H, W, C / 8, pad, s2d, total, amax, (__half*)xh, (__half*)xl);
  OG_RETURN_LAST_ERROR();
}

// *amax = float bits of max|x[0..n)| (x 16-byte aligned); used for the weight tensors before og_pack_weights_f16
OG_API int og_amax(const float* x, long long n, unsigned* amax, cudaStream_t stream) {
  OG_CHECK(cudaMemsetAsync(amax, 0, sizeof(unsigned), stream));
  if (n == 0) return 0;
  const long long n4 = n / 4;
  long long ab = (n4 + 255) / 256;
  if (ab < 1) ab = 1;
  if (ab > 148LL * 8) ab = 148LL * 8;
  amax_kernel<<<(int)ab, 256, 0, stream>>>(x, n4, (int)(n - n4 * 4), amax);
  OG_RETURN_LAST_ERROR();
}
