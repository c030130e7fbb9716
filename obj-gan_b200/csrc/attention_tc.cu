// Grid attention forward (GlobalAttentionGeneral / ATT_NET, ref: GlobalAttention.py:83-122) on the tensor cores:
// a persistent tcgen05 kernel for the large maps of the hot path (Q % 128 == 0, 48 channels, L <= 32 words).
//
//   scores[q][l] = sum_c h[q][c] * src[c][l]      (128 x 48) x (48 x 32)   -> TMEM columns [0, 32)
//   p = softmax_l(scores) (caption mask quirk: GlobalAttention.py:108)       -> attn[b][l][q]
//   wc[q][c]     = sum_l p[q][l] * src[c][l]      (128 x 32) x (32 x 48)   -> TMEM columns [32, 80)
//
// Why tensor cores for a bandwidth-bound op: the two tiny GEMMs cost 1920 fp32 FMAs per query; on the CUDA cores that
// instruction stream (plus the shared-memory operand traffic it needs) capped the kernel at 0.37 of the HBM
// bandwidth.  Here the FMA work is 15 small MMAs per 128 queries and the threads only move and convert data.
//
// fp32 parity: the convolution engine's error-compensated scheme (conv_tc.cu): x * 2^k = hi + lo with hi, lo fp16
// (22 mantissa bits), three products hi*hi + lo*hi + hi*lo into the fp32 TMEM accumulator.  The power of two comes from
// max|x| of the 32 rows a WARP stages (a register max + one shuffle reduction; no extra pass over the tensor, no block
// barrier: the same warp owns those rows in the softmax and undoes its factor there) resp. of the image's
// word projections; the accumulator is scaled back before the softmax / the store.  hi is the fp32 value truncated to
// 11 significant bits (one LOP3), lo the exact remainder; the only conversions are packed cvt.rn.f16x2 (no F2F).
//
// One CTA = 128 threads = 128 queries per trip; thread t owns query row t (= TMEM lane t).  Per trip:
//   (1) the h tile (24.5 KB, contiguous) arrives through coalesced 16-byte loads issued one trip AHEAD (registers),
//       is split and stored into two K-major 128B-swizzled A tiles;                        -> MMA 1 (9 instructions)
//   (2) tcgen05.ld of the thread's score row, softmax in registers, coalesced attn stores (lanes = adjacent queries),
//       the probabilities are split into the same two tiles (K = 32);                      -> MMA 2 (6 instructions)
//   (3) tcgen05.ld of the context row, staged through shared memory, written with coalesced 16-byte stores.
// CTAs are persistent over the query tiles of ONE image (the word-projection operand tiles are built once; tiles are
// dealt with a fixed stride -- drawing them from a per-image atomic counter was measured and lost 2.7 us to the extra
// memset node and the atomics: 35.6 vs 32.9 us).  52 KB of
// shared memory and <= 128 registers per thread: four CTAs per SM cover each other's MMA round trips.
#include "common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

constexpr int AT_Q = 128;          // queries per tile = threads per CTA
constexpr int AT_C = 48;           // channels (idf == row stride)
constexpr int AT_LP = 32;          // words padded to the MMA N / K extent
constexpr uint32_t AT_A_BYTES = 128 * 128;            // one fp16 copy of the A tile: 128 rows x 128-byte swizzle rows
constexpr uint32_t AT_B1_BYTES = AT_LP * 128;         // src as [l][c]  (N = 32 rows, K = 48)
constexpr uint32_t AT_B2_BYTES = AT_C * 128;          // src as [c][l]  (N = 48 rows, K = 32)
constexpr uint32_t AT_SMEM = 2 * AT_A_BYTES + 2 * AT_B1_BYTES + 2 * AT_B2_BYTES;     // 52 KB: four CTAs per SM
constexpr int AT_STAGE_PITCH = AT_C + 4;              // floats per staged context row (conflict-free 16-byte accesses)
constexpr int AT_MAXB = 128;                          // batch sizes whose caption masks fit the bit table
constexpr int AT_PSCALE = 13;                          // probabilities enter MMA 2 as p * 2^13 (fp16 subnormals start at 7e-9)

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(at_smem_u32(bar)), "r"(count));
}
// bounded wait: a protocol error traps after ~1 s instead of hanging the GPU
__device__ __forceinline__ void at_mbar_wait(uint64_t* bar, uint32_t parity) {
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
        : "r"(at_smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if ((++spins & 255) == 0 && clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void at_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void at_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void at_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void at_umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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
__device__ __forceinline__ void at_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(at_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void at_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void at_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void at_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (same form as conv_tc.cu)
__device__ __forceinline__ uint64_t at_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D = f32, A = B = f16 (format code 0), both K-major
__device__ __forceinline__ uint32_t at_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row r, 16-bit column k) in a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t at_sw_off(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// x (already scaled) = hi + lo: hi = x rounded to 11 significant bits in the fp32 domain (half an ulp added to the
// magnitude, then truncated: exact in fp16 while |x| >= 2^-14), lo = x - hi exactly (|lo| <= 2^-11 |x|)
__device__ __forceinline__ void at_hilo(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
// two floats -> packed f16x2 (first argument in the low half)
__device__ __forceinline__ uint32_t at_cvt2(float lo_el, float hi_el) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_el), "f"(lo_el));
  return r;
}
__device__ __forceinline__ void at_sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void at_sts64(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void at_sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float4 at_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float at_amax4(float m, const float4& v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// max over the CTA's 128 threads (4 warps) through a 4-word shared scratch; every thread gets the result
__device__ __forceinline__ float at_cta_max(float m, float* scratch) {
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = m;
  __syncthreads();
  m = fmaxf(fmaxf(scratch[0], scratch[1]), fmaxf(scratch[2], scratch[3]));
  return m;
}

// three products per k-step, small terms first: lo*hi, hi*lo, hi*hi
__device__ __forceinline__ void at_issue(uint32_t d_tmem, const uint64_t (&a)[2], const uint64_t (&b)[2], uint32_t idesc,
                                         int ksteps) {
  for (int k = 0; k < ksteps; ++k) {
    const uint64_t koff = (uint64_t)((k * 32) >> 4);
    at_umma(d_tmem, a[1] + koff, b[0] + koff, idesc, k > 0 ? 1u : 0u);
    at_umma(d_tmem, a[0] + koff, b[1] + koff, idesc, 1u);
    at_umma(d_tmem, a[0] + koff, b[0] + koff, idesc, 1u);
  }
}

// LQ: words handled in registers (20 covers the 12..18-word captions of the hot path; 32 = the MMA extent)
//
// Thread <-> data mapping of the two coalesced passes over a tile (1536 float4 = 128 rows x 12): warp w owns rows
// [32w, 32w + 32); it walks them in four groups g of 8 rows (= 96 float4 = three warp-wide accesses j), so float4
// number 32j + lane of a group is (row r8[j], quad c8[j]) with r8 / c8 fixed per thread, and a group step is exactly
// one 1024-byte swizzle atom of the operand tile: every shared-memory offset is base[j] + g * constant.
template <int LQ>
__global__ void __launch_bounds__(AT_Q, 4)
att_general_fwd_tc_kernel(const float* __restrict__ h, const float* __restrict__ src,
                          const unsigned char* __restrict__ mask, int B, int Q, int L, int nslots,
                          float* __restrict__ wc, float* __restrict__ attn) {
  extern __shared__ uint8_t at_smem_raw[];
  const uint32_t smem = (at_smem_u32(at_smem_raw) + 1023u) & ~1023u;   // shared-window address of the operand tiles
  const uint32_t sA = smem;                            // hi, lo; reused for the probabilities and the staged context
  const uint32_t sB1 = smem + 2 * AT_A_BYTES;
  const uint32_t sB2 = sB1 + 2 * AT_B1_BYTES;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[2][4];                          // CTA max scratch
  __shared__ uint32_t mbits[AT_MAXB];                  // caption masks of the batch, one bit per word

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int b = blockIdx.x / nslots, slot = blockIdx.x - b * nslots;
  const int ntiles = Q / AT_Q;

  if (t == 0) {
    at_mbar_init(&bar_s, 1);
    at_mbar_init(&bar_o, 1);
    at_fence_barrier_init();
  }
  if (warp == 0) at_tmem_alloc(&tmem_base_smem, 128);
  // per-thread constants of the tile walk
  uint32_t offA[3], offS[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int f = j * 32 + lane, r8 = (f * 43691) >> 19, c8 = f - r8 * 12;        // f / 12 for f < 96
    offA[j] = (uint32_t)(warp * 4 * 1024) + at_sw_off(r8, c8 * 4);
    offS[j] = (uint32_t)(((warp * 32 + r8) * AT_STAGE_PITCH + c8 * 4) * 4);
  }
  // prefetch the first tile
  float4 hv[12];
  int qt = slot;
  {
    const float4* hp = reinterpret_cast<const float4*>(h + ((long long)b * Q + (long long)qt * AT_Q) * AT_C) + warp * 384 + lane;
#pragma unroll
    for (int i = 0; i < 12; ++i) hv[i] = __ldg(hp + i * 32);
  }
  // word projections of this image -> the two operand layouts B1[l][c] and B2[c][l] as scaled fp16 hi / lo (zero beyond
  // L), straight from registers: thread t holds word l = t & 31 of channels (t >> 5) + 4j
  int ksrc;
  {
    const float* sb = src + (long long)b * AT_C * L;
    const int l = lane, c0 = warp;
    float v[12];
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      v[j] = l < L ? __ldg(sb + (c0 + 4 * j) * L + l) : 0.f;
      m = fmaxf(m, fabsf(v[j]));
    }
    if (mask && t < B) {
      uint32_t w = 0;
      for (int i = 0; i < L; ++i) w |= (mask[t * L + i] ? 1u : 0u) << i;
      mbits[t] = w;
    }
    m = at_cta_max(m, red[0]);
    ksrc = og_scale_exp(__float_as_uint(m));
    const float s = og_exp2i(ksrc);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int c = c0 + 4 * j;
      float hi, lo;
      at_hilo(v[j] * s, hi, lo);
      const uint32_t pk = at_cvt2(hi, lo);             // low half = hi, high half = lo
      const uint32_t o1 = at_sw_off(l, c), o2 = at_sw_off(c, l);
      at_sts16(sB1 + o1, pk & 0xFFFFu);
      at_sts16(sB1 + AT_B1_BYTES + o1, pk >> 16);
      at_sts16(sB2 + o2, pk & 0xFFFFu);
      at_sts16(sB2 + AT_B2_BYTES + o2, pk >> 16);
    }
  }
  at_tc_fence_before();
  __syncthreads();
  at_tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);

  uint64_t dA[2], dB1[2], dB2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    dA[i] = at_desc_sw128(sA + i * AT_A_BYTES);
    dB1[i] = at_desc_sw128(sB1 + i * AT_B1_BYTES);
    dB2[i] = at_desc_sw128(sB2 + i * AT_B2_BYTES);
  }
  const uint32_t idesc1 = at_idesc_f16(128, AT_LP), idesc2 = at_idesc_f16(128, AT_C);
  const uint32_t rowP = sA + (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);   // row t of an A tile; chunk ch at ((ch ^ (t & 7)) << 4)
  const uint32_t rowS = sA + (uint32_t)(t * AT_STAGE_PITCH) * 4u;
  const float so = og_exp2i(-AT_PSCALE) * og_exp2i(-ksrc);                   // context rows: undo both operand scales
  const float ss2 = og_exp2i(-ksrc);
  // caption-mask row of query (b, q): sample (b * Q + q) mod B (ref: GlobalAttention.py:108); advanced incrementally
  uint32_t mrow = ((uint32_t)b * (uint32_t)Q + (uint32_t)(qt * AT_Q + t)) % (uint32_t)B;
  const uint32_t mstep = ((uint32_t)nslots * AT_Q) % (uint32_t)B;

  uint32_t phase = 0;
  for (; qt < ntiles; qt += nslots, phase ^= 1) {
    const int q0 = qt * AT_Q;
    // ---------------- (1) scale by a power of two, split into the two A copies ----------------
    // The scale is per WARP: warp w stages rows [32w, 32w + 32) here and owns the same rows (TMEM lanes) in the softmax,
    // where it divides its scores by its own factor again -- a row scale of A is a row scale of A.B -- so a shuffle
    // reduction is enough (no block-wide barrier).
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) m = at_amax4(m, hv[i]);
    m = warp_max(m);
    const int kh = og_scale_exp(__float_as_uint(m));
    const float sh = og_exp2i(kh);
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 x = hv[g * 3 + j];
        float h0, l0, h1, l1, h2, l2, h3, l3;
        at_hilo(x.x * sh, h0, l0); at_hilo(x.y * sh, h1, l1);
        at_hilo(x.z * sh, h2, l2); at_hilo(x.w * sh, h3, l3);
        const uint32_t a = sA + offA[j] + g * 1024;
        at_sts64(a, at_cvt2(h0, h1), at_cvt2(h2, h3));
        at_sts64(a + AT_A_BYTES, at_cvt2(l0, l1), at_cvt2(l2, l3));
      }
    // the next tile's loads fly during the rest of this trip
    if (qt + nslots < ntiles) {
      const float4* hp = reinterpret_cast<const float4*>(h + ((long long)b * Q + (long long)(qt + nslots) * AT_Q) * AT_C) + warp * 384 + lane;
#pragma unroll
      for (int i = 0; i < 12; ++i) hv[i] = __ldg(hp + i * 32);
    }
    at_fence_proxy_async();
    at_tc_fence_before();
    __syncthreads();
    if (t == 0) {
      at_tc_fence_after();
      at_issue(tmem_base, dA, dB1, idesc1, AT_C / 16);
      at_umma_commit(&bar_s);
    }
    // ---------------- (2) softmax over the words ----------------
    at_mbar_wait(&bar_s, phase);
    at_tc_fence_after();
    float sc[LQ];
    {
      const float ss1 = og_exp2i(-kh);
      if (LQ == 32) {
        uint32_t sr[32];
        at_tmem_ld32(tmem_row, sr);
        at_tmem_ld_wait();
#pragma unroll
        for (int l = 0; l < LQ; ++l) sc[l] = __uint_as_float(sr[l]) * ss1 * ss2;
      } else {
        uint32_t sr[16], sr2[16];
        at_tmem_ld16(tmem_row, sr);
        at_tmem_ld16(tmem_row + 16, sr2);
        at_tmem_ld_wait();
#pragma unroll
        for (int l = 0; l < LQ; ++l) sc[l] = __uint_as_float(l < 16 ? sr[l & 15] : sr2[l & 15]) * ss1 * ss2;
      }
    }
    {
      // words beyond L and masked words leave the softmax (exp2(-inf) = 0)
      uint32_t dead = L >= 32 ? 0u : ~((1u << L) - 1u);
      if (mask) dead |= mbits[mrow];
      mrow += mstep;
      if (mrow >= (uint32_t)B) mrow -= (uint32_t)B;
      float mx = -INFINITY;
#pragma unroll
      for (int l = 0; l < LQ; ++l) {
        if ((dead >> l) & 1u) sc[l] = -INFINITY;
        mx = fmaxf(mx, sc[l]);
      }
      float sum = 0.f;
#pragma unroll
      for (int l = 0; l < LQ; ++l) {
        sc[l] = exp2f((sc[l] - mx) * 1.4426950408889634f);
        sum += sc[l];
      }
      const float inv = 1.f / sum;
      float* arow = attn + (long long)b * L * Q + (q0 + t);
#pragma unroll
      for (int l = 0; l < LQ; ++l) {
        sc[l] *= inv;
        if (l < L) *arow = sc[l];
        arow += Q;
      }
    }
    // probabilities * 2^13 -> the two A copies (row t, K = 32: four 16-byte chunks per copy; words >= LQ are zeros)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t yh[4], yl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l0 = ch * 8 + 2 * e;
        float ha = 0.f, la = 0.f, hb = 0.f, lb = 0.f;
        if (l0 < LQ) at_hilo(sc[l0 < LQ ? l0 : 0] * (float)(1 << AT_PSCALE), ha, la);
        if (l0 + 1 < LQ) at_hilo(sc[l0 + 1 < LQ ? l0 + 1 : 0] * (float)(1 << AT_PSCALE), hb, lb);
        yh[e] = at_cvt2(ha, hb);
        yl[e] = at_cvt2(la, lb);
      }
      const uint32_t addr = rowP + (uint32_t)(((ch ^ (t & 7)) & 7) << 4);
      at_sts128(addr, yh[0], yh[1], yh[2], yh[3]);
      at_sts128(addr + AT_A_BYTES, yl[0], yl[1], yl[2], yl[3]);
    }
    at_fence_proxy_async();
    at_tc_fence_before();
    __syncthreads();
    if (t == 0) {
      at_tc_fence_after();
      at_issue(tmem_base + 32, dA, dB2, idesc2, AT_LP / 16);
      at_umma_commit(&bar_o);
    }
    // ---------------- (3) context rows: TMEM -> shared staging -> coalesced stores ----------------
    at_mbar_wait(&bar_o, phase);
    at_tc_fence_after();
    // MMA 2 has completed: the A copies are free and take the staged context rows, 16 columns at a time
#pragma unroll
    for (int cb = 0; cb < AT_C; cb += 16) {
      uint32_t o[16];
      at_tmem_ld16(tmem_row + 32 + cb, o);
      at_tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        at_sts128(rowS + (cb + j) * 4, __float_as_uint(__uint_as_float(o[j]) * so), __float_as_uint(__uint_as_float(o[j + 1]) * so),
                  __float_as_uint(__uint_as_float(o[j + 2]) * so), __float_as_uint(__uint_as_float(o[j + 3]) * so));
    }
    at_tc_fence_before();
    __syncthreads();
    {
      float4* op = reinterpret_cast<float4*>(wc + ((long long)b * Q + q0) * AT_C) + warp * 384 + lane;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          op[g * 96 + j * 32] = at_lds128(sA + offS[j] + (uint32_t)(g * 8 * AT_STAGE_PITCH * 4));
    }
    __syncthreads();                                     // staging drained before the next trip overwrites the A copies
  }
  at_tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    at_tc_fence_after();
    at_tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward (g_attn = null: the attention maps are only visualised in training).  Per query q:
//   gP[l]   = sum_c g_wc[q][c] src[c][l]                          MMA 1:  Gw (128 x 48)   . B1           -> TMEM [0, 32)
//   gS[l]   = P[l] (gP[l] - sum_l' P[l'] gP[l'])                  registers (P read from the saved attention map)
//   g_h[c]  = sum_l gS[l] src[c][l]                               MMA 2:  gS (128 x 32)   . B2           -> TMEM [32, 80)
//   g_src[c][l] += sum_q g_wc[q][c] P[q][l] + h[q][c] gS[q][l]    MMA 3a: Gw^T . P, 3b: H^T . gS  (contraction over the
//       tile's 128 queries: the SAME shared-memory tiles read as MN-major operands, ref. conv_tc.cu make_desc_mn: a
//       slab = 128 query rows x 64 elements, k-step = 16 rows = 2048 bytes)                              -> TMEM [80, 112), [112, 144)
// The two g_src terms carry different operand scales, so each tile's pair of 48 x 32 results is read back, rescaled and
// summed in registers (thread c < 48 owns row c); one atomicAdd per (c, l) per CTA at the end.  M = 128 for MMA 3: the
// second 64-row slab of the A operand is whatever tile follows in shared memory -- rows 64..127 (and the pad rows
// 48..63) of the result are never read.
// Shared memory: X = {Gw, then H} hi / lo, Y = {P, then gS} hi / lo, B1, B2: 84 KB, two CTAs per SM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t saddr, uint32_t slab_pitch) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(slab_pitch >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D (+)= A^T-style contraction over the 128 rows of two MN-major tiles: 8 k-steps x (lo*hi, hi*lo, hi*hi)
__device__ __forceinline__ void at_issue_mn(uint32_t d_tmem, const uint64_t (&a)[2], const uint64_t (&b)[2], uint32_t idesc) {
  for (int k = 0; k < AT_Q / 16; ++k) {
    const uint64_t koff = (uint64_t)((k * 2048) >> 4);
    at_umma(d_tmem, a[1] + koff, b[0] + koff, idesc, k > 0 ? 1u : 0u);
    at_umma(d_tmem, a[0] + koff, b[1] + koff, idesc, 1u);
    at_umma(d_tmem, a[0] + koff, b[0] + koff, idesc, 1u);
  }
}
// two CTA-wide maxima at once
__device__ __forceinline__ void at_cta_max2(float& m0, float& m1, float (*scratch)[4]) {
  m0 = warp_max(m0);
  m1 = warp_max(m1);
  if ((threadIdx.x & 31) == 0) { scratch[0][threadIdx.x >> 5] = m0; scratch[1][threadIdx.x >> 5] = m1; }
  __syncthreads();
  m0 = fmaxf(fmaxf(scratch[0][0], scratch[0][1]), fmaxf(scratch[0][2], scratch[0][3]));
  m1 = fmaxf(fmaxf(scratch[1][0], scratch[1][1]), fmaxf(scratch[1][2], scratch[1][3]));
}

template <int LQ>
__global__ void __launch_bounds__(AT_Q, 2)
att_general_bwd_tc_kernel(const float* __restrict__ h, const float* __restrict__ src, const float* __restrict__ attn,
                          const float* __restrict__ g_wc, int B, int Q, int L, int nslots, float* __restrict__ g_h,
                          float* __restrict__ g_src) {
  extern __shared__ uint8_t at_smem_raw[];
  const uint32_t smem = (at_smem_u32(at_smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = smem;                            // Gw, then H (hi, lo); finally the staged g_h rows
  const uint32_t sY = smem + 2 * AT_A_BYTES;           // P, then gS (hi, lo)
  const uint32_t sB1 = sY + 2 * AT_A_BYTES;
  const uint32_t sB2 = sB1 + 2 * AT_B1_BYTES;
  __shared__ __align__(8) uint64_t bar1, bar2;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[4][4];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int b = blockIdx.x / nslots, slot = blockIdx.x - b * nslots;
  const int ntiles = Q / AT_Q;

  if (t == 0) {
    at_mbar_init(&bar1, 1);
    at_mbar_init(&bar2, 1);
    at_fence_barrier_init();
  }
  if (warp == 0) at_tmem_alloc(&tmem_base_smem, 256);
  uint32_t offA[3], offS[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int f = j * 32 + lane, r8 = (f * 43691) >> 19, c8 = f - r8 * 12;
    offA[j] = (uint32_t)(warp * 4 * 1024) + at_sw_off(r8, c8 * 4);
    offS[j] = (uint32_t)(((warp * 32 + r8) * AT_STAGE_PITCH + c8 * 4) * 4);
  }
  // word projections of this image -> B1[l][c], B2[c][l] (as in the forward kernel)
  int ksrc;
  {
    const float* sb = src + (long long)b * AT_C * L;
    const int l = lane, c0 = warp;
    float v[12];
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      v[j] = l < L ? __ldg(sb + (c0 + 4 * j) * L + l) : 0.f;
      m = fmaxf(m, fabsf(v[j]));
    }
    m = at_cta_max(m, red[0]);
    ksrc = og_scale_exp(__float_as_uint(m));
    const float s = og_exp2i(ksrc);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int c = c0 + 4 * j;
      float hi, lo;
      at_hilo(v[j] * s, hi, lo);
      const uint32_t pk = at_cvt2(hi, lo);
      const uint32_t o1 = at_sw_off(l, c), o2 = at_sw_off(c, l);
      at_sts16(sB1 + o1, pk & 0xFFFFu);
      at_sts16(sB1 + AT_B1_BYTES + o1, pk >> 16);
      at_sts16(sB2 + o2, pk & 0xFFFFu);
      at_sts16(sB2 + AT_B2_BYTES + o2, pk >> 16);
    }
  }
  at_tc_fence_before();
  __syncthreads();
  at_tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);

  uint64_t dX[2], dY[2], dB1[2], dB2[2], dXmn[2], dYmn[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    dX[i] = at_desc_sw128(sX + i * AT_A_BYTES);
    dY[i] = at_desc_sw128(sY + i * AT_A_BYTES);
    dB1[i] = at_desc_sw128(sB1 + i * AT_B1_BYTES);
    dB2[i] = at_desc_sw128(sB2 + i * AT_B2_BYTES);
    dXmn[i] = at_desc_mn(sX + i * AT_A_BYTES, AT_A_BYTES);
    dYmn[i] = at_desc_mn(sY + i * AT_A_BYTES, AT_A_BYTES);
  }
  const uint32_t idesc1 = at_idesc_f16(128, AT_LP), idesc2 = at_idesc_f16(128, AT_C);
  const uint32_t idesc3 = at_idesc_f16(128, AT_LP) | (1u << 15) | (1u << 16);      // both operands MN-major
  const uint32_t rowY = sY + (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
  const uint32_t rowS = sX + (uint32_t)(t * AT_STAGE_PITCH) * 4u;
  const float ss2 = og_exp2i(-ksrc);

  float gacc[LQ];                                      // g_src[c = t][l], summed over this CTA's tiles
#pragma unroll
  for (int l = 0; l < LQ; ++l) gacc[l] = 0.f;

  // g_wc rows and the saved probabilities of a tile are loaded one tile AHEAD (into the registers the current tile has
  // finished with); h is loaded at the top of its own tile (it is needed last)
  float4 gw[12];
  float pr[LQ];
  auto load_gw = [&](int tileq) {
    const float4* gp = reinterpret_cast<const float4*>(g_wc + ((long long)b * Q + (long long)tileq * AT_Q) * AT_C) + warp * 384 + lane;
#pragma unroll
    for (int i = 0; i < 12; ++i) gw[i] = __ldg(gp + i * 32);
  };
  auto load_pr = [&](int tileq) {
    const float* arow = attn + (long long)b * L * Q + ((long long)tileq * AT_Q + t);
#pragma unroll
    for (int l = 0; l < LQ; ++l) {
      pr[l] = l < L ? __ldg(arow) : 0.f;
      arow += Q;
    }
  };
  if (slot < ntiles) {
    load_gw(slot);
    load_pr(slot);
  }
  uint32_t phase = 0;
  for (int qt = slot; qt < ntiles; qt += nslots, phase ^= 1) {
    const int q0 = qt * AT_Q;
    const long long tile = ((long long)b * Q + q0) * AT_C;
    float4 hh[12];
    {
      const float4* hp = reinterpret_cast<const float4*>(h + tile) + warp * 384 + lane;
#pragma unroll
      for (int i = 0; i < 12; ++i) hh[i] = __ldg(hp + i * 32);
    }
    float prc[LQ];                                     // this tile's probabilities (pr is refilled below)
#pragma unroll
    for (int l = 0; l < LQ; ++l) prc[l] = pr[l];
    // ---- (1) Gw -> X, P -> Y ----
    float mg = 0.f, mh = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) { mg = at_amax4(mg, gw[i]); mh = at_amax4(mh, hh[i]); }
    at_cta_max2(mg, mh, red + 1);
    const int kg = og_scale_exp(__float_as_uint(mg)), kh = og_scale_exp(__float_as_uint(mh));
    {
      const float sg = og_exp2i(kg);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 x = gw[g * 3 + j];
          float h0, l0, h1, l1, h2, l2, h3, l3;
          at_hilo(x.x * sg, h0, l0); at_hilo(x.y * sg, h1, l1);
          at_hilo(x.z * sg, h2, l2); at_hilo(x.w * sg, h3, l3);
          const uint32_t a = sX + offA[j] + g * 1024;
          at_sts64(a, at_cvt2(h0, h1), at_cvt2(h2, h3));
          at_sts64(a + AT_A_BYTES, at_cvt2(l0, l1), at_cvt2(l2, l3));
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t yh[4], yl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l0 = ch * 8 + 2 * e;
        float ha = 0.f, la = 0.f, hb = 0.f, lb = 0.f;
        if (l0 < LQ) at_hilo(prc[l0 < LQ ? l0 : 0] * (float)(1 << AT_PSCALE), ha, la);
        if (l0 + 1 < LQ) at_hilo(prc[l0 + 1 < LQ ? l0 + 1 : 0] * (float)(1 << AT_PSCALE), hb, lb);
        yh[e] = at_cvt2(ha, hb);
        yl[e] = at_cvt2(la, lb);
      }
      const uint32_t addr = rowY + (uint32_t)(((ch ^ (t & 7)) & 7) << 4);
      at_sts128(addr, yh[0], yh[1], yh[2], yh[3]);
      at_sts128(addr + AT_A_BYTES, yl[0], yl[1], yl[2], yl[3]);
    }
    if (qt + nslots < ntiles) {                        // gw and pr are dead from here on: next tile's loads
      load_gw(qt + nslots);
      load_pr(qt + nslots);
    }
    at_fence_proxy_async();
    at_tc_fence_before();
    __syncthreads();
    if (t == 0) {
      at_tc_fence_after();
      at_issue(tmem_base, dX, dB1, idesc1, AT_C / 16);              // gP
      at_issue_mn(tmem_base + 80, dXmn, dYmn, idesc3);               // Gw^T . P
      at_umma_commit(&bar1);
    }
    // ---- (2) softmax backward in registers ----
    at_mbar_wait(&bar1, phase);
    at_tc_fence_after();
    float gs[LQ];
    {
      const float s1 = og_exp2i(-kg);
      if (LQ == 32) {
        uint32_t sr[32];
        at_tmem_ld32(tmem_row, sr);
        at_tmem_ld_wait();
#pragma unroll
        for (int l = 0; l < LQ; ++l) gs[l] = __uint_as_float(sr[l]) * s1 * ss2;
      } else {
        uint32_t sr[16], sr2[16];
        at_tmem_ld16(tmem_row, sr);
        at_tmem_ld16(tmem_row + 16, sr2);
        at_tmem_ld_wait();
#pragma unroll
        for (int l = 0; l < LQ; ++l) gs[l] = __uint_as_float(l < 16 ? sr[l & 15] : sr2[l & 15]) * s1 * ss2;
      }
      float dot = 0.f;
#pragma unroll
      for (int l = 0; l < LQ; ++l) dot = fmaf(prc[l], gs[l], dot);
      float mgs = 0.f;
#pragma unroll
      for (int l = 0; l < LQ; ++l) {
        gs[l] = prc[l] * (gs[l] - dot);
        mgs = fmaxf(mgs, fabsf(gs[l]));
      }
      // first g_src term of this tile (thread c = t < 48 reads row c): acc * 2^-(kg + 13)
      {
        uint32_t d3[32];
        at_tmem_ld32(tmem_row + 80, d3);
        at_tmem_ld_wait();
        const float s3a = og_exp2i(-kg), s3b = og_exp2i(-AT_PSCALE);    // two factors: their product may underflow
#pragma unroll
        for (int l = 0; l < LQ; ++l) gacc[l] = fmaf(__uint_as_float(d3[l]) * s3a, s3b, gacc[l]);
      }
      mgs = at_cta_max(mgs, red[3]);                   // (its __syncthreads also orders the TMEM reads above)
      const int ks = og_scale_exp(__float_as_uint(mgs));
      const float sgs = og_exp2i(ks), shh = og_exp2i(kh);
      // ---- (3) gS -> Y, H -> X (MMA 1 / 3a have completed: both regions are free) ----
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t yh[4], yl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int l0 = ch * 8 + 2 * e;
          float ha = 0.f, la = 0.f, hb = 0.f, lb = 0.f;
          if (l0 < LQ) at_hilo(gs[l0 < LQ ? l0 : 0] * sgs, ha, la);
          if (l0 + 1 < LQ) at_hilo(gs[l0 + 1 < LQ ? l0 + 1 : 0] * sgs, hb, lb);
          yh[e] = at_cvt2(ha, hb);
          yl[e] = at_cvt2(la, lb);
        }
        const uint32_t addr = rowY + (uint32_t)(((ch ^ (t & 7)) & 7) << 4);
        at_sts128(addr, yh[0], yh[1], yh[2], yh[3]);
        at_sts128(addr + AT_A_BYTES, yl[0], yl[1], yl[2], yl[3]);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 x = hh[g * 3 + j];
          float h0, l0, h1, l1, h2, l2, h3, l3;
          at_hilo(x.x * shh, h0, l0); at_hilo(x.y * shh, h1, l1);
          at_hilo(x.z * shh, h2, l2); at_hilo(x.w * shh, h3, l3);
          const uint32_t a = sX + offA[j] + g * 1024;
          at_sts64(a, at_cvt2(h0, h1), at_cvt2(h2, h3));
          at_sts64(a + AT_A_BYTES, at_cvt2(l0, l1), at_cvt2(l2, l3));
        }
      at_fence_proxy_async();
      at_tc_fence_before();
      __syncthreads();
      if (t == 0) {
        at_tc_fence_after();
        at_issue(tmem_base + 32, dY, dB2, idesc2, AT_LP / 16);       // g_h
        at_issue_mn(tmem_base + 112, dXmn, dYmn, idesc3);            // H^T . gS
        at_umma_commit(&bar2);
      }
      // ---- (4) g_h rows and the second g_src term ----
      at_mbar_wait(&bar2, phase);
      at_tc_fence_after();
      {
        uint32_t d3[32];
        at_tmem_ld32(tmem_row + 112, d3);
        at_tmem_ld_wait();
        const float s3a = og_exp2i(-kh), s3b = og_exp2i(-ks);
#pragma unroll
        for (int l = 0; l < LQ; ++l) gacc[l] = fmaf(__uint_as_float(d3[l]) * s3a, s3b, gacc[l]);
      }
      const float so = og_exp2i(-ks);
#pragma unroll
      for (int cb = 0; cb < AT_C; cb += 16) {
        uint32_t o[16];
        at_tmem_ld16(tmem_row + 32 + cb, o);
        at_tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          at_sts128(rowS + (cb + j) * 4, __float_as_uint(__uint_as_float(o[j]) * so * ss2),
                    __float_as_uint(__uint_as_float(o[j + 1]) * so * ss2), __float_as_uint(__uint_as_float(o[j + 2]) * so * ss2),
                    __float_as_uint(__uint_as_float(o[j + 3]) * so * ss2));
      }
    }
    at_tc_fence_before();
    __syncthreads();
    {
      float4* op = reinterpret_cast<float4*>(g_h + tile) + warp * 384 + lane;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          op[g * 96 + j * 32] = at_lds128(sX + offS[j] + (uint32_t)(g * 8 * AT_STAGE_PITCH * 4));
    }
    __syncthreads();
  }
  if (t < AT_C) {
    float* gp = g_src + ((long long)b * AT_C + t) * L;
#pragma unroll
    for (int l = 0; l < LQ; ++l)
      if (l < L) atomicAdd(gp + l, gacc[l]);
  }
  at_tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    at_tc_fence_after();
    at_tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// Returns 0 when the tensor-core kernel ran, -1 when the shape is outside its envelope (the caller then uses the SIMT
// kernels of attention.cu), else a cudaError_t.
extern "C" int og_att_general_fwd_tc(const float* h, const float* src, const unsigned char* mask, int B, int Q, int idf,
                                     int cs, int L, float* wc, float* attn, cudaStream_t stream) {
  if (idf != AT_C || cs != AT_C || L < 1 || L > AT_LP || Q % AT_Q != 0 || Q < AT_Q || B < 1 ||
      (long long)B * Q >= (1LL << 31) || B > AT_MAXB)
    return -1;
  const int ntiles = Q / AT_Q;
  static const int cps = getenv("OG_ATT_CPS") ? atoi(getenv("OG_ATT_CPS")) : 4;
  int nslots = (cps * 148) / B;        // four CTAs per SM, every CTA stays inside one image
  if (nslots < 1) nslots = 1;
  if (nslots > ntiles) nslots = ntiles;
  const size_t smem = AT_SMEM + 1024;
  static bool configured = false;
  if (!configured) {
    OG_CHECK(cudaFuncSetAttribute(att_general_fwd_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OG_CHECK(cudaFuncSetAttribute(att_general_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (L <= 20)
    att_general_fwd_tc_kernel<20><<<B * nslots, AT_Q, smem, stream>>>(h, src, mask, B, Q, L, nslots, wc, attn);
  else
    att_general_fwd_tc_kernel<32><<<B * nslots, AT_Q, smem, stream>>>(h, src, mask, B, Q, L, nslots, wc, attn);
  OG_RETURN_LAST_ERROR();
}

// Backward on the tensor cores (g_attn == null only).  g_src must be zeroed by the caller.  Same return convention.
extern "C" int og_att_general_bwd_tc(const float* h, const float* src, const float* attn, const float* g_wc, int B, int Q,
                                     int idf, int cs, int L, float* g_h, float* g_src, cudaStream_t stream) {
  if (idf != AT_C || cs != AT_C || L < 1 || L > AT_LP || Q % AT_Q != 0 || Q < AT_Q || B < 1) return -1;
  const int ntiles = Q / AT_Q;
  int nslots = (2 * 148) / B;          // two CTAs per SM
  if (nslots < 1) nslots = 1;
  if (nslots > ntiles) nslots = ntiles;
  const size_t smem = 4 * AT_A_BYTES + 2 * AT_B1_BYTES + 2 * AT_B2_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    OG_CHECK(cudaFuncSetAttribute(att_general_bwd_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OG_CHECK(cudaFuncSetAttribute(att_general_bwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (L <= 20)
    att_general_bwd_tc_kernel<20><<<B * nslots, AT_Q, smem, stream>>>(h, src, attn, g_wc, B, Q, L, nslots, g_h, g_src);
  else
    att_general_bwd_tc_kernel<32><<<B * nslots, AT_Q, smem, stream>>>(h, src, attn, g_wc, B, Q, L, nslots, g_h, g_src);
  OG_RETURN_LAST_ERROR();
}
