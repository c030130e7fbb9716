// Grid attention forward (GlobalAttentionGeneral / ATT_NET, ref: GlobalAttention.py:83-122) on the tensor cores:
// a persistent tcgen05 kernel for the large maps of the hot path (Q % 128 == 0, 48 channels, L <= 32 words).
//
//   scores[q][l] = sum_c h[q][c] * src[c][l]      (128 x 48) x (48 x 32)   -> TMEM columns [0, 32)
//   p = softmax_l(scores) (caption mask quirk: GlobalAttention.py:108)       -> attn[b][l][q]
//   wc[q][c]     = sum_l p[q][l] * src[c][l]      (128 x 32) x (32 x 48)   -> TMEM columns [32, 80)
//
// Why tensor cores for a bandwidth-bound op: the two tiny GEMMs cost 1920 fp32 FMAs per query; on the CUDA cores that
// instruction stream (plus the shared-memory operand traffic it needs) capped the kernel at 0.37 of the HBM
// bandwidth.  Here the FMA work is 30 small MMAs per 128 queries and the threads only move and convert data.
//
// fp32 parity: every operand is split into THREE bf16 terms (a = a1 + a2 + a3, 24 mantissa bits, bf16 keeps the fp32
// exponent so no scaling pass is needed) and six products (a1b1 + a1b2 + a2b1 + a2b2 + a1b3 + a3b1) accumulate in the
// fp32 TMEM accumulator: the dropped terms are below 2^-24 of the product, i.e. fp32 rounding level.
//
// One CTA = 128 threads = 128 queries per trip; thread t owns query row t (= TMEM lane t).  Per trip:
//   (1) the h tile (24.5 KB, contiguous) arrives through coalesced 16-byte loads issued one trip AHEAD (registers),
//       is split and stored into three K-major 128B-swizzled A tiles;                      -> MMA 1 (18 instructions)
//   (2) tcgen05.ld of the thread's score row, softmax in registers, coalesced attn stores (lanes = adjacent queries),
//       the probabilities are split into the same three tiles (K = 32);                    -> MMA 2 (12 instructions)
//   (3) tcgen05.ld of the context row, staged through shared memory, written with coalesced 16-byte stores.
// CTAs are persistent over the query tiles of ONE image (the word-projection operand tiles are built once).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int AT_Q = 128;          // queries per tile = threads per CTA
constexpr int AT_C = 48;           // channels (idf == row stride)
constexpr int AT_LP = 32;          // words padded to the MMA N / K extent
constexpr uint32_t AT_A_BYTES = 128 * 128;            // one bf16 copy of the A tile: 128 rows x 128-byte swizzle rows
constexpr uint32_t AT_B1_BYTES = AT_LP * 128;         // src as [l][c]  (N = 32 rows, K = 48)
constexpr uint32_t AT_B2_BYTES = AT_C * 128;          // src as [c][l]  (N = 48 rows, K = 32)
constexpr uint32_t AT_SMEM = 3 * AT_A_BYTES + 3 * AT_B1_BYTES + 3 * AT_B2_BYTES;
constexpr int AT_STAGE_PITCH = AT_C + 4;              // floats per staged context row (conflict-free 16-byte accesses)

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
// D = f32, A = B = bf16 (format code 1), both K-major
__device__ __forceinline__ uint32_t at_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row r, 16-bit column k) in a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t at_sw_off(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// a = a1 + a2 + a3 EXACTLY: a1 = the top 16 bits of the fp32 pattern (a bf16 by truncation), a2 = the top 16 bits of the
// exact residual, a3 = what is left (<= 8 significant bits: a bf16 as it stands).  Pure LOP3 / FADD work -- the rounding
// conversions (F2F) this replaces are quarter-rate instructions and dominated the first version of this kernel.
struct Split3 { uint32_t w1, w2, w3; };      // fp32 bit patterns whose HIGH halves are the three bf16 terms
__device__ __forceinline__ Split3 at_split3(float v) {
  Split3 s;
  s.w1 = __float_as_uint(v) & 0xFFFF0000u;
  const float r1 = v - __uint_as_float(s.w1);
  s.w2 = __float_as_uint(r1) & 0xFFFF0000u;
  s.w3 = __float_as_uint(r1 - __uint_as_float(s.w2));
  return s;
}
// {high half of lo_word, high half of hi_word} -> one 32-bit pair of bf16 (lo_word's term at the lower address)
__device__ __forceinline__ uint32_t at_pack_hi(uint32_t lo_word, uint32_t hi_word) {
  return __byte_perm(lo_word, hi_word, 0x7632);
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

// six products per k-step, small terms first: a3b1 a1b3 a2b2 a2b1 a1b2 a1b1
__device__ __forceinline__ void at_issue(uint32_t d_tmem, const uint64_t (&a)[3], const uint64_t (&b)[3], uint32_t idesc,
                                         int ksteps) {
  for (int k = 0; k < ksteps; ++k) {
    const uint64_t koff = (uint64_t)((k * 32) >> 4);
    at_umma(d_tmem, a[2] + koff, b[0] + koff, idesc, k > 0 ? 1u : 0u);
    at_umma(d_tmem, a[0] + koff, b[2] + koff, idesc, 1u);
    at_umma(d_tmem, a[1] + koff, b[1] + koff, idesc, 1u);
    at_umma(d_tmem, a[1] + koff, b[0] + koff, idesc, 1u);
    at_umma(d_tmem, a[0] + koff, b[1] + koff, idesc, 1u);
    at_umma(d_tmem, a[0] + koff, b[0] + koff, idesc, 1u);
  }
}

// LQ: words handled in registers (20 covers the 12..18-word captions of the hot path; 32 = the MMA extent)
template <int LQ>
__global__ void __launch_bounds__(AT_Q, 2)
att_general_fwd_tc_kernel(const float* __restrict__ h, const float* __restrict__ src,
                          const unsigned char* __restrict__ mask, int B, int Q, int L, int nslots,
                          float* __restrict__ wc, float* __restrict__ attn) {
  extern __shared__ uint8_t at_smem_raw[];
  const uint32_t smem = (at_smem_u32(at_smem_raw) + 1023u) & ~1023u;   // shared-window address of the operand tiles
  const uint32_t sA = smem;                            // 3 copies; reused for the probabilities and the staged context
  const uint32_t sB1 = smem + 3 * AT_A_BYTES;
  const uint32_t sB2 = sB1 + 3 * AT_B1_BYTES;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float ssrc[AT_C * AT_LP];                 // src[b] as [c][l], zero beyond L

  const int t = threadIdx.x, warp = t >> 5;
  const int b = blockIdx.x / nslots, slot = blockIdx.x - b * nslots;
  const int ntiles = Q / AT_Q;

  if (t == 0) {
    at_mbar_init(&bar_s, 1);
    at_mbar_init(&bar_o, 1);
    at_fence_barrier_init();
  }
  if (warp == 0) at_tmem_alloc(&tmem_base_smem, 128);
  // prefetch the first tile: 1536 float4 per tile, thread t takes t, t + 128, ...
  float4 hv[12];
  int qt = slot;
  {
    const float4* hp = reinterpret_cast<const float4*>(h + ((long long)b * Q + (long long)qt * AT_Q) * AT_C);
#pragma unroll
    for (int j = 0; j < 12; ++j) hv[j] = __ldg(hp + t + j * AT_Q);
  }
  // word projections of this image -> shared (independent coalesced loads), then the two operand layouts
  {
    const float* sb = src + (long long)b * AT_C * L;
    const int n = AT_C * L;
    float v[(AT_C * AT_LP + AT_Q - 1) / AT_Q];
#pragma unroll
    for (int j = 0; j < (AT_C * AT_LP + AT_Q - 1) / AT_Q; ++j) {
      const int i = t + j * AT_Q;
      v[j] = i < n ? __ldg(sb + i) : 0.f;
    }
    for (int i = t; i < AT_C * AT_LP; i += AT_Q) ssrc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < (AT_C * AT_LP + AT_Q - 1) / AT_Q; ++j) {
      const int i = t + j * AT_Q;
      if (i < n) {
        const int c = i / L;
        ssrc[c * AT_LP + (i - c * L)] = v[j];
      }
    }
    __syncthreads();
    for (int i = t; i < AT_LP * AT_C; i += AT_Q) {         // B1[l][c]: 32 rows x 48 k
      const int l = i / AT_C, c = i - l * AT_C;
      const Split3 x = at_split3(ssrc[c * AT_LP + l]);
      const uint32_t off = at_sw_off(l, c);
      at_sts16(sB1 + off, x.w1 >> 16);
      at_sts16(sB1 + AT_B1_BYTES + off, x.w2 >> 16);
      at_sts16(sB1 + 2 * AT_B1_BYTES + off, x.w3 >> 16);
    }
    for (int i = t; i < AT_C * AT_LP; i += AT_Q) {         // B2[c][l]: 48 rows x 32 k
      const int c = i / AT_LP, l = i - c * AT_LP;
      const Split3 x = at_split3(ssrc[i]);
      const uint32_t off = at_sw_off(c, l);
      at_sts16(sB2 + off, x.w1 >> 16);
      at_sts16(sB2 + AT_B2_BYTES + off, x.w2 >> 16);
      at_sts16(sB2 + 2 * AT_B2_BYTES + off, x.w3 >> 16);
    }
  }
  at_tc_fence_before();
  __syncthreads();
  at_tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);

  uint64_t dA[3], dB1[3], dB2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dA[i] = at_desc_sw128(sA + i * AT_A_BYTES);
    dB1[i] = at_desc_sw128(sB1 + i * AT_B1_BYTES);
    dB2[i] = at_desc_sw128(sB2 + i * AT_B2_BYTES);
  }
  const uint32_t idesc1 = at_idesc_bf16(128, AT_LP), idesc2 = at_idesc_bf16(128, AT_C);

  // tile-invariant shared-memory offsets of this thread's 12 float4 slots (A tile: swizzled; staging: padded rows)
  uint32_t offA[12], offS[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    const int idx = t + j * AT_Q;
    const int r = idx / 12, cq = idx - r * 12;
    offA[j] = at_sw_off(r, cq * 4);
    offS[j] = (uint32_t)(r * AT_STAGE_PITCH + cq * 4) * 4u;
  }
  const uint32_t rowP = sA + at_sw_off(t, 0) - (uint32_t)((t & 7) << 4);   // row base; chunk ch sits at ((ch ^ (t & 7)) << 4)
  const uint32_t rowS = sA + (uint32_t)(t * AT_STAGE_PITCH) * 4u;

  uint32_t phase = 0;
  for (; qt < ntiles; qt += nslots, phase ^= 1) {
    const int q0 = qt * AT_Q;
    // ---------------- (1) split the h tile into the three A copies ----------------
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const Split3 x0 = at_split3(hv[j].x), x1 = at_split3(hv[j].y), x2 = at_split3(hv[j].z), x3 = at_split3(hv[j].w);
      at_sts64(sA + offA[j], at_pack_hi(x0.w1, x1.w1), at_pack_hi(x2.w1, x3.w1));
      at_sts64(sA + AT_A_BYTES + offA[j], at_pack_hi(x0.w2, x1.w2), at_pack_hi(x2.w2, x3.w2));
      at_sts64(sA + 2 * AT_A_BYTES + offA[j], at_pack_hi(x0.w3, x1.w3), at_pack_hi(x2.w3, x3.w3));
    }
    // the next tile's loads fly during the rest of this trip
    if (qt + nslots < ntiles) {
      const float4* hp = reinterpret_cast<const float4*>(h + ((long long)b * Q + (long long)(qt + nslots) * AT_Q) * AT_C);
#pragma unroll
      for (int j = 0; j < 12; ++j) hv[j] = __ldg(hp + t + j * AT_Q);
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
    if (LQ == 32) {
      uint32_t sr[32];
      at_tmem_ld32(tmem_row, sr);
      at_tmem_ld_wait();
#pragma unroll
      for (int l = 0; l < LQ; ++l) sc[l] = __uint_as_float(sr[l]);
    } else {
      uint32_t sr[16], sr2[16];
      at_tmem_ld16(tmem_row, sr);
      at_tmem_ld16(tmem_row + 16, sr2);
      at_tmem_ld_wait();
#pragma unroll
      for (int l = 0; l < LQ; ++l) sc[l] = __uint_as_float(l < 16 ? sr[l & 15] : sr2[l & 15]);
    }
    const int q = q0 + t;
    if (mask) {
      const unsigned char* mr = mask + (((unsigned)b * (unsigned)Q + (unsigned)q) % (unsigned)B) * L;   // B * Q < 2^31
#pragma unroll
      for (int l = 0; l < LQ; ++l)
        if (l < L && mr[l]) sc[l] = -INFINITY;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < LQ; ++l)
      if (l < L) mx = fmaxf(mx, sc[l]);
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < LQ; ++l) {
      sc[l] = l < L ? expf(sc[l] - mx) : 0.f;
      sum += sc[l];
    }
    const float inv = 1.f / sum;
    float* arow = attn + (long long)b * L * Q + q;
#pragma unroll
    for (int l = 0; l < LQ; ++l) {
      sc[l] *= inv;
      if (l < L) arow[(long long)l * Q] = sc[l];
    }
    // probabilities -> the three A copies (row t, K = 32: four 16-byte chunks per copy; words >= LQ are zeros)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t y1[4], y2[4], y3[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l0 = ch * 8 + 2 * e;
        Split3 u = {0u, 0u, 0u}, w = {0u, 0u, 0u};
        if (l0 < LQ) u = at_split3(sc[l0 < LQ ? l0 : 0]);
        if (l0 + 1 < LQ) w = at_split3(sc[l0 + 1 < LQ ? l0 + 1 : 0]);
        y1[e] = at_pack_hi(u.w1, w.w1); y2[e] = at_pack_hi(u.w2, w.w2); y3[e] = at_pack_hi(u.w3, w.w3);
      }
      const uint32_t addr = rowP + (uint32_t)(((ch ^ (t & 7)) & 7) << 4);
      at_sts128(addr, y1[0], y1[1], y1[2], y1[3]);
      at_sts128(addr + AT_A_BYTES, y2[0], y2[1], y2[2], y2[3]);
      at_sts128(addr + 2 * AT_A_BYTES, y3[0], y3[1], y3[2], y3[3]);
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
    uint32_t o0[32], o1[16];
    at_tmem_ld32(tmem_row + 32, o0);
    at_tmem_ld16(tmem_row + 64, o1);
    at_tmem_ld_wait();
    // MMA 2 has completed: the A copies are free and take the staged context rows
#pragma unroll
    for (int j = 0; j < 32; j += 4) at_sts128(rowS + j * 4, o0[j], o0[j + 1], o0[j + 2], o0[j + 3]);
#pragma unroll
    for (int j = 0; j < 16; j += 4) at_sts128(rowS + (32 + j) * 4, o1[j], o1[j + 1], o1[j + 2], o1[j + 3]);
    at_tc_fence_before();
    __syncthreads();
    {
      float4* op = reinterpret_cast<float4*>(wc + ((long long)b * Q + q0) * AT_C);
#pragma unroll
      for (int j = 0; j < 12; ++j) op[t + j * AT_Q] = at_lds128(sA + offS[j]);
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

}  // namespace

// Returns 0 when the tensor-core kernel ran, -1 when the shape is outside its envelope (the caller then uses the SIMT
// kernels of attention.cu), else a cudaError_t.
extern "C" int og_att_general_fwd_tc(const float* h, const float* src, const unsigned char* mask, int B, int Q, int idf,
                                     int cs, int L, float* wc, float* attn, cudaStream_t stream) {
  if (idf != AT_C || cs != AT_C || L < 1 || L > AT_LP || Q % AT_Q != 0 || Q < AT_Q || B < 1 ||
      (long long)B * Q >= (1LL << 31))
    return -1;
  const int ntiles = Q / AT_Q;
  int nslots = (2 * 148) / B;          // two CTAs per SM, every CTA stays inside one image
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
