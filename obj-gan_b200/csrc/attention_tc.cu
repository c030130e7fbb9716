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

// a = a1 + a2 + a3 (three bf16 terms, 24 mantissa bits); returns the raw 16-bit patterns
__device__ __forceinline__ void at_split3(float v, unsigned short& a1, unsigned short& a2, unsigned short& a3) {
  const __nv_bfloat16 h1 = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(h1);
  const __nv_bfloat16 h2 = __float2bfloat16_rn(r);
  r -= __bfloat162float(h2);
  const __nv_bfloat16 h3 = __float2bfloat16_rn(r);
  a1 = __bfloat16_as_ushort(h1); a2 = __bfloat16_as_ushort(h2); a3 = __bfloat16_as_ushort(h3);
}
__device__ __forceinline__ uint32_t at_pack(unsigned short lo, unsigned short hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }

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

__global__ void __launch_bounds__(AT_Q, 2)
att_general_fwd_tc_kernel(const float* __restrict__ h, const float* __restrict__ src,
                          const unsigned char* __restrict__ mask, int B, int Q, int L, int nslots,
                          float* __restrict__ wc, float* __restrict__ attn) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                  // 3 copies; reused for the probabilities and the staged context
  uint8_t* sB1 = smem + 3 * AT_A_BYTES;
  uint8_t* sB2 = sB1 + 3 * AT_B1_BYTES;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_base_smem;

  const int t = threadIdx.x, warp = t >> 5;
  const int b = blockIdx.x / nslots, slot = blockIdx.x - b * nslots;
  const int ntiles = Q / AT_Q;

  if (t == 0) {
    at_mbar_init(&bar_s, 1);
    at_mbar_init(&bar_o, 1);
    at_fence_barrier_init();
  }
  if (warp == 0) at_tmem_alloc(&tmem_base_smem, 128);
  // word-projection operand tiles of this image: B1[l][c] and B2[c][l], zero beyond L
  {
    const float* sb = src + (long long)b * AT_C * L;
    for (int i = t; i < AT_LP * 64; i += AT_Q) {            // B1: 32 rows x 64 k (k >= 48 never read; zero anyway)
      const int l = i >> 6, c = i & 63;
      const float v = (l < L && c < AT_C) ? __ldg(sb + c * L + l) : 0.f;
      unsigned short a1, a2, a3;
      at_split3(v, a1, a2, a3);
      const uint32_t off = at_sw_off(l, c);
      *reinterpret_cast<unsigned short*>(sB1 + off) = a1;
      *reinterpret_cast<unsigned short*>(sB1 + AT_B1_BYTES + off) = a2;
      *reinterpret_cast<unsigned short*>(sB1 + 2 * AT_B1_BYTES + off) = a3;
    }
    for (int i = t; i < AT_C * 64; i += AT_Q) {             // B2: 48 rows x 64 k (k >= 32 never read)
      const int c = i >> 6, l = i & 63;
      const float v = l < L ? __ldg(sb + c * L + l) : 0.f;
      unsigned short a1, a2, a3;
      at_split3(v, a1, a2, a3);
      const uint32_t off = at_sw_off(c, l);
      *reinterpret_cast<unsigned short*>(sB2 + off) = a1;
      *reinterpret_cast<unsigned short*>(sB2 + AT_B2_BYTES + off) = a2;
      *reinterpret_cast<unsigned short*>(sB2 + 2 * AT_B2_BYTES + off) = a3;
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
    dA[i] = at_desc_sw128(at_smem_u32(sA + i * AT_A_BYTES));
    dB1[i] = at_desc_sw128(at_smem_u32(sB1 + i * AT_B1_BYTES));
    dB2[i] = at_desc_sw128(at_smem_u32(sB2 + i * AT_B2_BYTES));
  }
  const uint32_t idesc1 = at_idesc_bf16(128, AT_LP), idesc2 = at_idesc_bf16(128, AT_C);

  // prefetch the first tile: 1536 float4 per tile, thread t takes t, t + 128, ...
  float4 hv[12];
  int qt = slot;
  if (qt < ntiles) {
    const float4* hp = reinterpret_cast<const float4*>(h + ((long long)b * Q + (long long)qt * AT_Q) * AT_C);
#pragma unroll
    for (int j = 0; j < 12; ++j) hv[j] = __ldg(hp + t + j * AT_Q);
  }
  uint32_t phase = 0;
  for (; qt < ntiles; qt += nslots, phase ^= 1) {
    const int q0 = qt * AT_Q;
    // ---------------- (1) split the h tile into the three A copies ----------------
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int idx = t + j * AT_Q;
      const int r = idx / 12, cq = idx - r * 12;
      unsigned short x1[4], x2[4], x3[4];
      at_split3(hv[j].x, x1[0], x2[0], x3[0]);
      at_split3(hv[j].y, x1[1], x2[1], x3[1]);
      at_split3(hv[j].z, x1[2], x2[2], x3[2]);
      at_split3(hv[j].w, x1[3], x2[3], x3[3]);
      const uint32_t off = at_sw_off(r, cq * 4);
      *reinterpret_cast<uint2*>(sA + off) = make_uint2(at_pack(x1[0], x1[1]), at_pack(x1[2], x1[3]));
      *reinterpret_cast<uint2*>(sA + AT_A_BYTES + off) = make_uint2(at_pack(x2[0], x2[1]), at_pack(x2[2], x2[3]));
      *reinterpret_cast<uint2*>(sA + 2 * AT_A_BYTES + off) = make_uint2(at_pack(x3[0], x3[1]), at_pack(x3[2], x3[3]));
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
    uint32_t sr[32];
    at_tmem_ld32(tmem_row, sr);
    at_tmem_ld_wait();
    float sc[AT_LP];
#pragma unroll
    for (int l = 0; l < AT_LP; ++l) sc[l] = __uint_as_float(sr[l]);
    const int q = q0 + t;
    if (mask) {
      const unsigned char* mr = mask + (((long long)b * Q + q) % B) * L;
#pragma unroll
      for (int l = 0; l < AT_LP; ++l)
        if (l < L && mr[l]) sc[l] = -INFINITY;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < AT_LP; ++l)
      if (l < L) mx = fmaxf(mx, sc[l]);
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < AT_LP; ++l) {
      sc[l] = l < L ? expf(sc[l] - mx) : 0.f;
      sum += sc[l];
    }
    const float inv = 1.f / sum;
    float* arow = attn + (long long)b * L * Q + q;
#pragma unroll
    for (int l = 0; l < AT_LP; ++l) {
      sc[l] *= inv;
      if (l < L) arow[(long long)l * Q] = sc[l];
    }
    // probabilities -> the three A copies (row t, K = 32: four 16-byte chunks per copy)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      unsigned short y1[8], y2[8], y3[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) at_split3(sc[ch * 8 + e], y1[e], y2[e], y3[e]);
      const uint32_t off = at_sw_off(t, ch * 8);
      *reinterpret_cast<uint4*>(sA + off) =
          make_uint4(at_pack(y1[0], y1[1]), at_pack(y1[2], y1[3]), at_pack(y1[4], y1[5]), at_pack(y1[6], y1[7]));
      *reinterpret_cast<uint4*>(sA + AT_A_BYTES + off) =
          make_uint4(at_pack(y2[0], y2[1]), at_pack(y2[2], y2[3]), at_pack(y2[4], y2[5]), at_pack(y2[6], y2[7]));
      *reinterpret_cast<uint4*>(sA + 2 * AT_A_BYTES + off) =
          make_uint4(at_pack(y3[0], y3[1]), at_pack(y3[2], y3[3]), at_pack(y3[4], y3[5]), at_pack(y3[6], y3[7]));
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
    float* stage = reinterpret_cast<float*>(sA);        // MMA 2 has completed: the A copies are free
    {
      float* srow = stage + t * AT_STAGE_PITCH;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(srow + j) = make_uint4(o0[j], o0[j + 1], o0[j + 2], o0[j + 3]);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<uint4*>(srow + 32 + j) = make_uint4(o1[j], o1[j + 1], o1[j + 2], o1[j + 3]);
    }
    at_tc_fence_before();
    __syncthreads();
    {
      float4* op = reinterpret_cast<float4*>(wc + ((long long)b * Q + q0) * AT_C);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int idx = t + j * AT_Q;
        const int r = idx / 12, cq = idx - r * 12;
        op[idx] = *reinterpret_cast<const float4*>(stage + r * AT_STAGE_PITCH + cq * 4);
      }
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
  if (idf != AT_C || cs != AT_C || L < 1 || L > AT_LP || Q % AT_Q != 0 || Q < AT_Q || B < 1) return -1;
  const int ntiles = Q / AT_Q;
  int nslots = (2 * 148) / B;          // two CTAs per SM, every CTA stays inside one image
  if (nslots < 1) nslots = 1;
  if (nslots > ntiles) nslots = ntiles;
  const size_t smem = AT_SMEM + 1024;
  static bool configured = false;
  if (!configured) {
    OG_CHECK(cudaFuncSetAttribute(att_general_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  att_general_fwd_tc_kernel<<<B * nslots, AT_Q, smem, stream>>>(h, src, mask, B, Q, L, nslots, wc, attn);
  OG_RETURN_LAST_ERROR();
}
