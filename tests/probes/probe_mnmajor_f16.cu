// Hardware probe, fp16 variant (not part of the product): does tcgen05.mma kind::f16 accept MN-major operands (channel index
// contiguous, reduction index = smem row) and with which TMA swizzle / descriptor layout?  One config per process.
//   usage: probe_mnmajor <tma_swizzle 3|4> <layout_type 1|2> <lbo_bytes> <sbo_bytes> <which 1=A 2=B 3=both> [kstep_bytes]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#include <cuda_fp16.h>
constexpr int M = 128, N = 64, K = 64;
constexpr int SLABW = 64;   // channels per 128-byte row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}

struct Cfg {
  int type, lbo, sbo, which, kstep;
};

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Cfg c, float* d_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* sA = smem;              // 16 KB either way
  uint8_t* sB = smem + 16384;      // 8 KB
  const int warp = threadIdx.x >> 5;
  const bool tA = c.which & 1, tB = c.which & 2;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (M + N) * K * 2);
    if (tA) {   // global A[k][m]: slabs of 32 m, each [K rows][128 B]
      for (int s = 0; s < M / SLABW; ++s) tma_load_2d(sA + s * K * 128, &map_a, &full_bar, s * SLABW, 0);
    } else {    // global A[m][k]: [M rows][128 B]
      tma_load_2d(sA, &map_a, &full_bar, 0, 0);
    }
    if (tB) {
      for (int s = 0; s < N / SLABW; ++s) tma_load_2d(sB + s * K * 128, &map_b, &full_bar, s * SLABW, 0);
    } else {
      tma_load_2d(sB, &map_b, &full_bar, 0, 0);
    }
    mbar_wait(&full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tA) idesc |= 1u << 15;
    if (tB) idesc |= 1u << 16;
    for (int k = 0; k < K / 16; ++k) {
      uint64_t da = tA ? make_desc(smem_u32(sA) + k * c.kstep, c.lbo, c.sbo, c.type)
                       : make_desc(smem_u32(sA) + k * 32, 16, 1024, 2);
      uint64_t db = tB ? make_desc(smem_u32(sB) + k * c.kstep, c.lbo, c.sbo, c.type)
                       : make_desc(smem_u32(sB) + k * 32, 16, 1024, 2);
      umma_tf32(tmem, da, db, idesc, k > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
  }
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = threadIdx.x;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
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
    for (int j = 0; j < 32; ++j) d_out[row * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  int tswz = atoi(argv[1]);
  Cfg c;
  c.type = atoi(argv[2]);
  c.lbo = atoi(argv[3]);
  c.sbo = atoi(argv[4]);
  c.which = atoi(argv[5]);
  c.kstep = argc > 6 ? atoi(argv[6]) : 2048;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return 3;
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const bool tA = c.which & 1, tB = c.which & 2;
  static float hA[M * K], hB[N * K], hD[M * N], ref[M * N];
  static __half hAg[M * K], hBg[N * K];
  srand(1);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) hA[m * K + k] = (float)(rand() % 9 - 4);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hB[n * K + k] = (float)(rand() % 9 - 4);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += hA[m * K + k] * hB[n * K + k];
      ref[m * N + n] = s;
    }
  // global layouts
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) hAg[tA ? k * M + m : m * K + k] = __float2half(hA[m * K + k]);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hBg[tB ? k * N + n : n * K + k] = __float2half(hB[n * K + k]);
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, sizeof(hAg));
  cudaMalloc(&dB, sizeof(hBg));
  cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hAg, sizeof(hAg), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hBg, sizeof(hBg), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, sizeof(hD));
  CUtensorMap ma, mb;
  auto mk = [&](CUtensorMap* mp, __half* base, bool t, int R) -> int {
    cuuint64_t dims[2], str[1];
    cuuint32_t box[2], es[2] = {1, 1};
    CUtensorMapSwizzle sw;
    if (t) {  // [K][R], R contiguous
      dims[0] = R; dims[1] = K; str[0] = (cuuint64_t)R * 2; box[0] = SLABW; box[1] = K;
      sw = (CUtensorMapSwizzle)tswz;
    } else {  // [R][K]
      dims[0] = K; dims[1] = R; str[0] = (cuuint64_t)K * 2; box[0] = 64; box[1] = R;
      sw = CU_TENSOR_MAP_SWIZZLE_128B;
    }
    return (int)enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  int r1 = mk(&ma, dA, tA, M), r2 = mk(&mb, dB, tB, N);
  if (r1 || r2) { printf("map error %d %d\n", r1, r2); return 4; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(ma, mb, c, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("tswz=%d type=%d lbo=%d sbo=%d which=%d kstep=%d : CUDA ERROR %s\n", tswz, c.type, c.lbo, c.sbo, c.which, c.kstep, cudaGetErrorString(e)); return 5; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  double maxerr = 0; int zeros = 0, nan = 0, exact = 0;
  for (int i = 0; i < M * N; ++i) {
    if (hD[i] != hD[i]) { nan++; continue; }
    if (hD[i] == 0.f) zeros++;
    if (hD[i] == ref[i]) exact++;
    double d = fabs((double)hD[i] - ref[i]);
    if (d > maxerr) maxerr = d;
  }
  printf("tswz=%d type=%d lbo=%d sbo=%d which=%d kstep=%d : maxerr=%g exact=%d/%d zeros=%d nan=%d %s\n", tswz, c.type, c.lbo, c.sbo,
         c.which, c.kstep, maxerr, exact, M * N, zeros, nan, (exact == M * N) ? "MATCH" : "");
  return 0;
}
