// DAMSM word-region matching for ALL (image, caption) pairs of a batch in one launch
// (ref: miscc/losses.py:74-127 `words_loss`, which calls GlobalAttention.py:32-70 `func_attention` once per caption
// inside a Python loop, B times, each time on the caption tiled B x).
//
//   pair p = (image b, caption i):   n = cap_len[i] words,  q = words[i][:, :n]  (ndf x n),  ctx = img[b]  (ndf x S)
//     scores[s][l] = sum_c ctx[c][s] q[c][l];   P = softmax_l(scores);   P2[l][s] = softmax_s(gamma1 * P[s][l])
//     wc[c][l] = sum_s ctx[c][s] P2[l][s]                                        (func_attention)
//     cos[l]   = <q[:, l], wc[:, l]> / max(|q[:, l]| |wc[:, l]|, eps)            (losses.py:13-19, 101-108)
//     sim[b][i] = log sum_l exp(gamma2 * cos[l])                                 (losses.py:112-115, no max shift)
//
// One CTA per pair.  The image features (S x ndf = 296 KB per image at 17x17x256) are read straight from L2 (B
// images = 4.7 MB at B = 16 stay resident), coalesced along the region axis; the caption matrix and the S x n
// probability matrix live in shared memory.  Outputs: sim [B][NC], the weighted contexts wc [B*NC][ndf][T] (kept for
// the backward pass) and the attention maps attn [B*NC][T][S].
//
// Backward (gradient w.r.t. the image features only: the caption embeddings are constants of the generator update,
// ref: trainer.py:369): per pair, g_sim -> g_cos -> g_wc in shared memory, then the func_attention adjoint; the B
// captions of an image accumulate into g_ctx[b] with fp32 atomics.
#include "common.cuh"

namespace {

constexpr int WP_THREADS = 160;   // 5 warps; each thread owns up to two regions (S <= 320)

template <int LM>
__global__ void __launch_bounds__(WP_THREADS) words_pairs_fwd_kernel(
    const float* __restrict__ words, const float* __restrict__ ctx, const long long* __restrict__ lens, int NC, int ndf,
    int T, int S, float gamma1, float gamma2, float eps, float* __restrict__ wc, float* __restrict__ attn,
    float* __restrict__ sim) {
  extern __shared__ float smem[];
  const int p = blockIdx.x, b = p / NC, i = p - b * NC;
  const int Lq = min((int)lens[i], T);
  const int LP = Lq | 1;                 // odd row pitch of the probability matrix: conflict-free column walks
  float* sq = smem;                      // [ndf][Lq]
  float* sp = smem + ndf * T;            // [S][LP]
  float* red = sp + S * (T | 1);         // [5 warps][3][LM]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* qb = words + (long long)i * ndf * T;
  const float* cb = ctx + (long long)b * ndf * S;
  for (int k = t; k < ndf * Lq; k += WP_THREADS) {
    const int c = k / Lq, l = k - c * Lq;
    sq[k] = qb[c * T + l];
  }
  __syncthreads();
  // phase 1+2: scores of this thread's two regions against every word, softmax over the words
  {
    const int s0 = t, s1 = t + WP_THREADS;
    const bool v0 = s0 < S, v1 = s1 < S;
    float a0[LM], a1[LM];
#pragma unroll
    for (int l = 0; l < LM; ++l) a0[l] = a1[l] = 0.f;
    for (int c = 0; c < ndf; ++c) {
      const float c0 = v0 ? __ldg(cb + (long long)c * S + s0) : 0.f;
      const float c1 = v1 ? __ldg(cb + (long long)c * S + s1) : 0.f;
      const float* qr = sq + c * Lq;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          const float q = qr[l];
          a0[l] = fmaf(c0, q, a0[l]);
          a1[l] = fmaf(c1, q, a1[l]);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* a = h ? a1 : a0;
      const int s = h ? s1 : s0;
      if (!(h ? v1 : v0)) continue;
      float mx = -INFINITY;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) mx = fmaxf(mx, a[l]);
      float sum = 0.f;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          a[l] = expf(a[l] - mx);
          sum += a[l];
        }
      const float inv = 1.f / sum;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) sp[s * LP + l] = a[l] * inv;
    }
  }
  __syncthreads();
  // phase 3: per word, softmax over the regions of gamma1 * P (ref: GlobalAttention.py:57-62)
  float* ab = attn + (long long)p * T * S;
  for (int l = warp; l < Lq; l += WP_THREADS / 32) {
    float mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmaxf(mx, sp[s * LP + l] * gamma1);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float e = expf(sp[s * LP + l] * gamma1 - mx);
      sp[s * LP + l] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int s = lane; s < S; s += 32) {
      const float a = sp[s * LP + l] * inv;
      sp[s * LP + l] = a;
      ab[(long long)l * S + s] = a;
    }
  }
  __syncthreads();
  // phase 4: weighted context, two channels per warp pass (they share the probability loads); lane l collects the
  // three inner products the cosine of word l needs
  float p12 = 0.f, p11 = 0.f, p22 = 0.f;
  float* wb = wc + (long long)p * ndf * T;
  for (int c = warp * 2; c < ndf; c += 2 * (WP_THREADS / 32)) {
    float a0[LM], a1[LM];
#pragma unroll
    for (int l = 0; l < LM; ++l) a0[l] = a1[l] = 0.f;
    const bool two = c + 1 < ndf;
    for (int s = lane; s < S; s += 32) {
      const float c0 = __ldg(cb + (long long)c * S + s);
      const float c1 = two ? __ldg(cb + (long long)(c + 1) * S + s) : 0.f;
      const float* pr = sp + s * LP;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          const float pv = pr[l];
          a0[l] = fmaf(c0, pv, a0[l]);
          a1[l] = fmaf(c1, pv, a1[l]);
        }
    }
#pragma unroll
    for (int l = 0; l < LM; ++l)
      if (l < Lq) {
        const float v0 = warp_sum(a0[l]), v1 = warp_sum(a1[l]);
        if (lane == l) {
          const float q0 = sq[c * Lq + l];
          wb[c * T + l] = v0;
          p12 = fmaf(q0, v0, p12); p11 = fmaf(q0, q0, p11); p22 = fmaf(v0, v0, p22);
          if (two) {
            const float q1 = sq[(c + 1) * Lq + l];
            wb[(c + 1) * T + l] = v1;
            p12 = fmaf(q1, v1, p12); p11 = fmaf(q1, q1, p11); p22 = fmaf(v1, v1, p22);
          }
        }
      }
  }
  if (lane < LM) {
    red[(warp * 3 + 0) * LM + lane] = p12;
    red[(warp * 3 + 1) * LM + lane] = p11;
    red[(warp * 3 + 2) * LM + lane] = p22;
  }
  __syncthreads();
  if (warp == 0) {
    float e = 0.f;
    if (lane < Lq) {
      float w12 = 0.f, w1 = 0.f, w2 = 0.f;
      for (int w = 0; w < WP_THREADS / 32; ++w) {
        w12 += red[(w * 3 + 0) * LM + lane];
        w1 += red[(w * 3 + 1) * LM + lane];
        w2 += red[(w * 3 + 2) * LM + lane];
      }
      const float cs = w12 / fmaxf(sqrtf(w1) * sqrtf(w2), eps);
      e = expf(gamma2 * cs);
    }
    e = warp_sum(e);
    if (lane == 0) sim[p] = logf(e);
  }
}

// Backward of one pair; see the header.  g_ctx must be zero-filled by the launcher.
template <int LM>
__global__ void __launch_bounds__(WP_THREADS) words_pairs_bwd_kernel(
    const float* __restrict__ words, const float* __restrict__ ctx, const long long* __restrict__ lens,
    const float* __restrict__ wc, const float* __restrict__ attn, const float* __restrict__ g_sim, int NC, int ndf,
    int T, int S, float gamma1, float gamma2, float eps, float* __restrict__ g_ctx) {
  extern __shared__ float smem[];
  const int p = blockIdx.x, b = p / NC, i = p - b * NC;
  const int Lq = min((int)lens[i], T);
  const int LP = Lq | 1;
  float* sq = smem;                      // [ndf][Lq]   caption
  float* sg = sq + ndf * T;              // [ndf][Lq]   wc, then g_wc
  float* sP = sg + ndf * T;              // [S][LP]     P, then gS
  float* sP2 = sP + S * (T | 1);         // [Lq][S]     P2
  float* sG = sP2 + T * S;               // [Lq][S]     gP2, then gT
  float* red = sG + T * S;               // [5][3][LM] + [LM] coefficients x 2
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* qb = words + (long long)i * ndf * T;
  const float* cb = ctx + (long long)b * ndf * S;
  const float* wb = wc + (long long)p * ndf * T;
  const float go = g_sim[p];
  for (int k = t; k < ndf * Lq; k += WP_THREADS) {
    const int c = k / Lq, l = k - c * Lq;
    sq[k] = qb[c * T + l];
    sg[k] = wb[c * T + l];
  }
  for (int k = t; k < Lq * S; k += WP_THREADS) sP2[k] = attn[(long long)p * T * S + k];
  __syncthreads();
  // (0) g_wc from the cosine / log-sum-exp chain: per word l the three inner products over the channels
  {
    for (int l = warp; l < Lq; l += WP_THREADS / 32) {
      float w12 = 0.f, w1 = 0.f, w2 = 0.f;
      for (int c = lane; c < ndf; c += 32) {
        const float x = sq[c * Lq + l], y = sg[c * Lq + l];
        w12 = fmaf(x, y, w12); w1 = fmaf(x, x, w1); w2 = fmaf(y, y, w2);
      }
      w12 = warp_sum(w12); w1 = warp_sum(w1); w2 = warp_sum(w2);
      if (lane == 0) {
        red[l] = w12; red[LM + l] = w1; red[2 * LM + l] = w2;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    float cs = 0.f, e = 0.f;
    if (lane < Lq) {
      cs = red[lane] / fmaxf(sqrtf(red[LM + lane]) * sqrtf(red[2 * LM + lane]), eps);
      e = expf(gamma2 * cs);
    }
    const float tot = warp_sum(e);
    if (lane < Lq) {
      const float gc = go * gamma2 * e / tot;                  // d sim / d cos[l]
      const float n1 = sqrtf(red[LM + lane]), n2 = sqrtf(red[2 * LM + lane]), prod = n1 * n2;
      // cos = w12 / (n1 n2): d/dy_c = x_c / (n1 n2) - cos * y_c / n2^2 ; clamped denominator: x_c / eps
      float ka, kb;
      if (prod >= eps) { ka = gc / prod; kb = gc * (red[lane] / prod) / red[2 * LM + lane]; }
      else { ka = gc / eps; kb = 0.f; }
      red[3 * LM + lane] = ka;
      red[4 * LM + lane] = kb;
    }
  }
  __syncthreads();
  for (int k = t; k < ndf * Lq; k += WP_THREADS) {
    const int l = k % Lq;
    sg[k] = red[3 * LM + l] * sq[k] - red[4 * LM + l] * sg[k];
  }
  __syncthreads();
  // (1) recompute P (softmax over the words, per region) and gP2[l][s] = sum_c g_wc[c][l] ctx[c][s]
  {
    const int s0 = t, s1 = t + WP_THREADS;
    const bool v0 = s0 < S, v1 = s1 < S;
    float a0[LM], a1[LM], g0[LM], g1[LM];
#pragma unroll
    for (int l = 0; l < LM; ++l) a0[l] = a1[l] = g0[l] = g1[l] = 0.f;
    for (int c = 0; c < ndf; ++c) {
      const float c0 = v0 ? __ldg(cb + (long long)c * S + s0) : 0.f;
      const float c1 = v1 ? __ldg(cb + (long long)c * S + s1) : 0.f;
      const float* qr = sq + c * Lq;
      const float* gr = sg + c * Lq;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          const float q = qr[l], g = gr[l];
          a0[l] = fmaf(c0, q, a0[l]); a1[l] = fmaf(c1, q, a1[l]);
          g0[l] = fmaf(c0, g, g0[l]); g1[l] = fmaf(c1, g, g1[l]);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* a = h ? a1 : a0;
      float* g = h ? g1 : g0;
      const int s = h ? s1 : s0;
      if (!(h ? v1 : v0)) continue;
      float mx = -INFINITY;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) mx = fmaxf(mx, a[l]);
      float sum = 0.f;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          a[l] = expf(a[l] - mx);
          sum += a[l];
        }
      const float inv = 1.f / sum;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          sP[s * LP + l] = a[l] * inv;
          sG[l * S + s] = g[l];
        }
    }
  }
  __syncthreads();
  // (2) softmax backward over the regions, per word: gT = P2 * (gP2 - <gP2, P2>)
  for (int l = warp; l < Lq; l += WP_THREADS / 32) {
    float d = 0.f;
    for (int s = lane; s < S; s += 32) d = fmaf(sG[l * S + s], sP2[l * S + s], d);
    d = warp_sum(d);
    for (int s = lane; s < S; s += 32) sG[l * S + s] = sP2[l * S + s] * (sG[l * S + s] - d);
  }
  __syncthreads();
  // (3) softmax backward over the words, per region: gS = P * (gP - <gP, P>),  gP = gamma1 * gT^T
  for (int s = t; s < S; s += WP_THREADS) {
    float d = 0.f;
    for (int l = 0; l < Lq; ++l) d = fmaf(gamma1 * sG[l * S + s], sP[s * LP + l], d);
    for (int l = 0; l < Lq; ++l) sP[s * LP + l] = sP[s * LP + l] * (gamma1 * sG[l * S + s] - d);
  }
  __syncthreads();
  // (4) g_ctx[c][s] += sum_l g_wc[c][l] P2[l][s] + gS[s][l] q[c][l]
  float* gb = g_ctx + (long long)b * ndf * S;
  for (int c = warp; c < ndf; c += WP_THREADS / 32) {
    const float* qr = sq + c * Lq;
    const float* gr = sg + c * Lq;
    for (int s = lane; s < S; s += 32) {
      float a = 0.f;
#pragma unroll
      for (int l = 0; l < LM; ++l)
        if (l < Lq) {
          a = fmaf(gr[l], sP2[l * S + s], a);
          a = fmaf(sP[s * LP + l], qr[l], a);
        }
      atomicAdd(gb + (long long)c * S + s, a);
    }
  }
}

template <int LM>
int launch_fwd(const float* words, const float* ctx, const long long* lens, int B, int NC, int ndf, int T, int S,
               float gamma1, float gamma2, float eps, float* wc, float* attn, float* sim, cudaStream_t stream) {
  const size_t sm = sizeof(float) * ((size_t)ndf * T + (size_t)S * (T | 1) + 5 * 3 * LM);
  OG_CHECK(cudaFuncSetAttribute(words_pairs_fwd_kernel<LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  words_pairs_fwd_kernel<LM><<<B * NC, WP_THREADS, sm, stream>>>(words, ctx, lens, NC, ndf, T, S, gamma1, gamma2, eps,
                                                                wc, attn, sim);
  OG_RETURN_LAST_ERROR();
}
template <int LM>
int launch_bwd(const float* words, const float* ctx, const long long* lens, const float* wc, const float* attn,
               const float* g_sim, int B, int NC, int ndf, int T, int S, float gamma1, float gamma2, float eps,
               float* g_ctx, cudaStream_t stream) {
  const size_t sm = sizeof(float) * (2 * (size_t)ndf * T + (size_t)S * (T | 1) + 2 * (size_t)T * S + 5 * LM);
  OG_CHECK(cudaFuncSetAttribute(words_pairs_bwd_kernel<LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  words_pairs_bwd_kernel<LM><<<B * NC, WP_THREADS, sm, stream>>>(words, ctx, lens, wc, attn, g_sim, NC, ndf, T, S,
                                                                gamma1, gamma2, eps, g_ctx);
  OG_RETURN_LAST_ERROR();
}
}  // namespace

// words [NC][ndf][T], ctx [B][ndf][S], lens [NC] (int64, device) -> wc [B*NC][ndf][T], attn [B*NC][T][S], sim [B][NC]
// (entries of wc / attn beyond a caption's length are left untouched: zero-fill them if they are read)
OG_API int og_words_pairs_fwd(const float* words, const float* ctx, const long long* lens, int B, int NC, int ndf, int T,
                              int S, float gamma1, float gamma2, float eps, float* wc, float* attn, float* sim,
                              cudaStream_t stream) {
  if (T > 32 || S > 2 * WP_THREADS || T < 1) return (int)cudaErrorInvalidValue;
  if (B * NC == 0) return 0;
  return T <= 20 ? launch_fwd<20>(words, ctx, lens, B, NC, ndf, T, S, gamma1, gamma2, eps, wc, attn, sim, stream)
                 : launch_fwd<32>(words, ctx, lens, B, NC, ndf, T, S, gamma1, gamma2, eps, wc, attn, sim, stream);
}
OG_API int og_words_pairs_bwd(const float* words, const float* ctx, const long long* lens, const float* wc,
                              const float* attn, const float* g_sim, int B, int NC, int ndf, int T, int S, float gamma1,
                              float gamma2, float eps, float* g_ctx, cudaStream_t stream) {
  if (T > 32 || S > 2 * WP_THREADS || T < 1) return (int)cudaErrorInvalidValue;
  OG_CHECK(cudaMemsetAsync(g_ctx, 0, sizeof(float) * (size_t)B * ndf * S, stream));
  if (B * NC == 0) return 0;
  return T <= 20 ? launch_bwd<20>(words, ctx, lens, wc, attn, g_sim, B, NC, ndf, T, S, gamma1, gamma2, eps, g_ctx, stream)
                 : launch_bwd<32>(words, ctx, lens, wc, attn, g_sim, B, NC, ndf, T, S, gamma1, gamma2, eps, g_ctx, stream);
}

OG_API int og_zero_bytes(float* p, long long bytes, cudaStream_t stream) {
  if (bytes <= 0) return 0;
  OG_CHECK(cudaMemsetAsync(p, 0, (size_t)bytes, stream));
  return 0;
}
