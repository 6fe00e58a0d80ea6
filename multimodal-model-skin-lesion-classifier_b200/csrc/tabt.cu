// tabt.cu - the TabTransformer encoder as ONE kernel per pass: embedding gather + every
// nn.TransformerEncoderLayer of the stack, one sample per CTA, all activations of the sample in shared memory.
//
// Reference: models/tab_transformer.py:6-60.  82 categorical columns -> nn.Embedding(card, 32) each, stacked to
// [B, 82, 32] (:42-43); nn.TransformerEncoder of 2 post-norm layers (d_model 32, 4 heads, dim_feedforward 128, ReLU,
// dropout 0.3, batch_first) (:19-27, :46); flatten to [B, 82*32] (:47).  The numeric projection, the concatenation and the
// fc MLP that follow (:50-60) are plain row-major GEMMs and run on the library's GEMM engines (exec.cu, fb200_linear_*).
//
// Why one CTA per sample: a sample is 82 x 32 floats = 10 KB; its queries/keys/values (31 KB), feed-forward activations
// (42 KB) and every gradient buffer fit the 227 KB of shared memory of one SM, and a layer's weights (50 KB) are shared
// by all CTAs through L1/L2.  Stock PyTorch runs ~40 kernels per layer over [B*82, 32]-shaped tensors and writes the
// [B, 4, 82, 82] probabilities to HBM; here HBM sees the category codes, the layer inputs (one [B, 82*32] tensor per layer,
// saved for backward) and the output - everything else is recomputed on chip by the backward kernel, flash-attention style.
//
// Work split inside the CTA (512 threads): every phase is a "parallel for" over independent output elements, each
// computed serially by ONE thread (register-tiled 4 tokens x 4 outputs for the projections; one (head, row) per thread
// for the softmax), phases separated by __syncthreads().  No atomics anywhere: weight gradients accumulate across the
// samples of a CTA in that CTA's own slab in global memory (L2-resident, plain read-modify-write), and a second kernel
// adds the slabs in CTA order - the results are bit-reproducible.  All arithmetic is fp32 FFMA (parity bar 1e-5).
#include "common.cuh"
#include <mutex>

namespace fb200 {

constexpr int TABT_THREADS = 512;
constexpr int TABT_SITE0 = 16;            // Philox site ids 16 + 4 * layer + {0: attention probabilities, 1: after attention, 2: feed-forward, 3: after feed-forward}

// flat fp32 parameter block of one layer, in nn.TransformerEncoderLayer's state_dict order (offsets in floats)
struct TabtOff { int win, bin, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2, size; };
__host__ __device__ inline TabtOff tabt_off(int D, int F) {
  TabtOff o; int c = 0;
  o.win = c; c += 3 * D * D; o.bin = c; c += 3 * D; o.wo = c; c += D * D; o.bo = c; c += D;
  o.w1 = c; c += F * D; o.b1 = c; c += F; o.w2 = c; c += D * F; o.b2 = c; c += D;
  o.g1 = c; c += D; o.be1 = c; c += D; o.g2 = c; c += D; o.be2 = c; c += D;
  o.size = c;
  return o;
}
// row stride >= K with stride % 32 == 4: eight consecutive rows read with 128-bit loads cover the 32 banks exactly once
__host__ __device__ inline int tabt_ld(int K) { return ((K + 27) / 32) * 32 + 4; }

struct TabtSmem { int X, QKV, A, Xh1, X1, Hb, Xh2, rstd1, rstd2, lse, delta, bits, G, G2, dH, total; };
__host__ __device__ inline TabtSmem tabt_smem(int T, int D, int F, int H, bool bwd) {
  const int ldD = tabt_ld(D), ld3 = tabt_ld(3 * D), ldF = tabt_ld(F);
  const int W = (((T + 3) & ~3) + 31) / 32;
  TabtSmem s; int c = 0;
  auto take = [&](int n) { int at = c; c += (n + 3) & ~3; return at; };
  // Backward keeps its shared memory under 164 KB so that the SM's L1 keeps 92 KB for the layer's weights (50 KB, read by
  // every phase through L1): the feed-forward gradient dH overwrites Hb in place, dQKV later reuses the same region, dZ / dY
  // (G2) live in the buffer of the layer input X, which is re-read from global memory into the dead X1 buffer for the last phase.
  s.X = take(T * ldD); s.QKV = take(T * ld3); s.A = take(T * ldD); s.Xh1 = take(T * ldD); s.X1 = take(T * ldD);
  s.Hb = take(T * (bwd && ld3 > ldF ? ld3 : ldF)); s.Xh2 = take(T * ldD);
  s.rstd1 = take(T); s.rstd2 = take(T); s.lse = take(H * T); s.delta = take(H * T); s.bits = take(H * T * W);
  s.G = 0; s.G2 = s.X; s.dH = s.Hb;
  if (bwd) s.G = take(T * ldD);
  s.total = c;
  return s;
}

struct TabtArgs {
  int B, T, D, H, F, L;
  const long long* codes;        // [B, T] category codes (torch.long, tab_transformer.py:42)
  const int* emb_base;           // [T] first row of column t's table inside the stacked embedding table
  int n_emb_rows;
  const float* params;           // L layer blocks (TabtOff) then the stacked embedding table [n_emb_rows, D]
  float* out; int ldo;           // forward: [B, T*D] = transformer_encoder(tokens).flatten(1), row stride ldo
  float* saved;                  // [L-1][B][T*D]: inputs of layers 1..L-1 (forward writes, backward reads)
  const float* dout; int lddo;   // backward: gradient of out
  float* slab;                   // backward: [gridDim.x][p_total] per-CTA gradient partials (zeroed by the host)
  long long p_total;
  int train; float p; uint64_t seed, offset; const uint64_t* rng_state;
  const uint8_t* mask_attn;      // optional explicit keep-masks: [L][B][H][T][T]
  const uint8_t* mask_res1;      //                               [L][B*T][D]
  const uint8_t* mask_ff;        //                               [L][B*T][F]
  const uint8_t* mask_res2;      //                               [L][B*T][D]
  long long* trace;              // debug: clock64 stamps of CTA 0 after every phase (fb200_debug_tabt_trace), else nullptr
};

#define TABT_STAMP(a, k) do { if ((a).trace && blockIdx.x == 0 && threadIdx.x == 0) (a).trace[k] = clock64(); } while (0)

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// ---- register-tiled projections over the sample's tokens ---------------------------------------------------------------------
// Y[t, n0..n0+3] = bias + sum_k X[t, k] W[n, k]   (nn.Linear forward; W [N, K] row-major in global memory, X in shared memory).
// One work item = TT tokens x 4 outputs; the lanes of a warp take consecutive tokens of the SAME output group, so the
// weight loads are warp-wide broadcasts (L1) and the activation loads are conflict-free (tabt_ld).
template <int TT, class Epi>
__device__ __forceinline__ void tabt_lin_nt(const float* Xs, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                                            int T, int N, int K, Epi epi) {
  const int ntg = (T + TT - 1) / TT, items = ntg * (N >> 2);
  for (int it = threadIdx.x; it < items; it += TABT_THREADS) {
    const int tg = it % ntg, n0 = (it / ntg) << 2;
    float acc[TT][4];
    const float4 b4 = bias ? __ldg((const float4*)(bias + n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xr[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w;
      xr[j] = Xs + min(tg + j * ntg, T - 1) * ldx;               // rows past the end recompute the last row; never stored
    }
    const float* w0 = W + (size_t)n0 * K;
    for (int k = 0; k < K; k += 4) {
      const float4 wa = __ldg((const float4*)(w0 + k)), wb = __ldg((const float4*)(w0 + K + k));
      const float4 wc = __ldg((const float4*)(w0 + 2 * K + k)), wd = __ldg((const float4*)(w0 + 3 * K + k));
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        const float4 x = *(const float4*)(xr[j] + k);
        acc[j][0] = dot4(x, wa, acc[j][0]); acc[j][1] = dot4(x, wb, acc[j][1]);
        acc[j][2] = dot4(x, wc, acc[j][2]); acc[j][3] = dot4(x, wd, acc[j][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      const int t = tg + j * ntg;
      if (t < T) epi(t, n0, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
    }
  }
}

// dX[t, k0..k0+3] = sum_n dY[t, n] W[n, k]   (input gradient of nn.Linear; same W, read along its rows)
template <int TT, class Epi>
__device__ __forceinline__ void tabt_lin_nn(const float* Ys, int ldy, const float* __restrict__ W, int T, int N, int K, Epi epi) {
  const int ntg = (T + TT - 1) / TT, items = ntg * (K >> 2);
  for (int it = threadIdx.x; it < items; it += TABT_THREADS) {
    const int tg = it % ntg, k0 = (it / ntg) << 2;
    float acc[TT][4];
    const float* yr[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      yr[j] = Ys + min(tg + j * ntg, T - 1) * ldy;
    }
    for (int n = 0; n < N; n += 4) {
      const float* wr = W + (size_t)n * K + k0;
      const float4 wa = __ldg((const float4*)wr), wb = __ldg((const float4*)(wr + K));
      const float4 wc = __ldg((const float4*)(wr + 2 * K)), wd = __ldg((const float4*)(wr + 3 * K));
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        const float4 y = *(const float4*)(yr[j] + n);
        acc[j][0] = fmaf(y.x, wa.x, fmaf(y.y, wb.x, fmaf(y.z, wc.x, fmaf(y.w, wd.x, acc[j][0]))));
        acc[j][1] = fmaf(y.x, wa.y, fmaf(y.y, wb.y, fmaf(y.z, wc.y, fmaf(y.w, wd.y, acc[j][1]))));
        acc[j][2] = fmaf(y.x, wa.z, fmaf(y.y, wb.z, fmaf(y.z, wc.z, fmaf(y.w, wd.z, acc[j][2]))));
        acc[j][3] = fmaf(y.x, wa.w, fmaf(y.y, wb.w, fmaf(y.z, wc.w, fmaf(y.w, wd.w, acc[j][3]))));
      }
    }
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      const int t = tg + j * ntg;
      if (t < T) epi(t, k0, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
    }
  }
}

__device__ __forceinline__ void slab_add4(float* p, float a, float b, float c, float d) {
  float4 o = __ldcg((const float4*)p);
  o.x += a; o.y += b; o.z += c; o.w += d;
  __stcg((float4*)p, o);
}

// dW[n0..n0+3, k0..k0+3] += sum_t dY[t, n] X[t, k]  -> this CTA's slab.  One 4 x 4 tile per group of TS lanes: the lanes of a
// group split the tokens and are summed by shuffles in a fixed order.
template <int TS>
__device__ __forceinline__ void tabt_grad_tn(const float* Ys, int ldy, const float* Xs, int ldx, int T, int N, int K, float* slabW) {
  constexpr int WI = 32 / TS;                                  // tiles per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / WI, items = (N >> 2) * (K >> 2), kgs = K >> 2;
  for (int base = 0; base < items; base += (TABT_THREADS / 32) * WI) {    // uniform trip count: every lane reaches the shuffles
    const int it = base + warp * WI + (lane % WI);
    const bool valid = it < items;
    const int k0 = (it % kgs) << 2, n0 = (it / kgs) << 2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    if (valid) {
      for (int t = sub; t < T; t += TS) {
        const float4 y = *(const float4*)(Ys + t * ldy + n0), x = *(const float4*)(Xs + t * ldx + k0);
        acc[0][0] = fmaf(y.x, x.x, acc[0][0]); acc[0][1] = fmaf(y.x, x.y, acc[0][1]); acc[0][2] = fmaf(y.x, x.z, acc[0][2]); acc[0][3] = fmaf(y.x, x.w, acc[0][3]);
        acc[1][0] = fmaf(y.y, x.x, acc[1][0]); acc[1][1] = fmaf(y.y, x.y, acc[1][1]); acc[1][2] = fmaf(y.y, x.z, acc[1][2]); acc[1][3] = fmaf(y.y, x.w, acc[1][3]);
        acc[2][0] = fmaf(y.z, x.x, acc[2][0]); acc[2][1] = fmaf(y.z, x.y, acc[2][1]); acc[2][2] = fmaf(y.z, x.z, acc[2][2]); acc[2][3] = fmaf(y.z, x.w, acc[2][3]);
        acc[3][0] = fmaf(y.w, x.x, acc[3][0]); acc[3][1] = fmaf(y.w, x.y, acc[3][1]); acc[3][2] = fmaf(y.w, x.z, acc[3][2]); acc[3][3] = fmaf(y.w, x.w, acc[3][3]);
      }
    }
#pragma unroll
    for (int o = WI; o < 32; o <<= 1)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] += __shfl_xor_sync(0xffffffffu, acc[i][c], o);
    if (valid && sub == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) slab_add4(slabW + (size_t)(n0 + i) * K + k0, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

// slab[c] += sum_t Y[t, c]  (bias gradients), one column per thread
__device__ __forceinline__ void tabt_colsum(const float* Ys, int ldy, int T, int N, float* slab) {
  for (int c = threadIdx.x; c < N; c += TABT_THREADS) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += Ys[t * ldy + c];
    slab[c] = __ldcg(slab + c) + s;
  }
}
// LayerNorm parameter gradients: dgamma[c] += sum_t G[t, c] xhat[t, c], dbeta[c] += sum_t G[t, c]
__device__ __forceinline__ void tabt_ln_param_grads(const float* Gs, const float* Xh, int ld, int T, int D, float* dgamma, float* dbeta) {
  for (int c = threadIdx.x; c < D; c += TABT_THREADS) {
    float sg = 0.f, sb = 0.f;
    for (int t = 0; t < T; ++t) { const float g = Gs[t * ld + c]; sg = fmaf(g, Xh[t * ld + c], sg); sb += g; }
    dgamma[c] = __ldcg(dgamma + c) + sg;
    dbeta[c] = __ldcg(dbeta + c) + sb;
  }
}

// ---- dropout -------------------------------------------------------------------------------------------------------------------
struct TabtDrop { int active; float p, keep_scale; uint64_t seed, offset; uint32_t thr; };
__device__ __forceinline__ TabtDrop tabt_drop(const TabtArgs& a) {
  TabtDrop d;
  d.active = a.train && a.p > 0.f; d.p = a.p; d.keep_scale = d.active ? 1.0f / (1.0f - a.p) : 1.0f;
  d.seed = a.rng_state ? __ldg(a.rng_state) : a.seed;
  d.offset = a.rng_state ? __ldg(a.rng_state + 1) : a.offset;
  d.thr = (uint32_t)fminf(a.p * 4294967296.0f, 4294967040.0f);
  return d;
}
// keep flags (bit i = element g*4 + i kept) of Philox group g of `site`
__device__ __forceinline__ uint32_t tabt_keep4(const TabtDrop& d, uint64_t g, int site) {
  uint4 ctr = make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)d.offset, (uint32_t)(d.offset >> 32) ^ ((uint32_t)site << 24));
  const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
  return (r.x >= d.thr ? 1u : 0u) | (r.y >= d.thr ? 2u : 0u) | (r.z >= d.thr ? 4u : 0u) | (r.w >= d.thr ? 8u : 0u);
}
// keep-multipliers of the four elements (row, c..c+3) of a width-N element-wise site
__device__ __forceinline__ float4 tabt_mult4(const TabtDrop& d, const uint8_t* mask, int site, int64_t row, int c, int N) {
  if (!d.active) return make_float4(1.f, 1.f, 1.f, 1.f);
  uint32_t k;
  if (mask) { const uchar4 m = *(const uchar4*)(mask + row * N + c); k = (m.x ? 1u : 0u) | (m.y ? 2u : 0u) | (m.z ? 4u : 0u) | (m.w ? 8u : 0u); }
  else k = tabt_keep4(d, (uint64_t)(row * N + c) >> 2, site);
  const float s = d.keep_scale;
  return make_float4(k & 1u ? s : 0.f, k & 2u ? s : 0.f, k & 4u ? s : 0.f, k & 8u ? s : 0.f);
}

// ---- attention rows -------------------------------------------------------------------------------------------------------------
// One (head, query) per thread, ONE pass over the T keys with a running (max, sum, output row) - four keys per trip, the running
// maximum only moves (and the accumulators are only rescaled) when one of the four beats it.  Scores are kept in base-2 units
// (q is pre-multiplied by log2(e) / sqrt(hd)), so every exponential is one ex2 and the saved log-sum-exp is in base 2 as well.
// Dropout acts on the normalised probabilities (F.multi_head_attention_forward / SDPA dropout_p):
// A_i = sum_j P_ij keep_ij / (1 - p) V_j with the denominator of P taken over ALL keys.  keep bits of row (h, i) go to
// bits[(h*T + i) * W + j/32] for the backward phases.
constexpr float TABT_LOG2E = 1.4426950408889634f, TABT_LN2 = 0.6931471805599453f;
template <int HD>
__device__ __forceinline__ void tabt_attn_fwd(const float* QKV, int ld3, int T, int H, int D, float scale, float* A, int ldA,
                                              float* lse2, uint32_t* bits, int W, const TabtDrop& dr, const uint8_t* mask /*[H][T][T]*/,
                                              int site, int64_t b) {
  const int T4 = (T + 3) & ~3;
  for (int it = threadIdx.x; it < H * T; it += TABT_THREADS) {
    const int h = it / T, i = it - h * T;
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = QKV[i * ld3 + h * HD + d];              // pre-scaled by the in-projection epilogue
    const float* Kb = QKV + D + h * HD;
    const float* Vb = QKV + 2 * D + h * HD;
    float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    uint32_t word = 0;
    for (int j0 = 0; j0 < T; j0 += 4) {
      uint32_t keep = 0xfu;
      if (dr.active) {
        if (mask) {
          keep = 0;
          for (int jj = 0; jj < 4 && j0 + jj < T; ++jj) keep |= mask[((size_t)h * T + i) * T + j0 + jj] ? (1u << jj) : 0u;
        } else keep = tabt_keep4(dr, (uint64_t)((((int64_t)b * H + h) * T + i) * T4 + j0) >> 2, site);
      }
      word |= keep << (j0 & 31);
      float sc[4];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = min(j0 + jj, T - 1);                       // the tail re-reads the last key; masked below
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) a = fmaf(q[d], Kb[j * ld3 + d], a);
        sc[jj] = j0 + jj < T ? a : -INFINITY;
        mx = fmaxf(mx, sc[jj]);
      }
      if (mx > m) {                                              // exp2(-inf) = 0 on the first trip
        const float c = exp2f(m - mx);
        l *= c;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= c;
        m = mx;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = min(j0 + jj, T - 1);
        const float e = exp2f(sc[jj] - m);                       // 0 for the masked tail
        l += e;
        const float ek = (keep >> jj) & 1u ? e : 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = fmaf(ek, Vb[j * ld3 + d], o[d]);
      }
      if (bits && (((j0 + 4) & 31) == 0 || j0 + 4 >= T)) { bits[it * W + (j0 >> 5)] = word; word = 0; }
    }
    const float inv = dr.keep_scale / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) A[i * ldA + h * HD + d] = o[d] * inv;
    if (lse2) lse2[it] = m + log2f(l);
  }
}

// dQ rows and delta_i = dA_i . A_i  (one (head, query) per thread)
template <int HD>
__device__ __forceinline__ void tabt_attn_bwd_q(const float* QKV, int ld3, const float* A, const float* dA, int ldA, int T, int H, int D, float scale,
                                                const float* lse, const uint32_t* bits, int W, float keep_scale, float* delta, float* dQKV, int ldg) {
  for (int it = threadIdx.x; it < H * T; it += TABT_THREADS) {
    const int h = it / T, i = it - h * T;
    float q[HD], g[HD], dq[HD];
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      q[d] = QKV[i * ld3 + h * HD + d]; g[d] = dA[i * ldA + h * HD + d]; dq[d] = 0.f;
      dl = fmaf(g[d], A[i * ldA + h * HD + d], dl);
    }
    const float* Kb = QKV + D + h * HD;
    const float* Vb = QKV + 2 * D + h * HD;
    const float ls = lse[it];
#pragma unroll 2
    for (int j = 0; j < T; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) { s = fmaf(q[d], Kb[j * ld3 + d], s); dp = fmaf(g[d], Vb[j * ld3 + d], dp); }
      const float keep = (bits[it * W + (j >> 5)] >> (j & 31)) & 1u ? keep_scale : 0.f;
      const float ds = T == 1 ? 0.f : exp2f(s - ls) * (dp * keep - dl);     // one key: dS == 0 exactly, as in torch
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, Kb[j * ld3 + d], dq[d]);
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) dQKV[i * ldg + h * HD + d] = dq[d] * scale;
    delta[it] = dl;
  }
}

// dK and dV rows (one (head, key) per thread; the queries are walked in order)
template <int HD>
__device__ __forceinline__ void tabt_attn_bwd_kv(const float* QKV, int ld3, const float* dA, int ldA, int T, int H, int D, float scale,
                                                 const float* lse, const float* delta, const uint32_t* bits, int W, float keep_scale,
                                                 float* dQKV, int ldg) {
  for (int it = threadIdx.x; it < H * T; it += TABT_THREADS) {
    const int h = it / T, j = it - h * T;
    float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { k[d] = QKV[j * ld3 + D + h * HD + d]; v[d] = QKV[j * ld3 + 2 * D + h * HD + d]; dk[d] = 0.f; dv[d] = 0.f; }
#pragma unroll 2
    for (int i = 0; i < T; ++i) {
      float s = 0.f, dp = 0.f;
      float q[HD], g[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        q[d] = QKV[i * ld3 + h * HD + d]; g[d] = dA[i * ldA + h * HD + d];   // base-2 score units; dk is rescaled at the end
        s = fmaf(q[d], k[d], s); dp = fmaf(g[d], v[d], dp);
      }
      const int row = h * T + i;
      const float keep = (bits[row * W + (j >> 5)] >> (j & 31)) & 1u ? keep_scale : 0.f;
      const float pr = exp2f(s - lse[row]);
      const float pk = pr * keep;
      const float ds = T == 1 ? 0.f : pr * (dp * keep - delta[row]);
#pragma unroll
      for (int d = 0; d < HD; ++d) { dv[d] = fmaf(pk, g[d], dv[d]); dk[d] = fmaf(ds, q[d], dk[d]); }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) { dQKV[j * ldg + D + h * HD + d] = dk[d] * TABT_LN2; dQKV[j * ldg + 2 * D + h * HD + d] = dv[d]; }
  }
}

// ---- residual + LayerNorm (four adjacent lanes per token, sums by quad shuffles) -----------------------------------------------------
__device__ __forceinline__ float tabt_quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// v = R[t] + mult * Y[t];  xhat = (v - mean) rstd -> Y (in place);  dst[t] = xhat * gamma + beta (when dst != nullptr)
__device__ __forceinline__ void tabt_res_ln(const float* R, float* Y, int ld, int T, int D, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const TabtDrop& dr, const uint8_t* mask, int site, int64_t row0, float* dst, float* rstd_s) {
  const int sub = threadIdx.x & 3;
  for (int t0 = 0; t0 < T; t0 += TABT_THREADS / 4) {               // uniform trip count: every lane reaches the shuffles
    const int t = t0 + (threadIdx.x >> 2);
    const bool ok = t < T;
    const int tt = ok ? t : T - 1;
    float sum = 0.f;
    for (int k = 4 * sub; k < D; k += 16) {
      const float4 r = *(const float4*)(R + tt * ld + k), y = *(const float4*)(Y + tt * ld + k), m = tabt_mult4(dr, mask, site, row0 + tt, k, D);
      const float4 v = make_float4(fmaf(y.x, m.x, r.x), fmaf(y.y, m.y, r.y), fmaf(y.z, m.z, r.z), fmaf(y.w, m.w, r.w));
      if (ok) *(float4*)(Y + tt * ld + k) = v;
      sum += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = tabt_quad_sum(sum) / (float)D;
    __syncwarp();                                                    // the quad re-reads only its own writes; kept for clarity of the memory order
    float var = 0.f;
    for (int k = 4 * sub; k < D; k += 16) {
      const float4 v = *(const float4*)(Y + tt * ld + k);
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(tabt_quad_sum(var) / (float)D + 1e-5f);
    if (ok) {
      for (int k = 4 * sub; k < D; k += 16) {
        const float4 v = *(const float4*)(Y + tt * ld + k);
        const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
        *(float4*)(Y + tt * ld + k) = xh;
        if (dst) {
          const float4 g = __ldg((const float4*)(gamma + k)), be = __ldg((const float4*)(beta + k));
          *(float4*)(dst + tt * ld + k) = make_float4(fmaf(xh.x, g.x, be.x), fmaf(xh.y, g.y, be.y), fmaf(xh.z, g.z, be.z), fmaf(xh.w, g.w, be.w));
        }
      }
      if (sub == 0) rstd_s[tt] = rstd;
    }
  }
}
// G[t] <- dS = rstd (g*gamma - mean(g*gamma) - xhat mean(g*gamma*xhat))  (gradient of the LayerNorm input = of the residual sum);
// G2[t] <- dS * mult  (gradient of the branch that went through dropout)
__device__ __forceinline__ void tabt_ln_bwd(float* Gs, const float* Xh, int ld, int T, int D, const float* __restrict__ gamma, const float* rstd_s,
                                            const TabtDrop& dr, const uint8_t* mask, int site, int64_t row0, float* G2) {
  const int sub = threadIdx.x & 3;
  for (int t0 = 0; t0 < T; t0 += TABT_THREADS / 4) {
    const int t = t0 + (threadIdx.x >> 2);
    const bool ok = t < T;
    const int tt = ok ? t : T - 1;
    float m1 = 0.f, m2 = 0.f;
    for (int k = 4 * sub; k < D; k += 16) {
      const float4 g = *(const float4*)(Gs + tt * ld + k), ga = __ldg((const float4*)(gamma + k)), xh = *(const float4*)(Xh + tt * ld + k);
      const float a = g.x * ga.x, b = g.y * ga.y, c = g.z * ga.z, d = g.w * ga.w;
      m1 += (a + b) + (c + d);
      m2 += fmaf(a, xh.x, b * xh.y) + fmaf(c, xh.z, d * xh.w);
    }
    m1 = tabt_quad_sum(m1) / (float)D; m2 = tabt_quad_sum(m2) / (float)D;
    if (ok) {
      const float rstd = rstd_s[tt];
      for (int k = 4 * sub; k < D; k += 16) {
        const float4 g = *(const float4*)(Gs + tt * ld + k), ga = __ldg((const float4*)(gamma + k)), xh = *(const float4*)(Xh + tt * ld + k);
        const float4 m = tabt_mult4(dr, mask, site, row0 + tt, k, D);
        const float4 ds = make_float4(rstd * (g.x * ga.x - m1 - xh.x * m2), rstd * (g.y * ga.y - m1 - xh.y * m2),
                                      rstd * (g.z * ga.z - m1 - xh.z * m2), rstd * (g.w * ga.w - m1 - xh.w * m2));
        *(float4*)(Gs + tt * ld + k) = ds;
        *(float4*)(G2 + tt * ld + k) = make_float4(ds.x * m.x, ds.y * m.y, ds.z * m.z, ds.w * m.w);
      }
    }
  }
}

// ---- one encoder layer, forward, on the sample held in shared memory ---------------------------------------------------------------
// in: sm.X = layer input.  out: dst (when non-null) = layer output; Xh1, X1, Hb (post ReLU and dropout), Xh2, rstd1/2, lse and the
// keep bits stay in shared memory for the backward phases.
template <int HD>
__device__ __forceinline__ void tabt_layer_fwd(const TabtArgs& a, const TabtSmem& sm, float* S, const float* P, int layer, int64_t b,
                                               const TabtDrop& dr, float* dst, bool keep_for_bwd, int tb = -1) {
  if (tb >= 0) TABT_STAMP(a, tb);
  const int T = a.T, D = a.D, F = a.F, H = a.H;
  const int ldD = tabt_ld(D), ld3 = tabt_ld(3 * D), ldF = tabt_ld(F);
  const int W = (((T + 3) & ~3) + 31) / 32;
  const TabtOff o = tabt_off(D, F);
  const int site = TABT_SITE0 + 4 * layer;
  const int64_t row0 = b * T;
  const size_t lb = (size_t)layer * a.B;
  const uint8_t* m_attn = a.mask_attn ? a.mask_attn + (lb + b) * H * T * T : nullptr;
  const uint8_t* m_res1 = a.mask_res1 ? a.mask_res1 + lb * T * D : nullptr;
  const uint8_t* m_ff = a.mask_ff ? a.mask_ff + lb * T * F : nullptr;
  const uint8_t* m_res2 = a.mask_res2 ? a.mask_res2 + lb * T * D : nullptr;
  float* X = S + sm.X; float* QKV = S + sm.QKV; float* A = S + sm.A; float* Xh1 = S + sm.Xh1; float* X1 = S + sm.X1;
  float* Hb = S + sm.Hb; float* Xh2 = S + sm.Xh2;
  // packed in-projection: [Q | K | V] = X W_in^T + b_in
  // (the Q columns are stored pre-multiplied by log2(e) / sqrt(hd): every attention phase reads them in base-2 score units)
  const float qs = TABT_LOG2E / sqrtf((float)HD);
  tabt_lin_nt<4>(X, ldD, P + o.win, P + o.bin, T, 3 * D, D, [&](int t, int n, float4 v) {
    if (n < D) { v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs; }
    *(float4*)(QKV + t * ld3 + n) = v;
  });
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 1);
  tabt_attn_fwd<HD>(QKV, ld3, T, H, D, 1.0f / sqrtf((float)HD), A, ldD, keep_for_bwd ? S + sm.lse : nullptr,
                    keep_for_bwd ? (uint32_t*)(S + sm.bits) : nullptr, W, dr, m_attn, site + 0, b);
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 2);
  // output projection -> Xh1 (temporary), then X1 = LayerNorm1(X + dropout1(.))
  tabt_lin_nt<2>(A, ldD, P + o.wo, P + o.bo, T, D, D, [&](int t, int n, float4 v) { *(float4*)(Xh1 + t * ldD + n) = v; });
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 3);
  tabt_res_ln(X, Xh1, ldD, T, D, P + o.g1, P + o.be1, dr, m_res1, site + 1, row0, X1, S + sm.rstd1);
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 4);
  // feed-forward: Hb = dropout(relu(X1 W1^T + b1)),  Z = Hb W2^T + b2 -> Xh2 (temporary),  out = LayerNorm2(X1 + dropout2(Z))
  tabt_lin_nt<4>(X1, ldD, P + o.w1, P + o.b1, T, F, D, [&](int t, int n, float4 v) {
    const float4 m = tabt_mult4(dr, m_ff, site + 2, row0 + t, n, F);
    *(float4*)(Hb + t * ldF + n) = make_float4(fmaxf(v.x, 0.f) * m.x, fmaxf(v.y, 0.f) * m.y, fmaxf(v.z, 0.f) * m.z, fmaxf(v.w, 0.f) * m.w);
  });
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 5);
  tabt_lin_nt<2>(Hb, ldF, P + o.w2, P + o.b2, T, D, F, [&](int t, int n, float4 v) { *(float4*)(Xh2 + t * ldD + n) = v; });
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 6);
  tabt_res_ln(X1, Xh2, ldD, T, D, P + o.g2, P + o.be2, dr, m_res2, site + 3, row0, dst, S + sm.rstd2);
  __syncthreads();
  if (tb >= 0) TABT_STAMP(a, tb + 7);
}

__device__ __forceinline__ void tabt_gather(const TabtArgs& a, int64_t b, float* X, int ldD) {
  const int g = a.D >> 2;
  const float* table = a.params + (size_t)a.L * tabt_off(a.D, a.F).size;
  for (int idx = threadIdx.x; idx < a.T * g; idx += TABT_THREADS) {
    const int t = idx / g, k = (idx - t * g) << 2;
    long long row = (long long)__ldg(a.emb_base + t) + __ldg(a.codes + b * a.T + t);
    row = row < 0 ? 0 : (row >= a.n_emb_rows ? a.n_emb_rows - 1 : row);        // memory safety only: torch raises on an invalid code
    *(float4*)(X + t * ldD + k) = __ldg((const float4*)(table + (size_t)row * a.D + k));
  }
}

// the input of layer l of sample b: the embedding gather (layer 0) or the tensor the forward kernel saved (layers >= 1)
__device__ __forceinline__ void tabt_load_input(const TabtArgs& a, int64_t b, int l, float* X, int ldD) {
  if (l == 0) { tabt_gather(a, b, X, ldD); return; }
  const int g4 = a.D >> 2;
  const float* src = a.saved + ((size_t)(l - 1) * a.B + b) * a.T * a.D;
  for (int idx = threadIdx.x; idx < a.T * g4; idx += TABT_THREADS) {
    const int t = idx / g4, k = (idx - t * g4) << 2;
    *(float4*)(X + t * ldD + k) = __ldg((const float4*)(src + t * a.D + k));
  }
}

// ---- forward kernel ---------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(TABT_THREADS, 1) tabt_fwd_kernel(const TabtArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float tabt_sm[];
  float* S = tabt_sm;
  const TabtSmem sm = tabt_smem(a.T, a.D, a.F, a.H, false);
  const int ldD = tabt_ld(a.D), g = a.D >> 2;
  const int lsize = tabt_off(a.D, a.F).size;
  const TabtDrop dr = tabt_drop(a);
  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
    tabt_gather(a, b, S + sm.X, ldD);
    __syncthreads();
    for (int l = 0; l < a.L; ++l) {
      tabt_layer_fwd<HD>(a, sm, S, a.params + (size_t)l * lsize, l, b, dr, S + sm.X, false, (b == blockIdx.x && l == 0) ? 0 : -1);    // the output becomes the next layer's input
      const bool last = l == a.L - 1;
      float* dstg = last ? a.out + b * a.ldo : a.saved + ((size_t)l * a.B + b) * a.T * a.D;
      for (int idx = threadIdx.x; idx < a.T * g; idx += TABT_THREADS) {
        const int t = idx / g, k = (idx - t * g) << 2;
        *(float4*)(dstg + t * a.D + k) = *(const float4*)(S + sm.X + t * ldD + k);
      }
      // no barrier needed: the next layer's first phase only reads X, and its first write (QKV) is followed by one
    }
    __syncthreads();                                                   // the next sample's gather overwrites X
  }
}

// ---- backward kernel --------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(TABT_THREADS, 1) tabt_bwd_kernel(const TabtArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float tabt_sm[];
  float* S = tabt_sm;
  const int T = a.T, D = a.D, F = a.F, H = a.H;
  const TabtSmem sm = tabt_smem(T, D, F, H, true);
  const int ldD = tabt_ld(D), ld3 = tabt_ld(3 * D), ldF = tabt_ld(F), g4 = D >> 2;
  const int W = (((T + 3) & ~3) + 31) / 32;
  const TabtOff o = tabt_off(D, F);
  const TabtDrop dr = tabt_drop(a);
  float* slab = a.slab + (size_t)blockIdx.x * a.p_total;
  float* X = S + sm.X; float* QKV = S + sm.QKV; float* A = S + sm.A; float* Xh1 = S + sm.Xh1; float* X1 = S + sm.X1;
  float* Hb = S + sm.Hb; float* Xh2 = S + sm.Xh2; float* G = S + sm.G; float* G2 = S + sm.G2;      // G2 aliases X (see tabt_smem)
  const float scale = 1.0f / sqrtf((float)HD);
  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
    const int64_t row0 = b * T;
    for (int idx = threadIdx.x; idx < T * g4; idx += TABT_THREADS) {   // gradient of the encoder output
      const int t = idx / g4, k = (idx - t * g4) << 2;
      *(float4*)(G + t * ldD + k) = __ldg((const float4*)(a.dout + b * a.lddo + t * D + k));
    }
    for (int l = a.L - 1; l >= 0; --l) {
      const float* P = a.params + (size_t)l * o.size;
      float* gs = slab + (size_t)l * o.size;
      const int site = TABT_SITE0 + 4 * l;
      const size_t lb = (size_t)l * a.B;
      const uint8_t* m_res1 = a.mask_res1 ? a.mask_res1 + lb * T * D : nullptr;
      const uint8_t* m_res2 = a.mask_res2 ? a.mask_res2 + lb * T * D : nullptr;
      tabt_load_input(a, b, l, X, ldD);
      __syncthreads();
      const int tb = (b == blockIdx.x && l == a.L - 1) ? 0 : -1;          // phase timeline of CTA 0 (tools/tabt_trace.py)
      tabt_layer_fwd<HD>(a, sm, S, P, l, b, dr, nullptr, true, tb);        // recompute: Xh1, X1, Hb, Xh2, rstd, lse, keep bits
      // LayerNorm2: parameter gradients, then G <- dS2 (residual path), G2 <- dZ = dS2 * dropout2
      tabt_ln_param_grads(G, Xh2, ldD, T, D, gs + o.g2, gs + o.be2);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 8);
      tabt_ln_bwd(G, Xh2, ldD, T, D, P + o.g2, S + sm.rstd2, dr, m_res2, site + 3, row0, G2);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 9);
      // linear2: dW2 += dZ^T Hb, db2 += colsum(dZ); then, IN PLACE over Hb, dHpre = (dZ W2) * [Hb > 0] / (1 - p)
      // (Hb > 0 <=> ReLU active AND kept; every element is read and overwritten by the same thread, after all of dW2 is done)
      tabt_grad_tn<2>(G2, ldD, Hb, ldF, T, D, F, gs + o.w2);
      tabt_colsum(G2, ldD, T, D, gs + o.b2);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 10);
      tabt_lin_nn<4>(G2, ldD, P + o.w2, T, D, F, [&](int t, int k, float4 v) {
        const float4 h = *(const float4*)(Hb + t * ldF + k);
        const float s = dr.keep_scale;
        *(float4*)(Hb + t * ldF + k) = make_float4(h.x > 0.f ? v.x * s : 0.f, h.y > 0.f ? v.y * s : 0.f, h.z > 0.f ? v.z * s : 0.f, h.w > 0.f ? v.w * s : 0.f);
      });
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 11);
      // linear1: dW1 += dHpre^T X1, db1 += colsum(dHpre); G += dHpre W1  (G = gradient of X1)
      tabt_grad_tn<2>(Hb, ldF, X1, ldD, T, F, D, gs + o.w1);
      tabt_colsum(Hb, ldF, T, F, gs + o.b1);
      tabt_lin_nn<2>(Hb, ldF, P + o.w1, T, F, D, [&](int t, int k, float4 v) {
        float4 c = *(const float4*)(G + t * ldD + k);
        c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
        *(float4*)(G + t * ldD + k) = c;
      });
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 12);
      // LayerNorm1
      tabt_ln_param_grads(G, Xh1, ldD, T, D, gs + o.g1, gs + o.be1);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 13);
      tabt_ln_bwd(G, Xh1, ldD, T, D, P + o.g1, S + sm.rstd1, dr, m_res1, site + 1, row0, G2);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 14);
      // output projection: dWo += dY^T A, dbo += colsum(dY); dA = dY Wo -> Xh2 (free now).  The layer input comes back from
      // global memory into the X1 buffer (dead since linear1) for the in-projection gradients at the end.
      float* dA = Xh2;
      float* Xr = X1;
      tabt_load_input(a, b, l, Xr, ldD);
      tabt_grad_tn<8>(G2, ldD, A, ldD, T, D, D, gs + o.wo);
      tabt_colsum(G2, ldD, T, D, gs + o.bo);
      tabt_lin_nn<2>(G2, ldD, P + o.wo, T, D, D, [&](int t, int k, float4 v) { *(float4*)(dA + t * ldD + k) = v; });
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 15);
      // attention: dQ (+ delta), then dK and dV, into the Hb region reused as dQKV [T][3D]
      float* dQKV = Hb;
      tabt_attn_bwd_q<HD>(QKV, ld3, A, dA, ldD, T, H, D, scale, S + sm.lse, (const uint32_t*)(S + sm.bits), W, dr.keep_scale, S + sm.delta, dQKV, ld3);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 16);
      tabt_attn_bwd_kv<HD>(QKV, ld3, dA, ldD, T, H, D, scale, S + sm.lse, S + sm.delta, (const uint32_t*)(S + sm.bits), W, dr.keep_scale, dQKV, ld3);
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 17);
      // in-projection: dW_in += dQKV^T X, db_in += colsum(dQKV); G += dQKV W_in  (G = gradient of the layer input)
      tabt_grad_tn<2>(dQKV, ld3, Xr, ldD, T, 3 * D, D, gs + o.win);
      tabt_colsum(dQKV, ld3, T, 3 * D, gs + o.bin);
      tabt_lin_nn<2>(dQKV, ld3, P + o.win, T, 3 * D, D, [&](int t, int k, float4 v) {
        float4 c = *(const float4*)(G + t * ldD + k);
        c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
        *(float4*)(G + t * ldD + k) = c;
      });
      __syncthreads();
      if (tb >= 0) TABT_STAMP(a, 18);
    }
    // embedding rows: every token of a sample hits its own table, so the rows of one sample are distinct
    float* ge = slab + (size_t)a.L * o.size;
    for (int idx = threadIdx.x; idx < T * g4; idx += TABT_THREADS) {
      const int t = idx / g4, k = (idx - t * g4) << 2;
      long long row = (long long)__ldg(a.emb_base + t) + __ldg(a.codes + b * T + t);
      row = row < 0 ? 0 : (row >= a.n_emb_rows ? a.n_emb_rows - 1 : row);
      const float4 v = *(const float4*)(G + t * ldD + k);
      slab_add4(ge + (size_t)row * D + k, v.x, v.y, v.z, v.w);
    }
    __syncthreads();                                                   // G is reloaded for the next sample
  }
}

// dparams[i] = sum over CTAs of slab[cta][i], in CTA order
__global__ void __launch_bounds__(256) tabt_reduce_kernel(const float* __restrict__ slab, int nslab, long long n, float* __restrict__ out) {
  pdl_sync();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < nslab; ++c) {
      const float4 v = __ldcg((const float4*)(slab + (size_t)c * n) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    ((float4*)out)[i] = acc;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------------
namespace {
long long* g_tabt_trace = nullptr;
bool tabt_dev_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
int tabt_check(const fb200_tabt_desc* d) {
  if (!d || d->B < 1 || d->T < 1 || d->D < 4 || d->H < 1 || d->F < 4 || d->L < 1 || d->n_emb_rows < 1) return FB200_EBADARG;
  if (d->D % d->H != 0) return FB200_EBADARG;                  // nn.MultiheadAttention asserts embed_dim % num_heads == 0
  if (d->train && !(d->p >= 0.f && d->p < 1.f)) return FB200_EBADARG;
  const int hd = d->D / d->H;
  if (d->D % 4 != 0 || d->F % 4 != 0 || d->L > 8 || !(hd == 2 || hd == 4 || hd == 8 || hd == 16 || hd == 32)) return FB200_EUNSUPPORTED;
  if ((size_t)tabt_smem(d->T, d->D, d->F, d->H, true).total * sizeof(float) > 227 * 1024) return FB200_EUNSUPPORTED;   // the sample must fit one SM
  return FB200_OK;
}
int tabt_sms(int* sms) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return FB200_ECUDA;
  int maj = 0;
  if (cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return FB200_ECUDA;
  if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return FB200_ECUDA;
  return maj >= 10 ? FB200_OK : FB200_EUNSUPPORTED;
}
TabtArgs tabt_args(const fb200_tabt_desc& d, const int64_t* codes, const int32_t* emb_base, const float* params,
                   const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state) {
  TabtArgs a{};
  a.B = d.B; a.T = d.T; a.D = d.D; a.H = d.H; a.F = d.F; a.L = d.L;
  a.codes = (const long long*)codes; a.emb_base = emb_base; a.n_emb_rows = d.n_emb_rows; a.params = params;
  a.p_total = (long long)d.L * tabt_off(d.D, d.F).size + (long long)d.n_emb_rows * d.D;
  a.train = d.train; a.p = d.p; a.seed = seed; a.offset = offset; a.rng_state = (const uint64_t*)rng_state;
  if (masks) { a.mask_attn = masks[0]; a.mask_res1 = masks[1]; a.mask_ff = masks[2]; a.mask_res2 = masks[3]; }
  a.trace = g_tabt_trace;
  return a;
}
#define TABT_CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; } } while (0)
template <typename K>
int tabt_launch(K kern, int grid, size_t smem, cudaStream_t st, const TabtArgs& a) {
  TABT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // leave the rest of the unified 256 KB array to L1: the layer weights are read through it by every phase
  int carve = (int)(((smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)); if (carve > 100) carve = 100;
  TABT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  TABT_CUDA_OK(pdl_launch(kern, grid, TABT_THREADS, smem, st, a));
  return FB200_OK;
}
#define TABT_DISPATCH(hd, KERN, ...)                                   \
  switch (hd) {                                                        \
    case 2: return tabt_launch(KERN<2>, __VA_ARGS__);                  \
    case 4: return tabt_launch(KERN<4>, __VA_ARGS__);                  \
    case 8: return tabt_launch(KERN<8>, __VA_ARGS__);                  \
    case 16: return tabt_launch(KERN<16>, __VA_ARGS__);                \
    case 32: return tabt_launch(KERN<32>, __VA_ARGS__);                \
    default: return FB200_EUNSUPPORTED;                                \
  }
int tabt_launch_fwd(int hd, int grid, size_t smem, cudaStream_t st, const TabtArgs& a) { TABT_DISPATCH(hd, tabt_fwd_kernel, grid, smem, st, a) }
int tabt_launch_bwd(int hd, int grid, size_t smem, cudaStream_t st, const TabtArgs& a) { TABT_DISPATCH(hd, tabt_bwd_kernel, grid, smem, st, a) }
}  // namespace
}  // namespace fb200

using namespace fb200;

extern "C" {

/* debug: device int64[32] that CTA 0 fills with clock64 stamps after every phase of its first sample (nullptr: off) */
int fb200_debug_tabt_trace(void* buf) { g_tabt_trace = (long long*)buf; return FB200_OK; }

int fb200_tabt_param_elems(const fb200_tabt_desc* d, int64_t* layer_elems, int64_t* total_elems) {
  int rc = tabt_check(d); if (rc != FB200_OK) return rc;
  const int64_t ls = tabt_off(d->D, d->F).size;
  if (layer_elems) *layer_elems = ls;
  if (total_elems) *total_elems = ls * d->L + (int64_t)d->n_emb_rows * d->D;
  return FB200_OK;
}

int fb200_tabt_workspace_bytes(const fb200_tabt_desc* d, size_t* saved_bytes, size_t* bwd_ws_bytes) {
  int rc = tabt_check(d); if (rc != FB200_OK) return rc;
  if (saved_bytes) *saved_bytes = (size_t)(d->L - 1) * d->B * d->T * d->D * sizeof(float) + 256;
  if (bwd_ws_bytes) {
    const size_t ptot = (size_t)d->L * tabt_off(d->D, d->F).size + (size_t)d->n_emb_rows * d->D;
    *bwd_ws_bytes = (size_t)FB200_TABT_MAX_CTAS * ptot * sizeof(float) + 256;
  }
  return FB200_OK;
}

int fb200_tabt_forward(const fb200_tabt_desc* d, const int64_t* codes, const int32_t* emb_base, const float* params,
                       const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                       float* out, int ldo, void* saved, void* stream) {
  int rc = tabt_check(d); if (rc != FB200_OK) return rc;
  if (!codes || !emb_base || !params || !out || (d->L > 1 && !saved) || ldo < d->T * d->D || ldo % 4 != 0) return FB200_EBADARG;
  if (((uintptr_t)params | (uintptr_t)out | (uintptr_t)saved) & 15) return FB200_EBADARG;
  if (!tabt_dev_ptr(codes) || !tabt_dev_ptr(emb_base) || !tabt_dev_ptr(params) || !tabt_dev_ptr(out) || (saved && !tabt_dev_ptr(saved))) return FB200_EUNSUPPORTED;
  if (masks) for (int i = 0; i < 4; ++i) if (masks[i] && !tabt_dev_ptr(masks[i])) return FB200_EUNSUPPORTED;
  int sms = 0; rc = tabt_sms(&sms); if (rc != FB200_OK) return rc;
  TabtArgs a = tabt_args(*d, codes, emb_base, params, masks, seed, offset, rng_state);
  a.out = out; a.ldo = ldo; a.saved = (float*)saved;
  const size_t smem = (size_t)tabt_smem(d->T, d->D, d->F, d->H, false).total * sizeof(float);
  const int grid = d->B < sms ? d->B : sms;
  return tabt_launch_fwd(d->D / d->H, grid, smem, (cudaStream_t)stream, a);
}

int fb200_tabt_backward(const fb200_tabt_desc* d, const int64_t* codes, const int32_t* emb_base, const float* params,
                        const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                        const void* saved, const float* dout, int lddo, float* dparams, void* ws, void* stream) {
  int rc = tabt_check(d); if (rc != FB200_OK) return rc;
  if (!codes || !emb_base || !params || !dout || !dparams || !ws || (d->L > 1 && !saved) || lddo < d->T * d->D || lddo % 4 != 0) return FB200_EBADARG;
  if (((uintptr_t)params | (uintptr_t)dout | (uintptr_t)saved | (uintptr_t)dparams | (uintptr_t)ws) & 15) return FB200_EBADARG;
  if (!tabt_dev_ptr(codes) || !tabt_dev_ptr(emb_base) || !tabt_dev_ptr(params) || !tabt_dev_ptr(dout) || !tabt_dev_ptr(dparams) ||
      !tabt_dev_ptr(ws) || (saved && !tabt_dev_ptr(saved))) return FB200_EUNSUPPORTED;
  if (masks) for (int i = 0; i < 4; ++i) if (masks[i] && !tabt_dev_ptr(masks[i])) return FB200_EUNSUPPORTED;
  int sms = 0; rc = tabt_sms(&sms); if (rc != FB200_OK) return rc;
  TabtArgs a = tabt_args(*d, codes, emb_base, params, masks, seed, offset, rng_state);
  a.saved = (float*)saved; a.dout = dout; a.lddo = lddo; a.slab = (float*)ws;
  int grid = d->B < sms ? d->B : sms; if (grid > FB200_TABT_MAX_CTAS) grid = FB200_TABT_MAX_CTAS;
  cudaStream_t st = (cudaStream_t)stream;
  TABT_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)grid * a.p_total * sizeof(float), st));
  const size_t smem = (size_t)tabt_smem(d->T, d->D, d->F, d->H, true).total * sizeof(float);
  rc = tabt_launch_bwd(d->D / d->H, grid, smem, st, a); if (rc != FB200_OK) return rc;
  int rg = (int)((a.p_total / 4 + 255) / 256); if (rg > 4 * sms) rg = 4 * sms; if (rg < 1) rg = 1;
  TABT_CUDA_OK(pdl_launch(tabt_reduce_kernel, rg, 256, 0, st, (const float*)ws, grid, a.p_total, dparams));
  return FB200_OK;
}

}  // extern "C"
