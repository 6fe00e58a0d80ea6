// attention.cuh - fused multi-head attention core for token sequences (north star: "image tokens
// attending to metadata tokens, softmax held in registers and shared memory").
//
//   nn.MultiheadAttention(embed_dim = D, num_heads = H, batch_first = False) on (S, B, D) tensors, general
//   (S_q, S_kv):   P = softmax(Q_h K_h^T / sqrt(hd)),  O_h = P V_h          (reference call sites with real token
//   sequences: models/multimodalGated.py:118-206; the current model calls it with S = 1, which the head lowers to
//   two GEMMs - plan.cu).  The projections around the core run on the library's GEMM engines (exec.cu).
//
// The probability matrix never exists in memory.  Forward keeps one running (max, sum, output row) per query in
// registers (online softmax) while 32-key tiles of K and V stream through shared memory, and stores only O and the
// row log-sum-exp.  Backward recomputes P from Q, K and the log-sum-exp, flash-attention style, in two passes that
// need no atomics and are therefore bit-reproducible: one CTA per 16 queries produces dQ (and delta = rowsum(dO * O)),
// one CTA per 16 keys produces dK and dV.
//
// Work split inside a CTA (4 warps, 4 rows each): score-like dot products run lanes-over-keys (each lane owns one
// key of the tile and walks the head dimension with 128-bit loads; the tile rows are padded so that the lanes spread
// over all banks) for the warp's four rows at once - every tile element is loaded once and used four times;
// accumulations into a head-dim vector run lanes-over-d with the probabilities broadcast by shuffle.
// Everything is fp32 with expf/logf: the parity bar is 1e-5 against the float64 oracle.
#pragma once
#include "common.cuh"

namespace fb200 {

constexpr int ATT_WARPS = 4;          // warps per CTA
constexpr int ATT_R = 4;              // query (or key) rows per warp
constexpr int ATT_ROWS = ATT_WARPS * ATT_R;
constexpr int ATT_T = 32;             // rows of the streamed tile = one per lane

struct AttnArgs {
  const float* Q; const float* K; const float* V; int ldq, ldk, ldv;   // projected [S*B, D] views: row = s*B + b, head h = columns [h*hd, +hd)
  float* O; int ldo;                                                    // [Sq*B, D] heads concatenated (input of out_proj)
  float* lse;                                                           // [B, H, Sq] row log-sum-exp of the scaled scores
  const float* dO; int lddo;                                            // backward: gradient of O
  float* dQ; float* dK; float* dV; int lddq, lddk, lddv;
  float* delta;                                                         // [B, H, Sq] rowsum(dO * O)
  int Sq, Sk, B, H, hd;
  float scale;                                                          // 1 / sqrt(hd), applied to Q as PyTorch does
};

// Row stride (words) of the streamed tiles.  Lanes-over-keys dot products read one tile row per lane: with hd a
// multiple of 4 the rows are padded to hd + 4 words, so that 128-bit loads of eight consecutive lanes cover the 32
// banks exactly once; otherwise hd + 1 and scalar loads.
__host__ __device__ inline int attn_ldt(int hd) { return (hd % 4 == 0) ? hd + 4 : hd + 1; }
inline size_t attn_smem_bytes(int hd) { return (size_t)(2 * ATT_T * attn_ldt(hd) + 2 * ATT_ROWS * hd + 2 * ATT_T) * sizeof(float); }

// acc[r] += <rows[r], mine> for the warp's ATT_R rows at once: `mine` (this lane's tile row) is loaded ONCE per element
// and used for all four rows; rows[] are warp-wide broadcasts.  Before this blocking every row re-read the tile and the
// kernels were shared-memory bound (ncu r01: 66-77 % l1tex, 26 % FMA pipe).
__device__ __forceinline__ void attn_dot_rows(const float* __restrict__ rows, int row_stride, const float* __restrict__ mine, int hd, float (&acc)[ATT_R]) {
  if ((hd & 3) == 0) {
    for (int d = 0; d < hd; d += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(mine + d);
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        const float4 q4 = *reinterpret_cast<const float4*>(rows + r * row_stride + d);
        acc[r] = fmaf(q4.x, k4.x, fmaf(q4.y, k4.y, fmaf(q4.z, k4.z, fmaf(q4.w, k4.w, acc[r]))));
      }
    }
  } else {
    for (int d = 0; d < hd; ++d) {
      const float k = mine[d];
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) acc[r] = fmaf(rows[r * row_stride + d], k, acc[r]);
    }
  }
}


// Cooperative load of up to `nrows` rows (row index s0 + j < S, else zeros) of head h into shared memory, row stride `lds`
// words, optionally scaled: 128-bit global loads and shared stores when hd is a multiple of 4 (the views are 16-byte
// aligned: D % 4 == 0, hd % 4 == 0), scalar otherwise.
__device__ __forceinline__ void attn_load_rows(float* __restrict__ dst, int lds, const float* __restrict__ src, int ld, int B, int b, int col0,
                                               int s0, int S, int nrows, int hd, float scale) {
  if ((hd & 3) == 0 && (ld & 3) == 0 && (lds & 3) == 0) {
    const int g = hd >> 2;
    for (int idx = threadIdx.x; idx < nrows * g; idx += ATT_WARPS * 32) {
      const int j = idx / g, d = (idx - j * g) << 2, sidx = s0 + j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sidx < S) v = __ldg(reinterpret_cast<const float4*>(src + ((int64_t)sidx * B + b) * ld + col0 + d));
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      *reinterpret_cast<float4*>(dst + j * lds + d) = v;
    }
  } else {
    for (int idx = threadIdx.x; idx < nrows * hd; idx += ATT_WARPS * 32) {
      const int j = idx / hd, d = idx - j * hd, sidx = s0 + j;
      dst[j * lds + d] = sidx < S ? __ldg(src + ((int64_t)sidx * B + b) * ld + col0 + d) * scale : 0.f;
    }
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------
template <int NDL>                    // ceil(hd / 32): head-dim elements per lane
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_fwd_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float att_sm[];
  const int hd = a.hd, ldt = attn_ldt(hd);
  float* Ks = att_sm;                        // [T][ldt]
  float* Vs = Ks + ATT_T * ldt;              // [T][ldt]
  float* qs = Vs + ATT_T * ldt;              // [ROWS][hd], pre-scaled
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_ROWS;
  const int col0 = h * hd;
  attn_load_rows(qs, hd, a.Q, a.ldq, a.B, b, col0, q0, a.Sq, ATT_ROWS, hd, a.scale);
  float m[ATT_R], l[ATT_R], o[ATT_R][NDL];
#pragma unroll
  for (int r = 0; r < ATT_R; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int n = 0; n < NDL; ++n) o[r][n] = 0.f;
  }
  const float* qw = qs + warp * ATT_R * hd;  // this warp's ATT_R query rows (rows past Sq are zeros: computed, never stored)
  for (int k0 = 0; k0 < a.Sk; k0 += ATT_T) {
    __syncthreads();                         // previous tile fully consumed (first pass: qs visible)
    attn_load_rows(Ks, ldt, a.K, a.ldk, a.B, b, col0, k0, a.Sk, ATT_T, hd, 1.f);
    attn_load_rows(Vs, ldt, a.V, a.ldv, a.B, b, col0, k0, a.Sk, ATT_T, hd, 1.f);
    __syncthreads();
    const int nk = min(ATT_T, a.Sk - k0);
    float s[ATT_R], p[ATT_R];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) s[r] = 0.f;
    attn_dot_rows(qw, hd, Ks + lane * ldt, hd, s);
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) {
      if (lane >= nk) s[r] = -INFINITY;
      const float mn = fmaxf(m[r], warp_max(s[r]));
      p[r] = lane < nk ? expf(s[r] - mn) : 0.f;
      const float corr = expf(m[r] - mn);    // exp(-inf) = 0 on the first tile
      l[r] = l[r] * corr + warp_sum(p[r]);
      m[r] = mn;
#pragma unroll
      for (int n = 0; n < NDL; ++n) o[r][n] *= corr;
    }
    for (int jj = 0; jj < nk; ++jj) {
      float v[NDL];
#pragma unroll
      for (int n = 0; n < NDL; ++n) { const int d = lane + 32 * n; v[n] = d < hd ? Vs[jj * ldt + d] : 0.f; }
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        const float pj = __shfl_sync(0xffffffffu, p[r], jj);
#pragma unroll
        for (int n = 0; n < NDL; ++n) o[r][n] = fmaf(pj, v[n], o[r][n]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_R; ++r) {
    const int i = q0 + warp * ATT_R + r;
    if (i >= a.Sq) break;
    const float inv = 1.0f / l[r];
#pragma unroll
    for (int n = 0; n < NDL; ++n) {
      const int d = lane + 32 * n;
      if (d < hd) a.O[((int64_t)i * a.B + b) * a.ldo + col0 + d] = o[r][n] * inv;
    }
    if (lane == 0) a.lse[((int64_t)b * a.H + h) * a.Sq + i] = m[r] + logf(l[r]);
  }
}

// ---- backward, pass 1: dQ and delta (one CTA per 16 queries) -----------------------------------------------------
template <int NDL>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_bwd_dq_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float att_sm[];
  const int hd = a.hd, ldt = attn_ldt(hd);
  float* Ks = att_sm;                        // [T][ldt]
  float* Vs = Ks + ATT_T * ldt;              // [T][ldt]
  float* qs = Vs + ATT_T * ldt;              // [ROWS][hd], pre-scaled
  float* dos = qs + ATT_ROWS * hd;           // [ROWS][hd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_ROWS;
  const int col0 = h * hd;
  attn_load_rows(qs, hd, a.Q, a.ldq, a.B, b, col0, q0, a.Sq, ATT_ROWS, hd, a.scale);
  attn_load_rows(dos, hd, a.dO, a.lddo, a.B, b, col0, q0, a.Sq, ATT_ROWS, hd, 1.f);
  __syncthreads();
  float lse[ATT_R], dl[ATT_R], dq[ATT_R][NDL];
#pragma unroll
  for (int r = 0; r < ATT_R; ++r) {
    const int i = q0 + warp * ATT_R + r;
    lse[r] = 0.f; dl[r] = 0.f;
#pragma unroll
    for (int n = 0; n < NDL; ++n) dq[r][n] = 0.f;
    if (i < a.Sq) {
      float part = 0.f;
#pragma unroll
      for (int n = 0; n < NDL; ++n) {
        const int d = lane + 32 * n;
        if (d < hd) part = fmaf(dos[(warp * ATT_R + r) * hd + d], __ldg(a.O + ((int64_t)i * a.B + b) * a.ldo + col0 + d), part);
      }
      dl[r] = warp_sum(part);
      lse[r] = __ldg(a.lse + ((int64_t)b * a.H + h) * a.Sq + i);
      if (lane == 0) a.delta[((int64_t)b * a.H + h) * a.Sq + i] = dl[r];
    }
  }
  const float* qw = qs + warp * ATT_R * hd;
  const float* gw = dos + warp * ATT_R * hd;
  for (int k0 = 0; k0 < a.Sk; k0 += ATT_T) {
    __syncthreads();
    attn_load_rows(Ks, ldt, a.K, a.ldk, a.B, b, col0, k0, a.Sk, ATT_T, hd, 1.f);
    attn_load_rows(Vs, ldt, a.V, a.ldv, a.B, b, col0, k0, a.Sk, ATT_T, hd, 1.f);
    __syncthreads();
    const int nk = min(ATT_T, a.Sk - k0);
    float s[ATT_R], dp[ATT_R], ds[ATT_R];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) { s[r] = 0.f; dp[r] = 0.f; }
    attn_dot_rows(qw, hd, Ks + lane * ldt, hd, s);
    attn_dot_rows(gw, hd, Vs + lane * ldt, hd, dp);
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) {
      const float p = lane < nk ? expf(s[r] - lse[r]) : 0.f;
      // one key: P == 1 and dS == 0 EXACTLY (PyTorch: P * (dP - sum(dP * P)) = dP - dP), the fact behind the
      // exact-zero W_q / W_k gradients of the reference's S = 1 calls; dp and delta sum in different orders here
      ds[r] = a.Sk == 1 ? 0.f : p * (dp[r] - dl[r]);
    }
    for (int jj = 0; jj < nk; ++jj) {
      float kk[NDL];
#pragma unroll
      for (int n = 0; n < NDL; ++n) { const int d = lane + 32 * n; kk[n] = d < hd ? Ks[jj * ldt + d] : 0.f; }
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        const float dsj = __shfl_sync(0xffffffffu, ds[r], jj);
#pragma unroll
        for (int n = 0; n < NDL; ++n) dq[r][n] = fmaf(dsj, kk[n], dq[r][n]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_R; ++r) {
    const int i = q0 + warp * ATT_R + r;
    if (i >= a.Sq) break;
#pragma unroll
    for (int n = 0; n < NDL; ++n) {
      const int d = lane + 32 * n;
      if (d < hd) a.dQ[((int64_t)i * a.B + b) * a.lddq + col0 + d] = dq[r][n] * a.scale;
    }
  }
}

// ---- backward, pass 2: dK and dV (one CTA per 16 keys; queries stream through shared memory) --------------------
template <int NDL>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_bwd_dkv_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float att_sm[];
  const int hd = a.hd, ldt = attn_ldt(hd);
  float* Qs = att_sm;                        // [T][ldt], pre-scaled
  float* dOs = Qs + ATT_T * ldt;             // [T][ldt]
  float* ks = dOs + ATT_T * ldt;             // [ROWS][hd]
  float* vs = ks + ATT_ROWS * hd;            // [ROWS][hd]
  float* lse_s = vs + ATT_ROWS * hd;         // [T]
  float* dl_s = lse_s + ATT_T;               // [T]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * ATT_ROWS;
  const int col0 = h * hd;
  attn_load_rows(ks, hd, a.K, a.ldk, a.B, b, col0, j0, a.Sk, ATT_ROWS, hd, 1.f);
  attn_load_rows(vs, hd, a.V, a.ldv, a.B, b, col0, j0, a.Sk, ATT_ROWS, hd, 1.f);
  float dk[ATT_R][NDL], dv[ATT_R][NDL];
#pragma unroll
  for (int r = 0; r < ATT_R; ++r)
#pragma unroll
    for (int n = 0; n < NDL; ++n) { dk[r][n] = 0.f; dv[r][n] = 0.f; }
  const float* kw = ks + warp * ATT_R * hd;  // this warp's ATT_R key rows (rows past Sk are zeros: computed, never stored)
  const float* vw = vs + warp * ATT_R * hd;
  for (int i0 = 0; i0 < a.Sq; i0 += ATT_T) {
    __syncthreads();
    attn_load_rows(Qs, ldt, a.Q, a.ldq, a.B, b, col0, i0, a.Sq, ATT_T, hd, a.scale);
    attn_load_rows(dOs, ldt, a.dO, a.lddo, a.B, b, col0, i0, a.Sq, ATT_T, hd, 1.f);
    if (threadIdx.x < ATT_T) {
      const int i = i0 + threadIdx.x;
      lse_s[threadIdx.x] = i < a.Sq ? __ldg(a.lse + ((int64_t)b * a.H + h) * a.Sq + i) : 0.f;
      dl_s[threadIdx.x] = i < a.Sq ? __ldg(a.delta + ((int64_t)b * a.H + h) * a.Sq + i) : 0.f;
    }
    __syncthreads();
    const int nq = min(ATT_T, a.Sq - i0);
    float s[ATT_R], dp[ATT_R], p[ATT_R], ds[ATT_R];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) { s[r] = 0.f; dp[r] = 0.f; }
    attn_dot_rows(kw, hd, Qs + lane * ldt, hd, s);        // lanes over queries
    attn_dot_rows(vw, hd, dOs + lane * ldt, hd, dp);
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) {
      p[r] = lane < nq ? expf(s[r] - lse_s[lane]) : 0.f;
      ds[r] = a.Sk == 1 ? 0.f : p[r] * (dp[r] - dl_s[lane]);
    }
    for (int ii = 0; ii < nq; ++ii) {
      float qi[NDL], gi[NDL];
#pragma unroll
      for (int n = 0; n < NDL; ++n) {
        const int d = lane + 32 * n;
        qi[n] = d < hd ? Qs[ii * ldt + d] : 0.f; gi[n] = d < hd ? dOs[ii * ldt + d] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        const float pi = __shfl_sync(0xffffffffu, p[r], ii);
        const float dsi = __shfl_sync(0xffffffffu, ds[r], ii);
#pragma unroll
        for (int n = 0; n < NDL; ++n) { dv[r][n] = fmaf(pi, gi[n], dv[r][n]); dk[r][n] = fmaf(dsi, qi[n], dk[r][n]); }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_R; ++r) {
    const int j = j0 + warp * ATT_R + r;
    if (j >= a.Sk) break;
#pragma unroll
    for (int n = 0; n < NDL; ++n) {
      const int d = lane + 32 * n;
      if (d < hd) {
        a.dK[((int64_t)j * a.B + b) * a.lddk + col0 + d] = dk[r][n];
        a.dV[((int64_t)j * a.B + b) * a.lddv + col0 + d] = dv[r][n];
      }
    }
  }
}

// ---- launchers -----------------------------------------------------------------------------------------------
#define FB200_ATT_DISPATCH(hd, CALL)                 \
  do {                                               \
    if ((hd) <= 32) { CALL(1); }                     \
    else if ((hd) <= 64) { CALL(2); }                \
    else if ((hd) <= 128) { CALL(4); }               \
    else { CALL(8); }                                \
  } while (0)

template <typename K>
inline cudaError_t attn_set_smem(K kern, size_t bytes) {
  return bytes > 48 * 1024 ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) : cudaSuccess;
}

inline cudaError_t launch_attn_fwd(const AttnArgs& a, cudaStream_t st) {
  if (a.hd < 1 || a.hd > 256) return cudaErrorInvalidValue;
  const size_t smem = attn_smem_bytes(a.hd);
  const dim3 grid((a.Sq + ATT_ROWS - 1) / ATT_ROWS, a.H, a.B);
#define CALL(NDL) do { cudaError_t e = attn_set_smem(attn_fwd_kernel<NDL>, smem); if (e != cudaSuccess) return e; \
                       pdl_launch(attn_fwd_kernel<NDL>, grid, ATT_WARPS * 32, smem, st, a); } while (0)
  FB200_ATT_DISPATCH(a.hd, CALL);
#undef CALL
  return cudaGetLastError();
}

inline cudaError_t launch_attn_bwd(const AttnArgs& a, cudaStream_t st) {
  if (a.hd < 1 || a.hd > 256) return cudaErrorInvalidValue;
  const size_t smem = attn_smem_bytes(a.hd);
  const dim3 gq((a.Sq + ATT_ROWS - 1) / ATT_ROWS, a.H, a.B), gk((a.Sk + ATT_ROWS - 1) / ATT_ROWS, a.H, a.B);
#define CALL(NDL) do { cudaError_t e = attn_set_smem(attn_bwd_dq_kernel<NDL>, smem); if (e != cudaSuccess) return e;  \
                       e = attn_set_smem(attn_bwd_dkv_kernel<NDL>, smem); if (e != cudaSuccess) return e;             \
                       pdl_launch(attn_bwd_dq_kernel<NDL>, gq, ATT_WARPS * 32, smem, st, a);                          \
                       pdl_launch(attn_bwd_dkv_kernel<NDL>, gk, ATT_WARPS * 32, smem, st, a); } while (0)
  FB200_ATT_DISPATCH(a.hd, CALL);
#undef CALL
  return cudaGetLastError();
}

// ---- mean over the query tokens, folded IN FRONT of the output projection ---------------------------------------------
// The reference's sequence variants pool the attention output over its tokens right after the module
// (models/multimodalGated.py:200-205: cross_att.permute(1, 0, 2).mean(dim=1)).  The output projection is linear, so
// mean_s(O_s W_o^T + b_o) = (mean_s O_s) W_o^T + b_o: pooling first turns the S_q*B-row projection (and its two backward
// GEMMs) into B-row ones.  O is [S_q*B, D] with row = s*B + b.
__global__ void __launch_bounds__(256) attn_mean_pool_kernel(const float* __restrict__ O, int Sq, int B, int D, float* __restrict__ pooled) { pdl_sync();
  const int n4 = B * D / 4;
  const float inv = 1.f / (float)Sq;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < Sq; ++s) {                               // token order: bit-reproducible
      const float4 v = __ldg((const float4*)O + (size_t)s * n4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    ((float4*)pooled)[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}
// backward of the pooling: every token row of a batch element receives dPooled / S_q
__global__ void __launch_bounds__(256) attn_mean_pool_bwd_kernel(const float* __restrict__ dpooled, int Sq, int B, int D, float* __restrict__ dO) { pdl_sync();
  const int n4 = B * D / 4;
  const float inv = 1.f / (float)Sq;
  const size_t total = (size_t)Sq * n4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg((const float4*)dpooled + (i % n4));
    ((float4*)dO)[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

}  // namespace fb200
