// attention_tc.cuh - tensor-core core of the token attention ("0 tensor-core instructions").
//
// Same contract as attention.cuh (AttnArgs; O and the row log-sum-exp out of forward; dQ + delta, then dK / dV, out of two
// recomputing backward kernels; no atomics), but the four contractions of the core - Q K^T, P V, dO V^T and dS K (and
// their transposes in the key-side pass) - run on the tensor cores with warp-level mma.sync.m16n8k8 TF32 fragments.
//
// Why warp-level MMA and not tcgen05 here: one (batch, head) problem of the reference's sequence models is 197 x 85 x 64
// (multimodalGated.py:118-206) - two ragged 128-row tcgen05 tiles per problem with the softmax sitting between the two
// products; the accumulator would have to travel TMEM -> registers -> shared memory -> TMEM for every tile, and the
// operands need a hi/lo split on the way in (below).  With register fragments the score tile never leaves the warp's
// registers: the C fragment of S becomes the A fragment of P V by a fixed re-labelling of the key index inside each
// group of eight (see attn_tc_pv).  The dense projections around the core stay on tcgen05 (exec.cu).
//
// fp32 parity (1e-5 against the float64 oracle) rules out a single TF32 product (1e-3); every contraction is 3xTF32:
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and a.b ~ hi.hi + lo.hi + hi.lo with fp32 accumulation.  The
// streamed tiles (K, V or Q, dO) are split ONCE per CTA while they are staged into shared memory; the register-resident
// row operands are split on the fly (one split per k-step feeds all n-tiles of the tile).
//
// A CTA is 4 warps x 16 rows.  Forward keeps the running (max, sum) of each row in the four lanes that share the row
// (quad shuffles) - softmax never touches memory.
#pragma once
#include "attention.cuh"

namespace fb200 {

constexpr int ATC_WARPS = 4;
constexpr int ATC_ROWS = ATC_WARPS * 16;      // rows (queries, or keys in the key-side pass) per CTA

__device__ __forceinline__ uint32_t atc_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void atc_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = atc_tf32(x);
  lo = atc_tf32(x - __uint_as_float(hi));
}
// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void atc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void atc_mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  atc_mma(c, alo, bh0, bh1);                 // small terms first
  atc_mma(c, ahi, bl0, bl1);
  atc_mma(c, ahi, bh0, bh1);
}

__host__ __device__ inline int atc_ld(int hd) { return hd + 4; }   // hd % 8 == 0: ld % 32 in {4, 12, 20, 28} -> both fragment patterns are conflict-free

// Cooperative load of `nrows` rows (row s0 + j < S, else zeros) of head columns [col0, col0 + HD) into shared memory,
// scaled and split into tf32 hi / lo planes (row stride ld).
template <int HD>
__device__ __forceinline__ void atc_stage(float* __restrict__ hi, float* __restrict__ lo, int ld, const float* __restrict__ src, int lds, int B, int b,
                                          int col0, int s0, int S, int nrows, float scale) {
  constexpr int G = HD / 4;
  for (int idx = threadIdx.x; idx < nrows * G; idx += ATC_WARPS * 32) {
    const int j = idx / G, d = (idx - j * G) << 2, s = s0 + j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < S) v = __ldg(reinterpret_cast<const float4*>(src + ((int64_t)s * B + b) * lds + col0 + d));
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    atc_split(v.x, h0, l0); atc_split(v.y, h1, l1); atc_split(v.z, h2, l2); atc_split(v.w, h3, l3);
    *reinterpret_cast<uint4*>(hi + j * ld + d) = make_uint4(h0, h1, h2, h3);
    *reinterpret_cast<uint4*>(lo + j * ld + d) = make_uint4(l0, l1, l2, l3);
  }
}

// This warp's 16 rows (row0 + g, row0 + g + 8) of head columns as raw fp32 A fragments: a[kk] = {(g, 8kk+t), (g+8, 8kk+t),
// (g, 8kk+t+4), (g+8, 8kk+t+4)}; rows past S are zeros.
template <int HD>
__device__ __forceinline__ void atc_load_rows(float (&a)[HD / 8][4], const float* __restrict__ src, int lds, int B, int b, int col0, int row0, int S, float scale) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = row0 + g, r1 = r0 + 8;
  const float* p0 = src + ((int64_t)min(r0, S - 1) * B + b) * lds + col0 + t;
  const float* p1 = src + ((int64_t)min(r1, S - 1) * B + b) * lds + col0 + t;
  const float s0 = r0 < S ? scale : 0.f, s1 = r1 < S ? scale : 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 8; ++kk) {
    a[kk][0] = __ldg(p0 + 8 * kk) * s0; a[kk][1] = __ldg(p1 + 8 * kk) * s1;
    a[kk][2] = __ldg(p0 + 8 * kk + 4) * s0; a[kk][3] = __ldg(p1 + 8 * kk + 4) * s1;
  }
}

// acc[nt] += A (16 x HD, raw fp32 fragments) . Tile^T  for the NT 8-row groups of the staged tile: acc[nt][c] is the C
// fragment of rows (g, g+8) x tile rows 8nt + {2t, 2t+1}.
template <int HD, int NT>
__device__ __forceinline__ void atc_scores(float (&acc)[NT][4], const float (&a)[HD / 8][4], const float* __restrict__ thi, const float* __restrict__ tlo, int ld) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < HD / 8; ++kk) {
    uint32_t ahi[4], alo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) atc_split(a[kk][i], ahi[i], alo[i]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int o = (8 * nt + g) * ld + 8 * kk + t;                 // B[k = t][n = g] = Tile[8nt + g][8kk + t]
      atc_mma3(acc[nt], ahi, alo, __float_as_uint(thi[o]), __float_as_uint(thi[o + 4]), __float_as_uint(tlo[o]), __float_as_uint(tlo[o + 4]));
    }
  }
}

// out[nd] += P (16 x 8NT, given as C fragments p[nt]) . Tile  (Tile rows = the 8NT staged rows, columns = head dims).
// The C fragment holds columns (2t, 2t+1) of each 8-group; the A fragment wants k-indices (t, t+4).  The k-index is a dummy
// of the contraction, so tile row 8nt + 2t plays k = t and row 8nt + 2t + 1 plays k = t + 4: a = {c0, c2, c1, c3} and
// B[k = t][n = g] = Tile[8nt + 2t][8nd + g], B[k = t + 4][n = g] = Tile[8nt + 2t + 1][8nd + g].  No shuffles, no memory.
template <int HD, int NT>
__device__ __forceinline__ void atc_pv(float (&out)[HD / 8][4], const float (&p)[NT][4], const float* __restrict__ thi, const float* __restrict__ tlo, int ld) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    uint32_t ahi[4], alo[4];
    atc_split(p[nt][0], ahi[0], alo[0]); atc_split(p[nt][2], ahi[1], alo[1]);
    atc_split(p[nt][1], ahi[2], alo[2]); atc_split(p[nt][3], ahi[3], alo[3]);
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) {
      const int o = (8 * nt + 2 * t) * ld + 8 * nd + g;
      atc_mma3(out[nd], ahi, alo, __float_as_uint(thi[o]), __float_as_uint(thi[o + ld]), __float_as_uint(tlo[o]), __float_as_uint(tlo[o + ld]));
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Store C-fragment rows of a [16 x HD] result: thread holds (g, 8nd + 2t .. +1) and (g + 8, ...)
template <int HD>
__device__ __forceinline__ void atc_store_rows(float* __restrict__ dst, int ldd, int B, int b, int col0, int row0, int S, const float (&o)[HD / 8][4], float m0, float m1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) {
    if (r0 < S) *reinterpret_cast<float2*>(dst + ((int64_t)r0 * B + b) * ldd + col0 + 8 * nd + 2 * t) = make_float2(o[nd][0] * m0, o[nd][1] * m0);
    if (r1 < S) *reinterpret_cast<float2*>(dst + ((int64_t)r1 * B + b) * ldd + col0 + 8 * nd + 2 * t) = make_float2(o[nd][2] * m1, o[nd][3] * m1);
  }
}

template <int HD, int KT> constexpr size_t atc_smem_bytes() { return (size_t)(4 * KT * (HD + 4) + 2 * KT) * sizeof(float); }

// ---- forward ---------------------------------------------------------------------------------------------------------------
template <int HD, int KT>
__global__ void __launch_bounds__(ATC_WARPS * 32) attn_tc_fwd_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float atc_sm[];
  constexpr int LD = HD + 4, NT = KT / 8;
  float* Khi = atc_sm; float* Klo = Khi + KT * LD; float* Vhi = Klo + KT * LD; float* Vlo = Vhi + KT * LD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, col0 = h * HD;
  const int row0 = blockIdx.x * ATC_ROWS + warp * 16;
  float q[HD / 8][4];
  atc_load_rows<HD>(q, a.Q, a.ldq, a.B, b, col0, row0, a.Sq, a.scale);        // pre-scaled, as torch scales q
  float o[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;                   // rows g and g + 8
  for (int k0 = 0; k0 < a.Sk; k0 += KT) {
    __syncthreads();                                                          // previous tile fully consumed
    atc_stage<HD>(Khi, Klo, LD, a.K, a.ldk, a.B, b, col0, k0, a.Sk, KT, 1.f);
    atc_stage<HD>(Vhi, Vlo, LD, a.V, a.ldv, a.B, b, col0, k0, a.Sk, KT, 1.f);
    __syncthreads();
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    atc_scores<HD, NT>(s, q, Khi, Klo, LD);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int key = k0 + 8 * nt + 2 * t;
      if (key >= a.Sk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (key + 1 >= a.Sk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1])); mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));   // finite: every tile holds at least one valid key
    const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);                      // exp(-inf) = 0 on the first tile
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = expf(s[nt][0] - mn0); s[nt][1] = expf(s[nt][1] - mn0);
      s[nt][2] = expf(s[nt][2] - mn1); s[nt][3] = expf(s[nt][3] - mn1);
      ps0 += s[nt][0] + s[nt][1]; ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + quad_sum(ps0); l1 = l1 * c1 + quad_sum(ps1);
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) { o[nd][0] *= c0; o[nd][1] *= c0; o[nd][2] *= c1; o[nd][3] *= c1; }
    atc_pv<HD, NT>(o, s, Vhi, Vlo, LD);
  }
  atc_store_rows<HD>(a.O, a.ldo, a.B, b, col0, row0, a.Sq, o, 1.0f / l0, 1.0f / l1);
  if (t == 0) {
    const int r0 = row0 + g, r1 = r0 + 8;
    if (r0 < a.Sq) a.lse[((int64_t)b * a.H + h) * a.Sq + r0] = m0 + logf(l0);
    if (r1 < a.Sq) a.lse[((int64_t)b * a.H + h) * a.Sq + r1] = m1 + logf(l1);
  }
}

// ---- backward, pass 1: dQ and delta -------------------------------------------------------------------------------------------
template <int HD, int KT>
__global__ void __launch_bounds__(ATC_WARPS * 32) attn_tc_bwd_dq_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float atc_sm[];
  constexpr int LD = HD + 4, NT = KT / 8;
  float* Khi = atc_sm; float* Klo = Khi + KT * LD; float* Vhi = Klo + KT * LD; float* Vlo = Vhi + KT * LD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, col0 = h * HD;
  const int row0 = blockIdx.x * ATC_ROWS + warp * 16;
  const int r0 = row0 + g, r1 = r0 + 8;
  float q[HD / 8][4], go[HD / 8][4];
  atc_load_rows<HD>(q, a.Q, a.ldq, a.B, b, col0, row0, a.Sq, a.scale);
  atc_load_rows<HD>(go, a.dO, a.lddo, a.B, b, col0, row0, a.Sq, 1.f);
  // delta = rowsum(dO * O): this thread's columns of its two rows, then the four lanes of the row
  float d0 = 0.f, d1 = 0.f;
  {
    const float* o0 = a.O + ((int64_t)min(r0, a.Sq - 1) * a.B + b) * a.ldo + col0 + t;
    const float* o1 = a.O + ((int64_t)min(r1, a.Sq - 1) * a.B + b) * a.ldo + col0 + t;
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
      d0 = fmaf(go[kk][0], __ldg(o0 + 8 * kk), fmaf(go[kk][2], __ldg(o0 + 8 * kk + 4), d0));
      d1 = fmaf(go[kk][1], __ldg(o1 + 8 * kk), fmaf(go[kk][3], __ldg(o1 + 8 * kk + 4), d1));
    }
    d0 = quad_sum(d0); d1 = quad_sum(d1);
  }
  const int64_t sbase = ((int64_t)b * a.H + h) * a.Sq;
  const float ls0 = r0 < a.Sq ? __ldg(a.lse + sbase + r0) : 0.f, ls1 = r1 < a.Sq ? __ldg(a.lse + sbase + r1) : 0.f;
  if (t == 0) {
    if (r0 < a.Sq) a.delta[sbase + r0] = d0;
    if (r1 < a.Sq) a.delta[sbase + r1] = d1;
  }
  float dq[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.f;
  for (int k0 = 0; k0 < a.Sk; k0 += KT) {
    __syncthreads();
    atc_stage<HD>(Khi, Klo, LD, a.K, a.ldk, a.B, b, col0, k0, a.Sk, KT, 1.f);
    atc_stage<HD>(Vhi, Vlo, LD, a.V, a.ldv, a.B, b, col0, k0, a.Sk, KT, 1.f);
    __syncthreads();
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f; }
    atc_scores<HD, NT>(s, q, Khi, Klo, LD);
    atc_scores<HD, NT>(dp, go, Vhi, Vlo, LD);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int key = k0 + 8 * nt + 2 * t;
      const bool v0 = key < a.Sk, v1 = key + 1 < a.Sk;
      // one key: P == 1 and dS == 0 EXACTLY (torch: P * (dP - sum(dP * P)) = dP - dP)
      const bool one = a.Sk == 1;
      s[nt][0] = (v0 && !one) ? expf(s[nt][0] - ls0) * (dp[nt][0] - d0) : 0.f;
      s[nt][1] = (v1 && !one) ? expf(s[nt][1] - ls0) * (dp[nt][1] - d0) : 0.f;
      s[nt][2] = (v0 && !one) ? expf(s[nt][2] - ls1) * (dp[nt][2] - d1) : 0.f;
      s[nt][3] = (v1 && !one) ? expf(s[nt][3] - ls1) * (dp[nt][3] - d1) : 0.f;
    }
    atc_pv<HD, NT>(dq, s, Khi, Klo, LD);
  }
  atc_store_rows<HD>(a.dQ, a.lddq, a.B, b, col0, row0, a.Sq, dq, a.scale, a.scale);
}

// ---- backward, pass 2: dK and dV (rows = keys; query tiles stream through shared memory) ----------------------------------
template <int HD, int KT>
__global__ void __launch_bounds__(ATC_WARPS * 32) attn_tc_bwd_dkv_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) float atc_sm[];
  constexpr int LD = HD + 4, NT = KT / 8;
  float* Qhi = atc_sm; float* Qlo = Qhi + KT * LD; float* Ghi = Qlo + KT * LD; float* Glo = Ghi + KT * LD;
  float* lse_s = Glo + KT * LD; float* dl_s = lse_s + KT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, col0 = h * HD;
  const int row0 = blockIdx.x * ATC_ROWS + warp * 16;                         // this warp's 16 keys
  float kf[HD / 8][4], vf[HD / 8][4];
  atc_load_rows<HD>(kf, a.K, a.ldk, a.B, b, col0, row0, a.Sk, 1.f);
  atc_load_rows<HD>(vf, a.V, a.ldv, a.B, b, col0, row0, a.Sk, 1.f);
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) { dk[nd][0] = dk[nd][1] = dk[nd][2] = dk[nd][3] = 0.f; dv[nd][0] = dv[nd][1] = dv[nd][2] = dv[nd][3] = 0.f; }
  const int64_t sbase = ((int64_t)b * a.H + h) * a.Sq;
  for (int i0 = 0; i0 < a.Sq; i0 += KT) {
    __syncthreads();
    atc_stage<HD>(Qhi, Qlo, LD, a.Q, a.ldq, a.B, b, col0, i0, a.Sq, KT, a.scale);
    atc_stage<HD>(Ghi, Glo, LD, a.dO, a.lddo, a.B, b, col0, i0, a.Sq, KT, 1.f);
    if (threadIdx.x < KT) {
      const int i = i0 + threadIdx.x;
      lse_s[threadIdx.x] = i < a.Sq ? __ldg(a.lse + sbase + i) : 0.f;
      dl_s[threadIdx.x] = i < a.Sq ? __ldg(a.delta + sbase + i) : 0.f;
    }
    __syncthreads();
    float st[NT][4], dpt[NT][4];                                              // S^T and dP^T: rows = keys, columns = queries of the tile
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f; dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f; }
    atc_scores<HD, NT>(st, kf, Qhi, Qlo, LD);
    atc_scores<HD, NT>(dpt, vf, Ghi, Glo, LD);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = 8 * nt + 2 * t, i = i0 + c;
      const bool v0 = i < a.Sq, v1 = i + 1 < a.Sq;
      const float la = lse_s[c], lb = lse_s[c + 1], da = dl_s[c], db = dl_s[c + 1];
      const float p0 = v0 ? expf(st[nt][0] - la) : 0.f, p1 = v1 ? expf(st[nt][1] - lb) : 0.f;
      const float p2 = v0 ? expf(st[nt][2] - la) : 0.f, p3 = v1 ? expf(st[nt][3] - lb) : 0.f;
      st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
      const bool one = a.Sk == 1;
      dpt[nt][0] = one ? 0.f : p0 * (dpt[nt][0] - da); dpt[nt][1] = one ? 0.f : p1 * (dpt[nt][1] - db);
      dpt[nt][2] = one ? 0.f : p2 * (dpt[nt][2] - da); dpt[nt][3] = one ? 0.f : p3 * (dpt[nt][3] - db);
    }
    atc_pv<HD, NT>(dv, st, Ghi, Glo, LD);                                     // dV += P^T dO
    atc_pv<HD, NT>(dk, dpt, Qhi, Qlo, LD);                                    // dK += dS^T (scale * Q)
  }
  atc_store_rows<HD>(a.dK, a.lddk, a.B, b, col0, row0, a.Sk, dk, 1.f, 1.f);
  atc_store_rows<HD>(a.dV, a.lddv, a.B, b, col0, row0, a.Sk, dv, 1.f, 1.f);
}

// ---- launchers ----------------------------------------------------------------------------------------------------------------
inline bool attn_tc_ok(const AttnArgs& a) {
  static const bool off = [] { const char* e = getenv("FB200_ATTN_TC"); return e && e[0] == '0'; }();      // A/B measurements
  if (off) return false;
  if (!(a.hd == 16 || a.hd == 32 || a.hd == 64)) return false;
  // 128-bit staging loads and 64-bit fragment stores
  return a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 2 == 0;
}

template <int HD>
inline cudaError_t launch_attn_tc_fwd_hd(const AttnArgs& a, cudaStream_t st) {
  constexpr int KT = 64;
  const size_t smem = atc_smem_bytes<HD, KT>();
  cudaError_t e = attn_set_smem(attn_tc_fwd_kernel<HD, KT>, smem); if (e != cudaSuccess) return e;
  const dim3 grid((a.Sq + ATC_ROWS - 1) / ATC_ROWS, a.H, a.B);
  return pdl_launch(attn_tc_fwd_kernel<HD, KT>, grid, ATC_WARPS * 32, smem, st, a);
}
template <int HD>
inline cudaError_t launch_attn_tc_bwd_hd(const AttnArgs& a, cudaStream_t st) {
  constexpr int KT = 32;
  const size_t smem = atc_smem_bytes<HD, KT>();
  cudaError_t e = attn_set_smem(attn_tc_bwd_dq_kernel<HD, KT>, smem); if (e != cudaSuccess) return e;
  e = attn_set_smem(attn_tc_bwd_dkv_kernel<HD, KT>, smem); if (e != cudaSuccess) return e;
  const dim3 gq((a.Sq + ATC_ROWS - 1) / ATC_ROWS, a.H, a.B), gk((a.Sk + ATC_ROWS - 1) / ATC_ROWS, a.H, a.B);
  e = pdl_launch(attn_tc_bwd_dq_kernel<HD, KT>, gq, ATC_WARPS * 32, smem, st, a); if (e != cudaSuccess) return e;
  return pdl_launch(attn_tc_bwd_dkv_kernel<HD, KT>, gk, ATC_WARPS * 32, smem, st, a);
}
inline cudaError_t launch_attn_tc_fwd(const AttnArgs& a, cudaStream_t st) {
  switch (a.hd) {
    case 16: return launch_attn_tc_fwd_hd<16>(a, st);
    case 32: return launch_attn_tc_fwd_hd<32>(a, st);
    case 64: return launch_attn_tc_fwd_hd<64>(a, st);
    default: return cudaErrorInvalidValue;
  }
}
inline cudaError_t launch_attn_tc_bwd(const AttnArgs& a, cudaStream_t st) {
  switch (a.hd) {
    case 16: return launch_attn_tc_bwd_hd<16>(a, st);
    case 32: return launch_attn_tc_bwd_hd<32>(a, st);
    case 64: return launch_attn_tc_bwd_hd<64>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace fb200
