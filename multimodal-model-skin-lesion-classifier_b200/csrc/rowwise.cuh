// rowwise.cuh - the HBM-bound fused kernels of the fusion head.  One warp owns one row
// (rows are D = 512, D/2 = 256 or F <= 4096 wide), every global access is a 128-bit
// coalesced vector, row statistics are warp-shuffle reductions held in registers, and
// per-column parameter gradients are reduced warp -> CTA (smem) -> one atomic per column.
//
//   ln_relu_drop   : y = dropout(relu(LayerNorm(x)))                fc_fusion[1:4],[5:8]  (multimodalIntraInterModal.py:137-143)
//   gate_mul       : y = sigmoid(z) * x                              img_gate / txt_gate    (:219-229)
//   gated_residual : y = LN(g*drop(a) + (1-g)*q), g = sigmoid(z)     gatedResidualBlock.py:12-17
//   metablock      : y = sigmoid(tanh(v*LN(f)) + LN(g))              metablock.py:22-32
//   cross_entropy  : log-softmax + class-weighted NLL, fwd + dlogits train_pad_20.py:52,111
#pragma once
#include "common.cuh"

namespace fb200 {

constexpr int ROW_WARPS = 8;                 // warps per CTA in the row kernels
constexpr float LN_EPS = 1e-5f;

// A "group" of TPR threads owns one row: TPR = 32 (one warp, rows up to 512 wide) or
// TPR = 256 (the whole CTA, rows up to 4096 wide).  Slot i of thread t covers columns
// (i*TPR + t)*4 .. +3, so a warp-wide access is one contiguous 512-byte segment.
template <int TPR>
struct Grp {
  int t, row0, rstep;
  // bid / nblk: index of this CTA among the CTAs that share the op (blockIdx.x / gridDim.x for a stand-alone launch,
  // tile / tiles when the op is one task of the persistent step kernel - mega.cuh)
  __device__ __forceinline__ Grp(int bid, int nblk) {
    if (TPR == 32) { t = threadIdx.x & 31; row0 = bid * ROW_WARPS + (threadIdx.x >> 5); rstep = nblk * ROW_WARPS; }
    else { t = threadIdx.x; row0 = bid; rstep = nblk; }
  }
  // sums of a and b over the group (all threads of the group must call)
  __device__ __forceinline__ float2 sum2(float a, float b, float2* scratch) const {
    a = warp_sum(a); b = warp_sum(b);
    if (TPR == 32) return make_float2(a, b);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = make_float2(a, b);
    __syncthreads();
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) { float2 v = scratch[w]; r.x += v.x; r.y += v.y; }
    return r;
  }
  // four sums in ONE group reduction (scratch: >= ROW_WARPS float4 = the same 128 bytes viewed as float2[16])
  __device__ __forceinline__ float4 sum4(float a, float b, float c, float d, float2* scratch) const {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
    if (TPR == 32) return make_float4(a, b, c, d);
    float4* s4 = (float4*)scratch;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s4[threadIdx.x >> 5] = make_float4(a, b, c, d);
    __syncthreads();
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) { float4 v = s4[w]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    return r;
  }
};
#define COL(i) (((i) * TPR + G.t) * 4)

// F32 = true: the view is known to be plain fp32 at compile time (the persistent step kernel is fp32 only) - the three-format
// dispatch of ld4 / st4 disappears from the body, which matters for code that runs cold once per step.
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void row_load(const Grp<TPR>& G, const TRef& t, int64_t row, int N, float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = COL(i);
    if (F32) v[i] = (c < N) ? __ldcg((const float4*)((const float*)t.p + row * t.ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    else v[i] = (c < N) ? ld4(t, row, c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void row_store(const Grp<TPR>& G, const TRef& t, int64_t row, int N, const float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = COL(i);
    if (c < N) { if (F32) *(float4*)((float*)t.p + row * t.ld + c) = v[i]; else st4(t, row, c, v[i]); }
  }
}
template <int NV, int TPR>
__device__ __forceinline__ void row_load_param(const Grp<TPR>& G, const float* p, int N, float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = COL(i);
    v[i] = (c < N) ? __ldg((const float4*)(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// mean and 1/sqrt(var+eps) of one row held across the group (biased variance, two-pass)
template <int NV, int TPR>
__device__ __forceinline__ void row_stats(const Grp<TPR>& G, const float4 (&v)[NV], int N, float& mean, float& rstd, float2* scratch) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = G.sum2(s, 0.f, scratch).x / (float)N;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = COL(i);
    if (c < N) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  rstd = rsqrtf(G.sum2(q, 0.f, scratch).x / (float)N + LN_EPS);
}

// Add per-thread column partials (accumulated over the rows this CTA visited) to a global
// fp32 vector.  TPR = 32: the CTA's 8 warps hold partials for the same columns -> reduce in
// smem first, one atomic per column per CTA.  TPR = 256: columns are distinct per thread.
template <int NV, int TPR>
__device__ __forceinline__ void cta_colsum_atomic(const Grp<TPR>& G, const float4 (&part)[NV], int N, float* dst, float* smem /*[ROW_WARPS][512]*/) {
  if (TPR == 32) {
    static_assert(TPR != 32 || NV <= 4, "warp-per-row kernels cover rows up to 512 wide");
    const int warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) *(float4*)(smem + warp * 512 + COL(i)) = part[i];
    __syncthreads();
    for (int c = threadIdx.x; c < 128 * NV && c < N; c += ROW_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) s += smem[w * 512 + c];
      red_add(dst + c, s);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = COL(i);
      if (c < N) { red_add(dst + c, part[i].x); red_add(dst + c + 1, part[i].y); red_add(dst + c + 2, part[i].z); red_add(dst + c + 3, part[i].w); }
    }
  }
}

// ------------------------------------------------------------------ LN + ReLU + dropout
struct LnrdArgs {
  TRef x, y, dy, dx;
  const float *gamma, *beta;
  float *stats;              // [B,2] mean, rstd
  float *dgamma, *dbeta;
  DropSpec drop;
  int B, N;
};

template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void lnrd_fwd_body(const LnrdArgs& a, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  float4 gam[NV], bet[NV];
  row_load_param<NV, TPR>(G, a.gamma, a.N, gam);
  row_load_param<NV, TPR>(G, a.beta, a.N, bet);
  for (int64_t row = G.row0; row < a.B; row += G.rstep) {
    float4 v[NV];
    row_load<NV, TPR, F32>(G, a.x, row, a.N, v);
    float mean, rstd;
    row_stats<NV, TPR>(G, v, a.N, mean, rstd, scratch);
    if (G.t == 0) { a.stats[row * 2] = mean; a.stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = COL(i);
      if (c < a.N) {
        float4 m = drop_mult4(a.drop, row, c, a.N);
        v[i].x = fmaxf((v[i].x - mean) * rstd * gam[i].x + bet[i].x, 0.f) * m.x;
        v[i].y = fmaxf((v[i].y - mean) * rstd * gam[i].y + bet[i].y, 0.f) * m.y;
        v[i].z = fmaxf((v[i].z - mean) * rstd * gam[i].z + bet[i].z, 0.f) * m.z;
        v[i].w = fmaxf((v[i].w - mean) * rstd * gam[i].w + bet[i].w, 0.f) * m.w;
      }
    }
    row_store<NV, TPR, F32>(G, a.y, row, a.N, v);
  }
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) lnrd_fwd_kernel(const LnrdArgs a) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS];
  lnrd_fwd_body<NV, TPR>(a, blockIdx.x, gridDim.x, scratch, nullptr);
}

// LayerNorm backward core for one row: given dyh = d(out)/d(LN output) (already multiplied by
// everything downstream), returns dx in place of `g` and accumulates dgamma/dbeta partials.
template <int NV, int TPR>
__device__ __forceinline__ void ln_bwd_row(const Grp<TPR>& G, float2* scratch, const float4 (&xh)[NV], float4 (&g)[NV], const float4 (&gam)[NV],
                                           float rstd, int N, float4 (&dgam)[NV], float4 (&dbet)[NV]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dgam[i].x += g[i].x * xh[i].x; dgam[i].y += g[i].y * xh[i].y; dgam[i].z += g[i].z * xh[i].z; dgam[i].w += g[i].w * xh[i].w;
    dbet[i].x += g[i].x; dbet[i].y += g[i].y; dbet[i].z += g[i].z; dbet[i].w += g[i].w;
    g[i].x *= gam[i].x; g[i].y *= gam[i].y; g[i].z *= gam[i].z; g[i].w *= gam[i].w;
    s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
  }
  const float2 tot = G.sum2(s1, s2, scratch);
  s1 = tot.x / (float)N;
  s2 = tot.y / (float)N;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i].x = rstd * (g[i].x - s1 - xh[i].x * s2);
    g[i].y = rstd * (g[i].y - s1 - xh[i].y * s2);
    g[i].z = rstd * (g[i].z - s1 - xh[i].z * s2);
    g[i].w = rstd * (g[i].w - s1 - xh[i].w * s2);
  }
}
template <int NV>
__device__ __forceinline__ void zero4(float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// xh = (x - mean) * rstd, zero outside the row
template <int NV, int TPR>
__device__ __forceinline__ void normalize(const Grp<TPR>& G, float4 (&v)[NV], float mean, float rstd, int N) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = COL(i);
    if (c < N) { v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd; v[i].z = (v[i].z - mean) * rstd; v[i].w = (v[i].w - mean) * rstd; }
    else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void lnrd_bwd_body(const LnrdArgs& a, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  float4 gam[NV], dgam[NV], dbet[NV];
  row_load_param<NV, TPR>(G, a.gamma, a.N, gam);
  zero4<NV>(dgam); zero4<NV>(dbet);
  const float scale = a.drop.active ? 1.0f / (1.0f - a.drop.p) : 1.0f;
  for (int64_t row = G.row0; row < a.B; row += G.rstep) {
    float4 xh[NV], yv[NV], g[NV];
    row_load<NV, TPR, F32>(G, a.x, row, a.N, xh);
    row_load<NV, TPR, F32>(G, a.y, row, a.N, yv);
    row_load<NV, TPR, F32>(G, a.dy, row, a.N, g);
    const float mean = a.stats[row * 2], rstd = a.stats[row * 2 + 1];
    normalize<NV, TPR>(G, xh, mean, rstd, a.N);
    // y > 0  <=>  kept by dropout AND ReLU active, so no mask is needed here
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      g[i].x = yv[i].x > 0.f ? g[i].x * scale : 0.f; g[i].y = yv[i].y > 0.f ? g[i].y * scale : 0.f;
      g[i].z = yv[i].z > 0.f ? g[i].z * scale : 0.f; g[i].w = yv[i].w > 0.f ? g[i].w * scale : 0.f;
    }
    ln_bwd_row<NV, TPR>(G, scratch, xh, g, gam, rstd, a.N, dgam, dbet);
    row_store<NV, TPR, F32>(G, a.dx, row, a.N, g);
  }
  cta_colsum_atomic<NV, TPR>(G, dgam, a.N, a.dgamma, red);
  cta_colsum_atomic<NV, TPR>(G, dbet, a.N, a.dbeta, red);
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) lnrd_bwd_kernel(const LnrdArgs a) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS]; __shared__ __align__(16) float red[ROW_WARPS * 512];
  lnrd_bwd_body<NV, TPR>(a, blockIdx.x, gridDim.x, scratch, red);
}

// ------------------------------------------------------------------ gate: y = sigmoid(z) * x
struct GateArgs {
  TRef x, z, y, dy, dz, dx;
  int dx_accumulate;
  int B, N;
};
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void gate_fwd_body(const GateArgs& a, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  for (int64_t row = G.row0; row < a.B; row += G.rstep) {
    float4 x[NV], z[NV];
    row_load<NV, TPR, F32>(G, a.x, row, a.N, x);
    row_load<NV, TPR, F32>(G, a.z, row, a.N, z);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      x[i].x *= 1.f / (1.f + expf(-z[i].x)); x[i].y *= 1.f / (1.f + expf(-z[i].y));
      x[i].z *= 1.f / (1.f + expf(-z[i].z)); x[i].w *= 1.f / (1.f + expf(-z[i].w));
    }
    row_store<NV, TPR, F32>(G, a.y, row, a.N, x);
  }
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) gate_fwd_kernel(const GateArgs a) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS];
  gate_fwd_body<NV, TPR>(a, blockIdx.x, gridDim.x, scratch, nullptr);
}
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void gate_bwd_body(const GateArgs& a, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  for (int64_t row = G.row0; row < a.B; row += G.rstep) {
    float4 x[NV], z[NV], dy[NV], dx[NV];
    row_load<NV, TPR, F32>(G, a.x, row, a.N, x);
    row_load<NV, TPR, F32>(G, a.z, row, a.N, z);
    row_load<NV, TPR, F32>(G, a.dy, row, a.N, dy);
    if (a.dx_accumulate) row_load<NV, TPR, F32>(G, a.dx, row, a.N, dx); else zero4<NV>(dx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#define GATE1(f)                                              \
      { float g = 1.f / (1.f + expf(-z[i].f));               \
        dx[i].f += dy[i].f * g;                               \
        z[i].f = dy[i].f * x[i].f * g * (1.f - g); }
      GATE1(x) GATE1(y) GATE1(z) GATE1(w)
#undef GATE1
    }
    row_store<NV, TPR, F32>(G, a.dz, row, a.N, z);
    row_store<NV, TPR, F32>(G, a.dx, row, a.N, dx);
  }
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) gate_bwd_kernel(const GateArgs a) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS];
  gate_bwd_body<NV, TPR>(a, blockIdx.x, gridDim.x, scratch, nullptr);
}

// ------------------------------------------------------------------ gated residual + LN
struct GrbArgs {
  TRef q, a, z, y;           // fwd: inputs q (residual base), a (attention out), z (gate pre-activation); out y
  TRef dy, da, dz, dq;       // bwd
  int dq_accumulate;
  const float *gamma, *beta;
  float *stats, *dgamma, *dbeta;
  DropSpec drop;
  int B, N;
};
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void grb_fwd_body(const GrbArgs& p, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  float4 gam[NV], bet[NV];
  row_load_param<NV, TPR>(G, p.gamma, p.N, gam);
  row_load_param<NV, TPR>(G, p.beta, p.N, bet);
  for (int64_t row = G.row0; row < p.B; row += G.rstep) {
    float4 q[NV], a[NV], z[NV];
    row_load<NV, TPR, F32>(G, p.q, row, p.N, q);
    row_load<NV, TPR, F32>(G, p.a, row, p.N, a);
    row_load<NV, TPR, F32>(G, p.z, row, p.N, z);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = COL(i);
      float4 m = (c < p.N) ? drop_mult4(p.drop, row, c, p.N) : make_float4(0.f, 0.f, 0.f, 0.f);
#define GRB1(f)                                               \
      { float g = 1.f / (1.f + expf(-z[i].f));               \
        q[i].f = g * (a[i].f * m.f) + (1.f - g) * q[i].f; }
      GRB1(x) GRB1(y) GRB1(z) GRB1(w)
#undef GRB1
    }
    float mean, rstd;
    row_stats<NV, TPR>(G, q, p.N, mean, rstd, scratch);
    if (G.t == 0) { p.stats[row * 2] = mean; p.stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      q[i].x = (q[i].x - mean) * rstd * gam[i].x + bet[i].x; q[i].y = (q[i].y - mean) * rstd * gam[i].y + bet[i].y;
      q[i].z = (q[i].z - mean) * rstd * gam[i].z + bet[i].z; q[i].w = (q[i].w - mean) * rstd * gam[i].w + bet[i].w;
    }
    row_store<NV, TPR, F32>(G, p.y, row, p.N, q);
  }
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) grb_fwd_kernel(const GrbArgs p) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS];
  grb_fwd_body<NV, TPR>(p, blockIdx.x, gridDim.x, scratch, nullptr);
}
// LayerNorm backward for one row with gamma read where it is used (L1-resident) instead of held in registers across the row loop
template <int NV, int TPR>
__device__ __forceinline__ void ln_bwd_row_g(const Grp<TPR>& G, float2* scratch, const float4 (&xh)[NV], float4 (&g)[NV], const float* __restrict__ gamma,
                                             float rstd, int N, float4 (&dgam)[NV], float4 (&dbet)[NV]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = COL(i);
    const float4 gm = (c < N) ? __ldg((const float4*)(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    dgam[i].x += g[i].x * xh[i].x; dgam[i].y += g[i].y * xh[i].y; dgam[i].z += g[i].z * xh[i].z; dgam[i].w += g[i].w * xh[i].w;
    dbet[i].x += g[i].x; dbet[i].y += g[i].y; dbet[i].z += g[i].z; dbet[i].w += g[i].w;
    g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
    s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
  }
  const float2 tot = G.sum2(s1, s2, scratch);
  s1 = tot.x / (float)N;
  s2 = tot.y / (float)N;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i].x = rstd * (g[i].x - s1 - xh[i].x * s2);
    g[i].y = rstd * (g[i].y - s1 - xh[i].y * s2);
    g[i].z = rstd * (g[i].z - s1 - xh[i].z * s2);
    g[i].w = rstd * (g[i].w - s1 - xh[i].w * s2);
  }
}
// r02d: the row needs q, a, z, dy, the recomputed u and dq - with gamma and both parameter-gradient accumulators that was
// ~150 live registers per thread, one 8-warp CTA per SM (ncu: 12 % of the warp slots, 1.0 TB/s).  Only the difference
// w = a*m - q and the gate g are needed after u is formed, gamma is re-read from L1 where it is used: four row vectors live at
// a time, two CTAs per SM.
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void grb_bwd_body(const GrbArgs& p, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  float4 dgam[NV], dbet[NV];
  zero4<NV>(dgam); zero4<NV>(dbet);
  for (int64_t row = G.row0; row < p.B; row += G.rstep) {
    float4 u[NV], w[NV], g[NV], du[NV];
    row_load<NV, TPR, F32>(G, p.q, row, p.N, u);          // q, becomes u
    row_load<NV, TPR, F32>(G, p.a, row, p.N, w);          // a, becomes a*m - q
    row_load<NV, TPR, F32>(G, p.z, row, p.N, g);          // z, becomes the gate
    row_load<NV, TPR, F32>(G, p.dy, row, p.N, du);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = COL(i);
      float4 m = (c < p.N) ? drop_mult4(p.drop, row, c, p.N) : make_float4(0.f, 0.f, 0.f, 0.f);
#define GRB2(f)                                               \
      { const float gg = 1.f / (1.f + expf(-g[i].f));        \
        const float am = w[i].f * m.f, qq = u[i].f;           \
        g[i].f = gg; w[i].f = am - qq;                        \
        u[i].f = gg * am + (1.f - gg) * qq; }
      GRB2(x) GRB2(y) GRB2(z) GRB2(w)
#undef GRB2
    }
    const float mean = p.stats[row * 2], rstd = p.stats[row * 2 + 1];
    normalize<NV, TPR>(G, u, mean, rstd, p.N);
    ln_bwd_row_g<NV, TPR>(G, scratch, u, du, p.gamma, rstd, p.N, dgam, dbet);     // du now holds d(u)
    float4 (&dq)[NV] = u;                                                          // u is dead: its registers take dq
    if (p.dq_accumulate) row_load<NV, TPR, F32>(G, p.dq, row, p.N, dq); else zero4<NV>(dq);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = COL(i);
      float4 m = (c < p.N) ? drop_mult4(p.drop, row, c, p.N) : make_float4(0.f, 0.f, 0.f, 0.f);
#define GRB3(f)                                               \
      { const float gg = g[i].f, d = du[i].f;                 \
        dq[i].f += d * (1.f - gg);                            \
        g[i].f = d * w[i].f * gg * (1.f - gg);                \
        w[i].f = d * gg * m.f; }
      GRB3(x) GRB3(y) GRB3(z) GRB3(w)
#undef GRB3
    }
    row_store<NV, TPR, F32>(G, p.da, row, p.N, w);
    row_store<NV, TPR, F32>(G, p.dz, row, p.N, g);
    row_store<NV, TPR, F32>(G, p.dq, row, p.N, dq);
  }
  cta_colsum_atomic<NV, TPR>(G, dgam, p.N, p.dgamma, red);
  cta_colsum_atomic<NV, TPR>(G, dbet, p.N, p.dbeta, red);
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2) grb_bwd_kernel(const GrbArgs p) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS]; __shared__ __align__(16) float red[ROW_WARPS * 512];
  grb_bwd_body<NV, TPR>(p, blockIdx.x, gridDim.x, scratch, red);
}

// ------------------------------------------------------------------ MetaBlock modulation
struct MetaArgs {
  TRef v, f, g, y;            // v: visual features, f/g: Linear outputs of fb / gb, y: out
  TRef dy, df, dg, dv;        // dv.p == nullptr -> not needed
  int dv_accumulate;
  const float *gamma_f, *beta_f, *gamma_g, *beta_g;
  float *stats;               // [B,4] mean_f, rstd_f, mean_g, rstd_g
  float *dgamma_f, *dbeta_f, *dgamma_g, *dbeta_g;
  int B, N;
};
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void meta_fwd_body(const MetaArgs& p, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  // r02: the four LayerNorm parameter vectors are loaded ONCE per CTA (r01 re-read them for every row), the three row
  // operands are requested together, and the statistics of the two LayerNorms share their group reductions (2 instead of 4
  // rounds of __syncthreads per row for rows wider than 512): at F = 1664, B = 4096 the r01 body ran at 2.3 TB/s.
  float4 gf[NV], bf[NV], gg[NV], bg[NV];
  row_load_param<NV, TPR>(G, p.gamma_f, p.N, gf); row_load_param<NV, TPR>(G, p.beta_f, p.N, bf);
  row_load_param<NV, TPR>(G, p.gamma_g, p.N, gg); row_load_param<NV, TPR>(G, p.beta_g, p.N, bg);
  const float invN = 1.f / (float)p.N;
  for (int64_t row = G.row0; row < p.B; row += G.rstep) {
    float4 f[NV], g[NV], v[NV];
    row_load<NV, TPR, F32>(G, p.f, row, p.N, f);
    row_load<NV, TPR, F32>(G, p.g, row, p.N, g);
    row_load<NV, TPR, F32>(G, p.v, row, p.N, v);
    float sf = 0.f, sg = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { sf += (f[i].x + f[i].y) + (f[i].z + f[i].w); sg += (g[i].x + g[i].y) + (g[i].z + g[i].w); }
    const float2 m = G.sum2(sf, sg, scratch);
    const float mf = m.x * invN, mg = m.y * invN;
    float qf = 0.f, qg = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (COL(i) < p.N) {
        float a = f[i].x - mf, b = f[i].y - mf, c = f[i].z - mf, d = f[i].w - mf; qf += (a * a + b * b) + (c * c + d * d);
        a = g[i].x - mg; b = g[i].y - mg; c = g[i].z - mg; d = g[i].w - mg; qg += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float2 q = G.sum2(qf, qg, scratch);
    const float rf = rsqrtf(q.x * invN + LN_EPS), rg = rsqrtf(q.y * invN + LN_EPS);
    if (G.t == 0) *(float4*)(p.stats + row * 4) = make_float4(mf, rf, mg, rg);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#define MBF(c)                                                                         \
      { const float t1 = (f[i].c - mf) * rf * gf[i].c + bf[i].c;                       \
        const float t2 = (g[i].c - mg) * rg * gg[i].c + bg[i].c;                       \
        f[i].c = 1.f / (1.f + expf(-(tanhf(v[i].c * t1) + t2))); }
      MBF(x) MBF(y) MBF(z) MBF(w)
#undef MBF
    }
    row_store<NV, TPR, F32>(G, p.y, row, p.N, f);
  }
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) meta_fwd_kernel(const MetaArgs p) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS];
  meta_fwd_body<NV, TPR>(p, blockIdx.x, gridDim.x, scratch, nullptr);
}
// Backward in two sweeps per row so that at most ~6 row-vectors are live (F up to 4096).
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void meta_bwd_body_wide(const MetaArgs& p, const int bid, const int nblk, float2* scratch, float* red) {
  const Grp<TPR> G(bid, nblk);
  float4 dgf[NV], dbf[NV], dgg[NV], dbg[NV];
  zero4<NV>(dgf); zero4<NV>(dbf); zero4<NV>(dgg); zero4<NV>(dbg);
  for (int64_t row = G.row0; row < p.B; row += G.rstep) {
    const float4 st = *(const float4*)(p.stats + row * 4);
    float4 xf[NV], ds[NV], v[NV], pr[NV], pb[NV];
    // ds = dy * y (1 - y)
    row_load<NV, TPR, F32>(G, p.y, row, p.N, xf);
    row_load<NV, TPR, F32>(G, p.dy, row, p.N, ds);
#pragma unroll
    for (int i = 0; i < NV; ++i) { ds[i].x *= xf[i].x * (1.f - xf[i].x); ds[i].y *= xf[i].y * (1.f - xf[i].y); ds[i].z *= xf[i].z * (1.f - xf[i].z); ds[i].w *= xf[i].w * (1.f - xf[i].w); }
    // ---- f branch: t1 = LN_f(f); h = tanh(v t1); dt1 = ds (1-h^2) v; dv = ds (1-h^2) t1
    row_load<NV, TPR, F32>(G, p.f, row, p.N, xf);
    normalize<NV, TPR>(G, xf, st.x, st.y, p.N);
    row_load<NV, TPR, F32>(G, p.v, row, p.N, v);
    row_load_param<NV, TPR>(G, p.gamma_f, p.N, pr);
    row_load_param<NV, TPR>(G, p.beta_f, p.N, pb);
    float4 dt1[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#define MB1(c)                                                 \
      { float t1 = xf[i].c * pr[i].c + pb[i].c;               \
        float h = tanhf(v[i].c * t1);                         \
        float dh = ds[i].c * (1.f - h * h);                   \
        dt1[i].c = dh * v[i].c;                               \
        v[i].c = dh * t1; }
      MB1(x) MB1(y) MB1(z) MB1(w)
#undef MB1
    }
    if (p.dv.p) {
      if (p.dv_accumulate) {
        float4 o[NV];
        row_load<NV, TPR, F32>(G, p.dv, row, p.N, o);
#pragma unroll
        for (int i = 0; i < NV; ++i) { v[i].x += o[i].x; v[i].y += o[i].y; v[i].z += o[i].z; v[i].w += o[i].w; }
      }
      row_store<NV, TPR, F32>(G, p.dv, row, p.N, v);
    }
    ln_bwd_row<NV, TPR>(G, scratch, xf, dt1, pr, st.y, p.N, dgf, dbf);
    row_store<NV, TPR, F32>(G, p.df, row, p.N, dt1);
    // ---- g branch: dt2 = ds
    row_load<NV, TPR, F32>(G, p.g, row, p.N, xf);
    normalize<NV, TPR>(G, xf, st.z, st.w, p.N);
    row_load_param<NV, TPR>(G, p.gamma_g, p.N, pr);
    ln_bwd_row<NV, TPR>(G, scratch, xf, ds, pr, st.w, p.N, dgg, dbg);
    row_store<NV, TPR, F32>(G, p.dg, row, p.N, ds);
  }
  cta_colsum_atomic<NV, TPR>(G, dgf, p.N, p.dgamma_f, red);
  cta_colsum_atomic<NV, TPR>(G, dbf, p.N, p.dbeta_f, red);
  cta_colsum_atomic<NV, TPR>(G, dgg, p.N, p.dgamma_g, red);
  cta_colsum_atomic<NV, TPR>(G, dbg, p.N, p.dbeta_g, red);
}
// r02 body for rows up to 2048 wide (NV <= 2): parameters hoisted out of the row loop, every operand of the row requested up
// front, and the two LayerNorm backward passes share ONE four-value group reduction (r01: two sweeps, two reductions, the
// parameters re-read per row - 1.3 TB/s at F = 1664, B = 4096).  Wider rows keep the two-sweep body (register budget).
template <int NV, int TPR, bool F32 = false>
__device__ __forceinline__ void meta_bwd_body(const MetaArgs& p, const int bid, const int nblk, float2* scratch, float* red) {
  if (NV > 2) { meta_bwd_body_wide<NV, TPR, F32>(p, bid, nblk, scratch, red); return; }
  const Grp<TPR> G(bid, nblk);
  float4 gf[NV], bf[NV], gg[NV];
  row_load_param<NV, TPR>(G, p.gamma_f, p.N, gf); row_load_param<NV, TPR>(G, p.beta_f, p.N, bf); row_load_param<NV, TPR>(G, p.gamma_g, p.N, gg);
  float4 dgf[NV], dbf[NV], dgg[NV], dbg[NV];
  zero4<NV>(dgf); zero4<NV>(dbf); zero4<NV>(dgg); zero4<NV>(dbg);
  const float invN = 1.f / (float)p.N;
  for (int64_t row = G.row0; row < p.B; row += G.rstep) {
    const float4 st = *(const float4*)(p.stats + row * 4);
    float4 xf[NV], xg[NV], ds[NV], v[NV], y[NV];
    row_load<NV, TPR, F32>(G, p.y, row, p.N, y);
    row_load<NV, TPR, F32>(G, p.dy, row, p.N, ds);
    row_load<NV, TPR, F32>(G, p.f, row, p.N, xf);
    row_load<NV, TPR, F32>(G, p.g, row, p.N, xg);
    row_load<NV, TPR, F32>(G, p.v, row, p.N, v);
    float4 o[NV];
    if (p.dv.p && p.dv_accumulate) row_load<NV, TPR, F32>(G, p.dv, row, p.N, o); else zero4<NV>(o);
    normalize<NV, TPR>(G, xf, st.x, st.y, p.N);
    normalize<NV, TPR>(G, xg, st.z, st.w, p.N);
    float s1f = 0.f, s2f = 0.f, s1g = 0.f, s2g = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      // ds = dy y (1 - y); t1 = LN_f(f); h = tanh(v t1); dt1 = ds (1 - h^2) v; dv = ds (1 - h^2) t1; dt2 = ds
#define MBB(c)                                                                                   \
      { const float dsv = ds[i].c * y[i].c * (1.f - y[i].c);                                     \
        const float t1 = xf[i].c * gf[i].c + bf[i].c;                                            \
        const float h = tanhf(v[i].c * t1);                                                      \
        const float dh = dsv * (1.f - h * h);                                                    \
        const float d1 = dh * v[i].c;                                                            \
        o[i].c += dh * t1;                                                                       \
        dgf[i].c += d1 * xf[i].c; dbf[i].c += d1; dgg[i].c += dsv * xg[i].c; dbg[i].c += dsv;    \
        const float a1 = d1 * gf[i].c, a2 = dsv * gg[i].c;                                       \
        s1f += a1; s2f += a1 * xf[i].c; s1g += a2; s2g += a2 * xg[i].c;                          \
        y[i].c = a1; ds[i].c = a2; }
      MBB(x) MBB(y) MBB(z) MBB(w)
#undef MBB
    }
    if (p.dv.p) row_store<NV, TPR, F32>(G, p.dv, row, p.N, o);
    float4 t = G.sum4(s1f, s2f, s1g, s2g, scratch);
    t.x *= invN; t.y *= invN; t.z *= invN; t.w *= invN;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#define MBC(c)                                                                                   \
      { y[i].c = st.y * (y[i].c - t.x - xf[i].c * t.y); ds[i].c = st.w * (ds[i].c - t.z - xg[i].c * t.w); }
      MBC(x) MBC(y) MBC(z) MBC(w)
#undef MBC
    }
    row_store<NV, TPR, F32>(G, p.df, row, p.N, y);
    row_store<NV, TPR, F32>(G, p.dg, row, p.N, ds);
  }
  cta_colsum_atomic<NV, TPR>(G, dgf, p.N, p.dgamma_f, red);
  cta_colsum_atomic<NV, TPR>(G, dbf, p.N, p.dbeta_f, red);
  cta_colsum_atomic<NV, TPR>(G, dgg, p.N, p.dgamma_g, red);
  cta_colsum_atomic<NV, TPR>(G, dbg, p.N, p.dbeta_g, red);
}
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) meta_bwd_kernel(const MetaArgs p) { pdl_sync();
  __shared__ __align__(16) float2 scratch[2 * ROW_WARPS]; __shared__ __align__(16) float red[ROW_WARPS * 512];
  meta_bwd_body<NV, TPR>(p, blockIdx.x, gridDim.x, scratch, red);
}

// ------------------------------------------------------------------ cross entropy
// 8 lanes cooperate on one row (C <= 8 is one element per lane; larger C strides).
// Pass 1: per-row log-sum-exp -> unnormalised gradient w[y](softmax - onehot) into dlogits,
//         numerator / denominator reduced by shuffles then two atomics per warp.
// Pass 2: scale by 1/denominator (this batch's, or the global one under data parallelism).
__global__ void __launch_bounds__(256) ce_pass1_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                       const float* __restrict__ class_w, int B, int C,
                                                       float* __restrict__ loss_out, float* __restrict__ dlogits) { pdl_sync();
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  float num = 0.f, den = 0.f;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < (((int64_t)B + 3) & ~3LL); row += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const bool live = row < B;
    const float* z = logits + (live ? row : 0) * C;
    float mx = -INFINITY;
    for (int c = sub; c < C; c += 8) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = sub; c < C; c += 8) se += expf(z[c] - mx);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (live) {
      const int64_t yl = labels[row];
      const bool valid = yl >= 0 && yl < C;                 // anything else is ignored (covers ignore_index = -100; never indexes out of range)
      const int y = valid ? (int)yl : 0;
      const float w = valid ? (class_w ? class_w[y] : 1.f) : 0.f;
      const float inv = 1.f / se;
      if (dlogits) for (int c = sub; c < C; c += 8) dlogits[row * C + c] = w * (expf(z[c] - mx) * inv - (c == y ? 1.f : 0.f));
      if (sub == 0) { num += w * (logf(se) + mx - z[y]); den += w; }
    }
  }
  num = warp_sum(num); den = warp_sum(den);
  if (lane == 0) { red_add(loss_out + 1, num); red_add(loss_out + 2, den); }
}
__global__ void __launch_bounds__(256) ce_pass2_kernel(const float* __restrict__ denom, int n, float* __restrict__ loss_out,
                                                       float* __restrict__ dlogits) { pdl_sync();
  const float den = denom ? *denom : loss_out[2];
  const float inv = 1.f / den;
  if (blockIdx.x == 0 && threadIdx.x == 0) loss_out[0] = loss_out[1] * inv;
  if (dlogits) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dlogits[i] *= inv;
}


// ------------------------------------------------------------------ small-N Linear (the classifier heads)
// y[B,C] = x[B,K] W[C,K]^T + b with C <= 8 (num_classes): a GEMM tile would be >90% padding, so one warp
// owns one row, lanes split K in 128-bit chunks, W (C*K*4 bytes <= 16 KB) is served by L1, and the C dot
// products finish with warp shuffles.  Backward produces dX (coalesced 128-bit stores), dW and db in one pass.
struct SmallNArgs {
  TRef x, y, dy, dx;           // y / dy: [B,C] fp32 (ld = C)
  TRef mask_src;               // dX *= [mask_src > 0] when .p != nullptr
  const float *W, *bias;
  float *dW, *db;
  int dx_accumulate;
  int B, K, C;
};
template <int MAXC>
__device__ __forceinline__ void smalln_fwd_body(const SmallNArgs& a, const int bid, const int nblk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nch = a.K / 4;
  for (int64_t row = (int64_t)bid * ROW_WARPS + warp; row < a.B; row += (int64_t)nblk * ROW_WARPS) {
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
    for (int j = lane; j < nch; j += 32) {
      const float4 xv = ld4(a.x, row, j * 4);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < a.C) {
          const float4 w = __ldg((const float4*)(a.W + (int64_t)c * a.K + j * 4));
          acc[c] += (xv.x * w.x + xv.y * w.y) + (xv.z * w.z + xv.w * w.w);
        }
      }
    }
    float out = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { const float v = warp_sum(acc[c]); if (lane == c) out = v; }
    if (lane < a.C) ((float*)a.y.p)[row * a.y.ld + lane] = out + (a.bias ? __ldg(a.bias + lane) : 0.f);
  }
}
template <int MAXC>
__global__ void __launch_bounds__(ROW_WARPS * 32) smalln_fwd_kernel(const SmallNArgs a) { pdl_sync(); smalln_fwd_body<MAXC>(a, blockIdx.x, gridDim.x); }
template <int MAXC, int CH>      // CH = 128-bit chunks of K per lane (K <= 128 * CH); sm_dw: [C*K + C] floats of shared memory
__device__ __forceinline__ void smalln_bwd_body(const SmallNArgs& a, const int bid, const int nblk, float* sm_dw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nch = a.K / 4;
  float4 dw[MAXC][CH];
  float dbacc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { dbacc[c] = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) dw[c][i] = make_float4(0.f, 0.f, 0.f, 0.f); }
  for (int64_t row = (int64_t)bid * ROW_WARPS + warp; row < a.B; row += (int64_t)nblk * ROW_WARPS) {
    float dl[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { dl[c] = c < a.C ? __ldcg((const float*)a.dy.p + row * a.dy.ld + c) : 0.f; dbacc[c] += dl[c]; }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int j = lane + 32 * i;
      if (j < nch) {
        const float4 xv = ld4(a.x, row, j * 4);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          if (c < a.C) {
            const float4 w = __ldg((const float4*)(a.W + (int64_t)c * a.K + j * 4));
            g.x += dl[c] * w.x; g.y += dl[c] * w.y; g.z += dl[c] * w.z; g.w += dl[c] * w.w;
            dw[c][i].x += dl[c] * xv.x; dw[c][i].y += dl[c] * xv.y; dw[c][i].z += dl[c] * xv.z; dw[c][i].w += dl[c] * xv.w;
          }
        }
        if (a.dx.p) {
          if (a.mask_src.p) { const float4 m = ld4(a.mask_src, row, j * 4); g.x = m.x > 0.f ? g.x : 0.f; g.y = m.y > 0.f ? g.y : 0.f; g.z = m.z > 0.f ? g.z : 0.f; g.w = m.w > 0.f ? g.w : 0.f; }
          if (a.dx_accumulate) { const float4 o = ld4(a.dx, row, j * 4); g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w; }
          st4(a.dx, row, j * 4, g);
        }
      }
    }
  }
  // warp partials -> CTA totals in shared memory (shared atomics) -> one global atomic per element per CTA.
  // (r02e, measured and removed: a shared fp32 atomicAdd is a compare-and-swap loop - ATOMS.CAST.SPIN - so the eight warps were
  // given turns instead, warp 0 storing and the others adding with plain 128-bit read-modify-writes behind a barrier each.
  // Same-box A/B, whole step: +0.7 % at B = 4096, +1.5 % at B = 32 / 256 - eight barriers cost more than the spinning.)
  const int tot = a.C * a.K + a.C;                        // sm_dw: zeroed below
  for (int i = threadIdx.x; i < tot; i += blockDim.x) sm_dw[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < a.C) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int j = lane + 32 * i;
        if (j < nch) {
          float* d = sm_dw + c * a.K + j * 4;
          atomicAdd(d, dw[c][i].x); atomicAdd(d + 1, dw[c][i].y); atomicAdd(d + 2, dw[c][i].z); atomicAdd(d + 3, dw[c][i].w);
        }
      }
      if (lane == 0) atomicAdd(sm_dw + a.C * a.K + c, dbacc[c]);      // every lane saw every row of its warp: lane 0 speaks for the warp
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const float v = sm_dw[i];
    if (i < a.C * a.K) red_add(a.dW + i, v); else red_add(a.db + (i - a.C * a.K), v);
  }
}
template <int MAXC, int CH>
__global__ void __launch_bounds__(ROW_WARPS * 32) smalln_bwd_kernel(const SmallNArgs a) { pdl_sync();
  extern __shared__ __align__(16) float smalln_dyn_smem[];
  smalln_bwd_body<MAXC, CH>(a, blockIdx.x, gridDim.x, smalln_dyn_smem);
}
inline bool smalln_ok(int C, int K) { return C <= 8 && K % 4 == 0 && K <= 512; }
inline cudaError_t launch_smalln_fwd(const SmallNArgs& a, int num_sms, cudaStream_t st) {
  int grid = (a.B + ROW_WARPS - 1) / ROW_WARPS; if (grid > num_sms * 8) grid = num_sms * 8; if (grid < 1) grid = 1;      // one row per warp
  pdl_launch(smalln_fwd_kernel<8>, grid, ROW_WARPS * 32, 0, st, a);
  return cudaGetLastError();
}
// Backward of the small-N Linear for the stand-alone launch (r02e): a THREAD owns one column k of K (KPT columns when K > 256) for
// every class and walks the CTA's rows, so its C weight-gradient partials never meet another thread's - no shared-memory
// reduction at all (the warp-per-row body above ends in 48 shared fp32 atomicAdds per lane, compare-and-swap loops that eight
// warps contend on: 21 us at B = 4096 and 12 us at B = 256 on the critical path where the lanes have joined).  A warp reads and
// writes 128 contiguous bytes of a row per access; the loads of UR rows are issued before the first use; dX rows are stored as
// they are formed; one red.global per weight-gradient element per CTA at the end.
template <int MAXC, int KPT>
__global__ void __launch_bounds__(ROW_WARPS * 32) smalln_bwd_cols_kernel(const SmallNArgs a, const int rows_per_cta) { pdl_sync();
  constexpr int NT = ROW_WARPS * 32, UR = 4;
  const int t = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < (int64_t)a.B ? r0 + rows_per_cta : (int64_t)a.B;
  float w[MAXC][KPT], dw[MAXC][KPT], dbacc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    dbacc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      const int k = t + NT * j;
      w[c][j] = (c < a.C && k < a.K) ? __ldg(a.W + (int64_t)c * a.K + k) : 0.f;
      dw[c][j] = 0.f;
    }
  }
  const bool has_dx = a.dx.p != nullptr, has_mask = has_dx && a.mask_src.p != nullptr, acc_dx = has_dx && a.dx_accumulate != 0;
  for (int64_t r = r0; r < r1; r += UR) {
    float x[UR][KPT], mk[UR][KPT], o[UR][KPT], dl[UR][MAXC];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int64_t row = r + u; const bool ok = row < r1;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) dl[u][c] = (ok && c < a.C) ? __ldcg((const float*)a.dy.p + row * a.dy.ld + c) : 0.f;
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const int k = t + NT * j; const bool live = ok && k < a.K;
        x[u][j] = live ? ld1(a.x, row, k) : 0.f;
        mk[u][j] = (live && has_mask) ? ld1(a.mask_src, row, k) : 1.f;
        o[u][j] = (live && acc_dx) ? ld1(a.dx, row, k) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int64_t row = r + u; const bool ok = row < r1;
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const int k = t + NT * j;
        float g = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) { g += dl[u][c] * w[c][j]; dw[c][j] += dl[u][c] * x[u][j]; }
        g = (mk[u][j] > 0.f ? g : 0.f) + o[u][j];
        if (has_dx && ok && k < a.K) st1(a.dx, row, k, g);
      }
#pragma unroll
      for (int c = 0; c < MAXC; ++c) dbacc[c] += dl[u][c];
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < a.C) {
#pragma unroll
      for (int j = 0; j < KPT; ++j) { const int k = t + NT * j; if (k < a.K) red_add(a.dW + (int64_t)c * a.K + k, dw[c][j]); }
      if (t == c) red_add(a.db + c, dbacc[c]);            // every thread saw every row of the CTA: thread c speaks for class c
    }
  }
}
inline cudaError_t launch_smalln_bwd(const SmallNArgs& a, int num_sms, cudaStream_t st) {
  static const int legacy = [] { const char* e = getenv("FB200_SMALLN_ROWS"); return e ? atoi(e) : 0; }();      // 1: the warp-per-row kernel (A/B runs)
  if (!legacy) {
    // ~2 CTAs per SM at large batches, at least 8 rows (two trips) per CTA: the CTA count is also the number of atomics per address
    int rows = (a.B + 2 * num_sms - 1) / (2 * num_sms); if (rows < 8) rows = 8;
    const int grid = (a.B + rows - 1) / rows;
    if (a.K <= ROW_WARPS * 32) pdl_launch(smalln_bwd_cols_kernel<8, 1>, grid, ROW_WARPS * 32, 0, st, a, rows);
    else pdl_launch(smalln_bwd_cols_kernel<8, 2>, grid, ROW_WARPS * 32, 0, st, a, rows);
    return cudaGetLastError();
  }
  // two rows per warp: each CTA ends with one global atomic per weight-gradient element, so the grid is also the number of
  // atomics every address receives
  int grid = (a.B + ROW_WARPS * 2 - 1) / (ROW_WARPS * 2); if (grid > num_sms * 3) grid = num_sms * 3; if (grid < 1) grid = 1;
  const size_t smem = (size_t)(a.C * a.K + a.C) * sizeof(float);          // <= 8 * 512 * 4 + 32 = 16.4 KB
  if (a.K <= 128) pdl_launch(smalln_bwd_kernel<8, 1>, grid, ROW_WARPS * 32, smem, st, a);
  else if (a.K <= 256) pdl_launch(smalln_bwd_kernel<8, 2>, grid, ROW_WARPS * 32, smem, st, a);
  else pdl_launch(smalln_bwd_kernel<8, 4>, grid, ROW_WARPS * 32, smem, st, a);
  return cudaGetLastError();
}


// ------------------------------------------------------------------ fused Adam (multi-tensor)
// torch.optim.Adam semantics as the reference uses it (Adam(lr=5e-5, weight_decay=1e-4), train_pad_20.py:54):
// coupled L2 (g += wd * p), bias-corrected moments, eps outside the sqrt.  One launch updates up to 48 tensors.
struct AdamSeg { float* p; const float* g; float* m; float* v; int64_t n; };
struct AdamBatch { AdamSeg seg[48]; int nseg; float lr, b1, b2, eps, wd, bc1, bc2, grad_scale; };
__global__ void __launch_bounds__(256) adam_kernel(const AdamBatch a) { pdl_sync();
  const AdamSeg sg = a.seg[blockIdx.y];
  const float step_size = a.lr / a.bc1, inv_sqrt_bc2 = rsqrtf(a.bc2);
  const int64_t n4 = sg.n / 4;
  const bool vec = ((((uintptr_t)sg.p) | ((uintptr_t)sg.g) | ((uintptr_t)sg.m) | ((uintptr_t)sg.v)) & 15) == 0;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g = g * a.grad_scale + a.wd * p;
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
  };
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 p = ((float4*)sg.p)[i], m = ((float4*)sg.m)[i], v = ((float4*)sg.v)[i];
      const float4 g = __ldg((const float4*)sg.g + i);
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      ((float4*)sg.p)[i] = p; ((float4*)sg.m)[i] = m; ((float4*)sg.v)[i] = v;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += (int64_t)gridDim.x * blockDim.x) upd(sg.p[i], sg.g[i], sg.m[i], sg.v[i]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += (int64_t)gridDim.x * blockDim.x) upd(sg.p[i], sg.g[i], sg.m[i], sg.v[i]);
  }
}


// ------------------------------------------------------------------ the other losses of the reference loops
// kind 1: FocalLoss(alpha, gamma, reduction='mean')           (models/focalLoss.py:6-26, train_derm7pt.py:52)
//         ce = lse(z) - z_y ; pt = exp(-ce) ; F = alpha_y (1-pt)^gamma ce ; loss = mean_i F_i
// kind 2: SoftTargetCrossEntropy(weight)                       (models/softtargetsCrossEntropy.py:5-22)
//         loss = mean_i ( - sum_c t_ic w_c log_softmax(z_i)_c )
// One thread-group of 8 lanes per row like the weighted CE; the 1/B of the mean is folded in, so one pass suffices.
__global__ void __launch_bounds__(256) aux_loss_kernel(int kind, const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                       const float* __restrict__ soft, const float* __restrict__ wvec, float gamma,
                                                       int B, int C, float* __restrict__ loss_out, float* __restrict__ dlogits) { pdl_sync();
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const float invB = 1.f / (float)B;
  float acc = 0.f;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < (((int64_t)B + 3) & ~3LL); row += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const bool live = row < B;
    const float* z = logits + (live ? row : 0) * C;
    float mx = -INFINITY;
    for (int c = sub; c < C; c += 8) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = sub; c < C; c += 8) se += expf(z[c] - mx);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = logf(se) + mx, inv = 1.f / se;
    if (kind == 1) {
      const int64_t yl = live ? labels[row] : 0;
      const bool valid = live && yl >= 0 && yl < C;
      const int y = valid ? (int)yl : 0;
      const float ce = lse - z[y];
      const float pt = expf(-ce), om = 1.f - pt;
      const float a = wvec ? wvec[y] : 1.f;
      const float omg = powf(om, gamma);
      const float f = a * omg * ce;
      // dF/dce = a [ (1-pt)^g + ce g (1-pt)^(g-1) pt ]
      const float dfdce = a * (omg + (om > 0.f ? ce * gamma * powf(om, gamma - 1.f) * pt : 0.f));
      if (valid) {
        if (dlogits) for (int c = sub; c < C; c += 8) dlogits[row * C + c] = dfdce * invB * (expf(z[c] - mx) * inv - (c == y ? 1.f : 0.f));
        if (sub == 0) acc += f * invB;
      } else if (live && dlogits) { for (int c = sub; c < C; c += 8) dlogits[row * C + c] = 0.f; }
    } else {
      const float* t = soft + (live ? row : 0) * C;
      float tw = 0.f, part = 0.f;
      for (int c = sub; c < C; c += 8) { const float twc = t[c] * (wvec ? wvec[c] : 1.f); tw += twc; part += twc * (z[c] - lse); }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) { tw += __shfl_xor_sync(0xffffffffu, tw, o); part += __shfl_xor_sync(0xffffffffu, part, o); }
      if (live) {
        if (dlogits) for (int c = sub; c < C; c += 8) dlogits[row * C + c] = invB * (expf(z[c] - mx) * inv * tw - t[c] * (wvec ? wvec[c] : 1.f));
        if (sub == 0) acc -= part * invB;
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red_add(loss_out, acc);
}
// softmax probabilities + argmax for evaluation (utils/model_metrics.py:57-58, utils/save_predictions.py:93-94)
__global__ void __launch_bounds__(256) softmax_argmax_kernel(const float* __restrict__ logits, int B, int C, float* __restrict__ probs, int64_t* __restrict__ pred) { pdl_sync();
  const int lane = threadIdx.x & 31, sub = lane & 7;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < (((int64_t)B + 3) & ~3LL); row += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const bool live = row < B;
    const float* z = logits + (live ? row : 0) * C;
    float mx = -INFINITY; int am = 0x7fffffff;
    for (int c = sub; c < C; c += 8) if (z[c] > mx) { mx = z[c]; am = c; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o); const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }          // first maximum wins, like torch.argmax
    }
    float se = 0.f;
    for (int c = sub; c < C; c += 8) se += expf(z[c] - mx);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (live) {
      if (probs) for (int c = sub; c < C; c += 8) probs[row * C + c] = expf(z[c] - mx) / se;
      if (pred && sub == 0) pred[row] = am;
    }
  }
}

// ------------------------------------------------------------------ metadata one-hot + StandardScaler
// The dense metadata vector the datasets feed the head (models/skinLesionDatasets.py:133-176: np.hstack((OneHotEncoder
// (handle_unknown='ignore').transform(categorical), StandardScaler().transform(numerical))) built on the device from the
// per-column category codes (code < 0 = unknown category -> an all-zero group, like handle_unknown='ignore') and the raw
// numeric columns.  The scaler runs in float64 like scikit-learn and rounds once to fp32 (the dataset's
// torch.tensor(..., dtype=float32)): bit-identical to the reference pipeline.  One thread per output element, coalesced.
__global__ void __launch_bounds__(256) metadata_encode_kernel(const int32_t* __restrict__ codes, const int32_t* __restrict__ col_of,
                                                              const int32_t* __restrict__ col_base, const double* __restrict__ numeric,
                                                              const double* __restrict__ mean, const double* __restrict__ scale,
                                                              int B, int n_cat, int cat_total, int n_num, float* __restrict__ out) { pdl_sync();
  const int V = cat_total + n_num;
  const int64_t total = (int64_t)B * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / V; const int j = (int)(i - b * V);
    float v;
    if (j < cat_total) {
      const int c = __ldg(col_of + j);
      v = (__ldg(codes + b * n_cat + c) == j - __ldg(col_base + c)) ? 1.f : 0.f;
    } else {
      const int q = j - cat_total;
      v = (float)((__ldg(numeric + b * n_num + q) - __ldg(mean + q)) / __ldg(scale + q));
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------ format conversion
// fp32 [rows, cols] (ld_in) -> FMT_PAIR / FMT_BF16 / FMT_F32 copy (ld_out); any alignment.
__global__ void __launch_bounds__(256) convert_kernel(const float* __restrict__ in, int ld_in, TRef out, int64_t rows, int cols) { pdl_sync();
  int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols; int c = (int)(i - r * cols);
    st1(out, r, c, in[r * ld_in + c]);
  }
}

// column sums of a [rows, N] matrix into fp32 dst (atomic): bias gradients for the tcgen05 path
template <int NV, int TPR>
__global__ void __launch_bounds__(ROW_WARPS * 32) colsum_kernel(TRef x, int B, int N, float* dst) { pdl_sync();
  __shared__ __align__(16) float red[ROW_WARPS * 512];
  const Grp<TPR> G(blockIdx.x, gridDim.x);
  float4 acc[NV];
  zero4<NV>(acc);
  for (int64_t row = G.row0; row < B; row += G.rstep) {
    float4 v[NV];
    row_load<NV, TPR>(G, x, row, N, v);
#pragma unroll
    for (int i = 0; i < NV; ++i) { acc[i].x += v[i].x; acc[i].y += v[i].y; acc[i].z += v[i].z; acc[i].w += v[i].w; }
  }
  cta_colsum_atomic<NV, TPR>(G, acc, N, dst, red);
}


// Bias gradients of every tcgen05 Linear of the step in ONE launch: segment = (dY view, N, db).
// Thread t owns 4 columns of its segment and walks the rows with a CTA-wide stride, so each
// access is a coalesced 128-bit load; partial sums go out as fp32 atomics (db is pre-zeroed).
struct ColsumSeg { TRef x; int N; float* dst; };
struct ColsumBatch { ColsumSeg seg[24]; int nseg; int B; };
__global__ void __launch_bounds__(256) colsum_batch_kernel(const ColsumBatch a) { pdl_sync();
  __shared__ __align__(16) float4 red[256];
  const ColsumSeg sg = a.seg[blockIdx.y];
  const int nv = sg.N / 4;                                   // float4 columns
  const int lanes = nv < 256 ? nv : 256;                     // threads across one row (power of two for the widths in use, else rounded down)
  const int groups = 256 / lanes;                            // row groups working side by side
  const int t = threadIdx.x;
  const int r_off = t / lanes, c_lane = t % lanes;
  const bool active = t < lanes * groups;
  // this CTA's slab of rows
  const int64_t rows_per_cta = (a.B + gridDim.x - 1) / gridDim.x;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = r_begin + rows_per_cta < a.B ? r_begin + rows_per_cta : a.B;
  for (int cg0 = 0; cg0 < nv; cg0 += lanes) {
    const int cg = cg0 + c_lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && cg < nv) {
      int64_t r = r_begin + r_off;
      // 4 independent loads in flight per thread
      for (; r + 3 * groups < r_end; r += 4 * groups) {
        const float4 v0 = ld4(sg.x, r, cg * 4), v1 = ld4(sg.x, r + groups, cg * 4), v2 = ld4(sg.x, r + 2 * groups, cg * 4), v3 = ld4(sg.x, r + 3 * groups, cg * 4);
        acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
        acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r_end; r += groups) { const float4 v = ld4(sg.x, r, cg * 4); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    __syncthreads();
    red[t] = acc;
    __syncthreads();
    if (active && r_off == 0 && cg < nv) {
      for (int g2 = 1; g2 < groups; ++g2) { const float4 v = red[g2 * lanes + c_lane]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      float* d = sg.dst + cg * 4;
      red_add(d, acc.x); red_add(d + 1, acc.y); red_add(d + 2, acc.z); red_add(d + 3, acc.w);
    }
  }
}

// ---- launch helpers ------------------------------------------------------------------
inline int row_grid(int B, int num_sms) {
  // ONE row per warp while the rows fit one wave of resident CTAs (8 CTAs of 8 warps per SM): these kernels are a chain of
  // dependent latencies per row (load -> row statistics -> store), so rows in flight are what reaches HBM bandwidth.  r01
  // gave every warp 4 rows in turn: lnrd_fwd<4,32> at B = 4096 ran at 0.20 of the measured HBM peak (profiles/r01 launch list).
  int ctas = (B + ROW_WARPS - 1) / ROW_WARPS;
  int cap = num_sms * 8;
  if (ctas > cap) ctas = cap;
  return ctas < 1 ? 1 : ctas;
}
// dispatch on the row width: warp-per-row up to 512 columns, CTA-per-row up to 4096
#define FB200_ROW_DISPATCH(N, CALL)                       \
  do {                                                    \
    if ((N) <= 128) { CALL(1, 32); }                      \
    else if ((N) <= 256) { CALL(2, 32); }                 \
    else if ((N) <= 512) { CALL(4, 32); }                 \
    else if ((N) <= 1024) { CALL(1, 256); }               \
    else if ((N) <= 2048) { CALL(2, 256); }               \
    else { CALL(4, 256); }                                \
  } while (0)
inline int row_grid_for(int B, int N, int num_sms) {
  if (N <= 512) return row_grid(B, num_sms);
  int cap = num_sms * 8;
  return B < cap ? (B < 1 ? 1 : B) : cap;
}

}  // namespace fb200
