// mega.cuh - the whole fusion-head pass as ONE persistent cooperative kernel, for the batches the reference actually
// runs (conf/.env.test:2 BATCH_SIZE=32; train_pad_20.py:108-114 is the loop).
//
// At B = 32 a train step of the head is 0.3-0.9 GFLOP over 19-54 MB of weights: 3-8 us at the roofline, i.e. less than
// two kernel launches (SURVEY.md 7 "hard parts").  As ~40 separate launches it costs ~9 us per launch-bound kernel.  Here
// the op program of plan.cu (forward, cross entropy, backward) is flattened into STAGES of independent tile tasks; one
// CTA per SM walks the stages, a grid barrier (one release-add and an acquire-poll on an L2 counter, ~0.5 us) separates
// them, and nothing else ever touches the host or the launch path:
//
//   * Linear forward   y[B,N]  = x W^T + b   : task = 32 rows x 4 output columns, full K; warp = 4 rows, lanes split K in
//   * Linear dX        dx[B,K] = dy W          128-bit chunks (x/dy straight from L2 with ld.global.cg, W through L1), the 16
//                                              partial sums of a lane meet in a 16-shuffle transpose-reduction; no shared
//                                              memory, no atomics, bit-reproducible.  K > 1024 splits into K-slices whose
//                                              partial tiles a REDUCE task of the next stage sums in fixed order.
//   * Linear dW        dW[N,K] = dy^T x       : task = 32 x 128 outputs, the batch is the (short) reduction loop; db rides along.
//   * the row ops (LayerNorm+ReLU+dropout, gates, gated residual, MetaBlock, classifier head, cross entropy) are the
//     SAME device bodies the stand-alone kernels of rowwise.cuh run, one row per warp.
//   Independent chains (image / metadata lanes, weight gradients) share stages: a stage's tasks are dealt round-robin to
//   the CTAs.  Weights stay L2-resident across steps (18 MB of 126 MB); activations never leave L2.
// fp32 only (exact FFMA arithmetic, like the FFMA engine it replaces at these sizes); B <= MEGA_MAX_B.
#pragma once
#include "common.cuh"
#include "simt_gemm.cuh"
#include "rowwise.cuh"
#include <vector>
#include <algorithm>
#include <mutex>
#include <cstdio>

namespace fb200 {

constexpr int MEGA_THREADS = ROW_WARPS * 32;      // 256: the row bodies are written for 8 warps
constexpr int MEGA_MAX_B = 64;
constexpr int MEGA_MAX_GEMM = 128, MEGA_MAX_ROW = 26, MEGA_MAX_STAGES = 80;
constexpr int MEGA_SPLIT_K = 1024;                // forward Linears with K above this are split into K-slices of <= 512

struct MRef { float* p; int ld; int pad_; };

enum MGemmLayout : int { MG_NT = 0, MG_NN = 1, MG_TN = 2 };
struct MGemm {                 // one GEMM op of a stage (fp32, row-major views)
  MRef A, B, C, mask;          // mask.p: multiply the result by [mask > 0] (ReLU backward of the producer)
  const float* bias;           // NT only
  float* colsum;               // TN only: colsum[m] += sum_k A(k, m)   (bias gradient)
  int M, N, K;
  short layout, relu, accumulate, splits;
  int tiles;
  int stage;
};

enum MRowKind : int { MR_LNRD_FWD = 0, MR_LNRD_BWD, MR_GATE_FWD, MR_GATE_BWD, MR_GRB_FWD, MR_GRB_BWD, MR_META_FWD, MR_META_BWD,
                      MR_SMALLN_FWD, MR_SMALLN_BWD, MR_CE, MR_REDUCE };
struct CeArgs { const float* logits; const int64_t* labels; const float* class_w; const float* denom; float* loss_out; float* dlogits; int B, C;
                const float* den_local; /* optional: this batch's sum of w[y], precomputed (large batches) */ };
struct ReduceArgs { const float* part; float* y; const float* bias; int splits, M, N, ldy, relu; };
struct MRowOp {
  int kind, tiles, stage;
  int chain;                   // 1: runs in the SAME task as the previous row op, right after it (same rows, one __syncthreads between)
  union U { LnrdArgs lnrd; GateArgs gate; GrbArgs grb; MetaArgs meta; SmallNArgs smalln; CeArgs ce; ReduceArgs red; } u;
};
struct MStage { unsigned short g0, g1, r0, r1; };

struct MegaProg {
  MStage st[MEGA_MAX_STAGES];
  MRowOp r[MEGA_MAX_ROW];
  MGemm g[MEGA_MAX_GEMM];
  unsigned* barrier;           // 256 bytes, zero at launch: [0] arrival counter, [32] epoch flag
  long long* trace;            // debug: clock64 of CTA 0 at kernel entry ([0]) and after every stage ([1 + s]); nullptr = off
  int nstages;
  int ngemm, nrow;             // ops in use (the kernel copies only those to shared memory)
  int pad_;
};
static_assert(sizeof(MegaProg) <= 32000, "kernel parameters are limited to 32764 bytes");

// ----------------------------------------------------------------------------- grid barrier
// One L2 counter, monotonic over the launch (zero at launch): arrive with a release-add, poll with acquire loads until all
// G CTAs of this epoch are in.  Measured on B200 with empty stages: 0.93 us per barrier at one CTA per SM (148 arrivals);
// two CTAs per SM cost 1.4 us, and a separate epoch flag polled with relaxed loads + nanosleep 2.0 us.
// The kernel is launched cooperatively (all CTAs co-resident, or the launch fails), and the wait is bounded: a protocol
// bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mega_grid_barrier(unsigned* bar, unsigned epoch, unsigned G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned target = epoch * G;
    unsigned v;
    const long long t0 = clock64();
    for (unsigned n = 1;; ++n) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if ((n & 255u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------- 32-value transpose-reduction
// Every lane holds 32 partial sums v[i]; afterwards lane L holds in v[0] the total of index L
// (recursive halving: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float mega_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return v[0];
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

constexpr int MEGA_CT = 8;                         // output columns per forward / dX task (one 32-byte sector of a W row in dX)
constexpr int MEGA_NN_SLAB = 1024;                 // dX: reduction rows of W staged in shared memory at a time (32 KB)

// ----------------------------------------------------------------------------- Linear forward (NT)
// C[m, n] = sum_k A[m, k] B[n, k] (+ bias, ReLU).  Task = (k-slice, row group of 32, 8 columns); warp = 4 rows, lanes
// split K in 128-bit chunks.
template <bool VEC>
__device__ __forceinline__ void mega_gemm_nt(const MGemm& g, int tile, long long* tr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = (g.N + MEGA_CT - 1) / MEGA_CT, rgs = (g.M + 31) >> 5;
  const int c_idx = tile % ct; tile /= ct;
  const int rg = tile % rgs; const int kz = tile / rgs;
  const int n0 = c_idx * MEGA_CT, m0 = rg * 32 + warp * 4;
  int k0 = 0, k1 = g.K;
  if (g.splits > 1) { const int per = (((g.K + g.splits - 1) / g.splits) + 127) & ~127; k0 = kz * per; k1 = min(g.K, k0 + per); }
  if (m0 >= g.M) return;                                         // warp-uniform: this warp has no rows
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  // One pointer per row of A / B, advanced once per k-chunk (rows / columns past the edge alias a valid one and are never
  // stored): the loop body is the loads and the FMAs - address arithmetic recomputed per load made this loop issue-bound.
  const float* ap[4]; const float* bp[MEGA_CT];
  const int kfirst = k0 + lane * (VEC ? 4 : 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) ap[i] = g.A.p + (int64_t)(m0 + i < g.M ? m0 + i : m0) * g.A.ld + kfirst;
#pragma unroll
  for (int c = 0; c < MEGA_CT; ++c) bp[c] = g.B.p + (int64_t)(n0 + c < g.N ? n0 + c : n0) * g.B.ld + kfirst;
  if (VEC) {
    for (int k = kfirst; k < k1; k += 128) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = __ldcg((const float4*)ap[i]); ap[i] += 128; }
#pragma unroll
      for (int h = 0; h < MEGA_CT; h += 4) {                     // four columns at a time: 16 + 16 + 32 live registers
        float4 b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { b[c] = __ldg((const float4*)bp[h + c]); bp[h + c] += 128; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i * MEGA_CT + h + c] = dot4(a[i], b[c], acc[i * MEGA_CT + h + c]);
      }
    }
  } else {
#pragma unroll 2
    for (int k = kfirst; k < k1; k += 32) {
      float a[4], b[MEGA_CT];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = __ldcg(ap[i]); ap[i] += 32; }
#pragma unroll
      for (int c = 0; c < MEGA_CT; ++c) { b[c] = __ldg(bp[c]); bp[c] += 32; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < MEGA_CT; ++c) acc[i * MEGA_CT + c] = fmaf(a[i], b[c], acc[i * MEGA_CT + c]);
    }
  }
  if (tr && threadIdx.x == 0) tr[1] = clock64();
  float v = mega_reduce32(acc, lane);
  if (tr && threadIdx.x == 0) tr[2] = clock64();
  const int m = m0 + (lane >> 3), n = n0 + (lane & 7);
  if (m < g.M && n < g.N) {
    if (g.splits > 1) { g.C.p[((int64_t)kz * g.M + m) * g.C.ld + n] = v; return; }      // partial tile: bias / ReLU belong to the REDUCE task
    if (g.bias) v += __ldg(g.bias + n);
    if (g.relu) v = fmaxf(v, 0.f);
    g.C.p[(int64_t)m * g.C.ld + n] = v;
  }
}

// ----------------------------------------------------------------------------- Linear dX (NN)
// C[m, n] (+)= sum_k A[m, k] B[k, n], optionally masked by [mask > 0].  Task = (row group of 32, 8 columns).  The column
// slice B[:, n0 .. n0+7] (one 32-byte sector per row of W) is staged in shared memory once per CTA - read per warp it is
// 32 sectors per load instruction, eight times over; then warp = 4 rows, lane = k (stride 32), W rows from shared memory.
template <bool VEC>
__device__ __forceinline__ void mega_gemm_nn(const MGemm& g, int tile, float* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = (g.N + MEGA_CT - 1) / MEGA_CT;
  const int c_idx = tile % ct, rg = tile / ct;
  const int n0 = c_idx * MEGA_CT, m0 = rg * 32 + warp * 4;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const float* ap[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ap[i] = g.A.p + (int64_t)(m0 + i < g.M ? m0 + i : (m0 < g.M ? m0 : 0)) * g.A.ld + lane;
  for (int ks = 0; ks < g.K; ks += MEGA_NN_SLAB) {
    const int kn = min(MEGA_NN_SLAB, g.K - ks);
    __syncthreads();                                              // the previous slab / task is done with the staging buffer
    if (VEC) {                                                    // N % 8 == 0, 16-byte aligned rows: two 128-bit loads per row
      for (int i = threadIdx.x; i < kn * 2; i += MEGA_THREADS) {
        const int k = i >> 1, h = i & 1;
        *(float4*)(smem + k * MEGA_CT + h * 4) = __ldg((const float4*)(g.B.p + (int64_t)(ks + k) * g.B.ld + n0 + h * 4));
      }
    } else {
      for (int i = threadIdx.x; i < kn * MEGA_CT; i += MEGA_THREADS) {
        const int k = i / MEGA_CT, c = i % MEGA_CT;
        smem[i] = (n0 + c < g.N) ? __ldg(g.B.p + (int64_t)(ks + k) * g.B.ld + n0 + c) : 0.f;
      }
    }
    __syncthreads();
    if (m0 < g.M) {
#pragma unroll 4
      for (int k = lane; k < kn; k += 32) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = __ldcg(ap[i] + ks + k - lane);
        const float4 w0 = *(const float4*)(smem + k * MEGA_CT), w1 = *(const float4*)(smem + k * MEGA_CT + 4);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < MEGA_CT; ++c) acc[i * MEGA_CT + c] = fmaf(a[i], w[c], acc[i * MEGA_CT + c]);
      }
    }
  }
  if (m0 >= g.M) return;
  float v = mega_reduce32(acc, lane);
  const int m = m0 + (lane >> 3), n = n0 + (lane & 7);
  if (m < g.M && n < g.N) {
    if (g.mask.p && !(__ldcg(g.mask.p + (int64_t)m * g.mask.ld + n) > 0.f)) v = 0.f;
    float* dst = g.C.p + (int64_t)m * g.C.ld + n;
    if (g.accumulate) v += __ldcg(dst);
    *dst = v;
  }
}

// ----------------------------------------------------------------------------- Linear dW (TN)
// C[m, n] (+)= sum_k A[k, m] B[k, n] with k over the batch rows (<= 64); colsum[m] += sum_k A[k, m].  Task = 32 x 128
// outputs.  The two operand tiles (A[:, m0..+31], B[:, n0..+127]) are staged in shared memory with one round trip to L2,
// then warp = 4 rows m, lane = 4 columns n, the batch is the reduction loop over shared memory.
constexpr int MEGA_TN_LDA = 36, MEGA_TN_LDB = 132;               // padded rows: conflict-free 128-bit reads
constexpr int MEGA_TN_CPT = 4;                                   // 128-column chunks per task (vector path): the next chunk is prefetched under the current one
template <bool VEC>
__device__ __forceinline__ void mega_gemm_tn(const MGemm& g, int tile, float* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (g.N + 127) >> 7;
  const int cpt = VEC ? g.splits : 1;                             // the builder sets splits = chunks per task for TN ops
  const int ntask_n = (nt + cpt - 1) / cpt;
  const int n_task = tile % ntask_n, m_idx = tile / ntask_n;
  const int m0 = m_idx * 32;
  const int ch0 = n_task * cpt, ch1 = min(nt, ch0 + cpt);
  float* As = smem;                                               // [K][36]
  float* Bs = smem + MEGA_MAX_B * MEGA_TN_LDA;                    // [K][132]
  const int mw = warp * 4, nl = lane * 4;
  __syncthreads();
  if (VEC) {                                                      // M % 4 == 0, N % 4 == 0, aligned: whole float4s are in or out
    for (int i = threadIdx.x; i < g.K * 8; i += MEGA_THREADS) {
      const int k = i >> 3, c = (i & 7) * 4;
      *(float4*)(As + k * MEGA_TN_LDA + c) = (m0 + c < g.M) ? __ldcg((const float4*)(g.A.p + (int64_t)k * g.A.ld + m0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < g.K * 32; i += MEGA_THREADS) {
      const int k = i >> 5, c = (i & 31) * 4, n = ch0 * 128 + c;
      *(float4*)(Bs + k * MEGA_TN_LDB + c) = (n < g.N) ? __ldcg((const float4*)(g.B.p + (int64_t)k * g.B.ld + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int i = threadIdx.x; i < g.K * 32; i += MEGA_THREADS) {
      const int k = i >> 5, c = i & 31;
      As[k * MEGA_TN_LDA + c] = (m0 + c < g.M) ? __ldcg(g.A.p + (int64_t)k * g.A.ld + m0 + c) : 0.f;
    }
    for (int i = threadIdx.x; i < g.K * 128; i += MEGA_THREADS) {
      const int k = i >> 7, c = i & 127, n = ch0 * 128 + c;
      Bs[k * MEGA_TN_LDB + c] = (n < g.N) ? __ldcg(g.B.p + (int64_t)k * g.B.ld + n) : 0.f;
    }
  }
  __syncthreads();
  for (int ch = ch0; ch < ch1; ++ch) {
    const int n0 = ch * 128;
    // prefetch the next chunk's B tile into registers while this one is consumed from shared memory (K <= 64: <= 8 vectors per thread)
    float4 pre[MEGA_MAX_B / 8];
    const bool more = VEC && ch + 1 < ch1;
    if (more) {
#pragma unroll
      for (int q = 0; q < MEGA_MAX_B / 8; ++q) {
        const int i = threadIdx.x + q * MEGA_THREADS;
        const int k = i >> 5, n = (ch + 1) * 128 + (i & 31) * 4;
        pre[q] = (k < g.K && n < g.N) ? __ldcg((const float4*)(g.B.p + (int64_t)k * g.B.ld + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float acc[4][4], cs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { cs[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < g.K; ++k) {
      const float4 a = *(const float4*)(As + k * MEGA_TN_LDA + mw);
      const float4 b = *(const float4*)(Bs + k * MEGA_TN_LDB + nl);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { cs[i] += av[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(av[i], bv[c], acc[i][c]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + mw + i, n = n0 + nl;
      if (m >= g.M) break;
      float* dst = g.C.p + (int64_t)m * g.C.ld + n;
      if (VEC) {
        if (n < g.N) {
          float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          if (g.accumulate) { const float4 p = __ldcg((const float4*)dst); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
          *(float4*)dst = o;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < g.N) { float o = acc[i][c]; if (g.accumulate) o += __ldcg(dst + c); dst[c] = o; }
      }
      if (g.colsum && ch == 0 && lane == 0) red_add(g.colsum + m, cs[i]);
    }
    if (more) {
      __syncthreads();                                            // every warp is done reading the current B tile
#pragma unroll
      for (int q = 0; q < MEGA_MAX_B / 8; ++q) {
        const int i = threadIdx.x + q * MEGA_THREADS;
        const int k = i >> 5;
        if (k < g.K) *(float4*)(Bs + k * MEGA_TN_LDB + (i & 31) * 4) = pre[q];
      }
      __syncthreads();
    }
  }
}

// ----------------------------------------------------------------------------- Linear forward / dX, K split across the warps
// The first version gave every warp 4 rows and the whole K range: each of the 8 warps then streams the whole 8-column slice
// of W through L1 (128 KB of L1 wavefronts per task next to 64 KB of activations) and the task is bound by the ~64 B/clk of
// the L1 data path (measured: ~3 us per 32 x 8 x 512 task).  Here the 32 x 8 output tile is shared by all warps and the
// REDUCTION dimension is split: warp w takes the 32-wide k-chunks w, w+8, ...; inside a chunk lane (rq, kq) loads one
// float4 of k for the rows rq, rq+4, ..., rq+28 (8 lanes x 16 B = one 128-byte line per row: coalesced) and the matching
// 4 x 8 block of W.  Every activation and every weight element enters the SM once (80 KB per task).  The 64 partial sums of
// a lane are transposed-reduced over the 8 k-lanes (56 shuffles), the 8 warps meet in 8 KB of shared memory, and thread t
// finishes output (t / 8, t % 8).  LAYOUT 0: C = A B^T (B = W[n, k]); LAYOUT 1: C = A B (B = W[k, n]).
template <int LAYOUT>
__device__ __forceinline__ void mega_gemm_ksplit(const MGemm& g, int tile, float* smem, long long* tr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rq = lane >> 3, kq = lane & 7;
  const int ct = (g.N + MEGA_CT - 1) / MEGA_CT, rgs = (g.M + 31) >> 5;
  const int c_idx = tile % ct; tile /= ct;
  const int rg = tile % rgs; const int kz = tile / rgs;
  const int n0 = c_idx * MEGA_CT, m0 = rg * 32;
  int k0 = 0, k1 = g.K;
  if (LAYOUT == 0 && g.splits > 1) { const int per = (((g.K + g.splits - 1) / g.splits) + 127) & ~127; k0 = kz * per; k1 = min(g.K, k0 + per); }
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
  const float* arow[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int r = m0 + rq + 4 * j; arow[j] = g.A.p + (int64_t)(r < g.M ? r : m0) * g.A.ld; }
  const int nchunk = (k1 - k0 + 31) >> 5;
  for (int ch = warp; ch < nchunk; ch += ROW_WARPS) {
    const int k = k0 + ch * 32 + 4 * kq;
    if (k < k1) {                                                 // K % 4 == 0: a float4 of k is in or out as a whole
      float4 a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __ldcg((const float4*)(arow[j] + k));
      float b[4][8];                                              // b[i][c] = W element for reduction index k + i and output column n0 + c
      if (LAYOUT == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 w = __ldg((const float4*)(g.B.p + (int64_t)(n0 + c < g.N ? n0 + c : n0) * g.B.ld + k));
          b[0][c] = w.x; b[1][c] = w.y; b[2][c] = w.z; b[3][c] = w.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* wr = g.B.p + (int64_t)(k + i) * g.B.ld + n0;
          const float4 w0 = __ldg((const float4*)wr), w1 = __ldg((const float4*)(wr + 4));
          b[i][0] = w0.x; b[i][1] = w0.y; b[i][2] = w0.z; b[i][3] = w0.w; b[i][4] = w1.x; b[i][5] = w1.y; b[i][6] = w1.z; b[i][7] = w1.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          acc[j][c] = fmaf(a[j].x, b[0][c], fmaf(a[j].y, b[1][c], fmaf(a[j].z, b[2][c], fmaf(a[j].w, b[3][c], acc[j][c]))));
    }
  }
  if (tr && threadIdx.x == 0) tr[1] = clock64();
  // transpose-reduce the 64 sums over the 8 k-lanes: lane-bit 1 keeps rows j in {0..3} / {4..7}, bit 2 halves again, bit 4 again:
  // each lane ends with ONE row j = 4 (kq & 1) + (kq & 2) + (kq >> 2) and all 8 columns
  float v32[4][8], v16[2][8], v8[8];
  {
    const bool up = (kq & 1) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float send = up ? acc[j][c] : acc[j + 4][c], keep = up ? acc[j + 4][c] : acc[j][c];
        v32[j][c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
  }
  {
    const bool up = (kq & 2) != 0;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float send = up ? v32[j][c] : v32[j + 2][c], keep = up ? v32[j + 2][c] : v32[j][c];
        v16[j][c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
  }
  {
    const bool up = (kq & 4) != 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float send = up ? v16[0][c] : v16[1][c], keep = up ? v16[1][c] : v16[0][c];
      v8[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const int jrow = 4 * (kq & 1) + (kq & 2) + (kq >> 2);
  const int row_local = rq + 4 * jrow;                            // 0 .. 31
  __syncthreads();                                                // the previous task is done with the exchange buffer
  float* mine = smem + warp * 256 + row_local * 8;
  *(float4*)mine = make_float4(v8[0], v8[1], v8[2], v8[3]);
  *(float4*)(mine + 4) = make_float4(v8[4], v8[5], v8[6], v8[7]);
  __syncthreads();
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < ROW_WARPS; ++w) v += smem[w * 256 + threadIdx.x];      // warp order: bit-reproducible
  if (tr && threadIdx.x == 0) tr[2] = clock64();
  const int m = m0 + (threadIdx.x >> 3), n = n0 + (threadIdx.x & 7);
  if (m < g.M && n < g.N) {
    if (LAYOUT == 0) {
      if (g.splits > 1) { g.C.p[((int64_t)kz * g.M + m) * g.C.ld + n] = v; return; }    // partial tile: bias / ReLU belong to the REDUCE task
      if (g.bias) v += __ldg(g.bias + n);
      if (g.relu) v = fmaxf(v, 0.f);
      g.C.p[(int64_t)m * g.C.ld + n] = v;
    } else {
      if (g.mask.p && !(__ldcg(g.mask.p + (int64_t)m * g.mask.ld + n) > 0.f)) v = 0.f;
      float* dst = g.C.p + (int64_t)m * g.C.ld + n;
      if (g.accumulate) v += __ldcg(dst);
      *dst = v;
    }
  }
}

__device__ __forceinline__ void mega_run_gemm(const MGemm& g, int tile, float* smem, long long* tr) {
  const bool a16 = ((((uintptr_t)g.A.p) | ((uintptr_t)g.B.p)) & 15) == 0 && g.A.ld % 4 == 0 && g.B.ld % 4 == 0;
  if (g.layout == MG_NT) {
    if (a16 && g.K % 4 == 0) mega_gemm_ksplit<0>(g, tile, smem, tr); else mega_gemm_nt<false>(g, tile, tr);
  } else if (g.layout == MG_NN) {
    if (a16 && g.K % 4 == 0 && g.N % MEGA_CT == 0) mega_gemm_ksplit<1>(g, tile, smem, tr); else mega_gemm_nn<false>(g, tile, smem);
  } else {
    const bool c16 = (((uintptr_t)g.C.p) & 15) == 0 && g.C.ld % 4 == 0;
    if (a16 && c16 && g.M % 4 == 0 && g.N % 4 == 0) mega_gemm_tn<true>(g, tile, smem); else mega_gemm_tn<false>(g, tile, smem);
  }
}

// ----------------------------------------------------------------------------- row tasks
// weighted cross entropy + dlogits (nn.CrossEntropyLoss(weight), train_pad_20.py:52,111), ONE ROW PER WARP with the row -> CTA
// mapping of the warp-per-row bodies (row = bid * 8 + warp), so that it chains with the classifier head in one task.
// The denominator only needs the labels: every warp sums w[y] over the whole batch itself (B <= 64), dlogits leave scaled
// in one pass, and the loss is accumulated with one atomic per CTA (loss_out is zero at launch).
__device__ __forceinline__ void ce_rows_body(const CeArgs& a, int bid, int nblk, float* sm /* >= ROW_WARPS floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float den_local = 0.f;
  if (a.den_local) den_local = __ldcg(a.den_local);
  else {
    for (int r = lane; r < a.B; r += 32) {
      const int64_t yl = a.labels[r];
      den_local += (yl >= 0 && yl < a.C) ? (a.class_w ? __ldg(a.class_w + (int)yl) : 1.f) : 0.f;    // ignore_index and anything out of range weigh nothing
    }
    den_local = warp_sum(den_local);
  }
  const float den = a.denom ? __ldcg(a.denom) : den_local;
  const float inv_den = 1.f / den;
  float num = 0.f;
  for (int row = bid * ROW_WARPS + warp; row < a.B; row += nblk * ROW_WARPS) {
    const float* z = a.logits + (int64_t)row * a.C;
    float mx = -INFINITY;
    for (int c = lane; c < a.C; c += 32) mx = fmaxf(mx, __ldcg(z + c));
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < a.C; c += 32) se += expf(__ldcg(z + c) - mx);
    se = warp_sum(se);
    const int64_t yl = a.labels[row];
    const bool valid = yl >= 0 && yl < a.C;
    const int y = valid ? (int)yl : 0;
    const float w = valid ? (a.class_w ? __ldg(a.class_w + y) : 1.f) : 0.f;
    const float inv = 1.f / se;
    if (a.dlogits) for (int c = lane; c < a.C; c += 32) a.dlogits[(int64_t)row * a.C + c] = w * inv_den * (expf(__ldcg(z + c) - mx) * inv - (c == y ? 1.f : 0.f));
    if (lane == 0) num += w * (logf(se) + mx - __ldcg(z + y));
  }
  __syncthreads();
  if (lane == 0) sm[warp] = num;
  __syncthreads();
  if (threadIdx.x == 0) {
    float n = 0.f;
    for (int w = 0; w < ROW_WARPS; ++w) n += sm[w];
    red_add(a.loss_out + 1, n);
    red_add(a.loss_out + 0, n * inv_den);
    a.loss_out[2] = den_local;                                 // every CTA stores the same value
  }
}

// y = sum_s part[s] + bias (ReLU): the K-slices of a split forward Linear, summed in slice order
__device__ __forceinline__ void reduce_body(const ReduceArgs& a, int bid, int nblk) {
  const int n4 = a.N >> 2;
  const int64_t total = (int64_t)a.M * n4, plane = (int64_t)a.M * a.N;
  for (int64_t i = (int64_t)bid * MEGA_THREADS + threadIdx.x; i < total; i += (int64_t)nblk * MEGA_THREADS) {
    const int m = (int)(i / n4), n = (int)(i - (int64_t)m * n4) * 4;
    float4 s = a.bias ? __ldg((const float4*)(a.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < a.splits; ++z) {
      const float4 v = __ldcg((const float4*)(a.part + z * plane + (int64_t)m * a.N + n));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (a.relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
    *(float4*)(a.y + (int64_t)m * a.ldy + n) = s;
  }
}

// ----------------------------------------------------------------------------- the classifier tail as ONE kernel (large batches)
// dropout(relu(LN(z2))) -> classifier head -> weighted cross entropy -> head backward -> LayerNorm backward are all row-wise
// on rows up to 512 wide with the same row -> warp mapping, so one launch runs the five bodies back to back (a CTA only ever
// touches its own rows; __syncthreads makes its global writes visible to its next body): 6 launches -> 1 (+ the 1-CTA
// denominator kernel below, off the critical path), "classifier fused with the loss" of SURVEY.md 2.1 K5/K6.
struct TailArgs { LnrdArgs ln_fwd; SmallNArgs head_fwd; CeArgs ce; SmallNArgs head_bwd; LnrdArgs ln_bwd; };
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32) tail_chain_kernel(const __grid_constant__ TailArgs a) { pdl_sync();
  extern __shared__ __align__(16) float tail_smem[];
  float* red = tail_smem; float2* scratch = (float2*)(tail_smem + ROW_WARPS * 512); float* sm_dw = tail_smem + ROW_WARPS * 512 + 64;
  const int bid = blockIdx.x, nblk = gridDim.x;
  lnrd_fwd_body<NV, 32>(a.ln_fwd, bid, nblk, scratch, red); __syncthreads();
  smalln_fwd_body<8>(a.head_fwd, bid, nblk); __syncthreads();
  ce_rows_body(a.ce, bid, nblk, red); __syncthreads();
  if (a.head_bwd.K <= 128) smalln_bwd_body<8, 1>(a.head_bwd, bid, nblk, sm_dw);
  else if (a.head_bwd.K <= 256) smalln_bwd_body<8, 2>(a.head_bwd, bid, nblk, sm_dw);
  else smalln_bwd_body<8, 4>(a.head_bwd, bid, nblk, sm_dw);
  __syncthreads();
  lnrd_bwd_body<NV, 32>(a.ln_bwd, bid, nblk, scratch, red);
}
// this batch's sum of class weights over its labels (the denominator of nn.CrossEntropyLoss(weight, reduction='mean')): one CTA
__global__ void __launch_bounds__(256) ce_den_kernel(const int64_t* __restrict__ labels, const float* __restrict__ class_w, int B, int C, float* __restrict__ out) { pdl_sync();
  __shared__ float s[8];
  float d = 0.f;
  for (int r = threadIdx.x; r < B; r += 256) { const int64_t y = labels[r]; d += (y >= 0 && y < C) ? (class_w ? __ldg(class_w + (int)y) : 1.f) : 0.f; }
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += s[w]; *out = t; }
}
inline cudaError_t launch_tail_chain(const TailArgs& a, int num_sms, cudaStream_t st) {
  int grid = (a.ln_fwd.B + ROW_WARPS - 1) / ROW_WARPS; if (grid > num_sms * 4) grid = num_sms * 4; if (grid < 1) grid = 1;
  const size_t smem = (size_t)(ROW_WARPS * 512 + 64 + 8 * 512 + 8) * sizeof(float);
  const int N = a.ln_fwd.N;
  if (N <= 128) pdl_launch(tail_chain_kernel<1>, grid, ROW_WARPS * 32, smem, st, a);
  else if (N <= 256) pdl_launch(tail_chain_kernel<2>, grid, ROW_WARPS * 32, smem, st, a);
  else pdl_launch(tail_chain_kernel<4>, grid, ROW_WARPS * 32, smem, st, a);
  return cudaGetLastError();
}

#define MEGA_ROW_CASE(KIND, BODY, ARGS)                                                        \
  case KIND: {                                                                                 \
    const int N_ = (ARGS).N;                                                                   \
    if (N_ <= 128) BODY<1, 32, true>(ARGS, tile, op.tiles, scratch, red);                            \
    else if (N_ <= 256) BODY<2, 32, true>(ARGS, tile, op.tiles, scratch, red);                       \
    else if (N_ <= 512) BODY<4, 32, true>(ARGS, tile, op.tiles, scratch, red);                       \
    else if (N_ <= 1024) BODY<1, 256, true>(ARGS, tile, op.tiles, scratch, red);                     \
    else if (N_ <= 2048) BODY<2, 256, true>(ARGS, tile, op.tiles, scratch, red);                     \
    else BODY<4, 256, true>(ARGS, tile, op.tiles, scratch, red);                                     \
  } break;

__device__ __noinline__ void mega_run_row_fwd(const MRowOp& op, int tile, float2* scratch, float* red) {
  switch (op.kind) {
    MEGA_ROW_CASE(MR_LNRD_FWD, lnrd_fwd_body, op.u.lnrd)
    MEGA_ROW_CASE(MR_GATE_FWD, gate_fwd_body, op.u.gate)
    MEGA_ROW_CASE(MR_GRB_FWD, grb_fwd_body, op.u.grb)
    MEGA_ROW_CASE(MR_META_FWD, meta_fwd_body, op.u.meta)
    default: break;
  }
}
__device__ __noinline__ void mega_run_row_bwd(const MRowOp& op, int tile, float2* scratch, float* red) {
  switch (op.kind) {
    MEGA_ROW_CASE(MR_LNRD_BWD, lnrd_bwd_body, op.u.lnrd)
    MEGA_ROW_CASE(MR_GATE_BWD, gate_bwd_body, op.u.gate)
    MEGA_ROW_CASE(MR_GRB_BWD, grb_bwd_body, op.u.grb)
    MEGA_ROW_CASE(MR_META_BWD, meta_bwd_body, op.u.meta)
    default: break;
  }
}
__device__ __noinline__ void mega_run_row_misc(const MRowOp& op, int tile, float* red, float* sm_dw) {
  switch (op.kind) {
    case MR_SMALLN_FWD: smalln_fwd_body<8>(op.u.smalln, tile, op.tiles); break;
    case MR_SMALLN_BWD:
      if (op.u.smalln.K <= 128) smalln_bwd_body<8, 1>(op.u.smalln, tile, op.tiles, sm_dw);
      else if (op.u.smalln.K <= 256) smalln_bwd_body<8, 2>(op.u.smalln, tile, op.tiles, sm_dw);
      else smalln_bwd_body<8, 4>(op.u.smalln, tile, op.tiles, sm_dw);
      break;
    case MR_CE: ce_rows_body(op.u.ce, tile, op.tiles, red); break;
    case MR_REDUCE: reduce_body(op.u.red, tile, op.tiles); break;
    default: break;
  }
}

// shared memory: [ task scratch (row-body reductions / GEMM operand staging) | the program: stage table, row ops, GEMM ops ]
constexpr int MEGA_SCRATCH_BYTES = 44 * 1024;
constexpr int MEGA_PROG_BYTES = (int)(sizeof(MStage) * MEGA_MAX_STAGES + sizeof(MRowOp) * MEGA_MAX_ROW + sizeof(MGemm) * MEGA_MAX_GEMM);
constexpr int MEGA_DYN_SMEM = MEGA_SCRATCH_BYTES + MEGA_PROG_BYTES;          // ~69 KB per CTA
static_assert(MEGA_SCRATCH_BYTES >= (MEGA_MAX_B * (MEGA_TN_LDA + MEGA_TN_LDB)) * 4 && MEGA_SCRATCH_BYTES >= MEGA_NN_SLAB * MEGA_CT * 4 &&
              MEGA_SCRATCH_BYTES >= (ROW_WARPS * 512 + 64 + 8 * 512 + 8) * 4, "task scratch");

__global__ void __launch_bounds__(MEGA_THREADS, 1) mega_step_kernel(const __grid_constant__ MegaProg P) {
  extern __shared__ __align__(16) float mega_smem[];
  // one scratch region, two uses: row tasks (reduction scratch + classifier-head dW scratch) / GEMM operand staging
  float* red = mega_smem;                                         // [ROW_WARPS * 512]
  float2* scratch = (float2*)(mega_smem + ROW_WARPS * 512);       // [ROW_WARPS] (+ padding)
  float* sm_dw = mega_smem + ROW_WARPS * 512 + 64;                // classifier-head weight-gradient scratch [8 * 512 + 8]
  // The program lives in shared memory for the whole kernel: 24 KB of kernel parameters do not fit the constant cache, and
  // every stage would start with a chain of constant-cache misses (measured: ~1 us per stage before any task ran).
  const MStage* S_st = (const MStage*)((char*)mega_smem + MEGA_SCRATCH_BYTES);
  const MRowOp* S_r = (const MRowOp*)(S_st + MEGA_MAX_STAGES);
  const MGemm* S_g = (const MGemm*)(S_r + MEGA_MAX_ROW);
  {
    uint32_t* dst = (uint32_t*)S_st; const uint32_t* src = (const uint32_t*)&P.st[0];
    const int n_st = (int)(sizeof(MStage) * MEGA_MAX_STAGES / 4);
    const int n_r = (int)(sizeof(MRowOp) / 4) * P.nrow, n_g = (int)(sizeof(MGemm) / 4) * P.ngemm;
    for (int w = threadIdx.x; w < (int)(sizeof(MStage) / 4) * P.nstages; w += MEGA_THREADS) dst[w] = src[w];
    dst += n_st; src = (const uint32_t*)&P.r[0];
    for (int w = threadIdx.x; w < n_r; w += MEGA_THREADS) dst[w] = src[w];
    dst += (int)(sizeof(MRowOp) / 4) * MEGA_MAX_ROW; src = (const uint32_t*)&P.g[0];
    for (int w = threadIdx.x; w < n_g; w += MEGA_THREADS) dst[w] = src[w];
  }
  __syncthreads();
  const int G = gridDim.x, cta = blockIdx.x, nstages = P.nstages;
  unsigned* const barrier = P.barrier; long long* const trace = P.trace;
  if (trace && cta == 0 && threadIdx.x == 0) trace[0] = clock64();
  for (int s = 0; s < nstages; ++s) {
    if (trace && cta == 0 && threadIdx.x == 0) trace[256 + 8 * s + 3] = clock64();
    const MStage st = S_st[s];
    int base = 0;
    for (int i = st.r0; i < st.r1;) {
      int j = i + 1;
      while (j < st.r1 && S_r[j].chain) ++j;                      // ops i .. j-1 run back to back on the same rows, in one task
      const int tiles = S_r[i].tiles;
      int first = (cta - base) % G; if (first < 0) first += G;
      for (int tile = first; tile < tiles; tile += G) {
        for (int q = i; q < j; ++q) {
          const MRowOp& op = S_r[q];
          if (op.kind <= MR_META_BWD) { if (op.kind & 1) mega_run_row_bwd(op, tile, scratch, red); else mega_run_row_fwd(op, tile, scratch, red); }
          else mega_run_row_misc(op, tile, red, sm_dw);
          __syncthreads();                                        // global writes of this op are visible to the CTA's next op; the scratch is free again
        }
      }
      base += tiles; i = j;
    }
    for (int i = st.g0; i < st.g1; ++i) {
      const MGemm& g = S_g[i];
      const int tiles = g.tiles;
      int first = (cta - base) % G; if (first < 0) first += G;
      if (first < tiles) {
        long long* tr = (trace && cta == 0) ? trace + 256 + 8 * s : nullptr;      // debug: stamps inside the first GEMM task of CTA 0
        if (tr && threadIdx.x == 0) tr[4] = clock64();
        for (int tile = first; tile < tiles; tile += G) mega_run_gemm(g, tile, mega_smem, tr);
        if (tr && threadIdx.x == 0) tr[5] = clock64();
      }
      base += tiles;
    }
    if (trace && cta == 0) { __syncthreads(); if (threadIdx.x == 0) trace[1 + 2 * s] = clock64(); }     // own tasks done
    if (s + 1 < nstages) mega_grid_barrier(barrier, (unsigned)(s + 1), (unsigned)G);
    if (trace && cta == 0 && threadIdx.x == 0) trace[2 + 2 * s] = clock64();                             // barrier passed
  }
}

inline long long*& mega_trace_buffer() { static long long* p = nullptr; return p; }     // debug only (fb200_debug_mega_trace)

// ----------------------------------------------------------------------------- host: program builder
// The executors of exec.cu emit ops into a MegaBuilder instead of launching kernels; each op carries the stage in which
// its inputs are complete (tracked per activation / gradient buffer and per parameter slot by the caller).
struct MegaBuilder {
  std::vector<MGemm> g;
  std::vector<MRowOp> r;
  char* scratch = nullptr; size_t scratch_bytes = 0, scratch_used = 0;     // split-K partial tiles
  bool overflow = false;

  static MRef ref(const TRef& t) { MRef m; m.p = (float*)t.p; m.ld = t.ld; m.pad_ = 0; return m; }

  // returns the stage after which C is complete
  int add_gemm(const GemmArgs& a, int stage) {
    MGemm m{};
    m.A = ref(a.A); m.B = ref(a.B); m.C = ref(a.C); m.mask = ref(a.mask_src); m.mask.p = (float*)a.mask_src.p;
    m.bias = a.bias; m.colsum = a.colsum_a; m.M = a.M; m.N = a.N; m.K = a.K;
    m.relu = (short)a.relu; m.accumulate = (short)a.accumulate; m.splits = 1; m.stage = stage;

    const int rgs = (a.M + 31) / 32;
    if (a.a_kc && a.b_kc) {
      m.layout = MG_NT;
      m.tiles = rgs * ((a.N + MEGA_CT - 1) / MEGA_CT);
      if (a.K > MEGA_SPLIT_K && a.N % 4 == 0 && a.C.ld % 4 == 0 && !a.mask_src.p && !a.accumulate) {
        const int splits = (a.K + 511) / 512;
        const size_t need = (size_t)splits * a.M * a.N * sizeof(float);
        const size_t off = (scratch_used + 255) & ~size_t(255);
        if (scratch && off + need <= scratch_bytes && splits <= 16) {
          float* part = (float*)(scratch + off); scratch_used = off + need;
          MGemm pm = m; pm.splits = (short)splits; pm.tiles = m.tiles * splits; pm.bias = nullptr; pm.relu = 0;
          pm.C.p = part; pm.C.ld = a.N;
          g.push_back(pm);
          MRowOp ro{}; ro.kind = MR_REDUCE; ro.stage = stage + 1; ro.chain = 0;
          ro.u.red = ReduceArgs{part, (float*)a.C.p, a.bias, splits, a.M, a.N, a.C.ld, a.relu};
          ro.tiles = std::max(1, std::min(32, (a.M * a.N / 4 + MEGA_THREADS - 1) / MEGA_THREADS));
          r.push_back(ro);
          return stage + 1;
        }
      }
    } else if (a.a_kc && !a.b_kc) {
      m.layout = MG_NN; m.tiles = rgs * ((a.N + MEGA_CT - 1) / MEGA_CT);
    } else if (!a.a_kc && !a.b_kc) {
      // vector path (same test as mega_run_gemm): several 128-column chunks per task, the next one prefetched under the current
      const bool vec = ((((uintptr_t)a.A.p) | ((uintptr_t)a.B.p) | ((uintptr_t)a.C.p)) & 15) == 0 && a.A.ld % 4 == 0 && a.B.ld % 4 == 0 &&
                       a.C.ld % 4 == 0 && a.M % 4 == 0 && a.N % 4 == 0;
      const int nt = (a.N + 127) / 128, cpt = vec ? std::min(MEGA_TN_CPT, nt) : 1;
      m.layout = MG_TN; m.splits = (short)cpt; m.tiles = ((a.M + 31) / 32) * ((nt + cpt - 1) / cpt);
    } else { overflow = true; return stage; }
    g.push_back(m);
    return stage;
  }
  MRowOp& add_row(int kind, int stage, int tiles, bool chain = false) {
    MRowOp ro{}; ro.kind = kind; ro.stage = stage; ro.tiles = tiles < 1 ? 1 : tiles; ro.chain = chain ? 1 : 0;
    r.push_back(ro);
    return r.back();
  }
  // one row per warp (rows up to 512 wide) / per CTA (wider rows)
  static int row_tiles(int B, int N) { return N <= 512 ? (B + ROW_WARPS - 1) / ROW_WARPS : B; }

  // sort by stage, fill the stage table
  int finalize(MegaProg& P, unsigned* barrier) {
    if (overflow || g.size() > (size_t)MEGA_MAX_GEMM || r.size() > (size_t)MEGA_MAX_ROW) return FB200_EUNSUPPORTED;
    std::stable_sort(g.begin(), g.end(), [](const MGemm& a, const MGemm& b) { return a.stage < b.stage; });
    std::stable_sort(r.begin(), r.end(), [](const MRowOp& a, const MRowOp& b) { return a.stage < b.stage; });
    int last = 0;
    for (auto& x : g) last = std::max(last, x.stage);
    for (auto& x : r) last = std::max(last, x.stage);
    // compact the stage numbers that are actually used
    std::vector<int> used(last + 1, 0);
    for (auto& x : g) used[x.stage] = 1;
    for (auto& x : r) used[x.stage] = 1;
    std::vector<int> remap(last + 1, -1);
    int ns = 0;
    for (int s = 0; s <= last; ++s) if (used[s]) remap[s] = ns++;
    if (ns > MEGA_MAX_STAGES) return FB200_EUNSUPPORTED;
    for (int s = 0; s < ns; ++s) P.st[s] = MStage{0, 0, 0, 0};
    size_t gi = 0, ri = 0;
    for (int s = 0; s < ns; ++s) {
      P.st[s].g0 = (unsigned short)gi;
      while (gi < g.size() && remap[g[gi].stage] == s) { P.g[gi] = g[gi]; ++gi; }
      P.st[s].g1 = (unsigned short)gi;
      P.st[s].r0 = (unsigned short)ri;
      while (ri < r.size() && remap[r[ri].stage] == s) { P.r[ri] = r[ri]; ++ri; }
      P.st[s].r1 = (unsigned short)ri;
    }
    P.nstages = ns; P.ngemm = (int)g.size(); P.nrow = (int)r.size(); P.barrier = barrier; P.pad_ = 0; P.trace = mega_trace_buffer();
    static const bool dump = [] { const char* e = getenv("FB200_MEGA_DUMP"); return e && e[0] == '1'; }();
    if (dump) {
      static const char* const rk[] = {"lnrd_fwd", "lnrd_bwd", "gate_fwd", "gate_bwd", "grb_fwd", "grb_bwd", "meta_fwd", "meta_bwd", "smalln_fwd", "smalln_bwd", "ce", "reduce"};
      static const char* const gl[] = {"NT", "NN", "TN"};
      for (int s = 0; s < ns; ++s) {
        fprintf(stderr, "[mega] stage %2d:", s);
        int tasks = 0;
        for (int i = P.st[s].r0; i < P.st[s].r1; ++i) { fprintf(stderr, " %s%s(%d)", P.r[i].chain ? "+" : "", rk[P.r[i].kind], P.r[i].tiles); tasks += P.r[i].chain ? 0 : P.r[i].tiles; }
        for (int i = P.st[s].g0; i < P.st[s].g1; ++i) { fprintf(stderr, " %s[%dx%dx%d%s](%d)", gl[P.g[i].layout], P.g[i].M, P.g[i].N, P.g[i].K, P.g[i].splits > 1 ? " split" : "", P.g[i].tiles); tasks += P.g[i].tiles; }
        fprintf(stderr, "  = %d tasks\n", tasks);
      }
    }
    return FB200_OK;
  }
};

// Cooperative launch: every CTA is resident before any of them runs, which the grid barrier relies on.
inline int mega_launch(const MegaProg& P, int num_sms, cudaStream_t st) {
  static std::once_flag once; static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] { attr_rc = cudaFuncSetAttribute(mega_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_DYN_SMEM); });
  if (attr_rc != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  if (cudaMemsetAsync(P.barrier, 0, 256, st) != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  static int per_sm = 0;
  if (per_sm == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mega_step_kernel, MEGA_THREADS, MEGA_DYN_SMEM) != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
    per_sm = 1;                                   // one CTA per SM: cheaper barrier, 255 registers for the GEMM tasks (two per SM measured slower)
    if (occ < 1) return FB200_ECUDA;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms * per_sm); cfg.blockDim = dim3(MEGA_THREADS); cfg.dynamicSmemBytes = MEGA_DYN_SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, mega_step_kernel, P) != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  return FB200_OK;
}

}  // namespace fb200
