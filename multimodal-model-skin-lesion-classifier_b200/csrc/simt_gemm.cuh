// simt_gemm.cuh - exact-fp32 FFMA GEMM for the shapes tcgen05 does not take
// (K = 85/13/11 metadata widths, N = num_classes, tiny batches) and for the
// FB200_FLAG_FORCE_SIMT exact-fp32 mode.  One kernel, three operand layouts.
//
//   C[M,N] (+)= sum_k A(m,k) * B(k,n)      A(m,k) = A[m*a_rs + k*a_cs],  B(k,n) = B[k*b_rs + n*b_cs]
//   NT (Linear fwd)  : A = X [M,K]  (a_cs = 1),  B = W [N,K]   (b_rs = 1)   -> both K-contiguous
//   NN (dX = dY W)   : A = dY[M,K'] (a_cs = 1),  B = W [K',N]  (b_cs = 1)
//   TN (dW = dY^T X) : A = dY[K',M] (a_rs = 1),  B = X [K',N]  (b_cs = 1)
// Epilogue: + bias[n], ReLU, multiply by [mask_src > 0] (ReLU backward of the producer),
// accumulate into C, split-K via fp32 atomics, column-sum of A (bias gradient) in TN mode.
#pragma once
#include "common.cuh"

namespace fb200 {

struct GemmArgs {
  TRef A, B, C;
  int M, N, K;
  int a_kc, b_kc;          // 1: operand is K-contiguous (ld = distance between m / n), 0: M/N-contiguous (ld = distance between k)
  const float* bias;       // [N] added in the epilogue (fp32 parameters), or nullptr
  int relu;                // apply max(.,0)
  TRef mask_src;           // multiply result by [mask_src(m,n) > 0] when mask_src.p != nullptr
  int accumulate;          // C += result (non-atomic read-modify-write)
  int split_k;             // >1: grid.z slices of K, results added with fp32 atomics (C must be FMT_F32, pre-zeroed or accumulate semantics)
  float* colsum_a;         // TN mode: colsum_a[m] += sum_k A(m,k)  (atomic), or nullptr
  // split-K with fix-up (small batches: the only parallelism left is along K): every slice parks its partial tile
  // in `partial`, the last slice to arrive (per-tile counter) sums them and runs the normal epilogue
  float* partial;          // [split][tiles][BM*BN] scratch, or nullptr -> atomic split-K
  unsigned* counters;      // [tiles], zero before the launch; the finishing CTA resets its counter
  size_t partial_bytes;    // capacity of `partial`
};

template <int BM, int BN, int BK, int TM, int TN>
struct SimtCfg {
  static constexpr int THREADS = (BM / TM) * (BN / TN);
  static constexpr int PAD = 4;
};

// element loaders that tolerate any alignment / format
__device__ __forceinline__ float gemm_ld(const TRef& t, int64_t idx) {
  if (t.fmt == FMT_F32) return __ldg((const float*)t.p + idx);
  if (t.fmt == FMT_PAIR) return __ldg((const float*)t.p + idx) + __ldg((const float*)t.p + t.plane + idx);
  return __bfloat162float(((const __nv_bfloat16*)t.p)[idx]);
}
__device__ __forceinline__ float4 gemm_ld4(const TRef& t, int64_t idx) {
  if (t.fmt == FMT_F32) return __ldg((const float4*)((const float*)t.p + idx));
  if (t.fmt == FMT_PAIR) {
    float4 h = __ldg((const float4*)((const float*)t.p + idx));
    float4 l = __ldg((const float4*)((const float*)t.p + t.plane + idx));
    return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
  }
  uint2 raw = __ldg((const uint2*)((const __nv_bfloat16*)t.p + idx));
  float2 fa = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
  float2 fb = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// Load a (ROWS x BK) operand tile into smem laid out [BK][ROWS + PAD].
//   kc = 1: element (r,k) at base + r*ld + k ; kc = 0: at base + k*ld + r
template <int ROWS, int BK, int THREADS, int PAD>
__device__ __forceinline__ void load_tile(const TRef& t, int kc, bool vec_ok, int64_t r0, int rmax, int k0, int kmax,
                                          float (*regs)[4], int tid) {
  constexpr int NV = ROWS * BK / 4;                 // float4 slots in the tile
  constexpr int PER = (NV + THREADS - 1) / THREADS;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int s = tid + i * THREADS;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    if (s < NV) {
      if (kc) {
        int r = s / (BK / 4), kq = (s % (BK / 4)) * 4;
        int64_t row = r0 + r; int k = k0 + kq;
        if (row < rmax) {
          int64_t idx = row * t.ld + k;
          if (vec_ok && k + 3 < kmax) { float4 f = gemm_ld4(t, idx); v0 = f.x; v1 = f.y; v2 = f.z; v3 = f.w; }
          else {
            if (k < kmax) v0 = gemm_ld(t, idx);
            if (k + 1 < kmax) v1 = gemm_ld(t, idx + 1);
            if (k + 2 < kmax) v2 = gemm_ld(t, idx + 2);
            if (k + 3 < kmax) v3 = gemm_ld(t, idx + 3);
          }
        }
      } else {
        int k = s / (ROWS / 4), rq = (s % (ROWS / 4)) * 4;
        int64_t row = r0 + rq; int kk = k0 + k;
        if (kk < kmax) {
          int64_t idx = (int64_t)kk * t.ld + row;
          if (vec_ok && row + 3 < rmax) { float4 f = gemm_ld4(t, idx); v0 = f.x; v1 = f.y; v2 = f.z; v3 = f.w; }
          else {
            if (row < rmax) v0 = gemm_ld(t, idx);
            if (row + 1 < rmax) v1 = gemm_ld(t, idx + 1);
            if (row + 2 < rmax) v2 = gemm_ld(t, idx + 2);
            if (row + 3 < rmax) v3 = gemm_ld(t, idx + 3);
          }
        }
      }
    }
    regs[i][0] = v0; regs[i][1] = v1; regs[i][2] = v2; regs[i][3] = v3;
  }
}

template <int ROWS, int BK, int THREADS, int PAD>
__device__ __forceinline__ void store_tile(float* sm, int kc, float (*regs)[4], int tid) {
  constexpr int NV = ROWS * BK / 4;
  constexpr int PER = (NV + THREADS - 1) / THREADS;
  constexpr int LD = ROWS + PAD;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int s = tid + i * THREADS;
    if (s < NV) {
      if (kc) {
        int r = s / (BK / 4), kq = (s % (BK / 4)) * 4;
        sm[(kq + 0) * LD + r] = regs[i][0];
        sm[(kq + 1) * LD + r] = regs[i][1];
        sm[(kq + 2) * LD + r] = regs[i][2];
        sm[(kq + 3) * LD + r] = regs[i][3];
      } else {
        int k = s / (ROWS / 4), rq = (s % (ROWS / 4)) * 4;
        *(float4*)(sm + k * LD + rq) = make_float4(regs[i][0], regs[i][1], regs[i][2], regs[i][3]);
      }
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
simt_gemm_kernel(const GemmArgs g) { pdl_sync();
  constexpr int THREADS = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  constexpr int LDA = BM + PAD, LDB = BN + PAD;
  constexpr int HM = TM / 4, HN = TN / 4;      // float4 groups per thread along m / n (split layout when > 1)
  static_assert(TM % 4 == 0 && TN % 4 == 0, "tile");
  __shared__ __align__(16) float As[2][BK * LDA];
  __shared__ __align__(16) float Bs[2][BK * LDB];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int64_t n0 = (int64_t)blockIdx.x * BN;

  // K range of this split
  int kb = 0, ke = g.K;
  if (g.split_k > 1) {
    int per = ((g.K + g.split_k - 1) / g.split_k + BK - 1) / BK * BK;
    kb = blockIdx.z * per; ke = min(g.K, kb + per);
  }
  const bool a_vec = (g.A.ld % 4 == 0) && ((((uintptr_t)g.A.p) & 15) == 0 || g.A.fmt == FMT_BF16 && (((uintptr_t)g.A.p) & 7) == 0);
  const bool b_vec = (g.B.ld % 4 == 0) && ((((uintptr_t)g.B.p) & 15) == 0 || g.B.fmt == FMT_BF16 && (((uintptr_t)g.B.p) & 7) == 0);

  constexpr int PA = (BM * BK / 4 + THREADS - 1) / THREADS;
  constexpr int PB = (BN * BK / 4 + THREADS - 1) / THREADS;
  float ra[PA][4], rb[PB][4];
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float asum[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) asum[i] = 0.f;
  const bool do_colsum = (g.colsum_a != nullptr) && (blockIdx.x == 0) && (tx == 0);

  int buf = 0;
  if (kb < ke) {
    load_tile<BM, BK, THREADS, PAD>(g.A, g.a_kc, a_vec, m0, g.M, kb, ke, ra, tid);
    load_tile<BN, BK, THREADS, PAD>(g.B, g.b_kc, b_vec, n0, g.N, kb, ke, rb, tid);
    store_tile<BM, BK, THREADS, PAD>(As[0], g.a_kc, ra, tid);
    store_tile<BN, BK, THREADS, PAD>(Bs[0], g.b_kc, rb, tid);
  }
  __syncthreads();
  for (int k0 = kb; k0 < ke; k0 += BK) {
    const bool more = (k0 + BK) < ke;
    if (more) {
      load_tile<BM, BK, THREADS, PAD>(g.A, g.a_kc, a_vec, m0, g.M, k0 + BK, ke, ra, tid);
      load_tile<BN, BK, THREADS, PAD>(g.B, g.b_kc, b_vec, n0, g.N, k0 + BK, ke, rb, tid);
    }
    const float* as = As[buf];
    const float* bs = Bs[buf];
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int h = 0; h < HM; ++h) {
        float4 v = *(const float4*)(as + k * LDA + h * (BM / HM) + ty * 4);
        a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int h = 0; h < HN; ++h) {
        float4 v = *(const float4*)(bs + k * LDB + h * (BN / HN) + tx * 4);
        b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_colsum) {
#pragma unroll
        for (int i = 0; i < TM; ++i) asum[i] += a[i];
      }
    }
    if (more) {
      store_tile<BM, BK, THREADS, PAD>(As[buf ^ 1], g.a_kc, ra, tid);
      store_tile<BN, BK, THREADS, PAD>(Bs[buf ^ 1], g.b_kc, rb, tid);
    }
    __syncthreads();
    buf ^= 1;
  }

  // ---- split-K fix-up: park the partial tile, the last arriving slice reduces and continues to the epilogue
  const bool fixup = g.split_k > 1 && g.partial != nullptr;
  if (fixup) {
    __shared__ int s_last;
    const int tiles = gridDim.x * gridDim.y, tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    float* mine = g.partial + ((size_t)blockIdx.z * tiles + tile_id) * (BM * BN);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int h = 0; h < HN; ++h) {
        const int lr = (i / 4) * (BM / HM) + ty * 4 + (i % 4), lc = h * (BN / HN) + tx * 4;
        *(float4*)(mine + lr * BN + lc) = make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(g.counters + tile_id, 1u);
      s_last = (prev == (unsigned)g.split_k - 1);
      if (s_last) g.counters[tile_id] = 0;                 // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed summation order (slice 0, 1, 2, ...) whoever arrives last: results stay bit-reproducible run to run
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int z = 0; z < g.split_k; ++z) {
      const float* other = g.partial + ((size_t)z * tiles + tile_id) * (BM * BN);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int h = 0; h < HN; ++h) {
          const int lr = (i / 4) * (BM / HM) + ty * 4 + (i % 4), lc = h * (BN / HN) + tx * 4;
          const float4 o = __ldcg((const float4*)(other + lr * BN + lc));
          acc[i][h * 4] += o.x; acc[i][h * 4 + 1] += o.y; acc[i][h * 4 + 2] += o.z; acc[i][h * 4 + 3] += o.w;
        }
    }
  }
  // ---- epilogue
  const bool c_vec = (g.C.ld % 4 == 0) && ((((uintptr_t)g.C.p) & 15) == 0 || g.C.fmt == FMT_BF16 && (((uintptr_t)g.C.p) & 7) == 0) &&
                     (g.mask_src.p == nullptr || g.mask_src.ld % 4 == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + (i / 4) * (BM / HM) + ty * 4 + (i % 4);
    if (m >= g.M) continue;
#pragma unroll
    for (int h = 0; h < HN; ++h) {
      int64_t n = n0 + h * (BN / HN) + tx * 4;
      if (n >= g.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      if (g.split_k > 1 && !fixup) {
        float* c = (float*)g.C.p + m * g.C.ld + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n + j < g.N) atomicAdd(c + j, v[j]);
        continue;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < g.N) {
          if (g.bias) v[j] += __ldg(g.bias + n + j);
          if (g.relu) v[j] = fmaxf(v[j], 0.f);
        }
      }
      if (c_vec && n + 3 < g.N) {
        if (g.mask_src.p) {
          float4 mk = ld4(g.mask_src, m, (int)n);
          v[0] = mk.x > 0.f ? v[0] : 0.f; v[1] = mk.y > 0.f ? v[1] : 0.f;
          v[2] = mk.z > 0.f ? v[2] : 0.f; v[3] = mk.w > 0.f ? v[3] : 0.f;
        }
        if (g.accumulate) { float4 o = ld4(g.C, m, (int)n); v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w; }
        st4(g.C, m, (int)n, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < g.N) {
            float o = v[j];
            if (g.mask_src.p) o = ld1(g.mask_src, m, (int)(n + j)) > 0.f ? o : 0.f;
            if (g.accumulate) o += ld1(g.C, m, (int)(n + j));
            st1(g.C, m, (int)(n + j), o);
          }
        }
      }
    }
  }
  if (do_colsum) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int64_t m = m0 + (i / 4) * (BM / HM) + ty * 4 + (i % 4);
      if (m < g.M) atomicAdd(g.colsum_a + m, asum[i]);
    }
  }
}

// Host-side launcher: picks the tile shape and split-K factor.
inline cudaError_t launch_simt_gemm(GemmArgs g, int num_sms, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  bool big = (g.M >= 256 && g.N >= 128) || (g.M >= 128 && g.N >= 256);
  if (big && ((g.M + 127) / 128) * ((g.N + 127) / 128) < num_sms) big = false;      // too few 128x128 tiles to fill the chip: smaller tiles
  int bm = big ? 128 : 64, bn = big ? 128 : 64;
  int tiles = ((g.M + bm - 1) / bm) * ((g.N + bn - 1) / bn);
  int split = 1;
  if (g.partial && g.counters && tiles <= 2048) {       // 2048 arrival counters per lane (exec.cu)
    // fix-up split-K (any epilogue): fill ~2 waves, keep >= 2 k-steps of 16 per slice
    int want = (2 * num_sms + tiles - 1) / tiles;
    int maxs = g.K / 32; if (maxs < 1) maxs = 1;
    split = want > maxs ? maxs : want;
    const int cap = (int)(g.partial_bytes / ((size_t)tiles * bm * bn * sizeof(float)));
    if (split > cap) split = cap;
    if (split < 2) { split = 1; }
  } else if (g.split_k != 1 && g.C.fmt == FMT_F32 && !g.bias && !g.relu && !g.mask_src.p) {
    // atomic split-K only where the epilogue is a plain sum into fp32 (weight gradients)
    int want = (2 * num_sms + tiles - 1) / tiles;
    // keep >= 32 reduction elements (two k-steps) per slice, at most 32 slices.  (Until r02e: >= 256 per slice - at B = 256 the
    // deferred text_fc.0 weight gradient ran as 8 CTAs x 16 k-steps = 26 us beside a 23 us grouped tcgen05 launch and set the
    // length of the step's tail.)
#ifdef FB200_OLD_SIMT_SPLIT
    int maxs = (g.K + 255) / 256;
#else
    int maxs = g.K / 32; if (maxs < 1) maxs = 1; if (maxs > 32) maxs = 32;
#endif
    split = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (split < 1) split = 1;
  }
  if (split == 1) { g.partial = nullptr; }
  g.split_k = split;
  dim3 grid((g.N + bn - 1) / bn, (g.M + bm - 1) / bm, split);
  if (big) pdl_launch(simt_gemm_kernel<128, 128, 16, 8, 8>, grid, 256, 0, st, g);
  else     pdl_launch(simt_gemm_kernel<64, 64, 16, 4, 4>, grid, 256, 0, st, g);
  return cudaGetLastError();
}

}  // namespace fb200
