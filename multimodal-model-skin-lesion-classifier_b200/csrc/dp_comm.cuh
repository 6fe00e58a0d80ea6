// dp_comm.cuh - gradient all-reduce of the data-parallel head over NVLink 5 / NVSwitch, hand-written.
//
// The head's gradients live in ONE flat fp32 buffer (fb200_grad_offset); under data parallelism that buffer is allocated
// in symmetric memory (every rank maps every peer's copy, plus the NVSwitch MULTICAST address that aliases all copies).
// One kernel per step does the whole SUM all-reduce, in place, and only over the LIVE ranges of the buffer (the W_q / W_k
// rows of every S = 1 attention are structural zeros - a third of the bytes for `crossattention` - and packing them away
// with copy kernels cost more than it saved with NCCL):
//
//   rank r owns the r-th 1/N slice of the live elements;
//     multimem path : v = multimem.ld_reduce.add [mc + off]   (the SWITCH sums the N copies and returns one vector)
//                     multimem.st [mc + off], v               (the switch writes it back into all N copies)
//     peer path     : v = sum over ranks p of ld [peer[p] + off]; st [peer[p] + off], v for every p   (no multicast address)
//   Slice r is read and then overwritten by rank r only, in every copy, so the kernel needs no barrier inside; the caller
//   brackets it with two cross-rank barriers (all gradients written / all results stored).
// Per GPU ~|live| bytes leave and ~|live| bytes arrive whatever N is: 18 MB -> ~25 us at the measured 770 GB/s per direction.
#pragma once
#include "common.cuh"

namespace fb200 {

constexpr int DP_MAX_RANGES = 48, DP_MAX_RANKS = 16;
struct DpRanges {
  long long begin4[DP_MAX_RANGES];     // first float4 of every live range (element offset / 4)
  long long prefix4[DP_MAX_RANGES + 1];// float4s before range i in the packed (virtual) order
  int n;
};
struct DpPeers { float* p[DP_MAX_RANKS]; };

__device__ __forceinline__ long long dp_locate(const DpRanges& r, long long v) {        // virtual float4 index -> float4 offset in the buffer
  int lo = 0, hi = r.n - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (r.prefix4[mid] <= v) lo = mid; else hi = mid - 1; }
  return r.begin4[lo] + (v - r.prefix4[lo]);
}

// ---- cross-rank barriers inside the kernel ---------------------------------------------------------------------------
// Every rank owns a SIGNAL PAD in symmetric memory (zero-initialised uint32 slots, mapped by every peer).  Barrier slot
// protocol (no epochs, reusable): rank r raises slot [base + r] in every peer's pad (CAS 0 -> 1, release.sys: spins while
// the previous use has not been consumed) and lowers slot [base + p] of its own pad for every peer p (CAS 1 -> 0,
// acquire.sys).  Bracketing the reduction with two such barriers INSIDE the kernel replaces two extra kernel launches per
// all-reduce: block 0 runs the opening barrier and releases the other blocks through a local flag; the last block to
// finish its stores (local counter) runs the closing one.  `state`: 4 uint32 of local device memory, zero at launch.
struct DpSync {
  uint32_t* pads[DP_MAX_RANKS];        // signal pad of every rank (peer-mapped), nullptr = barriers are the caller's business
  uint32_t* state;                     // [0] go flag, [1] blocks done
  int slot_open, slot_close;           // first slot of the two barriers in a pad (each uses `world` slots)
};
__device__ __forceinline__ void dp_spin_guard(long long t0, unsigned n) { if ((n & 1023u) == 0 && clock64() - t0 > 8000000000LL) __trap(); }
__device__ __forceinline__ void dp_rank_barrier(const DpSync& sy, int base, int rank, int world) {        // threads 0 .. world-1 of ONE block
  const int p = threadIdx.x;
  if (p < world) {
    uint32_t old; const long long t0 = clock64(); unsigned n = 0;
    uint32_t* theirs = sy.pads[p] + base + rank;
    do { asm volatile("atom.cas.release.sys.global.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(theirs) : "memory"); dp_spin_guard(t0, ++n); } while (old != 0u);
    uint32_t* mine = sy.pads[rank] + base + p;
    do { asm volatile("atom.cas.acquire.sys.global.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(mine) : "memory"); dp_spin_guard(t0, ++n); } while (old != 1u);
  }
}
__device__ __forceinline__ void dp_open(const DpSync& sy, int rank, int world) {
  if (!sy.pads[0]) return;
  if (blockIdx.x == 0) {
    dp_rank_barrier(sy, sy.slot_open, rank, world);
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], 1;" ::"l"(sy.state) : "memory");
  }
  if (threadIdx.x == 0) {
    uint32_t v; const long long t0 = clock64(); unsigned n = 0;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sy.state) : "memory"); dp_spin_guard(t0, ++n); } while (v == 0u);
  }
  __syncthreads();
}
__device__ __forceinline__ void dp_close(const DpSync& sy, int rank, int world) {
  if (!sy.pads[0]) return;
  __shared__ uint32_t s_last;
  asm volatile("fence.acq_rel.sys;" ::: "memory");                   // this thread's peer / multicast stores are ordered before the signal
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(sy.state + 1) : "memory");
    s_last = (prev == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) dp_rank_barrier(sy, sy.slot_close, rank, world);
}

constexpr int DP_THREADS = 256, DP_UNROLL = 4;       // 256-thread CTAs with ~40 registers co-reside with a tcgen05 GEMM CTA on its SM

__global__ void __launch_bounds__(DP_THREADS) dp_allreduce_multimem_kernel(float* __restrict__ mc, const __grid_constant__ DpRanges r, const __grid_constant__ DpSync sy, int rank, int world) {
  dp_open(sy, rank, world);
  const long long total = r.prefix4[r.n];
  const long long per = (total + world - 1) / world;
  const long long v0 = (long long)rank * per, v1 = v0 + per < total ? v0 + per : total;
  const long long stride = (long long)gridDim.x * DP_THREADS;
  // DP_UNROLL independent in-switch reductions in flight per thread before the first store: the NVLink round trip is ~2 us
  for (long long v = v0 + (long long)blockIdx.x * DP_THREADS + threadIdx.x; v < v1; v += stride * DP_UNROLL) {
    float* addr[DP_UNROLL]; float4 s[DP_UNROLL];
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const long long vu = v + u * stride;
      addr[u] = vu < v1 ? mc + 4 * dp_locate(r, vu) : nullptr;
      if (addr[u]) asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                                : "=f"(s[u].x), "=f"(s[u].y), "=f"(s[u].z), "=f"(s[u].w) : "l"(addr[u]) : "memory");
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u)
      if (addr[u]) asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr[u]), "f"(s[u].x), "f"(s[u].y), "f"(s[u].z), "f"(s[u].w) : "memory");
  }
  dp_close(sy, rank, world);
}

__global__ void __launch_bounds__(DP_THREADS) dp_allreduce_peer_kernel(const __grid_constant__ DpPeers peers, const __grid_constant__ DpRanges r, const __grid_constant__ DpSync sy, int rank, int world) {
  dp_open(sy, rank, world);
  const long long total = r.prefix4[r.n];
  const long long per = (total + world - 1) / world;
  const long long v0 = (long long)rank * per, v1 = v0 + per < total ? v0 + per : total;
  const long long stride = (long long)gridDim.x * DP_THREADS;
  for (long long v = v0 + (long long)blockIdx.x * DP_THREADS + threadIdx.x; v < v1; v += stride * 2) {
    long long off[2]; float4 s[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long vu = v + u * stride;
      off[u] = vu < v1 ? 4 * dp_locate(r, vu) : -1;
      s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int p = 0; p < world; ++p) {                                   // rank order: every rank computes bit-identical sums
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (off[u] < 0) continue;
        float4 x;                                                       // peer memory: system-scope load that bypasses L1
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(peers.p[p] + off[u]) : "memory");
        s[u].x += x.x; s[u].y += x.y; s[u].z += x.z; s[u].w += x.w;
      }
    }
    for (int p = 0; p < world; ++p) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (off[u] < 0) continue;
        float* dst = peers.p[(rank + p) % world] + off[u];
        asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(s[u].x), "f"(s[u].y), "f"(s[u].z), "f"(s[u].w) : "memory");
      }
    }
  }
  dp_close(sy, rank, world);
}

// ranges: [begin, end) element offsets (multiples of 4) of the live slices, ascending
// max_ctas: 0 = size the grid for the slice (up to 4 CTAs per SM); a small positive value keeps the kernel out of the way of a
// GEMM it overlaps with
// signal_pads / state / slot: in-kernel barriers (signal_pads == nullptr: the caller brackets the launch with its own)
inline int dp_allreduce_launch(float* mc, float* const* peer_ptrs, const int64_t* ranges, int nranges, int rank, int world, int num_sms, int max_ctas,
                               uint32_t* const* signal_pads, uint32_t* state, int slot, cudaStream_t st) {
  if (nranges < 1 || nranges > DP_MAX_RANGES || world < 1 || world > DP_MAX_RANKS || rank < 0 || rank >= world || !ranges) return FB200_EBADARG;
  if (!mc && !peer_ptrs) return FB200_EBADARG;
  DpRanges r{}; r.n = nranges; long long acc = 0;
  for (int i = 0; i < nranges; ++i) {
    const int64_t b = ranges[2 * i], e = ranges[2 * i + 1];
    if (b < 0 || e <= b || (b & 3) || (e & 3)) return FB200_EALIGN;
    r.begin4[i] = b / 4; r.prefix4[i] = acc; acc += (e - b) / 4;
  }
  r.prefix4[nranges] = acc;
  const long long per = (acc + world - 1) / world;
  int grid = (int)((per + DP_THREADS * 2 - 1) / (DP_THREADS * 2));
  const int cap = max_ctas > 0 ? max_ctas : num_sms * 4;          // every block must be resident (the opening barrier releases them through a flag)
  if (grid > cap) grid = cap; if (grid < 1) grid = 1;
  DpSync sy{};
  if (signal_pads) {
    if (!state || slot < 0) return FB200_EBADARG;
    for (int p = 0; p < world; ++p) { if (!signal_pads[p]) return FB200_EBADARG; sy.pads[p] = signal_pads[p]; }
    sy.state = state; sy.slot_open = slot; sy.slot_close = slot + world;
    if (cudaMemsetAsync(state, 0, 16, st) != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  }
  if (mc) {
    dp_allreduce_multimem_kernel<<<grid, DP_THREADS, 0, st>>>(mc, r, sy, rank, world);
  } else {
    DpPeers pp{};
    for (int p = 0; p < world; ++p) { if (!peer_ptrs[p]) return FB200_EBADARG; pp.p[p] = peer_ptrs[p]; }
    dp_allreduce_peer_kernel<<<grid, DP_THREADS, 0, st>>>(pp, r, sy, rank, world);
  }
  return cudaGetLastError() == cudaSuccess ? FB200_OK : FB200_ECUDA;
}

}  // namespace fb200
