// tc_gemm.cuh - tcgen05 / TMEM / TMA GEMM for sm_100a, hand-written PTX.
//
//   C[M,N] (+)= A x B      fp32 accumulation in tensor memory
//   kind bf16   : tcgen05.mma.kind::f16, bf16 operands
//   kind tf32x3 : fp32-strict.  TMA lands raw fp32 tiles in shared memory.  The tensor core truncates a tf32
//                 operand to its top 19 bits, so the raw tile IS hi = trunc(x); the worker warps compute
//                 lo = rn_tf32(x - hi) (A: into tensor memory next to the raw values, B: into a second smem
//                 buffer at the same swizzled offsets) and the MMA thread issues three
//                 tcgen05.mma.kind::tf32 per k-slice (hi*hi + lo*hi + hi*lo): fp32 products to ~2^-21
//                 while HBM / L2 only ever carry 4 bytes per element and no operand is pre-processed in memory.
//   layouts     : each operand is either K-major (reduction dim contiguous) or MN-major, so
//                 forward (X W^T), dX (dY W) and dW (dY^T X) all read the SAME row-major
//                 tensors straight from HBM through TMA - nothing is ever transposed in memory.
//
// CTA = 576 threads: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..17 = workers (operand split, chunk promotion, epilogue:
// TMEM -> registers -> smem -> global, fused bias / ReLU / ReLU-mask / accumulate / split-K reduction).  One 128 x BLOCK_N output tile per CTA,
// STAGES-deep smem ring of 128-byte-swizzled tiles guarded by full/empty mbarriers.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace fb200 {

constexpr int TC_BM = 128;
constexpr int TC_NH = 4;                          // worker warps per TMEM lane quarter: each owns 1/TC_NH of the columns
constexpr int TC_WORKERS = 4 * TC_NH;             // split + promote + epilogue warps
constexpr int TC_GROUPS = 2;                      // fp32-strict: the workers split k-blocks in turn (group g takes i = g mod TC_GROUPS)
constexpr int TC_GWARPS = TC_WORKERS / TC_GROUPS; // warps of one group: all four lane quarters x TC_NH / TC_GROUPS column parts
constexpr int TC_THREADS = 64 + 32 * TC_WORKERS;  // + TMA producer warp + MMA warp
// bf16 kind, two CTAs per SM (r02e): a bf16 tile is mostly prologue + epilogue (8 k-blocks of ~0.25 us inside a ~4.8 us CTA), so a
// second resident CTA - half the ring each, the same bytes in flight per SM - lets one tile's epilogue overlap the other's k-loop
#ifndef FB200_BF16_OCC2
#define FB200_BF16_OCC2 1
#endif
template <int KIND, bool CL> constexpr bool TC_EARLY_RELEASE = KIND == 0 || CL;     // see the MMA warp in tc_gemm_tile
template <int KIND> constexpr int tc_min_ctas() { return (KIND == 0 && FB200_BF16_OCC2) ? 2 : 1; }

struct TcEpilogue {
  TRef C;                 // output view (any Fmt)
  const float* bias;      // [N] or nullptr
  int relu;
  TRef mask_src;          // multiply by [mask_src > 0] when .p != nullptr
  int accumulate;         // C += result
  int atomic;             // split-K: fp32 atomic add into C (FMT_F32 only)
  float* colsum;          // optional: colsum[n] += sum_m result(m,n)  (unused by the head; reserved)
  int csplit;             // set by the launcher: CTAs of one thread-block cluster (1,1,csplit) that share an output tile, each
                          // reducing a slice of K; the partial tiles meet through distributed shared memory (0 / 1 = none)
  int dbg;                // debug (timing experiments only, results are wrong): 1 skip B split, 2 skip A split, 16 no tcgen05.st, 32 no A TMA, 64 no B TMA, 128 no global stores in the epilogue
  long long* trace;       // debug: per-k-block clock64 stamps of CTA (0,0,0): [i*8 + {issue, full, mma_issued, empty_seen, split_done}]
  long long* timeline;    // debug: every CTA of this launch writes [cta * 8 + {entry, dependency wait passed, tile stored, SM id, accumulator in registers, tensor memory released, accumulator read by all workers, MMA warp left the k-loop}] (globaltimer ns)
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
// SPIN = true only for the single MMA-issuing thread.  Everybody else backs off with nanosleep between polls: ten warps
// polling flat out (try_wait returns at once on this hardware) compete with the MMA warp for issue slots on their
// schedulers - measured, the MMA thread needed ~800 cycles to issue the four k-slices of a k-block.
template <bool SPIN = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (uint32_t n = 1;; ++n) {
    if (!SPIN) __nanosleep(32);
    if (mbar_try_wait(bar, parity)) return;
    if ((n & 1023u) == 0 && clock64() - t0 > 4000000000LL) { __trap(); }
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// exactly one lane of the (converged) warp gets true
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// thread-block cluster: rank of this CTA, barrier over every thread of the cluster, peer shared-memory addresses
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");   // (not .aligned: callable after divergent code)
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t addr) {
  float4 v; asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory"); return v;
}
// sum of the float4 at the same shared-memory offset in the S CTAs of the cluster, in rank order (bit-reproducible); the S
// loads are issued back to back so that one remote latency covers them all
template <int S>
__device__ __forceinline__ float4 dsmem_sum_v4(uint32_t addr) {
  float4 w[S];
#pragma unroll
  for (int p = 0; p < S; ++p) w[p] = ld_dsmem_v4(cluster_map_shared(addr, (uint32_t)p));
  float4 v = w[0];
#pragma unroll
  for (int p = 1; p < S; ++p) { v.x += w[p].x; v.y += w[p].y; v.z += w[p].z; v.w += w[p].w; }
  return v;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// A operand from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <int N> __device__ __forceinline__ void tmem_stN(uint32_t taddr, const uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_stN<16>(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }
template <> __device__ __forceinline__ void tmem_stN<8>(uint32_t taddr, const uint32_t (&r)[8]) { tmem_st8(taddr, r); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int N> __device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ldN<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
template <> __device__ __forceinline__ void tmem_ldN<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// Shared-memory matrix descriptor (tcgen05 "SmemDescriptor"): start address, leading / stride
// byte offsets (all >> 4), version 1 (bits 46-47), 128-byte swizzle (layout type 2 in bits 61-63).
//   K-major  tile: rows of 128 B (the K slice), 8-row swizzle atoms 1024 B apart      -> SBO = 1024, LBO unused (1)
//   MN-major tile: 128 B of MN per K row, 8 K rows per atom (1024 B), MN chunks LBO apart
//   MN-major fp32 (tf32) tiles exist only in the "128B swizzle, 32-byte atom" flavour (layout type 1):
//   32-byte chunks swizzled by (k row % 4), 4 K rows per 512-byte atom -> SBO = 512
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// Instruction descriptor: fp32 accumulate, operand format, majors, N >> 3, M >> 4.
__host__ __device__ constexpr uint32_t make_idesc(int kind, int a_mn, int b_mn, int m, int n) {
  return (1u << 4) | ((uint32_t)(kind == 0 ? 1 : 2) << 7) | ((uint32_t)(kind == 0 ? 1 : 2) << 10) |
         ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int KIND, int BN, int ATM = 0>
struct TcCfg {
  static constexpr int ESIZE = KIND == 0 ? 2 : 4;
  static constexpr int BK = 128 / ESIZE;                 // elements of K per stage (one 128-byte swizzle row)
  static constexpr int UMMA_K = 32 / ESIZE;              // 16 (bf16) / 8 (tf32)
  static constexpr int PLANES = KIND == 0 ? 1 : 2;       // smem planes per operand tile (tf32x3: raw/hi + lo)
  static constexpr int A_BYTES = TC_BM * 128;            // one plane of the A tile
  static constexpr int B_BYTES = BN * 128;
  // ATM (fp32-strict): the split A tile goes to TENSOR MEMORY (tcgen05.st) instead of back to shared memory, and
  // the MMAs read A from TMEM - shared memory then only carries the raw A tile once and the B planes.
  // The raw A tiles and the B planes live in TWO rings: an A slot is free again as soon as the workers have read it
  // into registers, a B slot only when the MMAs that read it have retired.  Measured on B200, the mainloop was bound
  // by bytes in flight (TMA issue -> landed ~3400 cycles under load, over 4 coupled stages = 1080 cycles per k-block
  // with NO split work and one MMA instead of three); 4 A slots + 5 B slots use the same 224 KB better.
  static constexpr int A_PLANES = ATM ? 1 : PLANES;      // smem planes of A
  static constexpr int STAGES = (KIND == 0) ? (FB200_BF16_OCC2 ? 3 : (BN <= 128 ? 6 : 4)) : (ATM ? 5 : (BN <= 128 ? 3 : 2));   // ring of B (ATM) / of A+B stages
  static constexpr int A_STAGES = ATM ? 4 : STAGES;      // ATM: ring of raw A tiles in smem = ring of split A tiles in TMEM
  static constexpr int A_RING_BYTES = ATM ? A_STAGES * A_BYTES : 0;                 // ATM: A ring first, then the B ring
  static constexpr int B_OFF = ATM ? 0 : A_PLANES * A_BYTES;                         // B tile offset inside a stage
  static constexpr int STAGE_BYTES = (ATM ? 0 : A_PLANES * A_BYTES) + PLANES * B_BYTES;
  static constexpr int TMA_BYTES = A_BYTES + B_BYTES;    // bytes TMA delivers per k-block (one plane of each)
  static constexpr int RING_BYTES = A_RING_BYTES + STAGES * STAGE_BYTES;
  static constexpr int A_TMEM_COLS = 2 * (128 / ESIZE);  // hi + lo columns of one A stage in TMEM (ATM)
  static constexpr int EPC = 128 / ESIZE;                // elements per 128-byte chunk along MN (MN-major operands)
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align*/ + 512 /*barriers*/;
  // The tensor-core accumulator truncates on every accumulate, a bias that grows linearly with the number
  // of MMAs chained into one TMEM accumulator (measured: 3e-5 relative at K = 4096 in 3xTF32).  The
  // fp32-strict kind therefore accumulates at most CHUNK_KB k-blocks (K = 128) in TMEM, and the epilogue
  // warps promote each chunk into fp32 registers (round-to-nearest FADD) while the MMA warp already
  // fills the other of two TMEM accumulators.
  static constexpr int CHUNK_KB = KIND == 1 ? 4 : (1 << 30);
  static constexpr int ACC_BUFS = KIND == 1 ? 2 : 1;
  static constexpr int ACC_COLS = ACC_BUFS * BN;
  static constexpr int TMEM_NEED = ACC_COLS + (ATM ? A_STAGES * A_TMEM_COLS : 0);
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "tensor memory has 512 columns");
};

// One 128 x BN output tile: the whole pipeline described at the top of this file.
// CL: the cluster split-K variant (r02d).  It is a separate instantiation because merely carrying the runtime branches cost the
// ordinary kernels 2.9 % at the headline batch (same-box A/B against the r02c build: 0.736 -> 0.758 ms per step).
template <int KIND, int A_MN, int B_MN, int BN, bool CL = false>
__device__ __forceinline__ void tc_gemm_tile(const CUtensorMap* __restrict__ pmap_a, const CUtensorMap* __restrict__ pmap_b, const TcEpilogue& ep,
                                             const int M, const int N, const int m0, const int n0, const int kb_begin, const int num_kb,
                                             long long* tr) {
  constexpr int ATM = (KIND == 1) ? 1 : 0;
  using Cfg = TcCfg<KIND, BN, ATM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // 128B swizzle needs 1024-byte alignment
  uint8_t* ring_b = smem + Cfg::A_RING_BYTES;               // stages of B planes (ATM) / of A+B tiles
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::RING_BYTES); // [STAGES] TMA landed (ATM: the B tile)
  uint64_t* empty_bar = full_bar + Cfg::STAGES;             // [STAGES] the MMAs that read the stage have retired
  uint64_t* split_bar = empty_bar + Cfg::STAGES;            // [STAGES] tf32x3: tile split into hi/lo, ready for the MMA thread
  uint64_t* tmem_full = split_bar + Cfg::STAGES;            // [2]
  uint64_t* tmem_empty = tmem_full + 2;                     // [2]
  uint64_t* a_full = tmem_empty + 2;                        // [A_STAGES] ATM: raw A tile landed
  uint64_t* a_empty = a_full + Cfg::A_STAGES;               // [A_STAGES] ATM: every worker warp has read the raw A tile
  uint32_t* tmem_slot = (uint32_t*)(a_empty + Cfg::A_STAGES);
  static_assert((3 * Cfg::STAGES + 4 + 2 * Cfg::A_STAGES) * 8 + 4 <= 512, "barrier block");
  const CUtensorMap& map_a = *pmap_a;
  const CUtensorMap& map_b = *pmap_b;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (tr && threadIdx.x == 0) tr[5] = clock64();                                     // trace: kernel entry
  long long* tl = nullptr;
  if (ep.timeline && threadIdx.x == 0) {
    tl = ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
    uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[0] = global_ns(); tl[3] = smid;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], KIND == 1 ? TC_GWARPS : TC_WORKERS); }
    for (int a = 0; a < Cfg::A_STAGES; ++a) { mbar_init(&a_full[a], 1); mbar_init(&a_empty[a], TC_GWARPS); }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], TC_WORKERS); mbar_init(&tmem_empty[1], TC_WORKERS);   // one arrive per worker warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel in the stream; wait until every kernel this one depends on has completed and flushed before the
  // first global-memory access, then let the next kernel start ITS prologue.  (Releasing later - once the tile's
  // last MMA is issued - measured 1 % slower, with one stream and with two.)
  pdl_sync();
  if (tr && threadIdx.x == 0) tr[13] = clock64();                                    // trace: setup done
  if (tl) tl[1] = global_ns();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      auto load_b = [&](int i) {
        const int s = i % Cfg::STAGES; const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (tr) tr[i * 8 + 3] = clock64();
        if (ATM && (ep.dbg & 64)) { mbar_expect_tx(&full_bar[s], 0); return; }
        mbar_expect_tx(&full_bar[s], ATM ? Cfg::B_BYTES : Cfg::TMA_BYTES);
        uint8_t* st = ring_b + s * Cfg::STAGE_BYTES;
        const int k0 = (kb_begin + i) * Cfg::BK;
        uint8_t* b_dst = st + Cfg::B_OFF;
        if (!ATM) {
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < TC_BM / Cfg::EPC; ++c) tma_load_3d(&map_a, &full_bar[s], st + c * (Cfg::BK * 128), m0 + c * Cfg::EPC, k0, 0);
          } else {
            tma_load_3d(&map_a, &full_bar[s], st, k0, m0, 0);
          }
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / Cfg::EPC; ++c) tma_load_3d(&map_b, &full_bar[s], b_dst + c * (Cfg::BK * 128), n0 + c * Cfg::EPC, k0, 0);
        } else {
          tma_load_3d(&map_b, &full_bar[s], b_dst, k0, n0, 0);
        }
        if (tr) tr[i * 8 + 0] = clock64();
      };
      auto load_a = [&](int i) {                       // ATM only: the raw A tile of k-block i into its own ring
        const int sa = i % Cfg::A_STAGES; const uint32_t ph = (i / Cfg::A_STAGES) & 1;
        mbar_wait(&a_empty[sa], ph ^ 1);
        if (ep.dbg & 32) { mbar_expect_tx(&a_full[sa], 0); return; }
        mbar_expect_tx(&a_full[sa], Cfg::A_BYTES);
        uint8_t* a_dst = smem + sa * Cfg::A_BYTES;
        const int k0 = (kb_begin + i) * Cfg::BK;
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < TC_BM / Cfg::EPC; ++c) tma_load_3d(&map_a, &a_full[sa], a_dst + c * (Cfg::BK * 128), m0 + c * Cfg::EPC, k0, 0);
        } else {
          tma_load_3d(&map_a, &a_full[sa], a_dst, k0, m0, 0);
        }
      };
      // A runs one k-block ahead of B in issue order: the B ring is the tight one (its slots wait for the MMAs), and an
      // A tile must never queue behind a B slot that is not free yet
      if (ATM) load_a(0);
      for (int i = 0; i < num_kb; ++i) {
        load_b(i);
        if (ATM && i + 1 < num_kb) load_a(i + 1);
      }
    }
    if constexpr (CL) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }    // the two barriers of the workers' cluster epilogue
  } else if (warp == 1) {
    // ===== MMA issuer: the WHOLE warp walks the k-blocks in lock step, one ELECTED lane issues.  Uniform control flow keeps
    //       the descriptor arithmetic on the uniform datapath; under a plain `if (lane == 0)` the compiler wraps every
    //       tcgen05.mma in an elect / branch "waterfall" loop over the active lanes (~200 cycles per k-slice whatever the
    //       number of MMAs).  Issued back to back, the twelve TS-mode tf32 MMAs of a k-block take ~1170 cycles of tensor
    //       pipe (~97 each): that is the mainloop's bound now - a second issuing warp taking every other k-block (token
    //       passed through named barriers) hid the ~400-cycle handshake but did not shorten the period. =====
    {
      constexpr uint32_t idesc = make_idesc(KIND, ATM ? 0 : A_MN, B_MN, TC_BM, BN);
      constexpr uint32_t a_lt = (A_MN && KIND == 1) ? 1 : 2, b_lt = (B_MN && KIND == 1) ? 1 : 2;   // smem layout types
      constexpr uint32_t a_lbo = A_MN ? Cfg::BK * 128 : 16, a_sbo = (A_MN && KIND == 1) ? 512 : 1024;
      constexpr uint32_t b_lbo = B_MN ? Cfg::BK * 128 : 16, b_sbo = (B_MN && KIND == 1) ? 512 : 1024;
      constexpr uint32_t a_kstep = A_MN ? Cfg::UMMA_K * 128 : 32;                  // bytes per UMMA_K advance
      constexpr uint32_t b_kstep = B_MN ? Cfg::UMMA_K * 128 : 32;
      const uint64_t a_desc_base = make_smem_desc(0, a_lbo, a_sbo, a_lt), b_desc_base = make_smem_desc(0, b_lbo, b_sbo, b_lt);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % Cfg::STAGES; const uint32_t ph = (i / Cfg::STAGES) & 1;
        const int chunk = i / Cfg::CHUNK_KB, ab = chunk & (Cfg::ACC_BUFS - 1);
        const bool chunk_first = (i % Cfg::CHUNK_KB) == 0;
        if (tr && lane == 0) tr[i * 8 + 7] = clock64();
        if (chunk_first) { mbar_wait<true>(&tmem_empty[ab], ((chunk / Cfg::ACC_BUFS) & 1) ^ 1); }
        mbar_wait<true>(KIND == 1 ? &split_bar[s] : &full_bar[s], ph);
        if (tr && lane == 0) tr[i * 8 + 6] = clock64();
        tc_fence_after();
        if (tr && lane == 0) tr[i * 8 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * BN);
        // & 0x3FFFF: in a cluster launch the shared-window address of a CTA of rank > 0 carries the rank above bit 24; added
        // unmasked to a descriptor it spills out of the 14-bit start-address field into the LBO field (the chunk stride of
        // MN-major operands: every 32-column chunk but the first read garbage - found with cluster split-K, r02d)
        const uint32_t a_hi = smem_u32(ring_b + s * Cfg::STAGE_BYTES) & 0x3FFFFu;       // (A in smem: the non-ATM kinds)
        const uint32_t b_hi = a_hi + Cfg::B_OFF;
        // descriptors differ only in the 14-bit start-address field (>> 4): add to the low word
        const uint64_t da0 = a_desc_base + (uint64_t)(a_hi >> 4), db0 = b_desc_base + (uint64_t)(b_hi >> 4);
        const uint32_t a_tm = tmem_base + (uint32_t)(Cfg::ACC_COLS + (i % Cfg::A_STAGES) * Cfg::A_TMEM_COLS);      // ATM: hi columns, lo = +BK
        const bool last_of_chunk = (i % Cfg::CHUNK_KB) == Cfg::CHUNK_KB - 1 || i == num_kb - 1;
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < Cfg::BK / Cfg::UMMA_K; ++j) {
            const uint64_t da = da0 + (uint64_t)((j * a_kstep) >> 4);
            const uint64_t db = db0 + (uint64_t)((j * b_kstep) >> 4);
            const uint32_t first = (chunk_first && j == 0) ? 0u : 1u;
            if (ATM) {
              umma_ts_tf32(d_tmem, a_tm + j * Cfg::UMMA_K, db, idesc, first);
              umma_ts_tf32(d_tmem, a_tm + Cfg::BK + j * Cfg::UMMA_K, db, idesc, 1u);
              umma_ts_tf32(d_tmem, a_tm + j * Cfg::UMMA_K, db + (uint64_t)(Cfg::B_BYTES >> 4), idesc, 1u);
            } else {
              umma<KIND>(d_tmem, da, db, idesc, first);
            }
          }
          umma_commit(&empty_bar[s]);                                               // frees the smem slot once these MMAs retire
          if (last_of_chunk) umma_commit(&tmem_full[ab]);                           // chunk accumulator complete
        }
        __syncwarp();
        if (tr && lane == 0) tr[i * 8 + 2] = clock64();
      }
    }
    if (ep.timeline && lane == 0) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[7] = global_ns();   // timeline: MMA warp left the k-loop
    // Release tensor memory as soon as the workers have read the last accumulator chunk(s) out of it (every worker warp arrives
    // on tmem_empty after its tcgen05.ld + wait::ld), while they transpose and store the tile: tcgen05.dealloc takes ~0.5 us here,
    // which otherwise sits between the tile's last store and the CTA's exit (r02e, tools/tc_handover.py).  Same-box A/B of the
    // whole step: bf16 cfg5 B = 4096 0.531 -> 0.517 ms, fp32 B = 256 (cluster split-K) 0.2192 -> 0.2176, but the fp32 one-wave
    // GEMMs of the headline batch 0.7292 -> 0.7325 - those keep the release at the end of the CTA.
    if constexpr (TC_EARLY_RELEASE<KIND, CL>) {
      const int num_chunks = (num_kb + Cfg::CHUNK_KB - 1) / Cfg::CHUNK_KB;
      for (int ch = num_chunks > Cfg::ACC_BUFS ? num_chunks - Cfg::ACC_BUFS : 0; ch < num_chunks; ++ch)
        mbar_wait(&tmem_empty[ch & (Cfg::ACC_BUFS - 1)], (ch / Cfg::ACC_BUFS) & 1);
      if (ep.timeline && lane == 0) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[6] = global_ns();   // timeline: accumulator read out by every worker
      tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
      if (ep.timeline && lane == 0) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[5] = global_ns();   // timeline: tensor memory released
    }
    if constexpr (CL) { cluster_sync_all(); cluster_sync_all(); }
  } else {
    // ===== workers: warps 2..17.  A warp may only touch TMEM lanes [32 (warp % 4), +32); the TC_NH warps that share a
    //       lane quarter split the columns between them (h = which part).  Four warps per scheduler: the operand split is
    //       issue / latency bound - its slowest warp, not the tensor pipe, set the pace with one (1275 cycles per k-block)
    //       and with two warps per scheduler (1316; the twelve MMAs of a k-block issue in ~700). =====
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    constexpr int HN = BN / TC_NH;                                                   // accumulator columns per thread
    const int num_chunks = (num_kb + Cfg::CHUNK_KB - 1) / Cfg::CHUNK_KB;
    float acc[HN];
#pragma unroll
    for (int j = 0; j < HN; ++j) acc[j] = 0.f;
    // chunk accumulator (TMEM) -> fp32 registers, then hand the TMEM buffer back to the MMA thread
    auto promote = [&](int ch) {
      const int ab = ch & (Cfg::ACC_BUFS - 1);
      mbar_wait(&tmem_full[ab], (ch / Cfg::ACC_BUFS) & 1);
      tc_fence_after();
      constexpr int LDW = HN < 32 ? HN : 32;                                           // columns per tcgen05.ld
#pragma unroll
      for (int c0 = 0; c0 < HN; c0 += LDW) {
        uint32_t r[LDW];
        tmem_ldN<LDW>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + h * HN + c0), r);
#pragma unroll
        for (int j = 0; j < LDW; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[ab])) : "memory");
    };
    int next_promote = 0;
    if constexpr (KIND == 1) {
      // Two groups of eight warps take the k-blocks in turn: the split of one k-block is a chain of dependent
      // latencies (LDS -> ALU -> tcgen05.st / STS -> wait::st -> proxy fence -> arrive, ~1300 cycles under load whether 8 or
      // 16 warps share it), so each group gets two k-block periods for it and the tensor pipe sees one every period.
      constexpr int NHG = TC_NH / TC_GROUPS;                                        // column parts per lane quarter inside a group
      const int grp = h / NHG, hh = h % NHG;
      const int tw = (threadIdx.x - 64) - grp * (32 * TC_GWARPS);                   // 0 .. 32 * TC_GWARPS - 1 inside the group
      constexpr int KH = Cfg::BK / NHG;                                             // K values of the A tile this thread splits
      static_assert(ATM, "fp32-strict always stages A through tensor memory");
      for (int i = grp; i < num_kb; i += TC_GROUPS) {
        const int s = i % Cfg::STAGES; const uint32_t ph = (i / Cfg::STAGES) & 1;
        const int sa = i % Cfg::A_STAGES; const uint32_t pha = (i / Cfg::A_STAGES) & 1;
        // The tensor core TRUNCATES the low 13 mantissa bits of a tf32 operand.  So the raw fp32 value IS the hi
        // operand (hi = trunc(x), no ALU work and no smem rewrite), lo = x - trunc(x) is exact in fp32, and only lo
        // is rounded to nearest (ties away) on the bit pattern: a truncated lo would be a one-sided error that does
        // not average out over K.  x = hi + lo to 2^-22 |x|, unbiased.
        auto rnd = [](float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); };
        auto low = [&rnd](float x) { return rnd(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)); };
        // A: thread = output row m (TMEM lane 32q + lane); gather its KH values of K from the raw tile, split, store
        // hi | lo to tensor memory.
        mbar_wait(&a_full[sa], pha);
        const uint32_t sta = smem_u32(smem + sa * Cfg::A_BYTES);
        uint32_t hi[KH], lo[KH];
        if (ep.dbg & 2) {
#pragma unroll
          for (int rr = 0; rr < KH; ++rr) { hi[rr] = 0x3f800000u; lo[rr] = 0u; }
        } else if (A_MN) {
          // MN-major tile (dW: A = dY^T): chunk q holds m in [32q, 32q+32); K row r is 128 B with its four
          // 32-byte groups XOR-swizzled by (r & 3).  One 4-byte load per K value, conflict-free across the warp.
          const uint32_t base = sta + q * (Cfg::BK * 128) + (lane & 7) * 4;
          const int gsel = lane >> 3;
#pragma unroll
          for (int rr = 0; rr < KH; ++rr) {
            const int r = hh * KH + rr;
            float x;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(base + r * 128 + ((gsel ^ (r & 3)) << 5)));
            hi[rr] = __float_as_uint(x); lo[rr] = __float_as_uint(low(x));
          }
        } else {
          // K-major tile: row m at m*128 B with its eight 16-byte chunks XOR-swizzled by (m & 7)
          const int r = q * 32 + lane;
#pragma unroll
          for (int jj = 0; jj < KH / 4; ++jj) {
            const int j = hh * (KH / 4) + jj;
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(sta + r * 128 + ((j ^ (r & 7)) << 4)));
            hi[4 * jj] = __float_as_uint(x.x); hi[4 * jj + 1] = __float_as_uint(x.y); hi[4 * jj + 2] = __float_as_uint(x.z); hi[4 * jj + 3] = __float_as_uint(x.w);
            lo[4 * jj] = __float_as_uint(low(x.x)); lo[4 * jj + 1] = __float_as_uint(low(x.y)); lo[4 * jj + 2] = __float_as_uint(low(x.z)); lo[4 * jj + 3] = __float_as_uint(low(x.w));
          }
        }
        // tensor-memory slot sa was last read by the MMAs of k-block i - A_STAGES: they must have retired
        if (i >= Cfg::A_STAGES) { const int j = i - Cfg::A_STAGES; mbar_wait(&empty_bar[j % Cfg::STAGES], (j / Cfg::STAGES) & 1); tc_fence_after(); }
        const uint32_t a_tm = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::ACC_COLS + sa * Cfg::A_TMEM_COLS + hh * KH);
        static_assert(KH == 16 || KH == 8, "tcgen05.st shapes x16 / x8");
        if (!(ep.dbg & 16)) {
          tmem_stN<KH>(a_tm, hi);
          tmem_stN<KH>(a_tm + Cfg::BK, lo);
        }
        // the raw A tile is in registers / on its way to tensor memory: hand its smem slot back to the producer
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_empty[sa])) : "memory");
        // B: the raw tile stays where TMA put it (= hi); the lo plane goes to the second buffer at the same swizzled
        // offsets (elementwise, swizzle-oblivious); explicit ld/st.shared.  Issued between the TMEM stores and
        // their wait so that the store latency is covered.
        mbar_wait(&full_bar[s], ph);
        const uint32_t st = smem_u32(ring_b + s * Cfg::STAGE_BYTES);
        if (!(ep.dbg & 1)) {
#pragma unroll 4
          for (int v = tw; v < Cfg::B_BYTES / 16; v += 32 * TC_GWARPS) {
            const uint32_t hi_a = st + Cfg::B_OFF + v * 16, lo_a = hi_a + Cfg::B_BYTES;
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(hi_a));
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo_a), "f"(low(x.x)), "f"(low(x.y)), "f"(low(x.z)), "f"(low(x.w)) : "memory");
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");               // generic-proxy writes -> visible to tcgen05 (async proxy)
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_bar[s])) : "memory");
        if (tr && tw == 0) tr[i * 8 + 4] = clock64();
        // promote one chunk behind the split front so the MMA thread never starves (every warp of both groups takes part:
        // group g reaches the end of a chunk at its last k-block of that chunk)
        if ((i % Cfg::CHUNK_KB) >= Cfg::CHUNK_KB - TC_GROUPS && i / Cfg::CHUNK_KB >= 1) { promote(next_promote); ++next_promote; }
      }
    }
    for (; next_promote < num_chunks; ++next_promote) promote(next_promote);
    if (ep.timeline && threadIdx.x == 64) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[4] = global_ns();   // timeline: accumulator in registers
    if (tr && threadIdx.x == 64) tr[21] = clock64();                                 // trace: accumulator in registers
    // Coalesced epilogue.  A thread holds HN columns of one output ROW (TMEM lane) in registers; written as is, a warp
    // store would touch 32 rows x 16 B (measured: 5.5 us per tile, a third of a K=512 GEMM).  The warps transpose
    // through shared memory (the operand ring is idle by now: every MMA has retired) so that each store
    // instruction covers 512 contiguous bytes of ONE row, with bias / ReLU / mask / accumulate applied on the way.
    {
      constexpr int LDS_ROW = BN + 4;                                                // floats; +4 keeps the 128-bit accesses conflict free
      static_assert(4 * 32 * LDS_ROW * 4 <= Cfg::RING_BYTES, "epilogue staging must fit in the operand rings");
      const uint32_t wbase = smem_u32(smem) + (uint32_t)(q * 32 * LDS_ROW * 4);
#pragma unroll
      for (int g = 0; g < HN / 4; ++g)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(wbase + (uint32_t)((lane * LDS_ROW + h * HN + g * 4) * 4)),
                     "f"(acc[g * 4]), "f"(acc[g * 4 + 1]), "f"(acc[g * 4 + 2]), "f"(acc[g * 4 + 3]) : "memory");
      asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * TC_NH) : "memory");     // the TC_NH warps of this lane quarter
      // Cluster split-K (ep.csplit = S > 1): the S CTAs of the cluster hold partial sums of the SAME tile, parked at the same
      // shared-memory offsets.  After a cluster-wide barrier CTA `crank` finishes the rows r with (r / RPI) % S == crank: it reads
      // the S partial rows through distributed shared memory, adds them in rank order and runs the ordinary epilogue (bias, ReLU,
      // mask, accumulate, any output format) on the sum - no atomics, no zeroed output, bit-reproducible; a second barrier keeps
      // every CTA's shared memory alive until its peers have read it.
      const int S = CL ? ep.csplit : 1;                                              // CL kernels are only launched with csplit >= 2
      uint32_t crank = 0;
      if (CL && tr && threadIdx.x == 64) tr[45] = clock64();                         // trace: own partial tile parked, before the cluster barrier
      if constexpr (CL) { cluster_sync_all(); crank = cluster_ctarank(); }
      if (tr && threadIdx.x == 64) tr[37] = clock64();                               // trace: tile parked in smem
      const int64_t row_base = (int64_t)m0 + q * 32;
      const int rows_valid = (int)((M - row_base) < 32 ? (M - row_base) : 32);
      constexpr int RPW = 32 / TC_NH;                                                // rows each warp of the quarter stores
      const int r_begin = h * RPW;
      const int r_end = rows_valid < r_begin + RPW ? rows_valid : r_begin + RPW;
      const bool plain = !CL && !ep.atomic && !ep.mask_src.p && !ep.accumulate && (ep.C.fmt == FMT_F32 || (KIND == 0 && ep.C.fmt == FMT_BF16));
      // BN / 4 lanes cover one row with four columns each: a 128-wide tile is one row per warp store (512 contiguous
      // bytes), a 64-wide tile two rows of 256 bytes
      constexpr int CPR = BN / 4, RPI = 32 / CPR, UNR = RPW / RPI;
      static_assert(BN == 128 || BN == 64, "the epilogue maps BN/4 lanes to one row");
      const int c4 = lane % CPR, rsub = lane / CPR;
      const int n = n0 + c4 * 4;
      auto ld_tile = [&](uint32_t a) -> float4 {
        if constexpr (CL) {
          if (S == 2) return dsmem_sum_v4<2>(a);
          if (S == 4) return dsmem_sum_v4<4>(a);
          return dsmem_sum_v4<8>(a);
        }
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        return v;
      };
      auto mine = [&](int r) -> bool { return !CL || (uint32_t)((r / RPI) % S) == crank; };
      if (n < N && r_end > r_begin) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias) bv = __ldg((const float4*)(ep.bias + n));
        const uint32_t rbase = wbase + (uint32_t)(c4 * 16);
        const float floor_v = ep.relu ? 0.f : -INFINITY;
        if constexpr (CL) {
          // cluster split-K: gather the sums of the rows this CTA owns (u = crank, crank + S, ...: at most UNR / 2 of the warp's
          // rows) with every remote load in flight at once, tell the cluster that this CTA is done reading (arrive - the wait
          // comes after the stores, so a slow peer costs nothing here), then the ordinary epilogue on registers
          constexpr int OWN = UNR / 2;
          float4 v[OWN];
#pragma unroll
          for (int k = 0; k < OWN; ++k) {
            const int u = (int)crank + k * S, r = r_begin + u * RPI + rsub;
            v[k] = (u < UNR && r < r_end) ? ld_tile(rbase + (uint32_t)(r * LDS_ROW * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          cluster_arrive();
          if (tr && threadIdx.x == 64) tr[53] = clock64();                             // trace: partial sums gathered
          const bool has_mask = ep.mask_src.p != nullptr, acc_c = ep.accumulate != 0;
#pragma unroll
          for (int k = 0; k < OWN; ++k) {
            const int u = (int)crank + k * S, r = r_begin + u * RPI + rsub;
            if (u >= UNR || r >= r_end) continue;
            const int64_t row = row_base + r;
            float4 w;
            w.x = fmaxf(v[k].x + bv.x, floor_v); w.y = fmaxf(v[k].y + bv.y, floor_v);
            w.z = fmaxf(v[k].z + bv.z, floor_v); w.w = fmaxf(v[k].w + bv.w, floor_v);
            if (has_mask) { const float4 mk = ld4(ep.mask_src, row, n); w.x = mk.x > 0.f ? w.x : 0.f; w.y = mk.y > 0.f ? w.y : 0.f; w.z = mk.z > 0.f ? w.z : 0.f; w.w = mk.w > 0.f ? w.w : 0.f; }
            if (acc_c) { const float4 o = ld4(ep.C, row, n); w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
            st4(ep.C, row, n, w);
          }
        } else if (plain) {
          // hot path: all loads of the warp's rows first, then the stores (every data-dependent branch costs its full latency:
          // measured 280 cycles per row in a generic loop, 9 k cycles per tile)
          float* cp = (float*)ep.C.p + row_base * ep.C.ld + n;
          __nv_bfloat16* cb = (__nv_bfloat16*)ep.C.p + row_base * ep.C.ld + n;      // (bf16 outputs: r02d, 8-byte stores, same shape)
          const bool f32_out = KIND == 1 || ep.C.fmt == FMT_F32;                    // fp32-strict never stores bf16: the branch below folds away
          float4 v[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int r = r_begin + u * RPI + rsub;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "r"(rbase + (uint32_t)(r * LDS_ROW * 4)));
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int r = r_begin + u * RPI + rsub;
            v[u].x = fmaxf(v[u].x + bv.x, floor_v); v[u].y = fmaxf(v[u].y + bv.y, floor_v);
            v[u].z = fmaxf(v[u].z + bv.z, floor_v); v[u].w = fmaxf(v[u].w + bv.w, floor_v);
            if (r < r_end && !(ep.dbg & 128)) {
              if (f32_out) *(float4*)(cp + (int64_t)r * ep.C.ld) = v[u];
              else {
                const __nv_bfloat162 lo2 = __floats2bfloat162_rn(v[u].x, v[u].y), hi2 = __floats2bfloat162_rn(v[u].z, v[u].w);
                uint2 raw; raw.x = *reinterpret_cast<const uint32_t*>(&lo2); raw.y = *reinterpret_cast<const uint32_t*>(&hi2);
                *(uint2*)(cb + (int64_t)r * ep.C.ld) = raw;
              }
            }
          }
        } else if (ep.atomic) {
          // split-K slice: fp32 atomic add into the (pre-zeroed or accumulated-into) output; the bias rides on slice 0, a
          // ReLU-backward mask is linear and applies per slice
          const bool has_mask = ep.mask_src.p != nullptr;
          for (int r = r_begin + rsub; r < r_end; r += RPI) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(rbase + (uint32_t)(r * LDS_ROW * 4)));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (has_mask) {
              const float4 mk = ld4(ep.mask_src, row_base + r, n);
              v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
            }
            float* c = (float*)ep.C.p + (row_base + r) * ep.C.ld + n;
            atomicAdd(c, v.x); atomicAdd(c + 1, v.y); atomicAdd(c + 2, v.z); atomicAdd(c + 3, v.w);
          }
        } else {
          // any output format, ReLU-backward mask, accumulate-into-gradient, cluster split-K: every load of the trip issued
          // before the first use (the flags are warp-uniform: predicated, no divergence)
          const bool has_mask = ep.mask_src.p != nullptr, acc_c = ep.accumulate != 0;
          constexpr int U2 = UNR < 4 ? UNR : 4;
          for (int r0 = r_begin; r0 < r_end; r0 += U2 * RPI) {
            float4 v[U2], mk[U2], o[U2];
#pragma unroll
            for (int u = 0; u < U2; ++u) {
              const int r = r0 + u * RPI + rsub;
              const bool ok = r < r_end && mine(r);
              if (CL && !ok) { v[u] = mk[u] = o[u] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
              const int rr = ok ? r : r_begin;
              const int64_t row = row_base + rr;
              v[u] = ld_tile(rbase + (uint32_t)(rr * LDS_ROW * 4));
              mk[u] = has_mask ? ld4(ep.mask_src, row, n) : make_float4(1.f, 1.f, 1.f, 1.f);
              o[u] = acc_c ? ld4(ep.C, row, n) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U2; ++u) {
              const int r = r0 + u * RPI + rsub;
              if (r >= r_end || !mine(r)) continue;
              float4 w;
              w.x = fmaxf(v[u].x + bv.x, floor_v); w.y = fmaxf(v[u].y + bv.y, floor_v);
              w.z = fmaxf(v[u].z + bv.z, floor_v); w.w = fmaxf(v[u].w + bv.w, floor_v);
              w.x = (mk[u].x > 0.f ? w.x : 0.f) + o[u].x; w.y = (mk[u].y > 0.f ? w.y : 0.f) + o[u].y;
              w.z = (mk[u].z > 0.f ? w.z : 0.f) + o[u].z; w.w = (mk[u].w > 0.f ? w.w : 0.f) + o[u].w;
              if (!(ep.dbg & 128)) st4(ep.C, row_base + r, n, w);
            }
          }
        }
      }
      if constexpr (CL) {                                                             // peers may still be reading this CTA's tile
        if (!(n < N && r_end > r_begin)) cluster_arrive();                              // (threads with nothing to gather arrive here)
        cluster_wait();
      }
    }
    tc_fence_before();
    if (tr && threadIdx.x == 64) tr[29] = clock64();                                 // trace: tile stored
  }
  __syncthreads();
  // timeline: every warp has stored its rows.  (Stamped by a worker: thread 0 - the TMA producer, idle since its last load - was
  // observed to read the timer right after ARRIVING at this barrier, ~5 us before the workers got there: the SASS is
  // BAR.SYNC.DEFER_BLOCKING followed by CS2R SR_GLOBALTIMER, and the timer read evidently issues in the barrier's shadow - only
  // the store of the value waited.  A worker arrives last, so its stamp is right.)
  if (ep.timeline && threadIdx.x == 64) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[2] = global_ns();
  if constexpr (!TC_EARLY_RELEASE<KIND, CL>) {
    if (warp == 1) {
      if (ep.timeline && lane == 0) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[6] = global_ns();
      tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
      if (ep.timeline && lane == 0) (ep.timeline + 8 * (int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)))[5] = global_ns();
    }
  }
}

template <int KIND, int A_MN, int B_MN, int BN, bool CL = false>
__global__ void __launch_bounds__(TC_THREADS, tc_min_ctas<KIND>())
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const TcEpilogue ep, const int M, const int N, const int K, const int kb_per_split) {
  using Cfg = TcCfg<KIND, BN, KIND == 1 ? 1 : 0>;
  const int total_kb = (K + Cfg::BK - 1) / Cfg::BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int num_kb = min(total_kb, kb_begin + kb_per_split) - kb_begin;             // >= 1 by construction of the grid
  long long* tr = (ep.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? ep.trace : nullptr;
  tc_gemm_tile<KIND, A_MN, B_MN, BN, CL>(&map_a, &map_b, ep, M, N, blockIdx.y * TC_BM, blockIdx.x * BN, kb_begin, num_kb, tr);
}

// All weight gradients of one backward pass in ONE launch: problem p is dW_p[M_p, N_p] = dY_p^T X_p with the
// batch as the (shared) reduction length.  Tile -> problem by a prefix table; full-K tiles, plain stores:
// no split-K, no atomics, no memset, and ~2 waves of 128-k-block tiles instead of a dozen latency-bound launches.
constexpr int TC_MAX_GROUP = 24;
struct TcGroupArgs {
  CUtensorMap maps[2 * TC_MAX_GROUP];       // A (dY) and B (X) of every problem
  float* C[TC_MAX_GROUP]; int ldc[TC_MAX_GROUP]; int M[TC_MAX_GROUP]; int N[TC_MAX_GROUP];
  int tile_begin[TC_MAX_GROUP + 1]; int accumulate[TC_MAX_GROUP];
  int nprob; int K;
};
template <int KIND, int BN>
__global__ void __launch_bounds__(TC_THREADS, tc_min_ctas<KIND>()) tc_gemm_grouped_tn_kernel(const __grid_constant__ TcGroupArgs g) {
  using Cfg = TcCfg<KIND, BN, KIND == 1 ? 1 : 0>;
  int p = 0;
  while (p + 1 < g.nprob && (int)blockIdx.x >= g.tile_begin[p + 1]) ++p;
  const int t = blockIdx.x - g.tile_begin[p];
  const int tiles_n = (g.N[p] + BN - 1) / BN;
  TcEpilogue ep;
  ep.C = make_ref(g.C[p], g.ldc[p], FMT_F32); ep.bias = nullptr; ep.relu = 0; ep.mask_src.p = nullptr; ep.accumulate = g.accumulate[p];
  ep.atomic = 0; ep.colsum = nullptr; ep.trace = nullptr; ep.timeline = nullptr; ep.dbg = 0; ep.csplit = 0;
  tc_gemm_tile<KIND, 1, 1, BN>(&g.maps[2 * p], &g.maps[2 * p + 1], ep, g.M[p], g.N[p], (t / tiles_n) * TC_BM, (t % tiles_n) * BN, 0,
                               (g.K + Cfg::BK - 1) / Cfg::BK, nullptr);
}

// ----------------------------------------------------------------------------- host side
inline long long*& tc_trace_buffer() { static long long* p = nullptr; return p; }     // debug only (fb200_debug_tc_trace)
// debug only (fb200_debug_tc_timeline): launch l of the tcgen05 GEMM writes the stamps of its CTAs at buf + l * TC_TL_STRIDE
constexpr int TC_TL_STRIDE = 8 * 1024;
struct TcTimeline { long long* buf = nullptr; int max_launches = 0; int next = 0; };
inline TcTimeline& tc_timeline() { static TcTimeline t; return t; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// One GEMM operand as it lies in HBM: a row-major [outer, inner] matrix (inner contiguous),
// optionally a (hi, lo) pair of planes.  kmajor: inner is the reduction dimension.
struct TcOperand {
  const void* base; int64_t plane_elems; int ld; int inner, outer;
};
inline int encode_operand_map(CUtensorMap* map, int kind, const TcOperand& o, int box_inner, int box_outer, bool mn_major) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return FB200_ECUDA;
  const int es = kind == 0 ? 2 : 4;
  const int planes = 1;                        // tf32x3 reads raw fp32: the hi/lo split happens in shared memory
  if ((((uintptr_t)o.base) & 15) || ((int64_t)o.ld * es) % 16 || (planes == 2 && (o.plane_elems * es) % 16)) return FB200_EALIGN;
  cuuint64_t dims[3] = {(cuuint64_t)o.inner, (cuuint64_t)o.outer, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)o.ld * es, (cuuint64_t)(planes == 2 ? o.plane_elems : (int64_t)o.ld * o.outer) * es};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)o.base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, (kind == 1 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FB200_OK : FB200_ECUDA;
}

// cuTensorMapEncodeTiled costs ~1 us of host time; the eager drop-in route (model(x, meta) -> loss.backward(), one library
// call per pass, no CUDA graph) encodes two maps per GEMM per call.  Weights, workspace buffers and gradient slices recur
// with the same address and shape call after call, so encoded maps are kept per (pointer, geometry, box, kind) and per
// host thread (forward runs on the caller's thread, backward on autograd's worker: no lock, no sharing).  A tensor map
// holds no reference to the memory - it is an address plus strides - so a recycled address with the same geometry may
// reuse it safely.
struct TcMapKey {
  const void* base; int ld, inner, outer, box_inner, box_outer, kind_mn;
  bool operator==(const TcMapKey& o) const { return std::memcmp(this, &o, sizeof(TcMapKey)) == 0; }
};
struct TcMapKeyHash {
  size_t operator()(const TcMapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    auto mix = [&h](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix((uint64_t)k.ld << 32 | (uint32_t)k.inner); mix((uint64_t)k.outer << 32 | (uint32_t)k.box_inner); mix((uint64_t)k.box_outer << 8 | (uint32_t)k.kind_mn);
    return (size_t)h;
  }
};
inline int make_operand_map(CUtensorMap* map, int kind, const TcOperand& o, int box_inner, int box_outer, bool mn_major) {
  static thread_local std::unordered_map<TcMapKey, CUtensorMap, TcMapKeyHash> cache;
  TcMapKey key; std::memset(&key, 0, sizeof(key));
  key.base = o.base; key.ld = o.ld; key.inner = o.inner; key.outer = o.outer; key.box_inner = box_inner; key.box_outer = box_outer;
  key.kind_mn = kind * 2 + (mn_major ? 1 : 0);
  auto it = cache.find(key);
  if (it != cache.end()) { *map = it->second; return FB200_OK; }
  const int rc = encode_operand_map(map, kind, o, box_inner, box_outer, mn_major);
  if (rc != FB200_OK) return rc;
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *map);
  return FB200_OK;
}

struct TcGemmArgs {
  int kind;               // 0 bf16, 1 tf32x3
  int a_mn, b_mn;         // operand majors: 0 K-major, 1 MN-major
  TcOperand A, B;         // A covers (M, K), B covers (N, K) in the orientation given by the majors
  int M, N, K;
  TcEpilogue ep;
  int allow_split;        // weight-gradient GEMMs: split K across CTAs, atomically reduce into pre-zeroed fp32 C
  int alone;              // hint for cluster split-K: no GEMM of the other lane runs beside this one (it may fill the whole chip)
};

template <int KIND, int A_MN, int B_MN, int BN>
inline int tc_launch_one(const TcGemmArgs& g, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<KIND, BN, KIND == 1 ? 1 : 0>;
  CUtensorMap ma, mb;
  int rc = make_operand_map(&ma, KIND, g.A, A_MN ? Cfg::EPC : Cfg::BK, A_MN ? Cfg::BK : TC_BM, A_MN);
  if (rc != FB200_OK) return rc;
  rc = make_operand_map(&mb, KIND, g.B, B_MN ? Cfg::EPC : Cfg::BK, B_MN ? Cfg::BK : BN, B_MN);
  if (rc != FB200_OK) return rc;
  auto kern = tc_gemm_kernel<KIND, A_MN, B_MN, BN, false>;
  constexpr bool HAS_CL = KIND == 1 && !A_MN;      // cluster split-K variants exist for the fp32-strict forward / dX layouts (the only users)
  {                                                // forward runs on the caller's thread, backward on autograd's worker
    static std::once_flag once; static cudaError_t attr_rc = cudaSuccess;
    std::call_once(once, [&] {
      attr_rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
      if constexpr (HAS_CL) { if (attr_rc == cudaSuccess) attr_rc = cudaFuncSetAttribute(tc_gemm_kernel<KIND, A_MN, B_MN, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); }
    });
    if (attr_rc != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  }
  const int tiles_m = (g.M + TC_BM - 1) / TC_BM, tiles_n = (g.N + BN - 1) / BN;
  const int total_kb = (g.K + Cfg::BK - 1) / Cfg::BK;
  int split = 1;
  TcEpilogue ep = g.ep;
  if (g.allow_split && ep.C.fmt == FMT_F32 && !ep.bias && !ep.relu && !ep.mask_src.p) {
    const int tiles = tiles_m * tiles_n;
    int want = (num_sms + tiles - 1) / tiles;
    int maxs = total_kb / 8; if (maxs < 1) maxs = 1;            // keep >= 8 k-blocks per slice
    split = want < maxs ? want : maxs;
    if (split < 1) split = 1;
  }
  int kb_per = (total_kb + split - 1) / split;
  split = (total_kb + kb_per - 1) / kb_per;                       // no empty slices
  ep.atomic = split > 1 ? 1 : 0;
  // Cluster split-K: when the tiles of this GEMM leave most of the chip idle (batches up to ~512 rows against the 512-wide
  // layers) the k-loop - the only part of a latency-bound launch that parallelises - is cut across the CTAs of a cluster
  // (1, 1, S) and the partial tiles are summed through distributed shared memory in the epilogue (tc_gemm_tile).
  // S = the largest power of two that keeps the launch within HALF the chip (the other lane's GEMM runs beside it; measured:
  // 64 CTAs win 19-22 %, 128 CTAs are a wash in fp32 and lose 20 % in bf16), leaves every slice >= 2 k-blocks and divides the
  // epilogue's rows; and only when it removes enough of the k-loop to pay for two cluster barriers and the remote reads.
  // fp32-strict only: a bf16 k-block is ~0.2 us, splitting K = 512 measured no gain (cfg5, B = 32 .. 256: +-1 %).
  int csplit = 1;
  if (HAS_CL && split == 1) {
    static const int cap = [] { const char* e = getenv("FB200_TC_CSPLIT"); return e ? atoi(e) : 8; }();       // 0 / 1: off (A/B runs)
    const int tiles = tiles_m * tiles_n, max_s = BN == 128 ? 8 : 4, min_saved = 4;
    int s = 1;
    // chip share: a GEMM whose lane runs alone may fill the chip; beside the other lane's GEMM two slices may still fill it
    // (B = 1024: 64 tiles x 2 = 128 CTAs, 0.377 -> 0.358 ms per step), four or more stay within half (B = 512: 32 tiles x 4 =
    // 128 CTAs measured 0.287 against 0.274 ms with 64).  FB200_TC_CSPLIT_FILL = 1 / 2 forces whole / half for A/B runs.
    static const int fill_env = [] { const char* e = getenv("FB200_TC_CSPLIT_FILL"); return e ? atoi(e) : 0; }();
    auto room = [&](int s2) { const int div = fill_env > 0 ? fill_env : ((g.alone || s2 == 2) ? 1 : 2); return tiles * s2 <= num_sms / div; };
    while (s * 2 <= max_s && s * 2 <= cap && room(s * 2) && total_kb / (s * 2) >= 2) s *= 2;
    for (; s > 1; s >>= 1) { const int per = (total_kb + s - 1) / s; if ((total_kb + per - 1) / per == s) break; }   // no empty slices
    if (s > 1 && total_kb - (total_kb + s - 1) / s >= min_saved) { csplit = s; kb_per = (total_kb + s - 1) / s; }
  }
  ep.csplit = csplit;
  ep.trace = tc_trace_buffer();
  ep.timeline = nullptr;
  { TcTimeline& t = tc_timeline();
    if (t.buf && t.next < t.max_launches && tiles_m * tiles_n * (csplit > 1 ? csplit : split) <= TC_TL_STRIDE / 8) ep.timeline = t.buf + (int64_t)(t.next++) * TC_TL_STRIDE; }
  { static const int dbg = [] { const char* e = getenv("FB200_TC_DBG"); return e ? atoi(e) : 0; }(); ep.dbg = dbg; }
  dim3 grid(tiles_n, tiles_m, csplit > 1 ? csplit : split);
  cudaError_t lrc;
  if constexpr (HAS_CL) {
    lrc = csplit > 1 ? pdl_launch_cluster(tc_gemm_kernel<KIND, A_MN, B_MN, BN, true>, grid, dim3(TC_THREADS), Cfg::SMEM_BYTES, st, dim3(1, 1, csplit), ma, mb, ep, g.M, g.N, g.K, kb_per)
                     : pdl_launch(kern, grid, dim3(TC_THREADS), Cfg::SMEM_BYTES, st, ma, mb, ep, g.M, g.N, g.K, kb_per);
  } else lrc = pdl_launch(kern, grid, dim3(TC_THREADS), Cfg::SMEM_BYTES, st, ma, mb, ep, g.M, g.N, g.K, kb_per);
  if (lrc != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  return FB200_OK;
}

template <int KIND, int BN>
inline int tc_launch_major(const TcGemmArgs& g, int num_sms, cudaStream_t st) {
  if (!g.a_mn && !g.b_mn) return tc_launch_one<KIND, 0, 0, BN>(g, num_sms, st);
  if (!g.a_mn && g.b_mn) return tc_launch_one<KIND, 0, 1, BN>(g, num_sms, st);
  if (g.a_mn && g.b_mn) return tc_launch_one<KIND, 1, 1, BN>(g, num_sms, st);
  return FB200_EUNSUPPORTED;
}

// Shapes the tcgen05 path takes: 16-byte global strides (K, N multiples of 8) - ragged M/N/K
// tile edges are handled by TMA zero fill and epilogue predication.
inline bool tc_shape_ok(int layout, int M, int N, int K) {
  if (M < 1 || N < 8 || K < 1 || N % 8) return false;
  if (layout == 0) return K % 8 == 0;                 // A [M,K], B [N,K]
  if (layout == 1) return K % 8 == 0;                 // A [M,K], B [K,N]
  return M % 8 == 0;                                  // A [K,M], B [K,N]: the reduction length is free
}

// Tile width: 128 x 128 tiles are tensor-pipe-bound per k-block (twelve 128x128x8 TS-mode tf32 MMAs = ~1170 cycles against
// ~650 for the operand split), so when the 128-wide grid fills at most half the chip - batches up to ~2k rows against a
// 512-wide layer - 128 x 64 tiles double the CTAs and halve the per-tile MMA time: the k-loop of every CTA gets ~1.8x shorter.
inline int launch_tc_gemm(const TcGemmArgs& g, int num_sms, cudaStream_t st) {
  const int t128 = ((g.M + TC_BM - 1) / TC_BM) * ((g.N + 127) / 128);
  static const int force_bn = [] { const char* e = getenv("FB200_TC_BN"); return e ? atoi(e) : 0; }();     // A/B measurements
  const bool narrow = force_bn ? force_bn == 64 : (2 * t128 <= num_sms && g.N >= 64);
  if (g.kind == 0) return narrow ? tc_launch_major<0, 64>(g, num_sms, st) : tc_launch_major<0, 128>(g, num_sms, st);
  return narrow ? tc_launch_major<1, 64>(g, num_sms, st) : tc_launch_major<1, 128>(g, num_sms, st);
}


// ---- grouped weight-gradient launch ---------------------------------------------------------------
struct TcGroupProblem {
  TcOperand A, B;          // A = dY [K rows, M cols] (MN-major), B = X [K rows, N cols] (MN-major)
  int M, N;
  float* C; int ldc; int accumulate;
};
template <int KIND>
inline int tc_launch_grouped_tn(const TcGroupProblem* probs, int nprob, int K, cudaStream_t st) {
  constexpr int BN = 128;
  using Cfg = TcCfg<KIND, BN, KIND == 1 ? 1 : 0>;
  if (nprob < 1 || nprob > TC_MAX_GROUP) return FB200_EBADARG;
  TcGroupArgs g{};
  int tiles = 0;
  for (int p = 0; p < nprob; ++p) {
    int rc = make_operand_map(&g.maps[2 * p], KIND, probs[p].A, Cfg::EPC, Cfg::BK, true);
    if (rc != FB200_OK) return rc;
    rc = make_operand_map(&g.maps[2 * p + 1], KIND, probs[p].B, Cfg::EPC, Cfg::BK, true);
    if (rc != FB200_OK) return rc;
    g.C[p] = probs[p].C; g.ldc[p] = probs[p].ldc; g.M[p] = probs[p].M; g.N[p] = probs[p].N; g.accumulate[p] = probs[p].accumulate;
    g.tile_begin[p] = tiles;
    tiles += ((probs[p].M + TC_BM - 1) / TC_BM) * ((probs[p].N + BN - 1) / BN);
  }
  g.tile_begin[nprob] = tiles; g.nprob = nprob; g.K = K;
  auto kern = tc_gemm_grouped_tn_kernel<KIND, BN>;
  {
    static std::once_flag once; static cudaError_t attr_rc = cudaSuccess;
    std::call_once(once, [&] { attr_rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); });
    if (attr_rc != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  }
  if (pdl_launch(kern, dim3(tiles, 1, 1), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, g) != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; }
  return FB200_OK;
}
inline int launch_tc_grouped_tn(int kind, const TcGroupProblem* probs, int nprob, int K, cudaStream_t st) {
  return kind == 0 ? tc_launch_grouped_tn<0>(probs, nprob, K, st) : tc_launch_grouped_tn<1>(probs, nprob, K, st);
}

// ---- fp32-in / fp32-out wrapper used by the fb200_gemm primitive (unit tests, micro-benchmarks):
// converts A and B into the operand format inside `ws`, then runs the tcgen05 kernel.
__global__ void __launch_bounds__(256) tc_split_kernel(const float* __restrict__ in, int64_t rows, int cols, int ld_in, TRef out) { pdl_sync();
  const int64_t total4 = rows * (cols / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (cols / 4); const int c = (int)(i - r * (cols / 4)) * 4;
    st4(out, r, c, *(const float4*)(in + r * ld_in + c));
  }
}
inline size_t tc_operand_bytes(int64_t rows, int cols) { return (size_t)rows * cols * 2; }
inline int tc_gemm_workspace_bytes(int layout, int engine, int M, int N, int K, size_t* bytes) {
  if (engine != 1 && engine != 2) return FB200_EBADARG;
  if (!tc_shape_ok(layout, M, N, K)) return FB200_EUNSUPPORTED;
  *bytes = 256;
  if (engine == 2)      // bf16 copies of both operands
    *bytes = ((tc_operand_bytes(layout == 2 ? K : M, layout == 2 ? M : K) + 255) & ~size_t(255)) +
             ((tc_operand_bytes(layout == 0 ? N : K, layout == 0 ? K : N) + 255) & ~size_t(255)) + 256;
  return FB200_OK;
}
inline int tc_gemm_f32(int layout, int engine, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       const float* bias, int relu, int accumulate, void* ws, size_t ws_bytes, int num_sms, cudaStream_t st) {
  size_t need = 0;
  int rc = tc_gemm_workspace_bytes(layout, engine, M, N, K, &need);
  if (rc != FB200_OK) return rc;
  if ((lda % 4) || (ldb % 4) || (ldc % 4) || (((uintptr_t)A) & 15) || (((uintptr_t)B) & 15) || (((uintptr_t)C) & 15)) return FB200_EALIGN;
  const int kind = engine == 2 ? 0 : 1;
  const int64_t a_rows = layout == 2 ? K : M; const int a_cols = layout == 2 ? M : K;
  const int64_t b_rows = layout == 0 ? N : K; const int b_cols = layout == 0 ? K : N;
  TcGemmArgs g{};
  g.kind = kind; g.a_mn = layout == 2; g.b_mn = layout != 0;
  if (kind == 0) {
    if (!ws || ws_bytes < need || (((uintptr_t)ws) & 255)) return FB200_EBADARG;
    char* w = (char*)ws;
    TRef a_ref = make_ref(w, a_cols, FMT_BF16);
    TRef b_ref = make_ref(w + ((tc_operand_bytes(a_rows, a_cols) + 255) & ~size_t(255)), b_cols, FMT_BF16);
    pdl_launch(tc_split_kernel, 296, 256, 0, st, A, a_rows, a_cols, lda, a_ref);
    pdl_launch(tc_split_kernel, 296, 256, 0, st, B, b_rows, b_cols, ldb, b_ref);
    g.A = TcOperand{a_ref.p, 0, a_cols, a_cols, (int)a_rows};
    g.B = TcOperand{b_ref.p, 0, b_cols, b_cols, (int)b_rows};
  } else {
    g.A = TcOperand{A, 0, lda, a_cols, (int)a_rows};
    g.B = TcOperand{B, 0, ldb, b_cols, (int)b_rows};
  }
  g.M = M; g.N = N; g.K = K;
  g.ep.C = make_ref(C, ldc, FMT_F32); g.ep.bias = bias; g.ep.relu = relu; g.ep.mask_src.p = nullptr; g.ep.accumulate = accumulate;
  g.ep.atomic = 0; g.ep.colsum = nullptr; g.allow_split = 0;
  if (layout == 2 && !bias && !relu && ldc == N) {
    g.allow_split = 1;
    if (!accumulate && cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st) != cudaSuccess) return FB200_ECUDA;
  }
  return launch_tc_gemm(g, num_sms, st);
}

}  // namespace fb200
