// common.cuh - tensor references, format-generic 128-bit access, Philox, reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdlib>
#include "../../include/fb200.h"

namespace fb200 {

// Storage formats of activations inside the workspace.
//   FMT_F32  : plain fp32
//   FMT_PAIR : two fp32 planes (hi = tf32-rounded value, lo = exact remainder x - hi); hi + lo == x
//              bit-exactly, and the planes are directly the operands of the 3xTF32 tcgen05 GEMM
//   FMT_BF16 : bf16
enum Fmt : int { FMT_F32 = 0, FMT_PAIR = 1, FMT_BF16 = 2 };

struct TRef {           // a 2-D row-major view living in device memory
  void* p;              // first element (of the hi plane for FMT_PAIR)
  int64_t plane;        // FMT_PAIR: element distance hi -> lo plane
  int ld;               // row stride in elements
  int fmt;
};

__host__ __device__ inline TRef make_ref(void* p, int ld, int fmt = FMT_F32, int64_t plane = 0) {
  TRef r; r.p = p; r.plane = plane; r.ld = ld; r.fmt = fmt; return r;
}
__host__ __device__ inline size_t fmt_bytes(int fmt) { return fmt == FMT_BF16 ? 2 : 4; }

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- scalar access ----------------------------------------------------------------
// Activations and their gradients are read with ld.global.cg (L2 only): inside the persistent step kernel (mega.cuh) they
// were written by OTHER CTAs of the same launch a grid barrier earlier, which the non-coherent path (__ldg) does not
// allow; the stand-alone row kernels stream every element once, so bypassing L1 costs them nothing.
__device__ __forceinline__ float ld1(const TRef& t, int64_t r, int c) {
  int64_t i = r * t.ld + c;
  if (t.fmt == FMT_F32) return __ldcg((const float*)t.p + i);
  if (t.fmt == FMT_PAIR) return __ldcg((const float*)t.p + i) + __ldcg((const float*)t.p + t.plane + i);
  return __bfloat162float(__ushort_as_bfloat16(__ldcg((const unsigned short*)t.p + i)));
}
__device__ __forceinline__ void st1(const TRef& t, int64_t r, int c, float v) {
  int64_t i = r * t.ld + c;
  if (t.fmt == FMT_F32) { ((float*)t.p)[i] = v; }
  else if (t.fmt == FMT_PAIR) { float h = tf32_round(v); ((float*)t.p)[i] = h; ((float*)t.p)[t.plane + i] = v - h; }
  else { ((__nv_bfloat16*)t.p)[i] = __float2bfloat16_rn(v); }
}

// ---- 4-wide access (column multiple of 4, 16-byte aligned rows) ----------------------
__device__ __forceinline__ float4 ld4(const TRef& t, int64_t r, int c) {
  int64_t i = r * t.ld + c;
  if (t.fmt == FMT_F32) return __ldcg((const float4*)((const float*)t.p + i));
  if (t.fmt == FMT_PAIR) {
    float4 h = __ldcg((const float4*)((const float*)t.p + i));
    float4 l = __ldcg((const float4*)((const float*)t.p + t.plane + i));
    return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
  }
  uint2 raw = __ldcg((const uint2*)((const __nv_bfloat16*)t.p + i));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(const TRef& t, int64_t r, int c, float4 v) {
  int64_t i = r * t.ld + c;
  if (t.fmt == FMT_F32) { *(float4*)((float*)t.p + i) = v; }
  else if (t.fmt == FMT_PAIR) {
    float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
    *(float4*)((float*)t.p + i) = h;
    *(float4*)((float*)t.p + t.plane + i) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw; raw.x = *reinterpret_cast<uint32_t*>(&a); raw.y = *reinterpret_cast<uint32_t*>(&b);
    *(uint2*)((__nv_bfloat16*)t.p + i) = raw;
  }
}

// ---- warp / block reductions ---------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32-10 (counter-based; one call yields 4 uniforms for one float4 group) ---
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

struct DropSpec {            // one dropout site
  const uint8_t* mask;       // explicit keep-mask [rows, N] or nullptr -> Philox
  uint64_t seed, offset;
  const uint64_t* state;     // optional device pair {seed, offset} read at run time (CUDA-graph replays advance it)
  float p;                   // drop probability
  int site;
  int active;                // 0: identity
};

// keep-multipliers (0 or 1/(1-p)) for the 4 elements (row, c..c+3) of a width-N site
__device__ __forceinline__ float4 drop_mult4(const DropSpec& d, int64_t row, int c, int N) {
  if (!d.active) return make_float4(1.f, 1.f, 1.f, 1.f);
  float s = 1.0f / (1.0f - d.p);
  if (d.mask) {
    uchar4 m = *(const uchar4*)(d.mask + row * N + c);
    return make_float4(m.x ? s : 0.f, m.y ? s : 0.f, m.z ? s : 0.f, m.w ? s : 0.f);
  }
  uint64_t g = (uint64_t)(row * N + c) >> 2;
  const uint64_t seed = d.state ? __ldg(d.state) : d.seed;
  const uint64_t offset = d.state ? __ldg(d.state + 1) : d.offset;
  uint4 ctr = make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)offset, (uint32_t)(offset >> 32) ^ ((uint32_t)d.site << 24));
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t thr = (uint32_t)fminf(d.p * 4294967296.0f, 4294967040.0f);
  return make_float4(r.x >= thr ? s : 0.f, r.y >= thr ? s : 0.f, r.z >= thr ? s : 0.f, r.w >= thr ? s : 0.f);
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute: it may begin
// (block scheduling, prologue) while its predecessor in the stream drains.  pdl_sync() at kernel entry waits until
// every kernel this one depends on has completed and flushed its writes - no global-memory access may precede it -
// and only THEN lets the next kernel start its own prologue.  (Releasing the dependents before the wait lets a
// third kernel start while the first is still running; measured on B200: a reader two launches downstream then
// occasionally saw stale data.  With wait-then-release at most two kernels overlap.)
// FB200_PDL=0 launches without the attribute (A/B measurements).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_release(); }
// fp32 add into GLOBAL memory, result unused.  A plain atomicAdd through a pointer whose address space the compiler cannot
// prove (anything loaded from a program struct in memory, as in the persistent step kernel) compiles to a generic atomic:
// an isspacep test plus a shared-memory compare-and-swap loop beside the RED - 264 such loops in mega_step_kernel (r02e).
__device__ __forceinline__ void red_add(float* p, float v) {
#ifdef FB200_NO_RED_ASM
  atomicAdd(p, v);
#else
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
#endif
}
inline int& pdl_flag() { static int v = -1; return v; }
inline bool pdl_enabled() {
  int& v = pdl_flag();
  if (v < 0) { const char* e = getenv("FB200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same launch with a thread-block cluster shape (cluster split-K of the tcgen05 GEMM)
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster.x; at[0].val.clusterDim.y = cluster.y; at[0].val.clusterDim.z = cluster.z;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + __expf(-z)); }

}  // namespace fb200
