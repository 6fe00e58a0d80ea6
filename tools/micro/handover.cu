// Micro-benchmark (r02e): how long does an SM take to hand over from an exiting CTA to the next CTA of the same grid?
// grid = 2 waves of one-CTA-per-SM blocks; every CTA spins ~8 us, stamps globaltimer at entry and just before exit.
// Variants: threads, dynamic shared memory, tensor-memory allocation.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ long long gns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int TMEM>
__global__ void k(long long* out, int spin_ns, int touch_smem) {
  extern __shared__ __align__(16) uint8_t sm[];
  __shared__ uint32_t slot;
  long long t0 = 0;
  if (threadIdx.x == 0) { t0 = gns(); uint32_t s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); out[blockIdx.x * 4 + 3] = s; out[blockIdx.x * 4] = t0; }
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(TMEM) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  if (touch_smem) for (int i = threadIdx.x; i < touch_smem; i += blockDim.x) sm[i] = (uint8_t)i;
  if (threadIdx.x == 0) { while (gns() - t0 < spin_ns) {} }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x * 4 + 2] = gns();
  if (TMEM) { if (threadIdx.x < 32) { uint32_t a = slot; asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(TMEM) : "memory"); } }
}
template <int TMEM>
void run(const char* name, int threads, int smem, int touch) {
  int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int grid = nsm * 3;
  long long* d; cudaMalloc(&d, grid * 4 * sizeof(long long));
  cudaFuncSetAttribute(k<TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 3; ++rep) { cudaMemset(d, 0, grid * 4 * sizeof(long long)); k<TMEM><<<grid, threads, smem>>>(d, 8000, touch ? smem : 0); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<long long> h(grid * 4); cudaMemcpy(h.data(), d, grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
  struct Ev { long long sm, entry, exit; }; std::vector<Ev> ev;
  for (int i = 0; i < grid; ++i) ev.push_back({h[i * 4 + 3], h[i * 4], h[i * 4 + 2]});
  std::sort(ev.begin(), ev.end(), [](const Ev& a, const Ev& b) { return a.sm != b.sm ? a.sm < b.sm : a.entry < b.entry; });
  std::vector<double> gaps;
  for (size_t i = 0; i + 1 < ev.size(); ++i) if (ev[i].sm == ev[i + 1].sm) gaps.push_back((ev[i + 1].entry - ev[i].exit) / 1e3);
  std::sort(gaps.begin(), gaps.end());
  if (gaps.empty()) { printf("%s: no same-SM pairs\n", name); return; }
  printf("%-44s threads %4d smem %6d tmem %3d: hand-over us p10 %.2f median %.2f p90 %.2f (%zu pairs)\n", name, threads, smem, TMEM, gaps[gaps.size() / 10], gaps[gaps.size() / 2], gaps[gaps.size() * 9 / 10], gaps.size());
  cudaFree(d);
}
int main() {
  run<0>("small CTA", 256, 16 * 1024, 0);
  run<0>("576 threads, small smem", 576, 16 * 1024, 0);
  run<0>("256 threads, 224 KB smem (untouched)", 256, 224 * 1024, 0);
  run<0>("576 threads, 224 KB smem (untouched)", 576, 224 * 1024, 0);
  run<0>("576 threads, 224 KB smem (written)", 576, 224 * 1024, 1);
  run<0>("576 threads, 120 KB smem (written)", 576, 120 * 1024, 1);
  run<512>("576 threads, 224 KB smem, 512 TMEM cols", 576, 224 * 1024, 1);
  run<128>("576 threads, 224 KB smem, 128 TMEM cols", 576, 224 * 1024, 1);
  run<512>("576 threads, 16 KB smem, 512 TMEM cols", 576, 16 * 1024, 0);
  run<512>("128 threads, 16 KB smem, 512 TMEM cols", 128, 16 * 1024, 0);
  return 0;
}
