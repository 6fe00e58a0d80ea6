// tc_gemm.cuh - tcgen05 GEMM (placeholder until the kernel lands)
#pragma once
#include "common.cuh"
namespace fb200 {
inline int tc_gemm_workspace_bytes(int, int, int, int, int, size_t*) { return FB200_EUNSUPPORTED; }
inline int tc_gemm_f32(int, int, int, int, int, const float*, int, const float*, int, float*, int, const float*, int, int,
                       void*, size_t, int, cudaStream_t) { return FB200_EUNSUPPORTED; }
}
