/*
 * knn_ref_shim.cu — TEST INFRASTRUCTURE, not product code.
 *
 * C-ABI veneer over the UNMODIFIED reference simple-knn
 * (/root/reference/submodules/simple-knn/simple_knn.cu), which oracle/Makefile compiles
 * where the source lies (with `-include cfloat`: the file uses FLT_MAX without including
 * it, which only older toolkits tolerated) and links with this file into
 * oracle/_ref/libknn_ref.so.  No reference source is copied.  Used by the GPU parity
 * tests (bit-exact comparison) and tests/golden/make_knn_golden.py.
 */
#include <cuda_runtime.h>

#include "simple_knn.h"

extern "C" int knn_ref_dist_cuda2(int P, const float* points, float* mean_dist2) {
    /* spatial.cu:15-25 zero-fills the output, then calls SimpleKNN::knn on the legacy stream */
    if (cudaMemset(mean_dist2, 0, sizeof(float) * (size_t)P) != cudaSuccess) return -2;
    SimpleKNN::knn(P, (float3*)points, mean_dist2);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}
