// host_shim.h -- force-included when compiling the UNMODIFIED reference MatchGPULib.cpp.
// It changes no algorithm; it only makes two undefined behaviours of the reference defined so
// that a run is deterministic and cannot fault:
//   * malloc -> zero-filled: the coarsest-level disparity planes are read before they are
//     written (MatchGPULib.cpp:1247, 1764-1765);
//   * cudaMalloc -> padded + zero-filled: the shared-memory convolution kernels write up to
//     15 pixels / 7 rows past the end of W*H buffers and read below the image on partial tiles
//     (MatchLib.cu:97-100, 121-137, 220-223, 244-260).
// Test infrastructure only.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
static inline void* ugref_zmalloc(size_t n) { return calloc(1, n ? n : 1); }
#define malloc(n) ugref_zmalloc(n)
static inline cudaError_t ugref_padded_cudaMalloc(void** p, size_t n) {
    const size_t pad = 4u << 20;
    cudaError_t e = (cudaMalloc)(p, n + pad);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n + pad);
    return e;
}
#define cudaMalloc(p, n) ugref_padded_cudaMalloc((void**)(p), (n))
