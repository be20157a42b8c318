// Minimal stand-in for the CUDA samples' helper_cuda.h (not shipped with the toolkit).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define checkCudaErrors(val) ugref_check((val), #val, __FILE__, __LINE__)
static inline void ugref_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error %s at %s:%d: %s\n", cudaGetErrorString(e), file, line, what);
        exit(3);
    }
}
#define getLastCudaError(msg) ugref_check(cudaGetLastError(), msg, __FILE__, __LINE__)
static inline int findCudaDevice(int, const char**) {
    cudaSetDevice(0);
    return 0;
}
