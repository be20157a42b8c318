// texref_shim.h -- force-included when compiling the UNMODIFIED reference MatchLib.cu with CUDA 12:
// the texture-reference API it uses (texture<...> globals, cudaBindTextureToArray,
// cudaUnbindTexture, tex2D(texref, x, y)) was removed in CUDA 12.  This header maps it onto
// texture OBJECTS with the attributes a default-constructed texture reference had: point
// filtering, clamp addressing, unnormalised coordinates, element read mode.
// Test infrastructure only (second-opinion reference build, see oracle/refbuild/README.md).
#pragma once
#include <cuda_runtime.h>
#include <cstring>

template <class T, int D, int M>
struct ugref_texref {
    cudaTextureObject_t obj;
};
#define texture __device__ ugref_texref

template <class T, int D, int M>
static __device__ __forceinline__ T tex2D(const ugref_texref<T, D, M>& t, float x, float y) {
    return tex2D<T>(t.obj, x, y);
}

template <class T, int D, int M>
static inline cudaError_t cudaBindTextureToArray(const ugref_texref<T, D, M>& sym, cudaArray* arr) {
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    ugref_texref<T, D, M> h;
    // release the previously bound object of this reference, if any
    if (cudaMemcpyFromSymbol(&h, sym, sizeof h) == cudaSuccess && h.obj) cudaDestroyTextureObject(h.obj);
    h.obj = 0;
    cudaError_t e = cudaCreateTextureObject(&h.obj, &rd, &td, nullptr);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(sym, &h, sizeof h);
}

template <class T, int D, int M>
static inline cudaError_t cudaUnbindTexture(const ugref_texref<T, D, M>& sym) {
    ugref_texref<T, D, M> h;
    cudaError_t e = cudaMemcpyFromSymbol(&h, sym, sizeof h);
    if (e != cudaSuccess) return e;
    if (h.obj) cudaDestroyTextureObject(h.obj);
    h.obj = 0;
    return cudaMemcpyToSymbol(sym, &h, sizeof h);
}
