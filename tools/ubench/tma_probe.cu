// Probe: one 2-D TMA tile load (box 40 x 4 f32, origin may be negative / past the edge) completing on an mbarrier,
// issued by lane 0 of a warp, descriptor passed as a __grid_constant__ kernel parameter.  Prints the tile.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
struct P { int x, y; alignas(64) CUtensorMap tm; };
__global__ void probe(const __grid_constant__ P p, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* tile = reinterpret_cast<float*>(sm);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 1024);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(640) : "memory");
        const unsigned d = (unsigned)__cvta_generic_to_shared(tile);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(d), "l"(&p.tm), "r"(p.x), "r"(p.y), "r"(b) : "memory");
    }
    unsigned done = 0; int spins = 0;
    while (!done && spins < 1000000) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
        spins++;
    }
    __syncwarp();
    for (int i = lane; i < 160; i += 32) out[i] = done ? tile[i] : -999.f;
}
int main(int argc, char** argv) {
    const int W = 100, H = 50, pitch = 128;
    float* d; cudaMalloc(&d, pitch * H * 4);
    float* hbuf = new float[pitch * H];
    for (int y = 0; y < H; y++) for (int x = 0; x < pitch; x++) hbuf[y * pitch + x] = y * 1000 + x;
    cudaMemcpy(d, hbuf, pitch * H * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    P p;
    cuuint64_t dims[2] = {W, H}, strides[1] = {pitch * 4};
    cuuint32_t box[2] = {40, 4}, es[2] = {1, 1};
    CUresult r = ((enc)fn)(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    float* out; cudaMalloc(&out, 160 * 4);
    float ho[160];
    const int cases[1][2] = {{argc > 1 ? atoi(argv[1]) : 8, argc > 2 ? atoi(argv[2]) : 2}};
    for (auto& c : cases) {
        p.x = c[0]; p.y = c[1];
        probe<<<1, 32, 2048>>>(p, out);
        e = cudaDeviceSynchronize();
        cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
        printf("origin (%d,%d): %s; row0: %g %g %g %g ... %g | row3: %g %g ... %g\n", c[0], c[1], cudaGetErrorString(e), ho[0], ho[1], ho[2], ho[3], ho[39], ho[120], ho[121], ho[159]);
    }
    return 0;
}
