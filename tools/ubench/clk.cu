// Calibration: clock64() ticks vs %globaltimer nanoseconds for a long FFMA loop (is clock64 the SM clock, and how
// many FFMA / FFMA2 warp-instructions per clock does a sub-partition really issue?).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 gtime() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int PACKED>
__global__ void __launch_bounds__(256) k(const float2* in, float* out, long long* cyc, u64* ns, int iters) {
    float2 s = in[threadIdx.x & 31];
    const float b = s.x * 0.999f, c = s.y * 0.001f;
    u64 B, C;
    asm("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1,%2};" : "=l"(C) : "f"(c), "f"(c));
    float a[8];
    u64 p[8];
    for (int i = 0; i < 8; i++) { a[i] = s.x + i; asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(s.y)); }
    u64 g0 = gtime();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(B), "l"(C));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            }
    }
    long long t1 = clock64();
    u64 g1 = gtime();
    float acc = 0;
    for (int i = 0; i < 8; i++) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); acc += a[i] + x + y; }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; ns[blockIdx.x] = g1 - g0; }
}
int main() {
    float2* in; float* out; long long* cyc; u64* ns;
    cudaMalloc(&in, 64 * sizeof(float2)); cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8); cudaMalloc(&ns, 148 * 8 * 8);
    float2 h[64]; for (int i = 0; i < 64; i++) h[i] = make_float2(0.5f + i * 0.001f, 0.25f + i * 0.002f);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 400000;  // 32 FFMA per iteration
    for (int rep = 0; rep < 3; rep++)
        for (int packed = 0; packed < 2; packed++)
            for (int wps = 1; wps <= 4; wps *= 2) {
                const int grid = 148 * wps;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                if (packed) k<1><<<grid, 256>>>(in, out, cyc, ns, iters); else k<0><<<grid, 256>>>(in, out, cyc, ns, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                long long hc[148 * 8]; u64 hn[148 * 8];
                cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hn, ns, grid * 8, cudaMemcpyDeviceToHost);
                double ac = 0, an = 0; for (int i = 0; i < grid; i++) { ac += hc[i]; an += hn[i]; } ac /= grid; an /= grid;
                const double instr_per_smsp = 2.0 * wps * iters * 32.0;
                printf("%s warps/SMSP=%d: event %.3f ms, globaltimer %.3f ms, clock64 %.0f ticks -> %.1f MHz; %s per clock64 tick per SMSP %.3f; per ns*1.965 %.3f\n",
                       packed ? "FFMA2" : "FFMA ", 2 * wps, ms, an * 1e-6, ac, ac / an * 1e3, packed ? "FFMA2" : "FFMA", instr_per_smsp / ac, instr_per_smsp / (an * 1.965));
            }
    return 0;
}
