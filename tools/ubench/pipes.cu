// Micro-benchmark of sm_100a issue/pipe throughput for the instruction kinds the K1/K2 kernels
// are made of (tools only; not part of the product library).  Prints warp-instructions per clock
// per SM sub-partition for each mix, from clock64() deltas inside the kernel.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
#define DECL8 float a0 = s.x; float a1 = s.y; float a2 = s.x + 1.f; float a3 = s.y + 2.f; float a4 = s.x + 3.f; float a5 = s.y + 4.f; float a6 = s.x + 5.f; float a7 = s.y + 6.f
#define PK(a, b) pk((a), (b))
#define DECL8P u64 p0 = PK(s.x, s.y); u64 p1 = PK(s.y, s.x); u64 p2 = PK(s.x + 1.f, s.y); u64 p3 = PK(s.y + 2.f, s.x); u64 p4 = PK(s.x, 3.f); u64 p5 = PK(s.y, 4.f); u64 p6 = PK(5.f, s.x); u64 p7 = PK(s.y, 6.f)
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }

#define KERNEL(name, BODY, SETUP, FIN)                                                              \
    __global__ void __launch_bounds__(1024) name(const float2* in, float* out, long long* cyc) {      \
        __shared__ float sh[2048];                                                                   \
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = (float)i;                              \
        __syncthreads();                                                                             \
        float2 s = in[threadIdx.x & 31];                                                             \
        const float b = s.x * 0.999f, c = s.y * 0.001f;                                              \
        const u64 B = pk(b, b), C = pk(c, c);                                                        \
        (void)B; (void)C; (void)b; (void)c;                                                          \
        SETUP;                                                                                       \
        long long t0 = clock64();                                                                    \
        _Pragma("unroll 1") for (int it = 0; it < ITERS; it++) { BODY; }                             \
        long long t1 = clock64();                                                                    \
        out[blockIdx.x * 1024 + threadIdx.x] = FIN;                                                   \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                             \
    }

#define F(x) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(b), "f"(c))
#define F2(x) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(B), "l"(C))
#define M2(x) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(B))
#define MN(x) asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(b))
#define MN3(x) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(b), "f"(c))
#define RCP(x) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x))
#define SHF(x) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(x))
#define LDS64(x, o) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y##x) : "r"(saddr + (o)))
#define LDS32(x, o) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(saddr + (o)))
#define DF(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(db), "d"(dc))
#define CVT(d, x) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x))
#define CVTB(x, d) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x) : "d"(d))
#define MX(x) asm volatile("max.f32 %0, %0, %1;" : "+f"(x) : "f"(c))
#define FA(x) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c))
#define FM(x) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(b))
#define IA(x) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(ib))
#define LOP(x) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(ib))
#define MNAN(x) asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(x) : "f"(b))
#define A2(x) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(C))
#define SETSEL(x) asm volatile("{.reg .pred q; setp.gt.f32 q, %0, %1; selp.f32 %0, %1, %0, q;}" : "+f"(x) : "f"(b))
#define SEL(x) asm volatile("{.reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(x) : "f"(b), "r"(pr))

#define YA8 float ya0; float ya1; float ya2; float ya3; float ya4; float ya5; float ya6; float ya7
#define DD8 double d0 = s.x; double d1 = s.y; double d2 = 1 + s.x; double d3 = 2 + s.y; double d4 = 3; double d5 = 4; double d6 = 5; double d7 = 6; double db = b; double dc = c
#define DD4 double d0 = 0; double d1 = 1; double d2 = 2; double d3 = 3
#define ALL8(OP) OP(a0); OP(a1); OP(a2); OP(a3); OP(a4); OP(a5); OP(a6); OP(a7)
#define ALL8P(OP) OP(p0); OP(p1); OP(p2); OP(p3); OP(p4); OP(p5); OP(p6); OP(p7)
#define SUM8 (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7)
#define DECL8I int i0 = threadIdx.x; int i1 = i0 + 1; int i2 = i0 * 3; int i3 = i0 ^ 5; int i4 = i0 + 7; int i5 = i0 * 5; int i6 = i0 + 11; int i7 = i0 * 7; int ib = (int)(s.x * 100.f)
#define ALL8I(OP) OP(i0); OP(i1); OP(i2); OP(i3); OP(i4); OP(i5); OP(i6); OP(i7)
#define SUM8I (float)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7)
#define DECL8B float e0 = s.x; float e1 = s.y; float e2 = s.x * 2.f; float e3 = s.y * 2.f; float e4 = s.x * 3.f; float e5 = s.y * 3.f; float e6 = s.x * 4.f; float e7 = s.y * 4.f
#define ALL8B(OP) OP(e0); OP(e1); OP(e2); OP(e3); OP(e4); OP(e5); OP(e6); OP(e7)
#define SUM8B (e0 + e1 + e2 + e3 + e4 + e5 + e6 + e7)
#define SUM8P (lo(p0) + lo(p1) + lo(p2) + lo(p3) + lo(p4) + lo(p5) + lo(p6) + lo(p7))


KERNEL(k_ffma, ALL8(F); ALL8(F), DECL8, SUM8)
KERNEL(k_ffma2, ALL8P(F2); ALL8P(F2), DECL8P, SUM8P)
KERNEL(k_fadd, ALL8(FA); ALL8(FA), DECL8, SUM8)
KERNEL(k_fmul, ALL8(FM); ALL8(FM), DECL8, SUM8)
KERNEL(k_fadd2, ALL8P(A2); ALL8P(A2), DECL8P, SUM8P)
KERNEL(k_minmax, ALL8(MN); ALL8(MX), DECL8, SUM8)
KERNEL(k_mnan, ALL8(MNAN); ALL8(MX), DECL8, SUM8)
KERNEL(k_fmnmx3, ALL8(MN3); ALL8(MN3), DECL8, SUM8)
KERNEL(k_sel, ALL8(SEL); ALL8(SEL), DECL8; int pr = (s.x > 0.5f), SUM8)
KERNEL(k_setsel, ALL8(SETSEL), DECL8, SUM8)
KERNEL(k_iadd, ALL8I(IA); ALL8I(IA), DECL8I, SUM8I)
KERNEL(k_lop, ALL8I(LOP); ALL8I(LOP), DECL8I, SUM8I)
KERNEL(k_ffma_minmax, ALL8(F); ALL8B(MN); ALL8(F); ALL8B(MX), DECL8; DECL8B, SUM8 + SUM8B)
KERNEL(k_ffma_sel, ALL8(F); ALL8B(SEL); ALL8(F); ALL8B(SEL), DECL8; DECL8B; int pr = (s.x > 0.5f), SUM8 + SUM8B)
KERNEL(k_ffma_iadd, ALL8(F); ALL8I(IA); ALL8(F); ALL8I(IA), DECL8; DECL8I, SUM8 + SUM8I)
KERNEL(k_ffma2_minmax, ALL8P(F2); ALL8B(MN); ALL8P(F2); ALL8B(MX), DECL8P; DECL8B, SUM8P + SUM8B)
KERNEL(k_ffma2_2minmax, ALL8P(F2); ALL8B(MN); ALL8B(MX), DECL8P; DECL8B, SUM8P + SUM8B)
KERNEL(k_ffma2_sel, ALL8P(F2); ALL8B(SEL); ALL8P(F2); ALL8B(SEL), DECL8P; DECL8B; int pr = (s.x > 0.5f), SUM8P + SUM8B)
KERNEL(k_ffma2_iadd, ALL8P(F2); ALL8I(IA); ALL8P(F2); ALL8I(IA), DECL8P; DECL8I, SUM8P + SUM8I)
KERNEL(k_ffma2_2iadd, ALL8P(F2); ALL8I(IA); ALL8I(LOP), DECL8P; DECL8I, SUM8P + SUM8I)
KERNEL(k_ffma2_fadd, ALL8P(F2); ALL8B(FA); ALL8P(F2); ALL8B(FA), DECL8P; DECL8B, SUM8P + SUM8B)
KERNEL(k_ffma2_ffma, ALL8P(F2); ALL8B(F); ALL8P(F2); ALL8B(F), DECL8P; DECL8B, SUM8P + SUM8B)
KERNEL(k_ffma2_fmnmx3, ALL8P(F2); ALL8B(MN3); ALL8P(F2); ALL8B(MN3), DECL8P; DECL8B, SUM8P + SUM8B)
KERNEL(k_rcp, ALL8(RCP), DECL8, SUM8)
KERNEL(k_ffma2_rcp, ALL8P(F2); RCP(a0); RCP(a1); ALL8P(F2); RCP(a2); RCP(a3), DECL8; DECL8P, SUM8 + SUM8P)
KERNEL(k_shfl, ALL8(SHF), DECL8, SUM8)
KERNEL(k_ffma2_shfl, ALL8P(F2); SHF(a0); SHF(a1); ALL8P(F2); SHF(a2); SHF(a3), DECL8; DECL8P, SUM8 + SUM8P)
KERNEL(k_lds64,
       LDS64(a0, 0); LDS64(a1, 256); LDS64(a2, 512); LDS64(a3, 768); LDS64(a4, 1024); LDS64(a5, 1280); LDS64(a6, 1536); LDS64(a7, 1792); acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + ya0 + ya7,
       DECL8; YA8; float acc = 0.f; unsigned saddr = (unsigned)__cvta_generic_to_shared(sh) + 8 * (threadIdx.x & 31),
       acc + ya1 + ya2 + ya3 + ya4 + ya5 + ya6)
KERNEL(k_ffma2_lds64,
       ALL8P(F2); LDS64(a0, 0); LDS64(a1, 256); LDS64(a2, 512); LDS64(a3, 768); ALL8P(F2); LDS64(a4, 1024); LDS64(a5, 1280); LDS64(a6, 1536); LDS64(a7, 1792),
       DECL8; DECL8P; YA8; unsigned saddr = (unsigned)__cvta_generic_to_shared(sh) + 8 * (threadIdx.x & 31),
       SUM8 + SUM8P + ya0 + ya1 + ya2 + ya3 + ya4 + ya5 + ya6 + ya7)
KERNEL(k_dfma, DF(d0); DF(d1); DF(d2); DF(d3); DF(d4); DF(d5); DF(d6); DF(d7), DD8, (float)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7))
KERNEL(k_cvt, CVT(d0, a0); CVT(d1, a1); CVT(d2, a2); CVT(d3, a3); CVTB(a0, d1); CVTB(a1, d2); CVTB(a2, d3); CVTB(a3, d0), DECL8; DD4, (float)(d0 + d1 + d2 + d3) + SUM8)
KERNEL(k_ffma2_cvt, ALL8P(F2); CVT(d0, a0); CVTB(a1, d1); ALL8P(F2); CVT(d2, a2); CVTB(a3, d3), DECL8; DECL8P; DD4, (float)(d0 + d1 + d2 + d3) + SUM8 + SUM8P)
KERNEL(k_ffma2_dfma, ALL8P(F2); DF(d0); DF(d1); ALL8P(F2); DF(d2); DF(d3), DECL8P; DD8, (float)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7) + SUM8P)
struct Case { const char* name; void (*fn)(const float2*, float*, long long*); int ninstr; };
#define C(k) {#k, k, 0}
int main() {
    float2* in; float* out; long long* cyc;
    cudaMalloc(&in, 64 * sizeof(float2)); cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
    float2 h[64]; for (int i = 0; i < 64; i++) h[i] = make_float2(0.5f + i * 0.001f, 0.25f + i * 0.002f);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    Case cases[] = {C(k_ffma), C(k_ffma2), C(k_fadd), C(k_fmul), C(k_fadd2), C(k_minmax), C(k_mnan), C(k_fmnmx3), C(k_sel), C(k_setsel), C(k_iadd), C(k_lop), C(k_ffma_minmax), C(k_ffma_sel), C(k_ffma_iadd), C(k_ffma2_minmax), C(k_ffma2_2minmax), C(k_ffma2_sel), C(k_ffma2_iadd), C(k_ffma2_2iadd), C(k_ffma2_fadd), C(k_ffma2_ffma), C(k_ffma2_fmnmx3), C(k_rcp), C(k_ffma2_rcp), C(k_shfl), C(k_ffma2_shfl), C(k_lds64), C(k_ffma2_lds64), C(k_dfma), C(k_cvt), C(k_ffma2_cvt), C(k_ffma2_dfma)};
    for (int wps = 1; wps <= 4; wps *= 2) {            // CTAs of 8 warps per SM: 8, 16, 32 warps per SM
        printf("---- %d warps per SM (%d per sub-partition)\n", 8 * wps, 2 * wps);
        for (auto& c : cases) {
            const int grid = 148;
            c.fn<<<grid, 256 * wps>>>(in, out, cyc);
            cudaDeviceSynchronize();
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            c.fn<<<grid, 256 * wps>>>(in, out, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long hc[148 * 8]; cudaMemcpy(hc, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; i++) avg += (double)hc[i]; avg /= grid;
            // warp-instructions issued per sub-partition = (8*wps/4) warps * ITERS * ninstr
            const double cpi = avg / ITERS / (2.0 * wps);
            printf("%-18s %7.2f SMSP-cyc/body  (%.3f ms, err=%d)\n", c.name, cpi, ms, (int)cudaGetLastError());
        }
    }
    return 0;
}
