// Development probe: throughput of the packed add.rn.f32x2 / mul.rn.f32x2 instructions (FADD2 / FMUL2 / FFMA2) against scalar FADD / FMUL.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/f32x2_probe tools/exp/f32x2_probe.cu && ./tools/exp/f32x2_probe
// Result on a B200: 59 vs 34 G lane-ops/s; note that ptxas contracted the packed mul + add into FFMA2 despite the .rn qualifiers.
#include <cstdio>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}
template <int MODE>
__global__ void k(float* out, int n, float s) {
    float2 a = make_float2(threadIdx.x * 0.001f, threadIdx.x * 0.002f), b = make_float2(s, s * 0.5f), c = make_float2(1.0001f, 0.9999f);
    float2 a2 = make_float2(threadIdx.x * 0.003f, threadIdx.x * 0.004f);
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) {
            a = mul2(add2(a, b), c);
            a2 = mul2(add2(a2, b), c);
        } else {
            a.x = __fmul_rn(__fadd_rn(a.x, b.x), c.x); a.y = __fmul_rn(__fadd_rn(a.y, b.y), c.y);
            a2.x = __fmul_rn(__fadd_rn(a2.x, b.x), c.x); a2.y = __fmul_rn(__fadd_rn(a2.y, b.y), c.y);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y + a2.x + a2.y;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, 100000, 0.5f); else k<1><<<148 * 8, 256>>>(d, 100000, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d: %.3f ms  -> %.1f G lane-op/s\n", mode, ms, 148.0 * 8 * 256 * 100000.0 * 8 / ms / 1e6);
    }
    float h[4]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%f\n", h[1]);
}
