// Microbenchmark: throughput of the squared-difference accumulation  acc += (q - s)^2  on sm_100a,
// scalar (FADD + FFMA per element) vs packed (FADD2 + FFMA2 per TWO elements, PTX add/fma .f32x2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int NACC>
__global__ void __launch_bounds__(256) k_scalar(float* out, int iters, float q0) {
    float acc[NACC], s[NACC];
    for (int j = 0; j < NACC; ++j) { acc[j] = 0.f; s[j] = 0.001f * (threadIdx.x + j); }
    float q = q0;
    for (int it = 0; it < iters; ++it) {
        q = fmaf(q, 1.0001f, 0.01f);
#pragma unroll
        for (int j = 0; j < NACC; ++j) { float df = q - s[j]; acc[j] = fmaf(df, df, acc[j]); }
    }
    float r = 0.f;
    for (int j = 0; j < NACC; ++j) r += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_packed(float* out, int iters, float q0) {
    u64 acc[NACC / 2], s[NACC / 2];
    for (int j = 0; j < NACC / 2; ++j) { acc[j] = 0ull; s[j] = pack(0.001f * (threadIdx.x + 2 * j), 0.001f * (threadIdx.x + 2 * j + 1)); }
    float q = q0;
    for (int it = 0; it < iters; ++it) {
        q = fmaf(q, 1.0001f, 0.01f);
        const u64 qq = pack(q, q);
#pragma unroll
        for (int j = 0; j < NACC / 2; ++j) { u64 df = sub2(qq, s[j]); acc[j] = fma2(df, df, acc[j]); }
    }
    float r = 0.f;
    for (int j = 0; j < NACC / 2; ++j) { float a, b; unpack(acc[j], a, b); r += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    constexpr int NACC = 64;
    const double elems = (double)blocks * threads * iters * NACC;
    float ms_s = time_ms([&] { k_scalar<NACC><<<blocks, threads>>>(out, iters, 0.5f); });
    float ms_p = time_ms([&] { k_packed<NACC><<<blocks, threads>>>(out, iters, 0.5f); });
    printf("{\"scalar_Gelem_s\": %.1f, \"packed_f32x2_Gelem_s\": %.1f, \"speedup\": %.3f, \"scalar_ms\": %.3f, \"packed_ms\": %.3f, "
           "\"note\": \"elem = one (q-s)^2 accumulate = 3 flops; 148x8 CTAs x 256 thr x %d iters x %d acc\"}\n",
           elems / ms_s * 1e-6, elems / ms_p * 1e-6, ms_s / ms_p, ms_s, ms_p, iters, NACC);
    return 0;
}
