// Stand-alone probe for tcgen05.mma kind::tf32 descriptor conventions (one CTA, one 128x128x8 MMA).
// D[m][n] = sum_k A[m][k] B[k][n], A[m][k] = (k == m % 8), B[k][n] = 1000 k + n  => D[m][n] = 1000 (m%8) + n
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
struct Sh {
    float a[128 * 8];
    float b[8 * 128];
    uint64_t bar;
    uint32_t tmem;
};
// variant bit0: B MN-major (else K-major); bit1: use mask-operand instruction form; bit2: swap LBO/SBO of A; bit3: swap LBO/SBO of B
__global__ void probe(float* out, int variant) {
    __shared__ __align__(128) Sh sh;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const bool b_mn = variant & 1;
    // A image, K-major no swizzle: [kcore 2][mgroup 16][8 rows][4 k]
    for (int e = t; e < 1024; e += 128) {
        int kk = e & 3, row8 = (e >> 2) & 7, mg = (e >> 5) & 15, kc = e >> 9;
        int m = mg * 8 + row8, k = kc * 4 + kk;
        sh.a[e] = (k == (m % 8)) ? 1.f : 0.f;
    }
    for (int e = t; e < 1024; e += 128) {
        if (b_mn) {  // [ncore 32][8 k][4 n]
            int nn = e & 3, k = (e >> 2) & 7, nc = e >> 5;
            sh.b[e] = 1000.f * k + (nc * 4 + nn);
        } else {     // K-major: [kcore 2][ngroup 16][8 n-rows][4 k]
            int kk = e & 3, row8 = (e >> 2) & 7, ng = (e >> 5) & 15, kc = e >> 9;
            sh.b[e] = 1000.f * (kc * 4 + kk) + (ng * 8 + row8);
        }
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = sh.tmem;
    if (t == 0) {
        uint32_t a_lbo = 2048, a_sbo = 128, b_lbo, b_sbo;
        if (b_mn) { b_lbo = 4096; b_sbo = 128; } else { b_lbo = 2048; b_sbo = 128; }
        if (variant & 4) { uint32_t x = a_lbo; a_lbo = a_sbo; a_sbo = x; }
        if (variant & 8) { uint32_t x = b_lbo; b_lbo = b_sbo; b_sbo = x; }
        const uint64_t da = make_desc(smem_u32(sh.a), a_lbo, a_sbo);
        const uint64_t db = make_desc(smem_u32(sh.b), b_lbo, b_sbo);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn ? 1u : 0u) << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        if (variant & 2) {
            uint32_t z = 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                         ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh.bar)) : "memory");
    }
    // wait
    asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra WL;\n\tDN:\n\t}"
                 ::"r"(smem_u32(&sh.bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cc = 0; cc < 128; cc += 32) {
        uint32_t v[32];
        const uint32_t taddr = tm + ((uint32_t)(w * 32) << 16) + cc;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(w * 32 + lane) * 128 + cc + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128) : "memory");
}
int main() {
    float* d; cudaMalloc(&d, 128 * 128 * 4);
    static float h[128 * 128];
    for (int variant = 0; variant < 16; ++variant) {
        cudaMemset(d, 0xff, 128 * 128 * 4);
        probe<<<1, 128>>>(d, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int ok = 0, nz = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float ex = 1000.f * (m % 8) + n; ok += (h[m * 128 + n] == ex); nz += (h[m * 128 + n] != 0.f); }
        printf("variant %2d (B %s, %s form, A lbo/sbo %s, B lbo/sbo %s): err=%s correct=%d/16384 nonzero=%d  D[1][0..5]= %g %g %g %g %g %g  D[9][3]=%g D[77][100]=%g\n",
               variant, (variant & 1) ? "MN" : "K ", (variant & 2) ? "mask" : "plain", (variant & 4) ? "swapped" : "as-is", (variant & 8) ? "swapped" : "as-is",
               cudaGetErrorString(e), ok, nz, h[128], h[129], h[130], h[131], h[132], h[133], h[9 * 128 + 3], h[77 * 128 + 100]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
