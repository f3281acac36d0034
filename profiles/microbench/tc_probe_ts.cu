// Stand-alone probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (TS form) and B in shared memory.
//   1. correctness: A[m][k] written by tcgen05.st.32x32b (lane = row m, column = k), B K-major no-swizzle image,
//      D[m][n] = sum_k A[m][k] B[n][k] with A[m][k] = (m % 7) + 0.25 k, B[n][k] = (k == n % 8) + (k == 3) * 0.5 n
//   2. throughput: cycles per MMA for back-to-back SS-form and TS-form 128 x N x 8 MMAs (N = 128, 256), one CTA per SM.
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
    float a[256 * 8];  // K-major image of up to 256 rows x 8 k: [kcore 2][group 32][8 rows][4 k]
    float b[256 * 8];
    uint64_t bar;
    uint32_t tmem;
};
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// mode 0: correctness of the TS form (K = 16 as two MMAs, A columns 0-7 and 8-15 of the TMEM A region)
// mode != 0: timing of iters back-to-back MMAs; ts = A from TMEM, N = 64 / 128 / 256, naccs = accumulators used round robin
__global__ void probe(float* out, long long* cycles, int mode, int iters, int ts, int Narg, int naccs) {
    extern __shared__ __align__(128) unsigned char raw[];
    Sh& sh = *reinterpret_cast<Sh*>(raw);
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int N = mode == 0 ? 128 : Narg;
    // B image (rows = n): value B[n][k]
    for (int e = t; e < 256 * 8; e += 128) {
        const int kk = e & 3, row8 = (e >> 2) & 7, g = (e >> 5) % 32, kc = e / (32 * 32);
        const int n = g * 8 + row8, k = kc * 4 + kk;
        sh.b[e] = ((k == (n % 8)) ? 1.f : 0.f) + ((k == 3) ? 0.5f * n : 0.f);
        sh.a[e] = (float)(n % 7) + 0.25f * k;  // A image for the SS timing runs (same layout, rows = m)
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = sh.tmem;
    const uint32_t a_tm = tm + 496;  // A region: columns 496..511
    // ---- A into TMEM: thread = row m, 16 columns (k = 0..15; second k-step uses a different pattern: A2[m][k] = A[m][k] + 100)
    {
        uint32_t v[16];
        const int m = w * 32 + lane;
        for (int k = 0; k < 8; ++k) v[k] = __float_as_uint((float)(m % 7) + 0.25f * k);
        for (int k = 0; k < 8; ++k) v[8 + k] = __float_as_uint((float)(m % 7) + 0.25f * k + 100.f);
        const uint32_t taddr = a_tm + ((uint32_t)(w * 32) << 16);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
            "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
            "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t lbo = (uint32_t)(256 / 8) * 128;  // k-core stride of the 256-row images
    long long t0 = 0, t1 = 0;
    if (mode == 2 && w == 0) {
        // converged warp, elect.sync picks the issuing lane (the CUTLASS pattern)
        const uint64_t da = make_desc(smem_u32(sh.a), lbo, 128);
        const uint64_t db = make_desc(smem_u32(sh.b), lbo, 128);
        if (lane == 0) t0 = clock64();
        int acc = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tm + (uint32_t)(acc * N);
            uint32_t is_leader;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(is_leader));
            if (is_leader) {
                if (ts)
                    mma_ts(d, a_tm + (i & 1) * 8, db, idesc, i >= naccs);
                else
                    mma_ss(d, da, db, idesc, i >= naccs);
            }
            __syncwarp();
            if (++acc == naccs) acc = 0;
        }
        if (lane == 0) commit(&sh.bar);
    } else if (t == 0) {
        const uint64_t da = make_desc(smem_u32(sh.a), lbo, 128);
        const uint64_t db = make_desc(smem_u32(sh.b), lbo, 128);
        if (mode == 0) {
            mma_ts(tm, a_tm, db, idesc, 0);
            mma_ts(tm, a_tm + 8, db, idesc, 1);
            commit(&sh.bar);
        } else {
            t0 = clock64();
            int acc = 0;
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tm + (uint32_t)(acc * N);
                if (ts)
                    mma_ts(d, a_tm + (i & 1) * 8, db, idesc, i >= naccs);
                else
                    mma_ss(d, da, db, idesc, i >= naccs);
                if (++acc == naccs) acc = 0;
            }
            commit(&sh.bar);
        }
    }
    wait(&sh.bar, 0);
    if (t == 0 && mode != 0) {
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0) {
        for (int cc = 0; cc < 128; cc += 32) {
            uint32_t v[32];
            const uint32_t taddr = tm + ((uint32_t)(w * 32) << 16) + cc;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[(w * 32 + lane) * 128 + cc + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
int main() {
    float* d;
    long long* cyc;
    cudaMalloc(&d, 128 * 128 * 4);
    cudaMalloc(&cyc, 148 * 8);
    static float h[128 * 128];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sh) + 128);
    cudaMemset(d, 0xff, 128 * 128 * 4);
    probe<<<1, 128, sizeof(Sh) + 128>>>(d, cyc, 0, 0, 1, 128, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int ok = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            double ex = 0;
            for (int k = 0; k < 8; ++k) {
                const double b = ((k == (n % 8)) ? 1.0 : 0.0) + ((k == 3) ? 0.5 * n : 0.0);
                ex += ((m % 7) + 0.25 * k) * b + ((m % 7) + 0.25 * k + 100.0) * b;
            }
            const double er = fabs(h[m * 128 + n] - ex) / (fabs(ex) + 1e-9);
            if (er < 2e-3) ++ok;
            if (er > maxerr) maxerr = er;
        }
    printf("TS correctness: err=%s correct=%d/16384 max rel err %.3g  D[1][0..3]= %g %g %g %g D[77][100]=%g\n", cudaGetErrorString(e), ok,
           maxerr, h[128], h[129], h[130], h[131], h[77 * 128 + 100]);
    for (int N : {64, 128, 256})
        for (int ts = 0; ts < 2; ++ts)
            for (int naccs : {1, 2, 3}) {
                if (N * naccs > 384) continue;
                const int iters = 4092;
                for (int mode = 1; mode <= 2; ++mode) {
                probe<<<148, 128, sizeof(Sh) + 128>>>(d, cyc, mode, iters, ts, N, naccs);
                e = cudaDeviceSynchronize();
                long long hc[148];
                cudaMemcpy(hc, cyc, 148 * 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
                printf("%s N=%3d accumulators=%d %s: err=%s %.1f cycles per 128xNx8 MMA\n", ts ? "TS" : "SS", N, naccs,
                       mode == 1 ? "single thread" : "elect.sync   ", cudaGetErrorString(e), (double)mx / iters);
                }
            }
    return 0;
}
