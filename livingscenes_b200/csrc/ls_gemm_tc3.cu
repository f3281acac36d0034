// k_gemm_tc3: tcgen05 3xTF32 GEMM with the ACTIVATION operand in tensor memory (TS form).
//
//     C[r, (b,n)] = sum_k W[r,k] * X[b,k,n]          (same contract as ls_gemm.cu / ls_gemm_tc.cu)
//
// Why a third kernel (measurements: profiles/microbench/tc_probe_ts.cu, profiles/r02/experiments.md section 4):
//   * one tcgen05.mma kind::tf32 128 x N x 8 costs ~105 cycles for N <= 128 whatever its operands (issue floor),
//     171 cycles for N = 256 with both operands in shared memory (operand-fetch bound) and 138 cycles for N = 256
//     with A in tensor memory: only the last form gets near the tensor pipe's 128-cycle floor;
//   * k_gemm_tc2 (SS form, N = 128) moved ~250 B/cycle through shared memory per SM (TMA writes of both operands,
//     the raw -> hi/lo transform's LDS + STS, 6 operand reads per k-step) against 128 B/cycle available: its tensor
//     pipe stayed below 50 % active on the SDF decoder.
// Here the roles are swapped: D[point (TMEM lane)][output row (TMEM column)] = A[point][k] * B[row][k]^T
//   * A = activations: the transform warps read X straight from global memory (lanes = consecutive columns:
//     coalesced, any instance geometry, no tensor map), split hi/lo in registers and tcgen05.st them into a TMEM ring
//     (32 columns per k-block of 16: hi | lo) -- no shared-memory traffic at all for the activations;
//   * B = weights: pre-split, pre-tiled images ([kcore][row group][8 rows][4 k], K-major, no swizzle) of NT = 128 or
//     256 output rows, one cp.async.bulk per k-block into a 4-stage ring;
//   * one thread issues tcgen05.mma [d], [a_tmem], b_desc (lo*hi + hi*lo + hi*hi per k-step);
//   * epilogue: thread = point, registers = 32 consecutive output rows.  Point-major gather tables become eight
//     16-byte stores per thread (one full 128-byte line); channel-major outputs are 32 warp-coalesced 4-byte stores.
//   * NT = 128: two accumulator buffers (epilogue of tile i overlaps the MMAs of tile i+1), used for K < 128 where the
//     epilogue dominates; NT = 256 (K >= 128, R > 128): one 256-column accumulator drained by both epilogue groups.
// Warps: 0-7 epilogue (two groups of 4), 8-23 transform (4 groups of 4, k-blocks round robin), 24 weight producer,
// 25 MMA issuer.  TMEM: accumulators in columns 0-255, A ring in columns 256-383.
#include <cuda.h>

#include <atomic>

#include "ls_common.cuh"

namespace ls {
namespace {

constexpr int TM3 = 128, TKB3 = 16;
constexpr int G3_S = 4;          // weight stages == TMEM A stages
constexpr int G3_XF_GROUPS = 4;  // INVARIANT (see k_gemm_tc2): a ring stage is always served by the same group
static_assert(G3_S % G3_XF_GROUPS == 0, "a ring stage must always be served by the same transform group");
constexpr int G3_XF_WARP0 = 8, G3_PROD_WARP = G3_XF_WARP0 + 4 * G3_XF_GROUPS, G3_MMA_WARP = G3_PROD_WARP + 1;
constexpr int G3_THREADS = 32 * (G3_MMA_WARP + 1);
constexpr int G3_A_COL0 = 256, G3_A_COLS = 2 * TKB3;  // TMEM A ring: stage s at column 256 + 32 s (hi 16 | lo 16)
constexpr int G3_TMEM_COLS = 512;
constexpr int G3_EPI_LD = 36;  // padded row of the epilogue transpose tile: 16-byte aligned, conflict-free float4 rows

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "G3_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra G3_DONE;\n\t"
        "bra G3_WAIT;\n\t"
        "G3_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x 8 columns of tf32 (lane = row, column = k)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

template <int NT>
struct G3Shared {
    float w[G3_S][2 * NT * TKB3];  // [stage][hi image | lo image], image = [kcore 4][NT/8 row groups][8 rows][4 k]
    float epi[8][32 * G3_EPI_LD];  // per epilogue warp: 32 points x 32 channels (point-major stores go out transposed)
    long long epi_base[8][32];     // per epilogue warp: output offset of each of its 32 points (-1: column out of range)
    uint64_t full_w[G3_S], full_a[G3_S], empty[G3_S], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <bool PM, int NT>
__global__ void __launch_bounds__(G3_THREADS, 1) k_gemm_tc3(const GemmArgs a, const float* __restrict__ wpk, int n_kb, int n_nt,
                                                            int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    G3Shared<NT>& sh = *reinterpret_cast<G3Shared<NT>*>(smem_raw);
    constexpr int NACC = NT == 128 ? 2 : 1;  // accumulator buffers
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long ncols = (long long)a.B * a.n_per_b;
    const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (t == 0) {
        for (int s = 0; s < G3_S; ++s) {
            mbar_init(&sh.full_w[s], 1);
            mbar_init(&sh.full_a[s], 128);
            mbar_init(&sh.empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.tmem_full[i], 1);
            mbar_init(&sh.tmem_empty[i], NT == 128 ? 128 : 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == G3_MMA_WARP) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "r"(G3_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sh.tmem_base;

    if (w == G3_PROD_WARP) {
        // ================================================================ weight images: one bulk copy per k-block
        if (lane == 0) {
            constexpr uint32_t BYTES = 2 * NT * TKB3 * sizeof(float);
            int it = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int tile = blockIdx.x + lt * (int)gridDim.x;
                const float* wt = wpk + (size_t)(tile % n_nt) * n_kb * (2 * NT * TKB3);
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % G3_S;
                    if (it >= G3_S) mbar_wait(&sh.empty[s], ((it / G3_S) - 1) & 1);
                    mbar_arrive_expect_tx(&sh.full_w[s], BYTES);
                    bulk_g2s(&sh.w[s][0], wt + (size_t)kb * (2 * NT * TKB3), BYTES, &sh.full_w[s]);
                }
            }
        }
    } else if (w == G3_MMA_WARP) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((TM3 >> 4) << 24);
            constexpr uint32_t LBO = (NT / 8) * 128;  // k-core stride of a weight image
            int it = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int acc = NACC == 2 ? (lt & 1) : 0;
                if (lt >= NACC) {  // the epilogue must have drained this accumulator buffer
                    mbar_wait(&sh.tmem_empty[acc], ((lt / NACC) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t d = tmem + (uint32_t)(acc * 128);
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % G3_S;
                    mbar_wait(&sh.full_w[s], (it / G3_S) & 1);
                    mbar_wait(&sh.full_a[s], (it / G3_S) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_hi = smem_u32(&sh.w[s][0]), b_lo = b_hi + NT * TKB3 * 4;
                    const uint32_t a_hi = tmem + (uint32_t)(G3_A_COL0 + s * G3_A_COLS), a_lo = a_hi + TKB3;
#pragma unroll
                    for (int ks = 0; ks < TKB3 / 8; ++ks) {
                        const uint64_t dbh = make_desc(b_hi + ks * 2 * LBO, LBO, 128);
                        const uint64_t dbl = make_desc(b_lo + ks * 2 * LBO, LBO, 128);
                        umma_ts(d, a_lo + ks * 8, dbh, IDESC, (kb | ks) != 0);
                        umma_ts(d, a_hi + ks * 8, dbl, IDESC, 1);
                        umma_ts(d, a_hi + ks * 8, dbh, IDESC, 1);
                    }
                    umma_commit(&sh.empty[s]);  // frees the weight stage and the TMEM A stage when the MMAs have read them
                    if (kb == n_kb - 1) umma_commit(&sh.tmem_full[acc]);
                }
            }
        }
    } else if (w >= G3_XF_WARP0) {
        // ================================================================ transform: global fp32 -> hi/lo TF32 in TMEM
        const int q = (w - G3_XF_WARP0) & 3;     // TMEM lane quadrant (== warp id % 4: G3_XF_WARP0 is a multiple of 4)
        const int grp = (w - G3_XF_WARP0) >> 2;  // handles the k-blocks with it % G3_XF_GROUPS == grp
        const int total = my_tiles * n_kb;
        int cur_lt = -1;
        const float* xp = nullptr;
        auto load = [&](int it, float* r) {
            const int lt = it / n_kb, kb = it - lt * n_kb;
            if (lt != cur_lt) {  // this thread's column of the tile: flattened (instance, n) index
                cur_lt = lt;
                const int tile = blockIdx.x + lt * (int)gridDim.x;
                const long long j = (long long)(tile / n_nt) * TM3 + q * 32 + lane;
                xp = nullptr;
                if (j < ncols) {
                    const long long b = j / a.n_per_b;
                    xp = a.X + b * a.x_sb + (j - b * a.n_per_b);
                }
            }
            const int k0 = kb * TKB3;
#pragma unroll
            for (int i = 0; i < TKB3; ++i)
                r[i] = (xp != nullptr && k0 + i < a.K) ? __ldg(xp + (long long)(k0 + i) * a.x_sk) : 0.f;
        };
        float nxt[TKB3];
        if (grp < total) load(grp, nxt);
        for (int it = grp; it < total; it += G3_XF_GROUPS) {
            uint32_t v[2 * TKB3];
#pragma unroll
            for (int i = 0; i < TKB3; ++i) {
                // hi = x rounded to the nearest TF32 number, lo = x - hi exactly (signed: the tensor core's truncation of
                // lo does not accumulate a bias over K)
                const float x = nxt[i];
                const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
                v[i] = h;
                v[TKB3 + i] = __float_as_uint(x - __uint_as_float(h));
            }
            if (it + G3_XF_GROUPS < total) load(it + G3_XF_GROUPS, nxt);  // next k-block of this group: in flight during the wait
            const int s = it % G3_S;
            if (it >= G3_S) {
                mbar_wait(&sh.empty[s], ((it / G3_S) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(G3_A_COL0 + s * G3_A_COLS);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
                "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
                "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
                "r"(v[31])
                : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&sh.full_a[s]);
        }
    } else {
        // ================================================================ epilogue: TMEM -> registers -> global
        // thread = point (TMEM lane), registers = 32 consecutive output rows.  NT = 128: group eg owns accumulator buffer
        // eg (the CTA's tiles eg, eg + 2, ...); NT = 256: both groups drain every tile, group eg takes rows 128 eg ...
        const int eg = w >> 2, wq = w & 3;
        const long long r3 = (long long)a.R * 3;
        int use = 0;
        for (int lt = (NACC == 2 ? eg : 0); lt < my_tiles; lt += NACC, ++use) {
            const int tile = blockIdx.x + lt * (int)gridDim.x;
            const int nt = tile % n_nt;
            const long long j = (long long)(tile / n_nt) * TM3 + wq * 32 + lane;
            const bool col_ok = j < ncols;
            long long b = 0;
            int n = 0, axis = 0;
            if (col_ok) {
                b = j / a.n_per_b;
                n = (int)(j - b * a.n_per_b);
                axis = a.npts > 0 ? n / a.npts : 0;
            }
            const int acc = NACC == 2 ? eg : 0;
            const int rbase = nt * NT + (NACC == 2 ? 0 : eg * 128);
            float* obase;
            const float* bbase = nullptr;
            const float* mbase = nullptr;
            if (PM) {
                obase = a.out;
                __syncwarp();  // the previous tile's table is no longer read
                sh.epi_base[w][lane] = col_ok ? (b * a.npts + (n - axis * a.npts)) * r3 + (long long)axis * a.c_out : -1;
                __syncwarp();
            } else {
                obase = a.out + b * a.o_sb + n;
                if (a.bias) bbase = a.bias + b * a.bias_sb + (a.bias_axis ? axis : 0);
                if (a.mask) mbase = a.mask + b * a.o_sb + n;
            }
            mbar_wait(&sh.tmem_full[acc], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < 128; cc += 32) {
                const int r0 = rbase + cc;
                if (r0 >= a.R) {  // warp uniform: nothing left in this tile for this group
                    if (cc == 0) {  // (cc > 0: the previous chunk was the last valid one and has already arrived)
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(&sh.tmem_empty[acc]);
                    }
                    break;
                }
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * 128 + (NACC == 2 ? 0 : eg * 128) + cc);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == 96 || r0 + 32 >= a.R) {  // this thread's last read of the buffer: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&sh.tmem_empty[acc]);
                }
                if (PM) {
                    // R = parts * c_out with c_out % 32 == 0: the chunk is 32 consecutive channels of one part.  A thread
                    // holds one point's 32 channels; a 16-byte store per thread would touch 32 different lines per
                    // instruction (half a sector each: measured 2.4x slower than k_gemm_tc2's tables).  Transposed through
                    // shared memory, 8 lanes write one point's 128 contiguous bytes and an instruction covers 4 full lines.
                    const int part = r0 / a.c_out;
                    const long long coff = (long long)part * 3 * a.c_out + (r0 - part * a.c_out);
                    float* tl = &sh.epi[w][0];
                    float4* trow = reinterpret_cast<float4*>(tl + lane * G3_EPI_LD);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        trow[j4] = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]), __uint_as_float(v[4 * j4 + 2]),
                                               __uint_as_float(v[4 * j4 + 3]));
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int p = 4 * i + (lane >> 3);
                        const long long pb = sh.epi_base[w][p];
                        const float4 x = *reinterpret_cast<const float4*>(tl + p * G3_EPI_LD + 4 * (lane & 7));
                        if (pb >= 0) *reinterpret_cast<float4*>(obase + pb + coff + 4 * (lane & 7)) = x;
                    }
                    __syncwarp();
                    continue;
                }
                if (!col_ok) continue;
                {
                    const int nr = min(32, a.R - r0);
                    float* o = obase + (long long)r0 * a.o_sr;
                    const float* bp = bbase ? bbase + (long long)r0 * a.bias_sr : nullptr;
                    const float* mp = mbase ? mbase + (long long)r0 * a.o_sr : nullptr;
                    const bool relu = a.relu != 0;
                    if (nr == 32) {
#pragma unroll
                        for (int j8 = 0; j8 < 32; j8 += 8) {  // 8 rows at a time: bias (and mask) loads in flight together
                            float bv[8], mv[8];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                bv[jj] = bp ? __ldg(bp + (long long)(j8 + jj) * a.bias_sr) : 0.f;
                                mv[jj] = mp ? __ldg(mp + (long long)(j8 + jj) * a.o_sr) : 1.f;
                            }
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                float x = __uint_as_float(v[j8 + jj]) + bv[jj];
                                if (relu) x = fmaxf(x, 0.f);
                                if (!(mv[jj] > 0.f)) x = 0.f;
                                o[(long long)(j8 + jj) * a.o_sr] = x;  // lanes = consecutive columns: 128 B per warp store
                            }
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            if (jj < nr) {
                                float x = __uint_as_float(v[jj]);
                                if (bp) x += __ldg(bp + (long long)jj * a.bias_sr);
                                if (relu) x = fmaxf(x, 0.f);
                                if (mp && !(__ldg(mp + (long long)jj * a.o_sr) > 0.f)) x = 0.f;
                                o[(long long)jj * a.o_sr] = x;
                            }
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == G3_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(G3_TMEM_COLS) : "memory");
    }
}

// W [R][ldw] row-major -> per (256-row tile, k-block): hi image then lo image, each [kcore 4][row group 32][8 rows][4 k]
__global__ void k_tc3_pack_weights(const float* __restrict__ W, int R, int K, int ldw, float* __restrict__ out, int n_kb) {
    constexpr int IMG = 256 * TKB3;
    const int nt = blockIdx.y, kb = blockIdx.x;
    float* dst = out + ((size_t)nt * n_kb + kb) * (2 * IMG);
    for (int e = threadIdx.x; e < IMG; e += blockDim.x) {
        const int kk = e & 3, row8 = (e >> 2) & 7, g = (e >> 5) & 31, kc = e >> 10;
        const int r = nt * 256 + g * 8 + row8, k = kb * TKB3 + kc * 4 + kk;
        const float x = (r < R && k < K) ? W[(size_t)r * ldw + k] : 0.f;
        uint32_t u = __float_as_uint(x);
        uint32_t h = (u + 0x00000fffu + ((u >> 13) & 1u)) & 0xffffe000u;  // round to nearest even TF32
        if ((u & 0x7f800000u) == 0x7f800000u) h = u & 0xffffe000u;
        const float hi = __uint_as_float(h);
        dst[e] = hi;
        dst[IMG + e] = x - hi;
    }
}

int sm_count3() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

template <bool PM, int NT>
int launch3(const GemmArgs& a, const float* wpk, int n_kb, cudaStream_t st) {
    const long long ncols = (long long)a.B * a.n_per_b;
    const long long n_ct = (ncols + TM3 - 1) / TM3;
    const int n_nt = (a.R + NT - 1) / NT;
    LS_REQUIRE(n_ct * n_nt < (1LL << 31), "gemm_tc3: too many tiles");
    const int n_tiles = (int)(n_ct * n_nt);
    const int grid = n_tiles < sm_count3() ? n_tiles : sm_count3();
    const size_t smem = sizeof(G3Shared<NT>) + 128;
    LS_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc3<PM, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_tc3<PM, NT><<<grid, G3_THREADS, smem, st>>>(a, wpk, n_kb, n_nt, n_tiles);
    LS_CHECK_LAUNCH("k_gemm_tc3");
    return LS_OK;
}

}  // namespace

// floats of the 256-row image set appended to the 128-row images (0 when the GEMM never uses 256-row tiles)
size_t tc3_packed_floats(int R, int K) {
    if (R <= 128) return 0;
    const size_t nt = (R + 255) / 256, kb = (K + TKB3 - 1) / TKB3;
    return nt * kb * 2 * 256 * TKB3;
}

int tc3_pack_weights(const float* W, int R, int K, int ldw, float* packed256, cudaStream_t st) {
    if (R <= 128) return LS_OK;
    const int n_kb = (K + TKB3 - 1) / TKB3;
    dim3 grid(n_kb, (R + 255) / 256);
    k_tc3_pack_weights<<<grid, 256, 0, st>>>(W, R, K, ldw, packed256, n_kb);
    LS_CHECK_LAUNCH("k_tc3_pack_weights");
    return LS_OK;
}

// packed128: the k_gemm_tc / k_gemm_tc2 images (128-row tiles); packed256: tc3_pack_weights' images (or nullptr)
int launch_gemm_tc3(const GemmArgs& a, const float* packed128, const float* packed256, cudaStream_t st) {
    LS_REQUIRE(packed128 != nullptr, "gemm_tc3: packed weights missing");
    if (a.point_major)
        LS_REQUIRE(a.c_out % 32 == 0 && a.R % a.c_out == 0 && a.npts > 0 && a.n_per_b == 3 * a.npts &&
                       (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
                   "gemm_tc3: bad point-major geometry");
    const int n_kb = (a.K + TKB3 - 1) / TKB3;
    const bool wide = packed256 != nullptr && a.R > 128 && a.K >= 128;
    if (a.point_major) return wide ? launch3<true, 256>(a, packed256, n_kb, st) : launch3<true, 128>(a, packed128, n_kb, st);
    return wide ? launch3<false, 256>(a, packed256, n_kb, st) : launch3<false, 128>(a, packed128, n_kb, st);
}

}  // namespace ls
