// tcgen05 (5th-gen tensor core) GEMM with fp32-accurate 3xTF32 arithmetic, accumulators in TMEM.
//
//     C[r, (b,n)] = sum_k W[r,k] * X[b,k,n]          (same contract as ls_gemm.cu)
//
// Every fp32 operand is split x = hi + lo with hi = x rounded to the nearest TF32 number and lo = x - hi
// (exact in fp32); three MMAs  lo*hi + hi*lo + hi*hi  accumulate in one fp32 TMEM tile.  SURVEY.md 7.1
// fact 2: this keeps all 7 kNN graphs identical to the fp32 reference while single-pass TF32 does not.
//
// One CTA (128 threads) = one 128(rows) x 128(columns) output tile:
//   * weights are pre-split and pre-tiled by k_tc_pack_weights into the exact canonical UMMA smem image
//     (K-major, no swizzle: [kcore][mgroup][8 rows][4 k]); a k-block is one contiguous 16 KB copy;
//   * activations are read as float4 along the contiguous column axis, transposed 4x4 in registers,
//     split hi/lo and stored as the canonical K-major no-swizzle image ([kcore][ngroup][8 n][4 k]);
//   * 2-stage ring: fence.proxy.async + __syncthreads publishes a stage, ONE thread issues the
//     tcgen05.mma's (M=128, N=128, K=8, kind::tf32) and tcgen05.commit's to the stage's mbarrier so the
//     next use of that stage waits for the tensor core to have consumed it;
//   * epilogue: each warp tcgen05.ld's its 32 TMEM lanes (= 32 weight rows) and stores either the
//     point-major gather table (128 B coalesced across lanes) or the channel-major tensor (+bias, relu).
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
//
// Two kernels share the operand images and descriptors:
//   k_gemm_tc   (round 1) one CTA per tile, operands staged through registers, 2-stage ring, __syncthreads per
//               k-block.  Kept as the A/B reference (ls_set_gemm_variant(1)).
//   k_gemm_tc2  (default) persistent, warp-specialised: grid = #SMs, every CTA loops over output tiles;
//               warps 16-17 producers: cp.async.bulk (TMA bulk, mbarrier complete_tx) of the 16 KB weight image of
//                          the k-block (warp 17) and ONE cp.async.bulk.tensor.3d per k-block for the raw fp32
//                          activation tile [16 k][128 n] (warp 16, 9 stages of run-ahead);
//               warps 4-15 transform (3 groups of 4, k-blocks round robin): raw [16 k][128 n] tile -> hi/lo TF32 split in the canonical K-major UMMA
//                          image (the 4x4 register transpose of round 1, now smem -> smem);
//               warp 18    one thread issues the tcgen05.mma's; tcgen05.commit frees the operand stage / publishes
//                          the accumulator;
//               warps 0-3  epilogue: tcgen05.ld of accumulator buffer i while the MMAs of tile i+1 fill buffer
//                          i^1 (2 x 128 TMEM columns).
//               Rings: 3 operand stages (32 KB each), 9 raw stages (8 KB each); no __syncthreads in the tile loop.
#include <cuda.h>

#include <atomic>
#include <cstdlib>

#include "ls_common.cuh"

namespace ls {
namespace {

constexpr int TM = 128, TN = 128, TKB = 16;          // tile rows, tile columns, k-block
constexpr int TC_THREADS = 128, TC_STAGES = 2;
constexpr int A_STAGE_FLOATS = TM * TKB;              // per hi or lo image: 2048 floats (8 KB)
constexpr int B_STAGE_FLOATS = TKB * TN;              // 2048 floats (8 KB)
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
    // instead of polling -- in the round-2 profile 40 % of the persistent kernel's issued instructions were the
    // YIELD / TRYWAIT / BRA triplets of un-hinted polling loops, competing with the MMA-issuing thread for issue slots
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}

// smem matrix descriptor (no swizzle): start address, leading / stride byte offsets (>>4), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                 // layout_type (bits 61-63) = 0: SWIZZLE_NONE / interleave
}

// instruction descriptor: D=f32, A=B=tf32, both operands K-major, N=128, M=128
// (MN-major + SWIZZLE_NONE produced all-zero results in profiles/microbench/tc_probe.cu, K-major is verified)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((TN >> 3) << 17) |
                           ((TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_n(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

struct TcShared {
    float a[TC_STAGES][2][A_STAGE_FLOATS];  // [stage][hi/lo]
    float b[TC_STAGES][2][B_STAGE_FLOATS];
    long long col_base[TN];                 // output offset of every tile column (-1 = out of range)
    int col_axis[TN];
    long long col_bias[TN];
    uint64_t bar_empty[TC_STAGES];
    uint64_t bar_done;
    uint32_t tmem_base;
};

template <bool PM>
__global__ void __launch_bounds__(TC_THREADS) k_gemm_tc(const GemmArgs a, const float* __restrict__ wpk, int n_kb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcShared& sh = *reinterpret_cast<TcShared*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int r0 = blockIdx.y * TM;
    const long long c0 = (long long)blockIdx.x * TN;
    const long long ncols = (long long)a.B * a.n_per_b;

    // ---- operand staging through registers, two k-blocks ahead of the tensor core.  The global loads of
    //      k-blocks 0 and 1 are issued before the one-time setup so that their latency overlaps it.
    const float* wtile = wpk + (size_t)blockIdx.y * n_kb * (2 * A_STAGE_FLOATS);
    constexpr int W_F4 = (2 * A_STAGE_FLOATS / 4) / TC_THREADS;  // float4 of the weight image per thread
    const long long jcol = c0 + lane * 4;
    const bool col_ok = jcol < ncols;
    long long xoff = 0;
    if (col_ok) {
        const long long bb = jcol / a.n_per_b;
        xoff = bb * a.x_sb + (jcol - bb * a.n_per_b);
    }
    float4 wr[2][W_F4], xr[2][4];
    auto g_load = [&](int kb, float4* wreg, float4* xreg) {
        // weights: contiguous 16 KB image (hi then lo)
        const float4* src = reinterpret_cast<const float4*>(wtile + (size_t)kb * (2 * A_STAGE_FLOATS));
#pragma unroll
        for (int i = 0; i < W_F4; ++i) wreg[i] = __ldg(src + t + i * TC_THREADS);
        // activations: warp w owns k-core w (4 k rows), lane owns 4 consecutive columns
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int k = kb * TKB + w * 4 + r;
            xreg[r] = (col_ok && k < a.K) ? __ldg(reinterpret_cast<const float4*>(a.X + xoff + (long long)k * a.x_sk))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto s_store = [&](int stage, const float4* wreg, const float4* v) {
        float4* dst = reinterpret_cast<float4*>(&sh.a[stage][0][0]);
#pragma unroll
        for (int i = 0; i < W_F4; ++i) dst[t + i * TC_THREADS] = wreg[i];
        // Four coalesced float4 loads gave a 4(k) x 4(n) block per thread; its transpose is four 16-byte
        // core-matrix rows (one per column) of the canonical K-major image [kcore][ngroup][8 n][4 k].  The four
        // stores are issued in a lane-rotated order so that each quarter-warp hits 8 distinct 16-byte bank groups.
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int i = (s4 + (lane >> 1)) & 3;  // which of the thread's 4 columns goes out in this step
            float4 c;
            c.x = i == 0 ? v[0].x : (i == 1 ? v[0].y : (i == 2 ? v[0].z : v[0].w));
            c.y = i == 0 ? v[1].x : (i == 1 ? v[1].y : (i == 2 ? v[1].z : v[1].w));
            c.z = i == 0 ? v[2].x : (i == 1 ? v[2].y : (i == 2 ? v[2].z : v[2].w));
            c.w = i == 0 ? v[3].x : (i == 1 ? v[3].y : (i == 2 ? v[3].z : v[3].w));
            float4 hi, lo;
            // round to nearest TF32 (magnitude + half ulp, then clear 13 bits): lo = x - hi is exact and
            // signed, so the tensor core's truncation of lo does not accumulate a bias over K
            hi.x = __uint_as_float((__float_as_uint(c.x) + 0x1000u) & 0xffffe000u);
            hi.y = __uint_as_float((__float_as_uint(c.y) + 0x1000u) & 0xffffe000u);
            hi.z = __uint_as_float((__float_as_uint(c.z) + 0x1000u) & 0xffffe000u);
            hi.w = __uint_as_float((__float_as_uint(c.w) + 0x1000u) & 0xffffe000u);
            lo.x = c.x - hi.x;
            lo.y = c.y - hi.y;
            lo.z = c.z - hi.z;
            lo.w = c.w - hi.w;
            const int n = lane * 4 + i;
            const int off = ((w * (TN / 8) + (n >> 3)) * 8 + (n & 7)) * 4;
            *reinterpret_cast<float4*>(&sh.b[stage][0][off]) = hi;
            *reinterpret_cast<float4*>(&sh.b[stage][1][off]) = lo;
        }
    };
    g_load(0, wr[0], xr[0]);
    if (n_kb > 1) g_load(1, wr[1], xr[1]);

    // ---- one-time setup: barriers, TMEM allocation (warp 0), per-column output offsets
    if (t == 0) {
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(&sh.bar_empty[s], 1);
        mbar_init(&sh.bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const long long j = c0 + t;  // TC_THREADS == TN
        long long base = -1, boff = 0;
        int axis = 0;
        if (j < ncols) {
            const long long b = j / a.n_per_b;
            const int n = (int)(j - b * a.n_per_b);
            axis = a.npts > 0 ? n / a.npts : 0;
            if (PM) {
                const int pt = n - axis * a.npts;
                base = (b * a.npts + pt) * ((long long)a.R * 3) + (long long)axis * a.c_out;
            } else {
                base = b * a.o_sb + n;
                boff = b * a.bias_sb + (a.bias_axis ? axis : 0);
            }
        }
        sh.col_base[t] = base;
        sh.col_axis[t] = axis;
        sh.col_bias[t] = boff;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = sh.tmem_base;

    auto k_step = [&](int kb, float4* wreg, float4* xreg) {
        const int stage = kb % TC_STAGES;
        if (kb >= TC_STAGES) mbar_wait(&sh.bar_empty[stage], ((kb / TC_STAGES) - 1) & 1);
        s_store(stage, wreg, xreg);
        if (kb + 2 < n_kb) g_load(kb + 2, wreg, xreg);  // refill this register set: in flight for two k-steps
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(&sh.a[stage][0][0]), a_lo = smem_u32(&sh.a[stage][1][0]);
            const uint32_t b_hi = smem_u32(&sh.b[stage][0][0]), b_lo = smem_u32(&sh.b[stage][1][0]);
#pragma unroll
            for (int s = 0; s < TKB / 8; ++s) {
                // A: K-major, cores of 8 rows x 16 B; m-groups 128 B apart (SBO), k-cores 2048 B apart (LBO)
                const uint64_t dah = make_desc(a_hi + s * 2 * (TM / 8) * 128, (TM / 8) * 128, 128);
                const uint64_t dal = make_desc(a_lo + s * 2 * (TM / 8) * 128, (TM / 8) * 128, 128);
                // B: K-major as well (the loader transposes): n-groups 128 B apart, k-cores 2048 B apart
                const uint64_t dbh = make_desc(b_hi + s * 2 * (TN / 8) * 128, (TN / 8) * 128, 128);
                const uint64_t dbl = make_desc(b_lo + s * 2 * (TN / 8) * 128, (TN / 8) * 128, 128);
                umma_tf32(tmem_d, dal, dbh, (kb | s) != 0);
                umma_tf32(tmem_d, dah, dbl, 1);
                umma_tf32(tmem_d, dah, dbh, 1);
            }
            umma_commit(&sh.bar_empty[stage]);  // frees this smem stage when the MMAs have read it
            if (kb == n_kb - 1) umma_commit(&sh.bar_done);
        }
    };
    for (int kb = 0; kb < n_kb; kb += 2) {
        k_step(kb, wr[0], xr[0]);
        if (kb + 1 < n_kb) k_step(kb + 1, wr[1], xr[1]);
    }

    // ---- epilogue: TMEM -> registers -> global
    mbar_wait(&sh.bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = r0 + w * 32 + lane;  // TMEM lane == tile row
    const bool row_ok = r < a.R;
    long long row_off = 0;
    if (PM) {
        const int part = row_ok ? r / a.c_out : 0;
        row_off = (long long)part * 3 * a.c_out + (r - part * a.c_out);
    }
#pragma unroll 1
    for (int cc = 0; cc < TN; cc += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(w * 32) << 16) + (uint32_t)cc;
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
        if (PM) {
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const long long base = sh.col_base[cc + j];
                    if (base >= 0) a.out[base + row_off] = __uint_as_float(v[j]);  // lanes = consecutive channels
                }
            }
        } else {
            // channel-major: transpose the warp's 32 rows x 32 columns through shared memory (the operand
            // stages are free once bar_done has fired) so that lanes store consecutive columns of one row
            float* tile = &sh.a[0][0][0] + w * (32 * 33);
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]);
            __syncwarp();
            const long long base = sh.col_base[cc + lane];
            const long long bo = sh.col_bias[cc + lane];
            if (base >= 0) {
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = r0 + w * 32 + rr;
                    if (row >= a.R) break;
                    float val = tile[rr * 33 + lane];
                    if (a.bias) val += __ldg(a.bias + bo + (long long)row * a.bias_sr);
                    if (a.relu) val = fmaxf(val, 0.f);
                    if (a.mask && !(__ldg(a.mask + base + (long long)row * a.o_sr) > 0.f)) val = 0.f;
                    a.out[base + (long long)row * a.o_sr] = val;
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}


// =====================================================================================================
// k_gemm_tc2: persistent warp-specialised variant (see the header comment)
// =====================================================================================================
constexpr int G2_S = 3, G2_RS = 9;                 // operand stages, raw activation stages
// The raw -> hi/lo transform of one k-block is a ~600-cycle dependent chain per warp (LDS, split, 8 STS, proxy fence,
// arrive) against 384 cycles of MMA work per k-block: with one transform group the round-2 profile showed the epilogue
// and the MMA warp waiting on it (profiles/r02/ncu_gemm_tc2_v2a.txt).  G2_XF_GROUPS groups of 4 warps take the k-blocks
// round robin, so several k-blocks are transformed concurrently.
// INVARIANT: a ring stage is always served by the same group (G2_XF_GROUPS divides G2_S and G2_RS).  An mbarrier
// parity wait is only meaningful for a waiter that observes every phase of the barrier in order: a group that skipped a
// use of a stage could arrive two phases early, see the parity of the phase before last as "complete" and read a stage
// the TMA unit has not filled yet (measured: wrong tiles at N = 2048 with 3 groups over 4 stages).
constexpr int G2_XF_GROUPS = 3;
static_assert(G2_S % G2_XF_GROUPS == 0 && G2_RS % G2_XF_GROUPS == 0, "a ring stage must always be served by the same transform group");
// Two producer warps: the activation tiles come from HBM (first touch) and want a long run-ahead (9 raw stages =
// ~3 500 cycles of MMA work), the weight images come from L2 and share the 3-stage operand ring with the transform
// output.  With ONE producer loop the raw requests were throttled by the weight ring (3 k-blocks of run-ahead) and the
// round-2 profile showed 27 % of the stall samples on the transform warps' wait for the raw tile.
// Two epilogue groups of 4 warps: group g drains accumulator buffer g (tiles g, g+2, ... of the CTA).  With ONE group the
// round-2 profile showed its 4 warps (one per scheduler, nothing to hide their dependent-issue latency behind) busy 85 %
// of the time at ~11 us per channel-major tile while the tensor pipe idled (7-36 % active, 54 % on the SDF decoder).
constexpr int G2_EPI_GROUPS = 2;
constexpr int G2_XF_WARP0 = 4 * G2_EPI_GROUPS, G2_PRODX_WARP = G2_XF_WARP0 + 4 * G2_XF_GROUPS, G2_PRODA_WARP = G2_PRODX_WARP + 1,
              G2_MMA_WARP = G2_PRODA_WARP + 1;
constexpr int G2_THREADS = 32 * (G2_MMA_WARP + 1);  // warps 0-7 epilogue, 8.. transform groups, 2 producers, MMA
constexpr int G2_TMEM_COLS = 2 * TN;                // two accumulator buffers

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

struct G2Shared {
    float a[G2_S][2][A_STAGE_FLOATS];  // [stage][hi/lo] weight images (bulk copies)
    float b[G2_S][2][B_STAGE_FLOATS];  // [stage][hi/lo] activation images (written by the transform warps)
    float raw[G2_RS][TKB][TN];         // raw fp32 activation rows (bulk copies)
    float epi[4 * G2_EPI_GROUPS][32 * 36];  // per epilogue warp: 32 x 32 transpose tile (row stride 33 for the ragged path, 36 = 16-byte aligned rows for the fast path)
    long long col_base[G2_EPI_GROUPS][TN];  // per epilogue group: output offset of every tile column (-1 = out of range)
    long long col_bias[G2_EPI_GROUPS][TN];
    uint64_t full_a[G2_S], full_b[G2_S], empty[G2_S], raw_full[G2_RS], raw_empty[G2_RS], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

// one TMA tensor request: box {128 columns, 16 k rows, 1 instance} of the activation tensor -> smem [16][128]
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// tmap: activations as a 3-D tensor (n, k, instance) -- see make_act_map.  tile_per_b = column tiles per instance;
// tn = columns per tile (128, or 96 / 64 / 32 when an instance has fewer than 128 columns: UMMA N = tn).
template <bool PM>
__global__ void __launch_bounds__(G2_THREADS, 1) k_gemm_tc2(const GemmArgs a, const float* __restrict__ wpk, int n_kb, int n_mt,
                                                            int n_tiles, const __grid_constant__ CUtensorMap tmap,
                                                            int tile_per_b, int tn) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    G2Shared& sh = *reinterpret_cast<G2Shared*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long ncols = (long long)a.B * a.n_per_b;

    if (t == 0) {
        for (int s = 0; s < G2_S; ++s) {
            mbar_init(&sh.full_a[s], 1);
            mbar_init(&sh.full_b[s], 128);
            mbar_init(&sh.empty[s], 1);
        }
        for (int r = 0; r < G2_RS; ++r) {
            mbar_init(&sh.raw_full[r], 1);
            mbar_init(&sh.raw_empty[r], 128);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.tmem_full[i], 1);
            mbar_init(&sh.tmem_empty[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == G2_MMA_WARP) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "r"(G2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sh.tmem_base;

    if (w == G2_PRODX_WARP) {
        // ================================================================ producer 1: activation tiles (TMA tensor requests)
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int ct = tile / n_mt, tb = ct / tile_per_b, tn0 = (ct - tb * tile_per_b) * tn;
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    // ONE tensor request per k-block (rows k >= K and columns past the end are zero-filled by the TMA
                    // unit; the transaction count is always the full box)
                    const int r = it % G2_RS;
                    if (it >= G2_RS) mbar_wait(&sh.raw_empty[r], ((it / G2_RS) - 1) & 1);
                    mbar_arrive_expect_tx(&sh.raw_full[r], (uint32_t)(TKB * tn * 4));
                    tma_load_3d(&sh.raw[r][0][0], &tmap, tn0, kb * TKB, tb, &sh.raw_full[r]);
                }
            }
        }
    } else if (w == G2_PRODA_WARP) {
        // ================================================================ producer 2: weight images (TMA bulk copies)
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const float* wtile = wpk + (size_t)(tile % n_mt) * n_kb * (2 * A_STAGE_FLOATS);
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    // (m-tile, k-block): one contiguous 16 KB block (hi then lo)
                    const int s = it % G2_S;
                    if (it >= G2_S) mbar_wait(&sh.empty[s], ((it / G2_S) - 1) & 1);
                    mbar_arrive_expect_tx(&sh.full_a[s], 2 * A_STAGE_FLOATS * 4);
                    bulk_g2s(&sh.a[s][0][0], wtile + (size_t)kb * (2 * A_STAGE_FLOATS), 2 * A_STAGE_FLOATS * 4, &sh.full_a[s]);
                }
            }
        }
    } else if (w == G2_MMA_WARP) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            int it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                const int acc = lt & 1;
                if (lt >= 2) {  // the epilogue must have drained this accumulator buffer
                    mbar_wait(&sh.tmem_empty[acc], ((lt >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t d = tmem + (uint32_t)(acc * TN);
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(tn >> 3) << 17) | ((TM >> 4) << 24);
                const uint32_t b_lbo = (uint32_t)(tn / 8) * 128;  // k-core stride of the activation image
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % G2_S;
                    mbar_wait(&sh.full_a[s], (it / G2_S) & 1);
                    mbar_wait(&sh.full_b[s], (it / G2_S) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(&sh.a[s][0][0]), a_lo = smem_u32(&sh.a[s][1][0]);
                    const uint32_t b_hi = smem_u32(&sh.b[s][0][0]), b_lo = smem_u32(&sh.b[s][1][0]);
#pragma unroll
                    for (int ss = 0; ss < TKB / 8; ++ss) {
                        const uint64_t dah = make_desc(a_hi + ss * 2 * (TM / 8) * 128, (TM / 8) * 128, 128);
                        const uint64_t dal = make_desc(a_lo + ss * 2 * (TM / 8) * 128, (TM / 8) * 128, 128);
                        const uint64_t dbh = make_desc(b_hi + ss * 2 * b_lbo, b_lbo, 128);
                        const uint64_t dbl = make_desc(b_lo + ss * 2 * b_lbo, b_lbo, 128);
                        umma_tf32_n(d, dal, dbh, idesc, (kb | ss) != 0);
                        umma_tf32_n(d, dah, dbl, idesc, 1);
                        umma_tf32_n(d, dah, dbh, idesc, 1);
                    }
                    umma_commit(&sh.empty[s]);  // frees the operand stage when the MMAs have read it
                    if (kb == n_kb - 1) umma_commit(&sh.tmem_full[acc]);
                }
            }
        }
    } else if (w >= G2_XF_WARP0) {
        // ================================================================ transform: raw fp32 -> hi/lo K-major images
        const int kc = (w - G2_XF_WARP0) & 3;   // k-core (4 k rows) owned by this warp
        const int grp = (w - G2_XF_WARP0) >> 2;  // transform group: handles the k-blocks with it % G2_XF_GROUPS == grp
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                if (it % G2_XF_GROUPS != grp) continue;
                const int r = it % G2_RS;
                mbar_wait(&sh.raw_full[r], (it / G2_RS) & 1);
                const float* rawp = &sh.raw[r][0][0] + kc * 4 * tn;  // the TMA box is dense [16 k][tn columns] (row stride tn, not TN)
                // Column n = i * 32 + lane: the four k values of a column ARE one 16-byte core-matrix row of the K-major
                // image ([kcore][ngroup][8 n][4 k] = 4 * n floats into the k-core), so a thread reads 4 scalars (lanes =
                // consecutive columns: conflict free) and writes one float4 per image (lanes = consecutive 16-byte rows:
                // conflict free) -- no register transpose.  Out-of-range elements are TMA zero fill.
                float4 c[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = i * 32 + lane;
                    if (n < tn) {
                        c[i].x = rawp[n];
                        c[i].y = rawp[tn + n];
                        c[i].z = rawp[2 * tn + n];
                        c[i].w = rawp[3 * tn + n];
                    }
                }
                const int s = it % G2_S;
                if (it >= G2_S) mbar_wait(&sh.empty[s], ((it / G2_S) - 1) & 1);
                float* bh = &sh.b[s][0][kc * tn * 4];
                float* bl = &sh.b[s][1][kc * tn * 4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = i * 32 + lane;
                    if (n < tn) {
                        float4 hi, lo;
                        hi.x = __uint_as_float((__float_as_uint(c[i].x) + 0x1000u) & 0xffffe000u);
                        hi.y = __uint_as_float((__float_as_uint(c[i].y) + 0x1000u) & 0xffffe000u);
                        hi.z = __uint_as_float((__float_as_uint(c[i].z) + 0x1000u) & 0xffffe000u);
                        hi.w = __uint_as_float((__float_as_uint(c[i].w) + 0x1000u) & 0xffffe000u);
                        lo.x = c[i].x - hi.x;
                        lo.y = c[i].y - hi.y;
                        lo.z = c[i].z - hi.z;
                        lo.w = c[i].w - hi.w;
                        *reinterpret_cast<float4*>(bh + n * 4) = hi;
                        *reinterpret_cast<float4*>(bl + n * 4) = lo;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
                mbar_arrive(&sh.full_b[s]);
                mbar_arrive(&sh.raw_empty[r]);
            }
        }
    } else {
        // ================================================================ epilogue: TMEM -> registers -> global
        // Group eg (warps 4 eg .. 4 eg + 3) owns accumulator buffer eg: the CTA's tiles eg, eg + 2, ...  A warp reads the
        // 32 TMEM lanes (= tile rows) of its quadrant (w & 3).
        const int eg = w >> 2, wq = w & 3, te = t & 127;
        long long* col_base = sh.col_base[eg];
        long long* col_bias = sh.col_bias[eg];
        const long long r3 = (long long)a.R * 3;
        int use = 0;
        for (int tile = blockIdx.x + eg * (int)gridDim.x; tile < n_tiles; tile += G2_EPI_GROUPS * (int)gridDim.x, ++use) {
            const int mt = tile % n_mt;
            const long long c0 = (long long)(tile / n_mt) * tn;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");  // the previous tile's column tables are no longer read
            {
                const long long j = c0 + te;
                long long base = -1, boff = 0;
                if (te < tn && j < ncols) {
                    const long long b = j / a.n_per_b;
                    const int n = (int)(j - b * a.n_per_b);
                    const int axis = a.npts > 0 ? n / a.npts : 0;
                    if (PM) {
                        const int pt = n - axis * a.npts;
                        base = (b * a.npts + pt) * r3 + (long long)axis * a.c_out;
                    } else {
                        base = b * a.o_sb + n;
                        boff = b * a.bias_sb + (a.bias_axis ? axis : 0);
                    }
                }
                col_base[te] = base;
                col_bias[te] = boff;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
            const int r0 = mt * TM;
            const int r = r0 + wq * 32 + lane;  // TMEM lane == tile row
            const bool row_ok = r < a.R;
            long long row_off = 0;
            if (PM) {
                const int part = row_ok ? r / a.c_out : 0;
                row_off = (long long)part * 3 * a.c_out + (r - part * a.c_out);
            }
            mbar_wait(&sh.tmem_full[eg], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < tn; cc += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(eg * TN + cc);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                // the chunk's column geometry (warp uniform) and its bias are fetched while the TMEM load is in flight
                const long long b0 = col_base[cc], b31 = col_base[cc + 31];
                bool fast;
                float bias0 = 0.f;
                if (PM) {
                    // 32 valid columns = 32 consecutive points of one (instance, axis): rows r3 floats apart
                    fast = b0 >= 0 && b31 - b0 == 31 * r3;
                } else {
                    // 32 valid columns, contiguous in the output, 16-byte aligned, one bias offset for the whole chunk
                    const long long bo0 = col_bias[cc];
                    fast = b0 >= 0 && b31 - b0 == 31 && ((b0 | a.o_sr) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                           (a.mask == nullptr || (reinterpret_cast<uintptr_t>(a.mask) & 15) == 0) &&
                           (a.bias == nullptr || bo0 == col_bias[cc + 31]);
                    if (fast && a.bias && row_ok) bias0 = __ldg(a.bias + bo0 + (long long)r * a.bias_sr);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc + 32 >= tn) {  // this thread's last read of the buffer: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&sh.tmem_empty[eg]);
                }
                if (PM) {
                    if (row_ok) {
                        if (fast) {
                            float* o = a.out + b0 + row_off;  // lanes = consecutive channels: 128 B per warp store
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[j * r3] = __uint_as_float(v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const long long base = col_base[cc + j];
                                if (base >= 0) a.out[base + row_off] = __uint_as_float(v[j]);
                            }
                        }
                    }
                } else if (fast) {
                    // bias / ReLU while lane == row, then the warp's 32 rows x 32 columns go through shared memory so that
                    // 8 lanes store one row's 128 contiguous bytes (4 full lines per instruction).  A thread storing its
                    // own row as 8 float4 touches 32 different lines per instruction, half a sector each: the global
                    // conv of layers 2-3 ran at 2.5 TB/s of useful traffic with it.
                    float* tl = &sh.epi[w][0];
                    const bool relu = a.relu != 0;
                    float4* trow = reinterpret_cast<float4*>(tl + lane * 36);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float x0 = __uint_as_float(v[4 * j4]) + bias0, x1 = __uint_as_float(v[4 * j4 + 1]) + bias0;
                        float x2 = __uint_as_float(v[4 * j4 + 2]) + bias0, x3 = __uint_as_float(v[4 * j4 + 3]) + bias0;
                        if (relu) x0 = fmaxf(x0, 0.f), x1 = fmaxf(x1, 0.f), x2 = fmaxf(x2, 0.f), x3 = fmaxf(x3, 0.f);
                        trow[j4] = make_float4(x0, x1, x2, x3);
                    }
                    __syncwarp();
                    const int nrow = min(32, a.R - (r0 + wq * 32));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + (lane >> 3);
                        if (rr < nrow) {
                            const long long oo = b0 + (long long)(r0 + wq * 32 + rr) * a.o_sr + 4 * (lane & 7);
                            float4 x = *reinterpret_cast<const float4*>(tl + rr * 36 + 4 * (lane & 7));
                            if (a.mask) {
                                const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mask + oo));
                                x = make_float4(m4.x > 0.f ? x.x : 0.f, m4.y > 0.f ? x.y : 0.f, m4.z > 0.f ? x.z : 0.f, m4.w > 0.f ? x.w : 0.f);
                            }
                            *reinterpret_cast<float4*>(a.out + oo) = x;
                        }
                    }
                    __syncwarp();
                } else {
                    // ragged chunk (tile edge, instance boundary inside the chunk, unaligned output): bias + ReLU while
                    // lane == row, then transpose the warp's 32 x 32 block so that lanes store consecutive columns
                    float* tl = &sh.epi[w][0];
                    float bcache = 0.f;
                    long long last_bo = -1;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = __uint_as_float(v[j]);
                        if (a.bias) {
                            const long long bo = col_bias[cc + j];
                            if (bo != last_bo) {
                                bcache = row_ok ? __ldg(a.bias + bo + (long long)r * a.bias_sr) : 0.f;
                                last_bo = bo;
                            }
                            x += bcache;
                        }
                        if (a.relu) x = fmaxf(x, 0.f);
                        tl[lane * 33 + j] = x;
                    }
                    __syncwarp();
                    const long long base = col_base[cc + lane];
                    if (base >= 0) {
                        const int nrow = min(32, a.R - (r0 + wq * 32));
                        for (int rr = 0; rr < nrow; ++rr) {
                            const long long oo = base + (long long)(r0 + wq * 32 + rr) * a.o_sr;
                            float x = tl[rr * 33 + lane];
                            if (a.mask && !(__ldg(a.mask + oo) > 0.f)) x = 0.f;
                            a.out[oo] = x;
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == G2_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(G2_TMEM_COLS) : "memory");
    }
}

// W [R][ldw] row-major -> per (m-tile, k-block): hi image then lo image, each [kcore 4][mgroup 16][8 rows][4 k]
__global__ void k_tc_pack_weights(const float* __restrict__ W, int R, int K, int ldw, float* __restrict__ out, int n_kb) {
    const int mt = blockIdx.y, kb = blockIdx.x;
    float* dst = out + ((size_t)mt * n_kb + kb) * (2 * A_STAGE_FLOATS);
    for (int e = threadIdx.x; e < A_STAGE_FLOATS; e += blockDim.x) {
        const int kk = e & 3, row8 = (e >> 2) & 7, mg = (e >> 5) & 15, kc = e >> 9;
        const int r = mt * TM + mg * 8 + row8, k = kb * TKB + kc * 4 + kk;
        const float x = (r < R && k < K) ? W[(size_t)r * ldw + k] : 0.f;
        // round-to-nearest-even to TF32 for the hi part (the tensor core truncates, so hi must be exact TF32)
        uint32_t u = __float_as_uint(x);
        uint32_t h = (u + 0x00000fffu + ((u >> 13) & 1u)) & 0xffffe000u;
        if ((u & 0x7f800000u) == 0x7f800000u) h = u & 0xffffe000u;  // inf / nan: no rounding carry
        const float hi = __uint_as_float(h);
        dst[e] = hi;
        dst[A_STAGE_FLOATS + e] = x - hi;
    }
}

}  // namespace

// 128-row-tile images (k_gemm_tc, k_gemm_tc2, k_gemm_tc3 with NT = 128), followed by the 256-row-tile images of
// k_gemm_tc3 (ls_gemm_tc3.cu) when R > 128
static size_t tc128_floats(int R, int K) {
    const size_t mt = (R + TM - 1) / TM, kb = (K + TKB - 1) / TKB;
    return mt * kb * 2 * A_STAGE_FLOATS;
}
size_t tc_packed_floats(int R, int K) { return tc128_floats(R, K) + tc3_packed_floats(R, K); }

int tc_pack_weights(const float* W, int R, int K, int ldw, float* packed, cudaStream_t st) {
    LS_REQUIRE(W && packed && R > 0 && K > 0 && ldw >= K, "tc_pack_weights: bad arguments");
    const int n_kb = (K + TKB - 1) / TKB;
    dim3 grid(n_kb, (R + TM - 1) / TM);
    k_tc_pack_weights<<<grid, 256, 0, st>>>(W, R, K, ldw, packed, n_kb);
    LS_CHECK_LAUNCH("k_tc_pack_weights");
    return tc3_pack_weights(W, R, K, ldw, packed + tc128_floats(R, K), st);
}

bool gemm_tc_supported(const GemmArgs& a) {
    // float4 activation loads: 4-column groups must not straddle instances and must be 16 B aligned
    if (a.n_per_b % 4 != 0 || a.x_sb % 4 != 0 || a.x_sk % 4 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.X) & 15) != 0) return false;
    return true;
}

// 3: k_gemm_tc3 (activations in tensor memory, TS-form MMAs; default); 2: persistent warp-specialised k_gemm_tc2 (both
// operands in shared memory); 1: round-1 k_gemm_tc (env LS_GEMM_VARIANT for A/B runs)
int g_gemm_variant = [] {
    const char* e = getenv("LS_GEMM_VARIANT");
    const int v = e ? atoi(e) : 3;
    return (v >= 1 && v <= 3) ? v : 3;
}();

static int sm_count() {
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

// The persistent kernel feeds its activations with ONE TMA tensor request per (tile, k-block): the activation tensor
// X[b*x_sb + k*x_sk + n] must be expressible as a tiled tensor map whose 128-column boxes never straddle instances:
// either the columns of consecutive instances are contiguous (x_sb == n_per_b: the SDF decoder's feature-major
// layout -> one flat column axis), or n_per_b is a multiple of 128, or an instance has 32 / 64 / 96 columns (3 N = 96
// in encoder layers 5-6 and the head: one tile per instance, UMMA N = 96).  Other shapes and launches with few tiles
// run the per-tile kernel.
static int gemm_v2_tile_cols(const GemmArgs& a) {
    if (a.x_sb == a.n_per_b || a.n_per_b % TN == 0) return TN;
    if (a.n_per_b < TN && a.n_per_b % 32 == 0) return a.n_per_b;  // 96 (3 x 32 points), 64, 32: one tile per instance, UMMA N = tile
    return 0;
}
static bool gemm_v2_geometry(const GemmArgs& a) { return gemm_v2_tile_cols(a) != 0; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static std::atomic<void*> cached{nullptr};
    void* f = cached.load(std::memory_order_acquire);
    if (f == nullptr) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        cached.store(f, std::memory_order_release);
    }
    return reinterpret_cast<EncodeTiledFn>(f);
}

static int make_act_map(const GemmArgs& a, CUtensorMap* map, int* tile_per_b, int tn) {
    EncodeTiledFn enc = encode_tiled_fn();
    LS_REQUIRE(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    const bool flat = a.x_sb == a.n_per_b;  // one contiguous column axis over all instances
    const cuuint64_t n0 = flat ? (cuuint64_t)a.B * a.n_per_b : (cuuint64_t)a.n_per_b;
    const cuuint64_t nb = flat ? 1 : (cuuint64_t)a.B;
    const cuuint64_t dims[3] = {n0, (cuuint64_t)a.K, nb};
    const cuuint64_t strides[2] = {(cuuint64_t)a.x_sk * sizeof(float), (cuuint64_t)(flat ? n0 : a.x_sb) * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)tn, TKB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.X), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return LS_ERR_CUDA;
    }
    *tile_per_b = (int)((n0 + tn - 1) / tn);
    return LS_OK;
}

int launch_gemm_tc(const GemmArgs& a, const float* packed, cudaStream_t st) {
    LS_REQUIRE(packed != nullptr, "gemm_tc: packed weights missing");
    LS_REQUIRE(gemm_tc_supported(a), "gemm_tc: unsupported activation geometry");
    LS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "gemm_tc: packed weights must be 16-byte aligned");
    if (a.point_major)
        LS_REQUIRE(a.c_out % 32 == 0 && a.R % a.c_out == 0 && a.npts > 0 && a.n_per_b == 3 * a.npts,
                   "gemm_tc: bad point-major geometry");
    // Dispatch by geometry: short contractions (K < 128: the table GEMMs of layers 1-4, the global conv of layers 2-3)
    // are epilogue / store bound and measured 10-30 % faster on k_gemm_tc2; K >= 128 (SDF decoder, deep encoder layers,
    // head) is MMA bound and runs 1.2-1.7x faster with the TS form (profiles/r02/experiments.md).  LS_GEMM_V3_MIN_K: A/B.
    static const int v3_min_k = [] {
        const char* e = getenv("LS_GEMM_V3_MIN_K");
        return e ? atoi(e) : 128;
    }();
    if (g_gemm_variant == 3 && a.K >= v3_min_k)
        return launch_gemm_tc3(a, packed, a.R > 128 ? packed + tc128_floats(a.R, a.K) : nullptr, st);
    const long long ncols = (long long)a.B * a.n_per_b;
    const int n_kb = (a.K + TKB - 1) / TKB;
    // the dynamic shared-memory opt-in is per device: set it on every launch (cheap, legal under stream capture)
    // instead of caching a process-wide flag that a second device in the same process would never see
    const int tn = gemm_v2_tile_cols(a);
    const long long tiles_all = tn ? ((ncols + tn - 1) / tn) * ((a.R + TM - 1) / TM) : 0;
    if (g_gemm_variant == 1 || tn == 0 || tiles_all < 2LL * sm_count()) {  // (variant 3 below its K threshold runs as variant 2)
        dim3 grid((unsigned)((ncols + TN - 1) / TN), (unsigned)((a.R + TM - 1) / TM));
        const size_t smem = sizeof(TcShared) + 128;
        if (a.point_major) {
            LS_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_gemm_tc<true><<<grid, TC_THREADS, smem, st>>>(a, packed, n_kb);
        } else {
            LS_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_gemm_tc<false><<<grid, TC_THREADS, smem, st>>>(a, packed, n_kb);
        }
        LS_CHECK_LAUNCH("k_gemm_tc");
        return LS_OK;
    }
    const long long n_ct = (ncols + tn - 1) / tn;
    const int n_mt = (a.R + TM - 1) / TM;
    LS_REQUIRE(n_ct * n_mt < (1LL << 31), "gemm_tc: too many tiles");
    const int n_tiles = (int)(n_ct * n_mt);
    CUtensorMap tmap;
    int tile_per_b = 0;
    int rc = make_act_map(a, &tmap, &tile_per_b, tn);
    if (rc != LS_OK) return rc;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    const size_t smem = sizeof(G2Shared) + 128;
    if (a.point_major) {
        LS_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gemm_tc2<true><<<grid, G2_THREADS, smem, st>>>(a, packed, n_kb, n_mt, n_tiles, tmap, tile_per_b, tn);
    } else {
        LS_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gemm_tc2<false><<<grid, G2_THREADS, smem, st>>>(a, packed, n_kb, n_mt, n_tiles, tmap, tile_per_b, tn);
    }
    LS_CHECK_LAUNCH("k_gemm_tc2");
    return LS_OK;
}

}  // namespace ls
