// Batched FP32 SIMT GEMM with weights shared across instances:
//     C[r, (b,n)] = sum_k W[r,k] * X[b,k,n]        (VecLinear.forward, vec_layers.py:121-134;
//                                                    nn.Linear of deepsdf_decoder.py:104)
// The column index runs over ALL instances (b,n) flattened, so that deep encoder layers with only
// 3*N/32 = 96 columns per instance still fill 128-wide tiles.  FP32 FFMA keeps the fp32-exact
// arithmetic the kNN-graph parity needs; a tcgen05 3xTF32 variant is the
// planned replacement for the large contractions.
//
// Two store modes:
//   channel-major  out[b][r][n]                      (features / raw VecLNA pre-activations)
//   point-major    out[b][pt][(part*3+axis)*Co + c]  (gather tables read by the fused kNN+EdgeConv
//                                                     kernel: one contiguous row per point)
#include "ls_common.cuh"

namespace ls {

constexpr int BM = 128, BN = 128, BK = 8, GEMM_THREADS = 256, PADW = 4;

template <bool PM>
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_gemm(const GemmArgs a) {
    __shared__ __align__(16) float As[2][BK][BM + PADW];
    __shared__ __align__(16) float Xs[2][BK][BN + PADW];

    const int t = threadIdx.x;
    const int r0 = blockIdx.y * BM;
    const long long c0 = (long long)blockIdx.x * BN;
    const long long ncols = (long long)a.B * a.n_per_b;

    // ---- global -> register staging assignments
    const int a_row = t >> 1, a_kq = (t & 1) * 4;
    const bool a_ok = (r0 + a_row) < a.R;
    const float* a_ptr = a.W + (size_t)(r0 + a_row) * a.ldw + a_kq;

    const int x_col = t & (BN - 1), x_k0 = t >> 7;
    const long long xj = c0 + x_col;
    const bool x_ok = xj < ncols;
    long long x_off = 0;
    if (x_ok) {
        long long xb = xj / a.n_per_b;
        x_off = xb * a.x_sb + (xj - xb * a.n_per_b);
    }

    float4 a_reg;
    float x_reg[4];
    auto g_load = [&](int k0) {
        a_reg = a_ok ? *reinterpret_cast<const float4*>(a_ptr + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int k = k0 + x_k0 + 2 * i;
            x_reg[i] = (x_ok && k < a.K) ? __ldg(a.X + x_off + (long long)k * a.x_sk) : 0.f;
        }
    };
    auto s_store = [&](int buf) {
        As[buf][a_kq + 0][a_row] = a_reg.x;
        As[buf][a_kq + 1][a_row] = a_reg.y;
        As[buf][a_kq + 2][a_row] = a_reg.z;
        As[buf][a_kq + 3][a_row] = a_reg.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) Xs[buf][x_k0 + 2 * i][x_col] = x_reg[i];
    };

    const int tx = t & 15, ty = t >> 4;
    const int ra = PM ? tx : ty;  // row group of this thread
    const int cb = PM ? ty : tx;  // column group

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (a.K + BK - 1) / BK;
    g_load(0);
    s_store(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) g_load((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ra * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ra * 4]);
            float4 x0 = *reinterpret_cast<const float4*>(&Xs[buf][k][cb * 4]);
            float4 x1 = *reinterpret_cast<const float4*>(&Xs[buf][k][64 + cb * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            s_store(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = jh * 4 + jj;
            const long long col = c0 + jh * 64 + cb * 4 + jj;
            if (col >= ncols) continue;
            const long long b = col / a.n_per_b;
            const int n = (int)(col - b * a.n_per_b);
            const int axis = a.npts > 0 ? n / a.npts : 0;
            if (PM) {
                const int pt = n - axis * a.npts;
                float* orow = a.out + ((size_t)(b * a.npts + pt)) * ((size_t)a.R * 3);
#pragma unroll
                for (int ih = 0; ih < 2; ++ih) {
                    const int r = r0 + ih * 64 + ra * 4;
                    if (r >= a.R) continue;  // R % 4 == 0 enforced by the launcher
                    const int part = r / a.c_out, c = r - part * a.c_out;
                    float4 v = make_float4(acc[ih * 4 + 0][j], acc[ih * 4 + 1][j], acc[ih * 4 + 2][j],
                                           acc[ih * 4 + 3][j]);
                    *reinterpret_cast<float4*>(orow + (size_t)(part * 3 + axis) * a.c_out + c) = v;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = r0 + (i >> 2) * 64 + ra * 4 + (i & 3);
                    if (r >= a.R) continue;
                    float v = acc[i][j];
                    if (a.bias) v += __ldg(a.bias + b * a.bias_sb + (long long)r * a.bias_sr + (a.bias_axis ? axis : 0));
                    if (a.relu) v = fmaxf(v, 0.f);
                    const long long oo = b * a.o_sb + (long long)r * a.o_sr + n;
                    if (a.mask && !(__ldg(a.mask + oo) > 0.f)) v = 0.f;
                    a.out[oo] = v;
                }
            }
        }
    }
}

int launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
    LS_REQUIRE(a.ldw % 8 == 0 && a.ldw >= a.K, "gemm: ldw must be a multiple of 8 and >= K");
    LS_REQUIRE(a.R > 0 && a.K > 0 && a.B > 0 && a.n_per_b > 0, "gemm: empty problem");
    LS_REQUIRE((reinterpret_cast<uintptr_t>(a.W) & 15) == 0, "gemm: W must be 16-byte aligned");
    if (a.point_major) {
        LS_REQUIRE(a.c_out % 4 == 0 && a.R % a.c_out == 0 && a.npts > 0 && a.n_per_b == 3 * a.npts,
                   "gemm: bad point-major geometry");
        LS_REQUIRE((reinterpret_cast<uintptr_t>(a.out) & 15) == 0, "gemm: out must be 16-byte aligned");
    }
    const long long ncols = (long long)a.B * a.n_per_b;
    dim3 grid((unsigned)((ncols + BN - 1) / BN), (unsigned)((a.R + BM - 1) / BM));
    if (a.point_major)
        k_gemm<true><<<grid, GEMM_THREADS, 0, st>>>(a);
    else
        k_gemm<false><<<grid, GEMM_THREADS, 0, st>>>(a);
    LS_CHECK_LAUNCH("k_gemm");
    return LS_OK;
}

bool g_use_tensor_cores = true;

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
    if (g_use_tensor_cores && a.Wtc != nullptr && gemm_tc_supported(a)) return launch_gemm_tc(a, a.Wtc, st);
    return launch_gemm_simt(a, st);
}

}  // namespace ls
