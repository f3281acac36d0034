// Backward pass of the SDF query (FieldWrapper.forward + DeepSDF_Decoder.forward, model_utils.py:230-263,
// deepsdf_decoder.py:78-123): gradients of the decoder output with respect to the query points and the code
// (z_so3, z_inv, s, t).  The reference gets them from autograd in its optimisation loops
// (more_solver.py:118-179 `optim=True` registration, :191-228 `_optimize_code`); here they are explicit kernels:
//
//   forward recompute, every layer kept:  U -> H0 -> H1 -> H2 -> [H3 ; U] -> H4 -> ... -> H7 -> sdf
//   g    = grad_sdf * (1 - sdf^2)                                   (tanh')
//   dY7  = (H7 > 0) * w8 (x) g                                      k_sdf_out_bwd
//   dY_{l-1} = (H_{l-1} > 0) * W_l^T dY_l                           GEMMs on the transposed weights, ReLU mask in the epilogue
//   layer 4:  [dH3 ; dU4] = W4^T dY4,  dz_inv += W4z^T sum_cols dY4
//   layer 0:  dU0 = W0u^T dY0,         dz_inv += W0z^T sum_cols dY0
//   U = [Z_so3 q ; |q|], q = (x - t) / s  ->  grad_query, grad_z_so3, grad_s, grad_t     k_sdf_prep_bwd, k_sdf_dzso3
//
// Same GEMM kernels (tcgen05 3xTF32 / FP32 SIMT) as the forward; one pass over at most SDF_BWD_MAX_COLS columns
// (the Python wrapper chunks longer query lists and sums the code gradients).
#include "ls_common.cuh"

namespace ls {
namespace {

constexpr int SDF_BWD_MAX_COLS = 131072;

__global__ void __launch_bounds__(256) k_bsdf_bias(const float* __restrict__ wz0, const float* __restrict__ b0,
                                                   const float* __restrict__ wz4, const float* __restrict__ b4,
                                                   const float* __restrict__ z_inv, int hidden, int latent,
                                                   float* __restrict__ bias, int B) {
    extern __shared__ float sz[];
    const int b = blockIdx.x, which = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < latent; k += blockDim.x) sz[k] = z_inv[(size_t)b * latent + k];
    __syncthreads();
    const float* W = which == 0 ? wz0 : wz4;
    const float* bv = which == 0 ? b0 : b4;
    float* out = bias + ((size_t)which * B + b) * hidden;
    for (int r = w; r < hidden; r += 8) {
        float s = 0.f;
        for (int k = lane; k < latent; k += 32) s = fmaf(__ldg(W + (size_t)r * latent + k), sz[k], s);
        s = warp_sum(s);
        if (lane == 0) out[r] = s + bv[r];
    }
}

// U rows + the canonical query q (3 rows) and |q| for the backward of the input map
__global__ void __launch_bounds__(128) k_bsdf_prep(const float* __restrict__ query, const float* __restrict__ z_so3,
                                                   const float* __restrict__ s, const float* __restrict__ t, int M,
                                                   int latent, long long ncols, float* __restrict__ U,
                                                   float* __restrict__ Q) {
    extern __shared__ float sz[];
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < latent * 3; k += blockDim.x) sz[k] = z_so3[(size_t)b * latent * 3 + k];
    __syncthreads();
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= M) return;
    const float* qp = query + ((size_t)b * M + pt) * 3;
    const float sc = s[b];
    const float q0 = (qp[0] - t[b * 3 + 0]) / sc, q1 = (qp[1] - t[b * 3 + 1]) / sc, q2 = (qp[2] - t[b * 3 + 2]) / sc;
    const size_t col = (size_t)b * M + pt;
    for (int c = 0; c < latent; ++c) U[(size_t)c * ncols + col] = sz[c * 3] * q0 + sz[c * 3 + 1] * q1 + sz[c * 3 + 2] * q2;
    const float len = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);
    U[(size_t)latent * ncols + col] = len;
    Q[col] = q0;
    Q[ncols + col] = q1;
    Q[2 * ncols + col] = q2;
    Q[3 * ncols + col] = len;
}

// sdf = tanh(w8 . H7 + b8);  dY7[k][col] = (H7[k][col] > 0) * w8[k] * grad_sdf[col] * (1 - sdf^2)
__global__ void __launch_bounds__(128) k_bsdf_out(const float* __restrict__ H7, const float* __restrict__ w8,
                                                  const float* __restrict__ b8, const float* __restrict__ grad_sdf,
                                                  int hidden, long long ncols, float* __restrict__ dY7) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    float acc = 0.f;
    for (int k = 0; k < hidden; ++k) acc = fmaf(__ldg(w8 + k), H7[(size_t)k * ncols + col], acc);
    const float y = tanhf(acc + __ldg(b8));
    const float g = grad_sdf[col] * (1.f - y * y);
    for (int k = 0; k < hidden; ++k)
        dY7[(size_t)k * ncols + col] = H7[(size_t)k * ncols + col] > 0.f ? __ldg(w8 + k) * g : 0.f;
}

// dbias[b][r] = sum over the instance's columns of dY[r][col]
__global__ void __launch_bounds__(256) k_bsdf_rowsum(const float* __restrict__ dY, int rows, int M, long long ncols,
                                                     float* __restrict__ dbias) {
    const int b = blockIdx.y, r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* p = dY + (size_t)r * ncols + (size_t)b * M;
    float s = 0.f;
    for (int i = lane; i < M; i += 32) s += p[i];
    s = warp_sum(s);
    if (lane == 0) dbias[(size_t)b * rows + r] = s;
}

// dz_inv[b][k] (+)= sum_r Wz[r][k] * dbias[b][r]
__global__ void __launch_bounds__(256) k_bsdf_dzinv(const float* __restrict__ Wz, const float* __restrict__ dbias, int hidden,
                                                    int latent, int accumulate, float* __restrict__ dz) {
    extern __shared__ float sd[];  // [hidden]
    const int b = blockIdx.x;
    for (int r = threadIdx.x; r < hidden; r += blockDim.x) sd[r] = dbias[(size_t)b * hidden + r];
    __syncthreads();
    for (int k = threadIdx.x; k < latent; k += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < hidden; ++r) s = fmaf(__ldg(Wz + (size_t)r * latent + k), sd[r], s);
        float* o = dz + (size_t)b * latent + k;
        *o = accumulate ? *o + s : s;
    }
}

// dq = sum_c dU[c] z_so3[c] + dU[L] q / |q| ;  grad_query = dq / s ;  grad_t -= sum dq / s ;  grad_s -= sum <dq, q> / s
__global__ void __launch_bounds__(128) k_bsdf_prep_bwd(const float* __restrict__ dU4, const float* __restrict__ dU0,
                                                       const float* __restrict__ Q, const float* __restrict__ z_so3,
                                                       const float* __restrict__ s, int M, int latent, long long ncols,
                                                       float* __restrict__ grad_query, float* __restrict__ grad_s,
                                                       float* __restrict__ grad_t) {
    extern __shared__ float sz[];  // [latent][3]
    __shared__ float red[4][4];
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < latent * 3; k += blockDim.x) sz[k] = z_so3[(size_t)b * latent * 3 + k];
    __syncthreads();
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, dsv = 0.f;
    const float sc = s[b];
    if (pt < M) {
        const size_t col = (size_t)b * M + pt;
        for (int c = 0; c < latent; ++c) {
            const float g = dU4[(size_t)c * ncols + col] + dU0[(size_t)c * ncols + col];
            d0 = fmaf(g, sz[c * 3], d0);
            d1 = fmaf(g, sz[c * 3 + 1], d1);
            d2 = fmaf(g, sz[c * 3 + 2], d2);
        }
        const float q0 = Q[col], q1 = Q[ncols + col], q2 = Q[2 * ncols + col], len = Q[3 * ncols + col];
        const float gl = dU4[(size_t)latent * ncols + col] + dU0[(size_t)latent * ncols + col];
        if (len > 0.f) {
            const float f = gl / len;
            d0 = fmaf(f, q0, d0);
            d1 = fmaf(f, q1, d1);
            d2 = fmaf(f, q2, d2);
        }
        if (grad_query) {
            float* gq = grad_query + col * 3;
            gq[0] = d0 / sc;
            gq[1] = d1 / sc;
            gq[2] = d2 / sc;
        }
        dsv = -(d0 * q0 + d1 * q1 + d2 * q2) / sc;
        d0 = -d0 / sc;
        d1 = -d1 / sc;
        d2 = -d2 / sc;
    }
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    dsv = warp_sum(dsv);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[w][0] = d0;
        red[w][1] = d1;
        red[w][2] = d2;
        red[w][3] = dsv;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const float v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
        if (threadIdx.x < 3) {
            if (grad_t) atomicAdd(grad_t + b * 3 + threadIdx.x, v);
        } else if (grad_s) {
            atomicAdd(grad_s + b, v);
        }
    }
}

// grad_z_so3[b][c][a] (+)= sum_cols (dU4 + dU0)[c][col] * q_a[col]
__global__ void __launch_bounds__(256) k_bsdf_dzso3(const float* __restrict__ dU4, const float* __restrict__ dU0,
                                                    const float* __restrict__ Q, int M, int latent, long long ncols,
                                                    int accumulate, float* __restrict__ grad_z_so3) {
    const int b = blockIdx.y, c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= latent) return;
    const size_t base = (size_t)b * M;
    const float* g4 = dU4 + (size_t)c * ncols + base;
    const float* g0 = dU0 + (size_t)c * ncols + base;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = lane; i < M; i += 32) {
        const float g = g4[i] + g0[i];
        a0 = fmaf(g, Q[base + i], a0);
        a1 = fmaf(g, Q[ncols + base + i], a1);
        a2 = fmaf(g, Q[2 * ncols + base + i], a2);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
        float* o = grad_z_so3 + ((size_t)b * latent + c) * 3;
        o[0] = accumulate ? o[0] + a0 : a0;
        o[1] = accumulate ? o[1] + a1 : a1;
        o[2] = accumulate ? o[2] + a2 : a2;
    }
}

int check_bwd_desc(const ls_decoder_desc* d) {
    LS_REQUIRE(d != nullptr, "null decoder descriptor");
    LS_REQUIRE(d->n_layers == 9 && d->latent_in == 4, "only the shipped 9-layer / latent_in=[4] decoder is supported");
    for (int l = 0; l < 9; ++l) LS_REQUIRE(d->w[l] && d->b[l], "missing decoder weights");
    LS_REQUIRE(d->w0_zinv && d->w4_zinv, "missing z_inv column blocks");
    for (int l = 1; l < 8; ++l) LS_REQUIRE(l == 4 || d->wt[l], "missing transposed weights (backward)");
    LS_REQUIRE(d->wt[0] && d->wt4_h && d->wt4_u, "missing transposed weights of layers 0 / 4 (backward)");
    return LS_OK;
}

struct BwdPlan {
    float *H[8], *HU, *U, *Q, *dA, *dB, *dU4, *dU0, *bias, *dbias;
    size_t bytes;
};

void bwd_plan(const ls_decoder_desc* d, int B, long long ncols, void* ws, BwdPlan& p) {
    char* base = static_cast<char*>(ws);
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        off = (off + 255) & ~size_t(255);
        float* r = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off += n_floats * sizeof(float);
        return r;
    };
    const size_t H = d->hidden, L = d->latent, h3 = d->out_dims[3], c = (size_t)ncols;
    for (int l = 0; l < 8; ++l) p.H[l] = (l == 3) ? nullptr : take(H * c);
    p.HU = take((h3 + L + 1) * c);  // rows [0,h3): H3 ; rows [h3, h3+L+1): U
    p.H[3] = p.HU;
    p.U = p.HU ? p.HU + h3 * c : nullptr;
    p.Q = take(4 * c);
    p.dA = take(H * c);
    p.dB = take(H * c);
    p.dU4 = take((L + 1) * c);
    p.dU0 = take((L + 1) * c);
    p.bias = take(2 * (size_t)B * H);
    p.dbias = take((size_t)B * H);
    p.bytes = (off + 255) & ~size_t(255);
}

}  // namespace
}  // namespace ls

using namespace ls;

extern "C" {

int ls_sdf_backward_workspace_bytes(const ls_decoder_desc* d, int32_t B, int32_t M, size_t* bytes) {
    int rc = check_bwd_desc(d);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(bytes && B >= 1 && M >= 1, "bad arguments");
    LS_REQUIRE((long long)B * M <= SDF_BWD_MAX_COLS, "sdf backward: at most 131072 query points per call (chunk the list)");
    BwdPlan p;
    bwd_plan(d, B, (long long)B * M, nullptr, p);
    *bytes = p.bytes;
    return LS_OK;
}

int ls_sdf_backward(const ls_decoder_desc* d, const float* query, const float* z_so3, const float* z_inv, const float* s,
                    const float* t, int32_t B, int32_t M, const float* grad_sdf, float* grad_query, float* grad_z_so3,
                    float* grad_z_inv, float* grad_s, float* grad_t, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_bwd_desc(d);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(query && z_so3 && z_inv && s && t && grad_sdf && workspace, "null pointer");
    LS_REQUIRE(B >= 1 && B <= 65535 && M >= 1 && (long long)B * M <= SDF_BWD_MAX_COLS, "bad sizes (B * M <= 131072)");
    const long long ncols = (long long)B * M;
    BwdPlan p;
    bwd_plan(d, B, ncols, workspace, p);
    if (p.bytes > workspace_bytes) {
        set_error("sdf backward workspace too small");
        return LS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = d->hidden, L = d->latent, h3 = d->out_dims[3];

    // ---- forward, every activation kept
    k_bsdf_bias<<<dim3(B, 2), 256, (size_t)L * sizeof(float), st>>>(d->w0_zinv, d->b[0], d->w4_zinv, d->b[4], z_inv, H, L, p.bias, B);
    LS_CHECK_LAUNCH("k_bsdf_bias");
    k_bsdf_prep<<<dim3((M + 127) / 128, B), 128, (size_t)L * 3 * sizeof(float), st>>>(query, z_so3, s, t, M, L, ncols, p.U, p.Q);
    LS_CHECK_LAUNCH("k_bsdf_prep");
    auto gemm = [&](const float* W, const float* Wtc, int R, int K, const float* X, float* out, const float* bvec,
                    long long bias_sb, int relu, const float* mask) {
        GemmArgs g{};
        g.W = W;
        g.Wtc = Wtc;
        g.R = R;
        g.K = K;
        g.ldw = (K + 7) & ~7;
        g.B = B;
        g.n_per_b = M;
        g.X = X;
        g.x_sb = M;
        g.x_sk = ncols;
        g.out = out;
        g.o_sb = M;
        g.o_sr = ncols;
        g.bias = bvec;
        g.bias_sb = bias_sb;
        g.bias_sr = 1;
        g.relu = relu;
        g.mask = mask;
        return launch_gemm(g, st);
    };
    if ((rc = gemm(d->w[0], d->w_tc[0], H, L + 1, p.U, p.H[0], p.bias, H, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[1], d->w_tc[1], H, H, p.H[0], p.H[1], d->b[1], 0, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[2], d->w_tc[2], H, H, p.H[1], p.H[2], d->b[2], 0, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[3], d->w_tc[3], h3, H, p.H[2], p.HU, d->b[3], 0, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[4], d->w_tc[4], H, h3 + L + 1, p.HU, p.H[4], p.bias + (size_t)B * H, H, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[5], d->w_tc[5], H, H, p.H[4], p.H[5], d->b[5], 0, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[6], d->w_tc[6], H, H, p.H[5], p.H[6], d->b[6], 0, 1, nullptr)) != LS_OK) return rc;
    if ((rc = gemm(d->w[7], d->w_tc[7], H, H, p.H[6], p.H[7], d->b[7], 0, 1, nullptr)) != LS_OK) return rc;

    // ---- backward
    k_bsdf_out<<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(p.H[7], d->w[8], d->b[8], grad_sdf, H, ncols, p.dA);  // dY7
    LS_CHECK_LAUNCH("k_bsdf_out");
    if ((rc = gemm(d->wt[7], d->wt_tc[7], H, H, p.dA, p.dB, nullptr, 0, 0, p.H[6])) != LS_OK) return rc;  // dY6
    if ((rc = gemm(d->wt[6], d->wt_tc[6], H, H, p.dB, p.dA, nullptr, 0, 0, p.H[5])) != LS_OK) return rc;  // dY5
    if ((rc = gemm(d->wt[5], d->wt_tc[5], H, H, p.dA, p.dB, nullptr, 0, 0, p.H[4])) != LS_OK) return rc;  // dY4
    if (grad_z_inv) {
        k_bsdf_rowsum<<<dim3((H + 7) / 8, B), 256, 0, st>>>(p.dB, H, M, ncols, p.dbias);
        LS_CHECK_LAUNCH("k_bsdf_rowsum");
        k_bsdf_dzinv<<<B, 256, (size_t)H * sizeof(float), st>>>(d->w4_zinv, p.dbias, H, L, 0, grad_z_inv);
        LS_CHECK_LAUNCH("k_bsdf_dzinv");
    }
    if ((rc = gemm(d->wt4_u, d->wt4_u_tc, L + 1, H, p.dB, p.dU4, nullptr, 0, 0, nullptr)) != LS_OK) return rc;  // dU (layer 4)
    if ((rc = gemm(d->wt4_h, d->wt4_h_tc, h3, H, p.dB, p.dA, nullptr, 0, 0, p.HU)) != LS_OK) return rc;         // dY3 [h3 rows]
    if ((rc = gemm(d->wt[3], d->wt_tc[3], H, h3, p.dA, p.dB, nullptr, 0, 0, p.H[2])) != LS_OK) return rc;        // dY2
    if ((rc = gemm(d->wt[2], d->wt_tc[2], H, H, p.dB, p.dA, nullptr, 0, 0, p.H[1])) != LS_OK) return rc;         // dY1
    if ((rc = gemm(d->wt[1], d->wt_tc[1], H, H, p.dA, p.dB, nullptr, 0, 0, p.H[0])) != LS_OK) return rc;         // dY0
    if (grad_z_inv) {
        k_bsdf_rowsum<<<dim3((H + 7) / 8, B), 256, 0, st>>>(p.dB, H, M, ncols, p.dbias);
        LS_CHECK_LAUNCH("k_bsdf_rowsum");
        k_bsdf_dzinv<<<B, 256, (size_t)H * sizeof(float), st>>>(d->w0_zinv, p.dbias, H, L, 1, grad_z_inv);
        LS_CHECK_LAUNCH("k_bsdf_dzinv");
    }
    if ((rc = gemm(d->wt[0], d->wt_tc[0], L + 1, H, p.dB, p.dU0, nullptr, 0, 0, nullptr)) != LS_OK) return rc;  // dU (layer 0)
    if (grad_t) LS_CHECK_CUDA(cudaMemsetAsync(grad_t, 0, sizeof(float) * 3 * B, st));
    if (grad_s) LS_CHECK_CUDA(cudaMemsetAsync(grad_s, 0, sizeof(float) * B, st));
    k_bsdf_prep_bwd<<<dim3((M + 127) / 128, B), 128, (size_t)L * 3 * sizeof(float), st>>>(p.dU4, p.dU0, p.Q, z_so3, s, M, L, ncols,
                                                                                          grad_query, grad_s, grad_t);
    LS_CHECK_LAUNCH("k_bsdf_prep_bwd");
    if (grad_z_so3) {
        k_bsdf_dzso3<<<dim3((L + 7) / 8, B), 256, 0, st>>>(p.dU4, p.dU0, p.Q, M, L, ncols, 0, grad_z_so3);
        LS_CHECK_LAUNCH("k_bsdf_dzso3");
    }
    return LS_OK;
}

}  // extern "C"
