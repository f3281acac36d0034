// SDF query: FieldWrapper.forward (model_utils.py:230-263, inner_deepsdf branch) +
// DeepSDF_Decoder.forward (lib_shape_prior/core/lib/implicit_func/deepsdf_decoder.py:78-123).
//
// Activations are kept feature-major [feature][column] with column = (instance, query point) so
// that every layer is one GEMM with weights shared by all instances (ls_gemm.cu).  The z_inv
// columns of layer 0 and of the latent re-injection layer are constant per instance and collapse
// into per-instance bias vectors; the 513-wide decoder input is never
// materialised: only the 257 rows [Z_so3 q ; |q|] are.
#include "ls_common.cuh"

namespace ls {
namespace {

constexpr int SDF_MAX_COLS = 131072;  // columns (instance x point) per pass

// bias[which][b][r] = sum_k Wz[r][k] z_inv[b][k] + bvec[r]   (which = 0: layer 0, 1: latent_in layer)
__global__ void __launch_bounds__(256) k_sdf_bias(const float* __restrict__ wz0, const float* __restrict__ b0,
                                                  const float* __restrict__ wz4, const float* __restrict__ b4,
                                                  const float* __restrict__ z_inv, int hidden, int latent,
                                                  float* __restrict__ bias, int B) {
    extern __shared__ float sz[];  // [latent]
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

// q = (x - t) / s ; U[c][col] = <z_so3[c], q> (c < latent) ; U[latent][col] = |q|   (model_utils.py:236-240)
__global__ void __launch_bounds__(128) k_sdf_prep(const float* __restrict__ query, const float* __restrict__ z_so3,
                                                  const float* __restrict__ s, const float* __restrict__ t,
                                                  int M, int m0, int Mc, int latent, long long ncols,
                                                  float* __restrict__ U) {
    extern __shared__ float sz[];  // [latent][3]
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < latent * 3; k += blockDim.x) sz[k] = z_so3[(size_t)b * latent * 3 + k];
    __syncthreads();
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= Mc) return;
    const float* qp = query + ((size_t)b * M + m0 + pt) * 3;
    const float sc = s[b];
    const float q0 = (qp[0] - t[b * 3 + 0]) / sc, q1 = (qp[1] - t[b * 3 + 1]) / sc, q2 = (qp[2] - t[b * 3 + 2]) / sc;
    float* u = U + (size_t)b * Mc + pt;
    for (int c = 0; c < latent; ++c) u[(size_t)c * ncols] = sz[c * 3] * q0 + sz[c * 3 + 1] * q1 + sz[c * 3 + 2] * q2;
    u[(size_t)latent * ncols] = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);
}

// last linear (hidden -> 1) + tanh (deepsdf_decoder.py:104,120-121)
__global__ void __launch_bounds__(128) k_sdf_out(const float* __restrict__ H, const float* __restrict__ w8, const float* __restrict__ b8,
                                                 int hidden, int M, int m0, int Mc, long long ncols,
                                                 float* __restrict__ sdf) {
    const int b = blockIdx.y;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= Mc) return;
    const float* h = H + (size_t)b * Mc + pt;
    float acc = 0.f;
    for (int k = 0; k < hidden; ++k) acc = fmaf(__ldg(w8 + k), h[(size_t)k * ncols], acc);
    sdf[(size_t)b * M + m0 + pt] = tanhf(acc + __ldg(b8));
}

int check_dec(const ls_decoder_desc* d) {
    LS_REQUIRE(d != nullptr, "null decoder descriptor");
    LS_REQUIRE(d->n_layers == 9 && d->latent_in == 4, "only the shipped 9-layer / latent_in=[4] decoder is supported");
    LS_REQUIRE(d->latent >= 1 && d->latent <= 1024 && d->hidden >= 8, "bad decoder sizes");
    for (int l = 0; l < 9; ++l) LS_REQUIRE(d->w[l] && d->b[l], "missing decoder weights");
    LS_REQUIRE(d->w0_zinv && d->w4_zinv, "missing z_inv column blocks");
    LS_REQUIRE(d->in_dims[0] == d->latent + 1, "layer 0 must take [inner, |q|]");
    LS_REQUIRE(d->in_dims[4] == d->out_dims[3] + d->latent + 1, "latent re-injection layer has the wrong width");
    LS_REQUIRE(d->out_dims[8] == 1, "last layer must have one output");
    return LS_OK;
}

inline int chunk_points(int B, int M) {
    int c = SDF_MAX_COLS / B;
    if (c < 1) c = 1;
    return c < M ? c : M;
}

}  // namespace
}  // namespace ls

using namespace ls;

extern "C" {

int ls_sdf_workspace_bytes(const ls_decoder_desc* d, int32_t B, int32_t M, size_t* bytes) {
    int rc = check_dec(d);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(bytes && B >= 1 && M >= 1, "bad arguments");
    const size_t cols = (size_t)B * chunk_points(B, M);
    const size_t hu_rows = (size_t)d->out_dims[3] + d->latent + 1;
    *bytes = (2 * (size_t)d->hidden + hu_rows) * cols * sizeof(float) + 2 * (size_t)B * d->hidden * sizeof(float) + 1024;
    return LS_OK;
}

int ls_sdf_decode(const ls_decoder_desc* d, const float* query, const float* z_so3, const float* z_inv,
                  const float* s, const float* t, int32_t B, int32_t M, float* sdf, void* workspace,
                  size_t workspace_bytes, void* stream) {
    int rc = check_dec(d);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(query && z_so3 && z_inv && s && t && sdf && workspace, "null pointer");
    LS_REQUIRE(B >= 1 && B <= 65535 && M >= 1, "bad sizes");
    size_t need;
    rc = ls_sdf_workspace_bytes(d, B, M, &need);
    if (rc != LS_OK) return rc;
    if (need > workspace_bytes) {
        set_error("sdf workspace too small");
        return LS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = d->hidden, L = d->latent, h3 = d->out_dims[3];
    const int Mc_max = chunk_points(B, M);
    const size_t cols_max = (size_t)B * Mc_max;
    float* bufA = static_cast<float*>(workspace);
    float* bufB = bufA + (size_t)H * cols_max;
    float* HU = bufB + (size_t)H * cols_max;                       // rows [0,h3): h3 ; rows [h3, h3+L+1): U
    float* bias = HU + (size_t)(h3 + L + 1) * cols_max;             // [2][B][H]

    k_sdf_bias<<<dim3(B, 2), 256, (size_t)L * sizeof(float), st>>>(d->w0_zinv, d->b[0], d->w4_zinv, d->b[4], z_inv, H, L,
                                                                    bias, B);
    LS_CHECK_LAUNCH("k_sdf_bias");

    for (int m0 = 0; m0 < M; m0 += Mc_max) {
        const int Mc = (M - m0) < Mc_max ? (M - m0) : Mc_max;
        const long long ncols = (long long)B * Mc;
        float* U = HU + (size_t)h3 * ncols;  // feature-major with the CURRENT pass's column count
        k_sdf_prep<<<dim3((Mc + 127) / 128, B), 128, (size_t)L * 3 * sizeof(float), st>>>(query, z_so3, s, t, M, m0, Mc, L,
                                                                                          ncols, U);
        LS_CHECK_LAUNCH("k_sdf_prep");
        auto layer = [&](int l, const float* X, int K, float* out, int R, const float* bvec, long long bias_sb) {
            GemmArgs g{};
            g.W = d->w[l];
            g.Wtc = d->w_tc[l];
            g.R = R;
            g.K = K;
            g.ldw = (K + 7) & ~7;
            g.B = B;
            g.n_per_b = Mc;
            g.X = X;
            g.x_sb = Mc;
            g.x_sk = ncols;
            g.out = out;
            g.o_sb = Mc;
            g.o_sr = ncols;
            g.bias = bvec;
            g.bias_sb = bias_sb;
            g.bias_sr = 1;
            g.relu = 1;
            return launch_gemm(g, st);
        };
        if ((rc = layer(0, U, L + 1, bufA, H, bias, H)) != LS_OK) return rc;
        if ((rc = layer(1, bufA, H, bufB, H, d->b[1], 0)) != LS_OK) return rc;
        if ((rc = layer(2, bufB, H, bufA, H, d->b[2], 0)) != LS_OK) return rc;
        if ((rc = layer(3, bufA, H, HU, h3, d->b[3], 0)) != LS_OK) return rc;
        if ((rc = layer(4, HU, h3 + L + 1, bufA, H, bias + (size_t)B * H, H)) != LS_OK) return rc;
        if ((rc = layer(5, bufA, H, bufB, H, d->b[5], 0)) != LS_OK) return rc;
        if ((rc = layer(6, bufB, H, bufA, H, d->b[6], 0)) != LS_OK) return rc;
        if ((rc = layer(7, bufA, H, bufB, H, d->b[7], 0)) != LS_OK) return rc;
        k_sdf_out<<<dim3((Mc + 127) / 128, B), 128, 0, st>>>(bufB, d->w[8], d->b[8], H, M, m0, Mc, ncols, sdf);
        LS_CHECK_LAUNCH("k_sdf_out");
    }
    return LS_OK;
}

}  // extern "C"
