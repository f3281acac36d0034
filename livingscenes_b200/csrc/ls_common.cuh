// Shared device helpers and host-side error plumbing for the livingscenes_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/livingscenes_b200.h"

namespace ls {

void set_error(const std::string& msg);
void count_launch();

#define LS_CHECK_CUDA(expr)                                                                 \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ls::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            return LS_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define LS_CHECK_LAUNCH(name)                                                               \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            ls::set_error(std::string("launch of ") + name + ": " + cudaGetErrorString(_e)); \
            return LS_ERR_CUDA;                                                             \
        }                                                                                   \
        ls::count_launch();                                                                 \
    } while (0)

#define LS_REQUIRE(cond, msg)                                                               \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ls::set_error(std::string("invalid argument: ") + msg);                         \
            return LS_ERR_INVALID;                                                          \
        }                                                                                   \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;
constexpr float EPS_NRM = 1e-12f;   // F.normalize eps
constexpr float EPS_NRM2 = 1e-24f;  // its square, for the sqrt-free VN activation

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
// reductions inside each aligned group of 16 lanes
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// Packed FP32 pairs (Blackwell FADD2 / FFMA2, PTX add/fma .f32x2): two IEEE-rounded fp32 operations per
// issued instruction; results are bit-identical to the scalar form.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// VN leaky-ReLU on one 3-vector (vec_layers.py:241-268, so3 mode) in sqrt-free form:
//   out = q + (lrelu(p) - p) * khat,  p = <q, khat>,  khat = k / max(|k|, 1e-12)
//       = q - (1 - slope) * min(<q,k>, 0) / max(|k|^2, 1e-24) * k
__device__ __forceinline__ void vn_act(float q0, float q1, float q2, float k0, float k1, float k2,
                                       float one_minus_slope, float& o0, float& o1, float& o2) {
    float n2 = fmaf(k2, k2, fmaf(k1, k1, k0 * k0));
    float dt = fmaf(q2, k2, fmaf(q1, k1, q0 * k0));
    // num * rcp(den) instead of num / den: the numerator is exactly 0 whenever <q,k> >= 0, which sends the
    // IEEE division through its slow path (exponent check on the NUMERATOR) for practically every warp --
    // 14 % of the layer-2 kernel in the round-1 profile.  den >= 1e-24 is always a normal number.
    float t = (one_minus_slope * fminf(dt, 0.f)) * __frcp_rn(fmaxf(n2, EPS_NRM2));
    o0 = fmaf(-t, k0, q0);
    o1 = fmaf(-t, k1, q1);
    o2 = fmaf(-t, k2, q2);
}

__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// vn_act on TWO channels at once: the products and FMAs as packed f32x2 operations (FMUL2 / FFMA2), the clamp / reciprocal
// per channel.  Every operation and its order equal the scalar vn_act: bit-identical results, ~half the instructions.
__device__ __forceinline__ void vn_act2(f32x2_t q0, f32x2_t q1, f32x2_t q2, f32x2_t k0, f32x2_t k1, f32x2_t k2, float one_minus_slope,
                                        f32x2_t& o0, f32x2_t& o1, f32x2_t& o2) {
    const f32x2_t n2 = fma2(k2, k2, fma2(k1, k1, mul2(k0, k0)));
    const f32x2_t dt = fma2(q2, k2, fma2(q1, k1, mul2(q0, k0)));
    float n2a, n2b, dta, dtb;
    unpack2(n2, n2a, n2b);
    unpack2(dt, dta, dtb);
    const float ta = (one_minus_slope * fminf(dta, 0.f)) * __frcp_rn(fmaxf(n2a, EPS_NRM2));
    const float tb = (one_minus_slope * fminf(dtb, 0.f)) * __frcp_rn(fmaxf(n2b, EPS_NRM2));
    const f32x2_t nt = pack2(-ta, -tb);
    o0 = fma2(nt, k0, q0);
    o1 = fma2(nt, k1, q1);
    o2 = fma2(nt, k2, q2);
}
// the same on the 4 channels of three float4 (one per axis)
__device__ __forceinline__ void vn_act4(const float4* q, const float4* k, float one_minus_slope, float4* o) {
    f32x2_t a0, a1, a2, b0, b1, b2;
    vn_act2(pack2(q[0].x, q[0].y), pack2(q[1].x, q[1].y), pack2(q[2].x, q[2].y), pack2(k[0].x, k[0].y), pack2(k[1].x, k[1].y),
            pack2(k[2].x, k[2].y), one_minus_slope, a0, a1, a2);
    vn_act2(pack2(q[0].z, q[0].w), pack2(q[1].z, q[1].w), pack2(q[2].z, q[2].w), pack2(k[0].z, k[0].w), pack2(k[1].z, k[1].w),
            pack2(k[2].z, k[2].w), one_minus_slope, b0, b1, b2);
    unpack2(a0, o[0].x, o[0].y);
    unpack2(b0, o[0].z, o[0].w);
    unpack2(a1, o[1].x, o[1].y);
    unpack2(b1, o[1].z, o[1].w);
    unpack2(a2, o[2].x, o[2].y);
    unpack2(b2, o[2].z, o[2].w);
}
__device__ __forceinline__ float4 add4_packed(float4 x, float4 y) {
    float4 r;
    unpack2(add2(pack2(x.x, x.y), pack2(y.x, y.y)), r.x, r.y);
    unpack2(add2(pack2(x.z, x.w), pack2(y.z, y.w)), r.z, r.w);
    return r;
}

// ---- GEMM launcher shared by the encoder and the SDF decoder (ls_gemm.cu) -------------------
struct GemmArgs {
    const float* W;   // [R][ldw] row-major, ldw >= K, ldw % 4 == 0, zero padded beyond K
    const float* Wtc; // optional: the same weights packed by tc_pack_weights (tcgen05 3xTF32 path)
    const float* X;   // element (b, k, n) at X[b*x_sb + k*x_sk + n]
    float* out;
    int R, K, ldw;
    int n_per_b;      // columns per instance (n3)
    int B;
    long long x_sb, x_sk;
    // store mode 0 (channel major): out[b*o_sb + r*o_sr + n]
    long long o_sb, o_sr;
    // store mode 1 (point major gather table): out[((b*npts + pt) * R*3) + (part*3+axis)*c_out + c]
    int point_major, npts, c_out;
    // epilogue: + bias[b*bias_sb + r*bias_sr + axis*(bias_axis)] with axis = n / npts; relu
    const float* bias;
    long long bias_sb, bias_sr;
    int bias_axis;
    int relu;
    // optional (store mode 0 only): zero the result where mask[same offset as out] <= 0 -- the ReLU derivative of the
    // decoder's backward pass, with the forward activation as the mask
    const float* mask;
};
int launch_gemm(const GemmArgs& a, cudaStream_t st);       // dispatches to the tcgen05 path when possible
int launch_gemm_simt(const GemmArgs& a, cudaStream_t st);
// tcgen05 3xTF32 path (ls_gemm_tc.cu)
size_t tc_packed_floats(int R, int K);
int tc_pack_weights(const float* W, int R, int K, int ldw, float* packed, cudaStream_t st);
bool gemm_tc_supported(const GemmArgs& a);
int launch_gemm_tc(const GemmArgs& a, const float* packed, cudaStream_t st);
// TS-form variant (ls_gemm_tc3.cu): activations in tensor memory, 128- or 256-row weight tiles
size_t tc3_packed_floats(int R, int K);
int tc3_pack_weights(const float* W, int R, int K, int ldw, float* packed256, cudaStream_t st);
int launch_gemm_tc3(const GemmArgs& a, const float* packed128, const float* packed256, cudaStream_t st);
extern bool g_use_tensor_cores;
extern int g_gemm_variant;  // 3 = k_gemm_tc3 (TS form, default), 2 = persistent SS-form k_gemm_tc2, 1 = round-1 k_gemm_tc

}  // namespace ls
