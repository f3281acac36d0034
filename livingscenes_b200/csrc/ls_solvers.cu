// lib_more solvers behind the C ABI: sequential / mutual-NN matching on invariant codes and the
// batched weighted Kabsch SE(3) fit (one warp per pair, 3x3 one-sided Jacobi SVD in registers).
//   matcher_new.py:85-139, pose_estimation.py:29-121 (paths relative to the reference root).
#include <float.h>

#include "ls_common.cuh"

namespace ls {
namespace {

constexpr int MAX_PAIRS_PER_LAUNCH = 48;
struct PairTable {
    int n_pairs;
    int off0[MAX_PAIRS_PER_LAUNCH + 1];
    int off1[MAX_PAIRS_PER_LAUNCH + 1];
    long long ws[MAX_PAIRS_PER_LAUNCH];  // float offset of the pair's scratch in the workspace
};

__host__ __device__ inline size_t pair_ws_floats(int n, int m, int dim) {
    return ((size_t)(n + m) * dim + (size_t)n * m + 63) & ~size_t(63);
}

// block-wide arg-reductions (all threads get the result)
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
    return r;
}
__device__ __forceinline__ int block_min_int(int v, int* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int r = red[0];
    for (int i = 1; i < nw; ++i) r = min(r, red[i]);
    return r;
}

// rows / max(|row|, 1e-12) (F.normalize, matcher_new.py:110-111) and S = A B^T (:120).
// ``stage`` (optional shared memory, n*dim + dim*33 floats, only for n, m <= 32): the normalised rows are kept in
// shared memory -- A row-major (broadcast reads), B transposed with a padded stride (conflict-free) -- instead of
// being re-read from global memory with 32 scattered sectors per load; the FMA order over d is the same.
__device__ void normalize_and_score(const float* z0, const float* z1, int n, int m, int dim, float* an,
                                    float* bn, float* S, float* stage = nullptr) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* sA = stage;
    float* sBt = stage ? stage + (size_t)n * dim : nullptr;
    for (int r = w; r < n + m; r += nw) {
        const float* src = r < n ? z0 + (size_t)r * dim : z1 + (size_t)(r - n) * dim;
        float* dst = r < n ? an + (size_t)r * dim : bn + (size_t)(r - n) * dim;
        float s = 0.f;
        for (int d = lane; d < dim; d += 32) s = fmaf(src[d], src[d], s);
        s = warp_sum(s);
        const float den = fmaxf(sqrtf(s), EPS_NRM);
        for (int d = lane; d < dim; d += 32) {
            const float v = src[d] / den;
            if (stage) {
                if (r < n) sA[(size_t)r * dim + d] = v;
                else sBt[(size_t)d * 33 + (r - n)] = v;
            } else {
                dst[d] = v;
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
        const int i = e / m, j = e - i * m;
        float acc = 0.f;
        if (stage) {
            const float* a = sA + (size_t)i * dim;
            const float* b = sBt + j;
#pragma unroll 8
            for (int d = 0; d < dim; ++d) acc = fmaf(a[d], b[(size_t)d * 33], acc);
        } else {
            const float* a = an + (size_t)i * dim;
            const float* b = bn + (size_t)j * dim;
#pragma unroll 8
            for (int d = 0; d < dim; ++d) acc = fmaf(a[d], b[d], acc);  // loads run ahead, the FMA order is fixed
        }
        S[e] = acc;
    }
    __syncthreads();
}

// sequential_matcher (matcher_new.py:109-139).  Every round replays the reference's fp32 sequence:
// S <- S / (max(S) + 1e-5) on the surviving rows/cols, pick the FIRST (row-major) entry equal to the
// new max, record the pair, retire its row and column.
__global__ void __launch_bounds__(1024) k_seq_match(const float* __restrict__ z0, const float* __restrict__ z1,
                                                    int dim, const PairTable tab, float* __restrict__ wsf,
                                                    int64_t* __restrict__ m0, int64_t* __restrict__ m1, int stage_off,
                                                    int stage_floats, const float* __restrict__ res, int score_mode) {
    __shared__ float redf[32];
    __shared__ int redi[32];
    extern __shared__ unsigned char alive[];  // [n] rows then [m] cols
    const int p = blockIdx.x;
    const int o0 = tab.off0[p], o1 = tab.off1[p];
    const int n = tab.off0[p + 1] - o0, m = tab.off1[p + 1] - o1;
    float* an = wsf + tab.ws[p];
    float* bn = an + (size_t)n * dim;
    float* S = bn + (size_t)m * dim;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        alive[i] = 1;
        m0[o0 + i] = -1;
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        alive[n + j] = 1;
        m1[o1 + j] = -1;
    }
    if (n == 0 || m == 0) return;
    // small pairs: the launcher provides shared memory for the normalised rows behind the alive flags
    float* stage = (stage_floats > 0 && n <= 32 && m <= 32 && (size_t)n * dim + (size_t)dim * 33 <= (size_t)stage_floats)
                       ? reinterpret_cast<float*>(alive + stage_off)
                       : nullptr;
    normalize_and_score(z0 + (size_t)o0 * dim, z1 + (size_t)o1 * dim, n, m, dim, an, bn, S, stage);
    if (res != nullptr) {
        // sim3_seq_matcher / eq_seq_matcher (matcher_new.py:142-230): the cosine score is divided by (mode 1), or
        // replaced by the inverse of (mode 2), the mean Kabsch residual of the pair's equivariant codes
        for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
            const float r = res[e] + 1e-5f;
            S[e] = score_mode == 1 ? S[e] / r : 1.f / r;
        }
        __syncthreads();
    }
    const int rounds = min(n, m);
    if (n <= 32 && m <= 32) {
        // Small scenes (the common case: <= 32 instances per scan): one warp holds the whole score matrix in
        // registers (lane = row) and replays the same fp32 sequence without block-wide barriers -- the general
        // path below spends its time in 3 block reductions per round.
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x;
        float s[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = (lane < n && j < m) ? S[lane * m + j] : 0.f;
        bool row_alive = lane < n;
        unsigned col_alive = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
        for (int r = 0; r < rounds; ++r) {
            float mx = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (row_alive && ((col_alive >> j) & 1u)) mx = fmaxf(mx, s[j]);
            mx = warp_max(mx);
            const float den = mx + 1e-5f;
            float mx2 = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (row_alive && ((col_alive >> j) & 1u)) {
                    s[j] = s[j] / den;
                    mx2 = fmaxf(mx2, s[j]);
                }
            }
            mx2 = warp_max(mx2);
            int first = 0x7fffffff;
#pragma unroll
            for (int j = 31; j >= 0; --j)
                if (row_alive && ((col_alive >> j) & 1u) && s[j] == mx2) first = lane * m + j;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(FULL, first, o));
            if (first == 0x7fffffff) break;  // NaN scores: nothing compares equal; leave the rest unmatched
            const int bi = first / m, bj = first - bi * m;
            if (lane == 0) {
                m0[o0 + bi] = bj;
                m1[o1 + bj] = bi;
            }
            if (lane == bi) row_alive = false;
            col_alive &= ~(1u << bj);
        }
        return;
    }
    for (int r = 0; r < rounds; ++r) {
        float mx = -FLT_MAX;
        for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
            const int i = e / m, j = e - i * m;
            if (alive[i] && alive[n + j]) mx = fmaxf(mx, S[e]);
        }
        mx = block_max(mx, redf);
        const float den = mx + 1e-5f;
        float mx2 = -FLT_MAX;
        for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
            const int i = e / m, j = e - i * m;
            if (alive[i] && alive[n + j]) {
                const float v = S[e] / den;
                S[e] = v;
                mx2 = fmaxf(mx2, v);
            }
        }
        mx2 = block_max(mx2, redf);
        int first = 0x7fffffff;
        for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
            const int i = e / m, j = e - i * m;
            if (alive[i] && alive[n + j] && S[e] == mx2) {
                first = e;
                break;  // e increases along the thread's stride: the first hit is its minimum
            }
        }
        first = block_min_int(first, redi);
        if (first == 0x7fffffff) break;  // NaN scores: nothing compares equal; leave the rest unmatched
        const int bi = first / m, bj = first - bi * m;
        __syncthreads();
        if (threadIdx.x == 0) {
            m0[o0 + bi] = bj;
            m1[o1 + bj] = bi;
            alive[bi] = 0;
            alive[n + bj] = 0;
        }
        __syncthreads();
    }
}

// nn_matcher (matcher_new.py:85-105): cosine top-1 in both directions (first index on ties), keep
// mutual pairs.
__global__ void __launch_bounds__(1024) k_mutual_nn(const float* __restrict__ z0, const float* __restrict__ z1,
                                                    int dim, const PairTable tab, float* __restrict__ wsf,
                                                    int64_t* __restrict__ m0, int64_t* __restrict__ m1) {
    extern __shared__ int best[];  // [n] best col per row, then [m] best row per col
    const int p = blockIdx.x;
    const int o0 = tab.off0[p], o1 = tab.off1[p];
    const int n = tab.off0[p + 1] - o0, m = tab.off1[p + 1] - o1;
    float* an = wsf + tab.ws[p];
    float* bn = an + (size_t)n * dim;
    float* S = bn + (size_t)m * dim;
    if (n == 0 || m == 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) m0[o0 + i] = -1;
        for (int j = threadIdx.x; j < m; j += blockDim.x) m1[o1 + j] = -1;
        return;
    }
    normalize_and_score(z0 + (size_t)o0 * dim, z1 + (size_t)o1 * dim, n, m, dim, an, bn, S);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float bv = -FLT_MAX;
        int bj = 0;
        for (int j = 0; j < m; ++j) {
            const float v = S[(size_t)i * m + j];
            if (v > bv) {
                bv = v;
                bj = j;
            }
        }
        best[i] = bj;
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        float bv = -FLT_MAX;
        int bi = 0;
        for (int i = 0; i < n; ++i) {
            const float v = S[(size_t)i * m + j];
            if (v > bv) {
                bv = v;
                bi = i;
            }
        }
        best[n + j] = bi;
    }
    __syncthreads();
    // mutual_check(m0, m1) then mutual_check(m1, m0_checked) (:95-96)
    for (int i = threadIdx.x; i < n; i += blockDim.x) m0[o0 + i] = (best[n + best[i]] == i) ? best[i] : -1;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int i = best[n + j];
        const bool ok0 = best[n + best[i]] == i;        // m0_checked[i] survives
        m1[o1 + j] = (ok0 && best[i] == j) ? i : -1;
    }
}

// ------------------------------------------------------------------------------------------ Kabsch
struct Mat3 {
    double m[3][3];
};

__device__ __forceinline__ double det3(const Mat3& a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
           a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
           a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// One-sided (Hestenes) Jacobi: A V = U diag(sig), columns sorted by descending sig.
__device__ void svd3(Mat3 A, Mat3& U, Mat3& V, double sig[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) V.m[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 15; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double al = 0, be = 0, ga = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                al += A.m[i][p] * A.m[i][p];
                be += A.m[i][q] * A.m[i][q];
                ga += A.m[i][p] * A.m[i][q];
            }
            if (fabs(ga) > 1e-300 && fabs(ga) > 1e-17 * sqrt(al * be)) {
                off = fmax(off, fabs(ga) / sqrt(al * be));
                const double zeta = (be - al) / (2.0 * ga);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double ap = A.m[i][p], aq = A.m[i][q];
                    A.m[i][p] = c * ap - s * aq;
                    A.m[i][q] = s * ap + c * aq;
                    const double vp = V.m[i][p], vq = V.m[i][q];
                    V.m[i][p] = c * vp - s * vq;
                    V.m[i][q] = s * vp + c * vq;
                }
            }
        }
        if (off < 1e-15) break;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
        sig[j] = sqrt(A.m[0][j] * A.m[0][j] + A.m[1][j] * A.m[1][j] + A.m[2][j] * A.m[2][j]);
    // sort columns by descending singular value (3-element network)
#define LS_SWAPCOL(x, y)                                   \
    if (sig[x] < sig[y]) {                                 \
        double ts = sig[x]; sig[x] = sig[y]; sig[y] = ts;  \
        for (int i = 0; i < 3; ++i) {                      \
            double ta = A.m[i][x]; A.m[i][x] = A.m[i][y]; A.m[i][y] = ta; \
            double tv = V.m[i][x]; V.m[i][x] = V.m[i][y]; V.m[i][y] = tv; \
        }                                                  \
    }
    LS_SWAPCOL(0, 1)
    LS_SWAPCOL(1, 2)
    LS_SWAPCOL(0, 1)
#undef LS_SWAPCOL
    const double tiny = 1e-14 * fmax(sig[0], 1e-300);
    // U columns; complete a rank-deficient basis so that U stays orthonormal
    if (sig[0] > 1e-300) {
        for (int i = 0; i < 3; ++i) U.m[i][0] = A.m[i][0] / sig[0];
    } else {
        U.m[0][0] = 1; U.m[1][0] = 0; U.m[2][0] = 0;
    }
    if (sig[1] > tiny) {
        for (int i = 0; i < 3; ++i) U.m[i][1] = A.m[i][1] / sig[1];
    } else {
        // any unit vector orthogonal to u0
        const double ax = fabs(U.m[0][0]), ay = fabs(U.m[1][0]), az = fabs(U.m[2][0]);
        double e[3] = {0, 0, 0};
        if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
        const double dp = e[0] * U.m[0][0] + e[1] * U.m[1][0] + e[2] * U.m[2][0];
        double v[3], nn = 0;
        for (int i = 0; i < 3; ++i) { v[i] = e[i] - dp * U.m[i][0]; nn += v[i] * v[i]; }
        nn = sqrt(nn);
        for (int i = 0; i < 3; ++i) U.m[i][1] = v[i] / nn;
    }
    if (sig[2] > tiny) {
        for (int i = 0; i < 3; ++i) U.m[i][2] = A.m[i][2] / sig[2];
    } else {
        U.m[0][2] = U.m[1][0] * U.m[2][1] - U.m[2][0] * U.m[1][1];
        U.m[1][2] = U.m[2][0] * U.m[0][1] - U.m[0][0] * U.m[2][1];
        U.m[2][2] = U.m[0][0] * U.m[1][1] - U.m[1][0] * U.m[0][1];
    }
}

struct KabschArgs {
    // generic form
    const float* x1;
    const float* x2;
    const float* w;
    // from-codes form (more_solver.py:114-116)
    const float* za;
    const float* ta;
    const float* zb;
    const float* tb;
    const int64_t* match;
    int b, n, normalize_w;
    float eps;
    float *R, *t, *res;
};

template <bool CODES>
__global__ void __launch_bounds__(128) k_kabsch(const KabschArgs a) {
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= a.b) return;
    const int n = a.n;
    const float* p1;
    const float* p2;
    float o1[3] = {0.f, 0.f, 0.f}, o2[3] = {0.f, 0.f, 0.f};
    if (CODES) {
        const long long mi = a.match[pair];
        if (mi < 0) {  // unmatched: identity, zero translation, zero residuals
            if (lane < 9) a.R[(size_t)pair * 9 + lane] = (lane % 4 == 0) ? 1.f : 0.f;
            if (lane < 3) a.t[(size_t)pair * 3 + lane] = 0.f;
            if (a.res)
                for (int j = lane; j < n; j += 32) a.res[(size_t)pair * n + j] = 0.f;
            return;
        }
        p1 = a.za + (size_t)pair * n * 3;
        p2 = a.zb + (size_t)mi * n * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o1[i] = a.ta[(size_t)pair * 3 + i];
            o2[i] = a.tb[(size_t)mi * 3 + i];
        }
    } else {
        p1 = a.x1 + (size_t)pair * n * 3;
        p2 = a.x2 + (size_t)pair * n * 3;
    }
    const float* wp = (!CODES && a.w) ? a.w + (size_t)pair * n : nullptr;

    // weights (pose_estimation.py:49-56): w / (sum w + eps) when normalize_w
    float wsum = 0.f;
    for (int j = lane; j < n; j += 32) wsum += wp ? wp[j] : 1.f;
    wsum = warp_sum(wsum);
    const float wden = a.normalize_w ? (wsum + a.eps) : 1.f;
    // weighted means (:70-71): sum(w x) / (sum(w) + eps)
    float sw = 0.f, s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f};
    for (int j = lane; j < n; j += 32) {
        const float wj = (wp ? wp[j] : 1.f) / wden;
        sw += wj;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            s1[i] = fmaf(wj, p1[j * 3 + i] + o1[i], s1[i]);
            s2[i] = fmaf(wj, p2[j * 3 + i] + o2[i], s2[i]);
        }
    }
    sw = warp_sum(sw);
    float mu1[3], mu2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        mu1[i] = warp_sum(s1[i]) / (sw + a.eps);
        mu2[i] = warp_sum(s2[i]) / (sw + a.eps);
    }
    // covariance (:73-77): C = X1c^T diag(w) X2c
    float c[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) c[i] = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float wj = (wp ? wp[j] : 1.f) / wden;
        float a1[3], a2[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            a1[i] = (p1[j * 3 + i] + o1[i]) - mu1[i];
            a2[i] = ((p2[j * 3 + i] + o2[i]) - mu2[i]) * wj;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) c[r * 3 + q] = fmaf(a1[r], a2[q], c[r * 3 + q]);
    }
    Mat3 C;
#pragma unroll
    for (int i = 0; i < 9; ++i) C.m[i / 3][i % 3] = (double)warp_sum(c[i]);
    // SVD and the proper-rotation fix (:80-94): R = V diag(1,1,det(V U^T)) U^T
    Mat3 U, V;
    double sig[3];
    svd3(C, U, V, sig);
    const double dsign = det3(V) * det3(U) < 0 ? -1.0 : 1.0;
    float R[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            R[r * 3 + q] = (float)(V.m[r][0] * U.m[q][0] + V.m[r][1] * U.m[q][1] + dsign * V.m[r][2] * U.m[q][2]);
    float tt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) tt[r] = mu2[r] - (R[r * 3] * mu1[0] + R[r * 3 + 1] * mu1[1] + R[r * 3 + 2] * mu1[2]);
    if (lane < 9) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            if (i == lane) v = R[i];
        a.R[(size_t)pair * 9 + lane] = v;
    }
    if (lane < 3) {
        float v = lane == 0 ? tt[0] : (lane == 1 ? tt[1] : tt[2]);
        a.t[(size_t)pair * 3 + lane] = v;
    }
    if (a.res) {  // transformation_residuals (:105-121)
        for (int j = lane; j < n; j += 32) {
            const float x = p1[j * 3] + o1[0], y = p1[j * 3 + 1] + o1[1], z = p1[j * 3 + 2] + o1[2];
            const float e0 = (R[0] * x + R[1] * y + R[2] * z + tt[0]) - (p2[j * 3] + o2[0]);
            const float e1 = (R[3] * x + R[4] * y + R[5] * z + tt[1]) - (p2[j * 3 + 1] + o2[1]);
            const float e2 = (R[6] * x + R[7] * y + R[8] * z + tt[2]) - (p2[j * 3 + 2] + o2[2]);
            a.res[(size_t)pair * n + j] = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
        }
    }
}

// ============================================================================================
// ICP refinement of the pose fit (more_solver.py:182-187 -> pytorch3d.ops.iterative_closest_point with
// init_transform, estimate_scale=False, allow_reflection=False; restated in oracle/p3d_shim.py):
//   Xt = s X R + T (row vectors);  loop: nn = 1-NN of Xt in Y (direct-form fp32 distance, lowest index on ties);
//   (R, T) = corresponding_points_alignment(X, nn): means, XYcov = Xc^T Yc / N, SVD, R = U diag(1,1,det(U V^T)) V^T,
//   T = Ymu - Xmu R;  Xt = X R + T;  rmse = sqrt(mean |Xt - nn|^2);  stop when (prev - rmse) / prev <= thr.
// One CTA per pair: Y as float4 in shared memory (broadcast reads), every thread owns N/blockDim points of X.
// ============================================================================================
constexpr int ICP_THREADS = 512, ICP_PPT = 8;  // N <= 4096
struct IcpArgs {
    const float* X;   // [B][N][3]
    const float* Y;   // [B][M][3]
    const float* R0;  // optional [B][3][3] row-vector convention (Xt = X R + T)
    const float* T0;  // optional [B][3]
    int N, M, max_iter;
    float rel_thr;
    float *R, *T, *rmse, *Xt;  // Xt optional [B][N][3]
    int32_t* n_iter;           // iterations executed; negated when the convergence test never fired
};

__device__ __forceinline__ void icp_block_sum(float* v, int n, float* red, float* out) {
    // sums v[0..n) over the block (n <= 9); result broadcast through out[0..n)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = 0; i < n; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0)
        for (int i = 0; i < n; ++i) red[w * 9 + i] = v[i];
    __syncthreads();
    if (threadIdx.x < n) {
        float s = 0.f;
        for (int k = 0; k < nw; ++k) s += red[k * 9 + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(ICP_THREADS) k_icp(const IcpArgs a) {
    extern __shared__ float4 sY[];  // [M]
    __shared__ float red[(ICP_THREADS / 32) * 9];
    __shared__ float bc[12];  // broadcast slot: sums, then R (9) + T (3)
    __shared__ int s_stop;
    const int b = blockIdx.x, t = threadIdx.x, N = a.N, M = a.M;
    const float* Xb = a.X + (size_t)b * N * 3;
    const float* Yb = a.Y + (size_t)b * M * 3;
    for (int j = t; j < M; j += ICP_THREADS) sY[j] = make_float4(Yb[j * 3], Yb[j * 3 + 1], Yb[j * 3 + 2], 0.f);
    float x[ICP_PPT][3], xt[ICP_PPT][3], nn[ICP_PPT][3];
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, T[3] = {0.f, 0.f, 0.f};
    if (a.R0)
        for (int i = 0; i < 9; ++i) R[i] = a.R0[(size_t)b * 9 + i];
    if (a.T0)
        for (int i = 0; i < 3; ++i) T[i] = a.T0[(size_t)b * 3 + i];
    float v[9];
    v[0] = v[1] = v[2] = 0.f;
#pragma unroll
    for (int p = 0; p < ICP_PPT; ++p) {
        const int i = t + p * ICP_THREADS;
        const bool ok = i < N;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            x[p][c] = ok ? Xb[i * 3 + c] : 0.f;
            v[c] += x[p][c];
        }
    }
    icp_block_sum(v, 3, red, bc);
    const float xmu[3] = {bc[0] / (float)N, bc[1] / (float)N, bc[2] / (float)N};
    auto apply = [&]() {
#pragma unroll
        for (int p = 0; p < ICP_PPT; ++p)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                xt[p][c] = (x[p][0] * R[c] + x[p][1] * R[3 + c] + x[p][2] * R[6 + c]) + T[c];  // bmm(X, R) + T
    };
    apply();
    __syncthreads();
    float prev = -1.f, rmse = 0.f;
    int it = 0;
    bool converged = false;
    for (; it < a.max_iter; ++it) {
        // ---- nearest neighbour of every transformed point
        v[0] = v[1] = v[2] = 0.f;
#pragma unroll
        for (int p = 0; p < ICP_PPT; ++p) {
            const int i = t + p * ICP_THREADS;
            float best = FLT_MAX;
            int bj = 0;
            if (i < N) {
                for (int j = 0; j < M; ++j) {
                    const float4 y = sY[j];
                    const float dx = xt[p][0] - y.x, dy = xt[p][1] - y.y, dz = xt[p][2] - y.z;
                    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (d < best) {
                        best = d;
                        bj = j;
                    }
                }
                const float4 y = sY[bj];
                nn[p][0] = y.x;
                nn[p][1] = y.y;
                nn[p][2] = y.z;
            } else {
                nn[p][0] = nn[p][1] = nn[p][2] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] += nn[p][c];
        }
        icp_block_sum(v, 3, red, bc);
        const float ymu[3] = {bc[0] / (float)N, bc[1] / (float)N, bc[2] / (float)N};
        // ---- XYcov = Xc^T Yc / N
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = 0.f;
#pragma unroll
        for (int p = 0; p < ICP_PPT; ++p) {
            if (t + p * ICP_THREADS < N) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[r * 3 + c] = fmaf(x[p][r] - xmu[r], nn[p][c] - ymu[c], v[r * 3 + c]);
            }
        }
        icp_block_sum(v, 9, red, bc);
        if (t == 0) {
            Mat3 C, U, V;
            double sig[3];
            for (int i = 0; i < 9; ++i) C.m[i / 3][i % 3] = (double)(bc[i] / (float)N);
            svd3(C, U, V, sig);
            const double dsign = det3(U) * det3(V) < 0 ? -1.0 : 1.0;  // det(U V^T)
            float Rn[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    Rn[r * 3 + c] = (float)(U.m[r][0] * V.m[c][0] + U.m[r][1] * V.m[c][1] + dsign * U.m[r][2] * V.m[c][2]);
            for (int i = 0; i < 9; ++i) bc[i] = Rn[i];
            for (int c = 0; c < 3; ++c)
                bc[9 + c] = ymu[c] - (xmu[0] * Rn[c] + xmu[1] * Rn[3 + c] + xmu[2] * Rn[6 + c]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = bc[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) T[c] = bc[9 + c];
        __syncthreads();
        apply();
        // ---- rmse and the relative-improvement test
        v[0] = 0.f;
#pragma unroll
        for (int p = 0; p < ICP_PPT; ++p) {
            if (t + p * ICP_THREADS < N) {
                const float dx = xt[p][0] - nn[p][0], dy = xt[p][1] - nn[p][1], dz = xt[p][2] - nn[p][2];
                v[0] += dx * dx + dy * dy + dz * dz;
            }
        }
        icp_block_sum(v, 1, red, bc);
        rmse = sqrtf(bc[0] / (float)N);
        if (t == 0) s_stop = (prev >= 0.f && (prev - rmse) / prev <= a.rel_thr) ? 1 : 0;
        __syncthreads();
        if (s_stop) {
            converged = true;
            ++it;
            break;
        }
        prev = rmse;
    }
    if (t < 9) a.R[(size_t)b * 9 + t] = R[t];
    if (t < 3) a.T[(size_t)b * 3 + t] = T[t];
    if (t == 0) {
        a.rmse[b] = rmse;
        a.n_iter[b] = converged ? it : -it;
    }
    if (a.Xt) {
#pragma unroll
        for (int p = 0; p < ICP_PPT; ++p) {
            const int i = t + p * ICP_THREADS;
            if (i < N)
                for (int c = 0; c < 3; ++c) a.Xt[((size_t)b * N + i) * 3 + c] = xt[p][c];
        }
    }
}

// ------------------------------------------------------------------------------------------ Sinkhorn matcher
// sinkhorn_matcher (matcher_new.py:11-71): cosine scores / sqrt(desc_dim), log-space optimal transport with a
// dustbin row/column (alpha = 1), `iters` Sinkhorn iterations, then mutual arg-max on the [n,m] block with the
// exp(score) > threshold test.  One CTA; the (n+1) x (m+1) coupling matrix lives in shared memory.
struct SinkArgs {
    const float* z0;
    const float* z1;
    int n, m, dim, iters;
    float alpha, thr, inv_sqrt_dim;
    float* ws;  // normalised rows + raw scores (pair_ws_floats)
    int64_t *m0, *m1;
};

__device__ __forceinline__ float lse_finish(float mx, float sum) { return mx + logf(sum); }

__global__ void __launch_bounds__(1024) k_sinkhorn_match(const SinkArgs a) {
    extern __shared__ float sm[];
    const int n = a.n, m = a.m, ld = m + 1;
    float* Z = sm;                       // [(n+1)][(m+1)]
    float* u = Z + (size_t)(n + 1) * ld;  // [n+1]
    float* v = u + (n + 1);               // [m+1]
    int* best = reinterpret_cast<int*>(v + (m + 1));  // [n] then [m]
    float* bval = reinterpret_cast<float*>(best + n + m);  // [n]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* an = a.ws;
    float* bn = an + (size_t)n * a.dim;
    float* S = bn + (size_t)m * a.dim;
    normalize_and_score(a.z0, a.z1, n, m, a.dim, an, bn, S);
    for (int e = threadIdx.x; e < (n + 1) * ld; e += blockDim.x) {
        const int i = e / ld, j = e - i * ld;
        Z[e] = (i < n && j < m) ? S[i * m + j] * a.inv_sqrt_dim : a.alpha;
    }
    for (int i = threadIdx.x; i <= n; i += blockDim.x) u[i] = 0.f;
    for (int j = threadIdx.x; j <= m; j += blockDim.x) v[j] = 0.f;
    __syncthreads();
    // log_mu = [norm]*n + [log(m) + norm], log_nu = [norm]*m + [log(n) + norm], norm = -log(n + m)   (:30-33)
    const float norm = -logf((float)(n + m));
    const float mu_last = logf((float)m) + norm, nu_last = logf((float)n) + norm;
    for (int it = 0; it < a.iters; ++it) {
        // u = log_mu - logsumexp_j(Z + v)
        for (int i = w; i <= n; i += nw) {
            float mx = -FLT_MAX;
            for (int j = lane; j <= m; j += 32) mx = fmaxf(mx, Z[i * ld + j] + v[j]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int j = lane; j <= m; j += 32) sum += expf(Z[i * ld + j] + v[j] - mx);
            sum = warp_sum(sum);
            if (lane == 0) u[i] = (i < n ? norm : mu_last) - lse_finish(mx, sum);
        }
        __syncthreads();
        // v = log_nu - logsumexp_i(Z + u)
        for (int j = w; j <= m; j += nw) {
            float mx = -FLT_MAX;
            for (int i = lane; i <= n; i += 32) mx = fmaxf(mx, Z[i * ld + j] + u[i]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int i = lane; i <= n; i += 32) sum += expf(Z[i * ld + j] + u[i] - mx);
            sum = warp_sum(sum);
            if (lane == 0) v[j] = (j < m ? norm : nu_last) - lse_finish(mx, sum);
        }
        __syncthreads();
    }
    // scores = Z + u + v - norm on the [n,m] block; row / column arg-max (first index on ties)
    for (int i = w; i < n; i += nw) {
        float bv = -FLT_MAX;
        int bj = 0x7fffffff;
        for (int j = lane; j < m; j += 32) {
            const float x = Z[i * ld + j] + u[i] + v[j] - norm;
            if (x > bv) { bv = x; bj = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oj = __shfl_xor_sync(FULL, bj, o);
            if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        if (lane == 0) { best[i] = bj; bval[i] = bv; }
    }
    for (int j = w; j < m; j += nw) {
        float bv = -FLT_MAX;
        int bi = 0x7fffffff;
        for (int i = lane; i < n; i += 32) {
            const float x = Z[i * ld + j] + u[i] + v[j] - norm;
            if (x > bv) { bv = x; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) best[n + j] = bi;
    }
    __syncthreads();
    // mutual check + score threshold (:57-66)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool valid0 = best[n + best[i]] == i && expf(bval[i]) > a.thr;
        a.m0[i] = valid0 ? best[i] : -1;
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int i = best[n + j];
        const bool mutual1 = best[i] == j;
        const bool valid0 = best[n + best[i]] == i && expf(bval[i]) > a.thr;
        a.m1[j] = (mutual1 && valid0) ? i : -1;
    }
}

int fill_table(const int32_t* off0, const int32_t* off1, int first, int count, int dim, size_t ws_base,
               PairTable& tab, size_t& ws_end, int& max_nm, int& max_n_plus_m) {
    tab.n_pairs = count;
    size_t off = ws_base;
    max_nm = 0;
    max_n_plus_m = 0;
    for (int i = 0; i <= count; ++i) {
        tab.off0[i] = off0[first + i];
        tab.off1[i] = off1[first + i];
    }
    for (int i = 0; i < count; ++i) {
        const int n = tab.off0[i + 1] - tab.off0[i], m = tab.off1[i + 1] - tab.off1[i];
        LS_REQUIRE(n >= 0 && m >= 0, "offsets must be non-decreasing");
        LS_REQUIRE(n + m <= 32768, "pair too large");
        tab.ws[i] = (long long)off;
        off += pair_ws_floats(n, m, dim);
        max_nm = n * m > max_nm ? n * m : max_nm;
        max_n_plus_m = n + m > max_n_plus_m ? n + m : max_n_plus_m;
    }
    ws_end = off;
    return LS_OK;
}

template <bool SEQ>
int run_match(const float* z0, const float* z1, int dim, const int32_t* off0, const int32_t* off1, int n_pairs,
              int64_t* m0, int64_t* m1, void* ws, size_t ws_bytes, cudaStream_t st, const float* res = nullptr,
              int score_mode = 0) {
    LS_REQUIRE(z0 && z1 && off0 && off1 && m0 && m1, "null pointer");
    LS_REQUIRE(n_pairs >= 0 && dim >= 1, "bad sizes");
    size_t ws_off = 0;
    for (int first = 0; first < n_pairs; first += MAX_PAIRS_PER_LAUNCH) {
        const int count = n_pairs - first < MAX_PAIRS_PER_LAUNCH ? n_pairs - first : MAX_PAIRS_PER_LAUNCH;
        PairTable tab;
        size_t ws_end;
        int max_nm, max_npm;
        int rc = fill_table(off0, off1, first, count, dim, ws_off, tab, ws_end, max_nm, max_npm);
        if (rc != LS_OK) return rc;
        if (ws_end * sizeof(float) > ws_bytes || (ws_end > 0 && ws == nullptr)) {
            set_error("match workspace too small");
            return LS_ERR_WORKSPACE;
        }
        ws_off = ws_end;
        // small pairs: one score per thread (the greedy rounds then run in a single warp); mid-size pairs keep the
        // block reductions of the greedy loop cheap with 8 warps
        const int threads = max_nm <= 1024 ? 1024 : (max_nm <= 4096 ? 256 : 1024);
        if (SEQ) {
            // alive flags, then (small pairs only) the shared-memory stage of the normalised rows
            const int stage_off = (max_npm + 16 + 15) & ~15;
            int stage_floats = 0;
            if (max_nm <= 1024 && max_npm <= 64) stage_floats = 32 * dim + dim * 33;
            size_t smem = (size_t)stage_off + (size_t)stage_floats * sizeof(float);
            if (smem > 200 * 1024) {
                stage_floats = 0;
                smem = (size_t)stage_off;
            }
            if (smem > 48 * 1024)
                LS_CHECK_CUDA(cudaFuncSetAttribute(k_seq_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_seq_match<<<count, threads, smem, st>>>(z0, z1, dim, tab, static_cast<float*>(ws), m0, m1, stage_off,
                                                      stage_floats, res, score_mode);
            LS_CHECK_LAUNCH("k_seq_match");
        } else {
            const size_t smem = (size_t)max_npm * sizeof(int) + 16;
            if (smem > 48 * 1024)
                LS_CHECK_CUDA(cudaFuncSetAttribute(k_mutual_nn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_mutual_nn<<<count, threads, smem, st>>>(z0, z1, dim, tab, static_cast<float*>(ws), m0, m1);
            LS_CHECK_LAUNCH("k_mutual_nn");
        }
    }
    return LS_OK;
}

}  // namespace
}  // namespace ls

using namespace ls;

extern "C" {

int ls_match_workspace_bytes(const int32_t* off0_host, const int32_t* off1_host, int32_t n_pairs, size_t* bytes) {
    LS_REQUIRE(off0_host && off1_host && bytes && n_pairs >= 0, "bad arguments");
    size_t fl = 0;
    for (int i = 0; i < n_pairs; ++i) {
        const int n = off0_host[i + 1] - off0_host[i], m = off1_host[i + 1] - off1_host[i];
        LS_REQUIRE(n >= 0 && m >= 0, "offsets must be non-decreasing");
        fl += pair_ws_floats(n, m, 256);
    }
    *bytes = fl * sizeof(float) + 256;
    return LS_OK;
}

int ls_seq_match(const float* z0, const float* z1, int32_t dim, const int32_t* off0_host, const int32_t* off1_host,
                 int32_t n_pairs, int64_t* matches0, int64_t* matches1, void* workspace, size_t workspace_bytes,
                 void* stream) {
    LS_REQUIRE(dim <= 256, "descriptor dimension above 256 is not supported by ls_match_workspace_bytes");
    return run_match<true>(z0, z1, dim, off0_host, off1_host, n_pairs, matches0, matches1, workspace,
                           workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ls_mutual_nn(const float* z0, const float* z1, int32_t dim, const int32_t* off0_host, const int32_t* off1_host,
                 int32_t n_pairs, int64_t* matches0, int64_t* matches1, void* workspace, size_t workspace_bytes,
                 void* stream) {
    LS_REQUIRE(dim <= 256, "descriptor dimension above 256 is not supported by ls_match_workspace_bytes");
    return run_match<false>(z0, z1, dim, off0_host, off1_host, n_pairs, matches0, matches1, workspace,
                            workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ls_seq_match_scored(const float* z0, const float* z1, int32_t dim, int32_t n, int32_t m, const float* res,
                        int32_t score_mode, int64_t* matches0, int64_t* matches1, void* workspace, size_t workspace_bytes,
                        void* stream) {
    LS_REQUIRE(dim <= 256, "descriptor dimension above 256 is not supported by ls_match_workspace_bytes");
    LS_REQUIRE(res != nullptr && (score_mode == 1 || score_mode == 2), "score_mode must be 1 (sim3_seq) or 2 (eq_seq)");
    const int32_t off0[2] = {0, n}, off1[2] = {0, m};
    return run_match<true>(z0, z1, dim, off0, off1, 1, matches0, matches1, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream), res, score_mode);
}

int ls_sinkhorn_match(const float* z0, const float* z1, int32_t dim, int32_t n, int32_t m, int32_t iters, float alpha,
                      float match_threshold, int64_t* matches0, int64_t* matches1, void* workspace, size_t workspace_bytes,
                      void* stream) {
    LS_REQUIRE(z0 && z1 && matches0 && matches1 && workspace, "null pointer");
    LS_REQUIRE(n >= 1 && m >= 1 && dim >= 1 && dim <= 256 && iters >= 0, "bad sizes");
    LS_REQUIRE(pair_ws_floats(n, m, dim) * sizeof(float) <= workspace_bytes, "sinkhorn workspace too small");
    const size_t smem = sizeof(float) * ((size_t)(n + 1) * (m + 1) + (n + 1) + (m + 1) + n) + sizeof(int) * (size_t)(n + m) + 16;
    LS_REQUIRE(smem <= 220 * 1024, "sinkhorn: the (n+1) x (m+1) coupling matrix must fit shared memory (n, m <= ~230)");
    SinkArgs a{};
    a.z0 = z0;
    a.z1 = z1;
    a.n = n;
    a.m = m;
    a.dim = dim;
    a.iters = iters;
    a.alpha = alpha;
    a.thr = match_threshold;
    a.inv_sqrt_dim = 1.f / sqrtf((float)dim);
    a.ws = static_cast<float*>(workspace);
    a.m0 = matches0;
    a.m1 = matches1;
    LS_CHECK_CUDA(cudaFuncSetAttribute(k_sinkhorn_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sinkhorn_match<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
    LS_CHECK_LAUNCH("k_sinkhorn_match");
    return LS_OK;
}

int ls_kabsch_batched(const float* x1, const float* x2, const float* weights, int32_t b, int32_t n,
                      int32_t normalize_w, float eps, float* R, float* t, float* res, void* stream) {
    LS_REQUIRE(x1 && x2 && R && t, "null pointer");
    LS_REQUIRE(b >= 0 && n >= 1, "bad sizes");
    if (b == 0) return LS_OK;
    KabschArgs a{};
    a.x1 = x1;
    a.x2 = x2;
    a.w = weights;
    a.b = b;
    a.n = n;
    a.normalize_w = normalize_w;
    a.eps = eps;
    a.R = R;
    a.t = t;
    a.res = res;
    k_kabsch<false><<<(b + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    LS_CHECK_LAUNCH("k_kabsch");
    return LS_OK;
}

int ls_kabsch_from_codes(const float* z_so3_a, const float* t_a, const float* z_so3_b, const float* t_b,
                         const int64_t* match, int32_t n_pairs, int32_t c_dim, float* R, float* t, float* res,
                         void* stream) {
    LS_REQUIRE(z_so3_a && t_a && z_so3_b && t_b && match && R && t, "null pointer");
    LS_REQUIRE(n_pairs >= 0 && c_dim >= 1, "bad sizes");
    if (n_pairs == 0) return LS_OK;
    KabschArgs a{};
    a.za = z_so3_a;
    a.ta = t_a;
    a.zb = z_so3_b;
    a.tb = t_b;
    a.match = match;
    a.b = n_pairs;
    a.n = c_dim;
    a.normalize_w = 1;
    a.eps = 1e-7f;
    a.R = R;
    a.t = t;
    a.res = res;
    k_kabsch<true><<<(n_pairs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    LS_CHECK_LAUNCH("k_kabsch_codes");
    return LS_OK;
}

int ls_icp(const float* X, const float* Y, int32_t B, int32_t N, int32_t M, const float* R0, const float* T0,
           int32_t max_iterations, float relative_rmse_thr, float* R, float* T, float* rmse, int32_t* n_iter, float* Xt,
           void* stream) {
    LS_REQUIRE(X && Y && R && T && rmse && n_iter, "null pointer");
    LS_REQUIRE(B >= 1 && N >= 1 && N <= ICP_THREADS * ICP_PPT && M >= 1 && M <= 12288, "icp: need N <= 4096, M <= 12288");
    LS_REQUIRE(max_iterations >= 1, "icp: max_iterations must be >= 1");
    IcpArgs a{};
    a.X = X;
    a.Y = Y;
    a.R0 = R0;
    a.T0 = T0;
    a.N = N;
    a.M = M;
    a.max_iter = max_iterations;
    a.rel_thr = relative_rmse_thr;
    a.R = R;
    a.T = T;
    a.rmse = rmse;
    a.n_iter = n_iter;
    a.Xt = Xt;
    const size_t smem = (size_t)M * sizeof(float4);
    LS_CHECK_CUDA(cudaFuncSetAttribute(k_icp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_icp<<<B, ICP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
    LS_CHECK_LAUNCH("k_icp");
    return LS_OK;
}

}  // extern "C"
