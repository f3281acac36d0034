// Device kernels of the VN-DGCNN+attention encoder (sm_100a).  Host orchestration: ls_encoder.cu.
// Reference arithmetic: SURVEY.md Appendix A; file:line citations are relative to the reference root.
#pragma once
#include <float.h>

#include "ls_common.cuh"
#include "ls_knn_tc.cuh"

namespace ls {

// ============================================================================================
// Shape_Prior.encode pre-processing (model_utils.py:171-177): centroid removal and
// scale_0 = mean of the 5 largest entries of the flattened N x N distance matrix.
// One CTA per instance; the centred cloud lives in shared memory; every ordered pair (i,j) is
// visited, so each unordered pair is counted twice exactly like the reference's flattened
// symmetric matrix (=> scale_0 = (2 d1 + 2 d2 + d3) / 5 for distinct pair distances).
// ============================================================================================
constexpr int NORM_THREADS = 512;

__global__ void __launch_bounds__(NORM_THREADS) k_normalize(const float* __restrict__ x, int N,
                                                            float* __restrict__ xn,
                                                            float* __restrict__ centroid,
                                                            float* __restrict__ s0_out) {
    extern __shared__ float sm[];
    float* sx = sm;           // [3][N]
    float* red = sm + 3 * N;  // [NORM_THREADS/32 * 3] + merge scratch
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nw = NORM_THREADS / 32;
    const float* xb = x + (size_t)b * 3 * N;

    float s[3] = {0.f, 0.f, 0.f};
    for (int i = t; i < N; i += NORM_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = xb[a * N + i];
            sx[a * N + i] = v;
            s[a] += v;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        s[a] = warp_sum(s[a]);
        if (lane == 0) red[w * 3 + a] = s[a];
    }
    __syncthreads();
    float mu[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = 0.f;
        for (int i = 0; i < nw; ++i) v += red[i * 3 + a];
        mu[a] = v / (float)N;
    }
    __syncthreads();
    for (int i = t; i < N; i += NORM_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) sx[a * N + i] -= mu[a];
    }
    __syncthreads();

    // The reference takes the top-5 of the flattened SYMMETRIC matrix: every unordered pair appears twice with the same
    // bits (d(i,j) and d(j,i) differ by exact negations only), so the top-5 is (p1, p1, p2, p2, p3) for the three largest
    // unordered-pair values p1 >= p2 >= p3 (ties included).  Each thread therefore visits the pairs j > i of rows i and
    // N-1-i (N-1 pairs per row pair: balanced), four j per 16-byte shared-memory load, and keeps a top-3.
    float top[3] = {-1.f, -1.f, -1.f};
    auto push = [&](float d2) {
        if (d2 > top[2]) {
            top[2] = d2;
            if (top[2] > top[1]) {
                const float tmp = top[2];
                top[2] = top[1];
                top[1] = tmp;
            }
            if (top[1] > top[0]) {
                const float tmp = top[1];
                top[1] = top[0];
                top[0] = tmp;
            }
        }
    };
    auto row = [&](int i) {
        const float xi = sx[i], yi = sx[N + i], zi = sx[2 * N + i];
        auto one = [&](int j) {
            const float dx = xi - sx[j], dy = yi - sx[N + j], dz = zi - sx[2 * N + j];
            push(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        };
        int j = i + 1;
        if ((N & 3) == 0) {
            for (; j < N && (j & 3) != 0; ++j) one(j);
            for (; j + 3 < N; j += 4) {
                const float4 xj = *reinterpret_cast<const float4*>(sx + j);
                const float4 yj = *reinterpret_cast<const float4*>(sx + N + j);
                const float4 zj = *reinterpret_cast<const float4*>(sx + 2 * N + j);
                float dx = xi - xj.x, dy = yi - yj.x, dz = zi - zj.x;
                push(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                dx = xi - xj.y, dy = yi - yj.y, dz = zi - zj.y;
                push(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                dx = xi - xj.z, dy = yi - yj.z, dz = zi - zj.z;
                push(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                dx = xi - xj.w, dy = yi - yj.w, dz = zi - zj.w;
                push(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            }
        }
        for (; j < N; ++j) one(j);
    };
    for (int i = t; 2 * i < N; i += NORM_THREADS) {
        row(i);
        if (N - 1 - i != i) row(N - 1 - i);
    }
    // merge: 3 rounds of block-wide max with removal
    __shared__ float s_best[NORM_THREADS / 32];
    __shared__ int s_who[NORM_THREADS / 32];
    __shared__ float s_top3[3];
    __shared__ float s_sel[5];
    int p = 0;
    for (int r = 0; r < 3; ++r) {
        float cand = -1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k == p) cand = top[k];
        float v = cand;
        int who = t;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float v2 = __shfl_xor_sync(FULL, v, o);
            int w2 = __shfl_xor_sync(FULL, who, o);
            if (v2 > v || (v2 == v && w2 < who)) {
                v = v2;
                who = w2;
            }
        }
        if (lane == 0) {
            s_best[w] = v;
            s_who[w] = who;
        }
        __syncthreads();
        float bv = s_best[0];
        int bw = s_who[0];
        for (int i = 1; i < nw; ++i) {
            if (s_best[i] > bv || (s_best[i] == bv && s_who[i] < bw)) {
                bv = s_best[i];
                bw = s_who[i];
            }
        }
        if (t == bw) ++p;
        if (t == 0) s_top3[r] = bv;
        __syncthreads();
    }
    if (t == 0) {  // the flattened matrix holds every pair twice (a cloud of < 3 pairs pads with the diagonal's zeros)
        s_sel[0] = s_sel[1] = s_top3[0];
        s_sel[2] = s_sel[3] = s_top3[1];
        s_sel[4] = s_top3[2];
    }
    __syncthreads();
    float s0 = 0.f;
#pragma unroll
    for (int r = 0; r < 5; ++r) s0 += sqrtf(fmaxf(s_sel[r], 0.f));
    s0 = s0 / 5.f;
    if (t == 0) {
        s0_out[b] = s0;
#pragma unroll
        for (int a = 0; a < 3; ++a) centroid[b * 3 + a] = mu[a];
    }
    float* xo = xn + (size_t)b * 3 * N;
    for (int i = t; i < N; i += NORM_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) xo[a * N + i] = sx[a * N + i] / s0;
    }
}

// ============================================================================================
// Farthest point sampling chain (pytorch3d sample_farthest_points as used at
// vec_dgcnn_atten.py:163-175): start index 0, min_d update, arg-max with lowest index on ties;
// d = dx*dx + dy*dy + dz*dz with every operation rounded separately (no FMA contraction) so the
// selection is bit-identical to the oracle.  FPS only depends on xyz, so the three down-sampling
// steps of the shipped encoder (N -> N/2 -> N/8 -> N/32) run back to back in ONE launch, one CTA
// per instance, coordinates in shared memory, running min-distance in registers, one
// __syncthreads per selected point.
// ============================================================================================
constexpr int FPS_MAX_LEVELS = 4;
struct FpsArgs {
    const float* xyz;  // [B][3][N]
    int N;
    int n_levels;
    int n_out[FPS_MAX_LEVELS];
    int* sel32[FPS_MAX_LEVELS];          // optional [B][n_out] int32 (internal use)
    int64_t* sel64[FPS_MAX_LEVELS];    // optional [B][n_out] int64 (API tap)
    const int64_t* force[FPS_MAX_LEVELS];  // optional forced selections (teacher forcing)
    float* out_xyz;                      // optional [B][3][n_out[last]]
    const int64_t* start;                // optional [B] first selected index of level 0 (default 0)
    int fma;                             // squared distance as fma(dz,dz,fma(dy,dy,dx*dx)) instead of 3 mul + 2 add
};

// Squared distance of the FPS update.  pytorch3d's CUDA kernel accumulates `dist2 += diff * diff` over the 3
// coordinates, which nvcc contracts to FMAs by default; its CPU implementation and torch expressions round every
// product and sum.  The two differ in the last bit, which can flip a near-tie arg-max.  Default: separately rounded
// (what the oracle / fixtures use); ls_set_fps_fma(1) selects the contracted form.  pytorch3d is not vendored in the
// reference tree, so which of the two its 0.7.4 binary really executes is unpinned.
__device__ __forceinline__ float fps_dist(float dx, float dy, float dz, int fma) {
    return fma ? fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)))
               : __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <int PPT>
__global__ void __launch_bounds__(1024) k_fps(const FpsArgs a) {
    extern __shared__ float sm[];
    const int T = blockDim.x;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = T >> 5;
    const int N = a.N;
    float* cur = sm;                              // [3][N]
    float* nxt = sm + 3 * N;                      // [3][n_out[0]]  (ping-pong partner)
    int* ssel = (int*)(sm + 3 * N + 3 * a.n_out[0]);  // [n_out[0]]
    __shared__ float wb_v[2][32];
    __shared__ int wb_i[2][32];

    const float* xb = a.xyz + (size_t)b * 3 * N;
    for (int i = t; i < 3 * N; i += T) cur[i] = xb[i];
    __syncthreads();

    int n_cur = N;
    for (int lv = 0; lv < a.n_levels; ++lv) {
        const int n_out = a.n_out[lv];
        if (a.force[lv]) {
            for (int j = t; j < n_out; j += T) ssel[j] = (int)a.force[lv][(size_t)b * n_out + j];
            __syncthreads();
        } else {
            float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                int i = t + p * T;
                bool ok = i < n_cur;
                px[p] = ok ? cur[i] : 0.f;
                py[p] = ok ? cur[n_cur + i] : 0.f;
                pz[p] = ok ? cur[2 * n_cur + i] : 0.f;
                md[p] = ok ? FLT_MAX : -1.f;  // padding can never win the arg-max
            }
            int last = (lv == 0 && a.start) ? (int)a.start[b] : 0;
            if (t == 0) ssel[0] = last;
            for (int j = 1; j < n_out; ++j) {
                const float lx = cur[last], ly = cur[n_cur + last], lz = cur[2 * n_cur + last];
                float bv = -2.f;
                int bi = 0x7fffffff;
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    float dx = px[p] - lx, dy = py[p] - ly, dz = pz[p] - lz;
                    float d = fps_dist(dx, dy, dz, a.fma);
                    float m = fminf(md[p], d);
                    if (md[p] >= 0.f) md[p] = m;
                    if (md[p] > bv) {  // strict: lowest index within the thread wins
                        bv = md[p];
                        bi = t + p * T;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    float v2 = __shfl_xor_sync(FULL, bv, o);
                    int i2 = __shfl_xor_sync(FULL, bi, o);
                    if (v2 > bv || (v2 == bv && i2 < bi)) {
                        bv = v2;
                        bi = i2;
                    }
                }
                const int par = j & 1;
                if (lane == 0) {
                    wb_v[par][w] = bv;
                    wb_i[par][w] = bi;
                }
                __syncthreads();
                bv = lane < nw ? wb_v[par][lane] : -2.f;
                bi = lane < nw ? wb_i[par][lane] : 0x7fffffff;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    float v2 = __shfl_xor_sync(FULL, bv, o);
                    int i2 = __shfl_xor_sync(FULL, bi, o);
                    if (v2 > bv || (v2 == bv && i2 < bi)) {
                        bv = v2;
                        bi = i2;
                    }
                }
                last = bi;
                if (t == 0) ssel[j] = last;
            }
            __syncthreads();
        }
        // publish the selection and build the next level's coordinates (selection order)
        for (int j = t; j < n_out; j += T) {
            int s = ssel[j];
            if (a.sel32[lv]) a.sel32[lv][(size_t)b * n_out + j] = s;
            if (a.sel64[lv]) a.sel64[lv][(size_t)b * n_out + j] = s;
            nxt[j] = cur[s];
            nxt[n_out + j] = cur[n_cur + s];
            nxt[2 * n_out + j] = cur[2 * n_cur + s];
        }
        __syncthreads();
        float* tmp = cur;
        cur = nxt;
        nxt = tmp;
        n_cur = n_out;
    }
    if (a.out_xyz) {
        float* o = a.out_xyz + (size_t)b * 3 * n_cur;
        for (int i = t; i < 3 * n_cur; i += T) o[i] = cur[i];
    }
}

// Large / ragged clouds (N > 8192, or masked instances of different sizes: model_utils.py:199-205): one CTA per
// instance, the points and the running min-distance live in a caller-provided scratch as float4 {x,y,z,min_d}
// per point (L2 resident: one 16-byte load and one 4-byte store per point and step), same arithmetic and tie rule
// as k_fps.  With a mask the valid points are first compacted (stable order, like pc.T[mask]) into the scratch, so
// a whole batch of ragged instances is sampled by ONE launch; indices then refer to the compacted list.
__device__ __forceinline__ void fps_scratch_loop(float4* __restrict__ sb, int N, int n_out, int last, int b,
                                                 int64_t* __restrict__ sel64, float* __restrict__ out_xyz, int fma) {
    __shared__ float wb_v[2][32];
    __shared__ int wb_i[2][32];
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = T >> 5;
    auto publish = [&](int j, int sel) {
        if (sel64) sel64[(size_t)b * n_out + j] = sel;
        if (out_xyz) {
            const float4 p = sb[sel];
            float* o = out_xyz + (size_t)b * 3 * n_out + j;
            o[0] = p.x;
            o[n_out] = p.y;
            o[2 * n_out] = p.z;
        }
    };
    if (t == 0) publish(0, last);
    // fewer points than samples: every point is selected once, the remaining slots are padded the way pytorch3d pads
    // them (index -1, coordinates 0) -- the reference's encode_fps then encodes those zeros as points (model_utils.py:205)
    const int n_sel = n_out < N ? n_out : N;
    for (int j = n_sel + t; j < n_out; j += T) {
        if (sel64) sel64[(size_t)b * n_out + j] = -1;
        if (out_xyz) {
            float* o = out_xyz + (size_t)b * 3 * n_out + j;
            o[0] = 0.f;
            o[n_out] = 0.f;
            o[2 * n_out] = 0.f;
        }
    }
    for (int j = 1; j < n_sel; ++j) {
        const float4 lp = sb[last];
        const float lx = lp.x, ly = lp.y, lz = lp.z;
        float bv = -2.f;
        int bi = 0x7fffffff;
        for (int i = t; i < N; i += T) {
            const float4 p = sb[i];
            const float dx = p.x - lx, dy = p.y - ly, dz = p.z - lz;
            const float d = fps_dist(dx, dy, dz, fma);
            const float m = fminf(p.w, d);
            sb[i].w = m;
            if (m > bv) {  // strict: lowest index within the thread wins
                bv = m;
                bi = i;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(FULL, bv, o);
            const int i2 = __shfl_xor_sync(FULL, bi, o);
            if (v2 > bv || (v2 == bv && i2 < bi)) {
                bv = v2;
                bi = i2;
            }
        }
        const int par = j & 1;
        if (lane == 0) {
            wb_v[par][w] = bv;
            wb_i[par][w] = bi;
        }
        __syncthreads();
        bv = lane < nw ? wb_v[par][lane] : -2.f;
        bi = lane < nw ? wb_i[par][lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(FULL, bv, o);
            const int i2 = __shfl_xor_sync(FULL, bi, o);
            if (v2 > bv || (v2 == bv && i2 < bi)) {
                bv = v2;
                bi = i2;
            }
        }
        last = bi;
        if (t == 0) publish(j, last);
    }
}

__global__ void __launch_bounds__(1024) k_fps_large(const float* __restrict__ xyz, int N, int n_out,
                                                    const int64_t* __restrict__ start, float4* __restrict__ scr,
                                                    int64_t* __restrict__ sel64, float* __restrict__ out_xyz, int fma) {
    const int T = blockDim.x, b = blockIdx.x, t = threadIdx.x;
    const float* xb = xyz + (size_t)b * 3 * N;
    float4* sb = scr + (size_t)b * N;
    for (int i = t; i < N; i += T) sb[i] = make_float4(xb[i], xb[N + i], xb[2 * N + i], FLT_MAX);
    __syncthreads();
    fps_scratch_loop(sb, N, n_out, start ? (int)start[b] : 0, b, sel64, out_xyz, fma);
}

// xyz [B][3][Nmax], mask [B][Nmax] (bytes, non-zero = valid).  n_valid[b] receives the number of valid points; an
// instance with fewer valid points than n_out selects all of them and pads with index -1 / zero coordinates (pytorch3d).
__global__ void __launch_bounds__(1024) k_fps_masked(const float* __restrict__ xyz, const unsigned char* __restrict__ mask,
                                                     int Nmax, int n_out, const int64_t* __restrict__ start,
                                                     float4* __restrict__ scr, int32_t* __restrict__ n_valid,
                                                     int64_t* __restrict__ sel64, float* __restrict__ out_xyz, int fma) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int T = blockDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = T >> 5;
    const float* xb = xyz + (size_t)b * 3 * Nmax;
    const unsigned char* mb = mask + (size_t)b * Nmax;
    float4* sb = scr + (size_t)b * Nmax;
    if (t == 0) s_base = 0;
    __syncthreads();
    for (int c = 0; c < Nmax; c += T) {  // stable compaction, one block-wide prefix sum per chunk of T points
        const int i = c + t;
        const bool ok = i < Nmax && mb[i] != 0;
        const unsigned bal = __ballot_sync(FULL, ok);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < w; ++k) off += s_warp[k];
        if (ok) sb[off + __popc(bal & ((1u << lane) - 1u))] = make_float4(xb[i], xb[Nmax + i], xb[2 * Nmax + i], FLT_MAX);
        __syncthreads();
        if (t == 0) {
            int tot = 0;
            for (int k = 0; k < nw; ++k) tot += s_warp[k];
            s_base += tot;
        }
        __syncthreads();
    }
    const int N = s_base;
    if (t == 0 && n_valid) n_valid[b] = N;
    if (N == 0) return;
    int first = start ? (int)start[b] : 0;
    first = min(max(first, 0), N - 1);
    fps_scratch_loop(sb, N, n_out, first, b, sel64, out_xyz, fma);
}

// dst_f[b][r][j] = src_f[b][r][sel[b][j]]   (vec_dgcnn_atten.py:173; r runs over C*3 rows)
// thread = (row, selected point) flattened per instance, so that small n_out (32 in the deep layers) still fills warps
__global__ void __launch_bounds__(256) k_gather_points(const float* __restrict__ in, const int* __restrict__ sel, int rows,
                                                       int n_in, int n_out, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * n_out) return;
    const int r = e / n_out, j = e - r * n_out;
    const int s = __ldg(sel + (size_t)b * n_out + j);
    out[(size_t)b * rows * n_out + e] = __ldg(in + ((size_t)b * rows + r) * n_in + s);
}

// ============================================================================================
// Fused kNN graph + VN-EdgeConv + pooling   (get_graph_feature + V/K/Q VecLNA + mean / attention
// pool: vec_dgcnn_atten.py:124-161, 197-219; vec_layers.py:121-134, 241-268, 24-31).
//
// One CTA = 64 dst points of one instance, 256 threads.
//   phase 1  brute-force feature-space kNN: 8x8 register tiles of squared distances accumulated in
//            fp32 in the direct form sum_d (q_d - s_d)^2 over shared-memory staged [d][point] tiles
//            (64 queries x 256 sources x 8 dims per stage, double buffered); a warp owns 8 queries
//            and keeps each query's sorted top-16 distributed over lanes 0..15; candidates that
//            beat the current 16th are merged with ballot/shuffle insertion.  Ties -> lower index.
//   phase 2  per dst point (one warp, lanes = output channels): gather the 16 neighbours' rows of
//            the point-level GEMM tables, add the dst term, VN leaky-ReLU, then
//              MODE_MEAN : mean over the 16 edges                               (layers 0-1)
//              MODE_ATT  : cevn(K) . cevn(Q) logits per 16-channel head, softmax over the 16
//                          edges (half-warp shuffles), weighted sum of V        (layers >= 2)
//              MODE_L0   : layer 0 builds its 3-channel edge feature [cross, nn-dst, dst]
//                          directly from xyz (vec_dgcnn_atten.py:153-158).
// ============================================================================================
constexpr int QT = 64, ST = 256, DKC = 8, EDGE_THREADS = 256, QCAP = 32;
enum { MODE_L0 = 0, MODE_MEAN = 1, MODE_ATT = 2, MODE_KNN_ONLY = 3 };

struct EdgeArgs {
    const float* src_f;  // [B][D][Ns] kNN features of the sources (xyz for layer 0)
    const float* dst_f;  // [B][D][Nd] kNN features of the queries
    int B, D, Ns, Nd;
    int qpc;             // dst points handled per CTA (multiple of 8, <= QT)
    const float* psrc;   // [B][Ns][row_s]  gather table  (parts Vq,Vk[,Kq,Kk] x 3 axes x Co)
    const float* pdst;   // [B][Nd][row_d]  dst table     (parts Vq,Vk[,Kq,Kk,Qq,Qk])
    int row_s, row_d;
    const float* w0;     // layer 0: [2][Co][3]
    int Co;
    float oms;           // 1 - negative slope
    float* out;          // [B][Co][3][Nd]
    int64_t* idx_out;        // optional [B][Nd][16]
    const int64_t* idx_in;   // optional: graph given (teacher forcing, or built by k_knn_small)
    float* dist_out;         // optional [B][Nd][16] (MODE_KNN_ONLY)
};

// ---- warp-level top-k machinery.  A candidate is ONE 64-bit key: (float bits of the squared distance
//      << 32) | source index.  Squared distances are >= +0, so unsigned integer order on the key is
//      exactly the lexicographic (distance, index) order: ties resolve towards the LOWER source index,
//      as pytorch3d's strict '<' replacement does, and a NaN distance sorts after everything.
typedef unsigned long long u64;
constexpr u64 KEY_MAX = ~0ull;
__device__ __forceinline__ u64 make_key(float d, int s) {
    return ((u64)__float_as_uint(d) << 32) | (unsigned)s;
}
__device__ __forceinline__ float key_dist(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_idx(u64 k) { return (int)(unsigned)(k & 0xffffffffu); }

__device__ __forceinline__ void bitonic_sort32(u64& v, int lane, bool desc) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 o = __shfl_xor_sync(FULL, v, j);
            const bool up = (((lane & k) == 0) != desc);
            const bool lower = (lane & j) == 0;
            v = (lower == up) ? (o < v ? o : v) : (o > v ? o : v);
        }
    }
}
__device__ __forceinline__ void bitonic_merge32(u64& v, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const u64 o = __shfl_xor_sync(FULL, v, j);
        v = ((lane & j) == 0) ? (o < v ? o : v) : (o > v ? o : v);
    }
}
// 16th smallest of 32 lane values (value-only bitonic sort)
__device__ __forceinline__ float warp_kth16(float v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(FULL, v, j);
            const bool up = (lane & k) == 0;
            const bool lower = (lane & j) == 0;
            v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return __shfl_sync(FULL, v, 15);
}
// Out-of-line on purpose: the fused kernel calls these from unrolled per-query sites; inlining the
// sorting networks there blows the kernel up to >1 MB of SASS and thrashes the instruction cache.
__device__ __noinline__ float warp_kth16_nl(float v) { return warp_kth16(v, threadIdx.x & 31); }

// Merge a query's shared-memory queue (cnt <= 32 unsorted keys) into its sorted list (one key per lane,
// ascending; only entries 0..15 are ever consumed).  Returns the new list element of this lane.
__device__ __noinline__ u64 knn_flush(u64 lk, const u64* q, int cnt) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const bool empty = __shfl_sync(FULL, lk, 0) == KEY_MAX;
    if (empty || cnt <= 16) {
        // the 16 kept list entries and up to 16 (32 if the list is empty) queue entries fit one 32-sort
        const int qi = empty ? lane : lane - 16;
        u64 v = (!empty && lane < 16) ? lk : ((qi >= 0 && qi < cnt) ? q[qi] : KEY_MAX);
        bitonic_sort32(v, lane, false);
        return v;
    }
    u64 b = lane < cnt ? q[lane] : KEY_MAX;
    bitonic_sort32(b, lane, true);
    lk = b < lk ? b : lk;  // half-cleaner of the bitonic sequence (ascending list) ++ (descending queue)
    bitonic_merge32(lk, lane);
    return lk;
}

// Slow path of the per-tile selection (queue overflow): slot by slot, each slot holds at most 32
// candidates so it always fits after a flush.  keys[j] == KEY_MAX marks an invalid slot.
struct SelRet {
    u64 lk;
    int cnt;
};
__device__ __noinline__ SelRet knn_select_slow(u64 k0, u64 k1, u64 k2, u64 k3, u64 k4, u64 k5, u64 k6, u64 k7,
                                               u64 lk, int cnt, u64* q) {
    const int lane = threadIdx.x & 31;
    const u64 kv[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
    u64 tau = __shfl_sync(FULL, lk, 15);
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        u64 key = KEY_MAX;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) key = (jj == j) ? kv[jj] : key;
        bool ok = key < tau;
        unsigned m = __ballot_sync(FULL, ok);
        if (m == 0) continue;
        if (cnt + __popc(m) > QCAP) {
            lk = knn_flush(lk, q, cnt);
            cnt = 0;
            tau = __shfl_sync(FULL, lk, 15);
            ok = key < tau;
            m = __ballot_sync(FULL, ok);
        }
        if (ok) q[cnt + __popc(m & ((1u << lane) - 1u))] = key;
        cnt += __popc(m);
    }
    __syncwarp();
    return SelRet{lk, cnt};
}

// P1 = false: phase-2-only instantiation (the graph always comes from idx_in: k_knn_rerank / k_knn_small /
// teacher forcing) without the phase-1 tiles and queues, compiled for 3 CTAs per SM.
#ifndef LS_P2_CTAS
#define LS_P2_CTAS 3
#endif
#ifndef LS_P2_WIDE_CPL
#define LS_P2_WIDE_CPL 8  // c_out >= 256: 2 CTAs per SM (128 registers): these layers are issue bound and spill at 80
#endif
template <int MODE, int CPL, bool P1 = true>
__global__ void __launch_bounds__(EDGE_THREADS, (P1 || CPL >= LS_P2_WIDE_CPL) ? 2 : LS_P2_CTAS) k_knn_edge(const EdgeArgs a) {
    // phase-1 tiles and phase-2 scratch share one buffer
    constexpr int TILE_FLOATS = P1 ? 2 * DKC * (QT + ST) : 4;
    constexpr int SR_FLOATS = (MODE == MODE_ATT) ? 8 * (CPL > 4 ? CPL / 4 : 1) * LS_KNN_K * 8 : 0;
    constexpr int BUF_FLOATS = TILE_FLOATS > SR_FLOATS ? TILE_FLOATS : SR_FLOATS;
    __shared__ __align__(16) float sbuf[BUF_FLOATS];
    __shared__ int sIdx[QT][LS_KNN_K];
    __shared__ float sDist[(MODE == MODE_KNN_ONLY) ? QT : 1][LS_KNN_K];
    __shared__ u64 sQ[P1 ? 8 : 1][P1 ? 8 : 1][P1 ? QCAP : 1];  // [warp][query][slot] candidate queues (64-bit keys)

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int b = blockIdx.y;
    const int qpc = a.qpc;
    const int q0 = blockIdx.x * qpc;
    const int Ns = a.Ns, Nd = a.Nd, D = a.D;
    const int nq = min(qpc, Nd - q0);  // valid dst points of this CTA

    if (!P1 || a.idx_in != nullptr) {
        for (int e = t; e < nq * LS_KNN_K; e += EDGE_THREADS) {
            int ql = e >> 4, k = e & 15;
            sIdx[ql][k] = (int)a.idx_in[((size_t)b * Nd + q0 + ql) * LS_KNN_K + k];
        }
    } else if constexpr (P1) {
        // ------------------------------------------------------------------ phase 1: kNN
        float(*Qs)[DKC][QT] = reinterpret_cast<float(*)[DKC][QT]>(sbuf);
        float(*Ss)[DKC][ST] = reinterpret_cast<float(*)[DKC][ST]>(sbuf + 2 * DKC * QT);
        const float* srcb = a.src_f + (size_t)b * D * Ns;
        const float* dstb = a.dst_f + (size_t)b * D * Nd;
        const int n_tiles = (Ns + ST - 1) / ST;
        const int n_chunks = (D + DKC - 1) / DKC;
        const int n_it = n_tiles * n_chunks;
        const bool warp_on = w * 8 < nq;  // warps without valid queries only help staging the tiles

        float sreg[DKC], qreg[2];
        auto g_load = [&](int it) {
            const int tile = it / n_chunks, chunk = it - tile * n_chunks;
            const int s = tile * ST + t, d0 = chunk * DKC;
#pragma unroll
            for (int dd = 0; dd < DKC; ++dd) {
                int d = d0 + dd;
                sreg[dd] = (s < Ns && d < D) ? __ldg(srcb + (size_t)d * Ns + s) : 0.f;
            }
            const int ql = t & 63;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int d = d0 + (t >> 6) + 4 * i;
                qreg[i] = (ql < nq && d < D) ? __ldg(dstb + (size_t)d * Nd + q0 + ql) : 0.f;
            }
        };
        auto s_store = [&](int buf) {
#pragma unroll
            for (int dd = 0; dd < DKC; ++dd) Ss[buf][dd][t] = sreg[dd];
#pragma unroll
            for (int i = 0; i < 2; ++i) Qs[buf][(t >> 6) + 4 * i][t & 63] = qreg[i];
        };

        // per query: the best keys so far, ascending over the lanes (entry 15 is the running 16th)
        u64 lk[8];
        f32x2_t acc2[8][4];  // squared distances of 8 queries x 8 sources, as 4 packed fp32 pairs per query
        int cnt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            lk[i] = KEY_MAX;
            cnt[i] = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
        }

        g_load(0);
        s_store(0);
        __syncthreads();
        for (int it = 0; it < n_it; ++it) {
            const int buf = it & 1;
            if (it + 1 < n_it) g_load(it + 1);
            if (warp_on) {
#pragma unroll
                for (int dd = 0; dd < DKC; ++dd) {
                    const float4 qa = *reinterpret_cast<const float4*>(&Qs[buf][dd][w * 8]);
                    const float4 qb = *reinterpret_cast<const float4*>(&Qs[buf][dd][w * 8 + 4]);
                    const float4 sa = *reinterpret_cast<const float4*>(&Ss[buf][dd][lane * 4]);
                    const float4 sb = *reinterpret_cast<const float4*>(&Ss[buf][dd][128 + lane * 4]);
                    const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                    const f32x2_t sp[4] = {pack2(sa.x, sa.y), pack2(sa.z, sa.w), pack2(sb.x, sb.y), pack2(sb.z, sb.w)};
                    // acc += (q - s)^2 in the direct form, two sources per FADD2 / FFMA2
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const f32x2_t qq = pack2(qv[i], qv[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const f32x2_t df = sub2(qq, sp[j]);
                            acc2[i][j] = fma2(df, df, acc2[i][j]);
                        }
                    }
                }
                const int tile = it / n_chunks, chunk = it - tile * n_chunks;
                if (chunk == n_chunks - 1) {
                    // ---- this tile's 8 x (32 lanes x 8) distances are final.  Per query: candidates that beat
                    //      the running 16th-best key are compacted into the query's shared-memory queue
                    //      (8-bit pass mask per lane + one warp prefix sum, no per-slot ballots/branches);
                    //      the queue is merged into the sorted list by a warp bitonic network when it holds
                    //      more than 16 keys and after the last tile.  On the first tile the threshold
                    //      starts from the 16th smallest of the 32 lane minima (an upper bound of the
                    //      tile's 16th smallest distance).
                    const int sbase = tile * ST + lane * 4;
                    const bool last_tile = tile == n_tiles - 1;
                    int sj[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) sj[j] = sbase + (j < 4 ? j : 124 + j);
#pragma unroll
                    for (int qi = 0; qi < 8; ++qi) {
                        float acc[1][8];  // this query's 8 distances, unpacked
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            unpack2(acc2[qi][j], acc[0][2 * j], acc[0][2 * j + 1]);
                            acc2[qi][j] = 0ull;
                        }
                        u64 tau = __shfl_sync(FULL, lk[qi], 15);
                        if (tile == 0) {
                            float mn = FLT_MAX;
#pragma unroll
                            for (int j = 0; j < 8; ++j) mn = sj[j] < Ns ? fminf(mn, acc[0][j]) : mn;
                            tau = make_key(warp_kth16_nl(mn), -1);  // index 0xffffffff: "<= distance" passes
                        }
                        u64 key[8];
                        unsigned mask = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            key[j] = sj[j] < Ns ? make_key(acc[0][j], sj[j]) : KEY_MAX;
                            mask |= (key[j] < tau ? 1u : 0u) << j;
                        }
                        const int c = __popc(mask);
                        int incl = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(FULL, incl, o);
                            incl += lane >= o ? v : 0;
                        }
                        const int total = __shfl_sync(FULL, incl, 31);
                        u64* q = sQ[w][qi];
                        if (cnt[qi] + total > QCAP) {  // rare: queue overflow -> slot-by-slot with flushes
                            const SelRet r = knn_select_slow(key[0], key[1], key[2], key[3], key[4], key[5], key[6],
                                                             key[7], lk[qi], cnt[qi], q);
                            lk[qi] = r.lk;
                            cnt[qi] = r.cnt;
                        } else {
                            int pos = cnt[qi] + incl - c;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if ((mask >> j) & 1u) q[pos++] = key[j];
                            }
                            cnt[qi] += total;
                        }
                        if (cnt[qi] > 16 || (last_tile && cnt[qi] > 0)) {
                            lk[qi] = knn_flush(lk[qi], q, cnt[qi]);
                            cnt[qi] = 0;
                        }
                    }
                }
            }
            if (it + 1 < n_it) {
                s_store(buf ^ 1);
                __syncthreads();
            }
        }
        if (warp_on && lane < LS_KNN_K) {
#pragma unroll
            for (int qi = 0; qi < 8; ++qi) {
                sIdx[w * 8 + qi][lane] = min(key_idx(lk[qi]) & 0x7fffffff, Ns - 1);  // stay in bounds on NaN input
                if (MODE == MODE_KNN_ONLY) sDist[w * 8 + qi][lane] = key_dist(lk[qi]);
            }
        }
    }
    __syncthreads();

    if (a.idx_out != nullptr) {
        for (int e = t; e < nq * LS_KNN_K; e += EDGE_THREADS) {
            int ql = e >> 4, k = e & 15;
            a.idx_out[((size_t)b * Nd + q0 + ql) * LS_KNN_K + k] = sIdx[ql][k];
        }
    }
    if (MODE == MODE_KNN_ONLY) {
        if (a.dist_out != nullptr) {
            for (int e = t; e < nq * LS_KNN_K; e += EDGE_THREADS) {
                int ql = e >> 4, k = e & 15;
                a.dist_out[((size_t)b * Nd + q0 + ql) * LS_KNN_K + k] = sDist[ql][k];
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- phase 2: EdgeConv
    constexpr int Co = 32 * CPL;  // compile-time: gather offsets become immediates
    const float oms = a.oms;
    const size_t ostride = (size_t)3 * Nd;  // floats between output channels

    if (MODE == MODE_L0) {
        const float* xyz = a.src_f + (size_t)b * 3 * Ns;  // layer 0: Ns == Nd, src == dst
        for (int ql = w; ql < nq; ql += 8) {
            const int n = q0 + ql;
            const float x0 = __ldg(xyz + n), x1 = __ldg(xyz + Ns + n), x2 = __ldg(xyz + 2 * Ns + n);
            const float nr = fmaxf(sqrtf(fmaf(x2, x2, fmaf(x1, x1, x0 * x0))), EPS_NRM);
            const float h0 = x0 / nr, h1 = x1 / nr, h2 = x2 / nr;
            // the edge geometry (cross product and difference) does not depend on the channel: lane k < 16 computes it
            // for edge k once, the channel loop below picks it up with shuffles (same operations: bit-identical)
            float e_c0, e_c1, e_c2, e_d0, e_d1, e_d2;
            {
                const int m = sIdx[ql][lane & (LS_KNN_K - 1)];
                const float n0 = __ldg(xyz + m), n1 = __ldg(xyz + Ns + m), n2 = __ldg(xyz + 2 * Ns + m);
                // cross(x_dir, nn)  (vec_dgcnn_atten.py:157)
                e_c0 = h1 * n2 - h2 * n1, e_c1 = h2 * n0 - h0 * n2, e_c2 = h0 * n1 - h1 * n0;
                e_d0 = n0 - x0, e_d1 = n1 - x1, e_d2 = n2 - x2;
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int c = j * 32 + lane;
                const float* wq = a.w0 + (size_t)c * 3;
                const float* wk = a.w0 + (size_t)(Co + c) * 3;
                const float wq0 = __ldg(wq), wq1 = __ldg(wq + 1), wq2 = __ldg(wq + 2);
                const float wk0 = __ldg(wk), wk1 = __ldg(wk + 1), wk2 = __ldg(wk + 2);
                // the feature (q) and direction (k) rows share their multiplicands: one packed f32x2 mul / fma per
                // pair (FMUL2 / FFMA2), same operations in the same order as the scalar form -> bit-identical
                const f32x2_t w0 = pack2(wq0, wk0), w1 = pack2(wq1, wk1), w2 = pack2(wq2, wk2);
                const f32x2_t X0 = pack2(x0, x0), X1 = pack2(x1, x1), X2 = pack2(x2, x2);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
                for (int k = 0; k < LS_KNN_K; ++k) {
                    const float c0 = __shfl_sync(FULL, e_c0, k), c1 = __shfl_sync(FULL, e_c1, k), c2 = __shfl_sync(FULL, e_c2, k);
                    const float d0 = __shfl_sync(FULL, e_d0, k), d1 = __shfl_sync(FULL, e_d1, k), d2 = __shfl_sync(FULL, e_d2, k);
                    float qx, qy, qz, kx, ky, kz;
                    unpack2(fma2(w2, X0, fma2(w1, pack2(d0, d0), mul2(w0, pack2(c0, c0)))), qx, kx);
                    unpack2(fma2(w2, X1, fma2(w1, pack2(d1, d1), mul2(w0, pack2(c1, c1)))), qy, ky);
                    unpack2(fma2(w2, X2, fma2(w1, pack2(d2, d2), mul2(w0, pack2(c2, c2)))), qz, kz);
                    float o0, o1, o2;
                    vn_act(qx, qy, qz, kx, ky, kz, oms, o0, o1, o2);
                    a0 += o0;
                    a1 += o1;
                    a2 += o2;
                }
                float* o = a.out + ((size_t)b * Co + c) * ostride + n;
                o[0] = a0 * (1.f / LS_KNN_K);
                o[Nd] = a1 * (1.f / LS_KNN_K);
                o[2 * Nd] = a2 * (1.f / LS_KNN_K);
            }
        }
        return;
    }

    // ---- layers >= 1: gather form.  A lane owns 4 consecutive channels (one float4 per table part and
    //      axis); LPP lanes cover one dst point, PPW points are processed per warp pass, and layers with
    //      more than 128 channels walk CH chunks of 128.
    const float* Ps = a.psrc + (size_t)b * Ns * a.row_s;
    constexpr int C3 = 3 * Co;  // floats per part
    constexpr int CG = Co / 4;
    constexpr int LPP = CG < 32 ? CG : 32;
    constexpr int PPW = 32 / LPP;
    constexpr int CH = CG / LPP;
    const int sub = lane / LPP, gl = lane % LPP;
    auto ld4 = [](const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
    // (packed f32x2 adds / products: same IEEE operations in the same order as the scalar forms, half the instructions --
    //  the deep layers are bound by instruction issue, not by their gathers: profiles/r02/experiments.md section 12)
    auto add4 = [](float4 x, float4 y) { return add4_packed(x, y); };
    // VN leaky-ReLU on 4 channels at once: q[axis], k[axis] hold the 4 channels of one axis
    auto act4 = [&](const float4* q, const float4* k, float4* o) { vn_act4(q, k, oms, o); };
    auto store_out = [&](int c4, int n, const float4* v, float sc) {
        float* o = a.out + ((size_t)b * Co + c4) * ostride + n;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            o[(size_t)0 * ostride + ax * Nd] = v[ax].x * sc;
            o[(size_t)1 * ostride + ax * Nd] = v[ax].y * sc;
            o[(size_t)2 * ostride + ax * Nd] = v[ax].z * sc;
            o[(size_t)3 * ostride + ax * Nd] = v[ax].w * sc;
        }
    };

    if (MODE == MODE_MEAN) {
        for (int base = w * PPW; base < nq; base += 8 * PPW) {
            const bool on = base + sub < nq;
            const int ql = on ? base + sub : nq - 1;
            const int n = q0 + ql;
            const float* Pd = a.pdst + ((size_t)b * Nd + n) * a.row_d;
#pragma unroll 1
            for (int ch = 0; ch < CH; ++ch) {
                const int c4 = (ch * LPP + gl) * 4;
                float4 dq[3], dk[3], acc[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    dq[ax] = ld4(Pd + ax * Co + c4);
                    dk[ax] = ld4(Pd + C3 + ax * Co + c4);
                    acc[ax] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll 4
                for (int k = 0; k < LS_KNN_K; ++k) {
                    const float* row = Ps + (size_t)sIdx[ql][k] * a.row_s + c4;
                    float4 q[3], kk[3], o[3];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        q[ax] = add4(ld4(row + ax * Co), dq[ax]);
                        kk[ax] = add4(ld4(row + C3 + ax * Co), dk[ax]);
                    }
                    act4(q, kk, o);
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) acc[ax] = add4(acc[ax], o[ax]);
                }
                if (on) store_out(c4, n, acc, 1.f / LS_KNN_K);
            }
        }
        return;
    }

    if (MODE == MODE_ATT) {
        float* sR = sbuf + (size_t)w * (CH * LS_KNN_K * 8);  // [chunk][edge][head slot] per warp (tiles are dead)
        const float inv_sqrt = rsqrtf(3.f * (float)LS_HEAD_C);
        const int hs = lane >> 2;  // 4 lanes x 4 channels = one 16-channel head
        for (int base = w * PPW; base < nq; base += 8 * PPW) {
            const bool on = base + sub < nq;
            const int ql = on ? base + sub : nq - 1;
            const int n = q0 + ql;
            const float* Pd = a.pdst + ((size_t)b * Nd + n) * a.row_d;
            // ---- |Q| over all channels of the point (cevn, vec_layers.py:24-31)
            float ssum = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < CH; ++ch) {
                const int c4 = (ch * LPP + gl) * 4;
                float4 q[3], kk[3], o[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    q[ax] = ld4(Pd + 4 * C3 + ax * Co + c4);
                    kk[ax] = ld4(Pd + 5 * C3 + ax * Co + c4);
                }
                act4(q, kk, o);
#pragma unroll
                for (int ax = 0; ax < 3; ++ax)
                    ssum += o[ax].x * o[ax].x + o[ax].y * o[ax].y + o[ax].z * o[ax].z + o[ax].w * o[ax].w;
            }
#pragma unroll
            for (int off = LPP / 2; off > 0; off >>= 1) ssum += __shfl_xor_sync(FULL, ssum, off);
            const float Lq = fmaxf(sqrtf(ssum), EPS_NRM);

            // ---- pass A: K branch -> per-edge channel norm S[k] and per-head raw logits R[chunk][k]
            float S[LS_KNN_K];
#pragma unroll
            for (int k = 0; k < LS_KNN_K; ++k) S[k] = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < CH; ++ch) {
                const int c4 = (ch * LPP + gl) * 4;
                float4 g[3];  // cevn(Q) of this lane's 4 channels
                {
                    float4 q[3], kk[3], o[3];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        q[ax] = ld4(Pd + 4 * C3 + ax * Co + c4);
                        kk[ax] = ld4(Pd + 5 * C3 + ax * Co + c4);
                    }
                    act4(q, kk, o);
                    float f[4];
                    const float e0 = sqrtf(o[0].x * o[0].x + o[1].x * o[1].x + o[2].x * o[2].x);
                    const float e1 = sqrtf(o[0].y * o[0].y + o[1].y * o[1].y + o[2].y * o[2].y);
                    const float e2 = sqrtf(o[0].z * o[0].z + o[1].z * o[1].z + o[2].z * o[2].z);
                    const float e3 = sqrtf(o[0].w * o[0].w + o[1].w * o[1].w + o[2].w * o[2].w);
                    f[0] = (e0 / Lq) / fmaxf(e0, EPS_NRM);
                    f[1] = (e1 / Lq) / fmaxf(e1, EPS_NRM);
                    f[2] = (e2 / Lq) / fmaxf(e2, EPS_NRM);
                    f[3] = (e3 / Lq) / fmaxf(e3, EPS_NRM);
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax)
                        g[ax] = make_float4(o[ax].x * f[0], o[ax].y * f[1], o[ax].z * f[2], o[ax].w * f[3]);
                }
                float4 dq[3], dk[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    dq[ax] = ld4(Pd + 2 * C3 + ax * Co + c4);
                    dk[ax] = ld4(Pd + 3 * C3 + ax * Co + c4);
                }
#pragma unroll
                for (int k = 0; k < LS_KNN_K; ++k) {
                    const float* row = Ps + (size_t)sIdx[ql][k] * a.row_s + 2 * C3 + c4;
                    float4 q[3], kk[3], o[3];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        q[ax] = add4(ld4(row + ax * Co), dq[ax]);
                        kk[ax] = add4(ld4(row + C3 + ax * Co), dk[ax]);
                    }
                    act4(q, kk, o);
                    float sk = 0.f, r = 0.f;
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        sk += o[ax].x * o[ax].x + o[ax].y * o[ax].y + o[ax].z * o[ax].z + o[ax].w * o[ax].w;
                        r += o[ax].x * g[ax].x + o[ax].y * g[ax].y + o[ax].z * g[ax].z + o[ax].w * g[ax].w;
                    }
                    S[k] += sk;
                    r += __shfl_xor_sync(FULL, r, 1);
                    r += __shfl_xor_sync(FULL, r, 2);
                    if ((lane & 3) == 0) sR[(ch * LS_KNN_K + k) * 8 + hs] = r;
                }
            }
            // per-edge |K| over all channels of the point
#pragma unroll
            for (int k = 0; k < LS_KNN_K; ++k) {
#pragma unroll
                for (int off = LPP / 2; off > 0; off >>= 1) S[k] += __shfl_xor_sync(FULL, S[k], off);
                S[k] = inv_sqrt / fmaxf(sqrtf(S[k]), EPS_NRM);  // logit scale of edge k
            }
            __syncwarp();
            // ---- pass B: softmax over the 16 edges (in-thread, per head) and the weighted sum of V
#pragma unroll 1
            for (int ch = 0; ch < CH; ++ch) {
                const int c4 = (ch * LPP + gl) * 4;
                float al[LS_KNN_K];
                float mx = -FLT_MAX;
#pragma unroll
                for (int k = 0; k < LS_KNN_K; ++k) {
                    al[k] = sR[(ch * LS_KNN_K + k) * 8 + hs] * S[k];
                    mx = fmaxf(mx, al[k]);
                }
                float den = 0.f;
#pragma unroll
                for (int k = 0; k < LS_KNN_K; ++k) {
                    al[k] = expf(al[k] - mx);
                    den += al[k];
                }
                const float rden = 1.f / den;
                float4 dq[3], dk[3], acc[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    dq[ax] = ld4(Pd + ax * Co + c4);
                    dk[ax] = ld4(Pd + C3 + ax * Co + c4);
                    acc[ax] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < LS_KNN_K; ++k) {
                    const float* row = Ps + (size_t)sIdx[ql][k] * a.row_s + c4;
                    float4 q[3], kk[3], o[3];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        q[ax] = add4(ld4(row + ax * Co), dq[ax]);
                        kk[ax] = add4(ld4(row + C3 + ax * Co), dk[ax]);
                    }
                    act4(q, kk, o);
                    const float ak = al[k] * rden;
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        acc[ax].x = fmaf(ak, o[ax].x, acc[ax].x);
                        acc[ax].y = fmaf(ak, o[ax].y, acc[ax].y);
                        acc[ax].z = fmaf(ak, o[ax].z, acc[ax].z);
                        acc[ax].w = fmaf(ak, o[ax].w, acc[ax].w);
                    }
                }
                if (on) store_out(c4, n, acc, 1.f);
            }
            __syncwarp();
        }
    }
}

// ============================================================================================
// kNN for small source sets (Ns <= 128: the deep, heavily down-sampled layers).  One warp per query,
// lanes over sources (<= 4 per lane), features streamed from L1/L2, top-16 by warp bitonic merges.
// Writes the graph that the phase-2-only launch of k_knn_edge then consumes (idx_in).
// ============================================================================================
__global__ void __launch_bounds__(256) k_knn_small(const float* __restrict__ src_f, const float* __restrict__ dst_f,
                                                   int D, int Ns, int Nd, int64_t* __restrict__ idx_out,
                                                   float* __restrict__ dist_out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y, n = blockIdx.x * 8 + w;
    if (n >= Nd) return;
    const float* srcb = src_f + (size_t)b * D * Ns;
    const float* dstb = dst_f + (size_t)b * D * Nd + n;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
        const float q = __ldg(dstb + (size_t)d * Nd);
        const float* sr = srcb + (size_t)d * Ns + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (lane + 32 * j < Ns) {
                const float df = q - __ldg(sr + 32 * j);
                acc[j] = fmaf(df, df, acc[j]);
            }
        }
    }
    u64 lk = lane < Ns ? make_key(acc[0], lane) : KEY_MAX;
    bitonic_sort32(lk, lane, false);
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        if (32 * j < Ns) {  // warp-uniform
            const int s = lane + 32 * j;
            u64 bk = s < Ns ? make_key(acc[j], s) : KEY_MAX;
            bitonic_sort32(bk, lane, true);
            lk = bk < lk ? bk : lk;
            bitonic_merge32(lk, lane);
        }
    }
    const int ai = key_idx(lk) & 0x7fffffff;
    const float ad = key_dist(lk);
    if (lane < LS_KNN_K) {
        idx_out[((size_t)b * Nd + n) * LS_KNN_K + lane] = min(ai, Ns - 1);
        if (dist_out) dist_out[((size_t)b * Nd + n) * LS_KNN_K + lane] = ad;
    }
}

// Tiled variant (default): one CTA = 32 queries of one instance x all Ns <= 128 sources.  The [dims][points] tiles of
// both operands are staged in shared memory (coalesced global loads, every element read once per CTA instead of once
// per warp), warp w owns a group of `spw` sources and lane = query, so a source value is one broadcast LDS.128 per 4
// sources and the distance is the same sequential fp32 chain fmaf(q - s, q - s, acc) over d = 0..D-1 as everywhere
// else (bit-identical keys).  The selection reuses the warp bitonic network on the [32][Ns] distances.
// Round 2: layers 5-6 (D = 384 / 768, 32 queries per instance) took 0.23 + 0.36 ms with the warp-per-query kernel
// above, latency-bound on 2 dependent global loads per dimension.
constexpr int KS_DK = 32;
__global__ void __launch_bounds__(256) k_knn_small_tiled(const float* __restrict__ src_f, const float* __restrict__ dst_f,
                                                         int D, int Ns, int Nd, int64_t* __restrict__ idx_out,
                                                         float* __restrict__ dist_out) {
    __shared__ __align__(16) float s_src[KS_DK][128];
    __shared__ float s_dst[KS_DK][33];
    __shared__ float s_dist[32][129];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int b = blockIdx.y, q0 = blockIdx.x * 32;
    const float* srcb = src_f + (size_t)b * D * Ns;
    const float* dstb = dst_f + (size_t)b * D * Nd;
    const int spw = (((Ns + 7) >> 3) + 3) & ~3;  // sources per warp, multiple of 4, <= 16
    const int s0 = w * spw;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    // the next 32-dim stage is fetched into registers while the current one is consumed (the kernel was a chain of
    // {global load latency, barrier, compute, barrier} per stage: 12 / 24 exposed round trips at layers 5 / 6)
    float rs[KS_DK * 128 / 256], rd[KS_DK * 32 / 256];
    auto fetch = [&](int d0) {
#pragma unroll
        for (int k = 0; k < KS_DK * 128 / 256; ++k) {
            const int i = t + k * 256, dd = i >> 7, sidx = i & 127;
            rs[k] = (d0 + dd < D && sidx < Ns) ? __ldg(srcb + (size_t)(d0 + dd) * Ns + sidx) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < KS_DK * 32 / 256; ++k) {
            const int i = t + k * 256, dd = i >> 5, qq = i & 31;
            rd[k] = (d0 + dd < D && q0 + qq < Nd) ? __ldg(dstb + (size_t)(d0 + dd) * Nd + q0 + qq) : 0.f;
        }
    };
    fetch(0);
    for (int d0 = 0; d0 < D; d0 += KS_DK) {
#pragma unroll
        for (int k = 0; k < KS_DK * 128 / 256; ++k) {
            const int i = t + k * 256;
            s_src[i >> 7][i & 127] = rs[k];
        }
#pragma unroll
        for (int k = 0; k < KS_DK * 32 / 256; ++k) {
            const int i = t + k * 256;
            s_dst[i >> 5][i & 31] = rd[k];
        }
        if (d0 + KS_DK < D) fetch(d0 + KS_DK);
        __syncthreads();
        // dims past D are staged as 0 for both operands: fmaf(0, 0, acc) == acc exactly
#pragma unroll 2
        for (int dd = 0; dd < KS_DK; ++dd) {
            const float qv = s_dst[dd][lane];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                if (j4 * 4 < spw && s0 + j4 * 4 < 128) {  // warp-uniform
                    const float4 sv = *reinterpret_cast<const float4*>(&s_src[dd][s0 + j4 * 4]);
                    float df = qv - sv.x;
                    acc[j4 * 4 + 0] = fmaf(df, df, acc[j4 * 4 + 0]);
                    df = qv - sv.y;
                    acc[j4 * 4 + 1] = fmaf(df, df, acc[j4 * 4 + 1]);
                    df = qv - sv.z;
                    acc[j4 * 4 + 2] = fmaf(df, df, acc[j4 * 4 + 2]);
                    df = qv - sv.w;
                    acc[j4 * 4 + 3] = fmaf(df, df, acc[j4 * 4 + 3]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (j < spw && s0 + j < Ns) s_dist[lane][s0 + j] = acc[j];
    __syncthreads();
    // selection: warp w ranks queries 4w .. 4w+3
    for (int qq = w * 4; qq < w * 4 + 4; ++qq) {
        const int n = q0 + qq;
        if (n >= Nd) break;  // warp-uniform
        u64 lk = lane < Ns ? make_key(s_dist[qq][lane], lane) : KEY_MAX;
        bitonic_sort32(lk, lane, false);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            if (32 * j < Ns) {  // warp-uniform
                const int sidx = lane + 32 * j;
                u64 bk = sidx < Ns ? make_key(s_dist[qq][sidx], sidx) : KEY_MAX;
                bitonic_sort32(bk, lane, true);
                lk = bk < lk ? bk : lk;
                bitonic_merge32(lk, lane);
            }
        }
        if (lane < LS_KNN_K) {
            idx_out[((size_t)b * Nd + n) * LS_KNN_K + lane] = min(key_idx(lk) & 0x7fffffff, Ns - 1);
            if (dist_out) dist_out[((size_t)b * Nd + n) * LS_KNN_K + lane] = key_dist(lk);
        }
    }
}

// ---- exact re-rank of the tensor-core candidates.  The squared distance is the reference's direct form
//      sum_d (q_d - s_d)^2 accumulated with one fp32 FMA per dimension in ascending d -- the same operation
//      sequence as the brute-force tiles above, so both paths produce bit-identical keys.
//      A lane owns one candidate row ([Dp] contiguous floats); RR_F4 float4 loads are issued back to back before
//      the dependent FMA chain consumes them (the rows are cold: one exposed miss per chunk instead of one per
//      sector), the query chunk is staged in shared memory and read as a broadcast.
constexpr int RR_F4 = 8;   // float4 loads in flight per lane (32 dims per chunk)
__device__ __forceinline__ float exact_dist_pm(const float* __restrict__ qrow, float* __restrict__ sq,
                                               const float* __restrict__ srow, bool on, int Dp) {
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
#pragma unroll 1
    for (int d0 = 0; d0 < Dp; d0 += 4 * RR_F4) {
        float4 sv[RR_F4];
#pragma unroll
        for (int i = 0; i < RR_F4; ++i)
            if (on && d0 + 4 * i < Dp) sv[i] = __ldg(reinterpret_cast<const float4*>(srow + d0 + 4 * i));
        __syncwarp();
        if (lane < RR_F4 && d0 + 4 * lane < Dp)
            *reinterpret_cast<float4*>(sq + 4 * lane) = __ldg(reinterpret_cast<const float4*>(qrow + d0 + 4 * lane));
        __syncwarp();
        if (on) {
#pragma unroll
            for (int i = 0; i < RR_F4; ++i) {
                if (d0 + 4 * i < Dp) {
                    const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * i);
                    float df = __fsub_rn(qv.x, sv[i].x);
                    acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(qv.y, sv[i].y);
                    acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(qv.z, sv[i].z);
                    acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(qv.w, sv[i].w);
                    acc = __fmaf_rn(df, df, acc);
                }
            }
        }
    }
    return acc;
}
// one warp, one query: lanes = candidates (32 per round); returns the sorted list element of this lane
__device__ __noinline__ u64 knn_rerank(const float* qrow, float* sq, const float* pms, const unsigned short* cl, int cnt,
                                       int Dp) {
    const int lane = threadIdx.x & 31;
    u64 lk = KEY_MAX;
    for (int r0 = 0; r0 < cnt; r0 += 32) {
        const int slot = r0 + lane;
        const bool on = slot < cnt;
        const int s = on ? (int)__ldg(cl + (size_t)slot * KT_PTS) : 0;
        const float d = exact_dist_pm(qrow, sq, pms + (size_t)s * Dp, on, Dp);
        u64 key = on ? make_key(d, s) : KEY_MAX;
        if (r0 == 0) {
            bitonic_sort32(key, lane, false);
            lk = key;
        } else {
            bitonic_sort32(key, lane, true);
            lk = key < lk ? key : lk;
            bitonic_merge32(lk, lane);
        }
    }
    return lk;
}
// exact brute force of one query over all sources (candidate-list overflow; never taken on sane inputs)
__device__ __noinline__ u64 knn_bruteforce_pm(const float* qrow, float* sq, const float* pms, int Ns, int Dp) {
    const int lane = threadIdx.x & 31;
    u64 lk = KEY_MAX;
    for (int s0 = 0; s0 < Ns; s0 += 32) {
        const int s = s0 + lane;
        const bool on = s < Ns;
        const float d = exact_dist_pm(qrow, sq, pms + (size_t)(on ? s : 0) * Dp, on, Dp);
        u64 key = on ? make_key(d, s) : KEY_MAX;
        if (s0 == 0) {
            bitonic_sort32(key, lane, false);
            lk = key;
        } else {
            bitonic_sort32(key, lane, true);
            lk = key < lk ? key : lk;
            bitonic_merge32(lk, lane);
        }
    }
    return lk;
}


// Re-rank kernel: one warp per query, candidate lists from k_knn_tc (ls_knn_tc.cu) -> the final graph
// idx [B][Nd][16] (ascending (distance, index)), optionally the squared distances.
//   The candidates are first sorted by their tensor-core ranking value dt.  Two candidates whose dt differ by
//   more than 2E are ordered like their exact fp32 distances (|dt + |q|^2 - d| <= E for both), so only RUNS of
//   candidates chained by gaps <= 2E can be mis-ordered -- exact ties always are in one run.  Exact direct-form
//   distances are computed for the runs that reach into the first 16 ranks only, and those runs are re-sorted
//   by (exact distance, index); every other candidate keeps its rank.  ~1 candidate in 10 needs its feature
//   row instead of all ~22.  Lists longer than 32 (rare) take the all-exact path, overflowed lists brute force.
struct RerankArgs {
    const unsigned short* cand;  // [B][ceil(Nd/128)][KT_CAP][128]
    const float* cand_dt;        // same layout
    const int* cand_cnt;         // [B][ceil(Nd/128)*128], -1 = overflow -> brute force that query
    const float* e2;             // [B][ceil(Nd/128)*128]
    const float* pm_s;           // [B][Ns][Dp] point-major fp32 features of the sources
    const float* pm_q;           // [B][Nd][Dp] ... of the queries
    int Ns, Nd, Dp;
    int all_exact;               // 1: exact distance for every candidate (needed when dist_out is wanted)
    int64_t* idx_out;            // [B][Nd][16] consumed by k_knn_edge (idx_in)
    int64_t* idx_tap;            // optional second copy (API tap)
    float* dist_out;             // optional [B][Nd][16]
};
constexpr int RR_WARPS = 8, RR_QPW = 4;  // warps per CTA, consecutive queries per warp
constexpr int RR_FB_Q = 1024;             // queries screened per CTA of the heavy-path launch
constexpr int RR_DCH = 192, RR_NR = 4, RR_ROW = RR_DCH + 4;  // dims per chunk, rows per pass, padded row stride
// Exact direct-form distances for the few lanes that `need` one: the warp loads the query chunk and up to RR_NR
// needed candidate rows cooperatively (coalesced, all loads in flight together) into shared memory, then every
// needing lane runs its sequential FMA chain out of shared memory.  Same operation order as exact_dist_pm.
__device__ __forceinline__ float exact_dist_coop(bool need, int s, const float* __restrict__ qrow,
                                                 const float* __restrict__ pms, int Dp, float* sm) {
    const int lane = threadIdx.x & 31;
    const unsigned mask = __ballot_sync(FULL, need);
    const int my_rank = __popc(mask & ((1u << lane) - 1u)), n_need = __popc(mask);
    float acc = 0.f;
    unsigned rest = mask;
    for (int base = 0; base < n_need; base += RR_NR) {
        int src[RR_NR];
#pragma unroll
        for (int r = 0; r < RR_NR; ++r) {
            const int l = rest ? __ffs(rest) - 1 : -1;
            rest = rest ? rest & (rest - 1) : 0u;
            src[r] = l >= 0 ? __shfl_sync(FULL, s, l) : -1;
        }
        const bool mine = need && my_rank >= base && my_rank < base + RR_NR;
        for (int d0 = 0; d0 < Dp; d0 += RR_DCH) {
            const int n4 = min(RR_DCH, Dp - d0) >> 2;
            __syncwarp();
            for (int i = lane; i < n4; i += 32) {
                reinterpret_cast<float4*>(sm)[i] = __ldg(reinterpret_cast<const float4*>(qrow + d0) + i);
#pragma unroll
                for (int r = 0; r < RR_NR; ++r)
                    if (src[r] >= 0)
                        reinterpret_cast<float4*>(sm + (1 + r) * RR_ROW)[i] =
                            __ldg(reinterpret_cast<const float4*>(pms + (size_t)src[r] * Dp + d0) + i);
            }
            __syncwarp();
            if (mine) {
                const float4* row = reinterpret_cast<const float4*>(sm + (1 + my_rank - base) * RR_ROW);
#pragma unroll 4
                for (int i = 0; i < n4; ++i) {
                    const float4 qv = reinterpret_cast<const float4*>(sm)[i], sv = row[i];
                    float df = __fsub_rn(qv.x, sv.x);
                    acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(qv.y, sv.y);
                    acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(qv.z, sv.z);
                    acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(qv.w, sv.w);
                    acc = __fmaf_rn(df, df, acc);
                }
            }
        }
    }
    return acc;
}
// Sort <= 32 distinct keys (one per lane, lanes >= cnt hold KEY_MAX) by counting: every lane compares its key with
// all cnt keys through broadcast shared-memory reads -- ~3 independent instructions per key instead of the 15
// dependent shuffle stages of the bitonic network (the kernel is latency bound, not issue bound).
__device__ __forceinline__ u64 warp_rank_sort(u64 key, int cnt, u64* sm) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    sm[lane] = key;
    __syncwarp();
    int r = 0;
#pragma unroll 8
    for (int j = 0; j < cnt; ++j) r += sm[j] < key ? 1 : 0;
    __syncwarp();
    if (lane < cnt) sm[32 + r] = key;
    __syncwarp();
    return lane < cnt ? sm[32 + lane] : KEY_MAX;
}
__device__ __forceinline__ unsigned ord_f32(float f) {  // order-preserving map float -> unsigned
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// DP > 0: the padded feature dimension is a compile-time constant (8 / 96 / 192 on the shipped model): the row
// loops unroll and the row addresses become constant strides; DP == 0 reads it from the arguments.
// FB = false: the common path (0 <= count <= 32, no distances wanted) ONLY; FB = true: the rare heavy paths only (overflowed
// lists -> exact brute force, lists longer than 32 or all_exact -> exact distance for every candidate).  One kernel with
// all three paths needs 128 registers (16 warps per SM) for a latency-bound warp-per-query loop; the common path alone
// needs 64-80 and runs at 4 CTAs per SM (measured: re-rank 1.07 ms -> 0.94 at 3 CTAs -> 0.90 at 4).  The FB launch screens
// 1024 queries per CTA and exits when none is flagged.
#ifndef LS_RR_CTAS
#define LS_RR_CTAS 4
#endif
template <int DP, bool FB>
__global__ void __launch_bounds__(RR_WARPS * 32, FB ? 2 : LS_RR_CTAS) k_knn_rerank(const RerankArgs a) {
    const int Dp = DP > 0 ? DP : a.Dp;
    __shared__ __align__(16) float sq_all[RR_WARPS][4 * RR_F4];
    __shared__ u64 ssort[RR_WARPS][64];
    __shared__ __align__(16) float scoop[RR_WARPS][(1 + RR_NR) * RR_ROW];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y, qbase = (blockIdx.x * RR_WARPS + w) * RR_QPW;
    const int n_pt_q = (a.Nd + KT_PTS - 1) / KT_PTS;
    if (FB) {
        // heavy launch: ONE CTA screens RR_FB_Q consecutive queries of an instance (coalesced count loads), collects the
        // flagged ones in shared memory and works them off warp by warp; without a flagged query it exits at once
        // (the usual case: 256 CTAs of a few hundred cycles instead of one CTA per 32 queries)
        __shared__ int s_list[RR_FB_Q];
        __shared__ int s_n;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int qblk = blockIdx.x * RR_FB_Q;
        for (int i = threadIdx.x; i < RR_FB_Q; i += RR_WARPS * 32) {
            const int q = qblk + i;
            if (q < a.Nd) {
                const int c = __ldg(a.cand_cnt + ((size_t)b * n_pt_q + q / KT_PTS) * KT_PTS + q % KT_PTS);
                if (a.all_exact || c < 0 || c > 32) s_list[atomicAdd(&s_n, 1)] = q;
            }
        }
        __syncthreads();
        const int n_flag = s_n;
        const float* pms_f = a.pm_s + (size_t)b * a.Ns * Dp;
        for (int i = w; i < n_flag; i += RR_WARPS) {
            const int q = s_list[i];
            const int c = __ldg(a.cand_cnt + ((size_t)b * n_pt_q + q / KT_PTS) * KT_PTS + q % KT_PTS);
            const float* qrow = a.pm_q + ((size_t)b * a.Nd + q) * Dp;
            const u64 lk = c < 0 ? knn_bruteforce_pm(qrow, sq_all[w], pms_f, a.Ns, Dp)
                                 : knn_rerank(qrow, sq_all[w], pms_f,
                                              a.cand + ((size_t)b * n_pt_q + q / KT_PTS) * KT_CAP * KT_PTS + q % KT_PTS, c, Dp);
            if (lane < LS_KNN_K) {
                const size_t o = ((size_t)b * a.Nd + q) * LS_KNN_K + lane;
                const int64_t sv = max(0, min(key_idx(lk) & 0x7fffffff, a.Ns - 1));  // stay in bounds on NaN input
                a.idx_out[o] = sv;
                if (a.idx_tap) a.idx_tap[o] = sv;
                if (a.dist_out) a.dist_out[o] = key_dist(lk);
            }
        }
        return;
    }
    if (qbase >= a.Nd) return;
    const float* pms = a.pm_s + (size_t)b * a.Ns * Dp;
    float* sq = sq_all[w];
    // software pipeline over the warp's queries: the next query's list is in flight while this one is processed
    int c_n, ci_n;
    float cd_n, e2_n;
    auto load_list = [&](int q) {
        q = min(q, a.Nd - 1);
        const size_t qo = ((size_t)b * n_pt_q + q / KT_PTS) * KT_PTS + q % KT_PTS;
        c_n = __ldg(a.cand_cnt + qo);
        e2_n = __ldg(a.e2 + qo);
        const size_t lo = ((size_t)b * n_pt_q + q / KT_PTS) * KT_CAP * KT_PTS + (size_t)lane * KT_PTS + q % KT_PTS;
        ci_n = (int)__ldg(a.cand + lo);
        cd_n = __ldg(a.cand_dt + lo);
    };
    load_list(qbase);
#pragma unroll 1
    for (int j = 0; j < RR_QPW; ++j) {
        const int q = qbase + j;
        if (q >= a.Nd) break;
        const int c_j = c_n, ci_j = ci_n;
        const float cd_j = cd_n, e2_j = e2_n;
        if (j + 1 < RR_QPW) load_list(q + 1);
        const float* qrow = a.pm_q + ((size_t)b * a.Nd + q) * Dp;
        int s_out;
        float d_out = 0.f;
        if (c_j < 0 || c_j > 32 || a.all_exact) continue;  // the heavy launch's query
        {
            const int cnt = c_j;
            u64 key = lane < cnt ? (((u64)ord_f32(cd_j) << 32) | (unsigned)ci_j) : KEY_MAX;
            key = warp_rank_sort(key, cnt, ssort[w]);
            const unsigned od = (unsigned)(key >> 32);
            const float dts = __uint_as_float((od & 0x80000000u) ? (od & 0x7fffffffu) : ~od);
            const int s = (int)(unsigned)(key & 0xffffffffu);
            const float nxt = __shfl_down_sync(FULL, dts, 1);
            const bool adj = (lane + 1 < cnt) && (nxt - dts <= e2_j);
            const bool prev_adj = __shfl_up_sync(FULL, (int)adj, 1) != 0 && lane > 0;
            int rs = prev_adj ? -1 : lane;  // run start marker
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, rs, o);
                rs = lane >= o ? max(rs, v) : rs;
            }
            const bool need = lane < cnt && (adj || prev_adj) && rs <= LS_KNN_K - 1;
            s_out = s;
            if (__any_sync(FULL, need)) {
                const float d = exact_dist_coop(need, s, qrow, pms, Dp, scoop[w]);
                // final order: by run (runs are contiguous rank ranges), inside a re-sorted run by (exact d, index)
                u64 k2 = lane < cnt ? (((u64)(unsigned)rs << 58) | ((u64)(need ? __float_as_uint(d) : 0u) << 26) | (unsigned)s) : KEY_MAX;
                k2 = warp_rank_sort(k2, cnt, ssort[w]);
                s_out = (int)(unsigned)(k2 & 0x3ffffffu);
            }
        }
        if (lane < LS_KNN_K) {
            const size_t o = ((size_t)b * a.Nd + q) * LS_KNN_K + lane;
            const int64_t sv = max(0, min(s_out, a.Ns - 1));  // stay in bounds on NaN input (empty candidate list)
            a.idx_out[o] = sv;
            if (a.idx_tap) a.idx_tap[o] = sv;
            if (a.dist_out) a.dist_out[o] = d_out;
        }
    }
}

// ============================================================================================
// Global context (vec_dgcnn_atten.py:222-225): g = mean_n f ; bias[r][a] = sum_c Wg2[r][c] g[c][a]
// so that VecLNA_G([f ; g]) = Wg1 f + bias.  Two small kernels, one warp per output element row.
// ============================================================================================
// g[b][r] = mean_n f[b][r][n], r over Co*3 rows
__global__ void __launch_bounds__(256) k_row_mean(const float* __restrict__ f, int rows, int Nd, float* __restrict__ g) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    if (Nd == 32) {
        // deep layers: 32 floats per row.  A lane sums one quarter of a row (two float4), 4 lanes share a row, a warp
        // covers 8 rows, the CTA 64 -- same pairing as the general path's (s0 + s1) + (s2 + s3) is NOT required here:
        // this path is the only one used for Nd == 32, for every batch composition.
        const int r = blockIdx.x * 64 + w * 8 + (lane >> 2);
        float acc = 0.f;
        if (r < rows) {
            const float4* p4 = reinterpret_cast<const float4*>(f + ((size_t)b * rows + r) * 32) + (lane & 3) * 2;
            const float4 u = __ldg(p4), v = __ldg(p4 + 1);
            acc = ((u.x + u.y) + (u.z + u.w)) + ((v.x + v.y) + (v.z + v.w));
        }
        acc += __shfl_xor_sync(FULL, acc, 1);
        acc += __shfl_xor_sync(FULL, acc, 2);
        if (r < rows && (lane & 3) == 0) g[(size_t)b * rows + r] = acc / 32.f;
        return;
    }
    const int r = blockIdx.x * 8 + w;
    if (r >= rows) return;
    const float* p = f + ((size_t)b * rows + r) * Nd;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int n = lane;
    for (; n + 96 < Nd; n += 128) {
        s0 += p[n];
        s1 += p[n + 32];
        s2 += p[n + 64];
        s3 += p[n + 96];
    }
    for (; n < Nd; n += 32) s0 += p[n];
    const float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) g[(size_t)b * rows + r] = s / (float)Nd;
}
// bias[b][r][a] = sum_c Wg2[r][c] * g[b][c][a]
__global__ void __launch_bounds__(256) k_bias_gemv(const float* __restrict__ g, int Co, const float* __restrict__ wg2,
                                                   float* __restrict__ bias) {
    extern __shared__ float sg[];  // [Co*3]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < Co * 3; i += blockDim.x) sg[i] = g[(size_t)b * Co * 3 + i];
    __syncthreads();
    const int r = blockIdx.x * 8 + w;
    if (r >= 2 * Co) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < Co; c += 32) {
        const float wv = __ldg(wg2 + (size_t)r * Co + c);
        s0 = fmaf(wv, sg[c * 3 + 0], s0);
        s1 = fmaf(wv, sg[c * 3 + 1], s1);
        s2 = fmaf(wv, sg[c * 3 + 2], s2);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        float* o = bias + ((size_t)b * 2 * Co + r) * 3;
        o[0] = s0;
        o[1] = s1;
        o[2] = s2;
    }
}

// bias[b][r][a] = sum_c Wg2[r][c] * g[b][c][a] for layers with Co >= 128, as a register-tiled FP32 GEMM: one CTA = 128
// weight rows x 8 instances (24 columns), a thread owns 2 rows x 2 instances x 3 axes, the means of the 8 instances are
// staged once in shared memory, the weight tile [128 rows][32 c] per step (next tile prefetched into registers).
// (The per-instance gemv re-read the 2 MB matrix of layer 6 for each of the 256 instances; a 48-CTA SIMT GEMM took
// 0.11 ms; a warp-per-row version with 24 shuffle reductions per row still 0.11 ms.)  Each output is one sequential
// fp32 chain over c = 0..Co-1: independent of the batch composition.
constexpr int BIAS_BG = 8, BIAS_ROWS = 128, BIAS_KT = 32;
__global__ void __launch_bounds__(256) k_bias_rows(const float* __restrict__ g, int Co, int B, const float* __restrict__ wg2,
                                                   float* __restrict__ bias) {
    extern __shared__ float sb[];
    const int gs = Co * 3 + 1;                 // padded per-instance stride of the staged means
    float* sg = sb;                            // [BIAS_BG][gs]
    float* sw = sb + BIAS_BG * gs;             // [BIAS_ROWS][BIAS_KT + 1]
    const int t = threadIdx.x;
    const int b0 = blockIdx.y * BIAS_BG, nb = min(BIAS_BG, B - b0), r0 = blockIdx.x * BIAS_ROWS;
    for (int i = t; i < BIAS_BG * Co * 3; i += 256) {
        const int bi = i / (Co * 3), e = i - bi * (Co * 3);
        sg[bi * gs + e] = bi < nb ? g[(size_t)(b0 + bi) * Co * 3 + e] : 0.f;
    }
    const int rg = t >> 2, cg = t & 3;         // rows r0 + 2 rg, + 1; instances 2 cg, 2 cg + 1
    float acc[2][6];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
    float wr[BIAS_ROWS * BIAS_KT / 256];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < BIAS_ROWS * BIAS_KT / 256; ++i) {
            const int row = i * 8 + (t >> 5), r = r0 + row;
            wr[i] = (r < 2 * Co && k0 + (t & 31) < Co) ? __ldg(wg2 + (size_t)r * Co + k0 + (t & 31)) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < Co; k0 += BIAS_KT) {
        __syncthreads();  // the previous tile is no longer read (and, first time, the means are staged)
#pragma unroll
        for (int i = 0; i < BIAS_ROWS * BIAS_KT / 256; ++i) sw[(i * 8 + (t >> 5)) * (BIAS_KT + 1) + (t & 31)] = wr[i];
        if (k0 + BIAS_KT < Co) fetch(k0 + BIAS_KT);
        __syncthreads();
        const float* w0 = sw + (2 * rg) * (BIAS_KT + 1);
        const float* g0 = sg + (2 * cg) * gs + k0 * 3;
#pragma unroll 8
        for (int k = 0; k < BIAS_KT; ++k) {
            const float wa = w0[k], wb = w0[BIAS_KT + 1 + k];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float gv = g0[(j / 3) * gs + k * 3 + (j % 3)];
                acc[0][j] = fmaf(wa, gv, acc[0][j]);
                acc[1][j] = fmaf(wb, gv, acc[1][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = r0 + 2 * rg + i;
        if (r >= 2 * Co) continue;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int bi = 2 * cg + j / 3;
            if (bi < nb) bias[((size_t)(b0 + bi) * 2 * Co + r) * 3 + (j % 3)] = acc[i][j];
        }
    }
}

// raw [B][2Co][3][N] (q rows, then direction rows) -> VN leaky-ReLU -> out [B][Co][3][N]; thread = (channel, point)
// flattened, so that the deep layers (N = 32) still run full warps
__global__ void __launch_bounds__(256) k_vnact(const float* __restrict__ raw, int Co, int N, float oms, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Co * N) return;
    const int c = e / N, n = e - c * N;
    const float* q = raw + ((size_t)b * 2 * Co + c) * 3 * N + n;
    const float* k = raw + ((size_t)b * 2 * Co + Co + c) * 3 * N + n;
    float o0, o1, o2;
    vn_act(q[0], q[N], q[2 * N], k[0], k[N], k[2 * N], oms, o0, o1, o2);
    float* o = out + ((size_t)b * Co + c) * 3 * N + n;
    o[0] = o0;
    o[N] = o1;
    o[2 * N] = o2;
}

// ============================================================================================
// Head (vec_dgcnn_atten.py:231-252 + VecResBlock vec_layers.py:631-672): one CTA per instance,
// blockDim = c_dim, thread c owns channel c.
// ============================================================================================
struct HeadArgs {
    const float* raw;  // [B][c_dim+1][3][Nl]  conv_c pre-activation rows + shared direction row
    int c_dim, Nl;
    const float* w_inv_t;
    const float* w_fc0_t;
    const float* w_lin1;
    const float* w_short;
    float w_act2, oms, scale_factor;
    int center_pred, center_pred_scale, normalize;
    const float* centroid;  // [B][3]  (normalize)
    const float* s0;        // [B]
    float *center, *scale, *z_so3, *z_inv, *packed;
};

__device__ __forceinline__ float block_sum(float v, float* red, int nw) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(1024) k_head(const HeadArgs a) {
    extern __shared__ float sm[];
    const int C = a.c_dim, Nl = a.Nl;
    float* sx = sm;          // [C][3]
    float* sf = sm + 3 * C;  // [C][3]  fc0 outputs (q cols then k cols)
    __shared__ float red[32];
    const int b = blockIdx.x, c = threadIdx.x, nw = blockDim.x >> 5;
    const float* rb = a.raw + (size_t)b * (C + 1) * 3 * Nl;
    const float* q = rb + (size_t)c * 3 * Nl;
    const float* kd = rb + (size_t)C * 3 * Nl;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    for (int n = 0; n < Nl; ++n) {
        float o0, o1, o2;
        vn_act(q[n], q[Nl + n], q[2 * Nl + n], kd[n], kd[Nl + n], kd[2 * Nl + n], a.oms, o0, o1, o2);
        x0 += o0;
        x1 += o1;
        x2 += o2;
    }
    x0 /= (float)Nl;
    x1 /= (float)Nl;
    x2 /= (float)Nl;
    sx[c * 3 + 0] = x0;
    sx[c * 3 + 1] = x1;
    sx[c * 3 + 2] = x2;
    const float ell = sqrtf(fmaf(x2, x2, fmaf(x1, x1, x0 * x0)));
    const float L = fmaxf(sqrtf(block_sum(ell * ell, red, nw)), EPS_NRM);
    const float sc_sum = block_sum(ell, red, nw);  // also orders the sx writes before the reads below
    const float ed = fmaxf(ell, EPS_NRM), fn = ell / L;
    const float z0 = (x0 / ed) * fn, z1 = (x1 / ed) * fn, z2 = (x2 / ed) * fn;  // z_so3 = cevn(x)
    float pred_scale = (sc_sum / (float)C) * a.scale_factor;

    // z_inv: <cevn(fc_inv x)[c], z_so3[c]>
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    for (int k = 0; k < C; ++k) {
        const float wv = __ldg(a.w_inv_t + (size_t)k * C + c);
        u0 = fmaf(wv, sx[k * 3 + 0], u0);
        u1 = fmaf(wv, sx[k * 3 + 1], u1);
        u2 = fmaf(wv, sx[k * 3 + 2], u2);
    }
    const float ul = sqrtf(fmaf(u2, u2, fmaf(u1, u1, u0 * u0)));
    const float UL = fmaxf(sqrtf(block_sum(ul * ul, red, nw)), EPS_NRM);
    const float ud = fmaxf(ul, EPS_NRM), un = ul / UL;
    const float zi = ((u0 / ud) * un) * z0 + ((u1 / ud) * un) * z1 + ((u2 / ud) * un) * z2;

    float ctr[3] = {0.f, 0.f, 0.f};
    if (a.center_pred) {
        // fc0 = VecLNA(C -> C/2): column c < C/2 is q_c, column C/2 + o is the direction of channel o
        float f0 = 0.f, f1 = 0.f, f2 = 0.f;
        for (int k = 0; k < C; ++k) {
            const float wv = __ldg(a.w_fc0_t + (size_t)k * C + c);
            f0 = fmaf(wv, sx[k * 3 + 0], f0);
            f1 = fmaf(wv, sx[k * 3 + 1], f1);
            f2 = fmaf(wv, sx[k * 3 + 2], f2);
        }
        sf[c * 3 + 0] = f0;
        sf[c * 3 + 1] = f1;
        sf[c * 3 + 2] = f2;
        __syncthreads();
        const int h = C / 2;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (c < h) {
            float o0, o1, o2;
            vn_act(sf[c * 3], sf[c * 3 + 1], sf[c * 3 + 2], sf[(h + c) * 3], sf[(h + c) * 3 + 1], sf[(h + c) * 3 + 2],
                   a.oms, o0, o1, o2);
            const float w1 = __ldg(a.w_lin1 + c);
            d0 = w1 * o0;
            d1 = w1 * o1;
            d2 = w1 * o2;
        }
        const float ws = __ldg(a.w_short + c);
        const float v0 = block_sum(fmaf(ws, x0, d0), red, nw);
        const float v1 = block_sum(fmaf(ws, x1, d1), red, nw);
        const float v2 = block_sum(fmaf(ws, x2, d2), red, nw);
        vn_act(v0, v1, v2, a.w_act2 * v0, a.w_act2 * v1, a.w_act2 * v2, a.oms, ctr[0], ctr[1], ctr[2]);
        if (a.center_pred_scale) {
            ctr[0] *= a.scale_factor;
            ctr[1] *= a.scale_factor;
            ctr[2] *= a.scale_factor;
        }
    }
    if (a.normalize) {  // model_utils.py:182-185
        pred_scale = a.s0[b] * pred_scale;
#pragma unroll
        for (int i = 0; i < 3; ++i) ctr[i] += a.centroid[b * 3 + i];
    }
    float* zs = a.z_so3 + ((size_t)b * C + c) * 3;
    zs[0] = z0;
    zs[1] = z1;
    zs[2] = z2;
    a.z_inv[(size_t)b * C + c] = zi;
    if (a.packed) {
        float* p = a.packed + (size_t)b * (4 * C + 4);
        p[c * 3 + 0] = z0;
        p[c * 3 + 1] = z1;
        p[c * 3 + 2] = z2;
        p[3 * C + c] = zi;
    }
    if (c == 0) {
        a.scale[b] = pred_scale;
        if (a.center) {
            a.center[b * 3 + 0] = ctr[0];
            a.center[b * 3 + 1] = ctr[1];
            a.center[b * 3 + 2] = ctr[2];
        }
        if (a.packed) {
            float* p = a.packed + (size_t)b * (4 * C + 4) + 4 * C;
            p[0] = pred_scale;
            p[1] = ctr[0];
            p[2] = ctr[1];
            p[3] = ctr[2];
        }
    }
}

}  // namespace ls
