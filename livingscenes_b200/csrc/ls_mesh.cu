// Mesh extraction from the SDF on the GPU: the MISE octree refinement that drives the decoder queries
// (mesh_extractor2.py:88-131 + utils/libmise/mise.pyx) and marching cubes on the completed grid
// (mesh_extractor2.py:158-181 + utils/libmcubes).  The reference keeps the octree in Cython/C++ STL containers on the
// host and ships every refinement level's points and values through numpy; here the state is four dense device
// arrays at the final resolution R = resolution0 << depth and nothing but one counter per level returns to the host.
//
//   state[(R+1)^3]  u8   0 = no grid point, 1 = grid point created, value unknown, 2 = value known   (GridPoint.known)
//   val  [(R+1)^3]  f32  occupancy logit of a known point (the decoder returns fp32; mise.pyx stores it as double)
//   level[R^3]      u8   level of the LEAF voxel that contains the unit cell (Voxel.level / is_leaf / children)
//   pos/neg[R^3]    u8   next_to_positive / next_to_negative of the leaf voxel, kept at the voxel's origin cell
//
// One refinement round = mise.update + mise.subdivide_voxels + mise.query:
//   k_mise_update    scatter the new values, state -> 2
//   k_mise_mark      every known point marks the (up to 8) leaf voxels around it:  value >= thr -> pos, <= thr -> neg
//   k_mise_subdivide a leaf voxel below the maximum depth with pos && neg gets level + 1 (its 8 children become the
//                    leaves) and creates the 27 lattice points of its 3x3x3 sub-grid that do not exist yet
//   k_mise_collect   append the points with state == 1 to the query list (order irrelevant), count to the host
// and mise.to_dense = k_mise_dense + three k_mise_fill passes (NaN filled from index - 1 along x, then y, then z).
//
// Marching cubes: classic 256-case cubes with one vertex per sign-changing grid edge.  The per-case triangle lists are
// GENERATED at first use (build_tables: trace the iso-polygon loops over the cube faces, fan-triangulate, orient the
// normal towards the corners with value <= iso) instead of being a pasted table; vertices are the reference's:
// linear interpolation in double on the edge, midpoint when both ends are equal, corner test "value <= iso".
#include <math.h>

#include <mutex>
#include <vector>

#include "ls_common.cuh"

namespace ls {
namespace {

// ------------------------------------------------------------------------------------------ MISE
__global__ void k_mise_init(unsigned char* __restrict__ state, unsigned char* __restrict__ level, int R, int step) {
    const long long n1 = (long long)(R + 1) * (R + 1) * (R + 1), n0 = (long long)R * R * R;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n1; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % (R + 1)), y = (int)((i / (R + 1)) % (R + 1)), x = (int)(i / ((long long)(R + 1) * (R + 1)));
        state[i] = (x % step == 0 && y % step == 0 && z % step == 0) ? 1 : 0;
        if (i < n0) level[i] = 0;
    }
}

// canonical query coordinates of the listed grid points: box * (p / R - 0.5)   (mesh_extractor2.py:113-116, fp32)
__global__ void k_mise_points(const int* __restrict__ list, int n, int R, float box, float* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = list[i];
    const int z = id % (R + 1), y = (id / (R + 1)) % (R + 1), x = id / ((R + 1) * (R + 1));
    q[i * 3 + 0] = box * ((float)x / (float)R - 0.5f);
    q[i * 3 + 1] = box * ((float)y / (float)R - 0.5f);
    q[i * 3 + 2] = box * ((float)z / (float)R - 0.5f);
}

__global__ void k_mise_update(const int* __restrict__ list, const float* __restrict__ v, int n, float scale,
                              float* __restrict__ val, unsigned char* __restrict__ state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    val[list[i]] = scale * v[i];
    state[list[i]] = 2;
}

__device__ __forceinline__ long long voxel_origin(int cx, int cy, int cz, int R, int depth, const unsigned char* level) {
    const long long c = ((long long)cx * R + cy) * R + cz;
    const int sh = depth - level[c];
    return ((long long)((cx >> sh) << sh) * R + ((cy >> sh) << sh)) * R + ((cz >> sh) << sh);
}

__global__ void k_mise_mark(const unsigned char* __restrict__ state, const float* __restrict__ val,
                            const unsigned char* __restrict__ level, int R, int depth, float thr,
                            unsigned char* __restrict__ pos, unsigned char* __restrict__ neg) {
    const long long n1 = (long long)(R + 1) * (R + 1) * (R + 1);
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n1 || state[i] != 2) return;
    const int z = (int)(i % (R + 1)), y = (int)((i / (R + 1)) % (R + 1)), x = (int)(i / ((long long)(R + 1) * (R + 1)));
    const float v = val[i];
    const bool p = v >= thr, q = v <= thr;
    for (int a = -1; a <= 0; ++a)
        for (int b = -1; b <= 0; ++b)
            for (int c = -1; c <= 0; ++c) {
                const int cx = x + a, cy = y + b, cz = z + c;
                if (cx < 0 || cy < 0 || cz < 0 || cx >= R || cy >= R || cz >= R) continue;
                const long long o = voxel_origin(cx, cy, cz, R, depth, level);
                if (p) pos[o] = 1;
                if (q) neg[o] = 1;
            }
}

__global__ void k_mise_subdivide(unsigned char* __restrict__ level, const unsigned char* __restrict__ pos,
                                 const unsigned char* __restrict__ neg, unsigned char* __restrict__ state, int R, int depth) {
    const long long n0 = (long long)R * R * R;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n0) return;
    const int cz = (int)(i % R), cy = (int)((i / R) % R), cx = (int)(i / ((long long)R * R));
    const int lv = level[i];
    if (lv >= depth) return;
    const int sh = depth - lv, size = 1 << sh;
    const int ox = (cx >> sh) << sh, oy = (cy >> sh) << sh, oz = (cz >> sh) << sh;
    const long long o = ((long long)ox * R + oy) * R + oz;
    if (!(pos[o] && neg[o])) return;
    if (i == o) {  // the origin cell creates the 27 lattice points (mise.pyx subdivide_voxel)
        const int h = size >> 1;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c) {
                    const long long p = ((long long)(ox + a * h) * (R + 1) + (oy + b * h)) * (R + 1) + (oz + c * h);
                    if (state[p] == 0) state[p] = 1;
                }
    }
    level[i] = (unsigned char)(lv + 1);  // every cell of the voxel moves to the child level (pos/neg are read at the OLD origin)
}

__global__ void k_mise_collect(const unsigned char* __restrict__ state, long long n1, int* __restrict__ list, int cap,
                               int* __restrict__ count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool want = i < n1 && state[i] == 1;
    const unsigned bal = __ballot_sync(FULL, want);
    if (bal == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(bal));
    base = __shfl_sync(FULL, base, 0);
    if (want) {
        const int slot = base + __popc(bal & ((1u << lane) - 1u));
        if (slot < cap) list[slot] = (int)i;
    }
}

__global__ void k_mise_dense(const unsigned char* __restrict__ state, const float* __restrict__ val, long long n1,
                             float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n1) out[i] = state[i] == 2 ? val[i] : __int_as_float(0x7fc00000);
}

// NaN <- value at index - 1 along `axis` (mise.pyx to_dense), one thread per grid line
__global__ void k_mise_fill(float* __restrict__ g, int R, int axis) {
    const int n = R + 1;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n * n) return;
    const int u = l / n, v = l % n;
    long long base, stride;
    if (axis == 0) { base = (long long)u * n + v; stride = (long long)n * n; }        // (j,k) fixed, walk i
    else if (axis == 1) { base = (long long)u * n * n + v; stride = n; }              // (i,k) fixed, walk j
    else { base = ((long long)u * n + v) * n; stride = 1; }                           // (i,j) fixed, walk k
    float prev = g[base];
    for (int t = 1; t < n; ++t) {
        float x = g[base + t * stride];
        if (isnan(x)) {
            x = prev;
            g[base + t * stride] = x;
        }
        prev = x;
    }
}

// ------------------------------------------------------------------------------------------ marching cubes tables
// corner m of a cell at (i,j,k): the reference's order (marchingcubes.h:59-63); edge e joins EDGE_A[e] and EDGE_B[e]
const int CORNER[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
const int EDGE_A[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
const int EDGE_B[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
// the 6 faces as corner cycles
const int FACE[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};

struct McTables {
    signed char tri[256][16];   // edge indices, 3 per triangle, -1 terminated (at most 5 triangles)
    unsigned char ntri[256];
};

int edge_between(int a, int b) {
    for (int e = 0; e < 12; ++e)
        if ((EDGE_A[e] == a && EDGE_B[e] == b) || (EDGE_A[e] == b && EDGE_B[e] == a)) return e;
    return -1;
}

// Per case: connect the sign-changing edges of every face pairwise (an ambiguous face -- diagonal corners alike --
// isolates its SET corners, a rule that only depends on the face, so neighbouring cells agree and the surface is
// closed), follow the resulting loops, fan-triangulate, orient the normal towards the set (value <= iso) corners.
void build_tables(McTables& T) {
    for (int cs = 0; cs < 256; ++cs) {
        int n_out = 0;
        for (int k = 0; k < 16; ++k) T.tri[cs][k] = -1;
        int link[12][2], nl[12];
        for (int e = 0; e < 12; ++e) nl[e] = 0;
        auto set = [&](int c) { return (cs >> c) & 1; };
        for (int f = 0; f < 6; ++f) {
            int ce[4], nce = 0, edges[4];
            for (int q = 0; q < 4; ++q) {
                const int a = FACE[f][q], b = FACE[f][(q + 1) & 3];
                edges[q] = edge_between(a, b);
                if (set(a) != set(b)) ce[nce++] = q;
            }
            auto connect = [&](int e0, int e1) {
                link[e0][nl[e0]++] = e1;
                link[e1][nl[e1]++] = e0;
            };
            if (nce == 2) {
                connect(edges[ce[0]], edges[ce[1]]);
            } else if (nce == 4) {
                // corners alternate; pair the two edges that meet at each SET corner
                for (int q = 0; q < 4; ++q)
                    if (set(FACE[f][q])) connect(edges[(q + 3) & 3], edges[q]);
            }
        }
        bool used[12] = {false};
        for (int e0 = 0; e0 < 12; ++e0) {
            if (nl[e0] != 2 || used[e0]) continue;
            int loop[12], n = 0, prev = -1, cur = e0;
            while (!used[cur]) {
                used[cur] = true;
                loop[n++] = cur;
                const int nxt = link[cur][0] != prev ? link[cur][0] : link[cur][1];
                prev = cur;
                cur = nxt;
            }
            // orientation: Newell normal of the loop (edge midpoints) against sum of (set corner - unset corner)
            double P[12][3], N[3] = {0, 0, 0}, D[3] = {0, 0, 0};
            for (int i = 0; i < n; ++i) {
                const int a = EDGE_A[loop[i]], b = EDGE_B[loop[i]];
                for (int d = 0; d < 3; ++d) {
                    P[i][d] = 0.5 * (CORNER[a][d] + CORNER[b][d]);
                    D[d] += set(a) ? (CORNER[a][d] - CORNER[b][d]) : (CORNER[b][d] - CORNER[a][d]);
                }
            }
            for (int i = 0; i < n; ++i) {
                const double* p = P[i];
                const double* q = P[(i + 1) % n];
                N[0] += (p[1] - q[1]) * (p[2] + q[2]);
                N[1] += (p[2] - q[2]) * (p[0] + q[0]);
                N[2] += (p[0] - q[0]) * (p[1] + q[1]);
            }
            const bool flip = N[0] * D[0] + N[1] * D[1] + N[2] * D[2] < 0;
            for (int i = 1; i + 1 < n; ++i) {
                T.tri[cs][n_out++] = (signed char)loop[0];
                T.tri[cs][n_out++] = (signed char)loop[flip ? i + 1 : i];
                T.tri[cs][n_out++] = (signed char)loop[flip ? i : i + 1];
            }
        }
        T.ntri[cs] = (unsigned char)(n_out / 3);
    }
}

__constant__ McTables c_mc;
std::once_flag g_mc_once[16];

int ensure_tables() {
    int dev = 0;
    LS_CHECK_CUDA(cudaGetDevice(&dev));
    LS_REQUIRE(dev >= 0 && dev < 16, "marching cubes: device ordinal above 15");
    cudaError_t err = cudaSuccess;
    std::call_once(g_mc_once[dev], [&] {
        static McTables T;
        static std::once_flag host_once;
        std::call_once(host_once, [] { build_tables(T); });
        err = cudaMemcpyToSymbol(c_mc, &T, sizeof(McTables));
    });
    LS_CHECK_CUDA(err);
    return LS_OK;
}

// value of the padded grid (mesh_extractor2.py:172: np.pad(occ, 1, constant -1e6)); n = unpadded points per axis
__device__ __forceinline__ float padded(const float* __restrict__ g, int n, int x, int y, int z) {
    if (x < 1 || y < 1 || z < 1 || x > n || y > n || z > n) return -1e6f;
    return g[((long long)(x - 1) * n + (y - 1)) * n + (z - 1)];
}

// pass 1: per padded grid point the number of +x/+y/+z edges that change sign; per cell the number of triangles
__global__ void k_mc_count(const float* __restrict__ g, int n, float iso, int* __restrict__ vcount, int* __restrict__ tcount) {
    const int P = n + 2;
    const long long tot = (long long)P * P * P;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= tot) return;
    const int z = (int)(i % P), y = (int)((i / P) % P), x = (int)(i / ((long long)P * P));
    const bool s0 = padded(g, n, x, y, z) <= iso;
    int nv = 0;
    if (x + 1 < P && (padded(g, n, x + 1, y, z) <= iso) != s0) ++nv;
    if (y + 1 < P && (padded(g, n, x, y + 1, z) <= iso) != s0) ++nv;
    if (z + 1 < P && (padded(g, n, x, y, z + 1) <= iso) != s0) ++nv;
    vcount[i] = nv;
    int nt = 0;
    if (x + 1 < P && y + 1 < P && z + 1 < P) {
        int cs = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int dx = (m == 1 || m == 2 || m == 5 || m == 6), dy = (m == 2 || m == 3 || m == 6 || m == 7), dz = m >= 4;
            if (padded(g, n, x + dx, y + dy, z + dz) <= iso) cs |= 1 << m;
        }
        nt = c_mc.ntri[cs];
    }
    tcount[i] = nt;
}

// exclusive scan of int32 counts: per-block sums, scan of the block sums by one block, final offsets
constexpr int SCAN_T = 1024;
__global__ void __launch_bounds__(SCAN_T) k_scan_blocks(const int* __restrict__ in, long long n, int* __restrict__ out,
                                                        int* __restrict__ block_sum) {
    __shared__ int sw[32];
    const long long i = blockIdx.x * (long long)SCAN_T + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int v = i < n ? in[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sw[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = sw[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, s, o);
            if (lane >= o) s += y;
        }
        sw[lane] = s;
    }
    __syncthreads();
    const int incl = x + (w > 0 ? sw[w - 1] : 0);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == SCAN_T - 1) block_sum[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_sums(int* __restrict__ block_sum, int nb, int* __restrict__ total) {
    __shared__ int sw[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += SCAN_T) {
        const int i = base + threadIdx.x;
        int v = i < nb ? block_sum[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sw[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = sw[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, s, o);
                if (lane >= o) s += y;
            }
            sw[lane] = s;
        }
        __syncthreads();
        const int incl = x + (w > 0 ? sw[w - 1] : 0) + carry;
        if (i < nb) block_sum[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_T - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_add(int* __restrict__ out, long long n, const int* __restrict__ block_sum) {
    const long long i = blockIdx.x * (long long)SCAN_T + threadIdx.x;
    if (i < n) out[i] += block_sum[blockIdx.x];
}

int exclusive_scan(const int* in, long long n, int* out, int* block_sum, int* total, cudaStream_t st) {
    const int nb = (int)((n + SCAN_T - 1) / SCAN_T);
    k_scan_blocks<<<nb, SCAN_T, 0, st>>>(in, n, out, block_sum);
    LS_CHECK_LAUNCH("k_scan_blocks");
    k_scan_sums<<<1, SCAN_T, 0, st>>>(block_sum, nb, total);
    LS_CHECK_LAUNCH("k_scan_sums");
    k_scan_add<<<nb, SCAN_T, 0, st>>>(out, n, block_sum);
    LS_CHECK_LAUNCH("k_scan_add");
    return LS_OK;
}

// vertex on the edge p -> p + e_axis (libmcubes mc_isovalue_interpolation, double), in the reference's final frame:
// v = box * ((c + 0.5 - 0.5 - 1) / (n - 1) - 0.5)  with c the padded-grid coordinate   (mesh_extractor2.py:175-180)
__global__ void k_mc_vertices(const float* __restrict__ g, int n, float iso, const int* __restrict__ voff, float box,
                              float* __restrict__ verts) {
    const int P = n + 2;
    const long long tot = (long long)P * P * P;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= tot) return;
    const int z = (int)(i % P), y = (int)((i / P) % P), x = (int)(i / ((long long)P * P));
    const double f0 = padded(g, n, x, y, z);
    const bool s0 = f0 <= (double)iso;
    int slot = voff[i];
    const int c[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        if (c[a] + 1 >= P) continue;
        const double f1 = padded(g, n, x + (a == 0), y + (a == 1), z + (a == 2));
        if ((f1 <= (double)iso) == s0) continue;
        const double tpar = (f1 == f0) ? 0.5 : ((double)iso - f0) / (f1 - f0);
        double p[3] = {(double)x, (double)y, (double)z};
        p[a] += tpar;
        for (int d = 0; d < 3; ++d) verts[(long long)slot * 3 + d] = (float)((double)box * ((p[d] - 1.0) / (double)(n - 1) - 0.5));
        ++slot;
    }
}

__global__ void k_mc_faces(const float* __restrict__ g, int n, float iso, const int* __restrict__ voff,
                           const int* __restrict__ toff, int64_t* __restrict__ faces) {
    const int P = n + 2;
    const long long tot = (long long)P * P * P;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= tot) return;
    const int z = (int)(i % P), y = (int)((i / P) % P), x = (int)(i / ((long long)P * P));
    if (x + 1 >= P || y + 1 >= P || z + 1 >= P) return;
    bool s[8];
    int cs = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int dx = (m == 1 || m == 2 || m == 5 || m == 6), dy = (m == 2 || m == 3 || m == 6 || m == 7), dz = m >= 4;
        s[m] = padded(g, n, x + dx, y + dy, z + dz) <= iso;
        if (s[m]) cs |= 1 << m;
    }
    const int nt = c_mc.ntri[cs];
    if (nt == 0) return;
    // vertex id of cell edge e: the grid edge starts at its lower corner and runs along one axis; the id is the owner
    // point's offset plus the number of sign-changing edges of lower axes at that point (the order k_mc_vertices emits)
    constexpr int E_AXIS[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
    constexpr int E_LO[12] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3};
    auto vid = [&](int e) -> long long {
        const int a = E_AXIS[e], m = E_LO[e];
        const int lx = x + (m == 1 || m == 2 || m == 5 || m == 6), ly = y + (m == 2 || m == 3 || m == 6 || m == 7), lz = z + (m >= 4);
        const long long p = ((long long)lx * P + ly) * P + lz;
        const bool b0 = s[m];
        int rank = 0;
        if (a > 0 && lx + 1 < P && (padded(g, n, lx + 1, ly, lz) <= iso) != b0) ++rank;
        if (a > 1 && ly + 1 < P && (padded(g, n, lx, ly + 1, lz) <= iso) != b0) ++rank;
        return (long long)voff[p] + rank;
    };
    long long out = (long long)toff[i] * 3;
    for (int k = 0; k < nt * 3; ++k) faces[out + k] = vid(c_mc.tri[cs][k]);
}

}  // namespace
}  // namespace ls

using namespace ls;

extern "C" {

int ls_mise_init(int32_t resolution0, int32_t depth, uint8_t* state, uint8_t* level, void* stream) {
    LS_REQUIRE(state && level && resolution0 >= 1 && depth >= 0 && depth <= 6, "mise: bad arguments");
    const long long R = (long long)resolution0 << depth;
    LS_REQUIRE(R <= 512, "mise: final resolution above 512");
    k_mise_init<<<1184, 256, 0, static_cast<cudaStream_t>(stream)>>>(state, level, (int)R, 1 << depth);
    LS_CHECK_LAUNCH("k_mise_init");
    return LS_OK;
}

int ls_mise_collect(const uint8_t* state, int32_t R, int32_t* list, int32_t capacity, int32_t* count, void* stream) {
    LS_REQUIRE(state && list && count && R >= 1 && capacity >= 0, "mise: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n1 = (long long)(R + 1) * (R + 1) * (R + 1);
    LS_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
    k_mise_collect<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(state, n1, list, capacity, count);
    LS_CHECK_LAUNCH("k_mise_collect");
    return LS_OK;
}

int ls_mise_points(const int32_t* list, int32_t n, int32_t R, float box_size, float* query, void* stream) {
    LS_REQUIRE(list && query && n >= 0 && R >= 1, "mise: bad arguments");
    if (n == 0) return LS_OK;
    k_mise_points<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(list, n, R, box_size, query);
    LS_CHECK_LAUNCH("k_mise_points");
    return LS_OK;
}

int ls_mise_update(const int32_t* list, const float* values, int32_t n, float value_scale, int32_t R, int32_t depth,
                   float threshold, float* val, uint8_t* state, uint8_t* level, uint8_t* pos, uint8_t* neg, void* stream) {
    LS_REQUIRE(list && values && val && state && level && pos && neg && n >= 0 && R >= 1, "mise: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n1 = (long long)(R + 1) * (R + 1) * (R + 1), n0 = (long long)R * R * R;
    if (n > 0) {
        k_mise_update<<<(n + 255) / 256, 256, 0, st>>>(list, values, n, value_scale, val, state);
        LS_CHECK_LAUNCH("k_mise_update");
    }
    LS_CHECK_CUDA(cudaMemsetAsync(pos, 0, (size_t)n0, st));
    LS_CHECK_CUDA(cudaMemsetAsync(neg, 0, (size_t)n0, st));
    k_mise_mark<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(state, val, level, R, depth, threshold, pos, neg);
    LS_CHECK_LAUNCH("k_mise_mark");
    k_mise_subdivide<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(level, pos, neg, state, R, depth);
    LS_CHECK_LAUNCH("k_mise_subdivide");
    return LS_OK;
}

int ls_mise_to_dense(const uint8_t* state, const float* val, int32_t R, float* dense, void* stream) {
    LS_REQUIRE(state && val && dense && R >= 1, "mise: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n1 = (long long)(R + 1) * (R + 1) * (R + 1);
    k_mise_dense<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(state, val, n1, dense);
    LS_CHECK_LAUNCH("k_mise_dense");
    const int lines = (R + 1) * (R + 1);
    for (int axis = 0; axis < 3; ++axis) {
        k_mise_fill<<<(lines + 127) / 128, 128, 0, st>>>(dense, R, axis);
        LS_CHECK_LAUNCH("k_mise_fill");
    }
    return LS_OK;
}

int ls_mcubes_workspace_bytes(int32_t n, size_t* bytes) {
    LS_REQUIRE(bytes && n >= 2 && n <= 1022, "marching cubes: grid size out of range");
    const long long P = n + 2, tot = P * P * P;
    *bytes = (size_t)(4 * tot + 2 * ((tot + SCAN_T - 1) / SCAN_T) + 64) * sizeof(int);
    return LS_OK;
}

// pass 1 (counts + offsets): n_out[0] = vertices, n_out[1] = triangles (device ints the caller reads back to size the
// outputs); pass 2 (ls_mcubes_emit) fills vertices [V,3] fp32 and faces [F,3] int64 with the same workspace.
int ls_mcubes_count(const float* grid, int32_t n, float iso, void* workspace, size_t workspace_bytes, int32_t* n_out,
                    void* stream) {
    LS_REQUIRE(grid && workspace && n_out, "null pointer");
    size_t need;
    int rc = ls_mcubes_workspace_bytes(n, &need);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(need <= workspace_bytes, "marching cubes workspace too small");
    if ((rc = ensure_tables()) != LS_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long P = n + 2, tot = P * P * P;
    int* vcount = static_cast<int*>(workspace);
    int *tcount = vcount + tot, *voff = tcount + tot, *toff = voff + tot;
    int* bs = toff + tot;
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    k_mc_count<<<blocks, 256, 0, st>>>(grid, n, iso, vcount, tcount);
    LS_CHECK_LAUNCH("k_mc_count");
    if ((rc = exclusive_scan(vcount, tot, voff, bs, n_out, st)) != LS_OK) return rc;
    if ((rc = exclusive_scan(tcount, tot, toff, bs + (tot + SCAN_T - 1) / SCAN_T, n_out + 1, st)) != LS_OK) return rc;
    return LS_OK;
}

int ls_mcubes_emit(const float* grid, int32_t n, float iso, float box_size, const void* workspace, float* vertices,
                   int64_t* faces, void* stream) {
    LS_REQUIRE(grid && workspace, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long P = n + 2, tot = P * P * P;
    const int* vcount = static_cast<const int*>(workspace);
    const int *voff = vcount + 2 * tot, *toff = voff + tot;
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    if (vertices) {
        k_mc_vertices<<<blocks, 256, 0, st>>>(grid, n, iso, voff, box_size, vertices);
        LS_CHECK_LAUNCH("k_mc_vertices");
    }
    if (faces) {
        k_mc_faces<<<blocks, 256, 0, st>>>(grid, n, iso, voff, toff, faces);
        LS_CHECK_LAUNCH("k_mc_faces");
    }
    return LS_OK;
}

}  // extern "C"
