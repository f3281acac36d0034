// Interface of the tensor-core kNN candidate filter (ls_knn_tc.cu) to the encoder orchestration and to the
// exact re-rank prologue of k_knn_edge (ls_encoder_kernels.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ls {

constexpr int KT_PTS = 128;  // points per packed tile = MMA M and N
constexpr int KT_KB = 8;     // feature dims per k-block = one tcgen05 kind::tf32 K step
constexpr int KT_CAP = 64;   // candidate slots per query in the global list

inline int knn_tc_tiles(int n) { return (n + KT_PTS - 1) / KT_PTS; }
inline int knn_tc_kblocks(int d) { return (d + KT_KB - 1) / KT_KB; }
// floats of the packed hi/lo image, of the norms and of the point-major copy, per instance
inline size_t knn_tc_img_floats(int n, int d) { return (size_t)knn_tc_tiles(n) * knn_tc_kblocks(d) * 2 * KT_PTS * KT_KB; }
inline size_t knn_tc_nrm_floats(int n) { return (size_t)knn_tc_tiles(n) * KT_PTS; }
inline size_t knn_tc_pm_floats(int n, int d) { return (size_t)n * knn_tc_kblocks(d) * KT_KB; }
inline size_t knn_tc_cand_u16(int nd) { return (size_t)knn_tc_tiles(nd) * KT_CAP * KT_PTS; }
// error budget of the ranking value relative to |q|^2 + max|s|^2 (see ls_knn_tc.cu)
inline float knn_tc_kappa(int d) { return 2.f * (float)d * 1.1920929e-07f + 1.52587890625e-05f; }

struct KnnTcArgs {
    const float* img_s;  // packed source images   [B][n_pt_s][n_kb][hi 1024 | lo 1024]
    const float* nrm_s;  // squared norms          [B][n_pt_s * 128]  (+inf beyond Ns)
    const float* img_q;  // packed query images
    const float* nrm_q;
    int Ns, Nd, n_pt_s, n_pt_q, n_kb;
    float kappa;
    unsigned short* cand;  // [B][n_pt_q][KT_CAP][128]  candidate source indices, slot-major
    float* cand_dt;        // [B][n_pt_q][KT_CAP][128]  their ranking values dt = |s|^2 - 2<q,s>
    int* cnt;              // [B][n_pt_q * 128]         candidates per query, -1 = overflow (brute-force it)
    float* e2;             // [B][n_pt_q * 128]         2E of the query: two dt closer than this are "ambiguous"
};

int launch_knn_pack(const float* f, int B, int D, int N, float* img, float* nrm, float* pm, cudaStream_t st);
int launch_knn_tc(const KnnTcArgs& a, int B, cudaStream_t st);

}  // namespace ls
