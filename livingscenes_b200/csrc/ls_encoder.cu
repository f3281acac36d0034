// Host orchestration of the encoder path behind the C ABI (include/livingscenes_b200.h):
//   ls_encoder_forward   VecDGCNN_att.forward / Shape_Prior.encode
//   ls_knn, ls_fps       the pytorch3d boundary as stand-alone ops
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <memory>
#include <vector>

#include "ls_encoder_kernels.cuh"

namespace ls {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-stage CUDA-event timing of the last ls_encoder_forward call
struct ProfEntry {
    cudaEvent_t a, b;
    int stage, layer;
};
static std::vector<ProfEntry> g_prof;
static int g_prof_n = 0;
static bool g_prof_on = false;

struct ProfScope {
    cudaStream_t st;
    int slot = -1;
    ProfScope(int stage, int layer, cudaStream_t s) : st(s) {
        if (!g_prof_on) return;
        if (g_prof_n == (int)g_prof.size()) {
            ProfEntry e{};
            if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
            g_prof.push_back(e);
        }
        slot = g_prof_n++;
        g_prof[slot].stage = stage;
        g_prof[slot].layer = layer;
        cudaEventRecord(g_prof[slot].a, st);
    }
    ~ProfScope() {
        if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
    }
};

// ---- fork/join helper: independent stages of one forward (FPS chain, the point-level GEMM tables) run on a
//      library-owned side stream next to the kNN chain.  Works eagerly and under stream capture (the waits on
//      events recorded in the capturing stream pull the side stream into the capture as a parallel branch).
struct SideCtx {
    int dev = -1;
    cudaStream_t s_fps = nullptr, s_tab = nullptr;  // FPS chain / point-level table GEMMs
    cudaEvent_t fork_ev = nullptr, join_fps = nullptr, ev_tab[2] = {nullptr, nullptr}, ev_edge[2] = {nullptr, nullptr};
    bool ok = false;
    void destroy() {
        if (s_fps) cudaStreamDestroy(s_fps);
        if (s_tab) cudaStreamDestroy(s_tab);
        for (cudaEvent_t e : {fork_ev, join_fps, ev_tab[0], ev_tab[1], ev_edge[0], ev_edge[1]})
            if (e) cudaEventDestroy(e);
        *this = SideCtx();
    }
};
constexpr int MAX_SIDE_DEVICES = 16;
static thread_local SideCtx g_side[MAX_SIDE_DEVICES];  // one context per (thread, device): nothing leaks when a thread alternates devices
static bool g_overlap = true;

static SideCtx* side_ctx(cudaStream_t main_stream) {
    if (!g_overlap || g_prof_on) return nullptr;  // per-stage event timing needs the serial order
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_SIDE_DEVICES) return nullptr;
    SideCtx& c = g_side[dev];
    if (c.ok) return &c;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main_stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
        return nullptr;  // resources are only created outside a capture; this call runs serially
    c.dev = dev;
    bool good = cudaStreamCreateWithFlags(&c.s_fps, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&c.s_tab, cudaStreamNonBlocking) == cudaSuccess;
    cudaEvent_t* evs[] = {&c.fork_ev, &c.join_fps, &c.ev_tab[0], &c.ev_tab[1], &c.ev_edge[0], &c.ev_edge[1]};
    for (cudaEvent_t* e : evs) good = good && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!good) {
        c.destroy();  // nothing half-created is kept
        cudaGetLastError();
        return nullptr;
    }
    c.ok = true;
    return &c;
}

namespace {

struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

constexpr int SMALL_NS = 128;  // source sets up to this size take the k_knn_small path (brute-force graph only)
// with the tensor-core graph enabled, source sets down to this size still go through k_knn_tc
int small_tc_min() {
    static const int v = [] {
        const char* e = getenv("LS_SMALL_TC_MIN");
        return e ? atoi(e) : 129;
    }();
    return v;
}
// Wave scheduling of {table GEMMs -> EdgeConv} (OFF by default; ls_set_wave_bytes / LS_WAVE_MB): the point-level gather
// tables of a layer are 1.6-5.5 MB per instance (19.5 MB over the six layers); written for the whole batch they leave
// L2 long before the EdgeConv reads them back.  A wave is as many instances as fit g_wave_bytes of tables; two table
// slots alternate, so the GEMM of wave w+1 overlaps the EdgeConv of wave w and the slot stays dirty in L2.
// MEASURED (profiles/r02/experiments.md): at these table sizes an L2-sized wave is only 7-25 instances, the per-wave
// launches run at a fraction of the GPU and the step gets 37 % SLOWER (11.3 -> 15.4 ms at 28 MB, 22.8 ms at 14 MB),
// so the whole batch stays one launch per layer.  The schedule is kept, result-invariant and tested, for larger L2s.
long long g_wave_bytes = [] {
    const char* e = getenv("LS_WAVE_MB");
    return (long long)(e ? atof(e) : 0.0) * (1LL << 20);
}();
// small source sets (Ns <= 128): shared-memory tiled kernel (default) or the round-1 warp-per-query kernel (LS_KNN_SMALL_TILED=0)
bool g_knn_small_tiled = [] {
    const char* e = getenv("LS_KNN_SMALL_TILED");
    return !(e && atoi(e) == 0);
}();
int g_fps_fma = 0;             // FPS squared distance: 0 = every product / sum rounded (oracle, fixtures), 1 = FMA-contracted
bool g_use_knn_tc = true;      // tensor-core candidate filter for the larger source sets
float g_knn_tc_kappa_scale = 1.f;

struct Plan {
    int n_src[LS_MAX_LAYERS], n_dst[LS_MAX_LAYERS];
    float *xn, *centroid, *s0, *featA, *featB, *dstf, *pooled, *raw, *psrc, *pdst, *bias, *gmean;
    int64_t* small_idx;
    // tensor-core kNN filter (ls_knn_tc.cu): packed images / norms / point-major copies and candidate lists
    float *kimg_s, *knrm_s, *kpm_s, *kimg_q, *knrm_q, *kpm_q;
    unsigned short* kcand;
    float *kcand_dt, *ke2;
    int* kcnt;
    int64_t* kidx;
    int* sel[LS_MAX_LAYERS];
    size_t bytes;
};

int check_desc(const ls_encoder_desc* d, int N) {
    LS_REQUIRE(d != nullptr, "null encoder descriptor");
    LS_REQUIRE(d->num_layers >= 1 && d->num_layers <= LS_MAX_LAYERS, "num_layers out of range");
    LS_REQUIRE(d->c_dim % 32 == 0 && d->c_dim >= 32 && d->c_dim <= 1024, "c_dim must be a multiple of 32 <= 1024");
    int n = N;
    for (int i = 0; i < d->num_layers; ++i) {
        const ls_enc_layer_desc& L = d->layers[i];
        LS_REQUIRE(L.c_out % 32 == 0 && L.c_out >= 32 && L.c_out <= 512, "c_out must be a multiple of 32 in [32,512]");
        LS_REQUIRE(L.down_factor >= 1, "down_factor must be >= 1");
        if (i == 0) {
            LS_REQUIRE(L.c_in == 1 && !L.attention && !L.global_conv && L.down_factor == 1 && L.w0,
                       "layer 0 must be the plain xyz EdgeConv layer");
        } else {
            LS_REQUIRE(L.c_in == d->layers[i - 1].c_out, "c_in must equal the previous layer's c_out");
            LS_REQUIRE(L.c_in % 8 == 0, "c_in must be a multiple of 8");
            LS_REQUIRE(L.w_src && L.w_dst, "missing folded weights");
            LS_REQUIRE(!L.global_conv || (L.w_g1 && L.w_g2), "missing global-conv weights");
            int cpl = L.c_out / 32;
            LS_REQUIRE(cpl == 1 || cpl == 2 || cpl == 4 || cpl == 8 || cpl == 16, "c_out/32 must be a power of two <= 16");
        }
        n /= L.down_factor;  // floor, like N_ori // factor in vec_dgcnn_atten.py:166
        LS_REQUIRE(n >= LS_KNN_K, "too few points for K=16 neighbours at a deep layer (N too small)");
    }
    LS_REQUIRE(d->w_conv_c && d->w_inv_t, "missing head weights");
    LS_REQUIRE(!d->center_pred || (d->w_fc0_t && d->w_lin1 && d->w_short), "missing fc_center weights");
    return LS_OK;
}

void make_plan(const ls_encoder_desc* d, int B, int N, void* ws, Plan& p) {
    Carver c(ws);
    size_t feat = 0, dstf = 0, pooled = 0, raw = 0, psrc = 0, pdst = 0, bias = 0, small = 0;
    size_t kimg_s = 0, knrm_s = 0, kpm_s = 0, kimg_q = 0, knrm_q = 0, kpm_q = 0, kcand = 0;
    int n = N;
    for (int i = 0; i < d->num_layers; ++i) {
        const ls_enc_layer_desc& L = d->layers[i];
        p.n_src[i] = n;
        n /= L.down_factor;
        p.n_dst[i] = n;
        const size_t co = L.c_out, ci = L.c_in;
        if (p.n_src[i] <= SMALL_NS) small = std::max(small, (size_t)n * LS_KNN_K);
        {
            const int D = (int)ci * 3;
            kimg_s = std::max(kimg_s, knn_tc_img_floats(p.n_src[i], D));
            knrm_s = std::max(knrm_s, knn_tc_nrm_floats(p.n_src[i]));
            kpm_s = std::max(kpm_s, knn_tc_pm_floats(p.n_src[i], D));
            if (L.down_factor > 1) {
                kimg_q = std::max(kimg_q, knn_tc_img_floats(n, D));
                knrm_q = std::max(knrm_q, knn_tc_nrm_floats(n));
                kpm_q = std::max(kpm_q, knn_tc_pm_floats(n, D));
            }
            kcand = std::max(kcand, knn_tc_cand_u16(n));
        }
        feat = std::max(feat, co * 3 * (size_t)n);
        if (L.down_factor > 1) dstf = std::max(dstf, ci * 3 * (size_t)n);
        if (L.global_conv) {
            pooled = std::max(pooled, co * 3 * (size_t)n);
            raw = std::max(raw, 2 * co * 3 * (size_t)n);
            bias = std::max(bias, 2 * co * 3);
        }
        if (i > 0) {
            const size_t nb = L.attention ? 2 : 1;
            psrc = std::max(psrc, 2 * nb * co * 3 * (size_t)p.n_src[i]);
            pdst = std::max(pdst, (2 * nb + (L.attention ? 2 : 0)) * co * 3 * (size_t)n);
        }
    }
    raw = std::max(raw, (size_t)(d->c_dim + 1) * 3 * n);
    p.xn = c.take<float>((size_t)B * 3 * N);
    p.centroid = c.take<float>((size_t)B * 3);
    p.s0 = c.take<float>((size_t)B);
    p.featA = c.take<float>((size_t)B * feat);
    p.featB = c.take<float>((size_t)B * feat);
    p.dstf = c.take<float>((size_t)B * std::max<size_t>(dstf, 1));
    p.pooled = c.take<float>((size_t)B * std::max<size_t>(pooled, 1));
    p.raw = c.take<float>((size_t)B * raw);
    p.psrc = c.take<float>((size_t)B * std::max<size_t>(psrc, 1));
    p.pdst = c.take<float>((size_t)B * std::max<size_t>(pdst, 1));
    p.bias = c.take<float>((size_t)B * std::max<size_t>(bias, 1));
    p.gmean = c.take<float>((size_t)B * std::max<size_t>(bias, 1));
    p.small_idx = c.take<int64_t>((size_t)B * std::max<size_t>(small, 1));
    p.kimg_s = c.take<float>((size_t)B * std::max<size_t>(kimg_s, 1));
    p.knrm_s = c.take<float>((size_t)B * std::max<size_t>(knrm_s, 1));
    p.kpm_s = c.take<float>((size_t)B * std::max<size_t>(kpm_s, 1));
    p.kimg_q = c.take<float>((size_t)B * std::max<size_t>(kimg_q, 1));
    p.knrm_q = c.take<float>((size_t)B * std::max<size_t>(knrm_q, 1));
    p.kpm_q = c.take<float>((size_t)B * std::max<size_t>(kpm_q, 1));
    p.kcand = c.take<unsigned short>((size_t)B * std::max<size_t>(kcand, 1));
    p.kcand_dt = c.take<float>((size_t)B * std::max<size_t>(kcand, 1));
    p.kcnt = c.take<int>((size_t)B * std::max<size_t>(knrm_s + knrm_q, 1));
    p.ke2 = c.take<float>((size_t)B * std::max<size_t>(knrm_s + knrm_q, 1));
    p.kidx = c.take<int64_t>((size_t)B * std::max<size_t>(knrm_s + knrm_q, 1) * LS_KNN_K);
    for (int i = 0; i < d->num_layers; ++i)
        p.sel[i] = d->layers[i].down_factor > 1 ? c.take<int>((size_t)B * p.n_dst[i]) : nullptr;
    p.bytes = (c.off + 255) & ~size_t(255);
}

template <int MODE>
int launch_edge_cpl(const EdgeArgs& a, int cpl, dim3 grid, cudaStream_t st) {
    if (a.idx_in != nullptr) {  // graph given: phase-2-only instantiation (higher occupancy)
        switch (cpl) {
            case 1: k_knn_edge<MODE, 1, false><<<grid, EDGE_THREADS, 0, st>>>(a); break;
            case 2: k_knn_edge<MODE, 2, false><<<grid, EDGE_THREADS, 0, st>>>(a); break;
            case 4: k_knn_edge<MODE, 4, false><<<grid, EDGE_THREADS, 0, st>>>(a); break;
            case 8: k_knn_edge<MODE, 8, false><<<grid, EDGE_THREADS, 0, st>>>(a); break;
            case 16: k_knn_edge<MODE, 16, false><<<grid, EDGE_THREADS, 0, st>>>(a); break;
            default: set_error("unsupported c_out/32"); return LS_ERR_INVALID;
        }
        LS_CHECK_LAUNCH("k_knn_edge");
        return LS_OK;
    }
    switch (cpl) {
        case 1: k_knn_edge<MODE, 1><<<grid, EDGE_THREADS, 0, st>>>(a); break;
        case 2: k_knn_edge<MODE, 2><<<grid, EDGE_THREADS, 0, st>>>(a); break;
        case 4: k_knn_edge<MODE, 4><<<grid, EDGE_THREADS, 0, st>>>(a); break;
        case 8: k_knn_edge<MODE, 8><<<grid, EDGE_THREADS, 0, st>>>(a); break;
        case 16: k_knn_edge<MODE, 16><<<grid, EDGE_THREADS, 0, st>>>(a); break;
        default: set_error("unsupported c_out/32"); return LS_ERR_INVALID;
    }
    LS_CHECK_LAUNCH("k_knn_edge");
    return LS_OK;
}

// dst points per CTA: as many as the tile holds, fewer when the grid would not fill the GPU twice.
// Phase-2-only launches gather rows of the instance's point-level tables (table_bytes per instance); the CTAs of
// one instance are adjacent in the grid, so the number of instances in flight is (resident CTAs) / (CTAs per
// instance).  Keeping that working set inside L2 (126 MB) turns the repeated row gathers into L2 hits -- with 64
// points per CTA layers 2-4 had 55-220 instances (0.3-1.2 GB of tables) in flight and re-read every row from HBM.
int pick_qpc(int B, int Nd, bool phase2_only, int Co = 0, size_t table_bytes = 0) {
    int qpc = QT;
    int floor_q = phase2_only ? 8 : 16;
    if (phase2_only && Co > 0) {
        const int lpp = std::max(1, std::min(Co / 4, 32));   // lanes per point in phase 2
        floor_q = std::max(8, 8 * (32 / lpp));                // one pass of the CTA's 8 warps
    }
    while (qpc > floor_q && (long long)B * ((Nd + qpc - 1) / qpc) < 4 * 148) qpc >>= 1;
    if (phase2_only && table_bytes > 0) {
        const long long slots = 148LL * LS_P2_CTAS, budget = 64LL << 20;
        while (qpc > floor_q) {
            const long long per_inst = (Nd + qpc - 1) / qpc;
            const long long in_flight = (slots + per_inst - 1) / per_inst;
            if (in_flight * (long long)table_bytes <= budget) break;
            qpc >>= 1;
        }
    }
    return qpc;
}

template <bool FB>
int launch_rerank_one(const RerankArgs& ra, dim3 grid, cudaStream_t st) {
    switch (ra.Dp) {
        case 8: k_knn_rerank<8, FB><<<grid, RR_WARPS * 32, 0, st>>>(ra); break;
        case 96: k_knn_rerank<96, FB><<<grid, RR_WARPS * 32, 0, st>>>(ra); break;
        case 192: k_knn_rerank<192, FB><<<grid, RR_WARPS * 32, 0, st>>>(ra); break;
        default: k_knn_rerank<0, FB><<<grid, RR_WARPS * 32, 0, st>>>(ra); break;
    }
    LS_CHECK_LAUNCH("k_knn_rerank");
    return LS_OK;
}
// two launches: the common path at 3 CTAs per SM, then the rare heavy paths (CTAs without a flagged query exit at once)
int launch_rerank(const RerankArgs& ra, int B, cudaStream_t st) {
    const dim3 grid((ra.Nd + RR_WARPS * RR_QPW - 1) / (RR_WARPS * RR_QPW), B);
    if (!ra.all_exact) {
        const int rc = launch_rerank_one<false>(ra, grid, st);
        if (rc != LS_OK) return rc;
    }
    return launch_rerank_one<true>(ra, dim3((ra.Nd + RR_FB_Q - 1) / RR_FB_Q, B), st);
}

int launch_edge(int mode, const EdgeArgs& a, cudaStream_t st) {
    LS_REQUIRE(a.qpc >= 8 && a.qpc <= QT && a.qpc % 8 == 0, "bad qpc");
    dim3 grid((a.Nd + a.qpc - 1) / a.qpc, a.B);
    const int cpl = a.Co / 32;
    if (mode == MODE_L0) return launch_edge_cpl<MODE_L0>(a, cpl, grid, st);
    if (mode == MODE_MEAN) return launch_edge_cpl<MODE_MEAN>(a, cpl, grid, st);
    if (mode == MODE_ATT) return launch_edge_cpl<MODE_ATT>(a, cpl, grid, st);
    k_knn_edge<MODE_KNN_ONLY, 1><<<grid, EDGE_THREADS, 0, st>>>(a);
    LS_CHECK_LAUNCH("k_knn_only");
    return LS_OK;
}

int launch_fps(FpsArgs a, int B, cudaStream_t st) {
    a.fma = g_fps_fma;
    const int N = a.N;
    LS_REQUIRE(N >= 1 && N <= 8192, "fps: N must be in [1, 8192] (in-register running min distance)");
    for (int l = 0; l < a.n_levels; ++l) {
        int prev = l == 0 ? N : a.n_out[l - 1];
        LS_REQUIRE(a.n_out[l] >= 1 && a.n_out[l] <= prev, "fps: n_out must not exceed the number of points");
    }
    // few warps, several points per thread: every selected point costs one block-wide arg-max, whose
    // instruction count grows with the number of warps (each warp redoes the final reduction), so the
    // kernel is issue-bound with many threads.  8 points per thread up to N = 8192.
    int T = 128;
    while (T < 1024 && T * 8 < N) T *= 2;
    const int ppt = (N + T - 1) / T;
    const size_t smem = (size_t)(3 * N + 4 * a.n_out[0]) * sizeof(float);
#define LS_FPS_CASE(P)                                                                              \
    {                                                                                               \
        LS_CHECK_CUDA(cudaFuncSetAttribute(k_fps<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_fps<P><<<B, T, smem, st>>>(a);                                                            \
    }
    if (ppt <= 1) LS_FPS_CASE(1)
    else if (ppt <= 2) LS_FPS_CASE(2)
    else if (ppt <= 4) LS_FPS_CASE(4)
    else LS_FPS_CASE(8)
#undef LS_FPS_CASE
    LS_CHECK_LAUNCH("k_fps");
    return LS_OK;
}

}  // namespace
}  // namespace ls

using namespace ls;

extern "C" {

int ls_version(void) { return LS_ABI_VERSION; }
int64_t ls_kernel_launches(void) { return (int64_t)ls::g_launches.load(); }
int ls_profile_enable(int32_t on) {
    ls::g_prof_on = on != 0;
    ls::g_prof_n = 0;
    return LS_OK;
}
int ls_profile_read(int32_t* stage, int32_t* layer, float* ms, int32_t max_entries, int32_t* n_entries) {
    LS_REQUIRE(stage && layer && ms && n_entries, "null pointer");
    int n = ls::g_prof_n < max_entries ? ls::g_prof_n : max_entries;
    for (int i = 0; i < n; ++i) {
        stage[i] = ls::g_prof[i].stage;
        layer[i] = ls::g_prof[i].layer;
        LS_CHECK_CUDA(cudaEventElapsedTime(&ms[i], ls::g_prof[i].a, ls::g_prof[i].b));
    }
    *n_entries = n;
    return LS_OK;
}
const char* ls_last_error(void) { return ls::g_last_error.c_str(); }

int ls_encoder_workspace_bytes(const ls_encoder_desc* desc, int32_t B, int32_t N, size_t* bytes) {
    LS_REQUIRE(bytes != nullptr && B >= 1 && N >= 1, "bad arguments");
    int rc = check_desc(desc, N);
    if (rc != LS_OK) return rc;
    Plan p;
    make_plan(desc, B, N, nullptr, p);
    *bytes = p.bytes;
    return LS_OK;
}

int ls_encoder_forward(const ls_encoder_desc* d, const ls_encoder_io* io, void* workspace,
                       size_t workspace_bytes, void* stream) {
    LS_REQUIRE(io != nullptr && io->x != nullptr, "null io");
    const int B = io->B, N = io->N;
    LS_REQUIRE(B >= 1 && N >= 1 && B <= 65535, "B must be in [1, 65535]");
    int rc = check_desc(d, N);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(io->scale && io->z_so3 && io->z_inv, "missing output pointers");
    LS_REQUIRE(!d->center_pred || io->center, "center output required when center_pred is set");
    LS_REQUIRE(workspace != nullptr, "null workspace");
    Plan p;
    make_plan(d, B, N, workspace, p);
    if (p.bytes > workspace_bytes) {
        set_error("workspace too small");
        return LS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float oms = 1.f - d->neg_slope;
    g_prof_n = 0;

    // ---- pre-processing (Shape_Prior.encode) ------------------------------------------------
    const float* x = io->x;
    if (io->normalize) {
        ProfScope ps(0, -1, st);
        const size_t smem = (size_t)(3 * N + 64) * sizeof(float);
        LS_REQUIRE(smem <= 200 * 1024, "normalize: N too large for the shared-memory resident cloud");
        LS_CHECK_CUDA(cudaFuncSetAttribute(k_normalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_normalize<<<B, NORM_THREADS, smem, st>>>(io->x, N, p.xn, p.centroid, p.s0);
        LS_CHECK_LAUNCH("k_normalize");
        x = p.xn;
        if (io->scale0) LS_CHECK_CUDA(cudaMemcpyAsync(io->scale0, p.s0, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
        if (io->x_norm)
            LS_CHECK_CUDA(cudaMemcpyAsync(io->x_norm, p.xn, sizeof(float) * (size_t)B * 3 * N, cudaMemcpyDeviceToDevice, st));
    }

    SideCtx* sc = side_ctx(st);
    bool fps_pending = false;
    // ---- FPS chain: all down-sampling selections depend on xyz only --------------------------
    {
        FpsArgs fa{};
        fa.xyz = x;
        fa.N = N;
        fa.n_levels = 0;
        for (int i = 0; i < d->num_layers; ++i) {
            if (d->layers[i].down_factor > 1) {
                LS_REQUIRE(fa.n_levels < FPS_MAX_LEVELS, "too many down-sampling layers");
                fa.n_out[fa.n_levels] = p.n_dst[i];
                fa.sel32[fa.n_levels] = p.sel[i];
                fa.sel64[fa.n_levels] = io->fps_idx[i];
                fa.force[fa.n_levels] = io->force_fps_idx[i];
                ++fa.n_levels;
            }
        }
        if (fa.n_levels > 0) {
            if (sc) {  // FPS only needs xyz: it runs beside layers 0..first down-sampling layer
                LS_CHECK_CUDA(cudaEventRecord(sc->fork_ev, st));
                LS_CHECK_CUDA(cudaStreamWaitEvent(sc->s_fps, sc->fork_ev, 0));
                rc = launch_fps(fa, B, sc->s_fps);
                if (rc != LS_OK) {
                    cudaStreamWaitEvent(st, sc->fork_ev, 0);
                    return rc;
                }
                LS_CHECK_CUDA(cudaEventRecord(sc->join_fps, sc->s_fps));
                fps_pending = true;
            } else {
                ProfScope ps(1, -1, st);
                rc = launch_fps(fa, B, st);
                if (rc != LS_OK) return rc;
            }
        }
    }

    // ---- layers -------------------------------------------------------------------------------
    const float* src_f = x;  // [B][C*3][Ns]
    float* bufs[2] = {p.featA, p.featB};
    int cur = 0;
    for (int i = 0; i < d->num_layers; ++i) {
        const ls_enc_layer_desc& L = d->layers[i];
        const int Ns = p.n_src[i], Nd = p.n_dst[i], Ci = L.c_in, Co = L.c_out;
        const float* dst_f = src_f;
        if (L.down_factor > 1) {
            if (fps_pending) {
                LS_CHECK_CUDA(cudaStreamWaitEvent(st, sc->join_fps, 0));
                fps_pending = false;
            }
            ProfScope ps(2, i, st);
            dim3 g((Ci * 3 * Nd + 255) / 256, B);
            k_gather_points<<<g, 256, 0, st>>>(src_f, p.sel[i], Ci * 3, Ns, Nd, p.dstf);
            LS_CHECK_LAUNCH("k_gather_points");
            dst_f = p.dstf;
        }
        float* layer_out = bufs[cur];
        float* edge_out = L.global_conv ? p.pooled : layer_out;

        EdgeArgs ea{};
        ea.src_f = src_f;
        ea.dst_f = dst_f;
        ea.B = B;
        ea.D = Ci * 3;
        ea.Ns = Ns;
        ea.Nd = Nd;
        ea.Co = Co;
        ea.oms = oms;
        ea.out = edge_out;
        ea.idx_out = io->knn_idx[i];
        ea.idx_in = io->force_knn_idx[i];

        // the point-level table GEMMs (layers >= 1) only need the layer input: they may start now on the side stream
        const int nb = L.attention ? 2 : 1;
        const int r_src = 2 * nb * Co, r_dst = (2 * nb + (L.attention ? 2 : 0)) * Co;
        const bool side_tables = i > 0 && sc != nullptr;
        if (side_tables) {
            LS_CHECK_CUDA(cudaEventRecord(sc->fork_ev, st));
            LS_CHECK_CUDA(cudaStreamWaitEvent(sc->s_tab, sc->fork_ev, 0));
        }

        // ---- kNN graph (whole batch, caller's stream)
        if (ea.idx_in == nullptr && Ns <= SMALL_NS && !(g_use_knn_tc && Ns >= small_tc_min())) {
            ProfScope ps(4, i, st);
            if (g_knn_small_tiled) {
                k_knn_small_tiled<<<dim3((Nd + 31) / 32, B), 256, 0, st>>>(src_f, dst_f, Ci * 3, Ns, Nd, p.small_idx, nullptr);
            } else {
                k_knn_small<<<dim3((Nd + 7) / 8, B), 256, 0, st>>>(src_f, dst_f, Ci * 3, Ns, Nd, p.small_idx, nullptr);
            }
            LS_CHECK_LAUNCH("k_knn_small");
            ea.idx_in = p.small_idx;
        }
        if (ea.idx_in == nullptr && g_use_knn_tc && Ns <= 65535) {  // candidate indices are 16-bit; larger sets: brute force
            // tensor-core candidate filter + exact re-rank instead of the brute-force phase 1
            std::unique_ptr<ProfScope> ps(new ProfScope(7, i, st));
            const int D = Ci * 3;
            rc = launch_knn_pack(src_f, B, D, Ns, p.kimg_s, p.knrm_s, p.kpm_s, st);
            if (rc != LS_OK) return rc;
            KnnTcArgs ka{};
            ka.img_s = p.kimg_s;
            ka.nrm_s = p.knrm_s;
            ka.img_q = p.kimg_s;
            ka.nrm_q = p.knrm_s;
            RerankArgs ra{};
            ra.pm_s = p.kpm_s;
            ra.pm_q = p.kpm_s;
            if (L.down_factor > 1) {
                rc = launch_knn_pack(dst_f, B, D, Nd, p.kimg_q, p.knrm_q, p.kpm_q, st);
                if (rc != LS_OK) return rc;
                ka.img_q = p.kimg_q;
                ka.nrm_q = p.knrm_q;
                ra.pm_q = p.kpm_q;
            }
            ka.Ns = Ns;
            ka.Nd = Nd;
            ka.n_pt_s = knn_tc_tiles(Ns);
            ka.n_pt_q = knn_tc_tiles(Nd);
            ka.n_kb = knn_tc_kblocks(D);
            ka.kappa = g_knn_tc_kappa_scale * knn_tc_kappa(D);
            ka.cand = p.kcand;
            ka.cand_dt = p.kcand_dt;
            ka.cnt = p.kcnt;
            ka.e2 = p.ke2;
            rc = launch_knn_tc(ka, B, st);
            if (rc != LS_OK) return rc;
            ps.reset();
            ps.reset(new ProfScope(8, i, st));
            ra.cand = p.kcand;
            ra.cand_dt = p.kcand_dt;
            ra.cand_cnt = p.kcnt;
            ra.e2 = p.ke2;
            ra.Ns = Ns;
            ra.Nd = Nd;
            ra.Dp = ka.n_kb * KT_KB;
            ra.idx_out = p.kidx;
            ra.idx_tap = io->knn_idx[i];
            rc = launch_rerank(ra, B, st);
            if (rc != LS_OK) return rc;
            ea.idx_in = p.kidx;
            ea.idx_out = nullptr;
        }

        if (i == 0) {
            ea.qpc = pick_qpc(B, Nd, ea.idx_in != nullptr);
            ea.w0 = L.w0;
            ProfScope ps(4, i, st);
            rc = launch_edge(MODE_L0, ea, st);
            if (rc != LS_OK) return rc;
        } else {
            // ---- waves of {table GEMMs (side stream) -> EdgeConv + pooling (caller's stream)}
            const size_t tab_s = (size_t)r_src * 3 * Ns, tab_d = (size_t)r_dst * 3 * Nd;  // floats per instance
            const size_t tab_bytes = sizeof(float) * (tab_s + tab_d);
            int W = B;
            if (g_wave_bytes > 0) W = (int)std::max<long long>(1, std::min<long long>(B, g_wave_bytes / (long long)tab_bytes));
            if (W * 2 > B) W = B;  // fewer than two full waves: not worth splitting
            const int n_waves = (B + W - 1) / W;
            cudaStream_t ts = side_tables ? sc->s_tab : st;
            const EdgeArgs ea0 = ea;
            for (int wv = 0; wv < n_waves; ++wv) {
                const int b0 = wv * W, Bw = std::min(W, B - b0), slot = n_waves > 1 ? (wv & 1) : 0;
                float* psrc_w = p.psrc + (size_t)slot * W * tab_s;
                float* pdst_w = p.pdst + (size_t)slot * W * tab_d;
                if (side_tables && wv >= 2) LS_CHECK_CUDA(cudaStreamWaitEvent(ts, sc->ev_edge[slot], 0));  // slot drained
                {
                    ProfScope pg(3, i, ts);
                    GemmArgs g{};
                    g.K = Ci;
                    g.ldw = Ci;
                    g.B = Bw;
                    g.point_major = 1;
                    g.c_out = Co;
                    // source table
                    g.W = L.w_src;
                    g.Wtc = L.w_src_tc;
                    g.R = r_src;
                    g.X = src_f + (size_t)b0 * Ci * 3 * Ns;
                    g.n_per_b = 3 * Ns;
                    g.npts = Ns;
                    g.x_sb = (long long)Ci * 3 * Ns;
                    g.x_sk = 3LL * Ns;
                    g.out = psrc_w;
                    rc = launch_gemm(g, ts);
                    if (rc == LS_OK) {
                        // dst table
                        g.W = L.w_dst;
                        g.Wtc = L.w_dst_tc;
                        g.R = r_dst;
                        g.X = dst_f + (size_t)b0 * Ci * 3 * Nd;
                        g.n_per_b = 3 * Nd;
                        g.npts = Nd;
                        g.x_sb = (long long)Ci * 3 * Nd;
                        g.x_sk = 3LL * Nd;
                        g.out = pdst_w;
                        rc = launch_gemm(g, ts);
                    }
                }
                if (side_tables) {  // join the side stream (also on the error path: an un-joined fork would break a capture)
                    cudaEventRecord(sc->ev_tab[slot], ts);
                    cudaStreamWaitEvent(st, sc->ev_tab[slot], 0);
                }
                if (rc != LS_OK) return rc;
                ea = ea0;
                ea.B = Bw;
                ea.src_f = ea0.src_f + (size_t)b0 * Ci * 3 * Ns;
                ea.dst_f = ea0.dst_f + (size_t)b0 * Ci * 3 * Nd;
                ea.out = ea0.out + (size_t)b0 * Co * 3 * Nd;
                if (ea0.idx_in) ea.idx_in = ea0.idx_in + (size_t)b0 * Nd * LS_KNN_K;
                if (ea0.idx_out) ea.idx_out = ea0.idx_out + (size_t)b0 * Nd * LS_KNN_K;
                ea.psrc = psrc_w;
                ea.pdst = pdst_w;
                ea.row_s = r_src * 3;
                ea.row_d = r_dst * 3;
                ea.qpc = pick_qpc(Bw, Nd, ea.idx_in != nullptr, Co, n_waves > 1 ? 0 : tab_bytes);  // a wave's tables fit L2 by construction
                {
                    ProfScope ps(4, i, st);
                    rc = launch_edge(L.attention ? MODE_ATT : MODE_MEAN, ea, st);
                    if (rc != LS_OK) return rc;
                }
                if (side_tables && wv + 2 < n_waves) LS_CHECK_CUDA(cudaEventRecord(sc->ev_edge[slot], st));
            }
        }
        if (L.global_conv) {
            ProfScope ps(5, i, st);
            k_row_mean<<<dim3(Nd == 32 ? (Co * 3 + 63) / 64 : (Co * 3 + 7) / 8, B), 256, 0, st>>>(p.pooled, Co * 3, Nd, p.gmean);
            LS_CHECK_LAUNCH("k_row_mean");
            if (Co >= 128) {
                // bias[b][r][a] = sum_c Wg2[r][c] g[b][c][a]: register-tiled, 128 rows x 8 instances per CTA
                const size_t smem = ((size_t)BIAS_BG * (Co * 3 + 1) + (size_t)BIAS_ROWS * (BIAS_KT + 1)) * sizeof(float);
                LS_CHECK_CUDA(cudaFuncSetAttribute(k_bias_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_bias_rows<<<dim3((2 * Co + BIAS_ROWS - 1) / BIAS_ROWS, (B + BIAS_BG - 1) / BIAS_BG), 256, smem, st>>>(p.gmean, Co, B, L.w_g2,
                                                                                                                     p.bias);
                LS_CHECK_LAUNCH("k_bias_rows");
            } else {
                k_bias_gemv<<<dim3((2 * Co + 7) / 8, B), 256, (size_t)Co * 3 * sizeof(float), st>>>(p.gmean, Co, L.w_g2, p.bias);
                LS_CHECK_LAUNCH("k_bias_gemv");
            }
            GemmArgs g{};
            g.W = L.w_g1;
            g.Wtc = L.w_g1_tc;
            g.R = 2 * Co;
            g.K = Co;
            g.ldw = Co;
            g.B = B;
            g.X = p.pooled;
            g.n_per_b = 3 * Nd;
            g.npts = Nd;
            g.x_sb = (long long)Co * 3 * Nd;
            g.x_sk = 3LL * Nd;
            g.out = p.raw;
            g.o_sb = 2LL * Co * 3 * Nd;
            g.o_sr = 3LL * Nd;
            g.bias = p.bias;
            g.bias_sb = 2LL * Co * 3;
            g.bias_sr = 3;
            g.bias_axis = 1;
            rc = launch_gemm(g, st);
            if (rc != LS_OK) return rc;
            k_vnact<<<dim3((Co * Nd + 255) / 256, B), 256, 0, st>>>(p.raw, Co, Nd, oms, layer_out);
            LS_CHECK_LAUNCH("k_vnact");
        }
        if (io->feat[i])
            LS_CHECK_CUDA(cudaMemcpyAsync(io->feat[i], layer_out, sizeof(float) * (size_t)B * Co * 3 * Nd,
                                          cudaMemcpyDeviceToDevice, st));
        src_f = layer_out;
        cur ^= 1;
    }

    if (fps_pending) LS_CHECK_CUDA(cudaStreamWaitEvent(st, sc->join_fps, 0));
    // ---- head ------------------------------------------------------------------------------
    {
        ProfScope ps(6, -1, st);
        const int last = d->num_layers - 1;
        const int Cl = d->layers[last].c_out, Nl = p.n_dst[last], C = d->c_dim;
        GemmArgs g{};
        g.W = d->w_conv_c;
        g.Wtc = d->w_conv_c_tc;
        g.R = C + 1;
        g.K = Cl;
        g.ldw = Cl;
        g.B = B;
        g.X = src_f;
        g.n_per_b = 3 * Nl;
        g.npts = Nl;
        g.x_sb = (long long)Cl * 3 * Nl;
        g.x_sk = 3LL * Nl;
        g.out = p.raw;
        g.o_sb = (long long)(C + 1) * 3 * Nl;
        g.o_sr = 3LL * Nl;
        rc = launch_gemm(g, st);
        if (rc != LS_OK) return rc;
        HeadArgs h{};
        h.raw = p.raw;
        h.c_dim = C;
        h.Nl = Nl;
        h.w_inv_t = d->w_inv_t;
        h.w_fc0_t = d->w_fc0_t;
        h.w_lin1 = d->w_lin1;
        h.w_short = d->w_short;
        h.w_act2 = d->w_act2;
        h.oms = oms;
        h.scale_factor = d->scale_factor;
        h.center_pred = d->center_pred;
        h.center_pred_scale = d->center_pred_scale;
        h.normalize = io->normalize;
        h.centroid = p.centroid;
        h.s0 = p.s0;
        h.center = io->center;
        h.scale = io->scale;
        h.z_so3 = io->z_so3;
        h.z_inv = io->z_inv;
        h.packed = io->packed;
        LS_REQUIRE(!io->packed || C == 256, "packed records assume c_dim == 256");
        k_head<<<B, C, (size_t)6 * C * sizeof(float), st>>>(h);
        LS_CHECK_LAUNCH("k_head");
    }
    return LS_OK;
}

int ls_tc_packed_floats(int32_t R, int32_t K, size_t* n_floats) {
    LS_REQUIRE(n_floats && R > 0 && K > 0, "bad arguments");
    *n_floats = tc_packed_floats(R, K);
    return LS_OK;
}
int ls_tc_pack_weights(const float* W, int32_t R, int32_t K, int32_t ldw, float* packed, void* stream) {
    return tc_pack_weights(W, R, K, ldw, packed, static_cast<cudaStream_t>(stream));
}
int ls_set_knn_tensor_cores(int32_t on, float kappa_scale) {
    ls::g_use_knn_tc = on != 0;
    ls::g_knn_tc_kappa_scale = kappa_scale > 0.f ? kappa_scale : 1.f;
    return LS_OK;
}
int ls_set_overlap(int32_t on) {
    ls::g_overlap = on != 0;
    return LS_OK;
}
int ls_set_tensor_cores(int32_t on) {
    ls::g_use_tensor_cores = on != 0;
    return LS_OK;
}
int ls_set_fps_fma(int32_t on) {
    ls::g_fps_fma = on != 0;
    return LS_OK;
}
int ls_set_wave_bytes(int64_t bytes) {
    LS_REQUIRE(bytes >= 0, "wave bytes must be >= 0");
    ls::g_wave_bytes = bytes;
    return LS_OK;
}
int ls_set_gemm_variant(int32_t variant) {
    LS_REQUIRE(variant >= 1 && variant <= 3,
               "gemm variant must be 1 (per-tile CTAs), 2 (persistent, operands in shared memory) or 3 (activations in tensor memory)");
    ls::g_gemm_variant = variant;
    return LS_OK;
}
int ls_vn_linear(const float* W, const float* packed, const float* X, float* out, int32_t R, int32_t K, int32_t ldw,
                 int32_t B, int32_t n, void* stream) {
    LS_REQUIRE(W && X && out && R > 0 && K > 0 && B > 0 && n > 0, "bad arguments");
    GemmArgs g{};
    g.W = W;
    g.Wtc = packed;
    g.R = R;
    g.K = K;
    g.ldw = ldw;
    g.B = B;
    g.n_per_b = n;
    g.X = X;
    g.x_sb = (long long)K * n;
    g.x_sk = n;
    g.out = out;
    g.o_sb = (long long)R * n;
    g.o_sr = n;
    return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int ls_knn(const float* query, const float* source, int32_t B, int32_t D, int32_t Nq, int32_t Ns,
           int64_t* idx, float* dist2, void* stream) {
    LS_REQUIRE(query && source && idx, "null pointer");
    LS_REQUIRE(B >= 1 && B <= 65535 && D >= 1 && Nq >= 1 && Ns >= LS_KNN_K, "bad sizes (need Ns >= 16)");
    EdgeArgs ea{};
    ea.src_f = source;
    ea.dst_f = query;
    ea.B = B;
    ea.D = D;
    ea.Ns = Ns;
    ea.Nd = Nq;
    ea.Co = 32;
    ea.idx_out = idx;
    ea.dist_out = dist2;
    if (Ns <= SMALL_NS) {
        if (g_knn_small_tiled)
            k_knn_small_tiled<<<dim3((Nq + 31) / 32, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(source, query, D, Ns, Nq, idx, dist2);
        else
            k_knn_small<<<dim3((Nq + 7) / 8, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(source, query, D, Ns, Nq, idx, dist2);
        LS_CHECK_LAUNCH("k_knn_small");
        return LS_OK;
    }
    ea.qpc = pick_qpc(B, Nq, false);
    return launch_edge(MODE_KNN_ONLY, ea, static_cast<cudaStream_t>(stream));
}

// the same graph through the tensor-core candidate filter + exact re-rank (what ls_encoder_forward runs)
static void knn_tc_plan(int B, int D, int Nq, int Ns, void* ws, float** img_s, float** nrm_s, float** pm_s, float** img_q,
                        float** nrm_q, float** pm_q, unsigned short** cand, int** cnt, float** cand_dt, float** e2,
                        size_t* bytes) {
    Carver c(ws);
    *img_s = c.take<float>((size_t)B * knn_tc_img_floats(Ns, D));
    *nrm_s = c.take<float>((size_t)B * knn_tc_nrm_floats(Ns));
    *pm_s = c.take<float>((size_t)B * knn_tc_pm_floats(Ns, D));
    *img_q = c.take<float>((size_t)B * knn_tc_img_floats(Nq, D));
    *nrm_q = c.take<float>((size_t)B * knn_tc_nrm_floats(Nq));
    *pm_q = c.take<float>((size_t)B * knn_tc_pm_floats(Nq, D));
    *cand = c.take<unsigned short>((size_t)B * knn_tc_cand_u16(Nq));
    *cnt = c.take<int>((size_t)B * knn_tc_nrm_floats(Nq));
    *cand_dt = c.take<float>((size_t)B * knn_tc_cand_u16(Nq));
    *e2 = c.take<float>((size_t)B * knn_tc_nrm_floats(Nq));
    *bytes = (c.off + 255) & ~size_t(255);
}
int ls_knn_tc_workspace_bytes(int32_t B, int32_t D, int32_t Nq, int32_t Ns, size_t* bytes) {
    LS_REQUIRE(bytes && B >= 1 && D >= 1 && Nq >= 1 && Ns >= 1, "bad arguments");
    float *a, *b, *c, *d, *e, *f, *i, *j;
    unsigned short* g;
    int* h;
    knn_tc_plan(B, D, Nq, Ns, nullptr, &a, &b, &c, &d, &e, &f, &g, &h, &i, &j, bytes);
    return LS_OK;
}
int ls_knn_tc(const float* query, const float* source, int32_t B, int32_t D, int32_t Nq, int32_t Ns, int64_t* idx,
              float* dist2, int32_t* n_candidates, void* workspace, size_t workspace_bytes, void* stream) {
    LS_REQUIRE(query && source && idx && workspace, "null pointer");
    LS_REQUIRE(B >= 1 && B <= 65535 && D >= 1 && Nq >= 1 && Ns >= LS_KNN_K && Ns <= 65535, "bad sizes (need 16 <= Ns <= 65535)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KnnTcArgs ka{};
    float *img_s, *nrm_s, *pm_s, *img_q, *nrm_q, *pm_q, *cand_dt, *e2;
    unsigned short* cand;
    int* cnt;
    size_t need;
    knn_tc_plan(B, D, Nq, Ns, workspace, &img_s, &nrm_s, &pm_s, &img_q, &nrm_q, &pm_q, &cand, &cnt, &cand_dt, &e2, &need);
    if (need > workspace_bytes) {
        set_error("workspace too small");
        return LS_ERR_WORKSPACE;
    }
    int rc = launch_knn_pack(source, B, D, Ns, img_s, nrm_s, pm_s, st);
    if (rc != LS_OK) return rc;
    rc = launch_knn_pack(query, B, D, Nq, img_q, nrm_q, pm_q, st);
    if (rc != LS_OK) return rc;
    ka.img_s = img_s;
    ka.nrm_s = nrm_s;
    ka.img_q = img_q;
    ka.nrm_q = nrm_q;
    ka.Ns = Ns;
    ka.Nd = Nq;
    ka.n_pt_s = knn_tc_tiles(Ns);
    ka.n_pt_q = knn_tc_tiles(Nq);
    ka.n_kb = knn_tc_kblocks(D);
    ka.kappa = g_knn_tc_kappa_scale * knn_tc_kappa(D);
    ka.cand = cand;
    ka.cand_dt = cand_dt;
    ka.cnt = cnt;
    ka.e2 = e2;
    rc = launch_knn_tc(ka, B, st);
    if (rc != LS_OK) return rc;
    if (n_candidates)
        LS_CHECK_CUDA(cudaMemcpy2DAsync(n_candidates, sizeof(int) * Nq, cnt, sizeof(int) * ka.n_pt_q * KT_PTS, sizeof(int) * Nq, B,
                                        cudaMemcpyDeviceToDevice, st));
    RerankArgs ra{};
    ra.cand = cand;
    ra.cand_dt = cand_dt;
    ra.cand_cnt = cnt;
    ra.e2 = e2;
    ra.all_exact = dist2 != nullptr;
    ra.pm_s = pm_s;
    ra.pm_q = pm_q;
    ra.Ns = Ns;
    ra.Nd = Nq;
    ra.Dp = ka.n_kb * KT_KB;
    ra.idx_out = idx;
    ra.dist_out = dist2;
    return launch_rerank(ra, B, st);
}

int ls_fps_workspace_bytes(int32_t B, int32_t N, size_t* bytes) {
    LS_REQUIRE(bytes && B >= 1 && N >= 1, "bad arguments");
    *bytes = N > 8192 ? (size_t)B * N * sizeof(float4) : 0;
    return LS_OK;
}

int ls_fps_ex(const float* xyz, int32_t B, int32_t N, int32_t n_out, const int64_t* start_idx, int64_t* idx,
              float* out_xyz, void* workspace, size_t workspace_bytes, void* stream) {
    LS_REQUIRE(xyz && idx, "null pointer");
    LS_REQUIRE(B >= 1 && N >= 1 && n_out >= 1 && n_out <= N, "fps: need 1 <= n_out <= N");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N > 8192) {
        LS_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)B * N * sizeof(float4),
                   "fps: N > 8192 needs the scratch of ls_fps_workspace_bytes");
        LS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "fps: workspace must be 16-byte aligned");
        k_fps_large<<<B, 1024, 0, st>>>(xyz, N, n_out, start_idx, static_cast<float4*>(workspace), idx, out_xyz, g_fps_fma);
        LS_CHECK_LAUNCH("k_fps_large");
        return LS_OK;
    }
    FpsArgs fa{};
    fa.xyz = xyz;
    fa.N = N;
    fa.n_levels = 1;
    fa.n_out[0] = n_out;
    fa.sel64[0] = idx;
    fa.out_xyz = out_xyz;
    fa.start = start_idx;
    return launch_fps(fa, B, st);
}

int ls_fps_masked(const float* xyz, const uint8_t* mask, int32_t B, int32_t Nmax, int32_t n_out, const int64_t* start_idx,
                  int64_t* idx, float* out_xyz, int32_t* n_valid, void* workspace, size_t workspace_bytes, void* stream) {
    LS_REQUIRE(xyz && mask && (idx || out_xyz), "null pointer");
    LS_REQUIRE(B >= 1 && Nmax >= 1 && n_out >= 1, "fps_masked: bad sizes");
    LS_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)B * Nmax * sizeof(float4),
               "fps_masked: workspace must hold B * Nmax float4");
    LS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "fps_masked: workspace must be 16-byte aligned");
    k_fps_masked<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(xyz, mask, Nmax, n_out, start_idx,
                                                                     static_cast<float4*>(workspace), n_valid, idx, out_xyz, g_fps_fma);
    LS_CHECK_LAUNCH("k_fps_masked");
    return LS_OK;
}

int ls_fps(const float* xyz, int32_t B, int32_t N, int32_t n_out, int64_t* idx, float* out_xyz, void* stream) {
    LS_REQUIRE(N <= 8192, "ls_fps: N > 8192 needs ls_fps_ex with a workspace");
    return ls_fps_ex(xyz, B, N, n_out, nullptr, idx, out_xyz, nullptr, 0, stream);
}

}  // extern "C"
