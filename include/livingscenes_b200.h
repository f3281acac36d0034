/*
 * livingscenes_b200.h -- C ABI of the B200-native LivingScenes inference hot path.
 *
 * The reference (GradientSpaces/LivingScenes) has NO FFI on this path: it is a plain
 * Python/PyTorch module API.  This header is therefore the boundary
 * the new build DEFINES; every entry point cites the reference Python interface it replaces
 * (paths relative to the reference root).  The Python package `livingscenes_b200` binds these
 * symbols with ctypes and re-exposes the reference's own names (VecDGCNN_att.forward,
 * Shape_Prior.encode/decoder, sequential_matcher, nn_matcher,
 * kabsch_transformation_estimation); INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (fp32 unless stated) unless the name ends in _host;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises the device, never allocates or frees caller memory; scratch memory is
 *    provided by the caller (see the *_workspace_bytes functions);
 *  - return value: 0 = ok, negative = error (LS_ERR_*); ls_last_error() gives the message of
 *    the last failing call on the calling thread;
 *  - indices at the boundary are int64 like the reference's (torch.long).
 */
#ifndef LIVINGSCENES_B200_H
#define LIVINGSCENES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LS_API __attribute__((visibility("default")))
#else
#define LS_API
#endif

#define LS_ABI_VERSION 3
#define LS_MAX_LAYERS 8
#define LS_KNN_K 16          /* num_knn of the shipped model (model_config.yaml:165) */
#define LS_HEAD_C 16         /* atten_multi_head_c (model_config.yaml:143)           */
#define LS_CODE_FLOATS 1028  /* packed embedding record: z_so3 768 | z_inv 256 | s 1 | t 3 */

#define LS_OK 0
#define LS_ERR_INVALID -1     /* bad argument / unsupported configuration */
#define LS_ERR_CUDA -2        /* a CUDA runtime call or launch failed     */
#define LS_ERR_WORKSPACE -3   /* workspace too small                      */

LS_API int ls_version(void);
LS_API const char* ls_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Encoder: VecDGCNN_att.forward  (lib_shape_prior/core/lib/vec_sim3/vec_dgcnn_atten.py:177-252)
 * and Shape_Prior.encode         (model_utils.py:165-197) when `normalize` != 0.
 *
 * Weights are passed pre-folded (fp64 products rounded to fp32, done once by the host):
 * with W = lin.weight = [W_a | W_b] (columns acting on nn-dst and on dst, vec_dgcnn_atten.py:160)
 * and Wd = act.lin_dir.weight (vec_layers.py:246), the per-edge VN-Linear pair
 *      q = W [nn-dst; dst],   k = Wd q
 * is evaluated as  q = (W_a src)[idx] + ((W_b-W_a) dst),  k = (Wd W_a src)[idx] + (Wd (W_b-W_a) dst)
 * i.e. two point-level GEMMs followed by a gather.
 * ------------------------------------------------------------------------------------------ */
typedef struct ls_enc_layer_desc {
    int32_t c_in;         /* channels of the incoming features (feat_dim[i-1]; 1 for layer 0)     */
    int32_t c_out;        /* feat_dim[i]; multiple of 32                                          */
    int32_t down_factor;  /* FPS down-sampling factor applied before this layer (1 = none)       */
    int32_t attention;    /* 0: mean pool over K (layers < atten_start_layer), 1: attention pool */
    int32_t global_conv;  /* 1: followed by the global-context VecLNA (use_res_global_conv)      */
    int32_t _pad;
    /* layer 0 (c_in == 1): [2][c_out][3] = { V.lin.weight , V.act.lin_dir.weight @ V.lin.weight } */
    const float* w0;
    /* layers >= 1, row blocks of c_out rows each, row-major [rows][c_in]:
     *   w_src: { Vq=W_a, Vk=(Wd W)_a [, Kq, Kk] }                        (2 or 4 blocks)
     *   w_dst: { Vq=W_b-W_a, Vk=(Wd W)_b-(Wd W)_a [, Kq, Kk, Qq=W_Q, Qk=Wd_Q W_Q] } (2 or 6 blocks) */
    const float* w_src;
    const float* w_dst;
    /* global conv: [2*c_out][c_out] each: { W_G[:, :c_out]; (Wd_G W_G)[:, :c_out] } and the same
     * for the [:, c_out:] columns that multiply the instance mean (vec_dgcnn_atten.py:222-225)   */
    const float* w_g1;
    const float* w_g2;
    /* optional (NULL = FP32 SIMT GEMM): w_src / w_dst / w_g1 packed by ls_tc_pack_weights for the
     * tcgen05 3xTF32 tensor-core GEMM */
    const float* w_src_tc;
    const float* w_dst_tc;
    const float* w_g1_tc;
} ls_enc_layer_desc;

typedef struct ls_encoder_desc {
    int32_t num_layers;
    int32_t c_dim;              /* 256 */
    int32_t center_pred;        /* fc_center present */
    int32_t center_pred_scale;  /* center *= scale_factor */
    float scale_factor;         /* 64000 */
    float neg_slope;            /* 0.2 */
    ls_enc_layer_desc layers[LS_MAX_LAYERS];
    const float* w_conv_c;   /* [c_dim+1][feat_last]: conv_c.lin.weight then the single shared
                                direction row conv_c.act.lin_dir.weight @ conv_c.lin.weight       */
    const float* w_inv_t;    /* [c_dim][c_dim]      fc_inv.weight TRANSPOSED ([in][out])          */
    const float* w_fc0_t;    /* [c_dim][2*(c_dim/2)] fc_center.fc0: {lin.weight, lin_dir@lin}^T   */
    const float* w_lin1;     /* [c_dim/2]  fc_center.lin1.weight                                  */
    const float* w_short;    /* [c_dim]    fc_center.shortcut.weight                              */
    float w_act2;            /* fc_center.act2.lin_dir.weight (1x1)                               */
    int32_t _pad;
    const float* w_conv_c_tc; /* optional: w_conv_c packed by ls_tc_pack_weights                          */
} ls_encoder_desc;

typedef struct ls_encoder_io {
    const float* x;        /* [B,3,N] input cloud (reference layout)                              */
    int32_t B, N;
    int32_t normalize;     /* 0: VecDGCNN_att.forward(x);  1: Shape_Prior.encode(x) semantics:
                              centroid removal, scale_0 = mean(top5(cdist)), t = center+centroid,
                              s = scale_0*scale (model_utils.py:171-195)                          */
    int32_t _pad;
    /* outputs (required) */
    float* center;         /* [B,3]   (normalize=1: "t")                                          */
    float* scale;          /* [B]     (normalize=1: "s")                                          */
    float* z_so3;          /* [B,c_dim,3]                                                         */
    float* z_inv;          /* [B,c_dim]                                                           */
    float* packed;         /* optional [B][LS_CODE_FLOATS] record for the embedding all-gather    */
    /* optional taps (NULL = not wanted) */
    int64_t* knn_idx[LS_MAX_LAYERS];  /* [B,Nd_l,16] ascending distance, ties -> lower index      */
    int64_t* fps_idx[LS_MAX_LAYERS];  /* [B,Nd_l] for layers with down_factor > 1                 */
    float* feat[LS_MAX_LAYERS];       /* [B,c_out_l,3,Nd_l] layer outputs                         */
    float* scale0;                    /* [B]   (normalize=1)                                      */
    float* x_norm;                    /* [B,3,N] (normalize=1)                                    */
    /* optional teacher forcing (NULL = compute): graph given by the caller                     */
    const int64_t* force_knn_idx[LS_MAX_LAYERS];
    const int64_t* force_fps_idx[LS_MAX_LAYERS];
} ls_encoder_io;

LS_API int ls_encoder_workspace_bytes(const ls_encoder_desc* desc, int32_t B, int32_t N, size_t* bytes);
LS_API int ls_encoder_forward(const ls_encoder_desc* desc, const ls_encoder_io* io,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Measurement hooks (no reference counterpart): when enabled, ls_encoder_forward brackets every
 * stage with CUDA events on the launching stream; after the caller has synchronised the stream,
 * ls_profile_read returns (stage, layer, milliseconds) of the LAST forward call.
 * stage: 0 normalize, 1 fps, 2 gather dst, 3 point-level GEMM tables, 4 fused kNN+EdgeConv+pool,
 *        5 global-context conv, 6 head, 7 tensor-core kNN filter (pack + filter), 8 kNN re-rank.  ls_kernel_launches counts every kernel this library launched. */
LS_API int ls_profile_enable(int32_t on);
LS_API int ls_profile_read(int32_t* stage, int32_t* layer, float* ms, int32_t max_entries, int32_t* n_entries);
LS_API int64_t ls_kernel_launches(void);

/* ------------------------------------------------------------------------------------------
 * VecLinear.forward (vec_layers.py:121-134) as a stand-alone op, and the weight packing of the
 * tensor-core path.  out[b][r][n] = sum_k W[r][k] X[b][k][n]  (X: [B,K,n], n = 3*points, channel-major).
 * `packed` = NULL runs the FP32 SIMT GEMM; otherwise the tcgen05 3xTF32 GEMM (fp32-accurate: every
 * operand is split into TF32 hi + lo parts, three MMAs accumulate in one fp32 TMEM tile).
 * ------------------------------------------------------------------------------------------ */
LS_API int ls_tc_packed_floats(int32_t R, int32_t K, size_t* n_floats);
LS_API int ls_tc_pack_weights(const float* W, int32_t R, int32_t K, int32_t ldw, float* packed, void* stream);
LS_API int ls_vn_linear(const float* W, const float* packed, const float* X, float* out, int32_t R, int32_t K,
                        int32_t ldw, int32_t B, int32_t n, void* stream);
/* kNN graph: 1 (default) = tcgen05 3xTF32 candidate filter + exact fp32 re-rank, 0 = exact fp32 brute force.
 * Both produce the same indices; kappa_scale (> 0, default 1) scales the filter's error budget (tests use a
 * huge value to force the candidate-overflow fallback). */
LS_API int ls_set_knn_tensor_cores(int32_t on, float kappa_scale);
/* 1 (default): ls_encoder_forward runs independent stages (FPS chain, point-level GEMM tables) on a library-owned
 * side stream next to the kNN chain (fork/join with events; capturable).  0: everything on the caller's stream. */
LS_API int ls_set_overlap(int32_t on);
LS_API int ls_set_tensor_cores(int32_t on);   /* 1 (default): use the tcgen05 path where packed weights exist */
/* 3 (default): tcgen05 GEMM with the activations in tensor memory (TS-form MMAs, 256-row weight tiles for K >= 128);
 * 2: persistent warp-specialised kernel with both operands in shared memory (bulk-TMA fed, double-buffered TMEM);
 * 1: the round-1 one-CTA-per-tile kernel (2 and 1 kept for A/B measurements).  Process-global, like the other
 * ls_set_* switches. */
LS_API int ls_set_gemm_variant(int32_t variant);
/* FPS squared distance: 0 (default) dx*dx + dy*dy + dz*dz with every product and sum rounded to fp32 (pytorch3d's
 * CPU path, the oracle and the committed fixtures); 1: fma(dz,dz,fma(dy,dy,dx*dx)), what nvcc's default contraction
 * makes of pytorch3d's CUDA kernel `dist2 += diff*diff`.  They differ in the last bit; only a near-tie arg-max can
 * tell them apart.  Which one the reference's pytorch3d 0.7.4 binary runs is unpinned (not vendored). */
LS_API int ls_set_fps_fma(int32_t on);
/* Table bytes per wave of the per-layer {point-level table GEMMs -> EdgeConv} schedule (default 0 = off, env LS_WAVE_MB; measured slower on B200, see profiles/r02):
 * a layer's batch is processed in waves of that many bytes of gather tables, two table slots alternating, so that
 * the tables are consumed out of L2 instead of HBM.  0: one launch per layer for the whole batch.  Results do not
 * depend on the setting. */
LS_API int ls_set_wave_bytes(int64_t bytes);

/* ------------------------------------------------------------------------------------------
 * Stand-alone graph ops (the pytorch3d boundary of the reference)
 * ------------------------------------------------------------------------------------------ */
/* pytorch3d.ops.knn_points as called at vec_dgcnn_atten.py:139: K = 16 nearest of every query
 * among the sources in D-dim feature space, squared L2 accumulated in fp32 over d = 0..D-1.
 * query [B,D,Nq], source [B,D,Ns] (channel-major, i.e. the reshape(B, C*3, N) view of :138).
 * idx [B,Nq,16] int64; dist2 optional [B,Nq,16]. */
LS_API int ls_knn(const float* query, const float* source, int32_t B, int32_t D, int32_t Nq, int32_t Ns,
           int64_t* idx, float* dist2, void* stream);

/* The same graph through the path ls_encoder_forward uses for Ns > 128: tcgen05 3xTF32 candidate filter
 * (ranking value |s|^2 - 2<q,s>, threshold = 16th smallest group minimum + error budget) followed by the exact
 * fp32 direct-form re-rank of the candidates; a query whose candidate list overflows is brute-forced exactly.
 * Results are identical to ls_knn.  n_candidates: optional [B,Nq] int32, candidates re-ranked per query
 * (-1 = overflow -> brute force).  workspace: ls_knn_tc_workspace_bytes. */
LS_API int ls_knn_tc_workspace_bytes(int32_t B, int32_t D, int32_t Nq, int32_t Ns, size_t* bytes);
LS_API int ls_knn_tc(const float* query, const float* source, int32_t B, int32_t D, int32_t Nq, int32_t Ns,
              int64_t* idx, float* dist2, int32_t* n_candidates, void* workspace, size_t workspace_bytes,
              void* stream);

/* pytorch3d.ops.sample_farthest_points(points, K=n_out), random_start_point=False
 * (vec_dgcnn_atten.py:169; model_utils.py:205; more_solver.py:67,107-108).
 * xyz [B,3,N] -> idx [B,n_out] int64, optional out_xyz [B,3,n_out]. */
LS_API int ls_fps(const float* xyz, int32_t B, int32_t N, int32_t n_out, int64_t* idx, float* out_xyz,
           void* stream);
/* ls_fps with an optional per-instance first index (start_idx [B] int64 on the device, NULL = 0: pytorch3d's
 * random_start_point=True picks it at random, model_utils.py:202-205) and support for large clouds: N > 8192
 * (scene instances of up to ~10^5 points) runs from a caller-provided scratch of ls_fps_workspace_bytes. */
LS_API int ls_fps_workspace_bytes(int32_t B, int32_t N, size_t* bytes);
LS_API int ls_fps_ex(const float* xyz, int32_t B, int32_t N, int32_t n_out, const int64_t* start_idx, int64_t* idx,
              float* out_xyz, void* workspace, size_t workspace_bytes, void* stream);
/* FPS of a whole batch of ragged, masked instances in ONE launch (Shape_Prior.encode_fps, model_utils.py:199-205:
 * valid_pc = pc.T[mask]; fps(valid_pc, K)): xyz [B,3,Nmax], mask [B,Nmax] bytes (non-zero = valid).  The valid
 * points of every instance are compacted in order into the workspace (B*Nmax float4) and sampled there; idx
 * (optional [B,n_out]) refers to the compacted list like pytorch3d's; out_xyz (optional) [B,3,n_out]; n_valid
 * (optional [B]) receives the number of valid points -- callers must check n_valid >= n_out.  start_idx as ls_fps_ex. */
LS_API int ls_fps_masked(const float* xyz, const uint8_t* mask, int32_t B, int32_t Nmax, int32_t n_out,
                  const int64_t* start_idx, int64_t* idx, float* out_xyz, int32_t* n_valid, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Matching: lib_more/matcher_new.py
 * ------------------------------------------------------------------------------------------ */
/* sequential_matcher(m0[n,dim], m1[m,dim]) (matcher_new.py:109-139) for `n_pairs` independent
 * scene pairs: pair p uses rows [off0[p], off0[p+1]) of z0 and [off1[p], off1[p+1]) of z1.
 * matches0 / matches1 are int64, -1 = unmatched, indices local to the pair.
 * off0_host / off1_host are HOST arrays of n_pairs+1 ints.  scratch: ls_match_workspace_bytes. */
LS_API int ls_match_workspace_bytes(const int32_t* off0_host, const int32_t* off1_host, int32_t n_pairs,
                             size_t* bytes);
LS_API int ls_seq_match(const float* z0, const float* z1, int32_t dim, const int32_t* off0_host,
                 const int32_t* off1_host, int32_t n_pairs, int64_t* matches0, int64_t* matches1,
                 void* workspace, size_t workspace_bytes, void* stream);
/* nn_matcher (matcher_new.py:85-105): cosine top-1 both ways + mutual check; same batching. */
LS_API int ls_mutual_nn(const float* z0, const float* z1, int32_t dim, const int32_t* off0_host,
                 const int32_t* off1_host, int32_t n_pairs, int64_t* matches0, int64_t* matches1,
                 void* workspace, size_t workspace_bytes, void* stream);
/* sim3_seq_matcher / eq_seq_matcher (matcher_new.py:142-230) for ONE scene pair: the greedy rounds of ls_seq_match on
 * score = cos / (res + 1e-5) (score_mode 1) or 1 / (res + 1e-5) (score_mode 2); res [n,m] = mean Kabsch residual of
 * the equivariant codes of every (src, tgt) pair (ls_kabsch_batched on the n*m pairs).  Workspace as ls_seq_match. */
LS_API int ls_seq_match_scored(const float* z0, const float* z1, int32_t dim, int32_t n, int32_t m, const float* res,
                        int32_t score_mode, int64_t* matches0, int64_t* matches1, void* workspace,
                        size_t workspace_bytes, void* stream);
/* sinkhorn_matcher (matcher_new.py:11-71) for ONE scene pair: cosine scores / sqrt(dim), log-space optimal transport
 * with a dustbin (alpha), `iters` iterations (reference: 100, alpha 1, threshold 0), mutual arg-max + exp(score) >
 * match_threshold.  n, m <= ~230 (the coupling matrix lives in shared memory).  Workspace as ls_seq_match. */
LS_API int ls_sinkhorn_match(const float* z0, const float* z1, int32_t dim, int32_t n, int32_t m, int32_t iters,
                      float alpha, float match_threshold, int64_t* matches0, int64_t* matches1, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pose: kabsch_transformation_estimation (lib_more/pose_estimation.py:29-121)
 * x1, x2 [b,n,3]; weights optional [b,n] (NULL = ones); normalize_w, eps as in the reference.
 * R [b,3,3], t [b,3] (the reference's [b,3,1]), res [b,n] residual norms.
 * ------------------------------------------------------------------------------------------ */
LS_API int ls_kabsch_batched(const float* x1, const float* x2, const float* weights, int32_t b, int32_t n,
                      int32_t normalize_w, float eps, float* R, float* t, float* res, void* stream);
/* Pose fit straight from two embedding sets as more_solver.py:114-116 does: x1 = z_so3_a[i] + t_a[i],
 * x2 = z_so3_b[match[i]] + t_b[match[i]] (pairs with match < 0 get identity / zeros). */
LS_API int ls_kabsch_from_codes(const float* z_so3_a, const float* t_a, const float* z_so3_b, const float* t_b,
                         const int64_t* match, int32_t n_pairs, int32_t c_dim, float* R, float* t,
                         float* res, void* stream);

/* ICP refinement after the pose fit (more_solver.py:182-187: pytorch3d.ops.iterative_closest_point(pc1, pc2,
 * init_transform=SimilarityTransform(R^T, t, 1)), estimate_scale=False, allow_reflection=False).  Row-vector
 * convention like pytorch3d: Xt = X R + T.  X [B,N,3] (N <= 4096), Y [B,M,3] (M <= 12288), R0 [B,3,3] / T0 [B,3]
 * optional initial transform; outputs R [B,3,3], T [B,3], rmse [B], n_iter [B] (iterations executed, negated when
 * the relative-rmse test never fired within max_iterations), optional Xt [B,N,3].  One CTA per pair. */
LS_API int ls_icp(const float* X, const float* Y, int32_t B, int32_t N, int32_t M, const float* R0, const float* T0,
           int32_t max_iterations, float relative_rmse_thr, float* R, float* T, float* rmse, int32_t* n_iter, float* Xt,
           void* stream);

/* ------------------------------------------------------------------------------------------
 * SDF query: FieldWrapper.forward (model_utils.py:230-263, inner_deepsdf branch) +
 * DeepSDF_Decoder.forward (lib_shape_prior/core/lib/implicit_func/deepsdf_decoder.py:78-123)
 * Weight-norm is pre-folded by the host (W = g * v / |v|_row).
 * ------------------------------------------------------------------------------------------ */
typedef struct ls_decoder_desc {
    int32_t latent;        /* 256 */
    int32_t hidden;        /* 768 */
    int32_t n_layers;      /* 9   */
    int32_t latent_in;     /* 4: layer whose input is cat[h, u]                                   */
    const float* w[12];    /* effective weights, row-major [out][in_padded]; in_padded = in rounded
                              up to a multiple of 8 (zero filled); layer 0 only holds the columns of
                              [inner, |q|] (257 -> 264); layer latent_in holds [h(255) | inner,|q| (257)] */
    const float* b[12];    /* biases [out]                                                       */
    const float* w0_zinv;  /* [hidden][latent] columns of layer 0 that multiply z_inv             */
    const float* w4_zinv;  /* [hidden][latent] columns of layer latent_in that multiply z_inv     */
    int32_t out_dims[12];  /* 768,768,768,255,768,768,768,768,1                                   */
    int32_t in_dims[12];   /* K of each GEMM (un-padded): 257,768,768,768,512,768,768,768,768     */
    const float* w_tc[12]; /* optional: w[l] packed by ls_tc_pack_weights (layers 0..7)                  */
    /* backward only (ls_sdf_backward; may be NULL for inference): transposed effective weights, row-major
       [in][out_padded8]: wt[l] for l = 0 (the [inner,|q|] columns: [257][768]), 1,2,3 ([768][255->256]),5,6,7;
       layer latent_in split into its h rows (wt4_h [255][768]) and its [inner,|q|] rows (wt4_u [257][768]);
       *_tc = the same packed by ls_tc_pack_weights (optional) */
    const float* wt[12];
    const float* wt_tc[12];
    const float* wt4_h;
    const float* wt4_u;
    const float* wt4_h_tc;
    const float* wt4_u_tc;
} ls_decoder_desc;

LS_API int ls_sdf_workspace_bytes(const ls_decoder_desc* desc, int32_t B, int32_t M, size_t* bytes);
/* query [B,M,3] world coordinates; codes z_so3 [B,latent,3], z_inv [B,latent], s [B], t [B,3];
 * sdf [B,M] (tanh output; occupancy logits of the reference are -sdf). */
LS_API int ls_sdf_decode(const ls_decoder_desc* desc, const float* query, const float* z_so3,
                  const float* z_inv, const float* s, const float* t, int32_t B, int32_t M,
                  float* sdf, void* workspace, size_t workspace_bytes, void* stream);


/* Gradients of the SDF query (the reference back-propagates through FieldWrapper / DeepSDF_Decoder with autograd in
 * More_Solver._solve_pairwise_registration(optim=True), more_solver.py:153-158, and _optimize_code, :210-214).
 * grad_sdf [B,M] = dLoss/dsdf.  Outputs (each may be NULL): grad_query [B,M,3], grad_z_so3 [B,latent,3],
 * grad_z_inv [B,latent], grad_s [B], grad_t [B,3] -- written, not accumulated.  B * M <= 131072 per call. */
LS_API int ls_sdf_backward_workspace_bytes(const ls_decoder_desc* desc, int32_t B, int32_t M, size_t* bytes);
LS_API int ls_sdf_backward(const ls_decoder_desc* desc, const float* query, const float* z_so3, const float* z_inv,
                    const float* s, const float* t, int32_t B, int32_t M, const float* grad_sdf,
                    float* grad_query, float* grad_z_so3, float* grad_z_inv, float* grad_s, float* grad_t,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mesh extraction: MISE octree refinement + marching cubes
 *   mesh_extractor2.py:88-131,158-181; utils/libmise/mise.pyx; utils/libmcubes
 * R = resolution0 << depth.  Caller-owned device arrays: state u8 [(R+1)^3] (0 absent, 1 unknown, 2 known),
 * val f32 [(R+1)^3], level u8 [R^3], pos / neg u8 [R^3].  One refinement round:
 *   ls_mise_collect (query list + count) -> ls_mise_points (canonical coordinates box*(p/R-0.5)) -> SDF query ->
 *   ls_mise_update (values * value_scale stored, active leaf voxels subdivided, new lattice points created)
 * until the count is 0; ls_mise_to_dense completes the (R+1)^3 grid like MISE.to_dense.
 * ------------------------------------------------------------------------------------------ */
LS_API int ls_mise_init(int32_t resolution0, int32_t depth, uint8_t* state, uint8_t* level, void* stream);
LS_API int ls_mise_collect(const uint8_t* state, int32_t R, int32_t* list, int32_t capacity, int32_t* count, void* stream);
LS_API int ls_mise_points(const int32_t* list, int32_t n, int32_t R, float box_size, float* query, void* stream);
LS_API int ls_mise_update(const int32_t* list, const float* values, int32_t n, float value_scale, int32_t R,
                   int32_t depth, float threshold, float* val, uint8_t* state, uint8_t* level, uint8_t* pos,
                   uint8_t* neg, void* stream);
LS_API int ls_mise_to_dense(const uint8_t* state, const float* val, int32_t R, float* dense, void* stream);
/* Marching cubes on grid [n,n,n] padded with -1e6 (mesh_extractor2.py:172), corner test value <= iso, vertices in the
 * extractor's final frame box*((c-1)/(n-1)-0.5).  ls_mcubes_count writes {#vertices, #triangles} to the device ints
 * n_out[2]; after reading them the caller allocates vertices [V,3] fp32 / faces [F,3] int64 and calls ls_mcubes_emit
 * with the same workspace. */
LS_API int ls_mcubes_workspace_bytes(int32_t n, size_t* bytes);
LS_API int ls_mcubes_count(const float* grid, int32_t n, float iso, void* workspace, size_t workspace_bytes,
                    int32_t* n_out, void* stream);
LS_API int ls_mcubes_emit(const float* grid, int32_t n, float iso, float box_size, const void* workspace,
                   float* vertices, int64_t* faces, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIVINGSCENES_B200_H */
