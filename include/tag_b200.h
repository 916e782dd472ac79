/*
 * tag_b200.h — C ABI of the B200-native TAG scoring hot path (libtag_b200.so).
 *
 * The reference (XThomasBU/video-gen-evals) is pure Python and has no FFI; the boundary it
 * exposes is its Python call signatures. Each entry point below replaces the
 * reference code cited next to it; the Python shim in `video-gen-evals_b200/` keeps the
 * reference's names/arguments and calls these through ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer on the handle's device
 *    unless marked HOST. The caller (PyTorch) owns all tensors; the library owns only the
 *    opaque handle (packed weights + workspace allocated in tag_finalize_weights).
 *  - all work is enqueued asynchronously on the cudaStream_t passed as `void* stream`.
 *  - return 0 on success, a TAG_ERR_* code otherwise; tag_last_error() gives the message.
 *    No exceptions or aborts cross the ABI.
 *  - ONE IN-FLIGHT CALL PER HANDLE: a handle owns mutable state (workspace, z-score tables, error string, launch
 *    counter), so calls on one handle must be issued from one thread and onto one stream at a time (calls on the
 *    same stream serialise by themselves). Use one handle per stream / thread for concurrency.
 *  - tensor-core mode (TAG_PRECISION_FP16_TC) tiles windows into 128-row blocks: clip lengths T must divide 128 or be
 *    a multiple of 128 (checked up front, TAG_ERR_UNSUPPORTED); TAG_PRECISION_FP32 takes any T.
 *  - all floating point tensors are fp32, row-major, densely packed.
 */
#ifndef TAG_B200_H
#define TAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAG_MAX_MODALITIES 8
#define TAG_D_MODEL 256

enum {
  TAG_OK = 0,
  TAG_ERR_INVALID = 1,      /* bad argument / unsupported configuration */
  TAG_ERR_CUDA = 2,         /* a CUDA runtime/driver call failed        */
  TAG_ERR_STATE = 3,        /* call order (weights not finalized, ...)  */
  TAG_ERR_MISSING = 4,      /* a required weight was never loaded       */
  TAG_ERR_UNSUPPORTED = 5   /* device is not sm_100 / shape not covered */
};

/* how a modality's frame-to-frame delta is formed (reference utils.py:142-217, :455-470) */
enum {
  TAG_KIND_COSINE = 0,      /* vit / clip / dino: L2-normalise, first difference  (_vit_delta  :142-147) */
  TAG_KIND_ROTMAT = 1,      /* global / pose: J 3x3 matrices -> log(R_prev^T R)   (_rotmat_delta :165-174) */
  TAG_KIND_PLAIN = 2,       /* betas: first difference                             (_betas_delta :161-163) */
  TAG_KIND_PROCRUSTES = 3   /* kp2d: centre, scale, rotate-align, difference       (_procrustes_kp_delta :177-217) */
};

enum {
  TAG_PRECISION_FP32 = 0,   /* fp32 CUDA-core GEMMs: reference-precision mode                     */
  TAG_PRECISION_FP16_TC = 1 /* fp16 operands, fp32 accumulate on tcgen05 tensor cores (default)   */
};

typedef struct tag_handle tag_handle;

/* mirrors HumanActionScorer.__init__ (reference model.py:103-146) */
typedef struct tag_config {
  int32_t n_modalities;                       /* M, in concat order (utils.py:496-510)          */
  int32_t raw_dims[TAG_MAX_MODALITIES];       /* dims_map_raw  values                           */
  int32_t diff_dims[TAG_MAX_MODALITIES];      /* dims_map_diff values (0 = no motion encoder)   */
  int32_t kinds[TAG_MAX_MODALITIES];          /* TAG_KIND_* per modality (used by feature fuse) */
  int32_t d_model;                            /* 256 (only value supported)                     */
  int32_t n_heads;                            /* 8   (head_dim must be 32)                      */
  int32_t n_layers;                           /* temporal transformer layers (4)                */
  int32_t ffn_dim;                            /* 4*d_model                                      */
  int32_t n_blocks;                           /* conv blocks per encoder, dilation 2^b (4)      */
  int32_t conv_kernel;                        /* 5                                              */
  int32_t precision;                          /* TAG_PRECISION_*                                */
  int32_t max_windows;                        /* windows per internal pass (workspace size)     */
  int32_t max_T;                              /* largest clip_len that will be passed           */
  int32_t device;                             /* CUDA device ordinal                            */
} tag_config;

/* per-modality packed frame arrays of V videos (what extract_mesh.py:25-44 / DWpose emit),
 * frames of video v are rows frame_offset[v] .. frame_offset[v+1]-1 of every array */
typedef struct tag_videos {
  const float* src[TAG_MAX_MODALITIES];       /* [F, raw_dims[m]] each                          */
  const int64_t* frame_offset;                /* [V+1]                                          */
  int64_t n_videos;
} tag_videos;

/* --- lifetime ---------------------------------------------------------------------------- */
int tag_create(tag_handle** out, const tag_config* cfg);
void tag_destroy(tag_handle* h);
const char* tag_last_error(const tag_handle* h);   /* h may be NULL: last error of tag_create */
int tag_abi_version(void);

/* --- weights: reference eval.py:136-165 `load_model` / state_dict key contract -------------
 * `key` is the reference state_dict key (e.g. "state_enc.vit.blocks.0.conv1.weight", with the
 * modality NAME replaced by its index: "state_enc.0.blocks..."); data may be HOST or DEVICE. */
int tag_load_weight(tag_handle* h, const char* key, const float* data, const int64_t* shape, int32_t ndim);
int tag_finalize_weights(tag_handle* h);
/* new parameter VALUES for a finalized handle (same architecture): tag_reload_weights_begin, then tag_load_weight for every
 * tensor, then tag_finalize_weights again — only the packed weights are rebuilt, the workspace (GBs) is kept. This is what a
 * training loop calls between optimiser steps before the next embedding pass (train.py:488-509). */
int tag_reload_weights_begin(tag_handle* h);

/* --- K1 feature fuse: WindowDataset._try_one compute part (utils.py:366-381, :396-404,
 *     :455-514): slice/pad, raw flatten, per-window deltas, z-score, concat [raw || diff].
 *     mean/stdv: [D] in feats column order, NULL = no normalisation (stats=None path).
 *     feats_out [n_windows, T, D]; flags_out[0] += #frames with det(H) < 0 (Procrustes
 *     reflection regime, not reproducible in closed form — SURVEY.md §8a A6), nullable. */
int tag_feature_fuse(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv,
                     const int32_t* win_video, const int32_t* win_start, int64_t n_windows, int32_t T,
                     float* feats_out, int32_t* flags_out, void* stream);

/* --- K2 encoder: HumanActionScorer.forward (model.py:162-193). feats [n_windows, T, D].
 *     outputs (any may be NULL except seq_embed): seq_embed [N,256], frame_embeds [N,T+1,256],
 *     tokens [N,T+1,256]; tc_window [N] = per-window temporal coherence of eval.py:218-224
 *     (fused, so frame_embeds need not leave the device). */
int tag_encode(tag_handle* h, const float* feats, int64_t n_windows, int32_t T,
               float* seq_embed, float* frame_embeds, float* tokens, float* tc_window, void* stream);
/* model.py:94, :185 `last_attn`: the next tag_encode call (only that one) also writes the fusion softmax over modalities,
 * attn [n_windows*T, M] fp32. NULL switches it off again. */
int tag_set_fusion_attn_out(tag_handle* h, float* attn);

/* --- K1+K2 fused over windows of resident videos: eval.py:168-206 `extract_window_features`
 *     without materialising feats for all windows (internally chunked by max_windows). */
int tag_encode_windows(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv,
                       const int32_t* win_video, const int32_t* win_start, int64_t n_windows, int32_t T,
                       float* seq_embed, float* frame_embeds, float* tokens, float* tc_window,
                       int32_t* flags_out, void* stream);

/* --- the same for clips of EQUAL length: every video has exactly L frames (frame_offset[v+1] - frame_offset[v] == L) and the
 *     windows are the reference's regular grid, starts 0, stride, 2*stride, ... <= L - T (eval.py:358-359 with
 *     WindowDataset utils.py:343-365), window-major by video: n_windows = n_videos * ((L - T) / stride + 1). Equivalent to
 *     tag_encode_windows on that window table. In tensor-core mode with T a power of two in 16..128 the library builds the
 *     features ONCE PER SOURCE FRAME (overlapping windows share frames) and lets the stem GEMMs gather the windows. */
int tag_encode_clips(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv, int64_t n_videos,
                     int32_t L, int32_t T, int32_t stride, float* seq_embed, float* frame_embeds, float* tokens,
                     float* tc_window, int32_t* flags_out, void* stream);

/* --- K3 centroids: build_train_centroids_subset (utils.py:1035-1043).
 *     sums_counts [C, 257] (256 sums || count) is accumulated INTO (zero it first); it is the
 *     buffer that is all-reduced across ranks before tag_centroid_finalize. labels int32 [n];
 *     labels outside [0, C) are ignored. */
int tag_centroid_accumulate(tag_handle* h, const float* z, const int32_t* labels, int64_t n, int32_t C,
                            float* sums_counts, void* stream);
int tag_centroid_finalize(tag_handle* h, const float* sums_counts, int32_t C, float* centroids,
                          float* counts, void* stream);

/* --- K4 scores: compute_action_consistency_scores (eval.py:235-255) and the per-video mean of
 *     compute_temporal_coherence_scores (eval.py:226). Windows of video v are rows
 *     seg_offsets[v] .. seg_offsets[v+1]-1. video_label outside [0, C) => ac_out = NaN (the
 *     reference silently skips such videos, eval.py:247-251). tc_out of a video with no
 *     windows (or T < 2) = NaN. tc_window may be NULL (then tc_out is not written); seq_embeds may
 *     be NULL (then only the TC aggregation runs and video_label / centroids / ac_out are ignored). */
int tag_score(tag_handle* h, const float* seq_embeds, const float* tc_window, const int64_t* seg_offsets,
              const int32_t* video_label, const float* centroids, int32_t C, int64_t n_videos,
              float* ac_out, float* tc_out, void* stream);

/* --- per-window temporal coherence from ALREADY NORMALISED frame embeddings [N, S, 256] (S = T+1,
 *     row 0 = CLS): mean_t ||f_{t+1} - f_t||_2 over the T-1 frame pairs (eval.py:218-224); NaN when
 *     the window has fewer than 2 frames (the reference skips those). */
int tag_window_tc(tag_handle* h, const float* frame_embeds, int64_t n_windows, int32_t S, float* tc_window, void* stream);

/* --- N2 stats: per-column float64 sum / sum-of-squares of a [rows, D] fp32 matrix, accumulated
 *     INTO sum/sumsq (compute_stats_from_npz, utils.py:589-593). */
int tag_stats_accumulate(tag_handle* h, const float* x, int64_t rows, int32_t D, double* sum, double* sumsq,
                         void* stream);

/* --- N1 TCL forward (losses.py:14-34): per-row loss terms of the [B,B] similarity matrix in
 *     one pass; loss_rows [B] (mean over rows = the reference's scalar). targets int32 [B]. */
int tag_tcl_forward(tag_handle* h, const float* z, const int32_t* targets, int64_t B, float temperature,
                    float k1, float k2, float* loss_rows, void* stream);
/*     For B >= 512 the similarity matrix Z Z^T is a tcgen05 GEMM (split-fp16 operands hi.hi + lo.hi + hi.lo, fp32
 *     accumulate: ~1e-7 on S) whose epilogue keeps the masked row sums; the [B,B] matrix never exists in memory.
 *     targets must not contain INT32_MIN. */

/* --- N1 SupConWithHardNegatives forward (losses.py:37-56): loss_rows [B] = CrossEntropy([a.p/t, a.h/t], 0) per sample
 *     (mean over rows = the reference's scalar); anchor / positive / hard_negative [B,256]. */
int tag_supcon_hard_forward(tag_handle* h, const float* anchor, const float* positive, const float* hard_negative,
                            int64_t B, float temperature, float* loss_rows, void* stream);

/* --- N1 hard-negative augmentations (utils.py:65-95: partial_shuffle_within_window, reverse_sequence, get_static_window)
 *     are frame gathers: out[b, t, :] = x[b, idx[b*T + t], :], x / out [B, T, D] fp32 (D % 4 == 0), idx int32 [B*T]. */
int tag_gather_frames(tag_handle* h, const float* x, const int32_t* idx, int64_t B, int32_t T, int32_t D, float* out,
                      void* stream);

/* --- introspection used by bench.py: kernels launched through this handle since creation; with
 *     profiling on, every encoder kernel is bracketed by CUDA events on the caller's stream and
 *     out12 = {other: ms, -, launches | conv GEMM: ms, flops, launches | other GEMM: ms, flops, launches |
 *     feature fuse: ms, algorithmic bytes, launches}, accumulated since tag_set_profiling(h, 1). */
int64_t tag_launch_count(const tag_handle* h);
int tag_set_profiling(tag_handle* h, int32_t on);
int tag_get_profile(tag_handle* h, double* out12);
/* the same split finer: out[3*k .. 3*k+2] = {ms, work, launches} of kind k < n_kinds, k = 0 other, 1 conv GEMM, 2 other GEMM,
 * 3 feature fuse (K1), 4 merge-fusion (A11 + A12 front half), 5 finalize (A14 + fused per-window TC), 6 attention,
 * 7 build-tokens (A8); work = FLOPs for the GEMM kinds, ALGORITHMIC BYTES for the bandwidth-bound kinds. */
int tag_get_profile_kinds(tag_handle* h, double* out, int32_t n_kinds);

/* --- test hooks: the two GEMM kernels in isolation (tests/test_gemm_gpu.py).
 *     C = act(sum_taps A[row+shift] W^T + bias + res); fp32: W [N, ldw]; tensor-core: A/W/res16/C16
 *     are fp16, W [N, taps*K], ld of res/C = N; gn_gamma/gn_beta non-NULL fuse GroupNorm(1,256) over each
 *     (T x 256) window after the activation (conv only, N == 256, T a power of two <= 128); ln_gamma/ln_beta
 *     non-NULL fuse LayerNorm over the 256 output columns after bias + fp32 residual (plain GEMM, N == 256; writes
 *     C32, which may alias res32, and its fp16 copy C16). */
int tag_debug_gemm_f32(tag_handle* h, const float* A, int32_t lda, const float* W, int32_t ldw, int64_t M, int32_t N,
                       int32_t K, int32_t taps, int32_t dil, int32_t T, const float* bias, const float* res, float* C,
                       int32_t act, void* stream);
/* K1 in its tensor-core-path form: the padded fp16 operand [n_windows*T, D16] the encoder's stem GEMMs read (every
 * modality block padded to a multiple of 64 columns: raw blocks in modality order, then diff blocks; pad columns zero).
 * This is what tag_encode_windows builds internally; exposed so the tests can compare it with tag_feature_fuse.
 * Returns D16 through *d16_out (call with feats16_out == NULL to query it). */
int tag_debug_feature_fuse16(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv,
                             const int32_t* win_video, const int32_t* win_start, int64_t n_windows, int32_t T,
                             void* feats16_out, int32_t* d16_out, int32_t* flags_out, void* stream);
int tag_debug_gemm_tc(tag_handle* h, const void* A, int32_t lda, const void* W, int64_t M, int32_t N, int32_t K,
                      int32_t taps, int32_t dil, int32_t T, const float* bias, const void* res16, const float* res32,
                      void* C16, float* C32, int32_t act, const float* gn_gamma, const float* gn_beta,
                      const float* ln_gamma, const float* ln_beta, void* stream);

/* test hook: fill every workspace buffer of the handle (activations, token stream, fp16 operand tables, staging rows) with
 * 0xFF bytes = NaN in fp16 and fp32, so that a kernel that reads something it did not write this call shows up as NaN. */
int tag_debug_poison_workspace(tag_handle* h, void* stream);

/* the fused tail of one transformer layer (model.py:145, post-norm): x32 <- LN2(x1 + relu(x1 W1^T + b1) W2^T + b2) with
 * x1 = LN1(x32 + att16 Wo^T + bo), in place, plus the fp16 copy x16; M > 128 rows, ffn_dim a multiple of 256; weights fp16
 * [out, in] row-major. This is the kernel tag_encode* runs once per layer in tensor-core mode. */
int tag_debug_tlayer_tail(tag_handle* h, const void* att16, float* x32, void* x16, int64_t M, int32_t ffn_dim, const void* Wo16,
                          const void* W1_16, const void* W2_16, const float* bo, const float* b1, const float* b2,
                          const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b, void* stream);

/* one fused TemporalConvBlock (model.py:22-41) in tensor-core mode: h16 <- GroupNorm(GELU(conv2(GELU(conv1(h16))) + h16)), in place,
 * [M, 256] fp16 rows of whole windows (M = windows * T); weights fp16 [256, 5 * 256], tap-major K; no conv bias (as the reference).
 * Returns TAG_ERR_UNSUPPORTED where the two-kernel path (tag_debug_gemm_tc twice) is used instead: T not a power of two <= 128, or a
 * dilation whose halo tile does not fit next to the weight ring (T = 32: dilation 8). Bit-identical to that path. */
int tag_debug_tcn_block(tag_handle* h, void* h16, int64_t M, int32_t T, int32_t dil, const void* W1_16, const void* W2_16,
                        const float* gn_gamma, const float* gn_beta, void* stream);
/* host-only: 1 if the fused block kernel covers (M rows, T frames per window, dilation), with its shared-memory plan: weight-ring
 * stages, dynamic shared memory of the launch, bytes of the six activation tiles and their zero halos. No GPU needed. */
int tag_debug_tcn_block_plan(int64_t M, int32_t T, int32_t dil, int32_t* weight_stages, int32_t* smem_bytes, int32_t* tile_bytes);

#ifdef __cplusplus
}
#endif
#endif /* TAG_B200_H */
