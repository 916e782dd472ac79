// Internal launcher interface between tag_api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/tag_b200.h"

// ------------------------------------------------------------------ K1 feature fuse
struct FuseParams {
  int M;
  int kind[TAG_MAX_MODALITIES];
  int raw_dim[TAG_MAX_MODALITIES], diff_dim[TAG_MAX_MODALITIES];
  int raw_off[TAG_MAX_MODALITIES], diff_off[TAG_MAX_MODALITIES];       // fp32 feats columns
  int raw_off16[TAG_MAX_MODALITIES], diff_off16[TAG_MAX_MODALITIES];   // padded fp16 operand columns
  const float* src[TAG_MAX_MODALITIES];
  const int64_t* frame_offset;
  const int64_t* total_frames;   // &frame_offset[V]: rows of every source array
  const float* mean;        // [D] z-score SCALE table 1/(std+1e-6) (launch_zscore_table; all ones without stats)
  const float* stdv;        // [D] z-score SHIFT table -mean*scale (zeros without stats)
  const int32_t* win_video;
  const int32_t* win_start;
  int64_t n_windows;
  int T, D, D16;
  float* feats;             // [N,T,D] or null
  __half* feats16;          // [N,T,D16] or null (pad columns must be pre-zeroed by the caller)
  int32_t* flags;           // [1] or null
};
cudaError_t launch_zscore_table(const float* mean, const float* stdv, float* scale, float* shift, int D, cudaStream_t s);
cudaError_t launch_feature_fuse(const FuseParams& p, cudaStream_t s);
// kernels launched by launch_feature_fuse: the staged kernel (fp16-only output) is one launch; the generic path is one
// for the cosine modalities plus one for the rest
inline int feature_fuse_launches(const FuseParams& p) {
  if (p.feats == nullptr && p.feats16 != nullptr) return 1;
  bool c = false, o = false;
  for (int m = 0; m < p.M; ++m) { if (p.kind[m] == TAG_KIND_COSINE) c = true; else o = true; }
  return (c ? 1 : 0) + (o ? 1 : 0);
}

// ------------------------------------------------------------------ K3 / K4 / N1 / N2
// deterministic two-pass K3; `scratch` holds centroid_scratch_floats(n, C) floats (per-CTA partial sums)
size_t centroid_scratch_floats(int64_t n, int C);
cudaError_t launch_centroid_accumulate(const float* z, const int32_t* labels, int64_t n, int C,
                                       float* sums_counts, float* scratch, cudaStream_t s);
cudaError_t launch_centroid_finalize(const float* sums_counts, int C, float* centroids, float* counts, cudaStream_t s);
cudaError_t launch_score(const float* seq, const float* tcw, const int64_t* seg, const int32_t* label,
                         const float* cen, int C, int64_t V, float* ac, float* tc, cudaStream_t s);
cudaError_t launch_stats_accumulate(const float* x, int64_t rows, int D, double* sum, double* sumsq, cudaStream_t s);
cudaError_t launch_tcl_forward(const float* z, const int32_t* y, int64_t B, float temperature, float k1, float k2,
                               float* loss_rows, cudaStream_t s);
// tensor-core TCL: z [B,256] fp32 -> split-fp16 operands A = [hi | lo | hi], W = [hi | hi | lo] ([Bp, 768], rows >= B zero), so
// that A W^T = hi.hi + lo.hi + hi.lo = Z Z^T to ~1e-7; then the row sums of tcl_part -> loss_rows
cudaError_t launch_tcl_split(const float* z, int64_t B, int64_t Bp, __half* A, __half* W, cudaStream_t s);
cudaError_t launch_tcl_finish(const float* part, int64_t B, int slices, float k1, float k2, float* loss_rows, cudaStream_t s);
// SupConWithHardNegatives forward (losses.py:37-56): rows [B] of softplus((a.h - a.p) / temperature)
cudaError_t launch_supcon_hard(const float* anchor, const float* positive, const float* hard, int64_t B, float temperature,
                               float* loss_rows, cudaStream_t s);
// out[b, t, :] = x[b, idx[b*T + t], :] (hard-negative augmentations, utils.py:65-95); D % 4 == 0
cudaError_t launch_gather_frames(const float* x, const int32_t* idx, int64_t B, int T, int D, float* out, cudaStream_t s);

// ------------------------------------------------------------------ encoder building blocks
// Generic fp32 CUDA-core GEMM with conv taps:  C[M,N] = act( sum_j A[row+shift_j, :K] . W[n, j*K : (j+1)*K] + bias + res )
struct GemmF32 {
  const float* A; int lda;          // [M, K] rows (row stride lda)
  const float* W; int ldw;          // [N, taps*K] (K contiguous per tap)
  const float* bias;                // [N] or null
  const float* res; int ldr;        // [M, N] or null
  float* C; int ldc;
  int M, N, K;
  int taps, dil, T;                 // taps>1: rows are (window, t) with T frames; shift_j = (j - taps/2)*dil, zero outside the window
  int act;                          // 0 none, 1 gelu(erf), 2 relu
};
cudaError_t launch_gemm_f32(const GemmF32& g, cudaStream_t s);

template <typename TA>
cudaError_t launch_groupnorm(const TA* z, const float* gamma, const float* beta, TA* out, int64_t n_windows, int T,
                             cudaStream_t s);

struct MergeParams {
  int M;
  const void* ps[TAG_MAX_MODALITIES];   // state proj outputs [R,256]
  const void* pm[TAG_MAX_MODALITIES];   // motion proj outputs or null
  float inv_tau[TAG_MAX_MODALITIES], lbias[TAG_MAX_MODALITIES];
  const float* kv_gamma; const float* kv_beta;
  const float* qk;                      // [256] = Wk^T (Wq LN_q(latent)) / sqrt(256)
  void* mix;                            // [R,256]  sum_m A_m kv_m
  float* attn;                          // [R,M] or null
  int64_t R;
};
template <typename TA> cudaError_t launch_merge_fusion(const MergeParams& p, cudaStream_t s);

// tokens[n,0] = cls + pe[0]; tokens[n,t+1] = fused[n,t] + pe[t+1]   (model.py:187-188)
template <typename TA>
cudaError_t launch_build_tokens(const TA* fused, const float* cls, const float* pe, float* x32, TA* x16_or_null,
                                int64_t n_windows, int T, cudaStream_t s);

// softmax(QK^T/sqrt(32)) V per (window, head); qkv [N*S, 768] -> out [N*S, 256]
template <typename TA>
cudaError_t launch_attention(const TA* qkv, TA* out, int64_t n_windows, int S, int n_heads, cudaStream_t s);

// y = LayerNorm(x) * gamma + beta over 256 columns (x fp32 pre-norm sum), writes fp32 and optionally TA copy
template <typename TA>
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* y32, TA* y16_or_null,
                             int64_t rows, cudaStream_t s);

// model.py:190-192 + eval.py:218-224: normalise tokens, seq embed, per-window TC
cudaError_t launch_finalize(const float* tokens, int64_t n_windows, int S, float* seq, float* frame_embeds,
                            float* tokens_out, float* tc_window, cudaStream_t s);

cudaError_t launch_window_tc(const float* frame_embeds, int64_t n_windows, int S, float* tc_window, cudaStream_t s);

// ------------------------------------------------------------------ tensor-core (tcgen05) GEMM
struct GemmTC {
  const __half* A;      // activations, [M, lda] (K contiguous)
  int64_t M; int lda;
  const __half* A2;     // optional second K-segment source (same shape conventions), null if unused
  int lda2; int K2;
  const __half* W;      // [N, Ktot] K-major, Ktot = taps*K + K2
  int N, K;
  int taps, dil, T;     // conv: rows are (window, t)
  const float* bias;    // [N] or null
  const __half* res16; int ldr;   // residual (fp16) or null
  const float* res32;             // residual (fp32) or null (ld = N)
  __half* C16; int ldc;           // fp16 output or null
  float* C32;                     // fp32 output or null (ld = N)
  int act;
  const float* gn_gamma;          // non-null: fuse GroupNorm(1,256) over each (T x 256) window after the activation
  const float* gn_beta;
  const float* ln_gamma;          // non-null: fuse LayerNorm over the 256 output columns after bias + fp32 residual;
  const float* ln_beta;           //   needs N == 256, res32, C32 (may alias res32) and C16
  // window gather (the stems in frame-table mode): A is a per-FRAME table [g_rows, lda] of clips of g_L frames each, and
  // logical row (window w, frame t) reads table row (w / g_wpv) * g_L + (w % g_wpv) * g_stride + t. Plain GEMM, T a power of
  // two in 16..128. row0_vec non-null: output rows with t == 0 are replaced by this [N] vector (a window's first frame has
  // zero motion whatever the table holds).
  int g_L, g_wpv, g_stride; int64_t g_rows;
  const float* row0_vec;
  // TCL forward (losses.py:14-34): C = Z Z^T is never stored; the epilogue keeps, per row and per 64-column slice, the five masked
  // sums {sum_pos exp(S/t), sum_pos exp(-S), sum_neg exp(S/t), sum_pos S/t, #pos} in tcl_part [M][N/64][5] (all outputs null)
  const int32_t* tcl_y; float* tcl_part; float tcl_inv_temp; int tcl_valid;
};
struct TcContext;   // opaque: driver entry points + cached tensor maps
TcContext* tc_context_create(int device, char* err, int errlen);
void tc_context_destroy(TcContext*);
bool tc_pair_enabled(const TcContext*);   // CTA pairs (cta_group::2) in use: fused GroupNorm then also covers T == 256
cudaError_t launch_gemm_tc(TcContext* ctx, const GemmTC& g, cudaStream_t s, char* err, int errlen);
void* tc_encode_fn(const TcContext*);     // cuTensorMapEncodeTiled entry point
int tc_num_sms(const TcContext*);

// ------------------------------------------------------------------ fused TemporalConvBlock (tcn_block_tc.cu)
// h <- GroupNorm(GELU(conv2(GELU(conv1(h) + b1)) + b2 + h)) for [M, 256] fp16 rows of whole windows (M = windows * T), in place; the
// intermediate activation never leaves shared memory. Bit-identical to conv1 + GELU followed by conv2 + residual + GELU + GroupNorm
// through launch_gemm_tc. tcn_block_supported(): T a power of two <= 128, more than 128 rows, and a dilation whose halo tiles leave room for three weight stages.
struct TcnBlock {
  int64_t M; int T; int dil;
  __half* h16;                    // [M,256] block input = residual = output
  const __half* W1_16;            // [256, 5*256] tap-major K
  const __half* W2_16;
  const float* gn_gamma; const float* gn_beta;
};
bool tcn_block_supported(int64_t M, int T, int dil);
bool tcn_block_plan(int64_t M, int T, int dil, int* weight_stages, int* smem_bytes, int* tile_bytes);   // host only
cudaError_t launch_tcn_block(void* encode_fn, int num_sms, const TcnBlock& t, cudaStream_t s, char* err, int errlen);

// ------------------------------------------------------------------ fused transformer-layer tail (tlayer_tc.cu)
// x <- LN2(x1 + relu(x1 W1^T + b1) W2^T + b2), x1 = LN1(x + att Wo^T + bo): out-proj + norm1 + FFN + norm2 of one post-norm
// layer (model.py:145) in one tcgen05 kernel; x32 is updated in place, x16 receives its fp16 copy
struct TlayerTail {
  int64_t M; int ffn_dim;
  const __half* att16;            // [M,256] attention output (fp16)
  float* x32; __half* x16;        // [M,256] token stream (fp32, in/out) and its fp16 copy (out)
  const __half* Wo16;             // [256,256]
  const __half* W1_16;            // [ffn,256]
  const __half* W2_16;            // [256,ffn]
  const float *bo, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};
bool tlayer_tail_supported(int64_t M, int ffn_dim);
cudaError_t launch_tlayer_tail(void* encode_fn, int num_sms, const TlayerTail& t, cudaStream_t s, char* err, int errlen);
