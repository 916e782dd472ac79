// C ABI of libtag_b200.so (include/tag_b200.h): handle, weight packing, and the kernel schedule of
// the encoder forward (reference model.py:162-193) in both precision modes.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kD = TAG_D_MODEL;
char g_create_error[512] = "";

struct EncWeights {                 // one MovementConvEncoder (model.py:43-58)
  int d_in = 0, ldw = 0, k16 = 0;
  float* stem = nullptr;            // [256, ldw]  (ldw = d_in rounded up to 4, zero padded)
  float* conv[8][2] = {};           // [256, 5*256]: W[co][j*256 + ci] = weight[co, ci, j]
  float* gn_g[8] = {};
  float* gn_b[8] = {};
  float* proj = nullptr;            // [256, 256]
  __half* stem16 = nullptr;         // [256, k16] (k16 = d_in rounded up to 64)
  __half* conv16[8][2] = {};
  __half* proj16 = nullptr;
  __half* projcat16 = nullptr;      // state encoders only: [256, 512] = [proj_state | proj_motion] (one GEMM for s = state + motion)
};

struct LayerWeights {               // nn.TransformerEncoderLayer (model.py:145)
  float *in_w, *in_b, *out_w, *out_b, *l1_w, *l1_b, *l2_w, *l2_b, *n1_g, *n1_b, *n2_g, *n2_b;
  __half *in_w16, *out_w16, *l1_w16, *l2_w16;
};

// kind: 0 other, 1 conv GEMM, 2 other GEMM, 3 feature fuse (K1), 4 merge-fusion, 5 finalize (+ per-window TC), 6 attention,
// 7 build-tokens; `flops` = FLOPs for the GEMM kinds, algorithmic bytes for the bandwidth-bound kinds
constexpr int kProfKinds = 8;
struct ProfEvent { cudaEvent_t a, b; double flops; int kind; };

}  // namespace

struct tag_handle {
  tag_config cfg;
  int M = 0, D = 0, D16 = 0, raw_total = 0;
  int raw_off[TAG_MAX_MODALITIES], diff_off[TAG_MAX_MODALITIES];
  int raw_off16[TAG_MAX_MODALITIES], diff_off16[TAG_MAX_MODALITIES];
  bool finalized = false;
  char err[512] = "";
  int64_t launches = 0;

  std::map<std::string, std::vector<float>> staged;
  std::map<std::string, std::vector<int64_t>> staged_shape;
  std::vector<void*> allocs;        // workspace and tables (live as long as the handle)
  std::vector<void*> w_allocs;      // packed weights (replaced by tag_reload_weights_begin + load + finalize)
  bool packing_weights = false;     // dev_alloc sink selector
  bool reloading = false;

  EncWeights state[TAG_MAX_MODALITIES], motion[TAG_MAX_MODALITIES];
  std::vector<LayerWeights> layers;
  float *kv_g = nullptr, *kv_b = nullptr, *qk = nullptr, *Wv = nullptr, *Wo = nullptr, *cls = nullptr, *pe = nullptr;
  __half* Wov16 = nullptr;
  float inv_tau[TAG_MAX_MODALITIES], lbias[TAG_MAX_MODALITIES];
  int pe_rows = 0;

  // workspace (sized for max_windows x max_T)
  void *bufH = nullptr, *bufY1 = nullptr, *bufY2 = nullptr, *mix = nullptr, *fusedA = nullptr, *fusedB = nullptr;
  void* P[2 * TAG_MAX_MODALITIES] = {};
  float *X = nullptr, *TMP = nullptr;
  void *X16 = nullptr, *QKV = nullptr, *ATT = nullptr, *FF = nullptr;
  float* feats = nullptr;           // [max_windows, max_T, D] fp32 (fp32 mode, fused entry)
  __half* feats16 = nullptr;        // [max_windows, max_T, D16] (tensor-core mode)
  TcContext* tc = nullptr;

  // profiling
  bool profiling = false;
  std::vector<ProfEvent> prof;
  size_t prof_used = 0;
  double prof_acc[3 * kProfKinds] = {};   // per kind: ms, flops (bandwidth-bound kinds: bytes), launches
  int* col_tab = nullptr;           // device [M][6] column map fp32 feats -> fp16 operand layout
  float* zs_scale = nullptr;        // [D] z-score tables, rebuilt from (mean, std) at every feature-fuse call
  float* zs_shift = nullptr;
  // frame-table mode (tag_encode_clips): window index helpers and the motion stems' zero-motion output rows
  int32_t* iota = nullptr;          // [max_windows] 0, 1, 2, ...
  int32_t* zeros = nullptr;         // [max_windows] 0
  int32_t *clip_wv = nullptr, *clip_ws = nullptr;   // [max_windows] window table of one pass (fallback path)
  float* row0 = nullptr;            // [M][256]: motion stem applied to the z-scored zero difference
  int fuse_tcn = 1;                 // fused TemporalConvBlock kernel where it fits (TAG_FUSE_TCN=0: conv1 and conv2 + GroupNorm as two GEMM launches)
  int fuse_tail = 1;                // fused transformer-layer tail (TAG_FUSE_TAIL=0 selects the three separate GEMMs in TAG_EXPERIMENTS builds)
  int frame_table = 1;              // frame-table mode of tag_encode_clips (TAG_FRAME_TABLE=0 disables it in TAG_EXPERIMENTS builds)
  float* attn_out = nullptr;        // tag_set_fusion_attn_out: [rows, M] fusion softmax of the NEXT tag_encode call (model.py:94 last_attn)
  __half *tcl_A = nullptr, *tcl_W = nullptr;   // tensor-core TCL: split-fp16 operands [Bp, 768] and the per-slice partial sums
  float* tcl_part = nullptr;
  int64_t tcl_cap = 0;              // rows (Bp) the three buffers are sized for
  float* k3_scratch = nullptr;      // per-CTA partial sums of the deterministic centroid reduction (grown on demand)
  size_t k3_scratch_floats = 0;
};

namespace {

int fail(tag_handle* h, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h ? h->err : g_create_error, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(h, TAG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

template <typename T>
int dev_alloc(tag_handle* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
  if (e != cudaSuccess) return fail(h, TAG_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  (h->packing_weights ? h->w_allocs : h->allocs).push_back(q);
  *p = reinterpret_cast<T*>(q);
  return TAG_OK;
}

template <typename T>
int upload(tag_handle* h, T** p, const std::vector<T>& v) {
  int rc = dev_alloc(h, p, v.size());
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TAG_OK;
}

std::vector<__half> to_half(const std::vector<float>& v) {
  std::vector<__half> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2half_rn(v[i]);
  return o;
}

const std::vector<float>* find_w(tag_handle* h, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = h->staged.find(key);
  if (it == h->staged.end()) {
    fail(h, TAG_ERR_MISSING, "weight '%s' was never loaded", key.c_str());
    return nullptr;
  }
  const auto& s = h->staged_shape[key];
  bool ok = s.size() == shape.size();
  size_t i = 0;
  for (int64_t d : shape) { if (ok && s[i] != d) ok = false; ++i; }
  if (!ok) {
    fail(h, TAG_ERR_INVALID, "weight '%s' has the wrong shape", key.c_str());
    return nullptr;
  }
  return &it->second;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

int pack_encoder(tag_handle* h, const std::string& prefix, int d_in, EncWeights* e) {
  const bool tc = h->cfg.precision == TAG_PRECISION_FP16_TC;
  const int nb = h->cfg.n_blocks, kk = h->cfg.conv_kernel;
  e->d_in = d_in;
  e->ldw = round_up(d_in, 4);
  e->k16 = round_up(d_in, 64);
  const auto* w = find_w(h, prefix + ".stem.weight", {kD, d_in, 1});
  if (!w) return TAG_ERR_MISSING;
  {
    std::vector<float> p((size_t)kD * e->ldw, 0.f);
    for (int co = 0; co < kD; ++co)
      for (int k = 0; k < d_in; ++k) p[(size_t)co * e->ldw + k] = (*w)[(size_t)co * d_in + k];
    int rc = upload(h, &e->stem, p);
    if (rc) return rc;
    if (tc) {
      std::vector<__half> q((size_t)kD * e->k16, __float2half_rn(0.f));
      for (int co = 0; co < kD; ++co)
        for (int k = 0; k < d_in; ++k) q[(size_t)co * e->k16 + k] = __float2half_rn((*w)[(size_t)co * d_in + k]);
      rc = upload(h, &e->stem16, q);
      if (rc) return rc;
    }
  }
  for (int b = 0; b < nb; ++b) {
    for (int c = 0; c < 2; ++c) {
      const std::string key = prefix + ".blocks." + std::to_string(b) + (c == 0 ? ".conv1.weight" : ".conv2.weight");
      const auto* cw = find_w(h, key, {kD, kD, kk});
      if (!cw) return TAG_ERR_MISSING;
      std::vector<float> p((size_t)kD * kk * kD);
      for (int co = 0; co < kD; ++co)
        for (int ci = 0; ci < kD; ++ci)
          for (int j = 0; j < kk; ++j)
            p[(size_t)co * kk * kD + (size_t)j * kD + ci] = (*cw)[((size_t)co * kD + ci) * kk + j];
      int rc = upload(h, &e->conv[b][c], p);
      if (rc) return rc;
      if (tc) { rc = upload(h, &e->conv16[b][c], to_half(p)); if (rc) return rc; }
    }
    const auto* g = find_w(h, prefix + ".blocks." + std::to_string(b) + ".norm.weight", {kD});
    const auto* bb = find_w(h, prefix + ".blocks." + std::to_string(b) + ".norm.bias", {kD});
    if (!g || !bb) return TAG_ERR_MISSING;
    int rc = upload(h, &e->gn_g[b], *g); if (rc) return rc;
    rc = upload(h, &e->gn_b[b], *bb); if (rc) return rc;
  }
  const auto* pw = find_w(h, prefix + ".proj.weight", {kD, kD});
  if (!pw) return TAG_ERR_MISSING;
  int rc = upload(h, &e->proj, *pw); if (rc) return rc;
  if (tc) { rc = upload(h, &e->proj16, to_half(*pw)); if (rc) return rc; }
  return TAG_OK;
}

// ------------------------------------------------------------------------------------------------
// profiling helpers
struct ProfScope {
  tag_handle* h; ProfEvent* ev = nullptr; cudaStream_t s;
  ProfScope(tag_handle* h_, cudaStream_t s_, int kind, double flops) : h(h_), s(s_) {
    if (h->profiling && h->prof_used < h->prof.size()) {
      ev = &h->prof[h->prof_used++];
      ev->kind = kind; ev->flops = flops;
      cudaEventRecord(ev->a, s);
    }
  }
  ~ProfScope() { if (ev) cudaEventRecord(ev->b, s); }
};

// ------------------------------------------------------------------------------------------------
// encoder schedule. Mode = float (fp32 CUDA-core GEMMs) or __half (tcgen05 GEMMs).
template <typename TA> struct Gemm;

template <> struct Gemm<float> {
  // C = act(A W^T + bias + res); conv when taps > 1
  static int run(tag_handle* h, cudaStream_t s, const float* A, int lda, const float* W, int ldw, int64_t M, int N,
                 int K, int taps, int dil, int T, const float* bias, const float* res, float* C, int act) {
    GemmF32 g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.res = res; g.ldr = N; g.C = C; g.ldc = N;
    g.M = (int)M; g.N = N; g.K = K; g.taps = taps; g.dil = dil; g.T = T; g.act = act;
    ProfScope ps(h, s, taps > 1 ? 1 : 2, 2.0 * (double)M * N * K * taps);
    cudaError_t e = launch_gemm_f32(g, s);
    h->launches++;
    if (e != cudaSuccess) return fail(h, TAG_ERR_CUDA, "gemm_f32 launch failed: %s", cudaGetErrorString(e));
    return TAG_OK;
  }
};

#define LAUNCH_TRY(h, expr)                                                                               \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    (h)->launches++;                                                                                      \
    if (e__ != cudaSuccess) return fail(h, TAG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

int gemm_tc_run(tag_handle* h, cudaStream_t s, const GemmTC& g, double flops) {
  ProfScope ps(h, s, g.taps > 1 ? 1 : 2, flops);
  cudaError_t e = launch_gemm_tc(h->tc, g, s, h->err, 512);
  h->launches++;
  if (e != cudaSuccess) {
    if (h->err[0] == 0) fail(h, TAG_ERR_CUDA, "gemm_tc launch failed: %s", cudaGetErrorString(e));
    return TAG_ERR_CUDA;
  }
  return TAG_OK;
}

// ---- fp32 mode ----------------------------------------------------------------------------------
int encode_chunk_f32(tag_handle* h, cudaStream_t s, const float* feats, int64_t W, int T, float* seq, float* frame,
                     float* tokens, float* tcw, float* attn = nullptr) {
  const int64_t R = W * T, R2 = W * (T + 1);
  const int M = h->M;
  float* bufH = (float*)h->bufH; float* bufY1 = (float*)h->bufY1; float* bufY2 = (float*)h->bufY2;
  int rc;
  int e_idx = 0;
  MergeParams mp{};
  mp.M = M;
  for (int m = 0; m < M; ++m) {
    for (int side = 0; side < 2; ++side) {
      if (side == 1 && h->cfg.diff_dims[m] <= 0) { mp.pm[m] = nullptr; continue; }
      const EncWeights& e = side == 0 ? h->state[m] : h->motion[m];
      const float* A = feats + (side == 0 ? h->raw_off[m] : h->diff_off[m]);
      rc = Gemm<float>::run(h, s, A, h->D, e.stem, e.ldw, R, kD, e.d_in, 1, 1, T, nullptr, nullptr, bufH, 0);
      if (rc) return rc;
      for (int b = 0; b < h->cfg.n_blocks; ++b) {
        const int dil = 1 << b;
        rc = Gemm<float>::run(h, s, bufH, kD, e.conv[b][0], h->cfg.conv_kernel * kD, R, kD, kD, h->cfg.conv_kernel, dil, T,
                              nullptr, nullptr, bufY1, 1);
        if (rc) return rc;
        rc = Gemm<float>::run(h, s, bufY1, kD, e.conv[b][1], h->cfg.conv_kernel * kD, R, kD, kD, h->cfg.conv_kernel, dil, T,
                              nullptr, bufH, bufY2, 1);
        if (rc) return rc;
        { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_groupnorm<float>(bufY2, e.gn_g[b], e.gn_b[b], bufH, W, T, s)); }
      }
      float* P = (float*)h->P[e_idx++];
      rc = Gemm<float>::run(h, s, bufH, kD, e.proj, kD, R, kD, kD, 1, 1, T, nullptr, nullptr, P, 0);
      if (rc) return rc;
      if (side == 0) mp.ps[m] = P; else mp.pm[m] = P;
    }
    mp.inv_tau[m] = h->inv_tau[m];
    mp.lbias[m] = h->lbias[m];
  }
  mp.kv_gamma = h->kv_g; mp.kv_beta = h->kv_b; mp.qk = h->qk; mp.mix = h->mix; mp.attn = attn; mp.R = R;
  { ProfScope ps(h, s, 4, (double)R * (2 * M + 1) * kD * 4); LAUNCH_TRY(h, launch_merge_fusion<float>(mp, s)); }
  float* fusedA = (float*)h->fusedA; float* fusedB = (float*)h->fusedB;
  rc = Gemm<float>::run(h, s, (const float*)h->mix, kD, h->Wv, kD, R, kD, kD, 1, 1, T, nullptr, nullptr, fusedA, 0);
  if (rc) return rc;
  rc = Gemm<float>::run(h, s, fusedA, kD, h->Wo, kD, R, kD, kD, 1, 1, T, nullptr, nullptr, fusedB, 0);
  if (rc) return rc;
  { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_build_tokens<float>(fusedB, h->cls, h->pe, h->X, nullptr, W, T, s)); }
  float* QKV = (float*)h->QKV; float* ATT = (float*)h->ATT; float* FF = (float*)h->FF;
  const int F = h->cfg.ffn_dim;
  for (const LayerWeights& L : h->layers) {
    rc = Gemm<float>::run(h, s, h->X, kD, L.in_w, kD, R2, 3 * kD, kD, 1, 1, 1, L.in_b, nullptr, QKV, 0);
    if (rc) return rc;
    { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_attention<float>(QKV, ATT, W, T + 1, h->cfg.n_heads, s)); }
    rc = Gemm<float>::run(h, s, ATT, kD, L.out_w, kD, R2, kD, kD, 1, 1, 1, L.out_b, h->X, h->TMP, 0);
    if (rc) return rc;
    { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_layernorm<float>(h->TMP, L.n1_g, L.n1_b, h->X, nullptr, R2, s)); }
    rc = Gemm<float>::run(h, s, h->X, kD, L.l1_w, kD, R2, F, kD, 1, 1, 1, L.l1_b, nullptr, FF, 2);
    if (rc) return rc;
    rc = Gemm<float>::run(h, s, FF, F, L.l2_w, F, R2, kD, F, 1, 1, 1, L.l2_b, h->X, h->TMP, 0);
    if (rc) return rc;
    { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_layernorm<float>(h->TMP, L.n2_g, L.n2_b, h->X, nullptr, R2, s)); }
  }
  { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_finalize(h->X, W, T + 1, seq, frame, tokens, tcw, s)); }
  return TAG_OK;
}

// ---- tensor-core mode -----------------------------------------------------------------------------
// frame-table mode: feats16 holds ONE row per source frame of the pass's clips (L rows per clip); windows are gathered by
// the stem GEMMs (GemmTC::g_*), and the first frame of every window takes the motion stems' zero-motion row
struct ClipGather { int L = 0, wpv = 0, stride = 0; int64_t rows = 0; };

int encode_chunk_tc(tag_handle* h, cudaStream_t s, const __half* feats16, int64_t W, int T, float* seq, float* frame,
                    float* tokens, float* tcw, const ClipGather* cg = nullptr, float* attn = nullptr) {
  const int64_t R = W * T, R2 = W * (T + 1);
  const int M = h->M;
  __half* bufY1 = (__half*)h->bufY1; __half* bufY2 = (__half*)h->bufY2;
  int rc;
  int e_idx = 0;
  MergeParams mp{};
  mp.M = M;
  const int kk = h->cfg.conv_kernel;
  for (int m = 0; m < M; ++m) {
    for (int side = 0; side < 2; ++side) {
      if (side == 1 && h->cfg.diff_dims[m] <= 0) continue;
      const EncWeights& e = side == 0 ? h->state[m] : h->motion[m];
      const bool has_motion = h->cfg.diff_dims[m] > 0;
      // the state encoder's final hidden state waits in its own buffer for the motion encoder's: s = state + motion
      // (model.py:174) is ONE projection GEMM over K = [h_state | h_motion]
      __half* bufH = (side == 0 && has_motion) ? (__half*)h->fusedB : (__half*)h->bufH;
      GemmTC g{};
      g.A = feats16 + (side == 0 ? h->raw_off16[m] : h->diff_off16[m]); g.M = R; g.lda = h->D16;
      g.W = e.stem16; g.N = kD; g.K = e.k16; g.taps = 1; g.dil = 1; g.T = T; g.C16 = bufH; g.ldc = kD;
      if (cg != nullptr) {
        g.g_L = cg->L; g.g_wpv = cg->wpv; g.g_stride = cg->stride; g.g_rows = cg->rows;
        if (side == 1) g.row0_vec = h->row0 + (size_t)m * kD;
      }
      rc = gemm_tc_run(h, s, g, 2.0 * R * kD * e.d_in); if (rc) return rc;
      for (int b = 0; b < h->cfg.n_blocks; ++b) {
        if (h->fuse_tcn && !(h->fuse_tcn == 2 && b >= 3) && kk == 5 && tc_pair_enabled(h->tc) && tcn_block_supported(R, T, 1 << b)) {   // 2: A/B switch of the experiments build (dilation >= 8 unfused)
          // the whole TemporalConvBlock in ONE kernel (tcn_block_tc.cu): GELU(conv1) stays in shared memory as conv2's operand
          TcnBlock tb{};
          tb.M = R; tb.T = T; tb.dil = 1 << b; tb.h16 = bufH; tb.W1_16 = e.conv16[b][0]; tb.W2_16 = e.conv16[b][1];
          tb.gn_gamma = e.gn_g[b]; tb.gn_beta = e.gn_b[b];
          ProfScope ps(h, s, 1, 2.0 * (2.0 * R * kD * kD * kk));
          h->err[0] = 0;
          cudaError_t ce = launch_tcn_block(tc_encode_fn(h->tc), tc_num_sms(h->tc), tb, s, h->err, 512);
          h->launches++;
          if (ce != cudaSuccess) {
            if (h->err[0] == 0) fail(h, TAG_ERR_CUDA, "tcn_block launch failed: %s", cudaGetErrorString(ce));
            return TAG_ERR_CUDA;
          }
          continue;
        }
        GemmTC c1{};
        c1.A = bufH; c1.M = R; c1.lda = kD; c1.W = e.conv16[b][0]; c1.N = kD; c1.K = kD; c1.taps = kk; c1.dil = 1 << b;
        c1.T = T; c1.C16 = bufY1; c1.ldc = kD; c1.act = 1;
        rc = gemm_tc_run(h, s, c1, 2.0 * R * kD * kD * kk); if (rc) return rc;
        GemmTC c2 = c1;
        c2.A = bufY1; c2.W = e.conv16[b][1]; c2.res16 = bufH; c2.ldr = kD;
        // a tile (T <= 128) or a CTA pair (T == 256) owns whole windows: GroupNorm fuses into conv2's epilogue
        const bool fuse_gn = (T & (T - 1)) == 0 && (T <= 128 || (T == 256 && tc_pair_enabled(h->tc)));
        if (fuse_gn) {
          // conv2 + residual + GELU + GroupNorm in one kernel; in place over the residual (a thread reads and
          // later overwrites only its own row/column block)
          c2.C16 = bufH; c2.gn_gamma = e.gn_g[b]; c2.gn_beta = e.gn_b[b];
          rc = gemm_tc_run(h, s, c2, 2.0 * R * kD * kD * kk); if (rc) return rc;
        } else {
          c2.C16 = bufY2;
          rc = gemm_tc_run(h, s, c2, 2.0 * R * kD * kD * kk); if (rc) return rc;
          { ProfScope ps(h, s, 0, 0); LAUNCH_TRY(h, launch_groupnorm<__half>(bufY2, e.gn_g[b], e.gn_b[b], bufH, W, T, s)); }
        }
      }
      if (side == 0 && has_motion) continue;
      __half* P = (__half*)h->P[e_idx++];
      GemmTC pj{};
      pj.M = R; pj.N = kD; pj.K = kD; pj.taps = 1; pj.dil = 1; pj.T = T; pj.C16 = P; pj.ldc = kD;
      if (side == 1) {
        pj.A = (const __half*)h->fusedB; pj.lda = kD; pj.A2 = bufH; pj.lda2 = kD; pj.K2 = kD; pj.W = h->state[m].projcat16;
        rc = gemm_tc_run(h, s, pj, 2.0 * R * kD * 2 * kD); if (rc) return rc;
      } else {
        pj.A = bufH; pj.lda = kD; pj.W = e.proj16;
        rc = gemm_tc_run(h, s, pj, 2.0 * R * kD * kD); if (rc) return rc;
      }
      mp.ps[m] = P; mp.pm[m] = nullptr;
    }
    mp.inv_tau[m] = h->inv_tau[m];
    mp.lbias[m] = h->lbias[m];
  }
  mp.kv_gamma = h->kv_g; mp.kv_beta = h->kv_b; mp.qk = h->qk; mp.mix = h->mix; mp.attn = attn; mp.R = R;
  { ProfScope ps(h, s, 4, (double)R * (M + 1) * kD * 2); LAUNCH_TRY(h, launch_merge_fusion<__half>(mp, s)); }
  __half* fused = (__half*)h->fusedA;
  {
    GemmTC g{};
    g.A = (const __half*)h->mix; g.M = R; g.lda = kD; g.W = h->Wov16; g.N = kD; g.K = kD; g.taps = 1; g.dil = 1; g.T = T;
    g.C16 = fused; g.ldc = kD;
    // FLOPs credited as the reference's two GEMMs (Wv then Wo) — the merge is an algebraic saving
    rc = gemm_tc_run(h, s, g, 2.0 * R * kD * kD); if (rc) return rc;
  }
  __half* X16 = (__half*)h->X16;
  { ProfScope ps(h, s, 7, (double)R * kD * 2 + (double)R2 * kD * 6); LAUNCH_TRY(h, launch_build_tokens<__half>(fused, h->cls, h->pe, h->X, X16, W, T, s)); }
  __half* QKV = (__half*)h->QKV; __half* ATT = (__half*)h->ATT; __half* FF = (__half*)h->FF;
  const int F = h->cfg.ffn_dim;
  for (const LayerWeights& L : h->layers) {
    GemmTC q{};
    q.A = X16; q.M = R2; q.lda = kD; q.W = L.in_w16; q.N = 3 * kD; q.K = kD; q.taps = 1; q.dil = 1; q.T = 1; q.bias = L.in_b;
    q.C16 = QKV; q.ldc = 3 * kD;
    rc = gemm_tc_run(h, s, q, 2.0 * R2 * 3 * kD * kD); if (rc) return rc;
    { ProfScope ps(h, s, 6, (double)R2 * kD * 8); LAUNCH_TRY(h, launch_attention<__half>(QKV, ATT, W, T + 1, h->cfg.n_heads, s)); }
    if (h->fuse_tail && tlayer_tail_supported(R2, F)) {
      // out-proj + norm1 + FFN1 + ReLU + FFN2 + norm2 in ONE kernel (tlayer_tc.cu): the FFN hidden state, x1 and its fp16
      // copy stay in TMEM / shared memory
      TlayerTail t{};
      t.M = R2; t.ffn_dim = F; t.att16 = ATT; t.x32 = h->X; t.x16 = X16; t.Wo16 = L.out_w16; t.W1_16 = L.l1_w16; t.W2_16 = L.l2_w16;
      t.bo = L.out_b; t.b1 = L.l1_b; t.b2 = L.l2_b; t.ln1_g = L.n1_g; t.ln1_b = L.n1_b; t.ln2_g = L.n2_g; t.ln2_b = L.n2_b;
      ProfScope ps(h, s, 2, 2.0 * R2 * kD * kD + 4.0 * R2 * F * kD);
      h->err[0] = 0;
      cudaError_t e = launch_tlayer_tail(tc_encode_fn(h->tc), tc_num_sms(h->tc), t, s, h->err, 512);
      h->launches++;
      if (e != cudaSuccess) {
        if (h->err[0] == 0) fail(h, TAG_ERR_CUDA, "tlayer_tail launch failed: %s", cudaGetErrorString(e));
        return TAG_ERR_CUDA;
      }
      continue;
    }
    GemmTC o{};
    o.A = ATT; o.M = R2; o.lda = kD; o.W = L.out_w16; o.N = kD; o.K = kD; o.taps = 1; o.dil = 1; o.T = 1; o.bias = L.out_b;
    // out-proj + bias + residual + LayerNorm(norm1) in one kernel, in place over the fp32 token stream
    o.res32 = h->X; o.C32 = h->X; o.C16 = X16; o.ldc = kD; o.ln_gamma = L.n1_g; o.ln_beta = L.n1_b;
    rc = gemm_tc_run(h, s, o, 2.0 * R2 * kD * kD); if (rc) return rc;
    GemmTC f1{};
    f1.A = X16; f1.M = R2; f1.lda = kD; f1.W = L.l1_w16; f1.N = F; f1.K = kD; f1.taps = 1; f1.dil = 1; f1.T = 1; f1.bias = L.l1_b;
    f1.C16 = FF; f1.ldc = F; f1.act = 2;
    rc = gemm_tc_run(h, s, f1, 2.0 * R2 * F * kD); if (rc) return rc;
    GemmTC f2{};
    f2.A = FF; f2.M = R2; f2.lda = F; f2.W = L.l2_w16; f2.N = kD; f2.K = F; f2.taps = 1; f2.dil = 1; f2.T = 1; f2.bias = L.l2_b;
    f2.res32 = h->X; f2.C32 = h->X; f2.C16 = X16; f2.ldc = kD; f2.ln_gamma = L.n2_g; f2.ln_beta = L.n2_b;
    rc = gemm_tc_run(h, s, f2, 2.0 * R2 * F * kD); if (rc) return rc;
  }
  {
    // algorithmic bytes: the fp32 token stream in, seq embeds (+ frame embeds / tokens when asked for) and the window TC out
    const double bytes = (double)R2 * kD * 4 * (1 + (frame ? 1 : 0) + (tokens ? 1 : 0)) + (double)W * (kD * 4 + 4);
    ProfScope ps(h, s, 5, bytes); LAUNCH_TRY(h, launch_finalize(h->X, W, T + 1, seq, frame, tokens, tcw, s));
  }
  return TAG_OK;
}

// fp32 feats -> padded fp16 operand layout (tensor-core mode of tag_encode)
__global__ void k_feats_to_half(const float* __restrict__ f, __half* __restrict__ o, int64_t rows, int D, int D16, int M,
                                const int* __restrict__ tab /* [M][6]: raw_off, raw_dim, raw_off16, diff_off, diff_dim, diff_off16 */) {
  const int64_t r = blockIdx.x;
  if (r >= rows) return;
  const float* fr = f + r * D;
  __half* orow = o + r * D16;
  for (int i = threadIdx.x; i < D16; i += blockDim.x) orow[i] = __float2half_rn(0.f);
  __syncthreads();
  for (int m = 0; m < M; ++m) {
    const int* t = tab + m * 6;
    for (int i = threadIdx.x; i < t[1]; i += blockDim.x) orow[t[2] + i] = __float2half_rn(fr[t[0] + i]);
    for (int i = threadIdx.x; i < t[4]; i += blockDim.x) orow[t[5] + i] = __float2half_rn(fr[t[3] + i]);
  }
}

int check_common(tag_handle* h, int64_t n_windows, int T) {
  if (!h) return TAG_ERR_INVALID;
  if (!h->finalized) return fail(h, TAG_ERR_STATE, "tag_finalize_weights has not been called");
  if (n_windows < 0) return fail(h, TAG_ERR_INVALID, "n_windows < 0");
  if (T < 1 || T > h->cfg.max_T) return fail(h, TAG_ERR_INVALID, "T=%d outside [1, max_T=%d]", T, h->cfg.max_T);
  if (T + 1 > h->pe_rows) return fail(h, TAG_ERR_INVALID, "T+1=%d exceeds the positional table (%d rows)", T + 1, h->pe_rows);
  if (h->cfg.precision == TAG_PRECISION_FP16_TC && !((T <= 128 && 128 % T == 0) || T % 128 == 0))
    return fail(h, TAG_ERR_UNSUPPORTED, "clip length T=%d: the tensor-core mode tiles windows into 128-row blocks and needs T dividing 128 "
                "(1, 2, 4, ..., 128) or a multiple of 128; use precision fp32 (TAG_PRECISION_FP32) for other clip lengths", T);
  return TAG_OK;
}

void prof_begin(tag_handle*) {}
// events are only recorded on the hot path; they are read back (after a device sync) in tag_get_profile
int prof_end(tag_handle*, cudaStream_t) { return TAG_OK; }

int prof_collect(tag_handle* h) {
  if (h->prof_used == 0) return TAG_OK;
  CUDA_TRY(h, cudaDeviceSynchronize());
  for (size_t i = 0; i < h->prof_used; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->prof[i].a, h->prof[i].b);
    const int k = h->prof[i].kind;
    h->prof_acc[k * 3 + 0] += ms; h->prof_acc[k * 3 + 1] += h->prof[i].flops; h->prof_acc[k * 3 + 2] += 1.0;
  }
  h->prof_used = 0;
  return TAG_OK;
}

// row0[n] = sum_k half(shift[k]) * W16[n][k]: what a motion stem GEMM produces for a row whose z-scored difference is zero
// (K1 writes half(0 * scale + shift) there); same fp16 operands as the tensor-core GEMM, fp32 accumulation
__global__ void k_row0_vec(const float* __restrict__ shift, const __half* __restrict__ W16, int k_valid, int k16,
                           float* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= kD) return;
  float acc = 0.f;
  for (int k = lane; k < k_valid; k += 32)
    acc = fmaf(__half2float(__float2half_rn(shift[k])), __half2float(W16[(size_t)n * k16 + k]), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[n] = acc;
}

__global__ void k_iota_zero(int32_t* iota, int32_t* zeros, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { iota[i] = i; zeros[i] = 0; }
}

// window table of clips of equal length: window w of the pass = clip v0 + w / wpv, start (w % wpv) * stride
__global__ void k_clip_windows(int32_t* wv, int32_t* ws, int64_t w0, int n, int wpv, int stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int64_t w = w0 + i; wv[i] = (int32_t)(w / wpv); ws[i] = (int32_t)(w % wpv) * stride; }
}

int fill_fuse_params(tag_handle* h, FuseParams* p, const tag_videos* vids, const float* mean, const float* stdv,
                     const int32_t* win_video, const int32_t* win_start, int64_t n, int T) {
  if (!vids || !vids->frame_offset) return fail(h, TAG_ERR_INVALID, "vids / frame_offset is NULL");
  if ((mean == nullptr) != (stdv == nullptr)) return fail(h, TAG_ERR_INVALID, "mean and stdv must both be given or both NULL");
  memset(p, 0, sizeof(*p));
  p->M = h->M;
  int n_proc = 0;
  for (int m = 0; m < h->M; ++m) {
    p->kind[m] = h->cfg.kinds[m];
    p->raw_dim[m] = h->cfg.raw_dims[m]; p->diff_dim[m] = h->cfg.diff_dims[m];
    p->raw_off[m] = h->raw_off[m]; p->diff_off[m] = h->diff_off[m];
    p->raw_off16[m] = h->raw_off16[m]; p->diff_off16[m] = h->diff_off16[m];
    p->src[m] = vids->src[m];
    if (!vids->src[m]) return fail(h, TAG_ERR_INVALID, "vids->src[%d] is NULL", m);
    if (p->kind[m] == TAG_KIND_PROCRUSTES) ++n_proc;
  }
  if (n_proc > 1) return fail(h, TAG_ERR_UNSUPPORTED, "at most one TAG_KIND_PROCRUSTES modality is supported");
  p->frame_offset = vids->frame_offset;
  p->mean = h->zs_scale;                       // the kernel reads the (scale, shift) tables (identity without stats)
  p->stdv = h->zs_shift;
  p->total_frames = vids->frame_offset + vids->n_videos;
  p->win_video = win_video; p->win_start = win_start; p->n_windows = n; p->T = T; p->D = h->D; p->D16 = h->D16;
  return TAG_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int tag_abi_version(void) { return 2; }

const char* tag_last_error(const tag_handle* h) { return h ? h->err : g_create_error; }

int tag_create(tag_handle** out, const tag_config* cfg) {
  if (!out || !cfg) return fail(nullptr, TAG_ERR_INVALID, "tag_create: NULL argument");
  *out = nullptr;
  if (cfg->n_modalities < 1 || cfg->n_modalities > TAG_MAX_MODALITIES)
    return fail(nullptr, TAG_ERR_INVALID, "n_modalities=%d outside [1,%d]", cfg->n_modalities, TAG_MAX_MODALITIES);
  if (cfg->d_model != kD) return fail(nullptr, TAG_ERR_UNSUPPORTED, "d_model=%d (only 256 is built)", cfg->d_model);
  if (cfg->n_heads < 1 || cfg->d_model / cfg->n_heads != 32 || cfg->d_model % cfg->n_heads)
    return fail(nullptr, TAG_ERR_UNSUPPORTED, "n_heads=%d (head_dim must be 32)", cfg->n_heads);
  if (cfg->conv_kernel != 5 || cfg->n_blocks < 1 || cfg->n_blocks > 8)
    return fail(nullptr, TAG_ERR_UNSUPPORTED, "conv_kernel=%d n_blocks=%d unsupported", cfg->conv_kernel, cfg->n_blocks);
  if (cfg->ffn_dim < 4 || cfg->ffn_dim % 64) return fail(nullptr, TAG_ERR_UNSUPPORTED, "ffn_dim=%d must be a multiple of 64", cfg->ffn_dim);
  if (cfg->n_layers < 0 || cfg->n_layers > 64) return fail(nullptr, TAG_ERR_INVALID, "n_layers=%d", cfg->n_layers);
  if (cfg->max_windows < 1 || cfg->max_T < 1) return fail(nullptr, TAG_ERR_INVALID, "max_windows / max_T must be >= 1");
  if (cfg->precision != TAG_PRECISION_FP32 && cfg->precision != TAG_PRECISION_FP16_TC)
    return fail(nullptr, TAG_ERR_INVALID, "precision=%d", cfg->precision);
  for (int m = 0; m < cfg->n_modalities; ++m) {
    const int rd = cfg->raw_dims[m], dd = cfg->diff_dims[m], k = cfg->kinds[m];
    if (rd < 1 || dd < 0) return fail(nullptr, TAG_ERR_INVALID, "modality %d: raw_dim=%d diff_dim=%d", m, rd, dd);
    bool ok = true;
    if (k == TAG_KIND_COSINE) ok = rd <= 1024 && rd % 2 == 0 && (dd == 0 || dd == rd);
    else if (k == TAG_KIND_ROTMAT) ok = rd % 9 == 0 && rd / 9 <= 256 && (dd == 0 || dd == rd / 3);
    else if (k == TAG_KIND_PLAIN) ok = (dd == 0 || dd == rd);
    else if (k == TAG_KIND_PROCRUSTES) ok = rd % 2 == 0 && rd <= 128 && (dd == 0 || dd == rd);
    else ok = false;
    if (!ok) return fail(nullptr, TAG_ERR_INVALID, "modality %d: kind=%d inconsistent with raw_dim=%d diff_dim=%d", m, k, rd, dd);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, TAG_ERR_CUDA, "no CUDA device");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TAG_ERR_INVALID, "device=%d of %d", cfg->device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, TAG_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, TAG_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(nullptr, TAG_ERR_CUDA, "cudaSetDevice failed");

  tag_handle* h = new tag_handle();
  h->cfg = *cfg;
  h->M = cfg->n_modalities;
  int off = 0, off16 = 0;
  for (int m = 0; m < h->M; ++m) { h->raw_off[m] = off; off += cfg->raw_dims[m]; h->raw_off16[m] = off16; off16 += round_up(cfg->raw_dims[m], 64); }
  h->raw_total = off;
  for (int m = 0; m < h->M; ++m) { h->diff_off[m] = off; off += cfg->diff_dims[m]; h->diff_off16[m] = off16; off16 += round_up(cfg->diff_dims[m], 64); }
  h->D = off; h->D16 = off16;
  if (cudaMalloc((void**)&h->zs_scale, (size_t)h->D * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&h->zs_shift, (size_t)h->D * sizeof(float)) != cudaSuccess) {
    delete h;
    return fail(nullptr, TAG_ERR_CUDA, "cudaMalloc of the z-score tables failed");
  }
  h->allocs.push_back(h->zs_scale);
  h->allocs.push_back(h->zs_shift);
  *out = h;
  return TAG_OK;
}

void tag_destroy(tag_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (void* p : h->allocs) cudaFree(p);
  for (void* p : h->w_allocs) cudaFree(p);
  if (h->k3_scratch) cudaFree(h->k3_scratch);
  if (h->tcl_A) cudaFree(h->tcl_A);
  if (h->tcl_W) cudaFree(h->tcl_W);
  if (h->tcl_part) cudaFree(h->tcl_part);
  for (auto& e : h->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  if (h->tc) tc_context_destroy(h->tc);
  delete h;
}

int tag_load_weight(tag_handle* h, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  if (!h || !key || !data || !shape || ndim < 1 || ndim > 4) return fail(h, TAG_ERR_INVALID, "tag_load_weight: bad argument");
  if (h->finalized) return fail(h, TAG_ERR_STATE, "weights already finalized");
  size_t n = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) { if (shape[i] < 1) return fail(h, TAG_ERR_INVALID, "'%s': bad shape", key); n *= (size_t)shape[i]; shp.push_back(shape[i]); }
  std::vector<float> v(n);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaMemcpy(v.data(), data, n * sizeof(float), cudaMemcpyDefault));
  h->staged[key] = std::move(v);
  h->staged_shape[key] = shp;
  return TAG_OK;
}

int tag_reload_weights_begin(tag_handle* h) {
  if (!h) return TAG_ERR_INVALID;
  if (!h->finalized) return fail(h, TAG_ERR_STATE, "tag_reload_weights_begin: the handle has no finalized weights yet");
  h->finalized = false;
  h->reloading = true;
  return TAG_OK;
}

int tag_finalize_weights(tag_handle* h) {
  if (!h) return TAG_ERR_INVALID;
  if (h->finalized) return fail(h, TAG_ERR_STATE, "weights already finalized");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const bool tc = h->cfg.precision == TAG_PRECISION_FP16_TC;
  int rc;
  if (h->reloading) {
    // new parameter values for an existing handle (a training loop between optimiser steps, BASELINE config 5): only the
    // packed weights are rebuilt; workspace, tensor-core context and tables stay
    CUDA_TRY(h, cudaDeviceSynchronize());          // launches that still read the old weights
    for (void* p : h->w_allocs) cudaFree(p);
    h->w_allocs.clear();
    for (int m = 0; m < TAG_MAX_MODALITIES; ++m) { h->state[m] = EncWeights{}; h->motion[m] = EncWeights{}; }
    h->layers.clear();
  } else if (tc && !h->tc) {
    h->tc = tc_context_create(h->cfg.device, h->err, 512);
    if (!h->tc) return TAG_ERR_CUDA;
  }
  h->packing_weights = true;
  struct Unset { tag_handle* h; ~Unset() { h->packing_weights = false; } } unset{h};
  for (int m = 0; m < h->M; ++m) {
    rc = pack_encoder(h, "state_enc." + std::to_string(m), h->cfg.raw_dims[m], &h->state[m]); if (rc) return rc;
    if (h->cfg.diff_dims[m] > 0) {
      rc = pack_encoder(h, "motion_enc." + std::to_string(m), h->cfg.diff_dims[m], &h->motion[m]); if (rc) return rc;
      if (h->cfg.precision == TAG_PRECISION_FP16_TC) {
        const auto* ps = find_w(h, "state_enc." + std::to_string(m) + ".proj.weight", {kD, kD});
        const auto* pm = find_w(h, "motion_enc." + std::to_string(m) + ".proj.weight", {kD, kD});
        if (!ps || !pm) return TAG_ERR_MISSING;
        std::vector<float> cat((size_t)kD * 2 * kD);
        for (int n = 0; n < kD; ++n)
          for (int k = 0; k < kD; ++k) { cat[(size_t)n * 2 * kD + k] = (*ps)[(size_t)n * kD + k]; cat[(size_t)n * 2 * kD + kD + k] = (*pm)[(size_t)n * kD + k]; }
        rc = upload(h, &h->state[m].projcat16, to_half(cat)); if (rc) return rc;
      }
    }
  }
  // ---- fusion (model.py:61-98): the query is input independent, fold it into one 256-vector
  const auto* latent = find_w(h, "fusion.latent", {1, 1, kD});
  const auto* qg = find_w(h, "fusion.q_ln.weight", {kD}); const auto* qb = find_w(h, "fusion.q_ln.bias", {kD});
  const auto* kg = find_w(h, "fusion.kv_ln.weight", {kD}); const auto* kb = find_w(h, "fusion.kv_ln.bias", {kD});
  const auto* Wq = find_w(h, "fusion.Wq.weight", {kD, kD}); const auto* Wk = find_w(h, "fusion.Wk.weight", {kD, kD});
  const auto* Wv = find_w(h, "fusion.Wv.weight", {kD, kD}); const auto* Wo = find_w(h, "fusion.Wo.weight", {kD, kD});
  const auto* lt = find_w(h, "fusion.logit_temp", {h->M}); const auto* lb = find_w(h, "fusion.logit_bias", {h->M});
  const auto* cls = find_w(h, "cls", {1, 1, kD});
  if (!latent || !qg || !qb || !kg || !kb || !Wq || !Wk || !Wv || !Wo || !lt || !lb || !cls) return TAG_ERR_MISSING;
  {
    double mean = 0, var = 0;
    for (int i = 0; i < kD; ++i) mean += (*latent)[i];
    mean /= kD;
    for (int i = 0; i < kD; ++i) { double d = (*latent)[i] - mean; var += d * d; }
    var /= kD;
    std::vector<double> q(kD), Q(kD);
    for (int i = 0; i < kD; ++i) q[i] = ((*latent)[i] - mean) / sqrt(var + 1e-5) * (*qg)[i] + (*qb)[i];
    for (int o = 0; o < kD; ++o) { double a = 0; for (int i = 0; i < kD; ++i) a += (double)(*Wq)[(size_t)o * kD + i] * q[i]; Q[o] = a; }
    std::vector<float> qk(kD);
    for (int k = 0; k < kD; ++k) { double a = 0; for (int o = 0; o < kD; ++o) a += (double)(*Wk)[(size_t)o * kD + k] * Q[o]; qk[k] = (float)(a / sqrt((double)kD)); }
    rc = upload(h, &h->qk, qk); if (rc) return rc;
    for (int m = 0; m < h->M; ++m) {
      const double t = (*lt)[m];
      const double sp = t > 20 ? t : log1p(exp(t));          // F.softplus
      h->inv_tau[m] = (float)(1.0 / (sp + 1e-3));
      h->lbias[m] = (*lb)[m];
    }
  }
  rc = upload(h, &h->kv_g, *kg); if (rc) return rc;
  rc = upload(h, &h->kv_b, *kb); if (rc) return rc;
  rc = upload(h, &h->Wv, *Wv); if (rc) return rc;
  rc = upload(h, &h->Wo, *Wo); if (rc) return rc;
  rc = upload(h, &h->cls, *cls); if (rc) return rc;
  if (tc) {   // Wo (Wv x) == (Wo Wv) x : one GEMM instead of two
    std::vector<float> wov((size_t)kD * kD);
    for (int o = 0; o < kD; ++o)
      for (int i = 0; i < kD; ++i) {
        double a = 0;
        for (int k = 0; k < kD; ++k) a += (double)(*Wo)[(size_t)o * kD + k] * (double)(*Wv)[(size_t)k * kD + i];
        wov[(size_t)o * kD + i] = (float)a;
      }
    rc = upload(h, &h->Wov16, to_half(wov)); if (rc) return rc;
  }
  // ---- positional table (model.py:8-16): use the checkpoint buffer when given, else rebuild it
  {
    const int need = h->cfg.max_T + 1;
    std::vector<float> pe((size_t)need * kD);
    auto it = h->staged.find("pos_enc.pe");
    if (it != h->staged.end()) {
      const auto& s = h->staged_shape["pos_enc.pe"];
      const int64_t rows = s.size() == 3 ? s[1] : (s.size() == 2 ? s[0] : 0);
      if (rows < need || s.back() != kD) return fail(h, TAG_ERR_INVALID, "pos_enc.pe has %lld rows, need %d", (long long)rows, need);
      memcpy(pe.data(), it->second.data(), pe.size() * sizeof(float));
    } else {
      for (int p = 0; p < need; ++p)
        for (int i = 0; i < kD; i += 2) {
          const float div = expf((float)i * (-logf(10000.0f) / (float)kD));
          pe[(size_t)p * kD + i] = sinf((float)p * div);
          pe[(size_t)p * kD + i + 1] = cosf((float)p * div);
        }
    }
    h->pe_rows = need;
    rc = upload(h, &h->pe, pe); if (rc) return rc;
  }
  // ---- temporal transformer
  const int F = h->cfg.ffn_dim;
  for (int l = 0; l < h->cfg.n_layers; ++l) {
    const std::string p = "temporal.layers." + std::to_string(l);
    LayerWeights L{};
    const auto* in_w = find_w(h, p + ".self_attn.in_proj_weight", {3 * kD, kD});
    const auto* in_b = find_w(h, p + ".self_attn.in_proj_bias", {3 * kD});
    const auto* out_w = find_w(h, p + ".self_attn.out_proj.weight", {kD, kD});
    const auto* out_b = find_w(h, p + ".self_attn.out_proj.bias", {kD});
    const auto* l1_w = find_w(h, p + ".linear1.weight", {F, kD}); const auto* l1_b = find_w(h, p + ".linear1.bias", {F});
    const auto* l2_w = find_w(h, p + ".linear2.weight", {kD, F}); const auto* l2_b = find_w(h, p + ".linear2.bias", {kD});
    const auto* n1_g = find_w(h, p + ".norm1.weight", {kD}); const auto* n1_b = find_w(h, p + ".norm1.bias", {kD});
    const auto* n2_g = find_w(h, p + ".norm2.weight", {kD}); const auto* n2_b = find_w(h, p + ".norm2.bias", {kD});
    if (!in_w || !in_b || !out_w || !out_b || !l1_w || !l1_b || !l2_w || !l2_b || !n1_g || !n1_b || !n2_g || !n2_b) return TAG_ERR_MISSING;
    if ((rc = upload(h, &L.in_w, *in_w)) || (rc = upload(h, &L.in_b, *in_b)) || (rc = upload(h, &L.out_w, *out_w)) ||
        (rc = upload(h, &L.out_b, *out_b)) || (rc = upload(h, &L.l1_w, *l1_w)) || (rc = upload(h, &L.l1_b, *l1_b)) ||
        (rc = upload(h, &L.l2_w, *l2_w)) || (rc = upload(h, &L.l2_b, *l2_b)) || (rc = upload(h, &L.n1_g, *n1_g)) ||
        (rc = upload(h, &L.n1_b, *n1_b)) || (rc = upload(h, &L.n2_g, *n2_g)) || (rc = upload(h, &L.n2_b, *n2_b)))
      return rc;
    if (tc) {
      if ((rc = upload(h, &L.in_w16, to_half(*in_w))) || (rc = upload(h, &L.out_w16, to_half(*out_w))) ||
          (rc = upload(h, &L.l1_w16, to_half(*l1_w))) || (rc = upload(h, &L.l2_w16, to_half(*l2_w))))
        return rc;
    }
    h->layers.push_back(L);
  }
  h->staged.clear();
  h->staged_shape.clear();
  h->packing_weights = false;
  if (h->reloading) {
    h->reloading = false;
    h->finalized = true;
    return TAG_OK;
  }

  // ---- workspace
  const size_t es = tc ? sizeof(__half) : sizeof(float);
  const size_t R = (size_t)h->cfg.max_windows * h->cfg.max_T, R2 = (size_t)h->cfg.max_windows * (h->cfg.max_T + 1);
  auto walloc = [&](void** p, size_t elems) { char* q = nullptr; int r = dev_alloc(h, &q, elems * es); *p = q; return r; };
  if ((rc = walloc(&h->bufH, R * kD)) || (rc = walloc(&h->bufY1, R * kD)) || (rc = walloc(&h->bufY2, R * kD)) ||
      (rc = walloc(&h->mix, R * kD)) || (rc = walloc(&h->fusedA, R * kD)) || (rc = walloc(&h->fusedB, R * kD)))
    return rc;
  for (int e = 0; e < 2 * h->M; ++e) if ((rc = walloc(&h->P[e], R * kD))) return rc;
  if ((rc = dev_alloc(h, &h->X, R2 * kD)) || (rc = dev_alloc(h, &h->TMP, R2 * kD))) return rc;
  if ((rc = walloc(&h->QKV, R2 * 3 * kD)) || (rc = walloc(&h->ATT, R2 * kD)) || (rc = walloc(&h->FF, R2 * F))) return rc;
  if (tc) {
    if ((rc = walloc(&h->X16, R2 * kD))) return rc;
    if ((rc = dev_alloc(h, &h->feats16, R * h->D16))) return rc;
    CUDA_TRY(h, cudaMemset(h->feats16, 0, R * h->D16 * sizeof(__half)));
    std::vector<int> t;
    for (int m = 0; m < h->M; ++m) {
      t.push_back(h->raw_off[m]); t.push_back(h->cfg.raw_dims[m]); t.push_back(h->raw_off16[m]);
      t.push_back(h->diff_off[m]); t.push_back(h->cfg.diff_dims[m]); t.push_back(h->diff_off16[m]);
    }
    if ((rc = upload(h, &h->col_tab, t))) return rc;
    const int mw = h->cfg.max_windows;
    if ((rc = dev_alloc(h, &h->iota, mw)) || (rc = dev_alloc(h, &h->zeros, mw)) || (rc = dev_alloc(h, &h->row0, (size_t)h->M * kD))) return rc;
    k_iota_zero<<<(mw + 255) / 256, 256>>>(h->iota, h->zeros, mw);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaDeviceSynchronize());
#ifdef TAG_EXPERIMENTS
    const char* env = getenv("TAG_FRAME_TABLE");
    if (env != nullptr) h->frame_table = atoi(env);
    env = getenv("TAG_FUSE_TAIL");
    if (env != nullptr) h->fuse_tail = atoi(env);
    env = getenv("TAG_FUSE_TCN");
    if (env != nullptr) h->fuse_tcn = atoi(env);
#endif
  } else {
    if ((rc = dev_alloc(h, &h->feats, R * h->D))) return rc;
  }
  if ((rc = dev_alloc(h, &h->clip_wv, h->cfg.max_windows)) || (rc = dev_alloc(h, &h->clip_ws, h->cfg.max_windows))) return rc;
  h->finalized = true;
  return TAG_OK;
}

int tag_feature_fuse(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv,
                     const int32_t* win_video, const int32_t* win_start, int64_t n_windows, int32_t T,
                     float* feats_out, int32_t* flags_out, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (n_windows < 0 || T < 1) return fail(h, TAG_ERR_INVALID, "n_windows=%lld T=%d", (long long)n_windows, T);
  if (n_windows == 0) return TAG_OK;
  if (!win_video || !win_start || !feats_out) return fail(h, TAG_ERR_INVALID, "tag_feature_fuse: NULL argument");
  FuseParams p;
  int rc = fill_fuse_params(h, &p, vids, mean, stdv, win_video, win_start, n_windows, T);
  if (rc) return rc;
  p.feats = feats_out; p.feats16 = nullptr; p.flags = flags_out;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_zscore_table(mean, stdv, h->zs_scale, h->zs_shift, h->D, (cudaStream_t)stream));
  LAUNCH_TRY(h, launch_feature_fuse(p, (cudaStream_t)stream));
  h->launches += feature_fuse_launches(p) - 1;
  return TAG_OK;
}

int tag_debug_feature_fuse16(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv,
                             const int32_t* win_video, const int32_t* win_start, int64_t n_windows, int32_t T,
                             void* feats16_out, int32_t* d16_out, int32_t* flags_out, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (d16_out) *d16_out = h->D16;
  if (!feats16_out) return TAG_OK;
  if (n_windows < 0 || T < 1) return fail(h, TAG_ERR_INVALID, "n_windows=%lld T=%d", (long long)n_windows, T);
  if (n_windows == 0) return TAG_OK;
  if (!win_video || !win_start) return fail(h, TAG_ERR_INVALID, "tag_debug_feature_fuse16: NULL argument");
  FuseParams p;
  int rc = fill_fuse_params(h, &p, vids, mean, stdv, win_video, win_start, n_windows, T);
  if (rc) return rc;
  p.feats = nullptr; p.feats16 = reinterpret_cast<__half*>(feats16_out); p.flags = flags_out;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  // the generic kernel leaves pad columns untouched (the encoder's own buffer is zeroed once at allocation)
  CUDA_TRY(h, cudaMemsetAsync(feats16_out, 0, (size_t)n_windows * T * h->D16 * sizeof(__half), (cudaStream_t)stream));
  LAUNCH_TRY(h, launch_zscore_table(mean, stdv, h->zs_scale, h->zs_shift, h->D, (cudaStream_t)stream));
  LAUNCH_TRY(h, launch_feature_fuse(p, (cudaStream_t)stream));
  h->launches += feature_fuse_launches(p) - 1;
  return TAG_OK;
}

int tag_set_fusion_attn_out(tag_handle* h, float* attn) {
  if (!h) return TAG_ERR_INVALID;
  h->attn_out = attn;
  return TAG_OK;
}

int tag_encode(tag_handle* h, const float* feats, int64_t n_windows, int32_t T, float* seq_embed, float* frame_embeds,
               float* tokens, float* tc_window, void* stream) {
  int rc = check_common(h, n_windows, T);
  if (rc) return rc;
  if (n_windows == 0) return TAG_OK;
  if (!feats || !seq_embed) return fail(h, TAG_ERR_INVALID, "tag_encode: feats / seq_embed is NULL");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc = h->cfg.precision == TAG_PRECISION_FP16_TC;
  prof_begin(h);
  const int* tab = h->col_tab;
  const int S = T + 1;
  for (int64_t w0 = 0; w0 < n_windows; w0 += h->cfg.max_windows) {
    const int64_t W = (n_windows - w0 < h->cfg.max_windows) ? n_windows - w0 : h->cfg.max_windows;
    float* seq = seq_embed + w0 * kD;
    float* fr = frame_embeds ? frame_embeds + w0 * S * kD : nullptr;
    float* tk = tokens ? tokens + w0 * S * kD : nullptr;
    float* tw = tc_window ? tc_window + w0 : nullptr;
    const float* f = feats + w0 * (int64_t)T * h->D;
    float* at = h->attn_out ? h->attn_out + w0 * (int64_t)T * h->M : nullptr;
    if (tc) {
      k_feats_to_half<<<(unsigned)(W * T), 256, 0, s>>>(f, h->feats16, W * T, h->D, h->D16, h->M, tab);
      h->launches++;
      CUDA_TRY(h, cudaGetLastError());
      rc = encode_chunk_tc(h, s, h->feats16, W, T, seq, fr, tk, tw, nullptr, at);
    } else {
      rc = encode_chunk_f32(h, s, f, W, T, seq, fr, tk, tw, at);
    }
    if (rc) break;
  }
  h->attn_out = nullptr;                       // one-shot: applies to this call only
  if (rc) return rc;
  return prof_end(h, s);
}

int tag_encode_windows(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv,
                       const int32_t* win_video, const int32_t* win_start, int64_t n_windows, int32_t T,
                       float* seq_embed, float* frame_embeds, float* tokens, float* tc_window, int32_t* flags_out,
                       void* stream) {
  int rc = check_common(h, n_windows, T);
  if (rc) return rc;
  if (n_windows == 0) return TAG_OK;
  if (!win_video || !win_start || !seq_embed) return fail(h, TAG_ERR_INVALID, "tag_encode_windows: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc = h->cfg.precision == TAG_PRECISION_FP16_TC;
  prof_begin(h);
  const int S = T + 1;
  LAUNCH_TRY(h, launch_zscore_table(mean, stdv, h->zs_scale, h->zs_shift, h->D, s));
  for (int64_t w0 = 0; w0 < n_windows; w0 += h->cfg.max_windows) {
    const int64_t W = (n_windows - w0 < h->cfg.max_windows) ? n_windows - w0 : h->cfg.max_windows;
    FuseParams p;
    rc = fill_fuse_params(h, &p, vids, mean, stdv, win_video + w0, win_start + w0, W, T);
    if (rc) return rc;
    p.feats = tc ? nullptr : h->feats;
    p.feats16 = tc ? h->feats16 : nullptr;
    p.flags = flags_out;
    { ProfScope ps(h, s, 3, (double)W * T * ((double)h->raw_total * 4 + (tc ? (double)h->D16 * 2 : (double)h->D * 4))); LAUNCH_TRY(h, launch_feature_fuse(p, s)); h->launches += feature_fuse_launches(p) - 1; }
    float* seq = seq_embed + w0 * kD;
    float* fr = frame_embeds ? frame_embeds + w0 * S * kD : nullptr;
    float* tk = tokens ? tokens + w0 * S * kD : nullptr;
    float* tw = tc_window ? tc_window + w0 : nullptr;
    rc = tc ? encode_chunk_tc(h, s, h->feats16, W, T, seq, fr, tk, tw) : encode_chunk_f32(h, s, h->feats, W, T, seq, fr, tk, tw);
    if (rc) return rc;
  }
  return prof_end(h, s);
}

int tag_encode_clips(tag_handle* h, const tag_videos* vids, const float* mean, const float* stdv, int64_t n_videos,
                     int32_t L, int32_t T, int32_t stride, float* seq_embed, float* frame_embeds, float* tokens,
                     float* tc_window, int32_t* flags_out, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (n_videos < 0 || L < 1 || T < 1 || stride < 1 || L < T)
    return fail(h, TAG_ERR_INVALID, "tag_encode_clips: n_videos=%lld L=%d T=%d stride=%d (need L >= T)", (long long)n_videos, L, T, stride);
  const int wpv = (L - T) / stride + 1;                  // windows per clip: starts 0, stride, ... <= L - T (eval.py:358-359 grid)
  const int64_t n_windows = n_videos * wpv;
  int rc = check_common(h, n_windows, T);
  if (rc) return rc;
  if (n_windows == 0) return TAG_OK;
  if (!vids || !vids->frame_offset || !seq_embed) return fail(h, TAG_ERR_INVALID, "tag_encode_clips: NULL argument");
  if (vids->n_videos < n_videos) return fail(h, TAG_ERR_INVALID, "tag_encode_clips: vids holds %lld videos, %lld requested", (long long)vids->n_videos, (long long)n_videos);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc = h->cfg.precision == TAG_PRECISION_FP16_TC;
  const int S = T + 1;
  const int64_t cap_rows = (int64_t)h->cfg.max_windows * h->cfg.max_T;        // rows of feats16
  int64_t vpp = h->cfg.max_windows / wpv;                                    // clips per pass
  if (vpp > cap_rows / L) vpp = cap_rows / L;
  const bool frame_mode = tc && h->frame_table && T >= 16 && T <= 128 && (T & (T - 1)) == 0 && vpp >= 1;
  prof_begin(h);
  LAUNCH_TRY(h, launch_zscore_table(mean, stdv, h->zs_scale, h->zs_shift, h->D, s));
  if (frame_mode) {
    for (int m = 0; m < h->M; ++m) {
      if (h->cfg.diff_dims[m] <= 0) continue;
      k_row0_vec<<<kD / 8, 256, 0, s>>>(h->zs_shift + h->diff_off[m], h->motion[m].stem16, h->cfg.diff_dims[m], h->motion[m].k16,
                                        h->row0 + (size_t)m * kD);
      h->launches++;
    }
    CUDA_TRY(h, cudaGetLastError());
    for (int64_t v0 = 0; v0 < n_videos; v0 += vpp) {
      const int64_t V = (n_videos - v0 < vpp) ? n_videos - v0 : vpp;
      const int64_t W = V * wpv, w0 = v0 * wpv;
      // K1 over whole clips: "window" i = clip v0 + i, start 0, L frames -> one table row per source frame
      tag_videos sub = *vids;
      sub.frame_offset = vids->frame_offset + v0;
      sub.n_videos = vids->n_videos - v0;
      FuseParams p;
      rc = fill_fuse_params(h, &p, &sub, mean, stdv, h->iota, h->zeros, V, L);
      if (rc) return rc;
      p.feats = nullptr; p.feats16 = h->feats16; p.flags = flags_out;
      { ProfScope ps(h, s, 3, (double)V * L * ((double)h->raw_total * 4 + (double)h->D16 * 2)); LAUNCH_TRY(h, launch_feature_fuse(p, s)); h->launches += feature_fuse_launches(p) - 1; }
      ClipGather cg; cg.L = L; cg.wpv = wpv; cg.stride = stride; cg.rows = V * L;
      rc = encode_chunk_tc(h, s, h->feats16, W, T, seq_embed + w0 * kD, frame_embeds ? frame_embeds + w0 * S * kD : nullptr,
                           tokens ? tokens + w0 * S * kD : nullptr, tc_window ? tc_window + w0 : nullptr, &cg);
      if (rc) return rc;
    }
    return prof_end(h, s);
  }
  // general path: build each pass's window table on the device and run the window pipeline
  for (int64_t w0 = 0; w0 < n_windows; w0 += h->cfg.max_windows) {
    const int64_t W = (n_windows - w0 < h->cfg.max_windows) ? n_windows - w0 : h->cfg.max_windows;
    k_clip_windows<<<(unsigned)((W + 255) / 256), 256, 0, s>>>(h->clip_wv, h->clip_ws, w0, (int)W, wpv, stride);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    FuseParams p;
    rc = fill_fuse_params(h, &p, vids, mean, stdv, h->clip_wv, h->clip_ws, W, T);
    if (rc) return rc;
    p.feats = tc ? nullptr : h->feats;
    p.feats16 = tc ? h->feats16 : nullptr;
    p.flags = flags_out;
    { ProfScope ps(h, s, 3, (double)W * T * ((double)h->raw_total * 4 + (tc ? (double)h->D16 * 2 : (double)h->D * 4))); LAUNCH_TRY(h, launch_feature_fuse(p, s)); h->launches += feature_fuse_launches(p) - 1; }
    float* seq = seq_embed + w0 * kD;
    float* fr = frame_embeds ? frame_embeds + w0 * S * kD : nullptr;
    float* tk = tokens ? tokens + w0 * S * kD : nullptr;
    float* tw = tc_window ? tc_window + w0 : nullptr;
    rc = tc ? encode_chunk_tc(h, s, h->feats16, W, T, seq, fr, tk, tw) : encode_chunk_f32(h, s, h->feats, W, T, seq, fr, tk, tw);
    if (rc) return rc;
  }
  return prof_end(h, s);
}

int tag_centroid_accumulate(tag_handle* h, const float* z, const int32_t* labels, int64_t n, int32_t C,
                            float* sums_counts, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (n < 0 || C < 1 || C > 200) return fail(h, TAG_ERR_INVALID, "n=%lld C=%d (1..200 classes)", (long long)n, C);
  if (n == 0) return TAG_OK;
  if (!z || !labels || !sums_counts) return fail(h, TAG_ERR_INVALID, "tag_centroid_accumulate: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const size_t need = centroid_scratch_floats(n, C);
  if (need > h->k3_scratch_floats) {          // rare: the first call, or a larger (n, C) than any before
    if (h->k3_scratch != nullptr) {
      CUDA_TRY(h, cudaDeviceSynchronize());   // earlier launches may still read the old buffer
      CUDA_TRY(h, cudaFree(h->k3_scratch));
      h->k3_scratch = nullptr; h->k3_scratch_floats = 0;
    }
    CUDA_TRY(h, cudaMalloc((void**)&h->k3_scratch, need * sizeof(float)));
    h->k3_scratch_floats = need;
  }
  LAUNCH_TRY(h, launch_centroid_accumulate(z, labels, n, C, sums_counts, h->k3_scratch, (cudaStream_t)stream));
  h->launches++;                              // two kernels: per-CTA partials, then the fixed-order combine
  return TAG_OK;
}

int tag_centroid_finalize(tag_handle* h, const float* sums_counts, int32_t C, float* centroids, float* counts, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (C < 1 || !sums_counts || !centroids) return fail(h, TAG_ERR_INVALID, "tag_centroid_finalize: bad argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_centroid_finalize(sums_counts, C, centroids, counts, (cudaStream_t)stream));
  return TAG_OK;
}

int tag_score(tag_handle* h, const float* seq_embeds, const float* tc_window, const int64_t* seg_offsets,
              const int32_t* video_label, const float* centroids, int32_t C, int64_t n_videos, float* ac_out,
              float* tc_out, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (n_videos < 0 || C < 1) return fail(h, TAG_ERR_INVALID, "n_videos=%lld C=%d", (long long)n_videos, C);
  if (n_videos == 0) return TAG_OK;
  if (!seg_offsets) return fail(h, TAG_ERR_INVALID, "tag_score: seg_offsets is NULL");
  if (seq_embeds && (!video_label || !centroids || !ac_out)) return fail(h, TAG_ERR_INVALID, "tag_score: AC needs video_label, centroids and ac_out");
  if (!seq_embeds && !(tc_window && tc_out)) return fail(h, TAG_ERR_INVALID, "tag_score: nothing to compute");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_score(seq_embeds, tc_window, seg_offsets, video_label, centroids, C, n_videos, ac_out, tc_out,
                             (cudaStream_t)stream));
  return TAG_OK;
}

int tag_window_tc(tag_handle* h, const float* frame_embeds, int64_t n_windows, int32_t S, float* tc_window, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (n_windows < 0 || S < 1) return fail(h, TAG_ERR_INVALID, "n_windows=%lld S=%d", (long long)n_windows, S);
  if (n_windows == 0) return TAG_OK;
  if (!frame_embeds || !tc_window) return fail(h, TAG_ERR_INVALID, "tag_window_tc: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_window_tc(frame_embeds, n_windows, S, tc_window, (cudaStream_t)stream));
  return TAG_OK;
}

int tag_stats_accumulate(tag_handle* h, const float* x, int64_t rows, int32_t D, double* sum, double* sumsq, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (rows < 0 || D < 1) return fail(h, TAG_ERR_INVALID, "rows=%lld D=%d", (long long)rows, D);
  if (rows == 0) return TAG_OK;
  if (!x || !sum || !sumsq) return fail(h, TAG_ERR_INVALID, "tag_stats_accumulate: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_stats_accumulate(x, rows, D, sum, sumsq, (cudaStream_t)stream));
  return TAG_OK;
}

int tag_tcl_forward(tag_handle* h, const float* z, const int32_t* targets, int64_t B, float temperature, float k1,
                    float k2, float* loss_rows, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (B < 0 || temperature <= 0.f) return fail(h, TAG_ERR_INVALID, "B=%lld temperature=%f", (long long)B, temperature);
  if (B == 0) return TAG_OK;
  if (!z || !targets || !loss_rows) return fail(h, TAG_ERR_INVALID, "tag_tcl_forward: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (B < 512) {            // a handful of tiles: one warp per anchor row on the CUDA cores (fp32 dot products)
    LAUNCH_TRY(h, launch_tcl_forward(z, targets, B, temperature, k1, k2, loss_rows, s));
    return TAG_OK;
  }
  // Z Z^T on the tensor cores (tcgen05, split-fp16 operands: hi.hi + lo.hi + hi.lo, fp32 accumulate), the masked row sums in
  // the GEMM epilogue — the [B,B] similarity matrix never exists in memory
  if (!h->tc) {
    h->tc = tc_context_create(h->cfg.device, h->err, 512);
    if (!h->tc) return TAG_ERR_CUDA;
  }
  const int64_t Bp = (B + 255) / 256 * 256;
  if (Bp > h->tcl_cap) {
    if (h->tcl_cap) CUDA_TRY(h, cudaDeviceSynchronize());
    if (h->tcl_A) cudaFree(h->tcl_A);
    if (h->tcl_W) cudaFree(h->tcl_W);
    if (h->tcl_part) cudaFree(h->tcl_part);
    h->tcl_A = h->tcl_W = nullptr; h->tcl_part = nullptr; h->tcl_cap = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->tcl_A, (size_t)Bp * 3 * kD * sizeof(__half)));
    CUDA_TRY(h, cudaMalloc((void**)&h->tcl_W, (size_t)Bp * 3 * kD * sizeof(__half)));
    CUDA_TRY(h, cudaMalloc((void**)&h->tcl_part, (size_t)Bp * (size_t)(Bp / 64) * 5 * sizeof(float)));
    h->tcl_cap = Bp;
  }
  LAUNCH_TRY(h, launch_tcl_split(z, B, Bp, h->tcl_A, h->tcl_W, s));
  GemmTC g{};
  g.A = h->tcl_A; g.M = B; g.lda = 3 * kD; g.W = h->tcl_W; g.N = (int)Bp; g.K = 3 * kD; g.taps = 1; g.dil = 1; g.T = 1;
  g.tcl_y = targets; g.tcl_part = h->tcl_part; g.tcl_inv_temp = 1.0f / temperature; g.tcl_valid = (int)B;
  h->err[0] = 0;
  int rc = gemm_tc_run(h, s, g, 2.0 * (double)B * (double)B * kD);
  if (rc) return rc;
  LAUNCH_TRY(h, launch_tcl_finish(h->tcl_part, B, (int)(Bp / 64), k1, k2, loss_rows, s));
  return TAG_OK;
}

int tag_supcon_hard_forward(tag_handle* h, const float* anchor, const float* positive, const float* hard_negative, int64_t B,
                            float temperature, float* loss_rows, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (B < 0 || temperature <= 0.f) return fail(h, TAG_ERR_INVALID, "B=%lld temperature=%f", (long long)B, temperature);
  if (B == 0) return TAG_OK;
  if (!anchor || !positive || !hard_negative || !loss_rows) return fail(h, TAG_ERR_INVALID, "tag_supcon_hard_forward: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_supcon_hard(anchor, positive, hard_negative, B, temperature, loss_rows, (cudaStream_t)stream));
  return TAG_OK;
}

int tag_gather_frames(tag_handle* h, const float* x, const int32_t* idx, int64_t B, int32_t T, int32_t D, float* out, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (B < 0 || T < 1 || D < 4 || D % 4) return fail(h, TAG_ERR_INVALID, "B=%lld T=%d D=%d (D must be a multiple of 4)", (long long)B, T, D);
  if (B == 0) return TAG_OK;
  if (!x || !idx || !out || x == out) return fail(h, TAG_ERR_INVALID, "tag_gather_frames: NULL argument or in-place call");
  if (B * (int64_t)T >= (1ll << 31)) return fail(h, TAG_ERR_INVALID, "tag_gather_frames: B*T too large");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LAUNCH_TRY(h, launch_gather_frames(x, idx, B, T, D, out, (cudaStream_t)stream));
  return TAG_OK;
}

int tag_debug_gemm_f32(tag_handle* h, const float* A, int32_t lda, const float* W, int32_t ldw, int64_t M, int32_t N,
                       int32_t K, int32_t taps, int32_t dil, int32_t T, const float* bias, const float* res, float* C,
                       int32_t act, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  return Gemm<float>::run(h, (cudaStream_t)stream, A, lda, W, ldw, M, N, K, taps, dil, T, bias, res, C, act);
}

int tag_debug_gemm_tc(tag_handle* h, const void* A, int32_t lda, const void* W, int64_t M, int32_t N, int32_t K,
                      int32_t taps, int32_t dil, int32_t T, const float* bias, const void* res16, const float* res32,
                      void* C16, float* C32, int32_t act, const float* gn_gamma, const float* gn_beta,
                      const float* ln_gamma, const float* ln_beta, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->tc) {
    h->tc = tc_context_create(h->cfg.device, h->err, 512);
    if (!h->tc) return TAG_ERR_CUDA;
  }
  GemmTC g{};
  g.A = (const __half*)A; g.M = M; g.lda = lda; g.W = (const __half*)W; g.N = N; g.K = K; g.taps = taps; g.dil = dil; g.T = T;
  g.bias = bias; g.res16 = (const __half*)res16; g.ldr = N; g.res32 = res32; g.C16 = (__half*)C16; g.ldc = N; g.C32 = C32; g.act = act;
  g.gn_gamma = gn_gamma; g.gn_beta = gn_beta; g.ln_gamma = ln_gamma; g.ln_beta = ln_beta;
  h->err[0] = 0;
  return gemm_tc_run(h, (cudaStream_t)stream, g, 2.0 * (double)M * N * K * taps);
}

int tag_debug_tlayer_tail(tag_handle* h, const void* att16, float* x32, void* x16, int64_t M, int32_t ffn_dim, const void* Wo16,
                          const void* W1_16, const void* W2_16, const float* bo, const float* b1, const float* b2, const float* ln1_g,
                          const float* ln1_b, const float* ln2_g, const float* ln2_b, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->tc) {
    h->tc = tc_context_create(h->cfg.device, h->err, 512);
    if (!h->tc) return TAG_ERR_CUDA;
  }
  TlayerTail t{};
  t.M = M; t.ffn_dim = ffn_dim; t.att16 = (const __half*)att16; t.x32 = x32; t.x16 = (__half*)x16; t.Wo16 = (const __half*)Wo16;
  t.W1_16 = (const __half*)W1_16; t.W2_16 = (const __half*)W2_16; t.bo = bo; t.b1 = b1; t.b2 = b2;
  t.ln1_g = ln1_g; t.ln1_b = ln1_b; t.ln2_g = ln2_g; t.ln2_b = ln2_b;
  h->err[0] = 0;
  cudaError_t e = launch_tlayer_tail(tc_encode_fn(h->tc), tc_num_sms(h->tc), t, (cudaStream_t)stream, h->err, 512);
  h->launches++;
  if (e != cudaSuccess) {
    if (h->err[0] == 0) fail(h, TAG_ERR_CUDA, "tlayer_tail launch failed: %s", cudaGetErrorString(e));
    return e == cudaErrorInvalidValue ? TAG_ERR_INVALID : TAG_ERR_CUDA;
  }
  return TAG_OK;
}

int tag_debug_tcn_block_plan(int64_t M, int32_t T, int32_t dil, int32_t* weight_stages, int32_t* smem_bytes, int32_t* tile_bytes) {
  int ws = 0, sb = 0, tb = 0;
  const bool ok = tcn_block_plan(M, T, dil, &ws, &sb, &tb);
  if (weight_stages) *weight_stages = ws;
  if (smem_bytes) *smem_bytes = sb;
  if (tile_bytes) *tile_bytes = tb;
  return ok ? 1 : 0;
}

int tag_debug_tcn_block(tag_handle* h, void* h16, int64_t M, int32_t T, int32_t dil, const void* W1_16, const void* W2_16,
                        const float* gn_gamma, const float* gn_beta, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->tc) {
    h->tc = tc_context_create(h->cfg.device, h->err, 512);
    if (!h->tc) return TAG_ERR_CUDA;
  }
  if (!tcn_block_supported(M, T, dil)) return fail(h, TAG_ERR_UNSUPPORTED, "tcn_block: shape not supported (M=%lld T=%d dil=%d)", (long long)M, T, dil);
  TcnBlock tb{};
  tb.M = M; tb.T = T; tb.dil = dil; tb.h16 = (__half*)h16; tb.W1_16 = (const __half*)W1_16; tb.W2_16 = (const __half*)W2_16;
  tb.gn_gamma = gn_gamma; tb.gn_beta = gn_beta;
  h->err[0] = 0;
  cudaError_t e = launch_tcn_block(tc_encode_fn(h->tc), tc_num_sms(h->tc), tb, (cudaStream_t)stream, h->err, 512);
  h->launches++;
  if (e != cudaSuccess) {
    if (h->err[0] == 0) fail(h, TAG_ERR_CUDA, "tcn_block launch failed: %s", cudaGetErrorString(e));
    return e == cudaErrorInvalidValue ? TAG_ERR_INVALID : TAG_ERR_CUDA;
  }
  return TAG_OK;
}

int tag_debug_poison_workspace(tag_handle* h, void* stream) {
  if (!h) return TAG_ERR_INVALID;
  if (!h->finalized) return fail(h, TAG_ERR_STATE, "tag_finalize_weights has not been called");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc = h->cfg.precision == TAG_PRECISION_FP16_TC;
  const size_t es = tc ? sizeof(__half) : sizeof(float);
  const size_t R = (size_t)h->cfg.max_windows * h->cfg.max_T, R2 = (size_t)h->cfg.max_windows * (h->cfg.max_T + 1);
  const size_t F = (size_t)h->cfg.ffn_dim;
  auto poison = [&](void* p, size_t bytes) { return p == nullptr ? cudaSuccess : cudaMemsetAsync(p, 0xFF, bytes, s); };   // 0xFF.. = NaN (fp16 and fp32)
  void* act[] = {h->bufH, h->bufY1, h->bufY2, h->mix, h->fusedA, h->fusedB};
  for (void* p : act) CUDA_TRY(h, poison(p, R * kD * es));
  for (int e = 0; e < 2 * h->M; ++e) CUDA_TRY(h, poison(h->P[e], R * kD * es));
  CUDA_TRY(h, poison(h->X, R2 * kD * sizeof(float)));
  CUDA_TRY(h, poison(h->TMP, R2 * kD * sizeof(float)));
  CUDA_TRY(h, poison(h->QKV, R2 * 3 * kD * es));
  CUDA_TRY(h, poison(h->ATT, R2 * kD * es));
  CUDA_TRY(h, poison(h->FF, R2 * F * es));
  CUDA_TRY(h, poison(h->X16, R2 * kD * es));
  CUDA_TRY(h, poison(h->feats16, R * h->D16 * sizeof(__half)));
  CUDA_TRY(h, poison(h->feats, R * h->D * sizeof(float)));
  CUDA_TRY(h, poison(h->row0, (size_t)h->M * kD * sizeof(float)));
  return TAG_OK;
}

int64_t tag_launch_count(const tag_handle* h) { return h ? h->launches : 0; }

int tag_set_profiling(tag_handle* h, int32_t on) {
  if (!h) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (on && h->prof.empty()) {
    h->prof.resize(16384);
    for (auto& e : h->prof) { CUDA_TRY(h, cudaEventCreate(&e.a)); CUDA_TRY(h, cudaEventCreate(&e.b)); }
  }
  h->profiling = on != 0;
  h->prof_used = 0;
  for (double& v : h->prof_acc) v = 0;
  return TAG_OK;
}

int tag_get_profile(tag_handle* h, double* out12) {
  if (!h || !out12) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc = prof_collect(h);
  if (rc) return rc;
  for (int i = 0; i < 12; ++i) out12[i] = h->prof_acc[i];
  for (int k = 4; k < kProfKinds; ++k) { out12[0] += h->prof_acc[3 * k]; out12[2] += h->prof_acc[3 * k + 2]; }   // "other" = every non-GEMM, non-K1 kernel
  return TAG_OK;
}

int tag_get_profile_kinds(tag_handle* h, double* out, int32_t n_kinds) {
  if (!h || !out || n_kinds < 1) return TAG_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc = prof_collect(h);
  if (rc) return rc;
  for (int k = 0; k < n_kinds; ++k)
    for (int j = 0; j < 3; ++j) out[3 * k + j] = k < kProfKinds ? h->prof_acc[3 * k + j] : 0.0;
  return TAG_OK;
}

}  // extern "C"
