// extern "C" surface of libtransception_sm100.so: one entry point per reference forward (see
// include/transception_sm100.h).  Each function only carves the caller's workspace and enqueues kernels.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "../../include/transception_sm100.h"
#include "common.cuh"
#include "attention.cuh"
#include "misc.cuh"
#include "fused16.cuh"
#include "ea16.cuh"
#include "fuse.cuh"
#include "loss.cuh"
#include "bwd.cuh"
#include <algorithm>
#include <mutex>
#include <unordered_map>

// ---- error state -----------------------------------------------------------------------
static thread_local char g_err[512] = "";
void tcx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static long long g_launches = 0;
int tcx_check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    tcx_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ---- per-kernel event timing --------------------------------------------------------------
#include <vector>
#include <string>
bool g_tcx_prof_on = false;
static std::string g_prof_name;
static std::vector<cudaEvent_t> g_prof_ev;   // start/stop pairs
static size_t g_prof_used = 0;
static bool g_prof_open = false;
static double g_prof_work = 0.0;
void tcx_prof_begin(const char* name, cudaStream_t st) {
  if (g_prof_name != name) return;
  if (g_prof_used + 2 > g_prof_ev.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    g_prof_ev.push_back(a); g_prof_ev.push_back(b);
  }
  cudaEventRecord(g_prof_ev[g_prof_used], st);
  g_prof_open = true;
}
void tcx_prof_end(const char* name, cudaStream_t st, double work) {
  if (!g_prof_open || g_prof_name != name) return;
  cudaEventRecord(g_prof_ev[g_prof_used + 1], st);
  g_prof_used += 2;
  g_prof_open = false;
  g_prof_work += work;
}

static int g_flag_gemm_tc = 1;
static int g_flag_flash_tc = 1;
static int g_flag_f16 = 1;
static int g_flag_ea_tc = 1;     // efficient-attention context on the tensor core (packT + gemm_tc)
static int g_flag_wgrad_tc = 1;  // Linear backward with operands read in place: MN-major wgrad kernel + MN-major-W dgrad (0 = round-1 packT path)
static int g_flag_mixtail = 0;   // fused dw+LN+GELU+fc2: bit-identical but slower (8 producer warps vs 16 in dwln), see DESIGN.md §4
int g_tcx_pdl = 1;
int g_tcx_pdl_chain = 0;      // PDL attribute on the training-row kernels launched through tcx_launch_chain (the Python host sets it: ops.PDL_CHAIN)
int g_tcx_smem_kb = 0;
int g_tcx_wgrad_ctas = 0;
int g_tcx_wgrad_idle = 0;
int g_tcx_max_ctas = 0;          // > 0: cap on the grid of the persistent tcgen05 kernels (lets kernels of parallel graph branches co-run)
bool tcx_flag_gemm_tc() { return g_flag_gemm_tc != 0; }
bool flash_tc_enabled() { return g_flag_flash_tc != 0; }

// Prepared-weight registry: fp32 weight pointer -> caller-owned fp16 copy (tcx_prepare_weight_f16).  The fp16
// pipeline is taken only when every matrix of an op has a prepared copy; otherwise the op runs its fp32/TF32 form.
static std::mutex g_w16_mu;
static std::unordered_map<const void*, const void*> g_w16;
static const __half* w16_of(const void* w32) {
  if (!g_flag_f16 || !g_flag_gemm_tc) return nullptr;
  std::lock_guard<std::mutex> lk(g_w16_mu);
  auto it = g_w16.find(w32);
  return it == g_w16.end() ? nullptr : reinterpret_cast<const __half*>(it->second);
}

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline const float* F(const void* p) { return reinterpret_cast<const float*>(p); }

struct Carver {
  float* base;
  size_t off = 0;
  explicit Carver(void* ws) : base(reinterpret_cast<float*>(ws)) {}
  float* take(size_t nfloats) {
    float* p = base + off;
    off += (nfloats + 63) / 64 * 64;   // keep 256-byte alignment
    return p;
  }
};
inline size_t rnd(size_t nfloats) { return (nfloats + 63) / 64 * 64; }

struct BridgeGeom {
  int hw[4], ch[4], off[5], ntok;
  int P, pp, red_off[4], nred;
};
inline bool bridge_geom(int S0, BridgeGeom& g) {
  if (S0 <= 0 || S0 % 8) return false;
  const int ch[4] = {64, 128, 320, 512};
  g.off[0] = 0;
  for (int k = 0; k < 4; k++) {
    g.hw[k] = S0 >> k;
    g.ch[k] = ch[k];
    g.off[k + 1] = g.off[k] + g.hw[k] * g.hw[k] * (ch[k] / 64);
  }
  g.ntok = g.off[4];
  g.P = S0 / 8;
  g.pp = g.P * g.P;
  g.red_off[0] = 0;
  g.red_off[1] = g.pp;
  g.red_off[2] = g.pp * 3;
  g.red_off[3] = g.pp * 8;
  g.nred = g.pp * 8 + (g.off[4] - g.off[3]);
  return true;
}

GemmParams gemm1(const float* A, const float* W, float* C, int M, int N, int K) {
  GemmParams p{};
  p.groups = 1; p.batch = 1;
  p.M = M; p.N = N; p.K = K; p.lda = K; p.ldw = K; p.ldc = N;
  p.g[0].A = A; p.g[0].W = W; p.g[0].C = C;
  p.g[0].epi.ldr = N;
  return p;
}

// Mix-FFN on G groups: xn -> y (+residual). hbuf/abuf: [G][B*N*C4]
int run_mixffn(int G, const float* const* xn, const void* const* const* pp, float eps, const float* const* residual,
               float* const* y, int B, int H, int W, int C, int C4, float* hbuf, float* abuf, cudaStream_t st) {
  const int M = B * H * W;
  const size_t per = (size_t)M * C4;
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, M, C4, C);
    g.groups = G;
    for (int i = 0; i < G; i++) {
      g.g[i].A = xn[i]; g.g[i].W = F(pp[i][0]); g.g[i].C = hbuf + i * per;
      g.g[i].epi.bias = F(pp[i][1]); g.g[i].epi.ldr = C4;
    }
    TCX_TRY(launch_gemm(g, st));
  }
  {
    MixMidGroup mg[TCX_MAX_GROUPS];
    for (int i = 0; i < G; i++)
      mg[i] = MixMidGroup{hbuf + i * per, F(pp[i][2]), F(pp[i][3]), F(pp[i][4]), F(pp[i][5]), abuf + i * per};
    TCX_TRY(launch_mixffn_mid(mg, G, B, H, W, C4, eps, st));
  }
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, M, C, C4);
    g.groups = G;
    for (int i = 0; i < G; i++) {
      g.g[i].A = abuf + i * per; g.g[i].W = F(pp[i][6]); g.g[i].C = y[i];
      g.g[i].epi.bias = F(pp[i][7]); g.g[i].epi.residual = residual ? residual[i] : nullptr; g.g[i].epi.ldr = C;
    }
    TCX_TRY(launch_gemm(g, st));
  }
  return 0;
}

// MB attention on G groups: xn -> y = residual + proj(attn). qkv [G][M*3C], ctx [G][B*C*Ch], att [G][M*C]
int run_mb_attn(int G, const float* const* xn, const void* const* const* pp, const float* const* residual,
                float* const* y, int B, int H, int W, int C, int heads, float* qkv, float* ctx, float* att,
                cudaStream_t st) {
  const int M = B * H * W, Ch = C / heads;
  const size_t pq = (size_t)M * 3 * C, pc = (size_t)B * C * Ch, pa = (size_t)M * C;
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, M, 3 * C, C);
    g.groups = G;
    for (int i = 0; i < G; i++) {
      g.g[i].A = xn[i]; g.g[i].W = F(pp[i][0]); g.g[i].C = qkv + i * pq;
      g.g[i].epi.bias = F(pp[i][1]); g.g[i].epi.ldr = 3 * C;
    }
    TCX_TRY(launch_gemm(g, st));
  }
  {
    MbAttnArgs a{};
    a.groups = G; a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads;
    a.scale = 1.0f / sqrtf((float)Ch);
    for (int i = 0; i < G; i++) {
      a.qkv[i] = qkv + i * pq; a.ctx[i] = ctx + i * pc; a.out[i] = att + i * pa;
      for (int j = 0; j < 3; j++) { a.cw[i][j] = F(pp[i][2 + 2 * j]); a.cb[i][j] = F(pp[i][3 + 2 * j]); }
    }
    TCX_TRY(launch_mb_attention(a, st));
  }
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, M, C, C);
    g.groups = G;
    for (int i = 0; i < G; i++) {
      g.g[i].A = att + i * pa; g.g[i].W = F(pp[i][8]); g.g[i].C = y[i];
      g.g[i].epi.bias = F(pp[i][9]); g.g[i].epi.residual = residual ? residual[i] : nullptr; g.g[i].epi.ldr = C;
    }
    TCX_TRY(launch_gemm(g, st));
  }
  return 0;
}

int run_scale_reduce(const float* x, const void* const* p, float eps, float* out, int B, const BridgeGeom& g, float* ws,
                     cudaStream_t st) {
  Carver c(ws);
  const int ratio[3] = {8, 4, 2};
  SrPackArgs a{};
  const long long xs_b = (long long)g.ntok * 64;
  for (int k = 0; k < 3; k++) {
    const int r = ratio[k], Cin = g.ch[k];
    const int K = Cin * r * r, M = B * g.pp;
    float* A = c.take((size_t)M * K);
    float* conv = c.take((size_t)M * Cin);
    TCX_TRY(launch_sr_im2row(x + (long long)g.off[k] * 64, xs_b, g.hw[k], Cin, r, B, A, st));
    GemmParams gp = gemm1(A, F(p[2 * k]), conv, M, Cin, K);
    gp.g[0].epi.bias = F(p[2 * k + 1]);
    TCX_TRY(launch_gemm(gp, st));
    a.conv[k] = conv;
    a.gmul[k] = Cin / 64;
    a.pp[k] = g.pp;
  }
  a.x = x; a.xs_b = xs_b; a.raw_tok0 = g.off[3];
  for (int i = 0; i < 4; i++) a.red_off[i] = g.red_off[i];
  a.nred = g.nred; a.B = B;
  a.lnw = F(p[6]); a.lnb = F(p[7]); a.eps = eps; a.out = out;
  return launch_sr_pack_ln(a, st);
}
size_t scale_reduce_ws_floats(int B, const BridgeGeom& g) {
  const int ratio[3] = {8, 4, 2};
  size_t n = 0;
  for (int k = 0; k < 3; k++) {
    const size_t M = (size_t)B * g.pp;
    n += rnd(M * g.ch[k] * ratio[k] * ratio[k]) + rnd(M * g.ch[k]);
  }
  return n;
}


// ---- fork / join onto auxiliary streams ---------------------------------------------------------------------------
// Independent kernel chains of one forward (the four per-scale Mix-FFNs of a bridge layer, the three spatial-reduction
// convolutions) are enqueued on per-device auxiliary streams between a fork and a join event.  Eagerly they overlap on
// the GPU; under stream capture the events become fork/join edges, so the replayed CUDA graph has parallel branches.
constexpr int TCX_AUX = 3;
struct AuxStreams {
  cudaStream_t s[TCX_AUX] = {};
  cudaEvent_t fork = nullptr, join[TCX_AUX] = {};
  bool ok = false;
};
static std::mutex g_aux_mu;
// one set per (device, calling stream): the legacy default stream has the same handle on every device, and one process may
// drive several GPUs (nn.DataParallel, trainer.py:110-111)
static std::unordered_map<unsigned long long, AuxStreams> g_aux;
static int g_flag_fork = 1;
AuxStreams* aux_streams(cudaStream_t st) {
  if (!g_flag_fork) return nullptr;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  AuxStreams& a = g_aux[(unsigned long long)reinterpret_cast<uintptr_t>(st) * 64ull + (unsigned long long)(dev & 63)];
  if (!a.ok) {
    bool good = cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < TCX_AUX && good; i++)
      good = cudaStreamCreateWithFlags(&a.s[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&a.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!good) { cudaGetLastError(); return nullptr; }
    a.ok = true;
  }
  return &a;
}
// aux streams [0, n) start after everything enqueued on `st` so far
inline int fork_streams(AuxStreams* a, cudaStream_t st, int n) {
  if (cudaEventRecord(a->fork, st) != cudaSuccess) { tcx_set_error("fork: cudaEventRecord failed"); return -1; }
  for (int i = 0; i < n; i++)
    if (cudaStreamWaitEvent(a->s[i], a->fork, 0) != cudaSuccess) { tcx_set_error("fork: cudaStreamWaitEvent failed"); return -1; }
  return 0;
}
// `waiter` continues after everything enqueued on aux stream i
inline int join_stream(AuxStreams* a, int i, cudaStream_t waiter) {
  if (cudaEventRecord(a->join[i], a->s[i]) != cudaSuccess || cudaStreamWaitEvent(waiter, a->join[i], 0) != cudaSuccess) {
    tcx_set_error("join: event record/wait failed");
    return -1;
  }
  return 0;
}

// ---- fp16-intermediate pipeline ---------------------------------------------------------------------------------
inline __half* H16(float* p) { return reinterpret_cast<__half*>(p); }

// optional LayerNorm second output of a GEMM (fused in the tensor-core epilogue when the row is one 64-column group)
struct LnOut {
  const float* w = nullptr;
  const float* b = nullptr;
  float eps = 0.f;
  __half* out = nullptr;
};
inline void set_ln(GemmEpi& e, const LnOut& ln, int ld, long long stride) {
  e.ln_w = ln.w; e.ln_b = ln.b; e.ln_eps = ln.eps; e.ln_out = ln.out; e.ld_ln = ld; e.stride_ln = stride;
}

struct Mix16 {
  const __half* xn;  long long xn_bs;     // [B*N, C] dense (xn_bs = 0) or per-image slabs with batch stride xn_bs
  const __half* w1;  const float* b1;
  const float* dww;  const float* dwb;  const float* lnw;  const float* lnb;
  const __half* w2;  const float* b2;
  const float* res;  long long res_bs;
  float* y;          long long y_bs;
  LnOut ln;          long long ln_bs;     // optional fp16 LayerNorm (64-column groups) of y, fused in fc2's epilogue
  float* u_save;                          // training: fp32 copy of the LayerNorm input dw3x3(h)+b+h, kept for backward
};
// all matrices of a Mix-FFN parameter block {fc1_w,fc1_b,dw_w,dw_b,ln_w,ln_b,fc2_w,fc2_b} prepared?
inline bool mix16_fill(const void* const* p, Mix16& m) {
  m.w1 = w16_of(p[0]); m.w2 = w16_of(p[6]);
  m.b1 = F(p[1]); m.dww = F(p[2]); m.dwb = F(p[3]); m.lnw = F(p[4]); m.lnb = F(p[5]); m.b2 = F(p[7]);
  return m.w1 && m.w2;
}
// y = res + fc2(GELU(LN(dw3x3(fc1 xn) + fc1 xn))) with fp16 h / a buffers ([G][B*N*C4] halfs each)
int run_mixffn16(int G, const Mix16* m, float eps, int B, int H, int W, int C, int C4, __half* hbuf, __half* abuf,
                 cudaStream_t st) {
  const int N = H * W;
  const bool strided = m[0].xn_bs != 0;
  const size_t per = (size_t)B * N * C4;
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, strided ? N : B * N, C4, C);
    g.groups = G; g.ab16 = 1; g.out16 = 1;
    if (strided) { g.batch = B; g.strideA = m[0].xn_bs; g.strideC = (long long)N * C4; }
    for (int i = 0; i < G; i++) {
      g.g[i].A = F(m[i].xn); g.g[i].W = F(m[i].w1); g.g[i].C = reinterpret_cast<float*>(hbuf + i * per);
      g.g[i].epi.bias = m[i].b1;
    }
    TCX_TRY(launch_gemm(g, st));
  }
  if (g_flag_mixtail && !m[0].ln.out && !m[0].u_save && mixtail_eligible(G, C4, (long long)B * N)) {
    // dw3x3 + skip + LN + GELU feed fc2's A tile through shared memory: one kernel, no [M, C4] round trip
    MixTailDesc d[TCX_MAX_GROUPS];
    for (int i = 0; i < G; i++)
      d[i] = MixTailDesc{hbuf + i * per, m[i].dww, m[i].dwb, m[i].lnw, m[i].lnb, m[i].w2, m[i].b2, m[i].res, m[i].y};
    return launch_mixtail(d, G, B, H, W, C4, eps, strided ? m[0].res_bs : 0, strided ? m[0].y_bs : 0, st);
  }
  {
    DwLnArgs a{};
    a.B = B; a.H = H; a.W = W; a.C = C4; a.eps = eps; a.gelu = 1;
    for (int i = 0; i < G; i++)
      a.g[i] = DwLnGroup{hbuf + i * per, m[i].dww, m[i].dwb, m[i].lnw, m[i].lnb, m[i].u_save, abuf + i * per};
    TCX_TRY(launch_dwln(a, G, true, st));
  }
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, strided ? N : B * N, C, C4);
    g.groups = G; g.ab16 = 1;
    if (strided) { g.batch = B; g.strideA = (long long)N * C4; g.strideC = m[0].y_bs; }
    for (int i = 0; i < G; i++) {
      g.g[i].A = F(abuf + i * per); g.g[i].W = F(m[i].w2); g.g[i].C = m[i].y;
      g.g[i].epi.bias = m[i].b2; g.g[i].epi.residual = m[i].res; g.g[i].epi.ldr = C; g.g[i].epi.strideR = m[i].res_bs;
      if (m[i].ln.out) set_ln(g.g[i].epi, m[i].ln, C, m[i].ln_bs);
    }
    TCX_TRY(launch_gemm(g, st));
  }
  return 0;
}

int run_ln16(int G, const float* const* x, const float* const* w, const float* const* b, __half* const* y16,
             float* const* y32, long long M, int C, float eps, cudaStream_t st) {
  Ln16Args a{};
  a.M = M; a.C = C; a.eps = eps;
  for (int i = 0; i < G; i++) a.g[i] = Ln16Group{x[i], w[i], b[i], y16 ? y16[i] : nullptr, y32 ? y32[i] : nullptr};
  return launch_ln16(a, G, st);
}
int run_ln16_1(const float* x, const float* w, const float* b, __half* y16, float* y32, long long M, int C, float eps,
               cudaStream_t st) {
  const float* xs[1] = {x}; const float* ws[1] = {w}; const float* bs[1] = {b};
  __half* y16s[1] = {y16}; float* y32s[1] = {y32};
  return run_ln16(1, xs, ws, bs, y16 ? y16s : nullptr, y32 ? y32s : nullptr, M, C, eps, st);
}

static inline int fuse_np(int N) { return (N + 63) / 64 * 64; }
// floats carved by run_eff_attn16 from its workspace
static size_t eff_attn16_carve_floats(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  int KS, Ks;
  ea16_ctx_tc_splits(N, &KS, &Ks);
  const size_t bcp = (size_t)B * KS * C * Ks;
  return rnd(3 * bnc / 2 + 64) + 2 * rnd(bnc / 2 + 64) + rnd((size_t)B * C * C / 2 + 64) + rnd(ea16_workspace_floats(B, N, C)) +
         rnd(ea16_ctx_tc_stats_floats(B, N, C) + 64) + 2 * rnd(bcp / 2 + 64) + rnd((size_t)B * KS * C * C + 64);
}
// efficient / channel attention with fp16 K/Q/V, context and attention output; y = residual + reproj(att) in fp32
int run_eff_attn16(const __half* xn16, const void* const* p, const float* residual, float* y, int B, int N, int C,
                   int reinterpret, float* ws, cudaStream_t st, const LnOut& ln = LnOut()) {
  Carver c(ws);
  const size_t bnc = (size_t)B * N * C;
  __half* kqv = H16(c.take(3 * bnc / 2 + 64));
  __half* qsm = H16(c.take(bnc / 2 + 64));
  __half* att = H16(c.take(bnc / 2 + 64));
  __half* ctxT = H16(c.take((size_t)B * C * C / 2 + 64));
  float* part = c.take(ea16_workspace_floats(B, N, C));
  const int M = B * N;
  GemmParams g = gemm1(F(xn16), nullptr, nullptr, M, C, C);
  g.groups = 3; g.ab16 = 1; g.out16 = 1;
  Ea16View v{};
  for (int i = 0; i < 3; i++) {
    g.g[i].A = F(xn16); g.g[i].W = F(w16_of(p[2 * i])); g.g[i].epi.bias = F(p[2 * i + 1]);
    g.g[i].C = reinterpret_cast<float*>(reinterpret ? kqv + i * bnc : kqv + i * C);
  }
  if (reinterpret) {
    g.ldc = C;
    v = Ea16View{kqv, kqv + bnc, kqv + 2 * bnc, (long long)N * C, C, 1};
  } else {
    g.ldc = 3 * C;
    v = Ea16View{kqv, kqv + C, kqv + 2 * C, (long long)N * 3 * C, 3 * C, 0};
  }
  TCX_TRY(launch_gemm(g, st));
  {  // the context (partials + combine) and the query softmax are independent: run them side by side
    AuxStreams* aux = aux_streams(st);
    if (aux) TCX_TRY(fork_streams(aux, st, 1));
    TCX_TRY(launch_ea16_qsoftmax(v, B, N, C, qsm, aux ? aux->s[0] : st));
    if (g_flag_ea_tc && !reinterpret && C <= 512) {
      // context on the tensor core: column-softmax probabilities and values re-laid K-major and split-major
      // ([B][KS][C][Ks]), partial[b][s][cv][ck] = sum_n Vt * Pt as one batched GEMM, then a fold over the K-splits
      int KS, Ks;
      ea16_ctx_tc_splits(N, &KS, &Ks);
      float* stats = c.take(ea16_ctx_tc_stats_floats(B, N, C) + 64);
      __half* Pt = H16(c.take((size_t)B * KS * C * Ks / 2 + 64));
      __half* Vt = H16(c.take((size_t)B * KS * C * Ks / 2 + 64));
      float* cpart = c.take((size_t)B * KS * C * C + 64);
      TCX_TRY(launch_ea16_packT(v, B, N, C, stats, Pt, Vt, st));
      GemmParams cg = gemm1(F(Vt), F(Pt), cpart, C, C, Ks);
      cg.batch = B * KS; cg.strideA = (long long)C * Ks; cg.strideW = (long long)C * Ks; cg.strideC = (long long)C * C;
      cg.ab16 = 1;
      TCX_TRY(launch_gemm(cg, st));
      TCX_TRY(launch_ea16_splitk_combine(cpart, ctxT, B, KS, C, st));
    } else {
      TCX_TRY(launch_ea16_context(v, B, N, C, part, ctxT, st));
    }
    if (aux) TCX_TRY(join_stream(aux, 0, st));
  }
  {  // att[b] = qsm[b] (N x C) * ctx[b] (C x C): W = ctxT[b]
    GemmParams a = gemm1(F(qsm), F(ctxT), reinterpret_cast<float*>(att), N, C, C);
    a.batch = B; a.strideA = (long long)N * C; a.strideW = (long long)C * C; a.strideC = (long long)N * C;
    a.ab16 = 1; a.out16 = 1;
    TCX_TRY(launch_gemm(a, st));
  }
  GemmParams r = gemm1(F(att), F(w16_of(p[6])), y, M, C, C);
  r.ab16 = 1;
  r.g[0].epi.bias = F(p[7]);
  r.g[0].epi.residual = residual;
  if (ln.out) set_ln(r.g[0].epi, ln, C, 0);      // caller guarantees C == 64
  return launch_gemm(r, st);
}
inline bool eff_attn_prepared(const void* const* p, int N, int C, int reinterpret) {
  if (N < 32 || C % 64) return false;
  if (reinterpret && (C != 64 || N % 4)) return false;
  return w16_of(p[0]) && w16_of(p[2]) && w16_of(p[4]) && w16_of(p[6]);
}

// bridge spatial-reduction attention with fp16 q / reduced tokens / kv / attention output (MSTr.py:2267-2292)
inline bool bridge_sr_prepared(const void* const* p) {
  return w16_of(p[0]) && w16_of(p[2]) && w16_of(p[4]) && w16_of(p[6]) && w16_of(p[8]) && w16_of(p[10]);
}
int run_bridge_sr_attn16(const __half* xn16, const void* const* p, float scale, float ln_eps, const float* residual, float* y,
                         int B, const BridgeGeom& g, float* ws, cudaStream_t st, const LnOut& ln = LnOut()) {
  Carver c(ws);
  const size_t bn = (size_t)B * g.ntok * 64;
  __half* q = H16(c.take(bn / 2 + 64));
  __half* o = H16(c.take(bn / 2 + 64));
  __half* red = H16(c.take((size_t)B * g.nred * 32 + 64));
  __half* kv = H16(c.take((size_t)B * g.nred * 64 + 64));
  float* fws = c.take(flash_tc_workspace_bytes(B, g.nred) / 4 + 64);
  const int M = B * g.ntok;
  // q projection on st; the reduced-token chain (3 patchify convs in parallel -> pack+LN -> kv projection) on aux streams
  AuxStreams* aux = aux_streams(st);
  if (aux) TCX_TRY(fork_streams(aux, st, 3));
  cudaStream_t s0 = aux ? aux->s[0] : st;
  {
    GemmParams gq = gemm1(F(xn16), F(w16_of(p[0])), reinterpret_cast<float*>(q), M, 64, 64);
    gq.ab16 = 1; gq.out16 = 1;
    gq.g[0].epi.bias = F(p[1]);
    TCX_TRY(launch_gemm(gq, st));
  }
  {  // Scale_reduce: patchify convs as GEMMs on (ky,kx,cin)-ordered rows, then pack + LayerNorm
    const int ratio[3] = {8, 4, 2};
    SrPackArgs a{};
    const long long xs_b = (long long)g.ntok * 64;
    for (int k = 0; k < 3; k++) {
      const int r = ratio[k], Cin = g.ch[k];
      const int K = Cin * r * r, Mk = B * g.pp;
      __half* A = H16(c.take((size_t)Mk * K / 2 + 64));
      float* conv = c.take((size_t)Mk * Cin);
      cudaStream_t sk = aux ? aux->s[k] : st;
      TCX_TRY(launch_sr_im2row16(xn16 + (long long)g.off[k] * 64, xs_b, g.hw[k], Cin, r, B, A, sk));
      GemmParams gp = gemm1(F(A), F(w16_of(p[6 + 2 * k])), conv, Mk, Cin, K);
      gp.ab16 = 1;
      gp.g[0].epi.bias = F(p[7 + 2 * k]);
      TCX_TRY(launch_gemm(gp, sk));
      a.conv[k] = conv; a.gmul[k] = Cin / 64; a.pp[k] = g.pp;
    }
    if (aux) { TCX_TRY(join_stream(aux, 1, s0)); TCX_TRY(join_stream(aux, 2, s0)); }
    a.x = nullptr; a.xs_b = xs_b; a.raw_tok0 = g.off[3];
    for (int i = 0; i < 4; i++) a.red_off[i] = g.red_off[i];
    a.nred = g.nred; a.B = B;
    a.lnw = F(p[12]); a.lnb = F(p[13]); a.eps = ln_eps; a.out = nullptr;
    TCX_TRY(launch_sr_pack_ln16(a, xn16, red, s0));
  }
  {
    GemmParams gk = gemm1(F(red), F(w16_of(p[2])), reinterpret_cast<float*>(kv), B * g.nred, 128, 64);
    gk.ab16 = 1; gk.out16 = 1;
    gk.g[0].epi.bias = F(p[3]);
    TCX_TRY(launch_gemm(gk, s0));
  }
  if (aux) TCX_TRY(join_stream(aux, 0, st));
  TCX_TRY(launch_flash_tc16(q, kv, o, B, g.ntok, g.nred, scale, fws, st));
  GemmParams gp = gemm1(F(o), F(w16_of(p[4])), y, M, 64, 64);
  gp.ab16 = 1;
  gp.g[0].epi.bias = F(p[5]);
  gp.g[0].epi.residual = residual;
  if (ln.out) set_ln(gp.g[0].epi, ln, 64, 0);
  return launch_gemm(gp, st);
}

}  // namespace

extern "C" {

const char* tcx_version(void) { return "transception_sm100 0.1.0 (sm_100a)"; }
const char* tcx_last_error(void) { return g_err; }

int tcx_device_ok(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    tcx_set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (prop.major != 10) {
    tcx_set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    return 0;
  }
  return 1;
}

long long tcx_launch_count(void) { return g_launches; }

int tcx_profile_enable(const char* kernel_name) {
  g_prof_used = 0; g_prof_open = false; g_prof_work = 0.0;
  if (kernel_name && kernel_name[0]) { g_prof_name = kernel_name; g_tcx_prof_on = true; }
  else { g_prof_name.clear(); g_tcx_prof_on = false; }
  return 0;
}

int tcx_profile_read_work(double* total_ms, int* count, double* work) {
  const double w = g_prof_work;
  const int rc = tcx_profile_read(total_ms, count);
  *work = w;
  g_prof_work = 0.0;
  return rc;
}

int tcx_profile_read(double* total_ms, int* count) {
  double t = 0; int n = 0;
  for (size_t i = 0; i + 1 < g_prof_used; i += 2) {
    if (cudaEventSynchronize(g_prof_ev[i + 1]) != cudaSuccess) { tcx_set_error("profile: event sync failed"); return -1; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[i], g_prof_ev[i + 1]) != cudaSuccess) { tcx_set_error("profile: elapsed failed"); return -1; }
    t += ms; n++;
  }
  *total_ms = t; *count = n;
  g_prof_used = 0;
  return 0;
}

int tcx_set_flag(const char* name, int value) {
  int* f = nullptr;
  if (!strcmp(name, "gemm_tc")) f = &g_flag_gemm_tc;
  else if (!strcmp(name, "flash_tc")) f = &g_flag_flash_tc;
  else if (!strcmp(name, "f16_pipeline")) f = &g_flag_f16;
  else if (!strcmp(name, "fork")) f = &g_flag_fork;
  else if (!strcmp(name, "mixtail")) f = &g_flag_mixtail;
  else if (!strcmp(name, "ea_tc")) f = &g_flag_ea_tc;
  else if (!strcmp(name, "pdl")) f = &g_tcx_pdl;
  else if (!strcmp(name, "pdl_chain")) f = &g_tcx_pdl_chain;
  else if (!strcmp(name, "wgrad_tc")) f = &g_flag_wgrad_tc;
  else if (!strcmp(name, "max_ctas")) f = &g_tcx_max_ctas;
  else if (!strcmp(name, "smem_kb")) f = &g_tcx_smem_kb;
  else if (!strcmp(name, "wgrad_ctas")) f = &g_tcx_wgrad_ctas;
  else if (!strcmp(name, "wgrad_idle")) f = &g_tcx_wgrad_idle;
  if (!f) { tcx_set_error("unknown flag %s", name); return -1; }
  const int old = *f;
  *f = value;
  return old;
}

int tcx_layernorm_fwd(const float* x, const float* w, const float* b, float* y, long long M, int C, float eps,
                      void* stream) {
  if (C == 64 || C == 128 || C == 256 || C == 320 || C == 512) return run_ln16_1(x, w, b, nullptr, y, M, C, eps, S(stream));   // exact fp32, vectorised
  return launch_layernorm(x, w, b, y, M, C, eps, S(stream));
}

// LayerNorm with two outputs from one pass: fp32 (kept for the backward) and fp16 (the GEMM operand of the node that follows)
int tcx_layernorm_dual_fwd(const float* x, const float* w, const float* b, float* y32, void* y16, long long M, int C, float eps,
                           void* stream) {
  TCX_REQUIRE(x && w && b && y32 && y16, "layernorm_dual_fwd: null pointer");
  TCX_REQUIRE(C == 64 || C == 128 || C == 256 || C == 320 || C == 512, "layernorm_dual_fwd: C=%d is not one of 64/128/256/320/512", C);
  return run_ln16_1(x, w, b, reinterpret_cast<__half*>(y16), y32, M, C, eps, S(stream));
}

int tcx_linear_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N,
                   int K, int act, void* stream) {
  return launch_linear(x, w, bias, residual, y, M, N, K, act, S(stream));
}

int tcx_f32_to_f16(const float* src, void* dst, long long n, void* stream) {
  return launch_f32_to_f16(src, dst, n, S(stream));
}

int tcx_prepare_weight_f16(const float* w32, void* w16, long long numel, void* stream) {
  TCX_REQUIRE(w32 && w16 && numel > 0, "prepare_weight: null pointer or empty weight");
  TCX_TRY(launch_f32_to_f16(w32, w16, numel, S(stream)));
  std::lock_guard<std::mutex> lk(g_w16_mu);
  g_w16[w32] = w16;
  return 0;
}

int tcx_prepare_conv_weight_f16(const float* w32, void* w16, int N, int Cin, int r, void* stream) {
  TCX_REQUIRE(w32 && w16 && N > 0 && Cin > 0 && r > 0, "prepare_conv_weight: bad arguments");
  TCX_TRY(launch_conv_weight_perm16(w32, w16, N, Cin, r, S(stream)));
  std::lock_guard<std::mutex> lk(g_w16_mu);
  g_w16[w32] = w16;
  return 0;
}

int tcx_forget_weight(const float* w32) {
  std::lock_guard<std::mutex> lk(g_w16_mu);
  g_w16.erase(w32);
  return 0;
}

int tcx_linear_f16_fwd(const void* x16, const void* w16, const float* bias, const float* residual, void* y, int M, int N,
                       int K, int out_f16, void* stream) {
  GemmParams g = gemm1(F(x16), F(w16), reinterpret_cast<float*>(y), M, N, K);
  g.ab16 = 1; g.out16 = out_f16 ? 1 : 0;
  g.g[0].epi.bias = bias;
  g.g[0].epi.residual = residual;
  return launch_gemm(g, S(stream));
}

int tcx_linear_bn_act_fwd(const float* x, const float* w, const float* bn_w, const float* bn_b, const float* bn_rm,
                          const float* bn_rv, float bn_eps, int act, float* y, int M, int N, int K, void* stream) {
  GemmParams g = gemm1(x, w, y, M, N, K);
  g.g[0].epi.bn = BnParams{bn_w, bn_b, bn_rm, bn_rv, bn_eps};
  g.g[0].epi.act = act;
  return launch_gemm(g, S(stream));
}

int tcx_patch_embed_ln_fwd(const float* x, int B, int Cin, int H, int W, const float* w, const float* bias,
                           const float* lnw, const float* lnb, float eps, float* out, void* stream) {
  TCX_REQUIRE(Cin == 1 || Cin == 3, "patch_embed: Cin must be 1 or 3 (got %d)", Cin);
  const long long plane = (long long)H * W;
  return launch_patch_embed_ln(x, Cin * plane, Cin == 1 ? 0 : plane, B, H, W, w, bias, lnw, lnb, eps, out, S(stream));
}

int tcx_dwconv_tokens_fwd(const float* x, const float* w, const float* b, float* y, int B, int H, int W, int C,
                          int add_input, void* stream) {
  DwGroup g{x, w, b, y};
  return launch_dwconv3x3(&g, 1, B, H, W, C, 1, add_input ? DW_ADD_INPUT : DW_PLAIN, BnParams{}, S(stream));
}

// ---- K8 --------------------------------------------------------------------------------
size_t tcx_eff_attn_workspace_bytes(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  size_t part = ea_workspace_floats(B, N, C);
  if (ea16_workspace_floats(B, N, C) > part) part = ea16_workspace_floats(B, N, C);
  size_t fp32_form = rnd(3 * bnc) + rnd(bnc) + rnd(bnc) + rnd((size_t)B * C * C) + rnd(part);
  size_t f16_form = eff_attn16_carve_floats(B, N, C);
  return 4 * ((fp32_form > f16_form ? fp32_form : f16_form) + rnd(bnc) + 1024);
}

int tcx_eff_attn_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int N, int C,
                     int reinterpret, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  if (eff_attn_prepared(p, N, C, reinterpret)) {
    // standalone entry: the caller's LayerNorm output is fp32 -> one conversion pass, then the fp16 form
    const size_t bnc0 = (size_t)B * N * C;
    float* tail = reinterpret_cast<float*>(ws) + eff_attn16_carve_floats(B, N, C);
    TCX_TRY(launch_f32_to_f16(xn, tail, (long long)bnc0, st));
    return run_eff_attn16(H16(tail), p, residual, y, B, N, C, reinterpret, reinterpret_cast<float*>(ws), st);
  }
  Carver c(ws);
  const size_t bnc = (size_t)B * N * C;
  float* kqv = c.take(3 * bnc);
  float* qsm = c.take(bnc);
  float* att = c.take(bnc);
  float* ctxT = c.take((size_t)B * C * C);
  float* part = c.take(ea_workspace_floats(B, N, C));
  const int M = B * N;
  GemmParams g = gemm1(xn, nullptr, nullptr, M, C, C);
  g.groups = 3;
  EaView v{};
  for (int i = 0; i < 3; i++) {
    g.g[i].A = xn; g.g[i].W = F(p[2 * i]); g.g[i].epi.bias = F(p[2 * i + 1]);
    g.g[i].C = reinterpret ? kqv + i * bnc : kqv + i * C;
  }
  if (reinterpret) {
    g.ldc = C;
    v = EaView{kqv, kqv + bnc, kqv + 2 * bnc, (long long)N * C, (long long)N, 1};
  } else {
    g.ldc = 3 * C;
    v = EaView{kqv, kqv + C, kqv + 2 * C, (long long)N * 3 * C, 1, (long long)3 * C};
  }
  TCX_TRY(launch_gemm(g, st));
  TCX_TRY(launch_ea_context(v, reinterpret != 0, B, N, C, part, ctxT, st));
  TCX_TRY(launch_ea_qsoftmax(v, reinterpret != 0, B, N, C, qsm, st));
  {  // att[b] = qsm[b] (N x C) * ctx[b] (C x C): W = ctxT[b]
    GemmParams a = gemm1(qsm, ctxT, att, N, C, C);
    a.batch = B; a.strideA = (long long)N * C; a.strideW = (long long)C * C; a.strideC = (long long)N * C;
    TCX_TRY(launch_gemm(a, st));
  }
  GemmParams r = gemm1(att, F(p[6]), y, M, C, C);
  r.g[0].epi.bias = F(p[7]);
  r.g[0].epi.residual = residual;
  return launch_gemm(r, st);
}

// ---- K1 --------------------------------------------------------------------------------
size_t tcx_mixffn_skip_workspace_bytes(int B, int N, int C4) { return 4 * 3 * rnd((size_t)B * N * C4); }

int tcx_mixffn_skip_fwd(const float* xn, const void* const* p, float ln_eps, const float* residual, float* y, int B,
                        int H, int W, int C, int C4, void* ws, void* stream) {
  Carver c(ws);
  float* h = c.take((size_t)B * H * W * C4);
  float* a = c.take((size_t)B * H * W * C4);
  Mix16 m{};
  if (mix16_fill(p, m)) {
    __half* xn16 = H16(c.take((size_t)B * H * W * C4));
    TCX_TRY(launch_f32_to_f16(xn, xn16, (long long)B * H * W * C, S(stream)));
    m.xn = xn16; m.res = residual; m.y = y;
    return run_mixffn16(1, &m, ln_eps, B, H, W, C, C4, H16(h), H16(a), S(stream));
  }
  const float* xs[1] = {xn};
  const void* const* ps[1] = {p};
  const float* rs[1] = {residual};
  float* ys[1] = {y};
  return run_mixffn(1, xs, ps, ln_eps, residual ? rs : nullptr, ys, B, H, W, C, C4, h, a, S(stream));
}

// ---- K2 --------------------------------------------------------------------------------
size_t tcx_mb_factor_attn_workspace_bytes(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  return 4 * (rnd(3 * bnc) + rnd((size_t)B * C * C) + rnd(bnc));
}

int tcx_mb_factor_attn_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int H, int W,
                           int C, int heads, void* ws, void* stream) {
  Carver c(ws);
  const size_t bnc = (size_t)B * H * W * C;
  float* qkv = c.take(3 * bnc);
  float* ctx = c.take((size_t)B * C * C);
  float* att = c.take(bnc);
  const float* xs[1] = {xn};
  const void* const* ps[1] = {p};
  const float* rs[1] = {residual};
  float* ys[1] = {y};
  return run_mb_attn(1, xs, ps, residual ? rs : nullptr, ys, B, H, W, C, heads, qkv, ctx, att, S(stream));
}

// training forward of FactorAtt_ConvRelPosEnc on the fp16 pipeline (the kernels of the inference path): saved = fp16 q | k | v rows
// and the fp16 attention output; ws = fp16 xn + per-head contexts
size_t tcx_mb_factor_attn_saved_bytes(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  return 4 * (rnd(3 * bnc / 2 + 64) + rnd(bnc / 2 + 64) + 64);
}
size_t tcx_mb_factor_attn_train_workspace_bytes(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  return 4 * (rnd(bnc / 2 + 64) + rnd((size_t)B * C * C) + 64);
}
int tcx_mb_factor_attn_train_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int H, int W, int C,
                                 int heads, void* saved, void* ws, const void* xn16_in, void* stream) {
  TCX_REQUIRE(xn && p && y && saved && ws, "mb_factor_attn_train_fwd: null pointer");
  TCX_REQUIRE(heads == 8 && C % heads == 0, "mb_factor_attn_train_fwd: built for 8 heads");
  const __half* wqkv = w16_of(p[0]);
  const __half* wproj = w16_of(p[8]);
  TCX_REQUIRE(wqkv && wproj, "mb_factor_attn_train_fwd: qkv / proj weights are not prepared (tcx_prepare_weight_f16)");
  cudaStream_t st = S(stream);
  const int M = B * H * W, Ch = C / heads;
  const size_t bnc = (size_t)M * C;
  Carver sv(saved);
  __half* qkv16 = H16(sv.take(3 * bnc / 2 + 64));
  __half* att16 = H16(sv.take(bnc / 2 + 64));
  Carver c(ws);
  __half* xn16_ws = H16(c.take(bnc / 2 + 64));
  float* ctx = c.take((size_t)B * C * C);
  if (!xn16_in) TCX_TRY(launch_f32_to_f16(xn, xn16_ws, (long long)bnc, st));      // no fp16 twin from the LayerNorm in front
  const __half* xn16 = xn16_in ? reinterpret_cast<const __half*>(xn16_in) : xn16_ws;
  {
    GemmParams gp = gemm1(F(xn16), F(wqkv), reinterpret_cast<float*>(qkv16), M, 3 * C, C);
    gp.ab16 = 1; gp.out16 = 1;
    gp.g[0].epi.bias = F(p[1]);
    TCX_TRY(launch_gemm(gp, st));
  }
  {
    Mb16Args a{};
    a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.scale = 1.0f / sqrtf((float)Ch);
    a.qkv[0] = qkv16; a.ctx[0] = ctx; a.out[0] = att16;
    for (int j = 0; j < 3; j++) { a.cw[0][j] = F(p[2 + 2 * j]); a.cb[0][j] = F(p[3 + 2 * j]); }
    TCX_TRY(launch_mb_attention16(a, 1, st));
  }
  GemmParams gp = gemm1(F(att16), F(wproj), y, M, C, C);
  gp.ab16 = 1;
  gp.g[0].epi.bias = F(p[9]); gp.g[0].epi.residual = residual; gp.g[0].epi.ldr = C;
  return launch_gemm(gp, st);
}

// ---- MHCA blocks -------------------------------------------------------------------------
size_t tcx_mhca_blocks_workspace_bytes(int G, int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  // xa, ln, xb, att (1 each), qkv (3), h, ax (4 each), ctx
  return 4 * (size_t)G * (4 * rnd(bnc) + rnd(3 * bnc) + 2 * rnd(4 * bnc) + rnd((size_t)B * C * C));
}

int tcx_mhca_blocks_fwd(const float* x_in, float* x, const void* const* p, int G, int L, int B, int H, int W, int C,
                        int heads, float ln_eps, float mlp_ln_eps, void* ws, void* stream) {
  TCX_REQUIRE(G >= 1 && G <= TCX_MAX_GROUPS, "mhca_blocks: G=%d out of range", G);
  cudaStream_t st = S(stream);
  const int N = H * W, M = B * N;
  const size_t bnc = (size_t)M * C;
  Carver c(ws);
  float* xa = c.take(G * bnc);
  float* ln = c.take(G * bnc);
  float* xb = c.take(G * bnc);
  float* att = c.take(G * bnc);
  float* qkv = c.take(G * 3 * bnc);
  float* hb = c.take(G * 4 * bnc);
  float* ab = c.take(G * 4 * bnc);
  float* ctx = c.take((size_t)G * B * C * C);
  bool fast = true;
  for (int i = 0; i < G * L && fast; i++) {
    const void* const* b = p + (size_t)i * TCX_MHCA_NP;
    fast = w16_of(b[4]) && w16_of(b[12]) && w16_of(b[16]) && w16_of(b[22]);
  }
  if (fast) {
    // fp16 intermediates: cpe+norm1 fused (fp32 stream xa + fp16 LN), fp16 qkv / attention output / Mix-FFN hidden
    __half* ln16 = H16(ln);
    __half* qkv16 = H16(qkv);
    __half* att16 = H16(att);
    const int Ch = C / heads;
    for (int l = 0; l < L; l++) {
      const void* const* blk[TCX_MAX_GROUPS];
      for (int g = 0; g < G; g++) blk[g] = p + ((size_t)g * L + l) * TCX_MHCA_NP;
      {
        DwLnArgs a{};
        a.B = B; a.H = H; a.W = W; a.C = C; a.eps = ln_eps; a.gelu = 0;
        for (int g = 0; g < G; g++)
          a.g[g] = DwLnGroup{(l == 0 ? x_in : x) + g * bnc, F(blk[g][0]), F(blk[g][1]), F(blk[g][2]), F(blk[g][3]), xa + g * bnc,
                             ln16 + g * bnc};
        TCX_TRY(launch_dwln(a, G, false, st));
      }
      {
        GemmParams gp = gemm1(nullptr, nullptr, nullptr, M, 3 * C, C);
        gp.groups = G; gp.ab16 = 1; gp.out16 = 1;
        for (int g = 0; g < G; g++) {
          gp.g[g].A = F(ln16 + g * bnc); gp.g[g].W = F(w16_of(blk[g][4])); gp.g[g].C = reinterpret_cast<float*>(qkv16 + g * 3 * bnc);
          gp.g[g].epi.bias = F(blk[g][5]);
        }
        TCX_TRY(launch_gemm(gp, st));
      }
      {
        Mb16Args a{};
        a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.scale = 1.0f / sqrtf((float)Ch);
        for (int g = 0; g < G; g++) {
          a.qkv[g] = qkv16 + g * 3 * bnc; a.ctx[g] = ctx + (size_t)g * B * C * Ch; a.out[g] = att16 + g * bnc;
          for (int j = 0; j < 3; j++) { a.cw[g][j] = F(blk[g][6 + 2 * j]); a.cb[g][j] = F(blk[g][7 + 2 * j]); }
        }
        TCX_TRY(launch_mb_attention16(a, G, st));
      }
      {
        GemmParams gp = gemm1(nullptr, nullptr, nullptr, M, C, C);
        gp.groups = G; gp.ab16 = 1;
        for (int g = 0; g < G; g++) {
          gp.g[g].A = F(att16 + g * bnc); gp.g[g].W = F(w16_of(blk[g][12])); gp.g[g].C = xb + g * bnc;
          gp.g[g].epi.bias = F(blk[g][13]); gp.g[g].epi.residual = xa + g * bnc; gp.g[g].epi.ldr = C;
          if (C == 64) {       // norm2 in the projection's epilogue
            LnOut ln; ln.w = F(blk[g][14]); ln.b = F(blk[g][15]); ln.eps = ln_eps; ln.out = ln16 + g * bnc;
            set_ln(gp.g[g].epi, ln, C, 0);
          }
        }
        TCX_TRY(launch_gemm(gp, st));
      }
      if (C != 64) {
        const float* xs[TCX_MAX_GROUPS]; const float* ws[TCX_MAX_GROUPS]; const float* bs[TCX_MAX_GROUPS];
        __half* ys[TCX_MAX_GROUPS];
        for (int g = 0; g < G; g++) { xs[g] = xb + g * bnc; ws[g] = F(blk[g][14]); bs[g] = F(blk[g][15]); ys[g] = ln16 + g * bnc; }
        TCX_TRY(run_ln16(G, xs, ws, bs, ys, nullptr, M, C, ln_eps, st));
      }
      {
        Mix16 m[TCX_MAX_GROUPS];
        for (int g = 0; g < G; g++) {
          m[g] = Mix16{};
          mix16_fill(blk[g] + 16, m[g]);
          m[g].xn = ln16 + g * bnc; m[g].res = xb + g * bnc; m[g].y = x + g * bnc;
        }
        TCX_TRY(run_mixffn16(G, m, mlp_ln_eps, B, H, W, C, 4 * C, H16(hb), H16(ab), st));
      }
    }
    return 0;
  }
  for (int l = 0; l < L; l++) {
    const void* const* blk[TCX_MAX_GROUPS];
    for (int g = 0; g < G; g++) blk[g] = p + ((size_t)g * L + l) * TCX_MHCA_NP;
    {  // x = x + dw3x3(x) + b   (ConvPosEnc, shared weights, applied in every block)
      DwGroup dg[TCX_MAX_GROUPS];
      for (int g = 0; g < G; g++) dg[g] = DwGroup{(l == 0 ? x_in : x) + g * bnc, F(blk[g][0]), F(blk[g][1]), xa + g * bnc};
      TCX_TRY(launch_dwconv3x3(dg, G, B, H, W, C, 1, DW_ADD_INPUT, BnParams{}, st));
    }
    {
      LnGroup lg[TCX_MAX_GROUPS];
      for (int g = 0; g < G; g++) lg[g] = LnGroup{xa + g * bnc, F(blk[g][2]), F(blk[g][3]), ln + g * bnc};
      TCX_TRY(launch_layernorm_grouped(lg, G, M, C, ln_eps, st));
    }
    {
      const float* xs[TCX_MAX_GROUPS]; const void* const* ps[TCX_MAX_GROUPS];
      const float* rs[TCX_MAX_GROUPS]; float* ys[TCX_MAX_GROUPS];
      for (int g = 0; g < G; g++) { xs[g] = ln + g * bnc; ps[g] = blk[g] + 4; rs[g] = xa + g * bnc; ys[g] = xb + g * bnc; }
      TCX_TRY(run_mb_attn(G, xs, ps, rs, ys, B, H, W, C, heads, qkv, ctx, att, st));
    }
    {
      LnGroup lg[TCX_MAX_GROUPS];
      for (int g = 0; g < G; g++) lg[g] = LnGroup{xb + g * bnc, F(blk[g][14]), F(blk[g][15]), ln + g * bnc};
      TCX_TRY(launch_layernorm_grouped(lg, G, M, C, ln_eps, st));
    }
    {
      const float* xs[TCX_MAX_GROUPS]; const void* const* ps[TCX_MAX_GROUPS];
      const float* rs[TCX_MAX_GROUPS]; float* ys[TCX_MAX_GROUPS];
      for (int g = 0; g < G; g++) { xs[g] = ln + g * bnc; ps[g] = blk[g] + 16; rs[g] = xb + g * bnc; ys[g] = x + g * bnc; }
      TCX_TRY(run_mixffn(G, xs, ps, mlp_ln_eps, rs, ys, B, H, W, C, 4 * C, hb, ab, st));
    }
  }
  return 0;
}

// ---- K4 / K5 -----------------------------------------------------------------------------
size_t tcx_ripm_dwsep_bn_hs_workspace_bytes(int B, int H, int W, int C, int stride) {
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  return 4 * rnd((size_t)B * Ho * Wo * C);
}

int tcx_ripm_dwsep_bn_hs_fwd(const float* x, const float* dw_w, const float* pw_w, const float* bn_w,
                             const float* bn_b, const float* bn_rm, const float* bn_rv, float bn_eps, float* y, int B,
                             int H, int W, int C, int stride, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  float* t = reinterpret_cast<float*>(ws);
  DwGroup g{x, dw_w, nullptr, t};
  TCX_TRY(launch_dwconv3x3(&g, 1, B, H, W, C, stride, DW_PLAIN, BnParams{}, st));
  GemmParams gp = gemm1(t, pw_w, y, B * Ho * Wo, C, C);
  gp.g[0].epi.bn = BnParams{bn_w, bn_b, bn_rm, bn_rv, bn_eps};
  gp.g[0].epi.act = ACT_HARDSWISH;
  return launch_gemm(gp, st);
}

size_t tcx_resblock_workspace_bytes(int B, int H, int W, int C) { return 4 * 2 * rnd((size_t)B * H * W * C); }

int tcx_resblock_fwd(const float* x, const void* const* p, float bn_eps, float* y, int B, int H, int W, int C,
                     void* ws, void* stream) {
  cudaStream_t st = S(stream);
  Carver c(ws);
  const int M = B * H * W;
  float* t1 = c.take((size_t)M * C);
  float* t2 = c.take((size_t)M * C);
  GemmParams g1 = gemm1(x, F(p[0]), t1, M, C, C);
  g1.g[0].epi.bn = BnParams{F(p[1]), F(p[2]), F(p[3]), F(p[4]), bn_eps};
  g1.g[0].epi.act = ACT_HARDSWISH;
  TCX_TRY(launch_gemm(g1, st));
  DwGroup dg{t1, F(p[5]), nullptr, t2};
  TCX_TRY(launch_dwconv3x3(&dg, 1, B, H, W, C, 1, DW_BN_HS, BnParams{F(p[6]), F(p[7]), F(p[8]), F(p[9]), bn_eps}, st));
  GemmParams g2 = gemm1(t2, F(p[10]), y, M, C, C);
  g2.g[0].epi.bn = BnParams{F(p[11]), F(p[12]), F(p[13]), F(p[14]), bn_eps};
  g2.g[0].epi.residual = x;
  return launch_gemm(g2, st);
}

// ---- K6 ------------------------------------------------------------------------------------
size_t tcx_iff_coordatt_workspace_bytes(int B, int HW, int C, int mip) {
  const size_t inp = 4 * (size_t)C;
  return 4 * (rnd((size_t)B * 2 * HW * inp) + rnd((size_t)B * 2 * HW * mip) + 2 * rnd((size_t)B * HW * inp) +
              rnd((size_t)B * HW * HW * inp));
}

int tcx_iff_coordatt_fwd(const void* const* maps, const void* const* p, float bn_eps, float* y, int B, int HW, int C,
                         int mip, int Cout, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  Carver c(ws);
  const int inp = 4 * C;
  float* pooled = c.take((size_t)B * 2 * HW * inp);
  float* yv = c.take((size_t)B * 2 * HW * mip);
  float* ah = c.take((size_t)B * HW * inp);
  float* aw = c.take((size_t)B * HW * inp);
  float* gated = c.take((size_t)B * HW * HW * inp);
  IffSrc src{};
  for (int i = 0; i < 4; i++) src.p[i] = F(maps[i]);
  TCX_TRY(launch_iff_pool(src, B, HW, C, pooled, st));
  {
    GemmParams g = gemm1(pooled, F(p[0]), yv, B * 2 * HW, mip, inp);
    g.g[0].epi.bias = F(p[1]);
    g.g[0].epi.bn = BnParams{F(p[2]), F(p[3]), F(p[4]), F(p[5]), bn_eps};
    g.g[0].epi.act = ACT_SILU_SWISH;
    TCX_TRY(launch_gemm(g, st));
  }
  {
    GemmParams g = gemm1(nullptr, nullptr, nullptr, HW, inp, mip);
    g.groups = 2; g.batch = B;
    g.strideA = (long long)2 * HW * mip; g.strideW = 0; g.strideC = (long long)HW * inp;
    g.g[0].A = yv; g.g[0].W = F(p[6]); g.g[0].C = ah; g.g[0].epi.bias = F(p[7]); g.g[0].epi.act = ACT_SIGMOID;
    g.g[1].A = yv + (size_t)HW * mip; g.g[1].W = F(p[8]); g.g[1].C = aw; g.g[1].epi.bias = F(p[9]); g.g[1].epi.act = ACT_SIGMOID;
    TCX_TRY(launch_gemm(g, st));
  }
  TCX_TRY(launch_iff_gate(src, B, HW, HW, C, ah, aw, gated, st));
  GemmParams g = gemm1(gated, F(p[10]), y, B * HW * HW, Cout, inp);
  g.g[0].epi.bias = F(p[11]);
  return launch_gemm(g, st);
}

// ---- bridge ---------------------------------------------------------------------------------
int tcx_bridge_regroup_fwd(const void* const* maps, float* tokens, int B, int S0, void* stream) {
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  RegroupArgs a{};
  for (int i = 0; i < 4; i++) a.src[i] = F(maps[i]);
  for (int i = 0; i < 5; i++) a.tok_off[i] = g.off[i];
  a.dst = tokens; a.ntok = g.ntok; a.B = B;
  return launch_regroup(a, S(stream));
}

size_t tcx_scale_reduce_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  return 4 * scale_reduce_ws_floats(B, g);
}

int tcx_scale_reduce_fwd(const float* x, const void* const* p, float ln_eps, float* out, int B, int S0, void* ws,
                         void* stream) {
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  return run_scale_reduce(x, p, ln_eps, out, B, g, reinterpret_cast<float*>(ws), S(stream));
}

size_t tcx_bridge_sr_attn_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  const size_t bn = (size_t)B * g.ntok * 64;
  return 4 * (2 * rnd(bn) + rnd((size_t)B * g.nred * 64) + rnd((size_t)B * g.nred * 128) + scale_reduce_ws_floats(B, g) +
              rnd(flash_tc_workspace_bytes(B, g.nred) / 4 + 64));
}

int tcx_bridge_sr_attn_fwd(const float* xn, const void* const* p, float scale, float ln_eps, const float* residual,
                           float* y, int B, int S0, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  Carver c(ws);
  const size_t bn = (size_t)B * g.ntok * 64;
  float* q = c.take(bn);
  float* o = c.take(bn);
  float* red = c.take((size_t)B * g.nred * 64);
  float* kv = c.take((size_t)B * g.nred * 128);
  float* srws = c.take(scale_reduce_ws_floats(B, g));
  float* fws = c.take(flash_tc_workspace_bytes(B, g.nred) / 4 + 64);
  const int M = B * g.ntok;
  GemmParams gq = gemm1(xn, F(p[0]), q, M, 64, 64);
  gq.g[0].epi.bias = F(p[1]);
  TCX_TRY(launch_gemm(gq, st));
  TCX_TRY(run_scale_reduce(xn, p + 6, ln_eps, red, B, g, srws, st));
  GemmParams gk = gemm1(red, F(p[2]), kv, B * g.nred, 128, 64);
  gk.g[0].epi.bias = F(p[3]);
  TCX_TRY(launch_gemm(gk, st));
  if (flash_tc_enabled()) TCX_TRY(launch_flash_tc(q, kv, o, B, g.ntok, g.nred, scale, fws, st));
  else TCX_TRY(launch_flash_ffma(q, kv, o, B, g.ntok, g.nred, scale, st));
  GemmParams gp = gemm1(o, F(p[4]), y, M, 64, 64);
  gp.g[0].epi.bias = F(p[5]);
  gp.g[0].epi.residual = residual;
  return launch_gemm(gp, st);
}

size_t tcx_flash_attn_workspace_bytes(int B, int Nk) { return flash_tc_workspace_bytes(B, Nk) + 256; }

int tcx_flash_attn_fwd(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, void* ws,
                       void* stream) {
  TCX_REQUIRE(B >= 0 && Nq >= 0 && Nk >= 1, "flash_attn: bad sizes B=%d Nq=%d Nk=%d", B, Nq, Nk);
  if (B == 0 || Nq == 0) return 0;
  if (flash_tc_enabled()) return launch_flash_tc(q, kv, out, B, Nq, Nk, scale, ws, S(stream));
  return launch_flash_ffma(q, kv, out, B, Nq, Nk, scale, S(stream));
}

/* training forward: the fp32-io flash kernel + the row log2-sum-exp the flash backward consumes */
int tcx_flash_attn_train_fwd(const float* q, const float* kv, float* out, float* lse, int B, int Nq, int Nk, float scale, void* ws,
                             void* stream) {
  TCX_REQUIRE(q && kv && out && lse && ws, "flash_attn_train_fwd: null pointer");
  TCX_REQUIRE(flash_tc_enabled(), "flash_attn_train_fwd: needs the tcgen05 flash kernel (flag flash_tc)");
  return launch_flash_tc(q, kv, out, B, Nq, Nk, scale, ws, S(stream), lse);
}
size_t tcx_flash_attn_bwd_workspace_bytes(int B, int Nq, int Nk) { return 4 * (flash_bwd_workspace_floats(B, Nq, Nk) + 64); }
int tcx_flash_attn_bwd(const float* q, const float* kv, const float* out, const float* lse, const float* dout, float scale, float* dq,
                       float* dkv, int B, int Nq, int Nk, void* ws, void* stream) {
  TCX_REQUIRE(q && kv && out && lse && dout && dq && dkv && ws, "flash_attn_bwd: null pointer");
  if (B == 0 || Nq == 0 || Nk == 0) return 0;
  return launch_flash_bwd(q, kv, out, lse, dout, scale, dq, dkv, B, Nq, Nk, reinterpret_cast<float*>(ws), S(stream));
}

int tcx_flash_attn_f16_fwd(const void* q16, const void* kv16, void* out16, int B, int Nq, int Nk, float scale, void* ws,
                           void* stream) {
  TCX_REQUIRE(B >= 0 && Nq >= 0 && Nk >= 1, "flash_attn_f16: bad sizes B=%d Nq=%d Nk=%d", B, Nq, Nk);
  if (B == 0 || Nq == 0) return 0;
  return launch_flash_tc16(q16, kv16, out16, B, Nq, Nk, scale, ws, S(stream));
}

size_t tcx_bridge_mixffn_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  size_t n = 0;
  for (int k = 0; k < 4; k++) n += 2 * rnd((size_t)B * g.hw[k] * g.hw[k] * g.ch[k] * 4);
  return 4 * n;
}

// the four per-scale Mix-FFNs of one bridge layer on the fp16 LayerNorm output tx16 [B][ntok][64]
static int bridge_mixffn16(const __half* tx16, const float* tx1, const void* const* p, float ln_eps, float* y, int B,
                           const BridgeGeom& g, float* ws, cudaStream_t st, const LnOut* next = nullptr) {
  Carver c(ws);
  const long long sb = (long long)g.ntok * 64;
  AuxStreams* aux = aux_streams(st);
  if (aux) TCX_TRY(fork_streams(aux, st, 3));
  for (int k = 0; k < 4; k++) {       // the four scales are independent chains: scale 0 on st, 1..3 on aux streams
    const int hw = g.hw[k], C = g.ch[k], C4 = 4 * C, Mi = hw * hw;
    __half* h = H16(c.take((size_t)B * Mi * C4 / 2));
    __half* a = H16(c.take((size_t)B * Mi * C4 / 2));
    const long long off = (long long)g.off[k] * 64;
    Mix16 m{};
    TCX_REQUIRE(mix16_fill(p + 8 * k, m), "bridge_mixffn16: weights of scale %d are not prepared", k);
    m.xn = tx16 + off; m.xn_bs = sb; m.res = tx1 + off; m.res_bs = sb; m.y = y + off; m.y_bs = sb;
    // The LN epilogue leaves the GEMM only two smem stages; with fc2's long K loop (4C = 256..2048) that costs more than
    // the separate LayerNorm kernel it saves (measured: +5 us per GEMM), so it is taken only for short-K tiles.
    if (next && next->out && C4 <= 128) {   // next layer's norm1 over every 64-wide token of this slab, as fp16 tokens
      m.ln = *next; m.ln.out = next->out + off; m.ln_bs = sb;
    }
    cudaStream_t sk = (aux && k > 0) ? aux->s[k - 1] : st;
    TCX_TRY(run_mixffn16(1, &m, ln_eps, B, hw, hw, C, C4, h, a, sk));
  }
  if (aux)
    for (int k = 0; k < 3; k++) TCX_TRY(join_stream(aux, k, st));
  return 0;
}
static bool bridge_mix_prepared(const void* const* p) {
  for (int k = 0; k < 4; k++)
    if (!w16_of(p[8 * k]) || !w16_of(p[8 * k + 6])) return false;
  return true;
}

int tcx_bridge_mixffn_fwd(const float* tx, const float* tx1, const void* const* p, float ln_eps, float* y, int B,
                          int S0, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  if (bridge_mix_prepared(p)) {
    // fp16 copy of tx goes to the tail of the workspace (the fp16 h/a buffers need only half of the fp32 budget)
    size_t nfl = 0;
    for (int k = 0; k < 4; k++) nfl += 2 * rnd((size_t)B * g.hw[k] * g.hw[k] * g.ch[k] * 4 / 2);
    __half* tx16 = H16(reinterpret_cast<float*>(ws) + nfl);
    TCX_TRY(launch_f32_to_f16(tx, tx16, (long long)B * g.ntok * 64, st));
    return bridge_mixffn16(tx16, tx1, p, ln_eps, y, B, g, reinterpret_cast<float*>(ws), st);
  }
  Carver c(ws);
  const long long sb = (long long)g.ntok * 64;
  for (int k = 0; k < 4; k++) {
    const int hw = g.hw[k], C = g.ch[k], C4 = 4 * C, Mi = hw * hw;
    const void* const* pk = p + 8 * k;
    float* h = c.take((size_t)B * Mi * C4);
    float* a = c.take((size_t)B * Mi * C4);
    const long long off = (long long)g.off[k] * 64;
    {
      GemmParams gp = gemm1(tx + off, F(pk[0]), h, Mi, C4, C);
      gp.batch = B; gp.strideA = sb; gp.strideC = (long long)Mi * C4;
      gp.g[0].epi.bias = F(pk[1]);
      TCX_TRY(launch_gemm(gp, st));
    }
    MixMidGroup mg{h, F(pk[2]), F(pk[3]), F(pk[4]), F(pk[5]), a};
    TCX_TRY(launch_mixffn_mid(&mg, 1, B, hw, hw, C4, ln_eps, st));
    {
      GemmParams gp = gemm1(a, F(pk[6]), y + off, Mi, C, C4);
      gp.batch = B; gp.strideA = (long long)Mi * C4; gp.strideC = sb;
      gp.g[0].epi.bias = F(pk[7]);
      gp.g[0].epi.residual = tx1 + off; gp.g[0].epi.ldr = C; gp.g[0].epi.strideR = sb;
      TCX_TRY(launch_gemm(gp, st));
    }
  }
  return 0;
}

// ---- fused per-forward entries (one call per reference forward, fp16 intermediates when weights are prepared) ----
size_t tcx_eff_block_workspace_bytes(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  return 4 * (2 * rnd(bnc)) + tcx_eff_attn_workspace_bytes(B, N, C) + tcx_mixffn_skip_workspace_bytes(B, N, 4 * C) + 1024;
}

int tcx_eff_block_fwd(const float* x, const void* const* p, float ln_eps, float mlp_ln_eps, float* y, int B, int H, int W,
                      int C, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  const int N = H * W;
  const size_t bnc = (size_t)B * N * C;
  Carver c(ws);
  float* n = c.take(bnc);
  float* tx = c.take(bnc);
  float* aws = c.take(tcx_eff_attn_workspace_bytes(B, N, C) / 4);
  float* mws = c.take(tcx_mixffn_skip_workspace_bytes(B, N, 4 * C) / 4);
  Mix16 mprobe{};
  const bool fuse_ln2 = C == 64 && eff_attn_prepared(p + 2, N, C, 0) && mix16_fill(p + 12, mprobe);
  // n16b: fp16 norm2 output written by the reprojection GEMM's epilogue (second half of the n buffer)
  __half* n16b = H16(n) + bnc;
  if (fuse_ln2) {
    TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), H16(n), nullptr, (long long)B * N, C, ln_eps, st));
    LnOut ln; ln.w = F(p[10]); ln.b = F(p[11]); ln.eps = ln_eps; ln.out = n16b;
    TCX_TRY(run_eff_attn16(H16(n), p + 2, x, tx, B, N, C, 0, aws, st, ln));
    Carver mc(mws);
    __half* h = H16(mc.take(bnc * 2));
    __half* a = H16(mc.take(bnc * 2));
    mprobe.xn = n16b; mprobe.res = tx; mprobe.y = y;
    return run_mixffn16(1, &mprobe, mlp_ln_eps, B, H, W, C, 4 * C, h, a, st);
  }
  if (eff_attn_prepared(p + 2, N, C, 0)) {
    TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), H16(n), nullptr, (long long)B * N, C, ln_eps, st));
    TCX_TRY(run_eff_attn16(H16(n), p + 2, x, tx, B, N, C, 0, aws, st));
  } else {
    TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), nullptr, n, (long long)B * N, C, ln_eps, st));
    TCX_TRY(tcx_eff_attn_fwd(n, p + 2, x, tx, B, N, C, 0, aws, stream));
  }
  Mix16 m{};
  if (mix16_fill(p + 12, m)) {
    __half* n16 = H16(n);
    TCX_TRY(run_ln16_1(tx, F(p[10]), F(p[11]), n16, nullptr, (long long)B * N, C, ln_eps, st));
    Carver mc(mws);
    __half* h = H16(mc.take(bnc * 2));     // B*N*4C halfs
    __half* a = H16(mc.take(bnc * 2));
    m.xn = n16; m.res = tx; m.y = y;
    return run_mixffn16(1, &m, mlp_ln_eps, B, H, W, C, 4 * C, h, a, st);
  }
  TCX_TRY(run_ln16_1(tx, F(p[10]), F(p[11]), nullptr, n, (long long)B * N, C, ln_eps, st));
  return tcx_mixffn_skip_fwd(n, p + 12, mlp_ln_eps, tx, y, B, H, W, C, 4 * C, mws, stream);
}

size_t tcx_bridge_layer_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  const size_t bn = (size_t)B * g.ntok * 64;
  size_t att = tcx_eff_attn_workspace_bytes(B, g.ntok, 64);
  const size_t sr = tcx_bridge_sr_attn_workspace_bytes(B, S0);
  if (sr > att) att = sr;
  return 4 * (3 * rnd(bn)) + att + tcx_bridge_mixffn_workspace_bytes(B, S0) + 1024;
}

// n1_in: this layer's norm1 output already computed (fp16) by the previous layer's epilogues, or null.
// next: norm1 of the FOLLOWING layer to be produced by this layer's fc2 epilogues (fp16 pipeline only), or null;
// *next_done tells the caller whether that happened.
static int bridge_layer_impl(const float* x, const void* const* p, int channel_att, float scale, float ln_eps, float* y,
                             int B, int S0, void* ws, void* stream, const __half* n1_in, const LnOut* next, bool* next_done) {
  if (next_done) *next_done = false;
  cudaStream_t st = S(stream);
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  const size_t bn = (size_t)B * g.ntok * 64;
  const long long M = (long long)B * g.ntok;
  Carver c(ws);
  float* n1 = c.take(bn);
  float* tx1 = c.take(bn);
  float* tx = c.take(bn);
  size_t att = tcx_eff_attn_workspace_bytes(B, g.ntok, 64);
  const size_t sr = tcx_bridge_sr_attn_workspace_bytes(B, S0);
  if (sr > att) att = sr;
  float* aws = c.take(att / 4);
  float* mws = c.take(tcx_bridge_mixffn_workspace_bytes(B, S0) / 4);
  const bool mixp = bridge_mix_prepared(p + 18);
  LnOut ln2;       // norm2 fused into the attention's output projection (64-wide rows) when the Mix-FFNs take fp16
  if (mixp) { ln2.w = F(p[16]); ln2.b = F(p[17]); ln2.eps = ln_eps; ln2.out = H16(tx); }
  bool ln2_done = false;
  if (channel_att && eff_attn_prepared(p + 2, g.ntok, 64, 1)) {
    if (!n1_in) TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), H16(n1), nullptr, M, 64, ln_eps, st));
    TCX_TRY(run_eff_attn16(n1_in ? n1_in : H16(n1), p + 2, x, tx1, B, g.ntok, 64, 1, aws, st, ln2));
    ln2_done = mixp;
  } else if (!channel_att && flash_tc_enabled() && bridge_sr_prepared(p + 2)) {
    if (!n1_in) TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), H16(n1), nullptr, M, 64, ln_eps, st));
    TCX_TRY(run_bridge_sr_attn16(n1_in ? n1_in : H16(n1), p + 2, scale, ln_eps, x, tx1, B, g, aws, st, ln2));
    ln2_done = mixp;
  } else {
    TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), nullptr, n1, M, 64, ln_eps, st));
    if (channel_att) TCX_TRY(tcx_eff_attn_fwd(n1, p + 2, x, tx1, B, g.ntok, 64, 1, aws, stream));
    else TCX_TRY(tcx_bridge_sr_attn_fwd(n1, p + 2, scale, ln_eps, x, tx1, B, S0, aws, stream));
  }
  if (mixp) {
    __half* tx16 = H16(tx);
    if (!ln2_done) TCX_TRY(run_ln16_1(tx1, F(p[16]), F(p[17]), tx16, nullptr, M, 64, ln_eps, st));
    if (next_done) *next_done = false;     // see bridge_mixffn16: the fc2 epilogue fusion is not taken at these K
    return bridge_mixffn16(tx16, tx1, p + 18, ln_eps, y, B, g, mws, st, next);
  }
  TCX_TRY(run_ln16_1(tx1, F(p[16]), F(p[17]), nullptr, tx, M, 64, ln_eps, st));
  return tcx_bridge_mixffn_fwd(tx, tx1, p + 18, ln_eps, y, B, S0, mws, stream);
}

int tcx_bridge_layer_fwd(const float* x, const void* const* p, int channel_att, float scale, float ln_eps, float* y,
                         int B, int S0, void* ws, void* stream) {
  return bridge_layer_impl(x, p, channel_att, scale, ln_eps, y, B, S0, ws, stream, nullptr, nullptr, nullptr);
}

size_t tcx_bridge_block_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  const size_t bn = (size_t)B * g.ntok * 64;
  return 4 * (rnd(bn) + 2 * rnd(bn / 2 + 64)) + tcx_bridge_layer_workspace_bytes(B, S0) + 1024;
}

int tcx_bridge_block_fwd(const float* x, const void* const* p, const int* channel_att, int L, float scale, float ln_eps,
                         float* y, int B, int S0, void* ws, void* stream) {
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  TCX_REQUIRE(L >= 1, "bridge_block: L must be >= 1");
  const size_t bn = (size_t)B * g.ntok * 64;
  Carver c(ws);
  float* xb = c.take(bn);                      // ping-pong partner of y for the layer outputs
  __half* n1buf[2] = {H16(c.take(bn / 2 + 64)), H16(c.take(bn / 2 + 64))};
  float* lws = c.take(tcx_bridge_layer_workspace_bytes(B, S0) / 4);
  const float* cur = x;
  const __half* n1_in = nullptr;
  for (int l = 0; l < L; l++) {
    // layer outputs alternate so that the last one lands in y
    float* dst = ((L - 1 - l) & 1) ? xb : y;
    const void* const* pl = p + (size_t)l * TCX_BRIDGE_NP;
    LnOut next;
    if (l + 1 < L) {
      const void* const* pn = p + (size_t)(l + 1) * TCX_BRIDGE_NP;
      next.w = F(pn[0]); next.b = F(pn[1]); next.eps = ln_eps; next.out = n1buf[l & 1];
    }
    bool done = false;
    TCX_TRY(bridge_layer_impl(cur, pl, channel_att[l], scale, ln_eps, dst, B, S0, lws, stream, n1_in,
                              l + 1 < L ? &next : nullptr, &done));
    n1_in = done ? n1buf[l & 1] : nullptr;
    cur = dst;
  }
  return 0;
}

// ---- decoder ----------------------------------------------------------------------------------
int tcx_concat_linear_fwd(const float* x1, const float* x2, const float* w, const float* b, float* y, int M, int C1,
                          int C2, int N, int batch, long long x2_batch_stride, void* stream) {
  cudaStream_t st = S(stream);
  TCX_REQUIRE(batch >= 1, "concat_linear: batch must be >= 1");
  // y = x2 * W[:, C1:]^T + b, then y += x1 * W[:, :C1]^T   (the concatenation is never materialised).
  // batch > 1: M rows per image, x1 / y dense, x2 a per-image slab with pitch x2_batch_stride (a bridge output map
  // viewed in place inside the token buffer, MSTr.py:2432-2435 -> :2847-2850).
  GemmParams g = gemm1(x2, w + C1, y, M, N, C2);
  g.ldw = C1 + C2;
  g.g[0].epi.bias = b;
  if (batch > 1) { g.batch = batch; g.strideA = x2_batch_stride; g.strideC = (long long)M * N; }
  TCX_TRY(launch_gemm(g, st));
  GemmParams h = gemm1(x1, w, y, M, N, C1);
  h.ldw = C1 + C2;
  h.g[0].epi.residual = y;
  if (batch > 1) { h.batch = batch; h.strideA = (long long)M * C1; h.strideC = (long long)M * N; h.g[0].epi.strideR = (long long)M * N; }
  return launch_gemm(h, st);
}

size_t tcx_patch_expand_workspace_bytes(int B, int H, int W, int C, int scale) {
  const size_t nout = scale == 2 ? 2 * (size_t)C : 16 * (size_t)C;
  return 4 * rnd((size_t)B * H * W * nout);
}

int tcx_patch_expand_fwd(const float* x, const float* w, const float* lnw, const float* lnb, float eps, float* y,
                         int B, int H, int W, int C, int scale, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  TCX_REQUIRE(scale == 2 || scale == 4, "patch_expand: scale must be 2 or 4");
  const int nout = scale == 2 ? 2 * C : 16 * C;
  float* e = reinterpret_cast<float*>(ws);
  GemmParams g = gemm1(x, w, e, B * H * W, nout, C);
  TCX_TRY(launch_gemm(g, st));
  return launch_shuffle_ln(e, B, H, W, scale, nout / (scale * scale), lnw, lnb, eps, y, st);
}

size_t tcx_final_expand_head_workspace_bytes(int B, int H, int W) { return 4 * rnd((size_t)B * H * W * 1024); }

int tcx_final_expand_head_fwd(const float* x, const float* w, const float* lnw, const float* lnb, float eps,
                              const float* cls_w, const float* cls_b, int ncls, float* logits_nchw, int B, int H,
                              int W, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  if (head_tc_eligible(x, w, ncls))   // one kernel: the 1024-wide expand output never reaches memory
    return launch_head_tc(x, w, B, H, W, lnw, lnb, eps, cls_w, cls_b, ncls, logits_nchw, reinterpret_cast<float*>(ws), st);
  float* e = reinterpret_cast<float*>(ws);
  GemmParams g = gemm1(x, w, e, B * H * W, 1024, 64);
  TCX_TRY(launch_gemm(g, st));
  return launch_final_head(e, B, H, W, lnw, lnb, eps, cls_w, cls_b, ncls, logits_nchw, st);
}

// ---- networks/Transception.py variant: fp16 pipeline only ----------------------------
static inline int conv_out(int H, int k, int stride, int pad, int dil) { return (H + 2 * pad - dil * (k - 1) - 1) / stride + 1; }

// FuseEfficientAttention (Transception.py:49-87, head_count = 1) on fp16 LayerNorm output xn16 [B*N][C]:
// y = residual + reprojection(att).  p = {keys w,b, queries w,b, values w,b, reprojection w,b}
static size_t fuse_ea_workspace_floats(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C, bcp = (size_t)B * C * fuse_np(N);
  return rnd(3 * bnc / 2 + 64) + 2 * rnd(bcp / 2 + 64) + rnd((size_t)B * C * C / 2 + 64) + 2 * rnd(bnc / 2 + 64);
}
static int run_fuse_ea16(const __half* xn16, const void* const* p, const float* residual, float* y, int B, int N, int C,
                         float* ws, cudaStream_t st) {
  TCX_REQUIRE(C % 64 == 0 && C <= 512, "fuse_eff_attn: C must be a multiple of 64, at most 512 (got %d)", C);
  for (int i = 0; i < 4; i++) TCX_REQUIRE(w16_of(p[2 * i]) != nullptr, "fuse_eff_attn: weight %d is not prepared (fp16 pipeline only)", i);
  Carver c(ws);
  const size_t bnc = (size_t)B * N * C;
  const int Np = fuse_np(N);
  __half* kqv = H16(c.take(3 * bnc / 2 + 64));
  __half* Pk = H16(c.take((size_t)B * C * Np / 2 + 64));
  __half* Vp = H16(c.take((size_t)B * C * Np / 2 + 64));
  __half* ctxT = H16(c.take((size_t)B * C * C / 2 + 64));
  __half* QsT = H16(c.take(bnc / 2 + 64));
  __half* att = H16(c.take(bnc / 2 + 64));
  {
    GemmParams g = gemm1(F(xn16), nullptr, nullptr, B * N, C, C);
    g.groups = 3; g.ab16 = 1; g.out16 = 1;
    for (int i = 0; i < 3; i++) {
      g.g[i].A = F(xn16); g.g[i].W = F(w16_of(p[2 * i])); g.g[i].epi.bias = F(p[2 * i + 1]);
      g.g[i].C = reinterpret_cast<float*>(kqv + i * bnc);
    }
    TCX_TRY(launch_gemm(g, st));
  }
  {  // the context chain and the query softmax are independent
    AuxStreams* aux = aux_streams(st);
    if (aux) TCX_TRY(fork_streams(aux, st, 1));
    TCX_TRY(launch_fea_qsoftmaxT(kqv + bnc, QsT, B, N, C, aux ? aux->s[0] : st));
    TCX_TRY(launch_fea_kpack(kqv, kqv + 2 * bnc, Pk, Vp, B, N, C, Np, st));
    // ctxT[b][cv][ck] = sum_n Vp[b][cv][n] * Pk[b][ck][n]
    GemmParams g = gemm1(F(Vp), F(Pk), reinterpret_cast<float*>(ctxT), C, C, Np);
    g.batch = B; g.strideA = (long long)C * Np; g.strideW = (long long)C * Np; g.strideC = (long long)C * C;
    g.ab16 = 1; g.out16 = 1;
    TCX_TRY(launch_gemm(g, st));
    if (aux) TCX_TRY(join_stream(aux, 0, st));
  }
  {  // att[b] = QsT[b] (N x C) * ctx[b] (C x C): W = ctxT[b]
    GemmParams a = gemm1(F(QsT), F(ctxT), reinterpret_cast<float*>(att), N, C, C);
    a.batch = B; a.strideA = (long long)N * C; a.strideW = (long long)C * C; a.strideC = (long long)N * C;
    a.ab16 = 1; a.out16 = 1;
    TCX_TRY(launch_gemm(a, st));
  }
  GemmParams r = gemm1(F(att), F(w16_of(p[6])), y, B * N, C, C);
  r.ab16 = 1;
  r.g[0].epi.bias = F(p[7]);
  r.g[0].epi.residual = residual;
  return launch_gemm(r, st);
}

size_t tcx_fuse_eff_attn_workspace_bytes(int B, int N, int C) {
  return 4 * (fuse_ea_workspace_floats(B, N, C) + rnd((size_t)B * N * C / 2 + 64)) + 1024;
}
int tcx_fuse_eff_attn_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int N, int C, void* ws,
                          void* stream) {
  cudaStream_t st = S(stream);
  Carver c(ws);
  __half* xn16 = H16(c.take((size_t)B * N * C / 2 + 64));
  float* rest = c.take(fuse_ea_workspace_floats(B, N, C));
  TCX_TRY(launch_f32_to_f16(xn, xn16, (long long)B * N * C, st));
  return run_fuse_ea16(xn16, p, residual, y, B, N, C, rest, st);
}

// EfficientTransformerBlockFuse (Transception.py:213-250, two-branch case) on the concatenated token buffer
// x [B][n1+n2][C] (n1 = H1*W1 tokens of the dilated 3x3 branch, n2 = H2*W2 of the 1x1 branch).
// p = {norm1 w,b, keys w,b, queries w,b, values w,b, reprojection w,b, norm2 w,b, mlp1[8], mlp2[8]} (Mix-FFN blocks as in
// tcx_mixffn_skip_fwd)
size_t tcx_fuse_block_workspace_bytes(int B, int N, int C) {
  const size_t bnc = (size_t)B * N * C;
  return 4 * (rnd(bnc / 2 + 64) + rnd(bnc) + fuse_ea_workspace_floats(B, N, C) + 2 * 2 * rnd(bnc * 2 + 64)) + 1024;
}
int tcx_fuse_block_fwd(const float* x, const void* const* p, float ln_eps, float mlp_ln_eps, float* y, int B, int H1, int W1, int H2,
                       int W2, int C, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  const int n1 = H1 * W1, n2 = H2 * W2, N = n1 + n2;
  const size_t bnc = (size_t)B * N * C;
  Carver c(ws);
  __half* n16 = H16(c.take(bnc / 2 + 64));
  float* tx = c.take(bnc);
  float* aws = c.take(fuse_ea_workspace_floats(B, N, C));
  TCX_TRY(run_ln16_1(x, F(p[0]), F(p[1]), n16, nullptr, (long long)B * N, C, ln_eps, st));
  TCX_TRY(run_fuse_ea16(n16, p + 2, x, tx, B, N, C, aws, st));
  TCX_TRY(run_ln16_1(tx, F(p[10]), F(p[11]), n16, nullptr, (long long)B * N, C, ln_eps, st));   // shared norm2, :230-231
  AuxStreams* aux = aux_streams(st);
  if (aux) TCX_TRY(fork_streams(aux, st, 1));
  const long long sb = (long long)N * C;
  for (int k = 0; k < 2; k++) {
    const int h = k ? H2 : H1, w = k ? W2 : W1;
    const long long off = k ? (long long)n1 * C : 0;
    __half* hb = H16(c.take(bnc * 2 + 64));     // sized for the whole sequence: B*n_k*4C halfs fit
    __half* ab = H16(c.take(bnc * 2 + 64));
    Mix16 m{};
    TCX_REQUIRE(mix16_fill(p + 12 + 8 * k, m), "fuse_block: Mix-FFN %d weights are not prepared (fp16 pipeline only)", k + 1);
    m.xn = n16 + off; m.xn_bs = sb; m.res = tx + off; m.res_bs = sb; m.y = y + off; m.y_bs = sb;
    TCX_TRY(run_mixffn16(1, &m, mlp_ln_eps, B, h, w, C, 4 * C, hb, ab, (aux && k) ? aux->s[0] : st));
  }
  if (aux) TCX_TRY(join_stream(aux, 0, st));
  return 0;
}

// The two OverlapPatchEmbeddings_fuse branches of one stage (EffSegformer.py:117-131, Transception.py:383-387, :412-419)
// on the NHWC fp32 map x [B][H][W][Cin]: k1 x k1 and k2 x k2 convs (stride / dilation shared, no padding when
// dil_conv = 1) + their LayerNorms, written into the concatenated token buffer tokens [B][n1+n2][C].
// p = {proj1 w,b, norm1 w,b, proj2 w,b, norm2 w,b}; proj weights prepared with tcx_prepare_conv_weight_f16.
size_t tcx_dual_patch_embed_workspace_bytes(int B, int H, int W, int Cin, int C, int k1, int k2, int stride, int pad1, int pad2, int dil) {
  const int H1 = conv_out(H, k1, stride, pad1, dil), W1 = conv_out(W, k1, stride, pad1, dil);
  const int H2 = conv_out(H, k2, stride, pad2, dil), W2 = conv_out(W, k2, stride, pad2, dil);
  const size_t r1 = (size_t)B * H1 * W1, r2 = (size_t)B * H2 * W2;
  return 4 * (rnd(r1 * k1 * k1 * Cin / 2 + 64) + rnd(r2 * k2 * k2 * Cin / 2 + 64) + rnd(r1 * C) + rnd(r2 * C)) + 1024;
}
int tcx_dual_patch_embed_fwd(const float* x, const void* const* p, float ln_eps, float* tokens, int B, int H, int W, int Cin, int C,
                             int k1, int k2, int stride, int pad1, int pad2, int dil, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  const int ks[2] = {k1, k2}, pads[2] = {pad1, pad2};
  int Ho[2], Wo[2];
  for (int i = 0; i < 2; i++) { Ho[i] = conv_out(H, ks[i], stride, pads[i], dil); Wo[i] = conv_out(W, ks[i], stride, pads[i], dil); }
  TCX_REQUIRE(Ho[0] > 0 && Wo[0] > 0 && Ho[1] > 0 && Wo[1] > 0, "dual_patch_embed: empty output map");
  const int n1 = Ho[0] * Wo[0], n2 = Ho[1] * Wo[1];
  Carver c(ws);
  AuxStreams* aux = aux_streams(st);
  if (aux) TCX_TRY(fork_streams(aux, st, 1));
  for (int i = 0; i < 2; i++) {
    cudaStream_t sk = (aux && i) ? aux->s[0] : st;
    const int n = i ? n2 : n1, K = ks[i] * ks[i] * Cin;
    const void* w16 = w16_of(p[4 * i]);
    TCX_REQUIRE(w16 != nullptr, "dual_patch_embed: conv weight %d is not prepared (fp16 pipeline only)", i + 1);
    __half* A = H16(c.take((size_t)B * n * K / 2 + 64));
    float* o = c.take((size_t)B * n * C);
    TCX_TRY(launch_im2row16(x, A, B, H, W, Cin, ks[i], stride, pads[i], dil, Ho[i], Wo[i], sk));
    GemmParams g = gemm1(F(A), F(w16), o, B * n, C, K);
    g.ab16 = 1;
    g.g[0].epi.bias = F(p[4 * i + 1]);
    TCX_TRY(launch_gemm(g, sk));
    TCX_TRY(launch_ln_scatter(o, F(p[4 * i + 2]), F(p[4 * i + 3]), tokens + (i ? (long long)n1 * C : 0), B, n, C,
                              (long long)(n1 + n2) * C, ln_eps, sk));
  }
  if (aux) TCX_TRY(join_stream(aux, 0, st));
  return 0;
}

// Stage tail of MiT_3inception (Transception.py:462-476, concat='original'): stage LayerNorm over all tokens, branch-1 map
// nearest-upsampled to the branch-2 size, channel concat, 1x1 conv (2C -> C).  out: [B][H2*W2][C] fp32 tokens (= NHWC map).
// p = {norm w,b, conv1_1 w,b}
size_t tcx_fuse_merge_workspace_bytes(int B, int N, int n2, int C) {
  return 4 * (rnd((size_t)B * N * C / 2 + 64) + rnd((size_t)B * n2 * C + 64)) + 1024;
}
int tcx_fuse_merge_fwd(const float* tokens, const void* const* p, float ln_eps, float* out, int B, int H1, int W1, int H2, int W2, int C,
                       void* ws, void* stream) {
  cudaStream_t st = S(stream);
  const int n1 = H1 * W1, n2 = H2 * W2, N = n1 + n2;
  const void* w16 = w16_of(p[2]);
  TCX_REQUIRE(w16 != nullptr, "fuse_merge: 1x1 conv weight is not prepared (fp16 pipeline only)");
  Carver c(ws);
  __half* t16 = H16(c.take((size_t)B * N * C / 2 + 64));
  __half* A = H16(c.take((size_t)B * n2 * C + 64));
  TCX_TRY(run_ln16_1(tokens, F(p[0]), F(p[1]), t16, nullptr, (long long)B * N, C, ln_eps, st));
  TCX_TRY(launch_upcat16(t16, A, B, H1, W1, H2, W2, C, st));
  GemmParams g = gemm1(F(A), F(w16), out, B * n2, C, 2 * C);
  g.ab16 = 1;
  g.g[0].epi.bias = F(p[3]);
  return launch_gemm(g, st);
}

// Stage tail with the selective-kernel fusion (Transception.py:477-481, SK_Block :328-358; concat != 'original'):
// stage LayerNorm, branch-1 map nearest-upsampled, S = mean(U), Z = fc(S), softmax over the two paths of fcs_i(Z),
// V = sum a_i * map_i, 1x1 conv -> ReLU -> BatchNorm(eval).  out [B][H2*W2][C].
// p = {norm_w,norm_b, fc_w,fc_b, fcs0_w,fcs0_b, fcs1_w,fcs1_b, conv_w,conv_b, bn_w,bn_b,bn_rm,bn_rv}
size_t tcx_fuse_merge_sk_workspace_bytes(int B, int N, int n2, int C) {
  return 4 * (rnd((size_t)B * N * C / 2 + 64) + rnd((size_t)B * n2 * C / 2 + 64) + 3 * rnd((size_t)B * C + 64)) + 1024;
}
int tcx_fuse_merge_sk_fwd(const float* tokens, const void* const* p, float ln_eps, float bn_eps, float* out, int B, int H1, int W1,
                          int H2, int W2, int C, int d, void* ws, void* stream) {
  cudaStream_t st = S(stream);
  const int n1 = H1 * W1, n2 = H2 * W2, N = n1 + n2;
  const void* w16 = w16_of(p[8]);
  TCX_REQUIRE(w16 != nullptr, "fuse_merge_sk: 1x1 conv weight is not prepared (fp16 pipeline only)");
  Carver c(ws);
  __half* t16 = H16(c.take((size_t)B * N * C / 2 + 64));
  __half* A = H16(c.take((size_t)B * n2 * C / 2 + 64));
  float* Sm = c.take((size_t)B * C + 64);
  float* att = c.take(2 * (size_t)B * C + 64);
  TCX_TRY(run_ln16_1(tokens, F(p[0]), F(p[1]), t16, nullptr, (long long)B * N, C, ln_eps, st));
  TCX_TRY(launch_sk_pool(t16, Sm, B, H1, W1, H2, W2, C, st));
  TCX_TRY(launch_sk_weights(Sm, F(p[2]), F(p[3]), F(p[4]), F(p[5]), F(p[6]), F(p[7]), att, B, C, d, st));
  TCX_TRY(launch_sk_mix(t16, att, A, B, H1, W1, H2, W2, C, st));
  GemmParams g = gemm1(F(A), F(w16), out, B * n2, C, C);
  g.ab16 = 1;
  g.g[0].epi.bias = F(p[9]);
  TCX_TRY(launch_gemm(g, st));
  BnParams bn{F(p[10]), F(p[11]), F(p[12]), F(p[13]), bn_eps};
  return launch_relu_bn(out, (long long)B * n2, C, bn, st);
}

// ---- fused training loss: 0.4 CE + 0.6 Dice of trainer.py:141-143 / utils.py:11-47 ----
static int seg_loss_args(SegLossArgs& a, const float* logits, const void* labels, int label_kind, int B, int K, long long HW,
                         int softmax, float w_ce, float w_dice, const float* class_w) {
  TCX_REQUIRE(logits && labels, "seg_loss: null input");
  a.logits = logits; a.labels = labels; a.kind = label_kind; a.B = B; a.K = K; a.HW = HW; a.softmax = softmax;
  a.w_ce = w_ce; a.w_dice = w_dice;
  for (int c = 0; c < SEG_LOSS_KMAX; c++) a.cw[c] = (class_w && c < K) ? class_w[c] : 1.f;     // class_w is a HOST array
  return 0;
}
size_t tcx_seg_loss_workspace_bytes(int B, int K, long long HW) { return 4 * seg_loss_workspace_floats(B, K, HW); }
int tcx_seg_loss_fwd(const float* logits, const void* labels, int label_kind, int B, int K, long long HW, int softmax, float w_ce,
                     float w_dice, const float* class_w, float* out, void* ws, void* stream) {
  SegLossArgs a;
  TCX_TRY(seg_loss_args(a, logits, labels, label_kind, B, K, HW, softmax, w_ce, w_dice, class_w));
  return launch_seg_loss_fwd(a, out, reinterpret_cast<float*>(ws), S(stream));
}
int tcx_seg_loss_bwd(const float* logits, const void* labels, int label_kind, int B, int K, long long HW, int softmax, float w_ce,
                     float w_dice, const float* class_w, const float* grad_out, float* dlogits, const void* ws, void* stream) {
  SegLossArgs a;
  TCX_TRY(seg_loss_args(a, logits, labels, label_kind, B, K, HW, softmax, w_ce, w_dice, class_w));
  return launch_seg_loss_bwd(a, reinterpret_cast<const float*>(ws), grad_out, dlogits, S(stream));
}

// per-pixel arg max over the class planes (utils.py:86), logits [B][K][HW] -> uint8 labels [B][HW]
int tcx_argmax_classes_fwd(const float* logits, unsigned char* labels, int B, int K, long long HW, void* stream) {
  TCX_REQUIRE(logits && labels, "argmax_classes: null pointer");
  return launch_argmax_classes(logits, labels, B, K, HW, S(stream));
}

}  // extern "C"

// ---- training row: backward entries ----------------------------------------------------------
namespace {

// in-place TF32 Linear backward (wgrad_tc.cu + the MN-major-W mode of gemm_tc): needs 16-byte fp32 row pitches
bool linear_bwd_mn_ok(long long M, int N, int K) {
  return g_flag_wgrad_tc && g_flag_gemm_tc && M >= 1 && N % 4 == 0 && K % 4 == 0 && N >= 16 && K >= 16 &&
         wgrad_tc_eligible(M, N, K, N, K, 4);
}

size_t linear_bwd_ws_floats(long long M, int N, int K) {
  int S, Ms;
  bwd_wgrad_splits(M, N, K, &S, &Ms);
  const size_t npad = (size_t)(N + 31) / 32 * 32;
  const size_t old = rnd(npad * K) + rnd((size_t)S * Ms * N) + rnd((size_t)S * Ms * K) + rnd((size_t)S * N * K) +
                     rnd((size_t)bwd_red_blocks(M) * N) + 64;
  const size_t mn = rnd((size_t)M * K) + rnd((size_t)N * K) + rnd(wgrad_tc_scratch_floats(M > 0 ? M : 1, N, K, 1, 4)) + 64;
  return std::max(old, mn);
}

// nn.Linear backward: y = x w^T + b with x [M][K] (fp32, or fp16 when x16), w [N][K], dy [M][N]:
//   dx [M][K] = dy w,  dw [N][K] = dy^T x,  db [N] = column sums of dy.   Any of dx / dw / db may be null.
// Default path (flag "wgrad_tc"): every operand is read IN PLACE as fp32 with TF32 MMAs — dx = dy w takes the [N][K] weight as
// an MN-major B operand, dw and db come from ONE launch of the MN-major weight-gradient kernel (dy and x as stored, tokens
// major).  2 launches instead of the 8-9 of the packT path below; same TF32 precision; no scale or range concerns (measured
// and dropped this round: fp16 gradient operands with a static scale saturate on large caller gradients, bf16 x fp16 mixed
// operands are rejected by the hardware, and bf16 x bf16 costs a conversion pass and 2.3e-3 relative error).
int run_linear_bwd(const void* x, int x16, const float* w, const float* dy, float* dx, float* dw, float* db, long long M, int N,
                   int K, float* ws, cudaStream_t st) {
  TCX_REQUIRE(M < (1ll << 31) && N % 4 == 0 && K % 4 == 0, "linear_bwd: N, K must be multiples of 4 (M=%lld N=%d K=%d)", M, N, K);
  Carver c(ws);
  if (M > 0 && linear_bwd_mn_ok(M, N, K)) {
    const bool need_w = dw != nullptr || db != nullptr;
    AuxStreams* aux = dx != nullptr && need_w ? aux_streams(st) : nullptr;
    if (aux) TCX_TRY(fork_streams(aux, st, 1));
    cudaStream_t sw = aux ? aux->s[0] : st;
    if (dx) {
      GemmParams g = gemm1(dy, w, dx, (int)M, K, N);       // contraction over the N output features; w [N][K] read as stored
      g.w_mn = 1; g.ldw = K;
      TCX_TRY(launch_gemm(g, st));
    }
    if (need_w) {
      const float* xx = F(x);
      if (x16) {           // a saved fp16 activation with no fp32 twin
        float* t = c.take((size_t)M * K);
        TCX_TRY(launch_f16_to_f32(reinterpret_cast<const __half*>(x), t, M * K, sw));
        xx = t;
      }
      float* dw_out = dw ? dw : c.take((size_t)N * K);      // bias gradient alone: the product is computed and dropped
      WgradArgs a{};
      a.A = dy; a.B = xx; a.fmt = 2; a.Mtok = M; a.NL = N; a.KL = K; a.lda = N; a.ldb = K; a.batch = 1;
      a.alpha = 1.0f; a.out = dw_out; a.ldo = K; a.db = db;
      a.scratch = c.take(wgrad_tc_scratch_floats(M, N, K, 1, 4));
      TCX_TRY(launch_wgrad_tc(a, sw));
    }
    if (aux) TCX_TRY(join_stream(aux, 0, st));
    return 0;
  }
  // dx, dw and db are independent: the weight-gradient chain and the bias sums run on auxiliary streams beside the input
  // gradient (parallel branches of a captured graph); all scratch regions are disjoint
  AuxStreams* aux = (M > 0 && ((dx != nullptr) + (dw != nullptr) + (db != nullptr)) > 1) ? aux_streams(st) : nullptr;
  if (aux) TCX_TRY(fork_streams(aux, st, 2));
  cudaStream_t sw = aux && dx ? aux->s[0] : st;                 // weight gradient
  cudaStream_t sb = aux && (dx || dw) ? aux->s[1] : st;         // bias gradient
  if (dx && M > 0) {
    const int npad = (N + 31) / 32 * 32;
    float* wT = c.take((size_t)npad * K);
    TCX_TRY(launch_bwd_packT_f32(w, N, K, K, 1, npad, wT, st));          // wT [K][npad]
    GemmParams g = gemm1(dy, wT, dx, (int)M, K, N);
    g.ldw = npad;
    TCX_TRY(launch_gemm(g, st));
  }
  if (dw) {
    if (M == 0) {
      TCX_REQUIRE(cudaMemsetAsync(dw, 0, sizeof(float) * N * K, st) == cudaSuccess, "linear_bwd: memset failed");
    } else {
      int S, Ms;
      bwd_wgrad_splits(M, N, K, &S, &Ms);
      float* dyT = c.take((size_t)S * Ms * N);
      float* xT = c.take((size_t)S * Ms * K);
      float* part = c.take((size_t)S * N * K);
      TCX_TRY(launch_bwd_packT_f32(dy, M, N, N, S, Ms, dyT, sw));
      if (x16) TCX_TRY(launch_bwd_packT_f16(reinterpret_cast<const __half*>(x), M, K, K, S, Ms, xT, sw));
      else TCX_TRY(launch_bwd_packT_f32(F(x), M, K, K, S, Ms, xT, sw));
      GemmParams g = gemm1(dyT, xT, S == 1 ? dw : part, N, K, Ms);
      g.batch = S; g.strideA = (long long)N * Ms; g.strideW = (long long)K * Ms; g.strideC = (long long)N * K;
      TCX_TRY(launch_gemm(g, sw));
      if (S > 1) TCX_TRY(launch_bwd_fold(part, S, (long long)N * K, dw, sw));
    }
  }
  if (db) {
    float* part = c.take((size_t)bwd_red_blocks(M) * N);
    TCX_TRY(launch_bwd_colsum(dy, M, N, N, part, db, sb));
  }
  if (aux) {      // both forked streams rejoin (an idle one too: stream capture requires every fork to be joined)
    TCX_TRY(join_stream(aux, 0, st));
    TCX_TRY(join_stream(aux, 1, st));
  }
  return 0;
}

struct MixSaved {
  __half* h16; __half* a16; __half* xn16; float* u;
  size_t floats;
};
MixSaved mix_saved(void* base, long long M, int C, int C4) {
  Carver c(base);
  MixSaved s;
  s.h16 = H16(c.take((size_t)M * C4 / 2 + 64));
  s.a16 = H16(c.take((size_t)M * C4 / 2 + 64));
  s.xn16 = H16(c.take((size_t)M * C / 2 + 64));
  s.u = c.take((size_t)M * C4);
  s.floats = c.off + 64;
  return s;
}

}  // namespace

// ---- efficient-attention core on given fp32 token-major K, Q, V [B][N][C] (the reinterpreted tensors of M_EfficientChannelAtten,
// MSTr.py:2312-2353): ctx = softmax_tokens(K)^T V, att = softmax_channels(Q) ctx ----
namespace {
struct EaCorePlan { int S, Ms; size_t bnc, bcc, pack, ch; };
EaCorePlan ea_core_plan(int B, int N, int C) {
  EaCorePlan p;
  bwd_wgrad_splits(N, C, C, &p.S, &p.Ms);
  p.bnc = (size_t)B * N * C; p.bcc = (size_t)B * C * C;
  p.pack = (size_t)B * p.S * C * p.Ms; p.ch = (size_t)B * ea_bwd_chunks(N) * C;
  return p;
}
struct EaCoreBufs { float *P, *Qs, *ctx, *ctxT, *packA, *packB, *part, *pm, *ps; };
size_t ea_core_common_floats(const EaCorePlan& p) {
  return 2 * rnd(p.bnc) + 2 * rnd(p.bcc) + 2 * rnd(p.pack) + rnd(p.bcc * p.S) + 2 * rnd(p.ch);
}
EaCoreBufs ea_core_carve(Carver& c, const EaCorePlan& p) {
  EaCoreBufs b;
  b.P = c.take(p.bnc); b.Qs = c.take(p.bnc); b.ctx = c.take(p.bcc); b.ctxT = c.take(p.bcc);
  b.packA = c.take(p.pack); b.packB = c.take(p.pack); b.part = c.take(p.bcc * p.S); b.pm = c.take(p.ch); b.ps = c.take(p.ch);
  return b;
}
// out[b] = a[b]^T b_[b] over the N tokens of image b (C x C), outT transposed
int ea_ctx_gemm(const float* a, const float* b_, int B, int N, int C, const EaCorePlan& p, const EaCoreBufs& w, float* out, float* outT,
                cudaStream_t st) {
  TCX_TRY(launch_bwd_packT_batched_f32(a, B, N, C, C, p.S, p.Ms, p.Ms, w.packA, st));
  TCX_TRY(launch_bwd_packT_batched_f32(b_, B, N, C, C, p.S, p.Ms, p.Ms, w.packB, st));
  GemmParams g = gemm1(w.packA, w.packB, w.part, C, C, p.Ms);
  g.batch = B * p.S; g.strideA = (long long)C * p.Ms; g.strideW = (long long)C * p.Ms; g.strideC = (long long)C * C;
  TCX_TRY(launch_gemm(g, st));
  return launch_bwd_fold_mask(w.part, B, p.S, C, C, 1.0f, out, outT, st);
}
// out[b] (N x C) = a[b] (N x C) w[b]^T
int ea_tok_gemm(const float* a, const float* w, float* out, int B, int N, int C, cudaStream_t st) {
  GemmParams g = gemm1(a, w, out, N, C, C);
  g.batch = B; g.strideA = (long long)N * C; g.strideW = (long long)C * C; g.strideC = (long long)N * C;
  return launch_gemm(g, st);
}
int ea_core_recompute(const float* k, const float* q, const float* v, int B, int N, int C, const EaCorePlan& p, const EaCoreBufs& w,
                      cudaStream_t st) {
  TCX_TRY(launch_bwd_ksoftmax32(k, C, B, N, C, w.pm, w.ps, w.P, st));
  TCX_TRY(launch_bwd_rowsoftmax_fwd(q, C, (long long)B * N, C, 1.0f, w.Qs, C, st));
  return ea_ctx_gemm(w.P, v, B, N, C, p, w, w.ctx, w.ctxT, st);
}
}  // namespace

extern "C" {
size_t tcx_ea_core_workspace_bytes(int B, int N, int C) {
  const EaCorePlan p = ea_core_plan(B, N, C);
  return 4 * (ea_core_common_floats(p) + 2 * rnd(p.bnc) + 2 * rnd(p.bcc) + rnd(p.ch) + 1024);
}
int tcx_ea_core_fwd(const float* k, const float* q, const float* v, float* out, int B, int N, int C, void* ws, void* stream) {
  TCX_REQUIRE(k && q && v && out && ws, "ea_core_fwd: null pointer");
  TCX_REQUIRE(C % 4 == 0 && C >= 16, "ea_core: C must be a multiple of 4, >= 16 (got %d)", C);
  const EaCorePlan p = ea_core_plan(B, N, C);
  Carver c(ws);
  const EaCoreBufs w = ea_core_carve(c, p);
  TCX_TRY(ea_core_recompute(k, q, v, B, N, C, p, w, S(stream)));
  return ea_tok_gemm(w.Qs, w.ctxT, out, B, N, C, S(stream));              // att = Qs ctx
}
int tcx_ea_core_bwd(const float* k, const float* q, const float* v, const float* dout, float* dk, float* dq, float* dv, int B, int N,
                    int C, void* ws, void* stream) {
  TCX_REQUIRE(k && q && v && dout && dk && dq && dv && ws, "ea_core_bwd: null pointer");
  TCX_REQUIRE(C % 4 == 0 && C >= 16, "ea_core: C must be a multiple of 4, >= 16 (got %d)", C);
  cudaStream_t st = S(stream);
  const EaCorePlan p = ea_core_plan(B, N, C);
  Carver c(ws);
  const EaCoreBufs w = ea_core_carve(c, p);
  float* dQs = c.take(p.bnc);
  float* dP = c.take(p.bnc);
  float* dctx = c.take(p.bcc);
  float* dctxT = c.take(p.bcc);
  float* sp = c.take(p.ch);
  TCX_TRY(ea_core_recompute(k, q, v, B, N, C, p, w, st));
  TCX_TRY(ea_tok_gemm(dout, w.ctx, dQs, B, N, C, st));                    // dQs = dout ctx^T
  TCX_TRY(launch_bwd_rowsoftmax_bwd(w.Qs, C, dQs, C, (long long)B * N, C, 1.0f, dq, C, st));
  TCX_TRY(ea_ctx_gemm(w.Qs, dout, B, N, C, p, w, dctx, dctxT, st));       // dctx = Qs^T dout
  TCX_TRY(ea_tok_gemm(w.P, dctxT, dv, B, N, C, st));                      // dv = P dctx
  TCX_TRY(ea_tok_gemm(v, dctx, dP, B, N, C, st));                         // dP = v dctx^T
  return launch_bwd_colsoftmax(w.P, dP, B, N, C, sp, dk, C, st);
}
}  // extern "C"

extern "C" {
// ---- encoder glue of the training row: BatchNorm (batch statistics), strided depthwise 3x3, CoordAtt pooling / gating ----
size_t tcx_bn_act_train_workspace_bytes(long long M, int C) { return 4 * bn_train_scratch_floats(M, C); }
int tcx_bn_act_train_fwd(const float* x, const float* w, const float* b, float* running_mean, float* running_var, float eps,
                         float momentum, int act, float* y, float* stat, long long M, int C, void* ws, void* stream) {
  TCX_REQUIRE(x && w && b && y && stat && ws, "bn_act_train_fwd: null pointer");
  TCX_REQUIRE(act == ACT_NONE || act == ACT_HARDSWISH || act == ACT_SILU_SWISH, "bn_act_train: activation %d not built", act);
  return launch_bn_train_fwd(x, w, b, eps, momentum, act, y, stat, running_mean, running_var, M, C, reinterpret_cast<float*>(ws), S(stream));
}
int tcx_bn_act_train_bwd(const float* x, const float* dy, const float* stat, const float* w, const float* b, int act, float* dx, float* dw,
                         float* db, long long M, int C, void* ws, void* stream) {
  TCX_REQUIRE(x && dy && stat && w && b && dx && dw && db && ws, "bn_act_train_bwd: null pointer");
  return launch_bn_train_bwd(x, dy, stat, w, b, act, dx, dw, db, M, C, reinterpret_cast<float*>(ws), S(stream));
}
int tcx_dwconv3x3_nhwc_fwd(const float* x, const float* w, float* y, int B, int H, int W, int C, int stride, void* stream) {
  TCX_REQUIRE(x && w && y, "dwconv3x3_nhwc_fwd: null pointer");
  DwGroup g{x, w, nullptr, y};
  return launch_dwconv3x3(&g, 1, B, H, W, C, stride, DW_PLAIN, BnParams{}, S(stream));
}
size_t tcx_dwconv3x3_nhwc_bwd_workspace_bytes(int B, int H, int W, int C, int stride) {
  return 4 * dw3s_scratch_floats((long long)B * ((H - 1) / stride + 1) * ((W - 1) / stride + 1), C);
}
int tcx_dwconv3x3_nhwc_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H, int W, int C, int stride,
                           void* ws, void* stream) {
  TCX_REQUIRE(x && w && dy && ws && (stride == 1 || stride == 2), "dwconv3x3_nhwc_bwd: null pointer or bad stride");
  return launch_dw3s_bwd(x, w, dy, dx, dw, B, H, W, C, stride, reinterpret_cast<float*>(ws), S(stream));
}
int tcx_coord_pool_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream) { return launch_coord_pool(x, B, H, W, C, y, S(stream)); }
int tcx_coord_pool_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
  return launch_coord_pool_bwd(dy, B, H, W, C, dx, S(stream));
}
int tcx_coord_gate_fwd(const float* x, const float* z, float* out, int B, int H, int W, int C, void* stream) {
  return launch_coord_gate(x, z, B, H, W, C, out, S(stream));
}
int tcx_coord_gate_bwd(const float* x, const float* z, const float* dout, float* dx, float* dz, int B, int H, int W, int C, void* stream) {
  return launch_coord_gate_bwd(x, z, dout, B, H, W, C, dx, dz, S(stream));
}
}  // extern "C"

extern "C" {

// floats of the LayerNorm-backward column partials: the two-pass kernels leave bwd_red_blocks(M) blocks, the one-pass ones
// ln_bwd_fused_blocks(M, C) (more than that on wide rows)
static size_t ln_part_floats(long long M, int C) {
  return 2 * (size_t)std::max(bwd_red_blocks(M), ln_bwd_fused_ok(M, C) ? ln_bwd_fused_blocks(M, C) : 0) * C;
}
size_t tcx_layernorm_bwd_workspace_bytes(long long M, int C) { return 4 * (rnd(2 * (size_t)M) + rnd(ln_part_floats(M, C)) + 64); }
int tcx_layernorm_bwd(const float* x, const float* w, const float* dy, const float* dres, float eps, float* dx, float* dw, float* db,
                      long long M, int C, void* ws, void* stream) {
  TCX_REQUIRE(x && w && dy && dx && dw && db && ws, "layernorm_bwd: null pointer");
  Carver c(ws);
  float* stats = c.take(2 * (size_t)M);
  float* part = c.take(ln_part_floats(M, C));
  return launch_bwd_ln(x, dy, w, nullptr, eps, 0, dx, dw, db, M, C, stats, part, S(stream), dres);
}

size_t tcx_linear_bwd_workspace_bytes(long long M, int N, int K) { return 4 * linear_bwd_ws_floats(M, N, K); }
int tcx_linear_bwd(const void* x, int x_f16, const float* w, const float* dy, float* dx, float* dw, float* db, long long M, int N,
                   int K, void* ws, void* stream) {
  TCX_REQUIRE(x && w && dy && ws, "linear_bwd: null pointer");
  return run_linear_bwd(x, x_f16, w, dy, dx, dw, db, M, N, K, reinterpret_cast<float*>(ws), S(stream));
}

size_t tcx_wgrad_mn_workspace_bytes(long long tokens, int NL, int KL, int batch, int fmt) {
  return 4 * (wgrad_tc_scratch_floats(tokens > 0 ? tokens : 1, NL, KL, batch, fmt == 2 ? 4 : 2) + 64);
}
int tcx_wgrad_mn(const void* a, const void* b, int fmt, long long tokens, int NL, int KL, int lda, int ldb, int batch, float alpha,
                 float* out, float* outT, float* db, int mask_ch, void* ws, void* stream) {
  TCX_REQUIRE(a && b && out && ws, "wgrad_mn: null pointer");
  WgradArgs g{};
  g.A = a; g.B = b; g.fmt = fmt;
  g.Mtok = tokens; g.NL = NL; g.KL = KL; g.lda = lda; g.ldb = ldb; g.batch = batch;
  g.strideA = batch > 1 ? tokens * lda : 0; g.strideB = batch > 1 ? tokens * ldb : 0;
  g.alpha = alpha; g.out = out; g.ldo = KL; g.stride_out = (long long)NL * KL;
  g.outT = outT; g.ldt = NL; g.stride_outT = (long long)NL * KL;
  g.db = db; g.mask_ch = mask_ch; g.scratch = reinterpret_cast<float*>(ws);
  return launch_wgrad_tc(g, S(stream));
}

// ---- EfficientAttention (MSTr.py:106-143) training forward / backward ----
size_t tcx_eff_attn_saved_bytes(int B, int N, int C) {
  return 4 * (eff_attn16_carve_floats(B, N, C) + rnd((size_t)B * N * C / 2 + 64) + 64);
}
int tcx_eff_attn_train_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int N, int C, void* saved,
                           void* stream) {
  TCX_REQUIRE(xn && p && y && saved, "eff_attn_train_fwd: null pointer");
  TCX_REQUIRE(eff_attn_prepared(p, N, C, 0), "eff_attn_train_fwd: needs prepared fp16 weights, N >= 32 and C %% 64 == 0 (N=%d C=%d)", N, C);
  float* base = reinterpret_cast<float*>(saved);
  __half* xn16 = H16(base + eff_attn16_carve_floats(B, N, C));
  TCX_TRY(launch_f32_to_f16(xn, xn16, (long long)B * N * C, S(stream)));
  return run_eff_attn16(xn16, p, residual, y, B, N, C, 0, base, S(stream));
}

static size_t eff_attn_bwd_ws_floats(int B, int N, int C) {
  const size_t M = (size_t)B * N, bnc = M * C, bcc = (size_t)B * C * C;
  int S, Ms;
  bwd_wgrad_splits(N, C, C, &S, &Ms);
  const size_t pack = (size_t)B * S * C * Ms, ch = (size_t)B * ea_bwd_chunks(N) * C;
  const size_t lin = std::max(linear_bwd_ws_floats(M, C, C), linear_bwd_ws_floats(M, 3 * C, C));
  return 6 * rnd(bnc) + rnd(3 * bnc) + 3 * rnd(bcc) + 2 * rnd(pack) + rnd((size_t)B * S * C * C) + 3 * rnd(ch) +
         2 * rnd(3 * (size_t)C * C) + rnd(3 * (size_t)C) + lin + 64;
}
size_t tcx_eff_attn_bwd_workspace_bytes(int B, int N, int C) { return 4 * eff_attn_bwd_ws_floats(B, N, C); }

int tcx_eff_attn_bwd(const float* dy, const void* const* p, const void* saved, float* dxn, void* const* dp, int B, int N, int C,
                     void* ws, void* stream) {
  TCX_REQUIRE(dy && p && saved && dp && ws, "eff_attn_bwd: null pointer");
  for (int i = 0; i < 8; i++) TCX_REQUIRE(dp[i] != nullptr, "eff_attn_bwd: gradient slot %d is null", i);
  cudaStream_t st = S(stream);
  const long long M = (long long)B * N;
  const size_t bnc = (size_t)M * C, bcc = (size_t)B * C * C;
  // the forward's buffers (run_eff_attn16 carve order) + the fp16 LayerNorm output behind them
  Carver sv(const_cast<void*>(saved));
  const __half* kqv = H16(sv.take(3 * bnc / 2 + 64));
  const __half* qsm = H16(sv.take(bnc / 2 + 64));
  const __half* att = H16(sv.take(bnc / 2 + 64));
  const __half* ctxT = H16(sv.take(bcc / 2 + 64));
  const __half* xn16 = H16(reinterpret_cast<float*>(const_cast<void*>(saved)) + eff_attn16_carve_floats(B, N, C));
  int SP, Ms;
  bwd_wgrad_splits(N, C, C, &SP, &Ms);
  const size_t ch = (size_t)B * ea_bwd_chunks(N) * C;
  Carver c(ws);
  float* datt = c.take(bnc);
  float* dQs = c.take(bnc);
  float* P32 = c.take(bnc);
  float* V32 = c.take(bnc);
  float* dP = c.take(bnc);
  float* spare = c.take(bnc);
  (void)spare;
  float* dkqv = c.take(3 * bnc);
  float* ctx32 = c.take(bcc);
  float* dctx = c.take(bcc);
  float* dctxT = c.take(bcc);
  float* QsT = c.take((size_t)B * SP * C * Ms);
  float* dattT = c.take((size_t)B * SP * C * Ms);
  float* part = c.take((size_t)B * SP * C * C);
  float* pm = c.take(ch);
  float* ps = c.take(ch);
  float* sp = c.take(ch);
  float* Wcat = c.take(3 * (size_t)C * C);
  float* dWcat = c.take(3 * (size_t)C * C);
  float* dbcat = c.take(3 * (size_t)C);
  float* lin = c.take(0);
  auto G = [&](int i) { return reinterpret_cast<float*>(dp[i]); };
  // reprojection: att [M][C] -> y
  TCX_TRY(run_linear_bwd(att, 1, F(p[6]), dy, datt, G(6), G(7), M, C, C, lin, st));
  // att[b] = Qs[b] ctx[b]:  dQs = datt ctx^T,  dctx = Qs^T datt
  TCX_TRY(launch_bwd_packT_batched_f16(ctxT, B, C, C, C, 1, C, C, ctx32, st));     // ctx32[b][ck][cv]
  {
    GemmParams g = gemm1(datt, ctx32, dQs, N, C, C);
    g.batch = B; g.strideA = (long long)N * C; g.strideW = (long long)C * C; g.strideC = (long long)N * C;
    TCX_TRY(launch_gemm(g, st));
  }
  TCX_TRY(launch_bwd_packT_batched_f16(qsm, B, N, C, C, SP, Ms, Ms, QsT, st));
  TCX_TRY(launch_bwd_packT_batched_f32(datt, B, N, C, C, SP, Ms, Ms, dattT, st));
  {
    GemmParams g = gemm1(QsT, dattT, part, C, C, Ms);
    g.batch = B * SP; g.strideA = (long long)C * Ms; g.strideW = (long long)C * Ms; g.strideC = (long long)C * C;
    TCX_TRY(launch_gemm(g, st));
  }
  TCX_TRY(launch_bwd_fold_batched(part, B, SP, C, dctx, dctxT, st));
  // ctx = P^T V with P = softmax over tokens of K:  dV = P dctx,  dP = V dctx^T
  TCX_TRY(launch_ea_bwd_prep(kqv, kqv + 2 * C, 3 * C, B, N, C, pm, ps, P32, V32, st));
  {
    GemmParams g = gemm1(P32, dctxT, dkqv + 2 * C, N, C, C);
    g.ldc = 3 * C;
    g.batch = B; g.strideA = (long long)N * C; g.strideW = (long long)C * C; g.strideC = (long long)N * 3 * C;
    TCX_TRY(launch_gemm(g, st));
  }
  {
    GemmParams g = gemm1(V32, dctx, dP, N, C, C);
    g.batch = B; g.strideA = (long long)N * C; g.strideW = (long long)C * C; g.strideC = (long long)N * C;
    TCX_TRY(launch_gemm(g, st));
  }
  TCX_TRY(launch_ea_bwd_softmax(P32, dP, qsm, dQs, B, N, C, sp, dkqv, st));
  // the three 1x1 convolutions as one Linear with the stacked weight [Wk; Wq; Wv]
  const size_t cc = (size_t)C * C * sizeof(float);
  for (int i = 0; i < 3; i++)
    TCX_REQUIRE(cudaMemcpyAsync(Wcat + (size_t)i * C * C, p[2 * i], cc, cudaMemcpyDeviceToDevice, st) == cudaSuccess,
                "eff_attn_bwd: weight copy failed");
  TCX_TRY(run_linear_bwd(xn16, 1, Wcat, dkqv, dxn, dWcat, dbcat, M, 3 * C, C, lin, st));
  for (int i = 0; i < 3; i++) {
    TCX_REQUIRE(cudaMemcpyAsync(dp[2 * i], dWcat + (size_t)i * C * C, cc, cudaMemcpyDeviceToDevice, st) == cudaSuccess &&
                    cudaMemcpyAsync(dp[2 * i + 1], dbcat + (size_t)i * C, C * sizeof(float), cudaMemcpyDeviceToDevice, st) == cudaSuccess,
                "eff_attn_bwd: gradient copy failed");
  }
  return 0;
}

// ---- FactorAtt_ConvRelPosEnc (MSTr.py:852-886) backward; forward = tcx_mb_factor_attn_fwd, whose workspace is kept ----
static size_t mb_attn_bwd_ws_floats(int B, int N, int C) {
  const size_t M = (size_t)B * N, bnc = M * C, bcc = (size_t)B * C * C;
  int S, Ms;
  bwd_wgrad_splits(N, C, C, &S, &Ms);
  const size_t pack = (size_t)B * S * C * Ms, ch = (size_t)B * ea_bwd_chunks(N) * C;
  const size_t lin = std::max(linear_bwd_ws_floats(M, C, C), linear_bwd_ws_floats(M, 3 * C, C));
  return 6 * rnd(bnc) + 2 * rnd(3 * bnc) + 3 * rnd(bcc) + 2 * rnd(pack) +
         rnd(std::max((size_t)B * S * C * C, wgrad_tc_scratch_floats(N, C, C, B, 4))) + 3 * rnd(ch) +
         rnd(std::max(bwd_dwk_wgrad_part_floats(7, M, C), bwd_dwk3_wgrad_part_floats(M, C))) + lin + 64;
}
size_t tcx_mb_factor_attn_bwd_workspace_bytes(int B, int N, int C) { return 4 * mb_attn_bwd_ws_floats(B, N, C); }

int tcx_mb_factor_attn_bwd(const float* dy, const float* xn, const void* const* p, const void* fwd_ws, int saved_f16, float* dxn,
                           void* const* dp, int B, int H, int W, int C, int heads, void* ws, void* stream) {
  TCX_REQUIRE(dy && xn && p && fwd_ws && dp && ws, "mb_factor_attn_bwd: null pointer");
  TCX_REQUIRE(heads == 8 && C % heads == 0 && C % 4 == 0, "mb_factor_attn_bwd: built for 8 heads (crpe windows 3/5/7 on 2/3/3 heads)");
  for (int i = 0; i < 10; i++) TCX_REQUIRE(dp[i] != nullptr, "mb_factor_attn_bwd: gradient slot %d is null", i);
  cudaStream_t st = S(stream);
  const int N = H * W, Ch = C / heads;
  const long long M = (long long)B * N;
  const size_t bnc = (size_t)M * C, bcc = (size_t)B * C * C;
  const float scale = 1.0f / sqrtf((float)Ch);
  // forward buffers: fp32 (tcx_mb_factor_attn_fwd carve order: q | k | v rows, context, attention output before the projection)
  // or the fp16 pair of tcx_mb_factor_attn_train_fwd (q | k | v rows, attention output), widened to fp32 here
  Carver f(const_cast<void*>(fwd_ws));
  Carver c(ws);
  const float* qkv;
  const void* att;
  if (saved_f16) {
    const __half* qkv16 = H16(f.take(3 * bnc / 2 + 64));
    att = H16(f.take(bnc / 2 + 64));
    float* q32 = c.take(3 * bnc);
    TCX_TRY(launch_f16_to_f32(qkv16, q32, (long long)(3 * bnc), st));
    qkv = q32;
  } else {
    qkv = f.take(3 * bnc);
    f.take(bcc);
    att = f.take(bnc);
  }
  const float* q = qkv; const float* k = qkv + C; const float* v = qkv + 2 * C;
  const int ld = 3 * C;
  int SP, Ms;
  bwd_wgrad_splits(N, C, C, &SP, &Ms);
  const size_t ch = (size_t)B * ea_bwd_chunks(N) * C;
  float* dxo = c.take(bnc);
  float* P = c.take(bnc);
  float* dP = c.take(bnc);
  float* dqfa = c.take(bnc);
  float* convv = c.take(bnc);
  float* dvconv = c.take(bnc);
  float* dqkv = c.take(3 * bnc);
  float* ctx = c.take(bcc);
  float* dctx = c.take(bcc);
  float* dctxT = c.take(bcc);
  float* packA = c.take((size_t)B * SP * C * Ms);
  float* packB = c.take((size_t)B * SP * C * Ms);
  float* part = c.take(std::max((size_t)B * SP * C * C, wgrad_tc_scratch_floats(N, C, C, B, 4)));
  float* pm = c.take(ch);
  float* ps = c.take(ch);
  float* sp = c.take(ch);
  float* wpart = c.take(std::max(bwd_dwk_wgrad_part_floats(7, M, C), bwd_dwk3_wgrad_part_floats(M, C)));
  float* lin = c.take(0);
  auto G = [&](int i) { return reinterpret_cast<float*>(dp[i]); };
  const int win[3] = {3, 5, 7}, c0[3] = {0, 2 * Ch, 5 * Ch}, cg[3] = {2 * Ch, 3 * Ch, 3 * Ch};
  const bool mn = g_flag_wgrad_tc && g_flag_gemm_tc && wgrad_tc_eligible(N, C, C, C, C, 4) && ld % 4 == 0;
  const float* cw3[3] = {F(p[2]), F(p[4]), F(p[6])};
  const float* cb3[3] = {F(p[3]), F(p[5]), F(p[7])};
  auto ctx_gemm = [&](const float* a, int lda_, const float* b_, int ldb_, float sc, float* out, float* outT) -> int {
    // out[b] = sc * mask(a[b]^T b_[b]) over the tokens of image b
    if (mn) {      // both operands read in place (MN-major), head mask, scale and transposed copy inside the kernel's fold
      WgradArgs w{};
      w.A = a; w.B = b_; w.fmt = 2; w.Mtok = N; w.NL = C; w.KL = C; w.lda = lda_; w.ldb = ldb_; w.batch = B;
      w.strideA = (long long)N * lda_; w.strideB = (long long)N * ldb_;
      w.alpha = sc; w.out = out; w.ldo = C; w.stride_out = (long long)C * C;
      w.outT = outT; w.ldt = C; w.stride_outT = (long long)C * C; w.mask_ch = Ch; w.scratch = part;
      return launch_wgrad_tc(w, st);
    }
    TCX_TRY(launch_bwd_packT_batched_f32(a, B, N, C, lda_, SP, Ms, Ms, packA, st));
    TCX_TRY(launch_bwd_packT_batched_f32(b_, B, N, C, ldb_, SP, Ms, Ms, packB, st));
    GemmParams g = gemm1(packA, packB, part, C, C, Ms);
    g.batch = B * SP; g.strideA = (long long)C * Ms; g.strideW = (long long)C * Ms; g.strideC = (long long)C * C;
    TCX_TRY(launch_gemm(g, st));
    return launch_bwd_fold_mask(part, B, SP, C, Ch, sc, out, outT, st);
  };
  auto tok_gemm = [&](const float* a, int lda_, const float* w, float* out, int ldo, const float* res) -> int {
    // out[b] (N x C) = a[b] (N x C) w[b]^T (+ res)
    GemmParams g = gemm1(a, w, out, N, C, C);
    g.lda = lda_; g.ldc = ldo;
    g.batch = B; g.strideA = (long long)N * lda_; g.strideW = (long long)C * C; g.strideC = (long long)N * ldo;
    if (res) { g.g[0].epi.residual = res; g.g[0].epi.ldr = C; g.g[0].epi.strideR = (long long)N * C; }
    return launch_gemm(g, st);
  };
  // projection: att [M][C] -> y
  TCX_TRY(run_linear_bwd(att, saved_f16, F(p[8]), dy, dxo, G(8), G(9), M, C, C, lin, st));
  // recompute: P = softmax over tokens of k, ctx = per-head P^T v, conv_v = crpe depthwise convolutions of v
  TCX_TRY(launch_bwd_ksoftmax32(k, ld, B, N, C, pm, ps, P, st));
  TCX_TRY(ctx_gemm(P, C, v, ld, 1.0f, ctx, nullptr));
  if (mn) {
    TCX_TRY(launch_bwd_dwk3(v, ld, cw3, cb3, 2 * Ch, 5 * Ch, convv, C, B, H, W, C, 0, 0, st));
  } else {
    for (int j = 0; j < 3; j++)
      TCX_TRY(launch_bwd_dwk(win[j], v + c0[j], ld, F(p[2 + 2 * j]), F(p[3 + 2 * j]), convv + c0[j], C, B, H, W, cg[j], 0, 0, st));
  }
  // x_out = scale * q ctx + q * conv_v
  TCX_TRY(tok_gemm(dxo, C, ctx, dqfa, C, nullptr));                       // d(q ctx)/dq, before the scale
  TCX_TRY(ctx_gemm(q, ld, dxo, C, scale, dctx, dctxT));                   // dctx = scale * mask(q^T dxo)
  TCX_TRY(launch_mb_bwd_dq(dxo, dqfa, convv, q, ld, scale, M, C, dqkv, ld, st));   // dq; convv <- d conv_v
  // the crpe filter / bias gradients need only d conv_v and v: they leave the critical stream here and rejoin at the end
  AuxStreams* auxw = mn ? aux_streams(st) : nullptr;
  if (auxw) {
    TCX_REQUIRE(cudaEventRecord(auxw->fork, st) == cudaSuccess && cudaStreamWaitEvent(auxw->s[2], auxw->fork, 0) == cudaSuccess,
                "mb_factor_attn_bwd: fork failed");
    float* dw3[3] = {G(2), G(4), G(6)};
    float* db3[3] = {G(3), G(5), G(7)};
    TCX_TRY(launch_bwd_dwk3_wgrad(convv, C, v, ld, B, H, W, C, 2 * Ch, 5 * Ch, dw3, db3, wpart, auxw->s[2]));
  }
  if (mn) {
    TCX_TRY(launch_bwd_dwk3(convv, C, cw3, nullptr, 2 * Ch, 5 * Ch, dvconv, C, B, H, W, C, 1, 0, st));
  } else {
    for (int j = 0; j < 3; j++)
      TCX_TRY(launch_bwd_dwk(win[j], convv + c0[j], C, F(p[2 + 2 * j]), nullptr, dvconv + c0[j], C, B, H, W, cg[j], 1, 0, st));
  }
  TCX_TRY(tok_gemm(P, C, dctxT, dqkv + 2 * C, ld, dvconv));               // dv = P dctx + conv^T(d conv_v)
  TCX_TRY(tok_gemm(v, ld, dctx, dP, C, nullptr));                         // dP = v dctx^T
  TCX_TRY(launch_bwd_colsoftmax(P, dP, B, N, C, sp, dqkv + C, ld, st));   // dk
  if (auxw) {
    // already running beside the main chain
  } else if (mn) {
    float* dw3[3] = {G(2), G(4), G(6)};
    float* db3[3] = {G(3), G(5), G(7)};
    TCX_TRY(launch_bwd_dwk3_wgrad(convv, C, v, ld, B, H, W, C, 2 * Ch, 5 * Ch, dw3, db3, wpart, st));
  } else {
    for (int j = 0; j < 3; j++)
      TCX_TRY(launch_bwd_dwk_wgrad(win[j], convv + c0[j], C, v + c0[j], ld, B, H, W, cg[j], G(2 + 2 * j), G(3 + 2 * j), wpart, st));
  }
  // qkv Linear
  TCX_TRY(run_linear_bwd(xn, 0, F(p[0]), dqkv, dxn, G(0), G(1), M, 3 * C, C, lin, st));
  if (auxw) TCX_TRY(join_stream(auxw, 2, st));
  return 0;
}

// ---- ConvPosEnc / DWConv (MSTr.py:744-752, :26-31) backward: y = dw3x3(x) + b (+ x) ----
size_t tcx_dwconv_tokens_bwd_workspace_bytes(int B, int H, int W, int C) {
  return 4 * (rnd(bwd_dwk_wgrad_part_floats(3, (long long)B * H * W, C)) + 64);
}
int tcx_dwconv_tokens_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int B, int H, int W, int C,
                          int add_input, void* ws, void* stream) {
  TCX_REQUIRE(x && w && dy && ws, "dwconv_tokens_bwd: null pointer");
  cudaStream_t st = S(stream);
  const long long M = (long long)B * H * W;
  AuxStreams* aux = dx && dw ? aux_streams(st) : nullptr;
  if (aux) TCX_TRY(fork_streams(aux, st, 1));
  if (dx) {
    if (add_input)
      TCX_REQUIRE(cudaMemcpyAsync(dx, dy, sizeof(float) * M * C, cudaMemcpyDeviceToDevice, st) == cudaSuccess, "dwconv_tokens_bwd: copy failed");
    TCX_TRY(launch_bwd_dwk(3, dy, C, w, nullptr, dx, C, B, H, W, C, 1, add_input ? 1 : 0, st));
  }
  if (dw) TCX_TRY(launch_bwd_dwk_wgrad(3, dy, C, x, C, B, H, W, C, dw, db, reinterpret_cast<float*>(ws), aux ? aux->s[0] : st));
  if (aux) TCX_TRY(join_stream(aux, 0, st));
  return 0;
}

// ---- bridge attention core softmax(q k^T scale) v (MSTr.py:2281-2285) backward; forward = tcx_flash_attn_fwd ----
static void attn_core_plan(int Nq, int Nk, int* S, int* Ms, int* Msk) {
  bwd_wgrad_splits(Nq, Nk, 64, S, Ms);
  *Msk = (Nk + 31) / 32 * 32;
}
size_t tcx_attn_core_bwd_workspace_bytes(int B, int Nq, int Nk) {
  int S, Ms, Msk;
  attn_core_plan(Nq, Nk, &S, &Ms, &Msk);
  const size_t sc = (size_t)B * Nq * Nk;
  return 4 * (2 * rnd(sc) + rnd((size_t)B * S * Nk * Ms) + rnd((size_t)B * S * 64 * Ms) +
              rnd(std::max((size_t)B * S * Nk * 64, wgrad_tc_scratch_floats(Nq, Nk, 64, B, 4))) + rnd((size_t)B * 64 * Msk) + 64);
}
int tcx_attn_core_bwd(const float* q, const float* kv, const float* dout, float scale, float* dq, float* dkv, int B, int Nq, int Nk,
                      void* ws, void* stream) {
  TCX_REQUIRE(q && kv && dout && dq && dkv && ws, "attn_core_bwd: null pointer");
  TCX_REQUIRE(Nk % 4 == 0 && Nk >= 16, "attn_core_bwd: Nk must be a multiple of 4 (got %d)", Nk);
  cudaStream_t st = S(stream);
  if (B == 0 || Nq == 0) return 0;
  int SP, Ms, Msk;
  attn_core_plan(Nq, Nk, &SP, &Ms, &Msk);
  const size_t sc = (size_t)B * Nq * Nk;
  Carver c(ws);
  float* P = c.take(sc);           // scores -> probabilities
  float* dS = c.take(sc);          // dP -> dS (scaled)
  float* big = c.take((size_t)B * SP * Nk * Ms);     // P^T / dS^T, K-major over the query tokens, split-major
  float* small = c.take((size_t)B * SP * 64 * Ms);   // dout^T / q^T
  float* part = c.take(std::max((size_t)B * SP * Nk * 64, wgrad_tc_scratch_floats(Nq, Nk, 64, B, 4)));
  float* kT = c.take((size_t)B * 64 * Msk);
  const float* k = kv; const float* v = kv + 64;
  const long long M = (long long)B * Nq;
  auto scores = [&](const float* a, const float* w, float* out) -> int {      // out[b] (Nq x Nk) = a[b] (Nq x 64) w[b]^T, w rows of pitch 128
    GemmParams g = gemm1(a, w, out, Nq, Nk, 64);
    g.ldw = 128;
    g.batch = B; g.strideA = (long long)Nq * 64; g.strideW = (long long)Nk * 128; g.strideC = (long long)Nq * Nk;
    return launch_gemm(g, st);
  };
  auto over_queries = [&](const float* big_src, const float* small_src, float sc_, float* out) -> int {
    // out[b] (Nk x 64, pitch 128) = sc_ * big_src[b]^T (Nk x Nq) small_src[b] (Nq x 64)
    TCX_TRY(launch_bwd_packT_batched_f32(big_src, B, Nq, Nk, Nk, SP, Ms, Ms, big, st));
    TCX_TRY(launch_bwd_packT_batched_f32(small_src, B, Nq, 64, 64, SP, Ms, Ms, small, st));
    GemmParams g = gemm1(big, small, part, Nk, 64, Ms);
    g.batch = B * SP; g.strideA = (long long)Nk * Ms; g.strideW = (long long)64 * Ms; g.strideC = (long long)Nk * 64;
    TCX_TRY(launch_gemm(g, st));
    return launch_bwd_fold_rows(part, B, SP, Nk, 64, sc_, out, 128, st);
  };
  TCX_TRY(scores(q, k, P));
  TCX_TRY(launch_bwd_rowsoftmax_fwd(P, Nk, M, Nk, scale, P, Nk, st));
  TCX_TRY(scores(dout, v, dS));                                         // dP = dout v^T
  const bool mn = g_flag_wgrad_tc && g_flag_gemm_tc && wgrad_tc_eligible(Nq, Nk, 64, Nk, 64, 4);
  // out[b] (Nk x 64, pitch 128) = big_src[b]^T (Nk x Nq) small_src[b] (Nq x 64): the score-sized operand is read in place by the
  // MN-major weight-gradient kernel (the round-1 path re-laid 305 MB of scores twice per layer)
  auto over_queries_mn = [&](const float* big_src, const float* small_src, float* out) -> int {
    WgradArgs a{};
    a.A = big_src; a.B = small_src; a.fmt = 2; a.Mtok = Nq; a.NL = Nk; a.KL = 64; a.lda = Nk; a.ldb = 64; a.batch = B;
    a.strideA = (long long)Nq * Nk; a.strideB = (long long)Nq * 64;
    a.alpha = 1.0f; a.out = out; a.ldo = 128; a.stride_out = (long long)Nk * 128;
    a.scratch = part;
    return launch_wgrad_tc(a, st);
  };
  if (mn) TCX_TRY(over_queries_mn(P, dout, dkv + 64));                  // dv = P^T dout
  else TCX_TRY(over_queries(P, dout, 1.0f, dkv + 64));
  TCX_TRY(launch_bwd_rowsoftmax_bwd(P, Nk, dS, Nk, M, Nk, scale, dS, Nk, st));
  if (mn) {
    GemmParams g = gemm1(dS, k, dq, Nq, 64, Nk);                         // dq = dS k, k [Nk][64] (pitch 128) read in place
    g.w_mn = 1; g.ldw = 128;
    g.batch = B; g.strideA = (long long)Nq * Nk; g.strideW = (long long)Nk * 128; g.strideC = (long long)Nq * 64;
    TCX_TRY(launch_gemm(g, st));
    return over_queries_mn(dS, q, dkv);                                  // dk = dS^T q
  }
  TCX_TRY(launch_bwd_packT_batched_f32(k, B, Nk, 64, 128, 1, Msk, Msk, kT, st));
  {
    GemmParams g = gemm1(dS, kT, dq, Nq, 64, Nk);                        // dq = dS k
    g.ldw = Msk;
    g.batch = B; g.strideA = (long long)Nq * Nk; g.strideW = (long long)64 * Msk; g.strideC = (long long)Nq * 64;
    TCX_TRY(launch_gemm(g, st));
  }
  return over_queries(dS, q, 1.0f, dkv);                                // dk = dS^T q
}

size_t tcx_mixffn_skip_saved_bytes(int B, int N, int C, int C4) { return 4 * mix_saved(nullptr, (long long)B * N, C, C4).floats; }
int tcx_mixffn_skip_train_fwd(const float* xn, const void* const* p, float ln_eps, const float* residual, float* y, int B, int H,
                              int W, int C, int C4, void* saved, const void* xn16, void* stream) {
  TCX_REQUIRE(xn && p && y && saved, "mixffn_skip_train_fwd: null pointer");
  Mix16 m{};
  TCX_REQUIRE(mix16_fill(p, m), "mixffn_skip_train_fwd: fc1 / fc2 weights are not prepared (tcx_prepare_weight_f16)");
  const long long M = (long long)B * H * W;
  MixSaved s = mix_saved(saved, M, C, C4);
  // xn16 (nullable): the fp16 twin of xn written by the LayerNorm in front (tcx_layernorm_dual_fwd), read in place — the copy in
  // `saved` is then NOT written, and tcx_mixffn_skip_bwd must be given xn32; without it xn is converted here
  if (!xn16) TCX_TRY(launch_f32_to_f16(xn, s.xn16, M * C, S(stream)));
  m.xn = xn16 ? reinterpret_cast<const __half*>(xn16) : s.xn16; m.res = residual; m.y = y; m.u_save = s.u;
  return run_mixffn16(1, &m, ln_eps, B, H, W, C, C4, s.h16, s.a16, S(stream));
}

size_t tcx_mixffn_skip_bwd_workspace_bytes(int B, int N, int C, int C4) {
  const long long M = (long long)B * N;
  const size_t lin = std::max(linear_bwd_ws_floats(M, C, C4), linear_bwd_ws_floats(M, C4, C));
  const size_t part = std::max((size_t)10 * bwd_red_blocks(M) * C4, (size_t)10 * dw_bwd_fused_blocks(M, C4) * C4);
  return 4 * (3 * rnd((size_t)M * C4) + rnd(2 * (size_t)M) + rnd(9 * (size_t)C4) + rnd(part) + 2 * rnd(ln_part_floats(M, C4)) +
              2 * lin + 64);
}
// xn32 (nullable): the fp32 LayerNorm output the forward was given (the TF32 operand of fc1's weight gradient); without it the
// saved fp16 copy is converted.
int tcx_mixffn_skip_bwd(const float* dy, const void* const* p, float ln_eps, const void* saved, const float* xn32, float* dxn,
                        void* const* dp, int B, int H, int W, int C, int C4, void* ws, void* stream) {
  TCX_REQUIRE(dy && p && saved && dp && ws, "mixffn_skip_bwd: null pointer");
  for (int i = 0; i < 8; i++) TCX_REQUIRE(dp[i] != nullptr, "mixffn_skip_bwd: gradient slot %d is null", i);
  cudaStream_t st = S(stream);
  const long long M = (long long)B * H * W;
  const MixSaved s = mix_saved(const_cast<void*>(saved), M, C, C4);
  Carver c(ws);
  float* da = c.take((size_t)M * C4);      // dL/d a (fc2 input), later dL/d h
  float* du = c.take((size_t)M * C4);      // dL/d u (LayerNorm input)
  float* a32 = c.take((size_t)M * C4);     // GELU(LN(u)) recomputed in fp32
  float* stats = c.take(2 * (size_t)M);
  float* wflip = c.take(9 * (size_t)C4);
  float* part = c.take(std::max((size_t)10 * bwd_red_blocks(M) * C4, (size_t)10 * dw_bwd_fused_blocks(M, C4) * C4));
  float* part_ln = c.take(ln_part_floats(M, C4));
  float* part_ln2 = c.take(ln_part_floats(M, C4));      // second-path LayerNorm partials (kept apart from `part`)
  float* lin = c.take(linear_bwd_ws_floats(M, C, C4) > linear_bwd_ws_floats(M, C4, C) ? linear_bwd_ws_floats(M, C, C4)
                                                                                       : linear_bwd_ws_floats(M, C4, C));
  float* lin2 = c.take(0);
  auto G = [&](int i) { return reinterpret_cast<float*>(dp[i]); };
  const bool fused = g_flag_wgrad_tc && ln_bwd_fused_ok(M, C4) && linear_bwd_mn_ok(M, C, C4) && linear_bwd_mn_ok(M, C4, C);
  if (fused) {
    // critical chain on `st`: dgrad fc2 -> LN/GELU backward -> depthwise backward -> dgrad fc1; the weight gradients and the
    // folds of the parameter sums hang off it on an auxiliary stream (nothing downstream of this call waits on them but the join)
    AuxStreams* aux = aux_streams(st);
    cudaStream_t sa = aux ? aux->s[1] : st;
    TCX_TRY(run_linear_bwd(nullptr, 0, F(p[6]), dy, da, nullptr, nullptr, M, C, C4, lin, st));               // da = dy W2
    TCX_TRY(launch_ln_bwd_fused(s.u, da, F(p[4]), F(p[5]), ln_eps, 1, du, a32, nullptr, M, C4, part_ln, st));
    // fc2's weight gradient and the LayerNorm parameter sums need only a32 / part_ln: they leave the chain HERE, beside the
    // depthwise backward; the depthwise filter fold follows on the same stream once that kernel is done (second event)
    if (aux) {
      TCX_REQUIRE(cudaEventRecord(aux->fork, st) == cudaSuccess && cudaStreamWaitEvent(sa, aux->fork, 0) == cudaSuccess,
                  "mixffn_skip_bwd: fork failed");
    }
    TCX_TRY(run_linear_bwd(a32, 0, F(p[6]), dy, nullptr, G(6), G(7), M, C, C4, lin2, sa));                   // dW2 = dy^T a, db2
    TCX_TRY(launch_bwd_ln_fold(part_ln, ln_bwd_fused_blocks(M, C4), C4, G(4), G(5), sa));
    float* dh = da;
    int dw_nblk = 0;
    TCX_TRY(launch_dw_bwd_fused(du, s.h16, F(p[2]), dh, B, H, W, C4, part, st, &dw_nblk));
    if (aux) {
      TCX_REQUIRE(cudaEventRecord(aux->join[2], st) == cudaSuccess && cudaStreamWaitEvent(sa, aux->join[2], 0) == cudaSuccess,
                  "mixffn_skip_bwd: second fork failed");
    }
    TCX_TRY(launch_bwd_dw_fold(part, dw_nblk, C4, G(2), G(3), sa));
    // fc1: dxn = dh W1 on the main stream, dW1 = dh^T xn beside it
    const void* xn = xn32 ? (const void*)xn32 : (const void*)s.xn16;
    TCX_TRY(run_linear_bwd(xn, xn32 ? 0 : 1, F(p[0]), dh, dxn, G(0), G(1), M, C4, C, lin, st));
    if (aux) TCX_TRY(join_stream(aux, 1, st));
    return 0;
  }
  // fc2: a16 [M][C4] -> y [M][C]
  TCX_TRY(run_linear_bwd(s.a16, 1, F(p[6]), dy, da, G(6), G(7), M, C, C4, lin, st));
  // GELU(LayerNorm(u))
  TCX_TRY(launch_bwd_ln(s.u, da, F(p[4]), F(p[5]), ln_eps, 1, du, G(4), G(5), M, C4, stats, part_ln2, st));
  // u = dw3x3(h) + b + h: input gradient and the filter / bias sums in one pass (row-sweep kernel), then the fold
  float* dh = da;
  {
    (void)wflip;
    int dw_nblk = 0;
    float* part_dw = part;
    TCX_TRY(launch_dw_bwd_fused(du, s.h16, F(p[2]), dh, B, H, W, C4, part_dw, st, &dw_nblk));
    TCX_TRY(launch_bwd_dw_fold(part_dw, dw_nblk, C4, G(2), G(3), st));
  }
  // fc1: xn16 [M][C] -> h [M][C4]
  return run_linear_bwd(s.xn16, 1, F(p[0]), dh, dxn, G(0), G(1), M, C4, C, lin, st);
}

// ---- training row of the bridge: slab split / merge and Scale_reduce as single nodes (no ATen slicing, zero fills or accumulation
// adds around them) ----
int tcx_bridge_split_fwd(const float* tokens, void* const* slabs, int B, int S0, void* stream) {
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  TCX_REQUIRE(tokens && slabs, "bridge_split: null pointer");
  UngroupArgs a{};
  a.src = tokens;
  for (int i = 0; i < 4; i++) { a.dst[i] = reinterpret_cast<float*>(slabs[i]); TCX_REQUIRE(a.dst[i], "bridge_split: slab %d is null", i); }
  for (int i = 0; i < 5; i++) a.tok_off[i] = g.off[i];
  a.ntok = g.ntok; a.B = B;
  return launch_ungroup(a, S(stream));
}
// tokens = regroup(slabs) (+ residual): BridgLayer_4's `tx1 + cat(...)` (MSTr.py:2403-2405) / the token buffer of :2380-2386
int tcx_bridge_merge_fwd(const void* const* slabs, const float* residual, float* tokens, int B, int S0, void* stream) {
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  TCX_REQUIRE(tokens && slabs, "bridge_merge: null pointer");
  RegroupArgs a{};
  for (int i = 0; i < 4; i++) { a.src[i] = F(slabs[i]); TCX_REQUIRE(a.src[i], "bridge_merge: slab %d is null", i); }
  for (int i = 0; i < 5; i++) a.tok_off[i] = g.off[i];
  a.dst = tokens; a.ntok = g.ntok; a.B = B; a.res = residual;
  return launch_regroup(a, S(stream));
}

namespace {
const int kSrRatio[3] = {8, 4, 2};
struct SrSaved { float* A[3]; size_t floats; };
SrSaved sr_saved(void* base, int B, const BridgeGeom& g) {
  Carver c(base);
  SrSaved s;
  for (int k = 0; k < 3; k++) s.A[k] = c.take((size_t)B * g.pp * g.ch[k] * kSrRatio[k] * kSrRatio[k]);
  s.floats = c.off + 64;
  return s;
}
}  // namespace

size_t tcx_scale_reduce_saved_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  return 4 * sr_saved(nullptr, B, g).floats;
}
size_t tcx_scale_reduce_train_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  size_t n = 64;
  for (int k = 0; k < 3; k++) n += rnd((size_t)B * g.pp * g.ch[k]);
  return 4 * n;
}
// Scale_reduce without its LayerNorm (MSTr.py:2225-2247): packed [B][nred][64]; the im2row matrices stay in `saved` for the
// weight gradients.  p = {sr0_w, sr0_b, sr1_w, sr1_b, sr2_w, sr2_b}
int tcx_scale_reduce_train_fwd(const float* x, const void* const* p, float* packed, int B, int S0, void* saved, void* ws, void* stream) {
  TCX_REQUIRE(x && p && packed && saved && ws, "scale_reduce_train_fwd: null pointer");
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  cudaStream_t st = S(stream);
  const SrSaved sv = sr_saved(saved, B, g);
  Carver c(ws);
  SrPackArgs a{};
  const long long xs_b = (long long)g.ntok * 64;
  AuxStreams* aux = aux_streams(st);
  if (aux) TCX_TRY(fork_streams(aux, st, 2));
  for (int k = 0; k < 3; k++) {          // three independent im2row -> GEMM chains
    const int r = kSrRatio[k], Cin = g.ch[k];
    const int K = Cin * r * r, M = B * g.pp;
    cudaStream_t sk = (aux && k > 0) ? aux->s[k - 1] : st;
    float* conv = c.take((size_t)M * Cin);
    TCX_TRY(launch_sr_im2row(x + (long long)g.off[k] * 64, xs_b, g.hw[k], Cin, r, B, sv.A[k], sk));
    GemmParams gp = gemm1(sv.A[k], F(p[2 * k]), conv, M, Cin, K);
    gp.g[0].epi.bias = F(p[2 * k + 1]);
    TCX_TRY(launch_gemm(gp, sk));
    a.conv[k] = conv; a.gmul[k] = Cin / 64; a.pp[k] = g.pp;
  }
  if (aux)
    for (int k = 0; k < 2; k++) TCX_TRY(join_stream(aux, k, st));
  a.x = x; a.xs_b = xs_b; a.raw_tok0 = g.off[3];
  for (int i = 0; i < 4; i++) a.red_off[i] = g.red_off[i];
  a.nred = g.nred; a.B = B; a.lnw = nullptr; a.lnb = nullptr; a.eps = 0.f; a.out = packed;
  return launch_sr_pack_ln(a, st);
}

size_t tcx_scale_reduce_bwd_workspace_bytes(int B, int S0) {
  BridgeGeom g;
  if (!bridge_geom(S0, g)) return 0;
  size_t n = 64;
  for (int k = 0; k < 3; k++) {
    const long long M = (long long)B * g.pp;
    const int Cin = g.ch[k], K = Cin * kSrRatio[k] * kSrRatio[k];
    n += rnd((size_t)M * Cin) + rnd((size_t)M * K) + rnd(linear_bwd_ws_floats(M, Cin, K));
  }
  return 4 * n;
}
// dpacked [B][nred][64] -> dx [B][ntok][64] (EVERY row written: three conv slabs + the raw stage-4 rows) and dp = {dsr0_w, dsr0_b,
// dsr1_w, dsr1_b, dsr2_w, dsr2_b}
int tcx_scale_reduce_bwd(const float* dpacked, const void* const* p, const void* saved, float* dx, void* const* dp, int B, int S0,
                         void* ws, void* stream) {
  TCX_REQUIRE(dpacked && p && saved && dx && dp && ws, "scale_reduce_bwd: null pointer");
  for (int i = 0; i < 6; i++) TCX_REQUIRE(dp[i] != nullptr, "scale_reduce_bwd: gradient slot %d is null", i);
  BridgeGeom g;
  TCX_REQUIRE(bridge_geom(S0, g), "bridge: stage-1 side %d must be a positive multiple of 8", S0);
  cudaStream_t st = S(stream);
  const SrSaved sv = sr_saved(const_cast<void*>(saved), B, g);
  Carver c(ws);
  const long long xs_b = (long long)g.ntok * 64;
  float* dconv[3]; float* dA[3]; float* lin[3];
  SrUnpackArgs u{};
  for (int k = 0; k < 3; k++) {
    const long long M = (long long)B * g.pp;
    const int Cin = g.ch[k], K = Cin * kSrRatio[k] * kSrRatio[k];
    dconv[k] = c.take((size_t)M * Cin); dA[k] = c.take((size_t)M * K); lin[k] = c.take(linear_bwd_ws_floats(M, Cin, K));
    u.dconv[k] = dconv[k]; u.gmul[k] = Cin / 64; u.pp[k] = g.pp;
  }
  u.dred = dpacked; u.dx = dx; u.xs_b = xs_b; u.raw_tok0 = g.off[3];
  for (int i = 0; i < 4; i++) u.red_off[i] = g.red_off[i];
  u.nred = g.nred; u.B = B;
  TCX_TRY(launch_sr_unpack(u, st));
  // aux stream 0 of `st` belongs to the weight-gradient launch inside run_linear_bwd(st): the side chains take 1 and 2
  AuxStreams* aux = aux_streams(st);
  if (aux) {
    TCX_REQUIRE(cudaEventRecord(aux->fork, st) == cudaSuccess && cudaStreamWaitEvent(aux->s[1], aux->fork, 0) == cudaSuccess &&
                cudaStreamWaitEvent(aux->s[2], aux->fork, 0) == cudaSuccess, "scale_reduce_bwd: fork failed");
  }
  for (int k = 0; k < 3; k++) {
    const long long M = (long long)B * g.pp;
    const int r = kSrRatio[k], Cin = g.ch[k], K = Cin * r * r;
    cudaStream_t sk = (aux && k > 0) ? aux->s[k] : st;
    TCX_TRY(run_linear_bwd(sv.A[k], 0, F(p[2 * k]), dconv[k], dA[k], reinterpret_cast<float*>(dp[2 * k]),
                           reinterpret_cast<float*>(dp[2 * k + 1]), M, Cin, K, lin[k], sk));
    TCX_TRY(launch_sr_row2im(dA[k], xs_b, g.hw[k], Cin, r, B, dx + (long long)g.off[k] * 64, sk));
  }
  if (aux) { TCX_TRY(join_stream(aux, 1, st)); TCX_TRY(join_stream(aux, 2, st)); }
  return 0;
}

// ---- training row of the class head: FinalPatchExpand_X4's pixel shuffle + LayerNorm(64) (MSTr.py:212-227) and the 1x1 conv to
// classes (:288-289) on the expand output e [B*H*W][1024] (the Linear node before it stays a Linear node) ----
int tcx_final_head_train_fwd(const float* e, const float* lnw, const float* lnb, float eps, const float* cls_w, const float* cls_b,
                             int ncls, float* logits_nchw, int B, int H, int W, void* stream) {
  TCX_REQUIRE(e && lnw && lnb && cls_w && cls_b && logits_nchw, "final_head_train_fwd: null pointer");
  return launch_final_head(e, B, H, W, lnw, lnb, eps, cls_w, cls_b, ncls, logits_nchw, S(stream));
}
size_t tcx_final_head_bwd_workspace_bytes(int B, int H, int W) { return 4 * (final_head_bwd_part_floats(B, H, W) + 64); }
// dlogits [B][ncls][4H][4W] -> de [B*H*W][1024] and d{ln_w, ln_b, cls_w, cls_b}
int tcx_final_head_bwd(const float* e, const float* dlogits, const float* lnw, const float* lnb, float eps, const float* cls_w, int ncls,
                       float* de, float* dlnw, float* dlnb, float* dcls_w, float* dcls_b, int B, int H, int W, void* ws, void* stream) {
  TCX_REQUIRE(e && dlogits && lnw && lnb && cls_w && de && dlnw && dlnb && dcls_w && dcls_b && ws, "final_head_bwd: null pointer");
  cudaStream_t st = S(stream);
  float* part = reinterpret_cast<float*>(ws);
  int nblk = 0;
  TCX_TRY(launch_final_head_bwd(e, dlogits, B, H, W, lnw, eps, cls_w, ncls, de, part, &nblk, st));
  AuxStreams* aux = aux_streams(st);
  if (aux) TCX_TRY(fork_streams(aux, st, 1));
  TCX_TRY(launch_final_head_bwd_fold(part, nblk, lnw, lnb, cls_w, ncls, dlnw, dlnb, dcls_w, dcls_b, aux ? aux->s[0] : st));
  if (aux) TCX_TRY(join_stream(aux, 0, st));
  return 0;
}

// ---- training row of the stem (MSTr.py:299-304): the patch matrix of the 7x7 / 4 conv, needed only by its weight gradient
// (dW = dy^T A through tcx_linear_bwd); the forward is tcx_patch_embed_ln_fwd with lnw = lnb = NULL (conv output, no LayerNorm) ----
int tcx_patch_im2row_fwd(const float* x, int B, int Cin, int H, int W, float* patches, int Kp, void* stream) {
  TCX_REQUIRE(x && patches, "patch_im2row: null pointer");
  TCX_REQUIRE(Cin == 1 || Cin == 3, "patch_im2row: Cin must be 1 or 3 (got %d)", Cin);
  const long long plane = (long long)H * W;
  return launch_patch_im2row(x, Cin * plane, Cin == 1 ? 0 : plane, B, H, W, Kp, patches, S(stream));
}

// out = sum of n <= 16 equally sized fp32 tensors, added in index order: the gradient of a parameter shared by the blocks of an
// MHCAEncoder (ConvPosEnc / ConvRelPosEnc, MSTr.py:966-978) from the gradients each block produced
int tcx_sum_tensors(const void* const* srcs, int n, long long numel, float* out, void* stream) {
  TCX_REQUIRE(srcs && out, "sum_tensors: null pointer");
  return launch_sum_tensors(reinterpret_cast<const float* const*>(srcs), n, numel, out, S(stream));
}

}  // extern "C"
