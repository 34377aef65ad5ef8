// Shared helpers for the transception_sm100 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define TCX_MAX_GROUPS 4

// ---- error plumbing (C-ABI: every export returns int, 0 = OK) ------------------------
void tcx_set_error(const char* fmt, ...);
int tcx_check_launch(const char* what);

// Optional per-kernel timing (bench.py's roofline leg): when profiling of `name` is enabled, the scope records a
// CUDA event pair on the launching stream around the launch. Disabled = one predictable branch.
extern bool g_tcx_prof_on;
void tcx_prof_begin(const char* name, cudaStream_t st);
void tcx_prof_end(const char* name, cudaStream_t st, double work);
struct ProfScope {
  const char* name;
  cudaStream_t st;
  double work;     // algorithmic bytes (HBM-bound kernels) or FLOPs (tensor-bound) of this launch, for the roofline leg
  ProfScope(const char* n, cudaStream_t s, double w = 0.0) : name(n), st(s), work(w) { if (g_tcx_prof_on) tcx_prof_begin(name, st); }
  ~ProfScope() { if (g_tcx_prof_on) tcx_prof_end(name, st, work); }
};

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// Kernels of the fp16 pipeline are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel calls
// pdl_trigger() at its top (the next kernel in the stream may then be scheduled as SMs free up) and pdl_wait() before
// its first access to activations / workspace (blocks until the preceding kernel has completed and flushed).  Only
// module parameters and prepared weights — never written inside a forward — may be read before pdl_wait(), so a
// kernel's prologue (barrier init, TMEM allocation, tensor-map prefetch, filter staging) overlaps its predecessor's
// tail.  Both instructions are no-ops for a kernel launched without the attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// First statement of a kernel launched through tcx_launch_chain: let the next kernel of the stream be scheduled, then block
// until the preceding kernel has completed and flushed.  Nothing is read before the wait, so the only thing that overlaps
// the predecessor is this kernel's own launch and block scheduling.
#define PDL_TOP()  \
  do {             \
    pdl_trigger(); \
    pdl_wait();    \
  } while (0)
#endif
extern int g_tcx_pdl;   // flag "pdl" (default 1)
extern int g_tcx_pdl_chain;   // flag "pdl_chain": the training-row kernels (backward, loss, optimizer, glue) carry the attribute too
extern int g_tcx_smem_kb;    // flag "smem_kb" (default 0 = whole SM): shared-memory budget of the tcgen05 GEMM kernels
extern int g_tcx_wgrad_ctas;   // flag "wgrad_ctas" (0 = resident-cluster capacity): CTA budget of one weight-gradient launch
extern int g_tcx_wgrad_idle;   // flag "wgrad_idle": allow splits that leave some CTAs without tokens
extern int g_tcx_max_ctas;   // flag "max_ctas" (default 0 = one CTA per SM)
template <typename... KArgs, typename... Args>
inline cudaError_t tcx_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_tcx_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Launch of a kernel that starts with PDL_TOP() (the CUDA-core kernels of the training row): same attribute as tcx_launch_pdl,
// switched by "pdl" AND "pdl_chain".  Errors are picked up by the tcx_check_launch that follows every launch.
template <typename... KArgs, typename... Args>
inline cudaError_t tcx_launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_tcx_pdl && g_tcx_pdl_chain) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define TCX_TRY(expr)                       \
  do {                                      \
    int _e = (expr);                        \
    if (_e != 0) return _e;                 \
  } while (0)

#define TCX_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      tcx_set_error(__VA_ARGS__);           \
      return -1;                            \
    }                                       \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute (dynamic shared memory opt-in) is per DEVICE: one process may drive several GPUs (nn.DataParallel), so
// the "done once" guards of the launchers are per device.  first() is true the first time it is called on the current device.
struct PerDeviceOnce {
  unsigned long long mask = 0;
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
  }
};
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- activations -----------------------------------------------------------------------
enum TcxAct { ACT_NONE = 0, ACT_GELU = 1, ACT_HARDSWISH = 2, ACT_SIGMOID = 3, ACT_SILU_SWISH = 4 };

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float hardswish(float x) { return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// reference MSTr.py:1275-1286: t * min(SiLU(t+3)/6, 1)
__device__ __forceinline__ float silu_swish(float t) {
  float u = t + 3.0f;
  float s = u * sigmoidf_(u) * (1.0f / 6.0f);
  return t * fminf(s, 1.0f);
}
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(v);
    case ACT_HARDSWISH: return hardswish(v);
    case ACT_SIGMOID: return sigmoidf_(v);
    case ACT_SILU_SWISH: return silu_swish(v);
    default: return v;
  }
}

// ---- warp helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// BatchNorm (eval) folded to y = x*scale + shift, computed on the fly from the four vectors.
struct BnParams {
  const float* w;
  const float* b;
  const float* rm;
  const float* rv;
  float eps;
};
__device__ __forceinline__ void bn_fold(const BnParams& bn, int c, float& scale, float& shift) {
  float s = bn.w[c] * rsqrtf(bn.rv[c] + bn.eps);
  scale = s;
  shift = bn.b[c] - bn.rm[c] * s;
}

// ---- GEMM front end (gemm.cu): C[M,N] = epi(A[M,K] * W[N,K]^T) -----------------------------
struct GemmEpi {
  const float* bias;      // [N] or null
  BnParams bn;            // bn.w == null -> none (applied after bias, before act)
  int act;                // TcxAct
  const float* residual;  // [M, ldr] or null (added after act)
  int ldr;
  long long strideR;      // batch stride of residual
  // optional second output (tensor-core kernel, fp16 operands, fp32 C): ln_out[m, 64g .. 64g+63] = fp16 LayerNorm over
  // each 64-column group of the finished row (affine ln_w / ln_b of length 64) — the consumer GEMM's A operand
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  void* ln_out;           // fp16 [M, ld_ln] or null
  int ld_ln;
  long long stride_ln;
};
struct GemmGroup {
  const float* A;
  const float* W;
  float* C;
  GemmEpi epi;
};
struct GemmParams {
  GemmGroup g[TCX_MAX_GROUPS];
  int groups;   // blockIdx.z = group * batch + b
  int batch;
  long long strideA, strideW, strideC;  // per-batch element strides (0 = shared)
  int M, N, K;
  int lda, ldw, ldc;
  // tensor-core kernel only: A and W point at fp16 data (kind::f16 MMA) / C is written as fp16.  Leading dimensions
  // and strides stay in elements of the respective type.  fp16 output: bias only (no BN / activation / residual).
  int ab16, out16;
  // tensor-core kernel only: W is stored [K][N] with N contiguous (row pitch ldw, batch stride strideW) and is read in place
  // as an MN-major tcgen05 operand — the input-gradient GEMM dx = dy W of a Linear reads the forward's [N][K] weight as is.
  int w_mn;
  // 16-bit operands only: A and W hold bf16 instead of fp16 (kind::f16 rejects mixed A / B formats on B200)
  int bf16;
  // accumulator scale applied before the bias (0 = 1.0): undoes the static scale of fp16 gradient operands
  float alpha;
};
int launch_gemm(const GemmParams& p, cudaStream_t st);

// convenience: single plain linear
int launch_linear(const float* A, const float* W, const float* bias, const float* residual, float* C, int M, int N,
                  int K, int act, cudaStream_t st);

int launch_f32_to_f16(const float* src, void* dst, long long n, cudaStream_t st);

// ---- elementwise / normalisation (elementwise.cu) -----------------------------------------
int launch_layernorm(const float* x, const float* w, const float* b, float* y, long long M, int C, float eps,
                     cudaStream_t st);
struct LnGroup {
  const float* x;
  const float* w;
  const float* b;
  float* y;
};
int launch_layernorm_grouped(const LnGroup* g, int groups, long long M, int C, float eps, cudaStream_t st);

enum DwEpi { DW_PLAIN = 0, DW_ADD_INPUT = 1, DW_BN_HS = 2 };
struct DwGroup {
  const float* x;   // [B,H,W,C]
  const float* w;   // [C,1,k,k]
  const float* b;   // [C] or null
  float* y;         // [B,Ho,Wo,C]
};
int launch_dwconv3x3(const DwGroup* g, int groups, int B, int H, int W, int C, int stride, int epi, BnParams bn,
                     cudaStream_t st);

struct MixMidGroup {
  const float* h;    // [B,H,W,C4] fc1 output
  const float* dww;  // [C4,1,3,3]
  const float* dwb;  // [C4]
  const float* lnw;
  const float* lnb;
  float* y;          // GELU(LN(dw(h)+b+h))
};
int launch_mixffn_mid(const MixMidGroup* g, int groups, int B, int H, int W, int C4, float eps, cudaStream_t st);
