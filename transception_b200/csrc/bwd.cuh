// Backward building blocks of the training row (kernels in bwd.cu).
// Gradients are fp32 in HBM (the per-pixel loss gradient of a bs16 224x224 step is ~1e-6: below fp16's normal range);
// the GEMM-shaped parts (dgrad, wgrad) run on the tcgen05 GEMM of gemm_tc.cu with TF32 operands.  Every reduction
// over tokens is two-pass (block partials in a fixed order, then a fold in index order): bit-reproducible.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

#ifdef __CUDACC__
// Ordered fold of S partials per output element by a (32 x 8) block: lane row y sums partials y, y+8, ... in order, then the
// eight row sums are added in row order — a fixed association, so the result is bit-reproducible, and the S loads of one
// element are spread over 8 threads with 4 independent loads in flight each (a single thread walking 592 partials is a
// 592-deep dependent-latency chain).  Valid on threadIdx.y == 0.
__device__ __forceinline__ float bwd_fold_sum(const float* __restrict__ part, int S, long long n, long long i, bool live) {
  __shared__ float sm[8][33];
  float a = 0.f;
  if (live) {
    int k = threadIdx.y;
    for (; k + 24 < S; k += 32) {
      const float v0 = part[(size_t)k * n + i], v1 = part[(size_t)(k + 8) * n + i];
      const float v2 = part[(size_t)(k + 16) * n + i], v3 = part[(size_t)(k + 24) * n + i];
      a += v0; a += v1; a += v2; a += v3;
    }
    for (; k < S; k += 8) a += part[(size_t)k * n + i];
  }
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int l = 0; l < 8; l++) s += sm[l][threadIdx.x];
  }
  return s;
}
#endif

// rows handled by one block of the token reductions; the partial buffers below are sized with it
int bwd_red_blocks(long long M);

// out[c] = sum_m x[m][c]   (bias gradients).  part: bwd_red_blocks(M) * C floats
int launch_bwd_colsum(const float* x, long long M, int C, int ld, float* part, float* out, cudaStream_t st);

// LayerNorm backward on rows u [M][C] (the LayerNorm input) with upstream gradient dz [M][C]:
//   gelu = 0: y = LN(u),        dz = dL/dy
//   gelu = 1: y = GELU(LN(u)),  dz = dL/dy (the GELU derivative is applied to the recomputed LN output)
// du [M][C] = dL/du, dgamma / dbeta [C].  stats: 2*M floats (mean, rstd written by the row pass, read by the column
// pass); part: 2 * bwd_red_blocks(M) * C floats.  du may not alias dz.
int launch_bwd_ln(const float* u, const float* dz, const float* gamma, const float* beta, float eps, int gelu, float* du,
                  float* dgamma, float* dbeta, long long M, int C, float* stats, float* part, cudaStream_t st,
                  const float* dres = nullptr);

// fused forms (bwd_mix.cu): one pass per tensor.  launch_bwd_ln uses the fused kernel by itself whenever ln_bwd_fused_ok.
bool ln_bwd_fused_ok(long long M, int C);
int ln_bwd_fused_blocks(long long M, int C);   // partial buffer: 2 * ln_bwd_fused_blocks(M, C) * C floats (may exceed bwd_red_blocks(M) for C > 512)
// du = LayerNorm (gelu: GELU o LayerNorm) backward of dz at u; act (nullable, gelu only) = fp32 GELU(LN(u)); part: 2 * blocks * C
// column partials to be folded with launch_bwd_ln_fold
// dres (nullable): added to du — the gradient arriving over the residual connection around the LayerNorm
int launch_ln_bwd_fused(const float* u, const float* dz, const float* gamma, const float* beta, float eps, int gelu, float* du,
                        float* act, const float* dres, long long M, int C, float* part, cudaStream_t st);
int launch_add_inplace(float* y, const float* x, long long n, cudaStream_t st);
int launch_sum_tensors(const float* const* srcs, int n, long long numel, float* out, cudaStream_t st);
int launch_bwd_ln_fold(const float* part, int nblk, int C, float* dgamma, float* dbeta, cudaStream_t st);
// Mix-FFN depthwise conv backward in one pass: dh = du + conv^T(du), part = 10 * blocks * C filter / bias partials (fold with
// launch_bwd_dw_fold); h is the fp16 fc1 output
int dw_bwd_fused_blocks(long long M, int C);
int launch_dw_bwd_fused(const float* du, const __half* h, const float* w, float* dh, int B, int H, int W, int C, float* part,
                        cudaStream_t st, int* nblk_out);
int launch_bwd_dw_fold(const float* part, int nblk, int C, float* dw, float* db, cudaStream_t st);

// depthwise 3x3 (stride 1, pad 1, DWConv MSTr.py:26-31) weight / bias gradient:
//   dw[c][ky][kx] = sum_p du[p][c] * h[p + (ky-1, kx-1)][c],  db[c] = sum_p du[p][c];   h is fp16 [B,H,W,C]
// part: 10 * bwd_red_blocks(B*H*W) * C floats
int launch_bwd_dwconv_wgrad(const float* du, const __half* h, int B, int H, int W, int C, float* dw, float* db, float* part,
                            cudaStream_t st);
// wflip[c][t] = w[c][8 - t]: the input gradient of the depthwise conv is the same conv with the taps mirrored
int launch_bwd_flip9(const float* w, float* wflip, int C, cudaStream_t st);

// K-major re-layout for the weight-gradient GEMM: src [M][C] (row pitch ld, fp32 or fp16) ->
// dst [S][C][Ms] fp32 with token m of split s = src row s*Ms + m (zero beyond M).  Also a plain transpose (S = 1, Ms = M).
int launch_bwd_packT_f32(const float* src, long long M, int C, int ld, int S, int Ms, float* dst, cudaStream_t st);
int launch_bwd_packT_f16(const __half* src, long long M, int C, int ld, int S, int Ms, float* dst, cudaStream_t st);

// batched forms: src [batch][N][ld] -> dst [batch][S][C][pitch] (pitch >= Ms: several packs may share one row pitch)
int launch_bwd_packT_batched_f32(const float* src, int batch, int N, int C, int ld, int S, int Ms, int pitch, float* dst, cudaStream_t st);
int launch_bwd_packT_batched_f16(const __half* src, int batch, int N, int C, int ld, int S, int Ms, int pitch, float* dst, cudaStream_t st);
// out[b] = sum_s part[b][s] for R x R matrices, plus the transposed copy outT[b] (may be null)
int launch_bwd_fold_batched(const float* part, int batch, int S, int R, float* out, float* outT, cudaStream_t st);

// ---- EfficientAttention backward pieces (MSTr.py:106-143; token-major fp16 K | Q | V rows of pitch ld) ----
int ea_bwd_chunks(int N);
// P32 [B*N][C] = softmax over the N tokens of each image of K (MSTr.py:118-122), V32 = float(V); pm / ps: B*chunks*C floats each
int launch_ea_bwd_prep(const __half* k, const __half* v, int ld, int B, int N, int C, float* pm, float* ps, float* P32, float* V32,
                       cudaStream_t st);
// dkqv [B*N][3C]: columns [0,C) = dK = P (dP - sum_n P dP), [C,2C) = dQ = Qs (dQs - sum_c Qs dQs); sp: B*chunks*C floats
int launch_ea_bwd_softmax(const float* P, const float* dP, const __half* qs, const float* dqs, int B, int N, int C, float* sp, float* dkqv,
                          cudaStream_t st);

// backward of a softmax over the N tokens of each image: dk[b*N+n][c] (row pitch ldo) = P (dP - sum_n P dP); sp: B*chunks*C floats
int launch_bwd_colsoftmax(const float* P, const float* dP, int B, int N, int C, float* sp, float* dk, int ldo, cudaStream_t st);

// ---- Multi-Branch attention / position-encoding convolutions (bwd_mb.cu) ----
// depthwise K x K (K in 3, 5, 7; stride 1, zero padding) on Cg channels of rows with pitch ldx -> ldy.
// flip = 0: y = b + conv(x) (forward); flip = 1: y = conv with mirrored taps (input gradient, no bias); add: y += instead of y =
int launch_bwd_dwk(int K, const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int B, int H, int W, int Cg, int flip,
                   int add, cudaStream_t st);
// dw [Cg][K*K] = sum_p g[p][c] x[p + off][c], db [Cg] = sum_p g (db may be null); part: bwd_dwk_wgrad_part_floats
size_t bwd_dwk_wgrad_part_floats(int K, long long M, int Cg);
int launch_bwd_dwk_wgrad(int K, const float* g, int ldg, const float* x, int ldx, int B, int H, int W, int Cg, float* dw, float* db,
                         float* part, cudaStream_t st);
// the three crpe windows (3x3 | 5x5 | 7x7 on channel ranges [0,c1) | [c1,c2) | [c2,C)) in one launch: w[j] [Cg_j][K_j^2], b[j] nullable
int launch_bwd_dwk3(const float* x, int ldx, const float* const* w, const float* const* b, int c1, int c2, float* y, int ldy, int B, int H,
                    int W, int C, int flip, int add, cudaStream_t st);
size_t bwd_dwk3_wgrad_part_floats(long long M, int C);
int launch_bwd_dwk3_wgrad(const float* g, int ldg, const float* x, int ldx, int B, int H, int W, int C, int c1, int c2, float* const* dw,
                          float* const* db, float* part, cudaStream_t st);
// out[b] = scale * sum_s part[b][s] restricted to the diagonal Ch x Ch blocks of the R x R matrix (zero elsewhere), outT transposed
int launch_bwd_fold_mask(const float* part, int batch, int S, int R, int Ch, float scale, float* out, float* outT, cudaStream_t st);
// dq[r][c] (pitch ldo) = scale * dqfa + dxo * convv;  convv <- dxo * q (the gradient of the conv output), q rows of pitch ldq
int launch_mb_bwd_dq(const float* dxo, const float* dqfa, float* convv, const float* q, int ldq, float scale, long long M, int C, float* dq,
                     int ldo, cudaStream_t st);
// P [B*N][C] = softmax over the N tokens of each image of fp32 k rows (pitch ld); pm / ps: B*chunks*C floats each
int launch_bwd_ksoftmax32(const float* k, int ld, int B, int N, int C, float* pm, float* ps, float* P, cudaStream_t st);

// row softmax y = softmax(scale * x) over n elements per row, and its backward ds = scale * P (dP - sum P dP) (ds may alias dP)
int launch_bwd_rowsoftmax_fwd(const float* x, int ld, long long M, int n, float scale, float* y, int ldy, cudaStream_t st);
int launch_bwd_rowsoftmax_bwd(const float* P, int ldp, const float* dP, int ldd, long long M, int n, float scale, float* ds, int lds,
                              cudaStream_t st);
// out[b][r][c] (row pitch ldo) = scale * sum_s part[b][s][r][c], r < R, c < Cc
int launch_bwd_fold_rows(const float* part, int batch, int S, int R, int Cc, float scale, float* out, int ldo, cudaStream_t st);

// ---- encoder glue (bwd_enc.cu) ----
// BatchNorm with batch statistics + activation (TcxAct: 0 none, 2 Hardswish, 4 silu_swish) on rows x [M][C]:
// y = act((x - mean) / sqrt(var + eps) * w + b); stat (2C floats) = mean | 1/std for backward; rm / rv (nullable) get the
// nn.BatchNorm2d running update (momentum, unbiased variance).  scratch: bn_train_scratch_floats
size_t bn_train_scratch_floats(long long M, int C);
int launch_bn_train_fwd(const float* x, const float* w, const float* b, float eps, float momentum, int act, float* y, float* stat, float* rm,
                        float* rv, long long M, int C, float* scratch, cudaStream_t st);
int launch_bn_train_bwd(const float* x, const float* dy, const float* stat, const float* w, const float* b, int act, float* dx, float* dw,
                        float* db, long long M, int C, float* scratch, cudaStream_t st);
// depthwise 3x3 (pad 1, stride 1 | 2, no bias) on NHWC x [B,H,W,C]: dx (nullable) and dw [C][9] (nullable) from dy [B,Ho,Wo,C]
size_t dw3s_scratch_floats(long long Mo, int C);
int launch_dw3s_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H, int W, int C, int stride,
                    float* scratch, cudaStream_t st);
// CoordAtt (MSTr.py:1322-1348): y [B][H+W][C] = (mean over W | mean over H); gate out = x sigmoid(z_h) sigmoid(z_w) with z [B][H+W][C]
int launch_coord_pool(const float* x, int B, int H, int W, int C, float* y, cudaStream_t st);
int launch_coord_pool_bwd(const float* dy, int B, int H, int W, int C, float* dx, cudaStream_t st);
int launch_coord_gate(const float* x, const float* z, int B, int H, int W, int C, float* out, cudaStream_t st);
int launch_coord_gate_bwd(const float* x, const float* z, const float* dout, int B, int H, int W, int C, float* dx, float* dz, cudaStream_t st);

// out[i] = sum_s part[s][i], i < n (split-K fold of the weight-gradient partials)
int launch_bwd_fold(const float* part, int S, long long n, float* out, cudaStream_t st);

// split plan of the weight-gradient GEMM dW[Nout][Kin] = dY^T X over M tokens
void bwd_wgrad_splits(long long M, int Nout, int Kin, int* S, int* Ms);

// ---- weight-gradient GEMM with MN-major operands read in place (wgrad_tc.cu) ----
// out[z][i][j] = alpha * sum_t A[z][t][i] * B[z][t][j]  (fp32 / fp16 / bf16 operands, fp32 result), optional db[i] = alpha * sum_t A[t][i],
// optional transposed copy outT[z][j][i], optional head mask (keep i / mask_ch == j / mask_ch).  Ordered split-K: bit-reproducible.
struct WgradArgs {
  const void* A;         // [batch][Mtok][lda], NL channels used
  int fmt;               // element format of BOTH operands: 0 fp16, 1 bf16, 2 fp32 (TF32 MMA).  tcgen05 kind::f16 rejects mixed
                         // A / B formats (illegal instruction on B200)
  const void* B;         // [batch][Mtok][ldb], KL channels used
  long long Mtok;
  int NL, KL, lda, ldb;
  int batch;
  long long strideA, strideB;     // elements between batch items (0 with batch 1)
  float alpha;
  float* out;            // [batch][NL][ldo]
  int ldo;
  long long stride_out;
  float* outT;           // nullable: [batch][KL][ldt]
  int ldt;
  long long stride_outT;
  float* db;             // nullable (batch 1): [NL]
  int mask_ch;           // 0 = no mask
  float* scratch;        // wgrad_tc_scratch_floats(...) floats
};
bool wgrad_tc_eligible(long long Mtok, int NL, int KL, int lda, int ldb, int elem_bytes);
size_t wgrad_tc_scratch_floats(long long Mtok, int NL, int KL, int batch, int elem_bytes);
int launch_wgrad_tc(const WgradArgs& a, cudaStream_t st);

// dst = fp16(src * scale), saturating at +-65504 (flag "grad_bf16" = 0: gradient operands carry a static power-of-two scale that
// the consuming GEMM's epilogue removes; only valid while |dy| * scale stays finite in fp16)
int launch_f32_to_f16_scaled(const float* src, __half* dst, long long n, float scale, cudaStream_t st);
int launch_f16_to_f32(const __half* src, float* dst, long long n, cudaStream_t st);
// up to three tensors -> bf16 (round to nearest even) in ONE launch; src fp32, or fp16 when src_f16
struct CvtSegs {
  const void* src[3];
  void* dst[3];
  long long count[3];
  long long chunks[3];   // filled by the launcher
  int src_f16[3];
  int n;
};
int launch_to_bf16_multi(CvtSegs segs, cudaStream_t st);
// dst = bf16(src), round to nearest even: the default gradient operand format (fp32's exponent range: no scale, no saturation)
int launch_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t st);
