// Fused segmentation loss of the reference training step (trainer.py:141-143, utils.py:11-47):
//   loss = w_ce * CrossEntropy(logits, label) + w_dice * DiceLoss(softmax(logits), label)
// forward (loss, both terms, class-wise Dice) and backward (d loss / d logits) without host synchronisation: the reference
// does a Python loop over classes with one `.item()` per class per step (utils.py:45).
//   pass 1  seg_loss_partial : per pixel softmax over the K class planes (NCHW, coalesced along HW), accumulate
//                              sum(-log p_label), I_c = sum p_c t_c, Z_c = sum p_c^2, Y_c = sum t_c per block (fixed order)
//   pass 2  seg_loss_final   : one block folds the block partials in index order -> statistics + outputs (deterministic)
//   pass 3  seg_loss_grad    : per pixel, recompute softmax, chain rule through Dice and softmax using the statistics
// HBM-bound: logits are read once per pass (K*4 bytes per pixel), the gradient written once.
#include "common.cuh"
#include "loss.cuh"

namespace {

constexpr int KMAX = SEG_LOSS_KMAX;
constexpr int NV = 2 + 3 * KMAX;          // ce, bad labels, I[K], Z[K], Y[K]
constexpr float SMOOTH = 1e-5f;            // utils.py:26

__device__ __forceinline__ int load_label(const void* p, int kind, long long i) {
  switch (kind) {
    case 0: return (int)reinterpret_cast<const long long*>(p)[i];
    case 1: { const float f = reinterpret_cast<const float*>(p)[i]; const int v = (int)f; return f == (float)v ? v : -1; }
    case 2: return reinterpret_cast<const int*>(p)[i];
    default: return (int)reinterpret_cast<const unsigned char*>(p)[i];
  }
}

// p[c] = softmax over classes (or the raw inputs when `softmax` is 0); returns log-sum-exp (0 without softmax)
__device__ __forceinline__ float pixel_probs(const float* __restrict__ base, int K, long long HW, int softmax, float (&p)[KMAX], float& zmax) {
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < KMAX; c++) {
    p[c] = c < K ? base[c * HW] : -INFINITY;
    m = fmaxf(m, p[c]);
  }
  zmax = m;
  if (!softmax) {
#pragma unroll
    for (int c = 0; c < KMAX; c++) if (c >= K) p[c] = 0.f;
    return 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < KMAX; c++) {
    p[c] = c < K ? __expf(p[c] - m) : 0.f;
    s += p[c];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < KMAX; c++) p[c] *= inv;
  return __logf(s);
}

__global__ void __launch_bounds__(256) seg_loss_partial_kernel(SegLossArgs a, float* __restrict__ partial) {
  PDL_TOP();
  __shared__ float red[8][NV];
  const long long npix = (long long)a.B * a.HW;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) acc[i] = 0.f;
  for (long long n = (long long)blockIdx.x * 256 + threadIdx.x; n < npix; n += (long long)gridDim.x * 256) {
    const long long b = n / a.HW, hw = n - b * a.HW;
    const float* base = a.logits + b * a.K * a.HW + hw;
    float p[KMAX], zmax;
    const float lse = pixel_probs(base, a.K, a.HW, a.softmax, p, zmax);
    const int lab = load_label(a.labels, a.kind, n);
    const bool ok = lab >= 0 && lab < a.K;
    if (!ok) acc[1] += 1.f;
#pragma unroll
    for (int c = 0; c < KMAX; c++) {
      if (c < a.K) {
        const bool hit = ok && c == lab;
        acc[2 + KMAX + c] = fmaf(p[c], p[c], acc[2 + KMAX + c]);
        if (hit) {
          acc[2 + c] += p[c];
          acc[2 + 2 * KMAX + c] += 1.f;
          if (a.softmax) acc[0] += lse + zmax - base[c * a.HW];      // -log softmax[label]
        }
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
    partial[(long long)blockIdx.x * NV + threadIdx.x] = v;
  }
}

// stats layout (floats): [0] ce sum, [1] bad labels, [2..) I, Z, Y (KMAX each); out: loss, ce, dice, bad, class-wise dice[K]
__global__ void __launch_bounds__(64) seg_loss_final_kernel(SegLossArgs a, const float* __restrict__ partial, int nblk, float* __restrict__ stats,
                                                            float* __restrict__ out) {
  PDL_TOP();
  __shared__ float tot[NV];
  if (threadIdx.x < NV) {
    float v = 0.f;
    for (int b = 0; b < nblk; b++) v += partial[(long long)b * NV + threadIdx.x];     // index order: deterministic
    tot[threadIdx.x] = v;
    stats[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float npix = (float)((long long)a.B * a.HW);
    const float ce = a.softmax ? tot[0] / npix : 0.f;
    float dice = 0.f;
    for (int c = 0; c < a.K; c++) {
      const float I = tot[2 + c], Z = tot[2 + KMAX + c], Y = tot[2 + 2 * KMAX + c];
      const float d = 1.f - (2.f * I + SMOOTH) / (Z + Y + SMOOTH);       // utils.py:24-31
      out[4 + c] = 1.f - d;                                               // class_wise_dice, utils.py:45
      dice += d * a.cw[c];
    }
    dice /= (float)a.K;
    out[0] = a.w_ce * ce + a.w_dice * dice;
    out[1] = ce;
    out[2] = dice;
    out[3] = tot[1];
  }
}

__global__ void __launch_bounds__(256) seg_loss_grad_kernel(SegLossArgs a, const float* __restrict__ stats, const float* __restrict__ grad_out,
                                                            float* __restrict__ dlogits) {
  PDL_TOP();
  __shared__ float ca[KMAX], cb[KMAX];     // d dice / d p_c = -ca[c] * t_c + cb[c] * p_c
  if (threadIdx.x < KMAX) {
    const int c = threadIdx.x;
    float A = 0.f, Bc = 0.f;
    if (c < a.K) {
      const float I = stats[2 + c], Z = stats[2 + KMAX + c], Y = stats[2 + 2 * KMAX + c];
      const float D = Z + Y + SMOOTH, s = a.w_dice * a.cw[c] / (float)a.K;
      A = s * 2.f / D;
      Bc = s * 2.f * (2.f * I + SMOOTH) / (D * D);
    }
    ca[c] = A;
    cb[c] = Bc;
  }
  __syncthreads();
  const long long npix = (long long)a.B * a.HW;
  const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
  if (n >= npix) return;
  const float go = grad_out ? __ldg(grad_out) : 1.f;
  const long long b = n / a.HW, hw = n - b * a.HW;
  const float* base = a.logits + b * a.K * a.HW + hw;
  float* dst = dlogits + b * a.K * a.HW + hw;
  float p[KMAX], zmax;
  pixel_probs(base, a.K, a.HW, a.softmax, p, zmax);
  const int lab = load_label(a.labels, a.kind, n);
  const bool ok = lab >= 0 && lab < a.K;
  float g[KMAX], dot = 0.f;
#pragma unroll
  for (int c = 0; c < KMAX; c++) {
    const float t = (ok && c == lab) ? 1.f : 0.f;
    g[c] = c < a.K ? fmaf(cb[c], p[c], -ca[c] * t) : 0.f;
    dot = fmaf(g[c], p[c], dot);
  }
  const float wce = a.w_ce / (float)npix;
#pragma unroll
  for (int c = 0; c < KMAX; c++) {
    if (c < a.K) {
      const float t = (ok && c == lab) ? 1.f : 0.f;
      const float d = a.softmax ? fmaf(p[c], g[c] - dot, (ok ? wce : 0.f) * (p[c] - t)) : g[c];
      dst[c * a.HW] = d * go;
    }
  }
}

// argmax over the K class planes -> uint8 label map (utils.py:86 `torch.argmax(torch.softmax(outputs, dim=1), dim=1)`: softmax is
// monotone, so the arg max is taken on the logits; first index wins ties like torch.argmax).  9x fewer bytes cross PCIe than logits.
__global__ void __launch_bounds__(256) argmax_classes_kernel(const float* __restrict__ logits, unsigned char* __restrict__ out, int B, int K,
                                                             long long HW) {
  PDL_TOP();
  const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
  if (n >= (long long)B * HW) return;
  const long long b = n / HW, hw = n - b * HW;
  const float* base = logits + b * K * HW + hw;
  float best = base[0];
  int arg = 0;
  for (int c = 1; c < K; c++) {
    const float v = base[c * HW];
    if (v > best) { best = v; arg = c; }
  }
  out[n] = (unsigned char)arg;
}

int grid_for(long long npix) {
  long long g = (npix + 255) / 256;
  return (int)(g < 592 ? (g > 0 ? g : 1) : 592);     // 148 SMs x 4 resident blocks
}

}  // namespace

size_t seg_loss_workspace_floats(int B, int K, long long HW) {
  (void)K;
  return (size_t)(grid_for((long long)B * HW) + 1) * NV + 64;
}

int launch_seg_loss_fwd(const SegLossArgs& a, float* out, float* ws, cudaStream_t st) {
  TCX_REQUIRE(a.K >= 1 && a.K <= KMAX, "seg_loss: 1 <= classes <= %d (got %d)", KMAX, a.K);
  TCX_REQUIRE(a.kind >= 0 && a.kind <= 3, "seg_loss: label kind must be 0 (int64), 1 (float32), 2 (int32) or 3 (uint8)");
  TCX_REQUIRE(a.softmax || a.w_ce == 0.f, "seg_loss: the cross-entropy term needs logits (softmax = 1)");
  const long long npix = (long long)a.B * a.HW;
  TCX_REQUIRE(npix > 0, "seg_loss: empty batch");
  const int nblk = grid_for(npix);
  float* stats = ws;
  float* partial = ws + NV;
  {
    ProfScope prof("seg_loss_partial", st, (double)npix * (a.K * 4.0 + 4.0));
    tcx_launch_chain(seg_loss_partial_kernel, dim3(nblk), dim3(256), 0, st, a, partial);
    TCX_TRY(tcx_check_launch("seg_loss_partial"));
  }
  tcx_launch_chain(seg_loss_final_kernel, dim3(1), dim3(64), 0, st, a, partial, nblk, stats, out);
  return tcx_check_launch("seg_loss_final");
}

int launch_seg_loss_bwd(const SegLossArgs& a, const float* ws, const float* grad_out, float* dlogits, cudaStream_t st) {
  TCX_REQUIRE(a.K >= 1 && a.K <= KMAX, "seg_loss: 1 <= classes <= %d (got %d)", KMAX, a.K);
  const long long npix = (long long)a.B * a.HW;
  if (npix == 0) return 0;
  ProfScope prof("seg_loss_grad", st, (double)npix * (a.K * 8.0 + 4.0));
  tcx_launch_chain(seg_loss_grad_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, st, a, ws, grad_out, dlogits);
  return tcx_check_launch("seg_loss_grad");
}

int launch_argmax_classes(const float* logits, unsigned char* out, int B, int K, long long HW, cudaStream_t st) {
  TCX_REQUIRE(K >= 1 && K <= 256, "argmax_classes: 1 <= classes <= 256");
  const long long npix = (long long)B * HW;
  if (npix == 0) return 0;
  ProfScope prof("argmax_classes", st, (double)npix * (K * 4.0 + 1.0));
  tcx_launch_chain(argmax_classes_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, st, logits, out, B, K, HW);
  return tcx_check_launch("argmax_classes");
}
