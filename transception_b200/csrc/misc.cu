// Stem, bridge regroup / spatial-reduction packing, IFF pooling+gating and decoder pixel-shuffle kernels.
#include <cuda_fp16.h>
#include "common.cuh"
#include "misc.cuh"

namespace {

// =====================================================================================
// stem: conv 7x7 / stride 4 / pad 3 (Cin 3 -> 64) + bias + LayerNorm(64)  (reference MSTr.py:299-304)
// block = 8x8 output pixels, 256 threads = 64 pixels x 4 channel groups of 16.
// =====================================================================================
constexpr int PE_T = 8, PE_K = 7, PE_S = 4, PE_IN = (PE_T - 1) * PE_S + PE_K;  // 35

// CIN = 3: three input planes.  CIN = 1: the reference repeats the grey plane three times (MSTr.py:2828-2829), which is
// the same as convolving it once with the filters summed over their input channels -> 49 taps instead of 147.
template <int CIN>
__global__ void __launch_bounds__(256) patch_embed_ln_kernel(const float* __restrict__ x, long long xs_b, long long xs_c,
                                                             int Hin, int Win, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ lnw,
                                                             const float* __restrict__ lnb, float eps, int Ho, int Wo,
                                                             float* __restrict__ out) {
  constexpr int NTAP = CIN * 49;
  extern __shared__ float sm[];
  float* ws = sm;                        // [NTAP][64]  (tap-major)
  float* tile = ws + NTAP * 64;          // [CIN][35][36]
  const int tid = threadIdx.x, b = blockIdx.z;
  const int oy0 = blockIdx.y * PE_T, ox0 = blockIdx.x * PE_T;
  pdl_trigger();
  // filters (module parameters): coalesced reads in their native [co][ci][ky][kx] order, transposed in shared memory
  for (int i = tid; i < 64 * NTAP; i += 256) {
    const int co = i / NTAP, tap = i - co * NTAP;
    float v;
    if (CIN == 1) v = __ldg(w + co * 147 + tap) + __ldg(w + co * 147 + 49 + tap) + __ldg(w + co * 147 + 98 + tap);
    else v = __ldg(w + i);
    ws[tap * 64 + co] = v;
  }
  pdl_wait();
  const int iy0 = oy0 * PE_S - 3, ix0 = ox0 * PE_S - 3;
  for (int i = tid; i < CIN * PE_IN * PE_IN; i += 256) {
    const int xx = i % PE_IN, yy = (i / PE_IN) % PE_IN, ci = i / (PE_IN * PE_IN);
    const int gy = iy0 + yy, gx = ix0 + xx;
    float v = 0.f;
    if (gy >= 0 && gy < Hin && gx >= 0 && gx < Win) v = x[(long long)b * xs_b + ci * xs_c + (long long)gy * Win + gx];
    tile[(ci * PE_IN + yy) * 36 + xx] = v;
  }
  __syncthreads();
  const int cg = tid & 3, pix = tid >> 2;
  const int py = pix >> 3, px = pix & 7;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = bias[cg * 16 + i];
#pragma unroll 1
  for (int ci = 0; ci < CIN; ci++)
#pragma unroll 1
    for (int ky = 0; ky < 7; ky++) {
      const float* trow = tile + (ci * PE_IN + py * PE_S + ky) * 36 + px * PE_S;
      const float* wrow = ws + ((ci * 7 + ky) * 7) * 64 + cg * 16;
#pragma unroll
      for (int kx = 0; kx < 7; kx++) {
        const float v = trow[kx];
        const float4* w4 = reinterpret_cast<const float4*>(wrow + kx * 64);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float4 ww = w4[j];
          acc[j * 4 + 0] = fmaf(v, ww.x, acc[j * 4 + 0]);
          acc[j * 4 + 1] = fmaf(v, ww.y, acc[j * 4 + 1]);
          acc[j * 4 + 2] = fmaf(v, ww.z, acc[j * 4 + 2]);
          acc[j * 4 + 3] = fmaf(v, ww.w, acc[j * 4 + 3]);
        }
      }
    }
  // LayerNorm over the 64 channels held by 4 adjacent lanes
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const float mean = s * (1.f / 64.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) { const float d = acc[i] - mean; q += d * d; }
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  q += __shfl_xor_sync(0xffffffffu, q, 2);
  const float rstd = rsqrtf(q * (1.f / 64.f) + eps);
  const int oy = oy0 + py, ox = ox0 + px;
  if (oy < Ho && ox < Wo) {
    float* __restrict__ o = out + (((long long)b * Ho + oy) * Wo + ox) * 64 + cg * 16;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float4 r;
      const int c = cg * 16 + j * 4;
      r.x = (acc[j * 4 + 0] - mean) * rstd * lnw[c + 0] + lnb[c + 0];
      r.y = (acc[j * 4 + 1] - mean) * rstd * lnw[c + 1] + lnb[c + 1];
      r.z = (acc[j * 4 + 2] - mean) * rstd * lnw[c + 2] + lnb[c + 2];
      r.w = (acc[j * 4 + 3] - mean) * rstd * lnw[c + 3] + lnb[c + 3];
      reinterpret_cast<float4*>(o)[j] = r;
    }
  }
}

// =====================================================================================
// bridge: NHWC maps -> [B][Ntok][64] token buffer (a C_k-channel pixel = C_k/64 consecutive tokens)
// =====================================================================================
__global__ void __launch_bounds__(256) regroup_kernel(RegroupArgs a) {
  const long long per_img = (long long)a.ntok * 16;   // float4 per image
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_img * a.B) return;
  const int b = (int)(idx / per_img);
  const long long r = idx % per_img;
  int k = 0;
#pragma unroll
  for (int i = 1; i < 4; i++) if (r >= (long long)a.tok_off[i] * 16) k = i;
  const long long local = r - (long long)a.tok_off[k] * 16;
  const long long n_k = (long long)(a.tok_off[k + 1] - a.tok_off[k]) * 16;
  float4 v = reinterpret_cast<const float4*>(a.src[k])[(long long)b * n_k + local];
  if (a.res) {
    const float4 r4 = reinterpret_cast<const float4*>(a.res)[idx];
    v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
  }
  reinterpret_cast<float4*>(a.dst)[idx] = v;
}

__global__ void __launch_bounds__(256) ungroup_kernel(UngroupArgs a) {
  const long long per_img = (long long)a.ntok * 16;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_img * a.B) return;
  const int b = (int)(idx / per_img);
  const long long r = idx % per_img;
  int k = 0;
#pragma unroll
  for (int i = 1; i < 4; i++) if (r >= (long long)a.tok_off[i] * 16) k = i;
  const long long local = r - (long long)a.tok_off[k] * 16;
  const long long n_k = (long long)(a.tok_off[k + 1] - a.tok_off[k]) * 16;
  reinterpret_cast<float4*>(a.dst[k])[(long long)b * n_k + local] = reinterpret_cast<const float4*>(a.src)[idx];
}

// =====================================================================================
// Scale_reduce (reference MSTr.py:2225-2249)
// im2row of the non-overlapping r x r patches in the conv weight's native (cin, ky, kx) K order
// =====================================================================================
__global__ void __launch_bounds__(256) sr_im2row_kernel(const float* __restrict__ x, long long xs_b, int HW, int Cin, int r,
                                                        int B, float* __restrict__ A) {
  // x: per image [HW][HW][Cin] at x + b*xs_b ; A: [B*P*P][Cin*r*r], P = HW/r
  const int P = HW / r;
  const long long K = (long long)Cin * r * r;
  const long long total = (long long)B * P * P * K;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // read-coalesced ordering: (b, i, j, ky, kx, cin)
  const int cin = (int)(idx % Cin);
  long long t = idx / Cin;
  const int kx = (int)(t % r); t /= r;
  const int ky = (int)(t % r); t /= r;
  const int j = (int)(t % P); t /= P;
  const int i = (int)(t % P);
  const int b = (int)(t / P);
  const float v = x[(long long)b * xs_b + ((long long)(i * r + ky) * HW + (j * r + kx)) * Cin + cin];
  A[((long long)(b * P + i) * P + j) * K + ((long long)cin * r + ky) * r + kx] = v;
}

// the same permutation backwards: dx[b][(i*r+ky)*HW + j*r+kx][cin] = dA[(b,i,j)][(cin,ky,kx)]  (coalesced on the dx side)
__global__ void __launch_bounds__(256) sr_row2im_kernel(const float* __restrict__ dA, long long xs_b, int HW, int Cin, int r,
                                                        int B, float* __restrict__ dx) {
  const int P = HW / r;
  const long long K = (long long)Cin * r * r;
  const long long total = (long long)B * P * P * K;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cin = (int)(idx % Cin);
  long long t = idx / Cin;
  const int kx = (int)(t % r); t /= r;
  const int ky = (int)(t % r); t /= r;
  const int j = (int)(t % P); t /= P;
  const int i = (int)(t % P);
  const int b = (int)(t / P);
  dx[(long long)b * xs_b + ((long long)(i * r + ky) * HW + (j * r + kx)) * Cin + cin] =
      dA[((long long)(b * P + i) * P + j) * K + ((long long)cin * r + ky) * r + kx];
}

// pack conv outputs + raw stage-4 tokens into the reduced sequence and LayerNorm(64) it.
// reduced token t of image b: scale k, t_local = (c % g)*49 + s, feature f = c // g  
__global__ void __launch_bounds__(256) sr_pack_ln_kernel(SrPackArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)a.B * a.nred) return;
  const int b = (int)(row / a.nred), t = (int)(row % a.nred);
  float v0, v1;
  if (t >= a.red_off[3]) {
    const float* src = a.x + (long long)b * a.xs_b + (long long)(a.raw_tok0 + t - a.red_off[3]) * 64;
    v0 = src[lane]; v1 = src[lane + 32];
  } else {
    int k = 0;
    if (t >= a.red_off[1]) k = 1;
    if (t >= a.red_off[2]) k = 2;
    const int g = a.gmul[k], pp = a.pp[k];
    const int tl = t - a.red_off[k];
    const int cm = tl / pp, s = tl % pp;
    const float* src = a.conv[k] + ((long long)b * pp + s) * (64 * g) + cm;
    v0 = src[lane * g]; v1 = src[(lane + 32) * g];
  }
  float* dst = a.out + row * 64;
  if (!a.lnw) {            // training row: the LayerNorm is its own autograd node
    dst[lane] = v0; dst[lane + 32] = v1;
    return;
  }
  const float mean = warp_sum(v0 + v1) * (1.f / 64.f);
  const float d0 = v0 - mean, d1 = v1 - mean;
  const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + a.eps);
  dst[lane] = d0 * rstd * a.lnw[lane] + a.lnb[lane];
  dst[lane + 32] = d1 * rstd * a.lnw[lane + 32] + a.lnb[lane + 32];
}

// gradient of the packing above (no LayerNorm): one warp per reduced row scatters it to its conv output / raw token row
__global__ void __launch_bounds__(256) sr_unpack_kernel(SrUnpackArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)a.B * a.nred) return;
  const int b = (int)(row / a.nred), t = (int)(row % a.nred);
  const float* src = a.dred + row * 64;
  const float v0 = src[lane], v1 = src[lane + 32];
  if (t >= a.red_off[3]) {
    float* dst = a.dx + (long long)b * a.xs_b + (long long)(a.raw_tok0 + t - a.red_off[3]) * 64;
    dst[lane] = v0; dst[lane + 32] = v1;
  } else {
    int k = 0;
    if (t >= a.red_off[1]) k = 1;
    if (t >= a.red_off[2]) k = 2;
    const int g = a.gmul[k], pp = a.pp[k];
    const int tl = t - a.red_off[k];
    const int cm = tl / pp, s = tl % pp;
    float* dst = a.dconv[k] + ((long long)b * pp + s) * (64 * g) + cm;
    dst[lane * g] = v0; dst[(lane + 32) * g] = v1;
  }
}

// ---- fp16 forms -------------------------------------------------------------------------------------------------
// im2row with K ordered (ky, kx, cin): every (patch, ky) piece is r*Cin contiguous halfs of the NHWC slab -> pure
// 16-byte vector copies (the prepared conv weight is permuted to the same K order, tcx_prepare_conv_weight_f16).
__global__ void __launch_bounds__(256) sr_im2row16_kernel(const __half* __restrict__ x, long long xs_b, int HW, int Cin, int r,
                                                          int B, __half* __restrict__ A) {
  pdl_trigger();
  pdl_wait();
  const int P = HW / r;
  const int seg = r * Cin / 8;                       // vectors per (patch, ky) piece
  const long long total = (long long)B * P * P * r * seg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % seg);
  long long t = idx / seg;
  const int ky = (int)(t % r); t /= r;
  const int j = (int)(t % P); t /= P;
  const int i = (int)(t % P);
  const int b = (int)(t / P);
  const uint4 val = *reinterpret_cast<const uint4*>(x + (long long)b * xs_b + ((long long)(i * r + ky) * HW + j * r) * Cin + v * 8);
  *reinterpret_cast<uint4*>(A + (((long long)(b * P + i) * P + j) * r + ky) * (long long)(r * Cin) + v * 8) = val;
}

// [N][Cin][r][r] fp32 conv weight -> [N][(ky, kx, cin)] fp16
__global__ void __launch_bounds__(256) conv_weight_perm16_kernel(const float* __restrict__ w, __half* __restrict__ o, int N, int Cin,
                                                                 int r) {
  const long long total = (long long)N * Cin * r * r;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cin = (int)(idx % Cin);
  long long t = idx / Cin;
  const int kx = (int)(t % r); t /= r;
  const int ky = (int)(t % r);
  const int n = (int)(t / r);
  o[idx] = __float2half_rn(w[(((long long)n * Cin + cin) * r + ky) * r + kx]);
}

// pack conv outputs (fp32) + raw stage-4 tokens (fp16) into the reduced sequence, LayerNorm(64), write fp16
__global__ void __launch_bounds__(256) sr_pack_ln16_kernel(SrPackArgs a, const __half* __restrict__ x16, __half* __restrict__ out16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)a.B * a.nred) return;
  const int b = (int)(row / a.nred), t = (int)(row % a.nred);
  float v0, v1;
  if (t >= a.red_off[3]) {
    const __half* src = x16 + (long long)b * a.xs_b + (long long)(a.raw_tok0 + t - a.red_off[3]) * 64;
    v0 = __half2float(src[lane]); v1 = __half2float(src[lane + 32]);
  } else {
    int k = 0;
    if (t >= a.red_off[1]) k = 1;
    if (t >= a.red_off[2]) k = 2;
    const int g = a.gmul[k], pp = a.pp[k];
    const int tl = t - a.red_off[k];
    const int cm = tl / pp, s = tl % pp;
    const float* src = a.conv[k] + ((long long)b * pp + s) * (64 * g) + cm;
    v0 = src[lane * g]; v1 = src[(lane + 32) * g];
  }
  const float mean = warp_sum(v0 + v1) * (1.f / 64.f);
  const float d0 = v0 - mean, d1 = v1 - mean;
  const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + a.eps);
  __half* dst = out16 + row * 64;
  dst[lane] = __float2half_rn(d0 * rstd * a.lnw[lane] + a.lnb[lane]);
  dst[lane + 32] = __float2half_rn(d1 * rstd * a.lnw[lane + 32] + a.lnb[lane + 32]);
}

// =====================================================================================
// IFF / coordinate attention (reference MSTr.py:1322-1348)
// =====================================================================================
// pooled[b][h][k] = mean_w x, pooled[b][H+w][k] = mean_h x  for the 4 concatenated sources.
// grid (HW, 4, 2B), one thread per channel (coalesced): blockIdx.z < B -> the row mean of map row blockIdx.x,
// otherwise the column mean of map column blockIdx.x.  Fixed summation order: deterministic, no atomics.
__global__ void __launch_bounds__(512) iff_pool_kernel(IffSrc src, int B, int HW, int C, float* __restrict__ pooled) {
  const int i = blockIdx.x, s = blockIdx.y;
  const bool col = (int)blockIdx.z >= B;
  const int b = col ? blockIdx.z - B : blockIdx.z;
  const float inv = 1.f / (float)HW;
  const long long step = col ? (long long)HW * C : C;                       // walk down a column / along a row
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* __restrict__ x = src.p[s] + (long long)b * HW * HW * C + (col ? (long long)i * C : (long long)i * HW * C) + c;
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
    for (; j + 1 < HW; j += 2) { a0 += x[j * step]; a1 += x[(j + 1) * step]; }
    if (j < HW) a0 += x[j * step];
    pooled[(long long)b * 2 * HW * 4 * C + (long long)((col ? HW : 0) + i) * 4 * C + s * C + c] = (a0 + a1) * inv;
  }
}

// gated[b,h,w,k] = x_src(k)[b,h,w,k%C] * a_w[b,w,k] * a_h[b,h,k]
__global__ void __launch_bounds__(256) iff_gate_kernel(IffSrc src, int B, int H, int W, int C, const float* __restrict__ ah,
                                                       const float* __restrict__ aw, float* __restrict__ out) {
  const int c4n = C >> 2;
  const long long total = (long long)B * H * W * 4 * c4n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int kv = (int)(idx % (4 * c4n));
  long long pix = idx / (4 * c4n);
  const int w = (int)(pix % W);
  const int h = (int)((pix / W) % H);
  const int b = (int)(pix / ((long long)W * H));
  const int s = kv / c4n, c = (kv % c4n) * 4;
  const float4 xv = *reinterpret_cast<const float4*>(src.p[s] + (((long long)b * H + h) * W + w) * C + c);
  const float4 a1 = *reinterpret_cast<const float4*>(ah + ((long long)b * H + h) * 4 * C + s * C + c);
  const float4 a2 = *reinterpret_cast<const float4*>(aw + ((long long)b * W + w) * 4 * C + s * C + c);
  float4 o;
  o.x = xv.x * a2.x * a1.x; o.y = xv.y * a2.y * a1.y; o.z = xv.z * a2.z * a1.z; o.w = xv.w * a2.w * a1.w;
  *reinterpret_cast<float4*>(out + pix * 4 * C + s * C + c) = o;
}

// =====================================================================================
// decoder: pixel shuffle + LayerNorm (PatchExpand MSTr.py:184-201); and the fused last stage
// (FinalPatchExpand_X4 :212-227 + 1x1 conv to classes :281), logits written NCHW.
// =====================================================================================
// in: [B*H*W][s*s*c]; out token (b, h*s+p1, w*s+p2) <- columns (p1*s+p2)*c .. +c ; one warp per out token
__global__ void __launch_bounds__(256) shuffle_ln_kernel(const float* __restrict__ in, int B, int H, int W, int s, int c,
                                                         const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                         float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int Ho = H * s, Wo = W * s;
  if (row >= (long long)B * Ho * Wo) return;
  const int wo = (int)(row % Wo), ho = (int)((row / Wo) % Ho), b = (int)(row / ((long long)Wo * Ho));
  const int h = ho / s, p1 = ho % s, w = wo / s, p2 = wo % s;
  const float* __restrict__ src = in + (((long long)b * H + h) * W + w) * (s * s * c) + (p1 * s + p2) * c;
  float v[8];  // c <= 256
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int idx = lane + i * 32;
    v[i] = idx < c ? src[idx] : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / (float)c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) if (lane + i * 32 < c) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)c + eps);
  float* __restrict__ dst = out + row * c;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int idx = lane + i * 32;
    if (idx < c) dst[idx] = (v[i] - mean) * rstd * lnw[idx] + lnb[idx];
  }
}

// in: [B*H*W][16*64] expand output; one thread per output pixel: LN(64) then ncls dot products, NCHW logits.
template <int MAXCLS>
__global__ void __launch_bounds__(128) final_head_kernel(const float* __restrict__ in, int B, int H, int W,
                                                         const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                         float eps, const float* __restrict__ cw, const float* __restrict__ cb,
                                                         int ncls, float* __restrict__ out) {
  __shared__ float wsm[MAXCLS * 64 + MAXCLS + 128];
  float* wl = wsm;                       // [ncls][64]  with LN weight folded in
  float* bl = wsm + MAXCLS * 64;         // [ncls]      bias + sum(lnb*cw)
  for (int i = threadIdx.x; i < ncls * 64; i += blockDim.x) wl[i] = cw[i] * lnw[i & 63];
  for (int k = threadIdx.x; k < ncls; k += blockDim.x) {
    float s = cb[k];
    for (int d = 0; d < 64; d++) s = fmaf(cw[k * 64 + d], lnb[d], s);
    bl[k] = s;
  }
  __syncthreads();
  const int Ho = H * 4, Wo = W * 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  const int wo = (int)(idx % Wo), ho = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  const int h = ho >> 2, p1 = ho & 3, w = wo >> 2, p2 = wo & 3;
  const float4* __restrict__ src =
      reinterpret_cast<const float4*>(in + (((long long)b * H + h) * W + w) * 1024 + (p1 * 4 + p2) * 64);
  float v[64];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const float4 t = src[i];
    v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    s += (t.x + t.y) + (t.z + t.w);
  }
  const float mean = s * (1.f / 64.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 64; i++) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
  const float rstd = rsqrtf(q * (1.f / 64.f) + eps);
  for (int k = 0; k < ncls; k++) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i++) a = fmaf(v[i], wl[k * 64 + i], a);
    out[(((long long)b * ncls + k) * Ho + ho) * Wo + wo] = fmaf(a, rstd, bl[k]);
  }
}

}  // namespace

int launch_patch_embed_ln(const float* x, long long xs_b, long long xs_c, int B, int Hin, int Win, const float* w,
                          const float* bias, const float* lnw, const float* lnb, float eps, float* out, cudaStream_t st) {
  const int Ho = (Hin + 6 - 7) / 4 + 1, Wo = (Win + 6 - 7) / 4 + 1;
  const bool grey = xs_c == 0;           // one plane read three times
  const int cin = grey ? 1 : 3;
  const size_t smem = (size_t)(cin * 49 * 64 + cin * PE_IN * 36) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first())
    cudaFuncSetAttribute(patch_embed_ln_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((147 * 64 + 3 * PE_IN * 36) * sizeof(float)));
  dim3 grid(cdiv(Wo, PE_T), cdiv(Ho, PE_T), B);
  if (grey) tcx_launch_pdl(patch_embed_ln_kernel<1>, grid, dim3(256), smem, st, x, xs_b, xs_c, Hin, Win, w, bias, lnw, lnb, eps, Ho, Wo, out);
  else tcx_launch_pdl(patch_embed_ln_kernel<3>, grid, dim3(256), smem, st, x, xs_b, xs_c, Hin, Win, w, bias, lnw, lnb, eps, Ho, Wo, out);
  return tcx_check_launch("patch_embed_ln");
}

int launch_regroup(const RegroupArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.ntok * 16;
  regroup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  return tcx_check_launch("bridge_regroup");
}

int launch_ungroup(const UngroupArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.ntok * 16;
  ungroup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  return tcx_check_launch("bridge_ungroup");
}

int launch_sr_unpack(const SrUnpackArgs& a, cudaStream_t st) {
  const long long rows = (long long)a.B * a.nred;
  sr_unpack_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a);
  return tcx_check_launch("sr_unpack");
}

int launch_sr_row2im(const float* dA, long long xs_b, int HW, int Cin, int r, int B, float* dx, cudaStream_t st) {
  const int P = HW / r;
  const long long total = (long long)B * P * P * Cin * r * r;
  sr_row2im_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dA, xs_b, HW, Cin, r, B, dx);
  return tcx_check_launch("sr_row2im");
}

int launch_sr_im2row(const float* x, long long xs_b, int HW, int Cin, int r, int B, float* A, cudaStream_t st) {
  const int P = HW / r;
  const long long total = (long long)B * P * P * Cin * r * r;
  sr_im2row_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, xs_b, HW, Cin, r, B, A);
  return tcx_check_launch("sr_im2row");
}

int launch_sr_pack_ln(const SrPackArgs& a, cudaStream_t st) {
  const long long rows = (long long)a.B * a.nred;
  sr_pack_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a);
  return tcx_check_launch("sr_pack_ln");
}

int launch_iff_pool(const IffSrc& src, int B, int HW, int C, float* pooled, cudaStream_t st) {
  dim3 grid(HW, 4, 2 * B);
  int threads = (C + 31) / 32 * 32;
  if (threads > 512) threads = 512;
  iff_pool_kernel<<<grid, threads, 0, st>>>(src, B, HW, C, pooled);
  return tcx_check_launch("iff_pool");
}

int launch_iff_gate(const IffSrc& src, int B, int H, int W, int C, const float* ah, const float* aw, float* out,
                    cudaStream_t st) {
  const long long total = (long long)B * H * W * C;
  iff_gate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, B, H, W, C, ah, aw, out);
  return tcx_check_launch("iff_gate");
}

int launch_shuffle_ln(const float* in, int B, int H, int W, int s, int c, const float* lnw, const float* lnb, float eps,
                      float* out, cudaStream_t st) {
  TCX_REQUIRE(c <= 256, "patch_expand: c=%d > 256", c);
  const long long rows = (long long)B * H * s * W * s;
  shuffle_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, B, H, W, s, c, lnw, lnb, eps, out);
  return tcx_check_launch("shuffle_ln");
}

int launch_final_head(const float* in, int B, int H, int W, const float* lnw, const float* lnb, float eps,
                      const float* cw, const float* cb, int ncls, float* out, cudaStream_t st) {
  TCX_REQUIRE(ncls >= 1 && ncls <= 32, "final_head: ncls=%d out of range (1..32)", ncls);
  const long long total = (long long)B * H * 4 * W * 4;
  final_head_kernel<32><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(in, B, H, W, lnw, lnb, eps, cw, cb, ncls, out);
  return tcx_check_launch("final_head");
}

int launch_sr_im2row16(const void* x16, long long xs_b, int HW, int Cin, int r, int B, void* A16, cudaStream_t st) {
  TCX_REQUIRE(HW % r == 0 && (r * Cin) % 8 == 0, "sr_im2row16: bad geometry HW=%d r=%d Cin=%d", HW, r, Cin);
  const int P = HW / r;
  const long long total = (long long)B * P * P * r * (r * Cin / 8);
  if (total == 0) return 0;
  tcx_launch_pdl(sr_im2row16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st,
                 reinterpret_cast<const __half*>(x16), xs_b, HW, Cin, r, B, reinterpret_cast<__half*>(A16));
  return tcx_check_launch("sr_im2row16");
}

int launch_conv_weight_perm16(const float* w, void* o16, int N, int Cin, int r, cudaStream_t st) {
  const long long total = (long long)N * Cin * r * r;
  if (total == 0) return 0;
  conv_weight_perm16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, reinterpret_cast<__half*>(o16), N, Cin, r);
  return tcx_check_launch("conv_weight_perm16");
}

int launch_sr_pack_ln16(const SrPackArgs& a, const void* x16, void* out16, cudaStream_t st) {
  const long long rows = (long long)a.B * a.nred;
  if (rows == 0) return 0;
  tcx_launch_pdl(sr_pack_ln16_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, a,
                 reinterpret_cast<const __half*>(x16), reinterpret_cast<__half*>(out16));
  return tcx_check_launch("sr_pack_ln16");
}
