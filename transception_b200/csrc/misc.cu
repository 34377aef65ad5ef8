// Stem, bridge regroup / spatial-reduction packing, IFF pooling+gating and decoder pixel-shuffle kernels.
#include <cuda_fp16.h>
#include "common.cuh"
#include "misc.cuh"
#include "bwd.cuh"

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
    if (!lnw) {        // training row: the conv output itself (the LayerNorm is its own autograd node)
#pragma unroll
      for (int j = 0; j < 4; j++)
        reinterpret_cast<float4*>(o)[j] = make_float4(acc[j * 4], acc[j * 4 + 1], acc[j * 4 + 2], acc[j * 4 + 3]);
      return;
    }
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

// patches of the stem conv for its weight gradient: A[(b, oy, ox)][(ci, ky, kx)], row pitch Kp >= 147 (zero beyond 147 and
// outside the image); xs_c = 0: the grey plane stands for all three input channels (MSTr.py:2828-2829)
__global__ void __launch_bounds__(256) patch_im2row_kernel(const float* __restrict__ x, long long xs_b, long long xs_c, int Hin, int Win,
                                                           int Ho, int Wo, int Kp, long long total, float* __restrict__ A) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % Kp);
  const long long r = idx / Kp;
  const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
  float v = 0.f;
  if (k < 147) {
    const int ci = k / 49, t = k - ci * 49, ky = t / 7, kx = t - ky * 7;
    const int gy = oy * PE_S - 3 + ky, gx = ox * PE_S - 3 + kx;
    if (gy >= 0 && gy < Hin && gx >= 0 && gx < Win) v = x[(long long)b * xs_b + ci * xs_c + (long long)gy * Win + gx];
  }
  A[idx] = v;
}

// =====================================================================================
// bridge: NHWC maps -> [B][Ntok][64] token buffer (a C_k-channel pixel = C_k/64 consecutive tokens)
// =====================================================================================
// four float4 per thread (independent loads in flight: these are pure 25 MB copies at bs16)
constexpr int RG_UNROLL = 4;
__device__ __forceinline__ void regroup_locate(const int* tok_off, long long per_img, long long idx, int& b, int& k, long long& slab_idx) {
  b = (int)(idx / per_img);
  const long long r = idx - (long long)b * per_img;
  k = 0;
#pragma unroll
  for (int i = 1; i < 4; i++) if (r >= (long long)tok_off[i] * 16) k = i;
  const long long local = r - (long long)tok_off[k] * 16;
  const long long n_k = (long long)(tok_off[k + 1] - tok_off[k]) * 16;
  slab_idx = (long long)b * n_k + local;
}
__global__ void __launch_bounds__(256) regroup_kernel(RegroupArgs a) {
  PDL_TOP();
  const long long per_img = (long long)a.ntok * 16;   // float4 per image
  const long long total = per_img * a.B;
  const long long base = ((long long)blockIdx.x * RG_UNROLL) * 256 + threadIdx.x;
  float4 v[RG_UNROLL], r4[RG_UNROLL];
#pragma unroll
  for (int u = 0; u < RG_UNROLL; u++) {
    const long long idx = base + (long long)u * 256;
    if (idx >= total) continue;
    int b, k; long long si;
    regroup_locate(a.tok_off, per_img, idx, b, k, si);
    v[u] = reinterpret_cast<const float4*>(a.src[k])[si];
    if (a.res) r4[u] = reinterpret_cast<const float4*>(a.res)[idx];
  }
#pragma unroll
  for (int u = 0; u < RG_UNROLL; u++) {
    const long long idx = base + (long long)u * 256;
    if (idx >= total) continue;
    float4 o = v[u];
    if (a.res) { o.x += r4[u].x; o.y += r4[u].y; o.z += r4[u].z; o.w += r4[u].w; }
    reinterpret_cast<float4*>(a.dst)[idx] = o;
  }
}

__global__ void __launch_bounds__(256) ungroup_kernel(UngroupArgs a) {
  PDL_TOP();
  const long long per_img = (long long)a.ntok * 16;
  const long long total = per_img * a.B;
  const long long base = ((long long)blockIdx.x * RG_UNROLL) * 256 + threadIdx.x;
  float4 v[RG_UNROLL];
#pragma unroll
  for (int u = 0; u < RG_UNROLL; u++) {
    const long long idx = base + (long long)u * 256;
    if (idx < total) v[u] = reinterpret_cast<const float4*>(a.src)[idx];
  }
#pragma unroll
  for (int u = 0; u < RG_UNROLL; u++) {
    const long long idx = base + (long long)u * 256;
    if (idx >= total) continue;
    int b, k; long long si;
    regroup_locate(a.tok_off, per_img, idx, b, k, si);
    reinterpret_cast<float4*>(a.dst[k])[si] = v[u];
  }
}

// =====================================================================================
// Scale_reduce (reference MSTr.py:2225-2249)
// im2row of the non-overlapping r x r patches in the conv weight's native (cin, ky, kx) K order
// =====================================================================================
// one thread = four consecutive input channels of one patch position: a 16-byte access on the NHWC side, four scalar accesses on the
// patch-matrix side (whose K order (cin, ky, kx) is the conv weight's own); Cin % 4 == 0
__global__ void __launch_bounds__(256) sr_im2row_kernel(const float* __restrict__ x, long long xs_b, int HW, int Cin, int r,
                                                        int B, float* __restrict__ A) {
  PDL_TOP();
  // x: per image [HW][HW][Cin] at x + b*xs_b ; A: [B*P*P][Cin*r*r], P = HW/r
  const int P = HW / r, C4 = Cin >> 2, rr = r * r;
  const long long K = (long long)Cin * rr;
  const long long total = (long long)B * P * P * rr * C4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // read-coalesced ordering: (b, i, j, ky, kx, cin / 4)
  const int cin = (int)(idx % C4) * 4;
  long long t = idx / C4;
  const int kx = (int)(t % r); t /= r;
  const int ky = (int)(t % r); t /= r;
  const int j = (int)(t % P); t /= P;
  const int i = (int)(t % P);
  const int b = (int)(t / P);
  const float4 v = *reinterpret_cast<const float4*>(x + (long long)b * xs_b + ((long long)(i * r + ky) * HW + (j * r + kx)) * Cin + cin);
  float* dst = A + ((long long)(b * P + i) * P + j) * K + ((long long)cin * r + ky) * r + kx;
  dst[0] = v.x; dst[rr] = v.y; dst[2 * rr] = v.z; dst[3 * rr] = v.w;
}

// the same permutation backwards: dx[b][(i*r+ky)*HW + j*r+kx][cin] = dA[(b,i,j)][(cin,ky,kx)]  (coalesced on the dx side)
__global__ void __launch_bounds__(256) sr_row2im_kernel(const float* __restrict__ dA, long long xs_b, int HW, int Cin, int r,
                                                        int B, float* __restrict__ dx) {
  PDL_TOP();
  const int P = HW / r, C4 = Cin >> 2, rr = r * r;
  const long long K = (long long)Cin * rr;
  const long long total = (long long)B * P * P * rr * C4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cin = (int)(idx % C4) * 4;
  long long t = idx / C4;
  const int kx = (int)(t % r); t /= r;
  const int ky = (int)(t % r); t /= r;
  const int j = (int)(t % P); t /= P;
  const int i = (int)(t % P);
  const int b = (int)(t / P);
  const float* src = dA + ((long long)(b * P + i) * P + j) * K + ((long long)cin * r + ky) * r + kx;
  *reinterpret_cast<float4*>(dx + (long long)b * xs_b + ((long long)(i * r + ky) * HW + (j * r + kx)) * Cin + cin) =
      make_float4(src[0], src[rr], src[2 * rr], src[3 * rr]);
}

// pack conv outputs + raw stage-4 tokens into the reduced sequence and LayerNorm(64) it.
// reduced token t of image b: scale k, t_local = (c % g)*49 + s, feature f = c // g  
__global__ void __launch_bounds__(256) sr_pack_ln_kernel(SrPackArgs a) {
  PDL_TOP();
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
  PDL_TOP();
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
  PDL_TOP();
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
  PDL_TOP();
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
  PDL_TOP();
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
  PDL_TOP();
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
  PDL_TOP();
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

// ---- training row of the class head (MSTr.py:212-227 pixel shuffle + LayerNorm(64), :288-289 1x1 conv to classes) ----
// Backward of final_head_kernel in ONE pass over the expand output e [B*H*W][16*64]: one thread per output pixel recomputes the
// LayerNorm statistics of its 64-float row, forms d(LN out) = sum_k dlogit_k cw[k][:] from the NCHW logit gradients (no
// [pixels][64] gradient tensor in memory), applies the LayerNorm backward and writes de in the UNSHUFFLED layout of e (the
// pixel shuffle is a row permutation).  Every parameter gradient follows from two per-block sums over pixels,
//   G[k][c] = sum_p dlogit[p][k] xhat[p][c],   s[k] = sum_p dlogit[p][k]:
//   d cw[k][c] = lnw[c] G[k][c] + lnb[c] s[k],  d cb[k] = s[k],  d lnw[c] = sum_k cw[k][c] G[k][c],  d lnb[c] = sum_k cw[k][c] s[k],
// accumulated per block through a shared-memory tile (xhat of the block's 128 pixels) in a fixed order, folded by
// final_head_fold_kernel in block order: bit-reproducible.
constexpr int FH_K = 16;         // class slots (ncls <= 16)
constexpr int FH_WARPS = 8;
constexpr int FH_ST = 6;         // quads in flight per warp (cp.async ring: 1 KB of rows + 256 B of logit gradients per stage)
constexpr int FH_DYN_SMEM = FH_WARPS * FH_ST * (64 * 16 + 64 * 4);
__device__ __forceinline__ void fh_cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fh_cp4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fh_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fh_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Work layout: a half-warp owns a pixel, lane (l & 15) its channels 4l .. 4l+3 (one float4 of the 256-byte row), so the 64-wide
// row reductions are 4 shuffle steps, the class weights cw[k][c] * lnw[c] and the G accumulators of the lane's four channels
// live in registers, and no shared memory is touched inside the loop.  A warp walks "quads" (the 4 horizontally adjacent output
// pixels of one (b, 4h + p1, w): 1 KB contiguous in e): half-warp h takes pixels p2 = h and h + 2; the next quad's rows are
// loaded before the current one is processed: every lane copies ITS two float4 and two logit gradients of the next FH_ST - 1 quads
// into a per-warp shared-memory ring with cp.async and later reads back only what it copied itself (no barrier needed) — the bytes
// in flight per SM (8 warps x 5 quads x 1.25 KB) no longer depend on registers.
template <int KC>
__global__ void __launch_bounds__(FH_WARPS * 32, 1) final_head_bwd_kernel(const float* __restrict__ e, const float* __restrict__ dlogits,
                                                                          int B, int H, int W, const float* __restrict__ lnw,
                                                                          const float* __restrict__ cw, float eps, int ncls,
                                                                          float* __restrict__ de, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float s_red[FH_WARPS][FH_K * 64 + FH_K];
  extern __shared__ float4 fh_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* sx = fh_dyn + (size_t)warp * FH_ST * 64;                                                   // [FH_ST][64] float4
  float* sd = reinterpret_cast<float*>(fh_dyn + (size_t)FH_WARPS * FH_ST * 64) + (size_t)warp * FH_ST * 64;   // [FH_ST][64] float
  const int hl = lane & 15, half = lane >> 4;
  const int c4 = hl * 4;
  const int Ho = H * 4, Wo = W * 4;
  const long long plane = (long long)Ho * Wo;
  const int nquad = B * Ho * W;            // < 2^31 (checked by the launcher): 32-bit index arithmetic in the loop
  float wr[KC][4], g[KC][4], sk[KC];
#pragma unroll
  for (int k = 0; k < KC; k++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      wr[k][j] = k < ncls ? cw[k * 64 + c4 + j] * lnw[c4 + j] : 0.f;
      g[k][j] = 0.f;
    }
    sk[k] = 0.f;
  }
  // quad qq -> float offset of its 256 floats in e, and the address of class hl's logit gradient of pixel p2 = half
  auto locate = [&](int qq, long long& eoff, const float*& dptr) {
    const unsigned t = (unsigned)qq / (unsigned)W;          // b * Ho + ho
    const unsigned w = (unsigned)qq - t * (unsigned)W;
    const unsigned b = t / (unsigned)Ho;
    const unsigned ho = t - b * (unsigned)Ho;
    eoff = ((((long long)b * H + (ho >> 2)) * W + w) * 16 + (ho & 3) * 4) * 64;
    dptr = dlogits + ((long long)b * ncls + hl) * plane + (long long)ho * Wo + 4 * w + half;
  };
  const int stride = (int)gridDim.x * FH_WARPS;
  int q = (int)blockIdx.x * FH_WARPS + warp;
  // lane hl < ncls of a half-warp fetches the logit gradient of class hl for the half-warp's two pixels (broadcast by shuffle when
  // it is needed), so the next quads' rows AND gradients are in flight while the current quad is processed
  long long offs[FH_ST];                   // e offsets of the quads in flight (ring, indexed like the smem slots)
  auto issue = [&](int qq, int slot_, long long& eoff) {
    if (qq < nquad) {
      const float* d;
      locate(qq, eoff, d);
      const float* src = e + eoff + c4;
      fh_cp16(sx + slot_ * 64 + lane * 2, src + half * 64);
      fh_cp16(sx + slot_ * 64 + lane * 2 + 1, src + (half + 2) * 64);
      if (hl < ncls) {
        fh_cp4(sd + slot_ * 64 + lane * 2, d);
        fh_cp4(sd + slot_ * 64 + lane * 2 + 1, d + 2);
      }
    }
    fh_commit();
  };
#pragma unroll
  for (int st = 0; st < FH_ST; st++) offs[st] = 0;
#pragma unroll
  for (int st = 0; st < FH_ST - 1; st++) issue(q + st * stride, st, offs[st]);
  // the loop body is unrolled FH_ST times so that ring slots (and offs[]) are compile-time indices
  for (; q < nquad;) {
#pragma unroll
    for (int slot = 0; slot < FH_ST; slot++) {
      if (q >= nquad) break;
      const int nslot = (slot + FH_ST - 1) % FH_ST;
      issue(q + (FH_ST - 1) * stride, nslot, offs[nslot]);
      fh_wait<FH_ST - 1>();                  // the oldest group (this quad) has landed
      const long long off = offs[slot];
      const float4 ca = sx[slot * 64 + lane * 2], cb = sx[slot * 64 + lane * 2 + 1];
      const float cda = hl < ncls ? sd[slot * 64 + lane * 2] : 0.f, cdb = hl < ncls ? sd[slot * 64 + lane * 2 + 1] : 0.f;
#pragma unroll
    for (int pp = 0; pp < 2; pp++) {        // pixel p2 = half + 2 * pp
      const float4 xv = pp ? cb : ca;
      const float dmine = pp ? cdb : cda;
      float dl[KC];
#pragma unroll
      for (int k = 0; k < KC; k++) dl[k] = __shfl_sync(0xffffffffu, dmine, (lane & 16) | k);     // 0 beyond ncls
      float x[4] = {xv.x, xv.y, xv.z, xv.w};
      float sum = (x[0] + x[1]) + (x[2] + x[3]);
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
      const float mean = sum * (1.f / 64.f);
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < 4; j++) { x[j] -= mean; var = fmaf(x[j], x[j], var); }
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) var += __shfl_xor_sync(0xffffffffu, var, m);
      const float rstd = rsqrtf(var * (1.f / 64.f) + eps);
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        x[j] *= rstd;                        // xhat
#pragma unroll
        for (int k = 0; k < KC; k++) {
          a[j] = fmaf(dl[k], wr[k][j], a[j]);
          g[k][j] = fmaf(dl[k], x[j], g[k][j]);
        }
        s1 += a[j];
        s2 = fmaf(a[j], x[j], s2);
      }
#pragma unroll
      for (int k = 0; k < KC; k++) sk[k] += dl[k];
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, m);
        s2 += __shfl_xor_sync(0xffffffffu, s2, m);
      }
      s1 *= (1.f / 64.f); s2 *= (1.f / 64.f);
      *reinterpret_cast<float4*>(de + off + (half + 2 * pp) * 64 + c4) =
          make_float4(rstd * (a[0] - s1 - x[0] * s2), rstd * (a[1] - s1 - x[1] * s2), rstd * (a[2] - s1 - x[2] * s2),
                      rstd * (a[3] - s1 - x[3] * s2));
    }
      q += stride;
    }
  }
  // warp sum = half 0 + half 1, then the eight warps in warp order
#pragma unroll
  for (int k = 0; k < KC; k++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float o = __shfl_xor_sync(0xffffffffu, g[k][j], 16);
      if (half == 0) s_red[warp][k * 64 + c4 + j] = g[k][j] + o;
    }
    const float o = __shfl_xor_sync(0xffffffffu, sk[k], 16);
    if (lane == 0) s_red[warp][FH_K * 64 + k] = sk[k] + o;
  }
  __syncthreads();
  float* pb = part + (size_t)blockIdx.x * (FH_K * 64 + FH_K);
  for (int i = threadIdx.x; i < FH_K * 64 + FH_K; i += FH_WARPS * 32) {
    const int k = i < FH_K * 64 ? i >> 6 : i - FH_K * 64;
    float s = 0.f;
    if (k < KC) {
      s = s_red[0][i];
#pragma unroll
      for (int wv = 1; wv < FH_WARPS; wv++) s += s_red[wv][i];
    }
    pb[i] = s;
  }
}

// fold of the block partials (ordered, bwd_fold_sum), then the four parameter gradients from the folded G | s
__global__ void __launch_bounds__(256) final_head_fold_kernel(const float* __restrict__ part, int nblk, float* __restrict__ Gs) {
  PDL_TOP();
  const int n = FH_K * 64 + FH_K;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const float s = bwd_fold_sum(part, nblk, n, i, i < n);
  if (threadIdx.y == 0 && i < n) Gs[i] = s;
}
__global__ void __launch_bounds__(256) final_head_params_kernel(const float* __restrict__ G, const float* __restrict__ lnw,
                                                                const float* __restrict__ lnb, const float* __restrict__ cw, int ncls,
                                                                float* __restrict__ dlnw, float* __restrict__ dlnb, float* __restrict__ dcw,
                                                                float* __restrict__ dcb) {
  PDL_TOP();
  const float* sk = G + FH_K * 64;
  for (int i = threadIdx.x; i < ncls * 64; i += 256) {
    const int k = i >> 6, c = i & 63;
    dcw[i] = fmaf(lnw[c], G[k * 64 + c], lnb[c] * sk[k]);
  }
  if (threadIdx.x < ncls) dcb[threadIdx.x] = sk[threadIdx.x];
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    float a = 0.f, b2 = 0.f;
    for (int k = 0; k < ncls; k++) { a = fmaf(cw[k * 64 + c], G[k * 64 + c], a); b2 = fmaf(cw[k * 64 + c], sk[k], b2); }
    dlnw[c] = a; dlnb[c] = b2;
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

int launch_patch_im2row(const float* x, long long xs_b, long long xs_c, int B, int Hin, int Win, int Kp, float* A, cudaStream_t st) {
  const int Ho = (Hin + 6 - 7) / 4 + 1, Wo = (Win + 6 - 7) / 4 + 1;
  const long long total = (long long)B * Ho * Wo * Kp;
  TCX_REQUIRE(Kp >= 147, "patch_im2row: row pitch %d < 147", Kp);
  if (total == 0) return 0;
  tcx_launch_chain(patch_im2row_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, xs_b, xs_c, Hin, Win, Ho, Wo, Kp, total, A);
  return tcx_check_launch("patch_im2row");
}

int launch_regroup(const RegroupArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.ntok * 16;
  tcx_launch_chain(regroup_kernel, dim3((unsigned)((total + 256 * RG_UNROLL - 1) / (256 * RG_UNROLL))), dim3(256), 0, st, a);
  return tcx_check_launch("bridge_regroup");
}

int launch_ungroup(const UngroupArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.ntok * 16;
  tcx_launch_chain(ungroup_kernel, dim3((unsigned)((total + 256 * RG_UNROLL - 1) / (256 * RG_UNROLL))), dim3(256), 0, st, a);
  return tcx_check_launch("bridge_ungroup");
}

int launch_sr_unpack(const SrUnpackArgs& a, cudaStream_t st) {
  const long long rows = (long long)a.B * a.nred;
  tcx_launch_chain(sr_unpack_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, a);
  return tcx_check_launch("sr_unpack");
}

int launch_sr_row2im(const float* dA, long long xs_b, int HW, int Cin, int r, int B, float* dx, cudaStream_t st) {
  const int P = HW / r;
  TCX_REQUIRE(Cin % 4 == 0 && (xs_b % 4) == 0, "sr_row2im: Cin and the image stride must be multiples of 4 (Cin=%d)", Cin);
  const long long total = (long long)B * P * P * Cin * r * r / 4;
  tcx_launch_chain(sr_row2im_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, dA, xs_b, HW, Cin, r, B, dx);
  return tcx_check_launch("sr_row2im");
}

int launch_sr_im2row(const float* x, long long xs_b, int HW, int Cin, int r, int B, float* A, cudaStream_t st) {
  const int P = HW / r;
  TCX_REQUIRE(Cin % 4 == 0 && (xs_b % 4) == 0, "sr_im2row: Cin and the image stride must be multiples of 4 (Cin=%d)", Cin);
  const long long total = (long long)B * P * P * Cin * r * r / 4;
  tcx_launch_chain(sr_im2row_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, xs_b, HW, Cin, r, B, A);
  return tcx_check_launch("sr_im2row");
}

int launch_sr_pack_ln(const SrPackArgs& a, cudaStream_t st) {
  const long long rows = (long long)a.B * a.nred;
  tcx_launch_chain(sr_pack_ln_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, a);
  return tcx_check_launch("sr_pack_ln");
}

int launch_iff_pool(const IffSrc& src, int B, int HW, int C, float* pooled, cudaStream_t st) {
  dim3 grid(HW, 4, 2 * B);
  int threads = (C + 31) / 32 * 32;
  if (threads > 512) threads = 512;
  tcx_launch_chain(iff_pool_kernel, dim3(grid), dim3(threads), 0, st, src, B, HW, C, pooled);
  return tcx_check_launch("iff_pool");
}

int launch_iff_gate(const IffSrc& src, int B, int H, int W, int C, const float* ah, const float* aw, float* out,
                    cudaStream_t st) {
  const long long total = (long long)B * H * W * C;
  tcx_launch_chain(iff_gate_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, src, B, H, W, C, ah, aw, out);
  return tcx_check_launch("iff_gate");
}

int launch_shuffle_ln(const float* in, int B, int H, int W, int s, int c, const float* lnw, const float* lnb, float eps,
                      float* out, cudaStream_t st) {
  TCX_REQUIRE(c <= 256, "patch_expand: c=%d > 256", c);
  const long long rows = (long long)B * H * s * W * s;
  tcx_launch_chain(shuffle_ln_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, in, B, H, W, s, c, lnw, lnb, eps, out);
  return tcx_check_launch("shuffle_ln");
}

int launch_final_head(const float* in, int B, int H, int W, const float* lnw, const float* lnb, float eps,
                      const float* cw, const float* cb, int ncls, float* out, cudaStream_t st) {
  TCX_REQUIRE(ncls >= 1 && ncls <= 32, "final_head: ncls=%d out of range (1..32)", ncls);
  const long long total = (long long)B * H * 4 * W * 4;
  tcx_launch_chain(final_head_kernel<32>, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, st, in, B, H, W, lnw, lnb, eps, cw, cb, ncls, out);
  return tcx_check_launch("final_head");
}

int final_head_bwd_blocks(int B, int H, int W) {
  const long long quads = (long long)B * H * 4 * W;
  long long nblk = 148;                           // one resident block per SM; every warp strides over the quads
  if (nblk * FH_WARPS > quads) nblk = (quads + FH_WARPS - 1) / FH_WARPS;
  return (int)(nblk < 1 ? 1 : nblk);
}
size_t final_head_bwd_part_floats(int B, int H, int W) { return (size_t)(final_head_bwd_blocks(B, H, W) + 1) * (FH_K * 64 + FH_K); }
int launch_final_head_bwd(const float* e, const float* dlogits, int B, int H, int W, const float* lnw, float eps, const float* cw,
                          int ncls, float* de, float* part, int* nblk_out, cudaStream_t st) {
  TCX_REQUIRE(ncls >= 1 && ncls <= FH_K, "final_head_bwd: ncls=%d out of range (1..%d)", ncls, FH_K);
  TCX_REQUIRE((long long)B * H * W > 0 && (long long)B * H * 4 * W * 4 < (1ll << 31), "final_head_bwd: empty or too large a batch");
  const int grid = final_head_bwd_blocks(B, H, W);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t ce = cudaFuncSetAttribute(final_head_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_DYN_SMEM);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(final_head_bwd_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_DYN_SMEM);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(final_head_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_DYN_SMEM);
    TCX_REQUIRE(ce == cudaSuccess, "final_head_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
  }
  if (ncls <= 4) tcx_launch_chain(final_head_bwd_kernel<4>, dim3(grid), dim3(FH_WARPS * 32), FH_DYN_SMEM, st, e, dlogits, B, H, W, lnw, cw, eps, ncls, de, part);
  else if (ncls <= 10) tcx_launch_chain(final_head_bwd_kernel<10>, dim3(grid), dim3(FH_WARPS * 32), FH_DYN_SMEM, st, e, dlogits, B, H, W, lnw, cw, eps, ncls, de, part);
  else tcx_launch_chain(final_head_bwd_kernel<16>, dim3(grid), dim3(FH_WARPS * 32), FH_DYN_SMEM, st, e, dlogits, B, H, W, lnw, cw, eps, ncls, de, part);
  *nblk_out = grid;
  return tcx_check_launch("final_head_bwd");
}
// part: the block partials of launch_final_head_bwd followed by one more slot (the folded sums)
int launch_final_head_bwd_fold(float* part, int nblk, const float* lnw, const float* lnb, const float* cw, int ncls, float* dlnw,
                               float* dlnb, float* dcw, float* dcb, cudaStream_t st) {
  float* Gs = part + (size_t)nblk * (FH_K * 64 + FH_K);
  tcx_launch_chain(final_head_fold_kernel, dim3((FH_K * 64 + FH_K + 31) / 32), dim3(dim3(32, 8)), 0, st, part, nblk, Gs);
  TCX_TRY(tcx_check_launch("final_head_fold"));
  tcx_launch_chain(final_head_params_kernel, dim3(1), dim3(256), 0, st, Gs, lnw, lnb, cw, ncls, dlnw, dlnb, dcw, dcb);
  return tcx_check_launch("final_head_params");
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
  tcx_launch_chain(conv_weight_perm16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, w, reinterpret_cast<__half*>(o16), N, Cin, r);
  return tcx_check_launch("conv_weight_perm16");
}

int launch_sr_pack_ln16(const SrPackArgs& a, const void* x16, void* out16, cudaStream_t st) {
  const long long rows = (long long)a.B * a.nred;
  if (rows == 0) return 0;
  tcx_launch_pdl(sr_pack_ln16_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, a,
                 reinterpret_cast<const __half*>(x16), reinterpret_cast<__half*>(out16));
  return tcx_check_launch("sr_pack_ln16");
}
