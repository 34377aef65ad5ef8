// HBM/L2-bound glue kernels of the fp16-intermediate pipeline (fp32 residual streams, fp16 tensor-core operands):
//  * dwln_kernel   — u = dw3x3(x) + b + x ; y = [GELU](LayerNorm(u))  (reference MSTr.py:59 Mix-FFN middle with fp16 h;
//                    MSTr.py:939-942 ConvPosEnc + norm1 with an fp32 stream).  One warp (or half warp) per token, the row
//                    stays in registers between the convolution and the normalisation; filter taps, bias and LN affine
//                    are staged once per block in shared memory (tap-major); a warp walks consecutive tokens so 6 of
//                    the 9 neighbour rows hit L1.
//  * ln16_kernel   — LayerNorm of an fp32 stream to fp16 (and/or fp32).
//  * mb_fused16    — Multi-Branch factorized attention + conv relative position encoding (MSTr.py:852-886, :801-823) on
//                    an fp16 qkv buffer, one block per (head, image, branch): column softmax, K^T V, q x ctx and the
//                    head's 3x3 / 5x5 / 7x7 depthwise filter from one shared-memory copy of the head's q / k / v.
#include <cuda_fp16.h>
#include "common.cuh"
#include "fused16.cuh"
#include <mutex>
#include <type_traits>
#include <unordered_map>

namespace {

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int LPT>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// dw3x3 + skip + LayerNorm (+GELU)
// ------------------------------------------------------------------------------------------------------------------
// GELU: tcx_gelu_fast (fused16.cuh)

// channel c of a token row -> position inside the per-tap smem vectors: lanes read contiguous 16-byte pieces
template <int LPT, int VEC>
__device__ __forceinline__ int wperm(int c) {
  if (VEC == 4) return c;
  const int blk = c / (LPT * 8), r = c - blk * (LPT * 8);      // NV block, offset inside it
  const int sl = r >> 3, q = (r >> 2) & 1, j = r & 3;
  return blk * (LPT * 8) + q * (LPT * 4) + sl * 4 + j;
}

template <bool IN16, int LPT, int NV>
__global__ void __launch_bounds__(256) dwln_kernel(const DwLnArgs a) {
  constexpr int VEC = IN16 ? 8 : 4;
  constexpr int SLOTS = 32 / LPT;
  constexpr int C = LPT * VEC * NV;          // the row width is fixed by the instantiation: all offsets are immediates
  using elem_t = typename std::conditional<IN16, __half, float>::type;
  using vec_t = typename std::conditional<IN16, uint4, float4>::type;
  extern __shared__ float wsm[];            // [9][C] taps (centre + 1 = skip), then bias[C], lnw[C], lnb[C] (permuted)
  const DwLnGroup& g = a.g[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  // (filters, bias and LN affine are module parameters: staged before pdl_wait, overlapping the previous kernel)
  // staging: loads are issued in batches of 8 before the dependent smem stores, so a thread has 8 requests in flight
  // (a one-load-per-iteration loop costs one L2 round trip per element and dominated the small-map launches)
  constexpr int NB = 8;
  for (int i0 = tid; i0 < 9 * C; i0 += 256 * NB) {
    float v[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) { const int i = i0 + j * 256; v[j] = i < 9 * C ? __ldg(g.dww + i) : 0.f; }
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int i = i0 + j * 256;
      if (i < 9 * C) { const int c = i / 9, t = i - c * 9; wsm[t * C + wperm<LPT, VEC>(c)] = v[j] + (t == 4 ? 1.f : 0.f); }
    }
  }
  float* bsm = wsm + 9 * C;
  {
    constexpr int PER = (C + 255) / 256;
    float vb[PER], vw[PER], vl[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
      const int i = tid + j * 256;
      vb[j] = (g.dwb && i < C) ? __ldg(g.dwb + i) : 0.f;
      vw[j] = i < C ? __ldg(g.lnw + i) : 0.f;
      vl[j] = i < C ? __ldg(g.lnb + i) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < PER; j++) {
      const int i = tid + j * 256;
      if (i < C) { const int pc = wperm<LPT, VEC>(i); bsm[pc] = vb[j]; bsm[C + pc] = vw[j]; bsm[2 * C + pc] = vl[j]; }
    }
  }
  __syncthreads();
  pdl_wait();
  const int slot = lane / LPT, sl = lane % LPT;
  const int total = a.B * a.H * a.W;
  const int H = a.H, W = a.W;
  constexpr float invC = 1.f / (float)C;
  const int chunk = SLOTS * a.tpw;                       // consecutive tokens per warp visit
  const int nchunks = (total + chunk - 1) / chunk;
  const float* wl = wsm + sl * 4;                        // this lane's slice of every per-channel smem vector
  const float* bl = bsm + sl * 4;
  const elem_t* __restrict__ xin = reinterpret_cast<const elem_t*>(g.x) + sl * VEC;
  const ptrdiff_t rowpitch = (ptrdiff_t)W * C;
  // the 9 taps of channel vector i of one token: [row pointer + immediate] under a predicate, so the loads are
  // independent and all in flight together; out-of-map taps read nothing and stay zero
  auto load9 = [&](int tok, int wq, int hq, int i, vec_t (&raw)[9]) {
    const bool live = tok < total;
    const elem_t* pc = xin + (size_t)(live ? tok : 0) * C + i * LPT * VEC;
    const elem_t* prow[3] = {pc - rowpitch, pc, pc + rowpitch};
    const bool rv[3] = {live && hq > 0, live, live && hq + 1 < H};
    const bool cv[3] = {wq > 0, true, wq + 1 < W};
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int ky = t / 3, kx = t % 3;
      if (IN16) *reinterpret_cast<uint4*>(&raw[t]) = make_uint4(0u, 0u, 0u, 0u);
      else *reinterpret_cast<float4*>(&raw[t]) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rv[ky] && cv[kx]) raw[t] = *reinterpret_cast<const vec_t*>(prow[ky] + (kx - 1) * C);
    }
  };
  for (int ch = blockIdx.x * 8 + warp; ch < nchunks; ch += gridDim.x * 8) {
    const int base = ch * chunk;
    int tok = base + slot;
    int wq = tok % W, hq = (tok / W) % H;
    vec_t cur[9], nxt[9];
    load9(tok, wq, hq, 0, cur);
    for (int it = 0; it < a.tpw; it++, tok += SLOTS) {
      if (base + it * SLOTS >= total) break;             // warp-uniform
      const bool live = tok < total;
      // position of this slot's next token (SLOTS further along the row-major token order)
      int wn = wq + SLOTS, hn = hq;
      if (wn >= W) { wn -= W; if (++hn >= H) hn = 0; }
      float acc[NV][VEC];
#pragma unroll
      for (int i = 0; i < NV; i++) {
#pragma unroll
        for (int q = 0; q < VEC / 4; q++) {
          const float4 b4 = *reinterpret_cast<const float4*>(bl + i * LPT * VEC + q * LPT * 4);
          acc[i][q * 4 + 0] = b4.x; acc[i][q * 4 + 1] = b4.y; acc[i][q * 4 + 2] = b4.z; acc[i][q * 4 + 3] = b4.w;
        }
      }
#pragma unroll
      for (int i = 0; i < NV; i++) {
        // software pipeline: the next batch (next channel vector, or the first one of the next token) is requested
        // before this one is consumed
        if (i + 1 < NV) load9(tok, wq, hq, i + 1, nxt);
        else if (it + 1 < a.tpw) load9(tok + SLOTS, wn, hn, 0, nxt);
#pragma unroll
        for (int t = 0; t < 9; t++) {
          float xv[VEC];
          if (IN16) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(&cur[t]), f);
#pragma unroll
            for (int j = 0; j < VEC; j++) xv[j] = f[j];
          } else {
            const float4 r = *reinterpret_cast<const float4*>(&cur[t]);
            xv[0] = r.x; xv[1] = r.y; xv[2] = r.z; xv[3] = r.w;
          }
#pragma unroll
          for (int q = 0; q < VEC / 4; q++) {
            const float4 w4 = *reinterpret_cast<const float4*>(wl + t * C + i * LPT * VEC + q * LPT * 4);
            acc[i][q * 4 + 0] = fmaf(xv[q * 4 + 0], w4.x, acc[i][q * 4 + 0]);
            acc[i][q * 4 + 1] = fmaf(xv[q * 4 + 1], w4.y, acc[i][q * 4 + 1]);
            acc[i][q * 4 + 2] = fmaf(xv[q * 4 + 2], w4.z, acc[i][q * 4 + 2]);
            acc[i][q * 4 + 3] = fmaf(xv[q * 4 + 3], w4.w, acc[i][q * 4 + 3]);
          }
        }
#pragma unroll
        for (int t = 0; t < 9; t++) cur[t] = nxt[t];
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; i++)
#pragma unroll
        for (int j = 0; j < VEC; j++) s += acc[i][j];
      const float mean = seg_sum<LPT>(s) * invC;
      float q2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; i++)
#pragma unroll
        for (int j = 0; j < VEC; j++) { const float d = acc[i][j] - mean; q2 = fmaf(d, d, q2); }
      const float rstd = rsqrtf(seg_sum<LPT>(q2) * invC + a.eps);
      if (live) {
        const size_t off0 = (size_t)tok * C + sl * VEC;
#pragma unroll
        for (int i = 0; i < NV; i++) {
          if (g.u) {
#pragma unroll
            for (int q = 0; q < VEC / 4; q++)
              *reinterpret_cast<float4*>(g.u + off0 + i * LPT * VEC + q * 4) =
                  make_float4(acc[i][q * 4], acc[i][q * 4 + 1], acc[i][q * 4 + 2], acc[i][q * 4 + 3]);
          }
          float o[VEC];
#pragma unroll
          for (int q = 0; q < VEC / 4; q++) {
            const float4 w4 = *reinterpret_cast<const float4*>(bl + C + i * LPT * VEC + q * LPT * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(bl + 2 * C + i * LPT * VEC + q * LPT * 4);
            o[q * 4 + 0] = fmaf((acc[i][q * 4 + 0] - mean) * rstd, w4.x, b4.x);
            o[q * 4 + 1] = fmaf((acc[i][q * 4 + 1] - mean) * rstd, w4.y, b4.y);
            o[q * 4 + 2] = fmaf((acc[i][q * 4 + 2] - mean) * rstd, w4.z, b4.z);
            o[q * 4 + 3] = fmaf((acc[i][q * 4 + 3] - mean) * rstd, w4.w, b4.w);
          }
          if (a.gelu) {
#pragma unroll
            for (int j = 0; j < VEC; j++) o[j] = tcx_gelu_fast(o[j]);
          }
          __half* yp = g.y + off0 + i * LPT * VEC;
          if (VEC == 8) {
            *reinterpret_cast<uint4*>(yp) = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[VEC - 4], o[VEC - 3]),
                                                       pack2(o[VEC - 2], o[VEC - 1]));
          } else {
            *reinterpret_cast<uint2*>(yp) = make_uint2(pack2(o[0], o[1]), pack2(o[2], o[3]));
          }
        }
      }
      wq = wn; hq = hn;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm fp32 -> fp16 (and/or fp32)
// ------------------------------------------------------------------------------------------------------------------
template <int LPT, int NV>
__global__ void __launch_bounds__(256) ln16_kernel(const Ln16Args a) {
  constexpr int SLOTS = 32 / LPT;
  const Ln16Group& g = a.g[blockIdx.y];
  const int C = a.C, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / LPT, sl = lane % LPT;
  const long long base = ((long long)blockIdx.x * 8 + warp) * (SLOTS * a.tpw);
  const float invC = 1.f / (float)C;
  pdl_trigger();
  float4 wv[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    wv[i] = *reinterpret_cast<const float4*>(g.w + (sl + i * LPT) * 4);
    bv[i] = *reinterpret_cast<const float4*>(g.b + (sl + i * LPT) * 4);
  }
  pdl_wait();
  for (int it = 0; it < a.tpw; it++) {
    const long long row = base + (long long)it * SLOTS + slot;
    if (base + (long long)it * SLOTS >= a.M) break;
    const bool live = row < a.M;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      v[i] = live ? *reinterpret_cast<const float4*>(g.x + row * C + (sl + i * LPT) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = seg_sum<LPT>(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float rstd = rsqrtf(seg_sum<LPT>(q) * invC + a.eps);
    if (!live) continue;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const int c0 = (sl + i * LPT) * 4;
      const float o0 = fmaf((v[i].x - mean) * rstd, wv[i].x, bv[i].x), o1 = fmaf((v[i].y - mean) * rstd, wv[i].y, bv[i].y);
      const float o2 = fmaf((v[i].z - mean) * rstd, wv[i].z, bv[i].z), o3 = fmaf((v[i].w - mean) * rstd, wv[i].w, bv[i].w);
      if (g.y16) *reinterpret_cast<uint2*>(g.y16 + row * C + c0) = make_uint2(pack2(o0, o1), pack2(o2, o3));
      if (g.y32) *reinterpret_cast<float4*>(g.y32 + row * C + c0) = make_float4(o0, o1, o2, o3);
    }
  }
}

// exact i / d for 0 <= i < 2^20 and small d via one multiply (inv = 1.0f / d): (i + 0.5) / d is never within 2^-20
// of an integer, so float rounding cannot cross a boundary.
__device__ __forceinline__ int fdiv(int i, float inv) { return (int)(((float)i + 0.5f) * inv); }

// ------------------------------------------------------------------------------------------------------------------
// Fused Multi-Branch attention, one block per (head, image, branch): column softmax of K over the tokens, the Ch x Ch
// context, the factorized-attention product and the conv relative position encoding of that head's channels, all from
// one shared-memory copy of the head's q / k / v slices (the whole H x W map of a head fits: <= 50 KB at 28 x 28).
// The window size is uniform inside a block (heads 0-1: 3x3, 2-4: 5x5, 5-7: 7x7, MSTr.py:958).
// ------------------------------------------------------------------------------------------------------------------
constexpr int MBF_THREADS = 256;

template <int WIN, int T>
__device__ __forceinline__ void mbf_apply(const __half* __restrict__ qs, const __half* __restrict__ vs, const float* __restrict__ ctx,
                                          const float* __restrict__ wt, const float* __restrict__ bs, __half* __restrict__ outp,
                                          int H, int W, int C, int Ch, int CHP, int tid) {
  constexpr int R = WIN / 2;
  const int np = Ch / 2, xg_n = W / T;
  const int items = H * xg_n * np;
  const float inv_np = 1.0f / (float)np, inv_xg = 1.0f / (float)xg_n;
  for (int i = tid; i < items; i += MBF_THREADS) {
    const int i1 = fdiv(i, inv_np), p = i - i1 * np;
    const int y = fdiv(i1, inv_xg), xg = i1 - y * xg_n;
    const int c = 2 * p, x0 = xg * T;
    float v0[T], v1[T];
#pragma unroll
    for (int t = 0; t < T; t++) { v0[t] = bs[c]; v1[t] = bs[c + 1]; }
#pragma unroll
    for (int ky = 0; ky < WIN; ky++) {
      const int yy = y + ky - R;
      if (yy < 0 || yy >= H) continue;
      const __half* vrow = vs + (size_t)(yy * W) * CHP + c;
      float2 win[T + WIN - 1];
#pragma unroll
      for (int j = 0; j < T + WIN - 1; j++) {
        const int xx = x0 + j - R;
        win[j] = (xx >= 0 && xx < W) ? __half22float2(*reinterpret_cast<const __half2*>(vrow + (size_t)xx * CHP))
                                     : make_float2(0.f, 0.f);
      }
      const float* wrow = wt + (size_t)(ky * WIN) * Ch + c;
#pragma unroll
      for (int kx = 0; kx < WIN; kx++) {
        const float2 ww = *reinterpret_cast<const float2*>(wrow + (size_t)kx * Ch);
#pragma unroll
        for (int t = 0; t < T; t++) {
          v0[t] = fmaf(win[t + kx].x, ww.x, v0[t]);
          v1[t] = fmaf(win[t + kx].y, ww.y, v1[t]);
        }
      }
    }
    float f0[T], f1[T];
#pragma unroll
    for (int t = 0; t < T; t++) { f0[t] = 0.f; f1[t] = 0.f; }
    const __half* qbase = qs + (size_t)(y * W + x0) * CHP;
    for (int k = 0; k < Ch; k += 2) {
      const float2 ca = *reinterpret_cast<const float2*>(ctx + (size_t)k * Ch + c);
      const float2 cb = *reinterpret_cast<const float2*>(ctx + (size_t)(k + 1) * Ch + c);
#pragma unroll
      for (int t = 0; t < T; t++) {
        const float2 q2 = __half22float2(*reinterpret_cast<const __half2*>(qbase + (size_t)t * CHP + k));
        f0[t] = fmaf(q2.x, ca.x, f0[t]); f1[t] = fmaf(q2.x, ca.y, f1[t]);
        f0[t] = fmaf(q2.y, cb.x, f0[t]); f1[t] = fmaf(q2.y, cb.y, f1[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < T; t++) {
      const float2 qc = __half22float2(*reinterpret_cast<const __half2*>(qbase + (size_t)t * CHP + c));
      *reinterpret_cast<uint32_t*>(outp + (size_t)(y * W + x0 + t) * C + c) = pack2(fmaf(qc.x, v0[t], f0[t]), fmaf(qc.y, v1[t], f1[t]));
    }
  }
}

template <int T>
__global__ void __launch_bounds__(MBF_THREADS) mb_fused16_kernel(const Mb16Args a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int H = a.H, W = a.W, C = a.C, Ch = a.C / a.heads, N = H * W;
  const int CHP = Ch + 2;                         // padded pitch (halfs) of the q / v tiles: spreads banks
  const int tid = threadIdx.x, h = blockIdx.x, b = blockIdx.y, gi = blockIdx.z;
  const int wi = h < 2 ? 0 : (h < 5 ? 1 : 2), win = 3 + 2 * wi;
  const int cl0 = (h - (wi == 0 ? 0 : (wi == 1 ? 2 : 5))) * Ch;     // first channel of this head inside its filter group
  float* E = reinterpret_cast<float*>(smraw);                       // [N][Ch] exp(k - max)
  float* ctx = E + (size_t)N * Ch;                                  // [Ch][Ch]
  float* wt = ctx + Ch * Ch;                                        // [win*win][Ch]
  float* bs = wt + 49 * Ch;                                         // [Ch]
  float* red = bs + Ch;                                             // [256]
  float* mx = red + 256;                                            // [Ch]
  float* ss = mx + Ch;                                              // [Ch]
  float* pbuf = ss + Ch;                                            // [1024] per-split partial contexts
  __half* qs = reinterpret_cast<__half*>(pbuf + 1024);              // [N][CHP]
  __half* vs = qs + (size_t)N * CHP;                                // [N][CHP]
  pdl_trigger();
  // filters / bias of this head (module parameters) before pdl_wait
  {
    const float* __restrict__ gw = a.cw[gi][wi] + (size_t)cl0 * win * win;
    const int nt = win * win;
    for (int i = tid; i < nt * Ch; i += MBF_THREADS) {
      const int ch = i / nt, t = i - ch * nt;
      wt[t * Ch + ch] = gw[i];
    }
    for (int i = tid; i < Ch; i += MBF_THREADS) bs[i] = a.cb[gi][wi][cl0 + i];
    for (int i = tid; i < Ch * Ch; i += MBF_THREADS) ctx[i] = 0.f;
  }
  pdl_wait();
  const __half* __restrict__ base = a.qkv[gi] + (long long)b * N * 3 * C + h * Ch;
  const int vpr = Ch / 8;
  const float inv_vpr = 1.0f / (float)vpr;
  for (int i = tid; i < N * vpr; i += MBF_THREADS) {
    const int n = fdiv(i, inv_vpr), j = i - n * vpr;
    const __half* src = base + (long long)n * 3 * C + j * 8;
    const uint4 rq = *reinterpret_cast<const uint4*>(src);
    const uint4 rk = *reinterpret_cast<const uint4*>(src + C);
    const uint4 rv = *reinterpret_cast<const uint4*>(src + 2 * C);
    float f[8];
    unpack8(rk, f);
    float* ed = E + n * Ch + j * 8;
    *reinterpret_cast<float4*>(ed) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(ed + 4) = make_float4(f[4], f[5], f[6], f[7]);
    uint32_t* qd = reinterpret_cast<uint32_t*>(qs + (size_t)n * CHP + j * 8);
    uint32_t* vd = reinterpret_cast<uint32_t*>(vs + (size_t)n * CHP + j * 8);
    qd[0] = rq.x; qd[1] = rq.y; qd[2] = rq.z; qd[3] = rq.w;
    vd[0] = rv.x; vd[1] = rv.y; vd[2] = rv.z; vd[3] = rv.w;
  }
  __syncthreads();
  const int nsl = MBF_THREADS / Ch;
  const int ck = tid % Ch, slc = tid / Ch;
  {
    float m = -INFINITY;
    if (slc < nsl)
      for (int n = slc; n < N; n += nsl) m = fmaxf(m, E[n * Ch + ck]);
    red[tid] = m;
  }
  __syncthreads();
  if (tid < Ch) {
    float mm = -INFINITY;
    for (int s2 = 0; s2 < nsl; s2++) mm = fmaxf(mm, red[s2 * Ch + tid]);
    mx[tid] = mm;
  }
  __syncthreads();
  {
    float sum = 0.f;
    if (slc < nsl) {
      const float mc = mx[ck];
      for (int n = slc; n < N; n += nsl) {
        const float e = __expf(E[n * Ch + ck] - mc);
        E[n * Ch + ck] = e;
        sum += e;
      }
    }
    red[tid] = slc < nsl ? sum : 0.f;
  }
  __syncthreads();
  if (tid < Ch) {
    float t = 0.f;
    for (int s2 = 0; s2 < nsl; s2++) t += red[s2 * Ch + tid];
    ss[tid] = t;
  }
  // context: 2 x 2 register tiles of (k, v) pairs, token range split over the remaining threads; the per-split partial
  // sums are combined in a fixed order (no atomics: the forward is bit-reproducible run to run)
  {
    const int tiles = (Ch / 2) * (Ch / 2);
    const int nsp = tiles >= MBF_THREADS ? 1 : MBF_THREADS / tiles;
    for (int idx = tid; idx < tiles * nsp; idx += MBF_THREADS) {
      const int tl = idx % tiles, sp = idx / tiles;
      const int k = (tl / (Ch / 2)) * 2, v = (tl % (Ch / 2)) * 2;
      float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
      for (int n = sp; n < N; n += nsp) {
        const float2 e = *reinterpret_cast<const float2*>(E + n * Ch + k);
        const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(vs + (size_t)n * CHP + v));
        a00 = fmaf(e.x, vv.x, a00); a01 = fmaf(e.x, vv.y, a01);
        a10 = fmaf(e.y, vv.x, a10); a11 = fmaf(e.y, vv.y, a11);
      }
      if (nsp > 1) {
        *reinterpret_cast<float4*>(pbuf + ((size_t)sp * tiles + tl) * 4) = make_float4(a00, a01, a10, a11);
      } else {
        ctx[k * Ch + v] = a00; ctx[k * Ch + v + 1] = a01;
        ctx[(k + 1) * Ch + v] = a10; ctx[(k + 1) * Ch + v + 1] = a11;
      }
    }
    if (nsp > 1) {
      __syncthreads();
      for (int tl = tid; tl < tiles; tl += MBF_THREADS) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp = 0; sp < nsp; sp++) {
          const float4 t = *reinterpret_cast<const float4*>(pbuf + ((size_t)sp * tiles + tl) * 4);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const int k = (tl / (Ch / 2)) * 2, v = (tl % (Ch / 2)) * 2;
        ctx[k * Ch + v] = acc.x; ctx[k * Ch + v + 1] = acc.y;
        ctx[(k + 1) * Ch + v] = acc.z; ctx[(k + 1) * Ch + v + 1] = acc.w;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < Ch * Ch; i += MBF_THREADS) ctx[i] = a.scale * ctx[i] / ss[i / Ch];
  __syncthreads();
  __half* __restrict__ outp = a.out[gi] + (long long)b * N * C + h * Ch;
  if (wi == 0) mbf_apply<3, T>(qs, vs, ctx, wt, bs, outp, H, W, C, Ch, CHP, tid);
  else if (wi == 1) mbf_apply<5, T>(qs, vs, ctx, wt, bs, outp, H, W, C, Ch, CHP, tid);
  else mbf_apply<7, T>(qs, vs, ctx, wt, bs, outp, H, W, C, Ch, CHP, tid);
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  static std::mutex mu;
  static std::unordered_map<unsigned long long, size_t> granted;       // per (device, kernel): the attribute is per device
  if (bytes <= 48 * 1024) return 0;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& g = granted[(unsigned long long)reinterpret_cast<uintptr_t>(reinterpret_cast<const void*>(kernel)) * 64ull + (unsigned long long)(dev & 63)];
  if (bytes > g) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { tcx_set_error("%s: cannot opt in to %zu bytes of shared memory: %s", what, bytes, cudaGetErrorString(e)); return -1; }
    g = bytes;
  }
  return 0;
}

int tokens_per_warp(long long rows, int slots) {
  // enough warps to fill the chip twice, at most 8 consecutive tokens per warp slot
  long long tpw = rows / ((long long)148 * 8 * 2 * slots);
  if (tpw < 1) tpw = 1;
  if (tpw > 8) tpw = 8;
  return (int)tpw;
}

}  // namespace

template <bool IN16, int LPT, int NV>
int dwln_launch(DwLnArgs a, int groups, cudaStream_t st) {
  constexpr int slots = 32 / LPT;
  const long long total = (long long)a.B * a.H * a.W;
  const size_t smem = (size_t)12 * a.C * sizeof(float);
  TCX_TRY(set_smem(dwln_kernel<IN16, LPT, NV>, smem, "dwln"));
  static int occ = 0;                               // resident blocks per SM of this instantiation (same smem every call)
  if (!occ) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dwln_kernel<IN16, LPT, NV>, 256, smem) != cudaSuccess || occ < 1) occ = 1;
  }
  // one visit per warp: tokens are split evenly over the resident warps, each warp walking `tpw` consecutive tokens
  // per slot with a one-deep software pipeline
  long long cap = (long long)148 * occ / groups;
  if (cap < 1) cap = 1;
  long long nblk = (total + 8 * slots - 1) / (8 * slots);        // at least one token per warp slot
  if (nblk > cap) nblk = cap;
  long long tpw = (total + nblk * 8 * slots - 1) / (nblk * 8 * slots);
  if (tpw > 16) tpw = 16;                                         // larger inputs take several visits per warp
  a.tpw = (int)tpw;
  dim3 grid((unsigned)nblk, groups);
  double bytes = 0.0;
  for (int i = 0; i < groups; i++) bytes += (double)total * a.C * ((IN16 ? 2.0 : 4.0) + 2.0 + (a.g[i].u ? 4.0 : 0.0));
  ProfScope prof("dwln", st, bytes);
  tcx_launch_pdl(dwln_kernel<IN16, LPT, NV>, grid, dim3(256), smem, st, a);
  return tcx_check_launch("dwln");
}

int launch_dwln(DwLnArgs a, int groups, bool in16, cudaStream_t st) {
  const int VEC = in16 ? 8 : 4;
  TCX_REQUIRE(a.C % (16 * VEC) == 0, "dwln: C=%d must be a multiple of %d", a.C, 16 * VEC);
  const int LPT = a.C % (32 * VEC) == 0 ? 32 : 16;
  const int NV = a.C / (LPT * VEC);
  const long long total = (long long)a.B * a.H * a.W;
  if (total == 0) return 0;
  TCX_REQUIRE(total * a.C < (1ll << 31), "dwln: tensor too large for 32-bit indexing");
  if (in16) {
    if (LPT == 32 && NV == 1) return dwln_launch<true, 32, 1>(a, groups, st);
    if (LPT == 32 && NV == 2) return dwln_launch<true, 32, 2>(a, groups, st);
    if (LPT == 32 && NV == 4) return dwln_launch<true, 32, 4>(a, groups, st);
    if (LPT == 32 && NV == 5) return dwln_launch<true, 32, 5>(a, groups, st);
    if (LPT == 32 && NV == 8) return dwln_launch<true, 32, 8>(a, groups, st);
    if (LPT == 16 && NV == 1) return dwln_launch<true, 16, 1>(a, groups, st);
    if (LPT == 16 && NV == 5) return dwln_launch<true, 16, 5>(a, groups, st);
  } else {
    if (LPT == 32 && NV == 1) return dwln_launch<false, 32, 1>(a, groups, st);
    if (LPT == 32 && NV == 2) return dwln_launch<false, 32, 2>(a, groups, st);
    if (LPT == 32 && NV == 4) return dwln_launch<false, 32, 4>(a, groups, st);
    if (LPT == 16 && NV == 1) return dwln_launch<false, 16, 1>(a, groups, st);
    if (LPT == 16 && NV == 5) return dwln_launch<false, 16, 5>(a, groups, st);
  }
  tcx_set_error("dwln: unsupported width C=%d (in16=%d)", a.C, (int)in16);
  return -1;
}

int launch_ln16(Ln16Args a, int groups, cudaStream_t st) {
  TCX_REQUIRE(a.C % 64 == 0, "ln16: C=%d must be a multiple of 64", a.C);
  if (a.M == 0) return 0;
  const int LPT = a.C % 128 == 0 ? 32 : 16;
  const int NV = a.C / (LPT * 4);
  const int slots = 32 / LPT;
  a.tpw = tokens_per_warp(a.M, slots);
  const long long per_block = (long long)8 * slots * a.tpw;
  dim3 grid((unsigned)((a.M + per_block - 1) / per_block), groups);
  ProfScope prof("ln16", st);
  if (LPT == 16 && NV == 1) tcx_launch_pdl(ln16_kernel<16, 1>, grid, dim3(256), 0, st, a);
  else if (LPT == 16 && NV == 5) tcx_launch_pdl(ln16_kernel<16, 5>, grid, dim3(256), 0, st, a);
  else if (LPT == 32 && NV == 1) tcx_launch_pdl(ln16_kernel<32, 1>, grid, dim3(256), 0, st, a);
  else if (LPT == 32 && NV == 2) tcx_launch_pdl(ln16_kernel<32, 2>, grid, dim3(256), 0, st, a);
  else if (LPT == 32 && NV == 4) tcx_launch_pdl(ln16_kernel<32, 4>, grid, dim3(256), 0, st, a);
  else { tcx_set_error("ln16: unsupported width C=%d", a.C); return -1; }
  return tcx_check_launch("ln16");
}

int launch_mb_attention16(const Mb16Args& a, int groups, cudaStream_t st) {
  TCX_REQUIRE(a.heads == 8 && a.C % a.heads == 0, "mb_attn16: needs 8 heads (crpe window map {3:2,5:3,7:3})");
  const int Ch = a.C / a.heads, N = a.H * a.W;
  TCX_REQUIRE(Ch % 8 == 0 && Ch <= 64, "mb_attn16: head dim %d must be a multiple of 8 and <= 64", Ch);
  const size_t smem = ((size_t)N * Ch + (size_t)Ch * Ch + 49 * Ch + Ch + 256 + 2 * Ch + 1024) * sizeof(float) +
                      (size_t)2 * N * (Ch + 2) * sizeof(__half);
  TCX_REQUIRE(smem <= 200 * 1024, "mb_attn16: %d tokens x head dim %d does not fit in shared memory", N, Ch);
  dim3 grid(a.heads, a.B, groups);
  ProfScope prof("mb_fused16", st, (double)groups * a.B * N * a.C * 2.0 * 4.0);   // qkv read + output written (fp16)
  if (a.W % 4 == 0) {
    TCX_TRY(set_smem(mb_fused16_kernel<4>, smem, "mb_fused16"));
    tcx_launch_pdl(mb_fused16_kernel<4>, grid, dim3(MBF_THREADS), smem, st, a);
  } else if (a.W % 2 == 0) {
    TCX_TRY(set_smem(mb_fused16_kernel<2>, smem, "mb_fused16"));
    tcx_launch_pdl(mb_fused16_kernel<2>, grid, dim3(MBF_THREADS), smem, st, a);
  } else {
    TCX_TRY(set_smem(mb_fused16_kernel<1>, smem, "mb_fused16"));
    tcx_launch_pdl(mb_fused16_kernel<1>, grid, dim3(MBF_THREADS), smem, st, a);
  }
  return tcx_check_launch("mb_fused16");
}
