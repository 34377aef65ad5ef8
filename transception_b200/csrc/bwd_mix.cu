// Fused backward of the Mix-FFN middle (MixFFN_skip.forward MSTr.py:59: a = GELU(LN(u)), u = dw3x3(h) + b + h) and of plain
// LayerNorm: the round-1 path read every [tokens][C] tensor three to four times (row pass, column pass, flip + conv as input
// gradient, weight-gradient pass) in seven launches; here
//   ln_bwd_fused  : ONE pass over (u, dz): per-row statistics, du, optionally a = GELU(LN(u)) in fp32 (the TF32 operand of fc2's
//                   weight gradient), and the per-column sums of d gamma / d beta carried in registers across the rows of a warp;
//   dw_bwd_fused  : ONE pass over (du, h): dh = du + conv^T(du) (input gradient incl. the skip) and the 9 + 1 per-channel filter /
//                   bias sums, 4 channels per thread with 16-byte accesses;
// each leaves one partial per block, folded in block order by the existing fold kernels (bit-reproducible, no atomics).
#include "bwd.cuh"

namespace {

constexpr int LF_WARPS = 8;
constexpr int LF_MAX_BLOCKS = 296;

__device__ __forceinline__ void gelu_parts(float z, float& gelu, float& dgelu) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * z * z);
  gelu = z * cdf;
  dgelu = fmaf(z, pdf, cdf);
}

// C <= 64: half a warp per row (16 lanes x float4), two rows per warp — the 64-channel maps (stage 1, bridge tokens, the last
// decoder layer: 50 176 to 97 216 rows) would leave half of every warp idle in the kernel below.
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <bool GELU>
__global__ void __launch_bounds__(LF_WARPS * 32) ln_bwd_fused64_kernel(const float* __restrict__ u, const float* __restrict__ dz,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      float eps, float* __restrict__ du, float* __restrict__ act,
                                                                      const float* __restrict__ dres, long long M, int C,
                                                                      float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[LF_WARPS * 2][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hl = lane & 15, hw = lane >> 4;
  const float invC = 1.0f / (float)C;
  const int c = hl * 4;
  const bool live = c < C;
  const float4 g4 = live ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b4 = (GELU && live) ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
  float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * LF_WARPS * 2;
  // both halves of a warp run the same number of iterations (full-mask shuffles): out-of-range rows compute on zeros
  // the next row's operands are loaded before the current row is reduced (one-deep software pipeline: the two shuffle reductions
  // of a row no longer wait for HBM)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  long long base = (long long)blockIdx.x * LF_WARPS * 2 + warp * 2;
  float4 nx4 = zero4, nd4 = zero4;
  if (base + hw < M && live) {
    nx4 = *reinterpret_cast<const float4*>(u + (base + hw) * C + c);
    nd4 = *reinterpret_cast<const float4*>(dz + (base + hw) * C + c);
  }
  for (; base < M; base += stride) {
    const long long row = base + hw;
    const bool rl = row < M && live;
    const float4 x4 = nx4, d4 = nd4;
    nx4 = zero4; nd4 = zero4;
    if (row + stride < M && live) {
      nx4 = *reinterpret_cast<const float4*>(u + (row + stride) * C + c);
      nd4 = *reinterpret_cast<const float4*>(dz + (row + stride) * C + c);
    }
    const float mean = half_sum((x4.x + x4.y) + (x4.z + x4.w)) * invC;
    float q = 0.f;
    if (live) {
      const float a = x4.x - mean, b = x4.y - mean, cc = x4.z - mean, d = x4.w - mean;
      q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(cc, cc, q); q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(half_sum(q) * invC + eps);
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
    float xh[4], g[4], ge[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      xh[k] = (xv[k] - mean) * rstd;
      float gg = dv[k];
      ge[k] = 0.f;
      if (GELU) {
        float dg;
        gelu_parts(fmaf(xh[k], gv[k], bv[k]), ge[k], dg);
        gg *= dg;
      }
      g[k] = rl ? gg : 0.f;
      const float dxh = g[k] * gv[k];
      s1 += dxh;
      s2 = fmaf(dxh, xh[k], s2);
    }
    s1 = half_sum(s1) * invC;
    s2 = half_sum(s2) * invC;
    if (rl) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        o[k] = rstd * (g[k] * gv[k] - s1 - xh[k] * s2);
        ag[k] = fmaf(g[k], xh[k], ag[k]);
        ab[k] += g[k];
      }
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + row * C + c);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      *reinterpret_cast<float4*>(du + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (GELU && act) *reinterpret_cast<float4*>(act + row * C + c) = make_float4(ge[0], ge[1], ge[2], ge[3]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    sm[warp * 2 + hw][0][c + k] = ag[k];
    sm[warp * 2 + hw][1][c + k] = ab[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += LF_WARPS * 32) {
    const int t = i / C, cc = i - t * C;
    float a = sm[0][t][cc];
#pragma unroll
    for (int w = 1; w < LF_WARPS * 2; w++) a += sm[w][t][cc];
    part[((size_t)blockIdx.x * 2 + t) * C + cc] = a;
  }
}

// One warp per row, NV float4 chunks per lane (columns lane*4 + 128*j); rows are dealt round-robin over all warps of the grid.
// part: [gridDim.x][2][C] = (sum g*xhat | sum g) of the rows this block handled.
template <int NV, bool GELU>
__global__ void __launch_bounds__(LF_WARPS * 32) ln_bwd_fused_kernel(const float* __restrict__ u, const float* __restrict__ dz,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                    float* __restrict__ du, float* __restrict__ act,
                                                                    const float* __restrict__ dres, long long M, int C,
                                                                    float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[LF_WARPS][2][NV * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invC = 1.0f / (float)C;
  float4 g4[NV], b4[NV];
  bool live[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const int c = lane * 4 + 128 * j;
    live[j] = c < C;
    g4[j] = live[j] ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    b4[j] = (GELU && live[j]) ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float ag[NV][4], ab[NV][4];
#pragma unroll
  for (int j = 0; j < NV; j++)
#pragma unroll
    for (int q = 0; q < 4; q++) { ag[j][q] = 0.f; ab[j][q] = 0.f; }

  const long long stride = (long long)gridDim.x * LF_WARPS;
  // one-deep software pipeline over rows: the next row's operands are in flight while this row is reduced
  float4 nx4[NV], nd4[NV];
  long long row = (long long)blockIdx.x * LF_WARPS + warp;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const int c = lane * 4 + 128 * j;
    const bool ld = live[j] && row < M;
    nx4[j] = ld ? *reinterpret_cast<const float4*>(u + row * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    nd4[j] = ld ? *reinterpret_cast<const float4*>(dz + row * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (; row < M; row += stride) {
    float4 x4[NV], d4[NV];
    float s = 0.f;
    const long long nrow = row + stride;
#pragma unroll
    for (int j = 0; j < NV; j++) {
      const int c = lane * 4 + 128 * j;
      x4[j] = nx4[j]; d4[j] = nd4[j];
      const bool ld = live[j] && nrow < M;
      nx4[j] = ld ? *reinterpret_cast<const float4*>(u + nrow * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      nd4[j] = ld ? *reinterpret_cast<const float4*>(dz + nrow * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (x4[j].x + x4[j].y) + (x4[j].z + x4[j].w);
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
      if (!live[j]) continue;
      const float a = x4[j].x - mean, b = x4[j].y - mean, c = x4[j].z - mean, d = x4[j].w - mean;
      q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    float xh[NV][4], g[NV][4], ge[NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
      const float xv[4] = {x4[j].x, x4[j].y, x4[j].z, x4[j].w}, dv[4] = {d4[j].x, d4[j].y, d4[j].z, d4[j].w};
      const float gv[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w}, bv[4] = {b4[j].x, b4[j].y, b4[j].z, b4[j].w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        xh[j][k] = (xv[k] - mean) * rstd;
        float gg = dv[k];
        ge[j][k] = 0.f;
        if (GELU) {
          float dg;
          gelu_parts(fmaf(xh[j][k], gv[k], bv[k]), ge[j][k], dg);
          gg *= dg;
        }
        g[j][k] = live[j] ? gg : 0.f;
        const float dxh = g[j][k] * gv[k];
        s1 += dxh;
        s2 = fmaf(dxh, xh[j][k], s2);
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int j = 0; j < NV; j++) {
      if (!live[j]) continue;
      const int c = lane * 4 + 128 * j;
      const float gv[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        o[k] = rstd * (g[j][k] * gv[k] - s1 - xh[j][k] * s2);
        ag[j][k] = fmaf(g[j][k], xh[j][k], ag[j][k]);
        ab[j][k] += g[j][k];
      }
      if (dres) {        // gradient arriving over the residual connection around this LayerNorm
        const float4 r = *reinterpret_cast<const float4*>(dres + row * C + c);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      *reinterpret_cast<float4*>(du + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (GELU && act) *reinterpret_cast<float4*>(act + row * C + c) = make_float4(ge[j][0], ge[j][1], ge[j][2], ge[j][3]);
    }
  }
  // block partial: the eight warps' register sums, added in warp order
#pragma unroll
  for (int j = 0; j < NV; j++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      sm[warp][0][lane * 4 + 128 * j + k] = ag[j][k];
      sm[warp][1][lane * 4 + 128 * j + k] = ab[j][k];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += LF_WARPS * 32) {
    const int t = i / C, c = i - t * C;
    float a = sm[0][t][c];
#pragma unroll
    for (int w = 1; w < LF_WARPS; w++) a += sm[w][t][c];
    part[((size_t)blockIdx.x * 2 + t) * C + c] = a;
  }
}

// Wide rows (512 < C <= 2048: the Mix-FFN hidden width of the 320- and 512-channel maps, 784 to 3 136 rows): one BLOCK per row,
// a thread owns the float4 chunks (t + 256 j) * 4 of every row its block handles, so the column sums of d gamma / d beta stay in
// its registers and are written as the block partial without any shared-memory fold; the three row reductions (mean, variance,
// the two LayerNorm-backward means) go through 8 per-warp partials added in warp order.
constexpr int LW_T = 256, LW_NV = 2;
__device__ __forceinline__ float lw_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
#pragma unroll
  for (int w = 1; w < LW_T / 32; w++) s += red[w];
  return s;
}
template <bool GELU>
__global__ void __launch_bounds__(LW_T) ln_bwd_wide_kernel(const float* __restrict__ u, const float* __restrict__ dz,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                           float* __restrict__ du, float* __restrict__ act, const float* __restrict__ dres,
                                                           long long M, int C, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float red[4][LW_T / 32];
  const float invC = 1.0f / (float)C;
  float4 g4[LW_NV], b4[LW_NV];
  bool live[LW_NV];
  float ag[LW_NV][4], ab[LW_NV][4];
#pragma unroll
  for (int j = 0; j < LW_NV; j++) {
    const int c = (threadIdx.x + LW_T * j) * 4;
    live[j] = c < C;
    g4[j] = live[j] ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    b4[j] = (GELU && live[j]) ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; k++) { ag[j][k] = 0.f; ab[j][k] = 0.f; }
  }
  for (long long row = blockIdx.x; row < M; row += gridDim.x) {
    float4 x4[LW_NV], d4[LW_NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LW_NV; j++) {
      const int c = (threadIdx.x + LW_T * j) * 4;
      x4[j] = live[j] ? *reinterpret_cast<const float4*>(u + row * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      d4[j] = live[j] ? *reinterpret_cast<const float4*>(dz + row * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (x4[j].x + x4[j].y) + (x4[j].z + x4[j].w);
    }
    const float mean = lw_block_sum(s, red[0]) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LW_NV; j++) {
      if (!live[j]) continue;
      const float a = x4[j].x - mean, b = x4[j].y - mean, c = x4[j].z - mean, d = x4[j].w - mean;
      q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(lw_block_sum(q, red[1]) * invC + eps);
    float xh[LW_NV][4], g[LW_NV][4], ge[LW_NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < LW_NV; j++) {
      const float xv[4] = {x4[j].x, x4[j].y, x4[j].z, x4[j].w}, dv[4] = {d4[j].x, d4[j].y, d4[j].z, d4[j].w};
      const float gv[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w}, bv[4] = {b4[j].x, b4[j].y, b4[j].z, b4[j].w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        xh[j][k] = live[j] ? (xv[k] - mean) * rstd : 0.f;
        float gg = dv[k];
        ge[j][k] = 0.f;
        if (GELU) {
          float dg;
          gelu_parts(fmaf(xh[j][k], gv[k], bv[k]), ge[j][k], dg);
          gg *= dg;
        }
        g[j][k] = live[j] ? gg : 0.f;
        const float dxh = g[j][k] * gv[k];
        s1 += dxh;
        s2 = fmaf(dxh, xh[j][k], s2);
      }
    }
    s1 = lw_block_sum(s1, red[2]) * invC;
    s2 = lw_block_sum(s2, red[3]) * invC;
#pragma unroll
    for (int j = 0; j < LW_NV; j++) {
      if (!live[j]) continue;
      const int c = (threadIdx.x + LW_T * j) * 4;
      const float gv[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        o[k] = rstd * (g[j][k] * gv[k] - s1 - xh[j][k] * s2);
        ag[j][k] = fmaf(g[j][k], xh[j][k], ag[j][k]);
        ab[j][k] += g[j][k];
      }
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + row * C + c);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      *reinterpret_cast<float4*>(du + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (GELU && act) *reinterpret_cast<float4*>(act + row * C + c) = make_float4(ge[j][0], ge[j][1], ge[j][2], ge[j][3]);
    }
  }
#pragma unroll
  for (int j = 0; j < LW_NV; j++) {
    if (!live[j]) continue;
    const int c = (threadIdx.x + LW_T * j) * 4;
    *reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * 2 + 0) * C + c) = make_float4(ag[j][0], ag[j][1], ag[j][2], ag[j][3]);
    *reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * 2 + 1) * C + c) = make_float4(ab[j][0], ab[j][1], ab[j][2], ab[j][3]);
  }
}

// depthwise 3x3 (stride 1, pad 1) backward on NHWC rows, C channels, in ONE pass:
//   dh[p][c] = du[p][c] + sum_t w[c][t] du[p - off(t)][c]                 (input gradient of u = conv(h) + b + h)
//   part[blk][t][c] = sum_{p in block} du[p][c] h[p + off(t)][c], t < 9;  part[blk][9][c] = sum du[p][c]
// Row sweep: a thread owns (2 channels, image row) and slides 3x3 register windows of du and h along the row — 6 loads per pixel
// (the first version, one thread per pixel x 4 channels, took 17), the channel's filter in registers; (32 channel pairs x 8 rows)
// per block, the block's (9 + 1) x 64 partial sums folded over its 8 row lanes in lane order.
constexpr int DS_C = 32, DS_L = 8;
__global__ void __launch_bounds__(DS_C * DS_L) dw_bwd_sweep_kernel(const float* __restrict__ du, const __half* __restrict__ h,
                                                                   const float* __restrict__ w, float* __restrict__ dh, int B, int H, int W,
                                                                   int C, int rows, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[DS_L][10][DS_C * 2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.y * DS_C + tx) * 2;
  const bool cl = c < C;                       // C is even
  const int nrow = B * H;
  const int row0 = blockIdx.x * rows;
  const int row1 = row0 + rows < nrow ? row0 + rows : nrow;
  float2 wt[9];
  float2 acc[10];
#pragma unroll
  for (int t = 0; t < 9; t++) wt[t] = cl ? make_float2(__ldg(w + (size_t)c * 9 + t), __ldg(w + (size_t)(c + 1) * 9 + t)) : make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 10; t++) acc[t] = make_float2(0.f, 0.f);
  if (cl) {
    for (int row = row0 + ty; row < row1; row += DS_L) {
      const int b = row / H, py = row - b * H;
      const float* gr[3];
      const __half* hr[3];
      bool rv[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const int yy = py + a - 1;
        rv[a] = yy >= 0 && yy < H;
        const size_t base = ((size_t)(b * H + (rv[a] ? yy : py)) * W) * C + c;
        gr[a] = du + base;
        hr[a] = h + base;
      }
      auto ldg2 = [&](int a, int xx) {
        return (rv[a] && xx >= 0 && xx < W) ? *reinterpret_cast<const float2*>(gr[a] + (size_t)xx * C) : make_float2(0.f, 0.f);
      };
      auto ldh2 = [&](int a, int xx) {
        return (rv[a] && xx >= 0 && xx < W) ? __half22float2(*reinterpret_cast<const __half2*>(hr[a] + (size_t)xx * C)) : make_float2(0.f, 0.f);
      };
      float2 gd[3][3], hh[3][3];               // [row a][col j] = value at (py + a - 1, px + j - 1)
#pragma unroll
      for (int a = 0; a < 3; a++) {
        gd[a][0] = make_float2(0.f, 0.f); hh[a][0] = make_float2(0.f, 0.f);
        gd[a][1] = ldg2(a, 0); hh[a][1] = ldh2(a, 0);
        gd[a][2] = ldg2(a, 1); hh[a][2] = ldh2(a, 1);
      }
      float* orow = dh + ((size_t)row * W) * C + c;
      for (int px = 0; px < W; px++) {
        float2 ng[3], nh[3];
#pragma unroll
        for (int a = 0; a < 3; a++) { ng[a] = ldg2(a, px + 2); nh[a] = ldh2(a, px + 2); }
        const float2 g = gd[1][1];
        float2 o = g;
        acc[9].x += g.x; acc[9].y += g.y;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            const int t = ky * 3 + kx;
            const float2 gn = gd[2 - ky][2 - kx];          // du at p - off(t)
            o.x = fmaf(wt[t].x, gn.x, o.x); o.y = fmaf(wt[t].y, gn.y, o.y);
            acc[t].x = fmaf(g.x, hh[ky][kx].x, acc[t].x);  // h at p + off(t)
            acc[t].y = fmaf(g.y, hh[ky][kx].y, acc[t].y);
          }
        *reinterpret_cast<float2*>(orow + (size_t)px * C) = o;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          gd[a][0] = gd[a][1]; gd[a][1] = gd[a][2]; gd[a][2] = ng[a];
          hh[a][0] = hh[a][1]; hh[a][1] = hh[a][2]; hh[a][2] = nh[a];
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 10; t++) { sm[ty][t][tx * 2] = acc[t].x; sm[ty][t][tx * 2 + 1] = acc[t].y; }
  __syncthreads();
  for (int i = ty * DS_C + tx; i < 10 * DS_C * 2; i += DS_C * DS_L) {
    const int t = i / (DS_C * 2), cc = i - t * (DS_C * 2);
    const int col = blockIdx.y * DS_C * 2 + cc;
    if (col >= C) continue;
    float a = sm[0][t][cc];
#pragma unroll
    for (int l = 1; l < DS_L; l++) a += sm[l][t][cc];
    part[((size_t)blockIdx.x * 10 + t) * C + col] = a;
  }
}

__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n4) {
  PDL_TOP();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(y)[i];
  const float4 b = reinterpret_cast<const float4*>(x)[i];
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  reinterpret_cast<float4*>(y)[i] = a;
}

// out = ((s0 + s1) + s2) + ... over n <= 16 equally sized tensors (the gradients a parameter shared by the blocks of an encoder
// receives from each of them): one launch instead of n - 1 accumulation kernels, fixed order
struct SumSrcs {
  const float* p[16];
};
__global__ void __launch_bounds__(256) sum_tensors_kernel(SumSrcs s, int n, long long numel, float* __restrict__ out) {
  PDL_TOP();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= numel) return;
  float a = s.p[0][i];
  for (int k = 1; k < n; k++) a += s.p[k][i];
  out[i] = a;
}

}  // namespace

int launch_sum_tensors(const float* const* srcs, int n, long long numel, float* out, cudaStream_t st) {
  TCX_REQUIRE(n >= 1 && n <= 16, "sum_tensors: 1..16 sources (got %d)", n);
  if (numel == 0) return 0;
  SumSrcs s{};
  for (int k = 0; k < n; k++) { TCX_REQUIRE(srcs[k] != nullptr, "sum_tensors: source %d is null", k); s.p[k] = srcs[k]; }
  tcx_launch_chain(sum_tensors_kernel, dim3((unsigned)((numel + 255) / 256)), dim3(256), 0, st, s, n, numel, out);
  return tcx_check_launch("sum_tensors");
}

int launch_add_inplace(float* y, const float* x, long long n, cudaStream_t st) {
  TCX_REQUIRE(n % 4 == 0 && (((uintptr_t)y | (uintptr_t)x) & 15) == 0, "add_inplace: n %% 4 != 0 or unaligned");
  if (n == 0) return 0;
  tcx_launch_chain(add_inplace_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), 0, st, y, x, n / 4);
  return tcx_check_launch("add_inplace");
}

bool ln_bwd_fused_ok(long long M, int C) { return M > 0 && C % 4 == 0 && C >= 4 && C <= LW_T * LW_NV * 4; }
int ln_bwd_fused_blocks(long long M, int C) {
  long long nb = C > 512 ? (M + 1) / 2                            // wide rows: one block per row, >= 2 rows per block
                         : (M + LF_WARPS * 4 - 1) / (LF_WARPS * 4);       // >= 4 rows per warp
  if (nb > LF_MAX_BLOCKS) nb = LF_MAX_BLOCKS;
  if (nb < 1) nb = 1;
  return (int)nb;
}

// part: 2 * ln_bwd_fused_blocks(M) * C floats.  act (nullable, gelu only): fp32 GELU(LN(u)).  The column sums still need
// launch_bwd_ln_fold(part, nblk, C, dgamma, dbeta).
int launch_ln_bwd_fused(const float* u, const float* dz, const float* gamma, const float* beta, float eps, int gelu, float* du,
                        float* act, const float* dres, long long M, int C, float* part, cudaStream_t st) {
  TCX_REQUIRE(ln_bwd_fused_ok(M, C), "ln_bwd_fused: C must be a multiple of 4, <= 2048 (M=%lld C=%d)", M, C);
  TCX_REQUIRE(du != dz, "ln_bwd_fused: du may not alias dz");
  const int nblk = ln_bwd_fused_blocks(M, C);
  if (C > 512) {
    ProfScope prof("ln_bwd_fused", st, (double)M * C * (gelu && act ? 16.0 : 12.0));
    if (gelu) tcx_launch_chain(ln_bwd_wide_kernel<true>, dim3(nblk), dim3(LW_T), 0, st, u, dz, gamma, beta, eps, du, act, dres, M, C, part);
    else tcx_launch_chain(ln_bwd_wide_kernel<false>, dim3(nblk), dim3(LW_T), 0, st, u, dz, gamma, beta, eps, du, nullptr, dres, M, C, part);
    return tcx_check_launch("ln_bwd_wide");
  }
  ProfScope prof("ln_bwd_fused", st, (double)M * C * (gelu && act ? 16.0 : 12.0));
#define LNB(NV)                                                                                                         \
  do {                                                                                                                  \
    if (gelu) tcx_launch_chain(ln_bwd_fused_kernel<NV, true>, dim3(nblk), dim3(LF_WARPS * 32), 0, st, u, dz, gamma, beta, eps, du, act, dres, M, C, part); \
    else tcx_launch_chain(ln_bwd_fused_kernel<NV, false>, dim3(nblk), dim3(LF_WARPS * 32), 0, st, u, dz, gamma, beta, eps, du, nullptr, dres, M, C, part); \
  } while (0)
  if (C <= 64) {
    if (gelu) tcx_launch_chain(ln_bwd_fused64_kernel<true>, dim3(nblk), dim3(LF_WARPS * 32), 0, st, u, dz, gamma, beta, eps, du, act, dres, M, C, part);
    else tcx_launch_chain(ln_bwd_fused64_kernel<false>, dim3(nblk), dim3(LF_WARPS * 32), 0, st, u, dz, gamma, beta, eps, du, nullptr, dres, M, C, part);
  } else if (C <= 128) LNB(1);
  else if (C <= 256) LNB(2);
  else LNB(4);
#undef LNB
  return tcx_check_launch("ln_bwd_fused");
}

// partial blocks of launch_dw_bwd_fused for a B x H x W map: one block per DS_L image rows, more rows per block only when that
// would exceed `cap` blocks.  (M, C) sizing form kept for the workspace formulas: an upper bound for any H, W with H*W*B = M, W >= 2.
static int dw_sweep_rows(int nrow, int cap) {
  int rows = DS_L;
  while ((nrow + rows - 1) / rows > cap) rows += DS_L;
  return rows;
}
constexpr int DW_SWEEP_MAX_BLOCKS = 1184;
int dw_bwd_fused_blocks(long long M, int C) {
  (void)C;
  long long nb = (M / 2 + DS_L - 1) / DS_L;            // image rows <= M / 2 for W >= 2
  if (nb > DW_SWEEP_MAX_BLOCKS) nb = DW_SWEEP_MAX_BLOCKS;
  if (nb < 1) nb = 1;
  return (int)nb;
}

// part: 10 * dw_bwd_fused_blocks(M, C) * C floats; fold with launch_bwd_dw_fold(part, *nblk_out, C, dw, db).
int launch_dw_bwd_fused(const float* du, const __half* h, const float* w, float* dh, int B, int H, int W, int C, float* part,
                        cudaStream_t st, int* nblk_out) {
  const long long M = (long long)B * H * W;
  TCX_REQUIRE(C % 4 == 0 && M > 0 && M < (1ll << 31), "dw_bwd_fused: C must be a multiple of 4 (C=%d)", C);
  TCX_REQUIRE(du != dh, "dw_bwd_fused: dh may not alias du");
  const int nrow = B * H;
  const int rows = dw_sweep_rows(nrow, W >= 2 ? dw_bwd_fused_blocks(M, C) : 1);
  const int nblk = (nrow + rows - 1) / rows;
  const dim3 grid(nblk, (C + DS_C * 2 - 1) / (DS_C * 2));
  ProfScope prof("dw_bwd_fused", st, (double)M * C * 10.0);
  tcx_launch_chain(dw_bwd_sweep_kernel, dim3(grid), dim3(dim3(DS_C, DS_L)), 0, st, du, h, w, dh, B, H, W, C, rows, part);
  if (nblk_out) *nblk_out = nblk;
  return tcx_check_launch("dw_bwd_fused");
}
