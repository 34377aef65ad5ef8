// Training kernels of the encoder's convolutional glue: BatchNorm with batch statistics (+ Hardswish / silu_swish) forward
// and backward (DWConv2d_BN MSTr.py:355-362, Conv2d_BN :399-404, CoordAtt.bn1 :1331), the strided depthwise 3x3 of RIPM
// (input and weight gradients), and the pooling / gating of CoordAtt (:1322-1348).  NHWC fp32, rows = pixels.
// Reductions over pixels are two-pass and ordered (bit-reproducible).
#include <algorithm>
#include "bwd.cuh"

namespace {

constexpr int RC = 64, RL = 4;

__device__ __forceinline__ float act_fwd(float z, int act) { return apply_act(z, act); }
// derivative of the activation at z
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == ACT_HARDSWISH) return z <= -3.f ? 0.f : (z >= 3.f ? 1.f : (2.f * z + 3.f) * (1.f / 6.f));
  if (act == ACT_SILU_SWISH) {                 // t * min(SiLU(t + 3) / 6, 1)   (MSTr.py:1270-1286)
    const float u = z + 3.f, sg = sigmoidf_(u);
    const float s = u * sg * (1.f / 6.f);
    if (s >= 1.f) return 1.f;
    const float ds = (sg + u * sg * (1.f - sg)) * (1.f / 6.f);
    return fmaf(z, ds, s);
  }
  return 1.f;
}

template <int K>
__device__ __forceinline__ void fold_lanes(float (&acc)[K], float (*sm)[RC]) {
#pragma unroll
  for (int k = 0; k < K; k++) {
    sm[threadIdx.y][threadIdx.x] = acc[k];
    __syncthreads();
    if (threadIdx.y == 0) {
      float s = sm[0][threadIdx.x];
#pragma unroll
      for (int l = 1; l < RL; l++) s += sm[l][threadIdx.x];
      acc[k] = s;
    }
    __syncthreads();
  }
}

// ---- batch statistics in ONE pass over x + one fold (6 -> 3 launches per BatchNorm forward).  A block leaves, per column, the
// sums of d = x - shift and d^2 over its rows with shift = the column's value in the block's first row (close to the data, so the
// shifted sums lose nothing) and its row count; the fold combines the blocks' (count, mean, M2) with Chan's update in block order —
// a fixed association, and the same variance as the two-pass form up to fp32 rounding.
__global__ void __launch_bounds__(RC * RL) bn_stats_kernel(const float* __restrict__ x, long long M, int C, int rows, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RL][RC];
  const int c = blockIdx.y * RC + threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows, r1 = r0 + rows < M ? r0 + rows : M;
  float acc[2] = {0.f, 0.f};
  float shift = 0.f;
  if (c < C && r0 < r1) {
    shift = x[r0 * C + c];
    for (long long r = r0 + threadIdx.y; r < r1; r += RL) {
      const float d = x[r * C + c] - shift;
      acc[0] += d;
      acc[1] = fmaf(d, d, acc[1]);
    }
  }
  fold_lanes<2>(acc, sm);
  if (threadIdx.y == 0 && c < C) {
    float* p = part + (size_t)blockIdx.x * 3 * C;
    p[c] = acc[0]; p[C + c] = acc[1]; p[2 * C + c] = shift;
  }
}
// one thread row of 8 lanes per column: lane l walks blocks l, l+8, ... (Chan), then the eight lane results are combined in lane order
__global__ void __launch_bounds__(256) bn_stats_fold_kernel(const float* __restrict__ part, int nblk, int rows, long long M, int C, float eps,
                                                            float momentum, float* __restrict__ stat, float* __restrict__ rm,
                                                            float* __restrict__ rv) {
  PDL_TOP();
  __shared__ float sn[8][32], smean[8][32], sm2[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x, l = threadIdx.y;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  if (c < C) {
    for (int b = l; b < nblk; b += 8) {
      const long long r0 = (long long)b * rows;
      const long long cnt = (r0 + rows < M ? r0 + rows : M) - r0;
      if (cnt <= 0) continue;
      const float* p = part + (size_t)b * 3 * C;
      const float nb = (float)cnt, s1 = p[c], s2 = p[C + c], sh = p[2 * C + c];
      const float mb = sh + s1 / nb, m2b = s2 - s1 * s1 / nb;
      const float tot = n + nb, delta = mb - mean;
      mean += delta * (nb / tot);
      m2 += m2b + delta * delta * (n * nb / tot);
      n = tot;
    }
  }
  sn[l][threadIdx.x] = n; smean[l][threadIdx.x] = mean; sm2[l][threadIdx.x] = m2;
  __syncthreads();
  if (l != 0 || c >= C) return;
  n = sn[0][threadIdx.x]; mean = smean[0][threadIdx.x]; m2 = sm2[0][threadIdx.x];
  for (int k = 1; k < 8; k++) {
    const float nb = sn[k][threadIdx.x];
    if (nb <= 0.f) continue;
    const float tot = n + nb, delta = smean[k][threadIdx.x] - mean;
    mean += delta * (nb / tot);
    m2 += sm2[k][threadIdx.x] + delta * delta * (n * nb / tot);
    n = tot;
  }
  const float var = m2 / (float)M;
  stat[c] = mean;
  stat[C + c] = rsqrtf(var + eps);
  if (rm) {
    rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
    rv[c] = (1.f - momentum) * rv[c] + momentum * (M > 1 ? m2 / (float)(M - 1) : var);
  }
}
// stat[0][c] = mean, stat[1][c] = 1/sqrt(var + eps) (written by the fold above; running statistics as nn.BatchNorm2d: momentum,
// unbiased variance)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stat, const float* __restrict__ w,
                                                       const float* __restrict__ b, int act, long long n, int C, float* __restrict__ y) {
  PDL_TOP();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  y[i] = act_fwd(fmaf((x[i] - stat[c]) * stat[C + c], w[c], b[c]), act);
}
// partials of dw = sum g xhat, db = sum g with g = dy act'(z)
__global__ void __launch_bounds__(RC * RL) bn_bwd_cols_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ stat, const float* __restrict__ w,
                                                              const float* __restrict__ b, int act, long long M, int C, int rows,
                                                              float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RL][RC];
  const int c = blockIdx.y * RC + threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows, r1 = r0 + rows < M ? r0 + rows : M;
  float acc[2] = {0.f, 0.f};
  if (c < C) {
    const float mean = stat[c], inv = stat[C + c], wc = w[c], bc = b[c];
    for (long long r = r0 + threadIdx.y; r < r1; r += RL) {
      const float xh = (x[r * C + c] - mean) * inv;
      const float g = dy[r * C + c] * act_grad(fmaf(xh, wc, bc), act);
      acc[0] = fmaf(g, xh, acc[0]);
      acc[1] += g;
    }
  }
  fold_lanes<2>(acc, sm);
  if (threadIdx.y == 0 && c < C) {
    part[((size_t)blockIdx.x * 2 + 0) * C + c] = acc[0];
    part[((size_t)blockIdx.x * 2 + 1) * C + c] = acc[1];
  }
}
// dwdb = [dw | db] (2C floats);  dx = w inv (g - db/M - xhat dw/M)
__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stat,
                                                        const float* __restrict__ w, const float* __restrict__ b, int act,
                                                        const float* __restrict__ dwdb, long long M, int C, float* __restrict__ dx) {
  PDL_TOP();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= M * C) return;
  const int c = (int)(i % C);
  const float inv = stat[C + c], xh = (x[i] - stat[c]) * inv;
  const float g = dy[i] * act_grad(fmaf(xh, w[c], b[c]), act);
  const float invM = 1.f / (float)M;
  dx[i] = w[c] * inv * (g - dwdb[C + c] * invM - xh * dwdb[c] * invM);
}

// depthwise 3x3, pad 1, stride s: input gradient  dx[b,yi,xi,c] = sum_t w[c][t] dy[b,(yi+1-ky)/s,(xi+1-kx)/s,c]
__global__ void __launch_bounds__(256) dw3s_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, int B, int H, int W, int C,
                                                         int s, int Ho, int Wo, float* __restrict__ dx) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W * C) return;
  const int c = (int)(idx % C);
  long long p = idx / C;
  const int xi = (int)(p % W); p /= W;
  const int yi = (int)(p % H);
  const int b = (int)(p / H);
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    const int ty = yi + 1 - ky;
    if (ty < 0 || ty % s) continue;
    const int yo = ty / s;
    if (yo >= Ho) continue;
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const int tx = xi + 1 - kx;
      if (tx < 0 || tx % s) continue;
      const int xo = tx / s;
      if (xo >= Wo) continue;
      acc = fmaf(__ldg(w + c * 9 + ky * 3 + kx), dy[(((long long)b * Ho + yo) * Wo + xo) * C + c], acc);
    }
  }
  dx[idx] = acc;
}
// weight gradient partials [blk][9][C]: dw[c][t] = sum_{b,yo,xo} dy[b,yo,xo,c] x[b, yo s - 1 + ky, xo s - 1 + kx, c]
__global__ void __launch_bounds__(RC * RL) dw3s_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int H, int W,
                                                             int C, int s, int Ho, int Wo, int rows, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RL][RC];
  const int c = blockIdx.y * RC + threadIdx.x;
  const long long M = (long long)B * Ho * Wo;
  const long long r0 = (long long)blockIdx.x * rows, r1 = r0 + rows < M ? r0 + rows : M;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; t++) acc[t] = 0.f;
  if (c < C) {
    for (long long r = r0 + threadIdx.y; r < r1; r += RL) {
      const int xo = (int)(r % Wo), yo = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
      const float g = dy[r * C + c];
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int yi = yo * s - 1 + ky;
        if (yi < 0 || yi >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const int xi = xo * s - 1 + kx;
          if (xi < 0 || xi >= W) continue;
          acc[ky * 3 + kx] = fmaf(g, x[(((long long)b * H + yi) * W + xi) * C + c], acc[ky * 3 + kx]);
        }
      }
    }
  }
  fold_lanes<9>(acc, sm);
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int t = 0; t < 9; t++) part[((size_t)blockIdx.x * 9 + t) * C + c] = acc[t];
  }
}
__global__ void __launch_bounds__(256) dw3s_fold_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ dw) {
  PDL_TOP();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const float s = bwd_fold_sum(part, nblk, 9 * C, i, i < 9 * C);
  if (threadIdx.y != 0 || i >= 9 * C) return;
  const int t = i / C, c = i - t * C;
  dw[c * 9 + t] = s;
}

// CoordAtt pooling: y [B][H+W][C]: rows 0..H-1 = mean over W, rows H.. = mean over H.  One thread per output element.
__global__ void __launch_bounds__(256) coord_pool_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * (H + W) * C) return;
  const int c = (int)(idx % C);
  const int r = (int)((idx / C) % (H + W));
  const int b = (int)(idx / ((long long)C * (H + W)));
  const float* xb = x + (size_t)b * H * W * C + c;
  float s = 0.f;
  if (r < H) { for (int j = 0; j < W; j++) s += xb[((size_t)r * W + j) * C]; s /= (float)W; }
  else { const int j = r - H; for (int i = 0; i < H; i++) s += xb[((size_t)i * W + j) * C]; s /= (float)H; }
  y[idx] = s;
}
__global__ void __launch_bounds__(256) coord_pool_bwd_kernel(const float* __restrict__ dy, int B, int H, int W, int C, float* __restrict__ dx) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W * C) return;
  const int c = (int)(idx % C);
  long long p = idx / C;
  const int j = (int)(p % W); p /= W;
  const int i = (int)(p % H);
  const int b = (int)(p / H);
  const float* d = dy + (size_t)b * (H + W) * C + c;
  dx[idx] = d[(size_t)i * C] / (float)W + d[(size_t)(H + j) * C] / (float)H;
}
// gate: out = x * sigmoid(zw[b,w,c]) * sigmoid(zh[b,h,c]);  z [B][H+W][C] (rows 0..H-1 = zh, H.. = zw)
__global__ void __launch_bounds__(256) coord_gate_kernel(const float* __restrict__ x, const float* __restrict__ z, int B, int H, int W, int C,
                                                         const float* __restrict__ dout, float* __restrict__ out) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W * C) return;
  const int c = (int)(idx % C);
  long long p = idx / C;
  const int j = (int)(p % W); p /= W;
  const int i = (int)(p % H);
  const int b = (int)(p / H);
  const float* zb = z + (size_t)b * (H + W) * C + c;
  const float a = sigmoidf_(zb[(size_t)i * C]) * sigmoidf_(zb[(size_t)(H + j) * C]);
  out[idx] = (dout ? dout[idx] : x[idx]) * a;       // forward: x a ; backward dx: dout a
}
// dz[b,r,c]: r < H: sum_w dout x a_w * a_h (1 - a_h);  r >= H: sum_h dout x a_h * a_w (1 - a_w)
__global__ void __launch_bounds__(256) coord_gate_bwd_z_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                               const float* __restrict__ dout, int B, int H, int W, int C,
                                                               float* __restrict__ dz) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * (H + W) * C) return;
  const int c = (int)(idx % C);
  const int r = (int)((idx / C) % (H + W));
  const int b = (int)(idx / ((long long)C * (H + W)));
  const float* zb = z + (size_t)b * (H + W) * C + c;
  const size_t base = (size_t)b * H * W * C + c;
  const float am = sigmoidf_(zb[(size_t)r * C]);
  float s = 0.f;
  if (r < H) {
    for (int j = 0; j < W; j++) {
      const size_t q = base + ((size_t)r * W + j) * C;
      s = fmaf(dout[q] * x[q], sigmoidf_(zb[(size_t)(H + j) * C]), s);
    }
  } else {
    const int j = r - H;
    for (int i = 0; i < H; i++) {
      const size_t q = base + ((size_t)i * W + j) * C;
      s = fmaf(dout[q] * x[q], sigmoidf_(zb[(size_t)i * C]), s);
    }
  }
  dz[idx] = s * am * (1.f - am);
}

inline int rows_for(long long M, int nblk) { return (int)((M + nblk - 1) / nblk + RL - 1) / RL * RL; }

}  // namespace

// BatchNorm (batch statistics) + activation forward.  stat: 2*C floats (mean | 1/std) kept for backward; scratch: C * (2 + 2*blocks)
size_t bn_train_scratch_floats(long long M, int C) { return (size_t)C * (2 + 3 * (size_t)bwd_red_blocks(M)) + 64; }
int launch_bn_train_fwd(const float* x, const float* w, const float* b, float eps, float momentum, int act, float* y, float* stat, float* rm,
                        float* rv, long long M, int C, float* scratch, cudaStream_t st) {
  TCX_REQUIRE(M > 0 && C > 0, "bn_train_fwd: empty input");
  // two blocks per SM are enough to stream x; the fold walks nblk / 8 partials per lane with a dependent Chan update each
  const int nblk = std::min(bwd_red_blocks(M), 296);
  const int rows = rows_for(M, nblk);
  float* part = scratch + 2 * C;
  tcx_launch_chain(bn_stats_kernel, dim3(dim3(nblk, cdiv(C, RC))), dim3(dim3(RC, RL)), 0, st, x, M, C, rows, part);
  TCX_TRY(tcx_check_launch("bn_stats"));
  tcx_launch_chain(bn_stats_fold_kernel, dim3(cdiv(C, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, rows, M, C, eps, momentum, stat, rm, rv);
  TCX_TRY(tcx_check_launch("bn_stats_fold"));
  tcx_launch_chain(bn_apply_kernel, dim3((unsigned)((M * C + 255) / 256)), dim3(256), 0, st, x, stat, w, b, act, M * C, C, y);
  return tcx_check_launch("bn_apply");
}
// dw and db adjacent in memory (db == dw + C: the binding allocates them as one [2][C] tensor): the fold writes them in place
int launch_bn_train_bwd(const float* x, const float* dy, const float* stat, const float* w, const float* b, int act, float* dx, float* dw,
                        float* db, long long M, int C, float* scratch, cudaStream_t st) {
  const int nblk = bwd_red_blocks(M);
  const bool adjacent = db == dw + C;
  float* dwdb = adjacent ? dw : scratch; float* part = scratch + 2 * C;
  tcx_launch_chain(bn_bwd_cols_kernel, dim3(dim3(nblk, cdiv(C, RC))), dim3(dim3(RC, RL)), 0, st, x, dy, stat, w, b, act, M, C, rows_for(M, nblk), part);
  TCX_TRY(tcx_check_launch("bn_bwd_cols"));
  TCX_TRY(launch_bwd_fold(part, nblk, 2 * (long long)C, dwdb, st));       // partial layout [blk][2][C] -> [dw | db]
  tcx_launch_chain(bn_bwd_dx_kernel, dim3((unsigned)((M * C + 255) / 256)), dim3(256), 0, st, x, dy, stat, w, b, act, dwdb, M, C, dx);
  TCX_TRY(tcx_check_launch("bn_bwd_dx"));
  if (!adjacent)
    TCX_REQUIRE(cudaMemcpyAsync(dw, dwdb, sizeof(float) * C, cudaMemcpyDeviceToDevice, st) == cudaSuccess &&
                    cudaMemcpyAsync(db, dwdb + C, sizeof(float) * C, cudaMemcpyDeviceToDevice, st) == cudaSuccess,
                "bn_train_bwd: gradient copy failed");
  return 0;
}

size_t dw3s_scratch_floats(long long Mo, int C) { return 9 * (size_t)bwd_red_blocks(Mo) * C + 64; }
int launch_dw3s_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H, int W, int C, int stride,
                    float* scratch, cudaStream_t st) {
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  if (dx) {
    tcx_launch_chain(dw3s_dgrad_kernel, dim3((unsigned)(((long long)B * H * W * C + 255) / 256)), dim3(256), 0, st, dy, w, B, H, W, C, stride, Ho, Wo, dx);
    TCX_TRY(tcx_check_launch("dw3s_dgrad"));
  }
  if (dw) {
    const long long Mo = (long long)B * Ho * Wo;
    const int nblk = bwd_red_blocks(Mo);
    tcx_launch_chain(dw3s_wgrad_kernel, dim3(dim3(nblk, cdiv(C, RC))), dim3(dim3(RC, RL)), 0, st, dy, x, B, H, W, C, stride, Ho, Wo, rows_for(Mo, nblk), scratch);
    TCX_TRY(tcx_check_launch("dw3s_wgrad"));
    tcx_launch_chain(dw3s_fold_kernel, dim3(cdiv(9 * C, 32)), dim3(dim3(32, 8)), 0, st, scratch, nblk, C, dw);
    TCX_TRY(tcx_check_launch("dw3s_fold"));
  }
  return 0;
}

int launch_coord_pool(const float* x, int B, int H, int W, int C, float* y, cudaStream_t st) {
  tcx_launch_chain(coord_pool_kernel, dim3((unsigned)(((long long)B * (H + W) * C + 255) / 256)), dim3(256), 0, st, x, B, H, W, C, y);
  return tcx_check_launch("coord_pool");
}
int launch_coord_pool_bwd(const float* dy, int B, int H, int W, int C, float* dx, cudaStream_t st) {
  tcx_launch_chain(coord_pool_bwd_kernel, dim3((unsigned)(((long long)B * H * W * C + 255) / 256)), dim3(256), 0, st, dy, B, H, W, C, dx);
  return tcx_check_launch("coord_pool_bwd");
}
int launch_coord_gate(const float* x, const float* z, int B, int H, int W, int C, float* out, cudaStream_t st) {
  tcx_launch_chain(coord_gate_kernel, dim3((unsigned)(((long long)B * H * W * C + 255) / 256)), dim3(256), 0, st, x, z, B, H, W, C, nullptr, out);
  return tcx_check_launch("coord_gate");
}
int launch_coord_gate_bwd(const float* x, const float* z, const float* dout, int B, int H, int W, int C, float* dx, float* dz, cudaStream_t st) {
  tcx_launch_chain(coord_gate_kernel, dim3((unsigned)(((long long)B * H * W * C + 255) / 256)), dim3(256), 0, st, x, z, B, H, W, C, dout, dx);
  TCX_TRY(tcx_check_launch("coord_gate_dx"));
  tcx_launch_chain(coord_gate_bwd_z_kernel, dim3((unsigned)(((long long)B * (H + W) * C + 255) / 256)), dim3(256), 0, st, x, z, dout, B, H, W, C, dz);
  return tcx_check_launch("coord_gate_dz");
}
