// Backward kernels of the Multi-Branch attention (FactorAtt_ConvRelPosEnc MSTr.py:852-886, ConvRelPosEnc :801-823) and of
// the depthwise position-encoding convolutions (ConvPosEnc :744-752).  fp32 tokens-major tensors with an explicit row
// pitch (q | k | v live side by side in one [M][3C] buffer); reductions over tokens are two-pass and ordered.
#include "bwd.cuh"

namespace {

constexpr int RC = 64, RL = 4;       // (64 columns x 4 row lanes) reduction block, as in bwd.cu

// depthwise K x K, stride 1, zero padding K/2, on the channel range this launch was given:
//   FLIP = false: y[p][c] = b[c] + sum_t w[c][t] x[p + off(t)][c]        (forward)
//   FLIP = true : y[p][c] =        sum_t w[c][t] x[p - off(t)][c]        (gradient w.r.t. the conv input)
// ADD: y += (accumulate into the destination)
template <int K, bool FLIP, bool ADD>
__global__ void __launch_bounds__(256) dwk_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                  const float* __restrict__ b, float* __restrict__ y, int ldy, int B, int H, int W, int Cg) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * H * W * Cg;
  if (idx >= total) return;
  const int c = (int)(idx % Cg);
  const long long p = idx / Cg;
  const int px = (int)(p % W), py = (int)((p / W) % H);
  const float* wc = w + (size_t)c * K * K;
  float acc = (b && !FLIP) ? b[c] : 0.f;
#pragma unroll
  for (int ky = 0; ky < K; ky++) {
    const int dy = FLIP ? K / 2 - ky : ky - K / 2;
    const int yy = py + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < K; kx++) {
      const int dx = FLIP ? K / 2 - kx : kx - K / 2;
      const int xx = px + dx;
      if (xx < 0 || xx >= W) continue;
      acc = fmaf(__ldg(wc + ky * K + kx), x[(p + (long long)dy * W + dx) * ldx + c], acc);
    }
  }
  if (ADD) y[p * ldy + c] += acc; else y[p * ldy + c] = acc;
}

// weight / bias gradient of the same conv: partials [blk][K*K + 1][Cg]
template <int K>
__global__ void __launch_bounds__(RC * RL) dwk_wgrad_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ x, int ldx,
                                                            int B, int H, int W, int Cg, int rows, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RL][RC];
  const int c = blockIdx.y * RC + threadIdx.x;
  const long long M = (long long)B * H * W;
  const long long r0 = (long long)blockIdx.x * rows;
  const long long r1 = r0 + rows < M ? r0 + rows : M;
  float acc[K * K + 1];
#pragma unroll
  for (int t = 0; t <= K * K; t++) acc[t] = 0.f;
  if (c < Cg) {
    for (long long r = r0 + threadIdx.y; r < r1; r += RL) {
      const int px = (int)(r % W), py = (int)((r / W) % H);
      const float gv = g[r * ldg + c];
      acc[K * K] += gv;
#pragma unroll
      for (int ky = 0; ky < K; ky++) {
        const int yy = py + ky - K / 2;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; kx++) {
          const int xx = px + kx - K / 2;
          if (xx < 0 || xx >= W) continue;
          acc[ky * K + kx] = fmaf(gv, x[(r + (long long)(ky - K / 2) * W + (kx - K / 2)) * ldx + c], acc[ky * K + kx]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t <= K * K; t++) {
    sm[threadIdx.y][threadIdx.x] = acc[t];
    __syncthreads();
    if (threadIdx.y == 0 && c < Cg) {
      float s = sm[0][threadIdx.x];
#pragma unroll
      for (int l = 1; l < RL; l++) s += sm[l][threadIdx.x];
      part[((size_t)blockIdx.x * (K * K + 1) + t) * Cg + c] = s;
    }
    __syncthreads();
  }
}
// partials [nblk][KK + 1][Cg] -> dw [Cg][KK], db [Cg] (db may be null)
__global__ void __launch_bounds__(256) dwk_fold_kernel(const float* __restrict__ part, int nblk, int KK, int Cg, float* __restrict__ dw,
                                                       float* __restrict__ db) {
  PDL_TOP();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int n = (KK + 1) * Cg;
  const float s = bwd_fold_sum(part, nblk, n, i, i < n);
  if (threadIdx.y != 0 || i >= n) return;
  const int t = i / Cg, c = i - t * Cg;
  if (t < KK) dw[(size_t)c * KK + t] = s; else if (db) db[c] = s;
}

// ---- the three crpe windows (3x3 on heads 0-1, 5x5 on heads 2-4, 7x7 on heads 5-7: channel ranges [0,c1), [c1,c2), [c2,C)) in ONE
// launch each: forward / input-gradient conv, and the filter / bias gradient with a single fold ----
struct Crpe3 {
  const float* w[3];     // [Cg_j][K_j*K_j]
  const float* b[3];     // [Cg_j] or null
  int c1, c2;            // channel boundaries
};
// ---- the three crpe windows in one launch, row-sweep form.  (The first version gave every output element its own thread, which
// loaded its K*K taps AND its K*K filter values — the latter at a stride of K*K floats across the lanes of a warp, 32 cache lines
// per load instruction: 28 us on the 14x14 maps against 13 us here.)  A thread owns (channel, image row):
// the channel's filter sits in registers, a K x K register window of x slides along the row, so an output costs K loads.
constexpr int SC = 32, SL = 8;
template <int K, bool FLIP, bool ADD>
__device__ __forceinline__ void dwk_sweep_row(const float* __restrict__ x, int ldx, const float* __restrict__ wc, float bias,
                                              float* __restrict__ y, int ldy, int H, int W, int c, int b, int py) {
  float wr[K][K], win[K][K];
#pragma unroll
  for (int a = 0; a < K; a++)
#pragma unroll
    for (int j = 0; j < K; j++) wr[a][j] = __ldg(wc + (FLIP ? (K - 1 - a) * K + (K - 1 - j) : a * K + j));
  const float* xr[K];
  bool rv[K];
#pragma unroll
  for (int a = 0; a < K; a++) {
    const int yy = py + a - K / 2;
    rv[a] = yy >= 0 && yy < H;
    xr[a] = x + ((size_t)(b * H + (rv[a] ? yy : py)) * W) * ldx + c;
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int xx = j - K / 2;
      win[a][j] = (rv[a] && xx >= 0 && xx < W) ? xr[a][(size_t)xx * ldx] : 0.f;
    }
  }
  float* yr = y + ((size_t)(b * H + py) * W) * ldy + c;
  for (int px = 0; px < W; px++) {
    const int xn = px + 1 + K / 2;
    float nx[K];
#pragma unroll
    for (int a = 0; a < K; a++) nx[a] = (rv[a] && xn < W) ? xr[a][(size_t)xn * ldx] : 0.f;
    float acc = bias;
#pragma unroll
    for (int a = 0; a < K; a++)
#pragma unroll
      for (int j = 0; j < K; j++) acc = fmaf(wr[a][j], win[a][j], acc);
    if (ADD) yr[(size_t)px * ldy] += acc; else yr[(size_t)px * ldy] = acc;
#pragma unroll
    for (int a = 0; a < K; a++) {
#pragma unroll
      for (int j = 0; j + 1 < K; j++) win[a][j] = win[a][j + 1];
      win[a][K - 1] = nx[a];
    }
  }
}
template <bool FLIP, bool ADD>
__global__ void __launch_bounds__(SC * SL) dwk3_sweep_kernel(const float* __restrict__ x, int ldx, Crpe3 f, float* __restrict__ y, int ldy,
                                                             int B, int H, int W, int C) {
  PDL_TOP();
  const int c = blockIdx.y * SC + threadIdx.x;
  const int row = blockIdx.x * SL + threadIdx.y;
  if (c >= C || row >= B * H) return;
  const int b = row / H, py = row - b * H;
  if (c < f.c1) {
    dwk_sweep_row<3, FLIP, ADD>(x, ldx, f.w[0] + (size_t)c * 9, (f.b[0] && !FLIP) ? f.b[0][c] : 0.f, y, ldy, H, W, c, b, py);
  } else if (c < f.c2) {
    dwk_sweep_row<5, FLIP, ADD>(x, ldx, f.w[1] + (size_t)(c - f.c1) * 25, (f.b[1] && !FLIP) ? f.b[1][c - f.c1] : 0.f, y, ldy, H, W, c, b, py);
  } else {
    dwk_sweep_row<7, FLIP, ADD>(x, ldx, f.w[2] + (size_t)(c - f.c2) * 49, (f.b[2] && !FLIP) ? f.b[2][c - f.c2] : 0.f, y, ldy, H, W, c, b, py);
  }
}

template <int K>
__device__ __forceinline__ void dwk_wgrad_rows(const float* __restrict__ g, int ldg, const float* __restrict__ x, int ldx, int H, int W,
                                               int c, long long r0, long long r1, float (&acc)[50]) {
  for (long long r = r0 + threadIdx.y; r < r1; r += RL) {
    const int px = (int)(r % W), py = (int)((r / W) % H);
    const float gv = g[r * ldg + c];
    acc[49] += gv;
#pragma unroll
    for (int ky = 0; ky < K; ky++) {
      const int yy = py + ky - K / 2;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < K; kx++) {
        const int xx = px + kx - K / 2;
        if (xx < 0 || xx >= W) continue;
        acc[ky * K + kx] = fmaf(gv, x[(r + (long long)(ky - K / 2) * W + (kx - K / 2)) * ldx + c], acc[ky * K + kx]);
      }
    }
  }
}
// ---- crpe filter gradients, row-sweep form.  (The first version gave every thread all K*K taps of its channel and a handful of
// pixels: 50 accumulators, 49 loads per pixel, a 50-round block fold and a partial buffer of M/16 * 50 * C floats — 128 us on the
// 28x28 maps.)  Partials [blk][50][C]: taps 0 .. K*K-1 of the channel's window, slot 49 = bias sum.  A thread owns ONE tap row ky at a time and sweeps whole image rows with a K-wide register window of x sliding
// along the row: 2 loads and K FMAs per pixel, K accumulators; (32 channels x 8 row lanes) per block, one block per 8 image rows,
// so the partial buffer is B*H/8 * 50 * C floats.  The association order is fixed (lane order, then block order in the fold).
constexpr int WC = 32, WL = 8;
template <int K>
__device__ __forceinline__ void dwk_wgrad_sweep(const float* __restrict__ g, int ldg, const float* __restrict__ x, int ldx, int H, int W,
                                                int c, int ky, int row0, int row1, float (&acc)[7], float& bsum) {
  for (int row = row0 + (int)threadIdx.y; row < row1; row += WL) {
    const int b = row / H, py = row - b * H;
    const int yy = py + ky - K / 2;
    if (yy < 0 || yy >= H) continue;
    const float* grow = g + (size_t)row * W * ldg + c;
    const float* xrow = x + (size_t)(b * H + yy) * W * ldx + c;
    float win[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int xx = j - K / 2;
      win[j] = (xx >= 0 && xx < W) ? xrow[(size_t)xx * ldx] : 0.f;
    }
#pragma unroll 4
    for (int px = 0; px < W; px++) {
      const float gv = grow[(size_t)px * ldg];
      const int xn = px + 1 + K / 2;
      const float nx = xn < W ? xrow[(size_t)xn * ldx] : 0.f;
      if (ky == K / 2) bsum += gv;
#pragma unroll
      for (int j = 0; j < K; j++) acc[j] = fmaf(gv, win[j], acc[j]);
#pragma unroll
      for (int j = 0; j + 1 < K; j++) win[j] = win[j + 1];
      win[K - 1] = nx;
    }
  }
}
// partials [blk][50][C] in the layout of the kernel above (tap ky*K + kx of the channel's own window, slot 49 = bias sum); only
// the slots of the channel's window are written (the fold reads only those)
__global__ void __launch_bounds__(WC * WL) dwk3_wgrad_sweep_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ x,
                                                                   int ldx, int B, int H, int W, int C, int c1, int c2, int rows,
                                                                   float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[WL][8][WC];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.y * WC + tx;
  const int K = c < c1 ? 3 : (c < c2 ? 5 : 7);
  const int nrow = B * H;
  const int row0 = blockIdx.x * rows;
  const int row1 = row0 + rows < nrow ? row0 + rows : nrow;
  float bsum = 0.f;
  for (int ky = 0; ky < 7; ky++) {
    float acc[7];
#pragma unroll
    for (int j = 0; j < 7; j++) acc[j] = 0.f;
    const bool live = c < C && ky < K;
    if (live) {
      if (K == 3) dwk_wgrad_sweep<3>(g, ldg, x, ldx, H, W, c, ky, row0, row1, acc, bsum);
      else if (K == 5) dwk_wgrad_sweep<5>(g, ldg, x, ldx, H, W, c, ky, row0, row1, acc, bsum);
      else dwk_wgrad_sweep<7>(g, ldg, x, ldx, H, W, c, ky, row0, row1, acc, bsum);
    }
#pragma unroll
    for (int j = 0; j < 7; j++) sm[ty][j][tx] = acc[j];
    if (ky == 6) sm[ty][7][tx] = bsum;
    __syncthreads();
    if (ty < 7 && live && ty < K) {          // lane ty folds tap kx = ty of this tap row
      float s = sm[0][ty][tx];
#pragma unroll
      for (int l = 1; l < WL; l++) s += sm[l][ty][tx];
      part[((size_t)blockIdx.x * 50 + ky * K + ty) * C + c] = s;
    }
    if (ky == 6 && ty == 7 && c < C) {
      float s = sm[0][7][tx];
#pragma unroll
      for (int l = 1; l < WL; l++) s += sm[l][7][tx];
      part[((size_t)blockIdx.x * 50 + 49) * C + c] = s;
    }
    __syncthreads();
  }
}
struct Crpe3Out {
  float* dw[3];
  float* db[3];
  int c1, c2;
};
__global__ void __launch_bounds__(256) dwk3_fold_kernel(const float* __restrict__ part, int nblk, int C, Crpe3Out o) {
  PDL_TOP();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int n = 50 * C;
  const int t = i / C, c = i - t * C;
  const int j = c < o.c1 ? 0 : (c < o.c2 ? 1 : 2);
  const int cc = c - (j == 0 ? 0 : (j == 1 ? o.c1 : o.c2));
  const int KK = j == 0 ? 9 : (j == 1 ? 25 : 49);
  const float s = bwd_fold_sum(part, nblk, n, i, i < n && (t == 49 || t < KK));     // slots outside the channel's window are never written
  if (threadIdx.y != 0 || i >= n) return;
  if (t == 49) { if (o.db[j]) o.db[j][cc] = s; }
  else if (t < KK) o.dw[j][(size_t)cc * KK + t] = s;
}

template <int K>
int dwk_launch(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int B, int H, int W, int Cg, bool flip, bool add,
               cudaStream_t st) {
  const long long total = (long long)B * H * W * Cg;
  if (total == 0) return 0;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (!flip && !add) tcx_launch_chain(dwk_kernel<K, false, false>, dim3(grid), dim3(256), 0, st, x, ldx, w, b, y, ldy, B, H, W, Cg);
  else if (!flip && add) tcx_launch_chain(dwk_kernel<K, false, true>, dim3(grid), dim3(256), 0, st, x, ldx, w, b, y, ldy, B, H, W, Cg);
  else if (flip && !add) tcx_launch_chain(dwk_kernel<K, true, false>, dim3(grid), dim3(256), 0, st, x, ldx, w, b, y, ldy, B, H, W, Cg);
  else tcx_launch_chain(dwk_kernel<K, true, true>, dim3(grid), dim3(256), 0, st, x, ldx, w, b, y, ldy, B, H, W, Cg);
  return tcx_check_launch("bwd_dwk");
}
template <int K>
int dwk_wgrad_launch(const float* g, int ldg, const float* x, int ldx, int B, int H, int W, int Cg, float* dw, float* db, float* part,
                     cudaStream_t st) {
  const long long M = (long long)B * H * W;
  if (M == 0 || Cg == 0) return 0;
  const int nblk = bwd_red_blocks(M);
  const int rows = (int)((M + nblk - 1) / nblk + RL - 1) / RL * RL;
  tcx_launch_chain(dwk_wgrad_kernel<K>, dim3(dim3(nblk, cdiv(Cg, RC))), dim3(dim3(RC, RL)), 0, st, g, ldg, x, ldx, B, H, W, Cg, rows, part);
  TCX_TRY(tcx_check_launch("bwd_dwk_wgrad"));
  tcx_launch_chain(dwk_fold_kernel, dim3(cdiv((K * K + 1) * Cg, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, K * K, Cg, dw, db);
  return tcx_check_launch("bwd_dwk_fold");
}

// out[b][i][j] = scale * sum_s part[b][s][i][j] inside the diagonal blocks of size Ch, 0 outside; outT = transposed copy
__global__ void __launch_bounds__(256) fold_mask_kernel(const float* __restrict__ part, int S, int R, int Ch, float scale,
                                                        float* __restrict__ out, float* __restrict__ outT) {
  PDL_TOP();
  const int n = R * R;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  const int b = blockIdx.y, i = idx / R, j = idx - i * R;
  float s = 0.f;
  if (i / Ch == j / Ch) {
    for (int k = 0; k < S; k++) s += part[((size_t)b * S + k) * n + idx];
    s *= scale;
  }
  out[(size_t)b * n + idx] = s;
  if (outT) outT[(size_t)b * n + (size_t)j * R + i] = s;
}

// dq = scale * dqfa + dxo * convv  -> dqkv[:, 0:C];   dconvv = dxo * q  (in place over convv)
__global__ void __launch_bounds__(256) mb_dq_kernel(const float* __restrict__ dxo, const float* __restrict__ dqfa, float* __restrict__ convv,
                                                    const float* __restrict__ q, int ldq, float scale, long long M, int C,
                                                    float* __restrict__ dq, int ldo) {
  PDL_TOP();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= M * C) return;
  const long long r = idx / C;
  const int c = (int)(idx - r * C);
  const float g = dxo[idx];
  dq[r * ldo + c] = fmaf(scale, dqfa[idx], g * convv[idx]);
  convv[idx] = g * q[r * ldq + c];
}

constexpr int CHUNK = 128;
// column softmax over the N tokens of each image (k rows of pitch ld): chunk partials, then P [B*N][C]
__global__ void __launch_bounds__(RC * RL) kstats32_kernel(const float* __restrict__ k, int ld, int N, int C, float* __restrict__ pm,
                                                           float* __restrict__ ps) {
  PDL_TOP();
  __shared__ float sm[RL][RC];
  __shared__ float bm[RC];
  const int c = blockIdx.z * RC + threadIdx.x;
  const int b = blockIdx.y, chunks = gridDim.x;
  const int r0 = blockIdx.x * CHUNK, r1 = min(r0 + CHUNK, N);
  const float* kb = k + (size_t)b * N * ld;
  float m = -INFINITY;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += RL) m = fmaxf(m, kb[(size_t)r * ld + c]);
  sm[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int l = 1; l < RL; l++) m = fmaxf(m, sm[l][threadIdx.x]);
    bm[threadIdx.x] = m;
  }
  __syncthreads();
  m = bm[threadIdx.x];
  float acc = 0.f;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += RL) acc += expf(kb[(size_t)r * ld + c] - m);
  __syncthreads();
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = sm[0][threadIdx.x];
    for (int l = 1; l < RL; l++) s += sm[l][threadIdx.x];
    pm[((size_t)b * chunks + blockIdx.x) * C + c] = m;
    ps[((size_t)b * chunks + blockIdx.x) * C + c] = s;
  }
}
__global__ void __launch_bounds__(RC * RL) ksoftmax32_kernel(const float* __restrict__ k, int ld, int N, int C, const float* __restrict__ pm,
                                                             const float* __restrict__ ps, float* __restrict__ P) {
  PDL_TOP();
  const int c = blockIdx.z * RC + threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.y, chunks = gridDim.x;
  float m = -INFINITY;
  for (int j = 0; j < chunks; j++) m = fmaxf(m, pm[((size_t)b * chunks + j) * C + c]);
  float sum = 0.f;
  for (int j = 0; j < chunks; j++) sum += ps[((size_t)b * chunks + j) * C + c] * expf(pm[((size_t)b * chunks + j) * C + c] - m);
  const float inv = 1.0f / sum;
  const int r0 = blockIdx.x * CHUNK, r1 = min(r0 + CHUNK, N);
  for (int r = r0 + threadIdx.y; r < r1; r += RL) {
    const size_t row = (size_t)b * N + r;
    P[row * C + c] = expf(k[row * ld + c] - m) * inv;
  }
}

}  // namespace

int launch_bwd_dwk(int K, const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int B, int H, int W, int Cg, int flip,
                   int add, cudaStream_t st) {
  if (K == 3) return dwk_launch<3>(x, ldx, w, b, y, ldy, B, H, W, Cg, flip, add, st);
  if (K == 5) return dwk_launch<5>(x, ldx, w, b, y, ldy, B, H, W, Cg, flip, add, st);
  if (K == 7) return dwk_launch<7>(x, ldx, w, b, y, ldy, B, H, W, Cg, flip, add, st);
  tcx_set_error("bwd_dwk: window %d not built (3, 5, 7)", K);
  return -1;
}
size_t bwd_dwk_wgrad_part_floats(int K, long long M, int Cg) { return (size_t)bwd_red_blocks(M) * (K * K + 1) * Cg; }
int launch_bwd_dwk_wgrad(int K, const float* g, int ldg, const float* x, int ldx, int B, int H, int W, int Cg, float* dw, float* db,
                         float* part, cudaStream_t st) {
  if (K == 3) return dwk_wgrad_launch<3>(g, ldg, x, ldx, B, H, W, Cg, dw, db, part, st);
  if (K == 5) return dwk_wgrad_launch<5>(g, ldg, x, ldx, B, H, W, Cg, dw, db, part, st);
  if (K == 7) return dwk_wgrad_launch<7>(g, ldg, x, ldx, B, H, W, Cg, dw, db, part, st);
  tcx_set_error("bwd_dwk_wgrad: window %d not built (3, 5, 7)", K);
  return -1;
}
int launch_bwd_dwk3(const float* x, int ldx, const float* const* w, const float* const* b, int c1, int c2, float* y, int ldy, int B, int H,
                    int W, int C, int flip, int add, cudaStream_t st) {
  const long long total = (long long)B * H * W * C;
  if (total == 0) return 0;
  Crpe3 f;
  for (int j = 0; j < 3; j++) { f.w[j] = w[j]; f.b[j] = b ? b[j] : nullptr; }
  f.c1 = c1; f.c2 = c2;
  const dim3 grid(cdiv(B * H, SL), cdiv(C, SC)), block(SC, SL);
  if (!flip && !add) tcx_launch_chain(dwk3_sweep_kernel<false, false>, dim3(grid), dim3(block), 0, st, x, ldx, f, y, ldy, B, H, W, C);
  else if (!flip && add) tcx_launch_chain(dwk3_sweep_kernel<false, true>, dim3(grid), dim3(block), 0, st, x, ldx, f, y, ldy, B, H, W, C);
  else if (flip && !add) tcx_launch_chain(dwk3_sweep_kernel<true, false>, dim3(grid), dim3(block), 0, st, x, ldx, f, y, ldy, B, H, W, C);
  else tcx_launch_chain(dwk3_sweep_kernel<true, true>, dim3(grid), dim3(block), 0, st, x, ldx, f, y, ldy, B, H, W, C);
  return tcx_check_launch("bwd_dwk3");
}
size_t bwd_dwk3_wgrad_part_floats(long long M, int C) { return (size_t)bwd_red_blocks(M) * 50 * C; }
int launch_bwd_dwk3_wgrad(const float* g, int ldg, const float* x, int ldx, int B, int H, int W, int C, int c1, int c2, float* const* dw,
                          float* const* db, float* part, cudaStream_t st) {
  const long long M = (long long)B * H * W;
  if (M == 0 || C == 0) return 0;
  // image rows per block: 8 (one per lane), more only if that would exceed the partial buffer sized by bwd_dwk3_wgrad_part_floats
  const int nrow = B * H, cap = bwd_red_blocks(M);
  int rows = WL;
  while (cdiv(nrow, rows) > cap) rows += WL;
  const int nblk = cdiv(nrow, rows);
  tcx_launch_chain(dwk3_wgrad_sweep_kernel, dim3(dim3(nblk, cdiv(C, WC))), dim3(dim3(WC, WL)), 0, st, g, ldg, x, ldx, B, H, W, C, c1, c2, rows, part);
  TCX_TRY(tcx_check_launch("bwd_dwk3_wgrad"));
  Crpe3Out o;
  for (int j = 0; j < 3; j++) { o.dw[j] = dw[j]; o.db[j] = db[j]; }
  o.c1 = c1; o.c2 = c2;
  tcx_launch_chain(dwk3_fold_kernel, dim3(cdiv(50 * C, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, C, o);
  return tcx_check_launch("bwd_dwk3_fold");
}
int launch_bwd_fold_mask(const float* part, int batch, int S, int R, int Ch, float scale, float* out, float* outT, cudaStream_t st) {
  if (batch == 0 || R == 0) return 0;
  tcx_launch_chain(fold_mask_kernel, dim3(dim3(cdiv(R * R, 256), batch)), dim3(256), 0, st, part, S, R, Ch, scale, out, outT);
  return tcx_check_launch("bwd_fold_mask");
}
int launch_mb_bwd_dq(const float* dxo, const float* dqfa, float* convv, const float* q, int ldq, float scale, long long M, int C, float* dq,
                     int ldo, cudaStream_t st) {
  if (M == 0) return 0;
  tcx_launch_chain(mb_dq_kernel, dim3((unsigned)((M * C + 255) / 256)), dim3(256), 0, st, dxo, dqfa, convv, q, ldq, scale, M, C, dq, ldo);
  return tcx_check_launch("mb_bwd_dq");
}
int launch_bwd_ksoftmax32(const float* k, int ld, int B, int N, int C, float* pm, float* ps, float* P, cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  const dim3 grid(cdiv(N, CHUNK), B, cdiv(C, RC)), block(RC, RL);
  tcx_launch_chain(kstats32_kernel, dim3(grid), dim3(block), 0, st, k, ld, N, C, pm, ps);
  TCX_TRY(tcx_check_launch("bwd_kstats32"));
  tcx_launch_chain(ksoftmax32_kernel, dim3(grid), dim3(block), 0, st, k, ld, N, C, pm, ps, P);
  return tcx_check_launch("bwd_ksoftmax32");
}

// ---- row softmax forward / backward (bridge attention scores MSTr.py:2281-2284; channel softmax of the queries :124-128) ----
namespace {
// one warp per row of n elements (pitch ld): y = softmax(scale * x)
__global__ void __launch_bounds__(256) rowsoftmax_fwd_kernel(const float* __restrict__ x, int ld, long long M, int n, float scale,
                                                             float* __restrict__ y, int ldy) {
  PDL_TOP();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m) * scale;
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s += expf(fmaf(xr[c], scale, -m));
  const float inv = 1.0f / warp_sum(s);
  float* yr = y + row * ldy;
  for (int c = lane; c < n; c += 32) yr[c] = expf(fmaf(xr[c], scale, -m)) * inv;
}
// ds = scale * P (dP - sum_c P dP)   (ds may alias dP)
__global__ void __launch_bounds__(256) rowsoftmax_bwd_kernel(const float* __restrict__ P, int ldp, const float* dP, int ldd, long long M,
                                                             int n, float scale, float* ds, int lds) {
  PDL_TOP();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* pr = P + row * ldp;
  const float* dr = dP + row * ldd;
  float dot = 0.f;
  for (int c = lane; c < n; c += 32) dot = fmaf(pr[c], dr[c], dot);
  dot = warp_sum(dot);
  float* o = ds + row * lds;
  for (int c = lane; c < n; c += 32) o[c] = scale * pr[c] * (dr[c] - dot);
}
// out[(b*R + r)*ldo + c] = scale * sum_s part[((b*S + s)*R + r)*Cc + c]
__global__ void __launch_bounds__(256) fold_rows_kernel(const float* __restrict__ part, int S, int R, int Cc, float scale,
                                                        float* __restrict__ out, int ldo) {
  PDL_TOP();
  const int n = R * Cc;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  const int b = blockIdx.y, r = idx / Cc, c = idx - r * Cc;
  float s = 0.f;
  for (int k = 0; k < S; k++) s += part[((size_t)b * S + k) * n + idx];
  out[((size_t)b * R + r) * ldo + c] = s * scale;
}
}  // namespace

int launch_bwd_rowsoftmax_fwd(const float* x, int ld, long long M, int n, float scale, float* y, int ldy, cudaStream_t st) {
  if (M == 0) return 0;
  tcx_launch_chain(rowsoftmax_fwd_kernel, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, x, ld, M, n, scale, y, ldy);
  return tcx_check_launch("bwd_rowsoftmax_fwd");
}
int launch_bwd_rowsoftmax_bwd(const float* P, int ldp, const float* dP, int ldd, long long M, int n, float scale, float* ds, int lds,
                              cudaStream_t st) {
  if (M == 0) return 0;
  tcx_launch_chain(rowsoftmax_bwd_kernel, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, P, ldp, dP, ldd, M, n, scale, ds, lds);
  return tcx_check_launch("bwd_rowsoftmax_bwd");
}
int launch_bwd_fold_rows(const float* part, int batch, int S, int R, int Cc, float scale, float* out, int ldo, cudaStream_t st) {
  if (batch == 0 || R == 0) return 0;
  tcx_launch_chain(fold_rows_kernel, dim3(dim3(cdiv(R * Cc, 256), batch)), dim3(256), 0, st, part, S, R, Cc, scale, out, ldo);
  return tcx_check_launch("bwd_fold_rows");
}
