// Backward kernels of the training row: column reductions, LayerNorm(+GELU) backward,
// depthwise-conv weight gradient, K-major packing for the weight-gradient GEMM.  See bwd.cuh for the contracts.
//
// Layout conventions: tokens-major [M][C] fp32 gradients, fp16 saved activations of the forward pipeline.  All token
// reductions use the (64 columns x 4 row lanes) block below: a warp reads 32 consecutive columns of one row (128 B),
// every block owns a contiguous run of rows and leaves one partial per column, a second kernel folds the partials in
// block order — no atomics, so gradients are bit-reproducible run to run.
#include "bwd.cuh"

namespace {

constexpr int RED_COLS = 64;     // columns per block
constexpr int RED_LANES = 4;     // row lanes per block
constexpr int RED_MAX_BLOCKS = 1184;  // 8 x 148 row blocks (x C/64 column groups): narrow maps (C = 64) still fill the SMs

inline int red_rows_per_block(long long M) {
  long long r = (M + RED_MAX_BLOCKS - 1) / RED_MAX_BLOCKS;
  if (r < 16) r = 16;
  return (int)((r + RED_LANES - 1) / RED_LANES * RED_LANES);
}

// fixed-order fold of the RED_LANES row lanes of a block; returns the block sum for column threadIdx.x (valid on lane row 0)
template <int K>
__device__ __forceinline__ void block_fold_lanes(float (&acc)[K], float (*sm)[K][RED_COLS]) {
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = 0; k < K; k++) sm[ty][k][tx] = acc[k];
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      float s = sm[0][k][tx];
#pragma unroll
      for (int l = 1; l < RED_LANES; l++) s += sm[l][k][tx];
      acc[k] = s;
    }
  }
}

__global__ void __launch_bounds__(RED_COLS * RED_LANES) bwd_colsum_kernel(const float* __restrict__ x, long long M, int C, int ld,
                                                                         int rows, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RED_LANES][1][RED_COLS];
  const int c = blockIdx.y * RED_COLS + threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows;
  const long long r1 = r0 + rows < M ? r0 + rows : M;
  float acc[1] = {0.f};
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += RED_LANES) acc[0] += x[r * ld + c];
  block_fold_lanes<1>(acc, sm);
  if (threadIdx.y == 0 && c < C) part[(size_t)blockIdx.x * C + c] = acc[0];
}

// out[i] = sum_s part[s*n + i]
__global__ void __launch_bounds__(256) bwd_fold_kernel(const float* __restrict__ part, int S, long long n, float* __restrict__ out) {
  PDL_TOP();
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  const float s = bwd_fold_sum(part, S, n, i, i < n);
  if (threadIdx.y == 0 && i < n) out[i] = s;
}

// d/dz GELU_erf(z) = Phi(z) + z phi(z)
__device__ __forceinline__ float gelu_grad(float z) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * z * z);
  return fmaf(z, pdf, cdf);
}

// one warp per row; the row (<= 8 KB) is re-read from L1 in each pass
template <bool GELU>
__global__ void __launch_bounds__(256) bwd_ln_rows_kernel(const float* __restrict__ u, const float* __restrict__ dz,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                          float* __restrict__ du, float* __restrict__ stats, long long M, int C) {
  PDL_TOP();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* ur = u + row * C;
  const float* dr = dz + row * C;
  const float invC = 1.0f / (float)C;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(ur + c);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) * invC;
  float q = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(ur + c);
    const float a = v.x - mean, b = v.y - mean, d = v.z - mean, e = v.w - mean;
    q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(d, d, q); q = fmaf(e, e, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * invC + eps);
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(ur + c);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + c);
    const float4 d4 = *reinterpret_cast<const float4*>(dr + c);
    const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (GELU) {
      const float4 b4 = *reinterpret_cast<const float4*>(beta + c);
      bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float xh = (xv[j] - mean) * rstd;
      float g = dv[j];
      if (GELU) g *= gelu_grad(fmaf(xh, gv[j], bv[j]));
      const float dxh = g * gv[j];
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
  }
  s1 = warp_sum(s1) * invC;
  s2 = warp_sum(s2) * invC;
  float* our = du + row * C;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(ur + c);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + c);
    const float4 d4 = *reinterpret_cast<const float4*>(dr + c);
    const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (GELU) {
      const float4 b4 = *reinterpret_cast<const float4*>(beta + c);
      bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float xh = (xv[j] - mean) * rstd;
      float g = dv[j];
      if (GELU) g *= gelu_grad(fmaf(xh, gv[j], bv[j]));
      o[j] = rstd * (g * gv[j] - s1 - xh * s2);
    }
    *reinterpret_cast<float4*>(our + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// dgamma[c] = sum_m g xhat, dbeta[c] = sum_m g with g = dz (* GELU'(LN(u))): block partials [blk][2][C]
template <bool GELU>
__global__ void __launch_bounds__(RED_COLS * RED_LANES) bwd_ln_cols_kernel(const float* __restrict__ u, const float* __restrict__ dz,
                                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                          const float* __restrict__ stats, long long M, int C, int rows,
                                                                          float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RED_LANES][2][RED_COLS];
  const int c = blockIdx.y * RED_COLS + threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows;
  const long long r1 = r0 + rows < M ? r0 + rows : M;
  float acc[2] = {0.f, 0.f};
  if (c < C) {
    const float gm = gamma[c], bt = GELU ? beta[c] : 0.f;
    for (long long r = r0 + threadIdx.y; r < r1; r += RED_LANES) {
      const float2 st = *reinterpret_cast<const float2*>(stats + 2 * r);
      const float xh = (u[r * C + c] - st.x) * st.y;
      float g = dz[r * C + c];
      if (GELU) g *= gelu_grad(fmaf(xh, gm, bt));
      acc[0] = fmaf(g, xh, acc[0]);
      acc[1] += g;
    }
  }
  block_fold_lanes<2>(acc, sm);
  if (threadIdx.y == 0 && c < C) {
    part[((size_t)blockIdx.x * 2 + 0) * C + c] = acc[0];
    part[((size_t)blockIdx.x * 2 + 1) * C + c] = acc[1];
  }
}
// partials [nblk][2][C] -> dgamma, dbeta
__global__ void __launch_bounds__(256) bwd_ln_fold_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta) {
  PDL_TOP();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const float s = bwd_fold_sum(part, nblk, 2 * C, i, i < 2 * C);
  if (threadIdx.y != 0 || i >= 2 * C) return;
  if (i < C) dgamma[i] = s; else dbeta[i - C] = s;
}

// depthwise 3x3 weight gradient: partials [blk][10][C] (9 taps + bias)
__global__ void __launch_bounds__(RED_COLS * RED_LANES) bwd_dw_wgrad_kernel(const float* __restrict__ du, const __half* __restrict__ h,
                                                                           int B, int H, int W, int C, int rows, float* __restrict__ part) {
  PDL_TOP();
  __shared__ float sm[RED_LANES][10][RED_COLS];
  const int c = blockIdx.y * RED_COLS + threadIdx.x;
  const long long M = (long long)B * H * W;
  const long long r0 = (long long)blockIdx.x * rows;
  const long long r1 = r0 + rows < M ? r0 + rows : M;
  float acc[10];
#pragma unroll
  for (int t = 0; t < 10; t++) acc[t] = 0.f;
  if (c < C) {
    for (long long r = r0 + threadIdx.y; r < r1; r += RED_LANES) {
      const int x = (int)(r % W);
      const int y = (int)((r / W) % H);
      const float g = du[r * C + c];
      acc[9] += g;
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const int xx = x + kx - 1;
          if (xx < 0 || xx >= W) continue;
          const long long nb = r + (long long)(ky - 1) * W + (kx - 1);
          acc[ky * 3 + kx] = fmaf(g, __half2float(h[nb * C + c]), acc[ky * 3 + kx]);
        }
      }
    }
  }
  block_fold_lanes<10>(acc, sm);
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int t = 0; t < 10; t++) part[((size_t)blockIdx.x * 10 + t) * C + c] = acc[t];
  }
}
// partials [nblk][10][C] -> dw [C][9], db [C]
__global__ void __launch_bounds__(256) bwd_dw_fold_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ dw,
                                                          float* __restrict__ db) {
  PDL_TOP();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const float s = bwd_fold_sum(part, nblk, 10 * C, i, i < 10 * C);
  if (threadIdx.y != 0 || i >= 10 * C) return;
  const int t = i / C, c = i - t * C;
  if (t < 9) dw[c * 9 + t] = s; else if (db) db[c] = s;
}

__global__ void __launch_bounds__(256) bwd_flip9_kernel(const float* __restrict__ w, float* __restrict__ wf, int n) {
  PDL_TOP();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = i / 9, t = i - c * 9;
  wf[i] = w[c * 9 + 8 - t];
}

// 32 x 32 tile transpose: src [batch][rows][ld] (rows = tokens), dst [batch][s][c][pitch]
template <typename T>
__global__ void __launch_bounds__(256) bwd_packT_kernel(const T* __restrict__ src, long long M, int C, int ld, int S, int Ms, int pitch,
                                                        float* __restrict__ dst) {
  PDL_TOP();
  __shared__ float tile[32][33];
  const long long t0 = (long long)blockIdx.x * 32;      // padded token index s*Ms + m (Ms % 32 == 0: a tile never straddles splits)
  const int c0 = blockIdx.y * 32;
  const int s = (int)(t0 / Ms);
  const int m0 = (int)(t0 - (long long)s * Ms);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  src += (size_t)blockIdx.z * M * ld;
  dst += (size_t)blockIdx.z * S * C * pitch;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const long long tok = t0 + ty + i * 8;
    const int c = c0 + tx;
    float v = 0.f;
    if (tok < M && c < C) v = (float)src[tok * ld + c];
    tile[ty + i * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = c0 + ty + i * 8;
    if (c < C) dst[((size_t)s * C + c) * pitch + m0 + tx] = tile[tx][ty + i * 8];
  }
}

template <typename T>
int launch_packT(const T* src, int batch, long long M, int C, int ld, int S, int Ms, int pitch, float* dst, cudaStream_t st) {
  TCX_REQUIRE(Ms % 32 == 0 && (long long)S * Ms >= M && pitch >= Ms, "bwd_packT: bad split plan (M=%lld S=%d Ms=%d pitch=%d)", M, S, Ms, pitch);
  if (M == 0 || C == 0 || batch == 0) return 0;
  dim3 grid((unsigned)((long long)S * Ms / 32), (unsigned)cdiv(C, 32), (unsigned)batch);
  ProfScope prof("bwd_packT", st, (double)batch * M * C * (sizeof(T) + 4.0));
  tcx_launch_chain(bwd_packT_kernel<T>, dim3(grid), dim3(256), 0, st, src, M, C, ld, S, Ms, pitch, dst);
  return tcx_check_launch("bwd_packT");
}

// ---- efficient attention backward (EfficientAttention.forward MSTr.py:106-143) -------------------------------------------
// Column softmax of the keys over the N tokens of an image: chunk partials (max, sum of exp) -> fold in the consumers.
constexpr int EA_CHUNK = 128;   // token rows per block

__global__ void __launch_bounds__(RED_COLS * RED_LANES) ea_bwd_kstats_kernel(const __half* __restrict__ k, int ld, int N, int C,
                                                                            float* __restrict__ pm, float* __restrict__ ps) {
  PDL_TOP();
  __shared__ float sm[RED_LANES][1][RED_COLS];
  __shared__ float bm[RED_COLS];
  const int c = blockIdx.z * RED_COLS + threadIdx.x;
  const int b = blockIdx.y, chunks = gridDim.x;
  const int r0 = blockIdx.x * EA_CHUNK, r1 = min(r0 + EA_CHUNK, N);
  const __half* kb = k + (size_t)b * N * ld;
  float m = -INFINITY;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += RED_LANES) m = fmaxf(m, __half2float(kb[(size_t)r * ld + c]));
  sm[threadIdx.y][0][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int l = 1; l < RED_LANES; l++) m = fmaxf(m, sm[l][0][threadIdx.x]);
    bm[threadIdx.x] = m;
  }
  __syncthreads();
  m = bm[threadIdx.x];
  float acc[1] = {0.f};
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += RED_LANES) acc[0] += expf(__half2float(kb[(size_t)r * ld + c]) - m);
  __syncthreads();
  block_fold_lanes<1>(acc, sm);
  if (threadIdx.y == 0 && c < C) {
    pm[((size_t)b * chunks + blockIdx.x) * C + c] = m;
    ps[((size_t)b * chunks + blockIdx.x) * C + c] = acc[0];
  }
}

// P32 [B*N][C] = column softmax of K, V32 [B*N][C] = float(V)
__global__ void __launch_bounds__(RED_COLS * RED_LANES) ea_bwd_prep_kernel(const __half* __restrict__ k, const __half* __restrict__ v, int ld,
                                                                          int N, int C, const float* __restrict__ pm,
                                                                          const float* __restrict__ ps, float* __restrict__ P32,
                                                                          float* __restrict__ V32) {
  PDL_TOP();
  const int c = blockIdx.z * RED_COLS + threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.y, chunks = gridDim.x;
  float m = -INFINITY;
  for (int j = 0; j < chunks; j++) m = fmaxf(m, pm[((size_t)b * chunks + j) * C + c]);
  float sum = 0.f;
  for (int j = 0; j < chunks; j++) sum += ps[((size_t)b * chunks + j) * C + c] * expf(pm[((size_t)b * chunks + j) * C + c] - m);
  const float inv = 1.0f / sum;
  const int r0 = blockIdx.x * EA_CHUNK, r1 = min(r0 + EA_CHUNK, N);
  for (int r = r0 + threadIdx.y; r < r1; r += RED_LANES) {
    const size_t row = (size_t)b * N + r;
    P32[row * C + c] = expf(__half2float(k[row * ld + c]) - m) * inv;
    V32[row * C + c] = __half2float(v[row * ld + c]);
  }
}

// chunk partials of sum_n P dP per (image, channel)
__global__ void __launch_bounds__(RED_COLS * RED_LANES) ea_bwd_pdp_kernel(const float* __restrict__ P, const float* __restrict__ dP, int N, int C,
                                                                         float* __restrict__ sp) {
  PDL_TOP();
  __shared__ float sm[RED_LANES][1][RED_COLS];
  const int c = blockIdx.z * RED_COLS + threadIdx.x;
  const int b = blockIdx.y, chunks = gridDim.x;
  const int r0 = blockIdx.x * EA_CHUNK, r1 = min(r0 + EA_CHUNK, N);
  float acc[1] = {0.f};
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += RED_LANES) {
      const size_t i = ((size_t)b * N + r) * C + c;
      acc[0] = fmaf(P[i], dP[i], acc[0]);
    }
  block_fold_lanes<1>(acc, sm);
  if (threadIdx.y == 0 && c < C) sp[((size_t)b * chunks + blockIdx.x) * C + c] = acc[0];
}
// dK = P (dP - sum_n P dP) -> dkqv[:, 0:C] (row pitch ldo)
__global__ void __launch_bounds__(RED_COLS * RED_LANES) ea_bwd_dk_kernel(const float* __restrict__ P, const float* __restrict__ dP, int N, int C,
                                                                        const float* __restrict__ sp, float* __restrict__ dk, int ldo) {
  PDL_TOP();
  const int c = blockIdx.z * RED_COLS + threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.y, chunks = gridDim.x;
  float s = 0.f;
  for (int j = 0; j < chunks; j++) s += sp[((size_t)b * chunks + j) * C + c];
  const int r0 = blockIdx.x * EA_CHUNK, r1 = min(r0 + EA_CHUNK, N);
  for (int r = r0 + threadIdx.y; r < r1; r += RED_LANES) {
    const size_t row = (size_t)b * N + r;
    dk[row * ldo + c] = P[row * C + c] * (dP[row * C + c] - s);
  }
}
// channel softmax backward, one warp per token: dQ = Qs (dQs - sum_c Qs dQs) -> dq (row pitch ldo)
__global__ void __launch_bounds__(256) ea_bwd_dq_kernel(const __half* __restrict__ qs, const float* __restrict__ dqs, long long M, int C,
                                                        float* __restrict__ dq, int ldo) {
  PDL_TOP();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot = fmaf(__half2float(qs[row * C + c]), dqs[row * C + c], dot);
  dot = warp_sum(dot);
  for (int c = lane; c < C; c += 32) dq[row * ldo + c] = __half2float(qs[row * C + c]) * (dqs[row * C + c] - dot);
}

// out[b][i] = sum_s part[(b*S + s)*n + i]; optional transposed copy outT[b][j*R + i'] of the R x R matrix
__global__ void __launch_bounds__(256) bwd_fold_batched_kernel(const float* __restrict__ part, int S, int R, float* __restrict__ out,
                                                               float* __restrict__ outT) {
  PDL_TOP();
  const int n = R * R;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int b = blockIdx.y;
  float s = 0.f;
  for (int k = 0; k < S; k++) s += part[((size_t)b * S + k) * n + i];
  out[(size_t)b * n + i] = s;
  if (outT) outT[(size_t)b * n + (i % R) * R + i / R] = s;
}

}  // namespace

int bwd_red_blocks(long long M) {
  const int rows = red_rows_per_block(M);
  return (int)((M + rows - 1) / rows);
}

int launch_bwd_fold(const float* part, int S, long long n, float* out, cudaStream_t st) {
  if (n == 0) return 0;
  tcx_launch_chain(bwd_fold_kernel, dim3((unsigned)((n + 31) / 32)), dim3(dim3(32, 8)), 0, st, part, S, n, out);
  return tcx_check_launch("bwd_fold");
}

int launch_bwd_colsum(const float* x, long long M, int C, int ld, float* part, float* out, cudaStream_t st) {
  if (C == 0) return 0;
  const int rows = red_rows_per_block(M), nblk = bwd_red_blocks(M);
  if (M == 0) return cudaMemsetAsync(out, 0, sizeof(float) * C, st) == cudaSuccess ? 0 : -1;
  ProfScope prof("bwd_colsum", st, (double)M * C * 4.0);
  tcx_launch_chain(bwd_colsum_kernel, dim3(dim3(nblk, cdiv(C, RED_COLS))), dim3(dim3(RED_COLS, RED_LANES)), 0, st, x, M, C, ld, rows, part);
  TCX_TRY(tcx_check_launch("bwd_colsum"));
  return launch_bwd_fold(part, nblk, C, out, st);
}

int launch_bwd_ln_fold(const float* part, int nblk, int C, float* dgamma, float* dbeta, cudaStream_t st) {
  tcx_launch_chain(bwd_ln_fold_kernel, dim3(cdiv(2 * C, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, C, dgamma, dbeta);
  return tcx_check_launch("bwd_ln_fold");
}
int launch_bwd_dw_fold(const float* part, int nblk, int C, float* dw, float* db, cudaStream_t st) {
  tcx_launch_chain(bwd_dw_fold_kernel, dim3(cdiv(10 * C, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, C, dw, db);
  return tcx_check_launch("bwd_dw_fold");
}

int launch_bwd_ln(const float* u, const float* dz, const float* gamma, const float* beta, float eps, int gelu, float* du,
                  float* dgamma, float* dbeta, long long M, int C, float* stats, float* part, cudaStream_t st, const float* dres) {
  TCX_REQUIRE(C % 4 == 0 && C > 0, "bwd_ln: C %% 4 != 0 (C=%d)", C);
  TCX_REQUIRE(du != dz, "bwd_ln: du may not alias dz");
  if (M == 0) {
    cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
    return 0;
  }
  if (ln_bwd_fused_ok(M, C)) {      // one pass over (u, dz) instead of a row pass + a column pass
    TCX_TRY(launch_ln_bwd_fused(u, dz, gamma, beta, eps, gelu, du, nullptr, dres, M, C, part, st));
    return launch_bwd_ln_fold(part, ln_bwd_fused_blocks(M, C), C, dgamma, dbeta, st);
  }
  const unsigned rb = (unsigned)((M + 7) / 8);
  const int rows = red_rows_per_block(M), nblk = bwd_red_blocks(M);
  {
    ProfScope prof("bwd_ln_rows", st, (double)M * C * 12.0);
    if (gelu) tcx_launch_chain(bwd_ln_rows_kernel<true>, dim3(rb), dim3(256), 0, st, u, dz, gamma, beta, eps, du, stats, M, C);
    else tcx_launch_chain(bwd_ln_rows_kernel<false>, dim3(rb), dim3(256), 0, st, u, dz, gamma, beta, eps, du, stats, M, C);
    TCX_TRY(tcx_check_launch("bwd_ln_rows"));
  }
  {
    ProfScope prof("bwd_ln_cols", st, (double)M * C * 8.0);
    const dim3 grid(nblk, cdiv(C, RED_COLS)), block(RED_COLS, RED_LANES);
    if (gelu) tcx_launch_chain(bwd_ln_cols_kernel<true>, dim3(grid), dim3(block), 0, st, u, dz, gamma, beta, stats, M, C, rows, part);
    else tcx_launch_chain(bwd_ln_cols_kernel<false>, dim3(grid), dim3(block), 0, st, u, dz, gamma, beta, stats, M, C, rows, part);
    TCX_TRY(tcx_check_launch("bwd_ln_cols"));
  }
  tcx_launch_chain(bwd_ln_fold_kernel, dim3(cdiv(2 * C, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, C, dgamma, dbeta);
  TCX_TRY(tcx_check_launch("bwd_ln_fold"));
  if (dres) return launch_add_inplace(du, dres, M * C, st);
  return 0;
}

int launch_bwd_dwconv_wgrad(const float* du, const __half* h, int B, int H, int W, int C, float* dw, float* db, float* part,
                            cudaStream_t st) {
  const long long M = (long long)B * H * W;
  if (M == 0 || C == 0) return 0;
  const int rows = red_rows_per_block(M), nblk = bwd_red_blocks(M);
  {
    ProfScope prof("bwd_dw_wgrad", st, (double)M * C * 6.0);
    tcx_launch_chain(bwd_dw_wgrad_kernel, dim3(dim3(nblk, cdiv(C, RED_COLS))), dim3(dim3(RED_COLS, RED_LANES)), 0, st, du, h, B, H, W, C, rows, part);
    TCX_TRY(tcx_check_launch("bwd_dw_wgrad"));
  }
  tcx_launch_chain(bwd_dw_fold_kernel, dim3(cdiv(10 * C, 32)), dim3(dim3(32, 8)), 0, st, part, nblk, C, dw, db);
  return tcx_check_launch("bwd_dw_fold");
}

int launch_bwd_flip9(const float* w, float* wflip, int C, cudaStream_t st) {
  tcx_launch_chain(bwd_flip9_kernel, dim3(cdiv(9 * C, 256)), dim3(256), 0, st, w, wflip, 9 * C);
  return tcx_check_launch("bwd_flip9");
}

int launch_bwd_packT_f32(const float* src, long long M, int C, int ld, int S, int Ms, float* dst, cudaStream_t st) {
  return launch_packT<float>(src, 1, M, C, ld, S, Ms, Ms, dst, st);
}
int launch_bwd_packT_f16(const __half* src, long long M, int C, int ld, int S, int Ms, float* dst, cudaStream_t st) {
  return launch_packT<__half>(src, 1, M, C, ld, S, Ms, Ms, dst, st);
}
int launch_bwd_packT_batched_f32(const float* src, int batch, int N, int C, int ld, int S, int Ms, int pitch, float* dst, cudaStream_t st) {
  return launch_packT<float>(src, batch, N, C, ld, S, Ms, pitch, dst, st);
}
int launch_bwd_packT_batched_f16(const __half* src, int batch, int N, int C, int ld, int S, int Ms, int pitch, float* dst, cudaStream_t st) {
  return launch_packT<__half>(src, batch, N, C, ld, S, Ms, pitch, dst, st);
}

int launch_bwd_fold_batched(const float* part, int batch, int S, int R, float* out, float* outT, cudaStream_t st) {
  if (batch == 0 || R == 0) return 0;
  tcx_launch_chain(bwd_fold_batched_kernel, dim3(dim3(cdiv(R * R, 256), batch)), dim3(256), 0, st, part, S, R, out, outT);
  return tcx_check_launch("bwd_fold_batched");
}

int ea_bwd_chunks(int N) { return cdiv(N, EA_CHUNK); }

int launch_ea_bwd_prep(const __half* k, const __half* v, int ld, int B, int N, int C, float* pm, float* ps, float* P32, float* V32,
                       cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  const dim3 grid(ea_bwd_chunks(N), B, cdiv(C, RED_COLS)), block(RED_COLS, RED_LANES);
  tcx_launch_chain(ea_bwd_kstats_kernel, dim3(grid), dim3(block), 0, st, k, ld, N, C, pm, ps);
  TCX_TRY(tcx_check_launch("ea_bwd_kstats"));
  tcx_launch_chain(ea_bwd_prep_kernel, dim3(grid), dim3(block), 0, st, k, v, ld, N, C, pm, ps, P32, V32);
  return tcx_check_launch("ea_bwd_prep");
}

int launch_bwd_colsoftmax(const float* P, const float* dP, int B, int N, int C, float* sp, float* dk, int ldo, cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  const dim3 grid(ea_bwd_chunks(N), B, cdiv(C, RED_COLS)), block(RED_COLS, RED_LANES);
  tcx_launch_chain(ea_bwd_pdp_kernel, dim3(grid), dim3(block), 0, st, P, dP, N, C, sp);
  TCX_TRY(tcx_check_launch("ea_bwd_pdp"));
  tcx_launch_chain(ea_bwd_dk_kernel, dim3(grid), dim3(block), 0, st, P, dP, N, C, sp, dk, ldo);
  return tcx_check_launch("ea_bwd_dk");
}

int launch_ea_bwd_softmax(const float* P, const float* dP, const __half* qs, const float* dqs, int B, int N, int C, float* sp, float* dkqv,
                          cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  TCX_TRY(launch_bwd_colsoftmax(P, dP, B, N, C, sp, dkqv, 3 * C, st));
  const long long M = (long long)B * N;
  tcx_launch_chain(ea_bwd_dq_kernel, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, qs, dqs, M, C, dkqv + C, 3 * C);
  return tcx_check_launch("ea_bwd_dq");
}

void bwd_wgrad_splits(long long M, int Nout, int Kin, int* S, int* Ms) {
  // enough (tile, split) pairs for two waves of the persistent GEMM, at least 256 tokens of K per split
  const long long tiles = (long long)cdiv(Nout, 128) * cdiv(Kin, 64);
  long long s = (2 * 148 + tiles - 1) / tiles;
  const long long smax = (M + 255) / 256;
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  long long ms = ((M + s - 1) / s + 31) / 32 * 32;
  if (ms < 32) ms = 32;
  *Ms = (int)ms;
  *S = (int)((M + ms - 1) / ms);
  if (*S < 1) *S = 1;
}
