// Attention kernels of the hot path (all fp32 in HBM):
//  * efficient / channel attention (reference MSTr.py:106-143, :2309-2353): softmax over tokens of K,
//    softmax over channels of Q, CxC context.  One set of kernels serves both the tokens-major stage-1 /
//    decoder form and the bridge's raw [N,C]->[C,N] reinterpretation through (channel, token) strides.
//  * Multi-Branch factorized attention + conv relative position encoding (MSTr.py:801-886).
//  * bridge spatial-reduction attention, FFMA flash kernel (the tcgen05 kernel lives in flash_tc.cu).
#include "common.cuh"
#include "attention.cuh"

namespace {

// =====================================================================================
// efficient attention
// =====================================================================================
constexpr int EA_TN = 32;   // tokens per smem sub-tile

template <bool REINT>
__device__ __forceinline__ void ea_load_tile(float (*tile)[65], const float* __restrict__ base, long long sc,
                                             long long sn, int c0, int n0, int nend, int tid) {
  // tile[n][c] for c in [c0,c0+64), n in [n0, n0+EA_TN); zero fill past nend
  for (int i = tid; i < EA_TN * 64; i += 256) {
    int n, c;
    if (REINT) { n = i % EA_TN; c = i / EA_TN; } else { c = i & 63; n = i >> 6; }
    const int nn = n0 + n;
    tile[n][c] = nn < nend ? base[(long long)(c0 + c) * sc + (long long)nn * sn] : 0.f;
  }
}

// grid (nchunks, B, (C/64)^2). Partial context over one token chunk for one 64x64 (ck,cv) tile.
template <bool REINT>
__global__ void __launch_bounds__(256) ea_ctx_partial_kernel(EaView v, int N, int C, int chunk, float* __restrict__ part_ctx,
                                                             float* __restrict__ part_m, float* __restrict__ part_s) {
  __shared__ float kt[EA_TN][65];
  __shared__ float vt[EA_TN][65];
  __shared__ float red[4][64];
  __shared__ float mx[64];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int tiles = C >> 6;
  const int tk = blockIdx.z / tiles, tv = blockIdx.z % tiles;
  const int n_begin = blockIdx.x * chunk;
  const int n_end = min(N, n_begin + chunk);
  const float* __restrict__ kb = v.k + (long long)b * v.sb;
  const float* __restrict__ vb = v.v + (long long)b * v.sb;

  // pass 1: per-channel max over the chunk
  float m = -INFINITY;
  for (int n0 = n_begin; n0 < n_end; n0 += EA_TN) {
    ea_load_tile<REINT>(kt, kb, v.sc, v.sn, tk * 64, n0, n_end, tid);
    __syncthreads();
    const int c = tid & 63, q = tid >> 6;
    for (int n = q; n < EA_TN && n0 + n < n_end; n += 4) m = fmaxf(m, kt[n][c]);
    __syncthreads();
  }
  red[tid >> 6][tid & 63] = m;
  __syncthreads();
  if (tid < 64) mx[tid] = fmaxf(fmaxf(red[0][tid], red[1][tid]), fmaxf(red[2][tid], red[3][tid]));
  __syncthreads();

  // pass 2: e = exp(k - max), partial sum and partial context
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  float ssum = 0.f;
  for (int n0 = n_begin; n0 < n_end; n0 += EA_TN) {
    ea_load_tile<REINT>(kt, kb, v.sc, v.sn, tk * 64, n0, n_end, tid);
    ea_load_tile<REINT>(vt, vb, v.sc, v.sn, tv * 64, n0, n_end, tid);
    __syncthreads();
    for (int i = tid; i < EA_TN * 64; i += 256) {
      const int c = i & 63, n = i >> 6;
      kt[n][c] = (n0 + n < n_end) ? __expf(kt[n][c] - mx[c]) : 0.f;
    }
    __syncthreads();
    if (tid < 64) {
      float s = 0.f;
#pragma unroll 8
      for (int n = 0; n < EA_TN; n++) s += kt[n][tid];
      ssum += s;
    }
#pragma unroll 4
    for (int n = 0; n < EA_TN; n++) {
      float e[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { e[i] = kt[n][ty * 4 + i]; w[i] = vt[n][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(e[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int nchunks = gridDim.x;
  float* __restrict__ pc = part_ctx + ((long long)b * nchunks + blockIdx.x) * C * C;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      pc[(long long)(tk * 64 + ty * 4 + i) * C + tv * 64 + tx * 4 + j] = acc[i][j];
  if (tv == 0 && tid < 64) {
    part_m[((long long)b * nchunks + blockIdx.x) * C + tk * 64 + tid] = mx[tid];
    part_s[((long long)b * nchunks + blockIdx.x) * C + tk * 64 + tid] = ssum;
  }
}

// ctxT[b][cv][ck] = sum_chunks part[ck][cv]*exp(m_chunk-M) / sum_chunks s*exp(m_chunk-M)
__global__ void __launch_bounds__(256) ea_ctx_finalize_kernel(const float* __restrict__ part_ctx, const float* __restrict__ part_m,
                                                              const float* __restrict__ part_s, int nchunks, int C,
                                                              float* __restrict__ ctxT) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * C) return;
  const int cv = idx / C, ck = idx % C;   // consecutive threads -> consecutive ck: coalesced write
  const float* pm = part_m + (long long)b * nchunks * C;
  const float* ps = part_s + (long long)b * nchunks * C;
  float M = -INFINITY;
  for (int j = 0; j < nchunks; j++) M = fmaxf(M, pm[j * C + ck]);
  float S = 0.f, a = 0.f;
  for (int j = 0; j < nchunks; j++) {
    const float f = __expf(pm[j * C + ck] - M);
    S = fmaf(ps[j * C + ck], f, S);
    a = fmaf(part_ctx[((long long)b * nchunks + j) * C * C + (long long)ck * C + cv], f, a);
  }
  ctxT[(long long)b * C * C + (long long)cv * C + ck] = a / S;
}

// Q softmax over channels, tokens-major source (pitch sn), one warp per token. dst dense [B*N][C].
__global__ void __launch_bounds__(256) ea_qsoftmax_rows_kernel(const float* __restrict__ q, long long sn, long long total, int C,
                                                               float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= total) return;
  const float* __restrict__ src = q + row * sn;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, src[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(src[c] - m);
  s = 1.f / warp_sum(s);
  for (int c = lane; c < C; c += 32) dst[row * C + c] = __expf(src[c] - m) * s;
}

// Q softmax over channels for the reinterpreted [C][N] view (C == 64): dst[n][c] tokens-major.
__global__ void __launch_bounds__(256) ea_qsoftmax_reint_kernel(const float* __restrict__ q, long long sb, int N,
                                                                float* __restrict__ dst) {
  __shared__ float t[64][65];   // [c][n_local]
  const int b = blockIdx.y, n0 = blockIdx.x * 64, tid = threadIdx.x;
  const float* __restrict__ qb = q + (long long)b * sb;
  for (int i = tid; i < 64 * 64; i += 256) {
    const int n = i & 63, c = i >> 6;
    t[c][n] = (n0 + n < N) ? qb[(long long)c * N + n0 + n] : 0.f;
  }
  __syncthreads();
  if (tid < 64) {
    float m = -INFINITY;
    for (int c = 0; c < 64; c++) m = fmaxf(m, t[c][tid]);
    float s = 0.f;
    for (int c = 0; c < 64; c++) { const float e = __expf(t[c][tid] - m); t[c][tid] = e; s += e; }
    s = 1.f / s;
    for (int c = 0; c < 64; c++) t[c][tid] *= s;
  }
  __syncthreads();
  float* __restrict__ db = dst + (long long)b * N * 64;
  for (int i = tid; i < 64 * 64; i += 256) {
    const int c = i & 63, n = i >> 6;
    if (n0 + n < N) db[(long long)(n0 + n) * 64 + c] = t[c][n];
  }
}

// =====================================================================================
// Multi-Branch factorized attention
// =====================================================================================
struct MbGroups {
  const float* qkv[TCX_MAX_GROUPS];   // [B*N][3C]
  float* ctx[TCX_MAX_GROUPS];         // [B][h][Ch][Ch]   (already * Ch^-0.5 / colsum)
  float* out[TCX_MAX_GROUPS];         // [B*N][C]
  const float* cw[TCX_MAX_GROUPS][3]; // crpe conv weights (3x3, 5x5, 7x7)
  const float* cb[TCX_MAX_GROUPS][3];
};

constexpr int MB_TN = 64;
constexpr int MB_MAXP = 7;  // ceil(40*40/256)

// grid (heads, B, G): ctx[k][v] = scale * sum_n softmax_n(K)[n,k] V[n,v]
__global__ void __launch_bounds__(256) mb_ctx_kernel(const MbGroups gs, int N, int C, int Ch, float scale) {
  extern __shared__ float sm[];
  float* et = sm;                    // [MB_TN][Ch]
  float* vt = et + MB_TN * Ch;       // [MB_TN][Ch]
  float* red = vt + MB_TN * Ch;      // [256]
  float* mx = red + 256;             // [Ch]
  float* ss = mx + Ch;               // [Ch]
  const int tid = threadIdx.x, h = blockIdx.x, b = blockIdx.y, gi = blockIdx.z;
  const int heads = gridDim.x;
  const float* __restrict__ base = gs.qkv[gi] + (long long)b * N * 3 * C;
  const float* __restrict__ kp = base + C + h * Ch;
  const float* __restrict__ vp = base + 2 * C + h * Ch;
  const int nsl = 256 / Ch;
  // column max
  float m = -INFINITY;
  if (tid < nsl * Ch) {
    const int ck = tid % Ch, sl = tid / Ch;
    for (int n = sl; n < N; n += nsl) m = fmaxf(m, kp[(long long)n * 3 * C + ck]);
  }
  red[tid] = m;
  __syncthreads();
  if (tid < Ch) {
    float mm = -INFINITY;
    for (int s = 0; s < nsl; s++) mm = fmaxf(mm, red[s * Ch + tid]);
    mx[tid] = mm;
    ss[tid] = 0.f;
  }
  __syncthreads();
  float acc[MB_MAXP];
#pragma unroll
  for (int i = 0; i < MB_MAXP; i++) acc[i] = 0.f;
  const int npairs = Ch * Ch;
  float ssum = 0.f;
  for (int n0 = 0; n0 < N; n0 += MB_TN) {
    const int tn = min(MB_TN, N - n0);
    for (int i = tid; i < MB_TN * Ch; i += 256) {
      const int n = i / Ch, c = i % Ch;
      const bool ok = n < tn;
      et[i] = ok ? __expf(kp[(long long)(n0 + n) * 3 * C + c] - mx[c]) : 0.f;
      vt[i] = ok ? vp[(long long)(n0 + n) * 3 * C + c] : 0.f;
    }
    __syncthreads();
    if (tid < Ch) {
      float s = 0.f;
      for (int n = 0; n < MB_TN; n++) s += et[n * Ch + tid];
      ssum += s;
    }
#pragma unroll
    for (int i = 0; i < MB_MAXP; i++) {
      const int p = tid + i * 256;
      if (p < npairs) {
        const int ck = p / Ch, cv = p % Ch;
        float a = acc[i];
#pragma unroll 8
        for (int n = 0; n < MB_TN; n++) a = fmaf(et[n * Ch + ck], vt[n * Ch + cv], a);
        acc[i] = a;
      }
    }
    __syncthreads();
  }
  if (tid < Ch) ss[tid] = ssum;
  __syncthreads();
  float* __restrict__ out = gs.ctx[gi] + ((long long)b * heads + h) * npairs;
#pragma unroll
  for (int i = 0; i < MB_MAXP; i++) {
    const int p = tid + i * 256;
    if (p < npairs) out[p] = scale * acc[i] / ss[p / Ch];
  }
}

// grid (ceil(N/TN), B, G), block 256: out[n,c] = sum_k q[n,hk] ctx[h][k][v] + q[n,c]*(dwconv_win(h)(V)[n,c] + b)
template <int TN>
__global__ void __launch_bounds__(256) mb_apply_kernel(const MbGroups gs, int H, int W, int C, int Ch) {
  extern __shared__ float sm[];
  float* ctx = sm;                 // [heads][Ch][Ch] = C*Ch
  float* qs = ctx + C * Ch;        // [TN][C]
  const int tid = threadIdx.x, b = blockIdx.y, gi = blockIdx.z;
  const int N = H * W;
  const int n0 = blockIdx.x * TN;
  const int heads = C / Ch;
  const float* __restrict__ base = gs.qkv[gi] + (long long)b * N * 3 * C;
  const float* __restrict__ cg = gs.ctx[gi] + (long long)b * heads * Ch * Ch;
  for (int i = tid; i < C * Ch; i += 256) ctx[i] = cg[i];
  for (int i = tid; i < TN * C; i += 256) {
    const int n = n0 + i / C;
    qs[i] = n < N ? base[(long long)n * 3 * C + (i % C)] : 0.f;
  }
  __syncthreads();
  const float* __restrict__ vbase = base + 2 * C;
  for (int i = tid; i < TN * C; i += 256) {
    const int tl = i / C, c = i % C;
    const int n = n0 + tl;
    if (n >= N) break;
    const int h = c / Ch, cv = c % Ch;
    // factorized attention term
    float fa = 0.f;
    const float* qrow = qs + tl * C + h * Ch;
    const float* crow = ctx + h * Ch * Ch + cv;
    for (int k = 0; k < Ch; k++) fa = fmaf(qrow[k], crow[k * Ch], fa);
    // conv relative position encoding: heads 0-1 -> 3x3, 2-4 -> 5x5, 5-7 -> 7x7 (MSTr.py:958)
    int win, wi, cl;
    if (h < 2) { win = 3; wi = 0; cl = c; }
    else if (h < 5) { win = 5; wi = 1; cl = c - 2 * Ch; }
    else { win = 7; wi = 2; cl = c - 5 * Ch; }
    const float* __restrict__ wt = gs.cw[gi][wi] + (long long)cl * win * win;
    float cvv = gs.cb[gi][wi][cl];
    const int y = n / W, x = n % W, r = win >> 1;
    for (int ky = 0; ky < win; ky++) {
      const int yy = y + ky - r;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < win; kx++) {
        const int xx = x + kx - r;
        if (xx < 0 || xx >= W) continue;
        cvv = fmaf(vbase[(long long)(yy * W + xx) * 3 * C + c], __ldg(wt + ky * win + kx), cvv);
      }
    }
    gs.out[gi][((long long)b * N + n) * C + c] = fa + qs[i] * cvv;
  }
}

// =====================================================================================
// bridge SR attention, FFMA flash (one head, d = 64)
// =====================================================================================
constexpr int FQ = 64, FK = 32;
// grid (ceil(Nq/64), B), 128 threads. q [B][Nq][64] (ld 64), kv [B][Nk][128] (k | v), out [B][Nq][64]
__global__ void __launch_bounds__(128) flash_ffma_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                         float* __restrict__ out, int Nq, int Nk, float scale) {
  __shared__ float Qs[FQ][65];
  __shared__ float Ks[FK][65];
  __shared__ float Vs[FK][65];
  __shared__ float Ps[FQ][FK + 1];
  const int tid = threadIdx.x, b = blockIdx.y, q0 = blockIdx.x * FQ;
  const float* __restrict__ qb = q + (long long)b * Nq * 64;
  const float* __restrict__ kb = kv + (long long)b * Nk * 128;
  for (int i = tid; i < FQ * 64; i += 128) {
    const int r = i >> 6, d = i & 63;
    Qs[r][d] = (q0 + r < Nq) ? qb[(long long)(q0 + r) * 64 + d] * scale : 0.f;
  }
  const int r = tid >> 1, half = tid & 1;
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; i++) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < Nk; k0 += FK) {
    __syncthreads();
    for (int i = tid; i < FK * 64; i += 128) {
      const int j = i >> 6, d = i & 63;
      const bool ok = k0 + j < Nk;
      Ks[j][d] = ok ? kb[(long long)(k0 + j) * 128 + d] : 0.f;
      Vs[j][d] = ok ? kb[(long long)(k0 + j) * 128 + 64 + d] : 0.f;
    }
    __syncthreads();
    float s[16];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 16; jj++) {
      const int j = half * 16 + jj;
      float a = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; d++) a = fmaf(Qs[r][d], Ks[j][d], a);
      s[jj] = (k0 + j < Nk) ? a : -INFINITY;
      tmax = fmaxf(tmax, s[jj]);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    const float mn = fmaxf(m, tmax);
    const float alpha = __expf(m - mn);
    float ps = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; jj++) {
      const float p = __expf(s[jj] - mn);
      Ps[r][half * 16 + jj] = p;
      ps += p;
    }
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    l = l * alpha + ps;
    m = mn;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i++) o[i] *= alpha;
    for (int j = 0; j < FK; j++) {
      const float p = Ps[r][j];
#pragma unroll
      for (int i = 0; i < 32; i++) o[i] = fmaf(p, Vs[j][half * 32 + i], o[i]);
    }
  }
  if (q0 + r < Nq) {
    const float inv = 1.f / l;
    float* __restrict__ ob = out + ((long long)b * Nq + q0 + r) * 64 + half * 32;
#pragma unroll
    for (int i = 0; i < 32; i++) ob[i] = o[i] * inv;
  }
}

}  // namespace

// =====================================================================================
// host launchers
// =====================================================================================
size_t ea_workspace_floats(int B, int N, int C) {
  const int chunk = 256;
  const int nchunks = cdiv(N, chunk);
  return (size_t)B * nchunks * ((size_t)C * C + 2 * C);
}

int launch_ea_context(const EaView& v, bool reinterpret, int B, int N, int C, float* ws, float* ctxT, cudaStream_t st) {
  TCX_REQUIRE(C % 64 == 0, "eff_attn: C must be a multiple of 64 (got %d)", C);
  const int chunk = 256;
  const int nchunks = cdiv(N, chunk);
  float* part_ctx = ws;
  float* part_m = part_ctx + (size_t)B * nchunks * C * C;
  float* part_s = part_m + (size_t)B * nchunks * C;
  const int tiles = C / 64;
  dim3 grid(nchunks, B, tiles * tiles);
  if (reinterpret) ea_ctx_partial_kernel<true><<<grid, 256, 0, st>>>(v, N, C, chunk, part_ctx, part_m, part_s);
  else ea_ctx_partial_kernel<false><<<grid, 256, 0, st>>>(v, N, C, chunk, part_ctx, part_m, part_s);
  TCX_TRY(tcx_check_launch("ea_ctx_partial"));
  dim3 g2(cdiv(C * C, 256), B);
  ea_ctx_finalize_kernel<<<g2, 256, 0, st>>>(part_ctx, part_m, part_s, nchunks, C, ctxT);
  return tcx_check_launch("ea_ctx_finalize");
}

int launch_ea_qsoftmax(const EaView& v, bool reinterpret, int B, int N, int C, float* dst, cudaStream_t st) {
  if (reinterpret) {
    TCX_REQUIRE(C == 64, "eff_attn(reinterpret): C must be 64");
    dim3 grid(cdiv(N, 64), B);
    ea_qsoftmax_reint_kernel<<<grid, 256, 0, st>>>(v.q, v.sb, N, dst);
  } else {
    const long long total = (long long)B * N;
    ea_qsoftmax_rows_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(v.q, v.sn, total, C, dst);
  }
  return tcx_check_launch("ea_qsoftmax");
}

int launch_mb_attention(const MbAttnArgs& a, cudaStream_t st) {
  TCX_REQUIRE(a.C % a.heads == 0, "mb_attn: C %% heads != 0");
  const int Ch = a.C / a.heads;
  TCX_REQUIRE(Ch * Ch <= MB_MAXP * 256 && Ch <= 256, "mb_attn: head dim %d too large", Ch);
  TCX_REQUIRE(a.heads == 8, "mb_attn: crpe window map {3:2,5:3,7:3} needs 8 heads");
  MbGroups gs{};
  for (int i = 0; i < a.groups; i++) {
    gs.qkv[i] = a.qkv[i]; gs.ctx[i] = a.ctx[i]; gs.out[i] = a.out[i];
    for (int j = 0; j < 3; j++) { gs.cw[i][j] = a.cw[i][j]; gs.cb[i][j] = a.cb[i][j]; }
  }
  const int N = a.H * a.W;
  {
    dim3 grid(a.heads, a.B, a.groups);
    const size_t smem = (size_t)(2 * MB_TN * Ch + 256 + 2 * Ch) * sizeof(float);
    mb_ctx_kernel<<<grid, 256, smem, st>>>(gs, N, a.C, Ch, a.scale);
    TCX_TRY(tcx_check_launch("mb_ctx"));
  }
  {
    constexpr int TN = 8;
    dim3 grid(cdiv(N, TN), a.B, a.groups);
    const size_t smem = (size_t)(a.C * Ch + TN * a.C) * sizeof(float);
    static PerDeviceOnce once;
    if (smem > 48 * 1024 && once.first())
      cudaFuncSetAttribute(mb_apply_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    TCX_REQUIRE(smem <= 100 * 1024, "mb_attn: smem too large");
    mb_apply_kernel<TN><<<grid, 256, smem, st>>>(gs, a.H, a.W, a.C, Ch);
    TCX_TRY(tcx_check_launch("mb_apply"));
  }
  return 0;
}

int launch_flash_ffma(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, cudaStream_t st) {
  dim3 grid(cdiv(Nq, FQ), B);
  ProfScope prof("flash_ffma", st);
  flash_ffma_kernel<<<grid, 128, 0, st>>>(q, kv, out, Nq, Nk, scale);
  return tcx_check_launch("flash_ffma");
}
