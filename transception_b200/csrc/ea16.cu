// Efficient ("linear") attention on fp16 K/Q/V (reference MSTr.py:106-143 EfficientAttention and, with the raw
// [N,C] -> [C,N] reinterpretation of MSTr.py:2312-2314, M_EfficientChannelAtten):
//   ctx[ck][cv] = sum_n softmax_n(K)[n,ck] V[n,cv]      (token-chunk partials + a combine kernel, exact max handling)
//   qsm[n][c]   = softmax_c(Q)[n,c]
// The two GEMMs around them (qsm x ctx, reprojection) run on the tensor-core kernel with fp16 operands.
#include <cuda_fp16.h>
#include "common.cuh"
#include "ea16.cuh"

namespace {

constexpr int EA_T = 112;          // tokens per chunk (3136 = 28 x 112)
constexpr int EA_EP = 65;          // padded pitch of the fp32 exp tile
constexpr int EA_VP = 72;          // padded pitch (halfs) of the V tile

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// grid (nchunks, B, (C/64)^2), 256 threads.  Partial context of one 64(ck) x 64(cv) tile over one token chunk,
// written TRANSPOSED ([cv][ck]) so the combine kernel reads and writes coalesced.
template <bool REINT>
__global__ void __launch_bounds__(256) ea16_ctx_partial_kernel(Ea16View v, int N, int C, float* __restrict__ part_ctx,
                                                               float* __restrict__ part_m, float* __restrict__ part_s) {
  __shared__ float E[EA_T * EA_EP];
  __shared__ __align__(16) __half V[EA_T * EA_VP];
  __shared__ float red[4][64];
  __shared__ float mx[64];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.y;
  const int tiles = C >> 6;
  const int tk = blockIdx.z / tiles, tv = blockIdx.z % tiles;
  const int n0 = blockIdx.x * EA_T;
  const int tn = min(EA_T, N - n0);
  const __half* __restrict__ kb = v.k + (long long)b * v.sb;
  const __half* __restrict__ vb = v.v + (long long)b * v.sb;
  if (!REINT) {
    for (int i = tid; i < EA_T * 8; i += 256) {
      const int n = i >> 3, j = i & 7;
      float f[8];
      uint4 rv = make_uint4(0u, 0u, 0u, 0u);
      if (n < tn) {
        unpack8(*reinterpret_cast<const uint4*>(kb + (long long)(n0 + n) * v.ldt + tk * 64 + j * 8), f);
        rv = *reinterpret_cast<const uint4*>(vb + (long long)(n0 + n) * v.ldt + tv * 64 + j * 8);
      } else {
#pragma unroll
        for (int q = 0; q < 8; q++) f[q] = -INFINITY;
      }
#pragma unroll
      for (int q = 0; q < 8; q++) E[n * EA_EP + j * 8 + q] = f[q];
      *reinterpret_cast<uint4*>(V + n * EA_VP + j * 8) = rv;
    }
  } else {
    // rows are channels c' (pitch N), 4 consecutive tokens per 8-byte load (N % 4 == 0 checked by the launcher)
    for (int i = tid; i < 64 * (EA_T / 4); i += 256) {
      const int c = i / (EA_T / 4), n = (i % (EA_T / 4)) * 4;
      float kf[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      __half vh[4] = {__half(0.f), __half(0.f), __half(0.f), __half(0.f)};
      if (n < tn) {   // tn is a multiple of 4
        const uint2 rk = *reinterpret_cast<const uint2*>(kb + (long long)(tk * 64 + c) * N + n0 + n);
        const uint2 rv = *reinterpret_cast<const uint2*>(vb + (long long)(tv * 64 + c) * N + n0 + n);
        const __half2* hk = reinterpret_cast<const __half2*>(&rk);
        const float2 a0 = __half22float2(hk[0]), a1 = __half22float2(hk[1]);
        kf[0] = a0.x; kf[1] = a0.y; kf[2] = a1.x; kf[3] = a1.y;
        const __half* hv = reinterpret_cast<const __half*>(&rv);
#pragma unroll
        for (int q = 0; q < 4; q++) vh[q] = hv[q];
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        E[(n + q) * EA_EP + c] = kf[q];
        V[(n + q) * EA_VP + c] = vh[q];
      }
    }
  }
  __syncthreads();
  {  // column max over the chunk
    const int c = tid & 63, sl = tid >> 6;
    float m = -INFINITY;
    for (int n = sl; n < EA_T; n += 4) m = fmaxf(m, E[n * EA_EP + c]);
    red[sl][c] = m;
  }
  __syncthreads();
  if (tid < 64) mx[tid] = fmaxf(fmaxf(red[0][tid], red[1][tid]), fmaxf(red[2][tid], red[3][tid]));
  __syncthreads();
  {  // exp in place + column sums
    const int c = tid & 63, sl = tid >> 6;
    const float mc = mx[c];
    float s = 0.f;
    for (int n = sl; n < EA_T; n += 4) {
      const float e = __expf(E[n * EA_EP + c] - mc);     // padded rows hold -inf -> 0
      E[n * EA_EP + c] = e;
      s += e;
    }
    red[sl][c] = s;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
#pragma unroll 4
  for (int n = 0; n < EA_T; n++) {
    float e[4];
#pragma unroll
    for (int i = 0; i < 4; i++) e[i] = E[n * EA_EP + ty * 4 + i];
    const uint2 rv = *reinterpret_cast<const uint2*>(V + n * EA_VP + tx * 4);
    const __half2* hv = reinterpret_cast<const __half2*>(&rv);
    const float2 w0 = __half22float2(hv[0]), w1 = __half22float2(hv[1]);
    const float w[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = fmaf(e[i], w[j], acc[i][j]);
  }
  const int nchunks = gridDim.x;
  float* __restrict__ pc = part_ctx + ((long long)b * nchunks + blockIdx.x) * C * C;
#pragma unroll
  for (int j = 0; j < 4; j++)
    *reinterpret_cast<float4*>(pc + (long long)(tv * 64 + tx * 4 + j) * C + tk * 64 + ty * 4) =
        make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
  if (tv == 0 && tid < 64) {
    part_m[((long long)b * nchunks + blockIdx.x) * C + tk * 64 + tid] = mx[tid];
    part_s[((long long)b * nchunks + blockIdx.x) * C + tk * 64 + tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
  }
}

// ctxT16[b][cv][ck] = sum_j part[j][cv][ck] e^(m_j - M) / sum_j s_j e^(m_j - M)
__global__ void __launch_bounds__(256) ea16_ctx_combine_kernel(const float* __restrict__ part_ctx, const float* __restrict__ part_m,
                                                               const float* __restrict__ part_s, int nchunks, int C,
                                                               __half* __restrict__ ctxT) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * C) return;
  const int ck = idx % C;
  const float* pm = part_m + (long long)b * nchunks * C + ck;
  const float* ps = part_s + (long long)b * nchunks * C + ck;
  const float* pc = part_ctx + (long long)b * nchunks * C * C + idx;
  float M = -INFINITY;
  for (int j = 0; j < nchunks; j++) M = fmaxf(M, pm[j * C]);
  float S = 0.f, a = 0.f;
  for (int j = 0; j < nchunks; j++) {
    const float f = __expf(pm[j * C] - M);
    S = fmaf(ps[j * C], f, S);
    a = fmaf(pc[(long long)j * C * C], f, a);
  }
  ctxT[(long long)b * C * C + idx] = __float2half_rn(a / S);
}

// Q softmax over channels, tokens-major source with pitch ldt: LPT lanes per token, NV 16-byte vectors per lane.
template <int LPT, int NV>
__global__ void __launch_bounds__(256) ea16_qsoftmax_kernel(const __half* __restrict__ q, int ldt, long long total, int C,
                                                            __half* __restrict__ dst) {
  constexpr int SLOTS = 32 / LPT;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int slot = lane / LPT, sl = lane % LPT;
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * SLOTS + slot;
  const bool live = row < total;
  float f[NV][8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (live) unpack8(*reinterpret_cast<const uint4*>(q + row * ldt + (sl + i * LPT) * 8), f[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; j++) f[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) m = fmaxf(m, f[i][j]);
  }
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) { f[i][j] = __expf(f[i][j] - m); s += f[i][j]; }
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  if (!live) return;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    __half2 h[4];
#pragma unroll
    for (int j = 0; j < 4; j++) h[j] = __floats2half2_rn(f[i][2 * j] * inv, f[i][2 * j + 1] * inv);
    *reinterpret_cast<uint4*>(dst + row * C + (sl + i * LPT) * 8) = *reinterpret_cast<uint4*>(h);
  }
}

// reinterpreted [C=64][N] view: softmax across the 64 rows for every column n, written tokens-major dst[n][c]
__global__ void __launch_bounds__(256) ea16_qsoftmax_reint_kernel(const __half* __restrict__ q, long long sb, int N,
                                                                  __half* __restrict__ dst) {
  __shared__ float t[64][65];   // [c][n_local]
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, n0 = blockIdx.x * 64, tid = threadIdx.x;
  const __half* __restrict__ qb = q + (long long)b * sb;
  for (int i = tid; i < 64 * 16; i += 256) {
    const int c = i >> 4, n = (i & 15) * 4;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 + n < N) {    // N % 4 == 0
      const uint2 r = *reinterpret_cast<const uint2*>(qb + (long long)c * N + n0 + n);
      const __half2* h = reinterpret_cast<const __half2*>(&r);
      const float2 a0 = __half22float2(h[0]), a1 = __half22float2(h[1]);
      f[0] = a0.x; f[1] = a0.y; f[2] = a1.x; f[3] = a1.y;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) t[c][n + j] = f[j];
  }
  __syncthreads();
  {  // 4 threads per column, 16 rows each
    const int n = tid & 63, part = tid >> 6;
    float m = -INFINITY;
    for (int c = part * 16; c < part * 16 + 16; c++) m = fmaxf(m, t[c][n]);
    __shared__ float rm[4][64], rs[4][64];
    rm[part][n] = m;
    __syncthreads();
    m = fmaxf(fmaxf(rm[0][n], rm[1][n]), fmaxf(rm[2][n], rm[3][n]));
    float s = 0.f;
    for (int c = part * 16; c < part * 16 + 16; c++) { const float e = __expf(t[c][n] - m); t[c][n] = e; s += e; }
    rs[part][n] = s;
    __syncthreads();
    s = 1.f / (rs[0][n] + rs[1][n] + rs[2][n] + rs[3][n]);
    for (int c = part * 16; c < part * 16 + 16; c++) t[c][n] *= s;
  }
  __syncthreads();
  __half* __restrict__ db = dst + (long long)b * N * 64;
  for (int i = tid; i < 64 * 8; i += 256) {
    const int n = i >> 3, j = i & 7;
    if (n0 + n < N) {
      __half2 h[4];
#pragma unroll
      for (int q2 = 0; q2 < 4; q2++) h[q2] = __floats2half2_rn(t[j * 8 + 2 * q2][n], t[j * 8 + 2 * q2 + 1][n]);
      *reinterpret_cast<uint4*>(db + (long long)(n0 + n) * 64 + j * 8) = *reinterpret_cast<uint4*>(h);
    }
  }
}

// ---- tensor-core context (token-major K/V): K-major operands for  ctxT[cv][ck] = sum_n Vt[cv][n] * Pt[ck][n] ------------
constexpr int EC_SPLIT_TOK = 64;       // tokens per statistics block (8 rows per warp, all loads in flight together)

// Partial column statistics of K over a token range: max and sum of exp(k - max) per channel.  grid (nsplit, B), 256
// threads; a warp walks rows, lanes own half2 channel pairs (C <= 512).  Fixed-order combines: deterministic.
__global__ void __launch_bounds__(256) ea16_colstats_kernel(const __half* __restrict__ k, long long sb, int ldt, int N, int C,
                                                            float* __restrict__ pm, float* __restrict__ ps) {
  __shared__ float red[8][512];
  __shared__ float bm[512];
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, b = blockIdx.y, nsplit = gridDim.x;
  const int r0 = split * EC_SPLIT_TOK, r1 = min(N, r0 + EC_SPLIT_TOK);
  const __half* kb = k + (long long)b * sb;
  const int npair = C >> 6;            // half2 pairs per lane
  // this warp's 8 rows (r0 + warp + 8 i), kept in registers for both passes
  __half2 rows[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int r = r0 + warp + 8 * i;
    const __half2* row = reinterpret_cast<const __half2*>(kb + (long long)(r < r1 ? r : r0) * ldt);
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < npair) rows[i][j] = r < r1 ? row[lane + 32 * j] : __float2half2_rn(-INFINITY);
  }
  float m[8][2];
#pragma unroll
  for (int j = 0; j < 8; j++) m[j][0] = m[j][1] = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < npair) {
        const float2 f = __half22float2(rows[i][j]);
        m[j][0] = fmaxf(m[j][0], f.x);
        m[j][1] = fmaxf(m[j][1], f.y);
      }
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (j < npair) { red[warp][2 * (lane + 32 * j)] = m[j][0]; red[warp][2 * (lane + 32 * j) + 1] = m[j][1]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float v = red[0][c];
#pragma unroll
    for (int w = 1; w < 8; w++) v = fmaxf(v, red[w][c]);
    bm[c] = v;
  }
  __syncthreads();
  float s[8][2];
#pragma unroll
  for (int j = 0; j < 8; j++) s[j][0] = s[j][1] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < npair) {
        const float2 f = __half22float2(rows[i][j]);          // -inf rows (past the range) contribute exp(-inf) = 0
        s[j][0] += __expf(f.x - bm[2 * (lane + 32 * j)]);
        s[j][1] += __expf(f.y - bm[2 * (lane + 32 * j) + 1]);
      }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (j < npair) { red[warp][2 * (lane + 32 * j)] = s[j][0]; red[warp][2 * (lane + 32 * j) + 1] = s[j][1]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float v = red[0][c];
#pragma unroll
    for (int w = 1; w < 8; w++) v += red[w][c];
    pm[((long long)b * nsplit + split) * C + c] = bm[c];
    ps[((long long)b * nsplit + split) * C + c] = v;
  }
}

// grid (ceil(N/64), B, C/64), 256 threads: fold the statistics partials of this channel group, then transpose a
// 64-token x 64-channel tile of K (as softmax probabilities) and of V through shared memory:
//   Pt[b][s][c][n'] = exp(k[n][c] - max_c) / sum_c,  Vt[b][s][c][n'] = v[n][c]  with n = s * Ks + n' (split-major, so a
//   (image, K-split) pair is one batch entry of the context GEMM); columns n >= N are written as zeros.
constexpr int EC_PITCH = 66;
__global__ void __launch_bounds__(256) ea16_packT_kernel(const __half* __restrict__ k, const __half* __restrict__ v, long long sb, int ldt,
                                                         int N, int C, int Ks, int KS, int nsplit, const float* __restrict__ pm,
                                                         const float* __restrict__ ps, __half* __restrict__ Pt, __half* __restrict__ Vt) {
  __shared__ float mx[64], inv[64];
  __shared__ __align__(16) __half tile[64 * EC_PITCH];
  pdl_trigger();
  pdl_wait();
  const int n0 = blockIdx.x * 64, b = blockIdx.y, c0 = blockIdx.z * 64;
  if (threadIdx.x < 64) {
    const int c = c0 + threadIdx.x;
    float M = -INFINITY;
    for (int sI = 0; sI < nsplit; sI++) M = fmaxf(M, pm[((long long)b * nsplit + sI) * C + c]);
    float S = 0.f;
    for (int sI = 0; sI < nsplit; sI++) S += ps[((long long)b * nsplit + sI) * C + c] * __expf(pm[((long long)b * nsplit + sI) * C + c] - M);
    mx[threadIdx.x] = M;
    inv[threadIdx.x] = 1.f / S;
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, j = threadIdx.x & 7;           // 32 rows x 8 channel vectors per pass
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const __half* src = (which ? v : k) + (long long)b * sb + c0 + j * 8;
    const int ks = n0 / Ks;
    __half* dst = (which ? Vt : Pt) + (((long long)b * KS + ks) * C + c0) * Ks + (n0 - ks * Ks);
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      const int n = r + pass * 32;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      const bool live = n0 + n < N;
      if (live) raw = *reinterpret_cast<const uint4*>(src + (long long)(n0 + n) * ldt);
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        float o = f[e];
        if (!which) o = live ? __expf(f[e] - mx[j * 8 + e]) * inv[j * 8 + e] : 0.f;
        tile[(j * 8 + e) * EC_PITCH + n] = __float2half_rn(o);
      }
    }
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      const int c = r + pass * 32;                                  // channel row of the tile; this thread writes 8 tokens
      const uint32_t* w = reinterpret_cast<const uint32_t*>(tile + c * EC_PITCH + j * 8);
      *reinterpret_cast<uint4*>(dst + (long long)c * Ks + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
  }
}

// ctxT16[b][i] = sum over the K-splits of the fp32 partial contexts (index order: deterministic)
__global__ void __launch_bounds__(256) ea16_splitk_combine_kernel(const float* __restrict__ part, __half* __restrict__ ctxT, int KS, int CC) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (i >= CC) return;
  float v = 0.f;
  for (int s2 = 0; s2 < KS; s2++) v += part[((long long)b * KS + s2) * CC + i];
  ctxT[(long long)b * CC + i] = __float2half_rn(v);
}

}  // namespace

size_t ea16_workspace_floats(int B, int N, int C) {
  const int nchunks = cdiv(N, EA_T);
  return (size_t)B * nchunks * ((size_t)C * C + 2 * C);
}

int launch_ea16_context(const Ea16View& v, int B, int N, int C, float* ws, __half* ctxT, cudaStream_t st) {
  TCX_REQUIRE(C % 64 == 0, "eff_attn16: C must be a multiple of 64 (got %d)", C);
  TCX_REQUIRE(!v.reint || N % 4 == 0, "eff_attn16(reinterpret): token count %d must be a multiple of 4", N);
  const int nchunks = cdiv(N, EA_T);
  float* part_ctx = ws;
  float* part_m = part_ctx + (size_t)B * nchunks * C * C;
  float* part_s = part_m + (size_t)B * nchunks * C;
  const int tiles = C / 64;
  dim3 grid(nchunks, B, tiles * tiles);
  ProfScope prof("ea16_ctx", st);
  if (v.reint) tcx_launch_pdl(ea16_ctx_partial_kernel<true>, grid, dim3(256), 0, st, v, N, C, part_ctx, part_m, part_s);
  else tcx_launch_pdl(ea16_ctx_partial_kernel<false>, grid, dim3(256), 0, st, v, N, C, part_ctx, part_m, part_s);
  TCX_TRY(tcx_check_launch("ea16_ctx_partial"));
  dim3 g2(cdiv(C * C, 256), B);
  tcx_launch_pdl(ea16_ctx_combine_kernel, g2, dim3(256), 0, st, part_ctx, part_m, part_s, nchunks, C, ctxT);
  return tcx_check_launch("ea16_ctx_combine");
}

int launch_ea16_qsoftmax(const Ea16View& v, int B, int N, int C, __half* dst, cudaStream_t st) {
  if (v.reint) {
    TCX_REQUIRE(C == 64 && N % 4 == 0, "eff_attn16(reinterpret): needs C == 64 and N %% 4 == 0");
    dim3 grid(cdiv(N, 64), B);
    tcx_launch_pdl(ea16_qsoftmax_reint_kernel, grid, dim3(256), 0, st, v.q, v.sb, N, dst);
    return tcx_check_launch("ea16_qsoftmax_reint");
  }
  const long long total = (long long)B * N;
  const int nv8 = C / 8;
  TCX_REQUIRE(C % 64 == 0, "eff_attn16: C must be a multiple of 64");
  int LPT = 8;
  if (nv8 % 32 == 0) LPT = 32;
  else if (nv8 % 16 == 0) LPT = 16;
  const int NV = nv8 / LPT;
  const long long per_block = 8 * (32 / LPT);
  const unsigned grid = (unsigned)((total + per_block - 1) / per_block);
  if (LPT == 8 && NV == 1) tcx_launch_pdl(ea16_qsoftmax_kernel<8, 1>, dim3(grid), dim3(256), 0, st, v.q, v.ldt, total, C, dst);
  else if (LPT == 8 && NV == 5) tcx_launch_pdl(ea16_qsoftmax_kernel<8, 5>, dim3(grid), dim3(256), 0, st, v.q, v.ldt, total, C, dst);
  else if (LPT == 16 && NV == 1) tcx_launch_pdl(ea16_qsoftmax_kernel<16, 1>, dim3(grid), dim3(256), 0, st, v.q, v.ldt, total, C, dst);
  else if (LPT == 32 && NV == 1) tcx_launch_pdl(ea16_qsoftmax_kernel<32, 1>, dim3(grid), dim3(256), 0, st, v.q, v.ldt, total, C, dst);
  else if (LPT == 32 && NV == 2) tcx_launch_pdl(ea16_qsoftmax_kernel<32, 2>, dim3(grid), dim3(256), 0, st, v.q, v.ldt, total, C, dst);
  else { tcx_set_error("eff_attn16: unsupported channel count %d", C); return -1; }
  return tcx_check_launch("ea16_qsoftmax");
}

// ---- tensor-core context path: statistics -> transposed split-major pack; the C x C contraction itself is a batched
// gemm_tc launch (one batch entry per (image, K-split)) and a small fold of the split partials ------------------------------
int ea16_ctx_tc_nsplit(int N) { return cdiv(N, EC_SPLIT_TOK); }
size_t ea16_ctx_tc_stats_floats(int B, int N, int C) { return 2 * (size_t)B * ea16_ctx_tc_nsplit(N) * C; }
void ea16_ctx_tc_splits(int N, int* KS, int* Ks) {
  const int tiles = cdiv(N, 64);
  const int ks = tiles < 8 ? tiles : 8;
  *KS = ks;
  *Ks = cdiv(tiles, ks) * 64;
}
int launch_ea16_packT(const Ea16View& v, int B, int N, int C, float* stats, void* Pt, void* Vt, cudaStream_t st) {
  TCX_REQUIRE(!v.reint, "ea16_packT: token-major K/V only");
  TCX_REQUIRE(C % 64 == 0 && C <= 512, "ea16_packT: C %% 64 == 0, C <= 512 (got C=%d)", C);
  TCX_REQUIRE(v.ldt % 8 == 0 && v.sb % 8 == 0, "ea16_packT: 16-byte row pitch needed");
  if (B == 0 || N == 0) return 0;
  int KS, Ks;
  ea16_ctx_tc_splits(N, &KS, &Ks);
  const int nsplit = ea16_ctx_tc_nsplit(N);
  float* pm = stats;
  float* ps = stats + (size_t)B * nsplit * C;
  {
    ProfScope prof("ea16_colstats", st, (double)B * N * C * 2);
    tcx_launch_pdl(ea16_colstats_kernel, dim3(nsplit, B), dim3(256), 0, st, v.k, v.sb, v.ldt, N, C, pm, ps);
    TCX_TRY(tcx_check_launch("ea16_colstats"));
  }
  ProfScope prof("ea16_packT", st, (double)B * C * (2.0 * N + 2.0 * KS * Ks) * 2);
  tcx_launch_pdl(ea16_packT_kernel, dim3(KS * Ks / 64, B, C / 64), dim3(256), 0, st, v.k, v.v, v.sb, v.ldt, N, C, Ks, KS, nsplit,
                 (const float*)pm, (const float*)ps, reinterpret_cast<__half*>(Pt), reinterpret_cast<__half*>(Vt));
  return tcx_check_launch("ea16_packT");
}
int launch_ea16_splitk_combine(const float* part, void* ctxT, int B, int KS, int C, cudaStream_t st) {
  if (B == 0) return 0;
  ProfScope prof("ea16_splitk_combine", st, (double)B * C * C * (4.0 * KS + 2.0));
  tcx_launch_pdl(ea16_splitk_combine_kernel, dim3(cdiv(C * C, 256), B), dim3(256), 0, st, part, reinterpret_cast<__half*>(ctxT), KS, C * C);
  return tcx_check_launch("ea16_splitk_combine");
}
