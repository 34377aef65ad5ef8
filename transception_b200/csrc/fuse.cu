// Glue kernels of the `networks/Transception.py` variant: MiT_3inception's dilated
// two-branch patch merging, FuseEfficientAttention's reinterpreted softmaxes and the nearest-upsample + concat in front
// of the 1x1 fusion conv.  Every contraction around them runs on the tcgen05 GEMM (gemm_tc.cu); these kernels only
// re-lay data so that the GEMM sees K-major fp16 operands with 16-byte pitches.  All are HBM/L2 streaming kernels.
#include <cuda_fp16.h>
#include "common.cuh"
#include "fuse.cuh"

namespace {

__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide reductions over 256 threads (8 warps); red: 8 floats of shared memory
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; i++) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; i++) r += red[i];      // fixed order: deterministic
  return r;
}

// x [B][H][W][Cin] fp32 (NHWC) -> rows [B*Ho*Wo][(ky, kx, cin)] fp16 of a k x k conv with stride / dilation / padding
// (EffSegformer.py:122 `nn.Conv2d(in_ch, dim, patch_size, stride, padding, dilation)`). One thread per 8 channels.
__global__ void __launch_bounds__(256) im2row16_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int H, int W, int Cin,
                                                       int k, int stride, int pad, int dil, int Ho, int Wo) {
  pdl_trigger();
  pdl_wait();
  const int c8n = Cin >> 3;
  const long long total = (long long)B * Ho * Wo * k * k * c8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int kx = (int)(t % k); t /= k;
  const int ky = (int)(t % k); t /= k;
  const int j = (int)(t % Wo); t /= Wo;
  const int i = (int)(t % Ho);
  const int b = (int)(t / Ho);
  const int yy = i * stride - pad + ky * dil, xx = j * stride - pad + kx * dil;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
    const float4* src = reinterpret_cast<const float4*>(x + (((long long)b * H + yy) * W + xx) * Cin + c8 * 8);
    const float4 a = src[0], c = src[1];
    o = make_uint4(pk2(a.x, a.y), pk2(a.z, a.w), pk2(c.x, c.y), pk2(c.z, c.w));
  }
  *reinterpret_cast<uint4*>(out + idx * 8) = o;
}

// LayerNorm of dense rows src [B*n][C] (fp32) scattered into a per-image token slab: dst + b*dst_bs + t*C.
// Warp per row, C <= 1024, C % 32 == 0 (EffSegformer.py:130 `nfx = self.norm(fx)`).
__global__ void __launch_bounds__(256) ln_scatter_kernel(const float* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bb,
                                                         float* __restrict__ dst, long long rows, int n, int C, long long dst_bs, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = src + row * C;
  float v[32];
  const int per = C >> 5;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    v[i] = i < per ? s[i * 32 + lane] : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i++)
    if (i < per) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  const long long b = row / n, t = row - b * n;
  float* d = dst + b * dst_bs + t * C;
#pragma unroll
  for (int i = 0; i < 32; i++)
    if (i < per) { const int c = i * 32 + lane; d[c] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(bb + c); }
}

// FuseEfficientAttention (Transception.py:54-57): the [N][C] key / value buffers are re-read as [C][N] ("rows" are
// runs of N consecutive halfs).  One block per (row c', image): Pk[b][c'][0..Np) = softmax over the run (fp16, zero
// padded to Np) and Vp[b][c'][0..Np) = the value run, zero padded — both K-major operands of the context GEMM.
__global__ void __launch_bounds__(256) fea_kpack_kernel(const __half* __restrict__ k16, const __half* __restrict__ v16, __half* __restrict__ Pk,
                                                        __half* __restrict__ Vp, int N, int C, int Np) {
  __shared__ float red[8];
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, b = blockIdx.y;
  const __half* kr = k16 + ((long long)b * C + c) * N;
  const __half* vr = v16 + ((long long)b * C + c) * N;
  __half* po = Pk + ((long long)b * C + c) * Np;
  __half* vo = Vp + ((long long)b * C + c) * Np;
  float m = -INFINITY;
  for (int n = threadIdx.x; n < N; n += 256) m = fmaxf(m, __half2float(kr[n]));
  m = block_max(m, red);
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) s += __expf(__half2float(kr[n]) - m);
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int n = threadIdx.x; n < Np; n += 256) {
    const bool live = n < N;
    po[n] = __float2half_rn(live ? __expf(__half2float(kr[n]) - m) * inv : 0.f);
    vo[n] = live ? vr[n] : __float2half_rn(0.f);
  }
}

// queries re-read as [C][N]; softmax over the C rows for every column n' (Transception.py:67-71), written transposed
// as QsT[b][n'][c'] — the A operand of  att = QsT x ctx.  Block = 32 columns, tile staged in shared memory.
constexpr int FQ_COLS = 32, FQ_PITCH = 34;
__global__ void __launch_bounds__(256) fea_qsoftmaxT_kernel(const __half* __restrict__ q16, __half* __restrict__ QsT, int N, int C) {
  extern __shared__ __half tile[];      // [C][FQ_PITCH]
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * FQ_COLS, b = blockIdx.y;
  const __half* q = q16 + (long long)b * C * N;
  const bool col_ok = n0 + lane < N;
  for (int c = warp; c < C; c += 8) tile[c * FQ_PITCH + lane] = col_ok ? q[(long long)c * N + n0 + lane] : __float2half_rn(0.f);
  __syncthreads();
  for (int j = warp * 4; j < warp * 4 + 4; j++) {
    if (n0 + j >= N) break;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, __half2float(tile[c * FQ_PITCH + j]));
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(__half2float(tile[c * FQ_PITCH + j]) - m);
    s = warp_sum(s);
    const float inv = 1.f / s;
    __half* o = QsT + ((long long)b * N + n0 + j) * C;
    for (int c = lane; c < C; c += 32) o[c] = __float2half_rn(__expf(__half2float(tile[c * FQ_PITCH + j]) - m) * inv);
  }
}

// t16 [B][n1+n2][C] fp16 tokens of the two branches -> A [B*H2*W2][2C]: channels [0,C) = branch-1 map (H1 x W1)
// nearest-upsampled to H2 x W2 (Transception.py:471 `F.interpolate`, default mode: src = min(floor(dst * in/out), in-1)
// evaluated in fp32 like ATen), channels [C,2C) = branch-2 map. One thread per 8 channels.
__global__ void __launch_bounds__(256) upcat16_kernel(const __half* __restrict__ t16, __half* __restrict__ A, int B, int H1, int W1, int H2,
                                                      int W2, int C, float sh, float sw) {
  pdl_trigger();
  pdl_wait();
  const int c8n = (2 * C) >> 3;
  const long long total = (long long)B * H2 * W2 * c8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int j = (int)(t % W2); t /= W2;
  const int i = (int)(t % H2);
  const int b = (int)(t / H2);
  const int n1 = H1 * W1, ntok = n1 + H2 * W2;
  const __half* src;
  if (c8 * 8 < C) {
    const int si = min((int)floorf((float)i * sh), H1 - 1), sj = min((int)floorf((float)j * sw), W1 - 1);
    src = t16 + ((long long)b * ntok + si * W1 + sj) * C + c8 * 8;
  } else {
    src = t16 + ((long long)b * ntok + n1 + i * W2 + j) * C + (c8 * 8 - C);
  }
  *reinterpret_cast<uint4*>(A + idx * 8) = *reinterpret_cast<const uint4*>(src);
}

// ---- SK_Block (Transception.py:306-358) on the two stage maps held as fp16 tokens t16 [B][n1+n2][C] ---------------------
__device__ __forceinline__ const __half* sk_src1(const __half* t16, int b, int i, int j, int H1, int W1, int ntok, int C, float sh, float sw) {
  const int si = min((int)floorf((float)i * sh), H1 - 1), sj = min((int)floorf((float)j * sw), W1 - 1);
  return t16 + ((long long)b * ntok + si * W1 + sj) * C;
}
// S[b][c] = mean over the H2 x W2 positions of (upsampled branch-1 + branch-2)  (:336-339).  grid (C/64, B), 256 threads =
// 64 channels x 4 position slices, fixed-order combine (deterministic).
__global__ void __launch_bounds__(256) sk_pool_kernel(const __half* __restrict__ t16, float* __restrict__ S, int H1, int W1, int H2, int W2,
                                                      int C, float sh, float sw) {
  __shared__ float part[4][64];
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), slice = threadIdx.x >> 6, b = blockIdx.y;
  const int n1 = H1 * W1, n2 = H2 * W2, ntok = n1 + n2;
  float acc = 0.f;
  for (int pos = slice; pos < n2; pos += 4) {
    const int i = pos / W2, j = pos - i * W2;
    acc += __half2float(sk_src1(t16, b, i, j, H1, W1, ntok, C, sh, sw)[c]) + __half2float(t16[((long long)b * ntok + n1 + pos) * C + c]);
  }
  part[slice][threadIdx.x & 63] = acc;
  __syncthreads();
  if (slice == 0) {
    const int l = threadIdx.x;
    S[(long long)b * C + c] = (part[0][l] + part[1][l] + part[2][l] + part[3][l]) / (float)n2;
  }
}
// Z = fc(S); w_i = fcs_i(Z); a = softmax over the two paths (:340-351).  One block per image; att [B][2][C].
__global__ void __launch_bounds__(256) sk_weights_kernel(const float* __restrict__ S, const float* __restrict__ fcw, const float* __restrict__ fcb,
                                                         const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, float* __restrict__ att, int C, int d) {
  extern __shared__ float sm[];      // S row [C] | Z [d]
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* srow = sm;
  float* z = sm + C;
  for (int c = threadIdx.x; c < C; c += 256) srow[c] = S[(long long)b * C + c];
  __syncthreads();
  for (int k = warp; k < d; k += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(srow[c], __ldg(fcw + (long long)k * C + c), a);
    a = warp_sum(a);
    if (lane == 0) z[k] = a + __ldg(fcb + k);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float u0 = __ldg(b0 + c), u1 = __ldg(b1 + c);
    for (int k = 0; k < d; k++) {
      u0 = fmaf(z[k], __ldg(w0 + (long long)c * d + k), u0);
      u1 = fmaf(z[k], __ldg(w1 + (long long)c * d + k), u1);
    }
    const float m = fmaxf(u0, u1), e0 = __expf(u0 - m), e1 = __expf(u1 - m), inv = 1.f / (e0 + e1);
    att[((long long)b * 2 + 0) * C + c] = e0 * inv;
    att[((long long)b * 2 + 1) * C + c] = e1 * inv;
  }
}
// V = a0 * upsampled branch-1 + a1 * branch-2 (:354) as the fp16 A operand [B*n2][C] of the 1x1 conv. Thread per 8 channels.
__global__ void __launch_bounds__(256) sk_mix_kernel(const __half* __restrict__ t16, const float* __restrict__ att, __half* __restrict__ A, int B,
                                                     int H1, int W1, int H2, int W2, int C, float sh, float sw) {
  pdl_trigger();
  pdl_wait();
  const int c8n = C >> 3;
  const long long total = (long long)B * H2 * W2 * c8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int j = (int)(t % W2); t /= W2;
  const int i = (int)(t % H2);
  const int b = (int)(t / H2);
  const int n1 = H1 * W1, ntok = n1 + H2 * W2;
  const uint4 r1 = *reinterpret_cast<const uint4*>(sk_src1(t16, b, i, j, H1, W1, ntok, C, sh, sw) + c8 * 8);
  const uint4 r2 = *reinterpret_cast<const uint4*>(t16 + ((long long)b * ntok + n1 + i * W2 + j) * C + c8 * 8);
  const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
  const __half2* h2 = reinterpret_cast<const __half2*>(&r2);
  const float* a0 = att + ((long long)b * 2 + 0) * C + c8 * 8;
  const float* a1 = att + ((long long)b * 2 + 1) * C + c8 * 8;
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float2 x1 = __half22float2(h1[q]), x2 = __half22float2(h2[q]);
    o[q] = pk2(a0[2 * q] * x1.x + a1[2 * q] * x2.x, a0[2 * q + 1] * x1.y + a1[2 * q + 1] * x2.y);
  }
  *reinterpret_cast<uint4*>(A + idx * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}
// y = BatchNorm_eval(ReLU(y)) in place on [rows][C] fp32 (:322-325: conv -> ReLU -> BatchNorm2d)
__global__ void __launch_bounds__(256) relu_bn_kernel(float* __restrict__ y, long long total, int C, BnParams bn) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  float scale, shift;
  bn_fold(bn, c, scale, shift);
  y[idx] = fmaxf(y[idx], 0.f) * scale + shift;
}

}  // namespace

int launch_im2row16(const float* x, void* out16, int B, int H, int W, int Cin, int k, int stride, int pad, int dil, int Ho, int Wo,
                    cudaStream_t st) {
  TCX_REQUIRE(Cin % 8 == 0, "im2row16: Cin must be a multiple of 8 (got %d)", Cin);
  const long long total = (long long)B * Ho * Wo * k * k * (Cin / 8);
  if (total == 0) return 0;
  ProfScope prof("im2row16", st, (double)total * 8 * 2 + (double)B * H * W * Cin * 4);
  tcx_launch_pdl(im2row16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, reinterpret_cast<__half*>(out16), B, H, W,
                 Cin, k, stride, pad, dil, Ho, Wo);
  return tcx_check_launch("im2row16");
}

int launch_ln_scatter(const float* src, const float* w, const float* b, float* dst, int B, int n, int C, long long dst_bs, float eps,
                      cudaStream_t st) {
  TCX_REQUIRE(C % 32 == 0 && C <= 1024, "ln_scatter: C must be a multiple of 32, at most 1024 (got %d)", C);
  const long long rows = (long long)B * n;
  if (rows == 0) return 0;
  ProfScope prof("ln_scatter", st, (double)rows * C * 8);
  tcx_launch_pdl(ln_scatter_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, src, w, b, dst, rows, n, C, dst_bs, eps);
  return tcx_check_launch("ln_scatter");
}

int launch_fea_kpack(const void* k16, const void* v16, void* Pk, void* Vp, int B, int N, int C, int Np, cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  ProfScope prof("fea_kpack", st, (double)B * C * (2.0 * N + 2.0 * Np) * 2);
  tcx_launch_pdl(fea_kpack_kernel, dim3(C, B), dim3(256), 0, st, reinterpret_cast<const __half*>(k16), reinterpret_cast<const __half*>(v16),
                 reinterpret_cast<__half*>(Pk), reinterpret_cast<__half*>(Vp), N, C, Np);
  return tcx_check_launch("fea_kpack");
}

int launch_fea_qsoftmaxT(const void* q16, void* QsT, int B, int N, int C, cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  const size_t smem = (size_t)C * FQ_PITCH * sizeof(__half);
  TCX_REQUIRE(smem <= 48 * 1024, "fea_qsoftmaxT: C = %d too large", C);
  ProfScope prof("fea_qsoftmaxT", st, (double)B * C * N * 4);
  tcx_launch_pdl(fea_qsoftmaxT_kernel, dim3(cdiv(N, FQ_COLS), B), dim3(256), smem, st, reinterpret_cast<const __half*>(q16),
                 reinterpret_cast<__half*>(QsT), N, C);
  return tcx_check_launch("fea_qsoftmaxT");
}

int launch_upcat16(const void* t16, void* A, int B, int H1, int W1, int H2, int W2, int C, cudaStream_t st) {
  TCX_REQUIRE(C % 8 == 0, "upcat16: C must be a multiple of 8");
  const long long total = (long long)B * H2 * W2 * (2 * C / 8);
  if (total == 0) return 0;
  ProfScope prof("upcat16", st, (double)total * 32);
  tcx_launch_pdl(upcat16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const __half*>(t16),
                 reinterpret_cast<__half*>(A), B, H1, W1, H2, W2, C, (float)H1 / (float)H2, (float)W1 / (float)W2);
  return tcx_check_launch("upcat16");
}

int launch_sk_pool(const void* t16, float* S, int B, int H1, int W1, int H2, int W2, int C, cudaStream_t st) {
  TCX_REQUIRE(C % 64 == 0, "sk_pool: C must be a multiple of 64");
  if (B == 0) return 0;
  ProfScope prof("sk_pool", st, (double)B * H2 * W2 * C * 4);
  tcx_launch_pdl(sk_pool_kernel, dim3(C / 64, B), dim3(256), 0, st, reinterpret_cast<const __half*>(t16), S, H1, W1, H2, W2, C,
                 (float)H1 / (float)H2, (float)W1 / (float)W2);
  return tcx_check_launch("sk_pool");
}
int launch_sk_weights(const float* S, const float* fcw, const float* fcb, const float* w0, const float* b0, const float* w1,
                      const float* b1, float* att, int B, int C, int d, cudaStream_t st) {
  if (B == 0) return 0;
  ProfScope prof("sk_weights", st, (double)B * C * 12 + 3.0 * C * d * 4);
  tcx_launch_pdl(sk_weights_kernel, dim3(B), dim3(256), (size_t)(C + d) * sizeof(float), st, S, fcw, fcb, w0, b0, w1, b1, att, C, d);
  return tcx_check_launch("sk_weights");
}
int launch_sk_mix(const void* t16, const float* att, void* A, int B, int H1, int W1, int H2, int W2, int C, cudaStream_t st) {
  TCX_REQUIRE(C % 8 == 0, "sk_mix: C must be a multiple of 8");
  const long long total = (long long)B * H2 * W2 * (C / 8);
  if (total == 0) return 0;
  ProfScope prof("sk_mix", st, (double)total * 48);
  tcx_launch_pdl(sk_mix_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const __half*>(t16), att,
                 reinterpret_cast<__half*>(A), B, H1, W1, H2, W2, C, (float)H1 / (float)H2, (float)W1 / (float)W2);
  return tcx_check_launch("sk_mix");
}
int launch_relu_bn(float* y, long long rows, int C, const BnParams& bn, cudaStream_t st) {
  const long long total = rows * C;
  if (total == 0) return 0;
  ProfScope prof("relu_bn", st, (double)total * 8);
  tcx_launch_pdl(relu_bn_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, y, total, C, bn);
  return tcx_check_launch("relu_bn");
}
