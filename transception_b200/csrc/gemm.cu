// GEMM front end: C[M,N] = epi(A[M,K] * W[N,K]^T), A and W both K-major (nn.Linear storage).
// Two back ends on the same device: the tcgen05/TMA kernel (gemm_tc.cu) for tensor-core sized
// problems and this FFMA kernel for the small/ragged ones (K not a multiple of 32, N < 16, M tiny).
#include "common.cuh"

int launch_gemm_tc(const GemmParams& p, cudaStream_t st);  // gemm_tc.cu
bool gemm_tc_eligible(const GemmParams& p);

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ float epilogue(float v, int n, const GemmEpi& e) {
  if (e.bias) v += e.bias[n];
  if (e.bn.w) {
    float s, t;
    bn_fold(e.bn, n, s, t);
    v = v * s + t;
  }
  return apply_act(v, e.act);
}

__global__ void __launch_bounds__(256) gemm_ffma_kernel(const GemmParams p) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Ws[BK][BN + PAD];
  const int z = blockIdx.z;
  const int gi = z / p.batch, bi = z - gi * p.batch;
  const GemmGroup& g = p.g[gi];
  const float* __restrict__ A = g.A + (long long)bi * p.strideA;
  const float* __restrict__ W = g.W + (long long)bi * p.strideW;
  float* __restrict__ C = g.C + (long long)bi * p.strideC;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row 0..63, k offset 0,4,8,12

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
    const int kk = k0 + lk;
    if (m0 + lr < p.M && kk < p.K) a = *reinterpret_cast<const float4*>(A + (long long)(m0 + lr) * p.lda + kk);
    if (n0 + lr < p.N && kk < p.K) w = *reinterpret_cast<const float4*>(W + (long long)(n0 + lr) * p.ldw + kk);
    As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
    Ws[lk + 0][lr] = w.x; Ws[lk + 1][lr] = w.y; Ws[lk + 2][lr] = w.z; Ws[lk + 3][lr] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }

  const GemmEpi& e = g.epi;
  const float* __restrict__ R = e.residual ? e.residual + (long long)bi * e.strideR : nullptr;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = epilogue(acc[i][j], n, e);
      if (R) v += R[(long long)m * e.ldr + n];
      C[(long long)m * p.ldc + n] = v;
    }
  }
}

}  // namespace

int launch_gemm_ffma(const GemmParams& p, cudaStream_t st) {
  TCX_REQUIRE(p.K % 4 == 0 && p.lda % 4 == 0 && p.ldw % 4 == 0, "gemm: K/lda/ldw must be multiples of 4 (K=%d lda=%d ldw=%d)",
              p.K, p.lda, p.ldw);
  for (int i = 0; i < p.groups; i++)
    TCX_REQUIRE((((uintptr_t)p.g[i].A | (uintptr_t)p.g[i].W) & 15) == 0, "gemm: A/W must be 16-byte aligned");
  TCX_REQUIRE(((p.strideA | p.strideW) & 3) == 0, "gemm: batch strides must be multiples of 4");
  dim3 grid(cdiv(p.M, BM), cdiv(p.N, BN), p.groups * p.batch);
  ProfScope prof("gemm_ffma", st);
  gemm_ffma_kernel<<<grid, 256, 0, st>>>(p);
  return tcx_check_launch("gemm_ffma");
}

int launch_gemm(const GemmParams& p, cudaStream_t st) {
  TCX_REQUIRE(p.groups >= 1 && p.groups <= TCX_MAX_GROUPS && p.batch >= 1, "gemm: bad groups/batch");
  if (p.M == 0 || p.N == 0) return 0;
  if (gemm_tc_eligible(p)) return launch_gemm_tc(p, st);
  TCX_REQUIRE(!p.ab16 && !p.out16, "gemm: fp16 operands/outputs need the tensor-core kernel (M=%d N=%d K=%d not eligible)",
              p.M, p.N, p.K);
  return launch_gemm_ffma(p, st);
}

int launch_linear(const float* A, const float* W, const float* bias, const float* residual, float* C, int M, int N,
                  int K, int act, cudaStream_t st) {
  GemmParams p{};
  p.groups = 1; p.batch = 1;
  p.M = M; p.N = N; p.K = K; p.lda = K; p.ldw = K; p.ldc = N;
  p.g[0].A = A; p.g[0].W = W; p.g[0].C = C;
  p.g[0].epi.bias = bias; p.g[0].epi.act = act; p.g[0].epi.residual = residual; p.g[0].epi.ldr = N;
  return launch_gemm(p, st);
}
