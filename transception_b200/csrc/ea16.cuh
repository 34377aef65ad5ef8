#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

// K/Q/V of the efficient attention in fp16.
//  tokens-major (reint = 0): element (n, c) at ptr + b*sb + n*ldt + c
//  reinterpreted (reint = 1, MSTr.py:2312-2314): element (c', n') at ptr + b*sb + c'*N + n'
struct Ea16View {
  const __half* k;
  const __half* q;
  const __half* v;
  long long sb;
  int ldt;
  int reint;
};
size_t ea16_workspace_floats(int B, int N, int C);
int launch_ea16_context(const Ea16View& v, int B, int N, int C, float* ws, __half* ctxT, cudaStream_t st);
int launch_ea16_qsoftmax(const Ea16View& v, int B, int N, int C, __half* dst, cudaStream_t st);

// tensor-core context path (token-major K/V): column-softmax statistics, then Pt / Vt = K-major split-major
// [B][KS][C][Ks] operands; ctx partial[b][s] = Vt[b][s] x Pt[b][s]^T is a batched gemm_tc launch in api.cu, folded over s
int ea16_ctx_tc_nsplit(int N);
size_t ea16_ctx_tc_stats_floats(int B, int N, int C);
void ea16_ctx_tc_splits(int N, int* KS, int* Ks);
int launch_ea16_packT(const Ea16View& v, int B, int N, int C, float* stats, void* Pt, void* Vt, cudaStream_t st);
int launch_ea16_splitk_combine(const float* part, void* ctxT, int B, int KS, int C, cudaStream_t st);
