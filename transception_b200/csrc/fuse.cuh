#pragma once
#include "common.cuh"

// glue kernels of the networks/Transception.py variant (fuse.cu)
int launch_im2row16(const float* x, void* out16, int B, int H, int W, int Cin, int k, int stride, int pad, int dil, int Ho, int Wo,
                    cudaStream_t st);
int launch_ln_scatter(const float* src, const float* w, const float* b, float* dst, int B, int n, int C, long long dst_bs, float eps,
                      cudaStream_t st);
int launch_fea_kpack(const void* k16, const void* v16, void* Pk, void* Vp, int B, int N, int C, int Np, cudaStream_t st);
int launch_fea_qsoftmaxT(const void* q16, void* QsT, int B, int N, int C, cudaStream_t st);
int launch_upcat16(const void* t16, void* A, int B, int H1, int W1, int H2, int W2, int C, cudaStream_t st);
