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
int launch_sk_pool(const void* t16, float* S, int B, int H1, int W1, int H2, int W2, int C, cudaStream_t st);
int launch_sk_weights(const float* S, const float* fcw, const float* fcb, const float* w0, const float* b0, const float* w1,
                      const float* b1, float* att, int B, int C, int d, cudaStream_t st);
int launch_sk_mix(const void* t16, const float* att, void* A, int B, int H1, int W1, int H2, int W2, int C, cudaStream_t st);
int launch_relu_bn(float* y, long long rows, int C, const BnParams& bn, cudaStream_t st);
