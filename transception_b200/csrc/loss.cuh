#pragma once
#include "common.cuh"

#define SEG_LOSS_KMAX 16
struct SegLossArgs {
  const float* logits;    // [B][K][HW] (NCHW) logits, or probabilities when softmax == 0
  const void* labels;     // [B][HW]
  int kind;               // 0 int64, 1 float32 (integer-valued, as the reference's label tensor), 2 int32, 3 uint8
  int B, K;
  long long HW;
  int softmax;
  float w_ce, w_dice;
  float cw[SEG_LOSS_KMAX];   // per-class Dice weights (utils.py:38-39)
};
size_t seg_loss_workspace_floats(int B, int K, long long HW);
int launch_seg_loss_fwd(const SegLossArgs& a, float* out, float* ws, cudaStream_t st);
int launch_seg_loss_bwd(const SegLossArgs& a, const float* ws, const float* grad_out, float* dlogits, cudaStream_t st);
int launch_argmax_classes(const float* logits, unsigned char* out, int B, int K, long long HW, cudaStream_t st);
