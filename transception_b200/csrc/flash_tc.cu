#include "common.cuh"
#include "attention.cuh"
int launch_flash_tc(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, cudaStream_t st) {
  return launch_flash_ffma(q, kv, out, B, Nq, Nk, scale, st);
}
