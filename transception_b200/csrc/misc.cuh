#pragma once
#include "common.cuh"

int launch_patch_embed_ln(const float* x, long long xs_b, long long xs_c, int B, int Hin, int Win, const float* w,
                          const float* bias, const float* lnw, const float* lnb, float eps, float* out, cudaStream_t st);

// lnw == null: the conv output without the LayerNorm.  launch_patch_im2row: the conv's patch matrix [B*Ho*Wo][Kp] (Kp >= 147)
int launch_patch_im2row(const float* x, long long xs_b, long long xs_c, int B, int Hin, int Win, int Kp, float* A, cudaStream_t st);

struct RegroupArgs {
  const float* src[4];   // NHWC maps, dense
  float* dst;            // [B][ntok][64]
  int tok_off[5];        // token offsets of the four slabs (+ total)
  int ntok, B;
  const float* res;      // optional [B][ntok][64] added to the regrouped tokens (the skip connection of BridgLayer_4, MSTr.py:2405)
};
int launch_regroup(const RegroupArgs& a, cudaStream_t st);
// the inverse: token buffer -> four dense per-scale slabs [B][n_k][64] (training row: the Mix-FFN inputs of a bridge layer, the
// maps handed to the decoder, and the gradient of launch_regroup)
struct UngroupArgs {
  const float* src;      // [B][ntok][64]
  float* dst[4];
  int tok_off[5];
  int ntok, B;
};
int launch_ungroup(const UngroupArgs& a, cudaStream_t st);

int launch_sr_im2row(const float* x, long long xs_b, int HW, int Cin, int r, int B, float* A, cudaStream_t st);

struct SrPackArgs {
  const float* conv[3];  // conv outputs [B*pp][64*g]
  const float* x;        // token buffer (raw stage-4 tokens are copied through)
  long long xs_b;
  int raw_tok0;
  int red_off[4];        // offsets of the 4 parts in the reduced sequence
  int gmul[3], pp[3];
  int nred, B;
  const float* lnw;
  const float* lnb;
  float eps;
  float* out;            // [B][nred][64]
};
int launch_sr_pack_ln(const SrPackArgs& a, cudaStream_t st);      // lnw == null: pack only (the training row normalises separately)
// gradient of the packing: d(reduced sequence) -> d(conv outputs) and the raw stage-4 rows of d(token buffer)
struct SrUnpackArgs {
  const float* dred;     // [B][nred][64]
  float* dconv[3];       // [B*pp][64*g]
  float* dx;             // token-buffer gradient: rows raw_tok0.. of every image are written
  long long xs_b;
  int raw_tok0;
  int red_off[4];
  int gmul[3], pp[3];
  int nred, B;
};
int launch_sr_unpack(const SrUnpackArgs& a, cudaStream_t st);
// gradient of launch_sr_im2row (non-overlapping patches: a permutation): dA [B*P*P][Cin*r*r] -> the slab of d(token buffer)
int launch_sr_row2im(const float* dA, long long xs_b, int HW, int Cin, int r, int B, float* dx, cudaStream_t st);

int launch_sr_im2row16(const void* x16, long long xs_b, int HW, int Cin, int r, int B, void* A16, cudaStream_t st);
int launch_conv_weight_perm16(const float* w, void* o16, int N, int Cin, int r, cudaStream_t st);
int launch_sr_pack_ln16(const SrPackArgs& a, const void* x16, void* out16, cudaStream_t st);

struct IffSrc {
  const float* p[4];
};
int launch_iff_pool(const IffSrc& src, int B, int HW, int C, float* pooled, cudaStream_t st);
int launch_iff_gate(const IffSrc& src, int B, int H, int W, int C, const float* ah, const float* aw, float* out,
                    cudaStream_t st);

int launch_shuffle_ln(const float* in, int B, int H, int W, int s, int c, const float* lnw, const float* lnb, float eps,
                      float* out, cudaStream_t st);
int launch_final_head(const float* in, int B, int H, int W, const float* lnw, const float* lnb, float eps,
                      const float* cw, const float* cb, int ncls, float* out, cudaStream_t st);

// backward of launch_final_head: de in the layout of e; part: final_head_bwd_part_floats floats (block partials of the parameter
// sums, folded into the four parameter gradients by launch_final_head_bwd_fold — any stream ordered after the first launch)
size_t final_head_bwd_part_floats(int B, int H, int W);
int launch_final_head_bwd(const float* e, const float* dlogits, int B, int H, int W, const float* lnw, float eps, const float* cw,
                          int ncls, float* de, float* part, int* nblk_out, cudaStream_t st);
int launch_final_head_bwd_fold(float* part, int nblk, const float* lnw, const float* lnb, const float* cw, int ncls, float* dlnw,
                               float* dlnb, float* dcw, float* dcb, cudaStream_t st);

// fused FinalPatchExpand_X4 + LayerNorm + class head (head_tc.cu)
bool head_tc_eligible(const float* x, const float* w, int ncls);
size_t head_tc_workspace_floats();
int launch_head_tc(const float* x, const float* w, int B, int H, int W, const float* lnw, const float* lnb, float eps,
                   const float* cw, const float* cb, int ncls, float* out, float* ws, cudaStream_t st);
