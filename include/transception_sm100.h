/* transception_sm100.h — C ABI of libtransception_sm100.so (sm_100a only).
 *
 * The reference (xmindflow/TransCeption) has no FFI: its plugin surface for the hot path is the nn.Module
 * forward of networks/MSTr.py.  Each entry point below replaces the ATen op sequence of ONE reference
 * forward (cited per function, lines of /root/reference/networks/MSTr.py) and is what a cgo/JNI/ctypes
 * style binding of that forward would call.  The Python binding used by this repo is
 * transception_b200/ops.py (ctypes); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every tensor is dense fp32 in device memory, tokens-major [B,N,C] / NHWC [B,H,W,C] unless stated;
 *  - the library never allocates, frees or retains device memory: outputs and `ws` scratch are caller-owned,
 *    sized by the matching *_workspace_bytes() call; pointers must be 16-byte aligned;
 *  - kernels are enqueued on `stream` (a cudaStream_t) of the current device, no host synchronisation
 *    (safe under CUDA-graph capture);
 *  - return value 0 = OK; non-zero = failure, message via tcx_last_error() (thread-local);
 *  - BatchNorm arguments are (weight, bias, running_mean, running_var, eps): inference statistics.
 *  - `p` arguments are host arrays of device pointers in the documented slot order.
 */
#ifndef TRANSCEPTION_SM100_H
#define TRANSCEPTION_SM100_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* tcx_version(void);
const char* tcx_last_error(void);
/* 1 when a usable sm_100 device is current, else 0 (with tcx_last_error set) */
int tcx_device_ok(void);
/* back-end switches for A/B measurement: name in {"gemm_tc","flash_tc","f16_pipeline","fork","pdl","pdl_chain","mixtail","ea_tc","wgrad_tc",
 * "max_ctas","smem_kb","wgrad_ctas","wgrad_idle"}; returns the previous value.  "pdl" = programmatic dependent launch on the tensor-core pipeline
 * kernels; "pdl_chain" (library default 0, the Python host sets ops.PDL_CHAIN) = the same attribute on the CUDA-core kernels of the training row. */
int tcx_set_flag(const char* name, int value);

/* number of kernels this library has enqueued so far in this process (bench.py's gpu_launches) */
long long tcx_launch_count(void);
/* per-kernel timing for bench.py's roofline leg: record a CUDA-event pair on the launching stream around every
 * launch of the named kernel ("flash_tc", "gemm_tc", "gemm_ffma", "flash_ffma", "mixffn_mid"); "" disables.
 * tcx_profile_read synchronises on the recorded events, returns their summed duration and count, and resets. */
int tcx_profile_enable(const char* kernel_name);
int tcx_profile_read(double* total_ms, int* count);
/* same, plus the summed algorithmic work of the timed launches: bytes for the HBM-bound kernels ("gemm_tc", "dwln",
 * "ln16", "mb_fused16", "ea16_ctx"), FLOPs for "flash_tc" */
int tcx_profile_read_work(double* total_ms, int* count, double* work);

/* K10 — nn.LayerNorm over the last dim (MSTr.py:153,156,1671,2360,2366; eps 1e-6 at :932-933) */
int tcx_layernorm_fwd(const float* x, const float* w, const float* b, float* y, long long M, int C, float eps,
                      void* stream);

/* nn.Linear: y[M,N] = act(x[M,K] w[N,K]^T + bias) + residual;  act: 0 none 1 GELU(erf) 2 Hardswish 3 sigmoid */
int tcx_linear_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N,
                   int K, int act, void* stream);
/* fp16-operand form of nn.Linear (tcgen05 kind::f16, fp32 accumulate; fp16 has TF32's 10-bit mantissa):
 * x16 [M,K] and w16 [N,K] are fp16 (tcx_f32_to_f16), y is fp32 (+bias, +fp32 residual) or, with out_f16, fp16 (+bias). */
int tcx_f32_to_f16(const float* src, void* dst, long long n, void* stream);
int tcx_linear_f16_fwd(const void* x16, const void* w16, const float* bias, const float* residual, void* y, int M, int N,
                       int K, int out_f16, void* stream);
/* Prepared weights: the fp16-intermediate pipeline (default) needs an fp16 copy of every GEMM weight matrix
 * (nn.Linear / 1x1 conv `*_w` slots of the `p` tables).  tcx_prepare_weight_f16 converts w32 into the caller-owned
 * buffer w16 and records the pair; an op whose matrices are all prepared runs with fp16 tensor-core operands and fp16
 * intermediates (fp32 residual streams, fp32 accumulation), otherwise it runs its fp32-storage / TF32 form.  This
 * registry is the only state the library keeps: w16 must stay valid until tcx_forget_weight(w32), and must be
 * re-prepared when the values at w32 change.  Flag "f16_pipeline" = 0 disables the lookup. */
int tcx_prepare_weight_f16(const float* w32, void* w16, long long numel, void* stream);
/* strided patchify conv weight [N][Cin][r][r] (Scale_reduce sr0/sr1/sr2): prepared as [N][(ky,kx,cin)] fp16 */
int tcx_prepare_conv_weight_f16(const float* w32, void* w16, int N, int Cin, int r, void* stream);
int tcx_forget_weight(const float* w32);
/* Conv2d_BN 1x1 (MSTr.py:399-404): y = act(BN(x w^T)) */
int tcx_linear_bn_act_fwd(const float* x, const float* w, const float* bn_w, const float* bn_b, const float* bn_rm,
                          const float* bn_rv, float bn_eps, int act, float* y, int M, int N, int K, void* stream);

/* K9 — OverlapPatchEmbeddings.forward (MSTr.py:299-304): conv7x7/4 pad 3 (3->64) + LayerNorm.
 * x is NCHW [B,Cin,H,W] with Cin in {1,3}; Cin==1 is read as three identical planes (MSTr.py:2828-2829). */
int tcx_patch_embed_ln_fwd(const float* x, int B, int Cin, int H, int W, const float* w, const float* bias,
                           const float* lnw, const float* lnb, float eps, float* out, void* stream);

/* K3 — ConvPosEnc / DWConv (MSTr.py:744-752, :26-31): y = dw3x3(x)+b (+x when add_input) */
int tcx_dwconv_tokens_fwd(const float* x, const float* w, const float* b, float* y, int B, int H, int W, int C,
                          int add_input, void* stream);

/* K8 — EfficientAttention.forward (MSTr.py:106-143) and, with reinterpret=1, M_EfficientChannelAtten.forward
 * (MSTr.py:2309-2353, raw [N,C]->[C,N] reinterpretation).  y = residual + reproj(att(xn)).
 * p = {k_w,k_b,q_w,q_b,v_w,v_b,reproj_w,reproj_b} */
size_t tcx_eff_attn_workspace_bytes(int B, int N, int C);
int tcx_eff_attn_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int N, int C,
                     int reinterpret, void* ws, void* stream);

/* K1 — MixFFN_skip.forward (MSTr.py:58-61): y = residual + fc2(GELU(LN(dw3x3(fc1 x)+fc1 x)))
 * p = {fc1_w,fc1_b,dw_w,dw_b,ln_w,ln_b,fc2_w,fc2_b} */
size_t tcx_mixffn_skip_workspace_bytes(int B, int N, int C4);
int tcx_mixffn_skip_fwd(const float* xn, const void* const* p, float ln_eps, const float* residual, float* y, int B,
                        int H, int W, int C, int C4, void* ws, void* stream);

/* K2 — FactorAtt_ConvRelPosEnc.forward (MSTr.py:852-886, ConvRelPosEnc :801-823)
 * p = {qkv_w,qkv_b,crpe_w3,crpe_b3,crpe_w5,crpe_b5,crpe_w7,crpe_b7,proj_w,proj_b}; y = residual + proj(...) */
size_t tcx_mb_factor_attn_workspace_bytes(int B, int N, int C);
int tcx_mb_factor_attn_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int H, int W,
                           int C, int heads, void* ws, void* stream);

/* K2+K3+K1 — G parallel branches x L chained MHCABlock.forward (MSTr.py:935-946): x_in[G][B][N][C] -> x (x_in may equal x).
 * p holds G*L blocks of TCX_MHCA_NP pointers:
 * {cpe_w,cpe_b,n1_w,n1_b,qkv_w,qkv_b,crpe_w3,crpe_b3,crpe_w5,crpe_b5,crpe_w7,crpe_b7,proj_w,proj_b,
 *  n2_w,n2_b,fc1_w,fc1_b,dw_w,dw_b,mlp_ln_w,mlp_ln_b,fc2_w,fc2_b}, block (g,l) at p[(g*L+l)*TCX_MHCA_NP] */
#define TCX_MHCA_NP 24
size_t tcx_mhca_blocks_workspace_bytes(int G, int B, int N, int C);
int tcx_mhca_blocks_fwd(const float* x_in, float* x, const void* const* p, int G, int L, int B, int H, int W, int C,
                        int heads, float ln_eps, float mlp_ln_eps, void* ws, void* stream);

/* K4 — DWConv2d_BN.forward (MSTr.py:355-362): Hardswish(BN(pw1x1(dw3x3_stride(x)))), NHWC */
size_t tcx_ripm_dwsep_bn_hs_workspace_bytes(int B, int H, int W, int C, int stride);
int tcx_ripm_dwsep_bn_hs_fwd(const float* x, const float* dw_w, const float* pw_w, const float* bn_w,
                             const float* bn_b, const float* bn_rm, const float* bn_rv, float bn_eps, float* y, int B,
                             int H, int W, int C, int stride, void* ws, void* stream);

/* K5 — ResBlock.forward (MSTr.py:1042-1050), NHWC.
 * p = {c1_w, bn1_w,bn1_b,bn1_rm,bn1_rv, dw_w, bn2_w,bn2_b,bn2_rm,bn2_rv, c2_w, bn3_w,bn3_b,bn3_rm,bn3_rv} */
size_t tcx_resblock_workspace_bytes(int B, int H, int W, int C);
int tcx_resblock_fwd(const float* x, const void* const* p, float bn_eps, float* y, int B, int H, int W, int C,
                     void* ws, void* stream);

/* K6 — IFF: CoordAtt.forward (MSTr.py:1322-1348) over the channel concatenation of nsrc NHWC maps of C channels
 * each (nsrc*C = inp; the concatenation is never materialised).
 * p = {conv1_w,conv1_b,bn_w,bn_b,bn_rm,bn_rv,convh_w,convh_b,convw_w,convw_b,out_w,out_b} */
size_t tcx_iff_coordatt_workspace_bytes(int B, int HW, int C, int mip);
int tcx_iff_coordatt_fwd(const void* const* maps, const void* const* p, float bn_eps, float* y, int B, int HW, int C,
                         int mip, int Cout, void* ws, void* stream);

/* K10 — BridgLayer_4 tokenisation (MSTr.py:2380-2386): four NHWC maps [B,S/2^k,S/2^k,{64,128,320,512}]
 * -> [B, Ntok, 64]; S = side of the stage-1 map (56 at 224x224). */
int tcx_bridge_regroup_fwd(const void* const* maps, float* tokens, int B, int S, void* stream);

/* Scale_reduce.forward (MSTr.py:2225-2249) -> [B, Nred, 64].  p = {sr0_w,sr0_b,sr1_w,sr1_b,sr2_w,sr2_b,ln_w,ln_b} */
size_t tcx_scale_reduce_workspace_bytes(int B, int S);
int tcx_scale_reduce_fwd(const float* x, const void* const* p, float ln_eps, float* out, int B, int S, void* ws,
                         void* stream);

/* K7 — M_EfficientSelfAtten.forward (MSTr.py:2267-2292): y = residual + proj(softmax(q k^T / 8) v)
 * p = {q_w,q_b,kv_w,kv_b,proj_w,proj_b, sr0_w,sr0_b,sr1_w,sr1_b,sr2_w,sr2_b,srln_w,srln_b} */
size_t tcx_bridge_sr_attn_workspace_bytes(int B, int S);
int tcx_bridge_sr_attn_fwd(const float* xn, const void* const* p, float scale, float ln_eps, const float* residual,
                           float* y, int B, int S, void* ws, void* stream);

/* the softmax(q k^T * scale) v core of K7 alone (MSTr.py:2281-2285), exposed for parity tests at ragged sizes:
 * q [B][Nq][64], kv [B][Nk][128] (k | v), out [B][Nq][64]; tcgen05 kernel unless the "flash_tc" flag is 0. */
size_t tcx_flash_attn_workspace_bytes(int B, int Nk);
int tcx_flash_attn_fwd(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, void* ws,
                       void* stream);

/* fp16 form used by the fp16 pipeline: q16 [B][Nq][64], kv16 [B][Nk][128] (k | v), out16 [B][Nq][64] are fp16; q and k
 * tiles arrive by TMA straight from these buffers (ws only holds the transposed V). tcgen05 kernel only. */
int tcx_flash_attn_f16_fwd(const void* q16, const void* kv16, void* out16, int B, int Nq, int Nk, float scale, void* ws,
                           void* stream);

/* BridgLayer_4.forward tail (MSTr.py:2394-2406): y = tx1 + cat_k MixFFN_k(tx slab k).  p = 4 x the K1 slots. */
size_t tcx_bridge_mixffn_workspace_bytes(int B, int S);
int tcx_bridge_mixffn_fwd(const float* tx, const float* tx1, const void* const* p, float ln_eps, float* y, int B,
                          int S, void* ws, void* stream);

/* EfficientTransformerBlock.forward (MSTr.py:164-173): tx = x + Attn(LN1 x); y = tx + MixFFN(LN2 tx).
 * p = {n1_w,n1_b, k_w,k_b,q_w,q_b,v_w,v_b,reproj_w,reproj_b, n2_w,n2_b, fc1_w,fc1_b,dw_w,dw_b,ln_w,ln_b,fc2_w,fc2_b} */
size_t tcx_eff_block_workspace_bytes(int B, int N, int C);
int tcx_eff_block_fwd(const float* x, const void* const* p, float ln_eps, float mlp_ln_eps, float* y, int B, int H, int W,
                      int C, void* ws, void* stream);

/* BridgLayer_4.forward (MSTr.py:2373-2409) on the token buffer x [B][Ntok][64] (tcx_bridge_regroup_fwd) -> y.
 * p (TCX_BRIDGE_NP slots) = {n1_w,n1_b, attn[14], n2_w,n2_b, 4 x {fc1_w,fc1_b,dw_w,dw_b,ln_w,ln_b,fc2_w,fc2_b}};
 * attn = the tcx_bridge_sr_attn_fwd table, or with channel_att the tcx_eff_attn_fwd table (8 slots, rest unused). */
#define TCX_BRIDGE_NP 50
size_t tcx_bridge_layer_workspace_bytes(int B, int S);
int tcx_bridge_layer_fwd(const float* x, const void* const* p, int channel_att, float scale, float ln_eps, float* y,
                         int B, int S, void* ws, void* stream);

/* BridgeBlock_4.forward (MSTr.py:2422-2431): L chained BridgLayer_4 on the token buffer, p = L x TCX_BRIDGE_NP slots,
 * channel_att[l] selects the attention of layer l.  Same arithmetic as L calls of tcx_bridge_layer_fwd; in the fp16
 * pipeline the norm1 of layer l+1 is produced by the Mix-FFN epilogues of layer l. */
size_t tcx_bridge_block_workspace_bytes(int B, int S);
int tcx_bridge_block_fwd(const float* x, const void* const* p, const int* channel_att, int L, float scale, float ln_eps,
                         float* y, int B, int S, void* ws, void* stream);

/* decoder — MyDecoderLayer.forward pieces (MSTr.py:273-290, :184-201, :212-227) */
int tcx_concat_linear_fwd(const float* x1, const float* x2, const float* w, const float* b, float* y, int M, int C1,
                          int C2, int N, int batch, long long x2_batch_stride, void* stream);
size_t tcx_patch_expand_workspace_bytes(int B, int H, int W, int C, int scale);
int tcx_patch_expand_fwd(const float* x, const float* w, const float* lnw, const float* lnb, float eps, float* y,
                         int B, int H, int W, int C, int scale, void* ws, void* stream);
size_t tcx_final_expand_head_workspace_bytes(int B, int H, int W);
int tcx_final_expand_head_fwd(const float* x, const float* w, const float* lnw, const float* lnb, float eps,
                              const float* cls_w, const float* cls_b, int ncls, float* logits_nchw, int B, int H,
                              int W, void* ws, void* stream);

/* ---- networks/Transception.py variant; fp16 pipeline only: every GEMM weight must have been
 * registered with tcx_prepare_weight_f16 / tcx_prepare_conv_weight_f16, otherwise the call fails (no fallback). ---- */

/* FuseEfficientAttention.forward (Transception.py:49-87, head_count = 1) on LayerNorm output xn [B][N][C]:
 * y = residual + reprojection(att) with the reference's raw [N][C] -> [C][N] re-reading of keys / queries / values.
 * p = {keys_w,keys_b, queries_w,queries_b, values_w,values_b, reprojection_w,reprojection_b}; C % 64 == 0, C <= 512. */
size_t tcx_fuse_eff_attn_workspace_bytes(int B, int N, int C);
int tcx_fuse_eff_attn_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int N, int C, void* ws,
                          void* stream);

/* EfficientTransformerBlockFuse.forward (Transception.py:213-250, two-branch case) on x [B][H1*W1 + H2*W2][C] -> y.
 * p = {n1_w,n1_b, k_w,k_b,q_w,q_b,v_w,v_b,reproj_w,reproj_b, n2_w,n2_b, mlp1[8], mlp2[8]} (Mix-FFN slots as in
 * tcx_mixffn_skip_fwd); N = H1*W1 + H2*W2 */
size_t tcx_fuse_block_workspace_bytes(int B, int N, int C);
int tcx_fuse_block_fwd(const float* x, const void* const* p, float ln_eps, float mlp_ln_eps, float* y, int B, int H1, int W1, int H2,
                       int W2, int C, void* ws, void* stream);

/* The two OverlapPatchEmbeddings_fuse branches of a stage (EffSegformer.py:117-131; Transception.py:412-419) on the NHWC
 * map x [B][H][W][Cin]: conv k1 x k1 / k2 x k2 (shared stride and dilation) + LayerNorm each, written to the concatenated
 * token buffer tokens [B][n1+n2][C].  p = {proj1_w,proj1_b,norm1_w,norm1_b, proj2_w,proj2_b,norm2_w,norm2_b}. */
size_t tcx_dual_patch_embed_workspace_bytes(int B, int H, int W, int Cin, int C, int k1, int k2, int stride, int pad1, int pad2,
                                            int dil);
int tcx_dual_patch_embed_fwd(const float* x, const void* const* p, float ln_eps, float* tokens, int B, int H, int W, int Cin, int C,
                             int k1, int k2, int stride, int pad1, int pad2, int dil, void* ws, void* stream);

/* Stage tail of MiT_3inception.forward (Transception.py:462-476, concat='original'): LayerNorm of all tokens, branch-1 map
 * nearest-upsampled (F.interpolate default) to H2 x W2, channel concat, 1x1 conv 2C -> C.  out [B][H2*W2][C] (NHWC map).
 * p = {norm_w,norm_b, conv1_1_w,conv1_1_b} */
size_t tcx_fuse_merge_workspace_bytes(int B, int N, int n2, int C);
int tcx_fuse_merge_fwd(const float* tokens, const void* const* p, float ln_eps, float* out, int B, int H1, int W1, int H2, int W2,
                       int C, void* ws, void* stream);

/* Same stage tail with the selective-kernel fusion SK_Block (Transception.py:477-481, :328-358; concat != 'original'):
 * S = mean(map1 + map2), Z = fc(S), softmax over the two paths of fcs_i(Z), V = sum a_i map_i, 1x1 conv -> ReLU -> BN(eval).
 * p = {norm_w,norm_b, fc_w,fc_b, fcs0_w,fcs0_b, fcs1_w,fcs1_b, conv_w,conv_b, bn_w,bn_b,bn_running_mean,bn_running_var};
 * d = width of fc (max(32, C/16)) */
size_t tcx_fuse_merge_sk_workspace_bytes(int B, int N, int n2, int C);
int tcx_fuse_merge_sk_fwd(const float* tokens, const void* const* p, float ln_eps, float bn_eps, float* out, int B, int H1, int W1,
                          int H2, int W2, int C, int d, void* ws, void* stream);

/* ---- fused segmentation loss of the training step ----
 * loss = w_ce * CrossEntropyLoss(logits, label) + w_dice * DiceLoss(n_classes)(logits, label, weight=class_w, softmax=True)
 * (trainer.py:141-143 with w = 0.4 / 0.6; utils.py:11-47), forward and d loss / d logits, no host synchronisation (the
 * reference reads one `.item()` per class per step, utils.py:45).  logits [B][K][HW] fp32 (NCHW); labels [B][HW] of
 * label_kind 0 int64 / 1 float32 (integer-valued) / 2 int32 / 3 uint8; K <= 16; class_w = HOST array of K floats or NULL
 * (all ones).  softmax = 0 treats `logits` as probabilities (DiceLoss(..., softmax=False)) and needs w_ce == 0.
 * out (device, 4 + K floats) = {loss, ce, dice, number of labels outside [0,K), class_wise_dice[K] (utils.py:45)}.
 * tcx_seg_loss_bwd needs the ws written by tcx_seg_loss_fwd on the same inputs; grad_out = device scalar or NULL (1). */
size_t tcx_seg_loss_workspace_bytes(int B, int K, long long HW);
int tcx_seg_loss_fwd(const float* logits, const void* labels, int label_kind, int B, int K, long long HW, int softmax, float w_ce,
                     float w_dice, const float* class_w, float* out, void* ws, void* stream);
int tcx_seg_loss_bwd(const float* logits, const void* labels, int label_kind, int B, int K, long long HW, int softmax, float w_ce,
                     float w_dice, const float* class_w, const float* grad_out, float* dlogits, const void* ws, void* stream);

/* Slice post-processing of utils.test_single_volume (utils.py:86): arg max over the class planes of logits [B][K][HW]
 * -> uint8 label map [B][HW] (softmax is monotone; first index wins ties). */
int tcx_argmax_classes_fwd(const float* logits, unsigned char* labels, int B, int K, long long HW, void* stream);

/* ---- training row: backward entries.  Gradients are fp32 in HBM; the GEMM-shaped parts run on the
 * tcgen05 kernels with TF32 operands read in place (flag "wgrad_tc" = 0: the round-1 re-layout path); every reduction over tokens is two-pass in a fixed order (bit-reproducible).  These are
 * what the autograd nodes of the drop-in modules (transception_b200/autograd.py) call where the reference relies on ATen's
 * autograd formulas for nn.LayerNorm / nn.Linear / MixFFN_skip (MSTr.py:58-61). ---- */

/* nn.LayerNorm backward: x [M][C] (the forward input), dy [M][C] -> dx [M][C], dw [C], db [C].  dres (nullable) [M][C] is added
 * to dx: the gradient arriving over the residual connection around the norm (x + f(LN(x)), MSTr.py:164-173, :935-946), so the
 * sum autograd would run as a separate pass happens in the same kernel. */
size_t tcx_layernorm_bwd_workspace_bytes(long long M, int C);
int tcx_layernorm_bwd(const float* x, const float* w, const float* dy, const float* dres, float eps, float* dx, float* dw, float* db,
                      long long M, int C, void* ws, void* stream);

/* nn.Linear backward for y = x w^T + b: x [M][K] (fp32, or fp16 when x_f16), w [N][K], dy [M][N] ->
 * dx [M][K] = dy w, dw [N][K] = dy^T x, db [N] = column sums of dy; any of dx / dw / db may be NULL (skipped). */
size_t tcx_linear_bwd_workspace_bytes(long long M, int N, int K);
int tcx_linear_bwd(const void* x, int x_f16, const float* w, const float* dy, float* dx, float* dw, float* db, long long M, int N,
                   int K, void* ws, void* stream);

/* Weight-gradient GEMM with both operands read in place (MN-major tcgen05 tiles; ordered split-K fold inside the kernel:
 * thread-block clusters over distributed shared memory, then cluster sums through HBM):
 *   out[z][i][j] = alpha * sum_t a[z][t][i] * b[z][t][j],  z < batch, t < tokens, i < NL, j < KL
 * a [batch][tokens][lda], b [batch][tokens][ldb]: fmt 2 = both fp32 (TF32 MMA; NL, KL, lda, ldb multiples of 4), fmt 0 = both
 * fp16, fmt 1 = both bf16 (multiples of 8); out [batch][NL][KL] fp32.
 * Optional: db [NL] = alpha * column sums of a (batch 1; the bias gradient of nn.Linear — ATen autograd of MSTr.py:58-61 and
 * every other nn.Linear of the path), outT [batch][KL][NL] = transposed copy, mask_ch > 0 keeps only the diagonal
 * mask_ch x mask_ch blocks (the per-head contexts of FactorAtt_ConvRelPosEnc, MSTr.py:868-872).  Bit-reproducible. */
size_t tcx_wgrad_mn_workspace_bytes(long long tokens, int NL, int KL, int batch, int fmt);
int tcx_wgrad_mn(const void* a, const void* b, int fmt, long long tokens, int NL, int KL, int lda, int ldb, int batch, float alpha,
                 float* out, float* outT, float* db, int mask_ch, void* ws, void* stream);

/* ---- multi-tensor optimizer step (trainer.py:125,148: clip_grad_norm_ [optional] + optim.SGD(momentum, weight_decay)) over a
 * device-resident tensor table.  All table arguments are DEVICE arrays built once by the caller: *_ptrs = one pointer per tensor,
 * numel / offsets = int64 per tensor, blocks = int32 pairs (tensor index, chunk index) — one per thread block, a chunk being
 * tcx_mt_chunk() elements.  lr, coef and out live in device memory (a learning-rate schedule is a 4-byte copy, not a re-capture).
 *   tcx_mt_gather : flat[offsets[t] + i] = src[t][i]                      (the gradient all-reduce bucket, no concatenation pass)
 *   tcx_mt_sqnorm : out[0] = ||all gradients||_2 (ordered fold, bit-reproducible), out[1] = min(1, max_norm / (out[0] + 1e-6))
 *                   (1 when max_norm <= 0); part = nblocks floats of scratch
 *   tcx_mt_sgd    : d = coef * g + weight_decay * p;  buf = momentum * buf + d;  p -= lr * buf;  w16[t] = half(p) where non-NULL
 *                   (the prepared fp16 GEMM copy of tcx_prepare_weight_f16 stays current without a conversion launch) ---- */
int tcx_mt_chunk(void);
int tcx_mt_gather(const void* src_ptrs, const void* numel, const void* offsets, const void* blocks, int nblocks, float* flat, void* stream);
int tcx_mt_sqnorm(const void* grad_ptrs, const void* numel, const void* blocks, int nblocks, float* part, float max_norm, float* out,
                  void* stream);
int tcx_mt_sgd(const void* param_ptrs, const void* grad_ptrs, const void* buf_ptrs, const void* w16_ptrs, const void* numel,
               const void* blocks, int nblocks, const float* lr, const float* coef, float momentum, float weight_decay, void* stream);

/* MixFFN_skip training forward: the arithmetic of tcx_mixffn_skip_fwd (fp16 pipeline only: fc1 / fc2 must be prepared),
 * keeping in `saved` (tcx_mixffn_skip_saved_bytes, opaque) what backward needs: fp16 xn, fc1 output, GELU output and the
 * fp32 LayerNorm input.  xn32 (nullable) = the fp32 xn the forward was given: the in-place TF32 operand of fc1's weight gradient.  tcx_mixffn_skip_bwd: dy [B*N][C] = dL/dy -> dxn [B*N][C] (NULL: skipped) and the eight parameter
 * gradients dp = {d fc1_w, d fc1_b, d dw_w, d dw_b, d ln_w, d ln_b, d fc2_w, d fc2_b} (device pointers, shapes of p). */
size_t tcx_mixffn_skip_saved_bytes(int B, int N, int C, int C4);
int tcx_mixffn_skip_train_fwd(const float* xn, const void* const* p, float ln_eps, const float* residual, float* y, int B, int H,
                              int W, int C, int C4, void* saved, const void* xn16, void* stream);   /* xn16 (nullable): fp16 twin of xn
                              from tcx_layernorm_dual_fwd, read in place (then tcx_mixffn_skip_bwd needs xn32) */
size_t tcx_mixffn_skip_bwd_workspace_bytes(int B, int N, int C, int C4);
int tcx_mixffn_skip_bwd(const float* dy, const void* const* p, float ln_eps, const void* saved, const float* xn32, float* dxn,
                        void* const* dp, int B, int H, int W, int C, int C4, void* ws, void* stream);

/* EfficientAttention (MSTr.py:106-143) training forward (arithmetic of tcx_eff_attn_fwd with reinterpret = 0; fp16 pipeline
 * only) and backward.  `saved` keeps the fp16 LayerNorm output, K | Q | V, the channel-softmax queries, the context and
 * the attention output; the token softmax of the keys is recomputed in fp32 by backward.
 * dy [B*N][C] -> dxn [B*N][C] (NULL: skipped), dp = gradients of {k_w,k_b,q_w,q_b,v_w,v_b,reproj_w,reproj_b}. */
size_t tcx_eff_attn_saved_bytes(int B, int N, int C);
int tcx_eff_attn_train_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int N, int C, void* saved,
                           void* stream);
size_t tcx_eff_attn_bwd_workspace_bytes(int B, int N, int C);
int tcx_eff_attn_bwd(const float* dy, const void* const* p, const void* saved, float* dxn, void* const* dp, int B, int N, int C,
                     void* ws, void* stream);

/* FactorAtt_ConvRelPosEnc (MSTr.py:852-886) training forward on the fp16 pipeline (qkv GEMM -> fused per-head attention + conv
 * relative position encoding -> projection GEMM + residual: the kernels of the inference path; qkv / proj weights must be
 * prepared).  `saved` (tcx_mb_factor_attn_saved_bytes) keeps the fp16 q | k | v rows and the fp16 attention output for backward. */
size_t tcx_mb_factor_attn_saved_bytes(int B, int N, int C);
size_t tcx_mb_factor_attn_train_workspace_bytes(int B, int N, int C);
int tcx_mb_factor_attn_train_fwd(const float* xn, const void* const* p, const float* residual, float* y, int B, int H, int W, int C,
                                 int heads, void* saved, void* ws, const void* xn16, void* stream);

/* FactorAtt_ConvRelPosEnc backward.  fwd_ws = the workspace of tcx_mb_factor_attn_fwd (saved_f16 = 0: fp32 q | k | v rows, context,
 * attention output) or the `saved` buffer of tcx_mb_factor_attn_train_fwd (saved_f16 = 1), with the same xn.  dy [B*N][C] -> dxn
 * (NULL: skipped), dp = gradients of {qkv_w,qkv_b,crpe_w3,crpe_b3,crpe_w5,crpe_b5,crpe_w7,crpe_b7,proj_w,proj_b}.  8 heads. */
size_t tcx_mb_factor_attn_bwd_workspace_bytes(int B, int N, int C);
int tcx_mb_factor_attn_bwd(const float* dy, const float* xn, const void* const* p, const void* fwd_ws, int saved_f16, float* dxn,
                           void* const* dp, int B, int H, int W, int C, int heads, void* ws, void* stream);

/* Bridge attention core softmax(q k^T * scale) v (MSTr.py:2281-2285) for training, flash style in both directions.
 * tcx_flash_attn_train_fwd = tcx_flash_attn_fwd + lse [B][Nq]: the row log2-sum-exp of q k^T * scale * log2(e).
 * tcx_flash_attn_bwd: q [B][Nq][64], kv [B][Nk][128] (k | v), out = the forward output, lse, dout [B][Nq][64] -> dq [B][Nq][64],
 * dkv [B][Nk][128]: one tcgen05 kernel per (image, 128-row kv tile) recomputes the probabilities tile by tile (no score-sized
 * tensor in HBM), dK / dV accumulate in tensor memory, the 7 dQ partials are folded in tile order.  fp16 operands; dout is
 * scaled by one power of two derived from its absolute maximum on the device.  Bit-reproducible. */
int tcx_flash_attn_train_fwd(const float* q, const float* kv, float* out, float* lse, int B, int Nq, int Nk, float scale, void* ws,
                             void* stream);
size_t tcx_flash_attn_bwd_workspace_bytes(int B, int Nq, int Nk);
int tcx_flash_attn_bwd(const float* q, const float* kv, const float* out, const float* lse, const float* dout, float scale, float* dq,
                       float* dkv, int B, int Nq, int Nk, void* ws, void* stream);

/* ConvPosEnc / DWConv (MSTr.py:744-752, :26-31) backward of y = dw3x3(x) + b (+ x when add_input): dx, dw [C][9], db [C]
 * (dx may be NULL; dw NULL skips dw and db) */
size_t tcx_dwconv_tokens_bwd_workspace_bytes(int B, int H, int W, int C);
int tcx_dwconv_tokens_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int B, int H, int W, int C,
                          int add_input, void* ws, void* stream);

/* Backward of the bridge attention core out = softmax(q k^T * scale) v (MSTr.py:2281-2285; forward: tcx_flash_attn_fwd).
 * q, dout [B][Nq][64], kv [B][Nk][128] (k | v) -> dq [B][Nq][64], dkv [B][Nk][128].  The probabilities are recomputed from q
 * and k as a batched tcgen05 GEMM (fp32 scores in ws), every contraction of the backward is a tcgen05 GEMM. */
size_t tcx_attn_core_bwd_workspace_bytes(int B, int Nq, int Nk);
int tcx_attn_core_bwd(const float* q, const float* kv, const float* dout, float scale, float* dq, float* dkv, int B, int Nq, int Nk,
                      void* ws, void* stream);

/* Efficient-attention core on given fp32 token-major K, Q, V [B][N][C]: out = softmax_channels(Q) (softmax_tokens(K)^T V).
 * With K, Q, V the transposed raw [C][N] re-readings of the k / q / v projections this is M_EfficientChannelAtten.forward
 * between its Linear layers (MSTr.py:2312-2353).  Backward recomputes the two softmaxes and the context. */
size_t tcx_ea_core_workspace_bytes(int B, int N, int C);
int tcx_ea_core_fwd(const float* k, const float* q, const float* v, float* out, int B, int N, int C, void* ws, void* stream);
int tcx_ea_core_bwd(const float* k, const float* q, const float* v, const float* dout, float* dk, float* dq, float* dv, int B, int N,
                    int C, void* ws, void* stream);

/* Encoder glue of the training row.
 * BatchNorm2d in train mode (batch statistics over the M = B*H*W rows of NHWC x [M][C]) fused with the activation that follows it
 * in DWConv2d_BN (MSTr.py:355-362, Hardswish = act 2), Conv2d_BN (:399-404, act 0 or 2) and CoordAtt.bn1 (:1331, silu_swish =
 * act 4).  stat (2*C floats: mean | 1/std) is kept for backward; running_mean / running_var (NULL to skip) receive
 * nn.BatchNorm2d's momentum update with the unbiased variance. */
size_t tcx_bn_act_train_workspace_bytes(long long M, int C);
int tcx_bn_act_train_fwd(const float* x, const float* w, const float* b, float* running_mean, float* running_var, float eps,
                         float momentum, int act, float* y, float* stat, long long M, int C, void* ws, void* stream);
int tcx_bn_act_train_bwd(const float* x, const float* dy, const float* stat, const float* w, const float* b, int act, float* dx, float* dw,
                         float* db, long long M, int C, void* ws, void* stream);
/* depthwise 3x3, pad 1, stride 1 | 2, no bias (RIPM / ResBlock, MSTr.py:340, :1031) on NHWC x [B,H,W,C] and its gradients */
int tcx_dwconv3x3_nhwc_fwd(const float* x, const float* w, float* y, int B, int H, int W, int C, int stride, void* stream);
size_t tcx_dwconv3x3_nhwc_bwd_workspace_bytes(int B, int H, int W, int C, int stride);
int tcx_dwconv3x3_nhwc_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H, int W, int C, int stride,
                           void* ws, void* stream);
/* CoordAtt (MSTr.py:1322-1348): y [B][H+W][C] = (mean over W | mean over H) of NHWC x; gate out = x * sigmoid(z_h) * sigmoid(z_w)
 * with z [B][H+W][C] the conv_h | conv_w outputs before the sigmoid */
int tcx_coord_pool_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream);
int tcx_coord_pool_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream);
int tcx_coord_gate_fwd(const float* x, const float* z, float* out, int B, int H, int W, int C, void* stream);
int tcx_coord_gate_bwd(const float* x, const float* z, const float* dout, float* dx, float* dz, int B, int H, int W, int C, void* stream);

/* Training row of the bridge (MSTr.py:2373-2409): the token buffer [B][Ntok][64] <-> its four dense per-scale slabs [B][n_k][64]
 * (n_k = (S/2^k)^2 * {1,2,5,8}) in ONE launch each way — the `tx[:, a:b].reshape(...)` slices and `torch.cat` of :2394-2403 and
 * their gradients (each is the other's adjoint).  tcx_bridge_merge_fwd adds `residual` [B][Ntok][64] when it is not NULL
 * (`tx1 + cat(...)`, :2405). */
int tcx_bridge_split_fwd(const float* tokens, void* const* slabs, int B, int S, void* stream);
int tcx_bridge_merge_fwd(const void* const* slabs, const float* residual, float* tokens, int B, int S, void* stream);
/* Scale_reduce (MSTr.py:2225-2247) WITHOUT its LayerNorm as one forward / one backward call: packed [B][Nred][64]; `saved`
 * (tcx_scale_reduce_saved_bytes) keeps the three im2row matrices for the weight gradients.  p = {sr0_w,sr0_b,sr1_w,sr1_b,sr2_w,sr2_b};
 * backward writes EVERY row of dx [B][Ntok][64] and dp = the six parameter gradients in the order of p. */
size_t tcx_scale_reduce_saved_bytes(int B, int S);
size_t tcx_scale_reduce_train_workspace_bytes(int B, int S);
int tcx_scale_reduce_train_fwd(const float* x, const void* const* p, float* packed, int B, int S, void* saved, void* ws, void* stream);
size_t tcx_scale_reduce_bwd_workspace_bytes(int B, int S);
int tcx_scale_reduce_bwd(const float* dpacked, const void* const* p, const void* saved, float* dx, void* const* dp, int B, int S,
                         void* ws, void* stream);

/* Training row of the class head: pixel shuffle x4 + LayerNorm(64) of FinalPatchExpand_X4 (MSTr.py:212-227) + the 1x1 conv to
 * classes (:288-289), on the expand output e [B*H*W][16*64] -> NCHW logits [B][ncls][4H][4W] (the arithmetic of the second half
 * of tcx_final_expand_head_fwd).  tcx_final_head_bwd: ONE pass over e and dlogits -> de (layout of e) and the gradients of
 * {ln_w, ln_b, cls_w [ncls][64], cls_b}; ncls <= 16.  No [pixels][64] tensor is written in either direction. */
int tcx_final_head_train_fwd(const float* e, const float* lnw, const float* lnb, float eps, const float* cls_w, const float* cls_b,
                             int ncls, float* logits_nchw, int B, int H, int W, void* stream);
size_t tcx_final_head_bwd_workspace_bytes(int B, int H, int W);
int tcx_final_head_bwd(const float* e, const float* dlogits, const float* lnw, const float* lnb, float eps, const float* cls_w, int ncls,
                       float* de, float* dlnw, float* dlnb, float* dcls_w, float* dcls_b, int B, int H, int W, void* ws, void* stream);

/* Training row of the stem conv (OverlapPatchEmbeddings.proj, MSTr.py:299-304): forward = tcx_patch_embed_ln_fwd with
 * lnw = lnb = NULL (the conv output; the LayerNorm is a separate node); its weight gradient is tcx_linear_bwd on the patch matrix
 * patches [B*Ho*Wo][Kp] (Kp >= 147, a multiple of 4; columns (ci, ky, kx), zero beyond 147 and outside the image) built here. */
int tcx_patch_im2row_fwd(const float* x, int B, int Cin, int H, int W, float* patches, int Kp, void* stream);

/* LayerNorm with two outputs from one pass: y32 (fp32, kept for the backward) and y16 (fp16: the GEMM operand of the node that
 * follows — the `xn16` argument of tcx_mixffn_skip_train_fwd / tcx_mb_factor_attn_train_fwd).  C in {64, 128, 256, 320, 512}. */
int tcx_layernorm_dual_fwd(const float* x, const float* w, const float* b, float* y32, void* y16, long long M, int C, float eps,
                           void* stream);

/* out = srcs[0] + srcs[1] + ... (n <= 16 fp32 tensors of `numel` elements, added in index order): the gradient of a parameter that
 * the blocks of an MHCAEncoder share (ConvPosEnc / ConvRelPosEnc, MSTr.py:966-978) from the per-block gradients, one launch. */
int tcx_sum_tensors(const void* const* srcs, int n, long long numel, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
