// Fused last decoder stage (reference MSTr.py:212-227 FinalPatchExpand_X4 + :281 1x1 conv to classes):
//   e = x W_e^T          [B*H*W, 64] x [1024, 64]^T : 16 output pixels x 64 channels per input token
//   logits(pixel) = conv1x1(LayerNorm_64(e(pixel)))  -> NCHW [B, ncls, 4H, 4W]
// in ONE kernel, so the 1024-wide expand output (205 MB fp32 at bs16) never exists.  Persistent CTAs; warp 0 = TMA
// producer (x tile of 128 tokens, then eight 2-pixel chunks of W_e through a 3-stage ring), warp 1 = tcgen05.mma issuer
// (kind::tf32; per chunk one 128 x 128 product for the expand channels and one 128 x 32 product against the class
// weights folded through W_e, so the class dot products are tensor-core columns too), warps 2-9 = epilogue: a thread
// owns one token row, pulls the 64 channels of one output pixel from TMEM for the LayerNorm statistics plus its 16
// class columns, and writes  rstd * (dot_k - mean * sum_k) + bias_k  to the NCHW logits.
#include "common.cuh"
#include "misc.cuh"
#include "tc.cuh"

bool tcx_flag_gemm_tc();

namespace {

constexpr int HT_BM = 128;
constexpr int HT_MAXCLS = 16;                      // class dot products ride on the tensor core as 16 extra columns / pixel
constexpr int HT_A_BYTES = 2 * HT_BM * 128;        // [2 k-blocks][128 rows][128 B]   (K = 64 fp32)
constexpr int HT_W_BYTES = 2 * 128 * 128;          // [2 k-blocks][128 rows][128 B]   W_e rows of 2 output pixels
constexpr int HT_G_BYTES = 2 * 32 * 128;           // [2 k-blocks][ 32 rows][128 B]   folded class rows of the same 2 pixels
constexpr int HT_STAGE = HT_W_BYTES + HT_G_BYTES;  // 40 KB
constexpr int HT_NSTAGE = 3;
constexpr int HT_OFF_W = 2 * HT_A_BYTES;           // 2 A buffers
constexpr int HT_OFF_CLS = HT_OFF_W + HT_NSTAGE * HT_STAGE;
constexpr int HT_OFF_BAR = HT_OFF_CLS + 2 * HT_MAXCLS * 4;
constexpr int HT_SMEM = HT_OFF_BAR + 16 * 8 + 16 + 1024;
constexpr int HT_THREADS = 64 + 8 * 32;
constexpr int HT_ACC = 160;                        // TMEM columns per accumulator set: 2 x 64 expand + 2 x 16 class

struct HeadMaps {
  CUtensorMap x;   // [M][64] fp32, box {32, 128}
  CUtensorMap w;   // [1024][64] fp32, box {32, 128}
  CUtensorMap g;   // [256][64] fp32 folded class rows, box {32, 32}
};

template <int O, int N>
__device__ __forceinline__ void ht_ld32(uint32_t taddr, uint32_t (&s)[N]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(s[O + 0]), "=r"(s[O + 1]), "=r"(s[O + 2]), "=r"(s[O + 3]), "=r"(s[O + 4]), "=r"(s[O + 5]), "=r"(s[O + 6]), "=r"(s[O + 7]), "=r"(s[O + 8]), "=r"(s[O + 9]), "=r"(s[O + 10]), "=r"(s[O + 11]), "=r"(s[O + 12]), "=r"(s[O + 13]), "=r"(s[O + 14]), "=r"(s[O + 15]), "=r"(s[O + 16]), "=r"(s[O + 17]), "=r"(s[O + 18]), "=r"(s[O + 19]), "=r"(s[O + 20]), "=r"(s[O + 21]), "=r"(s[O + 22]), "=r"(s[O + 23]), "=r"(s[O + 24]), "=r"(s[O + 25]), "=r"(s[O + 26]), "=r"(s[O + 27]), "=r"(s[O + 28]), "=r"(s[O + 29]), "=r"(s[O + 30]), "=r"(s[O + 31])
               : "r"(taddr)
               : "memory");
}

// G[p][k][j] = sum_c (cw[k][c] * lnw[c]) * W_e[p*64 + c][j]  (k < ncls, else 0): the class head folded through the
// expand weights, so that  sum_c e_c * lnw_c * cw[k][c]  of pixel p is one more output column of the expand GEMM.
__global__ void __launch_bounds__(64) head_fold_kernel(const float* __restrict__ we, const float* __restrict__ lnw,
                                                       const float* __restrict__ cw, int ncls, float* __restrict__ G) {
  pdl_trigger();
  const int p = blockIdx.x >> 4, k = blockIdx.x & 15, j = threadIdx.x;
  float a = 0.f;
  if (k < ncls)
    for (int c = 0; c < 64; c++) a = fmaf(cw[k * 64 + c] * lnw[c], we[(size_t)(p * 64 + c) * 64 + j], a);
  pdl_wait();          // G may still be read by the previous forward's head kernel
  G[(size_t)(p * 16 + k) * 64 + j] = a;
}

__global__ void __launch_bounds__(HT_THREADS, 1) head_tc_kernel(const __grid_constant__ HeadMaps maps, int M, int H, int W,
                                                                const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                                float eps, const float* __restrict__ cw,
                                                                const float* __restrict__ cb, int ncls, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bl = reinterpret_cast<float*>(smem + HT_OFF_CLS);   // [ncls] class bias + sum(lnb * cw)
  float* sl = bl + HT_MAXCLS;                                 // [ncls] sum_c lnw[c] * cw[k][c]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + HT_OFF_BAR);   // [2]
  uint64_t* a_empty = a_full + 2;                                      // [2]
  uint64_t* w_full = a_empty + 2;                                      // [HT_NSTAGE]
  uint64_t* w_empty = w_full + HT_NSTAGE;                              // [HT_NSTAGE]
  uint64_t* acc_full = w_empty + HT_NSTAGE;                            // [2]
  uint64_t* acc_empty = acc_full + 2;                                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (M + HT_BM - 1) / HT_BM;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&maps.x);
    tc::prefetch_tmap(&maps.w);
    tc::prefetch_tmap(&maps.g);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 8 * 32);
    }
    for (int i = 0; i < HT_NSTAGE; i++) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  if (threadIdx.x < ncls) {     // module parameters only: before pdl_wait
    float s = cb[threadIdx.x], t = 0.f;
    for (int d = 0; d < 64; d++) {
      s = fmaf(cw[threadIdx.x * 64 + d], lnb[d], s);
      t = fmaf(cw[threadIdx.x * 64 + d], lnw[d], t);
    }
    bl[threadIdx.x] = s;
    sl[threadIdx.x] = t;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ti = 0, wi = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
        const uint32_t ab = ti & 1;
        tc::mbar_wait(&a_empty[ab], ((ti >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&a_full[ab], HT_A_BYTES);
        for (int kb = 0; kb < 2; kb++)
          tc::tma_load_2d(smem + ab * HT_A_BYTES + kb * (HT_BM * 128), &maps.x, kb * 32, tile * HT_BM, &a_full[ab]);
        for (int g = 0; g < 8; g++, wi++) {          // group g = output pixels 2g, 2g+1
          const uint32_t ws = wi % HT_NSTAGE;
          tc::mbar_wait(&w_empty[ws], ((wi / HT_NSTAGE) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&w_full[ws], HT_STAGE);
          uint8_t* dst = smem + HT_OFF_W + ws * HT_STAGE;
          for (int kb = 0; kb < 2; kb++) {
            tc::tma_load_2d(dst + kb * (128 * 128), &maps.w, kb * 32, g * 128, &w_full[ws]);
            tc::tma_load_2d(dst + HT_W_BYTES + kb * (32 * 128), &maps.g, kb * 32, g * 32, &w_full[ws]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_e = tc::umma_idesc(2, HT_BM, 128);
      constexpr uint32_t idesc_c = tc::umma_idesc(2, HT_BM, 32);
      uint32_t ti = 0, wi = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
        const uint32_t ab = ti & 1;
        tc::mbar_wait(&a_full[ab], (ti >> 1) & 1);
        for (int g = 0; g < 8; g++, wi++) {
          const uint32_t ws = wi % HT_NSTAGE, acc = wi & 1;
          tc::mbar_wait(&w_full[ws], (wi / HT_NSTAGE) & 1);
          tc::mbar_wait(&acc_empty[acc], ((wi >> 1) & 1) ^ 1);
          tc::fence_after_sync();
          const uint32_t wbase = tc::smem_u32(smem + HT_OFF_W + ws * HT_STAGE);
#pragma unroll
          for (int kb = 0; kb < 2; kb++) {
            const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(smem + ab * HT_A_BYTES + kb * (HT_BM * 128)));
            const uint64_t bd = tc::umma_desc_sw128(wbase + kb * (128 * 128));
            const uint64_t gd = tc::umma_desc_sw128(wbase + HT_W_BYTES + kb * (32 * 128));
#pragma unroll
            for (int k = 0; k < 4; k++) {
              tc::umma_tf32(tmem_base + acc * HT_ACC, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_e, (kb | k) != 0);
              tc::umma_tf32(tmem_base + acc * HT_ACC + 128, ad + (uint64_t)(k * 2), gd + (uint64_t)(k * 2), idesc_c, (kb | k) != 0);
            }
          }
          tc::umma_commit(&w_empty[ws]);
          tc::umma_commit(&acc_full[acc]);
        }
        tc::umma_commit(&a_empty[ab]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3, half = ew >> 2;      // half = which of the group's two pixels this thread handles
    const int HW = H * W, Ho = 4 * H, Wo = 4 * W;
    uint32_t wi = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m = tile * HT_BM + quarter * 32 + lane;      // token row of this thread
      const bool live = m < M;
      const int b = live ? m / HW : 0, rem = live ? m % HW : 0;
      const int h = rem / W, w = rem % W;
      for (int g = 0; g < 8; g++, wi++) {
        const uint32_t acc = wi & 1;
        const int pix = 2 * g + half, p1 = pix >> 2, p2 = pix & 3;   // output pixel (4h + p1, 4w + p2)
        tc::mbar_wait(&acc_full[acc], (wi >> 1) & 1);
        tc::fence_after_sync();
        const uint32_t tacc = tmem_base + acc * HT_ACC + ((uint32_t)(quarter * 32) << 16);
        uint32_t v[64], d[16];
        ht_ld32<0>(tacc + half * 64, v);
        ht_ld32<32>(tacc + half * 64 + 32, v);
        tc::tmem_ld16(tacc + 128 + half * 16, d);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        tc::mbar_arrive(&acc_empty[acc]);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          s0 += __uint_as_float(v[i]); s1 += __uint_as_float(v[i + 1]);
          s2 += __uint_as_float(v[i + 2]); s3 += __uint_as_float(v[i + 3]);
        }
        const float mean = ((s0 + s1) + (s2 + s3)) * (1.f / 64.f);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float d0 = __uint_as_float(v[i]) - mean, d1 = __uint_as_float(v[i + 1]) - mean;
          const float d2 = __uint_as_float(v[i + 2]) - mean, d3 = __uint_as_float(v[i + 3]) - mean;
          q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
        }
        const float rstd = rsqrtf(((q0 + q1) + (q2 + q3)) * (1.f / 64.f) + eps);
        if (live) {
          float* o = out + ((long long)b * ncls * Ho + (4 * h + p1)) * Wo + 4 * w + p2;
          // sum_c (e_c - mean) * rstd * lnw_c * cw[k][c] + bl = rstd * (dot_k - mean * sl[k]) + bl[k]
#pragma unroll
          for (int k = 0; k < HT_MAXCLS; k++)
            if (k < ncls) o[(long long)k * Ho * Wo] = fmaf(rstd, __uint_as_float(d[k]) - mean * sl[k], bl[k]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool head_tc_eligible(const float* x, const float* w, int ncls) {
  return tcx_flag_gemm_tc() && ncls >= 1 && ncls <= HT_MAXCLS && tcx_get_encode_tiled() != nullptr &&
         ((((uintptr_t)x) | ((uintptr_t)w)) & 15) == 0;
}

size_t head_tc_workspace_floats() { return 256 * 64; }

int launch_head_tc(const float* x, const float* w, int B, int H, int W, const float* lnw, const float* lnb, float eps,
                   const float* cw, const float* cb, int ncls, float* out, float* ws, cudaStream_t st) {
  const int M = B * H * W;
  TCX_REQUIRE(ws != nullptr && (((uintptr_t)ws) & 15) == 0, "head_tc: workspace missing or misaligned");
  tcx_launch_pdl(head_fold_kernel, dim3(256), dim3(64), 0, st, w, lnw, cw, ncls, ws);
  TCX_TRY(tcx_check_launch("head_fold"));
  HeadMaps maps;
  {
    tcx_encode_tiled_fn enc = tcx_get_encode_tiled();
    TCX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    auto mk = [&](CUtensorMap* map, const float* base, int rows, int box_rows) -> int {
      cuuint64_t dims[2] = {64, (cuuint64_t)rows};
      cuuint64_t strides[1] = {64 * sizeof(float)};
      cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      TCX_REQUIRE(r == CUDA_SUCCESS, "head_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
      return 0;
    };
    TCX_TRY(mk(&maps.x, x, M, HT_BM));
    TCX_TRY(mk(&maps.w, w, 1024, 128));
    TCX_TRY(mk(&maps.g, ws, 256, 32));
  }
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM);
    TCX_REQUIRE(e == cudaSuccess, "head_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int ntiles = cdiv(M, HT_BM);
  // algorithmic bytes: x read, W_e read once, logits written
  ProfScope prof("head_tc", st, (double)M * 64 * 4 + 1024.0 * 64 * 4 + (double)M * 16 * ncls * 4);
  cudaError_t le = tcx_launch_pdl(head_tc_kernel, dim3(ntiles < sms ? ntiles : sms), dim3(HT_THREADS), (size_t)HT_SMEM, st, maps, M, H,
                                  W, lnw, lnb, eps, cw, cb, ncls, out);
  TCX_REQUIRE(le == cudaSuccess, "head_tc: launch failed: %s", cudaGetErrorString(le));
  return tcx_check_launch("head_tc");
}
