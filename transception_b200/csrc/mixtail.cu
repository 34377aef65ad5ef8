// Fused Mix-FFN tail (reference MSTr.py:59-61):  y = residual + fc2( GELU( LayerNorm( dw3x3(h) + b + h ) ) )
// for hidden widths C4 in {256, 512} (fc2: C4 -> C = C4 / 4), h = fc1 output in fp16.
//
// One persistent kernel replaces the dw+LN+GELU kernel, the [M, C4] fp16 round trip of its output and the fc2 GEMM:
//  * 8 producer warps compute a = GELU(LN(dw3x3(h)+b+h)) for a tile of 128 tokens on the CUDA cores (warp per token,
//    row in registers, predicated taps, one-deep software pipeline — the dwln_kernel inner loop) and write it as fp16
//    straight into the 128-byte-swizzled shared-memory A tile of a tcgen05 MMA;
//  * one thread issues  acc[128 x C] += A[128 x C4] * W2[C x C4]^T  (kind::f16, fp32 accumulate in TMEM); W2 is
//    resident in shared memory (C4 = 256) or streamed k-block by k-block through a TMA ring (C4 = 512);
//  * 4 epilogue warps add bias + fp32 residual and store y (rows may live in per-image slabs of a larger buffer: the
//    bridge token buffer).
// A tiles are double-buffered at C4 = 256, so the producers of tile i+1 overlap the MMA / epilogue of tile i.
#include <cuda_fp16.h>
#include "common.cuh"
#include "fused16.cuh"
#include "tc.cuh"

namespace {

constexpr int MT_BM = 128;
constexpr int MT_PROD_WARPS = 8;
constexpr int MT_THREADS = (MT_PROD_WARPS + 2 + 4) * 32;     // producers | W2 loader, MMA issuer | epilogue

struct MixTailGroup {
  const __half* h;      // [B*N][C4] fc1 output
  const float* dww;     // [C4][9]
  const float* dwb;     // [C4]
  const float* lnw;     // [C4]
  const float* lnb;     // [C4]
  const float* b2;      // [C]
  const float* res;     // residual rows (fp32) or null
  float* y;             // output rows (fp32)
};
struct MixTailArgs {
  MixTailGroup g[TCX_MAX_GROUPS];
  CUtensorMap w2[TCX_MAX_GROUPS];    // fp16 [C rows][C4 cols], box {64, C}
  int groups, B, H, W;
  long long res_bs, y_bs;            // per-image pitch of residual / y rows (0 = dense [B*N][C])
  float eps;
};

__device__ __forceinline__ void mt_unpack8(const uint4& r, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t mt_pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float mt_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// channel c -> position inside the per-tap smem vectors (lanes read contiguous 16-byte pieces), see fused16.cu::wperm
__device__ __forceinline__ int mt_wperm(int c) {
  const int blk = c >> 8, r = c & 255;
  const int sl = r >> 3, q = (r >> 2) & 1, j = r & 3;
  return blk * 256 + q * 128 + sl * 4 + j;
}
template <int C4>
struct MtCfg {
  static constexpr int C = C4 / 4;
  static constexpr int NV = C4 / 256;                 // 16-byte channel vectors per lane
  static constexpr int NKB = C4 / 64;                 // k-blocks of the fc2 product
  static constexpr int ABUF = C4 == 256 ? 2 : 1;      // A tile buffers
  static constexpr int A_BYTES = NKB * MT_BM * 128;   // 64 KB / 128 KB
  static constexpr bool W_RES = C4 == 256;            // W2 resident in smem
  static constexpr int WKB_BYTES = C * 128;           // one k-block of W2: C rows x 128 B
  static constexpr int WST = W_RES ? NKB : 4;         // W2 ring stages (resident: one per k-block)
  static constexpr int OFF_W = ABUF * A_BYTES;
  static constexpr int OFF_F = OFF_W + WST * WKB_BYTES;          // filters/affine: 12 x C4 floats
  static constexpr int OFF_BAR = OFF_F + 12 * C4 * 4;
  static constexpr int SMEM = OFF_BAR + 32 * 8 + 16 + 1024;
  static constexpr uint32_t TMEM_COLS = 2 * C < 32 ? 32 : 2 * C;  // two accumulators
};

template <int C4>
__global__ void __launch_bounds__(MT_THREADS, 1) mixtail_kernel(const __grid_constant__ MixTailArgs a) {
  using K = MtCfg<C4>;
  constexpr int C = K::C, NV = K::NV, NKB = K::NKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* wsm = reinterpret_cast<float*>(smem + K::OFF_F);   // [9][C4] taps (centre + 1), bias, lnw, lnb (permuted)
  float* bsm = wsm + 9 * C4;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);   // [2]
  uint64_t* a_empty = a_full + 2;                                      // [2]
  uint64_t* w_full = a_empty + 2;                                      // [8]
  uint64_t* w_empty = w_full + 8;                                      // [8]
  uint64_t* acc_full = w_empty + 8;                                    // [2]
  uint64_t* acc_empty = acc_full + 2;                                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int H = a.H, W = a.W, N = H * W;
  const int M = a.B * N;                                  // tokens per group
  const int tiles_per_group = (M + MT_BM - 1) / MT_BM;
  const int ntiles = tiles_per_group * a.groups;
  // groups may carry different filters: the per-CTA filter stage is (re)loaded when the group changes, which with the
  // static tile order below happens at most groups-1 times per CTA.

  if (warp == MT_PROD_WARPS && lane == 0) {
    for (int g = 0; g < a.groups; g++) tc::prefetch_tmap(&a.w2[g]);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&a_full[i], MT_PROD_WARPS * 32);
      tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4 * 32);
    }
    for (int i = 0; i < 8; i++) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    tc::fence_barrier_init();
  }
  if (warp == MT_PROD_WARPS + 1) {
    tc::tmem_alloc(tmem_slot, K::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // tiles are visited group-major per CTA so that the filter stage changes rarely
  auto tile_of = [&](int i, int& g, int& t) {          // i-th tile of this CTA
    const int lin = blockIdx.x + i * gridDim.x;
    g = lin / tiles_per_group;
    t = lin - g * tiles_per_group;
    return lin < ntiles;
  };

  if (warp < MT_PROD_WARPS) {
    // ===================== producers: dw3x3 + skip + LN + GELU -> fp16 A tile =====================
    int cur_g = -1;
    uint32_t ti = 0;
    for (int i = 0;; i++, ti++) {
      int g, t;
      if (!tile_of(i, g, t)) break;
      const MixTailGroup& G = a.g[g];
      if (g != cur_g) {
        // (module parameters: the first staging happens before pdl_wait and overlaps the previous kernel)
        asm volatile("bar.sync 1, %0;" ::"n"(MT_PROD_WARPS * 32) : "memory");      // everyone is done with the old stage
        constexpr int NB = 8;
        for (int i0 = tid; i0 < 9 * C4; i0 += MT_PROD_WARPS * 32 * NB) {
          float v[NB];
#pragma unroll
          for (int j = 0; j < NB; j++) { const int ii = i0 + j * MT_PROD_WARPS * 32; v[j] = ii < 9 * C4 ? __ldg(G.dww + ii) : 0.f; }
#pragma unroll
          for (int j = 0; j < NB; j++) {
            const int ii = i0 + j * MT_PROD_WARPS * 32;
            if (ii < 9 * C4) { const int c = ii / 9, tp = ii - c * 9; wsm[tp * C4 + mt_wperm(c)] = v[j] + (tp == 4 ? 1.f : 0.f); }
          }
        }
        for (int c = tid; c < C4; c += MT_PROD_WARPS * 32) {
          const int pc = mt_wperm(c);
          bsm[pc] = G.dwb ? __ldg(G.dwb + c) : 0.f;
          bsm[C4 + pc] = __ldg(G.lnw + c);
          bsm[2 * C4 + pc] = __ldg(G.lnb + c);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(MT_PROD_WARPS * 32) : "memory");
        if (cur_g < 0) pdl_wait();
        cur_g = g;
      }
      const uint32_t buf = K::ABUF == 2 ? (ti & 1) : 0;
      const uint32_t use = K::ABUF == 2 ? (ti >> 1) : ti;      // how many times this buffer has been used before
      tc::mbar_wait(&a_empty[buf], (use & 1) ^ 1);              // the MMAs that read the previous content retired
      uint8_t* abase = smem + buf * K::A_BYTES;
      const __half* __restrict__ xin = G.h + lane * 8;
      const float* wl = wsm + lane * 4;
      const float* bl = bsm + lane * 4;
      const ptrdiff_t rowpitch = (ptrdiff_t)W * C4;
      constexpr float invC = 1.f / (float)C4;
      // this warp's 16 rows of the tile
      const int r0 = warp * 16;
      int tok = t * MT_BM + r0;
      int wq = tok % W, hq = (tok / W) % H;
      auto load9 = [&](int tk, int wq_, int hq_, int iv, uint4 (&raw)[9]) {
        const bool live = tk < M;
        const __half* pc = xin + (size_t)(live ? tk : 0) * C4 + iv * 256;
        const __half* prow[3] = {pc - rowpitch, pc, pc + rowpitch};
        const bool rv[3] = {live && hq_ > 0, live, live && hq_ + 1 < H};
        const bool cv[3] = {wq_ > 0, true, wq_ + 1 < W};
#pragma unroll
        for (int tp = 0; tp < 9; tp++) {
          const int ky = tp / 3, kx = tp % 3;
          raw[tp] = make_uint4(0u, 0u, 0u, 0u);
          if (rv[ky] && cv[kx]) raw[tp] = *reinterpret_cast<const uint4*>(prow[ky] + (kx - 1) * C4);
        }
      };
      uint4 cur[9], nxt[9];
      load9(tok, wq, hq, 0, cur);
      for (int it = 0; it < 16; it++, tok++) {
        int wn = wq + 1, hn = hq;
        if (wn >= W) { wn = 0; if (++hn >= H) hn = 0; }
        float acc[NV][8];
#pragma unroll
        for (int iv = 0; iv < NV; iv++) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const float4 b4 = *reinterpret_cast<const float4*>(bl + iv * 256 + q * 128);
            acc[iv][q * 4 + 0] = b4.x; acc[iv][q * 4 + 1] = b4.y; acc[iv][q * 4 + 2] = b4.z; acc[iv][q * 4 + 3] = b4.w;
          }
        }
#pragma unroll
        for (int iv = 0; iv < NV; iv++) {
          if (iv + 1 < NV) load9(tok, wq, hq, iv + 1, nxt);
          else if (it + 1 < 16) load9(tok + 1, wn, hn, 0, nxt);
#pragma unroll
          for (int tp = 0; tp < 9; tp++) {
            float f[8];
            mt_unpack8(cur[tp], f);
#pragma unroll
            for (int q = 0; q < 2; q++) {
              const float4 w4 = *reinterpret_cast<const float4*>(wl + tp * C4 + iv * 256 + q * 128);
              acc[iv][q * 4 + 0] = fmaf(f[q * 4 + 0], w4.x, acc[iv][q * 4 + 0]);
              acc[iv][q * 4 + 1] = fmaf(f[q * 4 + 1], w4.y, acc[iv][q * 4 + 1]);
              acc[iv][q * 4 + 2] = fmaf(f[q * 4 + 2], w4.z, acc[iv][q * 4 + 2]);
              acc[iv][q * 4 + 3] = fmaf(f[q * 4 + 3], w4.w, acc[iv][q * 4 + 3]);
            }
          }
#pragma unroll
          for (int tp = 0; tp < 9; tp++) cur[tp] = nxt[tp];
        }
        float s = 0.f;
#pragma unroll
        for (int iv = 0; iv < NV; iv++)
#pragma unroll
          for (int j = 0; j < 8; j++) s += acc[iv][j];
        const float mean = mt_warp_sum(s) * invC;
        float q2 = 0.f;
#pragma unroll
        for (int iv = 0; iv < NV; iv++)
#pragma unroll
          for (int j = 0; j < 8; j++) { const float d = acc[iv][j] - mean; q2 = fmaf(d, d, q2); }
        const float rstd = rsqrtf(mt_warp_sum(q2) * invC + a.eps);
        // row r of the A tile, k-block kb = channel / 64, 16-byte chunk (channel % 64) / 8, 128-byte swizzle
        const int r = r0 + it;
#pragma unroll
        for (int iv = 0; iv < NV; iv++) {
          float o[8];
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const float4 w4 = *reinterpret_cast<const float4*>(bl + C4 + iv * 256 + q * 128);
            const float4 b4 = *reinterpret_cast<const float4*>(bl + 2 * C4 + iv * 256 + q * 128);
            o[q * 4 + 0] = tcx_gelu_fast(fmaf((acc[iv][q * 4 + 0] - mean) * rstd, w4.x, b4.x));
            o[q * 4 + 1] = tcx_gelu_fast(fmaf((acc[iv][q * 4 + 1] - mean) * rstd, w4.y, b4.y));
            o[q * 4 + 2] = tcx_gelu_fast(fmaf((acc[iv][q * 4 + 2] - mean) * rstd, w4.z, b4.z));
            o[q * 4 + 3] = tcx_gelu_fast(fmaf((acc[iv][q * 4 + 3] - mean) * rstd, w4.w, b4.w));
          }
          const int c0 = iv * 256 + lane * 8;                 // first channel of this lane's vector
          const int kb = c0 >> 6, chunk = (c0 & 63) >> 3;
          uint8_t* dst = abase + kb * (MT_BM * 128) + r * 128 + ((chunk ^ (r & 7)) << 4);
          // rows past the end of the group hold zeros (tok >= M never loads): harmless, never stored by the epilogue
          *reinterpret_cast<uint4*>(dst) = make_uint4(mt_pack2(o[0], o[1]), mt_pack2(o[2], o[3]), mt_pack2(o[4], o[5]), mt_pack2(o[6], o[7]));
        }
        wq = wn; hq = hn;
      }
      tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc::mbar_arrive(&a_full[buf]);
    }
  } else if (warp == MT_PROD_WARPS) {
    // ===================== W2 loader (TMA) =====================
    if (lane == 0) {
      pdl_wait();
      if (K::W_RES) {
        // one group only when resident (the launcher guarantees it): every k-block once
        for (int kb = 0; kb < NKB; kb++) {
          tc::mbar_arrive_expect_tx(&w_full[kb], K::WKB_BYTES);
          tc::tma_load_2d(smem + K::OFF_W + kb * K::WKB_BYTES, &a.w2[0], kb * 64, 0, &w_full[kb]);
        }
      } else {
        uint32_t wi = 0;
        for (int i = 0;; i++) {
          int g, t;
          if (!tile_of(i, g, t)) break;
          for (int kb = 0; kb < NKB; kb++, wi++) {
            const uint32_t s = wi % K::WST;
            tc::mbar_wait(&w_empty[s], ((wi / K::WST) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&w_full[s], K::WKB_BYTES);
            tc::tma_load_2d(smem + K::OFF_W + s * K::WKB_BYTES, &a.w2[g], kb * 64, 0, &w_full[s]);
          }
        }
      }
    }
  } else if (warp == MT_PROD_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc(0, MT_BM, C);
      uint32_t ti = 0, wi = 0;
      for (int i = 0;; i++, ti++) {
        int g, t;
        if (!tile_of(i, g, t)) break;
        const uint32_t buf = K::ABUF == 2 ? (ti & 1) : 0;
        const uint32_t use = K::ABUF == 2 ? (ti >> 1) : ti;
        const uint32_t acc = ti & 1;
        tc::mbar_wait(&a_full[buf], use & 1);
        tc::mbar_wait(&acc_empty[acc], ((ti >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        for (int kb = 0; kb < NKB; kb++, wi++) {
          const uint32_t s = K::W_RES ? (uint32_t)kb : wi % K::WST;
          if (K::W_RES) { if (ti == 0) tc::mbar_wait(&w_full[s], 0); }
          else tc::mbar_wait(&w_full[s], (wi / K::WST) & 1);
          tc::fence_after_sync();
          const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(smem + buf * K::A_BYTES + kb * (MT_BM * 128)));
          const uint64_t bd = tc::umma_desc_sw128(tc::smem_u32(smem + K::OFF_W + s * K::WKB_BYTES));
#pragma unroll
          for (int k = 0; k < 4; k++)
            tc::umma_f16(tmem_base + acc * C, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          if (!K::W_RES) tc::umma_commit(&w_empty[s]);
        }
        tc::umma_commit(&a_empty[buf]);
        tc::umma_commit(&acc_full[acc]);
      }
    }
  } else {
    // ===================== epilogue: + bias + residual -> y =====================
    pdl_wait();
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    uint32_t ti = 0;
    for (int i = 0;; i++, ti++) {
      int g, t;
      if (!tile_of(i, g, t)) break;
      const MixTailGroup& G = a.g[g];
      const uint32_t acc = ti & 1;
      const int m = t * MT_BM + quarter * 32 + lane;
      const bool live = m < M;
      long long roff = (long long)m * C, yoff = (long long)m * C;
      if (a.res_bs || a.y_bs) {
        const int b = m / N, n = m - b * N;
        if (a.res_bs) roff = (long long)b * a.res_bs + (long long)n * C;
        if (a.y_bs) yoff = (long long)b * a.y_bs + (long long)n * C;
      }
      tc::mbar_wait(&acc_full[acc], (ti >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t tacc = tmem_base + acc * C + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int s = 0; s < C / 32; s++) {
        uint32_t v[32];
        tc::tmem_ld32(tacc + s * 32, v);
        tc::tmem_ld_wait();
        if (s == C / 32 - 1) {
          tc::fence_before_sync();
          tc::mbar_arrive(&acc_empty[acc]);
        }
        if (live) {
          const float4* bp = reinterpret_cast<const float4*>(G.b2 + s * 32);
          const float4* rp = G.res ? reinterpret_cast<const float4*>(G.res + roff + s * 32) : nullptr;
          float4* yp = reinterpret_cast<float4*>(G.y + yoff + s * 32);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float4 bb = __ldg(bp + c);
            float4 o = make_float4(__uint_as_float(v[c * 4]) + bb.x, __uint_as_float(v[c * 4 + 1]) + bb.y,
                                   __uint_as_float(v[c * 4 + 2]) + bb.z, __uint_as_float(v[c * 4 + 3]) + bb.w);
            if (rp) { const float4 rr = rp[c]; o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
            yp[c] = o;
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == MT_PROD_WARPS + 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, K::TMEM_COLS);
  }
}

template <int C4>
int mixtail_launch(const MixTailArgs& a, cudaStream_t st) {
  using K = MtCfg<C4>;
  static_assert(K::SMEM <= 227 * 1024, "mixtail smem budget");
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(mixtail_kernel<C4>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    TCX_REQUIRE(e == cudaSuccess, "mixtail: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int M = a.B * a.H * a.W;
  const int ntiles = cdiv(M, MT_BM) * a.groups;
  const double bytes = (double)a.groups * M * (C4 * 2.0 + K::C * 4.0 * (a.g[0].res ? 2.0 : 1.0)) + (double)a.groups * K::C * C4 * 2.0;
  ProfScope prof("mixtail", st, bytes);
  cudaError_t le = tcx_launch_pdl(mixtail_kernel<C4>, dim3(ntiles < sms ? ntiles : sms), dim3(MT_THREADS), (size_t)K::SMEM, st, a);
  TCX_REQUIRE(le == cudaSuccess, "mixtail: launch failed: %s", cudaGetErrorString(le));
  return tcx_check_launch("mixtail");
}

}  // namespace

bool mixtail_eligible(int groups, int C4, long long tokens) {
  if (tcx_get_encode_tiled() == nullptr) return false;
  if (C4 == 256) return groups == 1;       // W2 resident: one group per launch
  return C4 == 512 && tokens > 0;
}

int launch_mixtail(const MixTailDesc* d, int groups, int B, int H, int W, int C4, float eps, long long res_bs, long long y_bs,
                   cudaStream_t st) {
  TCX_REQUIRE(groups >= 1 && groups <= TCX_MAX_GROUPS && (C4 == 256 || C4 == 512), "mixtail: unsupported configuration");
  TCX_REQUIRE((long long)B * H * W * C4 < (1ll << 31), "mixtail: tensor too large for 32-bit indexing");
  MixTailArgs a{};
  a.groups = groups; a.B = B; a.H = H; a.W = W; a.eps = eps; a.res_bs = res_bs; a.y_bs = y_bs;
  const int C = C4 / 4;
  tcx_encode_tiled_fn enc = tcx_get_encode_tiled();
  TCX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  for (int i = 0; i < groups; i++) {
    a.g[i] = MixTailGroup{reinterpret_cast<const __half*>(d[i].h), d[i].dww, d[i].dwb, d[i].lnw, d[i].lnb, d[i].b2, d[i].res, d[i].y};
    cuuint64_t dims[2] = {(cuuint64_t)C4, (cuuint64_t)C};
    cuuint64_t strides[1] = {(cuuint64_t)C4 * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)C};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&a.w2[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d[i].w2), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TCX_REQUIRE(r == CUDA_SUCCESS, "mixtail: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  for (int i = groups; i < TCX_MAX_GROUPS; i++) { a.g[i] = a.g[0]; a.w2[i] = a.w2[0]; }
  return C4 == 256 ? mixtail_launch<256>(a, st) : mixtail_launch<512>(a, st);
}
