// Bridge spatial-reduction attention, tcgen05 flash kernel (reference MSTr.py:2281-2285: softmax(q k^T * scale) v,
// one head, d = 64, Nq = 6076 queries against Nk = 784 reduced tokens per image at 224x224).
//
//  * persistent CTAs (one per SM), 10 warps: two softmax warpgroups (thread = one query row = one TMEM lane),
//    one TMA producer warp, one MMA issuer warp (single elected thread, tcgen05.mma kind::f16, fp32 accumulate
//    in TMEM).  A CTA owns a PAIR of 128-row query tiles of the same image so the K/V tiles it streams feed both
//    warpgroups; the tensor core computes S = Q K^T of one warpgroup while the other warpgroup is in its exp phase.
//  * K and V^T tiles (fp16) arrive by TMA (128-byte swizzle) through a 3-stage mbarrier ring; Q rows are read as
//    fp32, pre-multiplied by scale*log2(e), converted to fp16 and written to swizzled shared memory by their owner
//    threads; P = 2^(S - m) goes to swizzled shared memory as the A operand of the P V MMA.
//  * exact online softmax: running max / sum in fp32 registers; every P V product is written to a fresh TMEM
//    buffer and folded into the register accumulator O = O*alpha + PV, so no TMEM read-modify-write is needed.
//  Scores (N_q x N_k per image) never touch HBM: algorithmic traffic per image is q + out (fp32) + k,v (fp16).
#include <cuda_fp16.h>
#include "common.cuh"
#include "attention.cuh"
#include "tc.cuh"

namespace {

constexpr int FT_BM = 128;                             // query rows per warpgroup tile
constexpr int FT_BN = 112;                             // kv rows per tile (784 = 7 x 112; UMMA N = 112)
constexpr int FT_D = 64;
constexpr int FT_STAGES = 3;
constexpr int FT_Q_BYTES = FT_BM * FT_D * 2;           // 16 KB   [128][64] fp16, SW128 K-major
constexpr int FT_P_BYTES = FT_BM * 128 * 2;            // 32 KB   two [128][64] SW128 sub-tiles (kv 0-63 | 64-111)
constexpr int FT_K_BYTES = FT_BN * FT_D * 2;           // 14 KB   [112 kv][64 d]
constexpr int FT_V_BYTES = 2 * FT_D * 64 * 2;          // 16 KB   two [64 d][64 kv] sub-tiles of V^T
constexpr int FT_STAGE_BYTES = FT_K_BYTES + FT_V_BYTES;
constexpr int FT_OFF_P = 2 * FT_Q_BYTES;
constexpr int FT_OFF_KV = FT_OFF_P + 2 * FT_P_BYTES;
constexpr int FT_OFF_BAR = FT_OFF_KV + FT_STAGES * FT_STAGE_BYTES;
constexpr int FT_NBAR = 2 * FT_STAGES + 8;
constexpr int FT_SMEM = FT_OFF_BAR + FT_NBAR * 8 + 16 + 1024;
constexpr int FT_THREADS = 320;
constexpr uint32_t FT_TMEM_COLS = 512;                 // S[2] at 0,128 ; PV[2] at 256,320

struct FlashMaps {
  CUtensorMap k;    // k16  [B][Nk][64]      box {64, 112, 1}
  CUtensorMap vt;   // vt16 [B][64][Nkp]     box {64, 64, 1}
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tc::smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// pass 1 over NC columns starting at c0: running max (columns >= ncols are padding)
template <int NC, bool MASK>
__device__ __forceinline__ float chunk_max(uint32_t taddr, int c0, int ncols, float mx) {
  uint32_t v[NC];
  if constexpr (NC == 32) tc::tmem_ld32(taddr + c0, v); else tc::tmem_ld16(taddr + c0, v);
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < NC; j++) {
    const float s = __uint_as_float(v[j]);
    if (!MASK || c0 + j < ncols) mx = fmaxf(mx, s);
  }
  return mx;
}

// pass 2: p = 2^(s - m) -> fp16 -> swizzled smem (row base rowP, swizzle key sw); returns the partial row sum
template <int NC, bool MASK>
__device__ __forceinline__ float chunk_exp(uint32_t taddr, int c0, int ncols, float m, uint8_t* rowP, int sw) {
  uint32_t v[NC];
  if constexpr (NC == 32) tc::tmem_ld32(taddr + c0, v); else tc::tmem_ld16(taddr + c0, v);
  tc::tmem_ld_wait();
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < NC / 8; g++) {
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int col = c0 + g * 8 + j;
      float e = ex2f(__uint_as_float(v[g * 8 + j]) - m);
      if (MASK && col >= ncols) e = 0.f;
      p[j] = e;
      sum += e;
    }
    const int cc = (c0 >> 3) + g;                 // 16-byte chunk index along kv (0..13)
    uint8_t* dst = rowP + (cc >> 3) * (FT_BM * 128) + (((cc & 7) ^ sw) << 4);
    st_shared_v4(dst, pack_h2(p[0], p[1]), pack_h2(p[2], p[3]), pack_h2(p[4], p[5]), pack_h2(p[6], p[7]));
  }
  return sum;
}

__global__ void __launch_bounds__(FT_THREADS, 1) flash_tc_kernel(const __grid_constant__ FlashMaps maps,
                                                                 const float* __restrict__ q, float* __restrict__ out,
                                                                 int B, int Nq, int Nk, float qscale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sP = smem + FT_OFF_P;
  uint8_t* sKV = smem + FT_OFF_KV;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + FT_OFF_BAR);
  uint64_t* kv_empty = kv_full + FT_STAGES;
  uint64_t* q_full = kv_empty + FT_STAGES;   // [2]
  uint64_t* s_full = q_full + 2;             // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_full = p_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = (Nq + FT_BM - 1) / FT_BM;
  const int pairs_per_img = (tiles_per_img + 1) / 2;
  const int npairs = B * pairs_per_img;
  const int nkt = (Nk + FT_BN - 1) / FT_BN;

  if (warp == 8 && lane == 0) {
    tc::prefetch_tmap(&maps.k);
    tc::prefetch_tmap(&maps.vt);
    for (int s = 0; s < FT_STAGES; s++) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    for (int w = 0; w < 2; w++) {
      tc::mbar_init(&q_full[w], 128);
      tc::mbar_init(&s_full[w], 1);
      tc::mbar_init(&p_full[w], 128);
      tc::mbar_init(&o_full[w], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 9) {
    tc::tmem_alloc(tmem_slot, FT_TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t kvi = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int b = pair / pairs_per_img;
        for (int t = 0; t < nkt; t++, kvi++) {
          const int s = kvi % FT_STAGES;
          tc::mbar_wait(&kv_empty[s], ((kvi / FT_STAGES) & 1) ^ 1);
          uint8_t* dst = sKV + s * FT_STAGE_BYTES;
          tc::mbar_arrive_expect_tx(&kv_full[s], FT_STAGE_BYTES);
          tc::tma_load_3d(dst, &maps.k, 0, t * FT_BN, b, &kv_full[s]);
          tc::tma_load_3d(dst + FT_K_BYTES, &maps.vt, t * FT_BN, 0, b, &kv_full[s]);
          tc::tma_load_3d(dst + FT_K_BYTES + FT_D * 128, &maps.vt, t * FT_BN + 64, 0, b, &kv_full[s]);
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc_s = tc::umma_idesc(0, FT_BM, FT_BN);   // fp16 x fp16 -> fp32, 128 x 112
      constexpr uint32_t idesc_o = tc::umma_idesc(0, FT_BM, FT_D);    // 128 x 64
      const uint32_t q_addr = tc::smem_u32(sQ), p_addr = tc::smem_u32(sP), kv_addr = tc::smem_u32(sKV);
      uint32_t kvi = 0, it = 0, cnt = 0;   // cnt: kv tiles issued so far for this CTA (per warpgroup)
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, it++) {
        tc::mbar_wait(&q_full[0], it & 1);
        tc::mbar_wait(&q_full[1], it & 1);
        tc::fence_after_sync();
        for (int t = 0; t <= nkt; t++) {
          const uint32_t st_cur = (kvi + t) % FT_STAGES, st_prev = (kvi + t - 1 + FT_STAGES) % FT_STAGES;
          if (t < nkt) {
            tc::mbar_wait(&kv_full[st_cur], ((kvi + t) / FT_STAGES) & 1);
            tc::fence_after_sync();
          }
          for (int w = 0; w < 2; w++) {
            if (t >= 1) {   // P_w(t-1) is in smem, S_w and the PV_w buffer have been drained by the softmax warpgroup
              tc::mbar_wait(&p_full[w], (cnt + t - 1) & 1);
              tc::fence_after_sync();
            }
            if (t < nkt) {
              const uint64_t ad = tc::umma_desc_sw128(q_addr + w * FT_Q_BYTES);
              const uint64_t bd = tc::umma_desc_sw128(kv_addr + st_cur * FT_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < FT_D / 16; k++)
                tc::umma_f16(tmem_base + w * 128, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s, k != 0);
              tc::umma_commit(&s_full[w]);
            }
            if (t >= 1) {
              const uint32_t pa = p_addr + w * FT_P_BYTES;
              const uint32_t va = kv_addr + st_prev * FT_STAGE_BYTES + FT_K_BYTES;
#pragma unroll
              for (int k = 0; k < FT_BN / 16; k++) {
                const uint64_t ad = tc::umma_desc_sw128(pa + (k >> 2) * (FT_BM * 128)) + (uint64_t)((k & 3) * 2);
                const uint64_t bd = tc::umma_desc_sw128(va + (k >> 2) * (FT_D * 128)) + (uint64_t)((k & 3) * 2);
                tc::umma_f16(tmem_base + 256 + w * 64, ad, bd, idesc_o, k != 0);
              }
              tc::umma_commit(&o_full[w]);
            }
          }
          if (t >= 1) tc::umma_commit(&kv_empty[st_prev]);
        }
        kvi += nkt;
        cnt += nkt;
      }
    }
  } else {
    // ================= softmax warpgroups =================
    const int w = warp >> 2;
    const int r = threadIdx.x & 127;
    const int sw = r & 7;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + w * 128;
    const uint32_t tO = tmem_base + lane_sel + 256 + w * 64;
    uint8_t* rowQ = sQ + w * FT_Q_BYTES + r * 128;
    uint8_t* rowP = sP + w * FT_P_BYTES + r * 128;
    uint32_t cnt = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int b = pair / pairs_per_img;
      const int tile = (pair - b * pairs_per_img) * 2 + w;
      const int row = tile * FT_BM + r;
      const bool valid = row < Nq;
      {
        const float4* __restrict__ qrow = reinterpret_cast<const float4*>(q + ((long long)b * Nq + (valid ? row : 0)) * FT_D);
#pragma unroll
        for (int c = 0; c < 8; c++) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
          if (valid) { a = qrow[2 * c]; d = qrow[2 * c + 1]; }
          st_shared_v4(rowQ + ((c ^ sw) << 4), pack_h2(a.x * qscale, a.y * qscale), pack_h2(a.z * qscale, a.w * qscale),
                       pack_h2(d.x * qscale, d.y * qscale), pack_h2(d.z * qscale, d.w * qscale));
        }
      }
      tc::fence_before_sync();
      tc::fence_proxy_async();
      tc::mbar_arrive(&q_full[w]);

      float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float O[FT_D];
#pragma unroll
      for (int i = 0; i < FT_D; i++) O[i] = 0.f;

      for (int t = 0; t < nkt; t++, cnt++) {
        const int ncols = min(FT_BN, Nk - t * FT_BN);
        tc::mbar_wait(&s_full[w], cnt & 1);
        tc::fence_after_sync();
        float tmax = -INFINITY;
        if (ncols == FT_BN) {
          tmax = chunk_max<32, false>(tS, 0, ncols, tmax);
          tmax = chunk_max<32, false>(tS, 32, ncols, tmax);
          tmax = chunk_max<32, false>(tS, 64, ncols, tmax);
          tmax = chunk_max<16, false>(tS, 96, ncols, tmax);
        } else {
          tmax = chunk_max<32, true>(tS, 0, ncols, tmax);
          tmax = chunk_max<32, true>(tS, 32, ncols, tmax);
          tmax = chunk_max<32, true>(tS, 64, ncols, tmax);
          tmax = chunk_max<16, true>(tS, 96, ncols, tmax);
        }
        const float m_new = fmaxf(m, tmax);
        const float alpha = ex2f(m - m_new);
        if (t > 0) {   // fold P(t-1) V(t-1) into the register accumulator; also frees the P buffer
          tc::mbar_wait(&o_full[w], (cnt - 1) & 1);
          tc::fence_after_sync();
#pragma unroll
          for (int h = 0; h < 2; h++) {
            uint32_t v[32];
            tc::tmem_ld32(tO + h * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) O[h * 32 + j] = fmaf(O[h * 32 + j], alpha_prev, __uint_as_float(v[j]));
          }
        }
        float lsum = 0.f;
        if (ncols == FT_BN) {
          lsum += chunk_exp<32, false>(tS, 0, ncols, m_new, rowP, sw);
          lsum += chunk_exp<32, false>(tS, 32, ncols, m_new, rowP, sw);
          lsum += chunk_exp<32, false>(tS, 64, ncols, m_new, rowP, sw);
          lsum += chunk_exp<16, false>(tS, 96, ncols, m_new, rowP, sw);
        } else {
          lsum += chunk_exp<32, true>(tS, 0, ncols, m_new, rowP, sw);
          lsum += chunk_exp<32, true>(tS, 32, ncols, m_new, rowP, sw);
          lsum += chunk_exp<32, true>(tS, 64, ncols, m_new, rowP, sw);
          lsum += chunk_exp<16, true>(tS, 96, ncols, m_new, rowP, sw);
        }
        l = fmaf(l, alpha, lsum);
        m = m_new;
        alpha_prev = alpha;
        tc::fence_before_sync();
        tc::fence_proxy_async();
        tc::mbar_arrive(&p_full[w]);
      }
      // last P V product
      tc::mbar_wait(&o_full[w], (cnt - 1) & 1);
      tc::fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t v[32];
        tc::tmem_ld32(tO + h * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) O[h * 32 + j] = fmaf(O[h * 32 + j], alpha_prev, __uint_as_float(v[j]));
      }
      if (valid) {
        const float inv = 1.f / l;
        float4* __restrict__ orow = reinterpret_cast<float4*>(out + ((long long)b * Nq + row) * FT_D);
#pragma unroll
        for (int i = 0; i < FT_D / 4; i++)
          orow[i] = make_float4(O[4 * i] * inv, O[4 * i + 1] * inv, O[4 * i + 2] * inv, O[4 * i + 3] * inv);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, FT_TMEM_COLS);
  }
}

// kv fp32 [B][Nk][128] (k | v)  ->  k16 [B][Nk][64] fp16,  vt16 [B][64][Nkp] fp16 (V transposed, kv contiguous)
__global__ void __launch_bounds__(256) flash_pack_kv_kernel(const float* __restrict__ kv, __half* __restrict__ k16,
                                                            __half* __restrict__ vt16, int Nk, int Nkp) {
  __shared__ float vs[32][65];
  const int b = blockIdx.y, n0 = blockIdx.x * 32, tid = threadIdx.x;
  const float* __restrict__ src = kv + ((long long)b * Nk + n0) * 128;
  for (int i = tid; i < 32 * 128; i += 256) {
    const int n = i >> 7, c = i & 127;
    const float v = (n0 + n < Nk) ? src[(long long)n * 128 + c] : 0.f;
    if (c < 64) {
      if (n0 + n < Nk) k16[((long long)b * Nk + n0 + n) * 64 + c] = __float2half_rn(v);
    } else {
      vs[n][c - 64] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < 64 * 32; i += 256) {
    const int d = i >> 5, n = i & 31;
    if (n0 + n < Nkp) vt16[((long long)b * 64 + d) * Nkp + n0 + n] = __float2half_rn(n0 + n < Nk ? vs[n][d] : 0.f);
  }
}

}  // namespace

size_t flash_tc_workspace_bytes(int B, int Nk) {
  const int Nkp = (Nk + 7) / 8 * 8;
  return align_up((size_t)B * Nk * 64 * 2, 1024) + align_up((size_t)B * 64 * Nkp * 2, 1024);
}

int launch_flash_tc(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, void* ws,
                    cudaStream_t st) {
  TCX_REQUIRE(ws != nullptr && ((uintptr_t)ws & 127) == 0, "flash_tc: workspace must be 128-byte aligned");
  const int Nkp = (Nk + 7) / 8 * 8;
  __half* k16 = reinterpret_cast<__half*>(ws);
  __half* vt16 = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(ws) + align_up((size_t)B * Nk * 64 * 2, 1024));
  {
    dim3 grid(cdiv(Nkp, 32), B);
    flash_pack_kv_kernel<<<grid, 256, 0, st>>>(kv, k16, vt16, Nk, Nkp);
    TCX_TRY(tcx_check_launch("flash_pack_kv"));
  }
  FlashMaps maps;
  TCX_TRY(tcx_make_operand_map(&maps.k, k16, 2, 64, Nk, 64, B, (long long)Nk * 64, 64, FT_BN));
  TCX_TRY(tcx_make_operand_map(&maps.vt, vt16, 2, Nkp, 64, Nkp, B, (long long)64 * Nkp, 64, 64));
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    TCX_REQUIRE(e == cudaSuccess, "flash_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    done = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int tiles_per_img = cdiv(Nq, FT_BM);
  const int npairs = B * ((tiles_per_img + 1) / 2);
  const float qscale = scale * 1.4426950408889634f;
  ProfScope prof("flash_tc", st);
  flash_tc_kernel<<<min(npairs, sms), FT_THREADS, FT_SMEM, st>>>(maps, q, out, B, Nq, Nk, qscale);
  return tcx_check_launch("flash_tc");
}
