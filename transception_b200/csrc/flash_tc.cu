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
//  * exact online softmax in fp32: a thread pulls its whole 112-column S row into registers with four TMEM loads in
//    flight and immediately hands the S buffer back (Q K^T of the next kv tile overlaps the exp phase); O accumulates
//    in TMEM across kv tiles and is rescaled lazily, only when a row maximum grew by more than 2^8 (warp-uniform
//    vote), so P <= 2^8 stays well inside fp16 and the read-modify-write of O is rare.
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
constexpr int FT_OFF_P = 4 * FT_Q_BYTES;               // Q: [warpgroup][2 buffers] (fp16 q arrives by TMA one pair ahead)
constexpr int FT_OFF_KV = FT_OFF_P + 2 * FT_P_BYTES;
constexpr int FT_OFF_BAR = FT_OFF_KV + FT_STAGES * FT_STAGE_BYTES;
constexpr int FT_NBAR = 2 * FT_STAGES + 16;
constexpr float FT_LAZY = 8.0f;                        // rescale O only when the row max grew by > 2^8 (log2 units)
constexpr int FT_SMEM = FT_OFF_BAR + FT_NBAR * 8 + 16 + 1024;
constexpr int FT_THREADS = 384;                        // 2 softmax warpgroups + 1 service warpgroup (TMA, MMA, 2 idle)
constexpr uint32_t FT_TMEM_COLS = 512;                 // S[2] at 0,128 ; PV[2] at 256,320 ; P[2] (fp16 pairs) at 384,448
constexpr bool FT_P_TMEM = true;                       // P = 2^(S-m) handed to the P V MMA through tensor memory (A from TMEM)

struct FlashMaps {
  CUtensorMap k;    // k16  [B][Nk][64]      box {64, 112, 1}
  CUtensorMap vt;   // vt16 [B][64][Nkp]     box {64, 64, 1}
  CUtensorMap q;    // q16  [B][Nq][64]      box {64, 128, 1}   (fp16 path only)
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

// S tile row (112 fp32 columns) TMEM -> registers: four loads in flight, one wait.  Constant indices only, so the
// row stays in registers (taking the address of a sub-array would demote it to local memory).
template <int O>
__device__ __forceinline__ void ld32_at(uint32_t taddr, uint32_t (&s)[FT_BN]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(s[O + 0]), "=r"(s[O + 1]), "=r"(s[O + 2]), "=r"(s[O + 3]), "=r"(s[O + 4]), "=r"(s[O + 5]), "=r"(s[O + 6]), "=r"(s[O + 7]), "=r"(s[O + 8]), "=r"(s[O + 9]), "=r"(s[O + 10]), "=r"(s[O + 11]), "=r"(s[O + 12]), "=r"(s[O + 13]), "=r"(s[O + 14]), "=r"(s[O + 15]), "=r"(s[O + 16]), "=r"(s[O + 17]), "=r"(s[O + 18]), "=r"(s[O + 19]), "=r"(s[O + 20]), "=r"(s[O + 21]), "=r"(s[O + 22]), "=r"(s[O + 23]), "=r"(s[O + 24]), "=r"(s[O + 25]), "=r"(s[O + 26]), "=r"(s[O + 27]), "=r"(s[O + 28]), "=r"(s[O + 29]), "=r"(s[O + 30]), "=r"(s[O + 31])
               : "r"(taddr)
               : "memory");
}
template <int O>
__device__ __forceinline__ void ld16_at(uint32_t taddr, uint32_t (&s)[FT_BN]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(s[O + 0]), "=r"(s[O + 1]), "=r"(s[O + 2]), "=r"(s[O + 3]), "=r"(s[O + 4]), "=r"(s[O + 5]), "=r"(s[O + 6]), "=r"(s[O + 7]), "=r"(s[O + 8]), "=r"(s[O + 9]), "=r"(s[O + 10]), "=r"(s[O + 11]), "=r"(s[O + 12]), "=r"(s[O + 13]), "=r"(s[O + 14]), "=r"(s[O + 15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void load_s_row(uint32_t tS, uint32_t (&s)[FT_BN]) {
  ld32_at<0>(tS, s);
  ld32_at<32>(tS + 32, s);
  ld32_at<64>(tS + 64, s);
  ld16_at<96>(tS + 96, s);
  tc::tmem_ld_wait();
}

// Q16 / O16: q rows and the output are fp16 (the fp16 pipeline) instead of fp32.  q is staged UNSCALED; the softmax
// scale (times log2 e) is applied to the scores inside the exponent FFMA, which costs nothing extra.
template <bool Q16, bool O16>
__global__ void __launch_bounds__(FT_THREADS, 1) flash_tc_kernel(const __grid_constant__ FlashMaps maps,
                                                                 const void* __restrict__ qv, void* __restrict__ outv,
                                                                 float* __restrict__ lse, int B, int Nq, int Nk, float qscale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sP = smem + FT_OFF_P;
  uint8_t* sKV = smem + FT_OFF_KV;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + FT_OFF_BAR);
  uint64_t* kv_empty = kv_full + FT_STAGES;
  uint64_t* q_full = kv_empty + FT_STAGES;   // [warpgroup][buffer]
  uint64_t* q_empty = q_full + 4;            // [warpgroup][buffer]
  uint64_t* s_full = q_empty + 4;            // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_full = p_full + 2;             // [2]
  uint64_t* s_free = o_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = (Nq + FT_BM - 1) / FT_BM;
  const int pairs_per_img = (tiles_per_img + 1) / 2;
  const int npairs = B * pairs_per_img;
  const int nkt = (Nk + FT_BN - 1) / FT_BN;

  if (warp == 8 && lane == 0) {
    tc::prefetch_tmap(&maps.k);
    tc::prefetch_tmap(&maps.vt);
    if (Q16) tc::prefetch_tmap(&maps.q);
    for (int s = 0; s < FT_STAGES; s++) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 4; i++) { tc::mbar_init(&q_full[i], Q16 ? 1 : 128); tc::mbar_init(&q_empty[i], 1); }
    for (int w = 0; w < 2; w++) {
      tc::mbar_init(&s_full[w], 1);
      tc::mbar_init(&p_full[w], 128);
      tc::mbar_init(&o_full[w], 1);
      tc::mbar_init(&s_free[w], 128);
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
  pdl_wait();

  // register redistribution (the service warpgroup gives registers to the two softmax warpgroups); issued inside the
  // role branches so that ptxas allocates each region against its own budget.  The pool is the CTA's launch allocation
  // (168 x 384): 2 x (208 - 168) x 128 taken == (168 - 88) x 128 released, otherwise the second inc blocks forever.
  if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == 8) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t kvi = 0, it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, it++) {
        const int b = pair / pairs_per_img;
        if (Q16) {   // both query tiles of this pair, into buffer it & 1 (free once the previous user's Q K^T retired)
          const uint32_t buf = it & 1;
          for (int w = 0; w < 2; w++) {
            const int tile = (pair - b * pairs_per_img) * 2 + w;
            tc::mbar_wait(&q_empty[w * 2 + buf], ((it >> 1) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&q_full[w * 2 + buf], FT_Q_BYTES);
            tc::tma_load_3d(sQ + (w * 2 + buf) * FT_Q_BYTES, &maps.q, 0, tile * FT_BM, b, &q_full[w * 2 + buf]);
          }
        }
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
      uint32_t qbuf = 0;
      auto issue_qk = [&](int w, uint32_t stage, bool last) {
        const uint64_t ad = tc::umma_desc_sw128(q_addr + (w * 2 + qbuf) * FT_Q_BYTES);
        const uint64_t bd = tc::umma_desc_sw128(kv_addr + stage * FT_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < FT_D / 16; k++)
          tc::umma_f16(tmem_base + w * 128, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s, k != 0);
        tc::umma_commit(&s_full[w]);
        if (Q16 && last) tc::umma_commit(&q_empty[w * 2 + qbuf]);   // this Q buffer may be refilled once these retire
      };
      auto issue_pv = [&](int w, int t, uint32_t stage) {
        const uint32_t pa = p_addr + w * FT_P_BYTES;
        const uint32_t va = kv_addr + stage * FT_STAGE_BYTES + FT_K_BYTES;
#pragma unroll
        for (int k = 0; k < FT_BN / 16; k++) {
          const uint64_t bd = tc::umma_desc_sw128(va + (k >> 2) * (FT_D * 128)) + (uint64_t)((k & 3) * 2);
          if (FT_P_TMEM) {
            // A = P from tensor memory: lane = query row, 8 columns (16 fp16) per K step
            tc::umma_f16_ts(tmem_base + 256 + w * 64, tmem_base + 384 + w * 64 + k * 8, bd, idesc_o, (t | k) != 0);
          } else {
            const uint64_t ad = tc::umma_desc_sw128(pa + (k >> 2) * (FT_BM * 128)) + (uint64_t)((k & 3) * 2);
            tc::umma_f16(tmem_base + 256 + w * 64, ad, bd, idesc_o, (t | k) != 0);
          }
        }
        tc::umma_commit(&o_full[w]);
      };
      // Fixed issue order per kv tile (blocking waits; a polling scheduler was measured slower: every probe of an
      // mbarrier costs ~60-90 cycles):  QK_0(t+1), QK_1(t+1), PV_0(t), PV_1(t).
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, it++) {
        qbuf = Q16 ? (it & 1) : 0;
        tc::mbar_wait(&kv_full[kvi % FT_STAGES], (kvi / FT_STAGES) & 1);
        for (int w = 0; w < 2; w++) {
          tc::mbar_wait(&q_full[w * 2 + qbuf], Q16 ? ((it >> 1) & 1) : (it & 1));
          // with q arriving by TMA the first Q K^T of a pair no longer waits for the warpgroup's whole previous pair:
          // it only needs the S buffer, i.e. the previous pair's last S tile copied to registers
          if (Q16 && cnt > 0) tc::mbar_wait(&s_free[w], (cnt - 1) & 1);
          tc::fence_after_sync();
          issue_qk(w, kvi % FT_STAGES, nkt == 1);
        }
        for (int t = 0; t < nkt; t++) {
          const uint32_t st_cur = (kvi + t) % FT_STAGES, st_next = (kvi + t + 1) % FT_STAGES;
          if (t + 1 < nkt) {
            tc::mbar_wait(&kv_full[st_next], ((kvi + t + 1) / FT_STAGES) & 1);
            for (int w = 0; w < 2; w++) {     // S_w(t) is in registers -> the tensor core may overwrite it
              tc::mbar_wait(&s_free[w], (cnt + t) & 1);
              tc::fence_after_sync();
              issue_qk(w, st_next, t + 2 == nkt);
            }
          }
          for (int w = 0; w < 2; w++) {       // P_w(t) is in smem (and O_w has been rescaled if needed)
            tc::mbar_wait(&p_full[w], (cnt + t) & 1);
            tc::fence_after_sync();
            issue_pv(w, t, st_cur);
          }
          tc::umma_commit(&kv_empty[st_cur]);
        }
        kvi += nkt;
        cnt += nkt;
      }
    }
  } else if (warp < 8) {
    // ================= softmax warpgroups =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int w = warp >> 2;
    const int r = threadIdx.x & 127;
    const int sw = r & 7;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + w * 128;
    const uint32_t tO = tmem_base + lane_sel + 256 + w * 64;
    uint8_t* rowQ = sQ + (w * 2) * FT_Q_BYTES + r * 128;
    uint8_t* rowP = sP + w * FT_P_BYTES + r * 128;
    uint32_t cnt = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int b = pair / pairs_per_img;
      const int tile = (pair - b * pairs_per_img) * 2 + w;
      const int row = tile * FT_BM + r;
      const bool valid = row < Nq;
      if (!Q16) {   // fp32 queries: converted to fp16 and staged by their owner threads (the fp16 path uses TMA)
        const float4* __restrict__ qrow = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(qv) +
                                                                          ((long long)b * Nq + (valid ? row : 0)) * FT_D);
#pragma unroll
        for (int c = 0; c < 8; c++) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
          if (valid) { a = qrow[2 * c]; d = qrow[2 * c + 1]; }
          st_shared_v4(rowQ + ((c ^ sw) << 4), pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(d.x, d.y), pack_h2(d.z, d.w));
        }
        tc::fence_before_sync();
        tc::fence_proxy_async();
        tc::mbar_arrive(&q_full[w * 2]);
      }

      float m = -INFINITY, l = 0.f;   // m: the max the probabilities are currently expressed against (may lag)
      for (int t = 0; t < nkt; t++, cnt++) {
        const int ncols = min(FT_BN, Nk - t * FT_BN);
        uint32_t sv[FT_BN];
        tc::mbar_wait(&s_full[w], cnt & 1);
        tc::fence_after_sync();
        load_s_row(tS, sv);
        tc::fence_before_sync();
        tc::mbar_arrive(&s_free[w]);          // S is in registers: Q K^T of the next kv tile may start
        if (ncols < FT_BN) {
#pragma unroll
          for (int j = 0; j < FT_BN; j++)
            if (j >= ncols) sv[j] = 0xff800000u;   // -inf
        }
        // four independent chains for the row maximum and the row sum: a single chain of 112 dependent ops per thread
        // is latency-bound with only two softmax warps per scheduler
        float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < FT_BN; j += 4) {
          tm[0] = fmaxf(tm[0], __uint_as_float(sv[j]));
          tm[1] = fmaxf(tm[1], __uint_as_float(sv[j + 1]));
          tm[2] = fmaxf(tm[2], __uint_as_float(sv[j + 2]));
          tm[3] = fmaxf(tm[3], __uint_as_float(sv[j + 3]));
        }
        const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
        const float m_new = fmaxf(m, tmax * qscale);   // running max in log2 units (qscale > 0)
        if (t == 0) {
          m = m_new;
        } else {
          tc::mbar_wait(&o_full[w], (cnt - 1) & 1);   // P(t-1) V(t-1) retired: P buffer reusable, O stable
          tc::fence_after_sync();
          if (__any_sync(0xffffffffu, m_new - m > FT_LAZY)) {
            const float alpha = ex2f(m - m_new);
            m = m_new;
            l *= alpha;
#pragma unroll
            for (int h = 0; h < 2; h++) {
              uint32_t v[32];
              tc::tmem_ld32(tO + h * 32, v);
              tc::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
              tc::tmem_st32(tO + h * 32, v);
            }
            tc::tmem_st_wait();
          }
        }
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        if (FT_P_TMEM) {
          const uint32_t tP = tmem_base + lane_sel + 384 + w * 64;
          uint32_t pk[32];
#pragma unroll
          for (int g = 0; g < 32; g++) {        // columns 0..31  (scores 0..63)
            const float p0 = ex2f(fmaf(__uint_as_float(sv[2 * g]), qscale, -m));
            const float p1 = ex2f(fmaf(__uint_as_float(sv[2 * g + 1]), qscale, -m));
            ls[(2 * g) & 3] += p0; ls[(2 * g + 1) & 3] += p1;
            pk[g] = pack_h2(p0, p1);
          }
          tc::tmem_st32(tP, pk);
          uint32_t pk2[16];
#pragma unroll
          for (int g = 0; g < 16; g++) {        // columns 32..47 (scores 64..95)
            const float p0 = ex2f(fmaf(__uint_as_float(sv[64 + 2 * g]), qscale, -m));
            const float p1 = ex2f(fmaf(__uint_as_float(sv[64 + 2 * g + 1]), qscale, -m));
            ls[(2 * g) & 3] += p0; ls[(2 * g + 1) & 3] += p1;
            pk2[g] = pack_h2(p0, p1);
          }
          tc::tmem_st16(tP + 32, pk2);
          uint32_t pk3[8];
#pragma unroll
          for (int g = 0; g < 8; g++) {         // columns 48..55 (scores 96..111)
            const float p0 = ex2f(fmaf(__uint_as_float(sv[96 + 2 * g]), qscale, -m));
            const float p1 = ex2f(fmaf(__uint_as_float(sv[96 + 2 * g + 1]), qscale, -m));
            ls[(2 * g) & 3] += p0; ls[(2 * g + 1) & 3] += p1;
            pk3[g] = pack_h2(p0, p1);
          }
          tc::tmem_st8(tP + 48, pk3);
          tc::tmem_st_wait();
        } else {
#pragma unroll
        for (int g = 0; g < FT_BN / 8; g++) {
          float p[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            p[j] = ex2f(fmaf(__uint_as_float(sv[g * 8 + j]), qscale, -m));
            ls[j & 3] += p[j];
          }
          uint8_t* dst = rowP + (g >> 3) * (FT_BM * 128) + (((g & 7) ^ sw) << 4);
          st_shared_v4(dst, pack_h2(p[0], p[1]), pack_h2(p[2], p[3]), pack_h2(p[4], p[5]), pack_h2(p[6], p[7]));
        }
        }
        l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        tc::fence_before_sync();
        tc::fence_proxy_async();
        tc::mbar_arrive(&p_full[w]);
      }
      // all P V products have been accumulated in TMEM
      tc::mbar_wait(&o_full[w], (cnt - 1) & 1);
      tc::fence_after_sync();
      float O[FT_D];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t v[32];
        tc::tmem_ld32(tO + h * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) O[h * 32 + j] = __uint_as_float(v[j]);
      }
      if (valid) {
        const float inv = 1.f / l;
        // row log2-sum-exp of the scaled scores (training: the backward kernel recomputes P = 2^(s c - lse) from it)
        if (lse) lse[(long long)b * Nq + row] = m + log2f(l);
        if (O16) {
          uint4* __restrict__ orow = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(outv) + ((long long)b * Nq + row) * FT_D);
#pragma unroll
          for (int i = 0; i < FT_D / 8; i++)
            orow[i] = make_uint4(pack_h2(O[8 * i] * inv, O[8 * i + 1] * inv), pack_h2(O[8 * i + 2] * inv, O[8 * i + 3] * inv),
                                 pack_h2(O[8 * i + 4] * inv, O[8 * i + 5] * inv), pack_h2(O[8 * i + 6] * inv, O[8 * i + 7] * inv));
        } else {
          float4* __restrict__ orow = reinterpret_cast<float4*>(reinterpret_cast<float*>(outv) + ((long long)b * Nq + row) * FT_D);
#pragma unroll
          for (int i = 0; i < FT_D / 4; i++)
            orow[i] = make_float4(O[4 * i] * inv, O[4 * i + 1] * inv, O[4 * i + 2] * inv, O[4 * i + 3] * inv);
        }
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

// fp16 kv [B][Nk][128] (k | v) -> vt16 [B][64][Nkp] (V transposed, kv contiguous); K is consumed in place
__global__ void __launch_bounds__(256) flash_pack_vt16_kernel(const __half* __restrict__ kv, __half* __restrict__ vt16, int Nk,
                                                              int Nkp) {
  __shared__ __half vs[32][66];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, n0 = blockIdx.x * 32, tid = threadIdx.x;
  const __half* __restrict__ src = kv + ((long long)b * Nk + n0) * 128 + 64;
  {
    const int n = tid >> 3, j = tid & 7;      // 32 rows x 8 vectors of 8 halfs
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (n0 + n < Nk) r = *reinterpret_cast<const uint4*>(src + (long long)n * 128 + j * 8);
    const __half* h = reinterpret_cast<const __half*>(&r);
#pragma unroll
    for (int q = 0; q < 8; q++) vs[n][j * 8 + q] = h[q];
  }
  __syncthreads();
  for (int i = tid; i < 64 * 32; i += 256) {
    const int d = i >> 5, n = i & 31;
    if (n0 + n < Nkp) vt16[((long long)b * 64 + d) * Nkp + n0 + n] = vs[n][d];
  }
}

}  // namespace

size_t flash_tc_workspace_bytes(int B, int Nk) {
  const int Nkp = (Nk + 7) / 8 * 8;
  return align_up((size_t)B * Nk * 64 * 2, 1024) + align_up((size_t)B * 64 * Nkp * 2, 1024);
}

static int flash_tc_launch(const FlashMaps& maps, const void* q, void* out, float* lse, int B, int Nq, int Nk, float scale, bool f16io,
                           cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(flash_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(flash_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    TCX_REQUIRE(e == cudaSuccess, "flash_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
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
  ProfScope prof("flash_tc", st, 4.0 * B * (double)Nq * Nk * FT_D);   // QK^T + PV FLOPs
  cudaError_t le;
  if (f16io) le = tcx_launch_pdl(flash_tc_kernel<true, true>, dim3(min(npairs, sms)), dim3(FT_THREADS), (size_t)FT_SMEM, st, maps, q, out, lse, B, Nq, Nk, qscale);
  else le = tcx_launch_pdl(flash_tc_kernel<false, false>, dim3(min(npairs, sms)), dim3(FT_THREADS), (size_t)FT_SMEM, st, maps, q, out, lse, B, Nq, Nk, qscale);
  TCX_REQUIRE(le == cudaSuccess, "flash_tc: launch failed: %s", cudaGetErrorString(le));
  return tcx_check_launch("flash_tc");
}

int launch_flash_tc(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, void* ws,
                    cudaStream_t st, float* lse) {
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
  maps.q = maps.k;   // unused by the fp32-query kernel
  return flash_tc_launch(maps, q, out, lse, B, Nq, Nk, scale, false, st);
}

// fp16 form: q16 [B][Nq][64], kv16 [B][Nk][128] (k | v), out16 [B][Nq][64]; ws holds V^T only
int launch_flash_tc16(const void* q16, const void* kv16, void* out16, int B, int Nq, int Nk, float scale, void* ws,
                      cudaStream_t st) {
  TCX_REQUIRE(ws != nullptr && ((uintptr_t)ws & 127) == 0, "flash_tc: workspace must be 128-byte aligned");
  const int Nkp = (Nk + 7) / 8 * 8;
  __half* vt16 = reinterpret_cast<__half*>(ws);
  {
    dim3 grid(cdiv(Nkp, 32), B);
    tcx_launch_pdl(flash_pack_vt16_kernel, grid, dim3(256), 0, st, reinterpret_cast<const __half*>(kv16), vt16, Nk, Nkp);
    TCX_TRY(tcx_check_launch("flash_pack_vt16"));
  }
  FlashMaps maps;
  TCX_TRY(tcx_make_operand_map(&maps.k, kv16, 2, 64, Nk, 128, B, (long long)Nk * 128, 64, FT_BN));
  TCX_TRY(tcx_make_operand_map(&maps.vt, vt16, 2, Nkp, 64, Nkp, B, (long long)64 * Nkp, 64, 64));
  TCX_TRY(tcx_make_operand_map(&maps.q, q16, 2, 64, Nq, 64, B, (long long)Nq * 64, 64, FT_BM));
  return flash_tc_launch(maps, q16, out16, nullptr, B, Nq, Nk, scale, true, st);
}
