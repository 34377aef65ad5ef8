// Weight-gradient GEMM on tcgen05 with BOTH operands read in place as MN-major tiles:
//
//     D[i][j] = alpha * sum_t A[t][i] * B[t][j]          A = dy [tokens][NL], B = x [tokens][KL]; both fp32 (TF32 MMA),
//                                                        both fp16 or both bf16 (tcgen05 rejects mixed A / B formats)
//
// i.e. dW = dy^T x of y = x W^T (+ b).  The contraction runs over the token axis, which is the SLOW axis of both operands
// as the rest of the pipeline stores them, so the round-1 path transposed both (bwd_packT) before a K-major GEMM; here a
// TMA box {128 bytes of channels (64 x 16-bit | 32 x fp32), 64 tokens} with the 128-byte swizzle (32-byte base for fp32 / TF32) lands as the canonical MN-major UMMA tile
// (tc::umma_desc_mn_sw128) and instruction-descriptor bits 15 / 16 select MN-major A / B.  No re-layout kernels.
//
// Grid = output tiles (128 x BN) x token splits x batch; one CTA = one (tile, split): 6 warps — TMA producer through a
// 4..8-stage smem ring, one-thread tcgen05.mma issuer into one TMEM accumulator, 4 epilogue warps (TMEM lane quarters).
//  * split-K over tokens with an ORDERED fold through distributed shared memory: the S <= 8 splits of a tile are one
//    thread-block cluster; every CTA parks its fp32 tile in its own shared memory (the drained pipeline stages), and after a
//    cluster barrier rank r sums rows [r*128/S, (r+1)*128/S) over the S peers in rank order (ld.shared::cluster), applies
//    alpha and writes dW.  A fixed association: bit-reproducible run to run; no scratch, no counters, no second kernel.
//  * the bias gradient db[i] = sum_t dy[t][i] rides along as one extra N = 64 MMA per k-step against a constant tile of
//    ones (any layout of an all-ones tile is valid), accumulated in 64 spare TMEM columns and folded the same way.
//  * optional head mask (Multi-Branch attention: only the diagonal Ch x Ch blocks of the C x C context are real) and a
//    transposed second copy of the result for the consumers that need ctx^T.
#include <algorithm>
#include <mutex>
#include <cuda_fp16.h>
#include "bwd.cuh"
#include "tc.cuh"

namespace {

constexpr int WG_BM = 128;
constexpr int WG_KT = 64;                      // tokens per k-block = rows of one TMA box
constexpr int WG_BOX_BYTES = WG_KT * 128;      // 8 KB
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_THREADS = 192;
constexpr int WG_ONES_BYTES = 2048;            // 16 k rows x 128 bytes of fp16 1.0
constexpr int WG_DBN = 64;                     // columns of the bias-gradient MMA (one canonical 64-wide MN-major group)
constexpr int WG_MAX_SPLIT = 8;                // portable cluster size
constexpr int WG_RING_BYTES = 192 * 1024;      // pipeline stages (later: the parked fp32 tile, <= 128 KB)
constexpr int WG_MAX_S2 = 16;                  // clusters per tile
constexpr int WG_TICKETS = 1 << 16;

__device__ unsigned g_wgrad_tickets[WG_TICKETS];   // zero at load; the cluster that folds a tile resets its slot

struct WgradMaps {
  CUtensorMap a, b;
};

struct WgradParams {
  int Mtok;              // tokens per batch item
  int NL, KL;            // rows / columns of the result
  int S, Ms;             // token splits = CS * S2, tokens per split (multiple of WG_KT)
  int CS, S2;            // cluster size (splits folded through distributed shared memory), clusters per tile (folded through HBM)
  float* part;           // [batch][tiles][S2][128][BN] cluster sums (S2 > 1)
  float* part_db;        // [ntm][S2][128] (S2 > 1 and db)
  unsigned ticket_base;
  int BN;                // tile columns: 64 | 128 | 256
  int stages;
  int ntm, ntn;          // tiles along rows / columns
  int batch;
  float alpha;
  float* out;            // [batch][NL][ldo]
  float* outT;           // [batch][KL][ldt] or null
  float* db;             // [NL] or null (batch == 1)
  int ldo, ldt;
  long long stride_out, stride_outT;
  int mask_ch;           // > 0: keep only elements with i / mask_ch == j / mask_ch
  int fmt;               // 0 fp16, 1 bf16, 2 fp32 operands (TF32 MMA)
  int ring;              // bytes of the pipeline ring (flag "smem_kb": a smaller ring lets two CTAs of different kernels share an SM)
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float4 ld_peer_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra));
  return v;
}
__device__ __forceinline__ float ld_peer_f1(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

// EB = operand element bytes: 2 (fp16 / bf16, UMMA K = 16 tokens) or 4 (fp32 read as TF32, UMMA K = 8 tokens)
template <int EB>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradMaps maps,
                                                                 const __grid_constant__ WgradParams p) {
  constexpr int WG_BOX = 128 / EB;               // channels per 128-byte row: 64 | 32
  constexpr int UK = 32 / EB;                    // tokens per MMA: 16 | 8
  constexpr int A_STAGE = (WG_BM / WG_BOX) * WG_BOX_BYTES;     // 16 KB | 32 KB
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem base has the same offset in every CTA of the cluster (same kernel, same static smem), so the aligned
  // addresses below are valid peer addresses for mapa
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = p.BN;
  const int STAGES = p.stages;
  const int stage_b = (BN / WG_BOX) * WG_BOX_BYTES;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  float* sTile = reinterpret_cast<float*>(smem);                     // [128][BN] fp32, float4 index XOR-swizzled by (row & 7)
  uint8_t* sOnes = smem + p.ring;
  float* sDb = reinterpret_cast<float*>(sOnes + WG_ONES_BYTES);      // [128]
  uint64_t* full = reinterpret_cast<uint64_t*>(sDb + WG_BM);
  uint64_t* empty = full + WG_MAX_STAGES;
  uint64_t* acc_full = empty + WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint32_t* last_flag = tmem_slot + 1;

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.ntm * p.ntn;
  const int s = blockIdx.x % p.S;              // split; cluster = CS consecutive CTAs along x: rank = s % CS, cluster c2 = s / CS
  const int rank = s % p.CS, c2 = s / p.CS;
  const int t = (blockIdx.x / p.S) % tiles;
  const int z = blockIdx.x / (p.S * tiles);
  const int ti = t / p.ntn, tj = t - ti * p.ntn;
  const int i0 = ti * WG_BM, j0 = tj * BN;
  const int tok0 = s * p.Ms;
  const int tok1 = min(tok0 + p.Ms, p.Mtok);
  const int nkb = tok1 > tok0 ? (tok1 - tok0 + WG_KT - 1) / WG_KT : 0;
  const bool do_db = p.db != nullptr && tj == 0;
  const uint32_t tmem_cols = BN + WG_DBN <= 128 ? 128 : (BN + WG_DBN <= 256 ? 256 : 512);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&maps.a);
    tc::prefetch_tmap(&maps.b);
    for (int i = 0; i < STAGES; i++) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::tmem_relinquish();
  }
  // the constant ones tile of the bias-gradient MMA
  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += WG_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = p.fmt == 2 ? 0x3F800000u : (p.fmt == 1 ? 0x3F803F80u : 0x3C003C00u);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int nboxA = min(WG_BM / WG_BOX, (p.NL - i0 + WG_BOX - 1) / WG_BOX);     // channel groups of this tile that exist
  const int nboxB = min(BN / WG_BOX, (p.KL - j0 + WG_BOX - 1) / WG_BOX);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int kb = 0; kb < nkb; kb++) {
        tc::mbar_wait(&empty[st], ph ^ 1);
        tc::mbar_arrive_expect_tx(&full[st], (uint32_t)((nboxA + nboxB) * WG_BOX_BYTES));
        const int tok = tok0 + kb * WG_KT;
        for (int b = 0; b < nboxA; b++)
          tc::tma_load_3d(sA + st * A_STAGE + b * WG_BOX_BYTES, &maps.a, i0 + b * WG_BOX, tok, z, &full[st]);
        for (int b = 0; b < nboxB; b++)
          tc::tma_load_3d(sB + st * stage_b + b * WG_BOX_BYTES, &maps.b, j0 + b * WG_BOX, tok, z, &full[st]);
        if (++st == (uint32_t)STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = tc::umma_idesc(p.fmt, WG_BM, BN, 1, 1);
      const uint32_t idesc_db = tc::umma_idesc(p.fmt, WG_BM, WG_DBN, 1, 1);
      auto mkdesc = [](uint32_t addr, uint32_t lbo) {
        return EB == 2 ? tc::umma_desc_mn_sw128(addr, lbo) : tc::umma_desc_mn_sw128_base32(addr, lbo);
      };
      const uint64_t od = mkdesc(tc::smem_u32(sOnes), 1024);     // fp32: the second 32-column group 1 KB further
      uint32_t st = 0, ph = 0;
      for (int kb = 0; kb < nkb; kb++) {
        tc::mbar_wait(&full[st], ph);
        tc::fence_after_sync();
        const uint64_t ad = mkdesc(tc::smem_u32(sA + st * A_STAGE), WG_BOX_BYTES);
        const uint64_t bd = mkdesc(tc::smem_u32(sB + st * stage_b), WG_BOX_BYTES);
#pragma unroll
        for (int k = 0; k < WG_KT / UK; k++) {      // one MMA = UK tokens = UK rows of 128 bytes
          const uint64_t adv = (uint64_t)(k * (UK * 128 >> 4));
          if (EB == 2) {
            tc::umma_f16(tmem_base, ad + adv, bd + adv, idesc, (kb | k) != 0);
            if (do_db) tc::umma_f16(tmem_base + BN, ad + adv, od, idesc_db, (kb | k) != 0);
          } else {
            tc::umma_tf32(tmem_base, ad + adv, bd + adv, idesc, (kb | k) != 0);
            if (do_db) tc::umma_tf32(tmem_base + BN, ad + adv, od, idesc_db, (kb | k) != 0);
          }
        }
        tc::umma_commit(&empty[st]);
        if (++st == (uint32_t)STAGES) { st = 0; ph ^= 1; }
      }
      tc::umma_commit(acc_full);
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4, one result row per thread =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int i = i0 + row;
    if (nkb > 0) {
      tc::mbar_wait(acc_full, 0);           // every MMA has retired: the pipeline stages are free to hold the fp32 tile
      tc::fence_after_sync();
    }
    const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool direct = p.S == 1;
    float* orow = p.out + (size_t)z * p.stride_out + (size_t)i * p.ldo;
    const int q4 = BN / 4;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (j0 + c0 >= p.KL) break;
      uint32_t v[32];
      if (nkb > 0) {
        tc::tmem_ld32(tacc + c0, v);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int c = 0; c < 32; c++) v[c] = 0u;
      }
      if (!direct) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const int f4 = ((c0 + c) >> 2) ^ (row & 7);
          *reinterpret_cast<float4*>(sTile + ((size_t)row * q4 + f4) * 4) =
              make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
        }
      } else if (i < p.NL) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const int j = j0 + c0 + c;
          if (j >= p.KL) break;
          float y[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            y[q] = __uint_as_float(v[c + q]) * p.alpha;
            if (p.mask_ch > 0 && i / p.mask_ch != (j + q) / p.mask_ch) y[q] = 0.f;
          }
          if (j + 3 < p.KL) {
            *reinterpret_cast<float4*>(orow + j) = make_float4(y[0], y[1], y[2], y[3]);
          } else {
            for (int q = 0; q < 4 && j + q < p.KL; q++) orow[j + q] = y[q];
          }
          if (p.outT)
            for (int q = 0; q < 4 && j + q < p.KL; q++) p.outT[(size_t)z * p.stride_outT + (size_t)(j + q) * p.ldt + i] = y[q];
        }
      }
    }
    if (do_db) {
      uint32_t d[16];
      if (nkb > 0) {
        tc::tmem_ld16(tacc + BN, d);
        tc::tmem_ld_wait();
      } else {
        d[0] = 0u;
      }
      if (direct) {
        if (i < p.NL) p.db[i] = __uint_as_float(d[0]) * p.alpha;
      } else {
        sDb[row] = __uint_as_float(d[0]);
      }
    }
    tc::fence_before_sync();
  }

  // ===== ordered fold, level 1: over the cluster through distributed shared memory.  Rank r owns rows [r*128/CS, (r+1)*128/CS)
  // and adds the peers' tiles in rank order.  Level 2 (S2 > 1): the cluster sums go to HBM; the cluster that draws the last
  // ticket of the tile adds the S2 cluster sums in cluster order (each rank its own rows) =====
  __syncwarp();                                  // lanes of the producer / issuer warps reconverge before the aligned barriers
  if (p.S > 1) {
    const int q4 = BN / 4;
    const int rows_per = WG_BM / p.CS;
    const int r0 = rank * rows_per;
    const bool final1 = p.S2 == 1;
    auto finish = [&](int i, int j, float4 acc) {       // alpha, head mask, store (+ transposed copy)
      float y[4] = {acc.x * p.alpha, acc.y * p.alpha, acc.z * p.alpha, acc.w * p.alpha};
      if (p.mask_ch > 0) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (i / p.mask_ch != (j + q) / p.mask_ch) y[q] = 0.f;
      }
      float* orow = p.out + (size_t)z * p.stride_out + (size_t)i * p.ldo + j;
      if (j + 3 < p.KL) {
        *reinterpret_cast<float4*>(orow) = make_float4(y[0], y[1], y[2], y[3]);
      } else {
        for (int q = 0; q < 4 && j + q < p.KL; q++) orow[q] = y[q];
      }
      if (p.outT) {
        for (int q = 0; q < 4 && j + q < p.KL; q++) p.outT[(size_t)z * p.stride_outT + (size_t)(j + q) * p.ldt + i] = y[q];
      }
    };
    float* cpart = final1 ? nullptr : p.part + (((size_t)z * tiles + t) * p.S2 + c2) * WG_BM * BN;
    float* cpart_db = final1 || !do_db ? nullptr : p.part_db + ((size_t)ti * p.S2 + c2) * WG_BM;
    if (p.CS > 1) {
      cluster_sync_all();                        // every peer's tile is parked (release / acquire at cluster scope)
      const uint32_t tile_addr = tc::smem_u32(sTile);
      for (int e = threadIdx.x; e < rows_per * q4; e += WG_THREADS) {
        const int r = r0 + e / q4, c4 = e % q4;
        const int i = i0 + r, j = j0 + c4 * 4;
        if (i >= p.NL || j >= p.KL) continue;
        const uint32_t addr = tile_addr + (uint32_t)(((size_t)r * q4 + (c4 ^ (r & 7))) * 16);
        float4 pv[WG_MAX_SPLIT];
#pragma unroll
        for (int k = 0; k < WG_MAX_SPLIT; k++)       // all peer loads in flight before the first add
          if (k < p.CS) pv[k] = ld_peer_f4(addr, (uint32_t)k);
        float4 acc = pv[0];
#pragma unroll
        for (int k = 1; k < WG_MAX_SPLIT; k++)
          if (k < p.CS) { acc.x += pv[k].x; acc.y += pv[k].y; acc.z += pv[k].z; acc.w += pv[k].w; }
        if (final1) finish(i, j, acc);
        else *reinterpret_cast<float4*>(cpart + (size_t)r * BN + c4 * 4) = acc;
      }
      if (do_db) {
        const uint32_t db_addr = tc::smem_u32(sDb);
        for (int e = threadIdx.x; e < rows_per; e += WG_THREADS) {
          const int r = r0 + e;
          if (i0 + r >= p.NL) continue;
          float acc = ld_peer_f1(db_addr + r * 4, 0);
          for (int k = 1; k < p.CS; k++) acc += ld_peer_f1(db_addr + r * 4, (uint32_t)k);
          if (final1) p.db[i0 + r] = acc * p.alpha;
          else cpart_db[r] = acc;
        }
      }
    }
    if (!final1) {
      __threadfence();                           // this CTA's slice of the cluster sum is visible device-wide
      cluster_sync_all();                        // ... and so is every peer's: the cluster sum is complete
      if (rank == 0 && threadIdx.x == 0) {
        unsigned* tk = &g_wgrad_tickets[(p.ticket_base + (unsigned)(z * tiles + t)) & (WG_TICKETS - 1)];
        const unsigned old = atomicAdd(tk, 1u);
        const bool last = old == (unsigned)(p.S2 - 1);
        if (last) *tk = 0u;                      // every cluster has drawn: the slot is free for the next launch
        *last_flag = last ? 1u : 0u;
      }
      cluster_sync_all();
      uint32_t is_last;
      {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(tc::smem_u32(last_flag)), "r"(0));
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(is_last) : "r"(ra) : "memory");
      }
      if (is_last) {
        __threadfence();
        const float* base = p.part + ((size_t)z * tiles + t) * p.S2 * WG_BM * BN;
        const size_t cstride = (size_t)WG_BM * BN;
        for (int e = threadIdx.x; e < rows_per * q4; e += WG_THREADS) {
          const int r = r0 + e / q4, c4 = e % q4;
          const int i = i0 + r, j = j0 + c4 * 4;
          if (i >= p.NL || j >= p.KL) continue;
          const float* src = base + (size_t)r * BN + c4 * 4;
          float4 acc = __ldcg(reinterpret_cast<const float4*>(src));
          int k = 1;
          for (; k + 3 < p.S2; k += 4) {         // four independent loads in flight, added in cluster order
            const float4 a0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * cstride));
            const float4 a1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(k + 1) * cstride));
            const float4 a2 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(k + 2) * cstride));
            const float4 a3 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(k + 3) * cstride));
            acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
            acc.x += a1.x; acc.y += a1.y; acc.z += a1.z; acc.w += a1.w;
            acc.x += a2.x; acc.y += a2.y; acc.z += a2.z; acc.w += a2.w;
            acc.x += a3.x; acc.y += a3.y; acc.z += a3.z; acc.w += a3.w;
          }
          for (; k < p.S2; k++) {
            const float4 a0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * cstride));
            acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
          }
          finish(i, j, acc);
        }
        if (do_db) {
          for (int e = threadIdx.x; e < rows_per; e += WG_THREADS) {
            const int r = r0 + e;
            if (i0 + r >= p.NL) continue;
            const float* src = p.part_db + (size_t)ti * p.S2 * WG_BM + r;
            float acc = 0.f;
            for (int k = 0; k < p.S2; k++) acc += __ldcg(src + (size_t)k * WG_BM);
            p.db[i0 + r] = acc * p.alpha;
          }
        }
      }
    }
    cluster_sync_all();                          // no CTA leaves (and frees its shared memory) while a peer still reads it
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

inline int wg_bn(int KL, int eb) { return KL > 128 && eb == 2 ? 256 : (KL > 64 ? 128 : 64); }
unsigned g_ticket_next = 0;

}  // namespace

// ---- split plan ------------------------------------------------------------------------------------------------------------------
// Measured on B200 (tools/r2_wgrad_sweep.py, profiles/r02_wgrad_sweep.txt): one CTA streams its 48-64 KB k-blocks (64 tokens) at
// ~57 GB/s, i.e. ~0.8 us each (the TMA unit of one SM, 128-byte rows), the cluster fold costs ~2.5 us and the second (HBM) level
// ~3 us more; a launch has ~7 us of fixed cost.  So: pick the split count that minimises
//   waves * k-blocks-per-CTA * 0.3 us + fold overheads,   S in {1, 2, 4, 8} (one cluster) or 8 * S2 with S2 <= 16,
// keeping tiles * S * batch within about one wave of the 148 SMs.
// CTAs of this kernel that can be resident at once for cluster size cs (index log2 cs): a cluster needs cs free SMs inside ONE GPC, so
// with 8 GPCs of 18-20 SMs only 16 clusters of 8 fit — 128 CTAs, not 148 (cudaOccupancyMaxActiveClusters; measured per device once).
// A plan with more CTAs than that runs in two waves.
static int wgrad_capacity(int cs) {
  static int cap[64][4];                       // [device][log2 cs], 0 = not measured yet
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  const int li = cs == 1 ? 0 : (cs == 2 ? 1 : (cs == 4 ? 2 : 3));
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cap[dev & 63][li];
  if (c == 0) {
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 * 8); cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = 1024 + WG_RING_BYTES + WG_ONES_BYTES + WG_BM * 4 + 256;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(wgrad_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, wgrad_tc_kernel<4>, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = cs == 1 ? 148 : 128 / cs; }     // query failed: the B200 values
    c = n * cs;
  }
  return c;
}

void wgrad_tc_plan(long long Mtok, int NL, int KL, int batch, int eb, int* S, int* Ms, int* BN, int* CS, int* S2) {
  const int bn = wg_bn(KL, eb);
  const long long tiles = (long long)cdiv(NL, WG_BM) * cdiv(KL, bn) * batch;
  const long long kb_total = (Mtok + WG_KT - 1) / WG_KT;
  auto tokens_per = [&](int n) { return ((Mtok + n - 1) / n + WG_KT - 1) / WG_KT * WG_KT; };
  int best_cs = 1, best_s2 = 1;
  double best = 1e30;
  for (int idx = 0; idx < 4 + WG_MAX_S2 - 1; idx++) {
    const int cs = idx < 4 ? (1 << idx) : WG_MAX_SPLIT;
    const int s2 = idx < 4 ? 1 : idx - 2;          // 2 .. WG_MAX_S2
    const int s = cs * s2;
    long long cap = g_tcx_max_ctas > 0 ? g_tcx_max_ctas : wgrad_capacity(cs);
    if (g_tcx_wgrad_ctas > 0 && cap > g_tcx_wgrad_ctas) cap = g_tcx_wgrad_ctas;
    // one wave.  Splits are whole 64-token k-blocks, so a few CTAs of a fine split may get no tokens (they park a zero tile): that is
    // accepted as long as the split still shortens the longest CTA (49 k-blocks: 8 -> 7 each, 32 -> 2 each, 25 CTAs busy)
    if (s > 1 && tiles * s > cap) continue;
    if (g_tcx_wgrad_idle ? (s > 2 && (kb_total + s - 1) / s >= (kb_total + s / 2 - 1) / (s / 2)) : (s > 1 && (long long)(s - 1) * tokens_per(s) >= Mtok)) continue;
    const long long waves = (tiles * s + cap - 1) / cap;
    const double cost = (double)waves * (double)((kb_total + s - 1) / s) * 0.8 + (s > 1 ? 2.5 : 0.0) + (s2 > 1 ? 3.0 : 0.0);
    if (cost < best) { best = cost; best_cs = cs; best_s2 = s2; }
  }
  *CS = best_cs; *S2 = best_s2; *S = best_cs * best_s2; *Ms = (int)tokens_per(*S); *BN = bn;
}

size_t wgrad_tc_scratch_floats(long long Mtok, int NL, int KL, int batch, int eb) {
  int S, Ms, BN, CS, S2;
  wgrad_tc_plan(Mtok, NL, KL, batch, eb, &S, &Ms, &BN, &CS, &S2);
  if (S2 == 1) return 64;
  const size_t tiles = (size_t)cdiv(NL, WG_BM) * cdiv(KL, BN);
  return (size_t)batch * tiles * S2 * WG_BM * BN + (size_t)cdiv(NL, WG_BM) * S2 * WG_BM + 64;
}

bool wgrad_tc_eligible(long long Mtok, int NL, int KL, int lda, int ldb, int eb) {
  const int m = 16 / eb - 1;      // 16-byte row pitches
  return Mtok >= 1 && Mtok < (1ll << 31) && NL >= 8 && KL >= 8 && !((NL | KL | lda | ldb) & m) && tcx_get_encode_tiled() != nullptr;
}

int launch_wgrad_tc(const WgradArgs& a, cudaStream_t st) {
  const int eb = a.fmt == 2 ? 4 : 2;
  TCX_REQUIRE(a.fmt >= 0 && a.fmt <= 2, "wgrad_tc: bad operand format %d", a.fmt);
  TCX_REQUIRE(wgrad_tc_eligible(a.Mtok, a.NL, a.KL, a.lda, a.ldb, eb), "wgrad_tc: shape not eligible (tokens=%lld NL=%d KL=%d lda=%d ldb=%d)",
              (long long)a.Mtok, a.NL, a.KL, a.lda, a.ldb);
  TCX_REQUIRE((((uintptr_t)a.A | (uintptr_t)a.B) & 15) == 0 && (((uintptr_t)a.out) & 15) == 0 && a.ldo % 4 == 0 && a.stride_out % 4 == 0,
              "wgrad_tc: operands must be 16-byte aligned");
  TCX_REQUIRE(a.batch >= 1 && (a.db == nullptr || a.batch == 1), "wgrad_tc: bias gradient needs batch 1");
  WgradParams p{};
  wgrad_tc_plan(a.Mtok, a.NL, a.KL, a.batch, eb, &p.S, &p.Ms, &p.BN, &p.CS, &p.S2);
  p.Mtok = (int)a.Mtok; p.NL = a.NL; p.KL = a.KL; p.batch = a.batch;
  p.ntm = cdiv(a.NL, WG_BM); p.ntn = cdiv(a.KL, p.BN);
  p.alpha = a.alpha;
  p.out = a.out; p.outT = a.outT; p.db = a.db; p.ldo = a.ldo; p.ldt = a.ldt;
  p.stride_out = a.stride_out; p.stride_outT = a.stride_outT; p.mask_ch = a.mask_ch; p.fmt = a.fmt;
  const int boxc = 128 / eb;
  const int stage = (WG_BM / boxc + p.BN / boxc) * WG_BOX_BYTES;
  // the ring must hold the parked fp32 tile (128 x BN x 4 bytes) and at least two stages; 16-bit BN = 256 tiles own all 512
  // TMEM columns, so two of them can never share an SM: they keep the full ring
  p.ring = WG_RING_BYTES;
  if (g_tcx_smem_kb > 0 && p.BN <= 128) {
    int want = g_tcx_smem_kb * 1024 - (1024 + WG_ONES_BYTES + WG_BM * 4 + 256);
    const int need = std::max(WG_BM * p.BN * 4, 2 * stage);
    if (want < need) want = need;
    if (want < p.ring) p.ring = want / 1024 * 1024;
  }
  p.stages = p.ring / stage;
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  const int tiles = p.ntm * p.ntn;
  if (p.S2 > 1) {
    TCX_REQUIRE(a.scratch != nullptr, "wgrad_tc: scratch is null");
    TCX_REQUIRE(tiles * a.batch < WG_TICKETS / 8, "wgrad_tc: too many tiles (%d)", tiles * a.batch);
    p.part = a.scratch;
    p.part_db = a.scratch + (size_t)a.batch * tiles * p.S2 * WG_BM * p.BN;
    p.ticket_base = g_ticket_next;
    g_ticket_next = (g_ticket_next + (unsigned)(tiles * a.batch)) & (WG_TICKETS - 1);
  }
  WgradMaps maps;
  TCX_TRY(tcx_make_operand_map(&maps.a, a.A, eb, a.NL, a.Mtok, a.lda, a.batch, a.strideA, boxc, WG_KT, a.fmt == 1, eb == 4));
  TCX_TRY(tcx_make_operand_map(&maps.b, a.B, eb, a.KL, a.Mtok, a.ldb, a.batch, a.strideB, boxc, WG_KT, a.fmt == 1, eb == 4));
  const size_t smem = 1024 + (size_t)p.ring + WG_ONES_BYTES + WG_BM * 4 + 256;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    TCX_REQUIRE(e == cudaSuccess, "wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  const unsigned grid = (unsigned)(tiles * p.S * a.batch);
  ProfScope prof("wgrad_tc", st, (double)a.batch * a.Mtok * (a.NL + a.KL) * eb + (double)a.batch * a.NL * a.KL * 4.0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(WG_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)p.CS; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  na++;
  if (g_tcx_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = eb == 2 ? cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<2>, maps, p) : cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<4>, maps, p);
  TCX_REQUIRE(le == cudaSuccess, "wgrad_tc: launch failed: %s", cudaGetErrorString(le));
  return tcx_check_launch("wgrad_tc");
}
