// Bridge spatial-reduction attention BACKWARD as a tcgen05 flash kernel (reference MSTr.py:2281-2285: out = softmax(q k^T * scale) v,
// one 64-wide head, Nq = 6076 queries against Nk = 784 reduced tokens per image at 224x224).  The round-1 backward recomputed and
// MATERIALISED the scores (16 x 6076 x 784 fp32 = 305 MB, twice) and walked them with seven launches; here the probabilities are
// recomputed tile by tile from q, k and the forward's row log-sum-exp and never leave the SM.
//
// One CTA = one (image, kv tile of 128 reduced tokens); it keeps K_j and V_j in shared memory and dK_j, dV_j in tensor memory and
// walks the 48 query tiles of the image.  Per (query tile i, kv tile j), all on tcgen05 (fp16 operands, fp32 accumulate):
//     S  = Q_i K_j^T                (A = Q_i K-major,        B = K_j K-major)
//     dP = dO_i V_j^T               (A = dO_i K-major,       B = V_j K-major)
//     P  = 2^(S*c - L_i),  dS = P (dP - D_i) * scale      (one thread per query row, from tensor memory; written to shared
//                                                           memory as fp16 [q][kv] tiles)
//     dV_j += P^T dO_i              (A = P  read MN-major,   B = dO_i read MN-major — the same smem tiles, other descriptor)
//     dK_j += dS^T Q_i              (A = dS read MN-major,   B = Q_i  read MN-major)
//     dQ_ij = dS K_j                (A = dS K-major,         B = K_j  read MN-major)  -> partial [j][b][q][64] in HBM
// dQ needs the sum over the 7 kv tiles: the partials (7 x 25 MB) are folded in tile order by a small kernel — ordered, so the
// result is bit-reproducible.  Gradients are ~1e-6 per element: dO is converted to fp16 with ONE power-of-two scale per call
// derived from its absolute maximum on the device (max |dO'| = 16; dP', dS' stay far below the fp16 maximum); the outputs are
// unscaled in the epilogues.  D_i = rowsum(dO o O) is computed in fp32 by the same pre-pass.
#include <cuda_fp16.h>
#include "bwd.cuh"
#include "tc.cuh"

namespace {

constexpr int FB_T = 128;                     // query rows per tile = kv rows per tile
constexpr int FB_D = 64;
constexpr int FB_TILE = FB_T * FB_D * 2;      // 16 KB: [128 rows][64 halfs], 128-byte rows, SW128
constexpr int FB_PD = FB_T * FB_T * 2;        // 32 KB: [128 q][128 kv] fp16 as two 64-column SW128 sub-tiles
constexpr int FB_OFF_K = 0, FB_OFF_V = FB_TILE, FB_OFF_Q = 2 * FB_TILE;          // Q / dO: [2 buffers][Q | dO]
constexpr int FB_OFF_P = FB_OFF_Q + 4 * FB_TILE, FB_OFF_DS = FB_OFF_P + FB_PD;
constexpr int FB_OFF_BAR = FB_OFF_DS + FB_PD;
constexpr int FB_SMEM = FB_OFF_BAR + 256 + 1024;
constexpr int FB_THREADS = 192;               // warps 0-3: one thread per query row; warp 4: TMA; warp 5: MMA issuer
constexpr uint32_t FB_TMEM = 512;             // S 0..127 | dP 128..255 | dV 256..319 | dK 320..383 | dQ 384..447

struct FbMaps {
  CUtensorMap q, dout, kv;      // q16 / do16 [B][Nq][64] box {64, 128}; kv16 [B][Nk][128] box {64, 128}
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2s(float lo, float hi) {       // saturating
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(FB_THREADS, 1) flash_bwd_kernel(const __grid_constant__ FbMaps maps, const float* __restrict__ lse,
                                                                  const float* __restrict__ dsum, const float* __restrict__ scale_ptr,
                                                                  float* __restrict__ dq_part, float* __restrict__ dkv, int B, int Nq,
                                                                  int Nk, float scale, float qscale) {
  PDL_TOP();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + FB_OFF_BAR);
  uint64_t* q_full = kv_full + 1;        // [2]
  uint64_t* q_empty = q_full + 2;        // [2]
  uint64_t* sdp_full = q_empty + 2;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_free + 1;
  uint64_t* pds_free = pds_full + 1;
  uint64_t* dq_free = pds_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkt = (Nk + FB_T - 1) / FB_T;
  const int b = blockIdx.x / nkt, j = blockIdx.x - b * nkt;
  const int nq = (Nq + FB_T - 1) / FB_T;

  if (warp == 4 && lane == 0) {
    tc::prefetch_tmap(&maps.q);
    tc::prefetch_tmap(&maps.dout);
    tc::prefetch_tmap(&maps.kv);
    tc::mbar_init(kv_full, 1);
    for (int i = 0; i < 2; i++) { tc::mbar_init(&q_full[i], 1); tc::mbar_init(&q_empty[i], 1); }
    tc::mbar_init(sdp_full, 1);
    tc::mbar_init(sdp_free, 128);
    tc::mbar_init(pds_full, 128);
    tc::mbar_init(pds_free, 1);
    tc::mbar_init(dq_free, 128);
    tc::fence_barrier_init();
  }
  if (warp == 5) {
    tc::tmem_alloc(tmem_slot, FB_TMEM);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ================= TMA producer =================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(kv_full, 2 * FB_TILE);
      tc::tma_load_3d(smem + FB_OFF_K, &maps.kv, 0, j * FB_T, b, kv_full);
      tc::tma_load_3d(smem + FB_OFF_V, &maps.kv, 64, j * FB_T, b, kv_full);
      for (int i = 0; i < nq; i++) {
        const int buf = i & 1;
        tc::mbar_wait(&q_empty[buf], ((i >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&q_full[buf], 2 * FB_TILE);
        tc::tma_load_3d(smem + FB_OFF_Q + buf * 2 * FB_TILE, &maps.q, 0, i * FB_T, b, &q_full[buf]);
        tc::tma_load_3d(smem + FB_OFF_Q + buf * 2 * FB_TILE + FB_TILE, &maps.dout, 0, i * FB_T, b, &q_full[buf]);
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t id_s = tc::umma_idesc(0, FB_T, FB_T);              // 128 x 128, A and B K-major
      constexpr uint32_t id_kv = tc::umma_idesc(0, FB_T, FB_D, 1, 1);       // 128 x 64, A and B MN-major (dV, dK)
      constexpr uint32_t id_q = tc::umma_idesc(0, FB_T, FB_D, 0, 1);        // 128 x 64, A K-major, B MN-major (dQ)
      const uint32_t base = tc::smem_u32(smem);
      const uint32_t aK = base + FB_OFF_K, aV = base + FB_OFF_V, aP = base + FB_OFF_P, aDS = base + FB_OFF_DS;
      auto issue_sdp = [&](int i) {
        const uint32_t aQ = base + FB_OFF_Q + (i & 1) * 2 * FB_TILE, aDO = aQ + FB_TILE;
#pragma unroll
        for (int k = 0; k < FB_D / 16; k++) {
          tc::umma_f16(tmem_base + 0, tc::umma_desc_sw128(aQ) + (uint64_t)(k * 2), tc::umma_desc_sw128(aK) + (uint64_t)(k * 2), id_s, k != 0);
          tc::umma_f16(tmem_base + 128, tc::umma_desc_sw128(aDO) + (uint64_t)(k * 2), tc::umma_desc_sw128(aV) + (uint64_t)(k * 2), id_s,
                       k != 0);
        }
        tc::umma_commit(sdp_full);
      };
      tc::mbar_wait(kv_full, 0);
      tc::mbar_wait(&q_full[0], 0);
      tc::fence_after_sync();
      issue_sdp(0);
      for (int i = 0; i < nq; i++) {
        if (i + 1 < nq) {
          tc::mbar_wait(&q_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
          tc::mbar_wait(sdp_free, i & 1);                  // S(i), dP(i) have been read out of tensor memory
          tc::fence_after_sync();
          issue_sdp(i + 1);
        }
        tc::mbar_wait(pds_full, i & 1);                    // P(i), dS(i) are in shared memory
        if (i > 0) tc::mbar_wait(dq_free, (i - 1) & 1);    // dQ(i-1) has been read out of tensor memory
        tc::fence_after_sync();
        const uint32_t aQ = base + FB_OFF_Q + (i & 1) * 2 * FB_TILE, aDO = aQ + FB_TILE;
#pragma unroll
        for (int k = 0; k < FB_T / 16; k++) {              // contraction over the 128 query rows, 16 per MMA = 16 rows of 128 bytes
          const uint64_t adv = (uint64_t)(k * (16 * 128 >> 4));
          tc::umma_f16(tmem_base + 256, tc::umma_desc_mn_sw128(aP, FB_TILE) + adv, tc::umma_desc_mn_sw128(aDO, FB_TILE) + adv, id_kv,
                       (i | k) != 0);
          tc::umma_f16(tmem_base + 320, tc::umma_desc_mn_sw128(aDS, FB_TILE) + adv, tc::umma_desc_mn_sw128(aQ, FB_TILE) + adv, id_kv,
                       (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < FB_T / 16; k++) {              // contraction over the 128 kv rows: dS K-major (two 64-column sub-tiles)
          const uint64_t ad = tc::umma_desc_sw128(aDS + (k >> 2) * FB_TILE) + (uint64_t)((k & 3) * 2);
          const uint64_t bd = tc::umma_desc_mn_sw128(aK, FB_TILE) + (uint64_t)(k * (16 * 128 >> 4));
          tc::umma_f16(tmem_base + 384, ad, bd, id_q, k != 0);
        }
        tc::umma_commit(pds_free);
        tc::umma_commit(&q_empty[i & 1]);
      }
    }
  } else {
    // ================= compute warpgroup: thread = query row (S / dP / dQ) and kv row (dK / dV epilogue) =================
    const int r = threadIdx.x;
    const int sw = r & 7;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const float inv_scale = 1.0f / *scale_ptr;
    uint8_t* rowP = smem + FB_OFF_P + r * 128;
    uint8_t* rowDS = smem + FB_OFF_DS + r * 128;
    for (int i = 0; i < nq; i++) {
      const int row = i * FB_T + r;
      const bool valid = row < Nq;
      const float L = valid ? lse[(size_t)b * Nq + row] : 0.f;
      const float Dr = valid ? dsum[(size_t)b * Nq + row] : 0.f;
      tc::mbar_wait(sdp_full, i & 1);
      if (i > 0) tc::mbar_wait(pds_free, (i - 1) & 1);     // the MMAs that read P(i-1) / dS(i-1) from shared memory have retired
      tc::fence_after_sync();
      if (i > 0) {
        // dQ(i-1) partial: tensor memory -> HBM (raw, scaled by the dO scale; the fold kernel removes it)
        const int prow = (i - 1) * FB_T + r;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          uint32_t v[32];
          tc::tmem_ld32(tmem_base + lane_sel + 384 + h * 32, v);
          tc::tmem_ld_wait();
          if (prow < Nq) {
            float4* dst = reinterpret_cast<float4*>(dq_part + (((size_t)j * B + b) * Nq + prow) * FB_D + h * 32);
#pragma unroll
            for (int c = 0; c < 8; c++)
              dst[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                   __uint_as_float(v[4 * c + 3]));
          }
        }
        tc::fence_before_sync();
        tc::mbar_arrive(dq_free);
      }
#pragma unroll 1
      for (int c0 = 0; c0 < FB_T; c0 += 32) {
        uint32_t sv[32], dv[32];
        tc::tmem_ld32(tmem_base + lane_sel + c0, sv);
        tc::tmem_ld32(tmem_base + lane_sel + 128 + c0, dv);
        tc::tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int g = 0; g < 16; g++) {
          float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
          if (valid) {
            p0 = ex2f(fmaf(__uint_as_float(sv[2 * g]), qscale, -L));
            p1 = ex2f(fmaf(__uint_as_float(sv[2 * g + 1]), qscale, -L));
            d0 = p0 * (__uint_as_float(dv[2 * g]) - Dr) * scale;
            d1 = p1 * (__uint_as_float(dv[2 * g + 1]) - Dr) * scale;
          }
          pp[g] = pack_h2s(p0, p1);
          dd[g] = pack_h2s(d0, d1);
        }
        // 32 columns = 64 bytes = four 16-byte chunks of the row's 128-byte line in sub-tile c0 / 64
        const int sub = c0 >> 6, ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int phys = ((ch0 + q) ^ sw) << 4;
          *reinterpret_cast<uint4*>(rowP + sub * FB_TILE + phys) = make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
          *reinterpret_cast<uint4*>(rowDS + sub * FB_TILE + phys) = make_uint4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]);
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(sdp_free);
      tc::fence_proxy_async();
      tc::mbar_arrive(pds_full);
    }
    // last dQ partial, then dK_j / dV_j
    tc::mbar_wait(pds_free, (nq - 1) & 1);
    tc::fence_after_sync();
    {
      const int prow = (nq - 1) * FB_T + r;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + lane_sel + 384 + h * 32, v);
        tc::tmem_ld_wait();
        if (prow < Nq) {
          float4* dst = reinterpret_cast<float4*>(dq_part + (((size_t)j * B + b) * Nq + prow) * FB_D + h * 32);
#pragma unroll
          for (int c = 0; c < 8; c++)
            dst[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                 __uint_as_float(v[4 * c + 3]));
        }
      }
    }
    const int kvrow = j * FB_T + r;
#pragma unroll
    for (int h = 0; h < 4; h++) {            // columns 256..319 = dV, 320..383 = dK  ->  dkv row = [dk (64) | dv (64)]
      uint32_t v[32];
      tc::tmem_ld32(tmem_base + lane_sel + 256 + h * 32, v);
      tc::tmem_ld_wait();
      if (kvrow < Nk) {
        const int col = (h < 2 ? 64 + h * 32 : (h - 2) * 32);
        float4* dst = reinterpret_cast<float4*>(dkv + ((size_t)b * Nk + kvrow) * 128 + col);
#pragma unroll
        for (int c = 0; c < 8; c++)
          dst[c] = make_float4(__uint_as_float(v[4 * c]) * inv_scale, __uint_as_float(v[4 * c + 1]) * inv_scale,
                               __uint_as_float(v[4 * c + 2]) * inv_scale, __uint_as_float(v[4 * c + 3]) * inv_scale);
      }
    }
    tc::fence_before_sync();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, FB_TMEM);
  }
}

// ---- pre-pass 1: absolute maximum of dO (as ordered uint bits: atomicMax is order-independent, hence deterministic) ----
__global__ void __launch_bounds__(256) fb_absmax_kernel(const float* __restrict__ x, long long n4, unsigned* __restrict__ out) {
  PDL_TOP();
  __shared__ unsigned sm[8];
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = __float_as_uint(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned r = sm[0];
    for (int i = 1; i < 8; i++) r = max(r, sm[i]);
    atomicMax(out, r);
  }
}
// ---- pre-pass 2: one warp per query row: q16, do16 = fp16(s dO), dsum = s rowsum(dO o O); s = 2^(4 - ceil(log2 absmax)) is
// written to scale_out by block 0; the kv rows are converted by the tail blocks ----
__global__ void __launch_bounds__(256) fb_prep_kernel(const float* __restrict__ q, const float* __restrict__ dout, const float* __restrict__ out,
                                                      const float* __restrict__ kv, const unsigned* __restrict__ absmax, __half* __restrict__ q16,
                                                      __half* __restrict__ do16, __half* __restrict__ kv16, float* __restrict__ dsum,
                                                      float* __restrict__ scale_out, long long Mq, long long Mkv) {
  PDL_TOP();
  const float amax = __uint_as_float(*absmax);
  int e = 0;
  if (amax > 0.f) frexpf(amax, &e);                 // amax = f * 2^e, f in [0.5, 1)
  const float s = amax > 0.f ? ldexpf(1.0f, 4 - e) : 1.0f;
  if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = s;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row < Mq) {
    const float2 qv = reinterpret_cast<const float2*>(q + row * FB_D)[lane];
    const float2 gv = reinterpret_cast<const float2*>(dout + row * FB_D)[lane];
    const float2 ov = reinterpret_cast<const float2*>(out + row * FB_D)[lane];
    const float d = warp_sum(fmaf(gv.x, ov.x, gv.y * ov.y));
    reinterpret_cast<__half2*>(q16 + row * FB_D)[lane] = __floats2half2_rn(qv.x, qv.y);
    reinterpret_cast<uint32_t*>(do16 + row * FB_D)[lane] = pack_h2s(gv.x * s, gv.y * s);
    if (lane == 0) dsum[row] = d * s;
  } else if (row - Mq < Mkv) {
    const long long r = row - Mq;
    const float4 v = reinterpret_cast<const float4*>(kv + r * 128)[lane];
    reinterpret_cast<uint2*>(kv16 + r * 128)[lane] = make_uint2(pack_h2s(v.x, v.y), pack_h2s(v.z, v.w));
  }
}
// ---- dq = (1 / s) * sum_j part[j] in tile order ----
__global__ void __launch_bounds__(256) fb_fold_dq_kernel(const float* __restrict__ part, int nkt, long long n4, const float* __restrict__ scale_ptr,
                                                         float* __restrict__ dq) {
  PDL_TOP();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float inv = 1.0f / *scale_ptr;
  float4 a = reinterpret_cast<const float4*>(part)[i];
  for (int j = 1; j < nkt; j++) {
    const float4 v = reinterpret_cast<const float4*>(part)[(long long)j * n4 + i];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  reinterpret_cast<float4*>(dq)[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

}  // namespace

size_t flash_bwd_workspace_floats(int B, int Nq, int Nk) {
  const size_t mq = (size_t)B * Nq, mkv = (size_t)B * Nk;
  const size_t nkt = (Nk + FB_T - 1) / FB_T;
  return 2 * (mq * FB_D / 2 + 64) + (mkv * 128 / 2 + 64) + (mq + 64) + nkt * mq * FB_D + 256;
}

// q [B][Nq][64], kv [B][Nk][128] (k | v), out = the forward output, lse [B][Nq] = the forward's row log2-sum-exp of q k^T * scale * log2 e,
// dout -> dq [B][Nq][64], dkv [B][Nk][128]
int launch_flash_bwd(const float* q, const float* kv, const float* out, const float* lse, const float* dout, float scale, float* dq,
                     float* dkv, int B, int Nq, int Nk, float* ws, cudaStream_t st) {
  TCX_REQUIRE(tcx_get_encode_tiled() != nullptr, "flash_bwd: cuTensorMapEncodeTiled entry point not available");
  const long long mq = (long long)B * Nq, mkv = (long long)B * Nk;
  const int nkt = (Nk + FB_T - 1) / FB_T;
  float* p = ws;
  auto take = [&](size_t n) { float* r = p; p += (n + 63) / 64 * 64; return r; };
  __half* q16 = reinterpret_cast<__half*>(take((size_t)mq * FB_D / 2 + 64));
  __half* do16 = reinterpret_cast<__half*>(take((size_t)mq * FB_D / 2 + 64));
  __half* kv16 = reinterpret_cast<__half*>(take((size_t)mkv * 128 / 2 + 64));
  float* dsum = take((size_t)mq + 64);
  float* small = take(64);            // [0] = absmax bits, [1] = scale
  float* part = take((size_t)nkt * mq * FB_D);
  TCX_REQUIRE(cudaMemsetAsync(small, 0, 8, st) == cudaSuccess, "flash_bwd: memset failed");
  tcx_launch_chain(fb_absmax_kernel, dim3(296), dim3(256), 0, st, dout, mq * FB_D / 4, reinterpret_cast<unsigned*>(small));
  TCX_TRY(tcx_check_launch("fb_absmax"));
  tcx_launch_chain(fb_prep_kernel, dim3((unsigned)((mq + mkv + 7) / 8)), dim3(256), 0, st, q, dout, out, kv, reinterpret_cast<const unsigned*>(small), q16, do16, kv16,
                                                                  dsum, small + 1, mq, mkv);
  TCX_TRY(tcx_check_launch("fb_prep"));
  FbMaps maps;
  TCX_TRY(tcx_make_operand_map(&maps.q, q16, 2, FB_D, Nq, FB_D, B, (long long)Nq * FB_D, 64, FB_T));
  TCX_TRY(tcx_make_operand_map(&maps.dout, do16, 2, FB_D, Nq, FB_D, B, (long long)Nq * FB_D, 64, FB_T));
  TCX_TRY(tcx_make_operand_map(&maps.kv, kv16, 2, 128, Nk, 128, B, (long long)Nk * 128, 64, FB_T));
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
    TCX_REQUIRE(e == cudaSuccess, "flash_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  {
    ProfScope prof("flash_bwd", st, 10.0 * B * (double)Nq * Nk * FB_D);     // five 2*Nq*Nk*64 contractions
    tcx_launch_chain(flash_bwd_kernel, dim3(B * nkt), dim3(FB_THREADS), FB_SMEM, st, maps, lse, dsum, small + 1, part, dkv, B, Nq, Nk, scale,
                                                           scale * 1.4426950408889634f);
    TCX_TRY(tcx_check_launch("flash_bwd"));
  }
  tcx_launch_chain(fb_fold_dq_kernel, dim3((unsigned)((mq * FB_D / 4 + 255) / 256)), dim3(256), 0, st, part, nkt, mq * FB_D / 4, small + 1, dq);
  return tcx_check_launch("fb_fold_dq");
}
