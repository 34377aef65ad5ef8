// tcgen05 / TMA GEMM: C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 in HBM, TF32 tensor-core MMA, fp32 accumulate in TMEM.
//
// These GEMMs are skinny (K = 64..2048, N = 64..2048, M = tokens of the whole batch) and HBM/L2-bound, so the kernel
// is organised around keeping the memory system busy rather than around MMA issue:
//  * persistent CTAs (one per SM) walk a static tile list; 6 warps: warp 0 = TMA producer running ahead through a
//    multi-stage smem ring ACROSS tiles, warp 1 = MMA issuer (one elected thread, tcgen05.mma kind::tf32 reading the
//    fp32 bits straight from the 128-byte-swizzled TMA tiles), warps 2-5 = epilogue.
//  * two TMEM accumulators: the MMA of tile i+1 overlaps the epilogue of tile i.
//  * epilogue is row-per-thread (thread = TMEM lane): tcgen05.ld 32 columns -> per-column scale/shift (bias and
//    BatchNorm folded once per tile into smem) -> activation -> + residual -> swizzled smem slab -> TMA store.
//    Residual slabs are TMA-prefetched one slab ahead, so no thread ever waits on a global load, and every global
//    access of the kernel is a full-line TMA transaction; ragged M / N edges are clipped by the tensor maps.
//  * grouped (up to 4 independent problems of one shape, e.g. the three Multi-Branch encoders) and strided-batched
//    problems are folded into the tile list through rank-3 tensor maps.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc.cuh"

bool tcx_flag_gemm_tc();

// ---- host: driver entry point + tensor maps ----------------------------------------------------------
tcx_encode_tiled_fn tcx_get_encode_tiled() {
  static tcx_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tcx_encode_tiled_fn>(p);
  }
  return fn;
}

int tcx_make_operand_map(CUtensorMap* map, const void* base, int elem_bytes, long long K, long long rows, long long ld,
                         long long batch, long long batch_stride, int box_k, int box_rows, int bf16, int swizzle_base32) {
  tcx_encode_tiled_fn enc = tcx_get_encode_tiled();
  TCX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  if (batch_stride == 0 || batch < 1) { batch = 1; batch_stride = rows * ld; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)batch_stride * elem_bytes};
  if (strides[1] < strides[0]) strides[1] = strides[0];
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_base32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TCX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): K=%lld rows=%lld ld=%lld batch=%lld stride=%lld",
              (int)r, K, rows, ld, batch, batch_stride);
  return 0;
}

namespace {

constexpr int BM = 128;
constexpr int KB_BYTES = 128;                 // one k-block = one 128-byte swizzle row: 32 tf32 or 64 fp16 elements
constexpr int STAGE_A = BM * KB_BYTES;        // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int GT_THREADS = 64 + EPI_WARPS * 32;
constexpr int SUB_BYTES = 32 * 128;           // per-warp epilogue sub-slab: 32 rows x 128 bytes
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 227 * 1024;

struct TmaSet {
  CUtensorMap a[TCX_MAX_GROUPS];
  CUtensorMap w[TCX_MAX_GROUPS];
  CUtensorMap c[TCX_MAX_GROUPS];
  CUtensorMap r[TCX_MAX_GROUPS];   // residual (valid only when the group has one)
  CUtensorMap l[TCX_MAX_GROUPS];   // fp16 LayerNorm output (LN kernels only)
};

// dynamic shared memory plan (host-computed): [A stages][B stages][out 8x2x4K][res 8x2x4K]?[col 8x1K][barriers]
struct SmemPlan {
  int off_b, off_out, off_res, off_ln, off_col, off_bar, stages, total;
};
constexpr int COL_BYTES = 2048;               // per epilogue warp: [2][scale 64 | shift 64] + LN weight 64 + LN bias 64

struct TileCoord {
  int gi, bi, m0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int mt, int nt, int batch, int BN) {
  TileCoord t;
  const int ni = tile % nt;
  int rest = tile / nt;
  const int mi = rest % mt;
  rest /= mt;
  t.bi = rest % batch;
  t.gi = rest / batch;
  t.m0 = mi * BM;
  t.n0 = ni * BN;
  return t;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int ACT>
__device__ __forceinline__ float act_apply(float y) {
  if (ACT == ACT_GELU) return gelu_erf(y);
  if (ACT == ACT_HARDSWISH) return hardswish(y);
  if (ACT == ACT_SIGMOID) return sigmoidf_(y);
  if (ACT == ACT_SILU_SWISH) return silu_swish(y);
  return y;
}

// one epilogue row of a slab: y = act(acc * scale + shift) (+ residual) -> 128 swizzled bytes of shared memory.
// fp32 output: 32 columns; fp16 output: 64 columns.  All options are template parameters (branch-free body).
template <int ACT, bool SCALE, bool RES, bool OUT16, int NC>
__device__ __forceinline__ void slab_row(const uint32_t (&v)[NC], const float* __restrict__ scv,
                                         const float* __restrict__ shv, const uint8_t* __restrict__ rrow,
                                         uint8_t* __restrict__ orow, int sw) {
  static_assert(NC == (OUT16 ? 64 : 32), "slab width");
#pragma unroll
  for (int c = 0; c < 8; c++) {          // 16-byte output chunks
    constexpr int PER = OUT16 ? 8 : 4;
    float x[PER];
#pragma unroll
    for (int q = 0; q < PER / 4; q++) {
      const float4 sh = *reinterpret_cast<const float4*>(shv + c * PER + q * 4);
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if (SCALE) sc = *reinterpret_cast<const float4*>(scv + c * PER + q * 4);
      const float scs[4] = {sc.x, sc.y, sc.z, sc.w}, shs[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float a = __uint_as_float(v[c * PER + q * 4 + j]);
        x[q * 4 + j] = act_apply<ACT>(SCALE ? fmaf(a, scs[j], shs[j]) : a + shs[j]);
      }
    }
    const int phys = (c ^ sw) << 4;
    if (OUT16) {
      *reinterpret_cast<uint4*>(orow + phys) =
          make_uint4(pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[PER - 4], x[PER - 3]), pack_h2(x[PER - 2], x[PER - 1]));
    } else {
      if (RES) {
        const float4 rv = *reinterpret_cast<const float4*>(rrow + phys);
        x[0] += rv.x; x[1] += rv.y; x[2] += rv.z; x[3] += rv.w;
      }
      *reinterpret_cast<float4*>(orow + phys) = make_float4(x[0], x[1], x[2], x[3]);
    }
  }
}
template <int ACT, bool OUT16, int NC>
__device__ __forceinline__ void slab_act(bool scale, bool res, const uint32_t (&v)[NC], const float* scv, const float* shv,
                                         const uint8_t* rrow, uint8_t* orow, int sw) {
  if (scale) {
    if (res) slab_row<ACT, true, true, OUT16, NC>(v, scv, shv, rrow, orow, sw);
    else slab_row<ACT, true, false, OUT16, NC>(v, scv, shv, rrow, orow, sw);
  } else {
    if (res) slab_row<ACT, false, true, OUT16, NC>(v, scv, shv, rrow, orow, sw);
    else slab_row<ACT, false, false, OUT16, NC>(v, scv, shv, rrow, orow, sw);
  }
}
// FULL: every activation / BatchNorm-fold / residual combination (fp32-operand kernels).  !FULL: bias (+ residual for
// fp32 output) only — the fp16-operand kernels are used for plain nn.Linear sites.
template <bool FULL, bool OUT16, int NC>
__device__ __forceinline__ void slab_dispatch(int act, bool scale, bool res, const uint32_t (&v)[NC], const float* scv,
                                              const float* shv, const uint8_t* rrow, uint8_t* orow, int sw) {
  if constexpr (OUT16) {
    slab_row<ACT_NONE, false, false, true, NC>(v, scv, shv, rrow, orow, sw);
  } else if constexpr (!FULL) {
    slab_act<ACT_NONE, false, NC>(scale, res, v, scv, shv, rrow, orow, sw);
  } else {
    switch (act) {
      case ACT_GELU: slab_act<ACT_GELU, false, NC>(scale, res, v, scv, shv, rrow, orow, sw); break;
      case ACT_HARDSWISH: slab_act<ACT_HARDSWISH, false, NC>(scale, res, v, scv, shv, rrow, orow, sw); break;
      case ACT_SIGMOID: slab_act<ACT_SIGMOID, false, NC>(scale, res, v, scv, shv, rrow, orow, sw); break;
      case ACT_SILU_SWISH: slab_act<ACT_SILU_SWISH, false, NC>(scale, res, v, scv, shv, rrow, orow, sw); break;
      default: slab_act<ACT_NONE, false, NC>(scale, res, v, scv, shv, rrow, orow, sw); break;
    }
  }
}

template <int O, int N>
__device__ __forceinline__ void ld32_at(uint32_t taddr, uint32_t (&s)[N]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(s[O + 0]), "=r"(s[O + 1]), "=r"(s[O + 2]), "=r"(s[O + 3]), "=r"(s[O + 4]), "=r"(s[O + 5]), "=r"(s[O + 6]), "=r"(s[O + 7]), "=r"(s[O + 8]), "=r"(s[O + 9]), "=r"(s[O + 10]), "=r"(s[O + 11]), "=r"(s[O + 12]), "=r"(s[O + 13]), "=r"(s[O + 14]), "=r"(s[O + 15]), "=r"(s[O + 16]), "=r"(s[O + 17]), "=r"(s[O + 18]), "=r"(s[O + 19]), "=r"(s[O + 20]), "=r"(s[O + 21]), "=r"(s[O + 22]), "=r"(s[O + 23]), "=r"(s[O + 24]), "=r"(s[O + 25]), "=r"(s[O + 26]), "=r"(s[O + 27]), "=r"(s[O + 28]), "=r"(s[O + 29]), "=r"(s[O + 30]), "=r"(s[O + 31])
               : "r"(taddr)
               : "memory");
}

template <int BN, bool AB16, bool OUT16, bool LN>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TmaSet maps,
                                                                const __grid_constant__ GemmParams p,
                                                                const __grid_constant__ SmemPlan sp) {
  constexpr int STAGE_B = BN * KB_BYTES;
  constexpr int KBE = AB16 ? 64 : 32;                 // elements per k-block
  constexpr int SLABC = OUT16 ? 64 : 32;              // output columns per 128-byte slab row
  constexpr int NSLAB = BN / SLABC;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  // ALT: a tile is finished by ONE epilogue warp set (set index == accumulator index) and the two sets take alternate
  // tiles — used when a thread needs its whole row (LN) and when a tile has a single slab (otherwise one set would idle)
  constexpr bool ALT = LN || NSLAB == 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int STAGES = sp.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + sp.off_bar);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* res_full = acc_empty + 2;        // [EPI_WARPS][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2 * EPI_WARPS);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (p.M + BM - 1) / BM, nt = (p.N + BN - 1) / BN;
  const int ntiles = mt * nt * p.groups * p.batch;
  const int nkb = (p.K + KBE - 1) / KBE;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.groups; g++) {
      tc::prefetch_tmap(&maps.a[g]);
      tc::prefetch_tmap(&maps.w[g]);
      tc::prefetch_tmap(&maps.c[g]);
      if (p.g[g].epi.residual) tc::prefetch_tmap(&maps.r[g]);
    }
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; a++) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], ALT ? EPI_WARPS * 16 : EPI_WARPS * 32);   // ALT: one warp set per accumulator
    }
    for (int i = 0; i < 2 * EPI_WARPS; i++) tc::mbar_init(&res_full[i], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // everything above overlapped the previous kernel; no global access before this point

  if (warp == 0) {
    // ================= TMA producer: A / W k-blocks of every tile of this CTA, back to back =================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(tile, mt, nt, p.batch, BN);
        const int za = p.strideA ? t.bi : 0, zw = p.strideW ? t.bi : 0;
        for (int kb = 0; kb < nkb; kb++) {
          tc::mbar_wait(&empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&full[s], STAGE_A + STAGE_B);
          tc::tma_load_3d(smem + s * STAGE_A, &maps.a[t.gi], kb * KBE, t.m0, za, &full[s]);
          if (!p.w_mn) {
            tc::tma_load_3d(smem + sp.off_b + s * STAGE_B, &maps.w[t.gi], kb * KBE, t.n0, zw, &full[s]);
          } else {
            // MN-major W: boxes of {KBE columns of N (128 bytes), KBE rows of K}; out-of-range boxes are zero-filled
#pragma unroll
            for (int i = 0; i < BN / KBE; i++)
              tc::tma_load_3d(smem + sp.off_b + s * STAGE_B + i * (KBE * KB_BYTES), &maps.w[t.gi], t.n0 + i * KBE, kb * KBE, zw,
                              &full[s]);
          }
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const bool wmn = p.w_mn != 0;
      const uint32_t idesc = tc::umma_idesc(AB16 ? (p.bf16 ? 1 : 0) : 2, BM, BN, 0, wmn ? 1 : 0);
      // K advance of the B descriptor per MMA (16-byte units): 32 bytes inside the swizzle row when K-major; one MMA's worth of
      // 128-byte k rows (16 fp16 / 8 tf32) when MN-major
      const uint64_t bstep = wmn ? (uint64_t)((AB16 ? 16 : 8) * KB_BYTES >> 4) : 2;
      uint32_t s = 0, ph = 0, ti = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
        const uint32_t a = ti & 1;
        tc::mbar_wait(&acc_empty[a], ((ti >> 1) & 1) ^ 1);     // epilogue drained this accumulator
        tc::fence_after_sync();
        const uint32_t acc = tmem_base + a * BN;
        for (int kb = 0; kb < nkb; kb++) {
          tc::mbar_wait(&full[s], ph);
          tc::fence_after_sync();
          const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(smem + s * STAGE_A));
          const uint32_t baddr = tc::smem_u32(smem + sp.off_b + s * STAGE_B);
          const uint64_t bd = !wmn ? tc::umma_desc_sw128(baddr)
                                   : (AB16 ? tc::umma_desc_mn_sw128(baddr, KBE * KB_BYTES) : tc::umma_desc_mn_sw128_base32(baddr, KBE * KB_BYTES));
#pragma unroll
          for (int k = 0; k < 4; k++) {     // one MMA consumes 32 bytes of K: advance the start address inside the atom
            if (AB16) tc::umma_f16(acc, ad + (uint64_t)(k * 2), bd + (uint64_t)k * bstep, idesc, (kb | k) != 0);
            else tc::umma_tf32(acc, ad + (uint64_t)(k * 2), bd + (uint64_t)k * bstep, idesc, (kb | k) != 0);
          }
          tc::umma_commit(&empty[s]);        // frees the smem stage when these MMAs retire
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
        }
        tc::umma_commit(&acc_full[a]);
      }
    }
  } else {
    // ================= epilogue warps 2..9: independent pipelines, no CTA-wide barriers =================
    // warp -> TMEM lane quarter (hardware rule: warp % 4) and column half; each warp owns 32 rows x its slabs,
    // stages them in its own swizzled 4 KB buffers and issues its own TMA stores / residual prefetches.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int sw = lane & 7;
    uint8_t* out_w = smem + sp.off_out + ew * 2 * SUB_BYTES;
    uint8_t* res_w = smem + sp.off_res + ew * 2 * SUB_BYTES;
    float* col_w = reinterpret_cast<float*>(smem + sp.off_col + ew * COL_BYTES);   // [2][scale 64 | shift 64] | LN w, b
    uint8_t* ln_w_buf = smem + sp.off_ln + ew * SUB_BYTES;
    // LN kernels: the thread needs its whole 64-column row, so a tile is finished by ONE warp set (half == accumulator
    // index) instead of both sets splitting the slabs; the two sets then work on alternate tiles
    constexpr int S0_STEP = ALT ? 1 : 2;
    uint64_t* rbar = res_full + ew * 2;

    int pf_tile = blockIdx.x + (ALT ? half * (int)gridDim.x : 0), pf_slab = ALT ? 0 : half;
    const int pf_tile_step = ALT ? 2 * (int)gridDim.x : (int)gridDim.x;
    uint32_t pf_count = 0;
    auto prefetch_res = [&]() {      // lane 0: next live residual sub-slab of this warp
      while (pf_tile < ntiles && (ALT || half < NSLAB)) {
        const TileCoord t = tile_coord(pf_tile, mt, nt, p.batch, BN);
        const bool live = p.g[t.gi].epi.residual != nullptr && t.n0 + pf_slab * SLABC < p.N;
        if (live) {
          const uint32_t b = pf_count & 1;
          tc::mbar_arrive_expect_tx(&rbar[b], SUB_BYTES);
          tc::tma_load_3d(res_w + b * SUB_BYTES, &maps.r[t.gi], t.n0 + pf_slab * SLABC, t.m0 + quarter * 32,
                          p.g[t.gi].epi.strideR ? t.bi : 0, &rbar[b]);
          pf_count++;
        }
        pf_slab += S0_STEP;
        if (pf_slab >= NSLAB) { pf_slab = ALT ? 0 : half; pf_tile += pf_tile_step; }
        if (live) return;
      }
    };
    if (lane == 0 && !OUT16) prefetch_res();

    uint32_t ti = 0, out_count = 0, res_count = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
      if (ALT && (int)(ti & 1) != half) continue;
      const TileCoord t = tile_coord(tile, mt, nt, p.batch, BN);
      const GemmEpi& e = p.g[t.gi].epi;
      const bool has_res = !OUT16 && e.residual != nullptr;
      const bool has_scale = e.bn.w != nullptr || p.alpha != 0.f;
      const int act = e.act;
      const uint32_t a = ti & 1;
      tc::mbar_wait(&acc_full[a], (ti >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t tacc = tmem_base + a * BN + ((uint32_t)(quarter * 32) << 16);
      bool arrived = false;
      if constexpr (LN) {
        static_assert(!LN || (BN == 64 && !OUT16), "LN epilogue: BN = 64, fp32 C");
        float yv[BN];
        float* lnc = col_w + 256;                       // LN weight | bias of this tile's group
        lnc[lane] = __ldg(e.ln_w + lane); lnc[lane + 32] = __ldg(e.ln_w + lane + 32);
        lnc[64 + lane] = __ldg(e.ln_b + lane); lnc[96 + lane] = __ldg(e.ln_b + lane + 32);
#pragma unroll
        for (int s = 0; s < NSLAB; s++) {
          const int col0 = t.n0 + s * SLABC;
          const uint32_t ob = out_count & 1;
          uint32_t v[SLABC];
          ld32_at<0>(tacc + s * SLABC, v);
          float* cs = col_w + ob * 128;
          cs[64 + lane] = (e.bias && col0 + lane < p.N) ? __ldg(e.bias + col0 + lane) : 0.f;
          if (lane == 0) {
            bulk_wait_read<2>();                                 // groups per tile: slab, slab, LN (three buffers)
            if (has_res) prefetch_res();
          }
          __syncwarp();
          tc::tmem_ld_wait();
          if (s == NSLAB - 1) {
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_empty[a]);
            arrived = true;
          }
          uint8_t* orow = out_w + ob * SUB_BYTES + lane * 128;
          const uint8_t* rrow = res_w + (res_count & 1) * SUB_BYTES + lane * 128;
          if (has_res) tc::mbar_wait(&rbar[res_count & 1], (res_count >> 1) & 1);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float4 sh = *reinterpret_cast<const float4*>(cs + 64 + c * 4);
            const int phys = (c ^ sw) << 4;
            float4 y = make_float4(__uint_as_float(v[c * 4]) + sh.x, __uint_as_float(v[c * 4 + 1]) + sh.y,
                                   __uint_as_float(v[c * 4 + 2]) + sh.z, __uint_as_float(v[c * 4 + 3]) + sh.w);
            if (has_res) {
              const float4 rv = *reinterpret_cast<const float4*>(rrow + phys);
              y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
            }
            yv[s * 32 + c * 4] = y.x; yv[s * 32 + c * 4 + 1] = y.y; yv[s * 32 + c * 4 + 2] = y.z; yv[s * 32 + c * 4 + 3] = y.w;
            *reinterpret_cast<float4*>(orow + phys) = y;
          }
          if (has_res) res_count++;
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&maps.c[t.gi], out_w + ob * SUB_BYTES, col0, t.m0 + quarter * 32, p.strideC ? t.bi : 0);
            bulk_commit();
          }
          out_count++;
        }
        // LayerNorm of the finished 64-column row (exact two-pass statistics in registers) -> fp16 slab
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < BN; i += 4) { s0 += yv[i]; s1 += yv[i + 1]; s2 += yv[i + 2]; s3 += yv[i + 3]; }
        const float mean = ((s0 + s1) + (s2 + s3)) * (1.f / BN);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int i = 0; i < BN; i += 4) {
          const float d0 = yv[i] - mean, d1 = yv[i + 1] - mean, d2 = yv[i + 2] - mean, d3 = yv[i + 3] - mean;
          q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
        }
        const float rstd = rsqrtf(((q0 + q1) + (q2 + q3)) * (1.f / BN) + e.ln_eps);
        if (lane == 0) bulk_wait_read<2>();                      // the previous tile's LN store has left ln_w_buf
        __syncwarp();
        uint8_t* lrow = ln_w_buf + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float4 w0 = *reinterpret_cast<const float4*>(lnc + c * 8), w1 = *reinterpret_cast<const float4*>(lnc + c * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(lnc + 64 + c * 8), b1 = *reinterpret_cast<const float4*>(lnc + 64 + c * 8 + 4);
          *reinterpret_cast<uint4*>(lrow + ((c ^ sw) << 4)) = make_uint4(
              pack_h2(fmaf((yv[c * 8 + 0] - mean) * rstd, w0.x, b0.x), fmaf((yv[c * 8 + 1] - mean) * rstd, w0.y, b0.y)),
              pack_h2(fmaf((yv[c * 8 + 2] - mean) * rstd, w0.z, b0.z), fmaf((yv[c * 8 + 3] - mean) * rstd, w0.w, b0.w)),
              pack_h2(fmaf((yv[c * 8 + 4] - mean) * rstd, w1.x, b1.x), fmaf((yv[c * 8 + 5] - mean) * rstd, w1.y, b1.y)),
              pack_h2(fmaf((yv[c * 8 + 6] - mean) * rstd, w1.z, b1.z), fmaf((yv[c * 8 + 7] - mean) * rstd, w1.w, b1.w)));
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&maps.l[t.gi], ln_w_buf, t.n0, t.m0 + quarter * 32, e.stride_ln ? t.bi : 0);
          bulk_commit();
        }
      } else {
#pragma unroll 1
      for (int s = ALT ? 0 : half; s < NSLAB; s += S0_STEP) {
        const int col0 = t.n0 + s * SLABC;
        if (col0 >= p.N) break;
        const uint32_t ob = out_count & 1;
        uint32_t v[SLABC];
        ld32_at<0>(tacc + s * SLABC, v);
        if constexpr (OUT16) ld32_at<SLABC - 32>(tacc + s * SLABC + 32, v);
        // per-slab column constants y = act(acc * scale + shift): bias and BatchNorm folded, one column per lane
        float* cs = col_w + ob * 128;
#pragma unroll
        for (int j = lane; j < SLABC; j += 32) {
          const int col = col0 + j;
          float sc = 1.f, sh = 0.f;
          if (col < p.N) {
            const float bias = e.bias ? __ldg(e.bias + col) : 0.f;
            if (e.bn.w != nullptr) {
              float bs, bt;
              bn_fold(e.bn, col, bs, bt);
              sc = bs;
              sh = fmaf(bias, bs, bt);
            } else if (has_scale) {
              sc = p.alpha;
              sh = bias;
            } else {
              sh = bias;
            }
          }
          cs[j] = sc;
          cs[64 + j] = sh;
        }
        if (lane == 0) {
          bulk_wait_read<1>();                                   // the store that last read out buffer `ob` is done with it
          if (has_res) prefetch_res();                           // next live residual (the current one is already in flight)
        }
        __syncwarp();
        tc::tmem_ld_wait();
        if (s + S0_STEP >= NSLAB || col0 + S0_STEP * SLABC >= p.N) {   // last TMEM read of this tile by this warp
          tc::fence_before_sync();
          tc::mbar_arrive(&acc_empty[a]);
          arrived = true;
        }
        uint8_t* orow = out_w + ob * SUB_BYTES + lane * 128;
        const uint8_t* rrow = res_w + (res_count & 1) * SUB_BYTES + lane * 128;
        if (has_res) tc::mbar_wait(&rbar[res_count & 1], (res_count >> 1) & 1);
        slab_dispatch<!AB16, OUT16, SLABC>(act, has_scale, has_res, v, cs, cs + 64, rrow, orow, sw);
        if (has_res) res_count++;
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&maps.c[t.gi], out_w + ob * SUB_BYTES, col0, t.m0 + quarter * 32, p.strideC ? t.bi : 0);
          bulk_commit();
        }
        out_count++;
      }
      }
      if (!arrived) {
        tc::fence_before_sync();
        tc::mbar_arrive(&acc_empty[a]);
      }
    }
    if (lane == 0) bulk_wait_read<0>();   // smem may be released; the stores themselves complete before the grid does
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int g_sm_count = 0;
int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

SmemPlan make_plan(int bn, bool has_res, bool ln) {
  SmemPlan sp{};
  const int stage = STAGE_A + bn * KB_BYTES;
  const int sub = EPI_WARPS * 2 * SUB_BYTES;
  const int fixed = sub + (has_res ? sub : 0) + (ln ? EPI_WARPS * SUB_BYTES : 0) + EPI_WARPS * COL_BYTES + 512;
  int budget = SMEM_BUDGET;
  if (g_tcx_smem_kb > 0 && bn <= 128 && g_tcx_smem_kb * 1024 < budget) budget = g_tcx_smem_kb * 1024;   // BN = 256 owns all of TMEM anyway
  int stages = (budget - 1024 - fixed) / stage;
  if (stages < 2) stages = 2;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  sp.stages = stages;
  sp.off_b = stages * STAGE_A;
  sp.off_out = sp.off_b + stages * bn * KB_BYTES;
  sp.off_res = sp.off_out + sub;
  sp.off_ln = sp.off_res + (has_res ? sub : 0);
  sp.off_col = sp.off_ln + (ln ? EPI_WARPS * SUB_BYTES : 0);
  sp.off_bar = sp.off_col + EPI_WARPS * COL_BYTES;
  sp.total = sp.off_bar + 512 + 1024;   // + slack for the 1024-byte alignment of the dynamic base
  return sp;
}

template <int BN, bool AB16, bool OUT16, bool LN = false>
int launch_cfg(const TmaSet& maps, const GemmParams& p, const SmemPlan& sp, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, AB16, OUT16, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BUDGET);
    TCX_REQUIRE(e == cudaSuccess, "gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  TCX_REQUIRE(sp.stages >= 2 && sp.total <= SMEM_BUDGET, "gemm_tc: smem plan does not fit (BN=%d stages=%d)", BN, sp.stages);
  const int ntiles = cdiv(p.M, BM) * cdiv(p.N, BN) * p.groups * p.batch;
  int grid = ntiles < sm_count() ? ntiles : sm_count();
  if (g_tcx_max_ctas > 0 && grid > g_tcx_max_ctas) grid = g_tcx_max_ctas;
  // algorithmic bytes: every operand read once, the output written once
  const double ae = AB16 ? 2.0 : 4.0, ce = OUT16 ? 2.0 : 4.0;
  double bytes = 0.0;
  for (int i = 0; i < p.groups; i++) {
    const double nb = (double)p.batch;
    bytes += nb * p.M * p.K * ae + (p.strideW ? nb : 1.0) * p.N * p.K * ae + nb * p.M * p.N * ce +
             (p.g[i].epi.residual ? nb * p.M * p.N * 4.0 : 0.0) + (LN ? nb * p.M * p.N * 2.0 : 0.0);
  }
  ProfScope prof("gemm_tc", st, bytes);
  cudaError_t le = tcx_launch_pdl(gemm_tc_kernel<BN, AB16, OUT16, LN>, dim3(grid), dim3(GT_THREADS), (size_t)sp.total, st, maps, p, sp);
  TCX_REQUIRE(le == cudaSuccess, "gemm_tc: launch failed: %s", cudaGetErrorString(le));
  return tcx_check_launch("gemm_tc");
}

template <bool AB16, bool OUT16>
int launch_bn(int bn, const TmaSet& maps, const GemmParams& p, const SmemPlan& sp, cudaStream_t st) {
  if (bn == 256) return launch_cfg<256, AB16, OUT16>(maps, p, sp, st);
  if (bn == 128) return launch_cfg<128, AB16, OUT16>(maps, p, sp, st);
  return launch_cfg<64, AB16, OUT16>(maps, p, sp, st);
}

}  // namespace

bool gemm_tc_eligible(const GemmParams& p) {
  if (!tcx_flag_gemm_tc()) return false;
  if (p.N < 16 || p.K < 16 || p.M < 1) return false;   // ragged / tiny M is clipped by the tensor maps
  const int am = p.ab16 ? 7 : 3, cm = p.out16 ? 7 : 3;     // 16-byte row pitches
  if ((p.K | p.lda | p.ldw) & am) return false;
  if (p.w_mn && ((p.N & am) || p.alpha < 0.f)) return false;
  if (p.alpha != 0.f && (p.out16 || p.alpha < 0.f)) return false;
  if (p.ldc & cm) return false;
  if ((p.strideA | p.strideW) & am) return false;
  if (p.strideC & cm) return false;
  for (int i = 0; i < p.groups; i++) {
    if (((uintptr_t)p.g[i].A | (uintptr_t)p.g[i].W | (uintptr_t)p.g[i].C) & 15) return false;
    const GemmEpi& e = p.g[i].epi;
    if (e.residual && ((((uintptr_t)e.residual) & 15) || (e.ldr & 3) || (e.strideR & 3))) return false;
    if (p.out16 && (e.residual || e.bn.w || e.act != ACT_NONE)) return false;
    if (e.ln_out && (p.w_mn || p.alpha != 0.f)) return false;
    if (p.alpha != 0.f && e.bn.w) return false;
    if (e.ln_out && (!p.ab16 || p.out16 || (p.N & 63) || (e.ld_ln & 7) || (e.stride_ln & 7) || (((uintptr_t)e.ln_out) & 15) ||
                     !e.ln_w || !e.ln_b)) return false;
    if (p.ab16 && (e.bn.w || e.act != ACT_NONE)) return false;
  }
  return tcx_get_encode_tiled() != nullptr;
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  const long long mtiles = (long long)cdiv(p.M, BM) * p.groups * p.batch;
  bool has_res = false;
  for (int i = 0; i < p.groups; i++) has_res |= p.g[i].epi.residual != nullptr;
  bool ln = false;
  for (int i = 0; i < p.groups; i++) ln |= p.g[i].epi.ln_out != nullptr;
  if (ln)
    for (int i = 0; i < p.groups; i++) TCX_REQUIRE(p.g[i].epi.ln_out != nullptr, "gemm_tc: LN output must be set for every group");
  int bn = 64;
  if (ln) bn = 64;            // one 64-column LayerNorm group per tile
  else if (!has_res && p.N % 256 == 0 && mtiles * (p.N / 256) >= 2 * sm_count()) bn = 256;
  else if (p.N % 128 == 0 && mtiles * (p.N / 128) >= 2 * sm_count()) bn = 128;
  const int ae = p.ab16 ? 2 : 4, ce = p.out16 ? 2 : 4;
  const int kbe = KB_BYTES / ae, slabc = 128 / ce;
  TmaSet maps;
  for (int i = 0; i < p.groups; i++) {
    const GemmEpi& e = p.g[i].epi;
    TCX_TRY(tcx_make_operand_map(&maps.a[i], p.g[i].A, ae, p.K, p.M, p.lda, p.batch, p.strideA, kbe, BM));
    if (p.w_mn)   // storage [K][N]: inner dimension = N, rows = K; box = one 128-byte group of columns x one k-block of rows
      TCX_TRY(tcx_make_operand_map(&maps.w[i], p.g[i].W, ae, p.N, p.K, p.ldw, p.batch, p.strideW, kbe, kbe, p.bf16, ae == 4));
    else
      TCX_TRY(tcx_make_operand_map(&maps.w[i], p.g[i].W, ae, p.K, p.N, p.ldw, p.batch, p.strideW, kbe, bn));
    TCX_TRY(tcx_make_operand_map(&maps.c[i], p.g[i].C, ce, p.N, p.M, p.ldc, p.batch, p.strideC, slabc, 32));
    if (e.residual)
      TCX_TRY(tcx_make_operand_map(&maps.r[i], e.residual, 4, p.N, p.M, e.ldr, p.batch, e.strideR, 32, 32));
    else
      maps.r[i] = maps.c[i];
    if (e.ln_out)
      TCX_TRY(tcx_make_operand_map(&maps.l[i], e.ln_out, 2, p.N, p.M, e.ld_ln, p.batch, e.stride_ln, 64, 32));
    else
      maps.l[i] = maps.c[i];
  }
  for (int i = p.groups; i < TCX_MAX_GROUPS; i++) {
    maps.a[i] = maps.a[0]; maps.w[i] = maps.w[0]; maps.c[i] = maps.c[0]; maps.r[i] = maps.r[0]; maps.l[i] = maps.l[0];
  }
  const SmemPlan sp = make_plan(bn, has_res, ln);
  if (ln) return launch_cfg<64, true, false, true>(maps, p, sp, st);
  if (p.ab16) return p.out16 ? launch_bn<true, true>(bn, maps, p, sp, st) : launch_bn<true, false>(bn, maps, p, sp, st);
  TCX_REQUIRE(!p.out16, "gemm_tc: fp16 output needs fp16 operands");
  return launch_bn<false, false>(bn, maps, p, sp, st);
}
