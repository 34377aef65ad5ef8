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
                         long long batch, long long batch_stride, int box_k, int box_rows) {
  tcx_encode_tiled_fn enc = tcx_get_encode_tiled();
  TCX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  if (batch_stride == 0 || batch < 1) { batch = 1; batch_stride = rows * ld; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)batch_stride * elem_bytes};
  if (strides[1] < strides[0]) strides[1] = strides[0];
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TCX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): K=%lld rows=%lld ld=%lld batch=%lld stride=%lld",
              (int)r, K, rows, ld, batch, batch_stride);
  return 0;
}

namespace {

constexpr int BM = 128, BK = 32;              // BK fp32 = 128 bytes = one swizzle row
constexpr int STAGE_A = BM * BK * 4;          // 16 KB
constexpr int SLAB = 32;                      // epilogue slab: 128 rows x 32 fp32 columns (16 KB, one SW128 box)
constexpr int SLAB_BYTES = BM * SLAB * 4;
constexpr int GT_THREADS = 192;

struct TmaSet {
  CUtensorMap a[TCX_MAX_GROUPS];
  CUtensorMap w[TCX_MAX_GROUPS];
  CUtensorMap c[TCX_MAX_GROUPS];
  CUtensorMap r[TCX_MAX_GROUPS];   // residual (valid only when the group has one)
};

template <int BN, int STAGES>
struct Smem {
  static constexpr int STAGE_B = BN * BK * 4;
  static constexpr int OFF_B = STAGES * STAGE_A;
  static constexpr int OFF_OUT = OFF_B + STAGES * STAGE_B;          // 2 output slabs
  static constexpr int OFF_RES = OFF_OUT + 2 * SLAB_BYTES;          // 2 residual slabs
  static constexpr int OFF_COL = OFF_RES + 2 * SLAB_BYTES;          // [2 tiles][scale BN | shift BN] fp32
  static constexpr int OFF_BAR = OFF_COL + 2 * 2 * BN * 4;
  static constexpr int NBAR = 2 * STAGES + 6;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;                           // slack for 1024-byte alignment
  static constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // 128 / 256 / 512
};

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
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// one 128 x 32 epilogue slab row: y = act(acc * scale + shift) (+ residual) -> swizzled smem.  The activation and the
// optional terms are template parameters so the 32-element body is branch-free straight-line code.
template <int ACT, bool SCALE, bool RES>
__device__ __forceinline__ void slab_row(const uint32_t (&v)[32], const float* __restrict__ scv,
                                         const float* __restrict__ shv, const uint8_t* __restrict__ rrow,
                                         uint8_t* __restrict__ orow, int sw) {
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const float4 sh = *reinterpret_cast<const float4*>(shv + c * 4);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (SCALE) sc = *reinterpret_cast<const float4*>(scv + c * 4);
    float x[4] = {__uint_as_float(v[c * 4 + 0]), __uint_as_float(v[c * 4 + 1]), __uint_as_float(v[c * 4 + 2]),
                  __uint_as_float(v[c * 4 + 3])};
    const float scs[4] = {sc.x, sc.y, sc.z, sc.w}, shs[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float y = SCALE ? fmaf(x[j], scs[j], shs[j]) : x[j] + shs[j];
      if (ACT == ACT_GELU) y = gelu_erf(y);
      else if (ACT == ACT_HARDSWISH) y = hardswish(y);
      else if (ACT == ACT_SIGMOID) y = sigmoidf_(y);
      else if (ACT == ACT_SILU_SWISH) y = silu_swish(y);
      x[j] = y;
    }
    const int phys = (c ^ sw) << 4;
    if (RES) {
      const float4 rv = *reinterpret_cast<const float4*>(rrow + phys);
      x[0] += rv.x; x[1] += rv.y; x[2] += rv.z; x[3] += rv.w;
    }
    *reinterpret_cast<float4*>(orow + phys) = make_float4(x[0], x[1], x[2], x[3]);
  }
}
template <int ACT>
__device__ __forceinline__ void slab_act(bool scale, bool res, const uint32_t (&v)[32], const float* scv, const float* shv,
                                         const uint8_t* rrow, uint8_t* orow, int sw) {
  if (scale) {
    if (res) slab_row<ACT, true, true>(v, scv, shv, rrow, orow, sw);
    else slab_row<ACT, true, false>(v, scv, shv, rrow, orow, sw);
  } else {
    if (res) slab_row<ACT, false, true>(v, scv, shv, rrow, orow, sw);
    else slab_row<ACT, false, false>(v, scv, shv, rrow, orow, sw);
  }
}
__device__ __forceinline__ void slab_dispatch(int act, bool scale, bool res, const uint32_t (&v)[32], const float* scv,
                                           const float* shv, const uint8_t* rrow, uint8_t* orow, int sw) {
  switch (act) {
    case ACT_GELU: slab_act<ACT_GELU>(scale, res, v, scv, shv, rrow, orow, sw); break;
    case ACT_HARDSWISH: slab_act<ACT_HARDSWISH>(scale, res, v, scv, shv, rrow, orow, sw); break;
    case ACT_SIGMOID: slab_act<ACT_SIGMOID>(scale, res, v, scv, shv, rrow, orow, sw); break;
    case ACT_SILU_SWISH: slab_act<ACT_SILU_SWISH>(scale, res, v, scv, shv, rrow, orow, sw); break;
    default: slab_act<ACT_NONE>(scale, res, v, scv, shv, rrow, orow, sw); break;
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TmaSet maps,
                                                                const __grid_constant__ GemmParams p) {
  using L = Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint64_t* res_full = acc_empty + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (p.M + BM - 1) / BM, nt = (p.N + BN - 1) / BN;
  const int ntiles = mt * nt * p.groups * p.batch;
  const int nkb = (p.K + BK - 1) / BK;

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
      tc::mbar_init(&acc_empty[a], 128);
      tc::mbar_init(&res_full[a], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, L::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: A / W k-blocks of every tile of this CTA, back to back =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(tile, mt, nt, p.batch, BN);
        const int za = p.strideA ? t.bi : 0, zw = p.strideW ? t.bi : 0;
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % STAGES;
          tc::mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&full[s], STAGE_A + L::STAGE_B);
          tc::tma_load_3d(smem + s * STAGE_A, &maps.a[t.gi], kb * BK, t.m0, za, &full[s]);
          tc::tma_load_3d(smem + L::OFF_B + s * L::STAGE_B, &maps.w[t.gi], kb * BK, t.n0, zw, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc(2, BM, BN);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
        const uint32_t a = ti & 1;
        tc::mbar_wait(&acc_empty[a], ((ti >> 1) & 1) ^ 1);     // epilogue drained this accumulator
        tc::fence_after_sync();
        const uint32_t acc = tmem_base + a * BN;
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % STAGES;
          tc::mbar_wait(&full[s], (it / STAGES) & 1);
          tc::fence_after_sync();
          const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(smem + s * STAGE_A));
          const uint64_t bd = tc::umma_desc_sw128(tc::smem_u32(smem + L::OFF_B + s * L::STAGE_B));
#pragma unroll
          for (int k = 0; k < BK / 8; k++)   // UMMA_K = 8 tf32 = 32 bytes: advance the start address inside the atom
            tc::umma_tf32(acc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          tc::umma_commit(&empty[s]);        // frees the smem stage when these MMAs retire
        }
        tc::umma_commit(&acc_full[a]);
      }
    }
  } else {
    // ================= epilogue warps 2..5: TMEM lanes [32*(warp%4), +32) =================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // row inside the tile == TMEM lane
    const int et = threadIdx.x - 64;                   // 0..127
    const bool leader = (et == 0);
    const int sw = r & 7;
    uint8_t* out_slab = smem + L::OFF_OUT;
    uint8_t* res_slab = smem + L::OFF_RES;
    float* colc = reinterpret_cast<float*>(smem + L::OFF_COL);
    constexpr int NSLAB = BN / SLAB;

    // residual prefetch cursor (leader only): slabs are numbered consecutively over this CTA's tiles
    int pf_tile = blockIdx.x, pf_slab = 0;
    uint32_t pf_count = 0;
    auto prefetch_res = [&]() {
      while (pf_tile < ntiles) {
        const TileCoord t = tile_coord(pf_tile, mt, nt, p.batch, BN);
        const bool live = p.g[t.gi].epi.residual != nullptr && t.n0 + pf_slab * SLAB < p.N;
        if (live) {
          const uint32_t b = pf_count & 1;
          tc::mbar_arrive_expect_tx(&res_full[b], SLAB_BYTES);
          tc::tma_load_3d(res_slab + b * SLAB_BYTES, &maps.r[t.gi], t.n0 + pf_slab * SLAB, t.m0,
                          p.g[t.gi].epi.strideR ? t.bi : 0, &res_full[b]);
          pf_count++;
        }
        if (++pf_slab == NSLAB) { pf_slab = 0; pf_tile += gridDim.x; }
        if (live) return;
      }
    };
    if (leader) prefetch_res();

    uint32_t ti = 0, slab_count = 0, res_count = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
      const TileCoord t = tile_coord(tile, mt, nt, p.batch, BN);
      const GemmEpi& e = p.g[t.gi].epi;
      const bool has_res = e.residual != nullptr;
      const bool has_scale = e.bn.w != nullptr;
      const int act = e.act;
      const uint32_t a = ti & 1;
      // per-tile column constants: y = act(acc * scale + shift) with bias and BatchNorm folded
      float* cs = colc + a * 2 * BN;
      for (int j = et; j < BN; j += 128) {
        const int col = t.n0 + j;
        float sc = 1.f, sh = 0.f;
        if (col < p.N) {
          const float bias = e.bias ? __ldg(e.bias + col) : 0.f;
          if (has_scale) {
            float bs, bt;
            bn_fold(e.bn, col, bs, bt);
            sc = bs;
            sh = fmaf(bias, bs, bt);
          } else {
            sh = bias;
          }
        }
        cs[j] = sc;
        cs[BN + j] = sh;
      }
      tc::mbar_wait(&acc_full[a], (ti >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t tacc = tmem_base + a * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int s = 0; s < NSLAB; s++) {
        if (t.n0 + s * SLAB >= p.N) break;                      // uniform: slab entirely past N
        uint32_t v[32];
        tc::tmem_ld32(tacc + s * SLAB, v);
        const uint32_t ob = slab_count & 1;
        if (leader) bulk_wait_read<1>();                        // the store that last read out_slab[ob] is done with smem
        if (has_res && leader) prefetch_res();                  // next live residual slab (this one is already in flight)
        epi_bar(1);                                             // out_slab[ob] free; column constants visible
        tc::tmem_ld_wait();
        if (s == NSLAB - 1 || t.n0 + (s + 1) * SLAB >= p.N) {   // last TMEM read of this tile: release the accumulator
          tc::fence_before_sync();
          tc::mbar_arrive(&acc_empty[a]);
        }
        const float* scv = cs + s * SLAB;
        const float* shv = cs + BN + s * SLAB;
        uint8_t* orow = out_slab + ob * SLAB_BYTES + r * 128;
        const uint8_t* rrow = res_slab + (res_count & 1) * SLAB_BYTES + r * 128;
        if (has_res) tc::mbar_wait(&res_full[res_count & 1], (res_count >> 1) & 1);
        slab_dispatch(act, has_scale, has_res, v, scv, shv, rrow, orow, sw);
        if (has_res) res_count++;
        tc::fence_proxy_async();
        epi_bar(2);                                             // slab complete (and residual slab consumed)
        if (leader) {
          tma_store_3d(&maps.c[t.gi], out_slab + ob * SLAB_BYTES, t.n0 + s * SLAB, t.m0, p.strideC ? t.bi : 0);
          bulk_commit();
        }
        slab_count++;
      }
    }
    if (leader) bulk_wait_all();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, L::TMEM_COLS);
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

template <int BN, int STAGES>
int launch_cfg(const TmaSet& maps, const GemmParams& p, cudaStream_t st) {
  using L = Smem<BN, STAGES>;
  static_assert(L::DYN <= 227 * 1024, "smem budget");
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN);
    TCX_REQUIRE(e == cudaSuccess, "gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    done = true;
  }
  const int ntiles = cdiv(p.M, BM) * cdiv(p.N, BN) * p.groups * p.batch;
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  ProfScope prof("gemm_tc", st);
  gemm_tc_kernel<BN, STAGES><<<grid, GT_THREADS, L::DYN, st>>>(maps, p);
  return tcx_check_launch("gemm_tc");
}

}  // namespace

bool gemm_tc_eligible(const GemmParams& p) {
  if (!tcx_flag_gemm_tc()) return false;
  if (p.N < 16 || p.K < 16 || p.M < 32) return false;
  if ((p.K | p.lda | p.ldw | p.ldc) & 3) return false;
  if ((p.strideA | p.strideW | p.strideC) & 3) return false;
  for (int i = 0; i < p.groups; i++) {
    if (((uintptr_t)p.g[i].A | (uintptr_t)p.g[i].W | (uintptr_t)p.g[i].C) & 15) return false;
    const GemmEpi& e = p.g[i].epi;
    if (e.residual && ((((uintptr_t)e.residual) & 15) || (e.ldr & 3) || (e.strideR & 3))) return false;
  }
  return tcx_get_encode_tiled() != nullptr;
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  const long long mtiles = (long long)cdiv(p.M, BM) * p.groups * p.batch;
  int bn = 64;
  if (p.N % 256 == 0 && mtiles * (p.N / 256) >= 2 * sm_count()) bn = 256;
  else if (p.N % 128 == 0 && mtiles * (p.N / 128) >= 2 * sm_count()) bn = 128;
  TmaSet maps;
  for (int i = 0; i < p.groups; i++) {
    const GemmEpi& e = p.g[i].epi;
    TCX_TRY(tcx_make_operand_map(&maps.a[i], p.g[i].A, 4, p.K, p.M, p.lda, p.batch, p.strideA, BK, BM));
    TCX_TRY(tcx_make_operand_map(&maps.w[i], p.g[i].W, 4, p.K, p.N, p.ldw, p.batch, p.strideW, BK, bn));
    TCX_TRY(tcx_make_operand_map(&maps.c[i], p.g[i].C, 4, p.N, p.M, p.ldc, p.batch, p.strideC, SLAB, BM));
    if (e.residual)
      TCX_TRY(tcx_make_operand_map(&maps.r[i], e.residual, 4, p.N, p.M, e.ldr, p.batch, e.strideR, SLAB, BM));
    else
      maps.r[i] = maps.c[i];
  }
  for (int i = p.groups; i < TCX_MAX_GROUPS; i++) {
    maps.a[i] = maps.a[0]; maps.w[i] = maps.w[0]; maps.c[i] = maps.c[0]; maps.r[i] = maps.r[0];
  }
  if (bn == 256) return launch_cfg<256, 3>(maps, p, st);
  if (bn == 128) return launch_cfg<128, 4>(maps, p, st);
  return launch_cfg<64, 6>(maps, p, st);
}
