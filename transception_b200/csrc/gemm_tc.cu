// tcgen05 / TMA GEMM: C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 in HBM, TF32 tensor-core MMA, fp32 accumulate in TMEM.
//
//  * operands: both K-major (activations [M,K], nn.Linear weights [N,K]) -> TMA loads 128-byte-swizzled
//    [128 x 32] / [BN x 32] fp32 boxes straight into shared memory; no conversion pass, tf32 reads the fp32 bits.
//  * one CTA = one 128 x BN output tile; warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected
//    thread, tcgen05.mma kind::tf32, UMMA 128 x BN x 8), warps 2-5 = epilogue (tcgen05.ld -> smem transpose ->
//    coalesced bias / BatchNorm / activation / residual / store).
//  * grouped (blockIdx.z = group*batch + b): up to 4 independent problems of the same shape per launch — the three
//    Multi-Branch encoders run as one launch — plus a strided batch dimension through rank-3 tensor maps.
//  These GEMMs are HBM/L2-bound (K = 64..512, arithmetic intensity 30-250 FLOP/B), so the tile is sized for
//  2 CTAs/SM and the epilogue for fully coalesced 128-byte row segments rather than for peak MMA issue.
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

struct TmaSet {
  CUtensorMap a[TCX_MAX_GROUPS];
  CUtensorMap w[TCX_MAX_GROUPS];
};

template <int BN, int STAGES>
struct Smem {
  static constexpr int STAGE_B = BN * BK * 4;
  static constexpr int OFF_B = STAGES * STAGE_A;
  static constexpr int OFF_STAGE = OFF_B + STAGES * STAGE_B;        // 4 warps x [32][33] fp32
  static constexpr int OFF_BAR = OFF_STAGE + 4 * 32 * 33 * 4;
  static constexpr int TOTAL = OFF_BAR + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;                           // slack for 1024-byte alignment
};

__device__ __forceinline__ float epi_col(float v, float bias, float bscale, float bshift, bool has_bn, int act) {
  v += bias;
  if (has_bn) v = fmaf(v, bscale, bshift);
  return apply_act(v, act);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ TmaSet maps,
                                                      const __grid_constant__ GemmParams p) {
  using L = Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int gi = z / p.batch, bi = z - gi * p.batch;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nkb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&maps.a[gi]);
    tc::prefetch_tmap(&maps.w[gi]);
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, BN);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int wb = p.strideW ? bi : 0;
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&full[s], STAGE_A + L::STAGE_B);
        tc::tma_load_3d(smem + s * STAGE_A, &maps.a[gi], kb * BK, m0, p.strideA ? bi : 0, &full[s]);
        tc::tma_load_3d(smem + L::OFF_B + s * L::STAGE_B, &maps.w[gi], kb * BK, n0, wb, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc(2, BM, BN);
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::fence_after_sync();
        const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(smem + s * STAGE_A));
        const uint64_t bd = tc::umma_desc_sw128(tc::smem_u32(smem + L::OFF_B + s * L::STAGE_B));
#pragma unroll
        for (int k = 0; k < BK / 8; k++)   // UMMA_K = 8 tf32 = 32 bytes: advance the start address inside the atom
          tc::umma_tf32(tmem_acc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        tc::umma_commit(&empty[s]);        // frees the smem stage when these MMAs retire
      }
      tc::umma_commit(acc_full);
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), +32) -------------------------------------
    const int quarter = warp & 3;
    float* stage = reinterpret_cast<float*>(smem + L::OFF_STAGE) + quarter * (32 * 33);
    const GemmGroup& g = p.g[gi];
    const GemmEpi& e = g.epi;
    float* __restrict__ C = g.C + (long long)bi * p.strideC;
    const float* __restrict__ R = e.residual ? e.residual + (long long)bi * e.strideR : nullptr;
    const bool has_bn = e.bn.w != nullptr;
    tc::mbar_wait(acc_full, 0);
    tc::fence_after_sync();
    const int row0 = m0 + quarter * 32;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;
      uint32_t v[32];
      tc::tmem_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j++) stage[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int col = n0 + c0 + lane;
      if (col < p.N) {
        const float bias = e.bias ? e.bias[col] : 0.f;
        float bs = 1.f, bt = 0.f;
        if (has_bn) bn_fold(e.bn, col, bs, bt);
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
          const int m = row0 + r;
          if (m < p.M) {
            float x = epi_col(stage[r * 33 + lane], bias, bs, bt, has_bn, e.act);
            if (R) x += R[(long long)m * e.ldr + col];
            C[(long long)m * p.ldc + col] = x;
          }
        }
      }
      __syncwarp();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_acc, BN);
  }
}

template <int BN, int STAGES>
int launch_cfg(const TmaSet& maps, const GemmParams& p, cudaStream_t st) {
  using L = Smem<BN, STAGES>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN);
    TCX_REQUIRE(e == cudaSuccess, "gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    done = true;
  }
  dim3 grid(cdiv(p.M, BM), cdiv(p.N, BN), p.groups * p.batch);
  ProfScope prof("gemm_tc", st);
  gemm_tc_kernel<BN, STAGES><<<grid, 192, L::DYN, st>>>(maps, p);
  return tcx_check_launch("gemm_tc");
}

}  // namespace

bool gemm_tc_eligible(const GemmParams& p) {
  if (!tcx_flag_gemm_tc()) return false;
  if (p.N < 16 || p.K < 16 || p.M < 32) return false;
  if ((p.K | p.lda | p.ldw) & 3) return false;
  if ((p.strideA | p.strideW) & 3) return false;
  for (int i = 0; i < p.groups; i++)
    if (((uintptr_t)p.g[i].A | (uintptr_t)p.g[i].W) & 15) return false;
  return tcx_get_encode_tiled() != nullptr;
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  const long long tiles128 = (long long)cdiv(p.M, BM) * cdiv(p.N, 128) * p.groups * p.batch;
  const bool bn128 = (p.N % 128 == 0) && tiles128 >= 2 * 148;
  const int bn = bn128 ? 128 : 64;
  TmaSet maps;
  for (int i = 0; i < p.groups; i++) {
    TCX_TRY(tcx_make_operand_map(&maps.a[i], p.g[i].A, 4, p.K, p.M, p.lda, p.batch, p.strideA, BK, BM));
    TCX_TRY(tcx_make_operand_map(&maps.w[i], p.g[i].W, 4, p.K, p.N, p.ldw, p.batch, p.strideW, BK, bn));
  }
  for (int i = p.groups; i < TCX_MAX_GROUPS; i++) { maps.a[i] = maps.a[0]; maps.w[i] = maps.w[0]; }
  if (bn128) return launch_cfg<128, 2>(maps, p, st);
  return launch_cfg<64, 3>(maps, p, st);
}
