// Batched "NT" product on the tensor cores with fp32-equivalent accuracy (3xTF32 split), TMA in and TMA out:
//
//     C_z (=  |  -=)  A_z . B_z^T          A_z (M x K), B_z (N x K), C_z (M x N), all row-major fp32, z = 0 .. batch-1
//
// Used for (a) the deferred K = 256 trailing updates of the long-series factorisation and inverse sweep (chol_large.cu:
// C -= A B^T on sub-rectangles of the scratch, only the tiles that touch the lower triangle in the Cholesky sweep) and
// (b) the two products of the GPCV stage's K^-1 L_S = U (U^T L_S) ([GPyTorch] kl_mvn_mvn inside VariationalELBO,
// voltron/train_utils.py:46-56), which round 1 left to cuBLAS.
//
// One CTA per SM, persistent over 128 x BN output tiles (BN = 128 or 256), 320 threads with the three roles of the batched
// kernel's control-warp instance (chol_tc_dev.cuh):
//     warp 9  TMA producer   per k-tile t of 16 floats: wait done(t-4) -> cp.async.bulk.tensor.3d x2 (A 128 x 64 B, B BN x 64 B,
//                            SWIZZLE_64B) -> full(t)
//     warps 0-7 workers      wait full(t) -> A: own 32 B -> hi = raw / lo = a - trunc(a) -> tcgen05.st (TMEM stage t % 4);
//                            B: lo tile in shared memory (the raw tile is the hi operand)                         -> ready(t)
//     warp 8  MMA issuer     wait ready(t) -> 6 x tcgen05.mma.kind::tf32 (TS form, M128 x N BN x K8) -> commit    -> done(t)
// Epilogue (workers): accumulator -> registers (tcgen05.ld, 32 columns at a time) -> negated for "-=" -> 128 x 32 staging
// tile in shared memory (SWIZZLE_128B) -> one elected thread issues a TMA store (cp.async.bulk.tensor, "=") or a TMA
// reduction (cp.reduce.async.bulk.tensor .add, "-="): the read-modify-write of C happens in the L2, ragged edges are
// clipped by the tensor map, and no thread ever loads C.
#include <cuda.h>

#include <cstring>

#include "chol_tc_dev.cuh"

namespace volt {
namespace gnt {

using namespace tc;

constexpr int G_THREADS = NT + 64;
constexpr int G_RING = 4;
constexpr uint32_t G_A_TILE = 128u * 64u;                    // A k-tile: 128 rows x 16 floats
constexpr uint32_t G_STAGE = 128u * 128u;                    // output staging piece: 128 rows x 32 floats

struct GemmParams {
  int M, N, K, batch;
  int tri;        // 1: skip 64-column blocks that lie entirely above the diagonal of the (sub-)matrix (rows / columns share an origin)
  int mode;       // 0: C = A B^T   1: C -= A B^T
};

template <int BN>
struct Lay {
  static constexpr uint32_t B_TILE_ = (uint32_t)BN * 64u;                 // B k-tile: BN rows x 16 floats
  static constexpr uint32_t SLOT = G_A_TILE + B_TILE_;
  static constexpr uint32_t BL = G_RING * SLOT;                           // B lo tiles
  static constexpr uint32_t STG = BL + G_RING * B_TILE_;                  // two staging pieces (one per column half)
  static constexpr uint32_t BARS = STG + 2 * G_STAGE;
  static constexpr uint32_t BYTES = BARS + 256;
  static constexpr uint32_t ACC = 0, AHI = BN, ALO = BN + 64;             // TMEM columns: accumulator | 4 x 16 hi | 4 x 16 lo
  static constexpr uint32_t TCOLS = BN == 256 ? 512 : 256;
  static constexpr uint32_t IDESC_ = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)BN >> 3) << 17) | ((128u >> 4) << 24);
};

__device__ __forceinline__ void tma_load_3d(const void* tmap, uint32_t dst, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
template <int BN>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(Lay<BN>::IDESC_), "r"(accumulate)
      : "memory");
}

// tile t of the launch -> (batch, row block, column block); false when the tile is skipped (entirely above the diagonal)
__device__ __forceinline__ bool tile_of(const GemmParams& p, int BN, int t, int& z, int& r_base, int& c_base) {
  const int ncol = (p.N + BN - 1) / BN, nrow = (p.M + 127) / 128;
  z = t / (ncol * nrow);
  const int u = t - z * ncol * nrow;
  r_base = 128 * (u / ncol);
  c_base = BN * (u % ncol);
  return !(p.tri && c_base > r_base + 127);
}

template <int BN>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_nt_kernel(GemmParams p, const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC) {
  using L = Lay<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((s_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BARS);
  uint64_t* done = full + G_RING;
  uint64_t* ready = done + G_RING;
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(ready + G_RING);
  const int tid = threadIdx.x;
  const int wu = uniform_warp_id();
  if (wu == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(s_tmem_p)), "n"(L::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < G_RING; ++i) { mbar_init(full + i, 1); mbar_init(done + i, 1); mbar_init(ready + i, NT / 32); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = make_uniform(*s_tmem_p);
  const uint32_t xb = s_u32(smem);
  const int ncol = (p.N + BN - 1) / BN, nrow = (p.M + 127) / 128;
  const int ntiles = p.batch * ncol * nrow;
  const int nk = (p.K + 15) / 16;
  uint32_t g = 0;     // running k-tile counter (identical in every role): slot g % RING, use g / RING

  if (wu == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int z, r_base, c_base;
        if (!tile_of(p, BN, t, z, r_base, c_base)) continue;
        for (int kt = 0; kt < nk; ++kt, ++g) {
          const uint32_t s = g % G_RING;
          if (g >= G_RING) mbar_wait(done + s, ((g - G_RING) / G_RING) & 1u);       // MMAs of the slot's previous k-tile
          mbar_expect_tx(full + s, L::SLOT);
          tma_load_3d(&tmA, xb + s * L::SLOT, full + s, 16 * kt, r_base, z);
          tma_load_3d(&tmB, xb + s * L::SLOT + G_A_TILE, full + s, 16 * kt, c_base, z);
        }
      }
    }
    __syncwarp();
  } else if (wu == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int z, r_base, c_base;
        if (!tile_of(p, BN, t, z, r_base, c_base)) continue;
        for (int kt = 0; kt < nk; ++kt, ++g) {
          const uint32_t s = g % G_RING;
          mbar_wait(ready + s, (g / G_RING) & 1u);
          tc_fence_after();
          const uint64_t dbh = make_desc64(xb + s * L::SLOT + G_A_TILE), dbl = make_desc64(xb + L::BL + s * L::B_TILE_);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t adv = (uint64_t)(2 * ks);
            const uint32_t ah = tmem + L::AHI + 16 * s + 8 * ks, al = tmem + L::ALO + 16 * s + 8 * ks;
            umma_ts<BN>(tmem + L::ACC, al, dbh + adv, (kt == 0 && ks == 0) ? 0u : 1u);
            umma_ts<BN>(tmem + L::ACC, ah, dbl + adv, 1u);
            umma_ts<BN>(tmem + L::ACC, ah, dbh + adv, 1u);
          }
          umma_commit(done + s);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ workers: split, then the epilogue of the tile
    const int w = tid >> 5, lane = tid & 31;
    const int row = 32 * (w & 3) + lane, half_id = w >> 2;
    const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
    uint8_t* stage = smem + L::STG + half_id * G_STAGE;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int z, r_base, c_base;
      if (!tile_of(p, BN, t, z, r_base, c_base)) continue;
      for (int kt = 0; kt < nk; ++kt, ++g) {
        const uint32_t s = g % G_RING;
        const uint8_t* RAW = smem + s * L::SLOT;
        const uint8_t* BH = RAW + G_A_TILE;
        uint8_t* BL = smem + L::BL + s * L::B_TILE_;
        mbar_wait(full + s, (g / G_RING) & 1u);      // also: the MMAs of k-tile g - 4 are done (the producer waited for them)
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(RAW + swz64(row, 2 * half_id + q));
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            hi[4 * q + j] = __float_as_uint(e[j]);
            lo[4 * q + j] = __float_as_uint(e[j] - __uint_as_float(hi[4 * q + j] & 0xffffe000u));
          }
        }
        tmem_st8(tmem + lane_base + L::AHI + (uint32_t)(16 * s + 8 * half_id), hi);
        tmem_st8(tmem + lane_base + L::ALO + (uint32_t)(16 * s + 8 * half_id), lo);
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
          const int idx = tid + NT * i;
          const uint32_t off = swz64(idx >> 2, idx & 3);
          const float4 b = *reinterpret_cast<const float4*>(BH + off);
          float4 l;
          l.x = b.x - __uint_as_float(__float_as_uint(b.x) & 0xffffe000u);
          l.y = b.y - __uint_as_float(__float_as_uint(b.y) & 0xffffe000u);
          l.z = b.z - __uint_as_float(__float_as_uint(b.z) & 0xffffe000u);
          l.w = b.w - __uint_as_float(__float_as_uint(b.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(BL + off) = l;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ready + s);
      }
      {
        const uint32_t gl = g - 1u;                   // the last commit covers every MMA of the tile
        mbar_wait(done + (gl % G_RING), (gl / G_RING) & 1u);
      }
      tc_fence_after();
      // epilogue: the four warps of a column half share one staging piece; pieces of 32 columns, half_id takes every other one
      for (int piece = half_id; piece < BN / 32; piece += 2) {
        const int cc = 32 * piece, gc = c_base + cc;
        const bool live = gc < p.N && !(p.tri && (gc & ~63) > r_base + 127);
        if (!live) continue;                          // (uniform over the four warps)
        float sv[32];
        tmem_ld32(tmem + lane_base + L::ACC + (uint32_t)cc, sv);
        // the staging piece may still be read by the previous bulk store of this half: its issuer waited (below)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = make_float4(sv[4 * q], sv[4 * q + 1], sv[4 * q + 2], sv[4 * q + 3]);
          if (p.mode) { v.x = -v.x; v.y = -v.y; v.z = -v.z; v.w = -v.w; }
          *reinterpret_cast<float4*>(stage + row * 128 + ((q ^ (row & 7)) << 4)) = v;
        }
        fence_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(4 + half_id) : "memory");
        if ((w & 3) == 0 && lane == 0) {
          if (p.mode) tma_reduce_add_3d(&tmC, s_u32(stage), gc, r_base, z);
          else tma_store_3d(&tmC, s_u32(stage), gc, r_base, z);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the piece has been read: it may be overwritten
        }
        asm volatile("bar.sync %0, 128;" ::"r"(4 + half_id) : "memory");
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // every worker has read its accumulator rows before the next tile's first MMA
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (wu == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(L::TCOLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode3d(CUtensorMap* map, const float* base, int cols, int rows, int batch, long long ld, long long bstride, int box_cols,
                    int box_rows, CUtensorMapSwizzle swz) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return VOLT_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), (cuuint64_t)(batch > 1 ? bstride : (long long)rows * ld) * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%d rows=%d batch=%d ld=%lld)", (int)r, cols, rows, batch, ld);
    return VOLT_ERR_CUDA;
  }
  return VOLT_OK;
}

}  // namespace gnt

// C_z (= | -=) A_z B_z^T.  lda / ldb / ldc: row strides in floats (multiples of 4: 16-byte rows for the tensor maps);
// *_bstride: floats between consecutive batch members.  max_ctas <= 0: one CTA per SM.
int launch_gemm_nt(const float* A, long long lda, long long a_bstride, const float* B, long long ldb, long long b_bstride, float* C,
                   long long ldc, long long c_bstride, int M, int N, int K, int batch, int mode, int tri, int max_ctas, cudaStream_t st) {
  using namespace gnt;
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return VOLT_OK;
  if ((lda | ldb | ldc | a_bstride | b_bstride | c_bstride) & 3) {
    set_error("gemm_nt: row and batch strides must be multiples of 4 floats");
    return VOLT_ERR_ARG;
  }
  const int BN = (N > 128 && (long long)((M + 127) / 128) * ((N + 255) / 256) * batch >= sm_count() / 2) ? 256 : 128;
  CUtensorMap tmA, tmB, tmC;
  int s = encode3d(&tmA, A, K, M, batch, lda, a_bstride, 16, 128, CU_TENSOR_MAP_SWIZZLE_64B);
  if (s) return s;
  s = encode3d(&tmB, B, K, N, batch, ldb, b_bstride, 16, BN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (s) return s;
  s = encode3d(&tmC, C, N, M, batch, ldc, c_bstride, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (s) return s;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.tri = tri; p.mode = mode;
  const int ntiles = batch * ((M + 127) / 128) * ((N + BN - 1) / BN);
  int grid = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > ntiles) grid = ntiles;
  static bool attr_dev[16][2] = {};
  bool& attr = attr_dev[device_slot()][BN == 256];
  if (BN == 256) {
    if (!attr) {
      s = check_cuda(cudaFuncSetAttribute(gemm_nt_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<256>::BYTES),
                     "cudaFuncSetAttribute(gemm_nt_kernel)");
      if (s) return s;
      attr = true;
    }
    gemm_nt_kernel<256><<<grid, G_THREADS, Lay<256>::BYTES, st>>>(p, tmA, tmB, tmC);
  } else {
    if (!attr) {
      s = check_cuda(cudaFuncSetAttribute(gemm_nt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<128>::BYTES),
                     "cudaFuncSetAttribute(gemm_nt_kernel)");
      if (s) return s;
      attr = true;
    }
    gemm_nt_kernel<128><<<grid, G_THREADS, Lay<128>::BYTES, st>>>(p, tmA, tmB, tmC);
  }
  return check_cuda(cudaGetLastError(), "gemm_nt_kernel");
}

}  // namespace volt
