// Multi-CTA exact MLL + gradient for ONE long series (BASELINE config c5: T = 8192): right-looking blocked Cholesky
// and a right-looking inverse sweep, every step a set of independent 128x64 tiles spread over all SMs, the K = 64
// products on the tensor cores (tcgen05 kind::tf32, 3-pass hi/lo split) through the same building blocks as the
// batched kernel (chol_tc_dev.cuh).
//
//   W  (Tp x Tp): lower triangle = A, overwritten by L            Ut (Tp x Tp): upper triangle = R^T, overwritten by (L^-1)^T
//   for j:  diag + panel in one launch: every CTA factors W_jj (Linv_jj, z_j = Linv_jj r_j), then
//                  W[i,j] <- W[i,j] Linv_jj^T (i > j),  r_i -= W[i,j] z_j                      ((T-R0)/128 CTAs)
//           update W[i,c] -= W[i,j] W[c,j]^T  (j < c <= i)                                      (tiles)
//   for k:  final  Ut[0:k+1, k] <- Ut[0:k+1, k] Linv_kk^T;  tr += |.|^2, alpha += (.) z_k       ((k+1)/2 CTAs)
//           update Ut[n, i] -= Ut[n, k] W[i, k]^T  (n <= k < i)                                  (tiles)
// Replaces the same reference call chain as chol_tc.cu (train_utils.py:247-250) when a single series is too long for
// one CTA to be a sensible unit of work.
#include "chol_tc_dev.cuh"

#include <cmath>
#include <cstring>

namespace volt {
namespace large {

using namespace tc;

struct LargeParams {
  int T, Tp, nb, kind;
  const float* V;      // (T) cumtrapz prefix (KIND_VOL) or grid x (KIND_BM)
  const float* dense;  // (T, ldd) KIND_DENSE
  int ldd;
  float scale, dadd;
  const float* resid;  // (T)
  float* W; float* Ut; float* dinv;
  float* z;            // (Tp): residual, updated right-looking (r_i -= W[i,j] z_j)
  float* zf;           // (Tp): z = L^-1 r, block j written by CTA 0 of step j
  float* alpha;        // (Tp)
  float* acc;          // [0] sum log L_ii
  float* trp;          // (nb, trp_ld) per-(step, CTA) partial sums of |L^-1|_F^2, added in a fixed order by the finish kernel
  int trp_ld;
  float* origd;        // (Tp) original diagonal of A (pivot failure predicate)
  int* flag;           // first failing pivot index, -1 = none
  int* step_flag;      // (nb) step_flag[j] == epoch: the diagonal block of step j of this attempt has been published
  int epoch;
};

__device__ __forceinline__ float gen_large(const LargeParams& p, int i, int j) {
  if (i >= p.T || j >= p.T) return (i == j) ? 1.f : 0.f;
  float v;
  if (p.kind == KIND_VOL) v = p.V[min(i, j)];
  else if (p.kind == KIND_BM) v = p.scale * fminf(p.V[i], p.V[j]);
  else v = (i >= j) ? p.dense[(size_t)i * p.ldd + j] : p.dense[(size_t)j * p.ldd + i];
  if (i == j) v += p.dadd;
  return v;
}

// lower triangle of A -> W, residual -> z, alpha = 0, Ut = I (Ut is memset to 0 by the host first)
__global__ void __launch_bounds__(256) large_build_kernel(LargeParams p) {
  const int i = blockIdx.x;
  float* row = p.W + (size_t)i * p.Tp;
  for (int j = threadIdx.x; j <= i; j += 256) row[j] = gen_large(p, i, j);
  if (threadIdx.x == 0) {
    p.z[i] = (i < p.T) ? p.resid[i] : 0.f;
    p.alpha[i] = 0.f;
    p.origd[i] = gen_large(p, i, i);
    p.Ut[(size_t)i * p.Tp + i] = 1.f;
    if (i == 0) { p.acc[0] = 0.f; p.acc[1] = 0.f; *p.flag = -1; }
  }
}

struct Shared {
  Ctx c;
  float* LiT; float* tmpbuf;
};

// Programmatic dependent launch of the step kernels (a chain of ~256 short launches on one stream): a step may be
// scheduled while its predecessor is still running, so that its prologue (shared-memory carve-up, TMEM allocation,
// mbarrier init) is off the critical path; it touches global memory only after pdl_wait(), i.e. after the predecessor
// grid has completed and flushed.  Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// common prologue: carve shared memory like the batched kernel, allocate TMEM, init the mbarrier
__device__ __forceinline__ void cta_setup(Shared& sh, uint8_t* base, bool need_tmem) {
  Ctx& c = sh.c;
  c.X = base;
  c.Lr = base + L_OFF;
  c.Ct = reinterpret_cast<float*>(base + CT_OFF);
  c.Vs = reinterpret_cast<float*>(base + VEC_OFF);
  c.z = c.Vs; c.al = c.Vs; c.z2 = c.Vs;
  c.diagl = c.Vs;                 // 64
  c.tmp = c.diagl + NB;           // 192: [0,64) origd then z_j, [64,192) z residual / per-row half sums
  c.red = c.tmp + 3 * NB;         // 32
  c.flag = reinterpret_cast<int*>(c.red + 32);
  c.bar = reinterpret_cast<uint64_t*>(c.red + 36);
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(c.red + 40);   // c.bar holds two mbarriers (16 bytes)
  c.phase = 0;
  sh.LiT = reinterpret_cast<float*>(c.X + X_LIT);
  sh.tmpbuf = reinterpret_cast<float*>(c.X + X_TMP);
  if ((s_u32(base) & 1023u) != 0u) __trap();
  if (need_tmem) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s_u32(s_tmem_p)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
      mbar_init(c.bar, 1);
      mbar_init(c.bar + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *s_tmem_p;
  }
}
__device__ __forceinline__ void cta_teardown(Shared& sh) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(sh.c.tmem) : "memory");
}
constexpr size_t LARGE_SMEM = VEC_OFF + sizeof(float) * (NB + 3 * NB + 32 + 12);

// ---- step j of the factorisation: diagonal block + panel in ONE launch.  Every CTA factors and inverts the 64 x 64
// diagonal block itself (10 us of redundant work instead of a 1-CTA kernel and a launch boundary on the critical
// path: 128 of them at T = 8192), CTA 0 publishes Linv_jj / z_j / log-det / the failure flag, then CTA b solves its
// 128 panel rows W[i,j] <- W[i,j] Linv_jj^T and updates the residual r_i -= W[i,j] z_j.  L_jj itself is not written
// back: nothing reads the diagonal blocks of W again.
__global__ void __launch_bounds__(NT, 2) large_diagpanel_kernel(LargeParams p, int j, int kp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Shared sh;
  pdl_release();
  cta_setup(sh, smem_raw, true);
  pdl_wait();
  Ctx& c = sh.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half_id = warp >> 2, c0 = half_id * 32;
  const int R0 = j * NB, ld = p.Tp;
  const uint32_t t_lane = (uint32_t)(32 * (warp & 3)) << 16;
  // Inside a 256-column panel the updates are LEFT-looking: block column j receives the contributions of the panel's
  // earlier block columns [kp, R0) right here (K <= 192), so no separate update launch sits between two steps.
  // Round 2: the roles are split by CTA.  CTA 0 forms, factors and inverts the 64 x 64 diagonal block and publishes
  // Linv_jj / z_j / log-det / the failure flag behind a per-step flag; CTAs 1.. apply the earlier block columns to their 128
  // rows FIRST (that product does not need the diagonal block), then wait for the flag, then solve.  Before, every CTA
  // factored the block itself and only then started its own product: 26.6 -> 21 us per step at T = 8192.
  if (blockIdx.x == 0) {
    {
      float u[32];
      const bool have = gemm_tc<false>(c, p.W, ld, R0, p.Tp, R0, kp, R0, nullptr);   // rows R0.. (the diagonal block is rows < 64)
      if (have) {
        tmem_ld32(c.tmem + t_lane + (uint32_t)c0, u);
        tc_fence_before();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) u[q] = 0.f;
      }
      if (row < NB) {
        const float4* src = reinterpret_cast<const float4*>(p.W + (size_t)(R0 + row) * ld + R0 + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = src[q];
          *reinterpret_cast<float4*>(c.Ct + row * CLD + c0 + 4 * q) =
              make_float4(v.x - u[4 * q], v.y - u[4 * q + 1], v.z - u[4 * q + 2], v.w - u[4 * q + 3]);
        }
      }
    }
    if (tid == 0) *c.flag = -1;
    if (tid < NB) { c.tmp[tid] = p.origd[R0 + tid]; c.tmp[NB + tid] = p.z[R0 + tid]; }
    wsync();
    diag64_block_v2<CLD>(c.Ct, sh.LiT, sh.tmpbuf, c.diagl, c.tmp, c.flag, R0);
    wsync();
    float zj = 0.f;
    if (tid < NB)
      for (int k = 0; k <= tid; ++k) zj = fmaf(sh.LiT[k * CLD + tid], c.tmp[NB + k], zj);
    float* dj = p.dinv + (size_t)j * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += NT) dj[idx] = sh.LiT[(idx >> 6) * CLD + (idx & 63)];
    if (tid < NB) p.zf[R0 + tid] = zj;               // z_j = Linv_jj r_j
    float lg = (tid < NB && R0 + tid < p.T) ? logf(c.diagl[tid]) : 0.f;
    lg = block_sum(lg, c.red);
    if (tid == 0) {
      p.acc[0] += lg;                                  // one writer per step, launches are stream-ordered
      if (*c.flag >= 0 && *p.flag < 0) *p.flag = *c.flag;
    }
    __threadfence();
    wsync();
    if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.step_flag + j), "r"(p.epoch) : "memory");
  } else {
    const int r_base = R0 + NB + CM * ((int)blockIdx.x - 1), row_end = p.Tp;
    const int gr = r_base + row;
    float s[32], o[32];
    const bool have = gemm_tc<false>(c, p.W, ld, r_base, row_end, R0, kp, R0, nullptr);
    if (have) {
      tmem_ld32(c.tmem + t_lane + (uint32_t)c0, s);
      tc_fence_before();
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) s[q] = 0.f;
    }
    if (gr < row_end) {
      const float4* src = reinterpret_cast<const float4*>(p.W + (size_t)gr * ld + R0 + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = src[q];
        s[4 * q] = v.x - s[4 * q]; s[4 * q + 1] = v.y - s[4 * q + 1]; s[4 * q + 2] = v.z - s[4 * q + 2]; s[4 * q + 3] = v.w - s[4 * q + 3];
      }
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) s[q] = 0.f;
    }
    // the diagonal block of this step (CTA 0 of this launch; every CTA of the launch is resident: the grid is at most
    // Tp / 128 + 1 CTAs at two per SM)
    if (tid == 0) {
      int f = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.step_flag + j) : "memory");
        if (f == p.epoch) break;
        __nanosleep(64);
      }
    }
    wsync();
    stage_linv_from_dinv(c, p.dinv + (size_t)j * NB * NB);
    if (tid < NB) c.tmp[tid] = p.zf[R0 + tid];
    wsync();
    trsm_tc(c, s, o, row, half_id);
    {
      const int g0 = r_base + 32 * (warp & 3);
      if (g0 < row_end) store_block32(reinterpret_cast<float*>(c.X) + warp * 1152, o, p.W + (size_t)g0 * ld + R0 + c0, ld, lane);
    }
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) dot = fmaf(o[q], c.tmp[c0 + q], dot);
    if (half_id) c.tmp[NB + row] = dot;                // the two column halves of a row are combined in a fixed order
    wsync();
    if (half_id == 0 && gr < row_end) p.z[gr] -= dot + c.tmp[NB + row];
  }
  cta_teardown(sh);
}

// ---- panel (mode 0: Cholesky panel of step j; mode 1: finalise block row j of the inverse)
__global__ void __launch_bounds__(NT, 2) large_panel_kernel(LargeParams p, int j, int mode, int kp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Shared sh;
  pdl_release();
  cta_setup(sh, smem_raw, true);
  pdl_wait();
  Ctx& c = sh.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half_id = warp >> 2, c0 = half_id * 32;
  const int R0 = j * NB, ld = p.Tp;
  float* M = mode ? p.Ut : p.W;
  const int r_base = (mode ? 0 : R0 + NB) + CM * blockIdx.x;
  const int row_end = mode ? R0 + NB : p.Tp;
  const int gr = r_base + row;
  stage_linv_from_dinv(c, p.dinv + (size_t)j * NB * NB);
  float s[32], o[32];
  // left-looking inside the panel: contributions of the panel's earlier block columns [kp, R0) (B operand rows from W)
  const bool have = gemm_tc<false>(c, M, ld, r_base, row_end, R0, kp, R0, nullptr, p.W);
  if (have) {
    tmem_ld32(c.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, s);
    tc_fence_before();
  } else {
#pragma unroll
    for (int q = 0; q < 32; ++q) s[q] = 0.f;
  }
  if (gr < row_end) {
    const float4* src = reinterpret_cast<const float4*>(M + (size_t)gr * ld + R0 + c0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = src[q];
      s[4 * q] = v.x - s[4 * q]; s[4 * q + 1] = v.y - s[4 * q + 1]; s[4 * q + 2] = v.z - s[4 * q + 2]; s[4 * q + 3] = v.w - s[4 * q + 3];
    }
  } else {
#pragma unroll
    for (int q = 0; q < 32; ++q) s[q] = 0.f;
  }
  __syncthreads();
  trsm_tc(c, s, o, row, half_id);
  {
    const int g0 = r_base + 32 * (warp & 3);
    if (g0 < row_end) store_block32(reinterpret_cast<float*>(c.X) + warp * 1152, o, M + (size_t)g0 * ld + R0 + c0, ld, lane);
  }
  float dot = 0.f, sq = 0.f;
  if (gr < row_end) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      dot = fmaf(o[q], (mode ? p.zf : p.z)[R0 + c0 + q], dot);
      if (mode && gr < p.T && R0 + c0 + q < p.T) sq = fmaf(o[q], o[q], sq);
    }
  }
  if (half_id) c.tmp[row] = dot;
  __syncthreads();
  if (half_id == 0 && gr < row_end) {
    const float d2 = dot + c.tmp[row];
    if (mode) p.alpha[gr] += d2;      // alpha = X^T z, one owner per row per launch
    else p.z[gr] -= d2;               // right-looking forward substitution of the residual
  }
  if (mode) {
    sq = block_sum(sq, c.red);
    if (tid == 0) p.trp[(size_t)j * p.trp_ld + blockIdx.x] = sq;   // one slot per (step, CTA): no atomics, reproducible sum
  }
  cta_teardown(sh);
}

// ---- trailing update tiles: C[rows, cols] -= A[rows, k_lo:k_hi] B[cols, k_lo:k_hi]^T, 128 x 64 tiles.
//   mode 0 (Cholesky):  A = C = W, B = W; rows >= row_lo, only tiles touching the lower triangle.
//   mode 1 (inverse):   A = C = Ut, B = W; rows < row_end (Ut is upper triangular).
// The host calls it with K = 64 inside a 256-column panel and once with K = 256 for everything to the right of the
// panel, which cuts the read-modify-write traffic on C by 4x compared with a plain NB = 64 right-looking sweep.
// Three CTAs per SM: 128 TMEM columns, the 32 KB single-buffered stage of the batched kernel's small instance (gemm_tc1) and
// nothing else in shared memory -- a tile's fixed costs (TMEM allocation, first loads, MMA drain) overlap with the
// other two resident tiles.
constexpr size_t UPDATE_SMEM = Y_BYTES + 64;
__global__ void __launch_bounds__(NT, 3) large_update_kernel(LargeParams p, int mode, int row_lo, int row_end, int col_lo, int col_hi,
                                                             int k_lo, int k_hi) {
  const int ld = p.Tp;
  const int ncol = (col_hi - col_lo) / NB, nrow = (row_end - row_lo + CM - 1) / CM;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((s_u32(smem_raw) & 1023u) != 0u) __trap();
  Ctx c;
  c.X = smem_raw;
  c.bar = reinterpret_cast<uint64_t*>(smem_raw + Y_BYTES);
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(smem_raw + Y_BYTES + 16);
  c.phase = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(s_tmem_p)), "n"(T3_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(c.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *s_tmem_p;
  const int row = 32 * (warp & 3) + lane, half_id = warp >> 2, c0 = half_id * 32;
  float* Cm = mode ? p.Ut : p.W;
  // persistent over the launch's tiles (row-major within a tile row so that neighbouring CTAs share the A panel rows)
  for (int t = blockIdx.x; t < ncol * nrow; t += gridDim.x) {
    const int r_base = row_lo + CM * (t / ncol);      // rows of the C tile
    const int c_base = col_lo + NB * (t % ncol);      // columns of the C tile
    if (!mode && c_base > r_base + CM - 1) continue;  // tile entirely above the diagonal
    gemm_tc1<false>(c, Cm, ld, r_base, row_end, c_base, k_lo, k_hi, nullptr, p.W);
    float s[32];
    tmem_ld32(c.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, s);
    tc_fence_before();
    const int gr = r_base + row;
    if (gr < row_end) {
      float4* dst = reinterpret_cast<float4*>(Cm + (size_t)gr * ld + c_base + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = dst[q];
        v.x -= s[4 * q]; v.y -= s[4 * q + 1]; v.z -= s[4 * q + 2]; v.w -= s[4 * q + 3];
        dst[q] = v;
      }
    }
    __syncthreads();   // every thread has read its accumulator row before the next tile's first MMA overwrites it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "n"(T3_COLS) : "memory");
}

// ---- the same update on 128 x 256 tiles, one CTA per SM.  A 128 x 64 tile re-stages its 128-row A panel for every 64
// columns and reads 192 KB of operands for 4.2 MFLOP; four times the columns per staged A tile halve the operand
// traffic and the per-flop staging work.  TMEM: accumulator 256 columns + two A stages (hi | lo, 32 columns each);
// shared memory: raw A tile 16 KB + two B stages of 256 rows x 32 floats, hi and lo (2 x 64 KB); double-buffered like
// gemm_tc (the MMAs of k-tile kt overlap the staging of kt + 1).
constexpr uint32_t U2_B0 = A_TILE, U2_BSTAGE = 2u * 256u * 128u;               // 64 KB per stage: hi 32 KB | lo 32 KB
constexpr size_t UPDATE2_SMEM = A_TILE + 2 * U2_BSTAGE + 64;
constexpr uint32_t U2_ACC = 0, U2_A0 = 256;                                      // stage h: hi at 256 + 64 h, lo at 288 + 64 h
constexpr uint32_t IDESC256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_tf32_ts256(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(IDESC256), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(NT, 1) large_update256_kernel(LargeParams p, int mode, int row_lo, int row_end, int col_lo, int col_hi,
                                                                int k_lo, int k_hi) {
  const int ld = p.Tp;
  const int ncol = (col_hi - col_lo + 255) / 256, nrow = (row_end - row_lo + CM - 1) / CM;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((s_u32(smem_raw) & 1023u) != 0u) __trap();
  Ctx c;
  c.X = smem_raw;
  c.bar = reinterpret_cast<uint64_t*>(smem_raw + A_TILE + 2 * U2_BSTAGE);
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(smem_raw + A_TILE + 2 * U2_BSTAGE + 16);
  c.phase = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(s_tmem_p)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(c.bar, 1);
    mbar_init(c.bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *s_tmem_p;
  const int row = 32 * (warp & 3) + lane, half_id = warp >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  float* Cm = mode ? p.Ut : p.W;
  uint8_t* RAW = c.X;
  const int wu = uniform_warp_id();                     // MMA issue from one elected lane of warp 0, uniform operands (elect_one)
  const uint32_t tmem_u = make_uniform(c.tmem), xb = s_u32(c.X);
  const int nk = (k_hi - k_lo) / 32;
  for (int t = blockIdx.x; t < ncol * nrow; t += gridDim.x) {
    const int r_base = row_lo + CM * (t / ncol);      // rows of the C tile
    const int c_base = col_lo + 256 * (t % ncol);     // columns of the C tile
    if (!mode && c_base > r_base + CM - 1) continue;  // tile entirely above the diagonal
    float4 ra[4], rb[8];
    auto gload = [&](int k0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
        ra[i] = load_a<false>(Cm, ld, r_base + r, row_end, k0 + chunk * 4, nullptr);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
        rb[i] = (c_base + r < col_hi) ? *reinterpret_cast<const float4*>(p.W + (size_t)(c_base + r) * ld + k0 + chunk * 4)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    gload(k_lo);
    for (int kt = 0; kt < nk; ++kt) {
      const int h = kt & 1;
      uint8_t* BH = c.X + U2_B0 + h * U2_BSTAGE;
      uint8_t* BL = BH + U2_BSTAGE / 2;
      if (kt >= 2) wait_mma2(c, h);   // MMA group kt-2 has consumed TMEM / shared-memory stage h
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
        *reinterpret_cast<float4*>(RAW + r * 128 + ((chunk ^ (r & 7)) << 4)) = ra[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + NT * i;
        st_split(BH, BL, idx >> 3, idx & 7, rb[i]);
      }
      if (kt + 1 < nk) gload(k_lo + (kt + 1) * 32);
      __syncthreads();
      {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = 4 * half_id + q;
          const float4 v = *reinterpret_cast<const float4*>(RAW + row * 128 + ((chunk ^ (row & 7)) << 4));
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            hi[4 * q + j] = __float_as_uint(e[j]);
            lo[4 * q + j] = __float_as_uint(e[j] - __uint_as_float(hi[4 * q + j] & 0xffffe000u));
          }
        }
        tmem_st16(c.tmem + lane_base + U2_A0 + (uint32_t)(64 * h + 16 * half_id), hi);
        tmem_st16(c.tmem + lane_base + U2_A0 + 32u + (uint32_t)(64 * h + 16 * half_id), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (wu == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dbh = make_desc(xb + U2_B0 + h * U2_BSTAGE), dbl = make_desc(xb + U2_B0 + h * U2_BSTAGE + U2_BSTAGE / 2);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t adv = (uint64_t)(2 * ks);
            const uint32_t ah = tmem_u + U2_A0 + 64 * h + 8 * ks, al = ah + 32;
            umma_tf32_ts256(tmem_u + U2_ACC, al, dbh + adv, (kt == 0 && ks == 0) ? 0u : 1u);
            umma_tf32_ts256(tmem_u + U2_ACC, ah, dbl + adv, 1u);
            umma_tf32_ts256(tmem_u + U2_ACC, ah, dbh + adv, 1u);
          }
          umma_commit(c.bar + h);
        }
        __syncwarp();
      }
    }
    wait_mma2(c, (nk - 1) & 1);
    if (nk >= 2) wait_mma2(c, (nk - 2) & 1);
    tc_fence_after();
    const int gr = r_base + row;
#pragma unroll 1
    for (int piece = 0; piece < 4; ++piece) {
      const int cc = 128 * half_id + 32 * piece;          // column offset inside the tile
      const int gc = c_base + cc;
      float s[32];
      tmem_ld32(c.tmem + lane_base + U2_ACC + (uint32_t)cc, s);
      const bool live = gc < col_hi && (mode || (gc & ~63) <= r_base + CM - 1);   // same 64-column blocks as the 128 x 64 tiles
      if (live && gr < row_end) {
        float4* dst = reinterpret_cast<float4*>(Cm + (size_t)gr * ld + gc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = dst[q];
          v.x -= s[4 * q]; v.y -= s[4 * q + 1]; v.z -= s[4 * q + 2]; v.w -= s[4 * q + 3];
          dst[q] = v;
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // every thread has read its accumulator row before the next tile's first MMA overwrites it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(c.tmem) : "memory");
}

// ---- scalars
__global__ void __launch_bounds__(256) large_finish_kernel(LargeParams p, float jit_used, float* scalars, float* alpha_out, int* info) {
  __shared__ float red[32];
  float zz = 0.f, aa = 0.f, ar = 0.f, trs = 0.f;
  for (int i = threadIdx.x; i < p.nb * p.trp_ld; i += 256) trs += p.trp[i];
  for (int i = threadIdx.x; i < p.T; i += 256) {
    const float zi = p.zf[i], ai = p.alpha[i];
    zz = fmaf(zi, zi, zz);
    aa = fmaf(ai, ai, aa);
    ar = fmaf(ai, p.resid[i], ar);
    if (alpha_out) alpha_out[i] = ai;
  }
  const float inv_quad = block_sum(zz, red), alal = block_sum(aa, red), alr = block_sum(ar, red), tr_sum = block_sum(trs, red);
  if (threadIdx.x == 0) {
    const float Tf = (float)p.T, logdet = 2.f * p.acc[0], tr = tr_sum;
    scalars[0] = -0.5f * (inv_quad + logdet + Tf * 1.8378770664093453f) / Tf;
    scalars[1] = 0.5f * (alal - tr) / Tf;
    scalars[2] = logdet; scalars[3] = inv_quad; scalars[4] = tr; scalars[5] = alal; scalars[6] = alr; scalars[7] = jit_used;
    for (int q = 8; q < NSCALARS; ++q) scalars[q] = 0.f;
    if (info) *info = (*p.flag >= 0) ? *p.flag + 1 : 0;
  }
}

}  // namespace large

// One series per call (the dispatcher loops over the batch).  Returns VOLT_OK; numerical failure is reported in info.
int launch_mll_large(const MllParams& mp, int b, cudaStream_t st) {
  using namespace large;
  LargeParams p;
  memset(&p, 0, sizeof(p));
  p.T = mp.T;
  p.Tp = (mp.T + NB - 1) / NB * NB;
  p.nb = p.Tp / NB;
  p.kind = mp.kind;
  const size_t tp2 = (size_t)p.Tp * p.Tp;
  void *w = nullptr, *u = nullptr, *aux = nullptr;
  int s = get_workspace(tp2 * sizeof(float), &w, 9, st);
  if (s) return s;
  s = get_workspace(tp2 * sizeof(float), &u, 10, st);
  if (s) return s;
  p.trp_ld = (p.Tp + CM - 1) / CM + 1;
  const size_t aux_fl = (size_t)p.nb * NB * NB + 4 * (size_t)p.Tp + 16 + (size_t)p.nb * p.trp_ld + (size_t)p.nb;
  s = get_workspace(aux_fl * sizeof(float), &aux, 11, st);
  if (s) return s;
  p.W = (float*)w;
  p.Ut = (float*)u;
  p.dinv = (float*)aux;
  p.z = p.dinv + (size_t)p.nb * NB * NB;
  p.alpha = p.z + p.Tp;
  p.origd = p.alpha + p.Tp;
  p.zf = p.origd + p.Tp;
  p.acc = p.zf + p.Tp;
  p.flag = reinterpret_cast<int*>(p.acc + 4);
  p.trp = p.acc + 16;
  p.step_flag = reinterpret_cast<int*>(p.trp + (size_t)p.nb * p.trp_ld);
  VOLT_CUDA(cudaMemsetAsync(p.step_flag, 0, (size_t)p.nb * sizeof(int), st));
  if (mp.kind == KIND_VOL) p.V = mp.V + (size_t)b * mp.T;
  else if (mp.kind == KIND_BM) { p.V = mp.x; }
  else { p.dense = mp.dense + (size_t)b * mp.dense_bstride; p.ldd = mp.ldd; }
  p.resid = mp.resid + (size_t)b * mp.T;
  float scale = 1.f, dadd0 = 0.f;
  if (mp.kind == KIND_BM) VOLT_CUDA(cudaMemcpyAsync(&scale, mp.scale + (size_t)b * mp.scale_stride, 4, cudaMemcpyDeviceToHost, st));
  if (mp.diag_add) VOLT_CUDA(cudaMemcpyAsync(&dadd0, mp.diag_add + (size_t)b * mp.diag_stride, 4, cudaMemcpyDeviceToHost, st));
  if (mp.kind == KIND_BM || mp.diag_add) VOLT_CUDA(cudaStreamSynchronize(st));
  p.scale = scale;
  static bool attr_dev[16] = {};   // function attributes, streams and events are per device
  bool& attr = attr_dev[device_slot()];
  if (!attr) {
    VOLT_CUDA(cudaFuncSetAttribute(large_diagpanel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LARGE_SMEM));
    VOLT_CUDA(cudaFuncSetAttribute(large_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LARGE_SMEM));
    VOLT_CUDA(cudaFuncSetAttribute(large_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPDATE_SMEM));
    VOLT_CUDA(cudaFuncSetAttribute(large_update256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPDATE2_SMEM));
    attr = true;
  }
  // Look-ahead over two streams.  The deferred K = 256 update of panel P is split by columns: (a) the columns of panel
  // P+1, on the critical path (stream `st`), and (b) everything to the right of them, which only has to be finished
  // before the next deferred update touches the same tiles -- it runs on a second stream while panel P+1 (a chain of
  // 1-CTA diagonal kernels, TRSM panels and narrow updates that idles most of the GPU) is being factored.
  struct LookAhead { cudaStream_t sb; cudaEvent_t ev_panel, ev_rest, ev_join; };
  static LookAhead la_dev[16] = {};
  LookAhead& la = la_dev[device_slot()];
  cudaStream_t& sb = la.sb;
  cudaEvent_t& ev_panel = la.ev_panel;
  cudaEvent_t& ev_rest = la.ev_rest;
  cudaEvent_t& ev_join = la.ev_join;
  if (!sb) {
    VOLT_CUDA(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    VOLT_CUDA(cudaEventCreateWithFlags(&ev_panel, cudaEventDisableTiming));
    VOLT_CUDA(cudaEventCreateWithFlags(&ev_rest, cudaEventDisableTiming));
    VOLT_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  static const int use_pdl = [] { const char* e = getenv("VOLT_PDL"); return e ? atoi(e) : 1; }();   // 0: plain launches (A/B timing)
  auto launch_step = [&](auto kernel, int grid, auto... args) -> cudaError_t {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = LARGE_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
  };
  int upd_status = VOLT_OK;
  constexpr int PB = 4;  // blocks per panel: trailing updates outside the panel are deferred and applied with K = 256
  auto update = [&](cudaStream_t s2, int mode, int row_lo, int row_end, int col_lo, int col_hi, int k_lo, int k_hi) {
    const int nrow = (row_end - row_lo + CM - 1) / CM, ncol = (col_hi - col_lo) / NB;
    if (nrow <= 0 || ncol <= 0) return;
    static const int wide = [] { const char* e = getenv("VOLT_UPD256"); return e ? atoi(e) : 1; }();   // 0: 128 x 64 tiles (A/B timing)
    static const int wide_min = [] { const char* e = getenv("VOLT_UPD256_MIN"); return e ? atoi(e) : 1; }();
    const int ncol2w = (col_hi - col_lo + 255) / 256;
    const int live = mode ? nrow * ncol2w : (nrow * ncol2w + 1) / 2 + ncol2w;   // mode 0 keeps the tiles touching the lower triangle
    static const int gemm_min_cols = [] { const char* e = getenv("VOLT_UPD_GEMM_MINCOLS"); return e ? atoi(e) : 256; }();   // c5: 5.94 (256) vs 6.10 ms (512)
    static const int use_gemm0 = [] { const char* e = getenv("VOLT_UPD_GEMM"); return e ? atoi(e) : 1; }();
    if (wide && col_hi - col_lo >= (use_gemm0 ? gemm_min_cols : 512) && k_hi - k_lo >= 64 && live >= wide_min) {
      const int ncol2 = ncol2w;
      // the look-ahead part on the side stream leaves a quarter of the SMs to the step kernels of the critical path
      static const int cap_pct = [] { const char* e = getenv("VOLT_UPD256_CAP"); return e ? atoi(e) : 75; }();
      const int cap = (s2 == st) ? sm_count() : max(1, (cap_pct * sm_count()) / 100);
      // round 2: the TMA-fed product kernel (gemm_nt.cu: operand tiles by cp.async.bulk.tensor, C -= ... as a TMA reduction in
      // the L2).  VOLT_UPD_GEMM=0: the register-staged 128 x 256 tile kernel of round 1 (A/B timing).
      static const int use_gemm = [] { const char* e = getenv("VOLT_UPD_GEMM"); return e ? atoi(e) : 1; }();
      if (use_gemm) {
        const size_t ld = (size_t)p.Tp;
        int rc;
        if (mode == 0)
          rc = launch_gemm_nt(p.W + (size_t)row_lo * ld + k_lo, (long long)ld, 0, p.W + (size_t)col_lo * ld + k_lo, (long long)ld, 0,
                              p.W + (size_t)row_lo * ld + col_lo, (long long)ld, 0, row_end - row_lo, col_hi - col_lo, k_hi - k_lo, 1, 1,
                              row_lo == col_lo ? 1 : 0, cap, s2);
        else
          rc = launch_gemm_nt(p.Ut + k_lo, (long long)ld, 0, p.W + (size_t)col_lo * ld + k_lo, (long long)ld, 0, p.Ut + col_lo,
                              (long long)ld, 0, row_end, col_hi - col_lo, k_hi - k_lo, 1, 1, 0, cap, s2);
        if (rc) upd_status = rc;
        return;
      }
      large_update256_kernel<<<min(ncol2 * nrow, cap), NT, UPDATE2_SMEM, s2>>>(p, mode, row_lo, row_end, col_lo, col_hi, k_lo, k_hi);
      return;
    }
    const int grid = min(ncol * nrow, 3 * sm_count());
    large_update_kernel<<<grid, NT, UPDATE_SMEM, s2>>>(p, mode, row_lo, row_end, col_lo, col_hi, k_lo, k_hi);
  };
  // deferred update of the panel ending at `panel_end`: part (a) on st after the previous part (b) has left the tiles,
  // part (b) on sb once the panel is final
  auto deferred = [&](int mode, int panel_end, bool& rest_pending) -> int {
    const int next_end = min(p.Tp, panel_end + PB * NB);
    const int k_lo = panel_end - PB * NB;
    VOLT_CUDA(cudaEventRecord(ev_panel, st));
    if (rest_pending) VOLT_CUDA(cudaStreamWaitEvent(st, ev_rest, 0));
    if (mode == 0) update(st, 0, panel_end, p.Tp, panel_end, next_end, k_lo, panel_end);
    else update(st, 1, 0, panel_end, panel_end, next_end, k_lo, panel_end);
    if (next_end < p.Tp) {
      VOLT_CUDA(cudaStreamWaitEvent(sb, ev_panel, 0));
      if (mode == 0) update(sb, 0, next_end, p.Tp, next_end, p.Tp, k_lo, panel_end);
      else update(sb, 1, 0, panel_end, next_end, p.Tp, k_lo, panel_end);
      VOLT_CUDA(cudaEventRecord(ev_rest, sb));
      rest_pending = true;
    }
    return VOLT_OK;
  };
  auto join = [&](bool& rest_pending) -> int {
    if (rest_pending) VOLT_CUDA(cudaStreamWaitEvent(st, ev_rest, 0));
    rest_pending = false;
    return VOLT_OK;
  };
  float jit_used = 0.f;
  for (int attempt = 0;; ++attempt) {
    p.dadd = dadd0 + jit_used;
    p.epoch = attempt + 1;
    VOLT_CUDA(cudaMemsetAsync(p.Ut, 0, tp2 * sizeof(float), st));
    large_build_kernel<<<p.Tp, 256, 0, st>>>(p);
    bool rest_pending = false;
    for (int j = 0; j < p.nb; ++j) {
      const int R0 = j * NB, panel_end = min(p.Tp, (j / PB + 1) * PB * NB);
      const int rows = p.Tp - (R0 + NB);
      VOLT_CUDA(launch_step(large_diagpanel_kernel, 1 + max(0, (rows + CM - 1) / CM), p, j, (j / PB) * PB * NB));
      if (rows > 0 && R0 + NB == panel_end) { s = deferred(0, panel_end, rest_pending); if (s) return s; }
    }
    s = join(rest_pending);
    if (s) return s;
    int flag = -1;
    VOLT_CUDA(cudaMemcpyAsync(&flag, p.flag, 4, cudaMemcpyDeviceToHost, st));
    VOLT_CUDA(cudaStreamSynchronize(st));
    if (flag < 0 || attempt >= mp.max_tries || !(mp.jitter > 0.f)) break;
    jit_used = mp.jitter * powf(10.f, (float)attempt);
  }
  VOLT_CUDA(cudaMemsetAsync(p.trp, 0, (size_t)p.nb * p.trp_ld * sizeof(float), st));
  if (mp.do_inverse) {
    bool rest_pending = false;
    for (int k = 0; k < p.nb; ++k) {
      const int R0 = k * NB, rows_done = R0 + NB, panel_end = min(p.Tp, (k / PB + 1) * PB * NB);
      VOLT_CUDA(launch_step(large_panel_kernel, (rows_done + CM - 1) / CM, p, k, 1, (k / PB) * PB * NB));
      if (rows_done == panel_end && panel_end < p.Tp) { s = deferred(1, panel_end, rest_pending); if (s) return s; }
    }
    s = join(rest_pending);
    if (s) return s;
  }
  if (upd_status) return upd_status;
  large_finish_kernel<<<1, 256, 0, st>>>(p, jit_used, mp.scalars + (size_t)b * NSCALARS,
                                         (mp.alpha && mp.do_inverse) ? mp.alpha + (size_t)b * mp.T : nullptr,
                                         mp.info ? mp.info + b : nullptr);
  return check_cuda(cudaGetLastError(), "mll_large_kernels");
}

}  // namespace volt
