// Device building blocks shared by the SIMT (chol_batched.cu) and tcgen05 (chol_tc.cu) batched Cholesky kernels.
#pragma once
#include "params.cuh"

namespace volt {

constexpr int NB = 64;    // block-column width
constexpr int CM = 128;   // rows per chunk
constexpr int BK = 16;    // k-tile of the staged GEMM
constexpr int NT = 256;   // threads per CTA
constexpr int AS_LD = CM + 4;
constexpr int BS_LD = NB + 4;
constexpr int CT_LD = CM + 4;  // Ct[c][r]: chunk result, column-major ("transposed") so it can be re-used as a K-major A tile
constexpr int LI_LD = NB + 4;  // LiT[k][c] = Linv[c][k]

// CTA barrier of the NT = 256 worker threads: named barrier 1 with an explicit thread count.  Identical to wsync() in
// a 256-thread CTA; in the batched kernel's instance that adds two control warps (TMA producer / MMA issuer, chol_tc.cu)
// those warps never take part in the workers' barriers.
__device__ __forceinline__ void wsync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// block_sum (common.cuh) over the 256 worker threads
__device__ __forceinline__ float block_sum_w(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  wsync();
  if (lane == 0) red[w] = v;
  wsync();
  float r = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  wsync();
  r = red[0];
  wsync();
  return r;
}

struct Smem {
  float* As; float* Bs; float* Ct; float* LiT; float* Vs; float* z; float* al; float* z2;
  float* diagl; float* tmp; float* red; int* flag;
};

// ---------------------------------------------------------------------------------------------- GEMM micro-kernel
// acc[i][j] += sum_k A[ty*8+i][k] * Bt[tx*4+j][k], operands in shared memory as As[k][m], Bs[k][n].
__device__ __forceinline__ void mma_tile(float (&acc)[8][4], const float* __restrict__ As, int lda, const float* __restrict__ Bs,
                                         int ldb, int nk, int ty, int tx) {
#pragma unroll 4
  for (int kk = 0; kk < nk; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(As + kk * lda + ty * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(As + kk * lda + ty * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * ldb + tx * 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// A-operand loader.  PHASE_B reads U = X^T: zero below the block diagonal, the Dinv block on it, scratch above it.
template <bool PHASE_B>
__device__ __forceinline__ float4 load_a(const float* S, int ld, int gr, int row_end, int k, const float* dinv) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gr < row_end) {
    if (!PHASE_B) {
      v = *reinterpret_cast<const float4*>(S + (size_t)gr * ld + k);
    } else {
      const int mb = gr >> 6, kb = k >> 6;
      if (kb > mb) v = *reinterpret_cast<const float4*>(S + (size_t)gr * ld + k);
      else if (kb == mb) v = *reinterpret_cast<const float4*>(dinv + ((size_t)mb * NB + (gr & 63)) * NB + (k & 63));
    }
  }
  return v;
}

// acc += A[a_row0 + r, k_lo:k_hi] . Bm[b_row0 + n, k_lo:k_hi]^T   (r < 128, n < 64), all operands in the scratch.
template <bool PHASE_B>
__device__ void gemm_tn(float (&acc)[8][4], const float* S, int ld, int a_row0, int a_row_end, int b_row0, int k_lo, int k_hi,
                        const float* dinv, float* As, float* Bs) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int nk = (k_hi - k_lo) / BK;
  if (nk <= 0) return;
  const int arow0 = tid >> 2, arow1 = (tid + NT) >> 2, kq = (tid & 3) * 4;
  const int brow = tid >> 2;
  float4 ra0, ra1, rb;
  auto gload = [&](int k0) {
    ra0 = load_a<PHASE_B>(S, ld, a_row0 + arow0, a_row_end, k0 + kq, dinv);
    ra1 = load_a<PHASE_B>(S, ld, a_row0 + arow1, a_row_end, k0 + kq, dinv);
    rb = *reinterpret_cast<const float4*>(S + (size_t)(b_row0 + brow) * ld + k0 + kq);
  };
  auto sstore = [&](int buf) {
    float* a = As + buf * (BK * AS_LD);
    float* b = Bs + buf * (BK * BS_LD);
    a[(kq + 0) * AS_LD + arow0] = ra0.x; a[(kq + 1) * AS_LD + arow0] = ra0.y;
    a[(kq + 2) * AS_LD + arow0] = ra0.z; a[(kq + 3) * AS_LD + arow0] = ra0.w;
    a[(kq + 0) * AS_LD + arow1] = ra1.x; a[(kq + 1) * AS_LD + arow1] = ra1.y;
    a[(kq + 2) * AS_LD + arow1] = ra1.z; a[(kq + 3) * AS_LD + arow1] = ra1.w;
    b[(kq + 0) * BS_LD + brow] = rb.x; b[(kq + 1) * BS_LD + brow] = rb.y;
    b[(kq + 2) * BS_LD + brow] = rb.z; b[(kq + 3) * BS_LD + brow] = rb.w;
  };
  gload(k_lo);
  sstore(0);
  wsync();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(k_lo + (kt + 1) * BK);
    mma_tile(acc, As + buf * (BK * AS_LD), AS_LD, Bs + buf * (BK * BS_LD), BS_LD, BK, ty, tx);
    if (kt + 1 < nk) sstore(buf ^ 1);
    wsync();
  }
}

// ---------------------------------------------------------------------------------------------- diagonal block
// Unblocked right-looking Cholesky of the 64x64 block held column-major in Ct (Ct[c*CT_LD + r] = S[r][c], r >= c).
// Returns through *flag the first failing column (non-positive or NaN pivot, LAPACK's predicate) or -1.
template <int CLD>
__device__ void potrf64(float* Ct, float* diagl, int* flag, int col0) {
  const int tid = threadIdx.x;
  const int r = tid & 63, kg = tid >> 6;
  for (int c = 0; c < NB; ++c) {
    wsync();
    const float d = Ct[c * CLD + c];
    if (!(d > 0.f)) {
      if (tid == 0 && *flag < 0) *flag = col0 + c;
    }
    const float l = sqrtf(d);
    const float inv = 1.f / l;
    if (tid == c) diagl[c] = l;
    if (tid < NB && tid > c) Ct[c * CLD + tid] *= inv;
    wsync();
    const float lr = Ct[c * CLD + r];
    for (int k = c + 1 + kg; k < NB; k += 4)
      if (r >= k) Ct[k * CLD + r] = fmaf(-lr, Ct[c * CLD + k], Ct[k * CLD + r]);
  }
  wsync();
  if (tid < NB) Ct[tid * CLD + tid] = diagl[tid];
  wsync();
}

// LiT[k][c] = Linv[c][k], Linv = L^-1 for the 64x64 lower-triangular L in Ct, by recursive doubling:
//   Linv = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]  for block sizes s = 1, 2, ..., 32.  tmpbuf: 64*64 floats.
template <int CLD>
__device__ void trtri64(const float* Ct, float* LiT, float* tmpbuf) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * LI_LD; idx += NT) LiT[idx] = 0.f;
  wsync();
  if (tid < NB) LiT[tid * LI_LD + tid] = 1.f / Ct[tid * CLD + tid];
  wsync();
  for (int s = 1; s < NB; s <<= 1) {
    // step 1: Tm[r][c] = sum_k L_CA[r][k] Linv_AA[k][c], r,c in [0,s) per pair, k >= c
    for (int o = tid; o < 32 * s; o += NT) {
      const int pair = o / (s * s), rc = o % (s * s), rr = rc / s, cc = rc % s;
      const int a0 = pair * 2 * s, c0 = a0 + s;
      float acc = 0.f;
      for (int k = cc; k < s; ++k) acc = fmaf(Ct[(a0 + k) * CLD + c0 + rr], LiT[(a0 + cc) * LI_LD + a0 + k], acc);
      tmpbuf[(c0 + rr) * NB + a0 + cc] = acc;
    }
    wsync();
    // step 2: Linv_CA[r][c] = - sum_k Linv_CC[r][k] Tm[k][c], k <= r
    for (int o = tid; o < 32 * s; o += NT) {
      const int pair = o / (s * s), rc = o % (s * s), rr = rc / s, cc = rc % s;
      const int a0 = pair * 2 * s, c0 = a0 + s;
      float acc = 0.f;
      for (int k = 0; k <= rr; ++k) acc = fmaf(LiT[(c0 + k) * LI_LD + c0 + rr], tmpbuf[(c0 + k) * NB + a0 + cc], acc);
      LiT[(a0 + cc) * LI_LD + c0 + rr] = -acc;
    }
    wsync();
  }
}

// ---------------------------------------------------------------------------------------------- blocked diagonal block
// 64x64 factor + inverse as 4 panels of 16 columns.  The 16x16 pivot block is factored AND inverted by one warp
// entirely in registers (row / inverse column per lane, operands exchanged with shuffles: no shared-memory round
// trips on the 16-step dependency chain); the panel solve, the trailing update and the off-diagonal blocks of the
// inverse are small products spread over all 256 threads with rolled loops.  Measured: an earlier
// shared-memory-resident 32x32 version spent ~700 cycles per elimination step; this one ~100.
// D: row-major, stride RLD (lower triangle valid on entry; L in the lower triangle on exit, upper part untouched).
// LiT[k][c] = Linv[c][k] (stride RLD).  origd[64]: the original diagonal A_ii of these 64 rows (failure predicate).
// scratch: >= DIAG2_SCRATCH_FLOATS floats.  All CTA threads must call it.
// Failure predicate of the factorisation.  LAPACK's potrf (behind torch.linalg.cholesky_ex / psd_safe_cholesky) fails on a
// pivot <= 0 or NaN.  On an exactly singular matrix (e.g. a zero-volatility segment => duplicated rows of K) the computed
// pivot is rounding noise of either sign, so that test is a coin flip that depends on the summation order; here a pivot
// counts as failed when it is not larger than 8 eps times the ORIGINAL diagonal entry, which sends singular inputs down
// the jitter branch deterministically and never triggers on the well-posed matrices of this path (smallest pivot /
// diagonal >= 1e-4 for the noise-free rollout matrix at T = 8192).  NaNs fail the comparison as in LAPACK.
constexpr float PIVOT_RTOL = 8.f * 1.1920929e-07f;
constexpr int I16_LD = 20;
constexpr int DIAG2_SCRATCH_FLOATS = 4 * 16 * I16_LD + 16 * 52 + 48 * 20;   // I16 | XT | WT (diag64_block_v2)

template <int RLD>
__device__ __forceinline__ void pivot16_warp(float* D, float* LiT, float* I16p, float* diagl, const float* origd, int o, int lane,
                                             int& failc) {
  const int r = lane & 15;
  float a[16];
  {
    const float4* src = reinterpret_cast<const float4*>(D + (o + r) * RLD + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = src[q];
      a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
  }
  // Factor and invert in ONE unrolled loop: row c of the inverse (lane = column m: y[c] = Linv16[c][m]) only needs row
  // c of L, which is final once column c has been eliminated, so its shuffles / FMAs are independent of the next
  // elimination step and fill the latency gaps of that 16-step dependency chain (shfl -> rsqrt -> shfl -> fma).
  float y[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float d = __shfl_sync(0xffffffffu, a[c], c);
    if (!(d > PIVOT_RTOL * origd[o + c]) && failc < 0) failc = o + c;   // see PIVOT_RTOL
    float inv = rsqrtf(d);
    inv = inv * fmaf(-0.5f * d * inv, inv, 1.5f);  // one Newton step: 1/sqrt(d) to ~1 ulp
    const float l = d * inv;
    const float lrc = (r == c) ? l : a[c] * inv;
    a[c] = lrc;
#pragma unroll
    for (int k = c + 1; k < 16; ++k) a[k] = fmaf(-lrc, __shfl_sync(0xffffffffu, lrc, k), a[k]);
    float acc0 = (c == r) ? 1.f : 0.f, acc1 = 0.f;
#pragma unroll
    for (int t = 0; t < c; ++t) {
      const float lct = __shfl_sync(0xffffffffu, a[t], c);
      if (t & 1) acc1 = fmaf(-lct, y[t], acc1);
      else acc0 = fmaf(-lct, y[t], acc0);
    }
    y[c] = (acc0 + acc1) * inv;
  }
  __syncwarp();   // both half-warps have read their (duplicate) rows of D before the write-back below
  if (lane < 16) {
    float dl = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) dl = (k == r) ? a[k] : dl;
    diagl[o + r] = dl;
    float4* dst = reinterpret_cast<float4*>(D + (o + r) * RLD + o);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_float4(4 * q <= r ? a[4 * q] : 0.f, 4 * q + 1 <= r ? a[4 * q + 1] : 0.f, 4 * q + 2 <= r ? a[4 * q + 2] : 0.f,
                           4 * q + 3 <= r ? a[4 * q + 3] : 0.f);
  }
  if (lane < 16) {
#pragma unroll
    for (int k = 0; k < 16; ++k) I16p[k * I16_LD + r] = y[k];
    float4* dst = reinterpret_cast<float4*>(LiT + (o + r) * RLD + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
  }
}

// (A second version of this routine -- columns split between the two half-warps, column broadcast through a shared-memory
// line with broadcast LDS.128 instead of ten shuffles per step, MUFU.RCP on the chain and rsqrt + Newton off it -- was
// measured in round 2: correct, but 6.7 k instead of 3.4 k cycles per block: with in-order issue the STS -> __syncwarp ->
// LDS round trip of every step costs more than the shuffles it replaces.)
// (Also measured in round 2 and dropped: a split form -- the pivot warp only factors (55 % of the instructions), the panel below
// is solved by substitution, one thread per row, and the last warp inverts block p from shared memory while block p + 1 is
// being factored.  The pivot routine went from 3.4 k to 2.2 k cycles but the substitution costs 1.1 k against 0.46 k for the
// product with the inverse: diagonal block 22 k -> 20 k cycles alone, and no change at all on c2 / c3 / c5.)
__device__ __forceinline__ float dotn(const float* a, const float* b, int n) {  // n multiple of 4, 16-byte aligned
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int t = 0; t < n; t += 4) {
    const float4 u = *reinterpret_cast<const float4*>(a + t);
    const float4 v = *reinterpret_cast<const float4*>(b + t);
    s0 = fmaf(u.x, v.x, s0); s1 = fmaf(u.y, v.y, s1); s2 = fmaf(u.z, v.z, s2); s3 = fmaf(u.w, v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ float dot4(const float4 u, const float4 v, float acc) {
  acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
  return acc;
}
// 2 x 2 register tile of dot products: rows a0/a1 against rows b0/b1 (n floats, n multiple of 4, 16-byte aligned)
__device__ __forceinline__ void dot2x2(const float* a0, const float* a1, const float* b0, const float* b1, int n, float& r00, float& r01,
                                       float& r10, float& r11) {
  r00 = r01 = r10 = r11 = 0.f;
  for (int t = 0; t < n; t += 4) {
    const float4 u0 = *reinterpret_cast<const float4*>(a0 + t), u1 = *reinterpret_cast<const float4*>(a1 + t);
    const float4 v0 = *reinterpret_cast<const float4*>(b0 + t), v1 = *reinterpret_cast<const float4*>(b1 + t);
    r00 = dot4(u0, v0, r00); r01 = dot4(u0, v1, r01); r10 = dot4(u1, v0, r10); r11 = dot4(u1, v1, r11);
  }
}

// Shared-memory traffic of this routine was 31 % of the kernel's shared wavefronts with one output per thread
// (ncu, profiles/): every product below is register-tiled (2 x 2 dot tiles over rows {i, i+8} x {j, j+8}, 4 x 4 outer-
// product tiles for the trailing update) and the lane maps keep each 8-lane LDS.128 phase either on 8 distinct bank
// groups (row strides 68 and 20 floats) or on one broadcast address.
constexpr int XT_LD = 52;
constexpr int DIAG3_SCRATCH_FLOATS = 4 * 16 * I16_LD + 16 * XT_LD + 48 * 20;

// -DVOLT_PROFILE: clock64() deltas of the pivot warp's thread 0 in CTA 0 per sub-phase (read by tools/seg_probe.py)
#ifdef VOLT_PROFILE
static __device__ long long g_diag_prof[8];
#define DTICK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _n = clock64(); g_diag_prof[i] += _n - dlast; dlast = _n; } } while (0)
#else
#define DTICK(i) do { } while (0)
#endif

template <int RLD>
__device__ void diag64_block_v2(float* D, float* LiT, float* scratch, float* diagl, const float* origd, int* flag, int col0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef VOLT_PROFILE
  long long dlast = clock64();
#endif
  float* I16 = scratch;                          // 4 x (16 x I16_LD): row-major inverses of the pivot blocks
  float* XT = scratch + 4 * 16 * I16_LD;         // 16 x XT_LD: solved panel, transposed (XT[t][row] = X[row][t])
  float* WT = XT + 16 * XT_LD;                   // 48 x 20: per inverse block (16 x 20), WT[c][k] = W[k][c]
  const int ti = tid & 7, tj = (tid >> 3) & 7, tg = tid >> 6;   // 2 x 2 tile owner: rows {ti, ti+8}, cols {tj, tj+8}, group tg
  int failc = -1;
  // Look-ahead (round 2): the 16 x 16 pivot block of panel p + 1 only needs panel p's update of THAT block, so warp 0 applies
  // it itself right after the panel solve and goes on factoring while warps 1-7 write the panel back and apply the rest of the
  // trailing update.  The one-warp pivot chain (4 x 3.4 k cycles) then hides the trailing updates instead of alternating with them.
  if (warp == 0) pivot16_warp<RLD>(D, LiT, I16, diagl, origd, 0, lane, failc);
  else {
    for (int i = tid - 32; i < 64 * 64; i += NT - 32) {
      const int k = i >> 6, cc = i & 63;
      if ((k >> 4) != (cc >> 4)) LiT[k * RLD + cc] = 0.f;   // diagonal 16-blocks are written by the pivot warps
    }
  }
  DTICK(0);
  for (int p = 0; p < 3; ++p) {
    const int o = 16 * p;
    const int R = 48 - o;                        // rows below the pivot block
    wsync();                                     // pivot block p factored and inverted; trailing update of panel p - 1 complete
    DTICK(1);
    // ---- P2: panel solve X = S_panel Linv16^T, 16-row groups x (2 x 2 tiles); result kept transposed in XT
    if (tid < R * 4) {
      const float* a0 = D + (o + 16 + 16 * tg + ti) * RLD + o;
      const float* b0 = I16 + p * 16 * I16_LD + tj * I16_LD;
      float x00, x01, x10, x11;
      dot2x2(a0, a0 + 8 * RLD, b0, b0 + 8 * I16_LD, 16, x00, x01, x10, x11);
      const int r0 = 16 * tg + ti;
      XT[tj * XT_LD + r0] = x00; XT[(tj + 8) * XT_LD + r0] = x01;
      XT[tj * XT_LD + r0 + 8] = x10; XT[(tj + 8) * XT_LD + r0 + 8] = x11;
    }
    wsync();
    DTICK(2);
    if (warp == 0) {
      // next pivot block: D[o+16+i][o+16+j] -= X[i] . X[j], i, j < 16 (lane: row i, eight columns), then factor it
      const int i = lane & 15, jh = 8 * (lane >> 4);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
      for (int t = 0; t < 16; ++t) {
        const float a = XT[t * XT_LD + i];
        const float4 b0 = *reinterpret_cast<const float4*>(XT + t * XT_LD + jh), b1 = *reinterpret_cast<const float4*>(XT + t * XT_LD + jh + 4);
        acc[0] = fmaf(a, b0.x, acc[0]); acc[1] = fmaf(a, b0.y, acc[1]); acc[2] = fmaf(a, b0.z, acc[2]); acc[3] = fmaf(a, b0.w, acc[3]);
        acc[4] = fmaf(a, b1.x, acc[4]); acc[5] = fmaf(a, b1.y, acc[5]); acc[6] = fmaf(a, b1.z, acc[6]); acc[7] = fmaf(a, b1.w, acc[7]);
      }
      float4* dst = reinterpret_cast<float4*>(D + (o + 16 + i) * RLD + o + 16 + jh);
      float4 v0 = dst[0], v1 = dst[1];
      v0.x -= acc[0]; v0.y -= acc[1]; v0.z -= acc[2]; v0.w -= acc[3];
      v1.x -= acc[4]; v1.y -= acc[5]; v1.z -= acc[6]; v1.w -= acc[7];
      dst[0] = v0; dst[1] = v1;
      __syncwarp();
      pivot16_warp<RLD>(D, LiT, I16 + (p + 1) * 16 * I16_LD, diagl, origd, o + 16, lane, failc);
    } else {
      // ---- P3 (warps 1-7): write the panel back (row fastest) and apply the rest of the trailing update
      //      D[r][c] -= X[r].X[c] in 4 x 4 tiles that touch the lower triangle (the strictly-upper entries a diagonal tile also
      //      updates are never read); tiles 0-9 are the next pivot block (warp 0)
      const int t = tid - 32;
      if (t < R * 4) {
        const int rr = (t & 15) + 16 * (t >> 6), cq = ((t >> 4) & 3) * 4;
        *reinterpret_cast<float4*>(D + (o + 16 + rr) * RLD + o + cq) =
            make_float4(XT[cq * XT_LD + rr], XT[(cq + 1) * XT_LD + rr], XT[(cq + 2) * XT_LD + rr], XT[(cq + 3) * XT_LD + rr]);
      }
      const int nr = R >> 2, ntile = nr * (nr + 1) / 2;
      const int q = t + 10;
      if (q < ntile) {
        int ri = 0, ci = q;
        while (ci > ri) { ci -= ri + 1; ++ri; }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int tt = 0; tt < 16; ++tt) {
          const float4 a4 = *reinterpret_cast<const float4*>(XT + tt * XT_LD + 4 * ri);
          const float4 b4 = *reinterpret_cast<const float4*>(XT + tt * XT_LD + 4 * ci);
          const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4* dst = reinterpret_cast<float4*>(D + (o + 16 + 4 * ri + i) * RLD + o + 16 + 4 * ci);
          float4 v = *dst;
          v.x -= acc[i][0]; v.y -= acc[i][1]; v.z -= acc[i][2]; v.w -= acc[i][3];
          *dst = v;
        }
      }
    }
    DTICK(3);
  }
  wsync();
  if (tid == 0 && failc >= 0 && *flag < 0) *flag = col0 + failc;
  // ---- inverse: off-diagonal 16-blocks by block distance d = pb - qb.
  //   W[r][c] = sum_{k in [16 qb, 16 pb)} L[16 pb + r][k] Linv[k][16 qb + c]  (= D row . LiT row, both contiguous)
  //   Linv[16 pb + r][16 qb + c] = - sum_k Linv16_pb[r][k] W[k][c]
  for (int d = 1; d < 4; ++d) {
    const int nblk = 4 - d;
    const int qb = tg, pb = tg + d;
    if (tid < nblk * 64) {
      const float* a0 = D + (16 * pb + ti) * RLD + 16 * qb;
      const float* b0 = LiT + (16 * qb + tj) * RLD + 16 * qb;
      float w00, w01, w10, w11;
      dot2x2(a0, a0 + 8 * RLD, b0, b0 + 8 * RLD, 16 * d, w00, w01, w10, w11);
      float* wt = WT + (tg * 16 + tj) * 20 + ti;
      wt[0] = w00; wt[8 * 20] = w01; wt[8] = w10; wt[8 * 20 + 8] = w11;
    }
    wsync();
    if (tid < nblk * 64) {
      const float* a0 = I16 + pb * 16 * I16_LD + ti * I16_LD;
      const float* b0 = WT + (tg * 16 + tj) * 20;
      float l00, l01, l10, l11;
      dot2x2(a0, a0 + 8 * I16_LD, b0, b0 + 8 * 20, 16, l00, l01, l10, l11);
      float* lt = LiT + (16 * qb + tj) * RLD + 16 * pb + ti;
      lt[0] = -l00; lt[8 * RLD] = -l01; lt[8] = -l10; lt[8 * RLD + 8] = -l11;
    }
    wsync();
  }
  DTICK(4);
}

// ---------------------------------------------------------------------------------------------- generator
static __device__ __forceinline__ float gen_entry(const MllParams& p, int b, int i, int j, const float* Vs, float sc, float dadd) {
  if (i >= p.T || j >= p.T) return (i == j) ? 1.f : 0.f;
  float v;
  if (p.kind == KIND_VOL) v = Vs[min(i, j)];
  else if (p.kind == KIND_BM) v = sc * fminf(Vs[i], Vs[j]);
  else v = (i >= j) ? p.dense[(size_t)b * p.dense_bstride + (size_t)i * p.ldd + j]
                  : p.dense[(size_t)b * p.dense_bstride + (size_t)j * p.ldd + i];  // symmetric: only the lower triangle is read
  if (i == j) v += dadd;
  return v;
}


}  // namespace volt
