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
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(k_lo + (kt + 1) * BK);
    mma_tile(acc, As + buf * (BK * AS_LD), AS_LD, Bs + buf * (BK * BS_LD), BS_LD, BK, ty, tx);
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
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
    __syncthreads();
    const float d = Ct[c * CLD + c];
    if (!(d > 0.f)) {
      if (tid == 0 && *flag < 0) *flag = col0 + c;
    }
    const float l = sqrtf(d);
    const float inv = 1.f / l;
    if (tid == c) diagl[c] = l;
    if (tid < NB && tid > c) Ct[c * CLD + tid] *= inv;
    __syncthreads();
    const float lr = Ct[c * CLD + r];
    for (int k = c + 1 + kg; k < NB; k += 4)
      if (r >= k) Ct[k * CLD + r] = fmaf(-lr, Ct[c * CLD + k], Ct[k * CLD + r]);
  }
  __syncthreads();
  if (tid < NB) Ct[tid * CLD + tid] = diagl[tid];
  __syncthreads();
}

// LiT[k][c] = Linv[c][k], Linv = L^-1 for the 64x64 lower-triangular L in Ct, by recursive doubling:
//   Linv = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]  for block sizes s = 1, 2, ..., 32.  tmpbuf: 64*64 floats.
template <int CLD>
__device__ void trtri64(const float* Ct, float* LiT, float* tmpbuf) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * LI_LD; idx += NT) LiT[idx] = 0.f;
  __syncthreads();
  if (tid < NB) LiT[tid * LI_LD + tid] = 1.f / Ct[tid * CLD + tid];
  __syncthreads();
  for (int s = 1; s < NB; s <<= 1) {
    // step 1: Tm[r][c] = sum_k L_CA[r][k] Linv_AA[k][c], r,c in [0,s) per pair, k >= c
    for (int o = tid; o < 32 * s; o += NT) {
      const int pair = o / (s * s), rc = o % (s * s), rr = rc / s, cc = rc % s;
      const int a0 = pair * 2 * s, c0 = a0 + s;
      float acc = 0.f;
      for (int k = cc; k < s; ++k) acc = fmaf(Ct[(a0 + k) * CLD + c0 + rr], LiT[(a0 + cc) * LI_LD + a0 + k], acc);
      tmpbuf[(c0 + rr) * NB + a0 + cc] = acc;
    }
    __syncthreads();
    // step 2: Linv_CA[r][c] = - sum_k Linv_CC[r][k] Tm[k][c], k <= r
    for (int o = tid; o < 32 * s; o += NT) {
      const int pair = o / (s * s), rc = o % (s * s), rr = rc / s, cc = rc % s;
      const int a0 = pair * 2 * s, c0 = a0 + s;
      float acc = 0.f;
      for (int k = 0; k <= rr; ++k) acc = fmaf(LiT[(c0 + k) * LI_LD + c0 + rr], tmpbuf[(c0 + k) * NB + a0 + cc], acc);
      LiT[(a0 + cc) * LI_LD + c0 + rr] = -acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- warp-level diagonal block
// One warp factors and inverts the 64x64 diagonal block as a 2x2 grid of 32x32 blocks, one matrix row (or inverse
// column) per lane held in registers, operands broadcast through small shared-memory panels (no block barriers):
//   L11 = chol(S11); L21 = S21 L11^-T; S22 -= L21 L21^T; L22 = chol(S22);
//   Li11 = L11^-1, Li22 = L22^-1 (one column per lane, forward substitution); Li21 = -Li22 (L21 Li11).
// In/out: Ct (column-major, stride CLD; lower triangle valid on entry, L on exit), LiT[k][c] = Linv[c][k],
// diagl[r] = L[r][r]; tmp: >= 3*32*DW_LD + 128 floats of scratch.  All 32 lanes of the calling warp must be active.
constexpr int DW_LD = 36;  // row stride of the 32x32 broadcast panels (floats, 16-byte aligned rows)

#define VOLT_CBAR() asm volatile("" ::: "memory")

// sum_{t < N} row[t] * x[t] with four independent accumulators; `row` is a 16-byte aligned shared-memory row that
// every lane reads at the same address (broadcast).  N is a compile-time constant after unrolling.
template <int N>
__device__ __forceinline__ float dot_bcast(const float* row, const float (&x)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int t = 0; t < N; t += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + t);
    s0 = fmaf(v.x, x[t], s0);
    if (t + 1 < N) s1 = fmaf(v.y, x[t + 1], s1);
    if (t + 2 < N) s2 = fmaf(v.z, x[t + 2], s2);
    if (t + 3 < N) s3 = fmaf(v.w, x[t + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// right-looking Cholesky of a 32x32 block, lane = row, a[k] = S[lane][k] (k <= lane valid).  colbuf: 64 floats.
__device__ __forceinline__ void potrf32_warp(float (&a)[32], float* colbuf, int lane, int& failc, int off) {
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float d = __shfl_sync(0xffffffffu, a[c], c);
    if (!(d > 0.f) && failc < 0) failc = off + c;
    const float l = sqrtf(d);
    const float inv = 1.f / l;
    const float lrc = (lane == c) ? l : a[c] * inv;
    a[c] = lrc;
    float* cb = colbuf + (c & 1) * 32;
    cb[lane] = lrc;
    __syncwarp();
#pragma unroll
    for (int k4 = (c + 1) & ~3; k4 < 32; k4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(cb + k4);
      if (k4 > c) a[k4] = fmaf(-lrc, v.x, a[k4]);
      if (k4 + 1 > c) a[k4 + 1] = fmaf(-lrc, v.y, a[k4 + 1]);
      if (k4 + 2 > c) a[k4 + 2] = fmaf(-lrc, v.z, a[k4 + 2]);
      if (k4 + 3 > c) a[k4 + 3] = fmaf(-lrc, v.w, a[k4 + 3]);
    }
    VOLT_CBAR();
  }
}

// forward substitution with the row-major lower-triangular panel Lp (stride DW_LD) and reciprocal diagonal invd:
// x <- solution of (x' L^T = x) when UNIT_RHS == false (x holds the right-hand side), or column `lane` of L^-1 when
// UNIT_RHS == true (x is overwritten).
template <bool UNIT_RHS, int K>
struct FwdSub {
  static __device__ __forceinline__ void run(float (&x)[32], const float* Lp, const float* invd, int lane) {
    FwdSub<UNIT_RHS, K - 1>::run(x, Lp, invd, lane);
    constexpr int k = K - 1;
    const float rhs = UNIT_RHS ? ((k == lane) ? 1.f : 0.f) : x[k];
    const float acc = rhs - dot_bcast<k>(Lp + k * DW_LD, x);
    x[k] = acc * invd[k];
    VOLT_CBAR();
  }
};
template <bool UNIT_RHS>
struct FwdSub<UNIT_RHS, 0> {
  static __device__ __forceinline__ void run(float (&)[32], const float*, const float*, int) {}
};

// y[r] = sum_{t < (TRI ? r+1 : 32)} P[r][t] * x[t] for r = 0..31 (P row-major, stride DW_LD, broadcast reads)
template <bool TRI, int R>
struct MatVec {
  static __device__ __forceinline__ void run(float (&y)[32], const float* P, const float (&x)[32]) {
    MatVec<TRI, R - 1>::run(y, P, x);
    constexpr int r = R - 1;
    y[r] = dot_bcast<(TRI ? r + 1 : 32)>(P + r * DW_LD, x);
    VOLT_CBAR();
  }
};
template <bool TRI>
struct MatVec<TRI, 0> {
  static __device__ __forceinline__ void run(float (&)[32], const float*, const float (&)[32]) {}
};

template <int CLD>
__device__ void diag64_warp(float* Ct, float* LiT, float* tmp, float* diagl, int* flag, int col0) {
  const int lane = threadIdx.x & 31;
  float* P0 = tmp;                    // L11 (row-major), later L22
  float* P1 = tmp + 32 * DW_LD;       // L21
  float* P2 = tmp + 64 * DW_LD;       // Li22 (row-major)
  float* colbuf = tmp + 96 * DW_LD;   // 64 floats
  float* invd = colbuf + 64;          // 64 floats: 1 / L[r][r]
  int failc = -1;
  float a[32], x[32];

  // ---- L11 = chol(S11)   (lane = row)
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = Ct[k * CLD + lane];
  potrf32_warp(a, colbuf, lane, failc, 0);
  {
    float dl = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float v = (k <= lane) ? a[k] : 0.f;
      dl = (k == lane) ? v : dl;
      P0[lane * DW_LD + k] = v;
      Ct[k * CLD + lane] = v;
    }
    diagl[lane] = dl;
    invd[lane] = 1.f / dl;
  }
  __syncwarp();
  // ---- L21 = S21 L11^-T  (lane = row 32 + lane)
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = Ct[k * CLD + 32 + lane];
  FwdSub<false, 32>::run(x, P0, invd, lane);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    P1[lane * DW_LD + k] = x[k];
    Ct[k * CLD + 32 + lane] = x[k];
  }
  // ---- Li11 (lane = column m): a[k] = Linv11[k][m]
  FwdSub<true, 32>::run(a, P0, invd, lane);
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    *reinterpret_cast<float4*>(LiT + lane * CLD + k) = make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]);
    *reinterpret_cast<float4*>(LiT + (32 + lane) * CLD + k) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  // ---- S22 -= L21 L21^T ; L22 = chol(S22)   (lane = row 32 + lane; x = its L21 row)
  MatVec<false, 32>::run(a, P1, x);
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = Ct[(32 + k) * CLD + 32 + lane] - a[k];
  potrf32_warp(a, colbuf, lane, failc, 32);
  __syncwarp();
  {
    float dl = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float v = (k <= lane) ? a[k] : 0.f;
      dl = (k == lane) ? v : dl;
      P0[lane * DW_LD + k] = v;
      Ct[(32 + k) * CLD + 32 + lane] = v;
    }
    diagl[32 + lane] = dl;
    invd[32 + lane] = 1.f / dl;
  }
  __syncwarp();
  // ---- Li22 (lane = column m)
  FwdSub<true, 32>::run(a, P0, invd + 32, lane);
#pragma unroll
  for (int k = 0; k < 32; ++k) P2[k * DW_LD + lane] = a[k];
#pragma unroll
  for (int k = 0; k < 32; k += 4)
    *reinterpret_cast<float4*>(LiT + (32 + lane) * CLD + 32 + k) = make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]);
  __syncwarp();
  // ---- Li21 = -Li22 (L21 Li11)   (lane = column m): x = Li11[:, m]; a = L21 x; x = Li22 a
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(LiT + lane * CLD + k);
    x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
  }
  MatVec<false, 32>::run(a, P1, x);
  MatVec<true, 32>::run(x, P2, a);
#pragma unroll
  for (int k = 0; k < 32; k += 4)
    *reinterpret_cast<float4*>(LiT + lane * CLD + 32 + k) = make_float4(-x[k], -x[k + 1], -x[k + 2], -x[k + 3]);
  if (lane == 0 && failc >= 0 && *flag < 0) *flag = col0 + failc;
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- compact diagonal block
// Block-cooperative factor + inverse of the 64x64 diagonal block with ROLLED loops (small code: the fully unrolled
// warp version above is instruction-fetch bound) as a 2x2 grid of 32x32 blocks:
//   warp 0: L11 = chol(S11) | warp 0: L21 = S21 L11^-T, warp 1: Li11 | all: S22 -= L21 L21^T |
//   warp 0: L22 = chol(S22), warps 1-7: W^T = Li11^T-rows . L21-rows | warp 0: Li22 | all: Li21 = -Li22 W.
// D: row-major diagonal block, stride RLD (lower triangle valid on entry; L with a zeroed strict upper triangle in
// each 32x32 diagonal sub-block on exit).  LiT[k][c] = Linv[c][k] (stride RLD).  scratch: >= 2*32*36 + 32 + 64 floats.
// Must be called by all threads of the CTA (contains __syncthreads).
template <int RLD>
__device__ __forceinline__ void potrf32_rolled(float* D, int o, float* colbuf, float* invd, float* diagl, int lane, int& failc) {
  float* rowp = D + (o + lane) * RLD + o;
  for (int c = 0; c < 32; ++c) {
    const float d = D[(o + c) * RLD + o + c];
    if (!(d > 0.f) && failc < 0) failc = o + c;
    const float l = sqrtf(d);
    const float inv = 1.f / l;
    const float lrc = (lane == c) ? l : rowp[c] * inv;
    if (lane >= c) rowp[c] = lrc; else rowp[c] = 0.f;
    colbuf[lane] = lrc;
    if (lane == c) { diagl[o + c] = l; invd[o + c] = inv; }
    __syncwarp();
    for (int k4 = (c + 1) & ~3; k4 < 32; k4 += 4) {
      const float4 cb = *reinterpret_cast<const float4*>(colbuf + k4);
      float4 own = *reinterpret_cast<float4*>(rowp + k4);
      if (k4 > c) own.x = fmaf(-lrc, cb.x, own.x);
      if (k4 + 1 > c) own.y = fmaf(-lrc, cb.y, own.y);
      if (k4 + 2 > c) own.z = fmaf(-lrc, cb.z, own.z);
      own.w = fmaf(-lrc, cb.w, own.w);
      *reinterpret_cast<float4*>(rowp + k4) = own;
    }
    __syncwarp();
  }
}

// x[k] <- (rhs_k - sum_{t<k} L[k][t] x[t]) * invd[k], k = 0..31; L row-major (stride RLD, zero above the diagonal),
// x a private row per lane.  unit: rhs = e_lane (x must be zero-initialised), otherwise rhs = x[k] on entry.
template <int RLD>
__device__ __forceinline__ void fwdsub32_rolled(const float* L, float* x, const float* invd, bool unit, int lane) {
  for (int k = 0; k < 32; ++k) {
    const float* Lk = L + k * RLD;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int t4 = 0; t4 < k; t4 += 4) {
      const float4 l4 = *reinterpret_cast<const float4*>(Lk + t4);
      const float4 x4 = *reinterpret_cast<const float4*>(x + t4);
      a0 = fmaf(l4.x, x4.x, a0);
      if (t4 + 1 < k) a1 = fmaf(l4.y, x4.y, a1);
      if (t4 + 2 < k) a2 = fmaf(l4.z, x4.z, a2);
      if (t4 + 3 < k) a3 = fmaf(l4.w, x4.w, a3);
    }
    const float rhs = unit ? ((k == lane) ? 1.f : 0.f) : x[k];
    x[k] = (rhs - ((a0 + a1) + (a2 + a3))) * invd[k];
  }
}

__device__ __forceinline__ float dot32(const float* a, const float* b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int t = 0; t < 32; t += 4) {
    const float4 u = *reinterpret_cast<const float4*>(a + t);
    const float4 v = *reinterpret_cast<const float4*>(b + t);
    s0 = fmaf(u.x, v.x, s0); s1 = fmaf(u.y, v.y, s1); s2 = fmaf(u.z, v.z, s2); s3 = fmaf(u.w, v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

constexpr int DIAG_SCRATCH_FLOATS = 2 * 32 * 36 + 32 + 64;

template <int RLD>
__device__ void diag64_block(float* D, float* LiT, float* scratch, float* diagl, int* flag, int col0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* Li22r = scratch;              // 32 x 36 row-major Li22
  float* WT = scratch + 32 * 36;       // 32 x 36: WT[m][r] = (L21 Li11)[r][m]
  float* colbuf = WT + 32 * 36;        // 32
  float* invd = colbuf + 32;           // 64
  int failc = -1;
  // ---- S0: L11
  if (warp == 0) potrf32_rolled<RLD>(D, 0, colbuf, invd, diagl, lane, failc);
  else {
    for (int i = tid - 32; i < 64 * 64; i += NT - 32) LiT[(i >> 6) * RLD + (i & 63)] = 0.f;
    for (int i = tid - 32; i < 32 * 36; i += NT - 32) Li22r[i] = 0.f;
  }
  __syncthreads();
  // ---- S1: L21 (warp 0)  ||  Li11 (warp 1)
  if (warp == 0) fwdsub32_rolled<RLD>(D, D + (32 + lane) * RLD, invd, false, lane);
  else if (warp == 1) fwdsub32_rolled<RLD>(D, LiT + lane * RLD, invd, true, lane);
  __syncthreads();
  // ---- S2: S22 -= L21 L21^T (lower part only)
  {
    const int r = tid >> 3, k0 = (tid & 7) * 4;
    if (k0 <= r) {
      const float* xr = D + (32 + r) * RLD;
      float* out = D + (32 + r) * RLD + 32 + k0;
      float4 o4 = *reinterpret_cast<float4*>(out);
      o4.x -= dot32(xr, D + (32 + k0) * RLD);
      o4.y -= dot32(xr, D + (33 + k0) * RLD);
      o4.z -= dot32(xr, D + (34 + k0) * RLD);
      o4.w -= dot32(xr, D + (35 + k0) * RLD);
      *reinterpret_cast<float4*>(out) = o4;
    }
  }
  __syncthreads();
  // ---- S3: L22 (warp 0)  ||  WT (warps 1..7)
  if (warp == 0) potrf32_rolled<RLD>(D, 32, colbuf, invd, diagl, lane, failc);
  else {
    for (int i = tid - 32; i < 32 * 32; i += NT - 32) {
      const int m = i >> 5, r = i & 31;
      WT[m * 36 + r] = dot32(LiT + m * RLD, D + (32 + r) * RLD);
    }
  }
  __syncthreads();
  // ---- S4: Li22 (warp 0): column `lane` -> LiT row 32+lane (cols 32..63) and the row-major copy Li22r
  if (warp == 0) {
    float* y = LiT + (32 + lane) * RLD + 32;
    fwdsub32_rolled<RLD>(D + 32 * RLD + 32, y, invd + 32, true, lane);
    for (int k = 0; k < 32; ++k) Li22r[k * 36 + lane] = y[k];
    if (lane == 0 && failc >= 0 && *flag < 0) *flag = col0 + failc;
  }
  __syncthreads();
  // ---- S5: Li21 = -Li22 (L21 Li11): LiT[m][32 + r] = -sum_t Li22r[r][t] WT[m][t]
  for (int i = tid; i < 32 * 32; i += NT) {
    const int m = i >> 5, r = i & 31;
    LiT[m * RLD + 32 + r] = -dot32(Li22r + r * 36, WT + m * 36);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------- blocked diagonal block (v2)
// 64x64 factor + inverse as 4 panels of 16 columns.  The 16x16 pivot block is factored AND inverted by one warp
// entirely in registers (row / inverse column per lane, operands exchanged with shuffles: no shared-memory round
// trips on the 16-step dependency chain); the panel solve, the trailing update and the off-diagonal blocks of the
// inverse are small products spread over all 256 threads with rolled loops.  Measured: the 32x32 smem-resident
// version above spends ~700 cycles per elimination step; this one ~100.
// D: row-major, stride RLD (lower triangle valid on entry; L in the lower triangle on exit, upper part untouched).
// LiT[k][c] = Linv[c][k] (stride RLD).  scratch: >= DIAG2_SCRATCH_FLOATS floats.  All CTA threads must call it.
constexpr int I16_LD = 20;
constexpr int DIAG2_SCRATCH_FLOATS = 4 * 16 * I16_LD + 48 * 20 + 64;

template <int RLD>
__device__ __forceinline__ void pivot16_warp(float* D, float* LiT, float* I16p, float* diagl, int o, int lane, int& failc) {
  const int r = lane & 15;
  float a[16], inv16[16];
  {
    const float4* src = reinterpret_cast<const float4*>(D + (o + r) * RLD + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = src[q];
      a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float d = __shfl_sync(0xffffffffu, a[c], c);
    if (!(d > 0.f) && failc < 0) failc = o + c;
    float inv = rsqrtf(d);
    inv = inv * fmaf(-0.5f * d * inv, inv, 1.5f);  // one Newton step: 1/sqrt(d) to ~1 ulp
    const float l = d * inv;
    inv16[c] = inv;
    const float lrc = (r == c) ? l : a[c] * inv;
    a[c] = lrc;
#pragma unroll
    for (int k = c + 1; k < 16; ++k) a[k] = fmaf(-lrc, __shfl_sync(0xffffffffu, lrc, k), a[k]);
  }
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
  // inverse: lane = column m, y[k] = Linv16[k][m]
  float y[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float acc = (k == r) ? 1.f : 0.f;
#pragma unroll
    for (int t = 0; t < k; ++t) acc = fmaf(-__shfl_sync(0xffffffffu, a[t], k), y[t], acc);
    y[k] = acc * inv16[k];
  }
  if (lane < 16) {
#pragma unroll
    for (int k = 0; k < 16; ++k) I16p[k * I16_LD + r] = y[k];
    float4* dst = reinterpret_cast<float4*>(LiT + (o + r) * RLD + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
  }
}

__device__ __forceinline__ float dotn(const float* a, const float* b, int n) {  // n multiple of 4, 16-byte aligned
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int t = 0; t < n; t += 4) {
    const float4 u = *reinterpret_cast<const float4*>(a + t);
    const float4 v = *reinterpret_cast<const float4*>(b + t);
    s0 = fmaf(u.x, v.x, s0); s1 = fmaf(u.y, v.y, s1); s2 = fmaf(u.z, v.z, s2); s3 = fmaf(u.w, v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

template <int RLD>
__device__ void diag64_block_v2(float* D, float* LiT, float* scratch, float* diagl, int* flag, int col0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* I16 = scratch;                          // 4 x (16 x I16_LD): row-major inverses of the pivot blocks
  float* XP = scratch + 4 * 16 * I16_LD;         // 48 x 20: solved panel rows (row-major, 16 columns)
  int failc = -1;
  for (int p = 0; p < 4; ++p) {
    const int o = 16 * p;
    const int R = 48 - o;                        // rows below the pivot block
    // ---- P1: pivot block (warp 0); the other warps clear LiT once
    if (warp == 0) pivot16_warp<RLD>(D, LiT, I16 + p * 16 * I16_LD, diagl, o, lane, failc);
    else if (p == 0) {
      for (int i = tid - 32; i < 64 * 64; i += NT - 32) {
        const int k = i >> 6, cc = i & 63;
        if ((k >> 4) != (cc >> 4)) LiT[k * RLD + cc] = 0.f;   // diagonal 16-blocks are written by the pivot warps
      }
    }
    __syncthreads();
    if (R > 0) {
      // ---- P2: panel solve X = S_panel Linv16^T (thread per (row, 4 columns); rows read before anybody writes)
      const float* Ip = I16 + p * 16 * I16_LD;
      for (int task = tid; task < R * 4; task += NT) {
        const int rr = task >> 2, cq = (task & 3) * 4;
        const float* srow = D + (o + 16 + rr) * RLD + o;
        float4 out;
        out.x = dotn(srow, Ip + (cq + 0) * I16_LD, 16);
        out.y = dotn(srow, Ip + (cq + 1) * I16_LD, 16);
        out.z = dotn(srow, Ip + (cq + 2) * I16_LD, 16);
        out.w = dotn(srow, Ip + (cq + 3) * I16_LD, 16);
        *reinterpret_cast<float4*>(XP + rr * 20 + cq) = out;
      }
      __syncthreads();
      // ---- P3: write the panel back and apply the trailing update D[r][c] -= X[r].X[c] (c <= r)
      for (int task = tid; task < R * 4; task += NT) {
        const int rr = task >> 2, cq = (task & 3) * 4;
        *reinterpret_cast<float4*>(D + (o + 16 + rr) * RLD + o + cq) = *reinterpret_cast<const float4*>(XP + rr * 20 + cq);
      }
      for (int task = tid; task < R * (R / 4); task += NT) {
        const int rr = task / (R / 4), cq = (task - rr * (R / 4)) * 4;
        if (cq <= rr) {
          float4* dst = reinterpret_cast<float4*>(D + (o + 16 + rr) * RLD + o + 16 + cq);
          float4 v = *dst;
          const float* xr = XP + rr * 20;
          v.x -= dotn(xr, XP + (cq + 0) * 20, 16);
          v.y -= dotn(xr, XP + (cq + 1) * 20, 16);
          v.z -= dotn(xr, XP + (cq + 2) * 20, 16);
          v.w -= dotn(xr, XP + (cq + 3) * 20, 16);
          *dst = v;
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0 && failc >= 0 && *flag < 0) *flag = col0 + failc;
  // ---- inverse: off-diagonal 16-blocks by block distance d = pb - qb.
  //   W[r][c] = sum_{k in [16 qb, 16 pb)} L[16 pb + r][k] Linv[k][16 qb + c]  (= D row . LiT row, both contiguous)
  //   Linv[16 pb + r][16 qb + c] = - sum_k Linv16_pb[r][k] W[k][c]
  float* WT = XP;  // reuse: per block (16 x 20), WT[c][k] = W[k][c]; 3 blocks at most per level -> 48 x 20
  for (int d = 1; d < 4; ++d) {
    const int nblk = 4 - d;
    for (int task = tid; task < nblk * 256; task += NT) {
      const int bi = task >> 8, r = (task >> 4) & 15, cc = task & 15;
      const int qb = bi, pb = bi + d;
      WT[(bi * 16 + cc) * 20 + r] = dotn(D + (16 * pb + r) * RLD + 16 * qb, LiT + (16 * qb + cc) * RLD + 16 * qb, 16 * d);
    }
    __syncthreads();
    for (int task = tid; task < nblk * 256; task += NT) {
      const int bi = task >> 8, cc = (task >> 4) & 15, r = task & 15;
      const int qb = bi, pb = bi + d;
      LiT[(16 * qb + cc) * RLD + 16 * pb + r] = -dotn(I16 + pb * 16 * I16_LD + r * I16_LD, WT + (bi * 16 + cc) * 20, 16);
    }
    __syncthreads();
  }
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
