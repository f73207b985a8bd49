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

// ---------------------------------------------------------------------------------------------- generator
static __device__ __forceinline__ float gen_entry(const MllParams& p, int b, int i, int j, const float* Vs, float sc, float dadd) {
  if (i >= p.T || j >= p.T) return (i == j) ? 1.f : 0.f;
  float v;
  if (p.kind == KIND_VOL) v = Vs[min(i, j)];
  else if (p.kind == KIND_BM) v = sc * fminf(Vs[i], Vs[j]);
  else v = (i >= j) ? p.dense[(size_t)b * p.dense_bstride + (size_t)i * p.ldd + j] : 0.f;
  if (i == j) v += dadd;
  return v;
}


}  // namespace volt
