// Batched dense Cholesky / exact-MLL kernel, one CTA per series (persistent over the batch).
//
// Replaces, for B independent series of length T (fp32, dense, exact):
//   [GPyTorch] ExactMarginalLogLikelihood -> MultivariateNormal.log_prob -> psd_safe_cholesky / triangular solve /
//   log-det (call sites voltron/train_utils.py:89,136,249) and the autograd backward of the same chain
//   (train_utils.py:90,137,250), plus torch.linalg.cholesky_ex as used by voltron/rollout_utils.py:35.
//
// Algorithm per series (Tp = T rounded up to 64, padded with an identity block):
//   Phase A  left-looking blocked potrf, NB = 64 columns per step, rows processed in 128-row chunks:
//              S = A[rows, j] - L[rows, 0:j] L[j, 0:j]^T          (register-tiled fp32 GEMM, operands staged in smem)
//              L_jj = chol(S_jj) (shared memory), Linv_jj = L_jj^-1 (recursive doubling, shared memory)
//              L[rows>j, j] = S[rows>j] Linv_jj^T                  (same GEMM micro-kernel)
//              z_j = Linv_jj (r_j - L[j, 0:j] z_0:j)               (forward substitution fused into the sweep)
//   Phase B  in-place row-oriented trtri into the UPPER triangle of the same scratch (U = X^T, X = L^-1):
//              U[0:i, i] = -(U[0:i, 0:i] L[i, 0:i]^T) Linv_ii^T    (both products "TN": K-contiguous operands)
//            accumulating tr(A^-1) = ||X||_F^2 and alpha = X^T z on the fly.
//   Outputs  MLL = -1/2 (z.z + 2 sum log L_ii + T log 2pi)/T, dMLL/dnoise = 1/2 (alpha.alpha - tr A^-1)/T, alpha, info.
// The matrix A is generated on the fly (vol kernel: V[min(i,j)], BM kernel: s*min(x_i,x_j)) or read from a dense
// buffer, so in the fused path K never exists in HBM; L/U live in a per-CTA scratch that stays L2-resident.
#include "params.cuh"

#include <cmath>

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
__device__ void potrf64(float* Ct, float* diagl, int* flag, int col0) {
  const int tid = threadIdx.x;
  const int r = tid & 63, kg = tid >> 6;
  for (int c = 0; c < NB; ++c) {
    __syncthreads();
    const float d = Ct[c * CT_LD + c];
    if (!(d > 0.f)) {
      if (tid == 0 && *flag < 0) *flag = col0 + c;
    }
    const float l = sqrtf(d);
    const float inv = 1.f / l;
    if (tid == c) diagl[c] = l;
    if (tid < NB && tid > c) Ct[c * CT_LD + tid] *= inv;
    __syncthreads();
    const float lr = Ct[c * CT_LD + r];
    for (int k = c + 1 + kg; k < NB; k += 4)
      if (r >= k) Ct[k * CT_LD + r] = fmaf(-lr, Ct[c * CT_LD + k], Ct[k * CT_LD + r]);
  }
  __syncthreads();
  if (tid < NB) Ct[tid * CT_LD + tid] = diagl[tid];
  __syncthreads();
}

// LiT[k][c] = Linv[c][k], Linv = L^-1 for the 64x64 lower-triangular L in Ct, by recursive doubling:
//   Linv = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]  for block sizes s = 1, 2, ..., 32.  tmpbuf: 64*64 floats.
__device__ void trtri64(const float* Ct, float* LiT, float* tmpbuf) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * LI_LD; idx += NT) LiT[idx] = 0.f;
  __syncthreads();
  if (tid < NB) LiT[tid * LI_LD + tid] = 1.f / Ct[tid * CT_LD + tid];
  __syncthreads();
  for (int s = 1; s < NB; s <<= 1) {
    // step 1: Tm[r][c] = sum_k L_CA[r][k] Linv_AA[k][c], r,c in [0,s) per pair, k >= c
    for (int o = tid; o < 32 * s; o += NT) {
      const int pair = o / (s * s), rc = o % (s * s), rr = rc / s, cc = rc % s;
      const int a0 = pair * 2 * s, c0 = a0 + s;
      float acc = 0.f;
      for (int k = cc; k < s; ++k) acc = fmaf(Ct[(a0 + k) * CT_LD + c0 + rr], LiT[(a0 + cc) * LI_LD + a0 + k], acc);
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
__device__ __forceinline__ float gen_entry(const MllParams& p, int b, int i, int j, const float* Vs, float sc, float dadd) {
  if (i >= p.T || j >= p.T) return (i == j) ? 1.f : 0.f;
  float v;
  if (p.kind == KIND_VOL) v = Vs[min(i, j)];
  else if (p.kind == KIND_BM) v = sc * fminf(Vs[i], Vs[j]);
  else v = (i >= j) ? p.dense[(size_t)b * p.dense_bstride + (size_t)i * p.ldd + j] : 0.f;
  if (i == j) v += dadd;
  return v;
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NT, 2) mll_batched_kernel(MllParams p) {
  extern __shared__ __align__(16) float smem_f[];
  Smem sm;
  sm.As = smem_f;
  sm.Bs = sm.As + 2 * BK * AS_LD;
  sm.Ct = sm.Bs + 2 * BK * BS_LD;
  sm.LiT = sm.Ct + NB * CT_LD;
  sm.Vs = sm.LiT + NB * LI_LD;
  sm.z = sm.Vs + p.Tp;
  sm.al = sm.z + p.Tp;
  sm.z2 = sm.al + p.Tp;
  sm.diagl = sm.z2 + p.Tp;
  sm.tmp = sm.diagl + NB;
  sm.red = sm.tmp + 2 * NB;
  sm.flag = reinterpret_cast<int*>(sm.red + 32);

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int T = p.T, Tp = p.Tp, nb = p.nb, ld = p.Tp;
  float* S = p.scratch + (size_t)blockIdx.x * Tp * Tp;
  float* dinv = p.dinv + (size_t)blockIdx.x * nb * NB * NB;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    // ---- per-series vectors
    for (int i = tid; i < Tp; i += NT) {
      float v = 0.f;
      if (i < T) {
        if (p.kind == KIND_VOL) v = p.V[(size_t)b * T + i];
        else if (p.kind == KIND_BM) v = p.x[i];
      }
      sm.Vs[i] = v;
    }
    const float sc = (p.kind == KIND_BM) ? p.scale[(size_t)b * p.scale_stride] : 1.f;
    const float dadd0 = p.diag_add ? p.diag_add[(size_t)b * p.diag_stride] : 0.f;
    const float* rb = p.resid ? p.resid + (size_t)b * T : nullptr;
    const float* rb2 = p.resid2 ? p.resid2 + (size_t)b * T : nullptr;

    int fail = 0;          // 1-based index of the failing leading minor, 0 = success
    float jit_used = 0.f;
    float logdet_part = 0.f;
    for (int attempt = 0;; ++attempt) {
      const float dadd = dadd0 + jit_used;
      logdet_part = 0.f;
      if (tid == 0) *sm.flag = -1;
      for (int i = tid; i < Tp; i += NT) { sm.z[i] = 0.f; sm.al[i] = 0.f; sm.z2[i] = 0.f; }
      __syncthreads();
      fail = 0;
      // =============================== Phase A: potrf + forward substitution
      for (int j = 0; j < nb; ++j) {
        const int R0 = j * NB;
        const int nch = (Tp - R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int r_base = R0 + ch * CM;
          float acc[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
          gemm_tn<false>(acc, S, ld, r_base, Tp, R0, 0, R0, nullptr, sm.As, sm.Bs);
          // epilogue: Ct[c][r] = gen - acc
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int gc = R0 + tx * 4 + q;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int gr = r_base + ty * 8 + i;
              o[i] = (gr < Tp) ? gen_entry(p, b, gr, gc, sm.Vs, sc, dadd) - acc[i][q] : 0.f;
            }
            float* dst = sm.Ct + (tx * 4 + q) * CT_LD + ty * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
          }
          __syncthreads();
          if (ch == 0) {
            potrf64(sm.Ct, sm.diagl, sm.flag, R0);
            trtri64(sm.Ct, sm.LiT, sm.As);
            if (tid < NB && R0 + tid < T) logdet_part += logf(sm.diagl[tid]);
            // L_jj -> scratch (lower part), LiT -> dinv[j]
            for (int idx = tid; idx < NB * NB; idx += NT) {
              const int r = idx >> 6, c = idx & 63;
              S[(size_t)(R0 + r) * ld + R0 + c] = (c <= r) ? sm.Ct[c * CT_LD + r] : 0.f;
              dinv[((size_t)j * NB + r) * NB + c] = sm.LiT[r * LI_LD + c];
            }
            if (rb) {
              // z_j = Linv_jj (r_j - L[j,0:j] z)   (and the same for the optional second right-hand side)
              const int c = tid >> 2, part = tid & 3;
              float s = 0.f, s2 = 0.f;
              const float* Lrow = S + (size_t)(R0 + c) * ld;
              for (int k = part * 4; k < R0; k += 16) {
                const float4 lv = *reinterpret_cast<const float4*>(Lrow + k);
                s = fmaf(lv.x, sm.z[k], s); s = fmaf(lv.y, sm.z[k + 1], s);
                s = fmaf(lv.z, sm.z[k + 2], s); s = fmaf(lv.w, sm.z[k + 3], s);
                if (rb2) {
                  s2 = fmaf(lv.x, sm.z2[k], s2); s2 = fmaf(lv.y, sm.z2[k + 1], s2);
                  s2 = fmaf(lv.z, sm.z2[k + 2], s2); s2 = fmaf(lv.w, sm.z2[k + 3], s2);
                }
              }
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
              s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
              if (part == 0) {
                sm.tmp[c] = ((R0 + c < T) ? rb[R0 + c] : 0.f) - s;
                sm.tmp[NB + c] = ((rb2 && R0 + c < T) ? rb2[R0 + c] : 0.f) - s2;
              }
              __syncthreads();
              if (tid < 2 * NB) {
                const int cc = tid & 63, which = tid >> 6;
                float zz = 0.f;
                for (int k = 0; k <= cc; ++k) zz = fmaf(sm.LiT[k * LI_LD + cc], sm.tmp[which * NB + k], zz);
                (which ? sm.z2 : sm.z)[R0 + cc] = zz;
              }
            }
            __syncthreads();
          }
          // panel rows: out = Ct^T . LiT  (= S Linv^T); in chunk 0 the first 64 rows are the diagonal block itself
          if (!(ch == 0 && ty < 8)) {
            float out[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int q = 0; q < 4; ++q) out[i][q] = 0.f;
            mma_tile(out, sm.Ct, CT_LD, sm.LiT, LI_LD, NB, ty, tx);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int gr = r_base + ty * 8 + i;
              if (gr < Tp)
                *reinterpret_cast<float4*>(S + (size_t)gr * ld + R0 + tx * 4) = make_float4(out[i][0], out[i][1], out[i][2], out[i][3]);
            }
          }
          __syncthreads();
        }
      }
      const int fcol = *sm.flag;
      __syncthreads();
      if (fcol < 0) break;
      fail = fcol + 1;
      if (attempt >= p.max_tries || !(p.jitter > 0.f)) break;
      jit_used = p.jitter * __powf(10.f, (float)attempt);
    }

    float tr_part = 0.f;
    if (p.do_inverse) {
      // =============================== Phase B: U = (L^-1)^T, tr(A^-1), alpha = X^T z
      for (int i = 0; i < nb; ++i) {
        const int R0 = i * NB;
        for (int idx = tid; idx < NB * NB; idx += NT) {
          const int r = idx >> 6, c = idx & 63;
          sm.LiT[r * LI_LD + c] = dinv[((size_t)i * NB + r) * NB + c];
        }
        __syncthreads();
        // diagonal block X_ii = Linv_ii: X[c][m] = LiT[m][c]
        for (int idx = tid; idx < NB * NB; idx += NT) {
          const int m = idx >> 6, c = idx & 63;
          if (R0 + m < T && R0 + c < T) {
            const float v = sm.LiT[m * LI_LD + c];
            tr_part = fmaf(v, v, tr_part);
          }
        }
        if (tid < NB) {
          float a = 0.f;
          for (int c = tid; c < NB; ++c) a = fmaf(sm.LiT[tid * LI_LD + c], sm.z[R0 + c], a);
          sm.al[R0 + tid] += a;
        }
        const int nch = (R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int m_base = ch * CM;
          float acc[8][4];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[ii][q] = 0.f;
          gemm_tn<true>(acc, S, ld, m_base, R0, R0, m_base, R0, dinv, sm.As, sm.Bs);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* dst = sm.Ct + (tx * 4 + q) * CT_LD + ty * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(-acc[0][q], -acc[1][q], -acc[2][q], -acc[3][q]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(-acc[4][q], -acc[5][q], -acc[6][q], -acc[7][q]);
          }
          __syncthreads();
          float out[8][4];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii)
#pragma unroll
            for (int q = 0; q < 4; ++q) out[ii][q] = 0.f;
          mma_tile(out, sm.Ct, CT_LD, sm.LiT, LI_LD, NB, ty, tx);
          const float4 zv = *reinterpret_cast<const float4*>(sm.z + R0 + tx * 4);
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const int m = m_base + ty * 8 + ii;
            float dot = 0.f;
            if (m < R0) {
              *reinterpret_cast<float4*>(S + (size_t)m * ld + R0 + tx * 4) = make_float4(out[ii][0], out[ii][1], out[ii][2], out[ii][3]);
              if (m < T) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (R0 + tx * 4 + q < T) tr_part = fmaf(out[ii][q], out[ii][q], tr_part);
              }
              dot = out[ii][0] * zv.x + out[ii][1] * zv.y + out[ii][2] * zv.z + out[ii][3] * zv.w;
            }
            // reduce over the 16 tx lanes that share this row
            dot += __shfl_xor_sync(0xffffffffu, dot, 8);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            if (tx == 0 && m < R0) sm.al[m] += dot;
          }
          __syncthreads();
        }
        __syncthreads();
      }
    }

    // =============================== reductions and outputs
    float zz = 0.f, aa = 0.f, ar = 0.f, z22 = 0.f, z12 = 0.f;
    for (int i = tid; i < T; i += NT) {
      const float zi = sm.z[i], ai = sm.al[i], z2i = sm.z2[i];
      zz = fmaf(zi, zi, zz);
      z22 = fmaf(z2i, z2i, z22);
      z12 = fmaf(zi, z2i, z12);
      if (p.z_out) { p.z_out[((size_t)b * 2) * T + i] = zi; p.z_out[((size_t)b * 2 + 1) * T + i] = z2i; }
      aa = fmaf(ai, ai, aa);
      if (rb) ar = fmaf(ai, rb[i], ar);
      if (p.alpha && p.do_inverse) p.alpha[(size_t)b * T + i] = ai;
    }
    const float inv_quad = block_sum(zz, sm.red);
    const float logdet = 2.f * block_sum(logdet_part, sm.red);
    const float tr_inv = block_sum(tr_part, sm.red);
    const float alal = block_sum(aa, sm.red);
    const float alr = block_sum(ar, sm.red);
    const float sz22 = block_sum(z22, sm.red);
    const float sz12 = block_sum(z12, sm.red);
    if (tid == 0) {
      if (p.scalars) {
        float* o = p.scalars + (size_t)b * 16;
        const float Tf = (float)T;
        o[0] = -0.5f * (inv_quad + logdet + Tf * 1.8378770664093453f) / Tf;
        o[1] = 0.5f * (alal - tr_inv) / Tf;
        o[2] = logdet; o[3] = inv_quad; o[4] = tr_inv; o[5] = alal; o[6] = alr; o[7] = jit_used;
        o[8] = sz22; o[9] = sz12;
        for (int q = 10; q < 16; ++q) o[q] = 0.f;
      }
      if (p.info) p.info[b] = fail;
    }
    if (p.L_out) {
      float* Lo = p.L_out + (size_t)b * p.L_bstride;
      for (int idx = tid; idx < T * T; idx += NT) {
        const int r = idx / T, c = idx - r * T;
        Lo[(size_t)r * p.ldl + c] = (c <= r) ? S[(size_t)r * ld + c] : 0.f;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- host launcher
static size_t mll_smem_bytes(int Tp) {
  return sizeof(float) * (size_t)(2 * BK * AS_LD + 2 * BK * BS_LD + NB * CT_LD + NB * LI_LD + 4 * Tp + NB + 2 * NB + 32 + 4);
}

int launch_mll_batched(MllParams p, cudaStream_t st) {
  p.Tp = (p.T + NB - 1) / NB * NB;
  p.nb = p.Tp / NB;
  const size_t smem = mll_smem_bytes(p.Tp);
  if (smem > 227 * 1024) {
    set_error("mll_batched: T=%d needs %zu bytes of shared memory (max 227 KB)", p.T, smem);
    return VOLT_ERR_ARG;
  }
  static bool attr_done = false;
  static size_t attr_smem = 0;
  if (!attr_done || smem > attr_smem) {
    int s = check_cuda(cudaFuncSetAttribute(mll_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(mll_batched_kernel)");
    if (s) return s;
    attr_done = true;
    attr_smem = smem;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > p.B) grid = p.B;
  if (grid < 1) grid = 1;
  const size_t per_cta = (size_t)p.Tp * p.Tp + (size_t)p.nb * NB * NB;
  void* ws = nullptr;
  int s = get_workspace(per_cta * grid * sizeof(float), &ws, 0);
  if (s) return s;
  p.scratch = reinterpret_cast<float*>(ws);
  p.dinv = p.scratch + (size_t)grid * p.Tp * p.Tp;
  mll_batched_kernel<<<grid, NT, smem, st>>>(p);
  return check_cuda(cudaGetLastError(), "mll_batched_kernel");
}

}  // namespace volt
