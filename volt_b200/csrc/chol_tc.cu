// Batched exact-MLL kernel, tensor-core variant: same algorithm and outputs as chol_batched.cu (left-looking blocked
// potrf + in-place trtri, one CTA per series), but every GEMM-shaped product runs on the 5th-generation tensor cores:
//
//   tcgen05.mma.cta_group::1.kind::tf32, M=128, N=64, K=8 per instruction, accumulators in TMEM (fp32),
//   operands in shared memory in the canonical K-major SWIZZLE_128B layout, read back with tcgen05.ld.32x32b.x32.
//
// fp32 accuracy is kept with the 3xTF32 split: every operand tile is stored twice, hi = a & 0xffffe000 (exactly
// representable in TF32) and lo = a - hi, and each product is issued as hi*hi + hi*lo + lo*hi (SURVEY.md section 7,
// hard part 1: plain TF32 misses the 1e-3 posterior bar on the noise-free rollout matrix).
//
// Products on the tensor pipe, per 128-row chunk (all "TN": both operands K-contiguous in the scratch):
//   Phase A   S   = A[rows, j] - L[rows, 0:j] L[j, 0:j]^T          K = 64 j
//             L[rows, j] = S Linv_jj^T                              K = 64   (TRSM by the inverted diagonal block)
//   Phase B   G^T = U[0:i, 0:i] L[i, 0:i]^T                         K = 64 i (U = (L^-1)^T, upper triangle of the scratch)
//             U[0:i, i] = -G^T Linv_ii^T                            K = 64
// SIMT work that remains: the 64x64 diagonal potrf / trtri (chol_dev.cuh), the hi/lo split while staging operand
// tiles, the forward substitution for z, and the reductions.
#include "chol_dev.cuh"

namespace volt {
namespace tc {

constexpr int CLD = NB + 4;                   // Ct / LiT row stride (floats)
constexpr uint32_t A_TILE = 128u * 128u;      // bytes of one 128-row x 32-float operand tile
constexpr uint32_t B_TILE = 64u * 128u;       // bytes of one 64-row x 32-float operand tile
// shared-memory map (byte offsets from a 1024-aligned base)
constexpr uint32_t X_AHI = 0, X_ALO = A_TILE, X_BHI = 2 * A_TILE, X_BLO = 2 * A_TILE + B_TILE;
constexpr uint32_t X_BYTES = 2 * A_TILE + 2 * B_TILE;           // 48 KB GEMM stage; aliased by P (hi|lo) and LiT|tmp
constexpr uint32_t X_LIT = 0, X_TMP = 64 * CLD * 4;             // LiT 17408 B, diag scratch 14336 B,
constexpr uint32_t X_STASH = 32768;                             // 16 KB stash of the chunk-0 panel rows  (<= 48 KB)
constexpr uint32_t L_OFF = X_BYTES;                             // Linv operand: hi k-tile0, hi k-tile1, lo k-tile0, lo k-tile1
constexpr uint32_t L_BYTES = 4 * B_TILE;                        // 32 KB
constexpr uint32_t CT_OFF = L_OFF + L_BYTES;                    // diagonal block D, row-major, stride CLD
constexpr uint32_t CT_BYTES = 64 * CLD * 4;
constexpr uint32_t VEC_OFF = CT_OFF + CT_BYTES;

constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B operand descriptor: 8-row groups are 1024 B apart (SBO), LBO unused, version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane (warp w reads lanes 32 (w%4) .. +31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// store one 16-byte chunk (4 consecutive k) of row `row` into a hi and a lo K-major SW128 tile
__device__ __forceinline__ void st_split(uint8_t* hi_tile, uint8_t* lo_tile, int row, int chunk, float4 v) {
  const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

struct Ctx {
  uint8_t* X;        // 48 KB stage / alias region
  uint8_t* Lr;       // 32 KB Linv operand (hi0, hi1, lo0, lo1)
  float* Ct;
  float* Vs; float* z; float* al; float* z2;
  float* diagl; float* tmp; float* red; int* flag;
  uint64_t* bar;
  uint32_t tmem;     // TMEM base (128 columns: acc0 = [0,64), acc1 = [64,128))
  uint32_t phase;    // parity of the next mbarrier completion to wait for
};

__device__ __forceinline__ void wait_mma(Ctx& c) {
  mbar_wait(c.bar, c.phase);
  c.phase ^= 1u;
}

// 3xTF32 product of one k-tile (32 floats): D (+)= A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T.  One thread issues.
__device__ __forceinline__ void issue_ktile(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, bool first) {
  const uint64_t dah = make_desc(a_hi), dal = make_desc(a_lo), dbh = make_desc(b_hi), dbl = make_desc(b_lo);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t adv = (uint64_t)(2 * ks);  // 32 bytes per K=8 step, in 16-byte units
    umma_tf32(tmem_d, dal + adv, dbh + adv, (first && ks == 0) ? 0u : 1u);
    umma_tf32(tmem_d, dah + adv, dbl + adv, 1u);
    umma_tf32(tmem_d, dah + adv, dbh + adv, 1u);
  }
}

// acc0 = A[a_row0 + r, k_lo:k_hi] . Bm[b_row0 + n, k_lo:k_hi]^T  on the tensor cores (r < 128, n < 64).
// Returns false when the k-range is empty (acc0 untouched).
template <bool PHASE_B>
__device__ bool gemm_tc(Ctx& c, const float* S, int ld, int a_row0, int a_row_end, int b_row0, int k_lo, int k_hi, const float* dinv) {
  const int tid = threadIdx.x;
  const int nk = (k_hi - k_lo) / 32;
  if (nk <= 0) return false;
  float4 ra[4], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + NT * i, row = idx >> 3, chunk = idx & 7;
      ra[i] = load_a<PHASE_B>(S, ld, a_row0 + row, a_row_end, k0 + chunk * 4, dinv);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i, row = idx >> 3, chunk = idx & 7;
      rb[i] = *reinterpret_cast<const float4*>(S + (size_t)(b_row0 + row) * ld + k0 + chunk * 4);
    }
  };
  gload(k_lo);
  for (int kt = 0; kt < nk; ++kt) {
    if (kt > 0) wait_mma(c);  // the previous k-tile's MMAs have consumed the stage
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + NT * i;
      st_split(c.X + X_AHI, c.X + X_ALO, idx >> 3, idx & 7, ra[i]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i;
      st_split(c.X + X_BHI, c.X + X_BLO, idx >> 3, idx & 7, rb[i]);
    }
    if (kt + 1 < nk) gload(k_lo + (kt + 1) * 32);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t xb = s_u32(c.X);
      issue_ktile(c.tmem, xb + X_AHI, xb + X_ALO, xb + X_BHI, xb + X_BLO, kt == 0);
      umma_commit(c.bar);
    }
  }
  wait_mma(c);
  tc_fence_after();
  return true;
}

// out = P . Linv^T where P (128 x 64, one row per (thread, column half)) is in registers `s`, Linv hi/lo already in c.Lr.
// Two K halves (the warps holding columns 0..31 stage first, then the warps holding 32..63); result in acc1 -> `o`.
__device__ void trsm_tc(Ctx& c, const float (&s)[32], float (&o)[32], int row, int half_id) {
  const int tid = threadIdx.x;
  const uint32_t xb = s_u32(c.X), lb = s_u32(c.Lr);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (half_id == half) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        st_split(c.X, c.X + A_TILE, row, ch, make_float4(s[4 * ch], s[4 * ch + 1], s[4 * ch + 2], s[4 * ch + 3]));
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_ktile(c.tmem + 64, xb, xb + A_TILE, lb + half * B_TILE, lb + (2 + half) * B_TILE, half == 0);
      umma_commit(c.bar);
    }
    wait_mma(c);
  }
  tc_fence_after();
  const int w = tid >> 5;
  tmem_ld32(c.tmem + ((uint32_t)(32 * (w & 3)) << 16) + 64u + (uint32_t)(half_id * 32), o);
  tc_fence_before();
}

// Linv operand (B of the TRSM product): B[n][k] = Linv[n][k] = LiT[k][n], LiT with row stride CLD (floats).
__device__ __forceinline__ void stage_linv_from_lit(Ctx& c, const float* LiT) {
  for (int q = threadIdx.x; q < 64 * 16; q += NT) {
    const int n = q >> 4, kc = q & 15;  // kc: 16-byte chunk over k = 0..63
    const int k = kc * 4;
    const float4 v = make_float4(LiT[(k + 0) * CLD + n], LiT[(k + 1) * CLD + n], LiT[(k + 2) * CLD + n], LiT[(k + 3) * CLD + n]);
    const int kt = kc >> 3;
    st_split(c.Lr + kt * B_TILE, c.Lr + (2 + kt) * B_TILE, n, kc & 7, v);
  }
}
// same from the global Dinv block (Dinv[m][k'] = Linv[k'][m]  ->  Linv[n][k] = Dinv[k][n])
__device__ __forceinline__ void stage_linv_from_dinv(Ctx& c, const float* D) {
  for (int q = threadIdx.x; q < 64 * 16; q += NT) {
    const int n = q & 63, kc = q >> 6;
    const int k = kc * 4;
    const float4 v = make_float4(D[(k + 0) * NB + n], D[(k + 1) * NB + n], D[(k + 2) * NB + n], D[(k + 3) * NB + n]);
    const int kt = kc >> 3;
    st_split(c.Lr + kt * B_TILE, c.Lr + (2 + kt) * B_TILE, n, kc & 7, v);
  }
}

// A-generator for one accumulator row: s[q] <- A[gr][gc0 + q] - s[q], q = 0..31.  Fast path (no identity padding, on-the-fly
// kernels): the 32 column values come from 8 broadcast LDS.128 of the staged prefix vector.
__device__ __forceinline__ void gen_sub_row32(const MllParams& p, int b, int gr, int gc0, const float* Vs, float sc, float dadd,
                                              float (&s)[32]) {
  if (p.kind != KIND_DENSE && p.T == p.Tp) {
    const float vr = Vs[gr];
    const bool vol = (p.kind == KIND_VOL);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 cv = *reinterpret_cast<const float4*>(Vs + gc0 + 4 * q4);
      const float c4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gc = gc0 + 4 * q4 + e;
        float v = vol ? ((gc <= gr) ? c4[e] : vr) : sc * fminf(vr, c4[e]);
        if (gc == gr) v += dadd;
        s[4 * q4 + e] = v - s[4 * q4 + e];
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 32; ++q) s[q] = gen_entry(p, b, gr, gc0 + q, Vs, sc, dadd) - s[q];
  }
}

// Coalesced store of a warp's 32 x 32 block (lane = row, 32 columns in registers): transposed through a private
// 32 x 36 float shared-memory tile so that every STG.128 writes 4 full 128-byte rows instead of 32 row fragments.
__device__ __forceinline__ void store_block32(float* xs, const float (&o)[32], float* gdst /* row 0, col 0 of the block */, int ld,
                                              int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(xs + lane * 36 + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i, ch = (lane & 7) * 4;
    *reinterpret_cast<float4*>(gdst + (size_t)r * ld + ch) = *reinterpret_cast<const float4*>(xs + r * 36 + ch);
  }
}

__global__ void __launch_bounds__(NT, 2) mll_batched_tc_kernel(MllParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  if ((s_u32(smem_raw) & 1023u) != 0u) __trap();  // SWIZZLE_128B operand tiles need a 1024-byte aligned base
  Ctx c;
  c.X = base;
  c.Lr = base + L_OFF;
  c.Ct = reinterpret_cast<float*>(base + CT_OFF);
  c.Vs = reinterpret_cast<float*>(base + VEC_OFF);
  c.z = c.Vs + p.Tp;
  c.al = c.z + p.Tp;
  const bool has2 = (p.resid2 != nullptr);   // the second right-hand side (rollout prep) gets its own vector
  c.z2 = has2 ? c.al + p.Tp : c.z;
  c.diagl = c.al + (has2 ? 2 : 1) * p.Tp;
  c.tmp = c.diagl + NB;
  c.red = c.tmp + 2 * NB;
  c.flag = reinterpret_cast<int*>(c.red + 32);
  c.bar = reinterpret_cast<uint64_t*>(c.red + 36);
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(c.red + 38);
  c.phase = 0;
  float* LiT = reinterpret_cast<float*>(c.X + X_LIT);
  float* tmpbuf = reinterpret_cast<float*>(c.X + X_TMP);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;   // accumulator row (TMEM lane) owned by this thread
  const int half_id = warp >> 2;            // which 32-column half of the 64 accumulator columns
  const int c0 = half_id * 32;
  const int T = p.T, Tp = p.Tp, nb = p.nb, ld = p.Tp;
  float* S = p.scratch + (size_t)blockIdx.x * Tp * Tp;
  float* dinv = p.dinv + (size_t)blockIdx.x * nb * NB * NB;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s_u32(s_tmem_p)) : "memory");
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
  const uint32_t t_lane = (uint32_t)(32 * (warp & 3)) << 16;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    for (int i = tid; i < Tp; i += NT) {
      float v = 0.f;
      if (i < T) {
        if (p.kind == KIND_VOL) v = p.V[(size_t)b * T + i];
        else if (p.kind == KIND_BM) v = p.x[i];
      }
      c.Vs[i] = v;
    }
    const float sc = (p.kind == KIND_BM) ? p.scale[(size_t)b * p.scale_stride] : 1.f;
    const float dadd0 = p.diag_add ? p.diag_add[(size_t)b * p.diag_stride] : 0.f;
    const float* rb = p.resid ? p.resid + (size_t)b * T : nullptr;
    const float* rb2 = p.resid2 ? p.resid2 + (size_t)b * T : nullptr;

    int fail = 0;
    float jit_used = 0.f, logdet_part = 0.f;
    for (int attempt = 0;; ++attempt) {
      const float dadd = dadd0 + jit_used;
      logdet_part = 0.f;
      if (tid == 0) *c.flag = -1;
      for (int i = tid; i < Tp; i += NT) { c.z[i] = 0.f; c.al[i] = 0.f; if (has2) c.z2[i] = 0.f; }
      __syncthreads();
      fail = 0;
      // =============================== Phase A
      for (int j = 0; j < nb; ++j) {
        const int R0 = j * NB;
        const int nch = (Tp - R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int r_base = R0 + ch * CM;
          const int gr = r_base + row;
          const bool have = gemm_tc<false>(c, S, ld, r_base, Tp, R0, 0, R0, nullptr);
          float s[32];
          if (have) {
            tmem_ld32(c.tmem + t_lane + (uint32_t)c0, s);
            tc_fence_before();
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) s[q] = 0.f;
          }
          if (gr < Tp) {
            gen_sub_row32(p, b, gr, R0 + c0, c.Vs, sc, dadd, s);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) s[q] = 0.f;
          }
          if (ch == 0) {
            // rows < 64 are the diagonal block (-> Ct); the other rows park their 32 values in the free tail of X so
            // that no accumulator registers stay live across the warp-level factorisation below
            float4* stash = reinterpret_cast<float4*>(c.X + X_STASH);
            const int slot = (row - NB) + NB * half_id;
            if (row < NB) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(c.Ct + row * CLD + c0 + 4 * q) = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) stash[q * 128 + slot] = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
            }
            __syncthreads();
            diag64_block_v2<CLD>(c.Ct, LiT, tmpbuf, c.diagl, c.flag, R0);
            if (tid < NB && R0 + tid < T) logdet_part += logf(c.diagl[tid]);
            for (int idx = tid; idx < NB * NB; idx += NT) {
              const int r = idx >> 6, cc = idx & 63;
              S[(size_t)(R0 + r) * ld + R0 + cc] = (cc <= r) ? c.Ct[r * CLD + cc] : 0.f;
              dinv[((size_t)j * NB + r) * NB + cc] = LiT[r * CLD + cc];
            }
            stage_linv_from_lit(c, LiT);
            if (rb) {
              const int cz = tid >> 2, part = tid & 3;
              float a1 = 0.f, a2 = 0.f;
              const float* Lrow = S + (size_t)(R0 + cz) * ld;
              for (int k = part * 4; k < R0; k += 16) {
                const float4 lv = *reinterpret_cast<const float4*>(Lrow + k);
                a1 = fmaf(lv.x, c.z[k], a1); a1 = fmaf(lv.y, c.z[k + 1], a1);
                a1 = fmaf(lv.z, c.z[k + 2], a1); a1 = fmaf(lv.w, c.z[k + 3], a1);
                if (rb2) {
                  a2 = fmaf(lv.x, c.z2[k], a2); a2 = fmaf(lv.y, c.z2[k + 1], a2);
                  a2 = fmaf(lv.z, c.z2[k + 2], a2); a2 = fmaf(lv.w, c.z2[k + 3], a2);
                }
              }
              a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
              a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
              a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
              a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
              if (part == 0) {
                c.tmp[cz] = ((R0 + cz < T) ? rb[R0 + cz] : 0.f) - a1;
                c.tmp[NB + cz] = ((rb2 && R0 + cz < T) ? rb2[R0 + cz] : 0.f) - a2;
              }
              __syncthreads();
              if (tid < (has2 ? 2 : 1) * NB) {
                const int cc = tid & 63, which = tid >> 6;
                float zz = 0.f;
                for (int k = 0; k <= cc; ++k) zz = fmaf(LiT[k * CLD + cc], c.tmp[which * NB + k], zz);
                (which ? c.z2 : c.z)[R0 + cc] = zz;
              }
            }
            if (row >= NB) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 v = stash[q * 128 + slot];
                s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q) s[q] = 0.f;
            }
            __syncthreads();  // LiT / tmp / stash (aliasing X) are dead from here on; Linv operand staged
          }
          float o[32];
          trsm_tc(c, s, o, row, half_id);
          {
            // warp-uniform: the warp's 32 rows start at a multiple of 32 and Tp, R0 are multiples of 64
            const int g0 = r_base + 32 * (warp & 3);
            if (!(ch == 0 && (warp & 3) < 2) && g0 < Tp)
              store_block32(reinterpret_cast<float*>(c.X) + warp * 1152, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
          }
          __syncthreads();
        }
      }
      const int fcol = *c.flag;
      __syncthreads();
      if (fcol < 0) break;
      fail = fcol + 1;
      if (attempt >= p.max_tries || !(p.jitter > 0.f)) break;
      jit_used = p.jitter * __powf(10.f, (float)attempt);
    }

    float tr_part = 0.f;
    if (p.do_inverse) {
      // =============================== Phase B
      for (int i = 0; i < nb; ++i) {
        const int R0 = i * NB;
        const float* Di = dinv + (size_t)i * NB * NB;
        stage_linv_from_dinv(c, Di);
        for (int idx = tid; idx < NB * NB; idx += NT) {
          const int m = idx >> 6, cc = idx & 63;
          if (R0 + m < T && R0 + cc < T) {
            const float v = Di[idx];
            tr_part = fmaf(v, v, tr_part);
          }
        }
        if (tid < NB) {
          float a = 0.f;
          for (int cc = tid; cc < NB; ++cc) a = fmaf(Di[tid * NB + cc], c.z[R0 + cc], a);
          c.al[R0 + tid] += a;
        }
        __syncthreads();
        const int nch = (R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int m_base = ch * CM;
          const int m = m_base + row;
          gemm_tc<true>(c, S, ld, m_base, R0, R0, m_base, R0, dinv);
          float s[32], o[32];
          tmem_ld32(c.tmem + t_lane + (uint32_t)c0, s);
          tc_fence_before();
#pragma unroll
          for (int q = 0; q < 32; ++q) s[q] = -s[q];
          trsm_tc(c, s, o, row, half_id);
          float hdot = 0.f;
          {
            const int g0 = m_base + 32 * (warp & 3);
            if (g0 < R0) store_block32(reinterpret_cast<float*>(c.X) + warp * 1152, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
          }
          if (m < R0) {
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              if (m < T && R0 + c0 + q < T) tr_part = fmaf(o[q], o[q], tr_part);
              dot = fmaf(o[q], c.z[R0 + c0 + q], dot);
            }
            if (half_id) c.tmp[row] = dot;   // the two column halves of a row are combined in a fixed order
            else hdot = dot;
          }
          __syncthreads();
          if (m < R0 && half_id == 0) c.al[m] += hdot + c.tmp[row];
          __syncthreads();
        }
      }
    }

    // =============================== reductions and outputs (identical to the SIMT kernel)
    float zz = 0.f, aa = 0.f, ar = 0.f, z22 = 0.f, z12 = 0.f;
    for (int i = tid; i < T; i += NT) {
      const float zi = c.z[i], ai = c.al[i], z2i = has2 ? c.z2[i] : 0.f;
      zz = fmaf(zi, zi, zz);
      z22 = fmaf(z2i, z2i, z22);
      z12 = fmaf(zi, z2i, z12);
      if (p.z_out) { p.z_out[((size_t)b * 2) * T + i] = zi; p.z_out[((size_t)b * 2 + 1) * T + i] = z2i; }
      aa = fmaf(ai, ai, aa);
      if (rb) ar = fmaf(ai, rb[i], ar);
      if (p.alpha && p.do_inverse) p.alpha[(size_t)b * T + i] = ai;
    }
    const float inv_quad = block_sum(zz, c.red);
    const float logdet = 2.f * block_sum(logdet_part, c.red);
    const float tr_inv = block_sum(tr_part, c.red);
    const float alal = block_sum(aa, c.red);
    const float alr = block_sum(ar, c.red);
    const float sz22 = block_sum(z22, c.red);
    const float sz12 = block_sum(z12, c.red);
    if (tid == 0) {
      if (p.scalars) {
        float* o = p.scalars + (size_t)b * NSCALARS;
        const float Tf = (float)T;
        o[0] = -0.5f * (inv_quad + logdet + Tf * 1.8378770664093453f) / Tf;
        o[1] = 0.5f * (alal - tr_inv) / Tf;
        o[2] = logdet; o[3] = inv_quad; o[4] = tr_inv; o[5] = alal; o[6] = alr; o[7] = jit_used;
        o[8] = sz22; o[9] = sz12;
        for (int q = 10; q < NSCALARS; ++q) o[q] = 0.f;
      }
      if (p.info) p.info[b] = fail;
    }
    if (p.L_out) {
      float* Lo = p.L_out + (size_t)b * p.L_bstride;
      for (int idx = tid; idx < T * T; idx += NT) {
        const int r = idx / T, cc = idx - r * T;
        Lo[(size_t)r * p.ldl + cc] = (cc <= r) ? S[(size_t)r * ld + cc] : 0.f;
      }
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(c.tmem) : "memory");
}

}  // namespace tc

int launch_mll_batched_tc(MllParams p, cudaStream_t st) {
  p.Tp = (p.T + NB - 1) / NB * NB;
  p.nb = p.Tp / NB;
  const size_t smem = tc::VEC_OFF + sizeof(float) * (size_t)(4 * p.Tp + NB + 2 * NB + 32 + 8);
  if (smem > 227 * 1024) {
    set_error("mll_batched_tc: T=%d needs %zu bytes of shared memory (max 227 KB)", p.T, smem);
    return VOLT_ERR_ARG;
  }
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    int s = check_cuda(cudaFuncSetAttribute(tc::mll_batched_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(mll_batched_tc_kernel)");
    if (s) return s;
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
  tc::mll_batched_tc_kernel<<<grid, NT, smem, st>>>(p);
  return check_cuda(cudaGetLastError(), "mll_batched_tc_kernel");
}

}  // namespace volt
