// tcgen05 / TMEM building blocks shared by the batched (chol_tc.cu) and the multi-CTA large-matrix (chol_large.cu)
// Cholesky kernels: K-major SWIZZLE_128B operand staging with the 3xTF32 hi/lo split, UMMA issue, TMEM read-back.
#pragma once
#include "chol_dev.cuh"

namespace volt {
namespace tc {

constexpr int CLD = NB + 4;                   // Ct / LiT row stride (floats)
constexpr uint32_t A_TILE = 128u * 128u;      // bytes of one 128-row x 32-float operand tile
constexpr uint32_t B_TILE = 64u * 128u;       // bytes of one 64-row x 32-float operand tile
// shared-memory map (byte offsets from a 1024-aligned base)
constexpr uint32_t X_B0 = A_TILE;   // GEMM stage: raw A tile (16 KB) | B hi/lo buffer 0 (16 KB) | B hi/lo buffer 1 (16 KB)
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
// One lane of a fully converged warp.  The MMA / TMA issue sections run as `if (warp_uniform == k) { if (elect_one()) ... }`
// with operands derived from warp-uniform values: ptxas then keeps descriptors and addresses in uniform registers and emits
// one UTCHMMA per tcgen05.mma.  Issued from a thread-divergent branch (`if (tid == 0)`) every operand sits in a vector
// register and each instruction is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY "waterfall" loop -- measured: ~95 cycles
// of issue time per MMA, i.e. the issuing thread, not the tensor pipe, paced the GEMM loops.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0u;
}
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ uint32_t make_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or a time limit expires; with the default (short)
// limit a waiting warp re-issues the instruction continuously -- 13 % (register-staged instances) to 26 % (control-warp
// instance) of all executed instructions were these retries, taken from the issue slots of the warps that had work.  The
// explicit limit (suspendTimeHint, ns) keeps a waiter asleep until the mbarrier wakes it.
#ifndef VOLT_MBAR_HINT_NS
#define VOLT_MBAR_HINT_NS 200000
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(s_u32(bar)),
      "r"(parity), "r"((uint32_t)VOLT_MBAR_HINT_NS)
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
  uint32_t phase;    // bit h = parity of the next completion of mbarrier h to wait for
};

__device__ __forceinline__ void wait_mma2(Ctx& c, int h) {
  mbar_wait(c.bar + h, (c.phase >> h) & 1u);
  c.phase ^= (1u << h);
}
__device__ __forceinline__ void wait_mma(Ctx& c) { wait_mma2(c, 0); }

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

// 32 consecutive TMEM columns of this thread's lane <- registers (tcgen05.st, the mirror image of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A operand read from tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}

// TMEM column map of a CTA (256 columns allocated): accumulators and the TRSM A operand
constexpr uint32_t TM_ACC0 = 0, TM_ACC1 = 64, TM_PHI = 128, TM_PLO = 192, TM_COLS = 256;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// acc0 = A[a_row0 + r, k_lo:k_hi] . Bm[b_row0 + n, k_lo:k_hi]^T  on the tensor cores (r < 128, n < 64).
//
// The loop is bound by the shared-memory port when both operands are read from shared memory (an M128 x N64 x K8 SS
// MMA reads 6 KB per 32 ideal cycles and the 3-pass split reads every operand three times), so the A operand goes
// through TENSOR MEMORY instead: per 32-float k-tile the CTA stores the raw A tile once in shared memory (coalesced
// global loads), every thread reads back its own row / k-half (the tcgen05.st 32x32b layout: lane = row), splits it
// (hi = raw bits: the tensor core ignores the low 13 mantissa bits; lo = a - trunc(a)) and writes both to TMEM; only
// the 64-row B operand is staged hi/lo in shared memory and read by the MMAs (TS form).  A (in TMEM) and B (in
// shared memory) are double-buffered, so the MMAs of tile kt overlap the staging of tile kt+1; mbarrier h guards the
// reuse of buffer h (tcgen05.commit).  Returns false when the k-range is empty (acc0 untouched).
template <bool PHASE_B>
__device__ bool gemm_tc(Ctx& c, const float* S, int ld, int a_row0, int a_row_end, int b_row0, int k_lo, int k_hi, const float* dinv,
                        const float* SB = nullptr) {
  if (SB == nullptr) SB = S;  // B operand rows come from a second matrix in the large-matrix inverse sweep
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int nk = (k_hi - k_lo) / 32;
  if (nk <= 0) return false;
  const int row = 32 * (w & 3) + lane, half_id = w >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
  uint8_t* RAW = c.X;                                   // 16 KB raw A tile, rows of 128 B, 16-byte chunks XOR-swizzled
  const int wu = uniform_warp_id();                     // MMA issue from one elected lane of warp 0, uniform operands (elect_one)
  const uint32_t tmem_u = make_uniform(c.tmem), xb = s_u32(c.X);
  float4 ra[4], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
      ra[i] = load_a<PHASE_B>(S, ld, a_row0 + r, a_row_end, k0 + chunk * 4, dinv);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
      rb[i] = *reinterpret_cast<const float4*>(SB + (size_t)(b_row0 + r) * ld + k0 + chunk * 4);
    }
  };
  gload(k_lo);
  for (int kt = 0; kt < nk; ++kt) {
    const int h = kt & 1;
    uint8_t* BH = c.X + X_B0 + h * (2 * B_TILE);
    uint8_t* BL = BH + B_TILE;
    if (kt >= 2) wait_mma2(c, h);   // MMA group kt-2 has consumed TMEM / shared-memory buffer h
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
      *reinterpret_cast<float4*>(RAW + r * 128 + ((chunk ^ (r & 7)) << 4)) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i;
      st_split(BH, BL, idx >> 3, idx & 7, rb[i]);
    }
    if (kt + 1 < nk) gload(k_lo + (kt + 1) * 32);
    wsync();
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
      tmem_st16(c.tmem + lane_base + TM_PHI + (uint32_t)(32 * h + 16 * half_id), hi);
      tmem_st16(c.tmem + lane_base + TM_PLO + (uint32_t)(32 * h + 16 * half_id), lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    wsync();
    if (wu == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dbh = make_desc(xb + X_B0 + h * (2 * B_TILE)), dbl = make_desc(xb + X_B0 + h * (2 * B_TILE) + B_TILE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(2 * ks);
          const uint32_t ah = tmem_u + TM_PHI + 32 * h + 8 * ks, al = tmem_u + TM_PLO + 32 * h + 8 * ks;
          umma_tf32_ts(tmem_u + TM_ACC0, al, dbh + adv, (kt == 0 && ks == 0) ? 0u : 1u);
          umma_tf32_ts(tmem_u + TM_ACC0, ah, dbl + adv, 1u);
          umma_tf32_ts(tmem_u + TM_ACC0, ah, dbh + adv, 1u);
        }
        umma_commit(c.bar + h);
      }
      __syncwarp();
    }
  }
  wait_mma2(c, (nk - 1) & 1);
  if (nk >= 2) wait_mma2(c, (nk - 2) & 1);
  tc_fence_after();
  return true;
}

// out = P . Linv^T where P (128 x 64, one row per (thread, column half)) is in registers `s`, Linv hi/lo already in c.Lr.
// P is exactly in the layout tensor memory wants (lane = row), so it goes registers -> TMEM with tcgen05.st (raw fp32 as
// the hi operand: the tensor core ignores the low 13 mantissa bits; lo = s - trunc(s)) and the product runs in the
// TS form: no shared-memory staging of P and only the 2 KB B operand is read from shared memory per MMA.
static __device__ void trsm_tc(Ctx& c, const float (&s)[32], float (&o)[32], int row, int half_id) {
  const int tid = threadIdx.x, w = tid >> 5;
  const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
  {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      hi[q] = __float_as_uint(s[q]);
      lo[q] = __float_as_uint(s[q] - __uint_as_float(hi[q] & 0xffffe000u));
    }
    tmem_st32(c.tmem + lane_base + TM_PHI + (uint32_t)(half_id * 32), hi);
    tmem_st32(c.tmem + lane_base + TM_PLO + (uint32_t)(half_id * 32), lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  wsync();
  if (uniform_warp_id() == 0) {
    tc_fence_after();
    const uint32_t tmem_u = make_uniform(c.tmem);
    if (elect_one()) {
      const uint32_t lb = s_u32(c.Lr);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int kt = ks >> 2;
        const uint64_t adv = (uint64_t)(2 * (ks & 3));
        const uint64_t dbh = make_desc(lb + kt * B_TILE) + adv, dbl = make_desc(lb + (2 + kt) * B_TILE) + adv;
        const uint32_t ah = tmem_u + TM_PHI + 8 * ks, al = tmem_u + TM_PLO + 8 * ks;
        umma_tf32_ts(tmem_u + TM_ACC1, al, dbh, ks == 0 ? 0u : 1u);
        umma_tf32_ts(tmem_u + TM_ACC1, ah, dbl, 1u);
        umma_tf32_ts(tmem_u + TM_ACC1, ah, dbh, 1u);
      }
      umma_commit(c.bar);
    }
    __syncwarp();
  }
  wait_mma(c);
  tc_fence_after();
  tmem_ld32(c.tmem + lane_base + TM_ACC1 + (uint32_t)(half_id * 32), o);
  tc_fence_before();
}

// ---------------------------------------------------------------------------------------------------------------------
// "Three CTAs per SM" building blocks (series short enough for the small shared-memory map: T <= 832).  The kernel is latency bound (ncu: issue slots 36 %, L1
// data pipe ~35 %, tensor pipe 20 % with two resident CTAs), so the lever is more resident CTAs: 128 TMEM columns and
// ~71 KB of shared memory per CTA instead of 256 / 106 KB.  The price is single-buffered GEMM stages (the MMAs of a
// k-tile are waited for before the next tile is staged -- the other two CTAs fill the gap) and a two-pass TRSM.
//   TMEM:  [0,64) accumulator (GEMM result, then the TRSM result, and the parking place of the chunk-0 panel rows
//          during the diagonal factorisation) | [64,96) A hi | [96,128) A lo
//   smem:  Y: raw A tile 16 KB | B hi 8 KB | B lo 8 KB  (aliased by LiT | diag scratch and by the store tiles)
//          Linv operand 32 KB (aliased by the diagonal block D) | vectors
constexpr uint32_t T3_ACC = 0, T3_HI = 64, T3_LO = 96, T3_COLS = 128;
constexpr uint32_t Y_BH = A_TILE, Y_BL = A_TILE + B_TILE, Y_BYTES = A_TILE + 2 * B_TILE;   // 32 KB
constexpr uint32_t Y_LIT = 0, Y_TMP = 64 * CLD * 4;                                         // + 12288 B scratch <= 32 KB
constexpr uint32_t Y_L_OFF = Y_BYTES, Y_CT_OFF = Y_L_OFF, Y_VEC_OFF = Y_L_OFF + L_BYTES;
static_assert(Y_TMP + DIAG2_SCRATCH_FLOATS * 4 <= Y_BYTES, "diag scratch must fit the stage region");

template <bool PHASE_B>
__device__ bool gemm_tc1(Ctx& c, const float* S, int ld, int a_row0, int a_row_end, int b_row0, int k_lo, int k_hi, const float* dinv,
                         const float* SB = nullptr) {
  if (SB == nullptr) SB = S;  // B operand rows from a second matrix (large-matrix inverse sweep)
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int nk = (k_hi - k_lo) / 32;
  if (nk <= 0) return false;
  const int row = 32 * (w & 3) + lane, half_id = w >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
  uint8_t* RAW = c.X;
  uint8_t* BH = c.X + Y_BH;
  uint8_t* BL = c.X + Y_BL;
  const int wu = uniform_warp_id();                     // MMA issue from one elected lane of warp 0, uniform operands (elect_one)
  const uint32_t tmem_u = make_uniform(c.tmem), xb = s_u32(c.X);
  float4 ra[4], rb[2];
  auto gload_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
      rb[i] = *reinterpret_cast<const float4*>(SB + (size_t)(b_row0 + r) * ld + k0 + chunk * 4);
    }
  };
  auto gload_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
      ra[i] = load_a<PHASE_B>(S, ld, a_row0 + r, a_row_end, k0 + chunk * 4, dinv);
    }
  };
  gload_b(k_lo);
  gload_a(k_lo);
  for (int kt = 0; kt < nk; ++kt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // the raw tile is not an MMA operand: refill it while the previous MMAs still run
      const int idx = tid + NT * i, r = idx >> 3, chunk = idx & 7;
      *reinterpret_cast<float4*>(RAW + r * 128 + ((chunk ^ (r & 7)) << 4)) = ra[i];
    }
    if (kt >= 1) wait_mma(c);       // MMA group kt-1 has consumed the TMEM stage and B hi/lo
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i;
      st_split(BH, BL, idx >> 3, idx & 7, rb[i]);
    }
    if (kt + 1 < nk) {
      gload_b(k_lo + (kt + 1) * 32);
      gload_a(k_lo + (kt + 1) * 32);
    }
    wsync();
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
      tmem_st16(c.tmem + lane_base + T3_HI + (uint32_t)(16 * half_id), hi);
      tmem_st16(c.tmem + lane_base + T3_LO + (uint32_t)(16 * half_id), lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    wsync();
    if (wu == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dbh = make_desc(xb + Y_BH), dbl = make_desc(xb + Y_BL);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(2 * ks);
          const uint32_t ah = tmem_u + T3_HI + 8 * ks, al = tmem_u + T3_LO + 8 * ks;
          umma_tf32_ts(tmem_u + T3_ACC, al, dbh + adv, (kt == 0 && ks == 0) ? 0u : 1u);
          umma_tf32_ts(tmem_u + T3_ACC, ah, dbl + adv, 1u);
          umma_tf32_ts(tmem_u + T3_ACC, ah, dbh + adv, 1u);
        }
        umma_commit(c.bar);
      }
      __syncwarp();
    }
  }
  wait_mma(c);
  tc_fence_after();
  return true;
}

// out = P . Linv^T in two K = 32 passes: the warps holding columns 32 pass .. 32 pass + 31 of P put their hi / lo rows into
// the 64-column TMEM stage; the result accumulates in T3_ACC (the GEMM result it replaces is already in registers).
static __device__ void trsm_tc1(Ctx& c, const float (&s)[32], float (&o)[32], int row, int half_id) {
  const int tid = threadIdx.x, w = tid >> 5;
  const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
  const uint32_t lb = s_u32(c.Lr);
  const int wu = uniform_warp_id();
  const uint32_t tmem_u = make_uniform(c.tmem);
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    if (half_id == pass) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        hi[q] = __float_as_uint(s[q]);
        lo[q] = __float_as_uint(s[q] - __uint_as_float(hi[q] & 0xffffe000u));
      }
      tmem_st32(c.tmem + lane_base + T3_HI, hi);
      tmem_st32(c.tmem + lane_base + T3_LO, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    wsync();
    if (wu == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(2 * ks);
          const uint64_t dbh = make_desc(lb + pass * B_TILE) + adv, dbl = make_desc(lb + (2 + pass) * B_TILE) + adv;
          const uint32_t ah = tmem_u + T3_HI + 8 * ks, al = tmem_u + T3_LO + 8 * ks;
          umma_tf32_ts(tmem_u + T3_ACC, al, dbh, (pass == 0 && ks == 0) ? 0u : 1u);
          umma_tf32_ts(tmem_u + T3_ACC, ah, dbl, 1u);
          umma_tf32_ts(tmem_u + T3_ACC, ah, dbh, 1u);
        }
        umma_commit(c.bar);
      }
      __syncwarp();
    }
    wait_mma(c);
    tc_fence_after();
  }
  tmem_ld32(c.tmem + lane_base + T3_ACC + (uint32_t)(half_id * 32), o);
  tc_fence_before();
}

// store_block32 with a 32 x 20 float tile per warp (two passes of 16 columns): fits the 32 KB stage region
__device__ __forceinline__ void store_block32_2p(float* xs, const float (&o)[32], float* gdst, int ld, int lane) {
#pragma unroll
  for (int hcol = 0; hcol < 2; ++hcol) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(xs + lane * 20 + 4 * q) =
          make_float4(o[16 * hcol + 4 * q], o[16 * hcol + 4 * q + 1], o[16 * hcol + 4 * q + 2], o[16 * hcol + 4 * q + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (lane >> 2) + 8 * i, ch = (lane & 3) * 4;
      *reinterpret_cast<float4*>(gdst + (size_t)r * ld + 16 * hcol + ch) = *reinterpret_cast<const float4*>(xs + r * 20 + ch);
    }
    __syncwarp();
  }
}

// Linv operand (B of the TRSM product): B[n][k] = Linv[n][k] = LiT[k][n], LiT with row stride CLD (floats).
__device__ __forceinline__ void stage_linv_from_lit(Ctx& c, const float* LiT) {
  for (int q = threadIdx.x; q < 64 * 16; q += NT) {
    const int n = q & 63, kc = q >> 6;  // kc: 16-byte chunk over k = 0..63; n fastest: conflict-free LiT reads and tile stores
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

// ---------------------------------------------------------------------------------------------------------------------
// TMA-fed GEMM loop (the "W" instance of the batched kernel: two CTAs per SM, 256 TMEM columns).
//
// The register-staged loops above spend ~200 instructions per thread and k-tile on moving operands (LDG, address and
// bounds arithmetic, STS of the raw tile, a barrier, LDS, split, tcgen05.st): the loop is bound by instruction issue, not by
// the tensor pipe.  Here the raw fp32 operand tiles are brought in by the TMA unit (cp.async.bulk.tensor.2d, one
// elected thread, completion on an mbarrier) into a ring of shared-memory slots, 16 floats of K at a time:
//     slot s:  A raw 128 rows x 64 B | B raw 64 rows x 64 B      (SWIZZLE_64B, the layout the tensor map writes and the
//                                                                  UMMA descriptor reads)
// The B tile is used in place as the "hi" MMA operand (the tensor core ignores the low 13 mantissa bits of a raw fp32
// operand); threads only compute the "lo" tiles: every thread reads 32 bytes of its accumulator row of A, splits them and
// writes hi / lo to tensor memory (TS-form MMA, as above), and one 16-byte chunk of B, whose lo part goes to a second
// shared tile.  TMA of k-tile g + RING - 1 is issued as soon as the MMAs of k-tile g - 1 have released its slot, so up to
// RING - 1 tiles are in flight while one is being consumed.  Per k-tile and thread: 3 LDS.128, 1 STS.128, 24 ALU,
// 2 tcgen05.st, 1 barrier.
constexpr int W_RING = 4;
constexpr uint32_t HA_TILE = 128u * 64u, HB_TILE = 64u * 64u;      // bytes: 128 x 16 floats, 64 x 16 floats
constexpr uint32_t W_SLOT = HA_TILE + HB_TILE;                     // 12 KB
constexpr uint32_t W_BL = W_RING * W_SLOT;                         // two B lo tiles after the ring
constexpr int W_BLN = 4;                                           // B lo tiles / TMEM A stages: one per ring slot
constexpr uint32_t W_BYTES = W_BL + W_BLN * HB_TILE;               // 64 KB; aliased by LiT | diag scratch and the store tiles
constexpr uint32_t W_L_OFF = W_BYTES, W_CT_OFF = W_L_OFF, W_VEC_OFF = W_L_OFF + L_BYTES;   // D aliases the Linv operand
static_assert(X_TMP + DIAG2_SCRATCH_FLOATS * 4 <= W_BYTES && 8 * 1152 * 4 <= W_BYTES, "diag scratch / store tiles must fit the ring region");

struct TmaPipe {
  uint64_t* full;    // [W_RING] TMA landed
  uint64_t* done;    // [W_RING] MMAs of the tile that used the slot have completed
  uint64_t* ready;   // [W_RING] (control-warp instance) the workers have split the tile: one arrival per worker warp
  uint64_t* ringfree;   // (control-warp instance) the workers no longer use the ring region as scratch (diagonal block, preamble)
  uint64_t* depready;   // (control-warp instance) the global writes of the previous block step are visible to the async proxy
  uint32_t g;        // running k-tile counter of this CTA (slot = g % W_RING, use = g / W_RING)
  uint32_t rf_n, dep_n;   // producer side: phases of ringfree / depready consumed so far
#ifdef VOLT_PROFILE
  long long prof[8];
#endif
};

// K-major SWIZZLE_64B operand descriptor: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint32_t swz64(int row, int chunk) { return (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint32_t dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(s_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// generic-proxy writes -> async proxy (TMA): global (panel stores that a later tensor-map load reads) and this CTA's shared
// memory (ring slots that served as scratch)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// acc0 = A[a_row0 + r, k_lo:k_hi] . B[b_row0 + n, k_lo:k_hi]^T (r < 128, n < 64), both operands rows of this CTA's scratch
// square (tensor-map row = sq_row0 + local row).  PHASE_B: A = U = X^T, whose entries below the block diagonal are zero (the
// scratch holds L there) -- masked when the tile is split; its diagonal blocks were copied into the scratch beforehand.
#ifdef VOLT_PROFILE
static __device__ long long g_tma_prof[8];
#define WTICK(i) do { if (threadIdx.x == 0) { const long long _n = clock64(); tp.prof[i] += _n - wlast; wlast = _n; } } while (0)
#else
#define WTICK(i) do { } while (0)
#endif
template <bool PHASE_B>
__device__ bool gemm_tma(Ctx& c, TmaPipe& tp, const void* tmA, const void* tmB, int sq_row0, int a_row0, int a_row_end, int b_row0, int k_lo,
                         int k_hi) {
  const int nk = (k_hi - k_lo) / 16;
  if (nk <= 0) return false;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int row = 32 * (w & 3) + lane, half_id = w >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
  const uint32_t xb = s_u32(c.X);
  // writes of earlier phases (panel stores, diagonal blocks) must be visible to the async proxy before any TMA reads them,
  // and the ring region may have been used as scratch by the epilogue
#ifdef VOLT_PROFILE
  long long wlast = clock64();
#endif
  fence_proxy_async_global();
  fence_async_smem();
  wsync();
  WTICK(0);
  const uint32_t g0 = tp.g;
  // Two elected lanes in different warps share the asynchronous work of a k-tile: one lane of warp 0 issues the MMAs, one
  // lane of warp 4 waits for the slot of the tile RING - 1 ahead to be released and issues its TMA.  Nobody else waits for MMA
  // completions inside the loop: TMEM stage / B lo tile g % 3 was last read by the MMAs of tile g - 3, whose completion
  // thread 128 observed (wait on done[g - 3], issued in iteration g - 2) before the barrier of tile g - 1 -- which every
  // thread passes before it writes stage g % 3 again.
  auto issue_tma = [&](int kt) {       // one elected thread
    const uint32_t g = g0 + (uint32_t)kt, s = g % W_RING;
    mbar_expect_tx(tp.full + s, W_SLOT);
    tma_load_2d(tmA, xb + s * W_SLOT, tp.full + s, k_lo + 16 * kt, sq_row0 + a_row0);
    tma_load_2d(tmB, xb + s * W_SLOT + HA_TILE, tp.full + s, k_lo + 16 * kt, sq_row0 + b_row0);
  };
  const int wu = uniform_warp_id();
  const uint32_t tmem_u = make_uniform(c.tmem);
  if (wu == 4) {
    if (elect_one()) {
      const int pre = nk < W_RING - 1 ? nk : W_RING - 1;
      for (int kt = 0; kt < pre; ++kt) issue_tma(kt);
    }
    __syncwarp();
  }
  const bool row_ok = (a_row0 + row) < a_row_end;
  const int mb = (a_row0 + row) >> 6;
  for (int kt = 0; kt < nk; ++kt) {
    const uint32_t g = g0 + (uint32_t)kt, s = g % W_RING, ts = g % W_BLN;
    const uint8_t* RAW = c.X + s * W_SLOT;
    const uint8_t* BH = RAW + HA_TILE;
    uint8_t* BL = c.X + W_BL + ts * HB_TILE;
    WTICK(1);
    mbar_wait(tp.full + s, (g / W_RING) & 1u);
    WTICK(2);
    {
      uint32_t hi[8], lo[8];
      const bool live = row_ok && !(PHASE_B && ((k_lo + 16 * kt) >> 6) < mb);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(RAW + swz64(row, 2 * half_id + q));
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t u = live ? __float_as_uint(e[j]) : 0u;
          hi[4 * q + j] = u;
          lo[4 * q + j] = __float_as_uint(__uint_as_float(u) - __uint_as_float(u & 0xffffe000u));
        }
      }
      tmem_st8(c.tmem + lane_base + TM_PHI + (uint32_t)(16 * ts + 8 * half_id), hi);
      tmem_st8(c.tmem + lane_base + TM_PLO + (uint32_t)(16 * ts + 8 * half_id), lo);
      const uint32_t off = swz64(tid >> 2, tid & 3);
      const float4 b = *reinterpret_cast<const float4*>(BH + off);
      float4 l;
      l.x = b.x - __uint_as_float(__float_as_uint(b.x) & 0xffffe000u);
      l.y = b.y - __uint_as_float(__float_as_uint(b.y) & 0xffffe000u);
      l.z = b.z - __uint_as_float(__float_as_uint(b.z) & 0xffffe000u);
      l.w = b.w - __uint_as_float(__float_as_uint(b.w) & 0xffffe000u);
      *reinterpret_cast<float4*>(BL + off) = l;
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    WTICK(4);
    fence_async_smem();
    tc_fence_before();
    wsync();
    WTICK(5);
    if (wu == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dbh = make_desc64(xb + s * W_SLOT + HA_TILE), dbl = make_desc64(xb + W_BL + ts * HB_TILE);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t adv = (uint64_t)(2 * ks);
          const uint32_t ah = tmem_u + TM_PHI + 16 * ts + 8 * ks, al = tmem_u + TM_PLO + 16 * ts + 8 * ks;
          umma_tf32_ts(tmem_u + TM_ACC0, al, dbh + adv, (kt == 0 && ks == 0) ? 0u : 1u);
          umma_tf32_ts(tmem_u + TM_ACC0, ah, dbl + adv, 1u);
          umma_tf32_ts(tmem_u + TM_ACC0, ah, dbh + adv, 1u);
        }
        umma_commit(tp.done + s);
      }
      __syncwarp();
    } else if (wu == 4) {
      // refill: k-tile kt + RING - 1 goes into the slot k-tile kt - 1 used; its MMAs were issued a whole tile ago
      // (the wait is unconditional: it is also what makes the reuse of TMEM stage / B lo tile (g + 2) % 3 safe, see above)
      if (elect_one()) {
        if (kt >= 1) mbar_wait(tp.done + ((g - 1) % W_RING), ((g - 1) / W_RING) & 1u);
        if (kt + W_RING - 1 < nk) issue_tma(kt + W_RING - 1);
      }
      __syncwarp();
    }
  }
  WTICK(1);
  tp.g = g0 + (uint32_t)nk;
  // drain: the last commit covers every MMA issued before it
  {
    const uint32_t gl = g0 + (uint32_t)nk - 1u;
    mbar_wait(tp.done + (gl % W_RING), (gl / W_RING) & 1u);
  }
  WTICK(6);
  tc_fence_after();
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Control-warp form of the TMA-fed loop (the "W2" instance: 256 worker threads + warp 8 = MMA issuer + warp 9 = TMA producer).
// In gemm_tma the warp that issues the MMAs also has its share of the split work, so every k-tile pays for its issue
// section (~350 cycles: the UTCHMMAs queue behind the tensor pipe) on the critical path, and for a CTA barrier.  Here the
// three roles only meet on mbarriers:
//     TMA producer:  [t >= RING: wait done(t - RING)]  ->  cp.async.bulk.tensor of k-tile t into slot t % RING  -> full(t)
//     workers:       wait full(t) -> split (A: smem -> TMEM hi / lo, B: lo tile) -> one arrival per warp on ready(t)
//     MMA issuer:    wait ready(t) -> 6 x tcgen05.mma -> tcgen05.commit -> done(t)
// full(t) implies that the MMAs of tile t - RING have completed (the producer waited for them), which is what makes slot,
// TMEM stage and B lo tile t % RING free to be overwritten: the workers need no other wait inside the loop, and no barrier.
// The control warps follow the same deterministic schedule of GEMM calls as the workers (chol_tc.cu: w2_control).
//
// Across calls the producer runs AHEAD of the workers: t counts k-tiles over the whole kernel, so while the workers are in
// the epilogue of one call (TRSM, stores -- their scratch is the B-lo region, not the ring) the first RING tiles of the
// next call are already landing.  Two things hold it back, each an mbarrier the workers arrive on:
//     ringfree:  the ring region doubles as LiT | diagonal-block scratch in the first-chunk epilogue of a phase-A block step
//                and in phase B's per-step preamble; one phase per such use, consumed before the next call's first load
//     depready:  the first chunk of a block step reads the panel / inverse block column the step before wrote to global
//                memory; the workers fence (fence.proxy.async.global) and arrive once per block step, and the producer waits
//                before the first k-tile of the newest 64-column slab -- the older slabs (K ascending) are loaded before that
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_workers_tma() { asm volatile("bar.sync 3, 288;" ::: "memory"); }   // 256 workers + the TMA warp

// dep: the call reads global data written since the previous call with dep = true (first chunk of a block step: the
// panel / inverse block column of the step before); the other chunks of a step only read older columns.
// zv / z2v (phase A, first chunk of a block step): the B operand of this call is the block row L[j, 0:R0] that the forward
// substitution z_j = Linv_jj (r_j - L[j, 0:R0] z[0:R0]) needs, and the thread that splits B[row][4 part .. 4 part + 3] of a
// k-tile is exactly the (row, part) owner of that GEMV's partial sum: four FMAs per tile replace a second pass over the
// block row in global memory (same accumulation order as that pass: bit-identical).
template <bool PHASE_B>
__device__ bool gemm_w2_worker(Ctx& c, TmaPipe& tp, int a_row0, int a_row_end, int k_lo, int k_hi, bool dep,
                               const float* zv = nullptr, const float* z2v = nullptr, float* zacc = nullptr) {
  const int nk = (k_hi - k_lo) / 16;
  if (nk <= 0) return false;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int row = 32 * (w & 3) + lane, half_id = w >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
  // (Measured and rejected: warps 0-3 splitting the A tile and warps 4-7 the B tile, so that the two latency chains of a
  // k-tile run side by side -- the workers' own time per tile fell from 340 to 270 cycles but the wait for full(t) grew by
  // the same amount: the tile rate is set by the TMA -> split -> MMA -> done round trip over a 4-slot ring, c2 1.64 -> 1.73 ms.)
#ifdef VOLT_PROFILE
  long long wt0 = clock64(), wt1;
#define WTICK(i) do { if (tid == 0) { wt1 = clock64(); tp.prof[i] += wt1 - wt0; wt0 = wt1; } } while (0)
#else
#define WTICK(i) do { } while (0)
#endif
  if (dep) {   // earlier global writes (panel stores, diagonal blocks) -> visible to the async proxy, then tell the producer
    fence_proxy_async_global();
    wsync();
    if (tid == 0) mbar_arrive(tp.depready);
  }
  WTICK(0);
  const uint32_t g0 = tp.g;
  const bool row_ok = (a_row0 + row) < a_row_end;
  const int mb = (a_row0 + row) >> 6;
  for (int kt = 0; kt < nk; ++kt) {
    const uint32_t g = g0 + (uint32_t)kt, s = g % W_RING;
    const uint8_t* RAW = c.X + s * W_SLOT;
    const uint8_t* BH = RAW + HA_TILE;
    uint8_t* BL = c.X + W_BL + s * HB_TILE;
    mbar_wait(tp.full + s, (g / W_RING) & 1u);   // (probing the next tile's barrier early with test_wait was measured: +2 %)
    WTICK(kt == 0 ? 1 : 2);
    uint32_t hi[8], lo[8];
    const bool live = row_ok && !(PHASE_B && ((k_lo + 16 * kt) >> 6) < mb);
    if constexpr (PHASE_B) {
      // transposed tile (16 rows of k x 128 floats of m, no swizzle): this thread's row is a column of it, lanes read
      // consecutive floats
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float v = *reinterpret_cast<const float*>(RAW + (8 * half_id + kk) * 512 + 4 * row);
        const uint32_t u = live ? __float_as_uint(v) : 0u;
        hi[kk] = u;
        lo[kk] = __float_as_uint(__uint_as_float(u) - __uint_as_float(u & 0xffffe000u));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(RAW + swz64(row, 2 * half_id + q));
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t u = live ? __float_as_uint(e[j]) : 0u;
          hi[4 * q + j] = u;
          lo[4 * q + j] = __float_as_uint(__uint_as_float(u) - __uint_as_float(u & 0xffffe000u));
        }
      }
    }
    tmem_st8(c.tmem + lane_base + TM_PHI + (uint32_t)(16 * s + 8 * half_id), hi);
    tmem_st8(c.tmem + lane_base + TM_PLO + (uint32_t)(16 * s + 8 * half_id), lo);
    const uint32_t off = swz64(tid >> 2, tid & 3);
    const float4 b = *reinterpret_cast<const float4*>(BH + off);
    if (!PHASE_B && zv) {
      const int k0 = k_lo + 16 * kt + 4 * (tid & 3);
      const float4 zq = *reinterpret_cast<const float4*>(zv + k0);
      float a1 = zacc[0];
      a1 = fmaf(b.x, zq.x, a1); a1 = fmaf(b.y, zq.y, a1); a1 = fmaf(b.z, zq.z, a1); a1 = fmaf(b.w, zq.w, a1);
      zacc[0] = a1;
      if (z2v) {
        const float4 z2q = *reinterpret_cast<const float4*>(z2v + k0);
        float a2 = zacc[1];
        a2 = fmaf(b.x, z2q.x, a2); a2 = fmaf(b.y, z2q.y, a2); a2 = fmaf(b.z, z2q.z, a2); a2 = fmaf(b.w, z2q.w, a2);
        zacc[1] = a2;
      }
    }
    float4 l;
    l.x = b.x - __uint_as_float(__float_as_uint(b.x) & 0xffffe000u);
    l.y = b.y - __uint_as_float(__float_as_uint(b.y) & 0xffffe000u);
    l.z = b.z - __uint_as_float(__float_as_uint(b.z) & 0xffffe000u);
    l.w = b.w - __uint_as_float(__float_as_uint(b.w) & 0xffffe000u);
    *reinterpret_cast<float4*>(BL + off) = l;
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tp.ready + s);
    WTICK(3);
  }
  tp.g = g0 + (uint32_t)nk;
  {
    const uint32_t gl = g0 + (uint32_t)nk - 1u;     // the last commit covers every MMA issued before it
    mbar_wait(tp.done + (gl % W_RING), (gl / W_RING) & 1u);
  }
  tc_fence_after();
  WTICK(4);
#ifdef VOLT_PROFILE
  if (tid == 0) { tp.prof[5] += 1; tp.prof[6] += nk; }
#endif
  return true;
}

#ifdef VOLT_PROFILE
static __device__ long long g_w2_prof[8];   // (-DVOLT_PROFILE_CTRL: intrusive, +40 % on the chain) control warps of CTA 0: MMA wait ready | MMA issue | TMA wait ringfree | depready | done | issue
#endif
#ifdef VOLT_PROFILE_CTRL
#define CTICK(i) do { if (blockIdx.x == 0) { const long long _n = clock64(); g_w2_prof[i] += _n - ct0; ct0 = _n; } } while (0)
#define CTICK0() long long ct0 = clock64()
#else
#define CTICK(i) do { } while (0)
#define CTICK0() do { } while (0)
#endif

// the workers' side of ringfree: every worker calls this right after its last access to the ring region as scratch
__device__ __forceinline__ void w2_release_ring(TmaPipe& tp) {
  fence_async_smem();
  wsync();
  if (threadIdx.x == 0) mbar_arrive(tp.ringfree);
}

// (Measured and rejected: the producer pulling the whole NEXT call's tiles into the L2 with cp.async.bulk.prefetch.tensor --
// c2 1.64 -> 1.73 ms, c3 2.29 -> 2.83 ms: with 296 MB of scratch squares behind a 126 MB L2 the prefetched lines evict the
// lines the other CTAs are about to re-read.  A short look-ahead, 6 or 10 tiles inside the same call, changes nothing: 1.63 ms.)
// TMA producer warp, one elected lane issues the loads.  wait_ring: the workers used the ring region as scratch since the
// previous call; dep: the k-tiles from n_indep on read what the previous block step wrote.
// a_t (phase B): the A operand is stored transposed (the inverse lives IN PLACE of L, see chol_tc.cu): tmA is then the map
// with boxes of 128 floats x 16 rows, row_a the first row of the CTA's square and col_a the tile's first column.
__device__ __forceinline__ void w2_tma_call(TmaPipe& tp, const void* tmA, const void* tmB, uint32_t xb, int row_a, int row_b, int k_lo, int nk,
                                            bool wait_ring, bool dep, int n_indep, bool a_t = false, int col_a = 0) {
  const uint32_t g0 = tp.g;
  if (elect_one()) {
    CTICK0();
    if (wait_ring) mbar_wait(tp.ringfree, tp.rf_n & 1u);
    CTICK(2);
    for (int kt = 0; kt < nk; ++kt) {
      const uint32_t g = g0 + (uint32_t)kt, s = g % W_RING;
      if (dep && kt == n_indep) { mbar_wait(tp.depready, tp.dep_n & 1u); CTICK(3); }
      if (g >= W_RING) mbar_wait(tp.done + s, ((g - W_RING) / W_RING) & 1u);    // MMAs of the tile that used the slot before
      CTICK(4);
      mbar_expect_tx(tp.full + s, W_SLOT);
      if (a_t) tma_load_2d(tmA, xb + s * W_SLOT, tp.full + s, col_a, row_a + k_lo + 16 * kt);
      else tma_load_2d(tmA, xb + s * W_SLOT, tp.full + s, k_lo + 16 * kt, row_a);
      tma_load_2d(tmB, xb + s * W_SLOT + HA_TILE, tp.full + s, k_lo + 16 * kt, row_b);
      CTICK(5);
    }
  }
  __syncwarp();
  tp.g = g0 + (uint32_t)nk;
  if (wait_ring) ++tp.rf_n;
  if (dep) ++tp.dep_n;
}

// MMA issuer warp
__device__ __forceinline__ void w2_mma_call(TmaPipe& tp, uint32_t tmem_u, uint32_t xb, int nk) {
  const uint32_t g0 = tp.g;
  if (elect_one()) {
    CTICK0();
    for (int kt = 0; kt < nk; ++kt) {
      const uint32_t g = g0 + (uint32_t)kt, s = g % W_RING;
      mbar_wait(tp.ready + s, (g / W_RING) & 1u);
      CTICK(0);
      tc_fence_after();
      const uint64_t dbh = make_desc64(xb + s * W_SLOT + HA_TILE), dbl = make_desc64(xb + W_BL + s * HB_TILE);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t adv = (uint64_t)(2 * ks);
        const uint32_t ah = tmem_u + TM_PHI + 16 * s + 8 * ks, al = tmem_u + TM_PLO + 16 * s + 8 * ks;
        umma_tf32_ts(tmem_u + TM_ACC0, al, dbh + adv, (kt == 0 && ks == 0) ? 0u : 1u);
        umma_tf32_ts(tmem_u + TM_ACC0, ah, dbl + adv, 1u);
        umma_tf32_ts(tmem_u + TM_ACC0, ah, dbh + adv, 1u);
      }
      umma_commit(tp.done + s);
      CTICK(1);
    }
  }
  __syncwarp();
  tp.g = g0 + (uint32_t)nk;
}

// A-generator for one accumulator row: s[q] <- A[gr][gc0 + q] - s[q], q = 0..31.  On-the-fly kernels: the 32 column values
// come from 8 broadcast LDS.128 of the staged prefix vector (zero beyond T); the identity padding of rows / columns >= T is
// applied per entry only by the threads whose 32 entries touch it (T = 400 in Tp = 448: the last block row and column).
__device__ __forceinline__ void gen_sub_row32(const MllParams& p, int b, int gr, int gc0, const float* Vs, float sc, float dadd,
                                              float (&s)[32]) {
  if (p.kind != KIND_DENSE) {
    const float vr = Vs[gr];
    const bool vol = (p.kind == KIND_VOL);
    const bool interior = (gr < p.T) && (gc0 + 32 <= p.T);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 cv = *reinterpret_cast<const float4*>(Vs + gc0 + 4 * q4);
      const float c4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gc = gc0 + 4 * q4 + e;
        float v = vol ? ((gc <= gr) ? c4[e] : vr) : sc * fminf(vr, c4[e]);
        if (gc == gr) v += dadd;
        if (!interior && (gr >= p.T || gc >= p.T)) v = (gc == gr) ? 1.f : 0.f;
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

// store_block32 for the control-warp instance: the transposition tile lives in the B-lo region of the ring map (free between
// GEMM calls; the ring slots themselves may already be receiving the next call's tiles).  Warp w owns bytes [512 w, 512 w + 512)
// of each of the W_BLN lo tiles -- exactly the bytes its own threads write when they split a B tile, so no other warp ever
// touches them -- i.e. four pieces of 8 rows x 16 floats: the 32 x 32 block goes out in two passes of 16 columns.  The 16-byte
// chunk of a row is XOR-swizzled with (row / 2) % 4: each 8-lane phase of the STS.128 and of the LDS.128 hits 8 bank groups.
__device__ __forceinline__ void store_block32_bl(uint8_t* bl, int warp, const float (&o)[32], float* gdst, int ld, int lane) {
  float* mine = reinterpret_cast<float*>(bl + 512 * warp);
  const int wp = lane >> 3, wr = lane & 7, wsw = (wr >> 1) & 3;        // write: lane = row -> piece, row in piece
  const int rr = lane >> 2, rc = lane & 3, rsw = (rr >> 1) & 3;        // read: 8 rows x 4 chunks per pass
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(mine + wp * (HB_TILE / 4) + wr * 16 + 4 * (q ^ wsw)) =
          make_float4(o[16 * h + 4 * q], o[16 * h + 4 * q + 1], o[16 * h + 4 * q + 2], o[16 * h + 4 * q + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(gdst + (size_t)(8 * i + rr) * ld + 16 * h + 4 * rc) =
          *reinterpret_cast<const float4*>(mine + i * (HB_TILE / 4) + rr * 16 + 4 * (rc ^ rsw));
    __syncwarp();
  }
}

}  // namespace tc
}  // namespace volt
