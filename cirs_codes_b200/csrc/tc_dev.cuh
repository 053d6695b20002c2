// tcgen05 (5th-generation tensor core) building blocks for sm_100a: TMEM allocation, UMMA shared-memory / instruction
// descriptors, single-thread MMA issue, commit -> mbarrier, TMEM -> register loads, and the operand tile layout
// used by the actor-head kernels (head_tc.cu).
//
// Precision scheme ("3xTF32"): the north-star parity bar is 1e-5 relative on probabilities / losses, which one TF32
// pass (10-bit mantissa) cannot meet.  Every FP32 operand x is split exactly as x = hi + lo, hi = x with the low 13
// mantissa bits zero (cvt.rna.tf32: exactly representable in TF32 whatever the hardware's own input rounding),
// lo = rna_tf32(x - hi) (x - hi is exact in FP32, |lo| <= 2^-12 |x|, its rounding costs 2^-24 of x).
// a.b ~ hi.hi + hi.lo + lo.hi, three kind::tf32 MMAs accumulated in FP32 in TMEM; the dropped lo.lo term is 2^-24
// relative.  Measured (tests/tc_probe.cu, K = 64, |a|,|b| <= 1): max abs error ~1e-6 against FP64.
//
// Operand tile layout (no swizzle, "interleaved" core matrices).  A tile holds R x C floats, C the contiguous
// dimension of the source; element (r, c) lives at byte
//        (c / 4) * (R * 16)  +  (r / 8) * 128  +  (r % 8) * 16  +  (c % 4) * 4
// i.e. 16-byte chunks of 4 consecutive c, the 8 r of a core matrix contiguous (128 B), core matrices of one chunk
// column contiguous.  The tile is read as a K-major operand (K = c): LBO (K-chunk stride) = R*16, SBO (8-row group
// stride) = 128, k-step j (K = 8j..8j+7) starts at + 2*j*R*16.  Operands whose source is contiguous along the OTHER
// dimension are transposed while staging (tile_stage_T); the MN-major reading of the same bytes (SBO = R*16, LBO = 128)
// returned zeros for kind::tf32 on B200 in tests/tc_probe.cu and is not used.
// (canonical layouts: CUTLASS cute/atom/mma_traits_sm100.hpp make_umma_desc, LayoutType::INTERLEAVE.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cirs_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors ---------------------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (sm_100): start address, leading / stride byte offsets (all >> 4),
// version = 1 at bit 46, layout type 0 (no swizzle) at bits 61-63.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// 32-bit instruction descriptor, kind::tf32, FP32 accumulate: c_format = 1 (bits 4-5), a/b_format = 2 (TF32, bits 7-9 /
// 10-12), a/b major (bit 15 / 16: 0 = K-major, 1 = MN-major), N >> 3 (bits 17-22), M >> 4 (bits 24-28).
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM -----------------------------------------------------------------------------------------------------
// one full warp allocates ncols (power of two >= 32) columns; the base address is written to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the MMA's operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA issue (ONE thread) -----------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in TMEM (lane = row of A, one TF32 element per 32-bit column, K along the columns): D (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// transaction-count arrive (the thread that issues the bulk copies) and plain arrive
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy (1-D): global -> shared, completion counted in bytes on the mbarrier.  dst / src 16-byte aligned,
// bytes a multiple of 16.  The data lands through the async proxy: an MMA may read it after waiting on the barrier,
// without a proxy fence.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Bounded spin: returns false (instead of hanging the GPU) if the phase does not complete.
// CIRS_MBAR_MODE (compile time, experiments): 0 = try_wait (may suspend the thread for a system-dependent time), 1 =
// test_wait (pure polling), 2 = try_wait with a 32 ns suspend-time hint.
#ifndef CIRS_MBAR_MODE
#define CIRS_MBAR_MODE 0
#endif
__device__ __forceinline__ bool mbar_wait_n(uint64_t* bar, uint32_t parity, uint32_t max_spins) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0; spin < max_spins; ++spin) {
    uint32_t ok;
#if CIRS_MBAR_MODE == 1
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
#elif CIRS_MBAR_MODE == 2
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity), "r"(32u)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
#endif
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) { return mbar_wait_n(bar, parity, 1u << 28); }

// ---- TMEM -> registers: warp w reads lanes 32*(w%4).., thread = lane, 32 consecutive FP32 columns ---------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// split form: issue the load now, wait (for ALL outstanding tcgen05.ld of this thread) before the first use.  The
// wait takes the destination registers as in/out operands so that the compiler cannot move a use above it.
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8_wait(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// registers -> TMEM: thread = lane, 16 consecutive 32-bit columns; tmem_st_wait() before the data is consumed by an MMA
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMEM address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

// ---- operand tiles --------------------------------------------------------------------------------------------
// round to nearest TF32 (the result has the low 13 mantissa bits clear, so the MMA reads it exactly)
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// exp(x) with ~2 ulp error in 6 instructions: ex2.approx of the rounded product x * log2(e), corrected to first order
// for the product's rounding error and the low part of log2(e).  x <= 0 in all uses; exp(-1e30) = 0 (no NaN).
__device__ __forceinline__ float fast_exp(float x) {
  const float t = x * 1.4426950408889634f;
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  const float e = fmaf(x, 1.4426950408889634f, -t) + x * 1.925963033500011e-8f;
  return fmaf(r, e * 0.6931471805599453f, r);
}
// cheaper split for operands that only need ~2^-22 relative accuracy (gradients): hi by truncation, lo = x - hi exact
// (the MMA truncates lo to TF32)
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// byte offset of the 16-byte chunk (r, c4 = c / 4) inside an R-row tile
__device__ __forceinline__ uint32_t tile_chunk_off(int R, int r, int c4) {
  return (uint32_t)c4 * (uint32_t)(R * 16) + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
}
// store one chunk (4 consecutive c of row r) as hi / lo
__device__ __forceinline__ void tile_store_split(char* hi, char* lo, int R, int r, int c4, float4 v) {
  const uint32_t off = tile_chunk_off(R, r, c4);
  float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) =
      make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
}
__device__ __forceinline__ void tile_store_split_trunc(char* hi, char* lo, int R, int r, int c4, float4 v) {
  const uint32_t off = tile_chunk_off(R, r, c4);
  float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}
// Cooperative staging of an R x C tile (R % 8 == 0, C % 16 == 0) from a row-major source, split into a LOAD phase
// (global -> registers, all loads of a thread issued back to back so their latencies overlap, and early enough to
// prefetch the next tile behind the current tile's MMA + epilogue) and a STORE phase (registers -> hi / lo tiles).
// src(r, c4) returns the float4 of row r, columns 4*c4 .. 4*c4+3 (zero outside the matrix).  Lane mapping: 8 rows x 4
// chunks per warp step, so global reads are 64-byte row segments and shared stores conflict-free 16-byte vectors.
template <int R, int C, int NTHREADS>
struct TileV {
  static constexpr int RGS = R / 8, CQS = C / 16, NW = NTHREADS / 32;
  static constexpr int STEPS = (RGS * CQS + NW - 1) / NW;
  float4 buf[STEPS];
  template <class Src>
  __device__ __forceinline__ void load(int tid, Src src) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      const int w = warp + i * NW;
      if (RGS * CQS % NW == 0 || w < RGS * CQS) {
        const int rg = w % RGS, cq = w / RGS;
        buf[i] = src(rg * 8 + (lane & 7), cq * 4 + (lane >> 3));
      }
    }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int tid) const {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      const int w = warp + i * NW;
      if (RGS * CQS % NW == 0 || w < RGS * CQS) {
        const int rg = w % RGS, cq = w / RGS;
        tile_store_split(hi, lo, R, rg * 8 + (lane & 7), cq * 4 + (lane >> 3), buf[i]);
      }
    }
  }
};
// Transposing variant: the source is contiguous along r (the tile's row index), src(r, c) returns one element.
// Lane mapping: 4 consecutive c x 8 consecutive r per warp step -> global reads are 32-byte segments along r, the
// 4-byte shared stores of a warp cover 128 contiguous bytes (conflict free).
template <int R, int C, int NTHREADS>
struct TileT {
  static constexpr int RGS = R / 8, C4S = C / 4, NW = NTHREADS / 32;
  static constexpr int STEPS = (RGS * C4S + NW - 1) / NW;
  float buf[STEPS];
  template <class Src>
  __device__ __forceinline__ void load(int tid, Src src) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      const int w = warp + i * NW;
      if (RGS * C4S % NW == 0 || w < RGS * C4S) {
        const int rg = w % RGS, c4 = w / RGS;
        buf[i] = src(rg * 8 + (lane >> 2), c4 * 4 + (lane & 3));
      }
    }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int tid) const {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      const int w = warp + i * NW;
      if (RGS * C4S % NW == 0 || w < RGS * C4S) {
        const int rg = w % RGS, c4 = w / RGS;
        const float x = buf[i], h = tf32_hi(x);
        const uint32_t off = tile_chunk_off(R, rg * 8 + (lane >> 2), c4) + (uint32_t)(lane & 3) * 4u;
        *reinterpret_cast<float*>(hi + off) = h;
        *reinterpret_cast<float*>(lo + off) = tf32_hi(x - h);
      }
    }
  }
};
// run-time-shaped one-shot versions (tests/tc_probe.cu)
template <class Src>
__device__ __forceinline__ void tile_stage(char* hi, char* lo, int R, int C, int tid, int nthreads, Src src) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const int rgs = R >> 3, cqs = C >> 4;  // row groups of 8, chunk quads of 4 chunks
  for (int w = warp; w < rgs * cqs; w += nwarps) {
    const int rg = w % rgs, cq = w / rgs;
    const int r = rg * 8 + (lane & 7), c4 = cq * 4 + (lane >> 3);
    tile_store_split(hi, lo, R, r, c4, src(r, c4));
  }
}

// Issue the three TF32 products of one k-step range: D (+)= A.B with A, B given as hi / lo tiles.
//   a_hi/a_lo, b_hi/b_lo : shared addresses (u32) of the tiles' first k-step
//   a_step, b_step       : byte advance per k-step (8 K elements);  lbo / sbo per operand as in the header comment
// One descriptor per operand tile; k-step j is the same descriptor with the start-address field advanced by j * step / 16
// (shared addresses are < 256 KB, so the 14-bit field never carries): one add per MMA operand instead of rebuilding the
// descriptor -- the issuing lane's instruction stream paces the batch (measured in user_model.cu: 47 -> 37 cycles / MMA).
__device__ __forceinline__ void mma_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_step,
                                           uint32_t a_lbo, uint32_t a_sbo, uint32_t b_hi, uint32_t b_lo,
                                           uint32_t b_step, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc,
                                           int ksteps, bool accumulate_first) {
  uint32_t acc = accumulate_first ? 1u : 0u;
  const uint64_t ah0 = smem_desc(a_hi, a_lbo, a_sbo), al0 = smem_desc(a_lo, a_lbo, a_sbo);
  const uint64_t bh0 = smem_desc(b_hi, b_lbo, b_sbo), bl0 = smem_desc(b_lo, b_lbo, b_sbo);
  const uint64_t as = a_step >> 4, bs = b_step >> 4;
#pragma unroll
  for (int j = 0; j < ksteps; ++j) {
    mma_tf32(d_tmem, al0 + j * as, bh0 + j * bs, idesc, acc);   // small terms first
    mma_tf32(d_tmem, ah0 + j * as, bl0 + j * bs, idesc, 1u);
    mma_tf32(d_tmem, ah0 + j * as, bh0 + j * bs, idesc, 1u);
    acc = 1u;
  }
}

// One lane of a converged warp (elect.sync).  Code under it is provably executed by a single thread, which lets the
// compiler keep the MMA descriptors / TMEM addresses in uniform registers: under `if (tid == X)` every tcgen05.mma costs
// an R2UR + ELECT + BRA.U.ANY divergence waterfall (~8 SASS instructions, 72 cycles per MMA measured in user_model.cu
// against the tensor pipe's 31).  The same lane is elected every time for the same mask, so single-thread program order
// holds across consecutive elected regions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp-collective forms of the single-thread operations of a producer / issuer warp (all 32 lanes call them)
__device__ __forceinline__ void w_expect_tx(uint64_t* bar, uint32_t bytes) {
  if (elect_one()) mbar_expect_tx(bar, bytes);
}
__device__ __forceinline__ void w_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  if (elect_one()) bulk_g2s(smem_dst, gmem_src, bytes, bar);
}
__device__ __forceinline__ void w_commit(uint64_t* bar) {
  if (elect_one()) mma_commit(bar);
}

}  // namespace cirs_tc
