// FP32 tile GEMM with functor operands -- the dense-contraction core shared by the actor head, the PPO backward
// and the tracker's training pass.  C[M,N] (op)= sum_k A(m,k) * B(k,n).
//
// FP32 FFMA accumulation (not TF32/BF16 tensor-core MMA): the north-star parity bar is 1e-5 relative on
// probabilities / losses, which single-pass TF32 (10-bit mantissa) cannot meet.
//
//   LA / LB : operand functors.  `float operator()(int m, int k) const` returns the element (bounds already
//             checked by the tile loader);  `static constexpr bool INNER_IS_K` says which index is contiguous
//             in memory, which selects the thread->element mapping of the tile loader so that global loads
//             coalesce either way.
//   EP      : epilogue functor  `void operator()(int m, int n, float acc) const`.
// Tiles: BM x BN per CTA, BK per k-step, each thread TM x 4 outputs (TN fixed to 4 so every thread reads its
// B fragment as one float4), (BM/TM) x (BN/4) threads.  blockIdx.z splits K (epilogue must then accumulate).
#pragma once
#include "common.cuh"

namespace cirs {

// ---- operand functors ---------------------------------------------------------------------------------
// element (r, c) of a row-major matrix with optional row gather; used as A(m=r,k=c) [INNER_IS_K] ...
struct RowMajorA {  // A(m,k) = p[row(m)*ld + k]
  static constexpr bool INNER_IS_K = true;
  const float* p; int64_t ld; const int32_t* row;
  __device__ __forceinline__ float operator()(int m, int k) const {
    return __ldg(p + (int64_t)(row ? row[m] : m) * ld + k);
  }
};
struct ColMajorA {  // A(m,k) = p[row(k)*ld + m]   (the transpose of a row-major [K, M] matrix)
  static constexpr bool INNER_IS_K = false;
  const float* p; int64_t ld; const int32_t* row;
  __device__ __forceinline__ float operator()(int m, int k) const {
    return __ldg(p + (int64_t)(row ? row[k] : k) * ld + m);
  }
};
struct RowMajorB {  // B(k,n) = p[row(k)*ld + n]
  static constexpr bool INNER_IS_K = false;
  const float* p; int64_t ld; const int32_t* row;
  __device__ __forceinline__ float operator()(int k, int n) const {
    return __ldg(p + (int64_t)(row ? row[k] : k) * ld + n);
  }
};
struct ColMajorB {  // B(k,n) = p[n*ld + k]   (a row-major [N, K] matrix used transposed)
  static constexpr bool INNER_IS_K = true;
  const float* p; int64_t ld;
  __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(p + (int64_t)n * ld + k); }
};

// ---- epilogues ----------------------------------------------------------------------------------------
struct StoreEp {  // C[row(m)*ld + n] = act(acc + bias[n]) * (mask ? mask[m*ldm+n] > 0 : 1) + (res ? res[m*ldr+n] : 0)
  float* c; int64_t ld; const float* bias; int relu; const int32_t* row; const float* mask; int64_t ldm;
  const float* res; int64_t ldr;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    if (bias) v += __ldg(bias + n);
    if (relu) v = fmaxf(v, 0.f);
    if (mask && !(mask[(int64_t)m * ldm + n] > 0.f)) v = 0.f;
    if (res) v += res[(int64_t)m * ldr + n];
    c[(int64_t)(row ? row[m] : m) * ld + n] = v;
  }
};
struct AtomicEp {  // C[m*ld + n] += acc   (split-K partial sums)
  float* c; int64_t ld;
  __device__ __forceinline__ void operator()(int m, int n, float v) const { atomicAdd(c + (int64_t)m * ld + n, v); }
};

// ---- kernel -------------------------------------------------------------------------------------------
// the CTA-level body: tile (bx, by) of C, k-range of split bz.  gemm_kernel maps blockIdx onto it; grouped launches
// (tracker_fused.cuh) map a flat CTA index onto (problem, tile, split) themselves.
template <int BM, int BN, int BK, int TM, class LA, class LB, class EP>
__device__ __forceinline__ void gemm_tile(LA la, LB lb, EP ep, int M, int N, int K, int k_per_split,
                                          float* __restrict__ colsum, int bx, int by, int bz) {
  constexpr int TN = 4;
  constexpr int NTX = BN / TN, NTY = BM / TM, NT = NTX * NTY;
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;  // +4 keeps float4 alignment and staggers banks
  __shared__ __align__(16) float As[BK][LDA_S];
  __shared__ __align__(16) float Bs[BK][LDB_S];
  static_assert(TM % 4 == 0, "TM must be a multiple of 4");
  const int tid = threadIdx.x, tx = tid % NTX, ty = tid / NTX;
  const int m0 = by * BM, n0 = bx * BN;
  const int kb = bz * k_per_split, ke = min(K, kb + k_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float csum = 0.f;  // column sum of B over this CTA's k-range (bias gradient), column n0 + tid % BN

  // Register double buffering: the operand elements of k-step s+1 are loaded from global memory into registers
  // before the FMAs of k-step s, so the load latency hides behind the math instead of adding to every step.
  constexpr int NA = BM * BK / NT, NB = BN * BK / NT;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile must be a multiple of the CTA size");
  float ra[NA], rb[NB];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      const int i = tid + j * NT;
      const int kk = LA::INNER_IS_K ? i % BK : i / BM, mm = LA::INNER_IS_K ? i / BK : i % BM;
      const int m = m0 + mm, k = k0 + kk;
      ra[j] = (m < M && k < ke) ? la(m, k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int i = tid + j * NT;
      const int kk = LB::INNER_IS_K ? i % BK : i / BN, nn = LB::INNER_IS_K ? i / BK : i % BN;
      const int n = n0 + nn, k = k0 + kk;
      rb[j] = (n < N && k < ke) ? lb(k, n) : 0.f;
    }
  };
  if (kb < ke) load_tiles(kb);
  for (int k0 = kb; k0 < ke; k0 += BK) {
    // ---- registers -> shared: As[k][m], Bs[k][n]
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      const int i = tid + j * NT;
      const int kk = LA::INNER_IS_K ? i % BK : i / BM, mm = LA::INNER_IS_K ? i / BK : i % BM;
      As[kk][mm] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int i = tid + j * NT;
      const int kk = LB::INNER_IS_K ? i % BK : i / BN, nn = LB::INNER_IS_K ? i / BK : i % BN;
      Bs[kk][nn] = rb[j];
      if (!LB::INNER_IS_K && NT % BN == 0) csum += rb[j];  // nn == tid % BN for every j when NT % BN == 0
    }
    __syncthreads();
    if (k0 + BK < ke) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < N) ep(m, n, acc[i][j]);
    }
  }
  if (colsum != nullptr && !LB::INNER_IS_K && NT % BN == 0 && by == 0) {
    const int n = n0 + tid % BN;
    if (n < N) atomicAdd(colsum + n, csum);
  }
}

template <int BM, int BN, int BK, int TM, class LA, class LB, class EP>
__global__ void __launch_bounds__((BM / TM) * (BN / 4))
gemm_kernel(LA la, LB lb, EP ep, int M, int N, int K, int k_per_split, float* __restrict__ colsum) {
  gemm_tile<BM, BN, BK, TM>(la, lb, ep, M, N, K, k_per_split, colsum, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Host launcher.  split_k: number of K partitions (epilogue must accumulate when > 1).
template <int BM, int BN, int BK, int TM, class LA, class LB, class EP>
inline void launch_gemm(LA la, LB lb, EP ep, int M, int N, int K, int split_k, float* colsum, cudaStream_t st,
                        const char* tag = "gemm") {
  if (M <= 0 || N <= 0 || K <= 0) return;
  if (split_k < 1) split_k = 1;
  int kps = (K + split_k - 1) / split_k;
  kps = ((kps + BK - 1) / BK) * BK;
  split_k = (K + kps - 1) / kps;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split_k);
  const bool prof = cirs_profile_begin(tag, st);
  gemm_kernel<BM, BN, BK, TM, LA, LB, EP><<<grid, (BM / TM) * (BN / 4), 0, st>>>(la, lb, ep, M, N, K, kps, colsum);
  cirs_note_launch();
  if (prof) cirs_profile_end(st);
}

}  // namespace cirs
