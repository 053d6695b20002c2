// Microbenchmark: one matrix-vector stage of the rollout's token step, R rows x [n_in -> n_out], weights k-major in shared
// memory.  Variants: (a) thread per (row, output), scalar loads (tracker_cta_dev.cuh mv_rows); (b) the same with the
// weight pointer read from a struct in shared memory (generic loads); (c) warp per row, lane = output quad, float4 weight
// loads + k-groups (tracker_dev.cuh matvec); (d) thread per (row, output quad), float4 weights, float4 x.
#include <cstdio>
#include <cuda_runtime.h>
struct Ptrs { const float* w; const float* b; };
template <int V>
__global__ void __launch_bounds__(256, 1) k(const float* gw, const float* gb, float* out, long long* cyc, int R, int n_in, int n_out, int ldo, int reps) {
  extern __shared__ __align__(16) float sm[];
  __shared__ Ptrs P;
  float* w = sm;                 // [n_in][ldo]
  float* b = w + n_in * ldo;     // [ldo]
  float* x = b + ldo;            // [R][128]
  float* y = x + 8 * 128;        // [R][128]
  for (int i = threadIdx.x; i < n_in * ldo; i += 256) w[i] = gw[i];
  for (int i = threadIdx.x; i < ldo; i += 256) b[i] = gb[i];
  for (int i = threadIdx.x; i < 8 * 128; i += 256) x[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { P.w = w; P.b = b; }
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (V == 0 || V == 1) {
      const float* W = V == 1 ? P.w : w;
      const float* Bv = V == 1 ? P.b : b;
      for (int idx = tid; idx < R * n_out; idx += 256) {
        const int r = idx / n_out, o = idx - r * n_out;
        const float* xr = x + r * 128;
        float a = Bv[o];
#pragma unroll 8
        for (int i = 0; i < n_in; ++i) a = fmaf(W[i * ldo + o], xr[i], a);
        y[r * 128 + o] = a;
      }
    } else if (V == 2) {
      if (warp < R) {   // warp per row: quads x groups
        const float* xr = x + warp * 128;
        for (int o0 = 0; o0 < n_out; o0 += 128) {
          const int nq = min(32, (n_out - o0 + 3) >> 2);
          const int Pq = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 16 ? 16 : 32;
          const int G = 32 / Pq, q = lane & (Pq - 1), g = lane / Pq;
          float4 acc = make_float4(0, 0, 0, 0);
          if (q < nq) {
            const float* wp = w + o0 + 4 * q + g * ldo;
            const int step = G * ldo;
#pragma unroll 8
            for (int i = g; i < n_in; i += G) {
              const float xi = xr[i];
              const float4 w4 = *reinterpret_cast<const float4*>(wp);
              acc.x = fmaf(w4.x, xi, acc.x); acc.y = fmaf(w4.y, xi, acc.y); acc.z = fmaf(w4.z, xi, acc.z); acc.w = fmaf(w4.w, xi, acc.w);
              wp += step;
            }
          }
          for (int off = Pq; off < 32; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
          }
          if (g == 0 && q < nq) { float* yo = y + warp * 128 + o0 + 4 * q; yo[0] = acc.x + b[o0 + 4 * q]; yo[1] = acc.y; yo[2] = acc.z; yo[3] = acc.w; }
        }
      }
    } else if (V == 3) {   // thread per (row, quad): float4 w, float4 x
      const int nq = (n_out + 3) >> 2;
      for (int idx = tid; idx < R * nq; idx += 256) {
        const int r = idx / nq, q = idx - r * nq;
        const float* xr = x + r * 128;
        const float* wp = w + 4 * q;
        float4 acc = *reinterpret_cast<const float4*>(b + 4 * q);
#pragma unroll 2
        for (int i = 0; i < n_in; i += 4) {
          const float4 xv = *reinterpret_cast<const float4*>(xr + i);
          const float4 w0 = *reinterpret_cast<const float4*>(wp + (i + 0) * ldo), w1 = *reinterpret_cast<const float4*>(wp + (i + 1) * ldo);
          const float4 w2 = *reinterpret_cast<const float4*>(wp + (i + 2) * ldo), w3 = *reinterpret_cast<const float4*>(wp + (i + 3) * ldo);
          acc.x = fmaf(w0.x, xv.x, acc.x); acc.y = fmaf(w0.y, xv.x, acc.y); acc.z = fmaf(w0.z, xv.x, acc.z); acc.w = fmaf(w0.w, xv.x, acc.w);
          acc.x = fmaf(w1.x, xv.y, acc.x); acc.y = fmaf(w1.y, xv.y, acc.y); acc.z = fmaf(w1.z, xv.y, acc.z); acc.w = fmaf(w1.w, xv.y, acc.w);
          acc.x = fmaf(w2.x, xv.z, acc.x); acc.y = fmaf(w2.y, xv.z, acc.y); acc.z = fmaf(w2.z, xv.z, acc.z); acc.w = fmaf(w2.w, xv.z, acc.w);
          acc.x = fmaf(w3.x, xv.w, acc.x); acc.y = fmaf(w3.y, xv.w, acc.y); acc.z = fmaf(w3.z, xv.w, acc.z); acc.w = fmaf(w3.w, xv.w, acc.w);
        }
        *reinterpret_cast<float4*>(y + r * 128 + 4 * q) = acc;
      }
    } else if (V == 4) {   // thread per (row, output, k-half): 2-way split over k, combined by shuffle with the neighbour lane
      for (int idx = tid; idx < 2 * R * n_out; idx += 256) {
        const int pair = idx >> 1, half = idx & 1;
        const int r = pair / n_out, o = pair - r * n_out;
        const float* xr = x + r * 128;
        float a = half ? 0.f : b[o];
        const int kb = half * (n_in / 2), ke = half ? n_in : n_in / 2;
#pragma unroll 8
        for (int i = kb; i < ke; ++i) a = fmaf(w[i * ldo + o], xr[i], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        if (!half) y[r * 128 + o] = a;
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = (t1 - t0) / reps;
  if (tid < 128) out[tid] = y[tid];
}
template <int V> void run(const char* name, int R, int n_in, int n_out) {
  const int ldo = (n_out + 31) & ~31;
  float *gw, *gb, *out; long long* cyc;
  cudaMalloc(&gw, n_in * ldo * 4); cudaMalloc(&gb, ldo * 4); cudaMalloc(&out, 512); cudaMalloc(&cyc, 8);
  cudaMemset(gw, 0, n_in * ldo * 4); cudaMemset(gb, 0, ldo * 4);
  const size_t smem = (size_t)(n_in * ldo + ldo + 16 * 128) * 4;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<V><<<1, 256, smem>>>(gw, gb, out, cyc, R, n_in, n_out, ldo, 200);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s R=%d %3d->%3d : %6lld cycles/stage (%s)\n", name, R, n_in, n_out, h, cudaGetErrorString(cudaGetLastError()));
  cudaFree(gw); cudaFree(gb); cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int R : {1, 4, 8}) {
    for (auto sh : {std::pair<int,int>{32, 96}, {32, 32}, {32, 128}, {128, 32}}) {
      run<0>("thread/(row,out) scalar", R, sh.first, sh.second);
      run<1>("  same, ptr from smem struct", R, sh.first, sh.second);
      run<2>("warp/row quads x groups", R, sh.first, sh.second);
      run<3>("thread/(row,quad) f4 w, f4 x", R, sh.first, sh.second);
      run<4>("thread/(row,out,k-half)", R, sh.first, sh.second);
    }
  }
  return 0;
}
