// Microbenchmark: latency / throughput of cp.async.bulk (1-D TMA, SASS UBLKCP) global(L2) -> shared on sm_100a.
// One thread per CTA issues copies of BYTES bytes as PARTS bulk operations with DEPTH copies in flight and waits on the
// mbarrier; prints cycles per copy.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bench bulk_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)),
               "l"(s), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
}

__global__ void k(const char* src, size_t src_bytes, int bytes, int parts, int depth, int iters, long long* out) {
  extern __shared__ __align__(1024) char sm[];
  __shared__ __align__(8) uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t n_tiles = src_bytes / bytes;
  size_t tile = (size_t)blockIdx.x * 7919u % n_tiles;
  auto issue = [&](int i) {
    const int s = i % depth;
    expect_tx(&bar[s], bytes);
    const int pb = bytes / parts;
    for (int p = 0; p < parts; ++p) bulk(sm + (size_t)s * bytes + p * pb, src + tile * bytes + p * pb, pb, &bar[s]);
    tile = (tile + gridDim.x) % n_tiles;
  };
  for (int i = 0; i < depth; ++i) issue(i);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    wait(&bar[i % depth], (i / depth) & 1);
    if (i + depth < iters + depth) issue(i + depth);   // keep DEPTH in flight (drained after the loop)
  }
  const long long t1 = clock64();
  for (int i = iters; i < iters + depth; ++i) wait(&bar[i % depth], (i / depth) & 1);
  out[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t SRC = 11u << 20;   // two W3 images' worth: L2 resident after the first pass
  char* src;
  long long* out;
  cudaMalloc(&src, SRC);
  cudaMemset(src, 1, SRC);
  cudaMalloc(&out, 1024 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[1024];
  const int iters = 64;
  printf("%8s %6s %6s %6s %12s %12s\n", "bytes", "parts", "depth", "grid", "cyc/copy", "B/clk/SM");
  for (int grid : {1, 148, 296}) {
    for (int bytes : {32768, 65536}) {
      for (int parts : {1, 2, 4, 8, 32}) {
        for (int depth : {1, 2, 3}) {
          if ((size_t)bytes * depth > 96 * 1024 && grid == 296) continue;
          if ((size_t)bytes * depth > 196 * 1024) continue;
          for (int rep = 0; rep < 2; ++rep) {
            k<<<grid, 32, (size_t)bytes * depth>>>(src, SRC, bytes, parts, depth, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < grid; ++i) s += (double)h[i];
          const double cyc = s / grid / iters;
          const int per_sm = grid > 148 ? 2 : 1;
          printf("%8d %6d %6d %6d %12.0f %12.1f\n", bytes, parts, depth, grid, cyc, per_sm * bytes / cyc);
        }
      }
    }
  }
  return 0;
}
