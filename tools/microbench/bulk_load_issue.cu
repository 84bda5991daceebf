// How fast can bulk loads be issued / completed?  P producer threads (one per warp) each stream rows
// of row_bytes from a random place of a big buffer into their own ring of `depth` slots.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(const float* g, size_t n_rows_total, long long* out, int iters, int row_bytes, int depth, int P, int spin) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ alignas(8) unsigned long long bar[8][8];
  const int tid = threadIdx.x; int w = tid >> 5, lane = tid & 31;
  if (spin) { w = (tid < P) ? tid : 99; lane = 0; }   // spin==1: the producers are lanes of warp 0
  if (tid == 0) for (int p = 0; p < 8; ++p) for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[p][i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  long long t0 = clock64(), tissue = 0;
  if (w < P && lane == 0) {
    unsigned state = 12345u + blockIdx.x * 7919u + w * 104729u;
    unsigned char* ring = sm + (size_t)w * depth * row_bytes;
    for (int i = 0; i < iters + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) {
        if (true) asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(&bar[w][s])), "r"((unsigned)((i / depth - 1) & 1)) : "memory");
        else asm volatile("{ .reg .pred p; W2: mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D2; bra W2; D2: }" ::"r"(s32(&bar[w][s])), "r"((unsigned)((i / depth - 1) & 1)) : "memory");
      }
      if (i < iters) {
        state = state * 1664525u + 1013904223u;
        const float* src = g + ((size_t)(state >> 4) % n_rows_total) * (row_bytes / 4);
        long long a = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[w][s])), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)s * row_bytes)), "l"(src), "r"(row_bytes), "r"(s32(&bar[w][s])) : "memory");
        tissue += clock64() - a;
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = tissue; }
}
int main() {
  const size_t bytes = 48ull << 20;
  float* g; long long* out;
  cudaMalloc(&g, bytes); cudaMemset(g, 0, bytes);
  cudaMalloc(&out, 2 * 296 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int spin : {0, 1}) for (int rb : {4096, 16384}) for (int P : {1, 2, 4, 8}) for (int depth : {1, 2, 4, 8}) {
    if ((size_t)P * depth * rb > 200 * 1024 || depth == 1 || depth == 8) continue;
    const int iters = 1000, ctas = 148;
    for (int rep = 0; rep < 2; ++rep) k<<<ctas, 256, P * depth * rb>>>(g, bytes / rb, out, iters, rb, depth, P, spin);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[592]; cudaMemcpy(h, out, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost);
    double s = 0, si = 0; for (int i = 0; i < ctas; ++i) { s += h[2 * i]; si += h[2 * i + 1]; }
    const double cyc = s / ctas;
    printf("%s row %5d B  producers %d depth %d: %7.1f cycles per row per producer (issue alone %6.1f)  -> %6.0f GB/s  %s\n", spin ? "lanes of one warp" : "one lane per warp ", rb, P, depth,
           cyc / iters, si / ctas / iters, (double)ctas * P * iters * rb / (cyc / 1.965e9) / 1e9, cudaGetErrorString(e));
  }
  return 0;
}
