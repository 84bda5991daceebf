// cycles per iteration of the synchronisation primitives used by the RoIAlign backward row loop
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void __launch_bounds__(192, 2) k(float* g, long long* out, int iters, int row_bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4 * row_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.f;
  __syncthreads();
  float* dst = g + (size_t)blockIdx.x * 65536;
  long long t0 = clock64();
  if (tid < 128) {
    for (int i = 0; i < iters; ++i) {
      if (MODE == 0) { if (tid == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
      if (MODE == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (MODE == 2) asm volatile("bar.sync 1, 128;" ::: "memory");
      if (MODE == 3) {  // the full per-row protocol of the kernel: fence, wait_read, barrier, RED, commit
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (i & 7) * 8192), "r"(s32(sm + (i % 3) * row_bytes)), "r"(row_bytes) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (MODE == 4) {  // RED + commit only, thread 0
        if (tid == 0) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (i & 7) * 8192), "r"(s32(sm + (i % 3) * row_bytes)), "r"(row_bytes) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
      }
      if (MODE == 5) {  // RED, commit every 4th
        if (tid == 0) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (i & 7) * 8192), "r"(s32(sm + (i % 3) * row_bytes)), "r"(row_bytes) : "memory");
          if ((i & 3) == 3) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
        }
      }
      if (MODE == 6) { if (tid == 0) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } }
    }
    if (tid == 0) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, float* g, long long* out, int ctas, int row_bytes) {
  const int iters = 2000, smem = 4 * row_bytes;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE><<<ctas, 192, smem>>>(g, out, iters, row_bytes);
  cudaDeviceSynchronize();
  k<MODE><<<ctas, 192, smem>>>(g, out, iters, row_bytes);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, out, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < ctas; ++i) s += h[i];
  printf("%-34s ctas %3d row %5d B: %8.1f cycles/iter %s\n", name, ctas, row_bytes, s / ctas / iters, cudaGetErrorString(e));
}
int main() {
  float* g; long long* out;
  cudaMalloc(&g, (size_t)296 * 65536 * 4); cudaMemset(g, 0, (size_t)296 * 65536 * 4);
  cudaMalloc(&out, 296 * 8);
  for (int ctas : {148, 296}) for (int rb : {4096, 16384}) {
    run<0>("empty commit_group", g, out, ctas, rb);
    run<6>("empty commit + wait_read 0", g, out, ctas, rb);
    run<1>("fence.proxy.async (128 thr)", g, out, ctas, rb);
    run<2>("bar.sync 1,128", g, out, ctas, rb);
    run<3>("fence+wait+bar+RED+commit", g, out, ctas, rb);
    run<4>("RED+commit+wait_read 2", g, out, ctas, rb);
    run<5>("RED, commit+wait every 4th", g, out, ctas, rb);
  }
  return 0;
}
