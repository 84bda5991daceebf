// Microbenchmark: throughput of the bulk async-copy engine for the access shapes the RoIAlign
// kernels use (rows of 1..32 KB scattered over a few hundred MB):
//   mode 0  cp.async.bulk global->shared (mbarrier), D rows in flight per CTA
//   mode 1  cp.async.bulk shared->global store, D groups in flight
//   mode 2  cp.reduce.async.bulk .add.f32 shared->global, D groups in flight
//   mode 3  red.global.add.v4.f32 from the lanes (same bytes)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bulk tma_bulk.cu ; run: ./tma_bulk
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* g, size_t n_rows_total, int row_bytes, int depth, int rows_per_cta, unsigned seed) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar[16];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < depth * row_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  unsigned state = seed + blockIdx.x * 7919u;
  auto next_row = [&]() { state = state * 1664525u + 1013904223u; return (size_t)(state >> 4) % n_rows_total; };
  const size_t row_floats = row_bytes / 4;
  if (MODE == 0) {
    unsigned phase = 0;   // bit s = parity of stage s
    float acc = 0.f;
    for (int i = 0; i < rows_per_cta + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) {   // consume stage s (issued depth iterations ago)
        asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(&bar[s])), "r"((phase >> s) & 1u) : "memory");
        phase ^= 1u << s;
        acc += reinterpret_cast<float*>(sm + (size_t)s * row_bytes)[tid];
        __syncthreads();
      }
      if (i < rows_per_cta && tid == 0) {
        const float* src = g + next_row() * row_floats;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * row_bytes)), "l"(src), "r"(row_bytes), "r"(s32(&bar[s])) : "memory");
      }
    }
    if (acc == 12345.f) g[0] = acc;
  } else if (MODE == 1 || MODE == 2) {
    if (tid == 0) {
      for (int i = 0; i < rows_per_cta; ++i) {
        const int s = i % depth;
        float* dst = g + next_row() * row_floats;
        if (MODE == 1) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(sm + (size_t)s * row_bytes)), "r"(row_bytes) : "memory");
        else asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(s32(sm + (size_t)s * row_bytes)), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // keep at most `depth` groups in flight (reads of the source buffer)
        switch (depth) {
          case 1: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
          case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
          case 4: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
          default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    for (int i = 0; i < rows_per_cta; ++i) {
      float* dst = g + next_row() * row_floats;
      for (int j = tid * 4; j < (int)row_floats; j += blockDim.x * 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(dst + j), "f"(1.f) : "memory");
    }
  }
}

template <int MODE>
float run(float* g, size_t total_bytes, int row_bytes, int depth, int ctas, int rows_per_cta) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, depth * row_bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<ctas, 256, depth * row_bytes>>>(g, total_bytes / row_bytes, row_bytes, depth, rows_per_cta, 1u);
  cudaEventRecord(e0);
  k<MODE><<<ctas, 256, depth * row_bytes>>>(g, total_bytes / row_bytes, row_bytes, depth, rows_per_cta, 2u);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); exit(1); }
  return ms;
}

int main(int argc, char** argv) {
  const size_t total = (size_t)(argc > 1 ? atoi(argv[1]) : 400) << 20;   // MB; 400 ~ 8 images of FPN gradients, 48 ~ one image (L2-resident)
  printf("target region %zu MB\n", total >> 20);
  float* g;
  cudaMalloc(&g, total);
  cudaMemset(g, 0, total);
  const char* names[] = {"bulk load ", "bulk store", "bulk red  ", "red.v4    "};
  const bool quick = argc > 2;
  for (int row_kb : {2, 8, 16, 32}) {
    for (int depth : {1, 2, 4, 8}) {
      if (quick && (depth == 1 || depth == 8 || row_kb == 2 || row_kb == 32)) continue;
      if (depth * row_kb > 200) continue;
      for (int cps : {1, 2, 4}) {   // CTAs per SM
        if (cps * depth * row_kb > 220) continue;
        const int ctas = 148 * cps, rows = 2048 / cps * 16 / row_kb / 4 + 8;
        const double bytes = (double)ctas * rows * row_kb * 1024;
        float t[4];
        t[0] = run<0>(g, total, row_kb * 1024, depth, ctas, rows);
        t[1] = run<1>(g, total, row_kb * 1024, depth, ctas, rows);
        t[2] = run<2>(g, total, row_kb * 1024, depth, ctas, rows);
        t[3] = depth == 1 ? run<3>(g, total, row_kb * 1024, depth, ctas, rows) : 0.f;
        printf("row %2d KB depth %d ctas/SM %d: ", row_kb, depth, cps);
        for (int m = 0; m < 4; ++m)
          if (t[m] > 0) printf("%s %7.0f GB/s  ", names[m], bytes / t[m] / 1e6);
        printf("\n");
      }
    }
  }
  return 0;
}
