// Bitonic sort of (uint64 key, uint32 value) records in global memory, ascending by (key, value).
//
// The sort is the ordering step of batched NMS (descending score, ties by ascending index —
// the order of the CPU op's stable sort).  Records are 12 bytes; a tile of kSortTile records is
// sorted / merged entirely in shared memory, only compare distances >= kSortTile go through
// global memory.  The array length must be a power of two (callers pad with key = ~0).
#pragma once
#include "common.cuh"

namespace dgod {

constexpr int kSortTile = 8192;      // records per CTA tile (96 KB of shared memory)
constexpr int kSortThreads = 1024;
constexpr size_t kSortSmemBytes = (size_t)kSortTile * (sizeof(unsigned long long) + sizeof(uint32_t));

__device__ __forceinline__ bool rec_greater(unsigned long long ka, uint32_t va,
                                            unsigned long long kb, uint32_t vb) {
  return ka > kb || (ka == kb && va > vb);
}

// Sort network steps with compare distance j < tile inside shared memory.
// mode 0: full sort of every tile (k = 2..tile); mode 1: finish stage k_glob (j = tile/2..1).
__global__ void __launch_bounds__(kSortThreads)
bitonic_local_kernel(unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                     int n_pow2, int tile, int mode, int k_glob) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_raw);
  uint32_t* sv = reinterpret_cast<uint32_t*>(sk + tile);
  pdl_wait();                          // the steps of the network are launched behind one another (programmatic dependent launch)
  pdl_trigger();
  const long long base = (long long)blockIdx.x * tile;
  for (int i = threadIdx.x; i < tile; i += blockDim.x) {
    sk[i] = keys[base + i];
    sv[i] = vals[base + i];
  }
  __syncthreads();
  const int k_first = mode == 0 ? 2 : k_glob;
  const int k_last = mode == 0 ? tile : k_glob;
  for (int k = k_first; k <= k_last; k <<= 1) {
    int j0 = (mode == 0 ? k : tile) >> 1;
    for (int j = j0; j > 0; j >>= 1) {
      for (int q = threadIdx.x; q < (tile >> 1); q += blockDim.x) {
        int i = 2 * q - (q & (j - 1));
        int l = i + j;
        bool asc = (((base + i) & (long long)k) == 0);
        unsigned long long ka = sk[i], kb = sk[l];
        uint32_t va = sv[i], vb = sv[l];
        bool gt = rec_greater(ka, va, kb, vb);
        if (gt == asc) {
          sk[i] = kb; sk[l] = ka;
          sv[i] = vb; sv[l] = va;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < tile; i += blockDim.x) {
    keys[base + i] = sk[i];
    vals[base + i] = sv[i];
  }
}

// One network step (k, j) with j >= tile through global memory.
__global__ void __launch_bounds__(256)
bitonic_global_kernel(unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                      int n_pow2, int k, int j) {
  pdl_wait();
  pdl_trigger();
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (n_pow2 >> 1)) return;
  long long i = 2 * q - (q & (long long)(j - 1));
  long long l = i + j;
  bool asc = ((i & (long long)k) == 0);
  unsigned long long ka = keys[i], kb = keys[l];
  uint32_t va = vals[i], vb = vals[l];
  if (rec_greater(ka, va, kb, vb) == asc) {
    keys[i] = kb; keys[l] = ka;
    vals[i] = vb; vals[l] = va;
  }
}

static inline int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// Enqueue the whole sort; returns DGOD_OK or an error code.
static inline int bitonic_sort(unsigned long long* keys, uint32_t* vals, int n_pow2,
                               cudaStream_t st) {
  if (n_pow2 <= 1) return DGOD_OK;
  static bool attr_set = false;  // idempotent; racing threads set the same value
  if (!attr_set) {
    DGOD_CUDA(cudaFuncSetAttribute(bitonic_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kSortSmemBytes));
    attr_set = true;
  }
  const int tile = n_pow2 < kSortTile ? n_pow2 : kSortTile;
  const int threads = tile / 2 < kSortThreads ? (tile / 2 < 32 ? 32 : tile / 2) : kSortThreads;
  const size_t smem = (size_t)tile * 12;
  const int n_tiles = n_pow2 / tile;
  DGOD_CUDA(launch_pdl(bitonic_local_kernel, dim3(n_tiles), dim3(threads), smem, st, keys, vals, n_pow2, tile, 0, 0));
  DGOD_LAUNCHED();
  for (long long k = (long long)tile * 2; k <= n_pow2; k <<= 1) {
    for (long long j = k >> 1; j >= tile; j >>= 1) {
      DGOD_CUDA(launch_pdl(bitonic_global_kernel, dim3(cdiv(n_pow2 / 2, 256)), dim3(256), 0, st, keys, vals, n_pow2, (int)k, (int)j));
      DGOD_LAUNCHED();
    }
    DGOD_CUDA(launch_pdl(bitonic_local_kernel, dim3(n_tiles), dim3(threads), smem, st, keys, vals, n_pow2, tile, 1, (int)k));
    DGOD_LAUNCHED();
  }
  return DGOD_OK;
}

}  // namespace dgod
