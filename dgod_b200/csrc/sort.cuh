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

// ------------------------------------------------------------------------------------------------
// Rank sort for segmented inputs.  A live record (key != ~0) goes to
//     (#live records of earlier segments) + #{live records of its segment that compare smaller},
// dead records (key == ~0: masked-out or padding) go behind all live ones in index order — the same
// global order a full sort of the keys produces when the key carries the segment in its top bits.
// One launch, no dependent steps: O(sum seg_len^2) pair compares — for the NMS shapes of a training
// step (8 images x 4-8 k candidates) ~0.1-0.5 G compares, a few tens of microseconds, against ~20
// dependent launches of the global bitonic network.  seg_live[s] = live records of segment s.
constexpr int kRankSortMaxSeg = 16384;   // longer segments take the bitonic network
constexpr int kRankSortMaxNseg = 1024;
constexpr int kRankTile = 1024;          // keys staged per shared-memory pass

// Input segment s is [seg_offsets[s], seg_offsets[s+1]) or, when live_in is given (the input is the
// output of a previous rank sort), the live range of s in that output: [prefix(live_in)[s], +live_in[s]).
__global__ void __launch_bounds__(256)
rank_sort_kernel(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                 const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ live_in,
                 const int32_t* __restrict__ live_out, int n_seg, int n_total, int n_pow2,
                 unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ __align__(16) unsigned long long s_ck[kRankTile];
  __shared__ int s_base[4];                // live_out before this segment / in total; same for live_in
  const int seg = blockIdx.y;
  if (threadIdx.x < 32) {
    int before = 0, total = 0, before_in = 0, total_in = 0;
    for (int t = threadIdx.x; t < n_seg; t += 32) {
      const int c = live_out[t];
      total += c;
      if (t < seg) before += c;
      if (live_in) {
        const int ci = live_in[t];
        total_in += ci;
        if (t < seg) before_in += ci;
      }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, d);
      total += __shfl_xor_sync(0xffffffffu, total, d);
      before_in += __shfl_xor_sync(0xffffffffu, before_in, d);
      total_in += __shfl_xor_sync(0xffffffffu, total_in, d);
    }
    if (threadIdx.x == 0) { s_base[0] = before; s_base[1] = total; s_base[2] = before_in; s_base[3] = total_in; }
  }
  __syncthreads();
  const int s0 = live_in ? s_base[2] : seg_offsets[seg];
  const int s1 = live_in ? s_base[2] + live_in[seg] : seg_offsets[seg + 1];
  if (seg == 0) {   // records outside every segment stay dead: padding, and the dead tail of the previous sort
    const int tail0 = live_in ? s_base[3] : n_total;
    for (int p = tail0 + blockIdx.x * blockDim.x + threadIdx.x; p < n_pow2; p += gridDim.x * blockDim.x) {
      keys_out[p] = ~0ull;
      vals_out[p] = p < n_total ? vals_in[p] : (uint32_t)p;
    }
  }
  if (s0 + (int)(blockIdx.x * blockDim.x) >= s1) return;      // whole CTA beyond the segment (uniform)
  // One 64-bit compare per pair: inside a segment (key, value) order equals the order of
  //   ck = key[47:0] << 14 | (value - first index of the segment)        (segments <= 16384 records)
  // — the low 48 key bits are group | ~score, the value is the record's original index.
  const uint32_t v0 = (uint32_t)seg_offsets[seg];
  auto compare_key = [&](unsigned long long k, uint32_t v) {
    return ((k & 0xffffffffffffull) << 14) | (unsigned long long)((v - v0) & 0x3fffu);
  };
  const int i = s0 + blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < s1;
  const unsigned long long ki = active ? keys_in[i] : 0ull;
  const uint32_t vi = active ? vals_in[i] : 0u;
  const unsigned long long cki = compare_key(ki, vi);
  int rank = 0;
  for (int c0 = s0; c0 < s1; c0 += kRankTile) {
    const int cnt = min(kRankTile, s1 - c0);
    __syncthreads();
    for (int t = threadIdx.x; t < kRankTile; t += blockDim.x)
      s_ck[t] = t < cnt ? compare_key(keys_in[c0 + t], vals_in[c0 + t]) : ~0ull;   // padding never counts
    __syncthreads();
    if (active) {
      const ulonglong2* __restrict__ p2 = reinterpret_cast<const ulonglong2*>(s_ck);
      const int n2 = (cnt + 1) >> 1;
      int r0 = 0, r1 = 0;
#pragma unroll 8
      for (int t = 0; t < n2; ++t) {   // shared-memory broadcasts, two keys per load
        const ulonglong2 kk = p2[t];
        r0 += kk.x < cki ? 1 : 0;
        r1 += kk.y < cki ? 1 : 0;
      }
      rank += r0 + r1;
    }
  }
  if (active) {
    const int before = s_base[0], total = s_base[1];
    // dead records: behind all live ones; those of earlier segments first, then by index rank
    const int pos = ki != ~0ull ? before + rank : total + (s0 - before) + (rank - live_out[seg]);
    keys_out[pos] = ki;
    vals_out[pos] = vi;
  }
}

static inline int rank_sort(const unsigned long long* keys_in, const uint32_t* vals_in, const int32_t* seg_offsets,
                            const int32_t* live_in, const int32_t* live_out, int n_seg, int n_total, int n_pow2,
                            int max_seg_len, unsigned long long* keys_out, uint32_t* vals_out, cudaStream_t st) {
  dim3 grid(cdiv(max_seg_len > 0 ? max_seg_len : 1, 256), n_seg);
  rank_sort_kernel<<<grid, 256, 0, st>>>(keys_in, vals_in, seg_offsets, live_in, live_out, n_seg, n_total, n_pow2, keys_out,
                                        vals_out);
  DGOD_LAUNCHED();
  return DGOD_OK;
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
  bitonic_local_kernel<<<n_tiles, threads, smem, st>>>(keys, vals, n_pow2, tile, 0, 0);
  DGOD_LAUNCHED();
  for (long long k = (long long)tile * 2; k <= n_pow2; k <<= 1) {
    for (long long j = k >> 1; j >= tile; j >>= 1) {
      bitonic_global_kernel<<<cdiv(n_pow2 / 2, 256), 256, 0, st>>>(keys, vals, n_pow2, (int)k, (int)j);
      DGOD_LAUNCHED();
    }
    bitonic_local_kernel<<<n_tiles, threads, smem, st>>>(keys, vals, n_pow2, tile, 1, (int)k);
    DGOD_LAUNCHED();
  }
  return DGOD_OK;
}

}  // namespace dgod
