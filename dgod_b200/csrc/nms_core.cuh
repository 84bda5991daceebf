// Core of the bitmask NMS shared by the generic batched op (nms.cu) and the RPN pipeline
// (rpn.cu).  Both present their candidates in *processing order*: positions 0..n_pos-1 with a
// non-decreasing 32-bit run key; a run (maximal range of equal keys) is one independent greedy
// NMS problem whose members are already in descending score order.
#pragma once
#include "common.cuh"

namespace dgod {

constexpr uint32_t kNoRun = 0xffffffffu;  // run key of padding positions behind every run
// Masked-out positions that sit between runs carry (key of the preceding runs' segment) | 0xffff, so that the
// keys stay non-decreasing; a key whose low half is 0xffff is never a run (kNoRun included).
__host__ __device__ __forceinline__ bool is_norun(uint32_t k) { return (k & 0xffffu) == 0xffffu; }
__host__ __device__ __forceinline__ uint32_t dead_key(int segment) { return ((uint32_t)segment << 16) | 0xffffu; }

// Words per mask row: a row covers column chunks [p/64, p/64 + words).
static inline int nms_mask_row_words(int max_run_len) { return max_run_len / 64 + 2; }
static inline size_t nms_mask_rows(size_t n_pos) { return (n_pos + 63) / 64 * 64; }

// Chunk-major layout [chunk][word][row]: mask[((p/64)*row_words + w)*64 + p%64] bit b  <=>  position
// q = (p/64 + w)*64 + b is in p's run, q > p, and IoU(p,q) > threshold.  (The 64 rows of a chunk write one word
// index as 512 contiguous bytes, and the scan streams a chunk's words 1..n with one bulk copy.)
// diag_cols[q] (one word per position) = transpose of the diagonal block: the rows of q's own
// 64-chunk that suppress q.
int launch_nms_mask(const float4* sbox, const uint32_t* runkey, int n_pos, int max_run_len,
                    float thr_rounded_down, unsigned long long* mask, unsigned long long* diag_cols,
                    cudaStream_t st);

// Sequential greedy pass per run.  alive (optional, per position) = 0 removes a candidate
// before NMS.  keepbits (optional) must be zeroed by the caller.  compact_pos (optional): kept
// positions of a run written contiguously from the run's first position; run_count (optional) is
// indexed by run key.  The mask buffer must hold whole 64-row chunks: nms_mask_rows(n_pos) rows.
int launch_nms_scan(const unsigned long long* mask, const unsigned long long* diag_cols, const uint32_t* runkey,
                    const uint8_t* alive, int n_pos, int max_run_len, unsigned long long* keepbits,
                    int32_t* compact_pos, int32_t* run_count, cudaStream_t st);

}  // namespace dgod
