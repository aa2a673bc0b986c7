// fiss_pick_exchange.cuh -- the cross-GPU best-cost pick (BASELINE config 5 with few problems: the
// lattice's lateral rows are split across the GPUs of a box, each GPU picks the winner of its slab, and the GPUs agree
// on the global winner).
//
// Rule (frenet_optimal_planner.py:263-268, `min_cost >= cost` scan): the minimum cost, and among equal minima the LAST
// candidate in enumeration order = the largest global candidate id.  A float64 cost plus an id does not fit one 64-bit
// key without dropping cost bits -- and near-equal costs must not be mistaken for ties -- so the all-reduce runs over a
// slot table instead: problem b has one (cost key, id) slot per rank, a rank fills ITS slot and the identity of MIN
// everywhere else; ONE ncclAllReduce(MIN, uint64) then leaves every rank with every slab's winner (16 B x ranks per
// problem, latency-bound on NVLink), and each GPU applies the exact rule to the few slots itself.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fiss_abi.h"

namespace fiss {

// float64 -> uint64 whose unsigned order is the numeric order (-inf < ... < -0 < +0 < ... < +inf < NaNs with the sign bit clear)
__device__ __forceinline__ unsigned long long cost_key(double c) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(c);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double cost_from_key(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
constexpr unsigned long long kNoSlot = 0xffffffffffffffffull;

// table [B][nranks][2]: this rank's slot = (key of its slab's best cost, its GLOBAL candidate id); every other slot and
// "no feasible candidate" = the identity of MIN.
// Local candidate id c -> global id (c / id_inner) * id_outer + c % id_inner + id_offset: a slab of lateral rows is one
// contiguous block (id_inner >= the slab's size); a slab of horizons is nt_local * nv consecutive ids out of every nt * nv.
__global__ void fiss_pick_pack_kernel(int B, int nranks, int rank, long long id_inner, long long id_outer, long long id_offset,
                                      const int32_t* __restrict__ best_idx, const double* __restrict__ best_cost,
                                      unsigned long long* __restrict__ table) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B * nranks) return;
  const int b = q / nranks, r = q - b * nranks;
  unsigned long long key = kNoSlot, id = kNoSlot;
  if (r == rank && best_idx[b] >= 0) {
    key = cost_key(best_cost[b]);
    const long long c = best_idx[b];
    id = (unsigned long long)((c / id_inner) * id_outer + c % id_inner + id_offset);
  }
  table[2 * (size_t)q] = key;
  table[2 * (size_t)q + 1] = id;
}

// After the all-reduce: the global winner of problem b (one CTA per problem) by the exact rule over the nranks slots;
// best_idx / best_cost become global; the exchange block xch [B][rec_words + 1] receives this rank's record words and
// its packed (n, n') if it owns the winner, zeros otherwise.
__global__ void fiss_pick_select_kernel(int B, int nranks, int rank, const unsigned long long* __restrict__ table,
                                        int32_t* __restrict__ best_idx, double* __restrict__ best_cost,
                                        const int32_t* __restrict__ best_meta, const double* __restrict__ records,
                                        int rec_words, unsigned long long* __restrict__ xch) {
  const int b = blockIdx.x;
  __shared__ int s_owner;
  if (threadIdx.x == 0) {
    unsigned long long bk = kNoSlot, bi = kNoSlot;
    int owner = -1;
    for (int r = 0; r < nranks; ++r) {
      const unsigned long long k = table[2 * ((size_t)b * nranks + r)], i = table[2 * ((size_t)b * nranks + r) + 1];
      if (i == kNoSlot) continue;
      if (owner < 0 || k < bk || (k == bk && i > bi)) {  // min cost; ties: the largest id (last one wins)
        bk = k;
        bi = i;
        owner = r;
      }
    }
    s_owner = owner;
    best_idx[b] = owner >= 0 ? (int32_t)bi : -1;
    best_cost[b] = owner >= 0 ? cost_from_key(bk) : CUDART_INF;
  }
  __syncthreads();
  const bool mine = s_owner == rank;
  unsigned long long* x = xch + (size_t)b * (rec_words + 1);
  const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(records) + (size_t)b * rec_words;
  for (int q = threadIdx.x; q < rec_words; q += blockDim.x) x[q] = mine ? rec[q] : 0ull;
  if (threadIdx.x == 0) {
    unsigned long long m = 0ull;
    if (mine && best_meta)
      m = (unsigned long long)(uint32_t)best_meta[2 * b] | ((unsigned long long)(uint32_t)best_meta[2 * b + 1] << 32);
    x[rec_words] = m;
  }
}

// After the exchange: the winner's record and (n, n') on every rank; an all-NaN record where nothing is feasible.
__global__ void fiss_pick_unpack_kernel(int B, const unsigned long long* __restrict__ xch, const int32_t* __restrict__ best_idx,
                                        int32_t* __restrict__ best_meta, double* __restrict__ records, int rec_words) {
  const int b = blockIdx.x;
  const unsigned long long* x = xch + (size_t)b * (rec_words + 1);
  const bool any = best_idx[b] >= 0;
  unsigned long long* rec = reinterpret_cast<unsigned long long*>(records) + (size_t)b * rec_words;
  const unsigned long long nan_bits = (unsigned long long)__double_as_longlong(CUDART_NAN);
  for (int q = threadIdx.x; q < rec_words; q += blockDim.x) rec[q] = any ? x[q] : nan_bits;
  if (threadIdx.x == 0 && best_meta) {
    best_meta[2 * b] = any ? (int32_t)(uint32_t)(x[rec_words] & 0xffffffffull) : 0;
    best_meta[2 * b + 1] = any ? (int32_t)(uint32_t)(x[rec_words] >> 32) : 0;
  }
}

}  // namespace fiss
