// c2b_sort.cuh — device-wide exclusive scan (u32) and LSD radix sort of (u64 key, u32 value)
// pairs.  Hand-written for this library (no CUB/Thrust): used to order Morton codes for the
// LBVH build and to order candidate (camera, point) keys before ray traversal.
#pragma once
#include "c2b_common.cuh"

namespace c2b {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of one value per thread across a 256-thread block; returns the exclusive
// prefix, *total receives the block sum (valid in every thread)
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  __shared__ uint32_t block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
    if (lane == SCAN_THREADS / 32 - 1) block_total = wi;
  }
  __syncthreads();
  uint32_t r = warp_sums[warp] + incl - v;
  *total = block_total;
  __syncthreads();  // allow the shared arrays to be reused by a following call
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t *__restrict__ in,
                                                             uint64_t n,
                                                             uint32_t *__restrict__ sums) {
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t total;
  block_exclusive_scan_256(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// each thread owns SCAN_ITEMS consecutive elements of the tile
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const uint32_t *in, uint32_t *out,
                                                           uint64_t n, const uint32_t *offsets,
                                                           uint32_t *total_out) {
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan_256(s, &total);
  uint32_t off = offsets ? offsets[blockIdx.x] : 0u;
  uint32_t run = off + ex;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = off + total;
}

// out[i] = sum(in[0..i)), total_out (device, optional) = sum of everything.  in == out allowed.
// tmp: two scratch DevBufs for the upper levels.
inline int exclusive_scan_u32(cudaStream_t st, const uint32_t *in, uint32_t *out, uint64_t n,
                              uint32_t *total_out, DevBuf *tmp, int depth = 0) {
  if (n == 0) {
    if (total_out) C2B_CUDA(cudaMemsetAsync(total_out, 0, 4, st));
    return C2B_OK;
  }
  uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb == 1) {
    k_scan_down<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, total_out);
    C2B_KERNEL_CHECK();
    return C2B_OK;
  }
  if (depth > 2) return set_error(C2B_ERR_INVALID, "scan too deep");
  C2B_TRY(tmp[depth].ensure(nb * 4));
  uint32_t *sums = tmp[depth].as<uint32_t>();
  k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums);
  C2B_KERNEL_CHECK();
  C2B_TRY(exclusive_scan_u32(st, sums, sums, nb, nullptr, tmp, depth + 1));
  k_scan_down<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums, total_out);
  C2B_KERNEL_CHECK();
  return C2B_OK;
}

// ---- radix sort ------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;  // per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

// hist[d * nblk + b] = number of keys of tile b with digit d
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t *__restrict__ keys, uint64_t n,
                                                       int shift, uint32_t *__restrict__ hist,
                                                       uint32_t nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: warp w of the block owns the contiguous sub-range
// [tile + w*RS_ITEMS*32, tile + (w+1)*RS_ITEMS*32), visited in RS_ITEMS steps of 32 keys
__global__ void __launch_bounds__(RS_THREADS)
    k_rs_scatter(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                 uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint64_t n,
                 int shift, const uint32_t *__restrict__ hist_scanned, uint32_t nblk) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();

  uint64_t base = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * (RS_ITEMS * 32);
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint32_t dig[RS_ITEMS];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint64_t i = base + (uint64_t)k * 32 + lane;
    if (i < n) {
      key[k] = keys_in[i];
      val[k] = vals_in[i];
      dig[k] = (uint32_t)(key[k] >> shift) & 255u;
    } else {
      key[k] = 0;
      val[k] = 0;
      dig[k] = 256u;  // sentinel: never counted, never written
    }
  }
  // phase 1: per-warp digit counts
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint32_t peers = __match_any_sync(0xffffffffu, dig[k]);
    if (dig[k] < 256u && (peers & ((1u << lane) - 1u)) == 0u) cnt[warp][dig[k]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase 2: turn counts into global write bases (digit d handled by thread d)
  {
    uint32_t run = hist_scanned[(uint64_t)threadIdx.x * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t t = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
  // phase 3: ranks inside the warp step, then scatter
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint32_t peers = __match_any_sync(0xffffffffu, dig[k]);
    if (dig[k] < 256u) {
      uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t pos = cnt[warp][dig[k]] + rank;
      keys_out[pos] = key[k];
      vals_out[pos] = val[k];
    }
    __syncwarp();
    if (dig[k] < 256u && (peers & ((1u << lane) - 1u)) == 0u) cnt[warp][dig[k]] += __popc(peers);
    __syncwarp();
  }
}

// Sorts n (key, val) pairs by the low `bits` bits of key.  keys[0]/vals[0] hold the input;
// returns which of the two buffers (0/1) holds the sorted output.
inline int radix_sort_pairs(cudaStream_t st, uint64_t *keys[2], uint32_t *vals[2], uint64_t n,
                            int bits, DevBuf &hist, DevBuf *scan_tmp, int *result_buf) {
  *result_buf = 0;
  if (n <= 1 || bits <= 0) return C2B_OK;
  if (n >= 0xffffffffull) return set_error(C2B_ERR_INVALID, "radix sort: too many elements");
  uint32_t nblk = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
  C2B_TRY(hist.ensure((size_t)256 * nblk * 4));
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    k_rs_hist<<<nblk, RS_THREADS, 0, st>>>(keys[cur], n, shift, hist.as<uint32_t>(), nblk);
    C2B_KERNEL_CHECK();
    C2B_TRY(exclusive_scan_u32(st, hist.as<uint32_t>(), hist.as<uint32_t>(), (uint64_t)256 * nblk,
                               nullptr, scan_tmp));
    k_rs_scatter<<<nblk, RS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n,
                                              shift, hist.as<uint32_t>(), nblk);
    C2B_KERNEL_CHECK();
    cur ^= 1;
  }
  *result_buf = cur;
  return C2B_OK;
}

}  // namespace c2b
