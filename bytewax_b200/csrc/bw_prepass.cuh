// bw_prepass.cuh -- lateness verdict for one activation.
//
// The reference decides lateness per item against the *per-key* running
// watermark in arrival order (windowing.py:1120-1130 with the event clock of
// windowing.py:263-287).  With a frozen system clock that watermark is
//     wm_i(key) = max(UTC_MIN, max_{earlier items of key}(ts) - wait).
// A batch-parallel fold is exact iff no item of the batch can be late.  Since
// every per-key maximum is bounded by the global running maximum G_i (over all
// keys, all earlier batches included), the batch is provably clean when
//     ts_i >= G_i - wait     for every item i.
// This pass checks exactly that in one streaming read; a batch that fails is
// handed to the exact sort-based path (bw_slow.cuh).
#pragma once
#include "bw_common.cuh"
#include "bw_fold.cuh"

#define BW_RANGE_ROWS 2048  // rows per warp-range
#define BW_PRE_THREADS 128  // 128 x 62 registers fit beside three resident k_fold blocks: the verdict pass of the next activation co-runs with the fold
#define BW_PRE_MLP 8        // independent loads in flight per lane

__device__ __forceinline__ i64 bw_load_ts(const BatchView& bv, int seg, u64 off, const FoldParams& p) {
  if (p.ts_from_value == 2) return p.align_us;
  if (p.ts_from_value) {
    u64 raw = (p.val_dtype == 2) ? (u64)bw_ld_stream_u32((const u32*)bv.vals[seg] + off)
                                 : bw_ld_stream_u64((const u64*)bv.vals[seg] + off);
    return p.align_us + (i64)raw;
  }
  return (i64)bw_ld_stream_u64((const u64*)bv.ts[seg] + off) - p.now_us;  // (the frame where system time is 0: FoldParams)
}

// one warp per range: exact within-range check + (min, max) for the cross-range scan
__global__ void __launch_bounds__(BW_PRE_THREADS)
k_prepass_ranges(BatchView bv, FoldParams p, i64* range_min, i64* range_max, u32* range_bad) {
  __shared__ u64 seg_start[BW_MAX_WORLD + 1];
  if (threadIdx.x == 0) {
    u64 acc = 0;
    for (int j = 0; j < bv.nseg; ++j) {
      seg_start[j] = acc;
      acc += bw_seg_count(bv, j);
    }
    seg_start[bv.nseg] = acc;
  }
  __syncthreads();
  const u64 total = seg_start[bv.nseg];
  const u64 nranges = (total + BW_RANGE_ROWS - 1) / BW_RANGE_ROWS;
  const int lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 r = warp; r < nranges; r += nwarps) {
    i64 run_max = INT64_MIN, rmin = INT64_MAX;
    bool bad = false;
    for (int c0 = 0; c0 < BW_RANGE_ROWS / 32; c0 += BW_PRE_MLP) {
      // issue BW_PRE_MLP independent coalesced loads, then consume them in order
      i64 tsv[BW_PRE_MLP];
      bool val[BW_PRE_MLP];
#pragma unroll
      for (int j = 0; j < BW_PRE_MLP; ++j) {
        const u64 g = r * BW_RANGE_ROWS + (u64)(c0 + j) * 32 + lane;
        val[j] = g < total;
        tsv[j] = INT64_MIN;
        if (val[j]) {
          int seg = 0;
          u64 off = g;
          if (bv.nseg > 1) {
            // resolve the segment once per 32-row chunk when the chunk lies inside one segment
            const u64 g0 = g - lane;
            int cs = 0;
#pragma unroll
            for (int q = 1; q < BW_MAX_WORLD; ++q)
              if (q < bv.nseg && g0 >= seg_start[q]) cs = q;
            if (cs + 1 >= bv.nseg || g0 + 31 < seg_start[cs + 1]) {
              seg = cs;
              off = g - seg_start[cs];
            } else {
              bw_locate(bv, seg_start, g, seg, off);
            }
          }
          tsv[j] = bw_load_ts(bv, seg, off, p);
        }
      }
#pragma unroll
      for (int j = 0; j < BW_PRE_MLP; ++j) {
        const bool valid = val[j];
        const i64 ts = tsv[j];
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (vmask == 0u) continue;  // uniform
        // cheap case first: the 32 timestamps are already non-decreasing (in-order streams)
        const i64 prev = __shfl_up_sync(0xffffffffu, ts, 1);
        const bool in_order = !valid || lane == 0 || ts >= prev;
        if (__all_sync(0xffffffffu, in_order)) {
          const int last_lane = 31 - __clz(vmask);
          const i64 first = __shfl_sync(0xffffffffu, ts, 0);
          const i64 last = __shfl_sync(0xffffffffu, ts, last_lane);
          if (first < bw_sub_sat(run_max, p.wait_us)) bad = true;
          if (first < rmin) rmin = first;
          if (last > run_max) run_max = last;
          continue;
        }
        i64 incl = ts;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          i64 y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d && y > incl) incl = y;
        }
        i64 excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = INT64_MIN;
        i64 before = run_max > excl ? run_max : excl;
        if (valid && ts < bw_sub_sat(before, p.wait_us)) bad = true;
        i64 mn = valid ? ts : INT64_MAX;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
          i64 y = __shfl_xor_sync(0xffffffffu, mn, d);
          if (y < mn) mn = y;
        }
        if (mn < rmin) rmin = mn;
        i64 cm = __shfl_sync(0xffffffffu, incl, 31);
        if (cm > run_max) run_max = cm;
      }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      range_min[r] = rmin;
      range_max[r] = run_max;
      range_bad[r] = bad ? 1u : 0u;
    }
  }
}

// one block: chain the ranges (and the previous batches through gmax_ts)
__global__ void __launch_bounds__(1024)
k_prepass_scan(BatchView bv, FoldParams p, const i64* range_min, const i64* range_max, const u32* range_bad,
               Counters* ctr, u32* verdict_out, i64* span_out) {
  __shared__ i64 strip_max[1024];
  __shared__ unsigned long long span_min;  // i64 bits, biased by the sign bit so that unsigned min == signed min
  __shared__ int any_bad;
  u64 total = 0;
  for (int j = 0; j < bv.nseg; ++j) total += bw_seg_count(bv, j);
  const u64 nranges = (total + BW_RANGE_ROWS - 1) / BW_RANGE_ROWS;
  const u64 per = (nranges + blockDim.x - 1) / blockDim.x;
  const u64 lo = (u64)threadIdx.x * per, hi = (lo + per < nranges) ? lo + per : nranges;
  if (threadIdx.x == 0) {
    any_bad = 0;
    span_min = ~0ULL;
  }
  i64 m = INT64_MIN, mn = INT64_MAX;
  for (u64 r = lo; r < hi; ++r) {
    if (range_max[r] > m) m = range_max[r];
    if (range_min[r] < mn) mn = range_min[r];
  }
  strip_max[threadIdx.x] = m;
  __syncthreads();
  const i64 gprev = (i64)ctr->gmax_ts;
  // exclusive prefix max over strips (serial per thread over <= 1024 entries is
  // wasteful; do a Hillis-Steele scan instead)
  for (int d = 1; d < (int)blockDim.x; d <<= 1) {
    i64 y = (threadIdx.x >= (unsigned)d) ? strip_max[threadIdx.x - d] : INT64_MIN;
    __syncthreads();
    if (y > strip_max[threadIdx.x]) strip_max[threadIdx.x] = y;
    __syncthreads();
  }
  i64 running = (threadIdx.x == 0) ? gprev : strip_max[threadIdx.x - 1];
  if (running < gprev) running = gprev;
  bool bad = false;
  for (u64 r = lo; r < hi; ++r) {
    if (range_bad[r]) bad = true;
    if (range_min[r] < bw_sub_sat(running, p.wait_us)) bad = true;
    if (range_max[r] > running) running = range_max[r];
  }
  if (bad) atomicExch(&any_bad, 1);
  if (lo < hi) atomicMin(&span_min, (unsigned long long)mn ^ 0x8000000000000000ULL);
  __syncthreads();
  if (threadIdx.x == 0) {
    i64 tot = strip_max[blockDim.x - 1];
    if (tot < gprev) tot = gprev;
    ctr->gmax_ts = (unsigned long long)tot;
    u32 clean = any_bad ? 0u : 1u;
    ctr->batch_clean = clean;
    *verdict_out = clean;
    // time span of this activation alone (the combining fold keeps timestamps relative to its minimum)
    i64 amax = INT64_MIN;
    amax = strip_max[blockDim.x - 1];
    span_out[0] = (i64)(span_min ^ 0x8000000000000000ULL);
    span_out[1] = amax;
  }
}
