// bw_bucket.cuh -- K3, combining variant: bucket the activation by table segment,
// then fold every segment through shared memory.
//
// Same contract as k_fold (bw_fold.cuh; replaces the per-key `on_batch` loop of
// src/operators.rs:755-806 over `_WindowLogic.on_batch`, windowing.py:1115-1133),
// used when an activation brings many events per key.  The direct kernel pays
// one random 32-byte L2 read and two L2 reductions per EVENT; with E events per
// key per activation the L2 atomic units serialise on the same sectors (measured:
// 125 k keys per rank -> 1.05 ms for 2^24 rows against 0.54 ms at 1 M keys).
// Here the activation is first scattered into buckets of BW_BKT_SLOTS
// consecutive home slots (one streaming pass); a block then owns a bucket,
// accumulates its events in shared memory with native 32-bit atomics, and
// touches the table once per (key, activation) instead of once per event.
//
//   k_bkt_hist     per-tile histogram of home-slot buckets
//   k_bkt_scan     per bucket: exclusive scan over tiles, bucket totals
//   k_bkt_base     exclusive scan over buckets -> bucket offsets
//   k_bkt_scatter  shared-memory staged scatter (bucket-major runs); carries the
//                  arrival index of every row, so first-open order
//                  (windowing.py:1087-1108) is exactly that of the direct kernel
//   k_fold_seg     one block per bucket: shared-memory fold + one merge per slot
#pragma once
#include "bw_common.cuh"
#include "bw_fold.cuh"

#define BW_BKT_MAX 1024      // buckets per table (beyond: the direct kernel is used)
#define BW_BKT_THREADS 512
#define BW_SEG_THREADS 384
#define BW_SEG_WARPS (BW_SEG_THREADS / 32)
#define BW_SEG_UNROLL 4
#define BW_SEG_SINK_CAP 512
#define BW_SEG_SMEM ((size_t)BW_BKT_SLOTS * 24)

struct BktBufs {
  u64* keys;         // bucketed copy of the activation
  void* vals;
  i64* ts;
  u32* g;            // arrival index of every bucketed row
  u32* tile_counts;  // [nb][tiles_cap]
  u32* cnt;          // [nb]
  u32* off;          // [nb + 1]
  u32 nb;
  u32 tiles_cap;
  int val_bytes;     // 0, 4 or 8
};

__device__ __forceinline__ u32 bw_bucket_of(u64 key, u64 cap) {
  if (key == BW_EMPTY_KEY) return 0u;  // alias slot: never combined, any bucket will do
  return (u32)(bw_slot_of_hash(bw_mix64(key), cap) >> BW_BKT_SHIFT);
}
__device__ __forceinline__ void bw_seg_starts(const BatchView& bv, u64* seg_start) {
  if (threadIdx.x == 0) {
    u64 acc = 0;
    for (int j = 0; j < bv.nseg; ++j) {
      seg_start[j] = acc;
      acc += bw_seg_count(bv, j);
    }
    seg_start[bv.nseg] = acc;
  }
  __syncthreads();
}
__device__ __forceinline__ void bw_seg_of(const BatchView& bv, const u64* seg_start, u64 g, int& seg, u64& off) {
  seg = 0;
  off = g;
  if (bv.nseg > 1) {
#pragma unroll
    for (int j = 1; j < BW_MAX_WORLD; ++j)
      if (j < bv.nseg && g >= seg_start[j]) seg = j;
    off = g - seg_start[seg];
  }
}

template <int RPT>
__global__ void __launch_bounds__(BW_BKT_THREADS) k_bkt_hist(BatchView bv, u64 cap, BktBufs B) {
  __shared__ u32 cnt[BW_BKT_MAX];
  __shared__ u64 seg_start[BW_MAX_WORLD + 1];
  bw_seg_starts(bv, seg_start);
  const u64 total = seg_start[bv.nseg];
  const u64 T = (u64)BW_BKT_THREADS * RPT;
  const u64 ntiles = (total + T - 1) / T;
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (u32 d = threadIdx.x; d < B.nb; d += BW_BKT_THREADS) cnt[d] = 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const u64 g = tile * T + (u64)j * BW_BKT_THREADS + threadIdx.x;
      if (g < total) {
        int seg;
        u64 off;
        bw_seg_of(bv, seg_start, g, seg, off);
        atomicAdd(&cnt[bw_bucket_of(__ldg(bv.keys[seg] + off), cap)], 1u);
      }
    }
    __syncthreads();
    for (u32 d = threadIdx.x; d < B.nb; d += BW_BKT_THREADS) B.tile_counts[(size_t)d * B.tiles_cap + tile] = cnt[d];
    __syncthreads();
  }
}

// one block per bucket: tile counts -> exclusive offsets in place, bucket total
__global__ void __launch_bounds__(1024) k_bkt_scan(BatchView bv, BktBufs B, u32 tile_rows) {
  __shared__ u32 strip[1024];
  u64 total = 0;
  for (int j = 0; j < bv.nseg; ++j) total += bw_seg_count(bv, j);
  const u32 ntiles = (u32)((total + tile_rows - 1) / tile_rows);
  const u32 per = (ntiles + blockDim.x - 1) / blockDim.x;
  const u32 lo = threadIdx.x * per, hi = (lo + per < ntiles) ? lo + per : ntiles;
  for (u32 d = blockIdx.x; d < B.nb; d += gridDim.x) {
    u32* row = B.tile_counts + (size_t)d * B.tiles_cap;
    u32 s = 0;
    for (u32 t = lo; t < hi; ++t) s += row[t];
    strip[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      u32 y = (threadIdx.x >= (unsigned)o) ? strip[threadIdx.x - o] : 0u;
      __syncthreads();
      strip[threadIdx.x] += y;
      __syncthreads();
    }
    u32 run = threadIdx.x ? strip[threadIdx.x - 1] : 0u;
    for (u32 t = lo; t < hi; ++t) {
      u32 c = row[t];
      row[t] = run;
      run += c;
    }
    if (threadIdx.x == blockDim.x - 1) B.cnt[d] = strip[threadIdx.x];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_bkt_base(BktBufs B) {
  __shared__ u32 strip[1024];
  const u32 v = (threadIdx.x < B.nb) ? B.cnt[threadIdx.x] : 0u;
  strip[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    u32 y = (threadIdx.x >= (unsigned)o) ? strip[threadIdx.x - o] : 0u;
    __syncthreads();
    strip[threadIdx.x] += y;
    __syncthreads();
  }
  if (threadIdx.x < B.nb) B.off[threadIdx.x] = strip[threadIdx.x] - v;
  if (threadIdx.x == 1023) B.off[B.nb] = strip[1023];
}

// dynamic shared memory of k_bkt_scatter<RPT>
__host__ __device__ __forceinline__ size_t bw_bkt_scatter_smem(int rpt, int val_bytes, bool has_ts) {
  const size_t T = (size_t)BW_BKT_THREADS * rpt;
  return T * (8 + (has_ts ? 8 : 0) + (size_t)val_bytes + 4 + 2);
}

template <int RPT>
__global__ void __launch_bounds__(BW_BKT_THREADS) k_bkt_scatter(BatchView bv, u64 cap, BktBufs B) {
  __shared__ u32 cnt[BW_BKT_MAX];
  __shared__ u32 dstart[BW_BKT_MAX];
  __shared__ u32 wsum[BW_BKT_THREADS / 32];
  __shared__ u64 seg_start[BW_MAX_WORLD + 1];
  extern __shared__ __align__(16) unsigned char stage_raw[];
  constexpr u32 T = BW_BKT_THREADS * RPT;
  const bool has_ts = bv.ts[0] != nullptr;
  const int vb = B.val_bytes;
  u64* s_keys = (u64*)stage_raw;
  i64* s_ts = (i64*)(s_keys + T);
  unsigned char* s_vals = (unsigned char*)(s_ts + (has_ts ? T : 0));
  u32* s_g = (u32*)(s_vals + (size_t)T * vb);
  unsigned short* s_d = (unsigned short*)(s_g + T);
  bw_seg_starts(bv, seg_start);
  const u64 total = seg_start[bv.nseg];
  const u64 ntiles = (total + T - 1) / T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (u32 d = threadIdx.x; d < B.nb; d += BW_BKT_THREADS) cnt[d] = 0u;
    __syncthreads();
    u64 key[RPT];
    u32 meta[RPT];  // bucket << 16 | rank inside (tile, bucket)
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const u64 g = tile * T + (u64)j * BW_BKT_THREADS + threadIdx.x;
      key[j] = 0;
      meta[j] = 0;
      if (g < total) {
        int seg;
        u64 off;
        bw_seg_of(bv, seg_start, g, seg, off);
        key[j] = bw_ld_stream_u64(bv.keys[seg] + off);
        const u32 d = bw_bucket_of(key[j], cap);
        meta[j] = (d << 16) | atomicAdd(&cnt[d], 1u);
      }
    }
    __syncthreads();
    {  // exclusive scan of cnt[0, nb) -> dstart; two entries per thread
      const u32 i0 = 2u * threadIdx.x, i1 = i0 + 1u;
      const u32 a = (i0 < B.nb) ? cnt[i0] : 0u, b = (i1 < B.nb) ? cnt[i1] : 0u;
      const u32 v = a + b;
      u32 inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        u32 w = (lane < BW_BKT_THREADS / 32) ? wsum[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          u32 y = __shfl_up_sync(0xffffffffu, winc, o);
          if (lane >= o) winc += y;
        }
        if (lane < BW_BKT_THREADS / 32) wsum[lane] = winc - w;
      }
      __syncthreads();
      const u32 excl = inc - v + wsum[warp];
      if (i0 < B.nb) dstart[i0] = excl;
      if (i1 < B.nb) dstart[i1] = excl + a;
    }
    __syncthreads();
    // cnt[d] <- (global position of the bucket's run for this tile) - dstart[d]
    for (u32 d = threadIdx.x; d < B.nb; d += BW_BKT_THREADS)
      cnt[d] = B.off[d] + B.tile_counts[(size_t)d * B.tiles_cap + tile] - dstart[d];
    // stage bucket-major
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const u64 g = tile * T + (u64)j * BW_BKT_THREADS + threadIdx.x;
      if (g < total) {
        int seg;
        u64 off;
        bw_seg_of(bv, seg_start, g, seg, off);
        const u32 d = meta[j] >> 16;
        const u32 lp = dstart[d] + (meta[j] & 0xFFFFu);
        s_keys[lp] = key[j];
        if (vb == 8) ((u64*)s_vals)[lp] = bw_ld_stream_u64((const u64*)bv.vals[seg] + off);
        else if (vb == 4) ((u32*)s_vals)[lp] = bw_ld_stream_u32((const u32*)bv.vals[seg] + off);
        if (has_ts) s_ts[lp] = (i64)bw_ld_stream_u64((const u64*)bv.ts[seg] + off);
        s_g[lp] = (u32)g;
        s_d[lp] = (unsigned short)d;
      }
    }
    __syncthreads();
    const u32 n_tile = (u32)((total - tile * T < (u64)T) ? total - tile * T : (u64)T);
    for (u32 e = threadIdx.x; e < n_tile; e += BW_BKT_THREADS) {
      const u32 pos = cnt[s_d[e]] + e;
      B.keys[pos] = s_keys[e];
      if (vb == 8) ((u64*)B.vals)[pos] = ((const u64*)s_vals)[e];
      else if (vb == 4) ((u32*)B.vals)[pos] = ((const u32*)s_vals)[e];
      if (has_ts) B.ts[pos] = s_ts[e];
      B.g[pos] = s_g[e];
    }
    __syncthreads();
  }
}

struct SegSinks : DirtySink {
  u32 dirty[BW_SEG_SINK_CAP];
  u32 n_defer[BW_SEG_WARPS];
  u32 dq[BW_SEG_WARPS][32 * BW_SEG_UNROLL];  // bucketed row index | known-slot flag << 31
};

// One block per bucket (grid-strided).  Events whose key sits in its home slot with the
// event in pane 0 or pane 1 are combined in shared memory; everything else takes the
// general path of the direct kernel (which also combines when it lands in this segment).
template <class C>
__global__ void __launch_bounds__(BW_SEG_THREADS, 2)
k_fold_seg(BktBufs B, Table t, FoldParams p, u32 batch_no, i64 base_ts) {
  extern __shared__ __align__(16) unsigned char seg_raw[];
  __shared__ SegSinks sinks;
  SegSink sg;
  sg.acc0 = (u64*)seg_raw;
  sg.acc1 = sg.acc0 + BW_BKT_SLOTS;
  sg.mts = (u32*)(sg.acc1 + BW_BKT_SLOTS);
  sg.seq1 = sg.mts + BW_BKT_SLOTS;
  sg.base_ts = base_ts;
  sg.slot_base = 0;
  if (threadIdx.x == 0) {
    sinks.n_dirty = 0;
    sinks.n_new_keys = 0;
    sinks.cap = BW_SEG_SINK_CAP;
    sinks.buf = sinks.dirty;
  }
  if (threadIdx.x < BW_SEG_WARPS) sinks.n_defer[threadIdx.x] = 0;
  const int op = C::op(p);
  const u64 ident = (op <= BW_OP_ADD_F64) ? 0ULL : p.acc_identity;
  const u32 born = batch_no & 63u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr u32 TILE = BW_SEG_THREADS * BW_SEG_UNROLL;
  PaneCache pcache;
  pcache.lo = INT64_MAX;
  pcache.q = 0;
  for (u32 b = blockIdx.x; b < B.nb; b += gridDim.x) {
    sg.slot_base = (u64)b << BW_BKT_SHIFT;
    for (u32 i = threadIdx.x; i < BW_BKT_SLOTS; i += BW_SEG_THREADS) {
      sg.acc0[i] = ident;
      sg.acc1[i] = ident;
      sg.mts[i] = 0u;
      sg.seq1[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    const u32 e0 = B.off[b], e1 = B.off[b + 1];
    for (u32 base = e0; base < e1; base += TILE) {
      const u32 wbase = base + (u32)warp * (32 * BW_SEG_UNROLL);
      u64 key[BW_SEG_UNROLL], raw[BW_SEG_UNROLL];
      i64 tsv[BW_SEG_UNROLL];
      u32 gi[BW_SEG_UNROLL];
#pragma unroll
      for (int u = 0; u < BW_SEG_UNROLL; ++u) {
        const u32 e = wbase + (u32)u * 32 + lane;
        key[u] = 0;
        raw[u] = 0;
        tsv[u] = 0;
        gi[u] = 0;
        if (e < e1) {
          key[u] = bw_ld_stream_u64(B.keys + e);
          if (B.val_bytes == 8) raw[u] = bw_ld_stream_u64((const u64*)B.vals + e);
          else if (B.val_bytes == 4) raw[u] = (u64)bw_ld_stream_u32((const u32*)B.vals + e);
          tsv[u] = p.ts_from_value ? p.align_us + (i64)raw[u] : (i64)bw_ld_stream_u64((const u64*)B.ts + e);
          gi[u] = bw_ld_stream_u32(B.g + e);
        }
      }
      u64 k0[BW_SEG_UNROLL], a0[BW_SEG_UNROLL];
      i64 mts[BW_SEG_UNROLL], tag0[BW_SEG_UNROLL];
      u32 slot[BW_SEG_UNROLL];
#pragma unroll
      for (int u = 0; u < BW_SEG_UNROLL; ++u) {
        slot[u] = (u32)bw_home_slot(t, key[u]);
        bw_ld_slot(t.hot + slot[u], k0[u], mts[u], tag0[u], a0[u]);
      }
#pragma unroll
      for (int u = 0; u < BW_SEG_UNROLL; ++u) {
        const u32 e = wbase + (u32)u * 32 + lane;
        if (e >= e1) continue;
        const bool known = (k0[u] == key[u]);
        if (!(known && bw_try_fast<C, SegSink>(t, p, &sinks, sg, slot[u], tag0[u], mts[u], tsv[u], raw[u],
                                               ((u64)batch_no << 32) | gi[u], born, pcache))) {
          const u32 i = atomicAdd(&sinks.n_defer[warp], 1u);
          sinks.dq[warp][i] = e | (known ? 0x80000000u : 0u);
        }
      }
      __syncwarp();
      const u32 nd = sinks.n_defer[warp];
      for (u32 i = lane; i < nd; i += 32) {
        const u32 w = sinks.dq[warp][i];
        const u32 e = w & 0x7FFFFFFFu;
        const u64 kk = bw_ld_stream_u64(B.keys + e);
        u64 rw = 0;
        if (B.val_bytes == 8) rw = bw_ld_stream_u64((const u64*)B.vals + e);
        else if (B.val_bytes == 4) rw = (u64)bw_ld_stream_u32((const u32*)B.vals + e);
        const i64 ts = p.ts_from_value ? p.align_us + (i64)rw : (i64)bw_ld_stream_u64((const u64*)B.ts + e);
        const u64 seq = ((u64)batch_no << 32) | bw_ld_stream_u32(B.g + e);
        u32 ks = (w >> 31) ? (u32)bw_home_slot(t, kk) : BW_NO_SLOT;
        if (ks == BW_NO_SLOT) {
          i64 m2, w2;
          ks = bw_lookup_slot(t, kk, m2, w2);
          if (ks != BW_NO_SLOT && bw_try_fast<C, SegSink>(t, p, &sinks, sg, ks, w2, m2, ts, rw, seq, born, pcache)) continue;
        }
        u64 operand;
        bw_operand(p, rw, operand);
        bw_fold_event<C, SegSink>(t, p, &sinks, kk, ts, operand, seq, batch_no, ks, sg);
      }
      __syncwarp();
      if (lane == 0) sinks.n_defer[warp] = 0;
      __syncwarp();
    }
    __syncthreads();
    // merge: one table update per touched slot
    for (u32 ls = threadIdx.x; ls < BW_BKT_SLOTS; ls += BW_SEG_THREADS) {
      const u32 rel = sg.mts[ls];
      if (!rel) continue;
      const u64 s = sg.slot_base + ls;
      u64 kk, aa;
      i64 mts, tag0;
      bw_ld_slot(t.hot + s, kk, mts, tag0, aa);
      const u64 d0 = sg.acc0[ls], d1 = sg.acc1[ls];
      if (d0 != ident) bw_merge(op, &t.hot[s].acc0, d0);
      if (d1 != ident) bw_merge(op, &t.p1[s].acc1, d1);
      const u32 sq = sg.seq1[ls];
      if (sq != 0xFFFFFFFFu) bw_red_min_u64(&t.p1[s].seq1, ((u64)batch_no << 32) | sq);
      const i64 ts = base_ts + (i64)(rel - 1u);
      i64 rem;
      const i64 q = bw_pane_of_r(ts, p, rem);
      bw_after_fold<C>(t, p, &sinks, s, ts, mts, tag0, false, q, rem);
    }
    __syncthreads();
    if (sinks.n_dirty > BW_SEG_SINK_CAP / 2) bw_sinks_flush(&sinks, t);  // uniform: read after the barrier
  }
  __syncthreads();
  bw_sinks_flush(&sinks, t);
}
