// bw_snapshot.cuh -- dump / reload of the per-key window state.
//
// What `_WindowLogic.snapshot` (pysrc/bytewax/operators/windowing.py:1182-1190,
// `_WindowSnapshot` :1032-1037) captures per key -- the clock state (max event
// time -> watermark base, windowing.py:224-227), the open windows and their
// accumulators -- as one columnar row per live (key, pane).  Loading re-inserts
// the rows through the same structural paths as the fold (claim the key, pane 0 /
// pane 1 / overflow list), so the target may have any capacity and, with
// world > 1, keeps only the keys it owns: the reference's rescale-on-resume
// (src/recovery.rs:1701-1781 re-hashes snapshots to the new worker count).
#pragma once
#include "bw_common.cuh"
#include "bw_fold.cuh"
#include "bw_close.cuh"

struct SnapCols {
  u64* key;
  i64* pane;
  u64* acc;
  u64* cnt;
  u64* seq;
  i64* max_ts;
  i64* closed_upto;
  u64 cap;  // rows the columns can hold
};

__device__ __forceinline__ u32 bw_snap_panes_of(const Table& t, u64 s) {
  if (t.hot[s].wt0 == BW_EMPTY_WIDTAG) return 0u;
  u32 n = 1u + (t.p1[s].seq1 != ~0ULL ? 1u : 0u);
  for (u32 nd = t.aux[s].spill_head; nd; nd = t.nodes[nd].next) ++n;
  return n;
}

__global__ void k_snap_count(Table t, unsigned long long* total) {
  unsigned long long mine = 0;
  for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s <= t.cap; s += (u64)gridDim.x * blockDim.x) {
    if (s < t.cap && t.hot[s].key == BW_EMPTY_KEY) continue;
    mine += bw_snap_panes_of(t, s);
  }
  if (mine) atomicAdd(total, mine);
}

__global__ void k_snap_fill(Table t, SnapCols c, unsigned long long* cursor) {
  for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s <= t.cap; s += (u64)gridDim.x * blockDim.x) {
    if (s < t.cap && t.hot[s].key == BW_EMPTY_KEY) continue;
    const u32 n = bw_snap_panes_of(t, s);
    if (!n) continue;
    u64 i = atomicAdd(cursor, (unsigned long long)n);
    if (i + n > c.cap) continue;  // cannot happen: sized by k_snap_count on a quiescent table
    const HotSlot h = t.hot[s];
    const AuxSlot ax = t.aux[s];
    const i64 cu = t.closed_upto[s];
    auto put = [&](i64 q, u64 acc, u64 cnt, u64 seq) {
      c.key[i] = h.key;
      c.pane[i] = q;
      c.acc[i] = acc;
      c.cnt[i] = cnt;
      c.seq[i] = seq;
      c.max_ts[i] = h.max_ts;
      c.closed_upto[i] = cu;
      ++i;
    };
    put(bw_widtag_q(h.wt0), h.acc0, ax.cnt0, ax.seq0);
    const P1Slot p1 = t.p1[s];
    if (p1.seq1 != ~0ULL) put(bw_widtag_q1(h.wt0), p1.acc1, ax.cnt1, p1.seq1);
    for (u32 nd = ax.spill_head; nd; nd = t.nodes[nd].next) put(t.nodes[nd].wid, t.nodes[nd].acc, t.node_acc2[nd], t.nodes[nd].open_seq);
  }
}

// one thread per snapshot row
__global__ void __launch_bounds__(256) k_snap_load(Table t, FoldParams p, SnapCols c, u64 n, u32 batch_no, int world, int rank) {
  __shared__ DirtySink sink;
  __shared__ u32 sink_buf[256];
  if (threadIdx.x == 0) {
    sink.n_dirty = 0;
    sink.n_new_keys = 0;
    sink.cap = 256;
    sink.buf = sink_buf;
  }
  __syncthreads();
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const u64 key = c.key[i];
    if (world <= 1 || (int)bw_route_hash(bw_mix64(key), (u32)world) == rank) {
      const i64 q = c.pane[i];
      i64 mts, tag0;
      const u64 s = bw_find_slot(t, &sink, key, mts, tag0);
      if (s == ~0ULL) {
        bw_raise(t.ctr, 3u);
      } else {
        HotSlot* hs = t.hot + s;
        if (tag0 == BW_EMPTY_WIDTAG) {
          const i64 mine = bw_pack_widtag(q, BW_TAG_DELTA_MAX, BW_TAG_STALE);
          const i64 old = (i64)atomicCAS((unsigned long long*)&hs->wt0, (unsigned long long)BW_EMPTY_WIDTAG, (unsigned long long)mine);
          tag0 = (old == BW_EMPTY_WIDTAG) ? mine : old;
        }
        const int op = p.op;
        if (bw_widtag_q(tag0) == q) {
          bw_merge(op, &hs->acc0, c.acc[i]);
          bw_red_add_u64(&t.aux[s].cnt0, c.cnt[i]);
          bw_red_min_u64(&t.aux[s].seq0, c.seq[i]);
        } else if (bw_widtag_q1(tag0) == q) {
          bw_merge(op, &t.p1[s].acc1, c.acc[i]);
          bw_red_add_u64(&t.aux[s].cnt1, c.cnt[i]);
          bw_red_min_u64(&t.p1[s].seq1, c.seq[i]);
        } else {
          bool created = false;
          const u32 nd = bw_spill_node(t, p, s, q, batch_no, created);
          if (nd) {
            bw_merge(op, &t.nodes[nd].acc, c.acc[i]);
            bw_red_add_u64(&t.node_acc2[nd], c.cnt[i]);
            bw_red_min_u64(&t.nodes[nd].open_seq, c.seq[i]);
          }
        }
        bw_red_max_s64(&hs->max_ts, c.max_ts[i]);
        atomicMax((long long*)&t.closed_upto[s], (long long)c.closed_upto[i]);
        bw_mark_dirty(t, &sink, s);  // K4 re-ranks every restored key before the next activation
      }
    }
  }
  __syncthreads();
  bw_sinks_flush(&sink, t);
}

__global__ void k_snap_set_gmax(Counters* ctr, i64 gmax) {
  if ((i64)ctr->gmax_ts < gmax) ctr->gmax_ts = (unsigned long long)gmax;
}
