// bw_fold.cuh -- K3: window-assign + open-addressed hash-table fold.
//
// Replaces the per-key `on_batch` loop of src/operators.rs:755-806 running
// `_WindowLogic.on_batch` (pysrc/bytewax/operators/windowing.py:1115-1133) for
// numeric folds.  One thread per event; one 32-byte sector read and one or two
// fire-and-forget L2 reductions per event in the steady state.
#pragma once
#include "bw_common.cuh"


// Block-local staging of the rare global appends (dirty-key list, new-key
// count): one global atomic per block per tile instead of one per event, so
// the single list cursor in L2 never serialises the fold.
#define BW_SINK_CAP 1024
struct BlockSinks {
  u32 n_dirty;
  u32 n_new_keys;
  u32 dirty[BW_SINK_CAP];
};
__device__ __forceinline__ void bw_sinks_init(BlockSinks* sk) {
  if (threadIdx.x == 0) {
    sk->n_dirty = 0;
    sk->n_new_keys = 0;
  }
}
// call by all threads of the block, between __syncthreads()
__device__ __forceinline__ void bw_sinks_flush(BlockSinks* sk, const Table& t) {
  __shared__ u32 base;
  u32 n = sk->n_dirty < BW_SINK_CAP ? sk->n_dirty : BW_SINK_CAP;
  if (threadIdx.x == 0) {
    base = n ? atomicAdd(&t.ctr->dirty_count, n) : 0u;
    if (sk->n_new_keys) atomicAdd(&t.ctr->live_keys, (unsigned long long)sk->n_new_keys);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < n; i += blockDim.x) t.dirty[base + i] = sk->dirty[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    sk->n_dirty = 0;
    sk->n_new_keys = 0;
  }
  __syncthreads();
}

__device__ __forceinline__ u32 bw_ld_u32_coherent(const u32* p) {
  u32 v;
  asm volatile("ld.global.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ i64 bw_ld_i64_coherent(const i64* p) {
  i64 v;
  asm volatile("ld.global.relaxed.gpu.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Find (or create) the slot of `key`.  Returns the slot index, or ~0 on a full table.
__device__ __forceinline__ u64 bw_find_slot(const Table& t, BlockSinks* sk, u64 key, u64 h, i64& max_ts,
                                            i64& widtag) {
  u64 s = (key == BW_EMPTY_KEY) ? (t.mask + 1) : (h & t.mask);
  for (u64 probe = 0; probe <= t.mask; ++probe) {
    HotSlot* hs = t.hot + s;
    u64 k, acc_unused;
    bw_ld_slot(hs, k, max_ts, widtag, acc_unused);
    if (k == key) return s;
    if (k == BW_EMPTY_KEY) {
      u64 old = atomicCAS((unsigned long long*)&hs->key, (unsigned long long)BW_EMPTY_KEY, (unsigned long long)key);
      if (old == BW_EMPTY_KEY) {
        // a free slot is always in the reset state; nobody else touches it before the key is set
        atomicAdd(&sk->n_new_keys, 1u);
        return s;
      }
      if (old == key) {
        bw_ld_slot(hs, k, max_ts, widtag, acc_unused);
        return s;
      }
    }
    s = (s + 1) & t.mask;
  }
  return ~0ULL;
}

__device__ __forceinline__ void bw_mark_dirty(const Table& t, BlockSinks* sk, u64 s) {
  unsigned long long old =
      atomicOr((unsigned long long*)&t.hot[s].widtag, (unsigned long long)BW_TAG_DIRTY);
  if (!(old & (unsigned long long)BW_TAG_DIRTY)) {
    u32 i = atomicAdd(&sk->n_dirty, 1u);
    if (i < BW_SINK_CAP) {
      sk->dirty[i] = (u32)s;
    } else {  // staging full: append directly
      u32 g = atomicAdd(&t.ctr->dirty_count, 1u);
      t.dirty[g] = (u32)s;
    }
  }
}

// Find or create the extra pane node (key slot s, pane q).  Returns node index or 0 on failure.
__device__ __noinline__ u32 bw_spill_node(const Table& t, const FoldParams& p, u64 s, i64 q, u32 batch_no,
                                         bool& created) {
  ColdSlot* cs = t.cold + s;
  created = false;
  u32 n = bw_ld_u32_coherent(&cs->spill_head);
  while (n) {
    if (bw_ld_i64_coherent(&t.nodes[n].wid) == q) return n;
    n = bw_ld_u32_coherent(&t.nodes[n].next);
  }
  u32 result = 0;
  bool done = false;
  while (!done) {
    if (atomicCAS(&cs->lock, 0u, 1u) == 0u) {
      __threadfence();
      u32 head = bw_ld_u32_coherent(&cs->spill_head);
      n = head;
      while (n) {
        if (bw_ld_i64_coherent(&t.nodes[n].wid) == q) break;
        n = bw_ld_u32_coherent(&t.nodes[n].next);
      }
      if (!n) {
        int top = atomicSub(&t.ctr->free_top, 1);
        if (top > 0) {
          n = t.free_stack[top - 1];
        } else {
          atomicAdd(&t.ctr->free_top, 1);
          n = atomicAdd(&t.ctr->pool_next, 1u);
          if (n >= t.pool_cap) {
            n = 0;
            bw_raise(t.ctr, 3u /*BW_ERR_CAPACITY*/);
          }
        }
        if (n) {
          PaneNode nd;
          nd.wid = q;
          nd.acc = p.acc_identity;
          nd.open_seq = ~0ULL;
          nd.next = head;
          nd.born = batch_no;
          t.nodes[n] = nd;
          t.node_acc2[n] = 0;
          __threadfence();
          atomicExch(&cs->spill_head, n);
          created = true;
        }
      }
      result = n;
      __threadfence();
      atomicExch(&cs->lock, 0u);
      done = true;
    }
  }
  return result;
}

// Fold one non-late event.  `seq` = (batch_no << 32) | arrival index.
__device__ __forceinline__ void bw_fold_event(const Table& t, const FoldParams& p, BlockSinks* sk, u64 key, i64 ts,
                                              u64 operand, u64 seq, u32 batch_no) {
  i64 q = bw_pane_of(ts, p);
  if (q <= -BW_WID_LIMIT || q >= BW_WID_LIMIT) {
    bw_raise(t.ctr, 6u /*BW_ERR_RANGE*/);
    return;
  }
  i64 mts, tag;
  u64 s = bw_find_slot(t, sk, key, bw_mix64(key), mts, tag);
  if (s == ~0ULL) {
    bw_raise(t.ctr, 3u);
    return;
  }
  HotSlot* hs = t.hot + s;
  if (tag == BW_EMPTY_WIDTAG) {
    i64 mine = bw_pack_widtag(q, 0, batch_no & 63u);
    i64 old = (i64)atomicCAS((unsigned long long*)&hs->widtag, (unsigned long long)BW_EMPTY_WIDTAG,
                             (unsigned long long)mine);
    tag = (old == BW_EMPTY_WIDTAG) ? mine : old;
  }
  bool created = false;
  if (bw_widtag_q(tag) == q) {
    bw_apply(p.op, &hs->acc, operand);
    if (p.need_count) bw_red_add_u64(&t.cold[s].acc2, 1ULL);
    if (((u32)tag & 0x7Fu) == (batch_no & 63u)) bw_red_min_u64(&t.cold[s].open_seq, seq);
  } else {
    u32 n = bw_spill_node(t, p, s, q, batch_no, created);
    if (!n) return;
    bw_apply(p.op, &t.nodes[n].acc, operand);
    if (p.need_count) bw_red_add_u64(&t.node_acc2[n], 1ULL);
    if (t.nodes[n].born == batch_no) bw_red_min_u64(&t.nodes[n].open_seq, seq);
  }
  if (p.track_wm) {
    if (ts > mts) bw_red_max_s64(&hs->max_ts, ts);
    if (!(tag & BW_TAG_DIRTY)) {
      bool mark = created;
      if (!mark) {
        u32 delta = bw_widtag_delta(tag);
        mark = (delta == 255u) ||
               (bw_sub_sat(ts, p.wait_us) >= bw_pane_release(bw_widtag_q(tag) - (i64)delta, p));
      }
      if (mark) bw_mark_dirty(t, sk, s);
    }
  } else if (created) {
    // keep the newest pane inline for the next batch
    if (!(tag & BW_TAG_DIRTY)) bw_mark_dirty(t, sk, s);
  }
}

// value bits -> accumulator operand, and the event time
__device__ __forceinline__ void bw_load_event(const BatchView& bv, int seg, u64 i, const FoldParams& p, u64& key,
                                              i64& ts, u64& operand, u64& raw) {
  key = bw_ld_stream_u64(bv.keys[seg] + i);
  raw = 0;
  if (bv.vals[seg]) {
    if (p.val_dtype == 2 /*F32*/) {
      raw = bw_ld_stream_u32((const u32*)bv.vals[seg] + i);
    } else {
      raw = bw_ld_stream_u64((const u64*)bv.vals[seg] + i);
    }
  }
  if (p.ts_from_value) {
    ts = p.align_us + (i64)raw;
  } else {
    ts = (i64)bw_ld_stream_u64((const u64*)bv.ts[seg] + i);
  }
  // operand in accumulator representation
  if (p.val_dtype >= 2) {
    double d = (p.val_dtype == 2) ? (double)__uint_as_float((u32)raw) : __longlong_as_double((i64)raw);
    u64 b = (u64)__double_as_longlong(d);
    operand = (p.op == BW_OP_ADD_F64) ? b : bw_f64_to_ordered(b);
  } else {
    operand = raw;
    if (p.op == BW_OP_ADD_F64) {  // MEAN over integers accumulates in f64
      double d = (p.val_dtype == 1) ? (double)(i64)raw : (double)raw;
      operand = (u64)__double_as_longlong(d);
    }
  }
}

// Map a global arrival index to (segment, offset).
__device__ __forceinline__ bool bw_locate(const BatchView& bv, const u64* seg_start, u64 g, int& seg, u64& off) {
  seg = 0;
#pragma unroll
  for (int j = 1; j < BW_MAX_WORLD; ++j)
    if (j < bv.nseg && g >= seg_start[j]) seg = j;
  off = g - seg_start[seg];
  return true;
}

#define BW_FOLD_THREADS 256
#define BW_FOLD_UNROLL 4

// Fast path: every event of the batch is provably on time (prepass verdict), or
// the clock never advances on data (wait == forever).
__global__ void __launch_bounds__(BW_FOLD_THREADS)
k_fold(BatchView bv, Table t, FoldParams p, u32 batch_no, int check_clean) {
  if (check_clean && t.ctr->batch_clean == 0u) return;
  __shared__ u64 seg_start[BW_MAX_WORLD + 1];
  __shared__ BlockSinks sinks;
  bw_sinks_init(&sinks);
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
  const u64 tile = (u64)BW_FOLD_THREADS * BW_FOLD_UNROLL;
  for (u64 base = (u64)blockIdx.x * tile; base < total; base += (u64)gridDim.x * tile) {
    u64 key[BW_FOLD_UNROLL], operand[BW_FOLD_UNROLL], raw;
    i64 ts[BW_FOLD_UNROLL];
    bool ok[BW_FOLD_UNROLL];
#pragma unroll
    for (int u = 0; u < BW_FOLD_UNROLL; ++u) {
      u64 g = base + (u64)u * BW_FOLD_THREADS + threadIdx.x;
      ok[u] = g < total;
      if (ok[u]) {
        int seg = 0;
        u64 off = g;
        if (bv.nseg > 1) bw_locate(bv, seg_start, g, seg, off);
        bw_load_event(bv, seg, off, p, key[u], ts[u], operand[u], raw);
      }
    }
#pragma unroll
    for (int u = 0; u < BW_FOLD_UNROLL; ++u) {
      if (ok[u]) {
        u64 g = base + (u64)u * BW_FOLD_THREADS + threadIdx.x;
        bw_fold_event(t, p, &sinks, key[u], ts[u], operand[u], ((u64)batch_no << 32) | g, batch_no);
      }
    }
    __syncthreads();
    if (sinks.n_dirty > BW_SINK_CAP / 2) bw_sinks_flush(&sinks, t);  // uniform: read after the barrier
  }
  __syncthreads();
  bw_sinks_flush(&sinks, t);
}
