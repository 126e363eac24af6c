// bw_close.cuh -- K4: watermark / close scan and compaction of closed rows.
//
// Replaces `_WindowLogic._flush_queue` -> `_handle_closed`
// (pysrc/bytewax/operators/windowing.py:1087-1108), the windower's
// `close_for` (windowing.py:645-654), the discard of an empty logic
// (windowing.py:1110-1113, src/operators.rs:796-799) and, at end of input,
// `on_eof` for every live key (src/operators.rs:862-894).
//
// Runs between fold kernels, one thread per key whose state may have
// changed shape (the dirty list), so it never races with the fold.
#pragma once
#include "bw_common.cuh"
#include "bw_fold.cuh"

#define BW_TOMB_WID (INT64_MIN + 1)

__device__ __forceinline__ u64 bw_warp_reserve(unsigned long long* ctr, u32 n) {
  // opportunistic warp aggregation: one atomic per converged group
  unsigned m = __activemask();
  int lane = threadIdx.x & 31;
  // inclusive scan of n over the active lanes
  u32 pre = 0, tot = 0;
  for (int l = 0; l < 32; ++l) {
    if (m & (1u << l)) {
      u32 v = __shfl_sync(m, n, l);
      if (l < lane) pre += v;
      tot += v;
    }
  }
  int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(ctr, (unsigned long long)tot);
  base = __shfl_sync(m, base, leader);
  return base + pre;
}

__device__ __forceinline__ void bw_emit_closed(const EmitBufs& e, Counters* c, u64 key, i64 wid, u64 acc, u64 cnt,
                                               u64 seq, u64 epoch) {
  u64 i = bw_warp_reserve(&c->n_closed, 1u);
  if (i >= e.max_closed) {
    bw_raise(c, 3u);
    return;
  }
  e.c_key[i] = key;
  e.c_wid[i] = wid;
  e.c_acc[i] = acc;
  e.c_count[i] = cnt;
  e.c_seq[i] = seq;
  e.c_epoch[i] = epoch;
}

__device__ __forceinline__ u64 bw_combine(int op, u64 a, u64 b) {
  switch (op) {
    case BW_OP_ADD_ONE:
    case BW_OP_ADD_U64: return a + b;
    case BW_OP_ADD_F64:
      return (u64)__double_as_longlong(__longlong_as_double((i64)a) + __longlong_as_double((i64)b));
    case BW_OP_MIN_S64: return (u64)(((i64)a < (i64)b) ? (i64)a : (i64)b);
    case BW_OP_MIN_U64: return a < b ? a : b;
    case BW_OP_MAX_S64: return (u64)(((i64)a > (i64)b) ? (i64)a : (i64)b);
    default: return a > b ? a : b;
  }
}

// accumulator bits -> emitted bits (undo the ordered-float encoding)
__device__ __forceinline__ u64 bw_finish_acc(const FoldParams& p, u64 acc) {
  if (p.val_dtype >= 2 && (p.op == BW_OP_MIN_U64 || p.op == BW_OP_MAX_U64)) return bw_ordered_to_f64(acc);
  return acc;
}

struct PaneRef {
  i64 q;
  u64 acc, cnt, seq;
  u32 node;  // 0 == the inline pane
  bool valid;
};

__device__ __forceinline__ PaneRef bw_pane_inline(const Table& t, u64 s) {
  PaneRef r;
  i64 tag = t.hot[s].widtag;
  r.valid = (tag != BW_EMPTY_WIDTAG);
  r.q = bw_widtag_q(tag);
  r.acc = t.hot[s].acc;
  r.cnt = t.cold[s].acc2;
  r.seq = t.cold[s].open_seq;
  r.node = 0;
  return r;
}
__device__ __forceinline__ PaneRef bw_pane_node(const Table& t, u32 n) {
  PaneRef r;
  r.valid = true;
  r.q = t.nodes[n].wid;
  r.acc = t.nodes[n].acc;
  r.cnt = t.node_acc2[n];
  r.seq = t.nodes[n].open_seq;
  r.node = n;
  return r;
}

// Close everything the key's watermark allows; keep the newest pane inline.
__device__ void bw_close_key(const Table& t, const FoldParams& p, const EmitBufs& e, u64 s, bool eof, u64 epoch) {
  HotSlot* hs = t.hot + s;
  ColdSlot* cs = t.cold + s;
  i64 tag = hs->widtag;
  if (tag == BW_EMPTY_WIDTAG) return;
  const u64 key = hs->key;
  const i64 mts = hs->max_ts;
  i64 wm;
  if (eof) {
    wm = INT64_MAX;
  } else if (!p.track_wm) {
    wm = INT64_MIN;
  } else {
    wm = bw_sub_sat(mts, p.wait_us);
    if (wm < BW_UTC_MIN_US_DEV) wm = BW_UTC_MIN_US_DEV;
  }
  const i64 a = p.panes_per_offset, b = p.panes_per_window;
  const bool ordered_seq = p.ordered != 0;
  bool inline_alive = true;

  if (a == 1 && b == 1) {
    // tumbling: window id == pane id; a closed pane is emitted and dropped
    PaneRef in = bw_pane_inline(t, s);
    if (wm >= bw_pane_release(in.q, p)) {
      bw_emit_closed(e, t.ctr, key, in.q, bw_finish_acc(p, in.acc), in.cnt,
                     ordered_seq ? (u64)(in.q + (1LL << 62)) : in.seq, epoch);
      inline_alive = false;
    }
    for (u32 n = cs->spill_head; n; n = t.nodes[n].next) {
      PaneRef r = bw_pane_node(t, n);
      if (wm >= bw_pane_release(r.q, p)) {
        bw_emit_closed(e, t.ctr, key, r.q, bw_finish_acc(p, r.acc), r.cnt,
                       ordered_seq ? (u64)(r.q + (1LL << 62)) : r.seq, epoch);
        t.nodes[n].wid = BW_TOMB_WID;
      }
    }
  } else {
    // sliding: window w = panes [w*a, w*a + b); emit each newly closable
    // window once, from its smallest live pane; then drop dead panes.
    i64 c_new;  // largest closable window id
    if (eof) {
      c_new = INT64_MAX;
    } else if (wm == INT64_MIN) {
      c_new = INT64_MIN;
    } else {
      c_new = bw_floordiv(wm - p.align_us - p.length_us, p.offset_us);
    }
    const i64 c_prev = cs->closed_upto;  // windows <= c_prev were already emitted
    if (c_new > c_prev) {
      // outer walk over panes P
      PaneRef P = bw_pane_inline(t, s);
      u32 next = cs->spill_head;
      while (true) {
        if (P.valid) {
          i64 w_lo = bw_floordiv(P.q - b + a, a);  // ceil((q - b + 1)/a)
          i64 w_hi = bw_floordiv(P.q, a);
          if (w_lo <= c_prev) w_lo = c_prev + 1;
          if (w_hi > c_new) w_hi = c_new;
          for (i64 w = w_lo; w <= w_hi; ++w) {
            const i64 q0 = w * a, q1 = w * a + b;  // pane range [q0, q1)
            // is P the smallest live pane of w?  combine all panes of w on the way
            bool smallest = true;
            u64 acc = p.acc_identity, cnt = 0, seq = ~0ULL;
            PaneRef R = bw_pane_inline(t, s);
            u32 rn = cs->spill_head;
            while (true) {
              if (R.valid && R.q >= q0 && R.q < q1) {
                if (R.q < P.q) {
                  smallest = false;
                  break;
                }
                acc = bw_combine(p.op, acc, R.acc);
                cnt += R.cnt;
                seq = R.seq < seq ? R.seq : seq;
              }
              if (!rn) break;
              R = bw_pane_node(t, rn);
              rn = t.nodes[rn].next;
            }
            if (smallest)
              bw_emit_closed(e, t.ctr, key, w, bw_finish_acc(p, acc), cnt,
                             ordered_seq ? (u64)(w + (1LL << 62)) : seq, epoch);
          }
        }
        if (!next) break;
        P = bw_pane_node(t, next);
        next = t.nodes[next].next;
      }
      cs->closed_upto = c_new;
    }
    // drop panes whose last window is closed
    PaneRef in = bw_pane_inline(t, s);
    if (wm >= bw_pane_release(in.q, p)) inline_alive = false;
    for (u32 n = cs->spill_head; n; n = t.nodes[n].next)
      if (wm >= bw_pane_release(t.nodes[n].wid, p)) t.nodes[n].wid = BW_TOMB_WID;
  }

  // rebuild: unlink dead nodes, find newest / oldest survivors
  u32 best = 0;
  i64 best_q = INT64_MIN, min_q = INT64_MAX;
  u32 prev = 0;
  for (u32 n = cs->spill_head; n;) {
    u32 nx = t.nodes[n].next;
    if (t.nodes[n].wid == BW_TOMB_WID) {
      if (prev) t.nodes[prev].next = nx; else cs->spill_head = nx;
      int top = atomicAdd(&t.ctr->free_top, 1);
      t.free_stack[top] = n;
    } else {
      i64 q = t.nodes[n].wid;
      if (q > best_q) { best_q = q; best = n; }
      if (q < min_q) min_q = q;
      prev = n;
    }
    n = nx;
  }
  i64 in_q = bw_widtag_q(tag);
  if (!inline_alive && !best) {
    // no panes left: the reference discards the whole logic, watermark included
    hs->max_ts = INT64_MIN;
    hs->widtag = BW_EMPTY_WIDTAG;
    hs->acc = p.acc_identity;
    cs->open_seq = ~0ULL;
    cs->acc2 = 0;
    cs->closed_upto = INT64_MIN;
    return;
  }
  if (!inline_alive || best_q > in_q) {
    // move the newest node into the inline position (swap when inline is alive)
    PaneNode nd = t.nodes[best];
    u64 nd_cnt = t.node_acc2[best];
    if (inline_alive) {
      t.nodes[best].wid = in_q;
      t.nodes[best].acc = hs->acc;
      t.nodes[best].open_seq = cs->open_seq;
      t.nodes[best].born = 0xFFFFFFFFu;  // never "fresh" again
      t.node_acc2[best] = cs->acc2;
      if (in_q < min_q) min_q = in_q;
    } else {
      // unlink `best`
      u32 pv = 0;
      for (u32 n = cs->spill_head; n; n = t.nodes[n].next) {
        if (n == best) {
          if (pv) t.nodes[pv].next = t.nodes[n].next; else cs->spill_head = t.nodes[n].next;
          break;
        }
        pv = n;
      }
      int top = atomicAdd(&t.ctr->free_top, 1);
      t.free_stack[top] = best;
      // min over the remaining nodes
      min_q = INT64_MAX;
      for (u32 n = cs->spill_head; n; n = t.nodes[n].next)
        if (t.nodes[n].wid < min_q) min_q = t.nodes[n].wid;
    }
    hs->acc = nd.acc;
    cs->open_seq = nd.open_seq;
    cs->acc2 = nd_cnt;
    in_q = nd.wid;
  }
  if (min_q > in_q) min_q = in_q;
  u64 delta = (u64)(in_q - min_q);
  // The born tag only matters inside the batch that created the pane (arrival-order
  // minimum of open_seq); K4 runs after that batch, so the survivor is marked stale.
  hs->widtag = bw_pack_widtag(in_q, delta > 255 ? 255u : (u32)delta, BW_TAG_STALE);
}

__global__ void k_close_dirty(Table t, FoldParams p, EmitBufs e, u64 epoch) {
  const u32 n = t.ctr->dirty_count;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    bw_close_key(t, p, e, t.dirty[i], false, epoch);
}
// EOF: every slot (including the BW_EMPTY_KEY alias slot at index capacity)
__global__ void k_close_all(Table t, FoldParams p, EmitBufs e, u64 epoch) {
  const u64 n = t.mask + 2;
  for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (u64)gridDim.x * blockDim.x) {
    if (s <= t.mask && t.hot[s].key == BW_EMPTY_KEY) continue;
    bw_close_key(t, p, e, s, true, epoch);
  }
}
__global__ void k_reset_dirty(Table t) { t.ctr->dirty_count = 0; }
