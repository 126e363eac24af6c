// bw_stream.cuh -- K3, streaming form: fused lateness verdict + bucket scatter, then a
// shared-memory-resident segment fold that touches the table once per (key, activation).
//
// Same contract as k_fold (bw_fold.cuh): the per-key `on_batch` loop of
// src/operators.rs:755-806 over `_WindowLogic.on_batch` (windowing.py:1115-1133),
// followed for tumbling windows by `_flush_queue` -> `_handle_closed` / discard
// (windowing.py:1087-1113) for the keys the activation touched.
//
// The direct kernel pays one random 32-byte L2 read and two L2 reductions per EVENT
// (profiles/r01_fold_ncu_final.md: 45 G event/s ceiling).  Here every byte moves in streams:
//
//   k_scatter   one pass over the input columns (128-bit loads): lateness verdict of the
//               activation (what bw_prepass.cuh computed in its own pass) + multisplit of
//               the rows into one region per table SEGMENT (2048 consecutive home slots):
//               rank inside (tile, bucket) with a shared-memory histogram, one global
//               reservation per (tile, bucket), 16-byte records {key, ts - ts0, arrival index}.
//   k_verdict   one block: chains the per-tile (min, max, bad) triples -> clean?, event-time span.
//   k_segfold   one block per segment.  Probing wraps inside a segment (Table::seg_mask), so the
//               block owns every key that hashes into it: it keeps the segment's KEYS in shared
//               memory, claims new keys there, and accumulates per-(slot, local pane) deltas, the
//               newest timestamp and the first-open index with native 32-bit shared-memory atomics
//               (profiles/r02_ubench_smem.txt: a spread ATOMS costs about what an LDS costs).
//               One thread per touched slot then merges the deltas into the table (plain
//               loads / stores: nobody else touches the segment), closes what the key's new
//               watermark allows and re-ranks its panes -- the work of K4 for those keys.
//   k_spill     the few rows / partials that do not fit that scheme (alias key, full region,
//               a third live pane) through the general path of the direct kernel.
#pragma once
#include <type_traits>

#include "bw_close.cuh"
#include "bw_common.cuh"
#include "bw_fold.cuh"

#define BW_SEG_SHIFT 11
#define BW_SEG_SLOTS (1u << BW_SEG_SHIFT)  // slots per segment == per bucket
#define BW_STREAM_MAX_NB 8192              // buckets per table (beyond: the direct kernel)

#define BW_SC_THREADS 256
#define BW_SC_WARPS (BW_SC_THREADS / 32)
#define BW_SF_THREADS 512
#define BW_SF_UNROLL 4

struct __align__(8) SpillRec {  // a row, or a pre-combined partial, for the general path
  u64 key;
  i64 ts;      // newest event time it stands for (its pane is the pane of ts)
  u64 acc;     // operand in accumulator representation (ignored for counts of weight 1 rows: acc == weight)
  u64 seq;     // batch << 32 | arrival index of its first row
  u64 weight;  // number of values it stands for
};

struct StreamVerdict {
  u32 clean;   // no row of the activation can be late (prepass rule, bw_prepass.cuh)
  u32 flags;   // BW_SV_*
  i64 tmin, tmax;  // event-time span of the activation
  i64 ts0;     // base of the records' 32-bit relative timestamps (event time of row 0)
  u32 n_spill; // rows in this activation's spill list
  u32 pad;
};
#define BW_SV_RANGE 1u   // a timestamp is further than 2^31 us from row 0: the activation takes the direct kernel
#define BW_SV_LOST 2u    // the spill list overflowed during the scatter: rows were dropped from the buckets

struct StreamSide {    // one of two alternating sets (scatter of b+1 is queued before the fold of b)
  uint4* rec;          // [nb * region_cap]
  void* val;           // [nb * region_cap] values (only folds that need them)
  u32* cursor;         // [nb] rows per bucket; zero between activations
  SpillRec* spill;
  StreamVerdict* sv;   // device
};
struct StreamBufs {
  StreamSide side[2];
  u32 nb, region_cap, spill_cap;
  int val_bytes;       // value bytes stored beside the records: 0 (counts), 4 or 8
  i64 *tile_min, *tile_max;
  u32* tile_bad;
  u32 tiles_cap;
};

// ---------------------------------------------------------------------------
// lateness triples: (min, max, some row later than an earlier row by more than `wait`)
// ---------------------------------------------------------------------------
struct Trip {
  i64 mn, mx;
  u32 bad;
};
__device__ __forceinline__ Trip bw_trip_id() { return Trip{INT64_MAX, INT64_MIN, 0u}; }
__device__ __forceinline__ Trip bw_trip_of(i64 ts, bool valid) { return valid ? Trip{ts, ts, 0u} : bw_trip_id(); }
// a precedes b in arrival order
__device__ __forceinline__ Trip bw_trip_cat(const Trip& a, const Trip& b, i64 wait) {
  Trip r;
  r.mn = a.mn < b.mn ? a.mn : b.mn;
  r.mx = a.mx > b.mx ? a.mx : b.mx;
  r.bad = a.bad | b.bad | ((b.mn < bw_sub_sat(a.mx, wait)) ? 1u : 0u);
  return r;
}
__device__ __forceinline__ Trip bw_trip_shfl_up(const Trip& t, int d) {
  Trip r;
  r.mn = __shfl_up_sync(0xffffffffu, t.mn, d);
  r.mx = __shfl_up_sync(0xffffffffu, t.mx, d);
  r.bad = __shfl_up_sync(0xffffffffu, t.bad, d);
  return r;
}
// concatenation over the warp in lane order; the result is valid in lane 31
__device__ __forceinline__ Trip bw_trip_warp(Trip t, i64 wait) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Trip y = bw_trip_shfl_up(t, d);
    if (lane >= d) t = bw_trip_cat(y, t, wait);
  }
  return t;
}

// 128-bit streaming loads of two consecutive 8-byte column entries (read once: no L1, evict-first in L2)
__device__ __forceinline__ void bw_ld_stream_2u64(const u64* p, u64& a, u64& b) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
               : "=l"(a), "=l"(b)
               : "l"(p), "l"(bw_evict_first_policy()));
}
__device__ __forceinline__ void bw_ld_stream_2u32(const u32* p, u32& a, u32& b) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(a), "=r"(b)
               : "l"(p), "l"(bw_evict_first_policy()));
}
__device__ __forceinline__ uint4 bw_ld_stream_rec(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(bw_evict_first_policy()));
  return v;
}

struct ScatterArgs {
  const u64* keys;
  const void* vals;  // may be NULL (counts with a ts column)
  const i64* ts;     // NULL unless the fold has a ts column
  u64 n;
  StreamSide out;
  u32 nb, region_cap, spill_cap;
  i64 *tile_min, *tile_max;
  u32* tile_bad;
  u64 cap;           // table capacity (slots)
  u32 batch_no;
};

__device__ __forceinline__ void bw_spill_push(SpillRec* list, u32* n, u32 cap, u32* flags, u32 lost_flag, Counters* ctr, u64 key,
                                              i64 ts, u64 acc, u64 seq, u64 weight) {
  const u32 i = atomicAdd(n, 1u);
  if (i >= cap) {
    if (flags) atomicOr(flags, lost_flag);
    else bw_raise(ctr, 3u);
    return;
  }
  SpillRec r;
  r.key = key;
  r.ts = ts;
  r.acc = acc;
  r.seq = seq;
  r.weight = weight;
  list[i] = r;
}

// value bits of one row as stored in the input column (f32 widened to its 32 raw bits in the low word)
template <int VB_IN>
__device__ __forceinline__ void bw_ld_val_pair(const void* vals, u64 row, bool both, u64& a, u64& b) {
  a = 0;
  b = 0;
  if (VB_IN == 8) {
    const u64* p = (const u64*)vals + row;
    if (both) bw_ld_stream_2u64(p, a, b);
    else a = bw_ld_stream_u64(p);
  } else if (VB_IN == 4) {
    const u32* p = (const u32*)vals + row;
    u32 x = 0, y = 0;
    if (both) bw_ld_stream_2u32(p, x, y);
    else x = bw_ld_stream_u32(p);
    a = x;
    b = y;
  }
}

// TSM: 0 = ts column, 1 = ts from the (integer) value, 2 = none (the *_final folds).
// VB_IN: bytes per entry of the value column read here (0: not read).  VB_OUT: value bytes stored beside
// the records (0 for counts).  RPT rows per thread per tile (even).
template <int TSM, int VB_IN, int VB_OUT, int RPT>
__global__ void __launch_bounds__(BW_SC_THREADS) k_scatter(ScatterArgs A, FoldParams p) {
  extern __shared__ __align__(16) u32 sc_sm[];
  u32* cnt = sc_sm;
  u32* gbase = sc_sm + A.nb;
  constexpr int NPAIR = RPT / 2;
  constexpr int NCHUNK = NPAIR * BW_SC_WARPS;  // 64-row chunks per tile, in arrival order
  __shared__ i64 c_min[NCHUNK], c_max[NCHUNK];
  __shared__ u32 c_bad[NCHUNK];
  constexpr u32 T = BW_SC_THREADS * RPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u64 ntiles = (A.n + T - 1) / T;
  // base of the relative timestamps: event time of row 0
  i64 ts0 = p.align_us;
  if (TSM == 0) ts0 = A.ts[0];
  else if (TSM == 1) ts0 = p.align_us + (i64)((const u64*)A.vals)[0];
  u32* flags = &A.out.sv->flags;
  u32* n_spill = &A.out.sv->n_spill;
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (u32 d = threadIdx.x; d < A.nb; d += BW_SC_THREADS) cnt[d] = 0u;
    __syncthreads();
    u64 key[RPT], val[VB_OUT ? RPT : 1];
    u32 meta[RPT];  // bucket << 16 | rank inside (tile, bucket); 0xFFFFFFFF: not scattered
    int rel[RPT];
    const u64 tbase = tile * (u64)T;
#pragma unroll
    for (int j = 0; j < NPAIR; ++j) {
      const u64 row = tbase + 2ull * ((u64)j * BW_SC_THREADS + threadIdx.x);
      const bool va = row < A.n, vb = row + 1 < A.n;
      u64 ka = 0, kb = 0, xa = 0, xb = 0;
      i64 ta = p.align_us, tb = p.align_us;
      if (va) {
        if (vb) bw_ld_stream_2u64(A.keys + row, ka, kb);
        else ka = bw_ld_stream_u64(A.keys + row);
        if (VB_IN) bw_ld_val_pair<VB_IN>(A.vals, row, vb, xa, xb);
        if (TSM == 0) {
          u64 a, b = 0;
          if (vb) bw_ld_stream_2u64((const u64*)A.ts + row, a, b);
          else a = bw_ld_stream_u64((const u64*)A.ts + row);
          ta = (i64)a;
          tb = (i64)b;
        } else if (TSM == 1) {
          ta = p.align_us + (i64)xa;
          tb = p.align_us + (i64)xb;
        }
      }
      key[2 * j] = ka;
      key[2 * j + 1] = kb;
      if (VB_OUT) {
        val[VB_OUT ? 2 * j : 0] = xa;
        val[VB_OUT ? 2 * j + 1 : 0] = xb;
      }
      // lateness triple of this warp's 64 consecutive rows
      if (TSM != 2) {
        const i64 nxt = __shfl_down_sync(0xffffffffu, ta, 1);
        const bool ordered = va && vb && ta <= tb && (lane == 31 || tb <= nxt);
        Trip ct;
        if (__all_sync(0xffffffffu, ordered)) {
          ct.mn = __shfl_sync(0xffffffffu, ta, 0);
          ct.mx = __shfl_sync(0xffffffffu, tb, 31);
          ct.bad = 0u;
        } else {
          ct = bw_trip_warp(bw_trip_cat(bw_trip_of(ta, va), bw_trip_of(tb, vb), p.wait_us), p.wait_us);
          ct.mn = __shfl_sync(0xffffffffu, ct.mn, 31);
          ct.mx = __shfl_sync(0xffffffffu, ct.mx, 31);
          ct.bad = __shfl_sync(0xffffffffu, ct.bad, 31);
        }
        if (lane == 0) {
          const int c = j * BW_SC_WARPS + warp;
          c_min[c] = ct.mn;
          c_max[c] = ct.mx;
          c_bad[c] = ct.bad;
        }
      }
      // bucket + rank
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool v = h ? vb : va;
        const u64 k = h ? kb : ka;
        const i64 t = h ? tb : ta;
        u32 m = 0xFFFFFFFFu;
        int r = 0;
        if (v) {
          const i64 d = t - ts0;
          r = (int)d;
          if (d != (i64)r || r == INT32_MIN || r == INT32_MAX) atomicOr(flags, BW_SV_RANGE);
          if (k == BW_EMPTY_KEY) {
            // the alias slot lives outside every segment: general path
            const u64 x = h ? xb : xa;
            u64 operand;
            bw_operand(p, x, operand);
            bw_spill_push(A.out.spill, n_spill, A.spill_cap, flags, BW_SV_LOST, nullptr, k, t, (p.op == BW_OP_ADD_ONE) ? 1ULL : operand,
                          ((u64)A.batch_no << 32) | (row + h), 1ULL);
          } else {
            const u32 b = (u32)(bw_slot_of_hash(bw_mix64(k), A.cap) >> BW_SEG_SHIFT);
            m = (b << 16) | atomicAdd(&cnt[b], 1u);
          }
        }
        meta[2 * j + h] = m;
        rel[2 * j + h] = r;
      }
    }
    __syncthreads();
    // one reservation per (tile, bucket)
    for (u32 d = threadIdx.x; d < A.nb; d += BW_SC_THREADS) {
      const u32 c = cnt[d];
      if (c) gbase[d] = atomicAdd(&A.out.cursor[d], c);
    }
    if (TSM != 2 && warp == 0) {
      // triple of the tile: chunks in arrival order, NCHUNK / 32 per lane, then across the warp
      constexpr int E = (NCHUNK + 31) / 32;
      Trip tt = bw_trip_id();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int c = lane * E + e;
        if (c < NCHUNK) tt = bw_trip_cat(tt, Trip{c_min[c], c_max[c], c_bad[c]}, p.wait_us);
      }
      tt = bw_trip_warp(tt, p.wait_us);
      if (lane == 31) {
        A.tile_min[tile] = tt.mn;
        A.tile_max[tile] = tt.mx;
        A.tile_bad[tile] = tt.bad;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const u32 m = meta[i];
      if (m == 0xFFFFFFFFu) continue;
      const u32 b = m >> 16;
      const u32 pos = gbase[b] + (m & 0xFFFFu);
      const u64 row = tbase + 2ull * ((u64)(i >> 1) * BW_SC_THREADS + threadIdx.x) + (i & 1);
      if (pos < A.region_cap) {
        const size_t at = (size_t)b * A.region_cap + pos;
        A.out.rec[at] = make_uint4((u32)key[i], (u32)(key[i] >> 32), (u32)rel[i], (u32)row);
        if (VB_OUT == 8) ((u64*)A.out.val)[at] = val[VB_OUT ? i : 0];
        else if (VB_OUT == 4) ((u32*)A.out.val)[at] = (u32)val[VB_OUT ? i : 0];
      } else {
        // the bucket's region is full (skewed keys): general path for this row
        u64 operand = 1ULL;
        if (VB_OUT) bw_operand(p, val[VB_OUT ? i : 0], operand);
        bw_spill_push(A.out.spill, n_spill, A.spill_cap, flags, BW_SV_LOST, nullptr, key[i], ts0 + (i64)rel[i],
                      (p.op == BW_OP_ADD_ONE) ? 1ULL : operand, ((u64)A.batch_no << 32) | row, 1ULL);
      }
    }
    // cnt / gbase are rewritten two barriers from here: no barrier needed at the end of the tile
  }
}

// One block: chain the tiles (and the earlier activations through gmax_ts) as k_prepass_scan does
// for its ranges; publish the verdict, the span and the base timestamp.
__global__ void __launch_bounds__(1024)
k_verdict(const i64* tile_min, const i64* tile_max, const u32* tile_bad, u32 ntiles, FoldParams p, Counters* ctr, StreamVerdict* sv,
          const i64* ts_col, const u64* val_col) {
  __shared__ i64 s_mn[32], s_mx[32];
  __shared__ u32 s_bad[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 per = (ntiles + blockDim.x - 1) / blockDim.x;
  const u32 lo = threadIdx.x * per, hi = (lo + per < ntiles) ? lo + per : ntiles;
  Trip tt = bw_trip_id();
  for (u32 r = lo; r < hi; ++r) tt = bw_trip_cat(tt, Trip{tile_min[r], tile_max[r], tile_bad[r]}, p.wait_us);
  tt = bw_trip_warp(tt, p.wait_us);  // ordered concatenation: lane 31 holds the warp's
  if (lane == 31) {
    s_mn[warp] = tt.mn;
    s_mx[warp] = tt.mx;
    s_bad[warp] = tt.bad;
  }
  __syncthreads();
  if (warp == 0) {
    Trip act = bw_trip_warp(Trip{s_mn[lane], s_mx[lane], s_bad[lane]}, p.wait_us);
    if (lane == 31) {
      const i64 gprev = (i64)ctr->gmax_ts;
      // everything ingested before this activation: only its maximum matters
      const Trip all = bw_trip_cat(Trip{INT64_MAX, gprev, 0u}, act, p.wait_us);
      ctr->gmax_ts = (unsigned long long)all.mx;
      const u32 clean = (!p.track_wm || !all.bad) ? 1u : 0u;
      ctr->batch_clean = clean;
      sv->clean = clean;
      sv->tmin = act.mn;
      sv->tmax = act.mx;
      sv->ts0 = ts_col ? ts_col[0] : p.align_us + (i64)val_col[0];
    }
  }
}
// the *_final folds have no event time: everything is window 0 at align_to
__global__ void k_verdict_none(FoldParams p, Counters* ctr, StreamVerdict* sv) {
  ctr->batch_clean = 1u;
  sv->clean = 1u;
  sv->tmin = p.align_us;
  sv->tmax = p.align_us;
  sv->ts0 = p.align_us;
}

// ---------------------------------------------------------------------------
// segment fold
// ---------------------------------------------------------------------------
struct SegArgs {
  StreamSide in;
  u32 nb, region_cap, spill_cap;
  int val_bytes;
  i64 ts0;      // base of the records' relative timestamps
  i64 q_lo;     // pane of the activation's earliest timestamp
  u32 npass;    // the activation spans panes [q_lo, q_lo + 2 * npass): folded two panes at a time, closing in
                // between (what the direct path's sub-ranges do; exact because the activation is clean)
  u32 batch_no;
  u64 epoch;
};

// shared-memory accumulators of one segment: two local panes per slot
template <typename DT>
struct SegAcc {
  u64* key;    // [S] key of the slot (BW_EMPTY_KEY: free)
  DT* d[2];    // [S] delta of local pane 0 / 1
  int* mts;    // [S] newest relative timestamp, INT32_MIN: untouched
  u32* sq[2];  // [S] first arrival index per local pane (SEQ)
  u32* cn[2];  // [S] value counts (CNT)
  u32* tm;     // [S / 16] touched bits (2 per slot) for ops whose delta can equal the identity
};

template <int OP>
struct SegOp {
  static constexpr bool narrow = (OP == BW_OP_ADD_ONE);  // < 2^32 rows per activation: 32-bit deltas
  typedef typename std::conditional<narrow, u32, u64>::type DT;
  __device__ __forceinline__ static void apply(DT* a, u64 operand) {
    if (OP == BW_OP_ADD_ONE) {
      atomicAdd((u32*)a, 1u);
    } else if (OP == BW_OP_ADD_U64) {
      // exact 64-bit sum from two native 32-bit atomics: each add carries its own overflow up
      const u32 lo = (u32)operand, hi = (u32)(operand >> 32);
      const u32 old = atomicAdd((u32*)a, lo);
      const u32 carry = ((u32)(old + lo) < old) ? 1u : 0u;
      if (hi | carry) atomicAdd((u32*)a + 1, hi + carry);
    } else if (OP == BW_OP_ADD_F64) {
      atomicAdd((double*)a, __longlong_as_double((i64)operand));
    } else if (OP == BW_OP_MIN_S64) {
      atomicMin((long long*)a, (long long)operand);
    } else if (OP == BW_OP_MIN_U64) {
      atomicMin((unsigned long long*)a, (unsigned long long)operand);
    } else if (OP == BW_OP_MAX_S64) {
      atomicMax((long long*)a, (long long)operand);
    } else {
      atomicMax((unsigned long long*)a, (unsigned long long)operand);
    }
  }
};

__host__ __device__ __forceinline__ size_t bw_segfold_smem(int op, bool seq, bool cnt) {
  const size_t S = BW_SEG_SLOTS;
  size_t b = S * 8 + S * 4;                          // keys, newest timestamp
  b += 2 * S * (op == BW_OP_ADD_ONE ? 4 : 8);        // deltas
  if (op != BW_OP_ADD_ONE) b += S / 16 * 4;          // touched bits
  if (seq) b += 2 * S * 4;
  if (cnt) b += 2 * S * 4;
  return b;
}

// pane record used while one thread re-ranks a key
struct MPane {
  i64 q;
  u64 acc, cnt, seq;
};

// The fold kernel proper.  C = FoldCfg<op, wm, cnt> (compile-time), SEQ: keep first-open indices
// (folds whose emission order is not simply ascending window id).
template <class C, bool SEQ>
__global__ void __launch_bounds__(BW_SF_THREADS)
k_segfold(SegArgs A, Table t, FoldParams p, EmitBufs e) {
  constexpr int OP = C::kOp;
  constexpr bool CNT = C::kCnt != 0;
  const bool WM = p.track_wm != 0;
  typedef SegOp<OP> SO;
  typedef typename SO::DT DT;
  constexpr u32 S = BW_SEG_SLOTS;
  extern __shared__ __align__(16) unsigned char sf_raw[];
  __shared__ DirtySink sink;
  __shared__ u32 sink_buf[512];
  SegAcc<DT> sa;
  {
    unsigned char* q = sf_raw;
    sa.key = (u64*)q;
    q += S * 8;
    sa.d[0] = (DT*)q;
    q += S * sizeof(DT);
    sa.d[1] = (DT*)q;
    q += S * sizeof(DT);
    sa.mts = (int*)q;
    q += S * 4;
    sa.tm = (u32*)q;
    if (!SO::narrow) q += S / 16 * 4;
    sa.sq[0] = (u32*)q;
    if (SEQ) q += S * 4;
    sa.sq[1] = (u32*)q;
    if (SEQ) q += S * 4;
    sa.cn[0] = (u32*)q;
    if (CNT) q += S * 4;
    sa.cn[1] = (u32*)q;
  }
  if (threadIdx.x == 0) {
    sink.n_dirty = 0;
    sink.n_new_keys = 0;
    sink.cap = 512;
    sink.buf = sink_buf;
  }
  const DT ident = (OP <= BW_OP_ADD_F64) ? (DT)0 : (DT)p.acc_identity;
  const bool tumbling = p.panes_per_offset == 1 && p.panes_per_window == 1;
  for (u32 b = blockIdx.x; b < A.nb; b += gridDim.x) {
    const u64 slot_base = (u64)b << BW_SEG_SHIFT;
    const u32 n = min(A.in.cursor[b], A.region_cap);
    __syncthreads();  // the previous segment's merge is done with shared memory
    if (n == 0) continue;  // uniform
    for (u32 i = threadIdx.x; i < S; i += BW_SF_THREADS) sa.key[i] = t.hot[slot_base + i].key;
    for (u32 pass = 0; pass < A.npass; ++pass) {
    // relative start of the two local panes of this pass, and of the next pass
    const i64 pb0 = p.align_us + (A.q_lo + 2 * (i64)pass) * p.pane_us - A.ts0, pb1 = pb0 + p.pane_us, pb2 = pb1 + p.pane_us;
    const i64 q_pass = A.q_lo + 2 * (i64)pass;
    __syncthreads();
    for (u32 i = threadIdx.x; i < S; i += BW_SF_THREADS) {
      sa.d[0][i] = ident;
      sa.d[1][i] = ident;
      sa.mts[i] = INT32_MIN;
      if (!SO::narrow && i < S / 16) sa.tm[i] = 0u;
      if (SEQ) {
        sa.sq[0][i] = 0xFFFFFFFFu;
        sa.sq[1][i] = 0xFFFFFFFFu;
      }
      if (CNT) {
        sa.cn[0][i] = 0u;
        sa.cn[1][i] = 0u;
      }
    }
    __syncthreads();
    // ---- events of the bucket ----
    const uint4* rec = A.in.rec + (size_t)b * A.region_cap;
    for (u32 base = 0; base < n; base += BW_SF_THREADS * BW_SF_UNROLL) {
      uint4 r[BW_SF_UNROLL];
      u64 v[BW_SF_UNROLL];
#pragma unroll
      for (int u = 0; u < BW_SF_UNROLL; ++u) {
        const u32 i = base + (u32)u * BW_SF_THREADS + threadIdx.x;
        r[u] = make_uint4(0, 0, 0, 0);
        v[u] = 0;
        if (i < n) {
          r[u] = bw_ld_stream_rec(rec + i);
          if (OP != BW_OP_ADD_ONE) {
            const size_t at = (size_t)b * A.region_cap + i;
            v[u] = (A.val_bytes == 8) ? bw_ld_stream_u64((const u64*)A.in.val + at) : (u64)bw_ld_stream_u32((const u32*)A.in.val + at);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < BW_SF_UNROLL; ++u) {
        const u32 i = base + (u32)u * BW_SF_THREADS + threadIdx.x;
        if (i >= n) continue;
        const u64 key = (u64)r[u].x | ((u64)r[u].y << 32);
        const int rel = (int)r[u].z;
        if (A.npass > 1 && ((i64)rel < pb0 || (i64)rel >= pb2)) continue;  // another pass's rows
        u32 ls = (u32)bw_slot_of_hash(bw_mix64(key), t.cap) & (S - 1);
        // find or claim the key's slot: linear probing inside the segment, in shared memory
        bool found = false;
        for (u32 probe = 0; probe < S; ++probe) {
          const u64 k = ((volatile u64*)sa.key)[ls];
          if (k == key) {
            found = true;
            break;
          }
          if (k == BW_EMPTY_KEY) {
            const u64 old = atomicCAS((unsigned long long*)&sa.key[ls], (unsigned long long)BW_EMPTY_KEY, (unsigned long long)key);
            if (old == BW_EMPTY_KEY || old == key) {
              found = true;
              break;
            }
          }
          ls = (ls + 1) & (S - 1);
        }
        if (!found) {
          bw_raise(t.ctr, 3u);  // segment full: capacity_hint too small
          continue;
        }
        const int j = ((i64)rel >= pb1) ? 1 : 0;
        u64 operand = 0;
        if (OP != BW_OP_ADD_ONE) bw_operand(p, v[u], operand);
        SO::apply(&sa.d[j][ls], operand);
        if (!SO::narrow) atomicOr(&sa.tm[ls >> 4], 1u << (2 * (ls & 15) + j));
        atomicMax(&sa.mts[ls], rel);
        if (SEQ) atomicMin(&sa.sq[j][ls], r[u].w);
        if (CNT) atomicAdd(&sa.cn[j][ls], 1u);
      }
    }
    __syncthreads();
    // ---- merge: one thread per touched slot; the block owns the segment ----
    for (u32 ls = threadIdx.x; ls < S; ls += BW_SF_THREADS) {
      const int m = sa.mts[ls];
      if (m == INT32_MIN) continue;
      const u64 s = slot_base + ls;
      const u64 key = sa.key[ls];
      HotSlot h = t.hot[s];
      if (h.key == BW_EMPTY_KEY) {
        atomicAdd(&sink.n_new_keys, 1u);
        h.key = key;
      }
      bool touched[2];
      if (SO::narrow) {
        touched[0] = sa.d[0][ls] != 0;
        touched[1] = sa.d[1][ls] != 0;
      } else {
        const u32 bits = sa.tm[ls >> 4] >> (2 * (ls & 15));
        touched[0] = bits & 1u;
        touched[1] = bits & 2u;
      }
      const i64 ts_new = A.ts0 + (i64)m;  // newest event of the key in this activation
      if (WM && ts_new > h.max_ts) h.max_ts = ts_new;
      const i64 tag_in = h.wt0;
      const bool had_p1 = (tag_in & BW_TAG_HAS_P1) != 0;
      const u64 seq_hi = (u64)A.batch_no << 32;
      bool done = false;
      if (tumbling && !(tag_in & (BW_TAG_HAS_LIST | BW_TAG_DIRTY))) {
        // Every pane of the key is at hand: close what the new watermark allows and re-rank,
        // exactly what K4 (bw_close_key_simple) would do for it after the activation.  (A key
        // already on the dirty list stays K4's: its entry there must meet the DIRTY bit again.)
        MPane P[4];
        int np = 0;
        AuxSlot ax;
        const bool need_aux = (SEQ || CNT) && tag_in != BW_EMPTY_WIDTAG;
        if (need_aux) ax = t.aux[s];
        if (tag_in != BW_EMPTY_WIDTAG) {
          P[np++] = MPane{bw_widtag_q(tag_in), h.acc0, CNT ? ax.cnt0 : 0ULL, SEQ ? ax.seq0 : 0ULL};
          if (had_p1) {
            const P1Slot ps = t.p1[s];
            P[np++] = MPane{bw_widtag_q1(tag_in), ps.acc1, CNT ? ax.cnt1 : 0ULL, ps.seq1};
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (!touched[j]) continue;
          const i64 q = q_pass + j;
          const u64 d = (u64)sa.d[j][ls];
          const u64 dc = CNT ? (u64)sa.cn[j][ls] : 0ULL;
          const u64 ds = SEQ ? (seq_hi | sa.sq[j][ls]) : 0ULL;
          int at = -1;
          for (int i = 0; i < np; ++i)
            if (P[i].q == q) at = i;
          if (at < 0) {
            P[np++] = MPane{q, bw_combine(OP, p.acc_identity, d), dc, ds};
          } else {
            P[at].acc = bw_combine(OP, P[at].acc, d);
            P[at].cnt += dc;
            if (ds < P[at].seq) P[at].seq = ds;
          }
        }
        i64 wm = INT64_MIN;
        if (WM) {
          wm = bw_sub_sat(h.max_ts, p.wait_us);
          if (wm < BW_UTC_MIN_US_DEV) wm = BW_UTC_MIN_US_DEV;
        }
        // survivors: newest and second newest
        int i0 = -1, i1 = -1, alive = 0;
        bool range_ok = true;
        for (int i = 0; i < np; ++i) {
          if (P[i].q <= -BW_WID_LIMIT || P[i].q >= BW_WID_LIMIT) range_ok = false;
          if (WM && wm >= bw_pane_release(P[i].q, p)) continue;
          ++alive;
          if (i0 < 0 || P[i].q > P[i0].q) {
            i1 = i0;
            i0 = i;
          } else if (i1 < 0 || P[i].q > P[i1].q) {
            i1 = i;
          }
        }
        if (!range_ok) {
          bw_raise(t.ctr, 6u);
          continue;
        }
        if (alive <= 1 || (alive == 2 && P[i1].q == P[i0].q - 1)) {
          for (int i = 0; i < np; ++i)
            if (WM && wm >= bw_pane_release(P[i].q, p))
              bw_emit_closed(e, t.ctr, key, P[i].q, bw_finish_acc(p, P[i].acc), P[i].cnt, p.seq_by_id ? (u64)A.batch_no : P[i].seq,
                             A.epoch);
          const bool both = alive == 2;
          if (alive == 0) {
            // no panes left: the reference discards the whole logic, watermark included (windowing.py:1110-1113)
            h.max_ts = INT64_MIN;
            h.wt0 = BW_EMPTY_WIDTAG;
            h.acc0 = p.acc_identity;
            if (SEQ || CNT) {
              AuxSlot z = need_aux ? ax : t.aux[s];
              z.seq0 = ~0ULL;
              z.cnt0 = 0;
              z.cnt1 = 0;
              t.aux[s] = z;
            }
          } else {
            h.wt0 = bw_pack_widtag(P[i0].q, both ? 1u : 0u, BW_TAG_STALE, both, both);
            h.acc0 = P[i0].acc;
            if (SEQ || CNT) {
              AuxSlot z = need_aux ? ax : t.aux[s];
              z.seq0 = P[i0].seq;
              z.cnt0 = P[i0].cnt;
              z.cnt1 = both ? P[i1].cnt : 0;
              t.aux[s] = z;
            }
          }
          if (both) {
            P1Slot ps;
            ps.acc1 = P[i1].acc;
            ps.seq1 = SEQ ? P[i1].seq : 0ULL;  // present (any value but ~0)
            t.p1[s] = ps;
          } else if (had_p1) {
            P1Slot ps;
            ps.acc1 = p.acc_identity;
            ps.seq1 = ~0ULL;
            t.p1[s] = ps;
          }
          t.hot[s] = h;
          done = true;
        }
      }
      if (done) continue;
      // General shape (sliding windows, an overflow list, or more survivors than the two direct
      // panes hold): apply the deltas where the direct kernel would have, leave the closing to K4.
      i64 tag0 = tag_in;
      bool created = false;
      AuxSlot ax = t.aux[s];
      bool aux_dirty = false;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (!touched[j]) continue;
        const i64 q = q_pass + j;
        if (q <= -BW_WID_LIMIT || q >= BW_WID_LIMIT) {
          bw_raise(t.ctr, 6u);
          continue;
        }
        const u64 d = (u64)sa.d[j][ls];
        const u64 dc = CNT ? (u64)sa.cn[j][ls] : 0ULL;
        const u64 ds = seq_hi | (SEQ ? sa.sq[j][ls] : 0u);
        if (tag0 == BW_EMPTY_WIDTAG) {
          const i64 a = p.panes_per_offset, bb = p.panes_per_window;
          const i64 d0 = (a == 1) ? (bb - 1) : (q - a * bw_floordiv(q - bb + a, a));
          tag0 = bw_pack_widtag(q, d0 > (i64)BW_TAG_DELTA_MAX ? BW_TAG_DELTA_MAX : (u32)d0, A.batch_no & 63u);
          h.acc0 = p.acc_identity;
          ax.seq0 = ~0ULL;
          created = true;
        }
        if (bw_widtag_q(tag0) == q) {
          h.acc0 = bw_combine(OP, h.acc0, d);
          ax.cnt0 += dc;
          if (((u32)tag0 & 0x7Fu) == (A.batch_no & 63u) && ds < ax.seq0) ax.seq0 = ds;
          aux_dirty = true;
        } else if (bw_widtag_q1(tag0) == q) {
          P1Slot ps = t.p1[s];
          ps.acc1 = bw_combine(OP, ps.acc1, d);
          if (!(tag0 & BW_TAG_P1_PREV) && ds < ps.seq1) ps.seq1 = ds;
          t.p1[s] = ps;
          tag0 |= BW_TAG_HAS_P1;
          ax.cnt1 += dc;
          aux_dirty = true;
        } else {
          // a further pane: the overflow list, through the general path
          const i64 ts_j = (j == 0 && ts_new >= A.ts0 + pb1) ? A.ts0 + pb1 - 1 : ts_new;
          bw_spill_push(A.in.spill, &A.in.sv->n_spill, A.spill_cap, nullptr, 0u, t.ctr, key, ts_j, d, ds, dc);
        }
      }
      // watermark / closability: what bw_after_fold decides per event, once per key
      {
        i64 rem;
        const i64 qn = bw_pane_of_r(ts_new, p, rem);
        bool mark = created;
        if (WM && !mark) {
          const u32 delta = bw_widtag_delta(tag0);
          const i64 qc = qn - p.close_back - ((rem < p.wait_rem) ? 1 : 0);
          mark = (delta == BW_TAG_DELTA_MAX) || (qc >= bw_widtag_q(tag0) - (i64)delta);
        }
        if (mark && !(tag0 & BW_TAG_DIRTY)) {
          tag0 |= BW_TAG_DIRTY;
          const u32 i = atomicAdd(&sink.n_dirty, 1u);
          if (i < sink.cap) {
            sink.buf[i] = (u32)s;
          } else {
            const u32 g = atomicAdd(&t.ctr->dirty_count, 1u);
            t.dirty[g] = (u32)s;
          }
        }
      }
      h.wt0 = tag0;
      if (aux_dirty && (SEQ || CNT || created)) t.aux[s] = ax;
      t.hot[s] = h;
    }
    }  // pass
    if (threadIdx.x == 0) A.in.cursor[b] = 0u;  // ready for the next scatter into this side
    __syncthreads();
    if (sink.n_dirty > sink.cap / 2) bw_sinks_flush(&sink, t);  // uniform: read after the barrier
  }
  __syncthreads();
  bw_sinks_flush(&sink, t);
}

// Rows / partials that the streaming scheme set aside, through the general path of the direct kernel.
// Entries [lo, min(hi, count)): the rows the scatter set aside are applied BEFORE the segment fold (which closes
// windows on the assumption that it has seen the whole activation), the partials the fold sets aside after it.
__global__ void __launch_bounds__(256) k_spill(Table t, FoldParams p, const SpillRec* list, const StreamVerdict* sv, u32 cap, u32 batch_no,
                                               u32 lo, u32 hi) {
  __shared__ DirtySink sink;
  __shared__ u32 sink_buf[256];
  if (threadIdx.x == 0) {
    sink.n_dirty = 0;
    sink.n_new_keys = 0;
    sink.cap = 256;
    sink.buf = sink_buf;
  }
  __syncthreads();
  const u32 n = min(min(sv->n_spill, cap), hi);
  for (u32 i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const SpillRec r = list[i];
    bw_fold_event<FoldCfgRuntime, true>(t, p, &sink, r.key, r.ts, r.acc, r.seq, batch_no, BW_NO_SLOT, r.weight);
  }
  __syncthreads();
  bw_sinks_flush(&sink, t);
}
// end of an activation's fold stage: empty dirty list, empty spill list, no flags
__global__ void k_stream_reset(Table t, StreamVerdict* sv) {
  t.ctr->dirty_count = 0;
  if (sv) {
    sv->n_spill = 0;
    sv->flags = 0;
  }
}
