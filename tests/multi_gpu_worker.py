"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU).

Every rank ingests its own slice of each activation; the library partitions by
owning rank, exchanges (P2P stores over NVLink or NCCL send/recv) and folds.
Rank 0 replays the same arrival order -- per destination: source rank 0's rows,
then source rank 1's, ... (the order the exchange guarantees) -- through the C
oracle and compares per-destination row sequences exactly.
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bytewax_b200 import _native as N, gpu  # noqa: E402
from oracle import coracle, pyoracle as po  # noqa: E402

A = 1_640_995_200_000_000
S = 1_000_000


def make_batches(world, case):
    rnd = np.random.default_rng(123)
    out = []  # [batch][rank] -> (keys, ts, vals)
    nb, n = 5, 40_000
    for b in range(nb):
        per_rank = []
        for r in range(world):
            if case == "inorder":
                # global stream in time order; rank r holds the r-th slice of each activation
                base = A + (b * world + r) * 10 * S
                ts = base + (np.arange(n) * 10 * S) // n
            else:
                ts = A + b * 12 * S + rnd.integers(-6 * S, 12 * S, n)
            keys = rnd.integers(0, 3000, n).astype(np.uint64) * 7919
            vals = rnd.integers(-100, 100, n)
            per_rank.append((keys, ts.astype(np.int64), vals.astype(np.int64)))
        out.append(per_rank)
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idbuf = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        idbuf.copy_(torch.frombuffer(bytearray(gpu.Context.new_nccl_id()), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    ctx = gpu.Context(local, rank, world, bytes(idbuf.cpu().numpy().tobytes()))
    failures = []
    for xname, xchg in (("p2p", N.XCHG_P2P), ("nccl", N.XCHG_NCCL)):
        for case, red, length, offset, wait in (("inorder", "count", 10, None, 0), ("disorder", "sum", 10, 5, 2),
                                                ("disorder", "count", 10, None, 0)):
            batches = make_batches(world, case)
            fold = gpu.WindowFold(ctx, red, length * S, offset * S if offset else None, A, wait * S, val_dtype="i64",
                                  capacity_hint=8192, max_batch_rows=1 << 16, max_emit_rows=1 << 20, max_late_rows=1 << 21,
                                  exchange=xchg)
            for per_rank in batches:
                k, t, v = per_rank[rank]
                fold.ingest(k, v, t)
            em, em2 = fold.advance(), fold.eof()
            st = fold.stats()
            mine = dict(
                ck=np.concatenate([em.closed_key, em2.closed_key]), cw=np.concatenate([em.closed_window_id, em2.closed_window_id]),
                ca=np.concatenate([em.closed_acc, em2.closed_acc]).astype(np.int64), lk=em.late_key, lw=em.late_window_id,
                lv=em.late_val.astype(np.int64), slow=int(st.slow_batches), stream=int(st.combined_folds))
            fold.close()
            gathered = [None] * world
            dist.all_gather_object(gathered, pickle.dumps(mine))
            if rank == 0:
                for d in range(world):
                    got = pickle.loads(gathered[d])
                    orc = coracle.COracle(red, length * S, offset * S if offset else None, A, wait * S)
                    for per_rank in batches:
                        ks, tss, vs = [], [], []
                        for r in range(world):
                            k, t, v = per_rank[r]
                            m = np.array([po.dest_rank(int(x), world) == d for x in k])
                            ks.append(k[m]); tss.append(t[m]); vs.append(v[m])
                        orc.on_batch(np.concatenate(ks), np.concatenate(tss), np.concatenate(vs))
                    orc.on_eof()
                    ck, cw, ca, _, _ = orc.closed()
                    lk, lw, lv, _, _ = orc.late()
                    ok = (got["ck"].tolist() == ck.tolist() and got["cw"].tolist() == cw.tolist() and got["ca"].tolist() == ca.tolist()
                          and got["lk"].tolist() == lk.tolist() and got["lw"].tolist() == lw.tolist() and got["lv"].tolist() == lv.tolist())
                    if not ok:
                        failures.append((xname, case, red, d, len(ck), len(got["ck"]), len(lk), len(got["lk"])))
                    if case == "inorder" and got["slow"] != 0:
                        failures.append((xname, case, "unexpected slow path", d))
                    # in-order slices over P2P take the streaming path: combine at the source, partials over NVLink, merge
                    if case == "inorder" and xname == "p2p" and os.environ.get("BW_STREAM", "1") != "0" and got["stream"] != len(batches):
                        failures.append((xname, case, "streaming path not taken", d, got["stream"]))
    if rank == 0:
        print("MULTI_GPU_PARITY", "FAIL " + repr(failures) if failures else "OK", flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
