"""CPU restatement of the N>1 data path over gloo (world_size 2): route with the library's
bw_route, all-to-all the rows in (source rank, source order), fold with the C oracle.
Checks the host-side contract the GPU exchange implements: every key has one owner and
arrival order per destination is source-major."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bytewax_b200 import _native as N  # noqa: E402
from oracle import coracle  # noqa: E402

A, S = 1_640_995_200_000_000, 1_000_000


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    lib = N.load()
    rnd = np.random.default_rng(7)  # same stream on every rank
    orc = coracle.COracle("count", 10 * S, None, A, 2 * S)
    ref = [coracle.COracle("count", 10 * S, None, A, 2 * S) for _ in range(world)] if rank == 0 else None
    for b in range(4):
        slices = []
        for r in range(world):
            n = 3000
            keys = rnd.integers(0, 500, n).astype(np.uint64) * 104729
            ts = (A + b * 9 * S + rnd.integers(-4 * S, 9 * S, n)).astype(np.int64)
            slices.append((keys, ts))
        keys, ts = slices[rank]
        dest = np.array([lib.bw_route(int(k), world) for k in keys])
        send = [torch.from_numpy(np.stack([keys[dest == d].view(np.int64), ts[dest == d]])) for d in range(world)]
        counts = torch.tensor([s.shape[1] for s in send])
        allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allc, counts)
        recv = [torch.zeros((2, int(allc[r][rank])), dtype=torch.int64) for r in range(world)]
        # gloo has no all_to_all for CPU tensors in every build: use pairwise send/recv
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r] = send[r]
            else:
                reqs.append(dist.isend(send[r].contiguous(), r))
                reqs.append(dist.irecv(recv[r], r))
        for q in reqs:
            q.wait()
        k_in = np.concatenate([recv[r][0].numpy().view(np.uint64) for r in range(world)])  # source-major arrival order
        t_in = np.concatenate([recv[r][1].numpy() for r in range(world)])
        orc.on_batch(k_in, t_in)
        if rank == 0:
            for d in range(world):
                ks = np.concatenate([s[0][np.array([lib.bw_route(int(k), world) for k in s[0]]) == d] for s in slices])
                tss = np.concatenate([s[1][np.array([lib.bw_route(int(k), world) for k in s[0]]) == d] for s in slices])
                ref[d].on_batch(ks, tss)
    orc.on_eof()
    ck, cw, ca, _, _ = orc.closed()
    lk, lw, _, _, _ = orc.late()
    mine = [ck.tolist(), cw.tolist(), ca.tolist(), lk.tolist(), lw.tolist()]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        for d in range(world):
            ref[d].on_eof()
            rk, rw, ra, _, _ = ref[d].closed()
            rlk, rlw, _, _, _ = ref[d].late()
            ok &= gathered[d] == [rk.tolist(), rw.tolist(), ra.tolist(), rlk.tolist(), rlw.tolist()]
            ok &= all(lib.bw_route(int(k), world) == d for k in rk)
        # every key is owned by exactly one rank
        owners = {}
        for d in range(world):
            for k in set(gathered[d][0]):
                ok &= owners.setdefault(k, d) == d
        print("GLOO_PARITY", "OK" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
