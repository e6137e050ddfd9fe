"""Host-side logic of the N>1 path on CPU: 2 ranks, gloo.  The device work (probe of a part) is
stood in for by the oracle; what is tested is the sharding / exchange / merge choreography that
bench.py runs over NCCL: all-gather of sketches, every rank probes all reads against ITS part,
all-to-all of partial top hits by read slice, stable part-ordered merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden_util import G1, G2

MAXC = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nreads, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import mc_oracle as O
    g1, g2 = G1(), G2()
    reads = [r for r in g1.reads[:nreads]]
    per = (len(reads) + world - 1) // world
    mine = reads[rank * per:(rank + 1) * per]
    tab = O.Table(*g2.parts[rank])                      # rank r holds part r
    S, WMAX = 16, 8

    # 1. sketch my slice (fixed-shape tensors: [per, WMAX, S] padded with ~0, window counts)
    feats = torch.full((per, WMAX, S), -1, dtype=torch.int64)
    for i, (a, b) in enumerate(mine):
        sk = [x for x in O.sketch_sequence(a) if x is not None] + [x for x in O.sketch_sequence(b) if x is not None]
        for w, x in enumerate(sk[:WMAX]):
            feats[i, w, :len(x)] = torch.from_numpy(x.astype(np.int64))
    lens = torch.tensor([[len(a), len(b)] for a, b in mine] + [[0, 0]] * (per - len(mine)), dtype=torch.int64)
    # 2. all-gather sketches (+ read lengths for maxWindowsInRange)
    all_feats = [torch.empty_like(feats) for _ in range(world)]
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_feats, feats)
    dist.all_gather(all_lens, lens)
    # 3. probe ALL reads against my part -> partial tops [world, per, MAXC, 4]
    send = torch.zeros((world, per, MAXC, 4), dtype=torch.int64)
    for j in range(world):
        for i in range(per):
            fl = all_feats[j][i]
            locs = []
            for w in range(WMAX):
                for f in fl[w].tolist():
                    if f < 0:
                        continue
                    first = O.C.POINTER(O.C.c_uint64)()
                    O.lib().mco_table_find.restype = O.C.c_uint32
                    O.lib().mco_table_find.argtypes = [O.C.c_void_p, O.C.c_uint32, O.C.POINTER(O.C.POINTER(O.C.c_uint64))]
                    n = O.lib().mco_table_find(tab._h, f, O.C.byref(first))
                    locs += [first[k] for k in range(n)]
            l1, l2 = all_lens[j][i].tolist()
            top = O.candidates(sorted(locs), O.max_windows_in_range(l1, l2), MAXC)
            for c, t in enumerate(top):
                send[j, i, c] = torch.tensor(t, dtype=torch.int64)
    # 4. all-to-all by read slice, then stable part-ordered merge of the lists I received
    recv = torch.zeros_like(send)
    dist.all_to_all_single(recv, send)
    final = []
    for i in range(len(mine)):
        lists = [[tuple(int(x) for x in recv[p, i, c]) for c in range(MAXC) if recv[p, i, c, 1] > 0] for p in range(world)]
        final.append(O.merge_tops(lists, MAXC))
    torch.save(final, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_query_matches_per_part_reference(tmp_path):
    from oracle import mc_oracle as O
    world, nreads = 2, 120
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nreads, str(tmp_path)), nprocs=world, join=True)
    g2 = G2()
    e0, e1 = g2.expected(0), g2.expected(1)
    got = []
    for r in range(world):
        got += torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
    assert len(got) == nreads
    for i in range(nreads):
        assert got[i] == O.merge_tops([e0.top[i], e1.top[i]], MAXC), i
