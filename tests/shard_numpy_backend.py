"""CPU stand-in for the four device steps of the feature-sharded query (tests only): numpy + the
oracle (oracle/mc_oracle.py, pinned to the reference).  It lets the CPU tests run the REAL host logic
of metacache_b200.distributed.FeatureShardedQuery (split sizes, segments, pipelining, collectives)."""
import numpy as np
import torch

from oracle import mc_oracle as O


def hash32_np(x):
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    for _ in range(2):
        x = (((x >> np.uint64(16)) ^ x) * np.uint64(0x45d9f3b)) & m
    return ((x >> np.uint64(16)) ^ x) & m


def shard_of_np(keys, n_shards):
    """common.cuh: shard_of"""
    h = hash32_np(np.asarray(keys, np.uint64) ^ np.uint64(0x9E3779B9))
    return ((h * np.uint64(n_shards)) >> np.uint64(32)).astype(np.int64)


class NumpyBackend:
    device = torch.device("cpu")
    loc_dtype = torch.int64
    torch = torch

    def __init__(self, parts, shard, n_shards, maxc):
        """parts: [(keys, sizes, values)] of every database part, in part order"""
        self.N, self.shard, self.k = n_shards, shard, maxc
        self.table = {}
        # part-major target numbering: the reference merges per-part candidate lists in part order, so on
        # equal hits a target of an earlier part wins whatever its id (see Part::d_part_of in csrc/api.cu)
        nt = int(max(int((v >> np.uint64(32)).max()) for _, _, v in parts if len(v))) + 1
        part_of = np.full(nt, 255, np.int64)
        for p, (_, _, v) in enumerate(parts):
            part_of[(v >> np.uint64(32)).astype(np.int64)] = p
        self.orig = np.argsort(part_of, kind="stable")
        new_of_old = np.empty(nt, np.uint64)
        new_of_old[self.orig] = np.arange(nt, dtype=np.uint64)
        parts = [(k, s_, (new_of_old[(v >> np.uint64(32)).astype(np.int64)] << np.uint64(32)) | (v & np.uint64(0xFFFFFFFF)))
                 for k, s_, v in parts]
        for keys, sizes, values in parts:
            off = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
            mine = np.flatnonzero(shard_of_np(keys, n_shards) == shard)
            for i in mine:
                v = values[off[i]:off[i + 1]]
                k = int(keys[i])
                self.table[k] = np.concatenate([self.table[k], v]) if k in self.table else v
        self._lists = {}

    def empty(self, n, dtype):
        return torch.zeros(max(int(n), 1), dtype=dtype)

    def route(self, slot, feats, qwo, nq, S, pos, send):
        """feats: uint32 [nwin_total, S] numpy (0xFFFFFFFF padded); qwo: absolute first windows [nq + 1]"""
        N = self.N
        cnt = np.zeros(N * (nq + 1) + 1, np.int64)
        per = []
        for q in range(nq):
            f = feats[qwo[q]:qwo[q + 1]].reshape(-1)
            f = f[f != 0xFFFFFFFF]
            o = shard_of_np(f, N)
            per.append((f, o))
            for s in range(N):
                cnt[s * (nq + 1) + q] = int((o == s).sum())
        p = np.concatenate([[0], np.cumsum(cnt)])[:-1]
        pos[:len(p)] = torch.from_numpy(p.astype(np.int32))
        out = send.numpy()
        for q, (f, o) in enumerate(per):
            for s in range(N):
                sel = f[o == s]
                b = p[s * (nq + 1) + q]
                out[b:b + len(sel)] = sel.astype(np.uint32).view(np.int32)

    def probe(self, slot, feats, n, off, data):
        f = feats[:n].numpy().view(np.uint32)
        lists = [self.table.get(int(x), np.zeros(0, np.uint64)) for x in f]
        self._lists[slot] = lists
        sizes = np.array([len(x) for x in lists], np.int64)
        off[:n + 1] = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32))
        data[:n] = torch.arange(n, dtype=torch.int64)

    def gather(self, slot, off, data, n, locs):
        lists = self._lists[slot]
        if n and sum(len(x) for x in lists):
            cat = np.concatenate([lists[int(i)] for i in data[:n]])
            locs[:len(cat)] = torch.from_numpy(cat.astype(np.uint64).view(np.int64))

    def reduce(self, slot, pos, runs, max_win, nq, top, mean_locations=0.0):
        pos = pos.numpy().astype(np.int64)
        out = top.numpy().view(np.uint32)
        out[:, :, 0] = 0xFFFFFFFF
        out[:, :, 1:] = 0
        for q in range(nq):
            locs = []
            for o, (rl, ro, nf, nl) in enumerate(runs):
                seg = pos[o * (nq + 1)]
                i0, i1 = pos[o * (nq + 1) + q] - seg, pos[o * (nq + 1) + q + 1] - seg
                if i1 <= i0:
                    continue
                ro_ = ro.numpy().astype(np.int64)
                b = ro_[i0] - ro_[0]
                e = nl if i1 >= nf else ro_[i1] - ro_[0]
                locs += rl[b:e].numpy().view(np.uint64).tolist()
            for c, t in enumerate(O.candidates(sorted(locs), int(max_win[q]), self.k)):
                out[q, c] = (int(self.orig[t[0]]),) + tuple(t[1:])

    def check(self, slot):
        return False
