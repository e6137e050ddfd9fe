"""Host-side mirror of the reference's query seam over the C ABI (libmcb200.so).

Same names, argument meaning and error behaviour as the reference interface for
this path, so that parity tests read like the reference's own:

  reference                                       here
  ----------------------------------------------  -------------------------------
  sketching_options          hash_dna.hpp:99      SketchingOpt
  candidate_generation_rules candidate_structs:110 CandidateGenerationRules
  make_candidate_generation_rules      :134-151   make_candidate_generation_rules
  database::read             database.cpp:183-242 Database.read
  database::query_gpu_async  database.hpp:386-397 Database.query_gpu_async
  query_batch<location>      query_batch.cuh:346  QueryBatch
  query_host_data            query_batch.cuh:60   QueryHostData (host_data(hostId))
  match_candidate            candidate_structs:80 MatchCandidate

All compute happens in the CUDA library; nothing here touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Candidate, Sketching, check, check_ptr, lib
from .dbformat import DbMeta, RANK_NONE, RANK_SEQUENCE, read_meta

NO_TARGET = 0xFFFFFFFF


@dataclasses.dataclass(frozen=True)
class SketchingOpt:
    kmerlen: int = 16
    sketchlen: int = 16
    winlen: int = 127
    winstride: int = 112

    def c(self) -> Sketching:
        return Sketching(self.kmerlen, self.sketchlen, self.winlen, self.winstride)


@dataclasses.dataclass
class CandidateGenerationRules:
    maxWindowsInRange: int = 3
    maxCandidates: int = 2
    mergeBelow: int = RANK_SEQUENCE


def make_candidate_generation_rules(len1: int, len2: int = 0, insert_size_max: int = 0,
                                    winstride: int = 112, max_candidates: int = 2,
                                    lowest_rank: int = RANK_SEQUENCE) -> CandidateGenerationRules:
    """candidate_structs.hpp:134-151"""
    return CandidateGenerationRules(2 + max(len1 + len2, insert_size_max) // winstride,
                                    max_candidates, lowest_rank)


@dataclasses.dataclass(frozen=True)
class MatchCandidate:
    tgt: int
    hits: int
    beg: int
    end: int

    def as_tuple(self):
        return (self.tgt, self.hits, self.beg, self.end)


class Database:
    """Query half of mc::database with feature_store = the B200 table (database.hpp:183-189)."""

    def __init__(self, device: int = 0, n_parts: int = 1, devices: Optional[Sequence[int]] = None):
        """devices: one CUDA device per part (mcb200_db_open_multi: a store over several GPUs of this
        process, devices[0] = home device); default: all parts on `device`"""
        if devices is not None:
            import ctypes
            arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            self._h = check_ptr(lib().mcb200_db_open_multi(len(devices), arr))
            device = int(devices[0])
        else:
            self._h = check_ptr(lib().mcb200_db_open(device, n_parts))
        self.device = device
        self.meta: Optional[DbMeta] = None
        self._lowest_rank = RANK_SEQUENCE

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib().mcb200_db_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- loading (database.cpp:183-242) ----------------------------------------------
    @classmethod
    def read(cls, filename: str, single_part_id: int = -1, device: int = 0,
             max_load_factor: float = 0.0) -> "Database":
        meta = read_meta(filename + ".meta")
        if single_part_id >= 0:
            if single_part_id >= meta.num_parts:
                raise RuntimeError(f"Database part '.cache{single_part_id}' is not available. "
                                   f"Database has only {meta.num_parts} parts.")
            parts = [single_part_id]
        else:
            parts = list(range(meta.num_parts))
        db = cls(device, len(parts))
        db.meta = meta
        for i, p in enumerate(parts):
            path = f"{filename}.cache{p}"
            if not os.path.exists(path):
                raise FileNotFoundError(f"Could not read database file '{path}'")
            check(lib().mcb200_db_load_cache_file(db._h, i, path.encode(), max_load_factor))
        return db

    def load_part_arrays(self, part: int, keys, sizes, values, max_load_factor: float = 0.0,
                         batch: int = 1 << 20):
        """read_binary(istream, store, part) with the `.cache` arrays already in memory
        (streamed in batches of `batch` keys, as the file format is)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint8)
        values = np.ascontiguousarray(values, dtype=np.uint64)
        check(lib().mcb200_db_part_begin(self._h, part, len(keys), len(values), max_load_factor))
        off = np.zeros(len(keys) + 1, dtype=np.int64)
        np.cumsum(sizes, dtype=np.int64, out=off[1:])
        for b0 in range(0, len(keys), batch):
            b1 = min(len(keys), b0 + batch)
            k, s, v = keys[b0:b1], sizes[b0:b1], values[off[b0]:off[b1]]
            check(lib().mcb200_db_part_append(self._h, part, k.ctypes.data, s.ctypes.data,
                                              v.ctypes.data, len(k), len(v)))
        check(lib().mcb200_db_part_finish(self._h, part))

    def export_part(self, part: int):
        nk, nv = self.key_count(part), self.value_count(part)
        keys = np.zeros(nk, np.uint32)
        sizes = np.zeros(nk, np.uint8)
        values = np.zeros(nv, np.uint64)
        check(lib().mcb200_db_part_export(self._h, part, keys.ctypes.data, sizes.ctypes.data,
                                          values.ctypes.data))
        return keys, sizes, values

    # -- metadata ------------------------------------------------------------------
    def target_sketching(self) -> SketchingOpt:
        m = self.meta
        return SketchingOpt(m.kmerlen, m.sketchlen, m.winlen, m.winstride) if m else SketchingOpt()

    def part_count(self) -> int:
        return lib().mcb200_db_part_count(self._h)

    def key_count(self, part: int = 0) -> int:
        return lib().mcb200_db_key_count(self._h, part)

    def value_count(self, part: int = 0) -> int:
        return lib().mcb200_db_value_count(self._h, part)

    def bucket_count(self, part: int = 0) -> int:
        return lib().mcb200_db_bucket_count(self._h, part)

    def device_bytes(self, part: int = 0) -> int:
        return lib().mcb200_db_device_bytes(self._h, part)

    @staticmethod
    def max_supported_locations_per_feature() -> int:
        return lib().mcb200_max_supported_locations_per_feature()

    # -- `-lowest <rank>`: taxon key per target (taxonomy.hpp:576-597, 1260-1267) -----
    def ranked_lineage_keys(self, lowest_rank: int) -> np.ndarray:
        """lowest_ranked_ancestor(tgt, lowest) for every target as an opaque non-zero key
        (0 = none).  Built from the taxonomy stored in `.meta`."""
        m = self.meta
        by_id = {t.id: t for t in m.taxa}
        keys = np.zeros(m.target_count, dtype=np.uint64)
        for t in m.taxa:
            if t.id >= 0:
                continue
            tgt = -t.id - 1
            lin = [None] * RANK_NONE
            if t.rank != RANK_NONE:
                lin[t.rank] = t.id
            pid = t.parent
            while pid != 0:
                a = by_id.get(pid)
                if a is None or a.id < 0:
                    break
                if a.rank != RANK_NONE:
                    lin[a.rank] = a.id
                if a.parent == pid:
                    break
                pid = a.parent
            for r in range(lowest_rank, RANK_NONE):
                if lin[r] is not None:
                    keys[tgt] = np.uint64(lin[r] & 0xFFFFFFFFFFFFFFFF)
                    break
        return keys

    # -- ranked lineages (taxonomy.hpp:368, 576-597) for classify() ---------------------
    def target_lineages(self) -> np.ndarray:
        """[n_targets, 21] u32: taxon ordinal + 1 (index into meta.taxa) at every rank, 0 = none;
        `taxonomy::ranked_lineage` of each target as make_ranks builds it."""
        m = self.meta
        ordinal = {t.id: i + 1 for i, t in enumerate(m.taxa)}
        by_id = {t.id: t for t in m.taxa}
        lin = np.zeros((m.target_count, RANK_NONE), dtype=np.uint32)
        for t in m.taxa:
            if t.id >= 0:
                continue
            tgt = -t.id - 1
            if tgt >= m.target_count:
                continue
            if t.rank != RANK_NONE:
                lin[tgt, t.rank] = ordinal[t.id]
            pid = t.parent
            while pid != 0:
                a = by_id.get(pid)
                if a is None or a.id < 0:
                    break
                if a.rank != RANK_NONE:
                    lin[tgt, a.rank] = ordinal[a.id]
                if a.parent == pid:
                    break
                pid = a.parent
        return lin

    def copy_target_lineages_to_gpus(self, lineages: Optional[np.ndarray] = None):
        """gpu_hashmap::copy_target_lineages_to_gpus: ranked lineages for on-device classify()"""
        if lineages is None:
            lineages = self.target_lineages()
        lineages = np.ascontiguousarray(lineages, dtype=np.uint32)
        check(lib().mcb200_db_set_target_lineages(self._h, lineages.ctypes.data, lineages.shape[0]))
        self._lineages = lineages

    def set_lowest_rank(self, lowest_rank: int):
        """copy_target_lineages_to_gpus + the `lowestRank` argument of query_gpu_async."""
        if lowest_rank == self._lowest_rank:
            return
        if lowest_rank <= RANK_SEQUENCE:
            check(lib().mcb200_db_set_target_taxa(self._h, None, 0))
        else:
            keys = self.ranked_lineage_keys(lowest_rank)
            check(lib().mcb200_db_set_target_taxa(self._h, keys.ctypes.data, len(keys)))
        self._lowest_rank = lowest_rank

    def set_target_taxa(self, keys: Optional[np.ndarray]):
        if keys is None:
            check(lib().mcb200_db_set_target_taxa(self._h, None, 0))
            self._lowest_rank = RANK_SEQUENCE
        else:
            keys = np.ascontiguousarray(keys, dtype=np.uint64)
            check(lib().mcb200_db_set_target_taxa(self._h, keys.ctypes.data, len(keys)))
            self._lowest_rank = -1

    # -- the seam (database.hpp:386-397) -----------------------------------------
    def query_gpu_async(self, batch: "QueryBatch", host_id: int, sketching: SketchingOpt,
                        lowest_rank: int = RANK_SEQUENCE):
        if self._lowest_rank != -1:
            self.set_lowest_rank(lowest_rank)
        sk = sketching.c()
        check(lib().mcb200_batch_submit(batch._h, host_id, C.byref(sk)))


class QueryHostData:
    """query_batch<location>::query_host_data (query_batch.cuh:60-259)."""

    def __init__(self, batch: "QueryBatch", host_id: int):
        self._b, self._id = batch, host_id

    def num_queries(self) -> int:
        return lib().mcb200_batch_num_queries(self._b._h, self._id)

    def num_windows(self) -> int:
        return lib().mcb200_batch_num_windows(self._b._h, self._id)

    def wait_for_results(self):
        check(lib().mcb200_batch_wait(self._b._h, self._id))

    def clear(self):
        check(lib().mcb200_batch_clear(self._b._h, self._id))

    def top_candidates(self, i: int) -> List[MatchCandidate]:
        """span of maxCandidatesPerQuery entries; unused ones have hits == 0."""
        p = lib().mcb200_batch_top_candidates(self._b._h, self._id, i)
        if not p:
            return []
        return [MatchCandidate(p[c].tgt, p[c].hits, p[c].beg, p[c].end)
                for c in range(self._b.max_candidates)]

    def top_candidates_array(self) -> np.ndarray:
        """[num_queries, max_candidates, 4] u32 view copy of all top candidates."""
        n = self.num_queries()
        p = lib().mcb200_batch_top_candidates(self._b._h, self._id, 0)
        if not p or n == 0:
            return np.zeros((0, self._b.max_candidates, 4), np.uint32)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n, self._b.max_candidates, 4))
        return a.copy()

    def allhits(self, i: int) -> np.ndarray:
        """sorted locations as u64 (tgt << 32 | win); empty unless copy_all_hits."""
        n = C.c_uint64(0)
        p = lib().mcb200_batch_allhits(self._b._h, self._id, i, C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, np.uint64)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def sketches(self, i: int) -> List[np.ndarray]:
        """window sketches of query i (mate 1 windows first), one array per window."""
        w0 = lib().mcb200_batch_query_window_offset(self._b._h, self._id, i)
        w1 = lib().mcb200_batch_query_window_offset(self._b._h, self._id, i + 1)
        out = []
        for w in range(w0, w1):
            n = C.c_uint32(0)
            p = lib().mcb200_batch_sketch(self._b._h, self._id, w, C.byref(n))
            out.append(np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32))
        return out

    def classifications(self) -> np.ndarray:
        """[num_queries, 2] u32 (taxon ordinal + 1 or 0, rank) after wait_for_results, if enabled"""
        n = self.num_queries()
        p = lib().mcb200_batch_classifications(self._b._h, self._id)
        if not p or n == 0:
            return np.zeros((0, 2), np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n, 2)).copy()

    def last_timing(self):
        t, k = C.c_float(0), C.c_float(0)
        check(lib().mcb200_batch_last_timing(self._b._h, self._id, C.byref(t), C.byref(k)))
        return t.value, k.value


class QueryBatch:
    """query_batch<location> (query_batch.cuh:346-423): pinned host buffers + device buffers
    per host thread; results come back into the host buffers."""

    def __init__(self, db: Database, max_queries: int, max_bases: Optional[int] = None,
                 max_candidates: int = 2, copy_all_hits: bool = False, num_host_threads: int = 1):
        if max_bases is None:
            max_bases = max_queries * 512
        self.db = db
        self.max_candidates = max_candidates
        self.copy_all_hits = copy_all_hits
        self._h = check_ptr(lib().mcb200_batch_create(db._h, max_queries, max_bases, max_candidates,
                                                      int(copy_all_hits), num_host_threads))
        self._hosts = [QueryHostData(self, i) for i in range(num_host_threads)]

    def close(self):
        if getattr(self, "_h", None):
            lib().mcb200_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_classification(self, hits_min: int, hits_diff_fraction: float = 1.0,
                              lowest_rank: int = RANK_SEQUENCE, highest_rank: int = 19):
        """classify() (classification.cpp:146-189) on the device for every query of a submit;
        hitsMin default = sketchlen/3 (querying.cpp:256-260), highest default = domain"""
        check(lib().mcb200_batch_enable_classification(self._h, hits_min, hits_diff_fraction, lowest_rank,
                                                       highest_rank))

    def host_data(self, host_id: int) -> QueryHostData:
        return self._hosts[host_id]

    def add_paired_read(self, host_id: int, seq1: bytes, seq2: bytes = b"",
                        sketching: Optional[SketchingOpt] = None,
                        rules: Optional[CandidateGenerationRules] = None) -> bool:
        """query_batch.cuh:85-186.  Returns False if the batch is full (read not added)."""
        if rules is None:
            stride = (sketching or self.db.target_sketching()).winstride
            rules = make_candidate_generation_rules(len(seq1), len(seq2), 0, stride, self.max_candidates)
        rc = check(lib().mcb200_batch_add_read(self._h, host_id, seq1, len(seq1), seq2, len(seq2),
                                               rules.maxWindowsInRange))
        return rc == 1

    def add_reads(self, host_id: int, bases: np.ndarray, offsets: np.ndarray, paired: bool = False,
                  insert_size_max: int = 0, winstride: Optional[int] = None) -> int:
        """bulk add_paired_read over a concatenated base buffer; returns reads added."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nseq = len(offsets) - 1
        nq = nseq // 2 if paired else nseq
        if winstride is None:
            winstride = self.db.target_sketching().winstride
        return check(lib().mcb200_batch_add_reads(self._h, host_id, bases.ctypes.data, offsets.ctypes.data,
                                                  nq, int(paired), insert_size_max, winstride))


def default_hits_min(sketchlen: int) -> int:
    """querying.cpp:256-263"""
    if sketchlen >= 6:
        return int(sketchlen / 3.0)
    return 2 if sketchlen >= 4 else 1


def query_reads(db: Database, reads: Sequence, sketching: Optional[SketchingOpt] = None,
                max_candidates: int = 2, insert_size_max: int = 0, copy_all_hits: bool = True,
                lowest_rank: int = RANK_SEQUENCE, batch_queries: int = 8192, with_sketches: bool = False,
                classify: bool = False, hits_diff_fraction: float = 1.0, highest_rank: int = 19):
    """query_gpu (database_query.hpp:87-124) for a list of reads (bytes or (bytes, bytes)):
    fills batches, submits, waits, collects (allhits, top candidates[, sketches]) per read."""
    sk = sketching or db.target_sketching()
    pairs = [(r, b"") if isinstance(r, (bytes, bytearray)) else (r[0], r[1]) for r in reads]
    max_bases = max(1 << 20, 2 * max((len(a) + len(b) for a, b in pairs), default=0))
    qb = QueryBatch(db, batch_queries, max_bases, max_candidates, copy_all_hits, 1)
    if classify:
        qb.enable_classification(default_hits_min(sk.sketchlen), hits_diff_fraction, lowest_rank, highest_rank)
    hd = qb.host_data(0)
    results = []

    def flush():
        if hd.num_queries() == 0:
            return
        db.query_gpu_async(qb, 0, sk, lowest_rank)
        hd.wait_for_results()
        cls = hd.classifications() if classify else None
        for s in range(hd.num_queries()):
            top = [c.as_tuple() for c in hd.top_candidates(s) if c.hits > 0]
            item = [hd.allhits(s) if copy_all_hits else None, top]
            if with_sketches:
                item.append(hd.sketches(s))
            if classify:
                item.append((int(cls[s, 0]), int(cls[s, 1])))
            results.append(tuple(item))
        hd.clear()

    for a, b in pairs:
        rules = make_candidate_generation_rules(len(a), len(b), insert_size_max, db.target_sketching().winstride,
                                                max_candidates, lowest_rank)
        if not qb.add_paired_read(0, a, b, sk, rules):
            flush()
            if not qb.add_paired_read(0, a, b, sk, rules):
                raise RuntimeError("query batch is too small for a single read!")
    flush()
    qb.close()
    return results
