"""`.meta` / `.cache<N>` reader + writer (database.cpp:87-163,247-290; hash_multimap.hpp:970-1082)."""
import os

import numpy as np
import pytest

from metacache_b200 import dbformat
from tests.golden_util import C1, need_c1, G1


def test_cache_round_trip(tmp_path):
    g = G1()
    part = dbformat.CachePart(g.keys, g.sizes, g.values, batch_size=4096)   # several batches
    p = str(tmp_path / "x.cache0")
    dbformat.write_cache(p, part)
    assert os.path.getsize(p) == 24 + 5 * len(g.keys) + 8 * len(g.values)
    back = dbformat.read_cache(p)
    assert np.array_equal(back.keys, g.keys) and np.array_equal(back.sizes, g.sizes)
    assert np.array_equal(back.values, g.values) and back.batch_size == 4096
    hdr_and_batches = list(dbformat.iter_cache_batches(p))
    assert hdr_and_batches[0] == (len(g.keys), len(g.values), 4096)
    assert len(hdr_and_batches) - 1 == -(-len(g.keys) // 4096)


def test_bucket_invariants_of_reference_built_db():
    g = G1()
    assert g.sizes.min() >= 1 and g.sizes.max() == 254          # REP target hits the cap
    assert len(np.unique(g.keys)) == len(g.keys)
    off = np.zeros(len(g.sizes) + 1, np.int64)
    np.cumsum(g.sizes, out=off[1:])
    for i in np.flatnonzero(g.sizes > 1)[:2000]:
        b = g.values[off[i]:off[i + 1]]
        assert np.all(b[1:] > b[:-1])                           # sorted by (tgt, win), distinct


def test_meta_round_trip_synthetic(tmp_path):
    meta = dbformat.synthetic_meta([10, 20, 30], names=["a", "b", "c"], kmerlen=16, sketchlen=16,
                                   winlen=127, winstride=112)
    p = str(tmp_path / "s.meta")
    dbformat.write_meta(p, meta)
    back = dbformat.read_meta(p)
    assert back.target_count == 3 and back.target_names() == ["a", "b", "c"]
    assert list(back.target_windows()) == [10, 20, 30]
    assert (back.kmerlen, back.sketchlen, back.winlen, back.winstride) == (16, 16, 127, 112)
    assert back.max_locations_per_feature == 254


def test_reference_meta_byte_exact_round_trip(tmp_path):
    need_c1()
    src = os.path.join(C1, "bacteria1.meta")
    meta = dbformat.read_meta(src)
    assert meta.target_count == 20 and meta.num_parts == 1
    p = str(tmp_path / "rt.meta")
    dbformat.write_meta(p, meta)
    assert open(p, "rb").read() == open(src, "rb").read()
