"""The oracle (oracle/mc_oracle.c) against reference-generated golden vectors. CPU only."""
import os

import numpy as np
import pytest

from oracle import mc_oracle as O
from tests.golden_util import C1, need_c1, G1, G2, kat


def test_kat_hash_and_revcomp():
    z = kat()
    for x, y in zip(z["hash_in"], z["hash_out"]):
        assert O.hash32(int(x)) == int(y)
    assert O.revcomp32(0x1B, 4) == 27
    assert O.canonical32(0xFF, 4) == 0
    assert O.revcomp32(0x12345678, 16) == 0xD26AE37B
    assert O.canonical32(0x12345678, 16) == 0x12345678


def test_kat_sketches():
    z = kat()
    sk = O.sketch_sequence(z["S"].tobytes())
    assert len(sk) == 2
    assert np.array_equal(sk[0], z["S_win0"]) and np.array_equal(sk[1], z["S_win1"])
    assert np.array_equal(O.sketch_sequence(z["amb_seq"].tobytes())[0], z["amb_feats"])
    assert np.array_equal(O.sketch_sequence(z["low_seq"].tobytes())[0], z["low_feats"])


@pytest.mark.parametrize("n,expect", [(0, 1), (16, 1), (127, 1), (128, 2), (150, 2), (239, 3), (240, 3),
                                      (351, 4), (352, 4), (19000, 170)])
def test_num_windows(n, expect):
    # hash_dna.hpp:54-75
    assert O.num_windows(n) == expect


@pytest.fixture(scope="module")
def g1():
    g = G1()
    g.table = O.Table(g.keys, g.sizes, g.values)
    return g


@pytest.mark.parametrize("tag,maxc,insert", [("c2_", 2, 0), ("c5_", 5, 0), ("c2i_", 2, 1000)])
def test_oracle_matches_reference_g1(g1, tag, maxc, insert):
    exp = g1.expected(tag)
    assert len(exp.top) == len(g1.reads)
    for i, (a, b) in enumerate(g1.reads):
        sk = [x for x in O.sketch_sequence(a, g1.k, g1.s, g1.w, g1.stride) if x is not None]
        sk += [x for x in O.sketch_sequence(b, g1.k, g1.s, g1.w, g1.stride) if x is not None]
        assert len(sk) == len(exp.sketches[i]), i
        for x, y in zip(sk, exp.sketches[i]):
            assert np.array_equal(x, y), i
        allh, top = O.query(g1.table, a, b, g1.k, g1.s, g1.w, g1.stride, maxc=maxc, insert_size_max=insert)
        assert np.array_equal(allh, exp.allhits[i]), i
        assert top == exp.top[i], i


def test_oracle_two_parts_matches_reference_per_part(g1):
    g2 = G2()
    tabs = [O.Table(*p) for p in g2.parts]
    exps = [g2.expected(0), g2.expected(1)]
    for i, (a, b) in enumerate(g1.reads):
        tops = []
        for p in (0, 1):
            allh, top = O.query(tabs[p], a, b)
            assert np.array_equal(allh, exps[p].allhits[i]), (i, p)
            assert top == exps[p].top[i], (i, p)
            tops.append(top)
        # intended all-parts semantics == stable part-ordered merge of per-part tops
        allh, top = O.query(tabs, a, b)
        assert np.array_equal(allh, np.concatenate([exps[0].allhits[i], exps[1].allhits[i]])), i
        assert top == O.merge_tops(tops, 2), i


def test_candidates_tie_breaks_and_ranges():
    # candidate_generation.hpp:47-108: first strictly best range wins; stable top-k
    loc = lambda t, w: (t << 32) | w
    locs = [loc(1, 5), loc(1, 6), loc(1, 9), loc(1, 10), loc(2, 1), loc(2, 1), loc(3, 7), loc(3, 8)]
    assert O.candidates(locs, 3, 2) == [(1, 2, 5, 6), (2, 2, 1, 1)]
    assert O.candidates(locs, 3, 3) == [(1, 2, 5, 6), (2, 2, 1, 1), (3, 2, 7, 8)]
    assert O.candidates(locs, 6, 1) == [(1, 4, 5, 10)]
    assert O.candidates([], 3, 2) == []
    # taxon merge (-lowest above sequence): targets 1 and 3 share a taxon
    tax = np.asarray([0, 7, 8, 7], np.uint64)
    assert O.candidates(locs + [loc(3, 8)], 3, 2, tax) == [(3, 3, 7, 8), (2, 2, 1, 1)]


def test_oracle_matches_reference_golden_file():
    """oracle vs the reference's OWN golden test/data/classified.expected (single.fa section)."""
    need_c1()
    from metacache_b200 import dbformat, formatting
    from oracle import refio
    meta = dbformat.read_meta(os.path.join(C1, "bacteria1.meta"))
    c = dbformat.read_cache(os.path.join(C1, "bacteria1.cache0"))
    tab = O.Table(c.keys, c.sizes, c.values)
    names = meta.target_names()
    reads = refio.read_fasta(os.path.join(C1, "single.fa"))
    expected = {}
    section = None
    for line in open(os.path.join(C1, "classified.expected")):
        if line.startswith("# data/"):
            section = line.strip()[7:]
        if section != "single" or line.startswith("#"):
            continue
        cols = line.rstrip("\n").split("\t|\t")
        if len(cols) == 6:
            expected[cols[1]] = (cols[3], cols[4])
    assert len(expected) > 10000
    checked = 0
    for hdr, seq in reads:
        h = hdr.split()[0]
        if h not in expected:
            continue
        allh, top = O.query(tab, seq, b"")
        assert formatting.format_all_hits(allh, names) == expected[h][0], h
        want_top = expected[h][1] if expected[h][1] != "--" else ""
        assert formatting.format_top_hits(top, names) == want_top, h
        checked += 1
    assert checked > 10000


def test_oracle_classification_matches_reference_golden_file():
    """classify() restatement vs the last column of classified.expected (all three sections)"""
    need_c1()
    from metacache_b200 import dbformat, formatting
    from metacache_b200.database import Database
    from metacache_b200.statistics import ClassificationStatistics, TaxonCounts
    from oracle import refio
    meta = dbformat.read_meta(os.path.join(C1, "bacteria1.meta"))
    c = dbformat.read_cache(os.path.join(C1, "bacteria1.cache0"))
    tab = O.Table(c.keys, c.sizes, c.values)
    db = Database.__new__(Database)
    db.meta = meta
    lin = Database.target_lineages(db)
    expected, section = {}, None
    for line in open(os.path.join(C1, "classified.expected")):
        if line.startswith("# data/"):
            section = line.strip()[7:]
        if line.startswith("#"):
            continue
        cols = line.rstrip("\n").split("\t|\t")
        if len(cols) == 6 and cols[0].isdigit():
            expected.setdefault(section, {})[int(cols[0])] = cols[5]
    single = refio.read_fasta(os.path.join(C1, "single.fa"))
    pf = refio.read_fasta(os.path.join(C1, "pairs.fa"))
    runs = {"single": [(s, b"") for _, s in single],
            "pairs": [(pf[i][1], pf[i + 1][1]) for i in range(0, len(pf), 2)]}
    classified = 0
    summaries, abundances = {}, {}
    for sec, items in runs.items():
        cls = []
        for qid, (a, b) in enumerate(items, start=1):
            _, top = O.query(tab, a, b)
            t, r = O.classify(top, lin, hits_min=5)
            got = formatting.format_classification(t, r, meta.taxa)
            assert got == expected[sec][qid], (sec, qid, top)
            classified += t != 0
            cls.append((t, r))
        st = ClassificationStatistics()
        st.assign_batch(np.asarray(cls, np.uint32))
        summaries[sec] = st.summary_lines()
        # -abundances -abundance-per species (test/run_tests:153): per-taxon counts, then the estimate
        tc = TaxonCounts(meta.taxa)
        half = len(cls) // 2                                        # two workers' maps merged = one map
        tc.count_batch(np.asarray(cls[:half], np.uint32))
        other = TaxonCounts(meta.taxa)
        other.count_batch(np.asarray(cls[half:], np.uint32))
        tc.merge(other)
        lines = tc.abundance_lines(st)
        tc.estimate_abundance(4)
        abundances[sec] = lines + tc.estimate_lines(st, 4)
    assert classified > 150
    # the per-rank summary block the reference prints after each section (show_taxon_statistics)
    want, section = {}, None
    for line in open(os.path.join(C1, "classified.expected")):
        line = line.rstrip("\n")
        if line.startswith("# data/"):
            section = line[7:].split()[0]
        if line.startswith("# unclassified:") or line.startswith("# classified:") or line.startswith("#   "):
            want.setdefault(section, []).append(line)
    for sec in runs:
        key = next(k for k in want if k.startswith(sec))
        assert summaries[sec] == want[key], sec
    # the abundance tables, against what the unmodified CPU reference prints for its own CLI test
    # (oracle/_ref/c1/cli_cpu_reference.out; upstream's classified.expected predates a change of these rows)
    from tests.golden_util import reference_abundance_blocks
    blocks = reference_abundance_blocks(os.path.join(C1, "cli_cpu_reference.out"))
    assert len(blocks) == 3                                        # single, -pairseq, -pairfiles
    assert abundances["single"] == blocks[0]
    assert abundances["pairs"] == blocks[1]
    est = blocks[0][blocks[0].index("# estimated abundance (number of queries) per species"):]
    species = [x.split("\t|\t") for x in est if x.startswith("species:")]
    assert len(species) >= 3 and any("." in x[2] for x in species)   # the proportional split is exercised


@pytest.mark.parametrize("gi", range(5))
def test_oracle_matches_reference_for_other_geometries(gi):
    """(k, s, w, stride) other than the default: databases built and queried by the reference itself
    (tests/golden/g3.npz) - sketches, sorted all-hits and top candidates"""
    from tests.golden_util import G3
    g3 = G3()
    k, s, w, stride = g3.geometries[gi]
    exp = g3.expected(gi)
    tab = O.Table(*g3.part(gi))
    for i, (a, b) in enumerate(g3.reads):
        sk = [x for x in O.sketch_sequence(a, k, s, w, stride) if x is not None]
        sk += [x for x in O.sketch_sequence(b, k, s, w, stride) if x is not None]
        assert len(sk) == len(exp.sketches[i]), i
        for x, y in zip(sk, exp.sketches[i]):
            assert np.array_equal(x, y), i
        allh, top = O.query(tab, a, b, k, s, w, stride, maxc=2)
        assert np.array_equal(allh, exp.allhits[i]), i
        assert top == exp.top[i], i


def test_abundance_estimates_at_every_rank_match_the_reference_cli():
    """`-abundances -abundance-per <rank>` for eight ranks from sequence to domain: the tables of
    tests/golden/abundance.json were printed by the unmodified CPU reference (oracle/make_abundance_golden.py);
    here the classifications come from the oracle and the tables from metacache_b200.statistics"""
    need_c1()
    import json
    from metacache_b200 import dbformat
    from metacache_b200.database import Database
    from metacache_b200.formatting import RANK_NAMES
    from metacache_b200.statistics import ClassificationStatistics, TaxonCounts
    from oracle import refio
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "abundance.json")))["blocks"]
    meta = dbformat.read_meta(os.path.join(C1, "bacteria1.meta"))
    c = dbformat.read_cache(os.path.join(C1, "bacteria1.cache0"))
    tab = O.Table(c.keys, c.sizes, c.values)
    db = Database.__new__(Database)
    db.meta = meta
    lin = Database.target_lineages(db)
    single = refio.read_fasta(os.path.join(C1, "single.fa"))
    pf = refio.read_fasta(os.path.join(C1, "pairs.fa"))
    runs = {"single": [(s, b"") for _, s in single],
            "pairs": [(pf[i][1], pf[i + 1][1]) for i in range(0, len(pf), 2)]}
    for sec, items in runs.items():
        cls = np.asarray([O.classify(O.query(tab, a, b)[1], lin, hits_min=5) for a, b in items], np.uint32)
        st = ClassificationStatistics()
        st.assign_batch(cls)
        for rank_name, blocks in gold.items():
            rank = RANK_NAMES.index(rank_name)
            tc = TaxonCounts(meta.taxa)
            tc.count_batch(cls)
            lines = tc.abundance_lines(st)
            tc.estimate_abundance(rank)
            assert lines + tc.estimate_lines(st, rank) == blocks[sec], (sec, rank_name)


def test_oracle_candidates_and_classification_under_other_options_match_the_reference_cli():
    """-hitmin / -hitdiff / -maxcand / -lowest / -highest away from their defaults: top_hits and classification
    columns printed by the unmodified CPU reference for its own test reads (tests/golden/classify_options.json,
    oracle/make_classify_golden.py) against the oracle's candidates (taxon merge below `lowest`,
    candidate_generation.hpp:172-231) and classify() (classification.cpp:146-189)"""
    need_c1()
    import json
    from metacache_b200 import dbformat, formatting
    from metacache_b200.database import Database
    from oracle import refio
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "classify_options.json")))["sets"]
    meta = dbformat.read_meta(os.path.join(C1, "bacteria1.meta"))
    c = dbformat.read_cache(os.path.join(C1, "bacteria1.cache0"))
    tab = O.Table(c.keys, c.sizes, c.values)
    db = Database.__new__(Database)
    db.meta = meta
    lin = Database.target_lineages(db)
    names = meta.target_names()
    single = refio.read_fasta(os.path.join(C1, "single.fa"))
    pf = refio.read_fasta(os.path.join(C1, "pairs.fa"))
    runs = {"single": [(s, b"") for _, s in single],
            "pairs": [(pf[i][1], pf[i + 1][1]) for i in range(0, len(pf), 2)]}
    for name, g in gold.items():
        p = g["params"]
        keys = Database.ranked_lineage_keys(db, p["lowest"]) if p["lowest"] > 0 else None
        seen = 0
        for sec, items in runs.items():
            rows = g["reads"][sec]
            for qid, (a, b) in enumerate(items, start=1):
                top = O.query(tab, a, b, maxc=p["maxc"], tax_of_tgt=keys)[1]
                t, r = O.classify(top, lin, hits_min=p["hits_min"], hits_diff_fraction=p["frac"],
                                  lowest=p["lowest"], highest=p["highest"])
                want = rows.get(str(qid))
                if want is None:
                    assert t == 0, (name, sec, qid, top)                      # -mapped-only: not printed = not classified
                    continue
                seen += 1
                if keys is None:
                    got_top = formatting.format_top_hits(top, names)
                else:                                                         # show_candidates above sequence: taxid:hits
                    got_top = ",".join(f"{int(np.int64(keys[tgt]))}:{hits}" for tgt, hits, _b, _e in top if hits > 0)
                assert got_top == want[0], (name, sec, qid)
                assert formatting.format_classification(t, r, meta.taxa) == want[1], (name, sec, qid, top)
        assert seen == sum(len(v) for v in g["reads"].values()), name
