"""MetaCache on-disk database format (`<db>.meta` + `<db>.cache<N>`), version 20200820.

Host-side loader/writer for the files the probe kernel's table is built from.
Pure numpy; no device code here.

Reference (format definition, nothing copied):
  * `.meta`   writer `src/database.cpp:247-290`, reader `:87-163`
      u64 MC_DB_VERSION (20200820, `src/version.hpp`)
      7 x u8 type widths [feature, target_id, window_id, bucket_size, part_id, taxon_id, num_ranks]
      sketching options 4 x u64 {kmerlen, sketchlen, winlen, winstride}, WRITTEN TWICE
        (`src/hash_dna.hpp:118-147`, `database.cpp:275-277`)
      u64 max_locations_per_feature, u32 target_count, u32 num_parts
      taxonomy: u64 n, then n taxa (`src/taxonomy.hpp:320-342, 700-728`)
        i64 id, i64 parent, u8 rank, str name, str source.filename, u64 source.index,
        u64 source.windows (`taxonomy.hpp:264-278`);  str = u64 len + bytes.
        Target taxa have id = -(tgt)-1.
  * `.cache<N>` writer `src/hash_multimap.hpp:1037-1082`, readers `:970-1030`
      u64 nkeys, u64 nvalues, u64 batch_size (2^20), then per batch of <= batch_size non-empty
      buckets:  u32 keys[b] | u8 sizes[b] | location values[sum(sizes)]
      location = {u32 win, u32 tgt} packed, little endian (`src/database.hpp:136-166`).
"""
from __future__ import annotations

import dataclasses
import io
import struct
from typing import BinaryIO

import numpy as np

MC_DB_VERSION = 20200820
RANK_SEQUENCE = 0
RANK_NONE = 21          # taxonomy::rank::none  (src/taxonomy.hpp:67-90)
NUM_RANKS = 21
DEFAULT_BATCH = 1 << 20  # hash_multimap batch_size default


@dataclasses.dataclass
class Taxon:
    id: int
    parent: int
    rank: int
    name: str
    filename: str = ""
    index: int = 0
    windows: int = 0


@dataclasses.dataclass
class DbMeta:
    kmerlen: int = 16
    sketchlen: int = 16
    winlen: int = 127
    winstride: int = 112
    max_locations_per_feature: int = 254
    target_count: int = 0
    num_parts: int = 1
    taxa: list = dataclasses.field(default_factory=list)
    widths: tuple = (4, 4, 4, 1, 4, 8, NUM_RANKS)
    _source_field_fmt: str = "<QQ"

    def target_names(self) -> list:
        """target id -> sequence name (targets are the taxa with negative ids)."""
        names = [""] * self.target_count
        for t in self.taxa:
            if t.id < 0:
                tgt = -t.id - 1
                if tgt < len(names):
                    names[tgt] = t.name
        return names

    def target_windows(self) -> np.ndarray:
        w = np.zeros(self.target_count, dtype=np.uint32)
        for t in self.taxa:
            if t.id < 0 and -t.id - 1 < self.target_count:
                w[-t.id - 1] = t.windows
        return w


def _rd(f: BinaryIO, fmt: str):
    n = struct.calcsize(fmt)
    b = f.read(n)
    if len(b) != n:
        raise EOFError("truncated database metadata")
    return struct.unpack(fmt, b)


def _rd_str(f: BinaryIO) -> str:
    (n,) = _rd(f, "<Q")
    return f.read(n).decode("latin-1")


def read_meta(path: str) -> DbMeta:
    with open(path, "rb") as f:
        data = f.read()
    last_err = None
    # source.index / source.windows are `std::uint_least64_t` typedefs: 64 bit on the
    # platforms the reference builds on; fall back to narrower layouts just in case.
    for fmt in ("<QQ", "<II", "<QI", "<IQ"):
        try:
            return _parse_meta(io.BytesIO(data), len(data), fmt)
        except (EOFError, ValueError, struct.error, UnicodeDecodeError) as e:  # pragma: no cover
            last_err = e
    raise ValueError(f"cannot parse {path}: {last_err}")


def _parse_meta(f: BinaryIO, size: int, srcfmt: str) -> DbMeta:
    (ver,) = _rd(f, "<Q")
    if ver != MC_DB_VERSION:
        raise ValueError(f"database version {ver} != {MC_DB_VERSION}")
    widths = _rd(f, "<7B")
    if tuple(widths[:5]) != (4, 4, 4, 1, 4):
        raise ValueError(f"unsupported type widths {widths}")
    sk1 = _rd(f, "<4Q")
    sk2 = _rd(f, "<4Q")
    if sk1 != sk2:
        raise ValueError("sketching options mismatch")
    (maxloc,) = _rd(f, "<Q")
    tcount, nparts = _rd(f, "<II")
    (ntax,) = _rd(f, "<Q")
    if ntax > size:
        raise ValueError("implausible taxon count")
    taxa = []
    for _ in range(ntax):
        tid, par = _rd(f, "<qq")
        (rank,) = _rd(f, "<B")
        name = _rd_str(f)
        fname = _rd_str(f)
        idx, wins = _rd(f, srcfmt)
        taxa.append(Taxon(tid, par, rank, name, fname, idx, wins))
    if f.read(1) != b"":
        raise ValueError("trailing bytes in metadata")
    return DbMeta(*sk1, maxloc, tcount, nparts, taxa, tuple(widths), srcfmt)


def write_meta(path: str, meta: DbMeta) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", MC_DB_VERSION))
        f.write(struct.pack("<7B", *meta.widths))
        sk = struct.pack("<4Q", meta.kmerlen, meta.sketchlen, meta.winlen, meta.winstride)
        f.write(sk)
        f.write(sk)
        f.write(struct.pack("<Q", meta.max_locations_per_feature))
        f.write(struct.pack("<II", meta.target_count, meta.num_parts))
        f.write(struct.pack("<Q", len(meta.taxa)))
        # reference order: non-target taxa first, then targets (taxonomy.hpp:719-728);
        # within each group the store's iteration order is kept as given
        ordered = [t for t in meta.taxa if t.id >= 0] + [t for t in meta.taxa if t.id < 0]
        for t in ordered:
            f.write(struct.pack("<qqB", t.id, t.parent, t.rank))
            nb = t.name.encode("latin-1")
            f.write(struct.pack("<Q", len(nb)) + nb)
            fb = t.filename.encode("latin-1")
            f.write(struct.pack("<Q", len(fb)) + fb)
            f.write(struct.pack(meta._source_field_fmt, t.index, t.windows))


def synthetic_meta(target_windows, names=None, **sk) -> DbMeta:
    """Metadata for a taxonomy-less database: every target is a parentless
    sequence-level taxon (what `metacache build` writes without `-taxonomy`)."""
    n = len(target_windows)
    taxa = [Taxon(-(i + 1), 0, RANK_SEQUENCE,
                  names[i] if names is not None else f"T{i}", "synthetic.fa", i,
                  int(target_windows[i])) for i in range(n)]
    return DbMeta(target_count=n, num_parts=1, taxa=taxa, **sk)


@dataclasses.dataclass
class CachePart:
    """One `.cache<N>` file as flat arrays (bucket i = values[offsets[i]:offsets[i+1]])."""
    keys: np.ndarray      # u32 [nkeys]   feature values, file order
    sizes: np.ndarray     # u8  [nkeys]   bucket sizes (1..254)
    values: np.ndarray    # u64 [nvalues] location as (tgt << 32) | win
    batch_size: int = DEFAULT_BATCH

    @property
    def offsets(self) -> np.ndarray:
        off = np.zeros(len(self.sizes) + 1, dtype=np.int64)
        np.cumsum(self.sizes, dtype=np.int64, out=off[1:])
        return off


def iter_cache_batches(path: str):
    """Streams a `.cache<N>` file batch by batch: yields (keys u32, sizes u8, values u64)
    with values packed as (tgt << 32) | win.  First item yielded is the header tuple
    (nkeys, nvalues, batch_size)."""
    with open(path, "rb") as f:
        nkeys, nvalues, batch = struct.unpack("<QQQ", f.read(24))
        yield (nkeys, nvalues, batch)
        done = 0
        while done < nkeys:
            b = int(min(batch, nkeys - done))
            keys = np.fromfile(f, dtype="<u4", count=b)
            sizes = np.fromfile(f, dtype=np.uint8, count=b)
            nv = int(sizes.sum(dtype=np.int64))
            wt = np.fromfile(f, dtype="<u4", count=2 * nv).reshape(nv, 2)
            if len(keys) != b or len(sizes) != b or len(wt) != nv:
                raise EOFError(f"truncated cache file {path}")
            vals = (wt[:, 1].astype(np.uint64) << np.uint64(32)) | wt[:, 0].astype(np.uint64)
            yield (keys, sizes, vals)
            done += b


def read_cache(path: str) -> CachePart:
    it = iter_cache_batches(path)
    nkeys, nvalues, batch = next(it)
    ks, ss, vs = [], [], []
    for k, s, v in it:
        ks.append(k), ss.append(s), vs.append(v)
    keys = np.concatenate(ks) if ks else np.zeros(0, np.uint32)
    sizes = np.concatenate(ss) if ss else np.zeros(0, np.uint8)
    values = np.concatenate(vs) if vs else np.zeros(0, np.uint64)
    if len(keys) != nkeys or len(values) != nvalues:
        raise ValueError(f"{path}: header says {nkeys}/{nvalues}, found {len(keys)}/{len(values)}")
    return CachePart(keys.astype(np.uint32), sizes, values, int(batch))


def write_cache(path: str, part: CachePart) -> None:
    off = part.offsets
    nkeys = len(part.keys)
    with open(path, "wb") as f:
        f.write(struct.pack("<QQQ", nkeys, len(part.values), part.batch_size))
        for b0 in range(0, nkeys, part.batch_size):
            b1 = min(nkeys, b0 + part.batch_size)
            part.keys[b0:b1].astype("<u4").tofile(f)
            part.sizes[b0:b1].astype(np.uint8).tofile(f)
            v = part.values[off[b0]:off[b1]]
            wt = np.empty((len(v), 2), dtype="<u4")
            wt[:, 0] = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            wt[:, 1] = (v >> np.uint64(32)).astype(np.uint32)
            wt.tofile(f)
