"""Host-side 2-bit packing (csrc/pack.cpp, mcb200_pack_bases) against a numpy restatement of the
device layout encode_kernel writes (dna_encoding.hpp:38-62: A0 C1 G2 T3, U = T, lower case folded,
everything else ambiguous).  CPU only: no compute call touches a device."""
import ctypes as C

import numpy as np
import pytest

from metacache_b200 import _lib


def pack_numpy(flat: np.ndarray):
    """-> (codes u32[2 * units], amb u32[units]) for bases numbered from 0; tail bits zero"""
    n = len(flat)
    units = (n + 31) // 32
    up = flat & 0xDF
    code = np.zeros(n, np.uint64)
    ok = np.zeros(n, bool)
    for ch, c in ((ord("A"), 0), (ord("C"), 1), (ord("G"), 2), (ord("T"), 3), (ord("U"), 3)):
        m = up == ch
        code[m] = c
        ok |= m
    codes = np.zeros(units * 2, np.uint32)
    amb = np.zeros(units, np.uint32)
    i = np.arange(n)
    np.bitwise_or.at(codes, i // 16, (code << (30 - 2 * (i % 16)).astype(np.uint64)).astype(np.uint32))
    np.bitwise_or.at(amb, i // 32, ((~ok).astype(np.uint32) << (31 - (i % 32)).astype(np.uint32)))
    return codes, amb


def pack_lib(chunks, force_scalar):
    """force_scalar: 0 = best path of this CPU (AVX-512 or AVX2), 1 = scalar, 2 = at most AVX2"""
    L = _lib.lib()
    fn = L.mcb200_internal_pack_append
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    total = sum(len(c) for c in chunks)
    units = (total + 31) // 32 + 1
    codes = np.full(units * 2, 0xDEADBEEF, np.uint32)       # garbage: the packer must not rely on cleared memory
    amb = np.full(units, 0xDEADBEEF, np.uint32)
    pos = 0
    for c in chunks:
        buf = np.ascontiguousarray(c)
        fn(buf.ctypes.data if len(buf) else None, len(buf), pos, codes.ctypes.data, amb.ctypes.data, force_scalar)
        pos += len(buf)
    used = (total + 31) // 32
    return codes[:used * 2], amb[:used]


ALPHABET = np.frombuffer(b"ACGTacgtUuNnRYKM-*. \x00\xff", np.uint8)


@pytest.mark.parametrize("force_scalar", [1, 2, 0])
def test_appends_at_every_offset_match_the_layout(force_scalar):
    rng = np.random.default_rng(5)
    for trial in range(60):
        lens = rng.integers(0, 200, size=rng.integers(1, 12))
        if trial % 7 == 0:
            lens[rng.integers(0, len(lens))] = 0                  # empty reads / mates
        chunks = [ALPHABET[rng.integers(0, len(ALPHABET) if trial % 2 else 8, size=n)] for n in lens]
        flat = np.concatenate(chunks) if len(chunks) else np.zeros(0, np.uint8)
        codes, amb = pack_lib(chunks, force_scalar)
        rc, ra = pack_numpy(flat)
        assert np.array_equal(codes, rc), (trial, lens)
        assert np.array_equal(amb, ra), (trial, lens)


@pytest.mark.parametrize("force_scalar", [1, 2, 0])
def test_long_appends_of_arbitrary_bytes_at_aligned_and_odd_positions(force_scalar):
    """the bulk path (whole 64-base groups stored directly when the append starts on a word) and the
    shifted path, over every byte value incl. >= 0x80, with every tail length"""
    rng = np.random.default_rng(23)
    for trial, (first, n) in enumerate([(0, 70000), (32, 4096 + 63), (64, 129), (5, 65536 + 17), (31, 1000), (33, 64),
                                        (0, 64), (0, 63), (0, 65), (96, 127), (7, 191)]):
        body = rng.integers(0, 256, size=n).astype(np.uint8)
        mostly = ALPHABET[rng.integers(0, 8, size=n)]
        body = np.where(rng.random(n) < 0.9, mostly, body).astype(np.uint8)
        chunks = [ALPHABET[rng.integers(0, 8, size=first)], body, ALPHABET[rng.integers(0, 8, size=trial)]]
        codes, amb = pack_lib(chunks, force_scalar)
        rc, ra = pack_numpy(np.concatenate(chunks))
        assert np.array_equal(codes, rc), (first, n)
        assert np.array_equal(amb, ra), (first, n)


def test_one_bulk_append_equals_read_by_read():
    rng = np.random.default_rng(11)
    chunks = [ALPHABET[rng.integers(0, 10, size=150)] for _ in range(500)]
    a = pack_lib(chunks, 0)
    b = pack_lib([np.concatenate(chunks)], 0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_public_entry_point_reports_the_path_taken():
    L = _lib.lib()
    s = np.frombuffer(b"ACGTN" * 20, np.uint8)
    codes = np.zeros(2 * 5, np.uint32)
    amb = np.zeros(5, np.uint32)
    rc = L.mcb200_pack_bases(s.ctypes.data, len(s), 0, codes.ctypes.data, amb.ctypes.data)
    assert rc in (0, 1, 2)                               # scalar / AVX2 / AVX-512 available on this host
    rcodes, ramb = pack_numpy(s)
    assert np.array_equal(codes[:len(rcodes)], rcodes) and np.array_equal(amb[:len(ramb)], ramb)
    assert L.mcb200_pack_bases(None, 4, 0, codes.ctypes.data, amb.ctypes.data) < 0
