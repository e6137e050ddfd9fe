"""Synthetic benchmark workloads of BASELINE.md section 3 (DB-S / R150 / RLONG), generated
directly in HBM with torch (plumbing) - or with numpy for small test-sized instances.
Counter-based (splitmix64 of a position key), so every element is reproducible on its own
and the torch and numpy generators give identical bytes.

  DB-S   n_targets x target_len bases; families of `family` members, each member = family
         root with ~1 % substitutions                                   (seed 20260121)
  R150   n reads x 150 bp: 90 % sampled uniformly (target, offset, strand) with ~1 %
         substitutions and ~0.1 % N, 10 % iid random                     (seed 20260122)
  RLONG  lengths round(exp(N(ln 480, 1))) clipped to [200, 19000], 5 % substitutions
                                                                         (seed 20260123)
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_C0, _C1, _C2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
SEED_DB, SEED_R150, SEED_RLONG = 20260121, 20260122, 20260123
SUB_1PCT = 655        # of 65536
SUB_5PCT = 3277
N_01PCT = 66


def _mix_seed(seed: int, stream: int) -> int:
    """host-side splitmix64 of (seed, stream): well separated key spaces for nearby seeds"""
    z = (seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03 + _C0) & _M64
    z = ((z ^ (z >> 30)) * _C1) & _M64
    z = ((z ^ (z >> 27)) * _C2) & _M64
    return z ^ (z >> 31)


def _s64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


# ------------------------------------------------------------------ numpy backend
def _sm_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(_C0)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        return z ^ (z >> np.uint64(31))


# ------------------------------------------------------------------ torch backend
def _sm_t(x):
    import torch
    z = x + _s64(_C0)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(_C1)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(_C2)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


class _NP:
    @staticmethod
    def arange(n, device=None):
        return np.arange(n, dtype=np.uint64)

    sm = staticmethod(_sm_np)

    @staticmethod
    def u(x):
        return np.uint64(x & _M64)

    @staticmethod
    def shr(x, n):
        return x >> np.uint64(n)

    @staticmethod
    def to_u8(x):
        return x.astype(np.uint8)

    @staticmethod
    def mod(x, m):
        return x % np.uint64(m)

    where = staticmethod(np.where)


class _T:
    @staticmethod
    def arange(n, device=None):
        import torch
        return torch.arange(n, dtype=torch.int64, device=device)

    sm = staticmethod(_sm_t)

    @staticmethod
    def u(x):
        return _s64(x)

    @staticmethod
    def shr(x, n):
        return (x >> n) & ((1 << (64 - n)) - 1)

    @staticmethod
    def to_u8(x):
        import torch
        return x.to(torch.uint8)

    @staticmethod
    def mod(x, m):
        # x is a non-negative int64 here (callers shift first)
        return x % m

    @staticmethod
    def where(c, a, b):
        import torch
        return torch.where(c, a, b)


_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate(B, codes, h, sub_thresh):
    """codes 0..3; substitution when low 16 bits of h < sub_thresh: a guaranteed different base"""
    sub = (h & 0xFFFF) < sub_thresh if B is _T else (h & np.uint64(0xFFFF)) < np.uint64(sub_thresh)
    delta = B.mod(B.shr(h, 16) & (0xFFFF if B is _T else np.uint64(0xFFFF)), 3) + (1 if B is _T else np.uint64(1))
    new = (codes + delta) & (3 if B is _T else np.uint64(3))
    return B.where(sub, new, codes)


def target_codes(B, t0: int, t1: int, target_len: int, family: int, seed: int, device=None):
    """2-bit codes (as int64/uint64 0..3) of targets [t0, t1), shape [(t1-t0), target_len]"""
    n = t1 - t0
    pos = B.arange(target_len, device)
    tid = B.arange(n, device) + B.u(t0)
    fam = (tid // family) if B is _T else (tid // np.uint64(family))
    key_root = (fam[:, None] << (32 if B is _T else np.uint64(32))) + pos[None, :] + B.u(_mix_seed(seed, 1))
    root = B.sm(key_root) & (3 if B is _T else np.uint64(3))
    key_mut = (tid[:, None] << (32 if B is _T else np.uint64(32))) + pos[None, :] + B.u(_mix_seed(seed, 2))
    return _mutate(B, root, B.sm(key_mut), SUB_1PCT)


def make_targets(n_targets: int, target_len: int, family: int = 10, seed: int = SEED_DB,
                 device=None, chunk: int = 512):
    """ASCII bases of all targets, concatenated ([n_targets*target_len] uint8), + u64 offsets.
    device=None -> numpy; otherwise a torch device."""
    if device is None:
        c = target_codes(_NP, 0, n_targets, target_len, family, seed)
        bases = _ASCII[c.astype(np.int64)].reshape(-1)
        return bases, np.arange(n_targets + 1, dtype=np.uint64) * np.uint64(target_len)
    import torch
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    bases = torch.empty(n_targets * target_len, dtype=torch.uint8, device=device)
    for t0 in range(0, n_targets, chunk):
        t1 = min(n_targets, t0 + chunk)
        c = target_codes(_T, t0, t1, target_len, family, seed, device)
        bases[t0 * target_len:t1 * target_len] = lut[c].reshape(-1)
    off = torch.arange(n_targets + 1, dtype=torch.int64, device=device) * target_len
    return bases, off


def _reads_chunk(B, lut, comp, targets, n_targets, target_len, i0, i1, read_len, seed, sub_thresh,
                 n_thresh, frac_random_65536, device=None, pos_bits=12):
    n = i1 - i0
    idx = B.arange(n, device) + B.u(i0)
    h = B.sm(idx + B.u(_mix_seed(seed, 3)))
    low16 = h & (0xFFFF if B is _T else np.uint64(0xFFFF))
    is_random = low16 < (frac_random_65536 if B is _T else np.uint64(frac_random_65536))
    strand = B.shr(h, 16) & (1 if B is _T else np.uint64(1))
    h1 = B.sm(idx + B.u(_mix_seed(seed, 4)))
    tgt = B.mod(B.shr(h1, 1), n_targets)
    off = B.mod(B.shr(B.sm(h1), 1), target_len - read_len + 1)
    j = B.arange(read_len, device)
    # position in the target: forward j, reverse (read_len-1-j)
    fwd = strand[:, None] == 0
    rel = B.where(fwd, j[None, :] + 0 * off[:, None], (read_len - 1) - j[None, :] + 0 * off[:, None])
    src = tgt[:, None] * (target_len if B is _T else np.uint64(target_len)) + off[:, None] + rel
    if B is _T:
        raw = targets[src]
        base = B.where(fwd, raw, comp[raw.long()])
        codes = lut[base.long()].long()
    else:
        raw = targets[src.astype(np.int64)]
        base = np.where(fwd, raw, comp[raw])
        codes = lut[base].astype(np.uint64)
    hb = B.sm((idx[:, None] << (pos_bits if B is _T else np.uint64(pos_bits))) + j[None, :] + B.u(_mix_seed(seed, 5)))
    rnd = B.shr(hb, 48) & (3 if B is _T else np.uint64(3))
    codes = B.where(is_random[:, None], rnd, _mutate(B, codes, hb, sub_thresh))
    isn = (B.shr(hb, 32) & (0xFFFF if B is _T else np.uint64(0xFFFF))) < (n_thresh if B is _T else np.uint64(n_thresh))
    return codes, isn


def make_reads_150(n_reads: int, targets, n_targets: int, target_len: int, read_len: int = 150,
                   seed: int = SEED_R150, sub_thresh: int = SUB_1PCT, n_thresh: int = N_01PCT,
                   frac_random: float = 0.10, device=None, chunk: int = 1 << 19):
    """-> uint8 [n_reads, read_len] ASCII reads (R150 recipe)."""
    fr = int(round(frac_random * 65536))
    if device is None:
        lut = np.zeros(256, np.uint8)
        lut[_ASCII] = np.arange(4, dtype=np.uint8)
        comp = np.arange(256, dtype=np.uint8)
        comp[_ASCII] = _ASCII[::-1]
        codes, isn = _reads_chunk(_NP, lut, comp, targets, n_targets, target_len, 0, n_reads, read_len,
                                  seed, sub_thresh, n_thresh, fr)
        out = _ASCII[codes.astype(np.int64)]
        out[isn] = ord("N")
        return out
    import torch
    lut = torch.zeros(256, dtype=torch.uint8, device=device)
    asc = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    lut[asc.long()] = torch.arange(4, dtype=torch.uint8, device=device)
    comp = torch.arange(256, dtype=torch.uint8, device=device)
    comp[asc.long()] = asc.flip(0)
    out = torch.empty((n_reads, read_len), dtype=torch.uint8, device=device)
    for i0 in range(0, n_reads, chunk):
        i1 = min(n_reads, i0 + chunk)
        codes, isn = _reads_chunk(_T, lut, comp, targets, n_targets, target_len, i0, i1, read_len, seed,
                                  sub_thresh, n_thresh, fr, device)
        o = asc[codes]
        o[isn] = ord("N")
        out[i0:i1] = o
    return out


def make_long_reads(n_reads: int, targets, n_targets: int, target_len: int, seed: int = SEED_RLONG,
                    sub_thresh: int = SUB_5PCT, device=None, max_len: int = 19000):
    """RLONG recipe: reads of long_read_lengths(n_reads, seed) bases sampled from the targets with ~5 %
    substitutions (no random reads, ~0.1 % N).  -> (uint8 bases back to back, int64 offsets [n_reads+1]).
    Reads are generated per length class (cap = 256, 512, ... bases) at the cap length and cut to their
    own length, so a read is reproducible from (seed, index, length).  torch device or numpy (None)."""
    lens = np.minimum(long_read_lengths(n_reads, seed), min(max_len, target_len))
    offs = np.zeros(n_reads + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    caps = 1 << np.ceil(np.log2(np.maximum(lens, 256))).astype(np.int64)
    caps = np.minimum(caps, min(max_len, target_len))
    if device is None:
        out = np.empty(int(offs[-1]), np.uint8)
        lut = np.zeros(256, np.uint8)
        lut[_ASCII] = np.arange(4, dtype=np.uint8)
        comp = np.arange(256, dtype=np.uint8)
        comp[_ASCII] = _ASCII[::-1]
    else:
        import torch
        out = torch.empty(int(offs[-1]), dtype=torch.uint8, device=device)
        lut = torch.zeros(256, dtype=torch.uint8, device=device)
        asc = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
        lut[asc.long()] = torch.arange(4, dtype=torch.uint8, device=device)
        comp = torch.arange(256, dtype=torch.uint8, device=device)
        comp[asc.long()] = asc.flip(0)
    for cap in (int(c) for c in np.unique(caps)):
        ids = np.nonzero(caps == cap)[0]
        grp_size = max(1, min(4096, (1 << 24) // cap))
        for b in range(0, len(ids), grp_size):
            grp = ids[b:b + grp_size]
            if device is None:
                for i in grp:
                    c, isn = _reads_chunk(_NP, lut, comp, targets, n_targets, target_len, int(i), int(i) + 1, cap,
                                          seed, sub_thresh, N_01PCT, 0, None, 15)
                    r = _ASCII[c.astype(np.int64)]
                    r[isn] = ord("N")
                    out[offs[i]:offs[i + 1]] = r[0, :lens[i]]
            else:
                idx = torch.as_tensor(grp, dtype=torch.int64, device=device)
                c, isn = _reads_rows(_T, lut, comp, targets, n_targets, target_len, idx, cap, seed, sub_thresh,
                                     N_01PCT, device, 15)
                o = asc[c]
                o[isn] = ord("N")
                ln = torch.as_tensor(lens[grp], device=device)
                keep = torch.arange(cap, device=device)[None, :] < ln[:, None]
                out[_ragged_index(torch.as_tensor(offs[grp], device=device), cap, keep)] = o[keep]
    return out, (offs if device is None else torch.as_tensor(offs, device=device))


def _ragged_index(starts, cap, keep):
    import torch
    pos = starts[:, None] + torch.arange(cap, device=starts.device)[None, :]
    return pos[keep]


def _reads_rows(B, lut, comp, targets, n_targets, target_len, idx, read_len, seed, sub_thresh, n_thresh,
                device, pos_bits):
    """_reads_chunk for an arbitrary index vector (torch), no random reads"""
    h = B.sm(idx + B.u(_mix_seed(seed, 3)))
    strand = B.shr(h, 16) & 1
    h1 = B.sm(idx + B.u(_mix_seed(seed, 4)))
    tgt = B.mod(B.shr(h1, 1), n_targets)
    off = B.mod(B.shr(B.sm(h1), 1), target_len - read_len + 1)
    j = B.arange(read_len, device)
    fwd = strand[:, None] == 0
    rel = B.where(fwd, j[None, :] + 0 * off[:, None], (read_len - 1) - j[None, :] + 0 * off[:, None])
    src = tgt[:, None] * target_len + off[:, None] + rel
    raw = targets[src]
    base = B.where(fwd, raw, comp[raw.long()])
    codes = lut[base.long()].long()
    hb = B.sm((idx[:, None] << pos_bits) + j[None, :] + B.u(_mix_seed(seed, 5)))
    codes = _mutate(B, codes, hb, sub_thresh)
    isn = (B.shr(hb, 32) & 0xFFFF) < n_thresh
    return codes, isn


def long_read_lengths(n_reads: int, seed: int = SEED_RLONG) -> np.ndarray:
    """round(exp(N(ln 480, 1.0))) clipped to [200, 19000] (numpy Generator PCG64(seed))"""
    rng = np.random.Generator(np.random.PCG64(seed))
    ln = np.rint(np.exp(rng.normal(np.log(480.0), 1.0, n_reads)))
    return np.clip(ln, 200, 19000).astype(np.int64)
