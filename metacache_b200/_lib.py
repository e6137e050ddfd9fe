"""ctypes binding of libmcb200.so (the C ABI declared in include/mcb200.h).

The library is the product; this module only loads it.  There is no fallback:
if the shared object is missing, importing the compute API fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcb200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "mcb200.h")
CSRC = os.path.join(HERE, "csrc")


EAGAIN = -7          # MCB200_EAGAIN: a resource was grown, issue the same call again


class Mcb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmcb200 error {code}: {msg}")
        self.code = code


class Sketching(C.Structure):
    """hash_dna.hpp:99-163 sketching_options"""
    _fields_ = [("kmerlen", C.c_uint32), ("sketchlen", C.c_uint32),
                ("winlen", C.c_uint32), ("winstride", C.c_uint32)]


class Candidate(C.Structure):
    """candidate_structs.hpp:80-104 match_candidate (without the host `tax` pointer)"""
    _fields_ = [("tgt", C.c_uint32), ("hits", C.c_uint32), ("beg", C.c_uint32), ("end", C.c_uint32)]


class Classification(C.Structure):
    _fields_ = [("taxon", C.c_uint32), ("rank", C.c_uint32)]


class ShardRun(C.Structure):
    """mcb200_shard_run: what one owner shard returned for this rank's features"""
    _fields_ = [("locations", C.c_void_p), ("offsets", C.c_void_p), ("n_features", C.c_uint32),
                ("n_locations", C.c_uint32)]


class DevQueries(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("seq_offsets", C.c_void_p), ("seq_query", C.c_void_p),
                ("max_win", C.c_void_p), ("n_seqs", C.c_uint32), ("n_queries", C.c_uint32),
                ("n_bases", C.c_uint64)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libmcb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-j8"]
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    out = subprocess.run(args, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libmcb200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def declared_symbols(header: str = HEADER):
    """Every function name include/mcb200.h declares (used by the symbol test)."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb200_[a-z0-9_]+)\s*\(", txt)))


_P = C.c_void_p
_SIGS = {
    "mcb200_abi_version": (C.c_int, []),
    "mcb200_last_error": (C.c_char_p, []),
    "mcb200_device_count": (C.c_int, []),
    "mcb200_reader_open": (_P, [C.c_char_p, C.c_char_p]),
    "mcb200_reader_open_range": (_P, [C.c_char_p, C.c_uint64, C.c_uint64]),
    "mcb200_reader_close": (None, [_P]),
    "mcb200_reader_index": (C.c_uint64, [_P]),
    "mcb200_reader_next": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "mcb200_reader_skip": (C.c_int64, [_P, C.c_uint64, _P]),
    "mcb200_reader_fill_batch": (C.c_int64, [_P, _P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, _P, C.c_uint64, _P]),
    "mcb200_db_open": (_P, [C.c_int, C.c_uint32]),
    "mcb200_db_open_multi": (_P, [C.c_uint32, C.POINTER(C.c_int)]),
    "mcb200_db_part_device": (C.c_int, [_P, C.c_uint32]),
    "mcb200_db_close": (None, [_P]),
    "mcb200_db_part_begin": (C.c_int, [_P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_float]),
    "mcb200_db_part_append": (C.c_int, [_P, C.c_uint32, _P, _P, _P, C.c_uint64, C.c_uint64]),
    "mcb200_db_part_append_device": (C.c_int, [_P, C.c_uint32, _P, _P, _P, C.c_uint64, C.c_uint64]),
    "mcb200_db_part_finish": (C.c_int, [_P, C.c_uint32]),
    "mcb200_db_load_cache_file": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_float]),
    "mcb200_db_set_target_taxa": (C.c_int, [_P, _P, C.c_uint32]),
    "mcb200_db_set_target_lineages": (C.c_int, [_P, _P, C.c_uint32]),
    "mcb200_classify_device": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, _P, _P]),
    "mcb200_batch_enable_classification": (C.c_int, [_P, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32]),
    "mcb200_batch_classifications": (C.POINTER(Classification), [_P, C.c_uint32]),
    "mcb200_db_part_count": (C.c_uint32, [_P]),
    "mcb200_db_key_count": (C.c_uint64, [_P, C.c_uint32]),
    "mcb200_db_value_count": (C.c_uint64, [_P, C.c_uint32]),
    "mcb200_db_bucket_count": (C.c_uint64, [_P, C.c_uint32]),
    "mcb200_db_device_bytes": (C.c_uint64, [_P, C.c_uint32]),
    "mcb200_db_device": (C.c_int, [_P]),
    "mcb200_max_supported_locations_per_feature": (C.c_uint32, []),
    "mcb200_db_build_part_from_targets": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint32, C.c_uint32,
                                                    C.POINTER(Sketching), C.c_uint32, C.c_float, _P]),
    "mcb200_db_part_export": (C.c_int, [_P, C.c_uint32, _P, _P, _P]),
    "mcb200_batch_create": (_P, [_P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32]),
    "mcb200_batch_destroy": (None, [_P]),
    "mcb200_batch_add_read": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64,
                                        C.c_uint32]),
    "mcb200_batch_add_reads": (C.c_int64, [_P, C.c_uint32, _P, _P, C.c_uint32, C.c_int, C.c_uint64,
                                           C.c_uint32]),
    "mcb200_batch_submit": (C.c_int, [_P, C.c_uint32, C.POINTER(Sketching)]),
    "mcb200_batch_wait": (C.c_int, [_P, C.c_uint32]),
    "mcb200_batch_clear": (C.c_int, [_P, C.c_uint32]),
    "mcb200_batch_num_queries": (C.c_uint32, [_P, C.c_uint32]),
    "mcb200_batch_num_windows": (C.c_uint32, [_P, C.c_uint32]),
    "mcb200_batch_top_candidates": (C.POINTER(Candidate), [_P, C.c_uint32, C.c_uint32]),
    "mcb200_batch_allhits": (C.POINTER(C.c_uint64), [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "mcb200_batch_sketch": (C.POINTER(C.c_uint32), [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "mcb200_batch_query_window_offset": (C.c_uint32, [_P, C.c_uint32, C.c_uint32]),
    "mcb200_batch_span_ms": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "mcb200_batch_last_timing": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mcb200_workspace_create": (_P, [_P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int]),
    "mcb200_workspace_destroy": (None, [_P]),
    "mcb200_sketch_device": (C.c_int, [_P, C.POINTER(DevQueries), C.POINTER(Sketching), _P]),
    "mcb200_query_part_device": (C.c_int, [_P, C.c_uint32, _P, _P]),
    "mcb200_query_sketches_device": (C.c_int, [_P, C.c_uint32, _P, _P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    "mcb200_merge_candidates_device": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    "mcb200_query_device": (C.c_int, [_P, C.POINTER(DevQueries), C.POINTER(Sketching), _P, _P]),
    "mcb200_db_shard_begin": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "mcb200_db_shard_maxima": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mcb200_db_shard_finish": (C.c_int, [_P, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32]),
    "mcb200_db_location_bytes": (C.c_uint32, [_P, C.c_uint32]),
    "mcb200_shard_route_device": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P]),
    "mcb200_shard_probe_device": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, _P, _P, _P]),
    "mcb200_shard_gather_device": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64, _P, _P]),
    "mcb200_shard_reduce_device": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.POINTER(ShardRun), _P, C.c_uint32, _P, _P]),
    "mcb200_pack_bases": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P, _P]),
    "mcb200_sketch_packed_device": (C.c_int, [_P, C.POINTER(DevQueries), _P, _P, C.POINTER(Sketching), _P]),
    "mcb200_query_packed_device": (C.c_int, [_P, C.POINTER(DevQueries), _P, _P, C.POINTER(Sketching), _P, _P]),
    "mcb200_workspace_num_windows": (C.c_uint32, [_P]),
    "mcb200_workspace_sketches": (_P, [_P]),
    "mcb200_workspace_query_windows": (_P, [_P]),
    "mcb200_workspace_allhits": (_P, [_P]),
    "mcb200_workspace_allhits_offsets": (_P, [_P]),
    "mcb200_workspace_counters": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "mcb200_workspace_check": (C.c_int, [_P]),
    "mcb200_workspace_set_profiling": (C.c_int, [_P, C.c_int]),
    "mcb200_workspace_stage_times": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "mcb200_workspace_set_warp_capacity": (C.c_int, [_P, C.c_uint32]),
    "mcb200_kernel_launches": (C.c_uint64, []),
}

_lib = None


def lib():
    """The loaded library.  Raises if libmcb200.so has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the query path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise Mcb200Error(rc, lib().mcb200_last_error().decode(errors="replace"))
    return rc


def check_ptr(p):
    if not p:
        raise Mcb200Error(None, lib().mcb200_last_error().decode(errors="replace"))
    return p


def query_device_checked(ws, q, sk, d_top_ptr, stream=None, max_attempts=6):
    """mcb200_query_device followed by mcb200_workspace_check, re-issued while the library answers
    MCB200_EAGAIN (a read outgrew its scratch region and the pool was grown).  Synchronises the stream.
    Returns the number of attempts."""
    L = lib()
    for attempt in range(1, max_attempts + 1):
        check(L.mcb200_query_device(ws, C.byref(q), C.byref(sk), d_top_ptr, stream))
        rc = L.mcb200_workspace_check(ws)
        if rc == EAGAIN:
            continue
        check(rc)
        return attempt
    raise Mcb200Error(EAGAIN, "scratch pool still too small after %d attempts" % max_attempts)
