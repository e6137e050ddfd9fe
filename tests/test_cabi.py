"""The C-ABI library loads and exports everything include/mcb200.h declares (no GPU needed)."""
import ctypes as C
import os
import subprocess

import pytest

from metacache_b200 import _lib
from tests.conftest import _has_gpu


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    assert L.mcb200_abi_version() == 1
    assert L.mcb200_max_supported_locations_per_feature() == 254


def test_exports_every_declared_symbol():
    declared = _lib.declared_symbols()
    assert len(declared) >= 45
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # and the python binding covers the same set
    assert sorted(_lib._SIGS) == declared


def test_only_c_symbols_cross_the_boundary():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in exported if s.startswith("mcb200_")]
    assert set(_lib.declared_symbols()) <= set(ours)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device error path")
def test_fails_loudly_without_a_gpu():
    L = _lib.lib()
    assert L.mcb200_device_count() == 0
    assert not L.mcb200_db_open(0, 1)
    msg = L.mcb200_last_error().decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    from metacache_b200.database import Database
    with pytest.raises(_lib.Mcb200Error):
        Database(0, 1)
    # a store over several devices needs them too; argument errors come first
    with pytest.raises(_lib.Mcb200Error):
        Database(devices=[0, 1])
    assert not L.mcb200_db_open_multi(0, None)
    assert "n_parts" in L.mcb200_last_error().decode()
    # null handles are errors, not crashes, on every new entry point
    assert L.mcb200_db_shard_begin(None, 0, 0, 2, 0) < 0
    assert L.mcb200_db_shard_finish(None, 0, C.c_float(0), 0, 0) < 0
    assert L.mcb200_shard_route_device(None, None, None, 0, 16, 2, None, None, None) < 0
    assert L.mcb200_shard_probe_device(None, 0, None, 0, None, None, None) < 0
    assert L.mcb200_shard_reduce_device(None, 0, 2, None, None, None, 0, None, None) < 0
    assert L.mcb200_query_packed_device(None, None, None, None, None, None, None) < 0
    assert L.mcb200_db_location_bytes(None, 0) == 0 and L.mcb200_db_part_device(None, 0) == -1


def test_product_path_never_imports_the_oracle():
    root = os.path.dirname(_lib.HERE)
    for dirpath, _, files in os.walk(os.path.join(root, "metacache_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mc_oracle" not in txt.replace(
                    "oracle/mc_oracle.c:top_insert", ""), f


def test_loading_the_library_defaults_the_stream_channels_without_overriding_the_host():
    """the batch slots run on one stream each; the library raises CUDA_DEVICE_MAX_CONNECTIONS (8 by default,
    streams sharing a channel serialise) when it is loaded, unless the host program set it (api.cu)"""
    import sys
    code = ("import ctypes\nfrom metacache_b200 import _lib\n_lib.lib()\n"
            "g = ctypes.CDLL(None).getenv\ng.restype = ctypes.c_char_p\n"
            "print(g(b'CUDA_DEVICE_MAX_CONNECTIONS').decode())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for preset, want in ((None, "32"), ("4", "4")):
        env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}
        if preset:
            env["CUDA_DEVICE_MAX_CONNECTIONS"] = preset
        out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, check=True)
        assert out.stdout.strip() == want
