"""CPU-side checks of the drop-in boundary: libofb.so loads and exports every symbol that
include/*.h declares; no compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(ofb_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_header_declares_the_path():
    syms = _declared_symbols()
    for need in ("ofb_create", "ofb_destroy", "ofb_reset", "ofb_step", "ofb_step_host", "ofb_raster", "ofb_obs_vec",
                 "ofb_bot_actions", "ofb_state_export", "ofb_state_import", "ofb_last_error"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from ofighters_b200 import build, _lib
    build.build()
    lib = C.CDLL(build.LIB)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert missing == []
    assert _lib.load().ofb_abi_version() == 1


def test_default_config_matches_reference_constants():
    from ofighters_b200 import _lib, ArenaConfig
    c = _lib.OfbConfig()
    _lib.load().ofb_default_config(C.byref(c))
    d = ArenaConfig()
    assert (c.n_ships, c.width, c.height, c.max_time) == (d.n_ships, d.width, d.height, d.max_time) == (7, 400, 400, 200)
    assert (c.reward_kill, c.reward_death, c.reward_aim, c.reward_trajectory) == (0, 0, 2, 1)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ofighters_b200 import BatchedBattleground, OfbError
    with pytest.raises(OfbError, match="no CPU fallback"):
        BatchedBattleground(4, ships=7)
    # and the C ABI itself refuses too
    from ofighters_b200 import _lib
    lib = _lib.load()
    c = _lib.OfbConfig()
    lib.ofb_default_config(C.byref(c))
    h = C.c_void_p()
    dummy = (C.c_int32 * 56)()
    rc = lib.ofb_create(C.byref(c), 4, 0, C.cast(dummy, C.c_void_p), None, C.byref(h))
    assert rc < 0 and b"no CUDA device" in lib.ofb_last_error()


def test_product_does_not_import_oracle():
    for path in glob.glob(os.path.join(ROOT, "ofighters_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5",
                          "--warmup", "3"], capture_output=True, text=True, timeout=300, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
