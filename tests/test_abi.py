"""CPU: the C-ABI library loads, exports every symbol include/sslam_b200.h declares, and refuses
to compute without a GPU (no CPU fallback)."""

import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sslam_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from sslam_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib


DEBUG_HEADER = os.path.join(ROOT, "include", "sslam_b200_debug.h")


def declared_symbols(header=HEADER):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sslam_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 12
    handle = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert set(names) == set(lib.SIGNATURES), "ctypes signature table out of sync with the header"
    dbg = declared_symbols(DEBUG_HEADER)
    assert set(dbg) == set(lib.DEBUG_SIGNATURES)
    for n in dbg:
        assert hasattr(handle, n), f"{n} declared in the debug header but not exported"
    # nothing is exported that no header declares
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("sslam_")}
    assert exported == set(names) | set(dbg), exported ^ (set(names) | set(dbg))


def test_abi_version_and_workspace_queries(lib):
    l = lib.load()
    assert l.sslam_abi_version() == 4
    assert l.sslam_decode_workspace_bytes(2, 480, 640, 2048) >= 2 * 480 * 640 * 8
    assert l.sslam_match_workspace_bytes(3, 3, 3, 2048, 2048, 256, 0) >= 3 * 2048 * 8
    assert l.sslam_decode_workspace_bytes(0, 480, 640, 1) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    l = lib.load()
    assert l.sslam_device_check() == -5                       # SSLAM_ENODEVICE
    assert "no CPU fallback" in lib.last_error()
    from sslam_b200 import ops
    with pytest.raises(RuntimeError):
        ops.decode_topk(torch.rand(1, 8, 8), 4)
    with pytest.raises(RuntimeError):
        ops.gather_bilinear(torch.rand(1, 4, 4, 8), torch.zeros(1, 2, 2))
    from sslam_b200 import matchers
    import numpy as np
    with pytest.raises(RuntimeError):
        matchers.find_matches(np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "semantic-slam-master_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), \
                    f"{f} imports the oracle"
