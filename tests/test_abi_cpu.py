"""CPU-side checks of the drop-in boundary: libpms.so loads without a GPU, exports every symbol declared in
include/percept_smoother_c.h, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(pms_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import percept_b200 as pb
    for hdr in ("percept_smoother_c.h", "percept_fixtures_c.h"):
        names = declared(hdr)
        assert len(names) > 1
        for n in names:
            assert hasattr(pb.lib, n), f"{n} declared in {hdr} but not exported by libpms.so"


def test_library_contains_sm100a_kernels_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "percept_b200", "lib", "libpms.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    import percept_b200 as pb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pb.SmootherError) as ei:
        pb.create(0)
    assert "no CPU fallback" in str(ei.value)


def test_state_struct_layout_matches_header():
    from percept_b200.capi import State
    # 11 doubles, 4 ints, 6 uint64 (include/percept_smoother_c.h: pms_state)
    assert C.sizeof(State) == 11 * 8 + 4 * 4 + 6 * 8


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "percept_b200")):
        if "lib" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in txt and "orc_" not in txt and "oracle/" not in txt, f


def test_reference_arm_of_the_bench_runs_without_a_gpu():
    """`bench.py --impl reference` times the CPU restatement on the host cores and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "Melem/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
