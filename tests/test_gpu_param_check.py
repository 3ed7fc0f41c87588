"""The ray kernel's parametric node expansion cross-checked, node by node, against the slab
expansion (the reference's eight per-child tests + stable key order) by a library built with
-DVRT_PARAM_CHECK.  The check build is a test artefact (libvrt_check.so); it is produced with
nvcc on demand and never used by the product path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parametric_expansion_equals_slab_expansion(gpu):
    env = dict(os.environ, VRT_LIB_SUFFIX="_check", VRT_EXTRA_NVCC="-DVRT_PARAM_CHECK")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "param_check.py"), "8"], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["checked"] > 10_000_000, res
    assert res["mismatches"] == 0, res
