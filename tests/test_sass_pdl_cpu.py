"""SASS guard (no GPU needed): in every kernel that executes griddepcontrol.wait, the global loads placed ahead of the wait
must stay within the reviewed allowance (tools/sass_pdl_scan.py).  Round 2: nvcc hoisted four loads of predecessor-written
`qkv` above the wait in the decode attention kernel (`const __restrict__`), which made tensor-parallel decode
non-deterministic; this test fails if a rebuild brings loads back in front of the wait."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_new_global_loads_ahead_of_the_dependency_wait(built_lib):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_pdl_scan.py"), "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert "0 over their allowance" in r.stdout
