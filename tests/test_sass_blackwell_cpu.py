"""SASS evidence (no GPU needed): the kernels DESIGN.md §4 describes as tcgen05 / TMEM / TMA kernels really contain the
Blackwell instructions — UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor), UTCBAR (tcgen05.commit) —
and the programmatic-launch pair PREEXIT / ACQBULK (griddepcontrol.launch_dependents / .wait)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPECT = {
    # object -> {mnemonic regex: minimum occurrences}
    "gemm_tc.o": {r"UTCHMMA\b": 100, r"UTMALDG\.2D\b": 100, r"LDTM\.x16": 10, r"UTCBAR\b": 10, r"UTMAPF\.L2\.2D": 10, "ACQBULK": 10, "PREEXIT": 10},
    "gemm_tc2.o": {r"UTCHMMA\.2CTA": 50, r"UTMALDG\.2D\.2CTA": 20, r"UTCBAR\.2CTA\.MULTICAST": 10, r"LDTM\.x16": 4, "ACQBULK": 4},
    "attention_tc.o": {r"UTCHMMA\b": 8, r"UTMALDG\.5D": 8, r"LDTM\.x16": 4, r"UTCBAR\b": 2, "ACQBULK": 1},
    "gemv_ks.o": {"ACQBULK": 50, "PREEXIT": 50},
    "attention.o": {"ACQBULK": 2, r"UCGABAR_(ARV|WAIT)": 2},   # decode attention: PDL + cluster barrier (DSMEM merge)
}


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
@pytest.mark.parametrize("obj", sorted(EXPECT))
def test_blackwell_mnemonics_present(built_lib, obj):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "llmlb_b200", "_build", obj)], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    for pat, least in EXPECT[obj].items():
        n = len(re.findall(pat, sass))
        assert n >= least, "%s: %s occurs %d times, expected >= %d" % (obj, pat, n, least)
    assert not re.search(r"\bHGMMA\b", sass)            # no Hopper wgmma
    if obj in ("gemm_tc.o", "gemm_tc2.o", "attention_tc.o"):
        assert not re.search(r"\bHMMA\b", sass), obj    # the tcgen05 kernels carry no mma.sync
