"""SASS guard for programmatic dependent launch: lists, per kernel, the global loads that sit AHEAD of the first
ACQBULK (= griddepcontrol.wait).  A kernel that is launched as a programmatic secondary may only read data there that
its predecessor does not write (weights, tables, inputs uploaded at the start of the step).  Round 2 found four loads
of `qkv` hoisted above the wait in decode_attention_kernel — `const __restrict__` on a pointer to predecessor-written
data lets nvcc do that — which made tensor-parallel decode non-deterministic.  Runs without a GPU (cuobjdump only):

    python tools/sass_pdl_scan.py                 # table
    python tools/sass_pdl_scan.py --check         # exit 1 if a kernel exceeds its allowance (tests/test_sass_pdl_cpu.py)"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# loads allowed ahead of the wait: kernel-name prefix -> (max loads, max LDG...CONSTANT), with what they are
ALLOW = {
    "llmlb::gemv_ks_kernel": (11, 8, "8 weight loads of the first row batch (ld.global.nc) + trace-buffer words"),
    "llmlb::decode_attention_kernel": (22, 3, "seq_len, block-table row, rope row (CONSTANT); 16 K/V page loads of OLD tokens; trace words"),
    "llmlb::prefill_attention_kernel_tc": (1, 1, "the tile descriptor"),
    "llmlb::rope_append_kernel": (2, 2, "position and page of the token (uploaded at the start of the step)"),
    "llmlb::rmsnorm_parts_kernel": (0, 0, "nothing"),
    "llmlb::gemm_tc2_kernel": (3, 0, "trace-buffer words"),
    "llmlb::gemm_tc_kernel": (12, 0, "trace-buffer words, tensor-parallel descriptor fields"),
}


def scan():
    rows = []
    for o in sorted(glob.glob(os.path.join(ROOT, "llmlb_b200", "_build", "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        for f in re.split(r"\n\s*Function : ", txt)[1:]:
            name = f.split("\n", 1)[0].strip()
            lines = f.split("\n")
            acq = [i for i, l in enumerate(lines) if "ACQBULK" in l]
            if not acq:
                continue
            pre = [l for l in lines[: acq[0]] if re.search(r"\bLDG\.", l)]
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"^void ", "", re.sub(r"\(.*", "", dem))
            rows.append((dem, len(pre), sum("CONSTANT" in l for l in pre)))
    return rows


def main():
    rows = scan()
    bad = []
    for dem, n, nc in sorted(set(rows)):
        key = next((k for k in ALLOW if dem.startswith(k)), None)
        lim = ALLOW.get(key)
        ok = lim is not None and n <= lim[0] and nc <= lim[1]
        if not ok:
            bad.append(dem)
        if "--check" not in sys.argv or not ok:
            print("%-62s loads ahead of the wait: %2d (LDG.CONSTANT %d)  %s" % (dem[:62], n, nc, "ok: " + lim[2] if ok else "NOT ALLOWED (limit %s)" % (lim[:2] if lim else "none: add the kernel to ALLOW after review")))
    if "--check" in sys.argv:
        print("%d kernels with a dependency wait scanned, %d over their allowance" % (len(set(rows)), len(bad)))
        sys.exit(1 if bad or not rows else 0)


if __name__ == "__main__":
    main()
