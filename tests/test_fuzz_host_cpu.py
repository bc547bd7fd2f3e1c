"""A short pass of the mutation fuzzer (tools/fuzz) over every host-side parser of bytes that arrive from outside —
checkpoint files, tokenizer.json, request JSON, relayed SSE — built with AddressSanitizer + UndefinedBehaviorSanitizer.
The long runs (10^5 iterations per mode) are recorded in DESIGN.md; this keeps the harness building and the parsers
clean on every change."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _harness():
    """the sanitizer build of the harness is kept between runs (rebuilt when a parser or the harness changed)"""
    sys.path.insert(0, os.path.join(ROOT, "tools", "fuzz"))
    import run as fuzz_run
    out = os.path.join(ROOT, "tests", "support", "_build", "fuzz_host")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hd = os.path.join(ROOT, "llmlb_b200", "host")
    deps = [os.path.join(ROOT, "tools", "fuzz", "fuzz_host.cpp")] + [os.path.join(hd, f) for f in os.listdir(hd) if f.endswith((".cpp", ".hpp", ".inc"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        fuzz_run.build(out)
    return out


def test_short_fuzz_pass_is_clean():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz", "run.py"), "--iters", "1000", "--seed", "3", "--bin", _harness()],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    lines = r.stdout.strip().splitlines()
    assert sum(l.startswith("ckpt") for l in lines) == 3 and any(l.startswith("tokfile") and "clean" in l for l in lines), r.stdout
    # the seeds are valid files: a pass in which nothing ever opened would be testing the error path only
    for l in lines:
        if l.startswith("ckpt"):
            assert int(l.split(" opened")[0].split()[-1]) > 50, l
