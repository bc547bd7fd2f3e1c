"""The HTTP shim's start-up checks that need no GPU: unreadable / malformed checkpoints and
`--model auto` without a describing file stop the process with a message before the engine is
touched."""
import os
import subprocess

import pytest

from llmlb_b200 import build

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llmlb_b200", "llmlb_b200_server")


@pytest.fixture(scope="module")
def server_bin():
    build.build()
    build.build_host()
    assert os.path.exists(BIN)
    return BIN


def run(server_bin, *args):
    return subprocess.run([server_bin, "--port", "1", *args], capture_output=True, text=True, timeout=120)


def test_bad_checkpoints_are_rejected_before_the_engine_starts(server_bin, tmp_path):
    r = run(server_bin, "--weights", str(tmp_path / "missing.gguf"))
    assert r.returncode == 2 and "cannot read" in r.stderr
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"GGUF" + (7).to_bytes(4, "little") + b"\0" * 64)
    r = run(server_bin, "--weights", str(bad))
    assert r.returncode == 2 and "unsupported version" in r.stderr
    r = run(server_bin, "--model", "auto")
    assert r.returncode == 2 and "--model auto needs --weights" in r.stderr
