"""examples/inprocess_host.c: the gateway's hot path with the engine in the process, written in C99 over the three public
headers only — gate, TPS-EMA router, lease, submit/poll on the engine, SSE to the client, the relay's accounting over the
same bytes, lease completion and TPS update (the Rust crate under ffi/ does the same and cannot be compiled
here).  Built against the scripted engine of tests/support/fake_engine.cpp and checked against the gateway oracle: the
client-visible stream parses with the reference's accumulator to the engine's own usage, and the router state after N
requests is the oracle's, bit for bit, for the durations the program measured."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gateway_ref as G  # noqa: E402

BUILD = os.path.join(HERE, "support", "_build")


@pytest.fixture(scope="module")
def exe():
    from llmlb_b200 import build
    host = build.build_host()
    os.makedirs(BUILD, exist_ok=True)
    fake = os.path.join(BUILD, "libllmlb_b200.so")
    fake_src = os.path.join(HERE, "support", "fake_engine.cpp")
    if not os.path.exists(fake) or os.path.getmtime(fake) < os.path.getmtime(fake_src):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread", fake_src, "-o", fake])
    out = os.path.join(BUILD, "inprocess_host")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "inprocess_host.c"), "-L" + BUILD, "-lllmlb_b200", "-L" + os.path.dirname(host), "-lllmlb_host",
                           "-Wl,-rpath," + BUILD, "-Wl,-rpath," + os.path.dirname(host), "-o", out])
    return out


def _run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    streams = r.stdout.decode("utf-8").split("### request ")[1:]
    recs = [json.loads(l) for l in r.stderr.decode().splitlines() if l.startswith("{")]
    return streams, recs[:-1], recs[-1]


@pytest.mark.parametrize("api", ["chat", "responses"])
def test_in_process_path_against_the_gateway_oracle(exe, api):
    n_req, n_out = 4, 12
    streams, per, final = _run(exe, "--api", api, "--requests", str(n_req), "--max-tokens", str(n_out), "--prompt-len", "24")
    assert len(streams) == n_req == len(per)
    kind = "responses" if api == "responses" else "chat_completions"
    state = G.ModelTpsState()
    for i, (s, rec) in enumerate(zip(streams, per)):
        head, sse = s.split("\n", 1)
        assert int(head) == i == rec["request"] and rec["endpoint"] == "in-process"
        # the reference's own accounting over the bytes the client received
        acc = G.StreamingTokenAccumulator("llama-tiny")
        assert G.process_sse_lines(sse, acc) == "" and acc.done
        assert acc.finalize() == {"input_tokens": 24, "output_tokens": n_out, "total_tokens": 24 + n_out}
        assert rec["usage"] == [24, n_out, 24 + n_out] and rec["tokens"] == n_out and rec["done"] == 1 and rec["finish"] == 2
        assert len(acc.accumulated_content.encode()) == rec["content_bytes"] == rec["sent_bytes"]
        events = [json.loads(e[6:]) for e in sse.split("\n\n") if e.startswith("data: {")]
        if api == "chat":
            assert events[0]["choices"][0]["delta"] == {"role": "assistant"} and events[-2]["choices"][0]["finish_reason"] == "length"
            assert events[-1]["usage"]["completion_tokens"] == n_out and sse.endswith("data: [DONE]\n\n")
        else:
            types = [e["type"] for e in events]
            assert types[:3] == ["response.created", "response.output_item.added", "response.content_part.added"]
            assert types[-2:] == ["response.output_text.done", "response.done"] and types.count("response.output_text.delta") == n_out
        state.update_tps(n_out, rec["ms"])
    # router state == the oracle's for the same (tokens, ms) sequence: EMA alpha 0.2 (balancer/types.rs:102-118), bit for bit
    assert final["request_count"] == n_req and final["total_output_tokens"] == n_req * n_out
    assert final["total_duration_ms"] == sum(r["ms"] for r in per)
    assert all(r["ms"] >= 1 for r in per) and final["tps_ema"] == state.tps_ema          # durations are clamped to >= 1 ms (proxy.rs:157)
    active, assigned, success, errors, lat_sum, tin, tout, ttot = final["stats"]
    assert (active, assigned, success, errors) == (0, n_req, n_req, 0) and final["in_flight"] == 0
    assert (tin, tout, ttot) == (24 * n_req, n_out * n_req, (24 + n_out) * n_req) and lat_sum == sum(r["ms"] for r in per)
    assert kind in ("chat_completions", "responses")


def test_in_process_path_with_the_native_tokenizer(exe):
    """ids -> text through the streaming detokenizer: whatever the scripted ids decode to, the client only ever receives
    complete UTF-8 and the accumulator's content is exactly what was sent."""
    tj = os.path.join(HERE, "golden", "tokenizer_llama3_style.json")
    streams, per, final = _run(exe, "--api", "chat", "--requests", "2", "--max-tokens", "40", "--tokenizer", tj, "--vocab", "3072")
    for s, rec in zip(streams, per):
        sse = s.split("\n", 1)[1]
        acc = G.StreamingTokenAccumulator("llama-tiny")
        G.process_sse_lines(sse, acc)
        assert acc.done and rec["usage"][1] == 40 and rec["usage"][0] > 5          # the chat template's tokens, counted by the engine
        assert len(acc.accumulated_content.encode()) == rec["sent_bytes"] == rec["content_bytes"]
    assert final["stats"][2] == 2


def test_c_host_on_the_product_engine_library_over_the_fake_cuda_runtime(built_lib, tmp_path):
    """The same C program linked against the PRODUCT libllmlb_b200.so, run with tests/support/fake_cudart.cpp preloaded (host
    memory, no-op launches: every token id is 0): the engine's real submit / poll / release path under a C caller, with the
    gateway-side bookkeeping checked as above."""
    sys.path.insert(0, HERE)
    import test_engine_host_logic_cpu as HL
    from llmlb_b200 import build
    host = build.build_host()
    d = os.path.dirname(built_lib)
    exe_path = str(tmp_path / "inprocess_host_real")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "inprocess_host.c"),
                           "-L" + d, "-lllmlb_b200", "-L" + os.path.dirname(host), "-lllmlb_host", "-Wl,-rpath," + d, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe_path])
    r = subprocess.run([exe_path, "--api", "responses", "--requests", "3", "--max-tokens", "20", "--prompt-len", "30"], capture_output=True, timeout=120,
                       env=dict(os.environ, LD_PRELOAD=HL.build_fake()))
    assert r.returncode == 0, r.stderr.decode()[-1500:]
    streams = r.stdout.decode().split("### request ")[1:]
    recs = [json.loads(l) for l in r.stderr.decode().splitlines() if l.startswith("{")]
    state = G.ModelTpsState()
    for s, rec in zip(streams, recs[:-1]):
        acc = G.StreamingTokenAccumulator("llama-tiny")
        assert G.process_sse_lines(s.split("\n", 1)[1], acc) == "" and acc.done
        assert acc.finalize() == {"input_tokens": 30, "output_tokens": 20, "total_tokens": 50} and acc.accumulated_content == "<0> " * 20
        assert rec["usage"] == [30, 20, 50] and rec["finish"] == 2 and rec["ms"] >= 1
        state.update_tps(20, rec["ms"])
    assert recs[-1]["request_count"] == 3 and recs[-1]["tps_ema"] == state.tps_ema and recs[-1]["stats"][:4] == [0, 3, 3, 0]
    # without the preload the same binary refuses: there is no CPU path
    r = subprocess.run([exe_path, "--requests", "1"], capture_output=True, timeout=60)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode != 0 and b"no CUDA device" in r.stderr
