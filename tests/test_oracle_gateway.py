"""Pins oracle/gateway_ref.py against the known-answer vectors of the reference's own tests
(tests/golden/gateway_vectors.json, each citing its source test)."""
import json
import math
import os

import pytest

from oracle import gateway_ref as G

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gateway_vectors.json")))


@pytest.mark.parametrize("v", V["ema"], ids=lambda v: v["cite"].split()[-1])
def test_ema(v):
    s = G.ModelTpsState()
    for tok, dur in v["updates"]:
        s.update_tps(tok, dur)
    if v["ema"] is None:
        assert s.tps_ema is None
    else:
        assert abs(s.tps_ema - v["ema"]) < 0.01
    assert (s.request_count, s.total_output_tokens, s.total_duration_ms) == (
        v["request_count"], v["total_output_tokens"], v["total_duration_ms"])


def test_ema_properties():
    # llmlb/tests/unit/proptest_tests.rs:16-60
    import random
    rnd = random.Random(0)
    for _ in range(500):
        s = G.ModelTpsState()
        calls = [(rnd.randrange(0, 1000), rnd.randrange(1, 10000)) for _ in range(rnd.randrange(1, 20))]
        for t, d in calls:
            s.update_tps(t, d)
        assert s.tps_ema is not None and s.tps_ema >= 0 and s.request_count == len(calls)
        lo = min(t / (d / 1000.0) for t, d in calls)
        hi = max(t / (d / 1000.0) for t, d in calls)
        assert lo - 1e-9 <= s.tps_ema <= hi + 1e-9
    s = G.ModelTpsState()
    s.update_tps(5, 0)
    assert s.tps_ema is None and s.request_count == 0


@pytest.mark.parametrize("v", V["usage"], ids=lambda v: v["cite"])
def test_extract_usage(v):
    got = G.extract_usage_from_response(v["body"])
    if v["want"] is None:
        assert got is None
    else:
        assert [got["input_tokens"], got["output_tokens"], got["total_tokens"]] == v["want"]


@pytest.mark.parametrize("v", V["accumulator"], ids=lambda v: v["cite"].split()[-1])
def test_accumulator(v):
    acc = G.StreamingTokenAccumulator("m")
    for c in v["chunks"]:
        acc.process_chunk(c)
    assert acc.accumulated_content == v["content"] and acc.done == v["done"]
    if "usage" in v:
        u = acc.finalize()
        assert [u["input_tokens"], u["output_tokens"], u["total_tokens"]] == v["usage"]


def test_sse_line_buffering_across_chunk_boundaries():
    body = "".join(c + "\n" for c in V["accumulator"][-1]["chunks"])
    for cut in range(1, len(body)):
        acc = G.StreamingTokenAccumulator("m")
        buf = G.process_sse_lines(body[:cut], acc)
        buf = G.process_sse_lines(buf + body[cut:], acc)
        assert acc.accumulated_content == "Hello world" and acc.done and buf == ""


@pytest.mark.parametrize("v", V["routing"], ids=lambda v: v["cite"].split()[-1])
def test_routing(v):
    lm = G.LoadManager()
    for eid, models in v["endpoints"]:
        lm.add_endpoint(eid, models)
    for eid, model, kind, tok, dur in v["tps"]:
        lm.update_tps(eid, model, kind, tok, dur)
    for model, kind, want in v["selects"]:
        assert lm.select(model, kind) == want


def test_routing_exclusions_and_errors():
    lm = G.LoadManager()
    lm.add_endpoint("off", ["m"], status="offline")
    lm.add_endpoint("init", ["m"], initializing=True)
    with pytest.raises(LookupError) as ei:
        lm.select("m", "chat")
    assert str(ei.value) == "no_endpoints_available"   # balancer/mod.rs:1948-1950
    with pytest.raises(LookupError) as ei:
        lm.select("unknown", "chat")
    assert str(ei.value) == "no_capable_endpoints"      # balancer/mod.rs:1928-1933
    lm.add_endpoint("ok", ["m"])
    assert lm.select("m", "chat") == "ok"


def test_model_lookup_keys_and_alias_routing():
    maps = V["mappings"]
    assert G.model_lookup_keys("llama3.3:70b", maps) == ["llama3.3:70b", "meta-llama/Llama-3.3-70B-Instruct"]
    assert G.model_lookup_keys("META-LLAMA/llama-3.3-70b-instruct", maps)[1] == "meta-llama/Llama-3.3-70B-Instruct"
    assert G.model_lookup_keys("plain", maps) == ["plain"]
    lm = G.LoadManager(maps)
    lm.add_endpoint("ollama", ["llama3.3:70b"])
    assert lm.select("meta-llama/Llama-3.3-70B-Instruct", "chat") == "ollama"


@pytest.mark.parametrize("v", V["model_names"], ids=lambda v: v["in"])
def test_parse_quantized_model_name(v):
    if v.get("error"):
        with pytest.raises(ValueError):
            G.parse_quantized_model_name(v["in"])
    else:
        p = G.parse_quantized_model_name(v["in"])
        assert (p["base"], p["quantization"]) == (v["base"], v["quant"])


@pytest.mark.parametrize("v", V["errors"], ids=lambda v: v["cite"])
def test_error_bodies(v):
    assert G.openai_error_body(v["message"], v["type"], v["status"]) == v["body"]
    st, hdr, body = G.gate_rejection()
    assert st == 503 and hdr["retry-after"] == "30" and body == V["errors"][1]["body"]


@pytest.mark.parametrize("v", V["api_keys"], ids=lambda v: str(v["headers"]))
def test_extract_api_key(v):
    if "error" in v:
        with pytest.raises(PermissionError) as ei:
            G.extract_api_key(v["headers"])
        assert str(ei.value) == v["error"]
    else:
        assert G.extract_api_key(v["headers"]) == v["key"]
