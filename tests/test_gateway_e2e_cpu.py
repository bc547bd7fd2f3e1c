"""End to end the way the reference's own e2e tests read (tests/e2e_openai_proxy.rs, contract/test_proxy_completions.rs,
contract/responses_api_test.rs, integration/model_routing_balancing_test.rs): a gateway in front, endpoints behind, requests
through the front door.  The gateway is tests/support/mini_gateway.py — the oracle's restatements of the reference glued in
the reference's order, nothing else; the endpoints are this repo's HTTP shim (on the scripted engine, so it runs anywhere).
What it shows: an unmodified llmlb could register the shim, route to it by TPS, relay its streams and account for them."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "support"))
import test_server_fake_engine_cpu as F  # noqa: E402
import test_server_gpu as T  # noqa: E402
from mini_gateway import MiniGateway  # noqa: E402
from oracle import gateway_ref as G  # noqa: E402

fake_bin = F.fake_bin


@pytest.fixture(scope="module")
def two_endpoints(fake_bin):
    a = F._start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8", "--max-ctx", "512")
    b = F._start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8", "--max-ctx", "512", "--api-key", "sk-b")
    yield a[0], b[0]
    for _, proc in (a, b):
        proc.terminate(); proc.wait(timeout=20)


def _gw(two_endpoints):
    gw = MiniGateway()
    assert gw.register("A", two_endpoints[0])[0] == 201
    st, body = gw.register("B", two_endpoints[1], api_key="sk-b")
    assert st == 201 and body["endpoint_type"] == "xllm" and body["models"] == [
        {"model_id": "tiny-llama", "capabilities": ["chat"], "supported_apis": ["chat_completions"], "max_tokens": 512}]
    return gw


def test_registration_detects_syncs_and_rejects_what_does_not_answer(two_endpoints):
    gw = _gw(two_endpoints)
    assert gw.eps["A"]["reason"] == "xLLM: /api/system xllm_version=llmlb_b200-0.1"
    assert gw.register("dead", T._free_port())[0] == 502                       # endpoint_type_detection_test.rs:51-70
    assert gw.register("B-without-key", two_endpoints[1])[0] == 400           # answers 401 to every probe: not a supported type


def test_chat_non_stream_and_stream_are_relayed_and_accounted(two_endpoints):
    gw = _gw(two_endpoints)
    req = {"model": "tiny-llama:Q4_K_M", "messages": [{"role": "user", "content": "hello"}], "max_tokens": 9, "temperature": 0}
    st, hd, body = gw.post("/v1/chat/completions", req)
    j = json.loads(body)
    assert st == 200 and j["model"] == "tiny-llama:Q4_K_M" and j["usage"]["completion_tokens"] == 9          # the client's own model name comes back
    eid = hd["x-endpoint"]
    s = gw.book.state[eid]
    assert (s.assigned_active, s.total_assigned, s.success_count, s.error_count) == (0, 1, 1, 0)
    assert s.total_output_tokens == 9 and s.total_input_tokens == j["usage"]["prompt_tokens"] and s.total_tokens == j["usage"]["total_tokens"]
    tps = gw.lm.tps[(eid, "tiny-llama", "chat_completions")]
    assert tps.request_count == 1 and tps.total_output_tokens == 9 and tps.tps_ema == 9 / (tps.total_duration_ms / 1000.0)
    # streamed, the client did NOT ask for usage: the gateway injects stream_options.include_usage (openai.rs:977-992), the
    # endpoint honours it, the relay counts from the usage chunk; the bytes reach the client untouched
    st, hd, sse = gw.post("/v1/chat/completions", dict(req, stream=True))
    text = sse.decode()
    assert st == 200 and text.endswith("data: [DONE]\n\n") and hd["content-type"] == "text/event-stream"
    events = [json.loads(l[6:]) for l in text.split("\n") if l.startswith("data: {")]
    assert events[-1]["usage"]["completion_tokens"] == 9 and events[0]["choices"][0]["delta"].get("role") == "assistant"
    eid2 = hd["x-endpoint"]
    assert gw.book.state[eid2].total_output_tokens == (18 if eid2 == eid else 9)
    assert gw.lm.tps[(eid2, "tiny-llama", "chat_completions")].total_output_tokens == (18 if eid2 == eid else 9)


def test_tps_priority_routing_round_robin_ties_and_offline_endpoints(two_endpoints):
    """tests/unit/tps_routing_test.rs + integration/model_routing_balancing_test.rs: unmeasured endpoints tie at 0 and take
    turns; once measured the faster one wins every time; an offline endpoint loses its TPS state and is skipped; with none
    online the answer is 503 no_capable_nodes (the model is known), an unknown model is 404."""
    gw = _gw(two_endpoints)
    req = {"model": "tiny-llama", "prompt": "x", "max_tokens": 3, "temperature": 0}
    assert gw.post("/v1/completions", req)[1]["x-endpoint"] == "A"            # nobody measured: scores tie at 0.0, the cursor starts at A
    assert gw.post("/v1/completions", req)[1]["x-endpoint"] == "A"            # A now HAS a TPS, B still scores 0.0: measured beats unmeasured
    gw.lm.tps.clear()
    picks = [None, None]
    for i in range(2):                                                         # ties again: consecutive selections take turns
        picks[i] = gw.post("/v1/completions", req)[1]["x-endpoint"]
        gw.lm.tps.clear()
    assert sorted(picks) == ["A", "B"]
    gw.lm.tps.clear()
    gw.lm.update_tps("A", "tiny-llama", "completions", 100, 1000)              # A: 100 tok/s
    gw.lm.update_tps("B", "tiny-llama", "completions", 1000, 1000)             # B: 1000 tok/s
    gw.lm.tps[("B", "tiny-llama", "completions")].tps_ema = 1e9                # far above anything a real sample can pull A to
    assert [gw.post("/v1/completions", req)[1]["x-endpoint"] for _ in range(4)] == ["B"] * 4
    assert gw.post("/v1/chat/completions", dict(model="tiny-llama", messages=[{"role": "user", "content": "x"}], max_tokens=2))[0] == 200   # other api kind: own TPS table
    gw.set_status("B", "offline")
    assert ("B", "tiny-llama", "completions") not in gw.lm.tps
    assert [gw.post("/v1/completions", req)[1]["x-endpoint"] for _ in range(2)] == ["A", "A"]
    gw.set_status("A", "offline")
    st, _, body = gw.post("/v1/completions", req)
    assert st == 503 and json.loads(body) == G.model_unavailable_body("No available endpoints support model: tiny-llama", "no_capable_nodes")
    st, _, body = gw.post("/v1/completions", dict(req, model="gpt-nonexistent"))
    assert st == 404 and json.loads(body)["error"]["message"] == "The model 'gpt-nonexistent' does not exist"
    assert gw.post("/v1/completions", dict(req, model="a:b:c"))[0] == 400      # model_name.rs:19-40


def test_upstream_errors_502_on_the_openai_routes_pass_through_on_responses(two_endpoints):
    """openai.rs:1178-1213 vs responses.rs:411-424 (contract/test_proxy_completions.rs:236-320, contract/responses_api_test.rs:136):
    the endpoint refuses a request it cannot hold (prompt + max_tokens beyond its context) with 400."""
    gw = _gw(two_endpoints)
    too_long = {"model": "tiny-llama", "prompt_token_ids": [1] * 600, "max_tokens": 4}      # 600 prompt tokens against a 512-token context
    st, _, body = gw.post("/v1/completions", too_long)
    e = json.loads(body)["error"]
    assert st == 502 and e["type"] == "endpoint_upstream_error" and e["code"] == 502 and "context" in e["message"], e
    eid = next(k for k, s in gw.book.state.items() if s.error_count)
    assert gw.book.state[eid].error_count == 1 and gw.book.state[eid].assigned_active == 0
    st, _, body = gw.post("/v1/responses", {"model": "tiny-llama", "input": "x" * 3000, "max_output_tokens": 400})
    assert st == 400 and "error" in json.loads(body)                            # the endpoint's own status and body
    st, _, body = gw.post("/v1/responses", {"model": "tiny-llama", "input": "hi", "max_output_tokens": 5})
    j = json.loads(body)
    assert st == 200 and j["usage"]["output_tokens"] == 5 and j["object"] == "response"
    assert gw.lm.tps[(next(iter({k[0] for k in gw.lm.tps if k[2] == "responses"})), "tiny-llama", "responses")].total_output_tokens == 5
