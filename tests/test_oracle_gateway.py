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


# ---- Anthropic Messages front door (llmlb/src/api/anthropic.rs) ---------------------------------
@pytest.mark.parametrize("v", V["anthropic"]["request"], ids=lambda v: v["cite"].split()[-1])
def test_anthropic_request_conversion(v):
    if "error_status" in v:
        with pytest.raises(G.AnthropicError) as e:
            G.anthropic_request_to_openai(v["payload"])
        assert e.value.status == v["error_status"] and e.value.body()["type"] == "error"
        return
    body, text, stream = G.anthropic_request_to_openai(v["payload"])
    x = v["expect"]
    if "roles" in x:
        assert [m["role"] for m in body["messages"]] == x["roles"]
        assert (body["model"], body["max_tokens"], body["stream"], body["stop"]) == (x["model"], x["max_tokens"], x["stream"], x["stop"])
        assert stream is True and x["request_text_contains"] in text
    if "n_tools" in x:
        assert len(body["tools"]) == x["n_tools"] and body["tools"][0]["type"] == x["tool0"]["type"]
        assert body["tools"][0]["function"]["name"] == x["tool0"]["name"]
        assert body["tools"][0]["function"]["description"] == x["tool0"]["description"]
        assert body["tool_choice"] == x["tool_choice"]
    if "tool_message" in x:
        tm = [m for m in body["messages"] if m["role"] == "tool"]
        assert len(tm) == 1 and tm[0]["tool_call_id"] == x["tool_message"]["tool_call_id"] and tm[0]["content"] == x["tool_message"]["content"]


@pytest.mark.parametrize("v", V["anthropic"]["response"], ids=lambda v: v["cite"].split()[-1])
def test_anthropic_response_conversion(v):
    r = G.openai_to_anthropic_message_response(v["body"], v["model"], v["usage"][0], v["usage"][1])
    x = v["expect"]
    assert (r["type"], r["role"], r["stop_reason"]) == (x["type"], x["role"], x["stop_reason"])
    assert r["model"] == v["model"] and r["stop_sequence"] is None and r["id"] == v["body"]["id"]
    if "content0" in x:
        assert r["content"][0] == x["content0"]
        assert (r["usage"]["input_tokens"], r["usage"]["output_tokens"]) == (x["input_tokens"], x["output_tokens"])
    if "tool_use" in x:
        assert len(r["content"]) >= x["min_blocks"]
        tu = [c for c in r["content"] if c["type"] == "tool_use"][0]
        assert (tu["name"], tu["id"], tu["input"]) == (x["tool_use"]["name"], x["tool_use"]["id"], x["tool_use"]["input"])


@pytest.mark.parametrize("v", V["anthropic"]["stream"], ids=lambda v: v["cite"].split()[-1])
def test_anthropic_stream_transform(v):
    t = G.AnthropicStreamTransformer("test-model", input_tokens=None)
    # arbitrary network chunking must not matter
    up = v["upstream"]
    for i in range(0, len(up), 7):
        t.feed(up[i:i + 7])
    t.finish()
    wire = t.wire()
    for needle in v["contains"]:
        assert needle in wire
    names = [n for n, _ in t.out]
    assert names == ["message_start", "content_block_start", "content_block_delta", "content_block_delta",
                     "content_block_stop", "message_delta", "message_stop"]
    assert t.out[0][1]["message"]["id"] == "msg_123"
    assert t.out[-2][1]["delta"]["stop_reason"] == "end_turn"


def test_anthropic_errors_and_edges():
    for v in V["anthropic"]["errors"]:
        with pytest.raises(G.AnthropicError) as e:
            G.anthropic_required_header(v["headers"], "anthropic-version")
        b = e.value.body()
        assert (e.value.status, b["type"], b["error"]["type"]) == (v["status"], v["type"], v["error_type"])
        assert b["error"]["message"] == "Missing required header: anthropic-version"
    for payload, msg in [({}, "model is required"), ({"model": "  "}, "model must not be empty"),
                         ({"model": "m", "messages": []}, "max_tokens is required"),
                         ({"model": "m", "max_tokens": 1}, "messages must be an array"),
                         ({"model": "m", "max_tokens": 1, "messages": [{"content": "x"}]}, "messages[0].role is required"),
                         ({"model": "m", "max_tokens": 1, "messages": [{"role": "system", "content": "x"}]}, "messages[0].role must be 'user' or 'assistant'"),
                         ({"model": "m", "max_tokens": 1, "messages": [{"role": "user"}]}, "messages[0].content is required"),
                         ({"model": "m", "max_tokens": 1, "messages": [], "stop_sequences": "END"}, "stop_sequences must be an array of strings"),
                         ({"model": "m", "max_tokens": 1, "messages": [], "tool_choice": {"type": "nope"}}, "unknown tool_choice type: nope")]:
        with pytest.raises(G.AnthropicError) as e:
            G.anthropic_request_to_openai(payload)
        assert e.value.message == msg
    assert G.map_finish_reason_to_stop_reason("length") == "max_tokens"
    assert G.map_finish_reason_to_stop_reason("weird") == "end_turn"
    # empty completion still yields one empty text block; stream without [DONE] is closed by finish()
    r = G.openai_to_anthropic_message_response({"choices": [{"message": {"content": ""}, "finish_reason": "length"}]}, "m", None, None)
    assert r["content"] == [{"type": "text", "text": ""}] and r["stop_reason"] == "max_tokens" and r["usage"] == {"input_tokens": 0, "output_tokens": 0}
    t = G.AnthropicStreamTransformer("m", input_tokens=7)
    t.feed('data: {"id":"chatcmpl-9","choices":[{"delta":{"content":"x"}}],"usage":{"prompt_tokens":7,"completion_tokens":1}}\n')
    t.finish()
    assert [n for n, _ in t.out][-2:] == ["message_delta", "message_stop"] and t.out[-2][1]["usage"]["output_tokens"] == 1


# ---- outbound payload preparation (model_name.rs:43-108, openai.rs:977-992) -----------------------
@pytest.mark.parametrize("v", V["payload"]["rewrite"], ids=lambda v: v["cite"].split()[-1])
def test_rewrite_payload_model_for_endpoint(v):
    models = [(m, c) for m, c in v["endpoint_models"]]
    out = G.rewrite_payload_model_for_endpoint(v["payload"], v["selected"], v["endpoint_type"], models, V["payload"]["mappings"])
    assert out["model"] == v["model"]
    assert {k: x for k, x in out.items() if k != "model"} == {k: x for k, x in v["payload"].items() if k != "model"}
    if v["model"] == v["payload"]["model"]:
        assert out == v["payload"]


def test_prepare_upstream_payload_injects_include_usage():
    p = {"model": "client-name", "messages": [], "stream": True}
    assert G.prepare_upstream_payload(p, "m", True) == {"model": "m", "messages": [], "stream": True, "stream_options": {"include_usage": True}}
    assert "stream_options" not in G.prepare_upstream_payload(p, "m", False)
    keep = G.prepare_upstream_payload(dict(p, stream_options={"include_usage": False, "x": 1}), "m", True)
    assert keep["stream_options"] == {"include_usage": False, "x": 1}          # the client's explicit choice wins
    assert G.prepare_upstream_payload(dict(p, stream_options={}), "m", True)["stream_options"] == {"include_usage": True}
    assert G.prepare_upstream_payload(dict(p, stream_options="bogus"), "m", True)["stream_options"] == "bogus"
    assert p == {"model": "client-name", "messages": [], "stream": True}       # input untouched
    # no mapping, nothing advertised: the selected name passes through
    assert G.resolve_runtime_model_name_for_endpoint("a", "b", "vllm", [], []) == "b"
    assert G.resolve_runtime_model_name_for_endpoint("a", "b", "vllm", [("x", "b")], []) == "x"
    assert G.resolve_runtime_model_name_for_endpoint("a", "b", "vllm", [("a", None)], []) == "a"


# ---- 60-minute request history (balancer/mod.rs:2643-2658, 2973-3060) ---------------------------
NOW = 1749983445     # 2025-06-15T10:30:45Z


def _play(v, H):
    for outcome, dmin, count in v.get("records", []):
        for _ in range(count):
            H.record(outcome, NOW + 60 * dmin + 7)
    if "then" in v:
        H.record(v["then"][0], NOW + 60 * v["then"][1])


@pytest.mark.parametrize("v", V["history"], ids=lambda v: v["cite"].split()[-1])
def test_request_history(v):
    if "align" in v:
        assert G.align_to_minute(v["align"][0]) == v["align"][1] and v["align"][1] % 60 == 0
        return
    h = G.RequestHistory()
    _play(v, h)
    if "kept_minutes" in v:
        assert [p[0] for p in h.points] == [G.align_to_minute(NOW) + 60 * m for m in v["kept_minutes"]]
    if "totals" in v:
        assert [sum(p[1] for p in h.points), sum(p[2] for p in h.points)] == v["totals"]
    w = h.window(NOW + 10)
    assert len(w) == 60 and w[-1][0] == G.align_to_minute(NOW) and w[0][0] == w[-1][0] - 59 * 60
    assert sum(p[1] for p in w) == sum(p[1] for p in h.points if p[0] >= w[0][0])


@pytest.mark.parametrize("v", V["latency"], ids=lambda v: v["cite"].split()[-1])
def test_inference_latency_ema(v):
    l = G.InferenceLatency()
    assert l.ms is None and l.for_sort() == float("inf")
    for op, x in v["ops"]:
        l.update(x) if op == "update" else l.reset()
    assert l.ms == (float("inf") if v["ms"] == "inf" else v["ms"])
    for op, x in v.get("then", []):
        l.update(x)
    if "ms_then" in v:
        assert abs(l.ms - v["ms_then"]) < 1e-9
    if v["ms"] == "inf":
        l.update(50.0)                       # first sample after a reset replaces the infinity
        assert l.ms == 50.0


# ---- error conventions (row a1.16): LbError table, upstream failure classes, queue errors -----------------------------
def test_lb_error_table_matches_the_reference_tests():
    L = V["lb_errors"]
    for kind, status in L["status"]["cases"]:
        assert G.LB_ERRORS[kind][0] == status, kind
    for kind, etype in L["types"]["cases"]:
        assert G.LB_ERRORS[kind][1] == etype, kind
    for kind, msg, etype, code in L["openai"]["cases"]:
        status, body = G.lb_error_openai(kind)
        assert body == {"error": {"message": msg, "type": etype, "code": code}} and str(status) == code
    nl = L["external_no_leak"]
    assert G.LB_ERRORS[nl["kind"]][2] == nl["external"] and "192.168" not in G.app_error_response(nl["kind"], nl["detail"])[1]["error"]
    for kind, detail, status, shown in L["app"]["cases"]:
        assert G.app_error_response(kind, detail) == (status, {"error": shown}), (kind, detail)
    # a non-validation CommonError passes its text through only when it is the GPU-requirement message (api/error.rs:178-188)
    assert G.app_error_response("common_other", "Configuration error: GPU is required")[1]["error"].endswith("GPU is required")
    assert G.app_error_response("common_other", "Configuration error: bad port 70000")[1]["error"] == "Request error"


def test_queue_and_upstream_error_responses_match_the_reference():
    Q = V["queue_errors"]
    for c in Q["cases"]:
        status, headers, body = G.queue_error_response(c["status"], c["message"], c["type"], c["retry_after"])
        assert status == c["status"] and headers.get("retry-after") == c["header"]
        assert body == {"error": {"message": c["message"], "type": c["type"], "code": c["status"]}}
    for c in Q["call_sites"]:
        status, headers, body = G.queue_capacity_exceeded(c["queue_timeout_secs"]) if c["fn"] == "capacity" else G.queue_wait_timeout()
        assert (status, headers.get("retry-after")) == (c["status"], c["header"])
        assert body["error"] == {"message": c["message"], "type": c["type"], "code": c["status"]}
    for c in V["upstream_errors"]:
        assert G.classify_upstream_request_error(c["kind"], c["timeout_secs"], c["ollama_loading_model"]) == (c["status"], c["type"], c["message"])


# ---- the gateway as client of an endpoint's probe routes (detection / sync / metadata / health) -------------------
def _mock_fetch(routes, unreachable=False):
    """wiremock as the reference's tests use it: unmounted paths answer 404; `unreachable` = nothing listens."""
    def fetch(path, auth):
        if unreachable:
            return None
        r = routes.get(path)
        return (r["status"], r["headers"], r["json"]) if r else (404, {}, None)
    return fetch


@pytest.mark.parametrize("v", V["detection"], ids=[v["cite"][:60] for v in V["detection"]])
def test_endpoint_type_detection(v):
    fetch = _mock_fetch(v["routes"], v.get("unreachable", False))
    if v["type"] is None:
        with pytest.raises(ValueError) as ei:
            G.detect_endpoint_type(fetch)
        assert str(ei.value) == v["error"]
    else:
        assert G.detect_endpoint_type(fetch) == (v["type"], v["reason"])


def test_detection_marker_rules():
    # detection/lm_studio.rs tests: token-boundary matching, not substring matching
    for s, want in (("LM-Studio/0.3.5", True), ("lm studio", True), ("lm_studio", True), ("lmstudio-community", True), ("LMStudio", True),
                    ("film studio", False), ("calm-studio", False), ("organization_owner", False), ("vllm", False)):
        assert G._lm_studio_marker(s) is want, s
    # a models[] entry needs publisher + architecture + (state marker | LM Studio shape)
    assert not G._looks_like_lm_studio_model({"publisher": "x", "arch": "llama"})
    assert G._looks_like_lm_studio_model({"publisher": "x", "arch": "llama", "state": "loaded"})
    # xllm_version of the wrong type is a parse failure, not a detection (serde Option<String>)
    assert G.detect_xllm(_mock_fetch({"/api/system": {"status": 200, "headers": {}, "json": {"xllm_version": 3}}})) is None
    assert G.detect_xllm(_mock_fetch({"/api/system": {"status": 200, "headers": {}, "json": {"server_name": "other"}}})) is None
    assert G.detect_xllm(_mock_fetch({"/api/system": {"status": 500, "headers": {}, "json": {"xllm_version": "1"}}})) is None


@pytest.mark.parametrize("v", V["models_parse"], ids=[v["cite"][-40:] for v in V["models_parse"]])
def test_parse_models_response(v):
    assert G.parse_models_response(v["json"]) == (v["ids"], v["format"])


def test_detect_capabilities():
    for v in V["capabilities"]:
        assert G.detect_capabilities(v["name"]) == v["caps"], v["name"]
    assert G.detect_capabilities("org/embed-x") == ["embeddings"] and G.detect_capabilities("embed-org/chat-model") == ["chat"]    # leaf only


@pytest.mark.parametrize("v", V["xllm_model_info"], ids=[v["cite"][-44:] for v in V["xllm_model_info"]])
def test_xllm_model_info(v):
    r = (200, {}, v["json"])
    if v["want"] is None:
        with pytest.raises(ValueError):
            G.parse_xllm_model_info(r)
    else:
        assert G.parse_xllm_model_info(r) == v["want"]
    assert G.xllm_model_info_url("meta-llama/Llama 3:8b") == "/api/models/meta-llama%2FLlama%203%3A8b/info"      # metadata/xllm.rs:54-63
    with pytest.raises(ValueError):
        G.parse_xllm_model_info((404, {}, {"error": "model not found"}))                                         # tests/support/xllm.rs:35-44


def test_v0_health_fields_are_all_optional():
    # health/endpoint_checker.rs:515-557
    full = {"gpu": {"device_count": 8, "total_memory_bytes": 8 * 183 * 2 ** 30, "used_memory_bytes": 5, "capability_score": 99.5}, "load": {"active_requests": 3}}
    assert G.parse_v0_health((200, {}, full)) == {"gpu_device_count": 8, "gpu_total_memory_bytes": 8 * 183 * 2 ** 30, "gpu_used_memory_bytes": 5,
                                                 "gpu_capability_score": 99.5, "active_requests": 3}
    assert set(G.parse_v0_health((200, {}, {})).values()) == {None}
    assert G.parse_v0_health((200, {}, {"gpu": {"device_count": -1, "capability_score": 7}}))["gpu_device_count"] is None
    with pytest.raises(ValueError):
        G.parse_v0_health((503, {}, {}))
