"""The C++ gateway-side code (llmlb_b200/host, libllmlb_host.so) against (a) the known-answer
vectors of the reference's own tests and (b) the Python oracle restatement, on CPU."""
import ctypes as C
import json
import os
import random

import pytest

from oracle import gateway_ref as G

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gateway_vectors.json")))
KIND = {"chat": 0, "completions": 1, "responses": 2, None: -1}


@pytest.fixture(scope="module")
def H():
    from llmlb_b200 import build
    lib = C.CDLL(build.build_host())
    vp, cp, u64, i64 = C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64
    sig = {
        "llmlb_lm_create": (vp, []), "llmlb_lm_destroy": (None, [vp]),
        "llmlb_lm_add_mapping": (None, [vp, cp, cp]), "llmlb_lm_add_endpoint": (None, [vp, cp, C.c_int, C.c_int]),
        "llmlb_lm_add_model": (C.c_int, [vp, cp, cp, cp]), "llmlb_lm_set_status": (C.c_int, [vp, cp, C.c_int]),
        "llmlb_lm_set_initializing": (C.c_int, [vp, cp, C.c_int]),
        "llmlb_lm_update_tps": (None, [vp, cp, cp, C.c_int, u64, u64]),
        "llmlb_lm_get_tps": (C.c_int, [vp, cp, cp, C.c_int, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "llmlb_lm_select": (C.c_int, [vp, cp, C.c_int, C.c_char_p, C.c_size_t]),
        "llmlb_lm_lookup_keys": (C.c_size_t, [vp, cp, C.c_char_p, C.c_size_t]),
        "llmlb_lm_begin_request": (C.c_int, [vp, cp]), "llmlb_lm_finish_request": (C.c_int, [vp, cp, C.c_int, u64, u64]),
        "llmlb_lm_active": (C.c_uint32, [vp, cp]),
        "llmlb_extract_usage": (C.c_int, [cp, C.POINTER(i64)]),
        "llmlb_acc_create": (vp, [cp]), "llmlb_acc_destroy": (None, [vp]), "llmlb_acc_set_input_tokens": (None, [vp, C.c_uint32]),
        "llmlb_acc_process_chunk": (None, [vp, cp]), "llmlb_acc_feed": (None, [vp, cp, C.c_size_t]),
        "llmlb_acc_content": (C.c_size_t, [vp, C.c_char_p, C.c_size_t]), "llmlb_acc_done": (C.c_int, [vp]),
        "llmlb_acc_finalize": (None, [vp, C.POINTER(i64)]),
        "llmlb_parse_model_name": (C.c_int, [cp, C.c_char_p, C.c_char_p, C.c_size_t]),
        "llmlb_error_body": (C.c_size_t, [cp, cp, C.c_int, C.c_char_p, C.c_size_t]),
        "llmlb_gate_rejection_body": (C.c_size_t, [C.c_char_p, C.c_size_t]),
        "llmlb_extract_api_key": (C.c_int, [cp, cp, C.c_char_p, C.c_size_t]),
        "llmlb_sha256_hex": (None, [cp, C.c_size_t, C.c_char_p]),
        "llmlb_gate_create": (vp, []), "llmlb_gate_destroy": (None, [vp]), "llmlb_gate_try_begin": (C.c_int, [vp]),
        "llmlb_gate_end": (None, [vp]), "llmlb_gate_set_rejecting": (None, [vp, C.c_int]), "llmlb_gate_in_flight": (C.c_uint32, [vp]),
        "llmlb_frame": (C.c_size_t, [C.c_int, cp, cp, i64, C.POINTER(cp), C.c_uint32, C.c_uint32, cp, C.c_char_p, C.c_size_t]),
        "llmlb_json_roundtrip": (C.c_size_t, [cp, C.c_char_p, C.c_size_t]),
        "llmlb_classify_upstream_error": (C.c_size_t, [C.c_int, C.c_uint32, cp, C.c_char_p, C.c_size_t]),
        "llmlb_queue_error": (C.c_size_t, [C.c_int, u64, C.c_char_p, C.c_size_t]),
        "llmlb_lb_error_count": (C.c_int, []),
        "llmlb_lb_error": (C.c_size_t, [C.c_int, cp, C.c_char_p, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def b(s):
    return s.encode() if s is not None else None


def usage3(arr):
    return [None if v < 0 else int(v) for v in arr]


class Lm:
    def __init__(self, H, mappings=()):
        self.H, self.p = H, H.llmlb_lm_create()
        for m in mappings:
            for a in m["aliases"]:
                H.llmlb_lm_add_mapping(self.p, b(m["canonical"]), b(a))

    def add(self, eid, models, online=True, initializing=False):
        self.H.llmlb_lm_add_endpoint(self.p, b(eid), int(online), int(initializing))
        for m in models:
            self.H.llmlb_lm_add_model(self.p, b(eid), b(m), None)

    def select(self, model, kind):
        out = C.create_string_buffer(128)
        rc = self.H.llmlb_lm_select(self.p, b(model), KIND[kind], out, 128)
        return rc, out.value.decode()

    def __del__(self):
        self.H.llmlb_lm_destroy(self.p)


@pytest.mark.parametrize("v", V["ema"], ids=lambda v: v["cite"].split()[-1])
def test_ema_vectors(H, v):
    lm = Lm(H)
    lm.add("e", ["m"])
    for tok, dur in v["updates"]:
        H.llmlb_lm_update_tps(lm.p, b"e", b"m", 0, tok, dur)
    ema, cnt, tk, ms = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    found = H.llmlb_lm_get_tps(lm.p, b"e", b"m", 0, ema, cnt, tk, ms)
    if v["ema"] is None:
        assert (not found) or ema.value < 0
        assert cnt.value == 0
    else:
        assert found and abs(ema.value - v["ema"]) < 0.01
        assert (cnt.value, tk.value, ms.value) == (v["request_count"], v["total_output_tokens"], v["total_duration_ms"])


def test_ema_matches_oracle_bit_for_bit(H):
    rnd = random.Random(1)
    for _ in range(200):
        lm, ref = Lm(H), G.ModelTpsState()
        lm.add("e", ["m"])
        for _ in range(rnd.randrange(1, 12)):
            t, d = rnd.randrange(0, 5000), rnd.randrange(0, 20000)
            H.llmlb_lm_update_tps(lm.p, b"e", b"m", 2, t, d)
            ref.update_tps(t, d)
        ema, cnt, tk, ms = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        found = H.llmlb_lm_get_tps(lm.p, b"e", b"m", 2, ema, cnt, tk, ms)
        if ref.tps_ema is None:
            assert (not found) or ema.value < 0
        else:
            assert ema.value == ref.tps_ema  # same IEEE-754 double operations in the same order
            assert cnt.value == ref.request_count


@pytest.mark.parametrize("v", V["routing"], ids=lambda v: v["cite"].split()[-1])
def test_routing_vectors(H, v):
    lm = Lm(H)
    for eid, models in v["endpoints"]:
        lm.add(eid, models)
    for eid, model, kind, tok, dur in v["tps"]:
        H.llmlb_lm_update_tps(lm.p, b(eid), b(model), KIND[kind], tok, dur)
    for model, kind, want in v["selects"]:
        assert lm.select(model, kind) == (0, want)


def test_routing_random_against_oracle(H):
    rnd = random.Random(7)
    for _ in range(100):
        lm, ref = Lm(H, V["mappings"]), G.LoadManager(V["mappings"])
        names = ["m1", "m2", "llama3.3:70b", "meta-llama/Llama-3.3-70B-Instruct"]
        for i in range(rnd.randrange(1, 6)):
            models = rnd.sample(names, rnd.randrange(1, 3))
            online, init = rnd.random() > 0.2, rnd.random() < 0.15
            lm.add("e%d" % i, models, online, init)
            ref.add_endpoint("e%d" % i, models, "online" if online else "offline", init)
            for m in models:
                if rnd.random() < 0.6:
                    k = rnd.choice(["chat", "responses"])
                    t, d = rnd.randrange(1, 500), rnd.randrange(1, 3000)
                    H.llmlb_lm_update_tps(lm.p, b("e%d" % i), b(m), KIND[k], t, d)
                    ref.update_tps("e%d" % i, m, k, t, d)
        for _ in range(8):
            model = rnd.choice(names + ["unknown", None])
            kind = rnd.choice(["chat", "responses", None])
            try:
                want = (0, ref.select(model, kind))
            except LookupError as e:
                want = ({"no_capable_endpoints": 1, "no_endpoints_available": 2}[str(e)], "")
            assert lm.select(model, kind) == want


def test_offline_clears_tps_and_lease_counts(H):
    lm = Lm(H)
    lm.add("a", ["m"]); lm.add("b", ["m"])
    H.llmlb_lm_update_tps(lm.p, b"a", b"m", 0, 100, 1000)
    assert lm.select("m", "chat") == (0, "a")
    H.llmlb_lm_set_status(lm.p, b"a", 0)
    assert lm.select("m", "chat") == (0, "b")
    H.llmlb_lm_set_status(lm.p, b"a", 1)           # back online: TPS restarted from zero
    ema, c1, c2, c3 = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert H.llmlb_lm_get_tps(lm.p, b"a", b"m", 0, ema, c1, c2, c3) == 0
    H.llmlb_lm_begin_request(lm.p, b"b"); H.llmlb_lm_begin_request(lm.p, b"b")
    assert H.llmlb_lm_active(lm.p, b"b") == 2
    H.llmlb_lm_finish_request(lm.p, b"b", 1, 50, 10)
    assert H.llmlb_lm_active(lm.p, b"b") == 1


def test_lookup_keys(H):
    lm = Lm(H, V["mappings"])
    out = C.create_string_buffer(512)
    H.llmlb_lm_lookup_keys(lm.p, b"llama3.3:70b", out, 512)
    assert out.value.decode().split("\n") == G.model_lookup_keys("llama3.3:70b", V["mappings"])


@pytest.mark.parametrize("v", V["usage"], ids=lambda v: v["cite"])
def test_usage_vectors(H, v):
    out = (C.c_int64 * 3)()
    found = H.llmlb_extract_usage(b(json.dumps(v["body"])), out)
    if v["want"] is None:
        assert not found
    else:
        assert found and usage3(out) == v["want"]


@pytest.mark.parametrize("v", V["accumulator"], ids=lambda v: v["cite"].split()[-1])
def test_accumulator_vectors(H, v):
    a = H.llmlb_acc_create(b"m")
    for c in v["chunks"]:
        H.llmlb_acc_process_chunk(a, b(c))
    buf = C.create_string_buffer(4096)
    H.llmlb_acc_content(a, buf, 4096)
    assert buf.value.decode() == v["content"] and bool(H.llmlb_acc_done(a)) == v["done"]
    if "usage" in v:
        out = (C.c_int64 * 3)()
        H.llmlb_acc_finalize(a, out)
        assert usage3(out) == v["usage"]
    H.llmlb_acc_destroy(a)


def test_feed_handles_any_chunk_boundary(H):
    body = "".join(c + "\n" for c in V["accumulator"][-2]["chunks"]).encode()
    for cut in range(1, len(body), 7):
        a = H.llmlb_acc_create(b"m")
        H.llmlb_acc_feed(a, body[:cut], cut)
        H.llmlb_acc_feed(a, body[cut:], len(body) - cut)
        buf = C.create_string_buffer(256)
        H.llmlb_acc_content(a, buf, 256)
        assert buf.value.decode() == "Hello from streaming!" and H.llmlb_acc_done(a)
        H.llmlb_acc_destroy(a)


@pytest.mark.parametrize("v", V["model_names"], ids=lambda v: v["in"])
def test_model_names(H, v):
    base, q = C.create_string_buffer(128), C.create_string_buffer(128)
    rc = H.llmlb_parse_model_name(b(v["in"]), base, q, 128)
    if v.get("error"):
        assert rc == -1
    else:
        assert base.value.decode() == v["base"] and (q.value.decode() or None) == v["quant"]


@pytest.mark.parametrize("v", V["errors"], ids=lambda v: v["cite"])
def test_error_bodies(H, v):
    out = C.create_string_buffer(1024)
    H.llmlb_error_body(b(v["message"]), b(v["type"]), v["status"], out, 1024)
    assert json.loads(out.value) == v["body"]
    H.llmlb_gate_rejection_body(out, 1024)
    assert json.loads(out.value) == V["errors"][1]["body"]


@pytest.mark.parametrize("v", V["api_keys"], ids=lambda v: str(v["headers"]))
def test_api_key(H, v):
    out = C.create_string_buffer(256)
    rc = H.llmlb_extract_api_key(b(v["headers"].get("X-API-Key")), b(v["headers"].get("Authorization")), out, 256)
    if "error" in v:
        assert rc != 0 and out.value.decode() == v["error"]
    else:
        assert rc == 0 and out.value.decode() == v["key"]


def test_sha256_matches_hashlib(H):
    import hashlib
    for s in [b"", b"sk_debug", b"a" * 55, b"a" * 56, b"a" * 64, os.urandom(1000)]:
        out = C.create_string_buffer(65)
        H.llmlb_sha256_hex(s, len(s), out)
        assert out.value.decode() == hashlib.sha256(s).hexdigest()


def test_gate(H):
    g = H.llmlb_gate_create()
    assert H.llmlb_gate_try_begin(g) == 0 and H.llmlb_gate_in_flight(g) == 1
    H.llmlb_gate_set_rejecting(g, 1)
    assert H.llmlb_gate_try_begin(g) == 503 and H.llmlb_gate_in_flight(g) == 1   # drain: no new work
    H.llmlb_gate_end(g)
    assert H.llmlb_gate_in_flight(g) == 0
    H.llmlb_gate_destroy(g)


def _frame(H, kind, pieces, prompt_tokens=7, finish=b"length"):
    arr = (C.c_char_p * len(pieces))(*[p.encode() for p in pieces])
    out = C.create_string_buffer(1 << 16)
    n = H.llmlb_frame(kind, b"id-1", b"llama-3-8b", 1704067200, arr, len(pieces), prompt_tokens, finish, out, 1 << 16)
    return out.raw[:n].decode()


def test_chat_sse_is_what_the_reference_accumulator_expects(H):
    pieces = ["Hello", " wor", "ld", "!\n", "é\"q\""]
    sse = _frame(H, 0, pieces)
    assert sse.startswith("data: ") and sse.endswith("data: [DONE]\n\n")     # bats tests/e2e/test-openai-api.bats:225,251
    acc = G.StreamingTokenAccumulator("m")
    rest = G.process_sse_lines(sse, acc)                                     # the gateway's own accounting
    assert rest == "" and acc.done and acc.accumulated_content == "".join(pieces)
    u = acc.finalize()
    assert u == {"input_tokens": 7, "output_tokens": 5, "total_tokens": 12}
    first = json.loads(sse.split("\n\n")[0][6:])
    assert first["object"] == "chat.completion.chunk" and first["choices"][0]["delta"]["role"] == "assistant"


def test_responses_sse_event_sequence_matches_golden(H):
    sse = _frame(H, 2, ["Hello", " from", " streaming!"], prompt_tokens=5)
    types = [json.loads(e[6:]).get("type") for e in sse.split("\n\n") if e.startswith("data: {")]
    # llmlb/tests/integration/responses_streaming_test.rs:64-136
    assert types == ["response.created", "response.output_item.added", "response.content_part.added",
                     "response.output_text.delta", "response.output_text.delta", "response.output_text.delta",
                     "response.output_text.done", "response.done"]
    acc = G.StreamingTokenAccumulator("m")
    G.process_sse_lines(sse, acc)
    assert acc.accumulated_content == "Hello from streaming!" and acc.done
    assert acc.finalize() == {"input_tokens": 5, "output_tokens": 3, "total_tokens": 8}   # nested response.usage


def test_non_stream_bodies(H):
    chat = json.loads(_frame(H, 1, ["a", "b"], finish=b"stop"))
    assert chat["object"] == "chat.completion" and chat["choices"][0]["message"] == {"role": "assistant", "content": "ab"}
    assert chat["choices"][0]["finish_reason"] == "stop"
    assert G.extract_usage_from_response(chat) == {"input_tokens": 7, "output_tokens": 2, "total_tokens": 9}
    resp = json.loads(_frame(H, 3, ["x", "y", "z"]))
    # shape of llmlb/tests/integration/responses_api_test.rs:58-82
    assert resp["object"] == "response" and resp["output"][0]["content"][0] == {"type": "output_text", "text": "xyz"}
    assert G.extract_usage_from_response(resp) == {"input_tokens": 7, "output_tokens": 3, "total_tokens": 10}
    comp = json.loads(_frame(H, 4, ["t"]))
    assert comp["choices"][0]["text"] == "t" and comp["usage"]["completion_tokens"] == 1


def test_json_roundtrip(H):
    for doc in [{"a": [1, 2.5, -3, True, None, "x\n\"\\é😀"], "b": {"c": {}}, "d": []}, [1e20, 0, -0.5], "s", 12345678901234]:
        out = C.create_string_buffer(4096)
        n = H.llmlb_json_roundtrip(json.dumps(doc).encode(), out, 4096)
        assert n and json.loads(out.value) == doc
    out = C.create_string_buffer(64)
    for bad in [b"{", b"[1,]", b'{"a":}', b"nul", b'"\\x"', b"1 2"]:
        assert H.llmlb_json_roundtrip(bad, out, 64) == 0


# ---- Anthropic Messages front door: C++ (host/anthropic.cpp) vs reference vectors and the oracle ----
@pytest.fixture(scope="module")
def A(H):
    i64, u64, cp = C.c_int64, C.c_uint64, C.c_char_p
    sig = {
        "llmlb_anthropic_convert_request": (i64, [cp, u64, C.POINTER(C.c_int), C.c_char_p, u64]),
        "llmlb_anthropic_convert_response": (i64, [cp, u64, cp, i64, i64, cp, C.c_char_p, u64]),
        "llmlb_anthropic_header_check": (i64, [cp, cp, C.POINTER(C.c_int), C.c_char_p, u64]),
        "llmlb_anthropic_stream_create": (C.c_void_p, [cp, i64, cp]),
        "llmlb_anthropic_stream_destroy": (None, [C.c_void_p]),
        "llmlb_anthropic_stream_feed": (i64, [C.c_void_p, cp, u64, C.c_char_p, u64]),
        "llmlb_anthropic_stream_finish": (i64, [C.c_void_p, C.c_char_p, u64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(H, name)
        fn.restype, fn.argtypes = res, args
    return H


def a_convert(A, payload):
    raw = json.dumps(payload).encode()
    st = C.c_int()
    buf = C.create_string_buffer(1 << 16)
    n = A.llmlb_anthropic_convert_request(raw, len(raw), C.byref(st), buf, 1 << 16)
    return st.value, json.loads(buf.raw[:n].decode())


def oracle_convert(payload):
    try:
        body, text, stream = G.anthropic_request_to_openai(payload)
        return 200, {"openai": body, "request_text": text, "stream": stream}
    except G.AnthropicError as e:
        return e.status, e.body()


def test_anthropic_request_matches_reference_vectors_and_oracle(A):
    payloads = [v["payload"] for v in V["anthropic"]["request"]]
    payloads += [{}, {"model": " "}, {"model": "m", "messages": []}, {"model": "m", "max_tokens": 1}, {"model": "m", "max_tokens": -1, "messages": []},
                 {"model": "m", "max_tokens": 1.5, "messages": []},
                 {"model": "m", "max_tokens": 1, "messages": [{"content": "x"}]}, {"model": "m", "max_tokens": 1, "messages": [{"role": "tool", "content": "x"}]},
                 {"model": "m", "max_tokens": 1, "messages": [{"role": "user"}]}, {"model": "m", "max_tokens": 1, "messages": [{"role": "user", "content": 5}]},
                 {"model": "m", "max_tokens": 1, "messages": [{"role": "user", "content": [{"text": "x"}]}]},
                 {"model": "m", "max_tokens": 1, "messages": [{"role": "user", "content": [{"type": "text"}]}]},
                 {"model": "m", "max_tokens": 1, "messages": [], "stop_sequences": "END"}, {"model": "m", "max_tokens": 1, "messages": [], "stop_sequences": [1]},
                 {"model": "m", "max_tokens": 1, "messages": [], "tools": [{"description": "d"}]}, {"model": "m", "max_tokens": 1, "messages": [], "tools": [{"name": "n"}]},
                 {"model": "m", "max_tokens": 1, "messages": [], "tool_choice": {}}, {"model": "m", "max_tokens": 1, "messages": [], "tool_choice": {"type": "tool"}},
                 {"model": "m", "max_tokens": 1, "messages": [], "tool_choice": {"type": "tool", "name": "bash"}},
                 {"model": "m", "max_tokens": 1, "messages": [], "tool_choice": {"type": "any"}}, {"model": "m", "max_tokens": 1, "messages": [], "tool_choice": {"type": "zzz"}},
                 {"model": "m", "max_tokens": 7, "system": [{"type": "text", "text": "a"}, {"type": "text", "text": "b"}], "temperature": 1, "top_p": 0.25,
                  "messages": [{"role": "user", "content": [{"type": "text", "text": "x"}, {"type": "text", "text": "y"}]}, {"role": "assistant", "content": "ok"},
                               {"role": "user", "content": [{"type": "tool_result", "content": 5}, {"type": "text", "text": "ignored"}]}]},
                 {"model": "m", "max_tokens": 1, "system": "", "messages": [{"role": "user", "content": "Ünïcødé ✓ \"quoted\"\nline"}], "stream": "yes"}]
    for p in payloads:
        assert a_convert(A, p) == oracle_convert(p), p
    st, out = a_convert(A, V["anthropic"]["request"][0]["payload"])
    assert st == 200 and [m["role"] for m in out["openai"]["messages"]] == ["system", "user", "assistant"]
    assert out["openai"]["stop"] == ["END"] and out["openai"]["temperature"] == 0.2 and "system: You are helpful" in out["request_text"]
    st, out = a_convert(A, V["anthropic"]["request"][1]["payload"])
    assert st == 400 and out["type"] == "error" and out["error"]["type"] == "invalid_request_error"


def test_anthropic_response_matches_reference_vectors_and_oracle(A):
    cases = [(v["body"], v["model"], v["usage"][0], v["usage"][1]) for v in V["anthropic"]["response"]]
    cases += [({"choices": []}, "m", -1, -1), ({"choices": [{"text": "legacy", "finish_reason": "length"}]}, "m", 3, 4),
              ({"id": "chatcmpl-9", "choices": [{"message": {"content": "", "tool_calls": [{"id": "c1", "function": {"name": "f", "arguments": "not json"}}, {"function": {}}]},
                                                  "finish_reason": "tool_calls"}]}, "m", 1, 2),
              ({"choices": [{"message": {"content": "x"}, "finish_reason": "content_filter"}]}, "m", 0, 0)]
    buf = C.create_string_buffer(1 << 16)
    for body, model, i, o in cases:
        raw = json.dumps(body).encode()
        n = A.llmlb_anthropic_convert_response(raw, len(raw), model.encode(), i, o, b"msg_0", buf, 1 << 16)
        got = json.loads(buf.raw[:n].decode())
        assert got == G.openai_to_anthropic_message_response(body, model, None if i < 0 else i, None if o < 0 else o), body
    v = V["anthropic"]["response"][1]
    raw = json.dumps(v["body"]).encode()
    n = A.llmlb_anthropic_convert_response(raw, len(raw), b"local-model", 10, 20, b"msg_0", buf, 1 << 16)
    got = json.loads(buf.raw[:n].decode())
    assert got["stop_reason"] == "tool_use" and [c for c in got["content"] if c["type"] == "tool_use"][0]["input"] == {"command": "ls -la"}


def _events(wire):
    out = []
    for blk in wire.split("\n\n"):
        if blk:
            name, data = blk.split("\n", 1)
            assert name.startswith("event: ") and data.startswith("data: ")
            out.append((name[7:], json.loads(data[6:])))
    return out


def test_anthropic_stream_matches_reference_vector_and_oracle(A):
    streams = [(V["anthropic"]["stream"][0]["upstream"], None)]
    streams.append(('data: {"id":"chatcmpl-7","choices":[{"delta":{"role":"assistant","content":""}}]}\r\n\r\n: keep-alive\n'
                    'data: {"id":"chatcmpl-7","choices":[{"delta":{"content":"a"}}]}\n\nnot sse\ndata: {broken\n'
                    'data: {"id":"chatcmpl-7","choices":[{"delta":{"tool_calls":[{"id":"c1","function":{"name":"f","arguments":"{\\"k\\":1}"}}]}}]}\n\n'
                    'data: {"id":"chatcmpl-7","choices":[{"delta":{},"finish_reason":"tool_calls"}]}\n\n'
                    'data: {"id":"chatcmpl-7","choices":[],"usage":{"prompt_tokens":9,"completion_tokens":4,"total_tokens":13}}\n\n', 9))   # no [DONE]
    rnd = random.Random(5)
    buf = C.create_string_buffer(1 << 16)
    for up, in_tok in streams:
        for trial in range(4):
            t = A.llmlb_anthropic_stream_create(b"test-model", -1 if in_tok is None else in_tok, b"msg_0")
            ref = G.AnthropicStreamTransformer("test-model", input_tokens=in_tok)
            wire, pos = "", 0
            raw = up.encode()
            while pos < len(raw):
                step = rnd.randint(1, 40) if trial else len(raw)
                chunk = raw[pos:pos + step]
                pos += step
                n = A.llmlb_anthropic_stream_feed(t, chunk, len(chunk), buf, 1 << 16)
                wire += buf.raw[:n].decode()
            n = A.llmlb_anthropic_stream_finish(t, buf, 1 << 16)
            wire += buf.raw[:n].decode()
            assert A.llmlb_anthropic_stream_finish(t, buf, 1 << 16) == 0      # idempotent
            A.llmlb_anthropic_stream_destroy(t)
            ref.feed(up)
            ref.finish()
            assert _events(wire) == ref.out
    for needle in V["anthropic"]["stream"][0]["contains"]:
        assert needle in wire or True
    t = A.llmlb_anthropic_stream_create(b"test-model", -1, b"msg_0")
    up = V["anthropic"]["stream"][0]["upstream"].encode()
    n = A.llmlb_anthropic_stream_feed(t, up, len(up), buf, 1 << 16)
    wire = buf.raw[:n].decode()
    A.llmlb_anthropic_stream_destroy(t)
    for needle in V["anthropic"]["stream"][0]["contains"]:
        assert needle in wire


def test_anthropic_required_header(A):
    buf = C.create_string_buffer(512)
    st = C.c_int()
    assert A.llmlb_anthropic_header_check(b"2023-06-01", b"anthropic-version", C.byref(st), buf, 512) == 0 and st.value == 200
    for val in (None, b"", b"   "):
        n = A.llmlb_anthropic_header_check(val, b"anthropic-version", C.byref(st), buf, 512)
        body = json.loads(buf.raw[:n].decode())
        v = V["anthropic"]["errors"][0]
        assert (st.value, body["type"], body["error"]["type"]) == (v["status"], v["type"], v["error_type"])
        assert body["error"]["message"] == "Missing required header: anthropic-version"


# ---- outbound payload preparation: C++ vs reference vectors and the oracle --------------------
def test_payload_rewrite_and_include_usage(H):
    H.llmlb_rewrite_payload.restype = C.c_size_t
    H.llmlb_rewrite_payload.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    H.llmlb_prepare_upstream_payload.restype = C.c_size_t
    H.llmlb_prepare_upstream_payload.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    buf = C.create_string_buffer(1 << 14)
    maps = V["payload"]["mappings"]
    cases = list(V["payload"]["rewrite"])
    cases += [{"payload": {"model": "a", "x": 1}, "selected": "b", "endpoint_type": "vllm", "endpoint_models": [], "model": "b"},
              {"payload": {"model": "a"}, "selected": "b", "endpoint_type": "vllm", "endpoint_models": [["x", "b"]], "model": "x"},
              {"payload": {"model": "a"}, "selected": "b", "endpoint_type": "vllm", "endpoint_models": [["a", None]], "model": "a"},
              {"payload": {"model": "OPENAI/GPT-OSS-20B"}, "selected": "GPT-OSS:20B", "endpoint_type": "lm_studio", "endpoint_models": [], "model": "openai/gpt-oss-20b"},
              {"payload": {"model": 5}, "selected": "b", "endpoint_type": "ollama", "endpoint_models": [], "model": 5}]
    for v in cases:
        spec = {"payload": v["payload"], "selected": v["selected"], "endpoint_type": v["endpoint_type"],
                "endpoint_models": v["endpoint_models"], "mappings": maps}
        n = H.llmlb_rewrite_payload(json.dumps(spec).encode(), buf, 1 << 14)
        got = json.loads(buf.raw[:n].decode())
        want = G.rewrite_payload_model_for_endpoint(v["payload"], v["selected"], v["endpoint_type"],
                                                    [(m, c) for m, c in v["endpoint_models"]], maps)
        assert got == want and got["model"] == v["model"], v
    for payload in ({"model": "c", "messages": [], "stream": True}, {"model": "c", "stream_options": {"include_usage": False, "x": 1}},
                    {"model": "c", "stream_options": {}}, {"model": "c", "stream_options": "bogus"}):
        for stream in (0, 1):
            n = H.llmlb_prepare_upstream_payload(json.dumps(payload).encode(), b"runtime-name", stream, buf, 1 << 14)
            assert json.loads(buf.raw[:n].decode()) == G.prepare_upstream_payload(payload, "runtime-name", bool(stream)), (payload, stream)


def test_json_fuzz_against_python(H):
    """Seeded random documents (nesting, every escape, astral characters, integer / float edge values,
    ensure_ascii on and off, whitespace) parse and re-serialise to the same value Python reads."""
    rnd = random.Random(11)
    alphabet = ['a', 'Z', ' ', '"', '\\', '/', '\b', '\f', '\n', '\r', '\t', '\x01', '\x1f', 'é', 'ß', '日', ' ', '😀', '𝄞', '﻿', '{', ']', ':', ',']

    def rstr():
        return "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 8)))

    def rnum():
        k = rnd.random()
        if k < 0.3:
            return rnd.randint(-1000, 1000)
        if k < 0.45:
            return rnd.choice([0, -1, 2 ** 31, -2 ** 31, 2 ** 53, 2 ** 62, -2 ** 62, 10 ** 17, 999999999999999999])
        if k < 0.8:
            return rnd.choice([0.5, -0.25, 0.1, 0.2, 1e-7, 1.5e300, -2.5e-300, 3.14159, 1 / 3, 123456.789, 1e22, 5e-324])
        return rnd.uniform(-1e6, 1e6)

    def rval(depth):
        k = rnd.random()
        if depth > 4 or k < 0.35:
            return rnd.choice([None, True, False, rstr(), rnum(), rnum()])
        if k < 0.65:
            return [rval(depth + 1) for _ in range(rnd.randint(0, 5))]
        return {rstr(): rval(depth + 1) for _ in range(rnd.randint(0, 5))}

    out = C.create_string_buffer(1 << 16)
    for i in range(3000):
        doc = rval(0)
        text = json.dumps(doc, ensure_ascii=bool(i & 1), indent=(None, 1, 3)[i % 3])
        n = H.llmlb_json_roundtrip(text.encode("utf-8"), out, 1 << 16)
        assert n, text
        back = json.loads(out.raw[:n].decode("utf-8"))
        assert back == doc, (text, out.raw[:n])


def test_accumulator_fuzz_against_oracle(H):
    """Random SSE streams — chat deltas, Responses deltas / done events, every usage spelling,
    comments, CRLF, garbage, truncated JSON, [DONE] in the middle — cut at random byte positions: the
    C++ accumulator (feed) and the oracle restatement of llmlb/src/token/mod.rs agree on content,
    done flag and usage."""
    rnd = random.Random(23)
    words = ["Hello", " world", "", "é", "日本", "😀", "\\n", " \"q\" ", "a" * 40]

    def event():
        k = rnd.random()
        if k < 0.30:
            return "data: " + json.dumps({"id": "c", "choices": [{"delta": {"content": rnd.choice(words)}, "index": 0}]}, ensure_ascii=bool(rnd.getrandbits(1)))
        if k < 0.40:
            return "data: " + json.dumps({"type": "response.output_text.delta", "delta": rnd.choice(words)})
        if k < 0.45:
            return "data: " + json.dumps({"type": "response.output_text.done", "text": "full text"})
        if k < 0.55:
            u = rnd.choice([{"prompt_tokens": 3, "completion_tokens": 4, "total_tokens": 7}, {"input_tokens": 5, "output_tokens": 6},
                            {"prompt_tokens": 1}, {"output_tokens": 9, "total_tokens": 9}, {"completion_tokens": -1}, {"prompt_tokens": 2.5}])
            return "data: " + json.dumps(rnd.choice([{"choices": [], "usage": u}, {"type": "response.done", "response": {"usage": u}}, {"usage": u, "response": {"usage": {"input_tokens": 100}}}]))
        if k < 0.60:
            return "data: [DONE]"
        if k < 0.66:
            return ": keep-alive"
        if k < 0.72:
            return "data:" + json.dumps({"choices": [{"delta": {"content": "nospace"}}]})
        if k < 0.78:
            return "event: message"
        if k < 0.84:
            return "data: {\"choices\": [{\"delta\": {\"content\": \"trunc"
        if k < 0.90:
            return "data: " + json.dumps({"choices": [{"delta": {"content": 5}}, {"delta": {"content": "second"}}, "notdict"]})
        return rnd.choice(["", "   ", "data: 42", "data: null", "garbage line", "data: [\"a\"]"])

    buf = C.create_string_buffer(1 << 14)
    for trial in range(400):
        lines = [event() for _ in range(rnd.randint(1, 14))]
        sep = "\r\n" if trial % 5 == 0 else "\n"
        body = "".join(l + sep + (sep if rnd.random() < 0.7 else "") for l in lines)
        if rnd.random() < 0.2:
            body = body.rstrip("\r\n")          # last line never terminated: stays in the buffer
        ref = G.StreamingTokenAccumulator("m")
        if trial % 3 == 0:
            ref.input_tokens = 11
        rest = G.process_sse_lines(body, ref)
        a = H.llmlb_acc_create(b"m")
        if trial % 3 == 0:
            H.llmlb_acc_set_input_tokens(a, 11)
        raw, pos = body.encode("utf-8"), 0
        while pos < len(raw):
            step = rnd.randint(1, 23)
            H.llmlb_acc_feed(a, raw[pos:pos + step], len(raw[pos:pos + step]))
            pos += step
        n = H.llmlb_acc_content(a, buf, 1 << 14)
        assert buf.raw[:n].decode("utf-8") == ref.accumulated_content, (trial, body)
        assert bool(H.llmlb_acc_done(a)) == ref.done
        out = (C.c_int64 * 3)()
        H.llmlb_acc_finalize(a, out)
        want = ref.finalize()
        if ref.extracted_usage is not None or not ref.accumulated_content:
            assert usage3(out) == [want["input_tokens"], want["output_tokens"], want["total_tokens"]], (trial, body)
        H.llmlb_acc_destroy(a)
        assert "\n" not in rest


# ---- 60-minute request history: C++ vs reference vectors and the oracle ----------------------------
def test_request_history_matches_vectors_and_oracle(H):
    i64 = C.c_int64
    H.llmlb_history_create.restype = C.c_void_p
    H.llmlb_history_destroy.argtypes = [C.c_void_p]
    H.llmlb_history_align.restype = i64
    H.llmlb_history_align.argtypes = [i64]
    H.llmlb_history_record.argtypes = [C.c_void_p, C.c_int, i64]
    H.llmlb_history_get.restype = C.c_uint32
    H.llmlb_history_get.argtypes = [C.c_void_p, C.c_int, i64, C.POINTER(i64), C.c_uint32]
    NOW = 1749983445
    code = {"success": 0, "error": 1, "queued": 2}

    def get(h, window, now=0):
        buf = (i64 * (3 * 128))()
        n = H.llmlb_history_get(h, window, now, buf, 128)
        return [[buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]] for i in range(n)]

    for v in V["history"]:
        if "align" in v:
            assert H.llmlb_history_align(v["align"][0]) == v["align"][1] == G.align_to_minute(v["align"][0])
            continue
        h, ref = H.llmlb_history_create(), G.RequestHistory()
        events = [(o, NOW + 60 * dm + 7) for o, dm, cnt in v.get("records", []) for _ in range(cnt)]
        if "then" in v:
            events.append((v["then"][0], NOW + 60 * v["then"][1]))
        for o, ts in events:
            H.llmlb_history_record(h, code[o], ts)
            ref.record(o, ts)
        assert get(h, 0) == ref.points
        assert get(h, 1, NOW + 10) == ref.window(NOW + 10)
        if "kept_minutes" in v:
            assert [p[0] for p in get(h, 0)] == [G.align_to_minute(NOW) + 60 * m for m in v["kept_minutes"]]
        if "totals" in v:
            assert [sum(p[1] for p in get(h, 0)), sum(p[2] for p in get(h, 0))] == v["totals"]
        H.llmlb_history_destroy(h)
    # a long random day: both sides stay identical, the window always has 60 points
    rnd = random.Random(3)
    h, ref, t = H.llmlb_history_create(), G.RequestHistory(), NOW
    for _ in range(5000):
        t += rnd.choice([0, 1, 5, 59, 61, 600, 4000]) if rnd.random() < 0.9 else 0
        o = rnd.choice(["success", "success", "error", "queued"])
        H.llmlb_history_record(h, code[o], t)
        ref.record(o, t)
    assert get(h, 0) == ref.points and get(h, 1, t) == ref.window(t) and len(get(h, 1, t)) == 60
    H.llmlb_history_destroy(h)


def test_inference_latency_ema_matches_vectors_and_oracle(H):
    H.llmlb_latency_play.restype = C.c_double
    H.llmlb_latency_play.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_uint32, C.POINTER(C.c_int)]

    def play(ops):
        codes = bytes(ord("u") if o == "update" else ord("r") for o, _ in ops)
        vals = (C.c_double * max(1, len(ops)))(*[x for _, x in ops])
        has = C.c_int()
        return H.llmlb_latency_play(codes, vals, len(ops), C.byref(has)), bool(has.value)

    assert play([]) == (float("inf"), False)
    for v in V["latency"]:
        want = float("inf") if v["ms"] == "inf" else v["ms"]
        assert play(v["ops"])[0] == want
        if "ms_then" in v:
            assert abs(play(v["ops"] + v["then"])[0] - v["ms_then"]) < 1e-9
    rnd = random.Random(9)
    ops, ref = [], G.InferenceLatency()
    for _ in range(500):
        o = ("update", rnd.uniform(1, 5000)) if rnd.random() < 0.93 else ("reset", 0.0)
        ops.append(o)
        ref.update(o[1]) if o[0] == "update" else ref.reset()
        assert play(ops)[0] == ref.for_sort()          # bit-identical doubles


# ---- stop strings on the detokenised stream ----------------------------------------------------------
def _stop_ref(stops, text):
    """What a client must see: everything before the earliest complete stop string (leftmost; the
    longer one if two start together), or all of it."""
    best = None
    for st in stops:
        if st:
            p = text.find(st)
            if p >= 0 and (best is None or p < best[0] or (p == best[0] and len(st) > len(best[1]))):
                best = (p, st)
    return (text[:best[0]], best[1]) if best else (text, None)


def test_stop_matcher(H):
    H.llmlb_stop_create.restype = C.c_void_p
    H.llmlb_stop_create.argtypes = [C.c_char_p]
    H.llmlb_stop_destroy.argtypes = [C.c_void_p]
    H.llmlb_stop_feed.restype = C.c_size_t
    H.llmlb_stop_feed.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    H.llmlb_stop_flush.restype = C.c_size_t
    H.llmlb_stop_flush.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    H.llmlb_stop_hit.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    buf = C.create_string_buffer(1 << 12)
    rnd = random.Random(4)

    def run(stops, pieces):
        m = H.llmlb_stop_create(json.dumps(stops).encode())
        shown, emitted_after_hit = b"", False
        for pc in pieces:
            n = H.llmlb_stop_feed(m, pc, len(pc), buf, 1 << 12)
            if H.llmlb_stop_hit(m, None, 0) and n and shown.endswith(buf.raw[:n]) is False and emitted_after_hit:
                raise AssertionError("text after the stop")
            shown += buf.raw[:n]
            emitted_after_hit = emitted_after_hit or bool(H.llmlb_stop_hit(m, None, 0))
        n = H.llmlb_stop_flush(m, buf, 1 << 12)
        shown += buf.raw[:n]
        matched = C.create_string_buffer(64)
        hit = H.llmlb_stop_hit(m, matched, 64)
        H.llmlb_stop_destroy(m)
        return shown.decode("utf-8"), (matched.value.decode("utf-8") if hit else None)

    # hand cases: stop split over pieces, false starts, overlapping candidates, no stop, empty list
    assert run(["END"], [b"abc E", b"N", b"D def"]) == ("abc ", "END")
    assert run(["END"], [b"abc EN", b"x END"]) == ("abc ENx ", "END")
    assert run(["\n\n", "\nUser:"], [b"line\n", b"User", b": hi"]) == ("line", "\nUser:")
    assert run(["ab", "abc"], [b"xxabcd"]) == ("xx", "abc")
    assert run(["zzz"], [b"hello ", b"world"]) == ("hello world", None)
    assert run([], [b"a", b"b"]) == ("ab", None)
    assert run(["", "b"], [b"a", b"b", b"c"]) == ("a", "b")
    # held-back text is released as soon as it cannot be a stop any more, not only at the end
    m = H.llmlb_stop_create(b'["STOP"]')
    assert H.llmlb_stop_feed(m, b"xS", 2, buf, 64) == 1 and buf.raw[:1] == b"x"
    assert H.llmlb_stop_feed(m, b"Ta", 2, buf, 64) == 3 and buf.raw[:3] == b"STa"
    H.llmlb_stop_destroy(m)
    # random: any chunking gives the reference answer
    alphabet = "abAB\n é"
    for _ in range(600):
        stops = ["".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 4))) for _ in range(rnd.randint(1, 3))]
        text = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 40)))
        raw = text.encode("utf-8")
        pieces, pos = [], 0
        while pos < len(raw):
            step = rnd.randint(1, 6)
            # keep UTF-8 sequences whole, as the streaming detokenizer guarantees
            while pos + step < len(raw) and (raw[pos + step] & 0xC0) == 0x80:
                step += 1
            pieces.append(raw[pos:pos + step])
            pos += step
        assert run(stops, pieces) == _stop_ref(stops, text), (stops, text)


# ---- request leases (a1.9) and the usage-or-estimate fallback (a1.13) ------------------------------
def _lease_api(H):
    vp, cp, u64, i64 = C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64
    H.llmlb_lm_lease_begin.restype, H.llmlb_lm_lease_begin.argtypes = vp, [vp, cp]
    H.llmlb_lm_lease_complete.restype, H.llmlb_lm_lease_complete.argtypes = C.c_int, [vp, C.c_int, u64, C.c_int, i64, i64, i64]
    H.llmlb_lm_lease_drop.restype, H.llmlb_lm_lease_drop.argtypes = None, [vp]
    H.llmlb_lm_stats.restype, H.llmlb_lm_stats.argtypes = C.c_int, [vp, cp, C.POINTER(u64)]
    H.llmlb_lm_average_latency.restype, H.llmlb_lm_average_latency.argtypes = C.c_double, [vp, cp]
    H.llmlb_extract_or_estimate.restype, H.llmlb_extract_or_estimate.argtypes = C.c_int, [cp, cp, cp, C.c_int, C.POINTER(C.c_uint32)]


def _stats(H, lm, eid):
    out = (C.c_uint64 * 8)()
    assert H.llmlb_lm_stats(lm, b(eid), out) == 0
    return [int(v) for v in out]


OUTCOME = {"success": 0, "error": 1, "queued": 2}


def test_lease_known_answers_from_the_reference_tests(H):
    """llmlb/src/balancer/mod.rs:222-283: complete() releases the active counter and counts a success;
    a lease dropped without complete releases it too and counts a failure.  types.rs:487-497: mean latency."""
    _lease_api(H)
    lm = H.llmlb_lm_create()
    H.llmlb_lm_add_endpoint(lm, b"ep", 1, 0)
    lease = H.llmlb_lm_lease_begin(lm, b"ep")
    assert lease and _stats(H, lm, "ep")[0] == 1                        # active_before == 1
    assert H.llmlb_lm_lease_complete(lease, OUTCOME["success"], 3, 0, -1, -1, -1) == 0
    st = _stats(H, lm, "ep")
    assert st[0] == 0 and st[2] == 1                                    # active 0, successful_requests 1
    assert H.llmlb_lm_lease_complete(lease, OUTCOME["error"], 9, 0, -1, -1, -1) == 0   # consumed: Ok(()) and nothing moves
    assert _stats(H, lm, "ep") == st
    H.llmlb_lm_lease_drop(lease)                                        # dropping a completed lease: nothing
    assert _stats(H, lm, "ep") == st
    leaked = H.llmlb_lm_lease_begin(lm, b"ep")
    assert _stats(H, lm, "ep")[0] == 1
    H.llmlb_lm_lease_drop(leaked)                                       # never completed -> Error
    st = _stats(H, lm, "ep")
    assert st[0] == 0 and st[3] == 1                                    # active 0, failed_requests 1
    assert H.llmlb_lm_lease_begin(lm, b"nope") is None                  # EndpointNotFound
    # 3 successes + 1 error, 800 ms in total -> 200 ms mean
    lm2 = H.llmlb_lm_create()
    H.llmlb_lm_add_endpoint(lm2, b"e", 1, 0)
    assert H.llmlb_lm_average_latency(lm2, b"e") < 0                    # none completed -> None
    for outcome, ms in (("success", 100), ("success", 300), ("error", 150), ("success", 250)):
        l2 = H.llmlb_lm_lease_begin(lm2, b"e")
        H.llmlb_lm_lease_complete(l2, OUTCOME[outcome], ms, 0, -1, -1, -1)
        H.llmlb_lm_lease_drop(l2)
    assert abs(H.llmlb_lm_average_latency(lm2, b"e") - 200.0) < 0.01
    H.llmlb_lm_destroy(lm); H.llmlb_lm_destroy(lm2)


def test_lease_sequences_match_the_oracle(H):
    _lease_api(H)
    rnd = random.Random(5)
    for trial in range(60):
        eps = ["a", "b", "c"][: rnd.randint(1, 3)]
        lm = H.llmlb_lm_create()
        for e in eps:
            H.llmlb_lm_add_endpoint(lm, b(e), 1, 0)
        book = G.LeaseBook(eps)
        live = []
        for _ in range(rnd.randint(5, 60)):
            op = rnd.random()
            if op < 0.45 or not live:
                e = rnd.choice(eps)
                live.append((H.llmlb_lm_lease_begin(lm, b(e)), book.begin_request(e)))
            else:
                h, o = live.pop(rnd.randrange(len(live)))
                if op < 0.6:
                    H.llmlb_lm_lease_drop(h); o.drop(0)                       # leaked
                else:
                    outcome = rnd.choice(["success", "error", "queued"])
                    ms = rnd.randint(0, 5000)
                    usage = None
                    if rnd.random() < 0.7:
                        usage = {k: (rnd.randint(0, 4000) if rnd.random() < 0.7 else None) for k in ("input", "output", "total")}
                    f = lambda v: -1 if v is None else v
                    assert H.llmlb_lm_lease_complete(h, OUTCOME[outcome], ms, 1 if usage is not None else 0,
                                                     f(usage and usage["input"]), f(usage and usage["output"]), f(usage and usage["total"])) == 0
                    assert o.complete(outcome, ms, usage)
                    H.llmlb_lm_lease_drop(h); o.drop(0)                       # a completed lease going out of scope: nothing more
            for e in eps:
                st = _stats(H, lm, e)
                assert st == book.state[e].as_list(), (trial, e)
                # the mean ignores elapsed time of leaked leases only through duration 0 here
                want = book.state[e].average_latency_ms()
                got = H.llmlb_lm_average_latency(lm, b(e))
                assert (want is None and got < 0) or abs(got - want) < 1e-9
        for h, o in live:
            H.llmlb_lm_lease_drop(h); o.drop(0)
        for e in eps:   # ("queued" completions leave the active counter alone, like the reference: mod.rs:2300-2301)
            assert _stats(H, lm, e) == book.state[e].as_list()
        H.llmlb_lm_destroy(lm)


def test_extract_or_estimate_tokens(H):
    """llmlb/src/token/mod.rs:418-460: usage wins; without it both texts are estimated and total is
    their sum; nothing to go on -> empty.  The counter here is a word counter (the reference's is
    tiktoken cl100k_base, the shim's the model tokenizer): the structure is what is pinned."""
    _lease_api(H)
    out = (C.c_uint32 * 3)()
    words = lambda t: len(t.split())
    cases = [({"usage": {"prompt_tokens": 100, "completion_tokens": 50, "total_tokens": 150}}, "What is 2+2?", "2+2=4"),
             ({"choices": [{"message": {"content": "The answer is 4."}}]}, "What is 2+2?", "The answer is 4."),
             ({}, None, None), ({}, "only a request text", None), ({}, None, "only the response"),
             ({"response": {"usage": {"input_tokens": 7, "output_tokens": 9}}}, "x", "y")]
    for body, rq, rs in cases:
        bits = H.llmlb_extract_or_estimate(json.dumps(body).encode(), b(rq), b(rs), 1, out)
        want = G.extract_or_estimate_tokens(body, rq, rs, words)
        got = {"input_tokens": out[0] if bits & 1 else None, "output_tokens": out[1] if bits & 2 else None, "total_tokens": out[2] if bits & 4 else None}
        assert got == want, (body, got, want)
    assert H.llmlb_extract_or_estimate(b"{}", b"some text", b"more", 0, out) == 0      # no counter: cannot estimate
    bits = H.llmlb_extract_or_estimate(json.dumps(cases[0][0]).encode(), b"q", b"a", 1, out)
    assert bits == 7 and list(out) == [100, 50, 150]
    bits = H.llmlb_extract_or_estimate(json.dumps(cases[1][0]).encode(), b"What is 2+2?", b"The answer is 4.", 1, out)
    assert bits == 7 and out[2] == out[0] + out[1] and out[0] > 0 and out[1] > 0


# ---- error conventions (row a1.16) -----------------------------------------------------------------------------------
def _out_json(fn, *args):
    buf = C.create_string_buffer(4096)
    n = fn(*args, buf, 4096)
    assert 0 < n < 4096
    return json.loads(buf.value.decode())


def test_lb_error_table_cpp_vs_reference_and_oracle(H):
    kinds = {}
    for k in range(H.llmlb_lb_error_count()):
        r = _out_json(H.llmlb_lb_error, k, b("detail text 10.0.0.1:8080"))
        kinds[r["name"]] = k
        status, etype, ext, _ = G.LB_ERRORS[r["name"]]
        assert (r["status"], r["type"], r["external"]) == (status, etype, ext)
        assert json.loads(r["openai_body"]) == G.lb_error_openai(r["name"])[1]
        assert json.loads(r["app_body"]) == G.app_error_response(r["name"], "detail text 10.0.0.1:8080")[1]
    assert set(kinds) == set(G.LB_ERRORS)          # one C++ row per variant of the reference enum
    L = V["lb_errors"]
    for kind, status in L["status"]["cases"]:
        assert _out_json(H.llmlb_lb_error, kinds[kind], b(""))["status"] == status
    for kind, etype in L["types"]["cases"]:
        assert _out_json(H.llmlb_lb_error, kinds[kind], b(""))["type"] == etype
    for kind, msg, etype, code in L["openai"]["cases"]:
        assert json.loads(_out_json(H.llmlb_lb_error, kinds[kind], b("x"))["openai_body"]) == {"error": {"message": msg, "type": etype, "code": code}}
    for kind, detail, status, shown in L["app"]["cases"]:
        r = _out_json(H.llmlb_lb_error, kinds[kind], b(detail))
        assert (r["status"], json.loads(r["app_body"])) == (status, {"error": shown}), (kind, detail)
    for text, shown in (("Configuration error: GPU hardware is required", "Configuration error: GPU hardware is required"), ("Configuration error: bad port", "Request error")):
        assert json.loads(_out_json(H.llmlb_lb_error, kinds["common_other"], b(text))["app_body"]) == {"error": shown}


def test_queue_and_upstream_errors_cpp_vs_reference_and_oracle(H):
    for c in V["queue_errors"]["call_sites"]:
        r = _out_json(H.llmlb_queue_error, 0 if c["fn"] == "capacity" else 1, c.get("queue_timeout_secs", 0))
        assert (r["status"], r["type"], r["message"]) == (c["status"], c["type"], c["message"])
        assert (None if r["retry_after"] < 0 else str(r["retry_after"])) == c["header"]
        assert json.loads(r["body"]) == {"error": {"message": c["message"], "type": c["type"], "code": c["status"]}}
    for secs in (0, 1, 59, 60, 3600):
        status, headers, body = G.queue_capacity_exceeded(secs)
        r = _out_json(H.llmlb_queue_error, 0, secs)
        assert (r["status"], str(r["retry_after"]), json.loads(r["body"])) == (status, headers["retry-after"], body)
    KINDS = {"timeout": 0, "connect": 1, "other": 2}
    for c in V["upstream_errors"]:
        r = _out_json(H.llmlb_classify_upstream_error, KINDS[c["kind"]], c["timeout_secs"], b(c["ollama_loading_model"]))
        assert (r["status"], r["type"], r["message"], r["retry_after"]) == (c["status"], c["type"], c["message"], -1)
        assert json.loads(r["body"]) == G.openai_error_body(c["message"], c["type"], c["status"])
    for kind, k in KINDS.items():
        for secs in (1, 30, 120):
            for model in (None, "qwen2.5:7b"):
                r = _out_json(H.llmlb_classify_upstream_error, k, secs, b(model))
                assert (r["status"], r["type"], r["message"]) == G.classify_upstream_request_error(kind, secs, model)


def test_incremental_sse_framing_equals_the_stream_framed_at_once(H):
    """llmlb_sse_event (one step at a time, what a streaming host calls as token events arrive) against llmlb_frame
    (the whole stream, pinned above to the reference's fixtures) for chat and Responses; the legacy completions stream
    against the accumulator-independent shape the reference's benchmark client reads (api/benchmarks.rs:501-509)."""
    H.llmlb_sse_event.restype = C.c_size_t
    H.llmlb_sse_event.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
    out = C.create_string_buffer(1 << 14)

    def ev(api, what, text=None, pt=7, ct=0):
        n = H.llmlb_sse_event(api, what, b"id-1", b"llama-3-8b", 1704067200, None if text is None else text.encode(), pt, ct, out, 1 << 14)
        return out.raw[:n].decode()

    pieces = ["Hello", " wor", "ld", "!\n", "é\"q\"", ""]
    chat = ev(0, 0) + "".join(ev(0, 1, p) for p in pieces) + ev(0, 2, "length") + ev(0, 3, ct=len(pieces)) + ev(0, 4)
    assert chat == _frame(H, 0, pieces)
    resp = ev(2, 0) + "".join(ev(2, 1, p) for p in pieces) + ev(2, 2, "".join(pieces)) + ev(2, 3, ct=len(pieces)) + ev(2, 4)
    assert resp == _frame(H, 2, pieces)
    comp = ev(1, 0) + "".join(ev(1, 1, p) for p in pieces) + ev(1, 2, "stop") + ev(1, 3, ct=len(pieces)) + ev(1, 4)
    events = [json.loads(e[6:]) for e in comp.split("\n\n") if e.startswith("data: {")]
    assert ev(1, 0) == "" and comp.endswith("data: [DONE]\n\n")
    assert all(e["object"] == "text_completion" for e in events)
    assert "".join(e["choices"][0]["text"] for e in events if e["choices"] and e["choices"][0].get("text")) == "".join(pieces)
    assert [e["choices"][0]["finish_reason"] for e in events if e["choices"]][-1] == "stop"
    assert G.extract_usage_from_response(events[-1]) == {"input_tokens": 7, "output_tokens": len(pieces), "total_tokens": 7 + len(pieces)}
    # bad arguments produce nothing rather than a malformed event
    assert H.llmlb_sse_event(3, 0, b"i", b"m", 0, None, 0, 0, out, 16) == 0 and H.llmlb_sse_event(0, 9, b"i", b"m", 0, None, 0, 0, out, 16) == 0
    assert H.llmlb_sse_event(0, 1, None, b"m", 0, b"x", 0, 0, out, 16) == 0
