"""The HTTP shim (llmlb_b200/host/server.cpp) on a machine WITHOUT a GPU: the same protocol tests as tests/test_server_gpu.py
(they check framing, accounting, routing, error conventions, translation — nothing about what the tokens mean), run against
a server binary linked to tests/support/fake_engine.cpp, a scripted token source behind the C ABI.  Test infrastructure only:
the product library and the product server are untouched and still refuse to run without a CUDA device
(tests/test_bench_contract_cpu.py, tests/test_server_cli_cpu.py)."""
import http.client
import json
import os
import subprocess
import sys
import time

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import test_server_gpu as T  # noqa: E402  (its own tests carry the gpu mark; the functions are reused below)

BUILD = os.path.join(HERE, "support", "_build")
FAKE_LIB = os.path.join(BUILD, "libllmlb_b200.so")
FAKE_BIN = os.path.join(BUILD, "llmlb_b200_server_fake_engine")


def _newer(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


@pytest.fixture(scope="module")
def fake_bin():
    os.makedirs(BUILD, exist_ok=True)
    hd = os.path.join(ROOT, "llmlb_b200", "host")
    fake_src = os.path.join(HERE, "support", "fake_engine.cpp")
    hdr = os.path.join(ROOT, "include", "llmlb_b200.h")
    if _newer(FAKE_LIB, [fake_src, hdr]):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread", fake_src, "-o", FAKE_LIB])
    srcs = [os.path.join(hd, f) for f in ("server.cpp", "gateway.cpp", "tokenizer.cpp", "anthropic.cpp", "checkpoint.cpp", "download.cpp")]
    deps = srcs + [FAKE_LIB, hdr] + [os.path.join(hd, f) for f in os.listdir(hd) if f.endswith((".hpp", ".inc"))]
    if _newer(FAKE_BIN, deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-pthread", *srcs, "-o", FAKE_BIN, "-L" + BUILD, "-lllmlb_b200",
                               "-Wl,-rpath," + BUILD])
    return FAKE_BIN


def _start(binary, *args, env=None):
    port = T._free_port()
    proc = subprocess.Popen([binary, "--port", str(port), *args], stderr=subprocess.PIPE, env={**os.environ, **(env or {})})
    deadline = time.time() + 30
    while time.time() < deadline:
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
            return port, proc
        except OSError:
            assert proc.poll() is None, proc.stderr.read().decode()
            time.sleep(0.05)
    proc.kill()
    raise AssertionError("fake-engine server did not come up")


@pytest.fixture(scope="module")
def server(fake_bin):
    port, proc = _start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8", "--max-ctx", "512")
    yield port
    proc.terminate(); proc.wait(timeout=20)


@pytest.fixture(scope="module")
def tok_server(fake_bin):
    port, proc = _start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8", "--max-ctx", "512", "--vocab", "3072",
                        "--tokenizer", os.path.join(HERE, "golden", "tokenizer_llama3_style.json"))
    yield port
    proc.terminate(); proc.wait(timeout=20)


@pytest.fixture()
def built_lib():       # the reused tests only want "something is built"; here that is the fake-engine server
    return None


@pytest.fixture(autouse=True)
def _fake_binary_for_tests_that_start_their_own_server(fake_bin, monkeypatch):
    monkeypatch.setattr(T, "BIN", fake_bin)
    from llmlb_b200 import build
    monkeypatch.setattr(build, "build_host", lambda: None)     # those tests call build_host(): it would need the CUDA library


# ---- the protocol tests of tests/test_server_gpu.py, unchanged --------------------------------------------------------
test_probe_endpoints = T.test_probe_endpoints
test_chat_completion_non_stream = T.test_chat_completion_non_stream
test_chat_completion_stream_accounting = T.test_chat_completion_stream_accounting
test_responses_stream_and_body = T.test_responses_stream_and_body
test_prompt_token_ids_and_completions = T.test_prompt_token_ids_and_completions
test_errors = T.test_errors
test_drain_gate = T.test_drain_gate
test_concurrent_streams = T.test_concurrent_streams
test_chat_through_the_native_tokenizer = T.test_chat_through_the_native_tokenizer
test_stop_at_end_of_turn_token = T.test_stop_at_end_of_turn_token
test_messages_route_non_stream_and_stream = T.test_messages_route_non_stream_and_stream
test_messages_route_errors_in_anthropic_shape = T.test_messages_route_errors_in_anthropic_shape
test_server_from_a_single_gguf = T.test_server_from_a_single_gguf
test_stop_strings_end_the_text_before_the_match = T.test_stop_strings_end_the_text_before_the_match
test_model_download_routes = T.test_model_download_routes
test_an_unmodified_gateway_would_register_and_sync_this_endpoint = T.test_an_unmodified_gateway_would_register_and_sync_this_endpoint


# ---- only reachable with a scripted source: the gateway's queue conventions through the shim ---------------------------
def test_queue_full_is_429_with_retry_after(fake_bin):
    """openai.rs:841-861: capacity exceeded -> 429 rate_limit_exceeded, Retry-After = max(1, queue timeout in seconds)."""
    import threading
    port, proc = _start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "1", "--max-ctx", "4096",
                        "--queue-max", "1", "--queue-timeout-ms", "7000")
    try:
        body = {"model": "tiny-llama", "prompt_token_ids": [5, 6, 7], "max_tokens": 3000, "temperature": 0, "ignore_eos": True}
        res = []
        th = [threading.Thread(target=lambda: res.append(T.call(port, "POST", "/v1/completions", body))) for _ in range(6)]
        [t.start() for t in th]; [t.join() for t in th]
        full = [r for r in res if r[0] == 429]
        assert full and len(full) + sum(r[0] == 200 for r in res) == 6
        st, hdr, d = full[0]
        assert hdr.get("Retry-After") == "7" and json.loads(d) == {"error": {"message": "Request queue is full", "type": "rate_limit_exceeded", "code": 429}}
    finally:
        proc.terminate(); proc.wait(timeout=20)


def test_failures_of_the_engine_reach_the_client_as_the_gateway_would_report_them(fake_bin):
    """The shim maps engine outcomes the way the gateway maps an upstream's:
      deadline        -> 504 "timeout"  "Upstream endpoint request timed out after N seconds"   (openai_util.rs:105-115, types/endpoint.rs:389)
      queue wait      -> 504 "timeout"  "Queue wait timeout"                                     (openai.rs:863-882)
      engine failure  -> 502 "endpoint_request_error" "Failed to proxy request to upstream endpoint" (openai_util.rs:128-134)
    and on a stream whose headers are already out: an in-band error event and NO finish chunk / [DONE]."""
    import threading
    from oracle import gateway_ref as G
    port, proc = _start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "1", "--max-ctx", "4096",
                        "--request-timeout-ms", "1000", "--queue-timeout-ms", "300", env={"FAKE_ENGINE_TOKEN_US": "4000"})
    try:
        slow = {"model": "tiny-llama", "prompt_token_ids": [5, 6, 7], "max_tokens": 2000, "temperature": 0, "ignore_eos": True}
        # deadline, non-stream: 2000 tokens at 4 ms each cannot finish in 1 s
        st, _, d = T.call(port, "POST", "/v1/completions", slow)
        status, etype, msg = G.classify_upstream_request_error("timeout", 1)
        assert (st, json.loads(d)) == (status, G.openai_error_body(msg, etype, status))
        # deadline, streamed: chunks, then the error event, and the stream ends without [DONE]
        st, hdr, d = T.call(port, "POST", "/v1/chat/completions", {"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}],
                                                                     "max_tokens": 2000, "temperature": 0, "ignore_eos": True, "stream": True})
        text = d.decode()
        events = [json.loads(l[6:]) for l in text.split("\n") if l.startswith("data: {")]
        assert st == 200 and "[DONE]" not in text and events[-1] == G.openai_error_body(msg, etype, status)
        assert all("finish_reason" not in c or c["finish_reason"] is None for e in events[:-1] for c in e.get("choices", []))
        # queue wait: one sequence runs (max_seqs 1), the second waits longer than queue_timeout_ms
        res = {}
        t1 = threading.Thread(target=lambda: res.setdefault("a", T.call(port, "POST", "/v1/completions", slow)))
        t1.start(); time.sleep(0.1)
        res["b"] = T.call(port, "POST", "/v1/completions", dict(slow, prompt_token_ids=[9, 9, 9]))
        t1.join()
        st, hdrs, body = G.queue_wait_timeout()
        assert (res["b"][0], json.loads(res["b"][2])) == (st, body) and "Retry-After" not in res["b"][1]
        # engine failure after three tokens
        st, _, d = T.call(port, "POST", "/v1/completions", dict(slow, prompt_token_ids=[666, 666, 1], max_tokens=20))
        status, etype, msg = G.classify_upstream_request_error("other", 0)
        assert (st, json.loads(d)) == (status, G.openai_error_body(msg, etype, status))
    finally:
        proc.terminate(); proc.wait(timeout=20)


def _raw(port, payload, read=True, timeout=5):
    import socket
    s = socket.create_connection(("127.0.0.1", port), timeout=timeout)
    try:
        s.sendall(payload)
        if not read:
            return b""
        out = b""
        while True:
            try:
                d = s.recv(65536)
            except (socket.timeout, ConnectionResetError):
                break
            if not d:
                break
            out += d
        return out
    finally:
        s.close()


def test_malformed_http_is_answered_or_dropped_and_the_server_stays_up(server):
    """What hyper/axum do in front of the reference's handlers: 400 for a request that does not parse, 413 over the
    20 MiB body limit (DefaultBodyLimit, llmlb/src/api/mod.rs:58,536), 431 for an endless header block; random bytes
    never take the process down."""
    import random
    port = server
    st = lambda raw: raw.split(b"\r\n", 1)[0]
    assert st(_raw(port, b"NONSENSE\r\n\r\n")) == b"HTTP/1.1 400 Bad Request"
    assert st(_raw(port, b"GET /v1/models\r\n\r\n")) == b"HTTP/1.1 400 Bad Request"                     # no HTTP version
    for cl in (b"-1", b"abc", b"1e3", b"99999999999999999999999", b""):
        r = _raw(port, b"POST /v1/completions HTTP/1.1\r\nContent-Length: " + cl + b"\r\n\r\n{}")
        assert st(r) == b"HTTP/1.1 400 Bad Request", (cl, r[:80])
        assert json.loads(r.split(b"\r\n\r\n", 1)[1])["error"]["type"] == "invalid_request_error"
    assert st(_raw(port, b"POST /v1/completions HTTP/1.1\r\nTransfer-Encoding: chunked\r\n\r\n2\r\n{}\r\n0\r\n\r\n")) == b"HTTP/1.1 400 Bad Request"
    r = _raw(port, b"POST /v1/completions HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % ((20 << 20) + 1))
    assert st(r) == b"HTTP/1.1 413 Payload Too Large" and b"Connection: close" in r
    r = _raw(port, b"GET / HTTP/1.1\r\n" + b"X-Pad: " + b"a" * 8000 + b"\r\n" * 1 + (b"X-Pad: " + b"a" * 8000 + b"\r\n") * 140, timeout=10)
    assert st(r) == b"HTTP/1.1 431 Request Header Fields Too Large"
    # exactly at the limit is read and handled (it is not JSON: 400 from the handler, not 413)
    big = b"x" * (20 << 20)
    r = _raw(port, b"POST /v1/completions HTTP/1.1\r\nConnection: close\r\nContent-Length: %d\r\n\r\n" % len(big) + big, timeout=20)
    assert st(r) == b"HTTP/1.1 400 Bad Request"
    # Expect: 100-continue is honoured before the body is sent
    import socket
    body = json.dumps({"model": "tiny-llama", "prompt_token_ids": [1, 2, 3], "max_tokens": 4, "temperature": 0}).encode()
    s = socket.create_connection(("127.0.0.1", port), timeout=5)
    s.sendall(b"POST /v1/completions HTTP/1.1\r\nExpect: 100-continue\r\nConnection: Close\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body))
    assert s.recv(1024) == b"HTTP/1.1 100 Continue\r\n\r\n"
    s.sendall(body)
    out = b""
    while True:
        d = s.recv(65536)
        if not d:
            break                                               # "Connection: Close" (any case) ends the connection
        out += d
    s.close()
    assert st(out) == b"HTTP/1.1 200 OK" and json.loads(out.split(b"\r\n\r\n", 1)[1])["usage"]["completion_tokens"] == 4
    # random and half-valid byte soup, some connections abandoned mid-request
    rs = random.Random(5)
    valid = b"POST /v1/chat/completions HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: 62\r\n\r\n" \
            b'{"model":"tiny-llama","messages":[{"role":"user","content":1}]}'
    for i in range(300):
        kind = rs.randrange(4)
        if kind == 0:
            p = bytes(rs.randrange(256) for _ in range(rs.randrange(1, 400)))
        elif kind == 1:
            p = bytearray(valid)
            for _ in range(rs.randrange(1, 6)):
                p[rs.randrange(len(p))] = rs.randrange(256)
            p = bytes(p)
        elif kind == 2:
            p = valid[:rs.randrange(1, len(valid))]
        else:
            p = valid + valid[:rs.randrange(len(valid))] + bytes(rs.randrange(256) for _ in range(rs.randrange(64)))
        _raw(port, p, read=kind != 2, timeout=0.05)
    s_, _, d = T.call(port, "GET", "/v1/models")
    assert s_ == 200 and json.loads(d)["data"][0]["id"] == "tiny-llama"
    s_, _, d = T.call(port, "GET", "/api/health")
    assert s_ == 200


_ODD = [None, True, False, -1, 0, 1, 2 ** 31, 2 ** 40, 10 ** 30, 1e300, -1e300, 1.5, -0.0, "", " ", " ", "x" * 3000, [], {}, [[]], [None], {"type": "text"},
        {"type": "text", "text": 5}, {"type": "image_url", "image_url": {"url": "data:x"}}, [{"role": 3}], "퟿￿", ["a", 1, None]]


def _damage(rs, node, depth=0):
    """Replace / delete / insert at a random place of a JSON tree."""
    if isinstance(node, dict) and node and (depth == 0 or rs.random() < 0.7):
        k = rs.choice(list(node))
        r = rs.random()
        if r < 0.35:
            node[k] = rs.choice(_ODD)
        elif r < 0.45:
            del node[k]
        elif r < 0.55:
            node[rs.choice(["stream", "max_tokens", "stop", "n", "temperature", "top_p", "top_k", "seed", "stream_options", "tools", "input", "system",
                            "messages", "prompt", "model", "max_output_tokens", "ignore_eos", "prompt_token_ids"])] = rs.choice(_ODD)
        else:
            node[k] = _damage(rs, node[k], depth + 1)
        return node
    if isinstance(node, list) and node and rs.random() < 0.7:
        i = rs.randrange(len(node))
        node[i] = _damage(rs, node[i], depth + 1) if rs.random() < 0.6 else rs.choice(_ODD)
        return node
    return rs.choice(_ODD)


def test_request_bodies_with_wrong_types_never_take_the_server_down(tok_server):
    """Every front door (chat, completions, responses, Anthropic messages) with valid-JSON bodies whose fields have the
    wrong type, absurd values or are missing: each request gets an HTTP answer (2xx/4xx with a JSON error body), and the
    server still serves afterwards.  The same loop runs against an ASan/UBSan build of the shim in the long fuzz pass
    (hygiene)."""
    import copy
    import random
    port = tok_server
    rs = random.Random(int(os.environ.get("LLMLB_FUZZ_SEED", "11")))
    seeds = [
        ("/v1/chat/completions", {"model": "tiny-llama", "messages": [{"role": "system", "content": "s"}, {"role": "user", "content": [{"type": "text", "text": "hi"}]}],
                                  "max_tokens": 4, "temperature": 0.5, "top_p": 0.9, "seed": 3, "stop": ["x"], "stream": False, "stream_options": {"include_usage": True}}),
        ("/v1/completions", {"model": "tiny-llama", "prompt": "hello", "max_tokens": 4, "temperature": 0, "stream": False}),
        ("/v1/responses", {"model": "tiny-llama", "input": [{"role": "user", "content": [{"type": "input_text", "text": "hi"}]}], "max_output_tokens": 4, "stream": False}),
        ("/v1/messages", {"model": "tiny-llama", "max_tokens": 4, "system": [{"type": "text", "text": "s"}],
                          "messages": [{"role": "user", "content": [{"type": "text", "text": "hi"}]}], "stop_sequences": ["x"], "stream": False}),
    ]
    n_ok = n_err = 0
    for i in range(int(os.environ.get("LLMLB_FUZZ_REQUESTS", "400"))):
        path, body = seeds[i % 4]
        b = copy.deepcopy(body)
        for _ in range(rs.randint(1, 3)):
            b = _damage(rs, b)
            if not isinstance(b, dict):
                break
        if isinstance(b, dict) and rs.random() < 0.3:
            b["stream"] = True
        hdrs = {"anthropic-version": "2023-06-01"} if path == "/v1/messages" else {}
        try:
            raw = json.dumps(b)
        except (TypeError, ValueError):
            continue
        c = http.client.HTTPConnection("127.0.0.1", port, timeout=30)
        c.request("POST", path, raw.encode("utf-8", "surrogatepass"), {"Content-Type": "application/json", **hdrs})
        r = c.getresponse()
        data = r.read()
        c.close()
        assert r.status in (200, 400, 401, 404, 413, 422, 429, 502, 503, 504), (path, raw[:300], r.status)
        if r.status == 200:
            n_ok += 1
        else:
            n_err += 1
            assert "error" in json.loads(data), (path, raw[:300], data[:200])
    assert n_ok > 20 and n_err > 20, (n_ok, n_err)
    s_, _, d = T.call(port, "GET", "/v1/models")
    assert s_ == 200
    s_, _, d = T.call(port, "GET", "/api/health")
    assert s_ == 200 and json.loads(d)["load"]["active_requests"] == 0


def test_completions_prompt_forms(server):
    """OpenAI's `prompt` is a string, an array with one string, an array of token ids or an array with one such array.  The
    scripted engine's tokens are a hash of the prompt ids, so equal prompts give equal text: "hello" == ["hello"], and the id
    form equals `prompt_token_ids`.  Several prompts per request and non-prompts are refused with 400 — they used to be
    answered, silently, with the completion of an EMPTY prompt."""
    port = server
    ask = lambda p, **kw: T.call(port, "POST", "/v1/completions", dict({"model": "tiny-llama", "prompt": p, "max_tokens": 4, "temperature": 0}, **kw))
    text = lambda r: json.loads(r[2])["choices"][0]["text"]
    a, b = ask("hello"), ask(["hello"])
    assert a[0] == b[0] == 200 and text(a) == text(b) and json.loads(a[2])["usage"]["prompt_tokens"] == json.loads(b[2])["usage"]["prompt_tokens"]
    c, d = ask([5, 6, 7]), ask([[5, 6, 7]])
    e = T.call(port, "POST", "/v1/completions", {"model": "tiny-llama", "prompt_token_ids": [5, 6, 7], "max_tokens": 4, "temperature": 0})
    assert c[0] == d[0] == e[0] == 200 and text(c) == text(d) == text(e) and json.loads(c[2])["usage"]["prompt_tokens"] == 3
    assert text(a) != text(c)
    for bad in (["a", "b"], [[1, 2], [3, 4]], [], None, 5, {"text": "x"}, [1.5, 2], [None], ["a", 1]):
        st, _, body = ask(bad)
        assert st == 400 and json.loads(body)["error"]["type"] == "invalid_request_error", (bad, st, body[:120])
    st, _, body = ask([5, 999999999])                                            # an id outside the vocabulary
    assert st == 400


def test_empty_or_mistyped_conversations_are_refused(server):
    """`messages: []`, `messages: [5]`, a missing / numeric / empty `input`: 400, not a generation from the bare template."""
    port = server
    for path, body in (("/v1/chat/completions", {"messages": []}), ("/v1/chat/completions", {"messages": [5]}), ("/v1/chat/completions", {"messages": "hi"}),
                       ("/v1/responses", {}), ("/v1/responses", {"input": 5}), ("/v1/responses", {"input": []}), ("/v1/responses", {"input": None})):
        st, _, d = T.call(port, "POST", path, dict({"model": "tiny-llama", "max_tokens": 3, "max_output_tokens": 3}, **body))
        assert st == 400 and json.loads(d)["error"]["type"] == "invalid_request_error", (path, body, st, d[:100])
    st, _, d = T.call(port, "POST", "/v1/responses", {"model": "tiny-llama", "input": "hi", "max_output_tokens": 3})
    assert st == 200
    chat = {"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}], "max_tokens": 2}
    assert T.call(port, "POST", "/v1/chat/completions", dict(chat, n=1))[0] == 200
    assert T.call(port, "POST", "/v1/chat/completions", dict(chat, n=3))[0] == 400            # not one choice passed off as three


def test_messages_route_reports_engine_failures_in_anthropic_shape(fake_bin):
    """/v1/messages when the engine's deadline expires: 504 with the Anthropic error body when nothing was sent yet; on a
    stream whose headers are out, Anthropic's own streaming error event (`event: error`, {"type":"error","error":{...}}) and
    no message_stop — it used to be the OpenAI-shaped `{"error":{...}}` event in the middle of an Anthropic stream.  The
    message counts whole seconds, rounded up (a 600 ms limit is "1 seconds", not "0 seconds")."""
    port, proc = _start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "1", "--max-ctx", "4096", "--request-timeout-ms", "600",
                        env={"FAKE_ENGINE_TOKEN_US": "4000"})
    try:
        body = {"model": "tiny-llama", "max_tokens": 2000, "messages": [{"role": "user", "content": "x"}]}
        hdr = {"anthropic-version": "2023-06-01"}
        st, _, d = T.call(port, "POST", "/v1/messages", body, hdr)
        assert st == 504 and json.loads(d) == {"type": "error", "error": {"type": "api_error", "message": "Upstream endpoint request timed out after 1 seconds"}}
        st, _, d = T.call(port, "POST", "/v1/messages", dict(body, stream=True), hdr)
        text = d.decode()
        blocks = [b for b in text.split("\n\n") if b.strip()]
        assert st == 200 and blocks[0].startswith("event: message_start") and "message_stop" not in text and '"error":{"message"' not in text
        ev, data = blocks[-1].split("\n", 1)
        assert ev == "event: error" and json.loads(data[6:]) == {"type": "error", "error": {"type": "api_error", "message": "Upstream endpoint request timed out after 1 seconds"}}
    finally:
        proc.terminate(); proc.wait(timeout=20)


def test_token_budgets_that_are_not_positive_integers_are_refused(server):
    port = server
    chat = {"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}]}
    for bad in (0, -3, 2.5, "10", [4]):
        st, _, d = T.call(port, "POST", "/v1/chat/completions", dict(chat, max_tokens=bad))
        assert st == 400 and "max_tokens" in json.loads(d)["error"]["message"], (bad, st)
    assert T.call(port, "POST", "/v1/responses", {"model": "tiny-llama", "input": "x", "max_output_tokens": 0})[0] == 400
    assert T.call(port, "POST", "/v1/chat/completions", dict(chat, max_tokens=None))[0] == 200          # null = the default budget
    st, _, d = T.call(port, "POST", "/v1/chat/completions", dict(chat, max_completion_tokens=2))
    assert st == 200 and json.loads(d)["usage"]["completion_tokens"] == 2


def test_connection_cap_answers_503_and_recovers(fake_bin):
    """--max-connections 4: four idle keep-alive connections are held open; the fifth is answered 503 service_unavailable with
    Retry-After and closed at once (no thread is spent on it); when one of the four goes away the next connection is served."""
    import socket
    port, proc = _start(fake_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "2", "--max-ctx", "512", "--max-connections", "4")
    try:
        held = []
        for _ in range(4):
            s = socket.create_connection(("127.0.0.1", port), timeout=5)
            s.sendall(b"GET /v1/models HTTP/1.1\r\nHost: x\r\n\r\n")
            assert s.recv(4096).startswith(b"HTTP/1.1 200")
            held.append(s)
        r = _raw(port, b"GET /v1/models HTTP/1.1\r\nHost: x\r\n\r\n")
        assert r.startswith(b"HTTP/1.1 503") and b"Retry-After: 1" in r and json.loads(r.split(b"\r\n\r\n", 1)[1])["error"]["type"] == "service_unavailable"
        held.pop().close()
        deadline = time.time() + 5
        while True:
            r = _raw(port, b"GET /v1/models HTTP/1.1\r\nHost: x\r\nConnection: close\r\n\r\n")
            if r.startswith(b"HTTP/1.1 200"):
                break
            assert time.time() < deadline, r[:80]
            time.sleep(0.05)
        for s in held:
            s.close()
    finally:
        proc.terminate(); proc.wait(timeout=20)
