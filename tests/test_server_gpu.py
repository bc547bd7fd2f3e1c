"""-m gpu: the HTTP shim (llmlb_b200_server) that exposes the engine behind the endpoint
contract llmlb consumes.  The gateway's own accounting (oracle restatement of
llmlb/src/token/mod.rs + api/proxy.rs) is run over the bytes the server produces."""
import http.client
import json
import os
import socket
import subprocess
import time

import pytest

from oracle import gateway_ref as G

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "llmlb_b200", "llmlb_b200_server")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.fixture(scope="module")
def server(built_lib):
    from llmlb_b200 import build
    build.build_host()
    port = _free_port()
    proc = subprocess.Popen([BIN, "--port", str(port), "--model", "tiny", "--model-id", "tiny-llama",
                             "--max-seqs", "8", "--max-ctx", "512"], stderr=subprocess.PIPE)
    deadline = time.time() + 120
    while time.time() < deadline:
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
            break
        except OSError:
            assert proc.poll() is None, proc.stderr.read().decode()
            time.sleep(0.2)
    yield port
    proc.terminate()
    proc.wait(timeout=20)


def call(port, method, path, body=None, headers=None):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
    c.request(method, path, json.dumps(body) if body is not None else None,
              {"Content-Type": "application/json", **(headers or {})})
    r = c.getresponse()
    data = r.read()
    c.close()
    return r.status, dict(r.getheaders()), data


def test_probe_endpoints(server):
    st, _, d = call(server, "GET", "/v1/models")
    assert st == 200 and json.loads(d)["data"][0]["id"] == "tiny-llama"          # sync/parser.rs:78-110
    st, _, d = call(server, "GET", "/api/system")
    assert st == 200 and "xllm_version" in json.loads(d)                          # detection/xllm.rs:27-66
    st, _, d = call(server, "GET", "/api/health")
    h = json.loads(d)
    assert st == 200 and h["gpu"]["device_count"] == 1 and "active_requests" in h["load"]
    st, _, d = call(server, "GET", "/api/models/tiny-llama/info")
    assert st == 200 and json.loads(d)["context_length"] == 512                   # metadata/xllm.rs:48-61


def test_chat_completion_non_stream(server):
    st, _, d = call(server, "POST", "/v1/chat/completions",
                    {"model": "tiny-llama", "messages": [{"role": "user", "content": "Hi"}], "max_tokens": 9, "temperature": 0, "ignore_eos": True})
    j = json.loads(d)
    assert st == 200 and j["object"] == "chat.completion" and j["choices"][0]["message"]["role"] == "assistant"
    assert G.extract_usage_from_response(j)["output_tokens"] == 9 and j["choices"][0]["finish_reason"] == "length"


def test_chat_completion_stream_accounting(server):
    body = {"model": "tiny-llama", "messages": [{"role": "user", "content": "stream please"}], "max_tokens": 12,
            "temperature": 0, "stream": True, "stream_options": {"include_usage": True}, "ignore_eos": True}
    st, hdr, d = call(server, "POST", "/v1/chat/completions", body)
    assert st == 200 and hdr["Content-Type"].startswith("text/event-stream")
    text = d.decode()
    assert text.startswith("data: ") and text.rstrip().endswith("data: [DONE]")
    acc = G.StreamingTokenAccumulator("tiny-llama")
    assert G.process_sse_lines(text, acc) == "" and acc.done
    u = acc.finalize()
    assert u["output_tokens"] == 12 and u["total_tokens"] == u["input_tokens"] + 12
    # greedy stream == greedy non-stream
    st, _, d2 = call(server, "POST", "/v1/chat/completions", {**body, "stream": False})
    assert json.loads(d2)["choices"][0]["message"]["content"] == acc.accumulated_content


def test_responses_stream_and_body(server):
    body = {"model": "tiny-llama", "input": "hello", "max_output_tokens": 6, "temperature": 0, "stream": True, "ignore_eos": True}
    st, _, d = call(server, "POST", "/v1/responses", body)
    events = [json.loads(e[6:]) for e in d.decode().split("\n\n") if e.startswith("data: {")]
    assert [e["type"] for e in events][:3] == ["response.created", "response.output_item.added", "response.content_part.added"]
    assert events[-1]["type"] == "response.done" and events[-1]["response"]["usage"]["output_tokens"] == 6
    st, _, d = call(server, "POST", "/v1/responses", {**body, "stream": False})
    j = json.loads(d)
    assert st == 200 and j["object"] == "response" and j["usage"]["output_tokens"] == 6


def test_prompt_token_ids_and_completions(server):
    st, _, d = call(server, "POST", "/v1/completions", {"model": "tiny-llama", "prompt_token_ids": list(range(5, 45)), "max_tokens": 5, "temperature": 0, "ignore_eos": True})
    j = json.loads(d)
    assert st == 200 and j["usage"]["prompt_tokens"] == 40 and j["usage"]["completion_tokens"] == 5


def test_errors(server):
    st, _, d = call(server, "POST", "/v1/chat/completions", {"model": "nope", "messages": []})
    assert st == 404 and json.loads(d)["error"]["code"] == 404                      # openai.rs:813-817
    st, _, d = call(server, "POST", "/v1/chat/completions", {"messages": []})
    assert st == 400
    st, _, d = call(server, "POST", "/v1/chat/completions", {"model": "a:b:c", "messages": []})
    assert st == 400 and "quantization format" in json.loads(d)["error"]["message"]
    st, _, d = call(server, "POST", "/v1/chat/completions", {"model": "tiny-llama", "messages": [{"role": "user", "content": [{"type": "image_url", "image_url": {"url": "x"}}]}]})
    assert st == 400
    st, _, _ = call(server, "GET", "/nothing")
    assert st == 404


def test_drain_gate(server):
    call(server, "POST", "/admin/drain", True)
    st, hdr, d = call(server, "POST", "/v1/chat/completions", {"model": "tiny-llama", "messages": [], "max_tokens": 2})
    call(server, "POST", "/admin/drain", False)
    assert st == 503 and hdr.get("Retry-After") == "30"                              # inference_gate.rs:177-197
    assert json.loads(d) == G.gate_rejection()[2]


def test_concurrent_streams(server):
    import threading
    outs = [None] * 12

    def run(i):
        st, _, d = call(server, "POST", "/v1/responses", {"model": "tiny-llama", "prompt_token_ids": list(range(3 + i, 60 + 3 * i)),
                                                           "max_output_tokens": 16, "temperature": 0, "stream": True, "ignore_eos": True})
        acc = G.StreamingTokenAccumulator("m")
        G.process_sse_lines(d.decode(), acc)
        outs[i] = (st, acc.finalize()["output_tokens"], acc.done)
    th = [threading.Thread(target=run, args=(i,)) for i in range(12)]
    [t.start() for t in th]; [t.join() for t in th]
    assert all(o == (200, 16, True) for o in outs), outs


# ---- native tokenizer on the serving path ----
@pytest.fixture(scope="module")
def tok_server(built_lib):
    from llmlb_b200 import build
    build.build_host()
    port = _free_port()
    proc = subprocess.Popen([BIN, "--port", str(port), "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8",
                             "--max-ctx", "512", "--vocab", "3072",
                             "--tokenizer", os.path.join(HERE, "golden", "tokenizer_llama3_style.json")], stderr=subprocess.PIPE)
    deadline = time.time() + 120
    while time.time() < deadline:
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
            break
        except OSError:
            assert proc.poll() is None, proc.stderr.read().decode()
            time.sleep(0.2)
    yield port
    proc.terminate()
    proc.wait(timeout=20)


def test_chat_through_the_native_tokenizer(tok_server):
    gold = json.load(open(os.path.join(HERE, "golden", "tokenizer_vectors.json"), encoding="utf-8"))
    conv = gold["chats"][1]
    body = {"model": "tiny-llama", "messages": conv["messages"], "max_tokens": 48, "temperature": 0, "ignore_eos": True}
    st, _, d = call(tok_server, "POST", "/v1/chat/completions", body)
    assert st == 200
    j = json.loads(d)
    # the prompt is the chat template's ids (markers as control tokens, content as plain text)
    assert j["usage"]["prompt_tokens"] == len(conv["ids"])
    assert j["usage"]["completion_tokens"] == 48
    full = j["choices"][0]["message"]["content"]
    # same request streamed: every delta is complete UTF-8 and the pieces add up to the same text
    st, _, d = call(tok_server, "POST", "/v1/chat/completions", dict(body, stream=True, stream_options={"include_usage": True}))
    assert st == 200
    acc = G.StreamingTokenAccumulator("tiny-llama")
    pieces = []
    for line in d.decode("utf-8").split("\n"):
        if line.startswith("data: ") and line != "data: [DONE]":
            ch = json.loads(line[6:])
            for c in ch.get("choices", []):
                if c.get("delta", {}).get("content"):
                    pieces.append(c["delta"]["content"])
        if line:
            acc.process_chunk(line)
    assert "".join(pieces) == full == acc.accumulated_content
    assert acc.finalize()["output_tokens"] == 48 and acc.done


def test_stop_at_end_of_turn_token(tok_server):
    """Without ignore_eos the control tokens of the template end generation ('stop') and never
    show up in the text; a prompt outside the vocabulary is a 400."""
    st, _, d = call(tok_server, "POST", "/v1/completions", {"model": "tiny-llama", "prompt": "hello", "max_tokens": 400, "temperature": 1.0, "seed": 3})
    assert st == 200
    j = json.loads(d)
    assert "<|eot_id|>" not in j["choices"][0]["text"]
    assert j["choices"][0]["finish_reason"] in ("stop", "length")
    st, _, d = call(tok_server, "POST", "/v1/completions", {"model": "tiny-llama", "prompt_token_ids": [5, 999999], "max_tokens": 4})
    assert st == 400


# ---- Anthropic Messages front door (llmlb/src/api/anthropic.rs) ----
def test_messages_route_non_stream_and_stream(server):
    hdr = {"anthropic-version": "2023-06-01"}
    body = {"model": "tiny-llama", "max_tokens": 24, "temperature": 0, "system": "be brief",
            "messages": [{"role": "user", "content": [{"type": "text", "text": "Hello"}]}]}
    st, _, d = call(server, "POST", "/v1/messages", body, hdr)
    assert st == 200
    j = json.loads(d)
    assert (j["type"], j["role"], j["model"], j["stop_sequence"]) == ("message", "assistant", "tiny-llama", None)
    assert j["content"][0]["type"] == "text" and j["stop_reason"] in ("end_turn", "max_tokens")
    assert j["usage"]["input_tokens"] > 0 and 0 < j["usage"]["output_tokens"] <= 24
    # the same prompt through /v1/chat/completions gives the same text (the route is a translation)
    conv, _, _ = G.anthropic_request_to_openai(body)
    st, _, d2 = call(server, "POST", "/v1/chat/completions", conv)
    assert st == 200 and json.loads(d2)["choices"][0]["message"]["content"] == j["content"][0]["text"]
    # streamed: Anthropic event sequence, text adds up, output_tokens in message_delta
    st, h, d = call(server, "POST", "/v1/messages", dict(body, stream=True), hdr)
    assert st == 200 and h.get("Content-Type", h.get("content-type")) == "text/event-stream"
    evs = []
    for blk in d.decode("utf-8").split("\n\n"):
        if blk.startswith("event: "):
            name, data = blk.split("\n", 1)
            evs.append((name[7:], json.loads(data[6:])))
    names = [n for n, _ in evs]
    assert names[0] == "message_start" and names[1] == "content_block_start" and names[-3:] == ["content_block_stop", "message_delta", "message_stop"]
    assert set(names[2:-3]) <= {"content_block_delta"}
    text = "".join(e["delta"]["text"] for n, e in evs if n == "content_block_delta")
    assert text == j["content"][0]["text"]
    assert evs[0][1]["message"]["usage"]["input_tokens"] == j["usage"]["input_tokens"]
    assert evs[-2][1]["usage"]["output_tokens"] == j["usage"]["output_tokens"]
    assert evs[-2][1]["delta"]["stop_reason"] == j["stop_reason"]
    # and it is exactly what the reference's transformer makes of our own chat stream
    st, _, chat_sse = call(server, "POST", "/v1/chat/completions", dict(conv, stream=True, stream_options={"include_usage": True}))
    ref = G.AnthropicStreamTransformer("tiny-llama", input_tokens=j["usage"]["input_tokens"])
    ref.feed(chat_sse.decode("utf-8"))
    ref.finish()
    strip = lambda out: [(n, {k: v for k, v in e.items() if k != "message"} if n == "message_start" else e) for n, e in out]
    assert strip(evs) == strip(ref.out)


def test_messages_route_errors_in_anthropic_shape(server):
    ok_hdr = {"anthropic-version": "2023-06-01"}
    st, _, d = call(server, "POST", "/v1/messages", {"model": "tiny-llama", "max_tokens": 4, "messages": []})
    j = json.loads(d)   # llmlb/tests/contract/anthropic_messages_api_test.rs:223-247
    assert st == 400 and j["type"] == "error" and j["error"] == {"type": "invalid_request_error", "message": "Missing required header: anthropic-version"}
    st, _, d = call(server, "POST", "/v1/messages", {"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}]}, ok_hdr)
    assert st == 400 and json.loads(d)["error"]["message"] == "max_tokens is required"
    st, _, d = call(server, "POST", "/v1/messages", {"model": "tiny-llama", "max_tokens": 4, "messages": [{"role": "user", "content": [{"type": "image", "source": {}}]}]}, ok_hdr)
    assert st == 400 and "is not supported" in json.loads(d)["error"]["message"]
    st, _, d = call(server, "POST", "/v1/messages", {"model": "nope", "max_tokens": 4, "messages": [{"role": "user", "content": "x"}]}, ok_hdr)
    j = json.loads(d)
    assert st == 404 and j["type"] == "error" and j["error"]["type"] == "not_found_error"


# ---- a .gguf alone: geometry, weights and tokenizer from one file ----
def test_server_from_a_single_gguf(built_lib, tmp_path):
    pytest.importorskip("gguf")
    from gguf_util import write_tiny_llama_gguf
    from llmlb_b200 import build
    build.build_host()
    p = tmp_path / "tiny.gguf"
    write_tiny_llama_gguf(p)
    port = _free_port()
    proc = subprocess.Popen([BIN, "--port", str(port), "--model", "auto", "--model-id", "tiny-gguf", "--max-seqs", "4", "--max-ctx", "512",
                             "--weights", str(p)], stderr=subprocess.PIPE)
    try:
        deadline = time.time() + 120
        while time.time() < deadline:
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
                break
            except OSError:
                assert proc.poll() is None, proc.stderr.read().decode()
                time.sleep(0.2)
        gold = json.load(open(os.path.join(HERE, "golden", "tokenizer_vectors.json"), encoding="utf-8"))
        conv = gold["chats"][0]
        st, _, d = call(port, "POST", "/v1/chat/completions", {"model": "tiny-gguf", "messages": conv["messages"], "max_tokens": 8, "temperature": 0, "ignore_eos": True})
        assert st == 200
        j = json.loads(d)
        assert j["usage"]["prompt_tokens"] == len(conv["ids"]) and j["usage"]["completion_tokens"] == 8
        st, _, d = call(port, "GET", "/api/models/tiny-gguf/info")
        assert st == 200
    finally:
        proc.terminate()
        proc.wait(timeout=20)


def test_stop_strings_end_the_text_before_the_match(server):
    body = {"model": "tiny-llama", "prompt": "abc", "max_tokens": 40, "temperature": 0, "ignore_eos": True}
    st, _, d = call(server, "POST", "/v1/completions", body)
    assert st == 200
    full = json.loads(d)["choices"][0]["text"]
    assert len(full) >= 12
    stop = full[7:10]
    cut = full.find(stop)
    st, _, d = call(server, "POST", "/v1/completions", dict(body, stop=[stop, "\x00never"]))
    j = json.loads(d)
    assert st == 200 and j["choices"][0]["text"] == full[:cut] and j["choices"][0]["finish_reason"] == "stop"
    assert j["usage"]["completion_tokens"] < 40
    # chat + streaming: the deltas add up to the same truncated text and the stop string never shows
    cbody = {"model": "tiny-llama", "messages": [{"role": "user", "content": "abc"}], "max_tokens": 40, "temperature": 0, "ignore_eos": True}
    st, _, d = call(server, "POST", "/v1/chat/completions", cbody)
    cfull = json.loads(d)["choices"][0]["message"]["content"]
    cstop = cfull[5:8]
    st, _, d = call(server, "POST", "/v1/chat/completions", dict(cbody, stop=cstop, stream=True))
    acc = G.StreamingTokenAccumulator("tiny-llama")
    assert G.process_sse_lines(d.decode("utf-8"), acc) == "" and acc.done
    assert acc.accumulated_content == cfull[:cfull.find(cstop)]


def test_model_download_routes(server, built_lib, tmp_path):
    """POST /api/models/download + GET /api/download/progress (the gateway is the client: llmlb/src/xllm/download.rs:97,147).
    The default server has no mirror configured: the route answers 503 and queues nothing.  A server started with
    --mirror-root / --models-dir copies the file and reports DownloadProgressResponse-shaped documents (the manager
    itself is tested on CPU: tests/test_host_download.py)."""
    st, _, d = call(server, "POST", "/api/models/download", {"repo": "org/model-GGUF"})
    assert st == 503 and "not configured" in json.loads(d)["error"]
    st, _, d = call(server, "GET", "/api/download/progress?task_id=task-1")
    assert st == 404 and "error" in json.loads(d)
    root, models = tmp_path / "mirror", tmp_path / "models"
    (root / "org" / "model-GGUF").mkdir(parents=True)
    blob = os.urandom(1 << 20)
    (root / "org" / "model-GGUF" / "model-Q4_K_M.gguf").write_bytes(blob)
    (root / "org" / "model-GGUF" / "model-Q8_0.gguf").write_bytes(b"x" * 10)
    port = _free_port()
    proc = subprocess.Popen([BIN, "--port", str(port), "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "4", "--max-ctx", "256",
                             "--mirror-root", str(root), "--models-dir", str(models)], stderr=subprocess.PIPE)
    try:
        deadline = time.time() + 120
        while time.time() < deadline:
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
                break
            except OSError:
                assert proc.poll() is None, proc.stderr.read().decode()
                time.sleep(0.2)
        st, _, d = call(port, "POST", "/api/models/download", {"repo": "org/model-GGUF"})
        init = json.loads(d)
        assert st == 200 and set(init) == {"task_id", "model", "status"}
        for _ in range(400):
            st, _, d = call(port, "GET", "/api/download/progress?task_id=" + init["task_id"])
            p = json.loads(d)
            assert st == 200 and p["status"] in ("pending", "downloading", "completed") and 0.0 <= p["progress"] <= 100.0
            if p["status"] == "completed":
                break
            time.sleep(0.01)
        assert p["status"] == "completed" and p["filename"] == "model-Q4_K_M.gguf" and p["progress"] == 100.0
        assert (models / "org--model-GGUF" / "model-Q4_K_M.gguf").read_bytes() == blob
        st, _, d = call(port, "POST", "/api/models/download", {"repo": "../outside"})
        assert st == 400
    finally:
        proc.terminate()
        proc.wait(timeout=20)


def _gateway_fetch(port, api_key=None):
    """The HTTP the reference's detection / health / sync / metadata clients send (GET, optional bearer), as the `fetch`
    the restated clients in oracle/gateway_ref.py take."""
    def fetch(path, auth):
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=10)
            c.request("GET", path, headers={"Authorization": "Bearer " + api_key} if (auth and api_key) else {})
            r = c.getresponse()
            data = r.read()
            c.close()
        except OSError:
            return None
        try:
            j = json.loads(data)
        except ValueError:
            j = None
        return r.status, {k.lower(): v for k, v in r.getheaders()}, j
    return fetch


def test_an_unmodified_gateway_would_register_and_sync_this_endpoint(server, built_lib):
    """Registration, health and model sync as llmlb performs them, with the reference's CLIENT
    logic restated in the oracle and pinned to its own test vectors: the shim is typed `xllm` by the priority chain
    (detection/mod.rs:85-195), its /api/health parses into GpuInfo (health/endpoint_checker.rs:515-557), /v1/models syncs to
    one chat-capable model whose max_tokens comes from /api/models/{id}/info (sync/mod.rs:104-278, metadata/xllm.rs:48-99)."""
    fetch = _gateway_fetch(server)
    assert G.detect_endpoint_type(fetch) == ("xllm", "xLLM: /api/system xllm_version=llmlb_b200-0.1")
    hz = G.parse_v0_health(fetch("/api/health", True))
    assert hz["gpu_device_count"] == 1 and hz["gpu_total_memory_bytes"] > 0 and hz["gpu_used_memory_bytes"] is not None
    assert hz["gpu_capability_score"] == 100.0 and hz["active_requests"] is not None
    models, fmt = G.sync_models(fetch, "xllm")
    assert fmt == "openai" and models == [{"model_id": "tiny-llama", "capabilities": ["chat"], "supported_apis": ["chat_completions"], "max_tokens": 512}]
    info = G.parse_xllm_model_info(fetch(G.xllm_model_info_url("tiny-llama"), True))
    assert info["model"] == "tiny-llama" and info["context_length"] == 512
    with pytest.raises(ValueError):
        G.parse_xllm_model_info(fetch(G.xllm_model_info_url("nonexistent-model"), True))          # tests/support/xllm.rs:35-44: 404
    # an id with the three characters the gateway escapes, behind an API key: every probe that carries the key works, the two
    # probes the reference sends WITHOUT it (Ollama /api/tags, llama.cpp) are refused and the chain still ends at xllm
    port = _free_port()
    mid = "meta-llama/Tiny Llama:q8"
    proc = subprocess.Popen([BIN, "--port", str(port), "--model", "tiny", "--model-id", mid, "--max-seqs", "2", "--max-ctx", "256", "--api-key", "sk-test"],
                            stderr=subprocess.PIPE)
    try:
        deadline = time.time() + 120
        while time.time() < deadline:
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
                break
            except OSError:
                assert proc.poll() is None, proc.stderr.read().decode()
                time.sleep(0.2)
        fetch = _gateway_fetch(port, "sk-test")
        assert fetch("/api/tags", False)[0] == 401 and G.detect_endpoint_type(fetch)[0] == "xllm"
        models, _ = G.sync_models(fetch, "xllm")
        assert models == [{"model_id": mid, "capabilities": ["chat"], "supported_apis": ["chat_completions"], "max_tokens": 256}]
        # without the key the gateway gets 401s everywhere: it answers, so "unsupported", not "unreachable" (registration fails loudly)
        with pytest.raises(ValueError) as ei:
            G.detect_endpoint_type(_gateway_fetch(port, None))
        assert str(ei.value) == "unsupported"
    finally:
        proc.terminate()
        proc.wait(timeout=20)
