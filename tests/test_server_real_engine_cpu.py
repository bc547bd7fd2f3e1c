"""The PRODUCT server binary linked to the PRODUCT engine library, on a machine without a GPU: the server process is
started with tests/support/fake_cudart.cpp preloaded (a test double of libcudart: host memory, no-op launches, every token
id 0), so the shim, the C ABI and the engine's whole host side (queue, scheduler, paging, timeouts, event delivery) run
together exactly as shipped — only the arithmetic is missing.  tests/test_server_fake_engine_cpu.py covers the shim against
a scripted engine; this covers the integration with the real one, including two paths a GPU run cannot reach on purpose:
deadlines / queue timeouts expiring inside the real scheduler, and a device fault in the middle of serving."""
import http.client
import json
import os
import subprocess
import sys
import threading
import time

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import test_engine_host_logic_cpu as HL  # noqa: E402
import test_server_gpu as T  # noqa: E402
from oracle import gateway_ref as G  # noqa: E402


@pytest.fixture(scope="module")
def product_bin(built_lib):
    from llmlb_b200 import build
    build.build_host()
    alt = os.environ.get("LLMLB_SERVER_BIN")            # a sanitizer build of shim + engine (tools/sanitize_engine_host.py)
    assert os.path.exists(alt or T.BIN)
    return alt or T.BIN


def _start(binary, *args, env=None):
    port = T._free_port()
    errlog = os.environ.get("LLMLB_SERVER_STDERR")       # sanitizer reports of the server process are collected here
    proc = subprocess.Popen([binary, "--port", str(port), *args], stderr=open(errlog, "a") if errlog else subprocess.PIPE,
                            env={**os.environ, "LD_PRELOAD": ":".join(x for x in (os.environ.get("LLMLB_SERVER_PRELOAD_FIRST"), HL.build_fake()) if x), **(env or {})})
    deadline = time.time() + 30
    while time.time() < deadline:
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=2); c.request("GET", "/v1/models"); c.getresponse().read(); c.close()
            return port, proc
        except OSError:
            assert proc.poll() is None, proc.stderr.read().decode() if proc.stderr else "server exited"
            time.sleep(0.05)
    proc.kill()
    raise AssertionError("server did not come up")


@pytest.fixture(autouse=True)
def _servers_started_by_the_reused_tests_also_get_the_fake_runtime(monkeypatch, product_bin):
    # a few reused tests start their own server from T.BIN: child processes of this test inherit the preload (this process,
    # whose libraries are already loaded, is unaffected)
    monkeypatch.setenv("LD_PRELOAD", ":".join(x for x in (os.environ.get("LLMLB_SERVER_PRELOAD_FIRST"), HL.build_fake()) if x))
    monkeypatch.setattr(T, "BIN", product_bin)


@pytest.fixture(scope="module")
def server(product_bin):
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8", "--max-ctx", "512")
    yield port
    proc.terminate(); proc.wait(timeout=20)


@pytest.fixture(scope="module")
def tok_server(product_bin):
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "8", "--max-ctx", "512", "--vocab", "3072",
                        "--tokenizer", os.path.join(HERE, "golden", "tokenizer_llama3_style.json"))
    yield port
    proc.terminate(); proc.wait(timeout=20)


# the protocol tests of tests/test_server_gpu.py that do not depend on what the tokens are
test_probe_endpoints = T.test_probe_endpoints
test_chat_completion_non_stream = T.test_chat_completion_non_stream
test_chat_completion_stream_accounting = T.test_chat_completion_stream_accounting
test_responses_stream_and_body = T.test_responses_stream_and_body
test_prompt_token_ids_and_completions = T.test_prompt_token_ids_and_completions
test_errors = T.test_errors
test_drain_gate = T.test_drain_gate
test_concurrent_streams = T.test_concurrent_streams
test_messages_route_non_stream_and_stream = T.test_messages_route_non_stream_and_stream
test_messages_route_errors_in_anthropic_shape = T.test_messages_route_errors_in_anthropic_shape
test_chat_through_the_native_tokenizer = T.test_chat_through_the_native_tokenizer
test_stop_at_end_of_turn_token = T.test_stop_at_end_of_turn_token
test_server_from_a_single_gguf = T.test_server_from_a_single_gguf
test_stop_strings_end_the_text_before_the_match = T.test_stop_strings_end_the_text_before_the_match
test_model_download_routes = T.test_model_download_routes
test_an_unmodified_gateway_would_register_and_sync_this_endpoint = T.test_an_unmodified_gateway_would_register_and_sync_this_endpoint
# the structured request fuzz and the raw-HTTP abuse of tests/test_server_fake_engine_cpu.py, against the real engine
import test_server_fake_engine_cpu as _F  # noqa: E402
test_request_bodies_with_wrong_types_never_take_the_server_down = _F.test_request_bodies_with_wrong_types_never_take_the_server_down
test_malformed_http_is_answered_or_dropped_and_the_server_stays_up = _F.test_malformed_http_is_answered_or_dropped_and_the_server_stays_up


def test_deadline_and_queue_timeout_expire_inside_the_real_scheduler(product_bin):
    """request_timeout_ms / queue_timeout_ms are enforced by engine.cu's expire_requests(); with 3 ms decode steps a 2000-token
    request cannot finish in a second and a second request cannot be admitted on a 1-sequence engine.  Status, type and message
    are the gateway's for the same upstream failure (openai_util.rs:86-134, openai.rs:862-882)."""
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "1", "--max-ctx", "4096",
                        "--request-timeout-ms", "1000", "--queue-timeout-ms", "300", env={"FAKE_CUDART_STEP_US": "3000"})
    try:
        slow = {"model": "tiny-llama", "prompt_token_ids": [5, 6, 7], "max_tokens": 2000, "temperature": 0, "ignore_eos": True}
        st, _, d = T.call(port, "POST", "/v1/completions", slow)
        status, etype, msg = G.classify_upstream_request_error("timeout", 1)
        assert (st, json.loads(d)) == (status, G.openai_error_body(msg, etype, status))
        st, _, d = T.call(port, "POST", "/v1/chat/completions", {"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}], "max_tokens": 2000,
                                                                   "temperature": 0, "ignore_eos": True, "stream": True})
        text = d.decode()
        events = [json.loads(l[6:]) for l in text.split("\n") if l.startswith("data: {")]
        assert st == 200 and "[DONE]" not in text and events[-1] == G.openai_error_body(msg, etype, status) and len(events) > 20
        res = {}
        t1 = threading.Thread(target=lambda: res.setdefault("a", T.call(port, "POST", "/v1/completions", slow)))
        t1.start(); time.sleep(0.1)
        res["b"] = T.call(port, "POST", "/v1/completions", dict(slow, prompt_token_ids=[9, 9, 9]))
        t1.join()
        st, _, body = G.queue_wait_timeout()
        assert (res["b"][0], json.loads(res["b"][2])) == (st, body)
        hz = json.loads(T.call(port, "GET", "/api/health")[2])
        assert hz["load"]["active_requests"] == 0 and hz["kv"]["free_pages"] == hz["kv"]["total_pages"]       # expired requests gave everything back
    finally:
        proc.terminate(); proc.wait(timeout=20)


def test_a_device_fault_fails_every_request_and_the_server_keeps_answering(product_bin):
    """The 40th decode step's launch returns cudaErrorLaunchFailure.  The engine marks itself failed (engine.cu
    sched_iteration): requests in flight end with FINISH_ERROR -> the shim answers 502 endpoint_request_error ("Failed to
    proxy request to upstream endpoint", openai_util.rs:128-134), streamed ones get the in-band error event and no [DONE];
    later requests are refused at once; probes keep working (the gateway's health checker will take the endpoint offline)."""
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "4", "--max-ctx", "1024",
                        env={"FAKE_CUDART_FAIL_AFTER": "40", "FAKE_CUDART_STEP_US": "500"})
    try:
        body = {"model": "tiny-llama", "prompt_token_ids": [5, 6, 7], "max_tokens": 500, "temperature": 0, "ignore_eos": True}
        res = []
        th = [threading.Thread(target=lambda s=s: res.append((s, T.call(port, "POST", "/v1/completions", dict(body, stream=s))))) for s in (False, True, False)]
        [t.start() for t in th]; [t.join() for t in th]
        status, etype, msg = G.classify_upstream_request_error("other", 0)
        assert status == 502
        for streamed, (st, _, d) in res:
            if streamed:
                text = d.decode()
                events = [json.loads(l[6:]) for l in text.split("\n") if l.startswith("data: {")]
                assert st == 200 and "[DONE]" not in text and events[-1] == G.openai_error_body(msg, etype, status)
            else:
                assert (st, json.loads(d)) == (status, G.openai_error_body(msg, etype, status))
        st, _, d = T.call(port, "POST", "/v1/completions", dict(body, max_tokens=3))                        # after the fault
        assert st == 502 and json.loads(d)["error"]["type"] == etype
        assert T.call(port, "GET", "/v1/models")[0] == 200 and T.call(port, "GET", "/api/health")[0] == 200
        assert proc.poll() is None
    finally:
        proc.terminate(); proc.wait(timeout=20)


def test_clients_that_hang_up_mid_stream_leave_nothing_behind(product_bin):
    """60 streaming clients, 8 at a time, on a 4-sequence engine with 300 us decode steps: most of them close the socket
    after a few events (some before the first byte), a few read to the end.  The shim notices the dead socket, cancels and
    releases (the engine path that held the id-table bug); afterwards nothing is running or queued, every KV page is free,
    the in-flight gauge is back to zero and the server still answers.  Runs under ASan / TSan in the sanitizer pass."""
    import random
    import socket
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "4", "--max-ctx", "1024", "--queue-max", "64",
                        env={"FAKE_CUDART_STEP_US": "300"})
    try:
        body = json.dumps({"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}], "max_tokens": 600, "temperature": 0, "ignore_eos": True,
                           "stream": True}).encode()
        req = b"POST /v1/chat/completions HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body
        done = []

        def client(seed):
            rs = random.Random(seed)
            for _ in range(8):
                s = socket.create_connection(("127.0.0.1", port), timeout=20)
                try:
                    s.sendall(req)
                    mode = rs.random()
                    want = 0 if mode < 0.15 else rs.randrange(200, 4000) if mode < 0.9 else 10 ** 9
                    got = 0
                    while got < want:
                        d = s.recv(2048)
                        if not d:
                            break
                        got += len(d)
                        if b"[DONE]" in d:
                            break
                    done.append(mode >= 0.9)
                finally:
                    s.close()                                                  # abrupt: no reading to the end

        th = [threading.Thread(target=client, args=(i,)) for i in range(8)]
        [t.start() for t in th]; [t.join() for t in th]
        assert len(done) == 64
        deadline = time.time() + 20
        while True:
            hz = json.loads(T.call(port, "GET", "/api/health")[2])
            if hz["load"]["active_requests"] == 0 and hz["load"]["queued_requests"] == 0 and hz["load"]["in_flight_http"] == 0:
                break
            assert time.time() < deadline, hz
            time.sleep(0.05)
        assert hz["kv"]["free_pages"] == hz["kv"]["total_pages"]
        st, _, d = T.call(port, "POST", "/v1/completions", {"model": "tiny-llama", "prompt": "x", "max_tokens": 3})
        assert st == 200 and json.loads(d)["usage"]["completion_tokens"] == 3 and proc.poll() is None
    finally:
        proc.terminate(); proc.wait(timeout=20)


def test_a_client_that_stops_reading_does_not_pin_the_request_for_ever(product_bin):
    """A streaming client that keeps the connection open but never reads: the socket buffers fill, the server's send blocks.
    Without a send timeout that thread, the engine slot and the KV pages stayed taken until the client went away (never, for a
    wedged one).  With --send-timeout-ms the blocked send gives up, the request is cancelled and released like any hang-up."""
    import socket
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "2", "--max-ctx", "131072", "--request-timeout-ms", "600000",
                        "--send-timeout-ms", "400", env={"FAKE_CUDART_STEP_US": "20"})
    try:
        body = json.dumps({"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}], "max_tokens": 120000, "temperature": 0, "ignore_eos": True,
                           "stream": True}).encode()
        s = socket.create_connection(("127.0.0.1", port), timeout=30)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 4096)
        s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body)
        t0 = time.time()
        seen_active = False
        while True:                                        # never read from `s`
            hz = json.loads(T.call(port, "GET", "/api/health")[2])
            seen_active = seen_active or hz["load"]["active_requests"] == 1
            if seen_active and hz["load"]["active_requests"] == 0 and hz["load"]["in_flight_http"] == 0:
                break
            assert time.time() - t0 < 30, hz
            time.sleep(0.05)
        assert hz["kv"]["free_pages"] == hz["kv"]["total_pages"]
        s.close()
        st, _, d = T.call(port, "POST", "/v1/completions", {"model": "tiny-llama", "prompt": "x", "max_tokens": 3})
        assert st == 200
    finally:
        proc.terminate(); proc.wait(timeout=20)


def test_a_buffered_request_whose_client_went_away_is_cancelled(product_bin):
    """Not streamed: nothing is sent until the end, so a send can never notice the dead client.  The shim peeks at the socket
    between polls; a closed peer cancels the generation (the reference's handler future is dropped with the connection, and
    its upstream request with it) instead of running 100 000 tokens for nobody."""
    import socket
    port, proc = _start(product_bin, "--model", "tiny", "--model-id", "tiny-llama", "--max-seqs", "2", "--max-ctx", "131072", "--request-timeout-ms", "600000",
                        env={"FAKE_CUDART_STEP_US": "200"})
    try:
        body = json.dumps({"model": "tiny-llama", "messages": [{"role": "user", "content": "x"}], "max_tokens": 100000, "temperature": 0, "ignore_eos": True}).encode()
        s = socket.create_connection(("127.0.0.1", port), timeout=30)
        s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body)
        t0 = time.time()
        while json.loads(T.call(port, "GET", "/api/health")[2])["load"]["active_requests"] != 1:
            assert time.time() - t0 < 10
            time.sleep(0.02)
        s.close()                                             # 100 000 tokens x 200 us = 20 s of generation left
        t0 = time.time()
        while True:
            hz = json.loads(T.call(port, "GET", "/api/health")[2])
            if hz["load"]["active_requests"] == 0 and hz["load"]["in_flight_http"] == 0:
                break
            assert time.time() - t0 < 5, hz
            time.sleep(0.05)
        assert hz["kv"]["free_pages"] == hz["kv"]["total_pages"]
    finally:
        proc.terminate(); proc.wait(timeout=20)
