"""Oracle: Python restatement of the gateway-side hot path — test
infrastructure, not product.  The reference is Rust and cannot be compiled here (no cargo), so
each function below restates the cited reference code and is pinned against the known-answer
vectors of the reference's own tests, transcribed in tests/golden/gateway_vectors.json (each
vector cites the reference test it comes from).  The product implementation is the C++ in
llmlb_b200/host/, which tests/test_host_gateway.py compares against this file.

tiktoken's vocabulary (llmlb/src/token/mod.rs:217-223 — tiktoken-rs 0.11.0 cl100k_base, third-party data absent
here) is not restated: `extract_or_estimate_tokens` takes the counting function as an argument.  It is only the
fallback when an endpoint omits `usage`; the in-process engine always reports usage.

The last section restates the gateway as CLIENT of an endpoint's probe routes (type detection, /api/health,
/v1/models sync, xLLM model info): the conformance checker for the shim's responder side.
"""
import hashlib
import json


# --- llmlb/src/balancer/types.rs:102-118  ModelTpsState::update_tps (EMA alpha = 0.2) ---------
class ModelTpsState:
    ALPHA = 0.2

    def __init__(self):
        self.tps_ema = None
        self.request_count = 0
        self.total_output_tokens = 0
        self.total_duration_ms = 0

    def update_tps(self, output_tokens, duration_ms):
        if duration_ms == 0:
            return
        cur = float(output_tokens) / (float(duration_ms) / 1000.0)
        self.tps_ema = cur if self.tps_ema is None else self.ALPHA * cur + (1.0 - self.ALPHA) * self.tps_ema
        self.request_count += 1
        self.total_output_tokens += output_tokens
        self.total_duration_ms += duration_ms


# --- llmlb/src/models/mapping.rs:38-40,331-345 + registry/endpoints.rs:16-29 ------------------
def _id_eq(a, b):
    return a == b or a.lower() == b.lower()


def find_mapping(model_id, mappings):
    for m in mappings:
        if _id_eq(m["canonical"], model_id) or any(_id_eq(a, model_id) for a in m["aliases"]):
            return m
    return None


def model_lookup_keys(model_id, mappings):
    keys = [model_id]
    m = find_mapping(model_id, mappings)
    if m:
        for k in [m["canonical"]] + list(m["aliases"]):
            if k not in keys:
                keys.append(k)
    return keys


# --- llmlb/src/api/model_name.rs:19-40  parse_quantized_model_name ----------------------------
def parse_quantized_model_name(model):
    pos = model.find(":")
    if pos < 0:
        return {"raw": model, "base": model, "quantization": None}
    if ":" in model[pos + 1:] or pos == 0 or pos == len(model) - 1:
        raise ValueError("Invalid model name (quantization format): " + model)
    return {"raw": model, "base": model[:pos], "quantization": model[pos + 1:]}


# --- llmlb/src/balancer/mod.rs:1873-1985, 3006-3022  TPS-priority selection --------------------
class LoadManager:
    """endpoints keep registration order (the order find_by_model returns them)."""

    def __init__(self, mappings=()):
        self.endpoints = []          # dicts: id, status, initializing, models[list of {model_id, canonical}]
        self.tps = {}                # (endpoint_id, model_id, api_kind) -> ModelTpsState
        self.round_robin = 0
        self.mappings = list(mappings)

    def add_endpoint(self, eid, models, status="online", initializing=False):
        self.endpoints.append({"id": eid, "status": status, "initializing": initializing,
                               "models": [m if isinstance(m, dict) else {"model_id": m, "canonical": None} for m in models]})

    def update_tps(self, eid, model_id, api_kind, output_tokens, duration_ms):
        self.tps.setdefault((eid, model_id, api_kind), ModelTpsState()).update_tps(output_tokens, duration_ms)

    def clear_tps_for_endpoint(self, eid):  # health/endpoint_checker.rs:313-317
        for k in [k for k in self.tps if k[0] == eid]:
            del self.tps[k]

    def find_by_model(self, model_id):      # registry/endpoints.rs:209-231 (online only)
        keys = model_lookup_keys(model_id, self.mappings)
        out = []
        for ep in self.endpoints:
            if ep["status"] != "online":
                continue
            ep_keys = []
            for m in ep["models"]:
                ep_keys += model_lookup_keys(m["model_id"], self.mappings)
                if m.get("canonical"):
                    ep_keys += model_lookup_keys(m["canonical"], self.mappings)
            if any(k in ep_keys for k in keys):
                out.append(ep)
        return out

    def score(self, ep, model_id, api_kind):
        if model_id is not None:
            if api_kind is None:
                return 0.0
            best = 0.0
            for (eid, mid, kind), st in self.tps.items():
                if eid == ep["id"] and mid == model_id and kind == api_kind and st.tps_ema is not None:
                    best = max(best, st.tps_ema)
            return best
        tok = dur = 0
        for (eid, _mid, kind), st in self.tps.items():
            if eid == ep["id"] and (api_kind is None or kind == api_kind):
                tok += st.total_output_tokens
                dur += st.total_duration_ms
        return tok / (dur / 1000.0) if dur > 0 else 0.0

    def select(self, model_id, api_kind):
        eps = self.find_by_model(model_id) if model_id is not None else [e for e in self.endpoints if e["status"] == "online"]
        if not eps:
            raise LookupError("no_capable_endpoints" if model_id is not None else "no_endpoints_available")
        cands = [e for e in eps if not e["initializing"]]
        if not cands:
            raise LookupError("no_endpoints_available")
        cursor = self.round_robin
        self.round_robin += 1
        start = cursor % max(1, len(cands))
        rank = {cands[(start + off) % len(cands)]["id"]: off for off in range(len(cands))}
        ordered = sorted(cands, key=lambda e: (-self.score(e, model_id, api_kind), rank[e["id"]]))
        return ordered[0]["id"]


# --- llmlb/src/token/mod.rs:182-206  extract_usage_from_response ------------------------------
# ---- request leases and completion counters -------------------------------------------------------
# llmlb/src/balancer/mod.rs:2273-2425 (begin_request / finish_request / finish_request_with_tokens),
# llmlb/src/balancer/lease.rs:16-100 (RequestLease: complete* consume the lease; a lease DROPPED
# without complete finishes as Error with the elapsed time), balancer/types.rs:188-195 (average latency).
class EndpointLoadState:
    def __init__(self):
        self.assigned_active = 0
        self.total_assigned = 0
        self.success_count = 0
        self.error_count = 0
        self.total_latency_ms = 0
        self.total_input_tokens = 0
        self.total_output_tokens = 0
        self.total_tokens = 0

    def average_latency_ms(self):
        done = self.success_count + self.error_count
        return None if done == 0 else self.total_latency_ms / done

    def as_list(self):
        return [self.assigned_active, self.total_assigned, self.success_count, self.error_count, self.total_latency_ms,
                self.total_input_tokens, self.total_output_tokens, self.total_tokens]


class LeaseBook:
    """The lease side of LoadManager, keyed by endpoint id."""

    def __init__(self, endpoint_ids):
        self.state = {e: EndpointLoadState() for e in endpoint_ids}

    def begin_request(self, endpoint_id):
        if endpoint_id not in self.state:
            return None                                  # LbError::EndpointNotFound
        st = self.state[endpoint_id]
        st.assigned_active += 1
        st.total_assigned += 1
        return RequestLease(self, endpoint_id)

    def finish_request(self, endpoint_id, outcome, duration_ms, usage=None):
        """outcome: "success" | "error" | "queued"; usage: dict with optional input/output/total or None"""
        if endpoint_id not in self.state:
            return False
        st = self.state[endpoint_id]
        if outcome == "queued":
            return True
        if st.assigned_active > 0:
            st.assigned_active -= 1
        if outcome == "success":
            st.success_count += 1
        else:
            st.error_count += 1
        st.total_latency_ms += int(duration_ms)
        if usage is not None:
            i, o, t = usage.get("input"), usage.get("output"), usage.get("total")
            if i is not None:
                st.total_input_tokens += i
            if o is not None:
                st.total_output_tokens += o
            if t is None and (i is not None or o is not None):
                t = (i or 0) + (o or 0)
            if t is not None:
                st.total_tokens += t
        return True


class RequestLease:
    def __init__(self, book, endpoint_id):
        self.book, self.endpoint_id = book, endpoint_id

    def complete(self, outcome, duration_ms, usage=None):
        book, self.book = self.book, None                # take(): later complete / drop is a no-op
        if book is None:
            return True
        return book.finish_request(self.endpoint_id, outcome, duration_ms, usage)

    def drop(self, elapsed_ms=0):
        """Rust's Drop: a lease that was never completed finishes as Error."""
        book, self.book = self.book, None
        if book is not None:
            book.finish_request(self.endpoint_id, "error", elapsed_ms)


def extract_or_estimate_tokens(body, request_text, response_text, count):
    """llmlb/src/token/mod.rs:235-259.  `count(text) -> int | None` stands for estimate_tokens
    (:217-223; tiktoken cl100k_base in the reference — a third-party rank table that is not in this image;
    the in-process endpoint counts with the model's own tokenizer instead).  Returns dict(input, output, total)."""
    u = extract_usage_from_response(body)
    if u is not None:
        return u
    i = count(request_text) if (request_text is not None and count) else None
    o = count(response_text) if (response_text is not None and count) else None
    if i is not None and o is not None:
        t = i + o
    elif i is not None:
        t = i
    elif o is not None:
        t = o
    else:
        t = None
    return {"input_tokens": i, "output_tokens": o, "total_tokens": t}


def extract_usage_from_response(body):
    if not isinstance(body, dict):
        return None
    usage = body.get("usage")
    if usage is None:
        resp = body.get("response")
        usage = resp.get("usage") if isinstance(resp, dict) else None
    if usage is None:
        return None

    def u(*names):
        if not isinstance(usage, dict):
            return None
        for n in names:
            if n in usage:
                v = usage[n]
                return v if isinstance(v, int) and not isinstance(v, bool) and v >= 0 else None
        return None
    return {"input_tokens": u("prompt_tokens", "input_tokens"),
            "output_tokens": u("completion_tokens", "output_tokens"), "total_tokens": u("total_tokens")}


# --- llmlb/src/token/mod.rs:41-172  StreamingTokenAccumulator ----------------------------------
class StreamingTokenAccumulator:
    def __init__(self, model):
        self.model = model
        self.accumulated_content = ""
        self.input_tokens = None
        self.extracted_usage = None
        self.done = False

    def process_chunk(self, chunk):
        chunk = chunk.strip()
        if not chunk or chunk.startswith(":"):
            return
        if chunk.startswith("data: "):
            data = chunk[len("data: "):]
        elif chunk.startswith("data:"):
            data = chunk[len("data:"):].strip()
        else:
            return
        if data == "[DONE]":
            self.done = True
            return
        try:
            js = json.loads(data)
        except ValueError:
            return
        usage = extract_usage_from_response(js)
        if usage is not None:
            self.extracted_usage = usage
        if isinstance(js, dict):
            choices = js.get("choices")
            if isinstance(choices, list):
                for c in choices:
                    d = c.get("delta") if isinstance(c, dict) else None
                    content = d.get("content") if isinstance(d, dict) else None
                    if isinstance(content, str):
                        self.accumulated_content += content
            t = js.get("type")
            if t == "response.output_text.delta" and isinstance(js.get("delta"), str):
                self.accumulated_content += js["delta"]
            elif t == "response.output_text.done" and not self.accumulated_content and isinstance(js.get("text"), str):
                self.accumulated_content += js["text"]

    def finalize(self, estimate=None):
        """estimate: callable(text) -> tokens standing in for tiktoken (see module docstring)."""
        if self.extracted_usage is not None:
            return dict(self.extracted_usage)
        out = 0 if not self.accumulated_content else (estimate(self.accumulated_content) if estimate else None)
        i = self.input_tokens
        total = (i + out) if (i is not None and out is not None) else (i if i is not None else out)
        return {"input_tokens": i, "output_tokens": out, "total_tokens": total}


# --- llmlb/src/api/proxy.rs:104-116  process_sse_lines (buffer across chunk boundaries) --------
def process_sse_lines(buffer, acc):
    while "\n" in buffer:
        line, buffer = buffer.split("\n", 1)
        acc.process_chunk(line)
    return buffer


# --- llmlb/src/api/openai_util.rs:242-304, inference_gate.rs:177-197 ---------------------------
def openai_error_body(message, error_type="invalid_request_error", status=400):
    return {"error": {"message": message, "type": error_type, "code": status}}


def model_unavailable_body(message, code):
    return {"error": {"message": message, "type": "service_unavailable", "code": code}}


def gate_rejection():
    return 503, {"retry-after": "30"}, openai_error_body("Server is updating. Please retry.", "service_unavailable", 503)


# --- llmlb/src/common/error.rs:41-214 LbError: status_code / error_type / external_message / to_openai_error;
#     llmlb/src/api/error.rs:154-203 AppError::into_response (which message reaches the client) --------------------
# kind -> (status, OpenAI error type, external message, AppError exposes the detail text?)
LB_ERRORS = {
    "common_validation":    (400, "invalid_request_error", "Request error", True),
    "common_other":         (400, "invalid_request_error", "Request error", False),
    "endpoint_not_found":   (404, "not_found_error", "Endpoint not found", False),
    "not_found":            (404, "not_found_error", "Not found", True),
    "no_endpoints_available": (503, "service_unavailable", "No available endpoints", False),
    "no_capable_endpoints": (404, "not_found_error", "No capable endpoints", False),
    "database":             (500, "server_error", "Database error", False),
    "http":                 (502, "service_unavailable", "Backend service unavailable", False),
    "timeout":              (504, "server_error", "Request timeout", False),
    "service_unavailable":  (503, "service_unavailable", "Service temporarily unavailable", False),
    "internal":             (500, "server_error", "Internal server error", False),
    "endpoint_offline":     (503, "service_unavailable", "Endpoint offline", False),
    "invalid_model_name":   (400, "invalid_request_error", "Invalid model name", True),
    "insufficient_storage": (507, "server_error", "Insufficient storage", True),
    "password_hash":        (401, "authentication_error", "Authentication error", False),
    "jwt":                  (401, "authentication_error", "Authentication error", False),
    "authentication":       (401, "authentication_error", "Authentication failed", True),
    "authorization":        (403, "permission_error", "Access denied", True),
    "conflict":             (409, "invalid_request_error", "Resource conflict", True),
}


def lb_error_openai(kind):
    """LbError::to_openai_error (common/error.rs:206-214): the code is the status AS A STRING."""
    status, etype, ext, _ = LB_ERRORS[kind]
    return status, {"error": {"message": ext, "type": etype, "code": str(status)}}


def app_error_response(kind, detail=""):
    """AppError::into_response (api/error.rs:154-203): errors that may carry internal details (addresses, ports, DB text)
    answer with the generic external message; developer-crafted ones pass their text through.  A non-validation
    CommonError passes through only when it is the GPU-requirement message."""
    status, _, ext, expose = LB_ERRORS[kind]
    if kind == "common_other":
        expose = "GPU is required" in detail or "GPU hardware is required" in detail
    return status, {"error": detail if expose else ext}


# --- llmlb/src/api/openai_util.rs:86-134 classify_upstream_request_error ---------------------------------------------
def classify_upstream_request_error(kind, timeout_secs, ollama_loading_model=None):
    """kind: "timeout" | "connect" | anything else (reqwest's is_timeout / is_connect).  -> (status, type, client message)"""
    if kind == "timeout":
        if ollama_loading_model is not None:
            return 504, "model_loading", ("Ollama model '%s' is still loading. Retry after the initial load finishes or increase endpoint "
                                          "inference timeout above %d seconds." % (ollama_loading_model, timeout_secs))
        return 504, "timeout", "Upstream endpoint request timed out after %d seconds" % timeout_secs
    if kind == "connect":
        return 502, "connection_error", "Failed to connect to upstream endpoint"
    return 502, "endpoint_request_error", "Failed to proxy request to upstream endpoint"


# --- llmlb/src/api/openai_util.rs:264-292 queue_error_response; call sites openai.rs:841-882 ------------------------
def queue_error_response(status, message, error_type, retry_after=None):
    headers = {} if retry_after is None else {"retry-after": str(retry_after)}
    return status, headers, {"error": {"message": message, "type": error_type, "code": status}}


def queue_capacity_exceeded(queue_timeout_secs):
    """QueueSelection::CapacityExceeded (openai.rs:841-861): 429 rate_limit_exceeded, Retry-After = max(1, queue timeout)."""
    return queue_error_response(429, "Request queue is full", "rate_limit_exceeded", max(1, int(queue_timeout_secs)))


def queue_wait_timeout():
    """QueueSelection::Timeout (openai.rs:863-882): 504 timeout, no Retry-After."""
    return queue_error_response(504, "Queue wait timeout", "timeout", None)


# --- llmlb/src/auth/middleware.rs:292-321 extract_api_key; :254-289 SHA-256 lookup -------------
def extract_api_key(headers):
    h = {k.lower(): v for k, v in headers.items()}
    if "x-api-key" in h:
        return h["x-api-key"]
    if "authorization" in h:
        if h["authorization"].startswith("Bearer "):
            return h["authorization"][len("Bearer "):]
        raise PermissionError("Invalid Authorization header format. Expected 'Bearer <token>'")
    raise PermissionError("Missing X-API-Key header or Authorization header")


def api_key_hash(key):
    return hashlib.sha256(key.encode()).hexdigest()


# --- llmlb/src/api/benchmarks.rs:467-474, proxy.rs:154-160  TPS formula ------------------------
def request_tps(output_tokens, duration_ms):
    return None if duration_ms == 0 else output_tokens / (duration_ms / 1000.0)


# =============================================================================================
# Anthropic Messages front door — llmlb/src/api/anthropic.rs
# =============================================================================================
class AnthropicError(Exception):
    """anthropic_error_response (anthropic.rs:1559-1575): {"type":"error","error":{"type","message"}}"""

    def __init__(self, status, error_type, message):
        super().__init__(message)
        self.status, self.error_type, self.message = status, error_type, message

    def body(self):
        return {"type": "error", "error": {"type": self.error_type, "message": self.message}}


def _bad(msg):
    return AnthropicError(400, "invalid_request_error", msg)


# --- anthropic.rs:1388-1398 extract_required_header ---------------------------------------------
def anthropic_required_header(headers, name):
    v = headers.get(name)
    if v is None or not v.strip():
        raise _bad("Missing required header: %s" % name)
    return v


# --- anthropic.rs:1366-1386 extract_model ---------------------------------------------------------
def anthropic_extract_model(payload):
    m = payload.get("model") if isinstance(payload, dict) else None
    if not isinstance(m, str):
        raise _bad("model is required")
    if not m.strip():
        raise _bad("model must not be empty")
    return m


# --- anthropic.rs:1323-1364 flatten_anthropic_text_content ---------------------------------------
def flatten_anthropic_text_content(value, field):
    if isinstance(value, str):
        return value
    if isinstance(value, list):
        text = ""
        for item in value:
            t = item.get("type") if isinstance(item, dict) else None
            if not isinstance(t, str):
                raise _bad("%s content blocks must have a type" % field)
            if t != "text":
                raise _bad("%s content block type '%s' is not supported" % (field, t))
            bt = item.get("text")
            if not isinstance(bt, str):
                raise _bad("%s text content blocks must include text" % field)
            text += bt
        return text
    raise _bad("%s must be a string or text content array" % field)


# --- anthropic.rs:1218-1259 convert_anthropic_tool_to_openai --------------------------------------
def convert_anthropic_tool_to_openai(tool):
    name = tool.get("name") if isinstance(tool, dict) else None
    if not isinstance(name, str):
        raise _bad("tool.name is required")
    desc = tool.get("description")
    desc = desc if isinstance(desc, str) else ""
    if "input_schema" not in tool:
        raise _bad("tool.input_schema is required")
    schema = tool["input_schema"]
    params = {}
    if isinstance(schema, dict):
        for k in ("type", "properties", "required"):
            if k in schema:
                params[k] = schema[k]
    return {"type": "function", "function": {"name": name, "description": desc, "parameters": params}}


# --- anthropic.rs:1261-1298 convert_anthropic_tool_choice_to_openai -------------------------------
def convert_anthropic_tool_choice_to_openai(tc):
    t = tc.get("type") if isinstance(tc, dict) else None
    if not isinstance(t, str):
        raise _bad("tool_choice.type is required")
    if t == "auto":
        return "auto"
    if t == "any":
        return "required"
    if t == "tool":
        name = tc.get("name")
        if not isinstance(name, str):
            raise _bad("tool_choice.name is required when type is 'tool'")
        return {"type": "function", "function": {"name": name}}
    raise _bad("unknown tool_choice type: %s" % t)


# --- anthropic.rs:1048-1216 anthropic_request_to_openai -------------------------------------------
def anthropic_request_to_openai(payload):
    """-> (openai_payload, request_text, stream)"""
    model = anthropic_extract_model(payload)
    mt = payload.get("max_tokens")
    if isinstance(mt, bool) or not isinstance(mt, int) or mt < 0:       # Value::as_u64
        raise _bad("max_tokens is required")
    stream = payload.get("stream")
    stream = stream if isinstance(stream, bool) else False
    msgs = payload.get("messages")
    if not isinstance(msgs, list):
        raise _bad("messages must be an array")
    parts, out = [], []
    if "system" in payload:
        st = flatten_anthropic_text_content(payload["system"], "system")
        if st:
            out.append({"role": "system", "content": st})
            parts.append("system: %s" % st)
    for i, m in enumerate(msgs):
        role = m.get("role") if isinstance(m, dict) else None
        if not isinstance(role, str):
            raise _bad("messages[%d].role is required" % i)
        if role not in ("user", "assistant"):
            raise _bad("messages[%d].role must be 'user' or 'assistant'" % i)
        if "content" not in m:
            raise _bad("messages[%d].content is required" % i)
        content = m["content"]
        if role == "assistant" and isinstance(content, list) and any(
                isinstance(it, dict) and it.get("type") == "tool_use" for it in content):
            continue
        if role == "user" and isinstance(content, list) and any(
                isinstance(it, dict) and it.get("type") == "tool_result" for it in content):
            for it in content:
                if isinstance(it, dict) and it.get("type") == "tool_result":
                    tid = it.get("tool_use_id")
                    tid = tid if isinstance(tid, str) else "unknown"
                    rc = it.get("content")
                    rc = rc if isinstance(rc, str) else ""
                    out.append({"role": "tool", "tool_call_id": tid, "content": rc})
                    parts.append("tool_result[%s]: %s" % (tid, rc))
            continue
        text = flatten_anthropic_text_content(content, "messages[%d].content" % i)
        out.append({"role": role, "content": text})
        parts.append("%s: %s" % (role, text))
    body = {"model": model, "messages": out, "max_tokens": mt, "stream": stream}
    t = payload.get("temperature")
    if isinstance(t, (int, float)) and not isinstance(t, bool):
        body["temperature"] = float(t)
    tp = payload.get("top_p")
    if isinstance(tp, (int, float)) and not isinstance(tp, bool):
        body["top_p"] = float(tp)
    if "stop_sequences" in payload:
        ss = payload["stop_sequences"]
        if not isinstance(ss, list) or any(not isinstance(s, str) for s in ss):
            raise _bad("stop_sequences must be an array of strings")
        body["stop"] = list(ss)
    if isinstance(payload.get("tools"), list):
        body["tools"] = [convert_anthropic_tool_to_openai(t_) for t_ in payload["tools"]]
    if "tool_choice" in payload:
        body["tool_choice"] = convert_anthropic_tool_choice_to_openai(payload["tool_choice"])
    return body, "\n".join(parts), stream


# --- anthropic.rs:1526-1533 -------------------------------------------------------------------------
def map_finish_reason_to_stop_reason(fr):
    return {"length": "max_tokens", "stop": "end_turn", "tool_calls": "tool_use"}.get(fr, "end_turn")


# --- anthropic.rs:1415-1433 -------------------------------------------------------------------------
def convert_openai_tool_call_to_anthropic_tool_use(tc):
    func = tc.get("function") if isinstance(tc, dict) else None
    if not isinstance(func, dict):
        return None
    name, tid = func.get("name"), tc.get("id")
    if not isinstance(name, str) or not isinstance(tid, str):
        return None
    args = func.get("arguments")
    args = args if isinstance(args, str) else "{}"
    try:
        inp = json.loads(args)
    except ValueError:
        inp = {}
    return {"type": "tool_use", "id": tid, "name": name, "input": inp}


# --- anthropic.rs:1435-1524 openai_to_anthropic_message_response ------------------------------------
def openai_to_anthropic_message_response(body, model, input_tokens, output_tokens, fallback_id="msg_0"):
    choices = body.get("choices") if isinstance(body, dict) else None
    choice = choices[0] if isinstance(choices, list) and choices else None
    fr = choice.get("finish_reason") if isinstance(choice, dict) else None
    fr = fr if isinstance(fr, str) else None
    msg = choice.get("message") if isinstance(choice, dict) else None
    text = ""
    if isinstance(choice, dict):
        if isinstance(msg, dict) and isinstance(msg.get("content"), str):
            text = msg["content"]
        elif isinstance(choice.get("text"), str):
            text = choice["text"]
    content = []
    if text:
        content.append({"type": "text", "text": text})
    if isinstance(msg, dict) and isinstance(msg.get("tool_calls"), list):
        for tc in msg["tool_calls"]:
            blk = convert_openai_tool_call_to_anthropic_tool_use(tc)
            if blk is not None:
                content.append(blk)
    if not content:
        content.append({"type": "text", "text": ""})
    stop = "tool_use" if fr == "tool_calls" else (map_finish_reason_to_stop_reason(fr) if fr is not None else "end_turn")
    rid = body.get("id") if isinstance(body, dict) and isinstance(body.get("id"), str) else fallback_id
    return {"id": rid, "type": "message", "role": "assistant", "model": model, "content": content, "stop_reason": stop,
            "stop_sequence": None, "usage": {"input_tokens": input_tokens or 0, "output_tokens": output_tokens or 0}}


# --- anthropic.rs:813-1018 AnthropicStreamTracker (OpenAI chat SSE lines -> Anthropic events) -------
class AnthropicStreamTransformer:
    def __init__(self, model, input_tokens=None, response_id="msg_0"):
        self.acc = StreamingTokenAccumulator(model)
        self.acc.input_tokens = input_tokens
        self.model, self.response_id = model, response_id
        self.line_buf = ""
        self.started = self.block_started = self.block_stopped = self.stopped = False
        self.stop_reason = None
        self.out = []           # [(event_name, data dict)]

    def _emit(self, name, data):
        self.out.append((name, data))

    def feed(self, text):
        self.line_buf += text
        while "\n" in self.line_buf:
            line, self.line_buf = self.line_buf.split("\n", 1)
            self.process_line(line.rstrip("\r"))

    def _ensure_start(self):
        if self.started:
            return
        self.started = True
        self._emit("message_start", {"type": "message_start", "message": {
            "id": self.response_id, "type": "message", "role": "assistant", "content": [], "model": self.model,
            "stop_reason": None, "stop_sequence": None,
            "usage": {"input_tokens": self.acc.finalize()["input_tokens"] or 0, "output_tokens": 0}}})

    def _ensure_block(self):
        if self.block_started:
            return
        self.block_started = True
        self._emit("content_block_start", {"type": "content_block_start", "index": 0, "content_block": {"type": "text", "text": ""}})

    def process_line(self, line):
        self.acc.process_chunk(line)
        t = line.strip()
        if not t or t.startswith(":") or not t.startswith("data:"):
            return
        data = t[len("data:"):].strip()
        if data == "[DONE]":
            self.finish()
            return
        try:
            js = json.loads(data)
        except ValueError:
            return
        if isinstance(js, dict) and isinstance(js.get("id"), str):
            self.response_id = js["id"].replace("chatcmpl-", "msg_").replace("chatcmpl", "msg")
        self._ensure_start()
        choices = js.get("choices") if isinstance(js, dict) else None
        choice = choices[0] if isinstance(choices, list) and choices else None
        if not isinstance(choice, dict):
            return
        delta = choice.get("delta")
        content = delta.get("content") if isinstance(delta, dict) else None
        if isinstance(content, str):
            self._ensure_block()
            if content:
                self._emit("content_block_delta", {"type": "content_block_delta", "index": 0,
                                                   "delta": {"type": "text_delta", "text": content}})
        tcs = delta.get("tool_calls") if isinstance(delta, dict) else None
        if isinstance(tcs, list) and tcs:
            if self.block_started and not self.block_stopped:
                self.block_stopped = True
                self._emit("content_block_stop", {"type": "content_block_stop", "index": 0})
            for idx, tc in enumerate(tcs):
                blk = convert_openai_tool_call_to_anthropic_tool_use(tc)
                if blk is not None:
                    self._emit("content_block_start", {"type": "content_block_start", "index": 1 + idx, "content_block": blk})
                    self._emit("content_block_stop", {"type": "content_block_stop", "index": 1 + idx})
        fr = choice.get("finish_reason")
        if isinstance(fr, str):
            self.stop_reason = map_finish_reason_to_stop_reason(fr)

    def finish(self):
        if self.stopped:
            return
        self._ensure_start()
        self._ensure_block()
        if not self.block_stopped:
            self.block_stopped = True
            self._emit("content_block_stop", {"type": "content_block_stop", "index": 0})
        u = self.acc.finalize()
        self._emit("message_delta", {"type": "message_delta",
                                     "delta": {"stop_reason": self.stop_reason or "end_turn", "stop_sequence": None},
                                     "usage": {"output_tokens": u["output_tokens"] or 0}})
        self._emit("message_stop", {"type": "message_stop"})
        self.stopped = True

    def wire(self):
        """emit_event (anthropic.rs:1015-1018): 'event: <name>\\ndata: <json>\\n\\n'"""
        return "".join("event: %s\ndata: %s\n\n" % (n, json.dumps(d, separators=(",", ":"), ensure_ascii=False)) for n, d in self.out)


# =============================================================================================
# Outbound payload preparation
# =============================================================================================
# --- llmlb/src/models/mapping.rs:302-323 resolve_engine_name ----------------------------------
def resolve_engine_name(model, endpoint_type, mappings):
    """mappings: [{"canonical", "aliases": [...], "engines": {alias: endpoint_type}}] — the first
    alias registered for `endpoint_type` of the mapping that knows `model`, else None."""
    m = find_mapping(model, mappings)
    if not m:
        return None
    engines = m.get("engines", {})
    for a in m["aliases"]:
        if engines.get(a) == endpoint_type:
            return a
    return None


# --- llmlb/src/api/model_name.rs:43-80 resolve_runtime_model_name(_for_endpoint) -------------
def resolve_runtime_model_name_for_endpoint(requested, selected, endpoint_type, endpoint_models, mappings):
    """endpoint_models: [(model_id, canonical_name or None)] as the endpoint advertises them"""
    if any(mid == requested for mid, _ in endpoint_models):
        return requested
    for mid, canon in endpoint_models:
        if mid == selected:
            return mid
        if canon is not None and (canon == selected or canon == requested):
            return mid
    return resolve_engine_name(selected, endpoint_type, mappings) or selected


# --- llmlb/src/api/model_name.rs:82-108 rewrite_payload_model_for_endpoint -------------------
def rewrite_payload_model_for_endpoint(payload, selected, endpoint_type, endpoint_models, mappings):
    requested = payload.get("model") if isinstance(payload, dict) else None
    if not isinstance(requested, str):
        return payload
    runtime = resolve_runtime_model_name_for_endpoint(requested, selected, endpoint_type, endpoint_models, mappings)
    if runtime == requested:
        return payload
    out = dict(payload)
    out["model"] = runtime
    return out


# --- llmlb/src/api/openai.rs:977-992: upstream model + stream_options.include_usage -----------
def prepare_upstream_payload(payload, upstream_model, stream):
    out = dict(payload)
    out["model"] = upstream_model
    if stream:
        opts = out.get("stream_options")
        if opts is None or "stream_options" not in out:
            out["stream_options"] = {"include_usage": True}
        elif isinstance(opts, dict):
            opts = dict(opts)
            opts.setdefault("include_usage", True)
            out["stream_options"] = opts
        # a non-object stream_options is left as the client sent it (as_object_mut() is None)
    return out


# =============================================================================================
# 60-minute request history — llmlb/src/balancer/mod.rs:2643-2658, 2973-3060
# =============================================================================================
REQUEST_HISTORY_WINDOW_MINUTES = 60          # balancer/types.rs:22


def align_to_minute(ts):                     # mod.rs:2973-2975 (ts: unix seconds)
    return ts - ts % 60


class RequestHistory:
    """Per-minute (success, error) counters: the newest minute is incremented in place, a new minute
    appends a point, points older than the 60-minute window (relative to the newest) are dropped;
    'queued' outcomes leave the counters alone.  window(now) returns exactly 60 points, oldest
    first, ending at now's minute, zero-filled."""

    def __init__(self):
        self.points = []                     # [[minute, success, error]]

    def record(self, outcome, ts):           # mod.rs:2643-2658
        minute = align_to_minute(ts)
        if self.points and self.points[-1][0] == minute:
            self._inc(self.points[-1], outcome)
        else:
            p = [minute, 0, 0]
            self._inc(p, outcome)
            self.points.append(p)
        cutoff = minute - 60 * (REQUEST_HISTORY_WINDOW_MINUTES - 1)      # prune_history :2977-2986
        while self.points and self.points[0][0] < cutoff:
            self.points.pop(0)

    @staticmethod
    def _inc(p, outcome):                    # increment_history :2998-3004
        if outcome == "success":
            p[1] += 1
        elif outcome == "error":
            p[2] += 1

    def window(self, now):                   # build_history_window / fill_history :3024-3060
        now = align_to_minute(now)
        have = {p[0]: p for p in self.points}
        start = now - 60 * (REQUEST_HISTORY_WINDOW_MINUTES - 1)
        return [list(have.get(m, [m, 0, 0])) for m in range(start, now + 60, 60)]


# =============================================================================================
# Inference latency EMA — llmlb/src/types/endpoint.rs:419-440
# =============================================================================================
class InferenceLatency:
    ALPHA = 0.2

    def __init__(self):
        self.ms = None                       # Option<f64>

    def update(self, new_ms):                # endpoint.rs:419-427: first sample (or after a reset) = the sample
        cur = self.ms
        self.ms = (self.ALPHA * new_ms + (1.0 - self.ALPHA) * cur) if (cur is not None and cur != float("inf") and cur == cur) else new_ms

    def reset(self):                         # endpoint.rs:433-435: offline => sorts last
        self.ms = float("inf")

    def for_sort(self):                      # endpoint.rs:438-440
        return float("inf") if self.ms is None else self.ms


# =============================================================================================
# The gateway as CLIENT of an endpoint's probe routes: what an
# unmodified llmlb concludes about whatever answers at base_url.  The product is the RESPONDER
# (llmlb_b200/host/server.cpp); these restatements are the conformance checker: the shim's answers
# are fed through them in tests/test_endpoint_conformance*.py.
#
# `fetch(path, auth)` -> None on a connection error, else (status, headers, json_or_None);
# headers: lower-cased names; auth: True when the reference sends `Authorization: Bearer <key>`.
# =============================================================================================
def _ok(r):
    return r is not None and 200 <= r[0] < 300


def _lm_studio_marker(value):                # detection/lm_studio.rs:133-160 marker_tokens / has_lm_studio_marker
    toks, cur = [], ""
    for ch in value:
        if ch.isascii() and ch.isalnum():
            cur += ch.lower()
        else:
            if cur:
                toks.append(cur)
            cur = ""
    if cur:
        toks.append(cur)
    return "lmstudio" in toks or any(a == "lm" and b == "studio" for a, b in zip(toks, toks[1:]))


def _looks_like_lm_studio_model(m):          # detection/lm_studio.rs:115-131
    if not isinstance(m, dict):
        return False
    s = lambda k: isinstance(m.get(k), str)
    has_state = "state" in m or isinstance(m.get("loaded_instances"), list)
    has_shape = s("key") or s("display_name") or "format" in m or "compatibility_type" in m
    return s("publisher") and (s("arch") or s("architecture")) and (has_state or has_shape)


def detect_xllm(fetch):                      # detection/xllm.rs:27-66
    r = fetch("/api/system", True)
    if _ok(r) and isinstance(r[2], dict):
        v = r[2].get("xllm_version")
        sn = r[2].get("server_name")
        if (v is None or isinstance(v, str)) and (sn is None or isinstance(sn, str)):   # serde: Option<String> fields must be strings or null
            if isinstance(v, str):
                return "xLLM: /api/system xllm_version=%s" % v
    return None


def detect_lm_studio(fetch):                 # detection/lm_studio.rs:19-97
    r = fetch("/api/v1/models", True)
    if _ok(r) and isinstance(r[2], dict):
        for key in ("data", "models"):
            arr = r[2].get(key)
            if isinstance(arr, list) and any(_looks_like_lm_studio_model(m) for m in arr):
                return "LM Studio: /api/v1/models returned LM Studio format"
    r = fetch("/v1/models", True)
    if r is None:
        return None
    server = r[1].get("server")
    if server is not None and _lm_studio_marker(server):
        return "LM Studio: Server header contains lm-studio (%s)" % server
    if _ok(r) and isinstance(r[2], dict) and isinstance(r[2].get("data"), list):
        for m in r[2]["data"]:
            ob = m.get("owned_by") if isinstance(m, dict) else None
            if isinstance(ob, str) and _lm_studio_marker(ob):
                return "LM Studio: owned_by field contains LM Studio marker"
    return None


def detect_ollama(fetch):                    # detection/ollama.rs (no Authorization header on this probe)
    r = fetch("/api/tags", False)
    if not _ok(r) or not isinstance(r[2], dict):
        return None
    j = r[2]
    models = j.get("models")
    if models is not None:                   # serde: Option<Vec<OllamaModel{name: String, size: Option<i64>}>>
        if not isinstance(models, list):
            return None
        for m in models:
            if not isinstance(m, dict) or not isinstance(m.get("name"), str):
                return None
            if m.get("size") is not None and (isinstance(m["size"], bool) or not isinstance(m["size"], int)):
                return None
    if j.get("error") is not None:
        return None
    return "Ollama: /api/tags returned models" if models is not None else None


def detect_vllm(fetch):                      # detection/vllm.rs
    r = fetch("/v1/models", True)
    if r is None:
        return None
    server = r[1].get("server")
    if server is not None and "vllm" in server.lower():
        return "vLLM: Server header contains vllm (%s)" % server
    if _ok(r) and isinstance(r[2], dict) and isinstance(r[2].get("data"), list):
        for m in r[2]["data"]:
            ob = m.get("owned_by") if isinstance(m, dict) else None
            if isinstance(ob, str) and "vllm" in ob.lower():
                return "vLLM: owned_by field contains vllm"
    return None


def detect_llamacpp(fetch):                  # detection/llama_cpp.rs:29-96 (no Authorization header)
    r = fetch("/v1/models", False)
    if r is not None:
        server = r[1].get("server")
        if server is not None and "llama.cpp" in server:
            return "llama.cpp: Server header contains llama.cpp (%s)" % server
    r = fetch("/v1/version", False)
    if _ok(r) and isinstance(r[2], dict):
        s, v = r[2].get("server"), r[2].get("version")
        if (s is None or isinstance(s, str)) and (v is None or isinstance(v, str)) and isinstance(s, str) and "llama.cpp" in s:
            return "llama.cpp: /v1/version server field is '%s'" % s
    return None


def detect_endpoint_type(fetch):
    """detection/mod.rs:85-195: priority xLLM > LM Studio > Ollama > vLLM > llama.cpp > OpenAI-compatible.
    Returns (endpoint_type, reason); raises ValueError("unreachable" | "unsupported")."""
    for name, fn in (("xllm", detect_xllm), ("lm_studio", detect_lm_studio), ("ollama", detect_ollama), ("vllm", detect_vllm),
                     ("llamacpp", detect_llamacpp)):
        reason = fn(fetch)
        if reason:
            return name, reason
    r = fetch("/v1/models", True)            # detect_openai_compatible :208-238
    if r is not None:
        if _ok(r) and isinstance(r[2], dict) and ("data" in r[2] or "object" in r[2]):
            return "openai_compatible", "OpenAI-compatible: /v1/models responded 200"
        raise ValueError("unsupported")
    if fetch("/v1/models", False) is not None:
        raise ValueError("unsupported")
    raise ValueError("unreachable")


def parse_models_response(j):                # sync/parser.rs:78-110
    if isinstance(j, dict) and isinstance(j.get("data"), list):
        return [m["id"] for m in j["data"] if isinstance(m, dict) and isinstance(m.get("id"), str)], "openai"
    if isinstance(j, dict) and isinstance(j.get("models"), list):
        out = []
        for m in j["models"]:
            if not isinstance(m, dict):
                continue
            i = m.get("name") if isinstance(m.get("name"), str) else (m.get("model") if isinstance(m.get("model"), str) else None)
            if i:
                out.append(i)
        return out, "ollama"
    return [], "unknown"


def detect_capabilities(model_name):         # sync/capabilities.rs:47-57
    leaf = model_name.lower().rsplit("/", 1)[-1]
    return ["embeddings"] if (leaf.startswith("embed") or "-embed" in leaf or "_embed" in leaf) else ["chat"]


def parse_v0_health(r):
    """health/endpoint_checker.rs:515-557 try_v0_health: Err on non-2xx / non-JSON, else GpuInfo with every field optional."""
    if not _ok(r) or r[2] is None:
        raise ValueError("HTTP %s" % (r[0] if r is not None else "error"))
    body = r[2]
    gpu = body.get("gpu") if isinstance(body, dict) else None
    load = body.get("load") if isinstance(body, dict) else None

    def u64(o, k, bits=64):
        v = o.get(k) if isinstance(o, dict) else None
        if isinstance(v, bool) or not isinstance(v, int) or v < 0 or v >= 2 ** 64:
            return None
        return v & (2 ** bits - 1)                                   # `as u32` truncates

    def f32(o, k):
        v = o.get(k) if isinstance(o, dict) else None
        return float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else None

    return {"gpu_device_count": u64(gpu, "device_count", 32), "gpu_total_memory_bytes": u64(gpu, "total_memory_bytes"),
            "gpu_used_memory_bytes": u64(gpu, "used_memory_bytes"), "gpu_capability_score": f32(gpu, "capability_score"),
            "active_requests": u64(load, "active_requests", 32)}


def xllm_model_info_url(model):              # metadata/xllm.rs:54-63: three characters are escaped, nothing else
    return "/api/models/%s/info" % model.replace(" ", "%20").replace("/", "%2F").replace(":", "%3A")


def parse_xllm_model_info(r):
    """metadata/xllm.rs:11-99: `model` is a required string; context_length | n_ctx (u32), size_bytes | size | file_size (u64),
    quantization | quant | quantization_type, family, parameter_size | params | num_params are optional (null allowed)."""
    if r is None:
        raise ValueError("request failed")
    if not _ok(r):
        raise ValueError("endpoint error %d" % r[0])
    j = r[2]
    if not isinstance(j, dict) or not isinstance(j.get("model"), str):
        raise ValueError("invalid response")

    def pick(names, kind, hi=None):
        present = [n for n in names if n in j]
        if len(present) > 1:
            raise ValueError("invalid response")                    # serde: duplicate field through aliases
        if not present or j[present[0]] is None:
            return None
        v = j[present[0]]
        if kind is str:
            if not isinstance(v, str):
                raise ValueError("invalid response")
            return v
        if isinstance(v, bool) or not isinstance(v, int) or v < 0 or v > hi:
            raise ValueError("invalid response")
        return v

    return {"model": j["model"], "context_length": pick(("context_length", "n_ctx"), int, 2 ** 32 - 1),
            "size_bytes": pick(("size_bytes", "size", "file_size"), int, 2 ** 64 - 1),
            "quantization": pick(("quantization", "quant", "quantization_type"), str), "family": pick(("family",), str),
            "parameter_size": pick(("parameter_size", "params", "num_params"), str)}


def sync_models(fetch, endpoint_type):
    """sync/mod.rs:104-278 reduced to what it learns from the endpoint: GET /v1/models -> ids, capabilities by name,
    supported_apis = [chat_completions]; for xllm (also ollama / lm_studio, whose metadata clients are not restated)
    max_tokens = context_length of GET /api/models/{id}/info when that succeeds."""
    r = fetch("/v1/models", True)
    if not _ok(r) or r[2] is None:
        raise ValueError("sync failed")
    ids, fmt = parse_models_response(r[2])
    out = []
    for i in ids:
        m = {"model_id": i, "capabilities": detect_capabilities(i), "supported_apis": ["chat_completions"], "max_tokens": None}
        if endpoint_type == "xllm":
            try:
                m["max_tokens"] = parse_xllm_model_info(fetch(xllm_model_info_url(i), True))["context_length"]
            except ValueError:
                pass
        out.append(m)
    return out, fmt
