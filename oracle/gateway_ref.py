"""Oracle: Python restatement of the gateway-side hot path — test
infrastructure, not product.  The reference is Rust and cannot be compiled here (no cargo), so
each function below restates the cited reference code and is pinned against the known-answer
vectors of the reference's own tests, transcribed in tests/golden/gateway_vectors.json (each
vector cites the reference test it comes from).  The product implementation is the C++ in
llmlb_b200/host/, which tests/test_host_gateway.py compares against this file.

Not restated: tiktoken `estimate_tokens` (llmlb/src/token/mod.rs:217-223 — third-party BPE
tiktoken-rs 0.11.0 cl100k_base, absent here).  It is only the fallback when an endpoint omits
`usage`; the in-process engine always reports usage, so that branch is unreachable on this path.
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
