"""A gateway made of NOTHING BUT the oracle's restatements (oracle/gateway_ref.py) glued in the reference's order — TEST
INFRASTRUCTURE.  It stands where the unmodified llmlb would stand in front of an endpoint: registration by probing
(api/endpoints.rs:505-704 -> detection, sync), then `proxy_openai_post` (api/openai.rs:761-1338) and `post_responses`
(api/responses.rs:143-431): model lookup, TPS-priority selection, lease, payload model rewrite + stream_options.include_usage,
the HTTP POST with the endpoint's bearer key, byte-transparent SSE relay with the accumulator on the side, usage -> lease +
TPS update, upstream failures mapped the way the two routes map them.  tests/test_gateway_e2e_cpu.py drives the shim with it."""
import http.client
import json
import time

from oracle import gateway_ref as G


class MiniGateway:
    def __init__(self):
        self.lm = G.LoadManager()
        self.book = G.LeaseBook([])
        self.eps = {}                                       # id -> {"port", "api_key", "type", "models", "timeout_s"}

    # ---- POST /api/endpoints {"name","base_url","api_key"?}: detect the type, sync the models ------------------------
    def register(self, name, port, api_key=None, timeout_s=120):
        def fetch(path, auth):
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=5)            # DETECTION_TIMEOUT
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
        try:
            etype, reason = G.detect_endpoint_type(fetch)
        except ValueError as e:
            return (502 if str(e) == "unreachable" else 400), {"error": str(e)}         # endpoint_type_detection_test.rs:51-70
        models, _ = G.sync_models(fetch, etype)
        self.eps[name] = {"port": port, "api_key": api_key, "type": etype, "models": models, "timeout_s": timeout_s, "reason": reason}
        self.lm.add_endpoint(name, [m["model_id"] for m in models])
        self.book.state[name] = G.EndpointLoadState()
        return 201, {"id": name, "endpoint_type": etype, "models": models}

    def set_status(self, name, status):
        for e in self.lm.endpoints:
            if e["id"] == name:
                e["status"] = status
        if status != "online":
            self.lm.clear_tps_for_endpoint(name)                                          # health/endpoint_checker.rs:313-317

    # ---- the two inference routes --------------------------------------------------------------------------------------
    def post(self, path, payload):
        """-> (status, headers, body bytes).  path: /v1/chat/completions | /v1/completions | /v1/responses"""
        kind = {"/v1/chat/completions": "chat_completions", "/v1/completions": "completions", "/v1/responses": "responses"}[path]
        model = payload.get("model")
        if not isinstance(model, str) or not model:
            return self._json(400, G.openai_error_body("model is required", "invalid_request_error", 400))
        try:
            base = G.parse_quantized_model_name(model)["base"]
        except ValueError as e:
            return self._json(400, G.openai_error_body(str(e), "invalid_request_error", 400))
        stream = payload.get("stream") is True
        if not self.lm.find_by_model(base):
            known = any(m["model_id"] == base for e in self.eps.values() for m in e["models"])
            if not known:                                                                   # openai.rs:805-818
                return self._json(404, G.openai_error_body("The model '%s' does not exist" % model, "invalid_request_error", 404))
        try:
            eid = self.lm.select(base, kind)
        except LookupError as e:
            if str(e) == "no_capable_endpoints":                                            # openai.rs:908-913
                return self._json(503, G.model_unavailable_body("No available endpoints support model: %s" % base, "no_capable_nodes"))
            return self._json(503, G.openai_error_body("No endpoints available", "service_unavailable", 503))
        ep = self.eps[eid]
        lease = self.book.begin_request(eid)
        emodels = [(m["model_id"], None) for m in ep["models"]]
        up = G.rewrite_payload_model_for_endpoint(payload, base, ep["type"], emodels, [])
        if path != "/v1/responses":                                                         # responses.rs passes the payload through
            up = G.prepare_upstream_payload(up, up["model"], stream)
        t0 = time.monotonic()
        try:
            c = http.client.HTTPConnection("127.0.0.1", ep["port"], timeout=ep["timeout_s"])
            hdr = {"Content-Type": "application/json"}
            if ep["api_key"]:
                hdr["Authorization"] = "Bearer " + ep["api_key"]                            # proxy.rs:390-392
            c.request("POST", path, json.dumps(up).encode(), hdr)
            r = c.getresponse()
        except OSError as e:
            ms = int((time.monotonic() - t0) * 1000)
            lease.complete("error", ms)
            st, etype, msg = G.classify_upstream_request_error("timeout" if "timed out" in str(e) else "connect", ep["timeout_s"])
            return self._json(st, G.openai_error_body(msg, etype, st))
        status = r.status
        if not 200 <= status < 300:
            body = r.read()
            c.close()
            lease.complete("error", int((time.monotonic() - t0) * 1000))
            if path == "/v1/responses":                                                     # responses.rs:411-424: pass-through
                return status, {"content-type": r.getheader("content-type") or "application/json"}, body
            msg = body.decode("utf-8", "replace").strip() or str(status)                     # openai.rs:1178-1213
            return self._json(502, {"error": {"message": msg, "type": "endpoint_upstream_error", "code": 502}})
        ctype = (r.getheader("content-type") or "")
        if stream and ctype.startswith("text/event-stream"):
            acc = G.StreamingTokenAccumulator(model)
            relayed, pending = b"", ""
            while True:
                chunk = r.read1(4096) if hasattr(r, "read1") else r.read(4096)
                if not chunk:
                    break
                relayed += chunk                                                            # byte-transparent (proxy.rs:224-241)
                pending = G.process_sse_lines(pending + chunk.decode("utf-8", "replace"), acc)
            c.close()
            ms = max(1, int((time.monotonic() - t0) * 1000))
            u = acc.finalize()
            out = u.get("output_tokens") or 0
            lease.complete("success", ms, {"input": u.get("input_tokens"), "output": u.get("output_tokens"), "total": u.get("total_tokens")})
            if out > 0:
                self.lm.update_tps(eid, base, kind, out, ms)
            return 200, {"content-type": "text/event-stream", "x-endpoint": eid}, relayed
        body = r.read()
        c.close()
        ms = int((time.monotonic() - t0) * 1000)
        try:
            j = json.loads(body)
        except ValueError:
            lease.complete("error", ms)
            return self._json(502, G.openai_error_body("invalid JSON from endpoint", "endpoint_upstream_error", 502))
        if path != "/v1/responses" and isinstance(j, dict):
            j["model"] = model                                                              # openai.rs:1228: the client's name
        u = G.extract_usage_from_response(j)
        lease.complete("success", ms, None if u is None else {"input": u.get("input_tokens"), "output": u.get("output_tokens"), "total": u.get("total_tokens")})
        out = (u or {}).get("output_tokens") or 0
        if out > 0:
            self.lm.update_tps(eid, base, kind, out, max(1, ms))
        st, hd, bd = self._json(200, j)
        hd["x-endpoint"] = eid
        return st, hd, bd

    @staticmethod
    def _json(status, obj):
        return status, {"content-type": "application/json"}, json.dumps(obj).encode()
