// See anthropic.hpp.  Error texts and field handling follow llmlb/src/api/anthropic.rs so a client
// of the reference's /v1/messages sees the same bodies from the in-process engine.
#include "anthropic.hpp"

#include <cstring>

namespace llmlb_host {

std::string AnthropicError::body() const {
  Json e = Json::object();
  e.set("type", type);
  e.set("message", message);
  Json j = Json::object();
  j.set("type", "error");
  j.set("error", e);
  return j.dump();
}

namespace {
bool bad(AnthropicError* err, const std::string& msg) {
  if (err) { err->status = 400; err->type = "invalid_request_error"; err->message = msg; }
  return false;
}
bool is_type(const Json& item, const char* t) {
  const Json* ty = item.get("type");
  return ty && ty->is_string() && ty->str() == t;
}
// anthropic.rs:1323-1364
bool flatten_text(const Json& v, const std::string& field, std::string* out, AnthropicError* err) {
  if (v.is_string()) { *out = v.str(); return true; }
  if (v.is_array()) {
    out->clear();
    for (const Json& item : v.items()) {
      const Json* ty = item.get("type");
      if (!ty || !ty->is_string()) return bad(err, field + " content blocks must have a type");
      if (ty->str() != "text") return bad(err, field + " content block type '" + ty->str() + "' is not supported");
      const Json* tx = item.get("text");
      if (!tx || !tx->is_string()) return bad(err, field + " text content blocks must include text");
      *out += tx->str();
    }
    return true;
  }
  return bad(err, field + " must be a string or text content array");
}
// anthropic.rs:1218-1259
bool tool_to_openai(const Json& tool, Json* out, AnthropicError* err) {
  const Json* name = tool.get("name");
  if (!name || !name->is_string()) return bad(err, "tool.name is required");
  const Json* desc = tool.get("description");
  const Json* schema = tool.get("input_schema");
  if (!schema) return bad(err, "tool.input_schema is required");
  Json params = Json::object();
  for (const char* k : {"type", "properties", "required"})
    if (const Json* v = schema->get(k)) params.set(k, *v);
  Json fn = Json::object();
  fn.set("name", name->str());
  fn.set("description", desc && desc->is_string() ? desc->str() : std::string());
  fn.set("parameters", params);
  *out = Json::object();
  out->set("type", "function");
  out->set("function", fn);
  return true;
}
// anthropic.rs:1261-1298
bool tool_choice_to_openai(const Json& tc, Json* out, AnthropicError* err) {
  const Json* ty = tc.get("type");
  if (!ty || !ty->is_string()) return bad(err, "tool_choice.type is required");
  if (ty->str() == "auto") { *out = Json("auto"); return true; }
  if (ty->str() == "any") { *out = Json("required"); return true; }
  if (ty->str() == "tool") {
    const Json* name = tc.get("name");
    if (!name || !name->is_string()) return bad(err, "tool_choice.name is required when type is 'tool'");
    Json fn = Json::object();
    fn.set("name", name->str());
    *out = Json::object();
    out->set("type", "function");
    out->set("function", fn);
    return true;
  }
  return bad(err, "unknown tool_choice type: " + ty->str());
}
// anthropic.rs:1415-1433
bool tool_call_to_tool_use(const Json& tc, Json* out) {
  const Json* fn = tc.get("function");
  if (!fn) return false;
  const Json* name = fn->get("name");
  const Json* id = tc.get("id");
  if (!name || !name->is_string() || !id || !id->is_string()) return false;
  const Json* args = fn->get("arguments");
  Json input;
  if (!Json::parse(args && args->is_string() ? args->str() : std::string("{}"), &input)) input = Json::object();
  *out = Json::object();
  out->set("type", "tool_use");
  out->set("id", id->str());
  out->set("name", name->str());
  out->set("input", input);
  return true;
}
bool blank(const std::string& s) {
  for (unsigned char c : s)
    if (!(c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v')) return false;
  return true;
}
}  // namespace

bool anthropic_required_header(const char* value, const char* name, AnthropicError* err) {
  if (value && !blank(value)) return true;
  return bad(err, std::string("Missing required header: ") + name);
}

const char* map_finish_reason_to_stop_reason(const std::string& fr) {
  if (fr == "length") return "max_tokens";
  if (fr == "tool_calls") return "tool_use";
  return "end_turn";
}

bool anthropic_request_to_openai(const Json& payload, Json* openai, std::string* request_text, bool* stream_out,
                                 AnthropicError* err) {
  const Json* model = payload.get("model");
  if (!model || !model->is_string()) return bad(err, "model is required");
  if (blank(model->str())) return bad(err, "model must not be empty");
  const Json* mt = payload.get("max_tokens");
  uint64_t max_tokens = 0;
  if (!mt || !mt->as_u64(&max_tokens)) return bad(err, "max_tokens is required");
  const Json* st = payload.get("stream");
  const bool stream = st && st->type() == Json::Bool && st->as_bool();
  const Json* msgs = payload.get("messages");
  if (!msgs || !msgs->is_array()) return bad(err, "messages must be an array");

  std::vector<std::string> parts;
  Json out_msgs = Json::array();
  auto push_msg = [&](const std::string& role, const std::string& content) {
    Json m = Json::object();
    m.set("role", role);
    m.set("content", content);
    out_msgs.push(m);
  };
  if (const Json* sys = payload.get("system")) {
    std::string text;
    if (!flatten_text(*sys, "system", &text, err)) return false;
    if (!text.empty()) { push_msg("system", text); parts.push_back("system: " + text); }
  }
  size_t index = 0;
  for (const Json& m : msgs->items()) {
    const std::string idx = "messages[" + std::to_string(index++) + "]";
    const Json* role = m.get("role");
    if (!role || !role->is_string()) return bad(err, idx + ".role is required");
    if (role->str() != "user" && role->str() != "assistant") return bad(err, idx + ".role must be 'user' or 'assistant'");
    const Json* content = m.get("content");
    if (!content) return bad(err, idx + ".content is required");
    if (content->is_array()) {
      bool tool_use = false, tool_result = false;
      for (const Json& it : content->items()) { tool_use |= is_type(it, "tool_use"); tool_result |= is_type(it, "tool_result"); }
      if (role->str() == "assistant" && tool_use) continue;  // carried by tool_calls of the response, not replayed
      if (role->str() == "user" && tool_result) {
        for (const Json& it : content->items()) {
          if (!is_type(it, "tool_result")) continue;
          const Json* tid = it.get("tool_use_id");
          const Json* rc = it.get("content");
          const std::string id = tid && tid->is_string() ? tid->str() : "unknown";
          const std::string res = rc && rc->is_string() ? rc->str() : "";
          Json tm = Json::object();
          tm.set("role", "tool");
          tm.set("tool_call_id", id);
          tm.set("content", res);
          out_msgs.push(tm);
          parts.push_back("tool_result[" + id + "]: " + res);
        }
        continue;
      }
    }
    std::string text;
    if (!flatten_text(*content, idx + ".content", &text, err)) return false;
    push_msg(role->str(), text);
    parts.push_back(role->str() + ": " + text);
  }
  Json body = Json::object();
  body.set("model", model->str());
  body.set("messages", out_msgs);
  body.set("max_tokens", Json(max_tokens));
  body.set("stream", Json(stream));
  if (const Json* t = payload.get("temperature")) if (t->is_number()) body.set("temperature", Json(t->as_double()));
  if (const Json* t = payload.get("top_p")) if (t->is_number()) body.set("top_p", Json(t->as_double()));
  if (const Json* ss = payload.get("stop_sequences")) {
    if (!ss->is_array()) return bad(err, "stop_sequences must be an array of strings");
    Json stop = Json::array();
    for (const Json& s : ss->items()) {
      if (!s.is_string()) return bad(err, "stop_sequences must be an array of strings");
      stop.push(Json(s.str()));
    }
    body.set("stop", stop);
  }
  if (const Json* tools = payload.get("tools")) if (tools->is_array()) {
    Json arr = Json::array();
    for (const Json& t : tools->items()) {
      Json conv;
      if (!tool_to_openai(t, &conv, err)) return false;
      arr.push(conv);
    }
    body.set("tools", arr);
  }
  if (const Json* tc = payload.get("tool_choice")) {
    Json conv;
    if (!tool_choice_to_openai(*tc, &conv, err)) return false;
    body.set("tool_choice", conv);
  }
  *openai = body;
  if (request_text) {
    request_text->clear();
    for (size_t i = 0; i < parts.size(); ++i) { if (i) *request_text += "\n"; *request_text += parts[i]; }
  }
  if (stream_out) *stream_out = stream;
  return true;
}

Json openai_to_anthropic_message_response(const Json& body, const std::string& model, int64_t input_tokens,
                                          int64_t output_tokens, const std::string& fallback_id) {
  const Json* choices = body.get("choices");
  const Json* choice = choices && choices->is_array() && !choices->items().empty() ? &choices->items()[0] : nullptr;
  const Json* fr = choice ? choice->get("finish_reason") : nullptr;
  const Json* msg = choice ? choice->get("message") : nullptr;
  std::string text;
  if (choice) {
    const Json* c = msg ? msg->get("content") : nullptr;
    const Json* t = choice->get("text");
    if (c && c->is_string()) text = c->str();
    else if (t && t->is_string()) text = t->str();
  }
  Json content = Json::array();
  if (!text.empty()) {
    Json blk = Json::object();
    blk.set("type", "text");
    blk.set("text", text);
    content.push(blk);
  }
  if (msg)
    if (const Json* tcs = msg->get("tool_calls"))
      if (tcs->is_array())
        for (const Json& tc : tcs->items()) {
          Json blk;
          if (tool_call_to_tool_use(tc, &blk)) content.push(blk);
        }
  if (content.items().empty()) {
    Json blk = Json::object();
    blk.set("type", "text");
    blk.set("text", "");
    content.push(blk);
  }
  const char* stop = fr && fr->is_string() ? map_finish_reason_to_stop_reason(fr->str()) : "end_turn";
  const Json* id = body.get("id");
  Json usage = Json::object();
  usage.set("input_tokens", Json(input_tokens < 0 ? int64_t(0) : input_tokens));
  usage.set("output_tokens", Json(output_tokens < 0 ? int64_t(0) : output_tokens));
  Json out = Json::object();
  out.set("id", id && id->is_string() ? id->str() : fallback_id);
  out.set("type", "message");
  out.set("role", "assistant");
  out.set("model", model);
  out.set("content", content);
  out.set("stop_reason", stop);
  out.set("stop_sequence", Json());
  out.set("usage", usage);
  return out;
}

// ---- streaming ----------------------------------------------------------------------------------
void AnthropicStreamTransformer::emit(const char* name, const Json& data) {
  out_ += "event: ";
  out_ += name;
  out_ += "\ndata: ";
  out_ += data.dump();
  out_ += "\n\n";
  names_.push_back(name);
}

std::string AnthropicStreamTransformer::take_output() {
  std::string o;
  o.swap(out_);
  return o;
}

void AnthropicStreamTransformer::feed(const std::string& text) {
  line_buf_ += text;
  size_t nl;
  while ((nl = line_buf_.find('\n')) != std::string::npos) {
    std::string line = line_buf_.substr(0, nl);
    line_buf_.erase(0, nl + 1);
    while (!line.empty() && line.back() == '\r') line.pop_back();
    process_line(line);
  }
}

void AnthropicStreamTransformer::ensure_message_start() {
  if (started_) return;
  started_ = true;
  Json usage = Json::object();
  usage.set("input_tokens", Json(input_tokens_ < 0 ? int64_t(0) : input_tokens_));
  usage.set("output_tokens", Json(int64_t(0)));
  Json m = Json::object();
  m.set("id", response_id_);
  m.set("type", "message");
  m.set("role", "assistant");
  m.set("content", Json::array());
  m.set("model", model_);
  m.set("stop_reason", Json());
  m.set("stop_sequence", Json());
  m.set("usage", usage);
  Json ev = Json::object();
  ev.set("type", "message_start");
  ev.set("message", m);
  emit("message_start", ev);
}

void AnthropicStreamTransformer::ensure_content_block_start() {
  if (block_started_) return;
  block_started_ = true;
  Json blk = Json::object();
  blk.set("type", "text");
  blk.set("text", "");
  Json ev = Json::object();
  ev.set("type", "content_block_start");
  ev.set("index", Json(int64_t(0)));
  ev.set("content_block", blk);
  emit("content_block_start", ev);
}

void AnthropicStreamTransformer::process_line(const std::string& line) {
  acc_.process_chunk(line);
  size_t a = 0, b = line.size();
  while (a < b && isspace((unsigned char)line[a])) ++a;
  while (b > a && isspace((unsigned char)line[b - 1])) --b;
  if (a == b || line[a] == ':' || line.compare(a, 5, "data:") != 0) return;
  a += 5;
  while (a < b && isspace((unsigned char)line[a])) ++a;
  const std::string data = line.substr(a, b - a);
  if (data == "[DONE]") { finish(); return; }
  Json js;
  if (!Json::parse(data, &js)) return;
  if (const Json* id = js.get("id"))
    if (id->is_string()) {
      std::string r = id->str();
      for (const char* from : {"chatcmpl-", "chatcmpl"}) {
        const std::string to = strcmp(from, "chatcmpl-") == 0 ? "msg_" : "msg";
        size_t p = 0;
        while ((p = r.find(from, p)) != std::string::npos) { r.replace(p, strlen(from), to); p += to.size(); }
      }
      response_id_ = r;
    }
  ensure_message_start();
  const Json* choices = js.get("choices");
  if (!choices || !choices->is_array() || choices->items().empty()) return;
  const Json& choice = choices->items()[0];
  const Json* delta = choice.get("delta");
  const Json* content = delta ? delta->get("content") : nullptr;
  if (content && content->is_string()) {
    ensure_content_block_start();
    if (!content->str().empty()) {
      Json d = Json::object();
      d.set("type", "text_delta");
      d.set("text", content->str());
      Json ev = Json::object();
      ev.set("type", "content_block_delta");
      ev.set("index", Json(int64_t(0)));
      ev.set("delta", d);
      emit("content_block_delta", ev);
    }
  }
  const Json* tcs = delta ? delta->get("tool_calls") : nullptr;
  if (tcs && tcs->is_array() && !tcs->items().empty()) {
    if (block_started_ && !block_stopped_) {
      block_stopped_ = true;
      Json ev = Json::object();
      ev.set("type", "content_block_stop");
      ev.set("index", Json(int64_t(0)));
      emit("content_block_stop", ev);
    }
    int64_t idx = 1;
    for (const Json& tc : tcs->items()) {
      Json blk;
      if (tool_call_to_tool_use(tc, &blk)) {
        Json s = Json::object();
        s.set("type", "content_block_start");
        s.set("index", Json(idx));
        s.set("content_block", blk);
        emit("content_block_start", s);
        Json e = Json::object();
        e.set("type", "content_block_stop");
        e.set("index", Json(idx));
        emit("content_block_stop", e);
      }
      ++idx;
    }
  }
  const Json* fr = choice.get("finish_reason");
  if (fr && fr->is_string()) stop_reason_ = map_finish_reason_to_stop_reason(fr->str());
}

void AnthropicStreamTransformer::finish() {
  if (stopped_) return;
  ensure_message_start();
  ensure_content_block_start();
  if (!block_stopped_) {
    block_stopped_ = true;
    Json ev = Json::object();
    ev.set("type", "content_block_stop");
    ev.set("index", Json(int64_t(0)));
    emit("content_block_stop", ev);
  }
  const TokenUsage u = acc_.finalize();
  Json d = Json::object();
  d.set("stop_reason", stop_reason_ ? stop_reason_ : "end_turn");
  d.set("stop_sequence", Json());
  Json usage = Json::object();
  usage.set("output_tokens", Json(int64_t(u.has_out ? u.out : 0)));
  Json ev = Json::object();
  ev.set("type", "message_delta");
  ev.set("delta", d);
  ev.set("usage", usage);
  emit("message_delta", ev);
  Json stop = Json::object();
  stop.set("type", "message_stop");
  emit("message_stop", stop);
  stopped_ = true;
}

}  // namespace llmlb_host

// =============================================================================================
// extern "C" surface for ctypes tests (tests/test_host_gateway.py)
// =============================================================================================
using namespace llmlb_host;

static int64_t a_copy(const std::string& s, char* out, uint64_t cap) {
  const uint64_t n = s.size() < cap ? s.size() : cap;
  if (out && n) memcpy(out, s.data(), n);
  return int64_t(s.size());
}

#include "../../include/llmlb_host.h"   // the exported signatures are checked against the public header at compile time
extern "C" {
// *status = 200 and out = {"openai":{...},"request_text":"...","stream":bool}, or the HTTP status and the error body
int64_t llmlb_anthropic_convert_request(const char* json, uint64_t len, int* status, char* out, uint64_t cap) {
  Json payload, openai;
  AnthropicError err;
  std::string text;
  bool stream = false;
  if (!Json::parse(std::string(json, len), &payload)) {
    err.status = 400; err.type = "invalid_request_error"; err.message = "invalid JSON body";
    *status = 400;
    return a_copy(err.body(), out, cap);
  }
  if (!anthropic_request_to_openai(payload, &openai, &text, &stream, &err)) { *status = err.status; return a_copy(err.body(), out, cap); }
  Json r = Json::object();
  r.set("openai", openai);
  r.set("request_text", text);
  r.set("stream", Json(stream));
  *status = 200;
  return a_copy(r.dump(), out, cap);
}
int64_t llmlb_anthropic_convert_response(const char* body_json, uint64_t len, const char* model, int64_t input_tokens,
                                         int64_t output_tokens, const char* fallback_id, char* out, uint64_t cap) {
  Json body;
  if (!Json::parse(std::string(body_json, len), &body)) return -1;
  return a_copy(openai_to_anthropic_message_response(body, model, input_tokens, output_tokens, fallback_id).dump(), out, cap);
}
int64_t llmlb_anthropic_header_check(const char* value, const char* name, int* status, char* out, uint64_t cap) {
  AnthropicError err;
  if (anthropic_required_header(value, name, &err)) { *status = 200; return 0; }
  *status = err.status;
  return a_copy(err.body(), out, cap);
}
void* llmlb_anthropic_stream_create(const char* model, int64_t input_tokens, const char* response_id) {
  return new AnthropicStreamTransformer(model, input_tokens, response_id);
}
void llmlb_anthropic_stream_destroy(void* t) { delete static_cast<AnthropicStreamTransformer*>(t); }
int64_t llmlb_anthropic_stream_feed(void* t, const char* text, uint64_t len, char* out, uint64_t cap) {
  auto* tr = static_cast<AnthropicStreamTransformer*>(t);
  tr->feed(std::string(text, len));
  return a_copy(tr->take_output(), out, cap);
}
int64_t llmlb_anthropic_stream_finish(void* t, char* out, uint64_t cap) {
  auto* tr = static_cast<AnthropicStreamTransformer*>(t);
  tr->finish();
  return a_copy(tr->take_output(), out, cap);
}
}  // extern "C"
