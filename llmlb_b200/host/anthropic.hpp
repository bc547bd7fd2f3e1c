// Anthropic Messages front door onto the same engine: the JSON translation the
// reference does in llmlb/src/api/anthropic.rs —
//   anthropic_request_to_openai                :1048-1216  (+ flatten_anthropic_text_content :1323,
//                                                 tool / tool_choice / stop_sequences helpers :1218-1321)
//   openai_to_anthropic_message_response       :1435-1504
//   AnthropicStreamTracker (SSE transformation) :813-1018
//   anthropic_error_response                   :1559-1575, extract_required_header :1388-1398
// restated in oracle/gateway_ref.py and compared with it by tests/test_host_gateway.py.
#pragma once
#include <string>
#include <vector>

#include "gateway.hpp"
#include "json.hpp"

namespace llmlb_host {

struct AnthropicError {
  int status = 0;
  std::string type, message;
  std::string body() const;  // {"type":"error","error":{"type","message"}}
};

// false + *err when the payload is rejected (always 400 invalid_request_error)
bool anthropic_request_to_openai(const Json& payload, Json* openai_payload, std::string* request_text,
                                 bool* stream, AnthropicError* err);
bool anthropic_required_header(const char* value, const char* name, AnthropicError* err);
const char* map_finish_reason_to_stop_reason(const std::string& finish_reason);
// usage < 0 = absent (reported as 0, like Option::unwrap_or(0))
Json openai_to_anthropic_message_response(const Json& body, const std::string& model, int64_t input_tokens,
                                          int64_t output_tokens, const std::string& fallback_id);

// OpenAI chat-completions SSE text in, Anthropic event stream out ("event: X\ndata: {...}\n\n")
class AnthropicStreamTransformer {
 public:
  AnthropicStreamTransformer(const std::string& model, int64_t input_tokens, const std::string& response_id)
      : acc_(model), model_(model), response_id_(response_id) {
    if (input_tokens >= 0) { acc_.set_input_tokens(uint32_t(input_tokens)); input_tokens_ = input_tokens; }
  }
  void feed(const std::string& text);   // any chunking of the upstream bytes
  void finish();                        // idempotent; also run on "data: [DONE]"
  std::string take_output();            // wire bytes produced since the last call
  const std::vector<std::string>& event_names() const { return names_; }
  TokenUsage usage() const { return acc_.finalize(); }

 private:
  void process_line(const std::string& line);
  void ensure_message_start();
  void ensure_content_block_start();
  void emit(const char* name, const Json& data);
  StreamingTokenAccumulator acc_;
  std::string model_, response_id_, line_buf_, out_;
  std::vector<std::string> names_;
  int64_t input_tokens_ = -1;
  bool started_ = false, block_started_ = false, block_stopped_ = false, stopped_ = false;
  const char* stop_reason_ = nullptr;
};

}  // namespace llmlb_host
