/* llmlb_host.h — C ABI of libllmlb_host.so: the text side of the boundary (no GPU, no CUDA).
 *
 * The engine (llmlb_b200.h) takes and returns token ids; a gateway process needs text <-> ids and
 * the wire formats of its front doors.  The reference does these steps in Rust inside the
 * gateway (SSE accounting llmlb/src/token/mod.rs:41-223, Anthropic translation
 * llmlb/src/api/anthropic.rs:728-1216, 1435-1504) and leaves tokenisation to the engines it
 * proxies to; a maintainer binding the in-process engine can either keep its own crates or bind
 * these functions (same cgo / Rust-FFI / ctypes pattern as INTEGRATION.md §2).
 *
 * Conventions: strings are (pointer, length) or NUL-terminated as declared; every "out, cap" pair
 * returns the full length of the result (call again with a larger buffer if it exceeds cap);
 * handles are opaque and not thread-safe per object, the library itself is.                      */
#ifndef LLMLB_HOST_H
#define LLMLB_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Llama-3 byte-level BPE tokenizer (host/tokenizer.hpp) ---------------------------------- */
/* tokenizer.json text -> handle; NULL and a message in err on failure */
void* llmlb_tok_create(const char* tokenizer_json, uint64_t len, char* err, uint32_t err_cap);
void llmlb_tok_destroy(void* tok);
uint32_t llmlb_tok_vocab_size(void* tok);
int32_t llmlb_tok_bos_id(void* tok);
int32_t llmlb_tok_special_id(void* tok, const char* content);              /* -1 if absent */
/* parse_special: control tokens written in the text become their ids; otherwise they are text */
int64_t llmlb_tok_encode(void* tok, const char* text, uint64_t len, int add_bos, int parse_special,
                         int32_t* out_ids, uint64_t cap);
int64_t llmlb_tok_decode(void* tok, const int32_t* ids, uint64_t n, int skip_special, char* out, uint64_t cap);
/* byte ranges [begin,end) of the pre-tokenizer pieces, two uint32 per piece */
int64_t llmlb_tok_pretokenize(const char* text, uint64_t len, uint32_t* out_pairs, uint64_t cap_pairs);
/* streaming detokenizer: only complete UTF-8 is returned; flush replaces a truncated tail by U+FFFD */
void* llmlb_tok_stream_create(void);
void llmlb_tok_stream_destroy(void* stream);
int64_t llmlb_tok_stream_next(void* tok, void* stream, int32_t id, int skip_special, char* out, uint64_t cap);
int64_t llmlb_tok_stream_flush(void* stream, char* out, uint64_t cap);
/* messages_json: [{"role","content"},...] -> Llama-3 chat template (ids: markers as control tokens,
 * role/content as plain text; text: the rendered template).  -1 on malformed input. */
int64_t llmlb_tok_chat_ids(void* tok, const char* messages_json, uint64_t len, int32_t* out_ids, uint64_t cap);
int64_t llmlb_tok_chat_text(void* tok, const char* messages_json, uint64_t len, int add_generation_prompt,
                            char* out, uint64_t cap);

/* ---- Anthropic Messages <-> OpenAI chat (host/anthropic.hpp) --------------------------------- */
/* *status = 200 and {"openai":{...},"request_text":"...","stream":bool}, or the HTTP status and the
 * Anthropic error body {"type":"error","error":{"type","message"}} */
int64_t llmlb_anthropic_convert_request(const char* payload_json, uint64_t len, int* status, char* out, uint64_t cap);
/* OpenAI chat.completion body -> Anthropic message; usage < 0 = absent */
int64_t llmlb_anthropic_convert_response(const char* body_json, uint64_t len, const char* model, int64_t input_tokens,
                                         int64_t output_tokens, const char* fallback_id, char* out, uint64_t cap);
int64_t llmlb_anthropic_header_check(const char* value, const char* name, int* status, char* out, uint64_t cap);
/* OpenAI chat SSE bytes in (any chunking), Anthropic SSE events out */
void* llmlb_anthropic_stream_create(const char* model, int64_t input_tokens /* <0: unknown */, const char* response_id);
void llmlb_anthropic_stream_destroy(void* t);
int64_t llmlb_anthropic_stream_feed(void* t, const char* text, uint64_t len, char* out, uint64_t cap);
int64_t llmlb_anthropic_stream_finish(void* t, char* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* LLMLB_HOST_H */
