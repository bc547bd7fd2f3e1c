// TEST INFRASTRUCTURE — NOT PRODUCT.  A scripted token source behind the C ABI of include/llmlb_b200.h, so that the HTTP
// shim (llmlb_b200/host/server.cpp: routing, JSON / SSE framing, tokenizer, Anthropic translation, error conventions, drain
// gate, download routes) can be exercised on a machine WITHOUT a GPU (`pytest -m "not gpu"`).  It does NO model arithmetic:
// token i of a request is a hash of (prompt, sampling seed, i).  It is built only by tests/test_server_fake_engine_cpu.py
// into tests/support/_build/ next to a server binary linked against it; libllmlb_b200.so, the product, never links or loads
// it, and the product server still refuses to start without a CUDA device.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/llmlb_b200.h"

namespace {
thread_local std::string g_err;
int fail(int rc, const std::string& m) { g_err = m; return rc; }
using Clock = std::chrono::steady_clock;

struct Req {
  std::vector<int32_t> prompt, stop_ids;
  llmlb_sampling s{};
  std::deque<llmlb_token_event> events;
  uint32_t produced = 0;
  bool finished = false, cancel = false;
  uint64_t h = 0;
  bool admitted = false;
  Clock::time_point t_submit;
};
}  // namespace

struct llmlb_engine {
  llmlb_engine_config cfg{};
  std::mutex mu;
  std::condition_variable cv;
  std::map<uint64_t, std::shared_ptr<Req>> reqs;
  uint64_t next_id = 1;
  std::atomic<bool> stop{false};
  std::thread worker;
  uint64_t tokens = 0;
  unsigned token_us = 150;          // pace of the scripted source; FAKE_ENGINE_TOKEN_US slows it down for the timeout tests

  static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

  void finish(Req& r, uint32_t reason) {   // mu held
    if (r.finished) return;
    r.finished = true;
    if (!r.events.empty() && r.events.back().finish_reason == LLMLB_FINISH_NONE && (reason == LLMLB_FINISH_STOP || reason == LLMLB_FINISH_LENGTH)) {
      r.events.back().finish_reason = reason;
    } else {
      llmlb_token_event ev{};
      ev.token_id = -1; ev.index = r.produced; ev.finish_reason = reason;
      ev.prompt_tokens = uint32_t(r.prompt.size()); ev.completion_tokens = r.produced;
      ev.t_ms = std::chrono::duration<double, std::milli>(Clock::now() - r.t_submit).count();
      r.events.push_back(ev);
    }
  }

  void loop() {
    while (!stop.load()) {
      {
        std::lock_guard<std::mutex> lk(mu);
        uint32_t running = 0;
        for (auto& kv : reqs) {
          Req& r = *kv.second;
          if (r.finished) continue;
          const double age_ms = std::chrono::duration<double, std::milli>(Clock::now() - r.t_submit).count();
          if (r.cancel) { finish(r, LLMLB_FINISH_CANCELLED); continue; }
          if (cfg.request_timeout_ms && age_ms > cfg.request_timeout_ms) { finish(r, LLMLB_FINISH_DEADLINE); continue; }
          if (running++ >= cfg.max_seqs) {                      // the rest wait, like a full batch
            if (!r.admitted && cfg.queue_timeout_ms && age_ms > cfg.queue_timeout_ms) finish(r, LLMLB_FINISH_QUEUE_TIMEOUT);
            continue;
          }
          r.admitted = true;
          // scripted failure: a prompt that starts with 666 666 dies after three tokens (the engine's FINISH_ERROR path)
          if (r.prompt.size() >= 2 && r.prompt[0] == 666 && r.prompt[1] == 666 && r.produced >= 3) { finish(r, LLMLB_FINISH_ERROR); continue; }
          const uint64_t salt = r.s.temperature > 0 ? mix(r.s.seed + 0x9e37) : 0;
          const int32_t tok = int32_t(mix(r.h ^ salt ^ (uint64_t(r.produced) * 0x9e3779b97f4a7c15ull)) % cfg.model.vocab);
          llmlb_token_event ev{};
          ev.token_id = tok; ev.index = r.produced++;
          ev.prompt_tokens = uint32_t(r.prompt.size()); ev.completion_tokens = r.produced;
          ev.t_ms = std::chrono::duration<double, std::milli>(Clock::now() - r.t_submit).count();
          r.events.push_back(ev);
          ++tokens;
          bool stop_hit = false;
          if (!r.s.ignore_eos) for (int32_t sid : r.stop_ids) stop_hit |= sid == tok;
          if (stop_hit) finish(r, LLMLB_FINISH_STOP);
          else if (r.produced >= r.s.max_tokens || r.prompt.size() + r.produced >= cfg.max_ctx) finish(r, LLMLB_FINISH_LENGTH);
        }
      }
      cv.notify_all();
      std::this_thread::sleep_for(std::chrono::microseconds(token_us));
    }
  }
};

extern "C" {
uint32_t llmlb_abi_version(void) { return LLMLB_ABI_VERSION; }
const char* llmlb_last_error(void) { return g_err.c_str(); }

int llmlb_engine_create(const llmlb_engine_config* cfg, llmlb_engine** out) {
  if (!cfg || !out) return fail(LLMLB_E_INVALID_ARG, "null argument");
  if (cfg->abi_version != LLMLB_ABI_VERSION) return fail(LLMLB_E_INVALID_ARG, "abi_version mismatch");
  if (!cfg->model.vocab || !cfg->max_seqs || !cfg->max_ctx) return fail(LLMLB_E_INVALID_ARG, "bad geometry");
  auto* e = new llmlb_engine;
  e->cfg = *cfg;
  if (const char* v = getenv("FAKE_ENGINE_TOKEN_US")) e->token_us = unsigned(atoi(v));
  e->worker = std::thread([e] { e->loop(); });
  *out = e;
  return LLMLB_OK;
}
void llmlb_engine_destroy(llmlb_engine* e) {
  if (!e) return;
  e->stop.store(true);
  if (e->worker.joinable()) e->worker.join();
  delete e;
}
int llmlb_engine_model_info(const llmlb_engine* e, llmlb_model_info* out) {
  if (!e || !out) return fail(LLMLB_E_INVALID_ARG, "null argument");
  memset(out, 0, sizeof *out);
  memcpy(out->id, e->cfg.model_id, sizeof out->id);
  out->context_length = e->cfg.max_ctx; out->vocab = e->cfg.model.vocab; out->n_layers = e->cfg.model.n_layers; out->hidden = e->cfg.model.hidden;
  return LLMLB_OK;
}
int llmlb_engine_health(const llmlb_engine* ce, llmlb_health* out) {
  if (!ce || !out) return fail(LLMLB_E_INVALID_ARG, "null argument");
  llmlb_engine* e = const_cast<llmlb_engine*>(ce);
  memset(out, 0, sizeof *out);
  std::lock_guard<std::mutex> lk(e->mu);
  out->device_count = 1; out->total_memory_bytes = 1ull << 30; out->used_memory_bytes = 1ull << 20;
  for (auto& kv : e->reqs) if (!kv.second->finished) ++out->active_requests;
  out->total_kv_pages = e->cfg.max_seqs * ((e->cfg.max_ctx + 63) / 64); out->free_kv_pages = out->total_kv_pages;
  out->tokens_decode = e->tokens;
  return LLMLB_OK;
}
int llmlb_engine_load_tensor(llmlb_engine* e, const char* name, const void* host_bf16, uint64_t rows, uint64_t cols) {
  if (!e || !name || !host_bf16 || !rows || !cols) return fail(LLMLB_E_INVALID_ARG, "bad argument");
  return LLMLB_OK;   // the scripted source has no weights
}
int llmlb_request_submit(llmlb_engine* e, const int32_t* ids, uint32_t n, const llmlb_sampling* s, uint64_t* req_id) {
  if (!e || !ids || !n || !s || !req_id) return fail(LLMLB_E_INVALID_ARG, "bad argument");
  if (s->max_tokens == 0) return fail(LLMLB_E_INVALID_ARG, "max_tokens must be > 0");
  if (n >= e->cfg.max_ctx) return fail(LLMLB_E_INVALID_ARG, "prompt does not fit max_ctx");
  for (uint32_t i = 0; i < n; ++i)
    if (ids[i] < 0 || uint32_t(ids[i]) >= e->cfg.model.vocab) return fail(LLMLB_E_INVALID_ARG, "token id out of range");
  auto r = std::make_shared<Req>();
  r->prompt.assign(ids, ids + n);
  r->s = *s;
  if (s->stop_ids && s->n_stop_ids) r->stop_ids.assign(s->stop_ids, s->stop_ids + s->n_stop_ids);
  r->s.stop_ids = nullptr;
  uint64_t h = 0x243f6a8885a308d3ull;
  for (uint32_t i = 0; i < n; ++i) h = llmlb_engine::mix(h ^ uint64_t(uint32_t(ids[i])) ^ (uint64_t(i) << 32));
  r->h = h;
  r->t_submit = Clock::now();
  std::lock_guard<std::mutex> lk(e->mu);
  uint32_t waiting = 0;
  for (auto& kv : e->reqs) if (!kv.second->finished) ++waiting;
  if (e->cfg.queue_max && waiting >= e->cfg.max_seqs + e->cfg.queue_max) return fail(LLMLB_E_QUEUE_FULL, "request queue is full");
  *req_id = e->next_id++;
  e->reqs[*req_id] = r;
  return LLMLB_OK;
}
int llmlb_request_poll(llmlb_engine* e, uint64_t req_id, llmlb_token_event* out, uint32_t cap, uint32_t* n_out, int timeout_ms) {
  if (!e || !out || !n_out) return fail(LLMLB_E_INVALID_ARG, "bad argument");
  std::unique_lock<std::mutex> lk(e->mu);
  auto it = e->reqs.find(req_id);
  if (it == e->reqs.end()) return fail(LLMLB_E_INVALID_ARG, "unknown request");
  auto r = it->second;
  auto ready = [&] { return !r->events.empty(); };
  if (timeout_ms < 0) e->cv.wait(lk, ready);
  else if (timeout_ms > 0) e->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready);
  uint32_t n = 0;
  while (n < cap && !r->events.empty()) { out[n++] = r->events.front(); r->events.pop_front(); }
  *n_out = n;
  return LLMLB_OK;
}
int llmlb_request_cancel(llmlb_engine* e, uint64_t req_id) {
  if (!e) return fail(LLMLB_E_INVALID_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  auto it = e->reqs.find(req_id);
  if (it == e->reqs.end()) return fail(LLMLB_E_INVALID_ARG, "unknown request");
  it->second->cancel = true;
  return LLMLB_OK;
}
int llmlb_request_release(llmlb_engine* e, uint64_t req_id) {
  if (!e) return fail(LLMLB_E_INVALID_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  auto it = e->reqs.find(req_id);
  if (it == e->reqs.end()) return fail(LLMLB_E_INVALID_ARG, "unknown request");
  it->second->cancel = true;
  e->reqs.erase(it);
  return LLMLB_OK;
}
}
