// Mutation fuzzer for the host-side parsers that read bytes nobody here controls: checkpoint files (safetensors,
// GGUF), tokenizer.json, request JSON, SSE relayed from an upstream.  Built with -fsanitize=address,undefined by
// tools/fuzz/run.py; a finding is a sanitizer report or a crash, "clean" is N iterations without one.
//
//   fuzz_host ckpt  <seed-file> <iterations> <rng-seed> <scratch-path>
//   fuzz_host tok   <tokenizer.json> <iterations> <rng-seed>
//   fuzz_host tokfile <tokenizer.json> 1 0                     (no mutation: load + exercise this file)
//   fuzz_host json  <seed-file> <iterations> <rng-seed>
//   fuzz_host sse   <seed-file> <iterations> <rng-seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../llmlb_b200/host/checkpoint.hpp"
#include "../../llmlb_b200/host/gateway.hpp"
#include "../../llmlb_b200/host/json.hpp"
#include "../../llmlb_b200/host/tokenizer.hpp"

using namespace llmlb_host;

static uint64_t g_s = 88172645463325252ull;
static uint64_t rnd() { g_s ^= g_s << 13; g_s ^= g_s >> 7; g_s ^= g_s << 17; return g_s; }
static size_t below(size_t n) { return n ? size_t(rnd() % n) : 0; }

static std::string slurp(const char* p) {
  std::ifstream f(p, std::ios::binary);
  std::stringstream ss; ss << f.rdbuf();
  return ss.str();
}

// Byte-level mutations biased towards what breaks length-prefixed binary formats: interesting integers written
// over 4/8-byte windows, truncation, block duplication, plus the usual flips.
static std::string mutate(const std::string& seed, size_t header_bias) {
  std::string s = seed;
  const int n_mut = 1 + int(below(4));
  static const uint64_t kInteresting[] = {0, 1, 0x7F, 0x80, 0xFF, 0x7FFF, 0x8000, 0xFFFF, 0x7FFFFFFFull, 0x80000000ull, 0xFFFFFFFFull,
                                          0x100000000ull, 0x7FFFFFFFFFFFFFFFull, 0x8000000000000000ull, 0xFFFFFFFFFFFFFFFFull,
                                          0xFFFFFFFFFFFFFFF0ull, 1ull << 40, 1ull << 56};
  for (int m = 0; m < n_mut && !s.empty(); ++m) {
    // most mutations land in the header (where the structure is), some anywhere
    const size_t span = (header_bias && below(4)) ? std::min(header_bias, s.size()) : s.size();
    const size_t at = below(span);
    switch (below(8)) {
      case 0: s[at] = char(s[at] ^ (1u << below(8))); break;
      case 1: s[at] = char(rnd()); break;
      case 2: { const uint64_t v = kInteresting[below(sizeof kInteresting / 8)]; const size_t w = below(2) ? 8 : 4;
                if (at + w <= s.size()) memcpy(&s[at], &v, w); break; }
      case 3: s.resize(at); break;                                                       // truncate
      case 4: { const size_t len = 1 + below(64); if (at + len <= s.size()) s.insert(at, s.substr(at, len)); break; }   // duplicate
      case 5: { const size_t len = 1 + below(64); if (at + len <= s.size()) s.erase(at, len); break; }                  // delete
      case 6: { uint64_t v; if (at + 8 <= s.size()) { memcpy(&v, &s[at], 8); v += (rnd() % 65) - 32; memcpy(&s[at], &v, 8); } break; }
      case 7: { const size_t len = 1 + below(16); for (size_t i = 0; i < len && at + i < s.size(); ++i) s[at + i] = char(rnd()); break; }
    }
  }
  return s;
}

static int fuzz_ckpt(const std::string& seed, long iters, const char* scratch) {
  long opened = 0, read_ok = 0;
  for (long it = 0; it < iters; ++it) {
    const std::string m = mutate(seed, 2048);
    { std::ofstream f(scratch, std::ios::binary | std::ios::trunc); f.write(m.data(), std::streamsize(m.size())); }
    Checkpoint c;
    std::string err;
    if (!c.open(scratch, &err)) continue;
    ++opened;
    (void)c.geometry();
    (void)c.tokenizer_json();
    const auto& ts = c.tensors();
    for (size_t i = 0; i < ts.size() && i < 64; ++i) {
      if (ts[i].rows * ts[i].cols > (1u << 24)) continue;      // a mutated shape can ask for gigabytes; the loader's caller bounds that by the model geometry
      std::vector<uint16_t> out;
      if (c.read_bf16(i, &out, &err)) ++read_ok;
    }
  }
  printf("ckpt: %ld iterations, %ld opened, %ld tensors read\n", iters, opened, read_ok);
  return 0;
}

static int fuzz_tok(const std::string& seed, long iters, bool as_is) {
  long loaded = 0;
  for (long it = 0; it < iters; ++it) {
    const std::string m = as_is ? seed : mutate(seed, 0);
    BpeTokenizer t;
    std::string err;
    if (!t.load_json(m, &err)) continue;
    ++loaded;
    std::vector<int32_t> ids = t.encode("Hello, w\xC3\xB6rld! 12345 \xF0\x9F\x98\x80 <|eot_id|>\n\n  x", true, true);
    std::string out = t.decode(ids, false);
    BpeTokenizer::Stream st;
    for (int32_t id : ids) (void)t.decode_next(&st, id, true);
    for (int k = 0; k < 8; ++k) (void)t.decode_next(&st, int32_t(rnd() % 70000) - 100, false);   // ids outside the vocabulary too
    (void)t.encode_chat({{"system", "s"}, {"user", out}});
  }
  printf("tok: %ld iterations, %ld loaded\n", iters, loaded);
  return 0;
}

static int fuzz_json(const std::string& seed, long iters) {
  long ok = 0;
  for (long it = 0; it < iters; ++it) {
    const std::string m = mutate(seed, 0);
    Json j;
    if (!Json::parse(m, &j)) continue;
    ++ok;
    const std::string d = j.dump();
    Json k;
    if (!Json::parse(d, &k) || k.dump() != d) { fprintf(stderr, "json round trip differs\n%s\n", d.c_str()); abort(); }
    TokenUsage u; (void)extract_usage_from_response(j, &u);
  }
  printf("json: %ld iterations, %ld parsed\n", iters, ok);
  return 0;
}

static int fuzz_sse(const std::string& seed, long iters) {
  for (long it = 0; it < iters; ++it) {
    const std::string m = mutate(seed, 0);
    StreamingTokenAccumulator a("m");
    size_t at = 0;
    while (at < m.size()) {                       // arbitrary chunk boundaries, like a socket
      const size_t n = 1 + below(97);
      a.feed(m.data() + at, std::min(n, m.size() - at));
      at += n;
    }
    (void)a.finalize();
  }
  printf("sse: %ld iterations\n", iters);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: fuzz_host ckpt|tok|json|sse <seed-file> <iterations> <rng-seed> [scratch]\n"); return 2; }
  const std::string mode = argv[1], seed = slurp(argv[2]);
  const long iters = atol(argv[3]);
  g_s ^= strtoull(argv[4], nullptr, 10) * 0x9E3779B97F4A7C15ull;
  if (seed.empty()) { fprintf(stderr, "empty seed %s\n", argv[2]); return 2; }
  if (mode == "ckpt") return fuzz_ckpt(seed, iters, argc > 5 ? argv[5] : "/tmp/fuzz_ckpt.bin");
  if (mode == "tok") return fuzz_tok(seed, iters, false);
  if (mode == "tokfile") return fuzz_tok(seed, 1, true);        // a structurally mutated tokenizer.json written by run.py, loaded as it is
  if (mode == "json") return fuzz_json(seed, iters);
  if (mode == "sse") return fuzz_sse(seed, iters);
  return 2;
}
