// Minimal JSON value / parser / writer for the gateway-side host code (no third-party deps).
// Objects keep insertion order (the wire format the gateway relays is order-insensitive, but
// stable output makes golden tests exact).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace llmlb_host {

class Json {
 public:
  enum Type { Null, Bool, Int, Double, String, Array, Object };
  Json() : t_(Null) {}
  Json(std::nullptr_t) : t_(Null) {}
  Json(bool b) : t_(Bool), b_(b) {}
  Json(int v) : t_(Int), i_(v) {}
  Json(unsigned v) : t_(Int), i_(v) {}
  Json(int64_t v) : t_(Int), i_(v) {}
  Json(uint64_t v) : t_(Int), i_(int64_t(v)) {}
  Json(double v) : t_(Double), d_(v) {}
  Json(const char* s) : t_(String), s_(s) {}
  Json(const std::string& s) : t_(String), s_(s) {}
  static Json array() { Json j; j.t_ = Array; return j; }
  static Json object() { Json j; j.t_ = Object; return j; }

  Type type() const { return t_; }
  bool is_null() const { return t_ == Null; }
  bool is_string() const { return t_ == String; }
  bool is_array() const { return t_ == Array; }
  bool is_object() const { return t_ == Object; }
  bool is_number() const { return t_ == Int || t_ == Double; }
  // serde_json's as_u64: only non-negative integers
  bool as_u64(uint64_t* out) const {
    if (t_ == Int && i_ >= 0) { *out = uint64_t(i_); return true; }
    return false;
  }
  bool as_bool(bool dflt = false) const { return t_ == Bool ? b_ : dflt; }
  double as_double(double dflt = 0) const { return t_ == Int ? double(i_) : t_ == Double ? d_ : dflt; }
  int64_t as_int(int64_t dflt = 0) const {   // a double outside int64 (1e300, NaN) saturates instead of being cast (which is undefined)
    if (t_ == Int) return i_;
    if (t_ != Double) return dflt;
    if (!(d_ == d_)) return dflt;
    return d_ >= 9.2e18 ? INT64_MAX : d_ <= -9.2e18 ? INT64_MIN : int64_t(d_);
  }
  const std::string& str() const { return s_; }

  const Json* get(const std::string& key) const {
    if (t_ != Object) return nullptr;
    for (auto& kv : o_) if (kv.first == key) return &kv.second;
    return nullptr;
  }
  Json& set(const std::string& key, Json v) {
    t_ = Object;
    for (auto& kv : o_) if (kv.first == key) { kv.second = std::move(v); return kv.second; }
    o_.emplace_back(key, std::move(v));
    return o_.back().second;
  }
  // like set() without the duplicate-key scan: for building big objects (a 128k-entry vocabulary)
  Json& append(const std::string& key, Json v) { t_ = Object; o_.emplace_back(key, std::move(v)); return o_.back().second; }
  Json& push(Json v) { t_ = Array; a_.push_back(std::move(v)); return a_.back(); }
  const std::vector<Json>& items() const { return a_; }
  const std::vector<std::pair<std::string, Json>>& members() const { return o_; }

  std::string dump() const { std::string out; write(out); return out; }

  static bool parse(const std::string& text, Json* out) {
    Parser p{text.data(), text.data() + text.size()};
    p.ws();
    if (!p.value(out, 0)) return false;
    p.ws();
    return p.p == p.end;
  }

  static void escape(const std::string& s, std::string& out) {
    out.push_back('"');
    for (unsigned char c : s) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        case '\b': out += "\\b"; break;
        case '\f': out += "\\f"; break;
        default:
          if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
          else out.push_back(char(c));
      }
    }
    out.push_back('"');
  }

 private:
  void write(std::string& out) const {
    switch (t_) {
      case Null: out += "null"; break;
      case Bool: out += b_ ? "true" : "false"; break;
      case Int: out += std::to_string(i_); break;
      case Double: {
        if (!std::isfinite(d_)) { out += "null"; break; }
        char b[40];
        for (int prec = 1; prec <= 17; ++prec) {  // shortest text that reads back exactly (0.2, not 0.20000000000000001)
          snprintf(b, sizeof b, "%.*g", prec, d_);
          if (strtod(b, nullptr) == d_) break;
        }
        std::string s(b);
        if (s.find_first_of(".eE") == std::string::npos) s += ".0";
        out += s; break;
      }
      case String: escape(s_, out); break;
      case Array:
        out.push_back('[');
        for (size_t i = 0; i < a_.size(); ++i) { if (i) out.push_back(','); a_[i].write(out); }
        out.push_back(']'); break;
      case Object:
        out.push_back('{');
        for (size_t i = 0; i < o_.size(); ++i) {
          if (i) out.push_back(',');
          escape(o_[i].first, out); out.push_back(':'); o_[i].second.write(out);
        }
        out.push_back('}'); break;
    }
  }

  struct Parser {
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
    bool lit(const char* s) { size_t n = strlen_(s); if (size_t(end - p) < n) return false; for (size_t i = 0; i < n; ++i) if (p[i] != s[i]) return false; p += n; return true; }
    static size_t strlen_(const char* s) { size_t n = 0; while (s[n]) ++n; return n; }
    static void utf8(uint32_t cp, std::string& out) {
      if (cp < 0x80) out.push_back(char(cp));
      else if (cp < 0x800) { out.push_back(char(0xC0 | (cp >> 6))); out.push_back(char(0x80 | (cp & 0x3F))); }
      else if (cp < 0x10000) { out.push_back(char(0xE0 | (cp >> 12))); out.push_back(char(0x80 | ((cp >> 6) & 0x3F))); out.push_back(char(0x80 | (cp & 0x3F))); }
      else { out.push_back(char(0xF0 | (cp >> 18))); out.push_back(char(0x80 | ((cp >> 12) & 0x3F))); out.push_back(char(0x80 | ((cp >> 6) & 0x3F))); out.push_back(char(0x80 | (cp & 0x3F))); }
    }
    bool hex4(uint32_t* v) {
      if (end - p < 4) return false;
      uint32_t r = 0;
      for (int i = 0; i < 4; ++i) {
        char c = p[i]; r <<= 4;
        if (c >= '0' && c <= '9') r |= c - '0'; else if (c >= 'a' && c <= 'f') r |= c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') r |= c - 'A' + 10; else return false;
      }
      p += 4; *v = r; return true;
    }
    bool string(std::string* out) {
      if (p >= end || *p != '"') return false;
      ++p;
      while (p < end && *p != '"') {
        unsigned char c = *p++;
        if (c == '\\') {
          if (p >= end) return false;
          char e = *p++;
          switch (e) {
            case '"': out->push_back('"'); break; case '\\': out->push_back('\\'); break;
            case '/': out->push_back('/'); break; case 'b': out->push_back('\b'); break;
            case 'f': out->push_back('\f'); break; case 'n': out->push_back('\n'); break;
            case 'r': out->push_back('\r'); break; case 't': out->push_back('\t'); break;
            case 'u': {
              uint32_t cp; if (!hex4(&cp)) return false;
              if (cp >= 0xD800 && cp < 0xDC00) {
                uint32_t lo;
                if (end - p >= 6 && p[0] == '\\' && p[1] == 'u') { p += 2; if (!hex4(&lo)) return false; if (lo < 0xDC00 || lo > 0xDFFF) return false; cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); }
                else return false;
              }
              utf8(cp, *out); break;
            }
            default: return false;
          }
        } else if (c < 0x20) return false;
        else out->push_back(char(c));
      }
      if (p >= end) return false;
      ++p; return true;
    }
    bool number(Json* out) {
      const char* s = p; bool is_int = true;
      if (p < end && *p == '-') ++p;
      if (p >= end || !(*p >= '0' && *p <= '9')) return false;
      while (p < end && *p >= '0' && *p <= '9') ++p;
      if (p < end && *p == '.') { is_int = false; ++p; if (p >= end || !(*p >= '0' && *p <= '9')) return false; while (p < end && *p >= '0' && *p <= '9') ++p; }
      if (p < end && (*p == 'e' || *p == 'E')) { is_int = false; ++p; if (p < end && (*p == '+' || *p == '-')) ++p; if (p >= end || !(*p >= '0' && *p <= '9')) return false; while (p < end && *p >= '0' && *p <= '9') ++p; }
      std::string tok(s, p);
      if (is_int && tok.size() < 19) *out = Json(int64_t(strtoll(tok.c_str(), nullptr, 10)));
      else *out = Json(strtod(tok.c_str(), nullptr));
      return true;
    }
    bool value(Json* out, int depth) {
      if (depth > 128 || p >= end) return false;
      switch (*p) {
        case 'n': if (!lit("null")) return false; *out = Json(); return true;
        case 't': if (!lit("true")) return false; *out = Json(true); return true;
        case 'f': if (!lit("false")) return false; *out = Json(false); return true;
        case '"': { std::string s; if (!string(&s)) return false; *out = Json(s); return true; }
        case '[': {
          ++p; *out = Json::array(); ws();
          if (p < end && *p == ']') { ++p; return true; }
          for (;;) {
            Json v; ws(); if (!value(&v, depth + 1)) return false; out->push(std::move(v)); ws();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == ']') { ++p; return true; }
            return false;
          }
        }
        case '{': {
          ++p; *out = Json::object(); ws();
          if (p < end && *p == '}') { ++p; return true; }
          for (;;) {
            std::string k; ws(); if (!string(&k)) return false; ws();
            if (p >= end || *p != ':') return false;
            ++p; ws();
            Json v; if (!value(&v, depth + 1)) return false;
            // big objects (a 128k-entry vocab) skip the duplicate-key scan of set(): first key wins
            if (out->o_.size() < 64) out->set(k, std::move(v)); else out->o_.emplace_back(std::move(k), std::move(v));
            ws();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == '}') { ++p; return true; }
            return false;
          }
        }
        default: return number(out);
      }
    }
  };

  Type t_;
  bool b_ = false;
  int64_t i_ = 0;
  double d_ = 0;
  std::string s_;
  std::vector<Json> a_;
  std::vector<std::pair<std::string, Json>> o_;
};

}  // namespace llmlb_host
