// Endpoint side of POST /api/models/download and GET /api/download/progress: see download.hpp.
#include "download.hpp"

#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstring>

namespace llmlb_host {

const char* const kDownloadStatus[5] = {"pending", "downloading", "completed", "failed", "cancelled"};
enum { kPending = 0, kDownloading = 1, kCompleted = 2, kFailed = 3, kCancelled = 4 };

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static std::string lower(std::string s) { for (auto& c : s) c = char(tolower((unsigned char)c)); return s; }
static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }

bool DownloadManager::safe_component_path(const std::string& p) {
  if (p.empty() || p[0] == '/' || p.find('\\') != std::string::npos || p.find('\0') != std::string::npos) return false;
  size_t i = 0;
  while (i <= p.size()) {
    size_t j = p.find('/', i);
    if (j == std::string::npos) j = p.size();
    const std::string seg = p.substr(i, j - i);
    if (seg.empty() || seg == "." || seg == "..") return false;
    i = j + 1;
  }
  return true;
}

// The quantisation tag is the last '-' or '.' separated field of the stem that looks like one: Q4_K_M, Q8_0, IQ4_XS, F16, BF16, F32.
std::string DownloadManager::quantization_of(const std::string& filename) {
  std::string stem = filename;
  size_t slash = stem.rfind('/');
  if (slash != std::string::npos) stem = stem.substr(slash + 1);
  size_t dot = stem.rfind('.');
  if (dot != std::string::npos) stem = stem.substr(0, dot);
  size_t end = stem.size();
  while (end > 0) {
    size_t start = stem.find_last_of("-.", end - 1);
    const std::string f = stem.substr(start == std::string::npos ? 0 : start + 1, end - (start == std::string::npos ? 0 : start + 1));
    std::string u = f;
    for (auto& c : u) c = char(toupper((unsigned char)c));
    const bool is_q = (u.size() >= 2 && (u[0] == 'Q' || (u.size() >= 3 && u[0] == 'I' && u[1] == 'Q')) &&
                       isdigit((unsigned char)u[u[0] == 'Q' ? 1 : 2])) || u == "F16" || u == "BF16" || u == "F32" || u == "FP16";
    if (is_q) return u;
    if (start == std::string::npos) break;
    end = start;
  }
  return "";
}

// Preference among the files of a repository: GGUF first (one file carries weights, geometry and tokenizer), in the
// order a 4-bit-first default would pick them (what the engine's loader dequantises: gguf.py / checkpoint.cpp); then a
// single-file safetensors.  Ties: shorter name, then lexicographic — a deterministic choice.
std::string DownloadManager::choose_best(const std::vector<std::string>& names) {
  static const char* const pref[] = {"Q4_K_M", "Q4_K_S", "Q5_K_M", "Q5_K_S", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q6_K", "Q8_0", "BF16", "F16", "FP16", "F32"};
  const int n_pref = int(sizeof(pref) / sizeof(pref[0]));
  std::string best;
  int best_rank = 1 << 30;
  for (const std::string& n : names) {
    const std::string l = lower(n);
    int rank;
    if (ends_with(l, ".gguf")) {
      const std::string q = quantization_of(n);
      rank = n_pref;   // a gguf of a type the loader may not know: after the known ones
      for (int i = 0; i < n_pref; ++i) if (q == pref[i]) { rank = i; break; }
    } else if (ends_with(l, ".safetensors")) {
      rank = n_pref + 1;
    } else {
      continue;
    }
    if (rank < best_rank || (rank == best_rank && (n.size() < best.size() || (n.size() == best.size() && n < best)))) { best = n; best_rank = rank; }
  }
  return best;
}

DownloadManager::DownloadManager(std::string mirror_root, std::string models_dir, size_t chunk_bytes, unsigned throttle_us)
    : mirror_root_(std::move(mirror_root)), models_dir_(std::move(models_dir)), chunk_(chunk_bytes ? chunk_bytes : (4u << 20)), throttle_us_(throttle_us) {}

DownloadManager::~DownloadManager() {
  std::vector<std::shared_ptr<DownloadTask>> all;
  { std::lock_guard<std::mutex> lk(mu_); for (auto& kv : tasks_) all.push_back(kv.second); }
  for (auto& t : all) t->cancel.store(true);
  for (auto& t : all) if (t->worker.joinable()) t->worker.join();
}

static Json err_body(const std::string& m) { Json r = Json::object(); r.set("error", m); return r; }

static bool mkdirs(const std::string& path) {
  for (size_t i = 1; i <= path.size(); ++i)
    if (i == path.size() || path[i] == '/') {
      const std::string d = path.substr(0, i);
      if (mkdir(d.c_str(), 0755) != 0 && errno != EEXIST) return false;
    }
  return true;
}

int DownloadManager::start(const Json& request, Json* resp) {
  const Json* repo = request.get("repo");
  if (!request.is_object() || !repo || !repo->is_string() || repo->str().empty()) { *resp = err_body("repo is required"); return 400; }
  const Json* fn = request.get("filename");
  if (fn && !fn->is_null() && !fn->is_string()) { *resp = err_body("filename must be a string"); return 400; }
  const std::string r = repo->str();
  std::string file = fn && fn->is_string() ? fn->str() : "";
  if (!safe_component_path(r) || (!file.empty() && !safe_component_path(file))) { *resp = err_body("repo / filename must be relative paths without '..'"); return 400; }
  if (mirror_root_.empty() || models_dir_.empty()) { *resp = err_body("model download is not configured on this endpoint (no --mirror-root / --models-dir; the host has no network)"); return 503; }

  auto t = std::make_shared<DownloadTask>();
  t->repo = r;
  t->model = r.substr(r.rfind('/') == std::string::npos ? 0 : r.rfind('/') + 1);
  t->filename = file;
  {
    std::lock_guard<std::mutex> lk(mu_);
    t->task_id = "task-" + std::to_string(++seq_);
    tasks_[t->task_id] = t;
  }
  t->t_start = now_s();
  t->worker = std::thread([this, t] { run(t); });
  Json out = Json::object();
  out.set("task_id", t->task_id); out.set("model", t->model); out.set("status", std::string(kDownloadStatus[kPending]));
  *resp = out;
  return 200;
}

void DownloadManager::run(std::shared_ptr<DownloadTask> t) {
  auto fail = [&](const std::string& m) { std::lock_guard<std::mutex> lk(mu_); t->error = m; t->t_end = now_s(); t->status.store(kFailed); };
  const std::string repo_dir = mirror_root_ + "/" + t->repo;
  std::string file = t->filename;
  if (file.empty()) {
    std::vector<std::string> names;
    if (DIR* d = opendir(repo_dir.c_str())) {
      while (dirent* e = readdir(d)) if (e->d_name[0] != '.') names.push_back(e->d_name);
      closedir(d);
    } else { fail("repository not found in the local mirror (this host has no network): " + t->repo); return; }
    file = choose_best(names);
    if (file.empty()) { fail("no .gguf or .safetensors file in " + t->repo); return; }
    std::lock_guard<std::mutex> lk(mu_);
    t->filename = file;
  }
  const std::string src = repo_dir + "/" + file;
  struct stat st;
  if (stat(src.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) { fail("file not found in the local mirror (this host has no network): " + t->repo + "/" + file); return; }
  std::string flat = t->repo;
  for (size_t p; (p = flat.find('/')) != std::string::npos;) flat.replace(p, 1, "--");
  const std::string dst_dir = models_dir_ + "/" + flat;
  std::string dst = dst_dir + "/" + file;
  if (!mkdirs(dst.substr(0, dst.rfind('/')))) { fail("cannot create " + dst_dir + ": " + strerror(errno)); return; }
  { std::lock_guard<std::mutex> lk(mu_); t->src = src; t->dst = dst; }
  t->total.store(uint64_t(st.st_size));
  t->status.store(kDownloading);
  const std::string part = dst + ".part";
  int in = open(src.c_str(), O_RDONLY), out = in >= 0 ? open(part.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644) : -1;
  if (in < 0 || out < 0) { if (in >= 0) close(in); fail(std::string("open failed: ") + strerror(errno)); return; }
  std::vector<char> buf(chunk_);
  bool ok = true;
  while (ok) {
    if (t->cancel.load()) { close(in); close(out); unlink(part.c_str()); std::lock_guard<std::mutex> lk(mu_); t->t_end = now_s(); t->status.store(kCancelled); return; }
    ssize_t n = read(in, buf.data(), buf.size());
    if (n == 0) break;
    if (n < 0) { ok = false; break; }
    for (ssize_t off = 0; off < n;) {
      ssize_t w = write(out, buf.data() + off, size_t(n - off));
      if (w <= 0) { ok = false; break; }
      off += w;
    }
    t->done.fetch_add(uint64_t(n));
    if (throttle_us_) usleep(throttle_us_);
  }
  const int err = errno;
  close(in);
  if (fsync(out) != 0) ok = false;
  close(out);
  if (!ok || t->done.load() != t->total.load() || rename(part.c_str(), dst.c_str()) != 0) { unlink(part.c_str()); fail(std::string("copy failed: ") + strerror(err ? err : errno)); return; }
  std::lock_guard<std::mutex> lk(mu_);
  t->t_end = now_s();
  t->status.store(kCompleted);
}

Json DownloadManager::describe(const DownloadTask& t) const {   // mu_ held
  const int st = t.status.load();
  const uint64_t done = t.done.load(), total = t.total.load();
  Json r = Json::object();
  r.set("task_id", t.task_id); r.set("model", t.model); r.set("status", std::string(kDownloadStatus[st]));
  r.set("progress", st == kCompleted ? 100.0 : total ? 100.0 * double(done) / double(total) : 0.0);
  const double el = (t.t_end > 0 ? t.t_end : now_s()) - t.t_start;
  if ((st == kDownloading || st == kCompleted) && el > 0 && done > 0) {
    const double mbps = double(done) / 1e6 / el;
    r.set("speed_mbps", mbps);
    if (st == kDownloading && mbps > 0) r.set("eta_seconds", uint64_t(double(total - done) / 1e6 / mbps + 0.5));
  }
  if (st == kFailed) r.set("error", t.error);
  if (st == kCompleted) r.set("filename", t.filename);
  return r;
}

int DownloadManager::progress(const std::string& task_id, Json* resp) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = tasks_.find(task_id);
  if (it == tasks_.end()) { *resp = err_body("unknown task_id"); return 404; }
  *resp = describe(*it->second);
  return 200;
}

int DownloadManager::cancel(const std::string& task_id, Json* resp) {
  std::shared_ptr<DownloadTask> t;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = tasks_.find(task_id);
    if (it == tasks_.end()) { *resp = err_body("unknown task_id"); return 404; }
    t = it->second;
  }
  t->cancel.store(true);
  if (t->worker.joinable()) t->worker.join();
  std::lock_guard<std::mutex> lk(mu_);
  *resp = describe(*t);
  return 200;
}

}  // namespace llmlb_host

// ---- C exports for the CPU tests (ctypes) --------------------------------------------------------
using namespace llmlb_host;
static size_t dl_copy(const std::string& s, char* out, size_t cap) {
  if (out && cap) { size_t n = std::min(s.size(), cap - 1); memcpy(out, s.data(), n); out[n] = 0; }
  return s.size();
}
#include "../../include/llmlb_gateway.h"   // the exported signatures are checked against the public header at compile time
extern "C" {
void* llmlb_dl_create(const char* mirror_root, const char* models_dir, size_t chunk_bytes, unsigned throttle_us) {
  return new DownloadManager(mirror_root ? mirror_root : "", models_dir ? models_dir : "", chunk_bytes, throttle_us);
}
void llmlb_dl_destroy(void* m) { delete static_cast<DownloadManager*>(m); }
int llmlb_dl_start(void* m, const char* request_json, char* out, size_t cap) {
  Json req, resp;
  if (!Json::parse(request_json ? request_json : "", &req)) { dl_copy("{\"error\":\"invalid JSON body\"}", out, cap); return 400; }
  int st = static_cast<DownloadManager*>(m)->start(req, &resp);
  dl_copy(resp.dump(), out, cap);
  return st;
}
int llmlb_dl_progress(void* m, const char* task_id, char* out, size_t cap) {
  Json resp;
  int st = static_cast<DownloadManager*>(m)->progress(task_id ? task_id : "", &resp);
  dl_copy(resp.dump(), out, cap);
  return st;
}
int llmlb_dl_cancel(void* m, const char* task_id, char* out, size_t cap) {
  Json resp;
  int st = static_cast<DownloadManager*>(m)->cancel(task_id ? task_id : "", &resp);
  dl_copy(resp.dump(), out, cap);
  return st;
}
size_t llmlb_dl_choose_best(const char* names_json, char* out, size_t cap) {
  Json a;
  std::vector<std::string> names;
  if (Json::parse(names_json ? names_json : "[]", &a) && a.is_array()) for (auto& x : a.items()) if (x.is_string()) names.push_back(x.str());
  return dl_copy(DownloadManager::choose_best(names), out, cap);
}
size_t llmlb_dl_quantization_of(const char* filename, char* out, size_t cap) { return dl_copy(DownloadManager::quantization_of(filename ? filename : ""), out, cap); }
int llmlb_dl_safe_path(const char* p) { return DownloadManager::safe_component_path(p ? p : "") ? 1 : 0; }
}
