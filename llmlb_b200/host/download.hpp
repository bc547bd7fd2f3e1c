// Endpoint side of the gateway's model-download contract:
//
//   POST /api/models/download   {"repo": "<org>/<name>", "filename"?: "<file>"}
//        -> 200 {"task_id","model","status"}                     (llmlb/src/xllm/download.rs:31-41, 76-84, 97-139)
//   GET  /api/download/progress?task_id=<id>
//        -> 200 {"task_id","model","status","progress", "speed_mbps"?, "eta_seconds"?, "error"?, "filename"?}
//                                                                 (download.rs:43-74, 147-190)
//   status: "pending" | "downloading" | "completed" | "failed" | "cancelled"       (download.rs:52)
//
// The reference is the CLIENT of this contract (the server is the external xLLM engine); this file is the server.
// There is no network on the box, so "download" means: fetch <mirror_root>/<repo>/<filename> — a local mirror of the
// hub, e.g. a mounted model store — into <models_dir>/<repo with / -> -->/<filename>, in chunks, with the same
// progress bookkeeping.  Without a filename the best quantisation present in the repository directory is chosen
// (the reference's comment at download.rs:37-38: "xLLM will choose the best quantization").
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "json.hpp"

namespace llmlb_host {

struct DownloadTask {
  std::string task_id, repo, model, filename, src, dst, error;
  std::atomic<int> status{0};               // index into kDownloadStatus
  std::atomic<uint64_t> done{0}, total{0};
  std::atomic<bool> cancel{false};
  double t_start = 0, t_end = 0;            // steady-clock seconds
  std::thread worker;
};

class DownloadManager {
 public:
  // chunk_bytes / throttle_us exist for the tests (observe "downloading" on a small file)
  DownloadManager(std::string mirror_root, std::string models_dir, size_t chunk_bytes = 4u << 20, unsigned throttle_us = 0);
  ~DownloadManager();
  // returns the HTTP status; *resp is the JSON body (error bodies are {"error": "..."} like AppError's)
  int start(const Json& request, Json* resp);
  int progress(const std::string& task_id, Json* resp);
  int cancel(const std::string& task_id, Json* resp);
  // "Q4_K_M" from "Llama-3.2-1B-Instruct-Q4_K_M.gguf"; "" if the name carries none
  static std::string quantization_of(const std::string& filename);
  // the file a request without `filename` resolves to, among `names` (a repository listing); "" if nothing is loadable
  static std::string choose_best(const std::vector<std::string>& names);
  static bool safe_component_path(const std::string& p);   // relative, no "..", no empty / dot segments, no backslashes

 private:
  void run(std::shared_ptr<DownloadTask> t);
  Json describe(const DownloadTask& t) const;
  std::string mirror_root_, models_dir_;
  size_t chunk_;
  unsigned throttle_us_;
  std::mutex mu_;
  std::map<std::string, std::shared_ptr<DownloadTask>> tasks_;
  uint64_t seq_ = 0;
};

extern const char* const kDownloadStatus[5];

}  // namespace llmlb_host
