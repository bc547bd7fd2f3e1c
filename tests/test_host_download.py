"""Endpoint side of the gateway's model-download contract, C++ in llmlb_b200/host/download.cpp, on CPU.
The reference is the CLIENT of this contract (llmlb/src/xllm/download.rs); what is pinned to it here is the wire shape:
every body the server emits must deserialise into the reference's own structs — DownloadRequest (:31-41),
DownloadInitResponse (:76-84), DownloadProgressResponse (:43-74: optional fields are ABSENT, not null; progress is
0.0-100.0; status is one of five words) — and the golden JSON documents of its tests must be accepted as requests /
be producible as responses.  Behaviour behind the contract (a local mirror stands in for the hub: no network on the box)
is checked against a Python restatement in this file."""
import ctypes as C
import json
import os
import time

import pytest

STATUSES = {"pending", "downloading", "completed", "failed", "cancelled"}          # download.rs:52
PREF = ["Q4_K_M", "Q4_K_S", "Q5_K_M", "Q5_K_S", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q6_K", "Q8_0", "BF16", "F16", "FP16", "F32"]


@pytest.fixture(scope="module")
def H():
    from llmlb_b200 import build
    lib = C.CDLL(build.build_host())
    vp, cp = C.c_void_p, C.c_char_p
    sig = {"llmlb_dl_create": (vp, [cp, cp, C.c_size_t, C.c_uint]), "llmlb_dl_destroy": (None, [vp]),
           "llmlb_dl_start": (C.c_int, [vp, cp, C.c_char_p, C.c_size_t]), "llmlb_dl_progress": (C.c_int, [vp, cp, C.c_char_p, C.c_size_t]),
           "llmlb_dl_cancel": (C.c_int, [vp, cp, C.c_char_p, C.c_size_t]), "llmlb_dl_choose_best": (C.c_size_t, [cp, C.c_char_p, C.c_size_t]),
           "llmlb_dl_quantization_of": (C.c_size_t, [cp, C.c_char_p, C.c_size_t]), "llmlb_dl_safe_path": (C.c_int, [cp])}
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def call(fn, *args):
    buf = C.create_string_buffer(8192)
    st = fn(*args, buf, 8192)
    return st, json.loads(buf.value.decode())


def check_progress_shape(p):
    """serde would accept this as DownloadProgressResponse and re-serialise it unchanged (download.rs:43-74)."""
    assert set(p) <= {"task_id", "model", "status", "progress", "speed_mbps", "eta_seconds", "error", "filename"}
    assert isinstance(p["task_id"], str) and isinstance(p["model"], str) and p["status"] in STATUSES
    assert isinstance(p["progress"], (int, float)) and 0.0 <= p["progress"] <= 100.0
    for k, ty in (("speed_mbps", (int, float)), ("eta_seconds", int), ("error", str), ("filename", str)):
        assert k not in p or (p[k] is not None and isinstance(p[k], ty))          # skip_serializing_if = "Option::is_none"
    if "eta_seconds" in p:
        assert 0 <= p["eta_seconds"] < 2 ** 32                                        # u32


def wait_done(H, m, tid, timeout=20):
    t0 = time.time()
    seen = []
    while time.time() - t0 < timeout:
        st, p = call(H.llmlb_dl_progress, m, tid.encode())
        assert st == 200
        check_progress_shape(p)
        seen.append(p)
        if p["status"] in ("completed", "failed", "cancelled"):
            return p, seen
        time.sleep(0.005)
    raise AssertionError("download did not finish: %s" % seen[-1:])


def mirror(tmp_path, files):
    root = tmp_path / "mirror"
    for rel, data in files.items():
        f = root / rel
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_bytes(data)
    return str(root), str(tmp_path / "models")


def test_download_with_filename_copies_the_file_and_reports_progress(H, tmp_path):
    data = os.urandom(3 * 1024 * 1024 + 17)
    root, models = mirror(tmp_path, {"bartowski/Llama-3.2-1B-Instruct-GGUF/Llama-3.2-1B-Instruct-Q4_K_M.gguf": data})
    m = H.llmlb_dl_create(root.encode(), models.encode(), 64 * 1024, 300)        # small chunks + throttle: "downloading" is observable
    try:
        # the reference's own request document (download.rs test_download_request_serialization)
        st, init = call(H.llmlb_dl_start, m, json.dumps({"repo": "bartowski/Llama-3.2-1B-Instruct-GGUF", "filename": "Llama-3.2-1B-Instruct-Q4_K_M.gguf"}).encode())
        assert st == 200 and set(init) == {"task_id", "model", "status"} and init["status"] in STATUSES      # DownloadInitResponse (:76-84)
        final, seen = wait_done(H, m, init["task_id"])
        assert final["status"] == "completed" and final["progress"] == 100.0 and final["filename"] == "Llama-3.2-1B-Instruct-Q4_K_M.gguf"
        assert "error" not in final and "eta_seconds" not in final
        mid = [p for p in seen if p["status"] == "downloading"]
        assert mid and all(0.0 <= p["progress"] <= 100.0 for p in mid)
        assert [p["progress"] for p in seen] == sorted(p["progress"] for p in seen)                            # monotone
        assert any("speed_mbps" in p and p["speed_mbps"] > 0 for p in mid)
        out = os.path.join(models, "bartowski--Llama-3.2-1B-Instruct-GGUF", "Llama-3.2-1B-Instruct-Q4_K_M.gguf")
        assert open(out, "rb").read() == data and not os.path.exists(out + ".part")
    finally:
        H.llmlb_dl_destroy(m)


def test_download_without_filename_picks_the_best_quantisation(H, tmp_path):
    files = {"org/model-GGUF/model-Q8_0.gguf": b"8" * 1000, "org/model-GGUF/model-Q4_K_M.gguf": b"4" * 2000, "org/model-GGUF/model-IQ2_XS.gguf": b"2" * 10,
             "org/model-GGUF/README.md": b"readme", "org/model-GGUF/.hidden.gguf": b"x"}
    root, models = mirror(tmp_path, files)
    m = H.llmlb_dl_create(root.encode(), models.encode(), 0, 0)
    try:
        st, init = call(H.llmlb_dl_start, m, json.dumps({"repo": "org/model-GGUF"}).encode())      # test_download_request_without_filename: no "filename" key
        assert st == 200 and init["model"] == "model-GGUF"
        final, _ = wait_done(H, m, init["task_id"])
        assert final["status"] == "completed" and final["filename"] == "model-Q4_K_M.gguf"
        assert open(os.path.join(models, "org--model-GGUF", "model-Q4_K_M.gguf"), "rb").read() == b"4" * 2000
    finally:
        H.llmlb_dl_destroy(m)


def test_failures_are_reported_in_the_progress_document(H, tmp_path):
    root, models = mirror(tmp_path, {"org/empty/notes.txt": b"n", "org/present/a.gguf": b"a"})
    m = H.llmlb_dl_create(root.encode(), models.encode(), 0, 0)
    try:
        for req, needle in (({"repo": "org/missing"}, "repository not found"), ({"repo": "org/empty"}, "no .gguf or .safetensors"),
                            ({"repo": "org/present", "filename": "b.gguf"}, "file not found")):
            st, init = call(H.llmlb_dl_start, m, json.dumps(req).encode())
            assert st == 200
            final, _ = wait_done(H, m, init["task_id"])
            assert final["status"] == "failed" and needle in final["error"] and "filename" not in final and final["progress"] == 0.0
        for bad, code in (({}, 400), ({"repo": ""}, 400), ({"repo": 7}, 400), ({"repo": "../etc"}, 400), ({"repo": "/abs/path"}, 400),
                          ({"repo": "org/present", "filename": "../../x"}, 400), ({"repo": "org/present", "filename": 3}, 400)):
            st, body = call(H.llmlb_dl_start, m, json.dumps(bad).encode())
            assert st == code and set(body) == {"error"}, bad
        assert call(H.llmlb_dl_start, m, b"{not json")[0] == 400
        st, body = call(H.llmlb_dl_progress, m, b"task-999")
        assert st == 404 and "error" in body
    finally:
        H.llmlb_dl_destroy(m)
    m = H.llmlb_dl_create(b"", b"", 0, 0)                   # endpoint started without a mirror: the route answers 503, nothing is queued
    try:
        st, body = call(H.llmlb_dl_start, m, json.dumps({"repo": "org/present"}).encode())
        assert st == 503 and "not configured" in body["error"]
    finally:
        H.llmlb_dl_destroy(m)


def test_cancel_stops_the_copy_and_removes_the_partial_file(H, tmp_path):
    root, models = mirror(tmp_path, {"org/big/big-F16.gguf": os.urandom(4 * 1024 * 1024)})
    m = H.llmlb_dl_create(root.encode(), models.encode(), 16 * 1024, 2000)       # ~0.5 s for the whole file
    try:
        st, init = call(H.llmlb_dl_start, m, json.dumps({"repo": "org/big", "filename": "big-F16.gguf"}).encode())
        assert st == 200
        time.sleep(0.05)
        st, p = call(H.llmlb_dl_cancel, m, init["task_id"].encode())
        assert st == 200 and p["status"] == "cancelled" and p["progress"] < 100.0
        check_progress_shape(p)
        d = os.path.join(models, "org--big")
        assert not os.path.exists(os.path.join(d, "big-F16.gguf")) and not os.path.exists(os.path.join(d, "big-F16.gguf.part"))
        assert call(H.llmlb_dl_cancel, m, b"nope")[0] == 404
    finally:
        H.llmlb_dl_destroy(m)


def _quant_ref(name):
    stem = name.rsplit("/", 1)[-1]
    stem = stem.rsplit(".", 1)[0] if "." in stem else stem
    import re
    fields = re.split(r"[-.]", stem)
    for f in reversed(fields):
        u = f.upper()
        if re.fullmatch(r"(Q\d.*|IQ\d.*|F16|BF16|F32|FP16)", u):
            return u
    return ""


def _choose_ref(names):
    best, rank_best = "", 1 << 30
    for n in names:
        l = n.lower()
        if l.endswith(".gguf"):
            q = _quant_ref(n)
            rank = PREF.index(q) if q in PREF else len(PREF)
        elif l.endswith(".safetensors"):
            rank = len(PREF) + 1
        else:
            continue
        if rank < rank_best or (rank == rank_best and (len(n) < len(best) or (len(n) == len(best) and n < best))):
            best, rank_best = n, rank
    return best


def test_quantisation_tag_and_best_file_choice_match_the_restatement(H):
    import random
    buf = C.create_string_buffer(512)
    names = ["Llama-3.2-1B-Instruct-Q4_K_M.gguf", "llama-3-8b.Q8_0.gguf", "model-IQ4_XS.gguf", "model.BF16.gguf", "model-f16.gguf", "model.gguf",
             "tokenizer.json", "model-00001-of-00004.safetensors", "model.safetensors", "x-q5_k_s.gguf", "a.b-Q6_K.gguf", "weird-Q.gguf", "q4_0.gguf"]
    for n in names:
        H.llmlb_dl_quantization_of(n.encode(), buf, 512)
        assert buf.value.decode() == _quant_ref(n), n
    rnd = random.Random(5)
    for _ in range(300):
        pick = rnd.sample(names, rnd.randint(0, len(names)))
        H.llmlb_dl_choose_best(json.dumps(pick).encode(), buf, 512)
        assert buf.value.decode() == _choose_ref(pick), pick
    for p, ok in (("org/model", 1), ("a", 1), ("a/b/c.gguf", 1), ("", 0), ("/a", 0), ("a//b", 0), ("a/../b", 0), ("..", 0), ("a/./b", 0), ("a\\b", 0), ("a/", 0)):
        assert H.llmlb_dl_safe_path(p.encode()) == ok, p
