"""Loading real checkpoints through the C ABI: a dependency-free safetensors
reader (8-byte little-endian header length, JSON header, raw tensor bytes — the layout the
reference's PoC lists with safetensors-cpp, poc/nemotron-safetensors-cpp/main.cpp:101-113) that
feeds `llmlb_engine_load_tensor`.  bf16 tensors are passed through as they are; fp16/fp32 are
converted to bf16 (round to nearest even) on the host."""
import json
import mmap
import os
import struct

import numpy as np


def read_safetensors_header(path):
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n).decode("utf-8"))
    header.pop("__metadata__", None)
    return header, 8 + n


def _to_bf16_bits(arr, dtype):
    if dtype == "BF16":
        return arr.view(np.uint16)
    if dtype == "F16":
        arr = arr.view(np.float16).astype(np.float32)
    elif dtype == "F32":
        arr = arr.view(np.float32)
    else:
        raise ValueError("unsupported safetensors dtype " + dtype)
    u = arr.view(np.uint32)
    return ((u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))) >> np.uint32(16)).astype(np.uint16)


def load_safetensors(engine, paths):
    """Feeds every tensor the engine knows (HF Llama names) from one or more .safetensors shards.
    Returns the list of names loaded; unknown names are skipped (e.g. rotary inv_freq buffers)."""
    from . import ffi
    loaded = []
    for path in ([paths] if isinstance(paths, (str, os.PathLike)) else paths):
        header, base = read_safetensors_header(path)
        with open(path, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            try:
                for name, meta in header.items():
                    a, b = meta["data_offsets"]
                    raw = np.frombuffer(mm, dtype=np.uint8, count=b - a, offset=base + a)
                    bits = _to_bf16_bits(raw, meta["dtype"])
                    shape = meta["shape"]
                    rows, cols = (shape[0], shape[1]) if len(shape) == 2 else (1, shape[0])
                    try:
                        engine.load_tensor(name, bits.reshape(rows, cols))
                        loaded.append(name)
                    except ffi.LlmlbError as e:
                        if e.code != ffi.E_NOT_FOUND:
                            raise
                    del raw, bits
            finally:
                mm.close()
    return loaded


def write_safetensors(path, tensors_bf16_bits):
    """Test helper: writes {name: uint16 ndarray} as a BF16 safetensors file."""
    header, off, blobs = {}, 0, []
    for name, arr in tensors_bf16_bits.items():
        a = np.ascontiguousarray(arr, dtype=np.uint16)
        header[name] = {"dtype": "BF16", "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        off += a.nbytes
        blobs.append(a.tobytes())
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)
