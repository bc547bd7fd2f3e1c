"""Helpers for the -m gpu parity tests: torch only provides device memory and a stream."""
import ctypes as C

import numpy as np
import torch

from llmlb_b200 import ffi


def dev():
    return torch.device("cuda:0")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def ok(rc):
    ffi.check(rc)


def bf16_randn(shape, std=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * std).to(torch.bfloat16).to(dev())


def sync():
    torch.cuda.synchronize()
