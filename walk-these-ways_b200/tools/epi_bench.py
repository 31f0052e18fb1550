#!/usr/bin/env python
"""Which epilogue feature costs what: the learner's skinny-K tcgen05 products timed with the fused epilogue options switched on one by one
(plain, +ELU' operand, +column sums, +trailing-input backward; forward: plain, +bias, +ELU).  CUDA events, L2 flushed between runs."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from go1_b200 import capi  # noqa: E402


def run(L, M, N, K, ta, tb, feats, flush, reps=5):
    dev = "cuda"
    pad = (lambda n: (n + 31) // 32 * 32) if "pad128" in feats else (lambda n: n)      # row pitch a multiple of 128 bytes
    A = torch.randn(K, pad(M), device=dev)[:, :M] if ta else torch.randn(M, pad(K), device=dev)[:, :K]
    B = torch.randn(N, pad(K), device=dev)[:, :K] if tb else torch.randn(K, pad(N), device=dev)[:, :N]
    Cm = torch.empty(M, N, device=dev)
    y = torch.randn(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    cs = torch.zeros(N, device=dev)
    ex = torch.randn(M, 2, device=dev)
    wx = torch.randn(N, 2, device=dev)
    gwx = torch.zeros(N, 2, device=dev)
    dx = torch.zeros(M, 2, device=dev)
    ep = capi.Go1GemmEpilogue()
    ep.act = 2 if "dact" in feats else (1 if "elu" in feats else 0)
    ep.bias = bias.data_ptr() if "bias" in feats else None
    if "dact" in feats:
        ep.dact_y, ep.ld_dact_y = y.data_ptr(), N
    if "colsum" in feats:
        ep.colsum = cs.data_ptr()
    if "nbx" in feats:
        ep.bwd_extra, ep.ld_bwd_extra, ep.num_bwd_extra = ex.data_ptr(), 2, 2
        ep.bwd_w_extra, ep.ld_bwd_w_extra, ep.g_w_extra, ep.ld_g_w_extra = wx.data_ptr(), 2, gwx.data_ptr(), 2
        ep.d_extra, ep.ld_d_extra = dx.data_ptr(), 2
    if "nex" in feats:
        ep.extra, ep.ld_extra, ep.w_extra, ep.ld_w_extra, ep.num_extra = ex.data_ptr(), 2, wx.data_ptr(), 2, 2
    ts = []
    for it in range(reps + 2):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.check(L.go1_gemm_ex(ta, tb, M, N, K, capi.ptr(A), A.stride(0), capi.ptr(B), B.stride(0), capi.ptr(Cm), N, C.byref(ep), 1, capi.stream_ptr()), "gemm")
        e1.record(); torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


def main():
    L = capi.lib()
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    cases = [
        ("dgrad 24576x512x256", 24576, 512, 256, 0, 0, [(), ("dact",), ("dact", "colsum"), ("dact", "colsum", "nbx"), ("colsum",)]),
        ("dgrad 24576x256x128", 24576, 256, 128, 0, 0, [(), ("dact",), ("dact", "colsum")]),
        ("fwd 24576x256x512", 24576, 256, 512, 0, 1, [(), ("bias",), ("bias", "elu")]),
        ("fwd 24576x128x256", 24576, 128, 256, 0, 1, [(), ("bias", "elu")]),
        ("fwd 24576x1280x2100", 24576, 1280, 2100, 0, 1, [(), ("bias", "elu"), ("bias", "elu", "nex"), ("pad128",), ("pad128", "bias", "elu", "nex")]),
        ("fwd 24576x1280x2112", 24576, 1280, 2112, 0, 1, [()]),
        ("wgrad 1280x2100x24576", 1280, 2100, 24576, 1, 0, [(), ("pad128",)]),
        ("fwd 24576x256x2100", 24576, 256, 2100, 0, 1, [(), ("bias", "elu"), ("pad128", "bias", "elu")]),
        ("fwd 4096x768x2100", 4096, 768, 2100, 0, 1, [(), ("bias", "elu"), ("pad128", "bias", "elu")]),
        ("fwd 4096x256x512", 4096, 256, 512, 0, 1, [(), ("bias", "elu")]),
        ("wgrad 256x512x24576", 256, 512, 24576, 1, 0, [()]),
        ("wgrad 128x256x24576", 128, 256, 24576, 1, 0, [()]),
        ("wgrad 256x2100x24576", 256, 2100, 24576, 1, 0, [(), ("pad128",)]),
    ]
    for name, M, N, K, ta, tb, variants in cases:
        for feats in variants:
            us = run(L, M, N, K, ta, tb, feats, flush)
            bytes_ = 4.0 * (M * K + N * K + M * N * (2 if "dact" in feats else 1))
            print(f"{name:>24s} {'+'.join(feats) or 'plain':>20s} {us:8.1f} us {2.0 * M * N * K / us / 1e6:7.1f} TF/s  {bytes_ / us / 1e3:7.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
