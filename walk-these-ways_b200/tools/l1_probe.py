#!/usr/bin/env python
"""Epilogue cost of the fused first-layer product (M x 1280 x 2100): plain / bias / bias+ELU / +extra columns / +lead split."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa
import torch
from go1_b200 import capi
L = capi.lib()
for M in (24576, 4096):
    N, K, E = 1280, 2100, 2
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.02; y = torch.empty(M, N, device="cuda")
    bias = torch.randn(N, device="cuda"); ex = torch.randn(M, E, device="cuda"); wx = torch.randn(N, E, device="cuda")
    for name, kw in (("plain", {}), ("bias", dict(bias=1)), ("bias+elu", dict(bias=1, act=1)), ("bias+elu+extra", dict(bias=1, act=1, extra=1)),
                     ("bias+elu+extra+lead", dict(bias=1, act=1, extra=1, lead=768)), ("elu only", dict(act=1)), ("extra only", dict(extra=1))):
        ep = capi.Go1GemmEpilogue()
        ep.bias = bias.data_ptr() if kw.get("bias") else None
        ep.act, ep.accumulate = kw.get("act", 0), 0
        if kw.get("extra"):
            ep.extra, ep.ld_extra, ep.w_extra, ep.ld_w_extra, ep.num_extra = ex.data_ptr(), E, wx.data_ptr(), E, E
        ep.lead_cols = kw.get("lead", 0)
        ts = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.check(L.go1_gemm_ex(0, 1, M, N, K, capi.ptr(A), K, capi.ptr(W), K, capi.ptr(y), N, ep, 1, capi.stream_ptr()), "gemm")
            e1.record(); torch.cuda.synchronize()
            if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"M={M} {name:22s} {sum(ts)/len(ts):7.1f} us")
