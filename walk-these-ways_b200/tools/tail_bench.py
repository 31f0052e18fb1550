#!/usr/bin/env python
"""Times go1_mlp_tail_forward against the layer-by-layer tcgen05 products it replaces (CUDA events, L2 flushed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from go1_b200 import capi  # noqa: E402


def main():
    L = capi.lib()
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for M in (4096, 24576):
        for (K1, N2, N3, nh) in ((512, 256, 128, 12), (256, 128, 0, 2)):
            x = torch.randn(M, 1280, device="cuda")[:, :K1]
            W2, b2 = torch.randn(N2, K1, device="cuda") * 0.05, torch.randn(N2, device="cuda")
            W3, b3 = (torch.randn(N3, N2, device="cuda") * 0.05, torch.randn(N3, device="cuda")) if N3 else (None, None)
            NL = N3 or N2
            Wh, bh = torch.randn(nh, NL, device="cuda") * 0.05, torch.randn(nh, device="cuda")
            y2, y3, out = torch.empty(M, N2, device="cuda"), torch.empty(M, max(N3, 1), device="cuda"), torch.empty(M, nh, device="cuda")
            ts = []
            for it in range(6):
                flush.fill_(it)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                capi.check(L.go1_mlp_tail_forward(capi.ptr(x), x.stride(0), M, K1, capi.ptr(W2), capi.ptr(b2), N2, capi.ptr(y2), N2, capi.ptr(W3), capi.ptr(b3), N3,
                                                  capi.ptr(y3) if N3 else None, N3, capi.ptr(Wh), capi.ptr(bh), nh, capi.ptr(out), nh, capi.stream_ptr()), "tail")
                e1.record(); torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            # two problems of this shape in one grid (actor + critic bodies)
            tsg = []
            x2 = torch.randn(M, 1280, device="cuda")[:, :K1]
            y2b, y3b, outb = torch.empty_like(y2), torch.empty_like(y3), torch.empty(M, 1, device="cuda")
            qs = (capi.Go1TailProblem * 2)()
            for q, (xx, yy2, yy3, oo, hh) in zip(qs, ((x, y2, y3, out, nh), (x2, y2b, y3b, outb, 1))):
                q.x, q.ldx, q.W2, q.b2, q.y2, q.ldy2 = xx.data_ptr(), xx.stride(0), W2.data_ptr(), b2.data_ptr(), yy2.data_ptr(), N2
                q.W3, q.b3, q.y3, q.ldy3 = (W3.data_ptr(), b3.data_ptr(), yy3.data_ptr(), N3) if N3 else (None, None, None, 0)
                q.Wh, q.bh, q.nh, q.out, q.ldout = Wh.data_ptr(), bh.data_ptr(), hh, oo.data_ptr(), hh
            for it in range(6):
                flush.fill_(it)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                capi.check(L.go1_mlp_tail_forward_grouped(qs, 2, M, K1, N2, N3, capi.stream_ptr()), "tail2")
                e1.record(); torch.cuda.synchronize()
                if it >= 2:
                    tsg.append(e0.elapsed_time(e1) * 1e3)
            # the separate products
            ts2 = []
            for it in range(6):
                flush.fill_(it)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                capi.check(L.go1_gemm(0, 1, M, N2, K1, capi.ptr(x), x.stride(0), capi.ptr(W2), K1, capi.ptr(y2), N2, capi.ptr(b2), 1, 0, 1, capi.stream_ptr()), "g1")
                if N3:
                    capi.check(L.go1_gemm(0, 1, M, N3, N2, capi.ptr(y2), N2, capi.ptr(W3), N2, capi.ptr(y3), N3, capi.ptr(b3), 1, 0, 1, capi.stream_ptr()), "g2")
                yl = y3 if N3 else y2
                capi.check(L.go1_gemm(0, 1, M, nh, NL, capi.ptr(yl), NL, capi.ptr(Wh), NL, capi.ptr(out), nh, capi.ptr(bh), 0, 0, 1, capi.stream_ptr()), "g3")
                e1.record(); torch.cuda.synchronize()
                if it >= 2:
                    ts2.append(e0.elapsed_time(e1) * 1e3)
            print(f"M={M} tail {K1}-{N2}-{N3}-{nh}: fused {sum(ts) / len(ts):.1f} us, two problems in one grid {sum(tsg) / len(tsg):.1f} us, layer by layer {sum(ts2) / len(ts2):.1f} us", flush=True)


if __name__ == "__main__":
    main()
