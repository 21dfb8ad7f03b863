#!/usr/bin/env python
"""Times the single-operator entry points of libpcad at PlantCaduceus_l32 shapes (B windows x 512 bp):
CUDA events on the launch stream, inputs far larger than L2.  PCAD_LIB selects a library variant.

    python tools/bench_ops.py [--batch 256] [--ops scan,conv,norm,in_proj,out_proj,x_proj,dt_proj]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plantcaduceus_b200 import _lib  # noqa: E402

BF16 = 0


def ptr(t):
    return C.c_void_p(t.data_ptr())


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ops", default="scan,conv,norm,in_proj,out_proj,x_proj,dt_proj")
    ap.add_argument("--d", type=int, default=1024)
    ap.add_argument("--delta-final", type=int, default=0)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    d, L = args.d, 512
    E, R, N = 2 * d, d // 16, 16
    RP = (R + 2 * N + 15) // 16 * 16
    S = 2 * args.batch
    T = S * L
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(0)
    bf = dict(device=dev, dtype=torch.bfloat16)
    ops = args.ops.split(",")
    out = {"batch": args.batch, "lib": os.environ.get("PCAD_LIB", "default")}

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)

    if "scan" in ops:
        u_f, u_r = rnd(T, E), rnd(T, E)
        dl_f, dl_r = rnd(T, E), rnd(T, E)
        bc_f, bc_r = rnd(T, RP), rnd(T, RP)
        xz = rnd(T, 2 * E)
        z = xz[:, E:]
        A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float()).repeat(E, 1) + 0.1 * torch.randn(E, N, device=dev, generator=g))
        A2 = A.clone()
        Dp = torch.ones(E, device=dev)
        bias = torch.full((E,), -4.0, device=dev)
        y = torch.empty(T, E, **bf)
        def f():
            rc = lib.pcad_op_biscan(ptr(u_f), ptr(dl_f), ptr(bc_f), ptr(u_r), ptr(dl_r), ptr(bc_r), RP, R, ptr(z), 2 * E,
                                    ptr(A), ptr(Dp), ptr(bias), ptr(A2), ptr(Dp), ptr(bias), ptr(y), S, L, E, BF16, st)
            assert rc == 0, lib.pcad_last_error(None)
        ms = timeit(f)
        out["scan_ms"] = ms
        out["scan_algo_GBs"] = 817.9e6 / 32 * args.batch * (d / 1024) / ms / 1e6
        del u_f, u_r, dl_f, dl_r, bc_f, bc_r, xz, y
    if "conv" in ops:
        xz = rnd(T, 2 * E)
        w = torch.randn(E, 4, device=dev, generator=g)
        b = torch.randn(E, device=dev, generator=g)
        of, orv = torch.empty(T, E, **bf), torch.empty(T, E, **bf)
        def f():
            rc = lib.pcad_op_conv_silu(ptr(xz), 2 * E, ptr(w), ptr(b), ptr(w), ptr(b), ptr(of), ptr(orv), S, L, E, BF16, st)
            assert rc == 0
        ms = timeit(f)
        out["conv_ms"] = ms
        out["conv_GBs"] = T * E * 2 * 3 / ms / 1e6
        del xz, of, orv
    if "norm" in ops:
        x, r = rnd(T, d), rnd(T, d)
        w = torch.ones(d, device=dev)
        yv, ro = torch.empty(T, d, **bf), torch.empty(T, d, **bf)
        def f():
            rc = lib.pcad_op_add_rmsnorm(ptr(x), ptr(r), ptr(w), ptr(yv), ptr(ro), T, d, C.c_float(1e-5), BF16, BF16, st)
            assert rc == 0
        ms = timeit(f)
        out["norm_ms"] = ms
        out["norm_GBs"] = T * d * 2 * 4 / ms / 1e6
        del x, r, yv, ro
    gemms = {"in_proj": (2 * E, d, d, d, 2 * E), "out_proj": (d, E, E, E, d), "x_proj": (RP, E, E, E, RP),
             "dt_proj": (E, R, RP, R, E)}
    for name, (Nn, K, lda, ldw, ldc) in gemms.items():
        if name not in ops:
            continue
        Amat = rnd(T, lda)
        W = rnd(Nn, ldw, scale=K ** -0.5)
        Cm = torch.empty(T, ldc, **bf)
        gbias = torch.full((Nn,), -4.0, device=dev)
        def f():
            rc = lib.pcad_op_linear(ptr(Amat), ptr(W), ptr(Cm), T, Nn, K, lda, ldw, ldc, BF16, st)
            assert rc == 0, lib.pcad_last_error(None)
        ms = timeit(f)
        out[name + "_ms"] = ms
        out[name + "_TFLOPs"] = 2.0 * T * Nn * K / ms / 1e9
        out[name + "_GBs"] = (T * K + T * Nn + Nn * K) * 2 / ms / 1e6
        del Amat, W, Cm
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
