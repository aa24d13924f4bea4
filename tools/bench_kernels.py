#!/usr/bin/env python
"""Per-kernel timing through the C ABI at the north-star size (CUDA events, warm-up, inputs >> L2).

    python tools/bench_kernels.py [--m 4096] [--ks 1,2,4,8,15,16,30,31] [--out gpurun_out/kernels.json]

Prints one line per (kernel, k): ms, algorithmic GB/s (byte model of DESIGN.md) and fraction of the measured HBM
peak (MEASURED_PEAKS.json).  Development / profiling aid; bench.py is the judged entry point.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--ks", default="1,2,3,4,6,8,12,15,16,20,24,30")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import gauss_newton_via_generalized_krylov_subspaces_b200 as g
    from gauss_newton_via_generalized_krylov_subspaces_b200 import _lib
    from gauss_newton_via_generalized_krylov_subspaces_b200.device import ptr
    from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve

    rt = g.get_runtime()
    lib = rt.lib
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    pb = g.BratuPdeProblem(a.m + 1, 5, 10)
    d = pb.dev
    n, ld = d.fields["n_own"], d.ld
    ks = [int(x) for x in a.ks.split(",")]
    kmax = max(ks)
    gen = torch.Generator(device=rt.device).manual_seed(0)
    V = torch.randn(kmax * ld, dtype=torch.float64, device=rt.device, generator=gen) * 1e-3
    JV = torch.empty(kmax * n, dtype=torch.float64, device=rt.device)
    x, y, F, E, w = (d.new_col() for _ in range(5))
    x[d.fields["off"]:d.fields["off"] + n] = torch.randn(n, dtype=torch.float64, device=rt.device, generator=gen)
    y[d.fields["off"]:d.fields["off"] + n] = torch.randn(n, dtype=torch.float64, device=rt.device, generator=gen)
    cc = torch.randn(128, dtype=torch.float64, device=rt.device, generator=gen)
    dd = torch.randn(128, dtype=torch.float64, device=rt.device, generator=gen)
    h = rt.zeros(128)
    st = rt.zeros(2)
    blk = rt.zeros(512)
    flag = rt.zeros(1, dtype=torch.int32)
    lay, prm = C.byref(d.lay), C.byref(d.prm)
    d.residual_into(x, y, F, E, st, depth=1)

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        e[0].record()
        for i in range(a.reps):
            fn()
            e[i + 1].record()
        torch.cuda.synchronize()
        return min(e[i].elapsed_time(e[i + 1]) for i in range(a.reps))

    rows = []

    def report(name, k, ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append(dict(kernel=name, k=k, ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 3)))
        print(f"{name:12s} k={k:3d}  {ms:9.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.3f} of measured peak", flush=True)

    def want(name):
        return not a.only or name in a.only.split(",")

    if want("residual"):
        report("residual", 0, timeit(lambda: d.residual_into(x, y, F, E, st, depth=1)), 32.0 * n)
    if want("spmv_t"):
        report("spmv_t", 1, timeit(lambda: d.apply(E, F, ld, 1, 1.0, 1, w, ld, d.fields["off"])), 24.0 * n)
    if want("normalize"):
        lib.gnk_norm_stats(rt.ctx, lay, ptr(x), ptr(st), rt.stream)
        report("norm_stats", 1, timeit(lambda: lib.gnk_norm_stats(rt.ctx, lay, ptr(x), ptr(st), rt.stream)), 8.0 * n)
        report("normalize", 1, timeit(lambda: lib.gnk_normalize(rt.ctx, lay, ptr(x), ptr(st), 1e-8, ptr(w), ptr(flag),
                                                               rt.stream)), 16.0 * n)
    for k in ks:
        if want("spmm"):
            report("spmm", k, timeit(lambda: d.apply(E, V, ld, k, -1.0, 0, JV, n, 0)), 8.0 * n * (2 * k + 1))
        if want("tsqr"):
            d.apply(E, V, ld, k, -1.0, 0, JV, n, 0)
            report("tsqr", k, timeit(lambda: tsqr_solve(rt, JV, n, n, k, F[d.fields["off"]:], -1.0, blk)),
                   8.0 * n * (k + 1))
        if want("tsqr_hh"):  # the Householder TSQR pinned (gnk_tsqr_ls_method = 1) beside the default path
            d.apply(E, V, ld, k, -1.0, 0, JV, n, 0)
            report("tsqr_hh", k, timeit(lambda: tsqr_solve(rt, JV, n, n, k, F[d.fields["off"]:], -1.0, blk,
                                                           householder=True)), 8.0 * n * (k + 1))
        if want("combine"):
            report("combine", k, timeit(lambda: lib.gnk_combine(rt.ctx, lay, ptr(V), k, ptr(cc), ptr(dd), 1.0, ptr(x),
                                                                rt.stream)), 8.0 * n * (k + 1))
        if want("cgs_dots"):
            report("cgs_dots", k, timeit(lambda: lib.gnk_cgs_dots(rt.ctx, lay, ptr(V), k, ptr(F), ptr(h), rt.stream)),
                   8.0 * n * (k + 1))
        if want("cgs_update"):
            report("cgs_update", k, timeit(lambda: lib.gnk_cgs_update(rt.ctx, lay, ptr(V), k, ptr(h), ptr(w), ptr(st),
                                                                      rt.stream)), 8.0 * n * (k + 2))
        if want("spmm_ls") and k + 1 <= 32:
            F_col = F
            report("spmm_ls", k, timeit(lambda: lib.gnk_stencil_gram_ls(
                rt.ctx, lay, prm, ptr(E), ptr(V), ld, kmax, k, ptr(F_col), -1.0, ptr(JV), n, -1.0, ptr(blk),
                rt.stream)), 8.0 * n * (3 * k + 2))
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(dict(m=a.m, n=n, peak_gbs=peak, rows=rows), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
