"""What does a collective cost inside a kernel chain?  ("no microbenchmark of the mailbox
round trip".)

    python tools/microbench_mailbox.py                                   one GPU: the same chains without a collective
    torchrun --nproc-per-node N ... tools/microbench_mailbox.py          N = 2 / 4 / 8

Every line is the device time per launch of a chain of REPS identical, dependent launches on one stream (CUDA events
around the chain, max over the ranks), for two slab sizes: a 64-row-length toy grid (pure latency: nothing to stream) and
the slab a rank owns when 4096^2 is split over 8 GPUs (512 x 4096).  The difference between an N-rank line and the
one-GPU line of the same kernel is what the collective costs where the solver uses it.

  dots k=1 / k=30     dots_kernel      last CTA: flag-in-data all-reduce of k doubles        (p2p_tail_allreduce)
  update              update_kernel    last CTA: (sum, max) pair
  residual            residual_kernel  last CTA: one double
  normalize+halo      normalize_halo_kernel: border threads push 2 rows each way as lines, 32 CTAs receive
  allreduce(1 / 30)   stand-alone p2p_gather_kernel: data + __threadfence_system + st.release flag + ld.acquire spin
                      (the format ALL collectives used before the lines; one launch of its own per collective)
"""
import os, sys, json, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gauss_newton_via_generalized_krylov_subspaces_b200 as g
from gauss_newton_via_generalized_krylov_subspaces_b200 import _lib
from gauss_newton_via_generalized_krylov_subspaces_b200.device import ptr

REPS = int(os.environ.get("REPS", "2000"))
rt = g.get_runtime()
lib = rt.lib


def timed(fn):
    for _ in range(200):
        fn()
    rt.sync()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    rt.sync()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / REPS], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_grid(G):
    """chains on this rank's slab of BratuPdeProblem(G)"""
    pb = g.BratuPdeProblem(G, 5, 10)
    d = pb.dev
    lay, prm, ld, n = d.lay, d.prm, d.ld, d.fields["n_own"]
    kmax = 30
    V = rt.zeros(kmax * ld)
    V.normal_()
    w, x, y, F, eu, out = (rt.zeros(ld) for _ in range(6))
    w.normal_(); x.normal_()
    h, st = rt.zeros(256), rt.zeros(2)
    st.fill_(1.0)
    flag = rt.zeros(1, dtype=torch.int32)
    loss = rt.zeros(8)
    chk = _lib.check
    res = {}
    for k in (1, 30):
        res[f"dots k={k}"] = timed(lambda: chk(lib.gnk_cgs_dots(rt.ctx, C.byref(lay), ptr(V), k, ptr(w), ptr(h), rt.stream)))
    h.zero_()
    res["update k=1"] = timed(lambda: chk(lib.gnk_cgs_update(rt.ctx, C.byref(lay), ptr(V), 1, ptr(h), ptr(w), ptr(st), rt.stream)))
    res["residual"] = timed(lambda: chk(lib.gnk_bratu_residual(rt.ctx, C.byref(lay), C.byref(prm), ptr(x), ptr(y), ptr(F),
                                                               ptr(eu), 1, ptr(loss), rt.stream)))
    st.fill_(1.0)
    res["normalize+halo"] = timed(lambda: chk(lib.gnk_normalize_halo(rt.ctx, C.byref(lay), ptr(w), ptr(st), 1e-8, ptr(out),
                                                                     ptr(flag), rt.stream)))
    res["normalize (no halo)"] = timed(lambda: chk(lib.gnk_normalize(rt.ctx, C.byref(lay), ptr(w), ptr(st), 1e-8, ptr(out),
                                                                      ptr(flag), rt.stream)))
    if world > 1:
        for cnt in (1, 30):
            res[f"allreduce({cnt}) stand-alone, fence+flag"] = timed(
                lambda: chk(lib.gnk_comm_allreduce(rt.ctx, ptr(h), cnt, 0, rt.stream)))
    return dict(grid_nodes=G, rows_owned=int(d.fields["rows"]), row_length=int(pb.m), us_per_launch=res)


out = dict(n_gpus=world, reps=REPS, fused_reductions=bool(rt.fused_reductions),
           note="device time per launch of a chain of dependent launches, max over ranks; see the module docstring",
           toy=bench_grid(65))
# the slab of the 8-GPU run: 4096 columns x 512 rows per rank  ->  a (4096 x 512*world)-row problem does not exist as a
# square grid, so use the square grid whose slab has the same number of unknowns: G - 1 = 4096 / sqrt(8 / world)
G_slab = {1: 1449, 2: 2049, 4: 2897, 8: 4097}.get(world, 4097)
out["slab_2M_unknowns"] = bench_grid(G_slab)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
