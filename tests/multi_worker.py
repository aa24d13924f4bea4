"""torchrun worker for the N>1 CUDA+NCCL path (launched by tests/test_gpu_multi.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_worker.py

One rank per GPU; every rank builds the same problem, owns a slab of grid rows, and the solver exchanges one
2-row halo per outer iteration plus a few tiny all-gathers (NCCL owned by libgnk_b200.so).  Rank 0 checks the
result against the reference's golden traces; all ranks must hold bit-identical global results.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import gauss_newton_via_generalized_krylov_subspaces_b200 as g
    from golden_util import Golden, Recorder, bound_for, check_trace, rel
    from oracle import gnk_oracle as orc

    rt = g.get_runtime()
    assert rt.world == world and rt.lib.gnk_comm_size(rt.ctx) == world
    p2p = int(rt.lib.gnk_comm_p2p_enabled(rt.ctx))
    assert p2p == (0 if os.environ.get("GNK_P2P", "1") == "0" else 1), "peer-memory mailboxes did not attach"
    assert rt.fused_reductions == bool(p2p and os.environ.get("GNK_P2P_FUSED", "1") != "0")
    if rank == 0:
        fused = int(rt.lib.gnk_comm_fused_reductions(rt.ctx))
        print(f"[multi] collectives: {'peer-memory mailboxes (CUDA IPC over NVLink)' if p2p else 'NCCL'}"
              f"{', reductions finished inside the producing kernels' if fused else ''}", flush=True)

    def gather_equal(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        ts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        return all(torch.equal(ts[0], u) for u in ts[1:])

    # ---- small, uneven slabs, odd row length, restarts ------------------------------------------
    gd = Golden("bratu_g34")
    pb = g.BratuPdeProblem(34, 5, 10)
    assert rel(pb.pde_operator(pb.u_true), gd["y"]) < 1e-14
    res, jac, err = pb.make_res(gd["y"]), pb.make_jac(), pb.make_error()
    o = orc.BratuOracle(34, 5, 10)
    v = np.random.RandomState(0).normal(size=o.n)
    J, Jo = jac(gd["u0"]), o.make_jac()(gd["u0"])
    assert rel(J @ v, Jo @ v) < 1e-14 and rel(J.T @ v, Jo.T @ v) < 1e-14
    gr = gd.run("gnk_res_old")
    rec = Recorder(gr["sample_idx"], err)
    out = g.gauss_newton_krylow(res, gd["u0"], jac, callback=rec, krylow_restart=12, max_iter=40)
    check_trace(rec, gr, 1e-10, upto=12)
    check_trace(rec, gr, 1e-8)
    assert (out.nit, out.nrev, out.njev) == (int(gr["nit"]), int(gr["nfev"]), int(gr["njev"]))
    assert gather_equal(out.x), "ranks disagree on the global result"
    if rank == 0:
        print(f"[multi] G=34 world={world}: nit={out.nit} ok", flush=True)

    # ---- Bratu 1024^2, k <= 30: golden trace of the reference ----------------------------------------
    gd = Golden("bratu_g1025")
    o = orc.BratuOracle(1025, 5, 10)
    y, u0 = o.operator(o.u_true), o.start_vector(seed=42)
    pb = g.BratuPdeProblem(1025, 5, 10)
    res, jac, err = pb.make_res(y), pb.make_jac(), pb.make_error()
    gr = gd.run("gnk_k30")
    rec = Recorder(gr["sample_idx"], err)
    out = g.gauss_newton_krylow(res, u0, jac, callback=rec, max_iter=31)
    check_trace(rec, gr, bound_for("bratu_g1025", "gnk_k30"))
    assert (out.nit, out.nrev) == (30, 31) and gather_equal(out.x)
    if rank == 0:
        xs = np.array(rec.xs)
        d = np.max(np.abs(xs - gr["xs"]) / np.max(np.abs(gr["xs"]), axis=1, keepdims=True), axis=1)
        print(f"[multi] G=1025 world={world}: nit={out.nit} max dev {d.max():.2e} tail dev {d[4:].max():.2e} ok", flush=True)

    # ---- gauss_newton (full space, CGLS inner solve) and cg_least_squares on row slabs: golden traces of the reference
    # at grid_nodes=101 (m = 100: uneven slabs at 3+ ranks do not occur here, 100 = 2*50 = 4*25 = 8*12.5 -> 13/12 rows)
    gd = Golden("bratu_g101")
    pb = g.BratuPdeProblem(101, 5, 10)
    res, jac, err = pb.make_res(gd["y"]), pb.make_jac(), pb.make_error()
    for precond in (False, True):
        gr = gd.run("gn_precond" if precond else "gn")
        rec = Recorder(gr["sample_idx"], err)
        out = g.gauss_newton(res, gd["u0"], jac, callback=rec, cg_preconditioner=precond)
        assert (out.nit, out.nrev, out.njev, out.success) == (4, 5, 4, True), (out.nit, out.nrev, out.njev, out.success)
        assert rec.err[-1] < 1e-10 and rel(out.x, gr["x_final"]) < 1e-10 and gather_equal(out.x)
        for a, b in zip(rec.cg, gr["cg_iter"]):  # CG counts are rounding-sensitive: +-2 %
            assert abs(a - b) <= max(2, 0.02 * b), (rec.cg, list(gr["cg_iter"]))
    o = orc.BratuOracle(101, 5, 10)
    r0 = o.make_res(gd["y"])(gd["u0"])
    xo, ito = orc.cgls(-1 * o.make_jac()(gd["u0"]), r0, rtol=1e-6)
    xc, itc = g.cg_least_squares(-1 * jac(gd["u0"]), r0, cg_rtol=1e-6)
    assert rel(xc, xo) < 1e-6 and abs(itc - ito) <= max(2, 0.02 * ito) and gather_equal(xc)
    if rank == 0:
        print(f"[multi] G=101 gauss_newton world={world}: nit=4, cg iterations {list(rec.cg)} (reference "
              f"{list(gr['cg_iter'])}); cg_least_squares {itc} iterations (oracle {ito}) ok", flush=True)

    # ---- krylow_restart=50 with the QR least squares at 256^2: panels of 33..51 columns take the wide tensor-pipe path
    # (gnk_cholqr_wide_try: the Gram matrices and the refinement vectors are summed over the ranks inside the single-CTA
    # factor kernels) wherever every rank's slab has >= 16384 unknowns (world <= 4), the Householder TSQR otherwise.
    # Checked against the oracle (LAPACK QR) on the same inputs.
    o = orc.BratuOracle(257, 5, 10)
    y, u0 = o.operator(o.u_true), o.start_vector(seed=42)
    pb = g.BratuPdeProblem(257, 5, 10)
    res, jac = pb.make_res(y), pb.make_jac()
    idx = np.random.RandomState(0).choice(o.n, 64, replace=False)
    ref_trace, our_trace = [], []
    ref = orc.gnk(o.make_res(y), u0, o.make_jac(), restart=50, max_iter=61,
                  callback=lambda x, **kw: ref_trace.append(np.asarray(x)[idx].copy()))
    out = g.gauss_newton_krylow(res, u0, jac, krylow_restart=50, max_iter=61,
                                callback=lambda x, **kw: our_trace.append(np.asarray(x)[idx].copy()))
    assert (out.nit, out.nrev, out.njev) == (ref["nit"], ref["nfev"], ref["njev"]) and gather_equal(out.x)
    dev = [float(np.max(np.abs(a - b)) / np.max(np.abs(b))) for a, b in zip(our_trace, ref_trace)]
    assert len(dev) == 60 and max(dev[1:50]) < 1e-10 and max(dev) < 1e-5, dev
    if rank == 0:
        print(f"[multi] G=257 restart 50 (QR) world={world}: nit={out.nit} max dev first cycle {max(dev[:50]):.2e} "
              f"(wide panels {max(dev[32:50]):.2e}) after the restart {max(dev[50:]):.2e} ok", flush=True)

    # ---- the north-star workload itself: Bratu 4096^2, k = 1..30, golden trace of the reference --------------------
    # (skipped with GNK_MULTI_SKIP_4096=1 for quick plumbing checks)
    if os.environ.get("GNK_MULTI_SKIP_4096", "0") != "1":
        gd = Golden("bratu_g4097")
        o = orc.BratuOracle(4097, 5, 10)
        y, u0 = o.operator(o.u_true), o.start_vector(seed=42)
        pb = g.BratuPdeProblem(4097, 5, 10)
        res, jac, err = pb.make_res(y), pb.make_jac(), pb.make_error()
        gr = gd.run("gnk_k30")
        assert np.array_equal(u0[gr["sample_idx"]], gd["u0_sample"]) and np.array_equal(y[gr["sample_idx"]], gd["y_sample"])
        rec = Recorder(gr["sample_idx"], err)
        out = g.gauss_newton_krylow(res, u0, jac, callback=rec, max_iter=31)
        check_trace(rec, gr, bound_for("bratu_g4097", "gnk_k30"))
        assert (out.nit, out.nrev, out.njev) == (int(gr["nit"]), int(gr["nfev"]), int(gr["njev"])), (out.nit, out.nrev, out.njev)
        assert gather_equal(out.x)
        xs = np.array(rec.xs)
        d = np.max(np.abs(xs - gr["xs"]) / np.max(np.abs(gr["xs"]), axis=1, keepdims=True), axis=1)
        assert d[12:].max() < 1e-10, d[12:].max()                       # the plain bar from iteration 13 on
        loss = res.loss(out.x)
        assert abs(loss - gr["loss"][-1]) <= 2e-10 * gr["loss"][-1]
        if rank == 0:
            print(f"[multi] G=4097 world={world}: nit={out.nit} max dev {d.max():.2e} tail dev {d[12:].max():.2e} "
                  f"final loss dev {abs(loss - gr['loss'][-1]) / gr['loss'][-1]:.1e} ok", flush=True)
    if rank == 0:
        print("MULTI_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
