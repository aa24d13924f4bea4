"""world_size 2 and 3 over gloo on the CPU: the sharded (row-slab) path of the host logic -- slab layout, depth-2
halos, the single halo exchange per outer iteration, rank-ordered reductions, the gathered TSQR R factors -- with
the numpy mock standing in for the CUDA kernels and NCCL (tests/mock_backend.py).  The N>1 CUDA+NCCL path itself is
exercised on the GPU box (tests/test_gpu_multi.py via torchrun) and by bench.py --gpus N.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, G, restart, max_iter, version, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mock_backend
    mock_backend.install()
    import gauss_newton_via_generalized_krylov_subspaces_b200 as g
    from golden_util import Golden, Recorder

    gd = Golden(f"bratu_g{G}")
    pb = g.BratuPdeProblem(G, 5, 10)
    assert pb.dev.fields["rows"] >= 2 and g.get_runtime().world == world
    y = pb.pde_operator(pb.u_true)                      # sharded stencil + all-gather of the owned parts
    res, jac, err = pb.make_res(gd["y"]), pb.make_jac(), pb.make_error()
    J = jac(gd["u0"])
    v = np.random.RandomState(0).normal(size=pb.n)
    gr = gd.run("gnk_res_old")
    rec = Recorder(gr["sample_idx"], err)
    lib = g.get_runtime().lib
    pins, real_method = [], lib.gnk_tsqr_ls_method
    lib.gnk_tsqr_ls_method = lambda ctx, m: (pins.append(int(m)), real_method(ctx, m))[1]
    out = g.gauss_newton_krylow(res, gd["u0"], jac, callback=rec, krylow_restart=restart, max_iter=max_iter,
                                version=version)
    lib.gnk_tsqr_ls_method = real_method
    # slabs of 11-17 grid rows: not every rank qualifies for the tensor-pipe least squares, so the host pins the
    # Householder path (method 1) around every solve and releases it again -- all ranks issue the same collectives
    assert pins and pins[0::2] == [1] * (len(pins) // 2) and pins[1::2] == [0] * (len(pins) // 2)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), y=y, Jv=J @ v, JTv=J.T @ v, xs=np.array(rec.xs),
             xnorm=np.array(rec.xnorm), nfev=np.array(rec.nfev), x=out.x, nit=out.nit, nrev=out.nrev)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gnk_matches_reference_golden(world, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden, rel
    from oracle import gnk_oracle as orc
    G = 34  # m = 33 grid rows: uneven slabs (17+16 / 11+11+11), odd row length
    mp.spawn(_worker, args=(world, _free_port(), G, 12, 40, "res_old", str(tmp_path)), nprocs=world, join=True)
    gd = Golden(f"bratu_g{G}")
    gr = gd.run("gnk_res_old")
    o = orc.BratuOracle(G, 5, 10)
    Jo = o.make_jac()(gd["u0"])
    v = np.random.RandomState(0).normal(size=o.n)
    outs = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    for z in outs:
        assert rel(z["y"], gd["y"]) < 1e-14                      # halo rows were filled correctly
        assert rel(z["Jv"], Jo @ v) < 1e-14 and rel(z["JTv"], Jo.T @ v) < 1e-14
        assert int(z["nit"]) == int(gr["nit"]) and int(z["nrev"]) == int(gr["nfev"])
        assert list(z["nfev"]) == list(gr["nfev_cb"])
        scale = np.max(np.abs(gr["xs"]), axis=1, keepdims=True)
        assert np.max(np.abs(z["xs"][:12] - gr["xs"][:12]) / scale[:12]) < 1e-10       # before the first restart
        assert np.max(np.abs(z["xs"] - gr["xs"]) / scale) < 1e-8                       # restarts amplify rounding
    for z in outs[1:]:                                           # every rank holds the same global result, bit for bit
        assert np.array_equal(z["x"], outs[0]["x"]) and np.array_equal(z["xnorm"], outs[0]["xnorm"])


def _gn_worker(rank, world, port, G, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mock_backend
    mock_backend.install()
    import gauss_newton_via_generalized_krylov_subspaces_b200 as g
    from golden_util import Golden

    gd = Golden(f"bratu_g{G}")
    pb = g.BratuPdeProblem(G, 5, 10)
    res, jac = pb.make_res(gd["y"]), pb.make_jac()
    out = {}
    for tag, precond in (("gn", False), ("gnp", True)):
        xs, its, nf = [], [], []
        o = g.gauss_newton(res, gd["u0"], jac, max_iter=9, cg_preconditioner=precond,
                           callback=lambda x, nfev, cg_iter: (xs.append(np.asarray(x).copy()), its.append(cg_iter),
                                                              nf.append(nfev)))
        out.update({f"{tag}_xs": np.array(xs), f"{tag}_cg": np.array(its), f"{tag}_nfev": np.array(nf), f"{tag}_x": o.x,
                    f"{tag}_counts": np.array([o.nit, o.nrev, o.njev, int(o.success)])})
    # the user plug-in route of step_length_control (host objects, gauss_newton.py:118-120) on several ranks
    calls = []

    def plug(res_, x, r, J, args, d):
        calls.append(1)
        return g.armijo_goldstein(res_, x, r, J, args, d)
    o = g.gauss_newton(res, gd["u0"], jac, max_iter=4, step_length_control=plug, callback=lambda **k: None)
    out["plug_x"], out["plug_calls"] = o.x, np.array([len(calls)])
    # cg_least_squares on the sharded stencil operator, with and without an initial guess
    J = jac(gd["u0"])
    r0 = res(gd["u0"])
    x1, it1 = g.cg_least_squares(-1 * J, r0, cg_rtol=1e-6)
    x2, it2 = g.cg_least_squares(-1 * J, r0, x0=0.5 * x1, cg_rtol=1e-6, preconditioner=False)
    out.update(cg_x=x1, cg_it=np.array([it1, it2]), cg_x2=x2)
    np.savez(os.path.join(out_dir, f"gn{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gauss_newton_matches_the_single_rank_oracle(world, tmp_path):
    """gauss_newton (gauss_newton.py:63-138) and cg_least_squares (:11-60) on row slabs: halo exchange per operator
    application and rank-ordered dot products inside the CG solve, the step's halo rows once per outer iteration."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden, rel
    from oracle import gnk_oracle as orc
    G = 34
    mp.spawn(_gn_worker, args=(world, _free_port(), G, str(tmp_path)), nprocs=world, join=True)
    gd = Golden(f"bratu_g{G}")
    o = orc.BratuOracle(G, 5, 10)
    outs = [np.load(os.path.join(tmp_path, f"gn{r}.npz")) for r in range(world)]
    for tag, precond in (("gn", False), ("gnp", True)):
        xs, its, nf = [], [], []
        ref = orc.gn(o.make_res(gd["y"]), gd["u0"], o.make_jac(), max_iter=9, cg_preconditioner=precond,
                     callback=lambda x, nfev, cg_iter: (xs.append(x.copy()), its.append(cg_iter), nf.append(nfev)))
        for z in outs:
            assert list(z[f"{tag}_counts"]) == [ref["nit"], ref["nfev"], ref["njev"], int(ref["success"])]
            assert list(z[f"{tag}_nfev"]) == nf
            # CG counts are rounding-sensitive (another summation order): a count or two per solve
            assert np.all(np.abs(z[f"{tag}_cg"] - np.array(its)) <= np.maximum(2, 0.02 * np.array(its)))
            assert rel(z[f"{tag}_xs"], np.array(xs)) < 1e-5      # cg_rtol = 1e-4 solves: the iterates agree to the CG tolerance
            assert rel(z[f"{tag}_x"], ref["x"]) < 1e-5
    J0 = o.make_jac()(gd["u0"])
    r0 = o.make_res(gd["y"])(gd["u0"])
    x1, it1 = orc.cgls(-1 * J0, r0, rtol=1e-6)
    x2, it2 = orc.cgls(-1 * J0, r0, rtol=1e-6, preconditioner=False, x0=0.5 * x1)
    for z in outs:
        assert rel(z["cg_x"], x1) < 1e-6 and rel(z["cg_x2"], x2) < 1e-6
        assert abs(int(z["cg_it"][0]) - it1) <= 2 and abs(int(z["cg_it"][1]) - it2) <= 3
        assert int(z["plug_calls"][0]) == 3
    for z in outs[1:]:                                           # every rank holds the same global results, bit for bit
        for key in ("gn_x", "gnp_x", "plug_x", "cg_x", "cg_x2"):
            assert np.array_equal(z[key], outs[0][key])


def _replicated_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mock_backend
    mock_backend.install()
    import gauss_newton_via_generalized_krylov_subspaces_b200 as g
    from gauss_newton_via_generalized_krylov_subspaces_b200 import _lib, rosenbrock_problem as rp
    raised = []
    x0 = np.full(1000, 2.0)
    for name, call in (
            ("gnk", lambda: g.gauss_newton_krylow(rp.res, x0, rp.jac, callback=lambda **k: None, max_iter=4)),
            ("gn", lambda: g.gauss_newton(rp.res, x0, rp.jac, callback=lambda **k: None, max_iter=4)),
            ("lls", lambda: g.linear_least_squares(np.eye(4, 2), np.ones(4))),
            ("cg", lambda: g.cg_least_squares(rp.jac(x0), rp.res(x0))),
            ("krylow", lambda: g.GeneralizedKrylowSubspace().start(x0))):
        try:
            call()
            raised.append((name, ""))
        except _lib.GnkError as e:
            raised.append((name, str(e)))
    np.save(os.path.join(out_dir, f"g{rank}.npy"), np.array(raised))
    dist.barrier()
    dist.destroy_process_group()


def test_replicated_problems_refuse_a_multi_rank_process_group(tmp_path):
    """advisor (round 1, medium): problems that are not row-sharded hold FULL vectors on every rank while the library
    sums its reductions over all ranks (norms sqrt(W) too large, Armijo compares against W*g).  SURVEY 8e says
    "replicas only" for them: every such entry point raises a clear error under a process group with more than one
    rank instead of returning wrong numbers."""
    mp.spawn(_replicated_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(tmp_path, f"g{r}.npy"))
        assert len(got) == 5
        for name, msg in got:
            assert "not row-sharded" in msg, (name, msg)
