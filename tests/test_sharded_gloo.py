"""CPU (gloo, world_size 2) tests of the multi-GPU host logic in fastchebinterp.jl_b200/sharded.py: block
partition, coefficient broadcast, ragged point scatter, ordered gather, global out-of-domain index, sharded
fits.  The compute stand-in is the oracle (test infrastructure) — on the GPU box the same code paths run with
fastcheb.ChebPoly (tests/test_gpu_sharded.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import torch.distributed as dist

        import cheb_oracle as orc
        import fastcheb  # noqa: F401  (loads libfastcheb_cuda.so; no compute call is made without a GPU)
        from fastcheb import sharded as sh

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        rng = np.random.default_rng(5)
        coefs = rng.uniform(-1, 1, (7, 6, 5, 2)) + 1j * rng.uniform(-1, 1, (7, 6, 5, 2))
        lb, ub = np.array([-1.0, 0.0, 2.0]), np.array([1.0, 3.0, 2.5])
        pts = lb + (ub - lb) * rng.random((1001, 3))          # odd count: ragged blocks
        make = lambda c, l, u: orc.OracleChebPoly(c, l, u)    # noqa: E731
        # rank 0 owns the polynomial and all points
        sp = sh.ShardedChebPoly(coefs if rank == 0 else None, lb if rank == 0 else None, ub if rank == 0 else None,
                                src=0, make_poly=make)
        assert np.array_equal(sp.poly.coefs, coefs)            # broadcast is bit-exact, complex K-vector layout intact
        mine = sp.scatter_points(pts if rank == 0 else None, src=0)
        lo, hi = sh.block_range(len(pts), world, rank)
        assert np.array_equal(mine, pts[lo:hi])
        got = sp.eval_gather(mine, dst=0)
        if rank == 0:
            want = orc.OracleChebPoly(coefs, lb, ub)(pts)
            assert np.array_equal(got.numpy(), want)           # ordered gather == serial c.(ys)
        else:
            assert got is None
        # an out-of-domain point on the LAST rank is reported with its GLOBAL index on every rank
        bad = mine.copy()
        if rank == world - 1:
            bad[3, 1] = 99.0
        try:
            sp.eval_local(bad)
            raised = None
        except ValueError as e:
            raised = str(e)
        lo_last, _ = sh.block_range(len(pts), world, world - 1)
        assert raised is not None and f"point {lo_last + 3} " in raised, raised
        # sharded batch of independent fits (no collective inside the transform)
        vals = rng.uniform(-1, 1, (37, 17))
        fit = lambda v: np.stack([orc.chebcoefs(x, 1) for x in v]) if len(v) else v  # noqa: E731
        allc = sh.sharded_fit(vals if rank == 0 else None, ndim=1, src=0, fit_fn=fit)
        if rank == 0:
            assert np.array_equal(allc, fit(vals))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc() + repr(e)))


def test_block_range_partition():
    sys.path.insert(0, ROOT)
    import fastcheb  # noqa: F401
    from fastcheb import sharded as sh
    for n in (0, 1, 7, 8, 1001, 10**9):
        for w in (1, 2, 4, 8):
            blocks = [sh.block_range(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_sharded_eval_and_fit_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
