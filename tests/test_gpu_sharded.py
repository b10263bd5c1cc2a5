"""GPU: the multi-GPU host logic (fastcheb.sharded) driving the real device path over NCCL.  The driver's GPU
box has one GPU for this tier, so the process group has world_size 1 here (every collective and the
point-to-point gather still execute through NCCL); world_size 2 logic is covered on CPU by
tests/test_sharded_gloo.py and at 2-8 GPUs by bench.py under torchrun."""
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    import torch
    import torch.distributed as dist

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    yield dist
    dist.destroy_process_group()


def test_sharded_eval_matches_oracle(fc, orc, pg):
    import torch
    from fastcheb import sharded as sh

    rng = np.random.default_rng(11)
    coefs = rng.uniform(-1, 1, (13, 11, 9, 3))
    lb, ub = np.array([-1.0, 0.0, 2.0]), np.array([1.0, 3.0, 2.5])
    pts = lb + (ub - lb) * rng.random((5000, 3))
    sp = sh.ShardedChebPoly(coefs, lb, ub, src=0)
    assert isinstance(sp.poly, fc.ChebPoly)
    mine = sp.scatter_points(pts, src=0)                      # device tensor on the NCCL path
    assert isinstance(mine, torch.Tensor) and mine.is_cuda
    got = sp.eval_gather(mine, dst=0).cpu().numpy()
    want = orc.OracleChebPoly(coefs, lb, ub)(pts)
    tol = 16 * np.finfo(np.float64).eps * np.abs(coefs).sum()
    assert np.abs(got - want).max() <= tol
    bad = mine.clone()
    bad[77, 2] = 9.0
    with pytest.raises(fc.ArgumentError, match="point 77 "):
        sp.eval_local(bad)


def test_sharded_fit_matches_oracle(fc, orc, pg):
    from fastcheb import sharded as sh

    rng = np.random.default_rng(12)
    vals = rng.uniform(-1, 1, (3000, 65))
    got = sh.sharded_fit(vals, ndim=1, src=0)
    want = np.stack([orc.chebcoefs(v, 1) for v in vals[:200]])
    assert got.shape == vals.shape
    assert np.abs(got[:200] - want).max() <= 1e-13 * np.abs(want).max()
