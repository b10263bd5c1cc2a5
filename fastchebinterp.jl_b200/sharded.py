"""Multi-GPU host logic: one process per GPU over torch.distributed (NCCL on the B200 box, gloo in CPU tests).

SURVEY.md §8e.  Only the two naturally sharding paths are partitioned:

* batched evaluation — points are independent: the coefficients (<= 1.4 MB in every config) are broadcast from
  rank `src` over NVLink once, every rank builds its own device handle and evaluates a contiguous block of the
  points; results stay sharded or are gathered on `dst`.  No collective inside the kernel; the out-of-domain
  status is an all-reduce(MIN) of the first offending *global* index (the reference's broadcast aborts on the
  first throw, eval.jl:64).
* batches of independent fits — fit i goes to rank block(i); no collective.  A single N-d fit stays on one GPU.

The compute itself is always `fastcheb.ChebPoly` / `fastcheb.chebcoefs` on the rank's GPU.  `make_poly` /
`fit_fn` exist so the partition/collective logic can be tested on CPU (gloo) with the oracle as the stand-in.
"""
from __future__ import annotations

import numpy as np

NO_BAD = np.iinfo(np.int64).max


def block_range(n: int, world: int, rank: int):
    """Contiguous block partition of range(n): the first n % world ranks get one extra item."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun, one process per GPU)")
    return dist


def _comm_device(dist):
    import torch

    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_coefficients(coefs, lb, ub, src: int = 0):
    """Rank `src` passes numpy arrays, the others None; every rank returns (coefs, lb, ub) as numpy arrays.
    Shapes/dtypes travel as a small object broadcast, the coefficient payload as one tensor broadcast."""
    import torch

    dist = _dist()
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        coefs = np.ascontiguousarray(coefs)
        meta = [(coefs.shape, coefs.dtype.str, np.asarray(lb).tolist(), np.asarray(ub).tolist(),
                 np.asarray(lb).dtype.str, np.asarray(ub).dtype.str)]
    dist.broadcast_object_list(meta, src=src)
    shape, dts, lbl, ubl, lbd, ubd = meta[0]
    dt = np.dtype(dts)
    dev = _comm_device(dist)
    nbytes = int(np.prod(shape)) * dt.itemsize
    if rank == src:
        payload = torch.from_numpy(coefs.reshape(-1).view(np.uint8).copy()).to(dev)
    else:
        payload = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(payload, src=src)
    out = payload.cpu().numpy().view(dt).reshape(shape)
    return out, np.asarray(lbl, dtype=np.dtype(lbd)), np.asarray(ubl, dtype=np.dtype(ubd))


class ShardedChebPoly:
    """A ChebPoly replicated on every rank's GPU; evaluation is sharded over points."""

    def __init__(self, coefs=None, lb=None, ub=None, src: int = 0, device: int | None = None, make_poly=None):
        dist = _dist()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        c, l, u = broadcast_coefficients(coefs, lb, ub, src)
        if make_poly is None:
            import torch

            from . import ChebPoly

            dev = torch.cuda.current_device() if device is None else device
            make_poly = lambda c_, l_, u_: ChebPoly(c_, l_, u_, device=dev)  # noqa: E731
        self.poly = make_poly(c, l, u)
        self.ndim = len(l)

    # -- partition helpers
    def my_range(self, npts: int):
        return block_range(npts, self.world, self.rank)

    def scatter_points(self, pts, npts: int | None = None, src: int = 0):
        """Rank `src` holds all points (npts x N array); every rank receives its block (numpy)."""
        import torch

        dist = _dist()
        dev = _comm_device(dist)
        meta = [None]
        if self.rank == src:
            pts = np.ascontiguousarray(pts)
            meta = [(pts.shape[0], pts.dtype.str)]
        dist.broadcast_object_list(meta, src=src)
        n, dts = meta[0]
        dt = np.dtype(dts)
        lo, hi = block_range(n, self.world, self.rank)
        mine = torch.empty((hi - lo, self.ndim), dtype=torch.from_numpy(np.empty(0, dt)).dtype, device=dev)
        if self.rank == src:
            chunks = []
            for r in range(self.world):
                a, b = block_range(n, self.world, r)
                chunks.append(torch.from_numpy(pts[a:b].reshape(b - a, self.ndim)).to(dev))
            # ragged blocks: point-to-point sends (scatter needs equal sizes)
            reqs = [dist.isend(chunks[r], dst=r) for r in range(self.world) if r != src]
            mine.copy_(chunks[src])
            for q in reqs:
                q.wait()
        else:
            dist.recv(mine, src=src)
        return mine.cpu().numpy() if dev.type == "cpu" else mine

    # -- sharded evaluation
    def eval_local(self, pts_local, jac: bool = False):
        """Evaluate this rank's block.  Raises the reference's ArgumentError on *every* rank if any rank saw an
        out-of-domain point (first offending global index, as the serial broadcast would report)."""
        import torch

        dist = _dist()
        err, local_bad = None, NO_BAD
        res = None
        try:
            if jac:
                from . import chebjacobian

                res = chebjacobian(self.poly, pts_local) if hasattr(self.poly, "_eval") else self.poly.jacobian(pts_local)
            else:
                res = self.poly(pts_local)
        except ValueError as e:
            # only the out-of-domain ArgumentError (it carries the index of the first offending point) takes part in the
            # collective; a DimensionMismatch or any other ValueError is this rank's own error and propagates as is
            if not hasattr(e, "index"):
                raise
            err = e
            local_bad = e.index
        dev = _comm_device(dist)
        n_local = len(pts_local)
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[self.rank] = n_local
        dist.all_reduce(counts)
        offset = int(counts[: self.rank].sum().item())
        flag = torch.tensor([NO_BAD if err is None else offset + int(local_bad)], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        first = int(flag.item())
        if first != NO_BAD:
            from . import ArgumentError

            raise ArgumentError(f"point {first} not in domain {self.poly.lb} to {self.poly.ub}") from err
        return res

    def eval_gather(self, pts_local, dst: int = 0):
        """c.(ys) with ys block-partitioned over the ranks; the values are gathered on `dst` in global point order
        (None on the other ranks).

        NCCL backend with device-resident points (torch CUDA tensors): the FUSED path that bench.py times — rank `dst`
        owns one result buffer (PeerGather), every rank's evaluation kernel stores its block straight into it through
        the CUDA-IPC mapping over NVLink, and `dst` gets a torch tensor that aliases the buffer (no staging copy, no
        collective kernel; the out-of-domain status is an all-reduce(MIN) of the first offending global index).
        Other backends / host points: evaluate locally, then point-to-point gather."""
        import torch

        dist = _dist()
        if dist.get_backend() == "nccl" and isinstance(pts_local, torch.Tensor) and pts_local.is_cuda and hasattr(self.poly, "_handle"):
            return self._eval_gather_fused(pts_local, dst)
        v = self.eval_local(pts_local)
        dev = _comm_device(dist)
        t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v)).to(dev)
        cplx = t.is_complex()
        if cplx:
            t = torch.view_as_real(t.contiguous())
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[self.rank] = t.shape[0]
        dist.all_reduce(counts)
        counts = [int(x) for x in counts.tolist()]
        out = None
        if self.rank == dst:
            bufs = [torch.empty((counts[r],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for r in range(self.world)]
            reqs = [dist.irecv(bufs[r], src=r) for r in range(self.world) if r != dst]
            bufs[dst].copy_(t)
            for q in reqs:
                q.wait()
            out = torch.cat(bufs, dim=0)
            if cplx:
                out = torch.view_as_complex(out)
        else:
            dist.send(t.contiguous(), dst=dst)
        return out

    def _eval_gather_fused(self, pts, dst: int):
        import ctypes

        import torch

        from . import ArgumentError, _capi, _check

        dist = _dist()
        poly = self.poly
        N, K, cplx = poly.ndim, poly.K, poly.is_complex
        E = (K or 1) * (2 if cplx else 1)
        rdt = poly._compute_dtype(np.float32 if pts.dtype == torch.float32 else np.float64)
        tdt = torch.float32 if rdt == np.float32 else torch.float64
        p = pts.to(tdt).reshape(-1, N).contiguous()
        n_local = p.shape[0]
        dev = p.device
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[self.rank] = n_local
        dist.all_reduce(counts)
        counts = [int(x) for x in counts.tolist()]
        offset, total = sum(counts[: self.rank]), sum(counts)
        peer = PeerGather(total_rows=total, row_reals=E, itemsize=np.dtype(rdt).itemsize, dst=dst, device=dev.index)
        status = torch.empty(1, dtype=torch.int64, device=dev)
        sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _check(_capi.lib().fastcheb_eval_async(poly._handle(rdt), p.data_ptr(), n_local, peer.slice_ptr(offset), status.data_ptr(), sp),
               "fastcheb_eval_async: ")
        flag = torch.where(status == NO_BAD, status, status + offset)       # first offending GLOBAL index
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)                         # also orders every rank's kernel before the read
        torch.cuda.current_stream(dev).synchronize()
        dist.barrier()
        first = int(flag.item())
        if first != NO_BAD:
            peer.close()
            raise ArgumentError(f"point {first} not in domain {poly.lb} to {poly.ub}")
        out = None
        if self.rank == dst:
            out = peer.as_tensor(tdt, (total, E)).clone()     # one device-to-device copy out of the IPC buffer
            if cplx:
                out = torch.view_as_complex(out.reshape(total, -1, 2))
            out = out.reshape(total) if K is None else out.reshape(total, K)
        peer.close()
        return out


def sharded_fit(vals_all, ndim: int = 1, src: int = 0, fit_fn=None):
    """`howmany` independent fits (leading batch axis) sharded per rank: rank `src` holds `vals_all`
    (howmany x n1.. [x K]); every rank fits its block on its GPU; rank `src` returns all coefficients in order.
    No collective inside the transform; a single fit is never split."""
    import torch

    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _comm_device(dist)
    meta = [None]
    if rank == src:
        vals_all = np.ascontiguousarray(vals_all)
        meta = [(vals_all.shape, vals_all.dtype.str)]
    dist.broadcast_object_list(meta, src=src)
    shape, dts = meta[0]
    dt = np.dtype(dts)
    howmany = shape[0]
    lo, hi = block_range(howmany, world, rank)
    tdt = torch.from_numpy(np.empty(0, dt)).dtype
    mine = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=tdt, device=dev)
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = block_range(howmany, world, r)
            chunk = torch.from_numpy(vals_all[a:b]).to(dev)
            if r == src:
                mine.copy_(chunk)
            else:
                reqs.append(dist.isend(chunk, dst=r))
        for q in reqs:
            q.wait()
    else:
        dist.recv(mine, src=src)
    if fit_fn is None and dev.type == "cuda" and ndim == 1:
        # NCCL, batches of 1-D fits (configs[4]): the block never leaves the device — fastcheb_fit on the received tensor,
        # coefficients into a device tensor.  (howmany, n[, K]) C-order is the Julia order of back-to-back 1-D fits.
        import ctypes

        from . import _capi, _check

        mine = mine.contiguous()
        lt = torch.empty_like(mine)
        if hi > lo:
            body = tuple(shape[1:])
            K = body[1] if len(body) == 2 else 1
            rcode = _capi.F64 if mine.dtype in (torch.float64, torch.complex128) else _capi.F32
            sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _check(_capi.lib().fastcheb_fit(1, _capi.i64_array(body[:1]), rcode, int(mine.is_complex()), K, hi - lo, mine.data_ptr(),
                                            lt.data_ptr(), None, _capi.MEM_DEVICE, dev.index, sp), "fastcheb_fit: ")
    else:
        if fit_fn is None:
            from . import chebcoefs

            fit_fn = lambda v: chebcoefs(v, ndim=ndim, howmany=v.shape[0], device=torch.cuda.current_device())  # noqa: E731
        local = np.asarray(fit_fn(mine.cpu().numpy())) if hi > lo else np.empty((0,) + tuple(shape[1:]), dt)
        lt = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    if rank == src:
        outs = []
        reqs = []
        for r in range(world):
            a, b = block_range(howmany, world, r)
            buf = torch.empty((b - a,) + tuple(shape[1:]), dtype=tdt, device=dev)
            outs.append(buf)
            if r == src:
                buf.copy_(lt)
            else:
                reqs.append(dist.irecv(buf, src=r))
        for q in reqs:
            q.wait()
        return torch.cat(outs, dim=0).cpu().numpy()
    dist.send(lt, dst=src)
    return None


class PeerGather:
    """Fused evaluation + gather: rank `dst` owns one result buffer of `total_rows x row_reals` reals in its HBM
    (fastcheb_peer_alloc); every other rank maps it through CUDA IPC and hands `mapped + its row offset` to the
    evaluation kernel as the output pointer, so the kernel's own stores cross NVLink/NVSwitch — no staging buffer
    and no collective kernel.  The only synchronisation is the caller's end-of-step stream sync + barrier."""

    def __init__(self, total_rows: int, row_reals: int, itemsize: int = 8, dst: int = 0, device: int | None = None):
        import ctypes

        import torch

        from . import _capi
        from . import _check

        dist = _dist()
        self.rank, self.world, self.dst = dist.get_rank(), dist.get_world_size(), dst
        self.device = torch.cuda.current_device() if device is None else device
        self.row_bytes = int(row_reals) * int(itemsize)
        self.total_rows = int(total_rows)
        self._L = _capi.lib()
        self._own = ctypes.c_void_p()
        self._mapped = ctypes.c_void_p()
        handle = [None]
        if self.rank == dst:
            buf = ctypes.create_string_buffer(64)
            _check(self._L.fastcheb_peer_alloc(ctypes.byref(self._own), self.total_rows * self.row_bytes, buf, self.device),
                   "fastcheb_peer_alloc: ")
            handle = [buf.raw]
        dist.broadcast_object_list(handle, src=dst)
        if self.rank != dst:
            hb = ctypes.create_string_buffer(handle[0], 64)
            _check(self._L.fastcheb_peer_open(ctypes.byref(self._mapped), hb, self.device), "fastcheb_peer_open: ")
        self.base = self._own.value if self.rank == dst else self._mapped.value

    def slice_ptr(self, first_row: int) -> int:
        """Device pointer (valid on THIS rank) of row `first_row` of the gathered buffer."""
        return self.base + int(first_row) * self.row_bytes

    def as_tensor(self, torch_dtype, shape):
        """A torch tensor on the gathering rank that ALIASES the gathered buffer (no copy; valid until close())."""
        import torch

        if self.rank != self.dst:
            raise RuntimeError("only the gathering rank owns the buffer")
        typestr = {torch.float64: "<f8", torch.float32: "<f4"}[torch_dtype]

        class _Alias:  # the CUDA array interface of the library-owned allocation
            __cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(self.base), False),
                                        "version": 2}

        return torch.as_tensor(_Alias(), device=torch.device("cuda", self.device))

    def read(self, first_row: int, nrows: int, dtype=np.float64) -> np.ndarray:
        """Copy rows of the gathered buffer to the host (dst rank only)."""
        if self.rank != self.dst:
            raise RuntimeError("only the gathering rank reads the buffer")
        out = np.empty(int(nrows) * self.row_bytes // np.dtype(dtype).itemsize, dtype=dtype)
        from . import _check

        _check(self._L.fastcheb_peer_read(out.ctypes.data, self.slice_ptr(first_row), out.nbytes, self.device))
        return out

    def close(self):
        dist = _dist()
        dist.barrier()  # nobody unmaps while a peer may still be writing
        if self.rank == self.dst:
            if self._own.value:
                self._L.fastcheb_peer_free(self._own, self.device)
                self._own = type(self._own)()
        elif self._mapped.value:
            self._L.fastcheb_peer_close(self._mapped, self.device)
            self._mapped = type(self._mapped)()
