"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (omega3d_b200/device.py:
block partition of targets, equally sized zero-padded record streams, one all-gather, every rank evaluates its
own targets against all records). The CUDA engine is replaced by a stand-in that packs with numpy and evaluates
with the oracle, so what is tested is the sharding logic itself: sharded result == unsharded result, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

from omega3d_b200 import _lib
from omega3d_b200 import workloads as W
from omega3d_b200.device import REC_FLOATS, ShardedBiotSavart, ShardedConvection, shard_bounds

f32 = np.float32


class OracleEngine:
    """Stand-in for DeviceBiotSavart on CPU: same packed layout as pp_pack2_kernel (pairs interleaved, positions
    negated, r^2, zero-strength unit-radius padding), evaluation by the C restatement."""

    def __init__(self):
        from oracle import oracle_py
        self.o = oracle_py.Restatement()
        self.o.set_threads(2)
        self.lib = _lib.load()

    def packed_records(self, ns):
        return int(self.lib.o3d_cuda_packed_records(ns))

    def pack(self, x, s, r, out):
        x, s, r = x.numpy(), s.numpy(), r.numpy()
        nrec = out.numel() // REC_FLOATS
        ns = x.shape[1]
        rec = np.zeros((nrec, 8), f32)
        rec[:, 3] = 1.0
        rec[:ns, 0:3] = -x.T
        rec[:ns, 3] = r * r
        rec[:ns, 4:7] = s.T
        pair = rec.reshape(nrec // 2, 2, 8)
        q = np.empty((nrec // 2, 16), f32)
        q[:, 0:2], q[:, 2:4], q[:, 4:6], q[:, 6:8] = pair[:, :, 0], pair[:, :, 1], pair[:, :, 2], pair[:, :, 3]
        q[:, 8:10], q[:, 10:12], q[:, 12:14], q[:, 14:16] = pair[:, :, 4], pair[:, :, 5], pair[:, :, 6], 0
        out.copy_(torch.from_numpy(q.reshape(-1)))
        return out

    def pts_on_pts(self, packed, tx, tr, u, ug):
        q = packed.numpy().reshape(-1, 16)
        sx = np.ascontiguousarray(-np.stack([q[:, 0:2].reshape(-1), q[:, 2:4].reshape(-1), q[:, 4:6].reshape(-1)]))
        sr = np.ascontiguousarray(np.sqrt(q[:, 6:8].reshape(-1)))
        ss = np.ascontiguousarray(np.stack([q[:, 8:10].reshape(-1), q[:, 10:12].reshape(-1), q[:, 12:14].reshape(-1)]))
        self.o.pts_on_pts(sx, sr, ss, tx.numpy(), tr.numpy(), u.numpy(), ug.numpy())


    def finalize(self, u, ug, fs):
        self.o.finalize_vels(u.numpy(), None if ug is None else ug.numpy(), fs)

    def move(self, order, dt, wt, us, ugs, xin, sin, ein, xout, sout, eout, uout=None):
        if xout is not xin:
            xout.copy_(xin)
            sout.copy_(sin)
        self.o.move(order, dt, wt, [u.numpy() for u in us], [None if g is None else g.numpy() for g in ugs], xout.numpy(),
                    sout.numpy(), None if eout is None else eout.numpy(), None if uout is None else uout.numpy())


def _conv_worker(rank, world, port, n, order, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, s, r = W.random_cloud(n, seed=99, radius=0.125)   # 0.125^2 is exact: sqrt(r*r) round-trips through the packed records
    s = (s * f32(40.0)).astype(f32)
    sc = ShardedConvection(n, rank, world, OracleEngine(), order=order)
    lo, hi = sc.lo, sc.hi
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).copy())
    xs, ss, rs, es = t(x[:, lo:hi]), t(s[:, lo:hi]), t(r[lo:hi]), torch.ones(hi - lo)
    u, ug = torch.zeros((3, hi - lo)), torch.zeros((9, hi - lo))
    for _ in range(2):
        sc.advect(0.05, (0.1, 0.0, 0.2), xs, ss, rs, es, u, ug)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), x=xs.numpy(), s=ss.numpy(), e=es.numpy(), u=u.numpy(), ug=ug.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, n, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, s, r = W.random_cloud(n, seed=4242)
    r = (r * (1.0 + 0.25 * np.sin(np.arange(n)))).astype(f32)   # exactly representable? sqrt(r*r) must round-trip: use powers of two below
    r = np.exp2(np.round(np.log2(r))).astype(f32)
    sh = ShardedBiotSavart(n, rank, world, OracleEngine())
    lo, hi = sh.lo, sh.hi
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    u, ug = torch.zeros((3, hi - lo)), torch.zeros((9, hi - lo))
    sh.step(t(x[:, lo:hi]), t(s[:, lo:hi]), t(r[lo:hi]), u, ug)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), lo=lo, hi=hi, u=u.numpy(), ug=ug.numpy(), nrec=sh.rec_per_rank)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_every_target_once():
    for n, w in [(10, 3), (1 << 20, 8), (7, 8), (1000, 1), (1025, 2)]:
        spans = [shard_bounds(n, w, k) for k in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) <= -(-(-(-n // w)) // 512) * 512   # whole tiles: at most 511 more than n/w
        assert all(lo % 512 == 0 or lo == n for lo, _ in spans)


@pytest.mark.timeout(300)
def test_two_ranks_gloo_equal_single_rank(tmp_path):
    n, world = 1500, 2   # tile-aligned blocks: rank 0 owns [0,1024), rank 1 [1024,1500) + 548 padding records
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert [int(p["nrec"]) for p in parts] == [1024, 1024]
    u = np.concatenate([p["u"] for p in parts], axis=1)
    ug = np.concatenate([p["ug"] for p in parts], axis=1)
    assert int(parts[0]["lo"]) == 0 and int(parts[0]["hi"]) == int(parts[1]["lo"]) == 1024 and int(parts[1]["hi"]) == n

    from oracle import oracle_py
    x, s, r = W.random_cloud(n, seed=4242)
    r = (r * (1.0 + 0.25 * np.sin(np.arange(n)))).astype(f32)
    r = np.exp2(np.round(np.log2(r))).astype(f32)
    ru, rg = np.zeros((3, n), f32), np.zeros((9, n), f32)
    oracle_py.Restatement().pts_on_pts(x, r, s, x, r, ru, rg)
    # padding records contribute exactly zero and the gathered order is the global particle order, so the
    # sharded evaluation reproduces the single-rank one bit for bit
    assert np.array_equal(u, ru) and np.array_equal(ug, rg)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("order", [2, 3])
def test_two_ranks_gloo_convection_equals_single_rank(tmp_path, order):
    """ShardedConvection (device.py): two ranks, two Runge-Kutta steps, each rank moving only its own block and
    exchanging only packed records - equals the unsharded oracle step bit for bit."""
    n, world = 700, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_conv_worker, args=(world, port, n, order, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    from oracle import oracle_py
    x, s, r = W.random_cloud(n, seed=99, radius=0.125)
    s = (s * f32(40.0)).astype(f32)
    e = np.ones(n, f32)
    u, ug = oracle_py.Restatement().advect(order, 2, 0.05, (0.1, 0.0, 0.2), x, s, r, e)
    for key, ref in (("x", x), ("s", s), ("e", e), ("u", u), ("ug", ug)):
        mine = np.concatenate([p[key] for p in parts], axis=-1)
        assert np.array_equal(mine, ref), key
