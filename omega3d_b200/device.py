"""Device-resident evaluation: particle arrays already in HBM, one process per GPU.

PyTorch is plumbing here - device memory, the current CUDA stream and ``torch.distributed`` (NCCL over
NVLink) - and nothing else: every kernel launched is this repo's own, through the ``*_dev`` entry points of
``include/o3d_cuda.h``. The multi-GPU scheme is SURVEY.md section 8e: targets are block-partitioned across
ranks, each rank packs the sources it owns into 32-byte records, one all-gather replicates the packed
records, and every rank evaluates its own targets against all records. No reduction is needed.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int64, c_void_p

import torch

from . import _lib
from .influence import CudaContext

REC_FLOATS = 8  # one packed source record = 2 x float4


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


class O3DStage(ctypes.Structure):
    """include/o3d_cuda.h: o3d_stage"""
    _fields_ = [("u", c_void_p * 3), ("ug", c_void_p), ("ug_stride", c_int64)]


def _rows3(t):
    return None if t is None else (c_void_p * 3)(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr())


class DeviceBiotSavart:
    """particles -> points on one GPU with torch tensors as the SoA containers.

    Tensors are float32, C-contiguous rows: x (3,n), s (3,n), r (n,), u (3,n), ug (9,n)."""

    def __init__(self, device: int = 0, ctx: CudaContext | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("omega3d_b200.device needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = ctx or CudaContext((device,))
        self.lib = self.ctx.lib
        self.launches = 0

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def packed_records(self, ns: int) -> int:
        return int(self.lib.o3d_cuda_packed_records(ns))

    def pack(self, x, s, r, out=None):
        """SoA sources -> packed record stream; the whole of `out` is filled (zero-strength records past ns)."""
        ns = x.shape[1]
        if out is None:
            out = torch.empty(self.packed_records(ns) * REC_FLOATS, dtype=torch.float32, device=self.device)
        nrec = out.numel() // REC_FLOATS
        assert x.is_contiguous() and s.is_contiguous()
        self.ctx.check(self.lib.o3d_cuda_pack_sources_dev(self.ctx.h, self._stream(), ns, _p(x[0]), _p(x[1]), _p(x[2]), _p(r),
                                                          _p(s[0]), _p(s[1]), _p(s[2]), nrec, _p(out)))
        self.launches += 1
        return out

    def pts_on_pts(self, packed, tx, tr, u, ug):
        """u (3,nt) += , ug (9,nt) += (None: velocity only); tr None: singular targets."""
        nrec = packed.numel() // REC_FLOATS
        nt = tx.shape[1]
        self.ctx.check(self.lib.o3d_cuda_pts_on_pts_dev(self.ctx.h, self._stream(), nrec, _p(packed), nt, _p(tx[0]), _p(tx[1]),
                                                        _p(tx[2]), _p(tr), _p(u[0]), _p(u[1]), _p(u[2]), _p(ug),
                                                        ug.stride(0) if ug is not None else 0))
        self.launches += self.ctx.last_timing()["launches"]

    def finalize(self, u, ug, fs):
        """finalize_vels on device tensors: u (3,n) = fs + u/4pi, ug (9,n) *= 1/4pi (src/ElementBase.h:187-192, src/Points.h:265-277)."""
        n = u.shape[1]
        f = (c_double * 3)(*[float(v) for v in fs])
        self.ctx.check(self.lib.o3d_cuda_pts_finalize_dev(self.ctx.h, self._stream(), n, _p(u[0]), _p(u[1]), _p(u[2]), _p(ug),
                                                          ug.stride(0) if ug is not None else 0, f))
        self.launches += 1

    def move(self, order, dt, wt, us, ugs, xin, sin, ein, xout, sout, eout, uout=None):
        """Points::move with `order` stages on device tensors (src/ElementBase.h:253-336, src/Points.h:288-520).
        us / ugs: per-stage (3,n) velocities and (9,n) gradients (None: no stretching)."""
        n = xin.shape[1]
        w = (c_double * 3)(*([float(v) for v in wt] + [0.0] * (3 - len(wt))))
        st = (O3DStage * 3)()
        for k in range(order):
            st[k].u = (c_void_p * 3)(us[k][0].data_ptr(), us[k][1].data_ptr(), us[k][2].data_ptr())
            st[k].ug = None if ugs[k] is None else ugs[k].data_ptr()
            st[k].ug_stride = 0 if ugs[k] is None else ugs[k].stride(0)
        self.ctx.check(self.lib.o3d_cuda_pts_move_dev(self.ctx.h, self._stream(), n, order, float(dt), w, st, _rows3(xin), _rows3(sin),
                                                      _p(ein), _rows3(xout), _rows3(sout), _p(eout), _rows3(uout)))
        self.launches += 1

    def set_profiling(self, on: bool):
        self.ctx.check(self.lib.o3d_cuda_set_profiling(self.ctx.h, int(on)))

    def last_kernel_ms(self) -> float:
        ms = c_double()
        self.ctx.check(self.lib.o3d_cuda_dev_kernel_ms(self.ctx.h, byref(ms)))
        return ms.value

    def probe_fp32_peak(self):
        tf, ms = c_double(), c_double()
        self.ctx.check(self.lib.o3d_cuda_probe_fp32_peak(self.ctx.h, byref(tf), byref(ms)))
        return tf.value, ms.value


TILE = 512  # particles per shared-memory tile of the influence kernel (csrc/o3d_common.cuh: kTile)


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block partition in whole 512-particle tiles - the rule of the C ABI's resident collections
    (capi.cu: part_layout). Tile-aligned blocks make the all-gathered record stream identical to the single-GPU
    stream (padding only at the very end), so a sharded evaluation returns the same bits for any number of ranks."""
    per = -(-(-(-n // world)) // TILE) * TILE
    return min(n, per * rank), min(n, per * (rank + 1))


class ShardedBiotSavart:
    """One rank of a target-sharded evaluation. Each rank holds the particles [lo, hi) as BOTH its share
    of the sources and its targets; ``step`` = pack local sources -> all-gather packed records -> evaluate.

    ``backend`` collectives go through ``torch.distributed`` (NCCL on GPUs; the host-side bookkeeping of this
    class is exercised on CPU with gloo in tests/test_sharding.py)."""

    def __init__(self, n_total: int, rank: int, world: int, engine: DeviceBiotSavart | None):
        self.n, self.rank, self.world, self.engine = n_total, rank, world, engine
        self.lo, self.hi = shard_bounds(n_total, world, rank)
        per = shard_bounds(n_total, world, 0)[1] if world > 1 else n_total
        # every rank contributes the same number of records so one all_gather_into_tensor suffices;
        # short ranks pad with zero-strength records (exactly zero contribution)
        self.rec_per_rank = int(_lib.load().o3d_cuda_packed_records(max(per, 1)))
        self.local_packed = None
        self.gathered = None

    def buffers(self, device):
        if self.local_packed is None:
            self.local_packed = torch.empty(self.rec_per_rank * REC_FLOATS, dtype=torch.float32, device=device)
            self.gathered = (self.local_packed if self.world == 1 else
                             torch.empty(self.world * self.rec_per_rank * REC_FLOATS, dtype=torch.float32, device=device))
        return self.local_packed, self.gathered

    def step(self, x, s, r, u, ug):
        """x,s (3,nloc), r (nloc,): this rank's particles; u (3,nloc), ug (9,nloc) accumulate."""
        import torch.distributed as dist
        eng = self.engine
        local, gathered = self.buffers(x.device)
        eng.pack(x, s, r, out=local)   # fills all rec_per_rank records; those past this rank's count have zero strength
        if self.world > 1:
            dist.all_gather_into_tensor(gathered, local)
        eng.pts_on_pts(gathered, x, r, u, ug)


class ShardedConvection:
    """One rank of a target-sharded convection step: Convection::advect (src/Convection.h:208-556) for a particle-only
    system whose particles are block-partitioned over the ranks. Each stage is: zero -> pack own particles ->
    all-gather packed records -> own targets against all records -> finalize -> move own particles. Only the packed
    records ever cross NVLink; positions, strengths, velocities and gradients stay with their owner.

    State tensors (this rank's block): x, s (3,nloc); r, elong (nloc,); u (3,nloc); ug (9,nloc)."""

    def __init__(self, n_total: int, rank: int, world: int, engine, order: int = 2):
        assert 0 < order < 4
        self.order = order
        self.sh = ShardedBiotSavart(n_total, rank, world, engine)
        self.engine = engine
        self.lo, self.hi = self.sh.lo, self.sh.hi
        self._tmp = None

    def find_vels(self, x, s, r, u, ug, fs):
        """Convection::find_vels for (this rank's block of) one collection acting on itself (:130-184)."""
        u.zero_()
        ug.zero_()
        self.sh.step(x, s, r, u, ug)
        self.engine.finalize(u, ug, fs)

    def _interim(self, x, k):
        if self._tmp is None:
            mk = lambda rows: torch.empty((rows, x.shape[1]), dtype=x.dtype, device=x.device)
            self._tmp = [dict(x=mk(3), s=mk(3), u=mk(3), ug=mk(9)) for _ in range(2)]
        return self._tmp[k]

    def advect(self, dt, fs, x, s, r, elong, u, ug):
        """One step; x, s, elong, u are updated in place exactly as the reference updates the collection."""
        e = self.engine
        self.find_vels(x, s, r, u, ug, fs)
        if self.order == 1:
            e.move(1, dt, [1.0], [u], [ug], x, s, elong, x, s, elong)
        elif self.order == 2:
            a = self._interim(x, 0)
            e.move(1, (2.0 / 3.0) * dt, [1.0], [u], [ug], x, s, None, a["x"], a["s"], None)
            self.find_vels(a["x"], a["s"], r, a["u"], a["ug"], fs)
            e.move(2, dt, [0.25, 0.75], [u, a["u"]], [ug, a["ug"]], x, s, elong, x, s, elong, u)
        else:
            a, b = self._interim(x, 0), self._interim(x, 1)
            e.move(1, 0.5 * dt, [1.0], [u], [ug], x, s, None, a["x"], a["s"], None)
            self.find_vels(a["x"], a["s"], r, a["u"], a["ug"], fs)
            e.move(1, 0.75 * dt, [1.0], [a["u"]], [ug], x, s, None, b["x"], b["s"], None)   # vort1's velocity, own gradient
            self.find_vels(b["x"], b["s"], r, b["u"], b["ug"], fs)
            e.move(3, dt, [2.0 / 9.0, 3.0 / 9.0, 4.0 / 9.0], [u, a["u"], b["u"]], [ug, a["ug"], b["ug"]], x, s, elong, x, s, elong, u)
