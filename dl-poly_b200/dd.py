"""Native multi-GPU driver of the short-range path: DL_POLY's domain decomposition, one domain per GPU.

Mirrors, on top of :class:`engine.ShortRange` and ``torch.distributed``:

* ``map_domains``            -- domains.F90:63-258 (factor search minimising the domain surface, rank -> (idx,idy,idz),
                                the six face neighbours ``map(1:6)``)
* ``relocate_particles``     -- deport_data.F90:2870-3202 (six staged moves -x,+x,-y,+y,-z,+z)
* ``set_halo_particles``     -- halo.F90:153-355 + export_atomic_data (six staged exports, 6 doubles per atom)
* ``refresh_halo_positions`` -- halo.F90:47-113 (same atoms, same order, 3 doubles per atom)
* the per-step control flow of ``md_vv`` around the path (drivers.F90:1910-2290): VV stage 1, ``vnl_check`` + ``gmax``,
  rebuild or refresh, ``two_body_forces``, VV stage 2; the 16-double ``gsum`` is :meth:`Domain.gsum`.

The wire format is the reference's; the transport is NCCL send/recv (``batch_isend_irecv``) on device buffers -- or, for
a direction in which the decomposition has a single domain, the rank is its own neighbour (deport_data.F90:1884-1886)
and the buffer never leaves the GPU.  All device work is enqueued on the context's own CUDA stream (torch sees it as an
``ExternalStream``), so pack -> send/recv -> unpack are ordered without host synchronisation; the only host syncs are the
ones the reference has too (counts on rebuild steps, the ``gmax`` result).

The transport is duck-typed (anything with ``exchange(sendbuf, n_send_items, recvbuf, dst, src)`` and
``exchange_counts``), so the staging logic is exercised on CPU with the gloo backend in tests/test_dd_gloo.py.
"""
import numpy as np

MDIRS = (-1, 1, -2, 2, -3, 3)          # stage order of halo.F90:277-292 / deport_data.F90:3031-3052
HALO_WIDTH = 9                         # DLPGPU_HALO_WIDTH: doubles per atom of a halo-build message
HALF_MINUS = np.nextafter(0.5, 0.0)


# ---------------------------------------------------------------------------------------------- map_domains
def _divisors(p):
    return [d for d in range(1, p + 1) if p % d == 0]


def map_domains(mxnode, widths, imcon=1):
    """domains.F90:103-182: (nx, ny, nz) minimising S = 2(dx dy + dy dz + dz dx) with the reference's tie rules.
    ``widths`` = celprp(7:9), the perpendicular cell widths."""
    if mxnode == 1:
        return 1, 1, 1
    wx, wy, wz = (float(w) for w in widths)
    tol = 1.0e-6
    big = 1 << 30
    limx = big if imcon != 0 else 2
    limy = big if imcon != 0 else 2
    limz = big if (imcon != 0 and imcon != 6) else 2
    best, min_s = (-1, -1, -1), float("inf")
    for nx in _divisors(mxnode):
        if nx > limx:
            continue
        dx = wx / nx
        pyz = mxnode // nx
        for ny in _divisors(pyz):
            if ny > limy:
                continue
            nz = pyz // ny
            if nz > limz:
                continue
            dy, dz = wy / ny, wz / nz
            s = 2.0 * (dx * dy + dy * dz + dz * dx)
            if min_s - s > tol:
                min_s, best = s, (nx, ny, nz)
            elif abs(min_s - s) < tol:
                if max(nx, ny, nz) < max(best):
                    min_s, best = s, (nx, ny, nz)
                elif max(nx, ny, nz) == max(best):
                    if nx < best[0] or (nx == best[0] and ny < best[1]):
                        min_s, best = s, (nx, ny, nz)
    if -1 in best:
        raise RuntimeError("error 520: no domain decomposition found")
    return best


def domain_of_rank(rank, nx, ny, nz):
    """domains.F90:198-200."""
    idz = rank // (nx * ny)
    idy = rank // nx - idz * ny
    idx = rank % nx
    return idx, idy, idz


def idcube(i, j, k, nx, ny):
    return i + nx * (j + ny * k)


def face_neighbours(rank, nx, ny, nz):
    """domains.F90:206-211: map(1:6) = ranks in -x,+x,-y,+y,-z,+z."""
    idx, idy, idz = domain_of_rank(rank, nx, ny, nz)
    return [idcube((idx - 1) % nx, idy, idz, nx, ny), idcube((idx + 1) % nx, idy, idz, nx, ny),
            idcube(idx, (idy - 1) % ny, idz, nx, ny), idcube(idx, (idy + 1) % ny, idz, nx, ny),
            idcube(idx, idy, (idz - 1) % nz, nx, ny), idcube(idx, idy, (idz + 1) % nz, nx, ny)]


def dcell(cell):
    """numerics.F90:1344-1446: lengths, cosines, perpendicular widths and volume of the cell (bbb(1:10)), with the
    reference's operation order (so that Int(width / cutoff) decisions agree bit for bit with the other hosts)."""
    a = [float(v) for v in np.asarray(cell, dtype=np.float64).reshape(9)]
    sq = lambda x, y, z: float(np.sqrt(x * x + y * y + z * z))
    b = [0.0] * 10
    b[0], b[1], b[2] = sq(a[0], a[1], a[2]), sq(a[3], a[4], a[5]), sq(a[6], a[7], a[8])
    b[3] = (a[0] * a[3] + a[1] * a[4] + a[2] * a[5]) / (b[0] * b[1])
    b[4] = (a[0] * a[6] + a[1] * a[7] + a[2] * a[8]) / (b[0] * b[2])
    b[5] = (a[3] * a[6] + a[4] * a[7] + a[5] * a[8]) / (b[1] * b[2])
    axb = (a[1] * a[5] - a[2] * a[4], a[2] * a[3] - a[0] * a[5], a[0] * a[4] - a[1] * a[3])
    bxc = (a[4] * a[8] - a[5] * a[7], a[5] * a[6] - a[3] * a[8], a[3] * a[7] - a[4] * a[6])
    cxa = (a[7] * a[2] - a[8] * a[1], a[8] * a[0] - a[6] * a[2], a[6] * a[1] - a[7] * a[0])
    b[9] = abs(a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2])
    d = (b[9] / sq(*bxc), b[9] / sq(*cxa), b[9] / sq(*axb))
    x = [abs(a[3 * v]) / b[v] for v in range(3)]
    y = [abs(a[3 * v + 1]) / b[v] for v in range(3)]
    if x[0] >= x[1] and x[0] >= x[2]:
        first = 0
    elif x[1] >= x[0] and x[1] >= x[2]:
        first = 1
    else:
        first = 2
    p, q = (1 if first == 0 else 0), (1 if first == 2 else 2)
    b[6] = d[first]
    b[7], b[8] = (d[p], d[q]) if y[p] >= y[q] else (d[q], d[p])
    return b


def cell_widths(cell):
    """celprp(7:9) of numerics.F90::dcell for the lattice vectors in ``cell`` (rows)."""
    b = dcell(cell)
    return b[6], b[7], b[8]


def invert(cell):
    """numerics.F90:1448-1509 with its operation order (adjugate scaled by 1/det)."""
    a = [0.0] + [float(v) for v in np.asarray(cell, dtype=np.float64).reshape(9)]
    b = [0.0] * 10
    b[1] = a[5] * a[9] - a[6] * a[8]; b[2] = a[3] * a[8] - a[2] * a[9]; b[3] = a[2] * a[6] - a[3] * a[5]
    b[4] = a[6] * a[7] - a[4] * a[9]; b[5] = a[1] * a[9] - a[3] * a[7]; b[6] = a[3] * a[4] - a[1] * a[6]
    b[7] = a[4] * a[8] - a[5] * a[7]; b[8] = a[2] * a[7] - a[1] * a[8]; b[9] = a[1] * a[5] - a[2] * a[4]
    d = a[1] * b[1] + a[4] * b[2] + a[7] * b[3]
    r = 1.0 / d if abs(d) > 0.0 else 0.0
    return np.array([r * v for v in b[1:]])


def _anint(x):
    """Fortran Anint: round half away from zero."""
    r = np.rint(x)
    t = np.trunc(x)
    half = np.abs(x - t) == 0.5
    return np.where(half, t + np.sign(x), r)


def read_config_fold(xyz, cell, dims=(1, 1, 1)):
    """configuration.F90:1183-1205: fold every atom into the reduced cell [-0.5,0.5), RECOMPUTE its Cartesian position
    from cell.s, and assign it to domain idm = ipx + nx (ipy + ny ipz), ip = Int((s+0.5) n).  Returns (xyz_folded, owner).
    Same operation order as the reference, so the positions carry the bits a DL_POLY run starts from."""
    x = np.asarray(xyz, dtype=np.float64)
    c = np.asarray(cell, dtype=np.float64).reshape(9)
    rc = invert(c)
    ax, ay, az = x[:, 0], x[:, 1], x[:, 2]
    sx = rc[0] * ax + rc[3] * ay + rc[6] * az
    sy = rc[1] * ax + rc[4] * ay + rc[7] * az
    sz = rc[2] * ax + rc[5] * ay + rc[8] * az
    out = []
    for s in (sx, sy, sz):
        s = s - _anint(s)
        out.append(np.where(s >= HALF_MINUS, -s, s))
    sx, sy, sz = out
    f = np.empty_like(x)
    f[:, 0] = c[0] * sx + c[3] * sy + c[6] * sz
    f[:, 1] = c[1] * sx + c[4] * sy + c[7] * sz
    f[:, 2] = c[2] * sx + c[5] * sy + c[8] * sz
    nx, ny, nz = dims
    ipx = np.clip(((sx + 0.5) * float(nx)).astype(np.int64), 0, nx - 1)
    ipy = np.clip(((sy + 0.5) * float(ny)).astype(np.int64), 0, ny - 1)
    ipz = np.clip(((sz + 0.5) * float(nz)).astype(np.int64), 0, nz - 1)
    return f, (ipx + nx * (ipy + ny * ipz)).astype(np.int32)


def assign_domains(xyz, cell, nx, ny, nz):
    return read_config_fold(xyz, cell, (nx, ny, nz))[1]


# ---------------------------------------------------------------------------------------------- transports
class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class SelfTransport:
    """mxnode == 1 in every direction: never called with a remote peer."""
    rank, world = 0, 1

    def exchange_counts(self, n_send, dst, src):
        raise RuntimeError("SelfTransport cannot reach another rank")

    def allreduce_max(self, v):
        return v

    def allreduce_sum(self, arr):
        return arr

    def allgather_bytes(self, blob):
        return np.ascontiguousarray(blob, dtype=np.uint8).copy()

    def allreduce_sum_device(self, t, stream=None):
        return t

    def barrier(self):
        pass


class ThreadGroup:
    """Shared state of ranks that are THREADS of one process (several domains on one GPU, or one thread per GPU): the host
    side of the collectives is a barrier and a shared table.  The data plane is the library's own peer-memory exchange
    (handle blobs carry plain pointers for same-process peers); ctypes releases the GIL inside every library call, so a rank
    waiting in a stream synchronisation does not stop the others from enqueueing."""

    def __init__(self, world):
        import threading
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world

    def transport(self, rank, device=None):
        return ThreadTransport(self, rank, device)


class ThreadTransport:
    def __init__(self, group, rank, device=None):
        self.g, self.rank, self.world, self.device = group, rank, group.world, device

    def _all(self, v):
        self.g.slots[self.rank] = v
        self.g.bar.wait()
        out = list(self.g.slots)
        self.g.bar.wait()
        return out

    def allreduce_max(self, v):
        return max(float(x) for x in self._all(float(v)))

    def allreduce_sum(self, arr):
        parts = self._all(np.asarray(arr, dtype=np.float64).copy())
        tot = np.zeros_like(parts[0])
        for p in parts:                      # rank order: every rank forms the same bits
            tot = tot + p
        return tot

    def allgather_bytes(self, blob):
        return np.concatenate([np.ascontiguousarray(b, dtype=np.uint8) for b in self._all(np.array(blob, dtype=np.uint8))])

    def barrier(self):
        self.g.bar.wait()

    def allreduce_sum_device(self, t, stream=None):
        """Sum of the ranks' device tensors, in place, in rank order on every rank (same bits everywhere)."""
        import torch
        (stream or torch.cuda.current_stream()).synchronize()          # my contribution is complete
        posts = self._all(t)
        with torch.cuda.stream(stream) if stream is not None else _nullcontext():
            tot = posts[0].clone()
            for p in posts[1:]:
                tot += p
        (stream or torch.cuda.current_stream()).synchronize()
        self.g.bar.wait()                                               # everybody has read everybody's contribution
        with torch.cuda.stream(stream) if stream is not None else _nullcontext():
            t.copy_(tot)
        return t

    def exchange_counts(self, n_send, dst, src):
        """deport_data.F90:1888-1893 between threads: every rank posts (dst, count); the receiver picks the one addressed to it."""
        posts = self._all((int(dst), int(n_send)))
        return posts[src][1]

    def exchange(self, sendbuf, n_send, recvbuf, n_recv, dst, src):
        import torch
        torch.cuda.current_stream().synchronize()
        posts = self._all((sendbuf, int(n_send)))
        sb, ns = posts[src]
        assert ns == n_recv
        if n_recv > 0:
            recvbuf[:n_recv].copy_(sb[:n_recv])
        torch.cuda.current_stream().synchronize()
        self.g.bar.wait()                    # nobody reuses its send buffer before every receiver has copied


class TorchTransport:
    """NCCL (GPU) or gloo (CPU tests) point-to-point through torch.distributed."""

    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._cnt_s = torch.zeros(1, dtype=torch.int64, device=device)
        self._cnt_r = torch.zeros(1, dtype=torch.int64, device=device)
        self._red = torch.zeros(1, dtype=torch.float64, device=device)

    def _p2p(self, ops):
        for r in self.dist.batch_isend_irecv(ops):
            r.wait()

    def exchange_counts(self, n_send, dst, src):
        """Send my count to ``dst`` and receive the count ``src`` sends me (deport_data.F90:1888-1893)."""
        d = self.dist
        self._cnt_s.fill_(int(n_send))
        self._p2p([d.P2POp(d.isend, self._cnt_s, dst, self.group), d.P2POp(d.irecv, self._cnt_r, src, self.group)])
        return int(self._cnt_r.item())

    def exchange(self, sendbuf, n_send, recvbuf, n_recv, dst, src):
        """sendbuf[:n_send] -> dst ; recvbuf[:n_recv] <- src (flat float64 tensors on the transport's device)."""
        d = self.dist
        ops = []
        if n_send > 0:
            ops.append(d.P2POp(d.isend, sendbuf[:n_send], dst, self.group))
        if n_recv > 0:
            ops.append(d.P2POp(d.irecv, recvbuf[:n_recv], src, self.group))
        if ops:
            self._p2p(ops)

    def allreduce_max(self, v):
        self._red.fill_(float(v))
        self.dist.all_reduce(self._red, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(self._red.item())

    def allreduce_sum(self, arr):
        t = self.torch.as_tensor(np.asarray(arr, dtype=np.float64), device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allreduce_sum_device(self, t, stream=None):
        """In-place all-reduce of a device tensor (NCCL over NVLink; gloo in the CPU tests), enqueued on ``stream``."""
        if stream is not None:
            with self.torch.cuda.stream(stream):
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather_bytes(self, blob):
        t_blob = self.torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8).copy()).to(self.device)
        allb = [self.torch.empty_like(t_blob) for _ in range(self.world)]
        self.dist.all_gather(allb, t_blob, group=self.group)
        return self.torch.cat(allb).cpu().numpy()

    def barrier(self):
        self.dist.barrier(group=self.group)


# ---------------------------------------------------------------------------------------------- staging logic
def staged_exchange(transport, neighbours, dims, pack, unpack, alloc, width, self_copy=None):
    """The six-stage exchange pattern shared by halo build and migration.

    For mdir in (-1,+1,-2,+2,-3,+3): ``pack(mdir, buf, cap)`` -> (rc, n_atoms) fills ``buf`` with ``width`` doubles per
    atom (rc == 54: buffer too small, n_atoms = needed); the buffer goes to the neighbour in direction mdir and the
    one arriving from the opposite neighbour is handed to ``unpack(mdir, buf, n_atoms)``.
    ``alloc(n_doubles)`` returns a (buffer, raw pointer) pair on the right device.
    """
    bufs = {}

    def get(kind, n_atoms):
        need = max(n_atoms, 1) * width
        b = bufs.get(kind)
        if b is None or b[2] < need:
            cap = need + need // 4 + 64 * width
            t, p = alloc(cap)
            b = (t, p, cap)
            bufs[kind] = b
        return b

    for q, mdir in enumerate(MDIRS):
        axis = abs(mdir) - 1
        dst = neighbours[q]                       # map(1..6)[q]: the neighbour the data travels to
        src = neighbours[q ^ 1]                   # and the opposite one it arrives from
        st, sp, scap = get("s", 0)
        rc, n = pack(mdir, sp, scap // width)
        if rc == 54:
            st, sp, scap = get("s", n)
            rc, n2 = pack(mdir, sp, scap // width)
            assert rc == 0 and n2 == n
        if dims[axis] == 1:                       # the rank is its own neighbour: no transport
            unpack(mdir, sp, n)
            continue
        n_in = transport.exchange_counts(n, dst, src)
        rt, rp, rcap = get("r", n_in)
        transport.exchange(st, n * width, rt, n_in * width, dst, src)
        unpack(mdir, rp, n_in)
    return bufs


def exchange_capacities(sysm, dims, safety=2.0):
    """Per-stage receive capacities (atoms) of the fused device-side exchange, identical on every rank: the halo slab of
    the widest face (link-cell width w = domain width / Int(domain width / cutoff_extended), halo.F90:219-239, grown by the
    halo layers already received in the earlier directions) at the mean density, times ``safety``; migration: a layer of
    one padding thickness across that face (an atom moves less than padding / 2 between rebuilds, neighbours.F90:182)."""
    b = dcell(sysm.cell)
    rho = sysm.megatm / b[9]
    wid = [w / d for w, d in zip(b[6:9], dims)]
    rx = sysm.rcut + sysm.padding
    lw = [wd / max(int(wd / (rx + 1.0e-6)), 1) for wd in wid]
    full = [wd + 2.0 * l for wd, l in zip(wid, lw)]
    face = max(full[1] * full[2] * lw[0], full[0] * full[2] * lw[1], full[0] * full[1] * lw[2])
    cap_h = int(safety * rho * face) + 4096
    area = max(wid[1] * wid[2], wid[0] * wid[2], wid[0] * wid[1])
    cap_r = int(safety * rho * area * max(sysm.padding, 0.05 * rx)) + 4096
    return cap_r, cap_h


class Domain:
    """One DL_POLY domain resident on one GPU."""

    def __init__(self, sysm, device=0, transport=None, capacity_factor=1.6):
        import torch
        from . import engine
        self.torch = torch
        self.sys = sysm
        self.t = transport if transport is not None else SelfTransport()
        self.rank, self.world = self.t.rank, self.t.world
        self.dims = map_domains(self.world, cell_widths(sysm.cell), sysm.imcon)
        nx, ny, nz = self.dims
        self.idx = domain_of_rank(self.rank, nx, ny, nz)
        self.neigh = face_neighbours(self.rank, nx, ny, nz)
        self.device = torch.device("cuda", device)
        self.sr = engine.ShortRange(device, (nx, ny, nz) + tuple(self.idx))
        self.sr.dev_setup_system(sysm)
        self.stream = torch.cuda.ExternalStream(self.sr.stream(), device=self.device)
        xyz_f, owner = read_config_fold(sysm.xyz, sysm.cell, (nx, ny, nz))
        mine = np.nonzero(owner == self.rank)[0]
        self.natms0 = len(mine)
        cap = int(capacity_factor * max(sysm.megatm / self.world, 1.0) * 1.0) + 4096
        # halo thickness grows the resident count: capacity covers local + halo with head-room (bounds.F90 mxatms)
        ltg = (mine + 1).astype(np.int32)
        vel = None if sysm.vel is None else sysm.vel[mine]
        self.sr.dev_load_atoms(xyz_f[mine], vel, ltg, sysm.lsite[mine], capacity=cap * 2)
        self._refresh_bufs = None
        self.rebuilds = 0
        self.steps = 0
        self._pending = False
        self.last_out = None
        self.acc = {"pair_ms": 0.0, "force_ms": 0.0, "force_calls": 0, "list_ms": 0.0, "list_builds": 0, "xchg_ms": 0.0}
        import os
        # peer-memory refresh: publish buffers sized for the local atoms with head-room for migration
        self.p2p = os.environ.get("DLP_DD_STAGED_REFRESH") is None
        def all_ok(ok):
            """True only if the step succeeded on every rank (peer memory is all-or-nothing across the ranks)."""
            return ok if self.world == 1 else self.t.allreduce_max(0.0 if ok else 1.0) == 0.0

        gather_blobs = self.t.allgather_bytes

        from .lib import DlpError
        if self.p2p:
            blob = self.sr.dev_p2p_init(self.rank, self.world, int(1.25 * self.natms0) + 8192)
            if self.world > 1:
                allb = gather_blobs(blob)
                try:
                    self.sr.dev_p2p_open(allb); ok = True
                except DlpError:
                    ok = False          # no CUDA-IPC peer access between these GPUs: fall back to the NCCL message path
                self.p2p = all_ok(ok)
        # fused device-side exchange (migration + halo build + gmax over peer memory, no NCCL, one host sync per rebuild)
        self.xchg = self.p2p and os.environ.get("DLP_DD_STAGED_EXCHANGE") is None
        self.rseq = 0
        self.gseq = 0
        if self.xchg:
            cap_r, cap_h = exchange_capacities(sysm, self.dims, safety=float(os.environ.get("DLP_DD_CAP_SAFETY", "2.0")))
            blob = self.sr.dev_xchg_init(self.rank, self.world, cap_r, cap_h)
            if os.environ.get("DLP_DD_SCAN_MIGRATION") is not None:      # diagnostic: the all-atom migration stages (tests compare the two)
                self.sr.dev_xchg_set_migration(1)
            if os.environ.get("DLP_DD_XCHG_TIMEOUT") is not None:       # seconds a receive kernel waits for its peer (library default 60)
                self.sr.dev_xchg_set_timeout(float(os.environ["DLP_DD_XCHG_TIMEOUT"]))
            if self.world > 1:
                allb = gather_blobs(blob)
                try:
                    self.sr.dev_xchg_open(allb); ok = True
                except DlpError:
                    ok = False
                self.xchg = all_ok(ok)
                self.t.barrier()
        self.profile = {} if os.environ.get("DLP_DD_PROFILE") else None

    # ---- buffers on the context's device
    def _alloc(self, n_doubles):
        t = self.torch.empty(int(n_doubles), dtype=self.torch.float64, device=self.device)
        return t, t.data_ptr()

    # ---- the reference's three exchange routines
    def relocate(self):
        sr = self.sr
        with self.torch.cuda.stream(self.stream):
            if self.world == 1:
                sr.dev_relocate_serial()
                return
            sr.dev_relocate_begin()
            staged_exchange(self.t, self.neigh, self.dims, sr.dev_relocate_pack, sr.dev_relocate_unpack, self._alloc, 12)
            sr.dev_relocate_end()

    def set_halo(self):
        sr = self.sr
        with self.torch.cuda.stream(self.stream):
            if self.world == 1:
                sr.dev_halo_serial()
                return
            sr.dev_halo_begin()
            staged_exchange(self.t, self.neigh, self.dims, sr.dev_halo_pack, sr.dev_halo_unpack, self._alloc, HALO_WIDTH)
            sr.dev_halo_end()
            self._refresh_bufs = None

    def refresh_halo(self, staged=False):
        """refresh_halo_positions.  Default: one pull kernel over peer memory (needs publish() + a collective since the
        positions moved -- step() provides both); staged=True: the reference's six dependent messages."""
        sr = self.sr
        with self.torch.cuda.stream(self.stream):
            if self.p2p and not staged:
                sr.dev_refresh_pull()
                return
            if self.world == 1:
                sr.dev_refresh_serial()
                return
            if self._refresh_bufs is None:
                cs, cr = sr.dev_halo_stage_counts()
                ns, nr = max(max(cs), 1) * 3, max(max(cr), 1) * 3
                self._refresh_bufs = (self._alloc(ns), self._alloc(nr), cs, cr)
            (st, sp), (rt, rp), cs, cr = self._refresh_bufs
            for q, mdir in enumerate(MDIRS):
                axis = abs(mdir) - 1
                n = sr.dev_refresh_pack(mdir, sp)
                if self.dims[axis] == 1:
                    sr.dev_refresh_unpack(mdir, sp, n)
                    continue
                self.t.exchange(st, n * 3, rt, cr[q] * 3, self.neigh[q], self.neigh[q ^ 1])
                sr.dev_refresh_unpack(mdir, rp, cr[q])

    # ---- md_vv around the path
    def rebuild(self):
        if self.xchg:
            self.rseq += 1
            with self.torch.cuda.stream(self.stream):
                self.sr.dev_xchg_rebuild(self.neigh, self.rseq)
                self._refresh_bufs = None
                self.sr.dev_link_cell_pairs()
            self.rebuilds += 1
            self.acc["list_ms"] += self.sr.last_timings()["list_ms"]; self.acc["list_builds"] += 1
            return
        self.relocate()
        self.set_halo()
        with self.torch.cuda.stream(self.stream):
            self.sr.dev_link_cell_pairs()
        self.rebuilds += 1
        self.acc["list_ms"] += self.sr.last_timings()["list_ms"]; self.acc["list_builds"] += 1

    def publish(self):
        if self.p2p:
            with self.torch.cuda.stream(self.stream):
                self.sr.dev_publish()

    def vnl_update(self):
        if self.xchg:
            self.gseq += 1
            with self.torch.cuda.stream(self.stream):
                tol = self.sr.dev_xchg_gmax(self.gseq)      # gmax over the GPUs' mailboxes, neighbours.F90:176
            return self.sr.vnl_update(tol)
        with self.torch.cuda.stream(self.stream):
            tol = self.sr.dev_vnl_check()
            tol = self.t.allreduce_max(tol)                 # gmax, neighbours.F90:176
        return self.sr.vnl_update(tol)

    def forces(self):
        with self.torch.cuda.stream(self.stream):
            out = self.sr.dev_two_body_forces(zero_forces=True)
        self._account()
        return out

    def _account(self):
        t = self.sr.last_timings()
        self.acc["pair_ms"] += t["pair_kernel_ms"]; self.acc["force_ms"] += t["force_ms"]; self.acc["force_calls"] += 1

    def collect(self):
        """Energies / virial / stress of the last lazily issued force call (see step(lazy=True)); None if none is pending."""
        if not self._pending:
            return self.last_out
        with self.torch.cuda.stream(self.stream):
            self.last_out = self.sr.dev_fetch_results()
        self._pending = False
        self._account()
        return self.last_out

    def step(self, dt, lazy=False):
        """One velocity-Verlet step with the short-range path as the only force provider.  lazy=True does not wait for the
        step's energies: they are collected behind the NEXT host synchronisation (the gmax of the following step, or
        collect()), so a step costs one host round trip instead of two; the return value is then the previous step's sums."""
        sr = self.sr
        if self.profile is not None:
            return self._step_profiled(dt)
        if lazy and self.xchg and self.p2p:
            # the whole step is enqueued by the library (dlpgpu_dev_md_step): no interpreter between the gmax decision and
            # the force kernels
            self.gseq += 1
            reb, prev, list_ms = sr.dev_md_step(self.neigh, dt, self.gseq, self.rseq + 1)
            if prev is not None:
                self.last_out = prev
                self._account()
            if reb:
                self.rseq += 1
                self.rebuilds += 1
                self._refresh_bufs = None
                self.acc["list_ms"] += list_ms; self.acc["list_builds"] += 1
                self.acc["xchg_ms"] += sr.dev_xchg_last_ms()
            self._pending = True
            self.steps += 1
            return self.last_out
        with self.torch.cuda.stream(self.stream):
            sr.dev_vv(1, dt)
        self.publish()                       # before the gmax: the collective orders every rank's publish before any pull
        upd = self.vnl_update()              # host synchronisation: everything enqueued before it has completed
        prev = self.collect() if self._pending else self.last_out
        if upd:
            self.rebuild()
        else:
            self.refresh_halo()
        if lazy:
            with self.torch.cuda.stream(self.stream):
                sr.dev_two_body_forces_async(zero_forces=True)
                sr.dev_vv(2, dt)
            self._pending = True
            self.steps += 1
            return prev
        out = self.forces()
        self.last_out = out
        with self.torch.cuda.stream(self.stream):
            sr.dev_vv(2, dt)
        self.steps += 1
        return out

    def _step_profiled(self, dt):
        """step() with a device synchronisation after every phase (DLP_DD_PROFILE=1): wall time per phase in self.profile."""
        import time
        sr, prof = self.sr, self.profile

        def lap(name, t0):
            self.torch.cuda.synchronize()
            t1 = time.perf_counter()
            prof[name] = prof.get(name, 0.0) + (t1 - t0)
            return t1

        self.torch.cuda.synchronize()
        t = time.perf_counter()
        with self.torch.cuda.stream(self.stream):
            sr.dev_vv(1, dt)
        self.publish()
        t = lap("vv1", t)
        upd = self.vnl_update()
        t = lap("vnl_check+gmax", t)
        if upd and self.xchg:
            self.rseq += 1
            with self.torch.cuda.stream(self.stream):
                sr.dev_xchg_rebuild(self.neigh, self.rseq)
            t = lap("relocate+set_halo (fused)", t)
            with self.torch.cuda.stream(self.stream):
                sr.dev_link_cell_pairs()
            t = lap("link_cell_pairs", t)
            self.rebuilds += 1
        elif upd:
            self.relocate(); t = lap("relocate", t)
            self.set_halo(); t = lap("set_halo", t)
            with self.torch.cuda.stream(self.stream):
                sr.dev_link_cell_pairs()
            t = lap("link_cell_pairs", t)
            self.rebuilds += 1
        else:
            self.refresh_halo(); t = lap("refresh_halo", t)
        out = self.forces(); t = lap("two_body_forces", t)
        self.last_out = out
        with self.torch.cuda.stream(self.stream):
            sr.dev_vv(2, dt)
        t = lap("vv2", t)
        self.steps += 1
        return out

    def gsum(self, out):
        """two_body.F90:729 / drivers.F90:795."""
        return self.t.allreduce_sum(out)

    # ---- SPME reciprocal space over the domains (ewald_spole.F90:244-477 with a replicated grid, csrc/spme.cu)
    def set_spme(self, kdim, nsplines=8):
        self.sr.set_spme(kdim, nsplines)
        self._spme_grid = self.torch.empty(int(kdim[0]) * int(kdim[1]) * int(kdim[2]), dtype=self.torch.float64, device=self.device)

    def spme_forces(self):
        """ewald_spme_forces_coul for this rank's atoms: every rank spreads its charges onto a grid of the whole cell, the grids
        are summed over the ranks (one all-reduce), every rank transforms the whole grid and gathers its own forces; the net
        force is removed with the sum over all ranks.  Adds into the device force arrays; returns this rank's out[16] (the
        ranks' energies, virials and stresses add up under gsum)."""
        g = self._spme_grid
        with self.torch.cuda.stream(self.stream):
            self.sr.dev_spme_spread(g.data_ptr())
        self.t.allreduce_sum_device(g, self.stream)
        with self.torch.cuda.stream(self.stream):
            ftot = self.sr.dev_spme_solve_gather(g.data_ptr())
        ftot = self.t.allreduce_sum(ftot)
        with self.torch.cuda.stream(self.stream):
            return self.sr.dev_spme_finish(self.sys.megatm, ftot, self.world)

    def close(self):
        self.sr.close()
