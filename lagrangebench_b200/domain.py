"""Spatial domain decomposition of one large particle cloud across the GPUs of a box.

The reference is single-device; this is the B200-native analogue for the
1 M-particle clouds of BASELINE.json.  One process per GPU
(``torch.distributed``, NCCL over NVLink):

* the periodic box is cut into ``world`` slabs along one axis; a rank owns the particles whose
  most recent position lies in its slab;
* per rollout step a rank receives the *positions* of its neighbours' boundary particles
  (ghosts, within one cutoff of the slab faces), builds its local neighbor list (open along
  the cut axis, periodic along the others) and keeps the edges whose receiver it owns -- the
  edge latents never move;
* per message-passing step the ghost rows of the node projections ``P`` (1 KB per ghost) are
  exchanged with the two neighbours (``halo_fn`` hook of ``lb200_gns_forward``) -- 10 exchanges
  per rollout step, latency-bound;
* after integration, particles that left the slab migrate to their new owner.

The host logic (ownership, halo selection, migration, exchange order) is device-agnostic and is
covered on CPU with the gloo backend (``tests/test_domain_cpu.py``).
"""

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .case_setup import get_dataset_stats
from .models import gns_cfg, pack_params


class SlabDomain:
    """Ownership and halo geometry of a 1-D slab decomposition of a periodic box."""

    def __init__(self, box, axis, world, rank, halo):
        self.box = [float(b) for b in box]
        self.axis, self.world, self.rank = int(axis), int(world), int(rank)
        self.length = self.box[self.axis]
        self.width = self.length / self.world
        self.lo, self.hi = self.rank * self.width, (self.rank + 1) * self.width
        self.halo = float(halo)
        if self.world > 1 and self.width < 2 * self.halo:
            raise ValueError("slabs thinner than two halos: use fewer ranks")
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world

    def owner(self, coord):
        """Rank owning a particle from its coordinate along the cut axis."""
        return torch.clamp(torch.floor(coord / self.width).to(torch.int64), 0, self.world - 1)

    def halo_masks(self, coord):
        """Owned particles the left / right neighbour needs as ghosts."""
        return coord < self.lo + self.halo, coord >= self.hi - self.halo

    def ghost_shift(self, from_left):
        """Coordinate shift that places a neighbour's particles next to this slab across the
        periodic wrap (rank 0's left neighbour lives at the far end of the box)."""
        if from_left and self.rank == 0:
            return -self.length
        if not from_left and self.rank == self.world - 1:
            return self.length
        return 0.0


def exchange_rows(domain, to_left, to_right, group=None):
    """Send ``to_left`` / ``to_right`` (2-D, same trailing shape and dtype on all ranks) to the
    two neighbours; return ``(from_left, from_right)``.  Row counts are exchanged first.

    Posting order is (left, right) for sends and (from right, from left) for receives so that
    the messages pair up when both neighbours are the same rank (world == 2)."""
    if domain.world == 1:
        return to_left[:0], to_right[:0]
    dev = to_left.device
    counts = torch.tensor([to_left.shape[0], to_right.shape[0]], dtype=torch.int64, device=dev)
    all_counts = torch.empty(domain.world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(domain.world, 2).cpu()
    n_from_right = int(all_counts[domain.right, 0])  # the right neighbour's left-going set
    n_from_left = int(all_counts[domain.left, 1])
    return exchange_rows_sized(domain, to_left, to_right, n_from_left, n_from_right, group)


def exchange_rows_sized(domain, to_left, to_right, n_from_left, n_from_right, group=None, out_left=None,
                        out_right=None):
    """As :func:`exchange_rows` with known receive counts (the per-MP-step exchange of ``P``);
    optionally receives straight into ``out_left`` / ``out_right``."""
    tail = to_left.shape[1:]
    from_right = out_right if out_right is not None else to_left.new_empty((n_from_right,) + tuple(tail))
    from_left = out_left if out_left is not None else to_left.new_empty((n_from_left,) + tuple(tail))
    ops = [dist.P2POp(dist.isend, to_left.contiguous(), domain.left, group),
           dist.P2POp(dist.isend, to_right.contiguous(), domain.right, group),
           dist.P2POp(dist.irecv, from_right, domain.right, group),
           dist.P2POp(dist.irecv, from_left, domain.left, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return from_left, from_right


def halo_sets(domain, coord, group=None):
    """Indices (ascending) of the owned particles the left / right neighbour needs as ghosts, and
    the number of ghosts this rank will receive from each side: ``(send_left, send_right,
    n_from_left, n_from_right)``.  One host synchronisation (the all-gathered counts)."""
    m_left, m_right = domain.halo_masks(coord)
    if domain.world == 1:
        e = torch.empty(0, dtype=torch.int64, device=coord.device)
        return e, e, 0, 0
    counts = torch.stack([m_left.sum(), m_right.sum()]).to(torch.int64)
    all_counts = torch.empty(domain.world * 2, dtype=torch.int64, device=coord.device)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(domain.world, 2).cpu()
    n_left, n_right = int(all_counts[domain.rank, 0]), int(all_counts[domain.rank, 1])
    # stable sort of the negated mask: the selected rows first, in ascending index order -- no
    # second synchronisation (nonzero() would need one per mask)
    send_left = torch.argsort((~m_left).to(torch.uint8), stable=True)[:n_left]
    send_right = torch.argsort((~m_right).to(torch.uint8), stable=True)[:n_right]
    return send_left, send_right, int(all_counts[domain.left, 1]), int(all_counts[domain.right, 0])


def step_counts(domain, coord, group=None):
    """ONE collective + ONE host synchronisation for everything a step needs to know about the other
    ranks: the migration send-count matrix and every rank's (left, right) halo counts, both computed
    from the same coordinates.  Returns ``(matrix (world, world), halo_counts (world, 2), order_left,
    order_right)`` -- ``order_x[:halo_counts[rank, k]]`` are the ascending indices of the rows the
    left / right neighbour needs; the halo part is only valid when the matrix is diagonal (nobody
    changes owner)."""
    w = domain.world
    m_left, m_right = domain.halo_masks(coord)
    payload = torch.cat([torch.bincount(domain.owner(coord), minlength=w).to(torch.int64),
                         torch.stack([m_left.sum(), m_right.sum()]).to(torch.int64)])
    allp = torch.empty(w * (w + 2), dtype=torch.int64, device=coord.device)
    dist.all_gather_into_tensor(allp, payload, group=group)
    # enqueued BEFORE the host read, so that the device is not idle while the host launches them:
    # selected rows first, in ascending index order (the caller slices [:count])
    order_left = torch.argsort((~m_left).to(torch.uint8), stable=True)
    order_right = torch.argsort((~m_right).to(torch.uint8), stable=True)
    allp = allp.view(w, w + 2).cpu()
    return allp[:, :w], allp[:, w:], order_left, order_right


def migrate(domain, coord, tensors, group=None, matrix=None):
    """Move rows to the rank that now owns them.  ``tensors``: list of tensors with the same
    leading dimension; returns the list with departed rows removed and arrivals appended
    (stayers keep their relative order).  One host synchronisation (the all-gathered counts),
    none when the caller already holds the ``(world, world)`` send-count ``matrix`` (host tensor)."""
    if domain.world == 1:
        return tensors
    dest = domain.owner(coord)
    dev = coord.device
    if matrix is None:
        send_counts = torch.bincount(dest, minlength=domain.world)
        matrix = torch.empty(domain.world * domain.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(matrix, send_counts.to(torch.int64), group=group)
        matrix = matrix.view(domain.world, domain.world).cpu()
    row = [int(v) for v in matrix[domain.rank]]
    if sum(row) == row[domain.rank] and int(matrix[:, domain.rank].sum()) == row[domain.rank]:
        return tensors  # nobody leaves, nobody arrives
    order = torch.argsort(dest, stable=True)  # rows grouped by destination rank
    starts = [0]
    for v in row:
        starts.append(starts[-1] + v)
    out, ops, recv_bufs, keep = [], [], [], []
    for t in tensors:
        grouped = t.index_select(0, order)
        keep.append(grouped[starts[domain.rank]:starts[domain.rank + 1]])
        bufs = []
        for r in range(domain.world):
            if r == domain.rank:
                continue
            n_send, n_recv = row[r], int(matrix[r, domain.rank])
            if n_send:
                ops.append(dist.P2POp(dist.isend, grouped[starts[r]:starts[r + 1]].contiguous(), r, group))
            if n_recv:
                b = t.new_empty((n_recv,) + tuple(t.shape[1:]))
                ops.append(dist.P2POp(dist.irecv, b, r, group))
                bufs.append(b)
        recv_bufs.append(bufs)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for k, bufs in zip(keep, recv_bufs):
        out.append(torch.cat([k] + bufs, dim=0) if bufs else k)
    return out


class DistributedRollout:
    """Rollout of ONE cloud sharded over the ranks of ``group`` (GNS, periodic box, no
    kinematic particles -- the RPF-3D shape of BASELINE.json's 1 M-particle configuration)."""

    def __init__(self, box, metadata, params, num_mp_steps, force=None, axis=None, dtype=torch.float32,
                 multiplier=1.25, input_seq_length=6, group=None, noise_std=3.0e-4, timing=False):
        _cabi.require_cuda()
        self.lib = _cabi.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.box = [float(b) for b in box]
        self.dim = len(self.box)
        if not all(metadata["periodic_boundary_conditions"]):
            raise NotImplementedError("the slab decomposition covers fully periodic boxes")
        self.axis = int(np.argmax(self.box)) if axis is None else int(axis)
        self.tdtype = dtype
        npd = np.float64 if dtype == torch.float64 else np.float32
        self.radius = float(npd(metadata["default_connectivity_radius"]))
        self.domain = SlabDomain(self.box, self.axis, self.world, self.rank, self.radius)
        self.stats = get_dataset_stats(metadata, False, noise_std, npd)
        self.multiplier = float(multiplier)
        self.isl = int(input_seq_length)
        self.force = force
        self.packed = pack_params(params, num_mp_steps, self.dim)
        self.num_mp_steps = num_mp_steps
        # periodic along every axis but the cut one (a single rank keeps the full periodic box)
        full = (1 << self.dim) - 1
        self.pmask = full if self.world == 1 else full & ~(1 << self.axis)
        self._cap = None  # (n_cap, e_cap, cell_cap) the buffers are sized for
        self._halo_cb = _cabi.HALO_FN(self._halo_exchange)
        self.n_reallocations = 0
        self.n_migrations = 0
        self.edges_last = 0
        self._halo_error = None
        # optional per-phase CUDA-event timing of step(): phase name -> [ms summed, count]
        self.timing = bool(timing)
        self._marks = []
        self.phase_ms = {}

    # ------------------------------------------------------------------ state
    def scatter(self, positions, particle_type):
        """Keep this rank's share of a globally known initial state ``(N, T, d)``."""
        pos = torch.as_tensor(positions)
        owner = self.domain.owner(pos[:, self.isl - 1, self.axis].to(torch.float64))
        mine = (owner == self.rank).nonzero().squeeze(1)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.window = pos[mine, :self.isl].to(dev, self.tdtype).contiguous()
        self.ptype = torch.as_tensor(particle_type)[mine].to(dev, torch.int32).contiguous()
        self.gid = mine.to(dev)
        return self

    # ------------------------------------------------------------------ buffers
    def _ensure(self, n_loc, n_edges_hint=None):
        cap = self._cap
        if cap is not None and n_loc <= cap[0] and (n_edges_hint is None or n_edges_hint <= cap[1]):
            return
        lib, dev = self.lib, self.window.device
        n_cap = int(n_loc * 1.15) + 1024
        g = _cabi.Grid()
        box = (C.c_double * 3)(*(self.box + [1.0] * (3 - self.dim)))
        _cabi.check(lib.lb200_grid_init(C.byref(g), n_cap, self.dim, int(self.tdtype == torch.float64), self.pmask,
                                        box, self.radius))
        if not g.use_cells:
            raise NotImplementedError("decomposed clouds use the cell list")
        self.grid = g
        self.nbr_scratch = torch.empty(lib.lb200_nbr_scratch_bytes(C.byref(g)), dtype=torch.uint8, device=dev)
        self.stats_dev = torch.zeros(4, dtype=torch.int32, device=dev)
        e_cap = cap[1] if cap is not None else 0
        cell_cap = cap[2] if cap is not None else 0
        if n_edges_hint is not None:
            e_cap = max(e_cap, int(n_edges_hint * self.multiplier))
        self._cap = (n_cap, e_cap, cell_cap)
        if e_cap:
            self._alloc_edges(e_cap)

    def _alloc_edges(self, e_cap):
        lib, dev = self.lib, self.window.device
        n_cap = self._cap[0]
        i32 = dict(dtype=torch.int32, device=dev)
        self.idx = torch.empty((2, e_cap), **i32)
        self.rowptr = torch.empty(n_cap + 1, **i32)
        self.perm, self.snd, self.rcv = (torch.empty(e_cap, **i32) for _ in range(3))
        self.csr_scratch = torch.empty(lib.lb200_csr_scratch_bytes(n_cap, e_cap), dtype=torch.uint8, device=dev)
        self.edge_feat = torch.empty((e_cap, 4), dtype=torch.float32, device=dev)
        self._cap = (n_cap, e_cap, self._cap[2])

    def _gns_buffers(self, n_loc, e_cap):
        key = (n_loc, e_cap)
        if getattr(self, "_gns_key", None) != key:
            lib = self.lib
            nbytes = lib.lb200_gns_scratch_bytes(n_loc, e_cap)
            if getattr(self, "gns_scratch", None) is None or self.gns_scratch.numel() < nbytes:
                self.gns_scratch = torch.empty(int(nbytes * 1.1), dtype=torch.uint8, device=self.window.device)
            off_p = C.c_int64()
            lib.lb200_gns_scratch_layout(n_loc, e_cap, None, C.byref(off_p), None, None)
            self.P = self.gns_scratch[off_p.value:off_p.value + n_loc * 256 * 4].view(torch.float32).view(n_loc, 256)
            self._gns_key = key
        return self.gns_scratch

    # ------------------------------------------------------------------ halo of P (called from C)
    def _halo_exchange(self, _ctx, _mp_step):
        try:  # ctypes swallows exceptions raised inside callbacks: keep it and re-raise after the call
            n_own, nl, nr = self.n_own, self.n_ghost_left, self.n_ghost_right
            p = self.P
            exchange_rows_sized(self.domain, p.index_select(0, self.send_left), p.index_select(0, self.send_right),
                                nl, nr, self.group, out_left=p[n_own:n_own + nl],
                                out_right=p[n_own + nl:n_own + nl + nr])
            self.halo_bytes += (self.send_left.numel() + self.send_right.numel()) * 1024
        except BaseException as exc:  # noqa: BLE001
            self._halo_error = exc

    # ------------------------------------------------------------------ phase timing
    def _mark(self, name):
        if self.timing:
            import time

            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev, time.perf_counter()))

    def read_phase_ms(self):
        """Mean milliseconds per step of each phase since the last call (synchronises)."""
        torch.cuda.synchronize()
        acc = {}
        for (n0, e0, c0), (n1, e1, c1) in zip(self._marks[:-1], self._marks[1:]):
            if n1 == "begin":
                continue
            t = acc.setdefault(n1, [0.0, 0, 0.0])
            t[0] += e0.elapsed_time(e1)
            t[1] += 1
            t[2] += (c1 - c0) * 1e3
        self._marks = []
        steps = max((v[1] for v in acc.values()), default=1)
        self.phase_host_ms = {k: v[2] / steps for k, v in acc.items()}  # host wall time spent enqueuing the phase
        return {k: v[0] / steps for k, v in acc.items()}

    # ------------------------------------------------------------------ one rollout step
    def step(self):
        """Two host synchronisations per step: one all-gather carrying the migration matrix and the
        ghost counts, and the neighbor-list overflow flag (re-allocate and retry,
        evaluate/rollout.py:135-151 -- local, before any neighbour is involved)."""
        lib, dom, dev = self.lib, self.domain, self.window.device
        st = _cabi.stream()
        n_own = self.window.shape[0]
        self._mark("begin")
        pos_own = self.window[:, -1].contiguous()
        self.halo_bytes = 0
        if self.world > 1:
            # one collective: who changed owner during the last integrate, and the halo counts
            matrix, halo_counts, order_left, order_right = step_counts(dom, pos_own[:, self.axis], self.group)
            if int(matrix.sum()) != int(matrix.diagonal().sum()):  # somebody migrates (rare): move, then recount
                n = self.window.shape[0]
                w2, pt, gid = migrate(dom, self.window[:, -1, self.axis],
                                      [self.window.view(n, -1), self.ptype, self.gid], self.group, matrix=matrix)
                self.window = w2.view(-1, self.isl, self.dim).contiguous()
                self.ptype, self.gid = pt.contiguous(), gid.contiguous()
                n_own = self.window.shape[0]
                pos_own = self.window[:, -1].contiguous()
                self.n_migrations += 1
                self.send_left, self.send_right, n_fl, n_fr = halo_sets(dom, pos_own[:, self.axis], self.group)
            else:
                n_left, n_right = int(halo_counts[dom.rank, 0]), int(halo_counts[dom.rank, 1])
                self.send_left, self.send_right = order_left[:n_left], order_right[:n_right]
                n_fl, n_fr = int(halo_counts[dom.left, 1]), int(halo_counts[dom.right, 0])
            self._mark("owner + ghost counts (one all-gather, host sync)")
            pos_loc = torch.empty((n_own + n_fl + n_fr, self.dim), dtype=pos_own.dtype, device=dev)
            pos_loc[:n_own] = pos_own
            exchange_rows_sized(dom, pos_own.index_select(0, self.send_left), pos_own.index_select(0, self.send_right),
                                n_fl, n_fr, self.group, out_left=pos_loc[n_own:n_own + n_fl],
                                out_right=pos_loc[n_own + n_fl:])
            # place the neighbours' particles next to this slab across the periodic wrap
            if n_fl and dom.ghost_shift(True) != 0.0:
                pos_loc[n_own:n_own + n_fl, self.axis] += dom.ghost_shift(True)
            if n_fr and dom.ghost_shift(False) != 0.0:
                pos_loc[n_own + n_fl:, self.axis] += dom.ghost_shift(False)
            self.n_ghost_left, self.n_ghost_right = n_fl, n_fr
        else:
            pos_loc = pos_own
            self.n_ghost_left = self.n_ghost_right = 0
        self.n_own = n_own
        n_loc = pos_loc.shape[0]
        self._ensure(n_loc)
        self._mark("ghost positions")
        while True:
            n_cap, e_cap, cell_cap = self._cap
            self.grid.n = n_loc
            if e_cap == 0:  # first use, or after an overflow: size the list from this cloud (host read)
                _cabi.check(lib.lb200_nbr_build(C.byref(self.grid), _cabi.ptr(pos_loc), 0, None, 0,
                                                _cabi.ptr(self.stats_dev), _cabi.ptr(self.nbr_scratch),
                                                self.nbr_scratch.numel(), st))
                n_edges, max_occ = self.stats_dev[:2].tolist()
                self._cap = (n_cap, 0, max(1, int(max_occ * self.multiplier)))
                self._alloc_edges(max(1, int(n_edges * self.multiplier)))
                self.stats_dev.zero_()
                continue
            _cabi.check(lib.lb200_nbr_build(C.byref(self.grid), _cabi.ptr(pos_loc), cell_cap, _cabi.ptr(self.idx),
                                            e_cap, _cabi.ptr(self.stats_dev), _cabi.ptr(self.nbr_scratch),
                                            self.nbr_scratch.numel(), st))
            n_edges, _, overflow, _ = self.stats_dev.tolist()  # host read (rollout.py:135): the retry is local,
            if overflow:                                        # the neighbours are not involved yet
                self.n_reallocations += 1
                self._cap = (n_cap, 0, 0)
                self.stats_dev.zero_()
                continue
            break
        _cabi.check(lib.lb200_csr_build(_cabi.ptr(self.idx), n_loc, e_cap, _cabi.ptr(self.rowptr), _cabi.ptr(self.perm),
                                        _cabi.ptr(self.snd), _cabi.ptr(self.rcv), _cabi.ptr(self.csr_scratch),
                                        self.csr_scratch.numel(), st))
        self._mark("neighbor list + csr")
        # ---- features: nodes from the owned window, edges from the local (owned + ghost) positions
        fc = self._feature_cfg(n_own, self.isl)
        node_feat = torch.empty((n_own, fc.node_stride), dtype=torch.float32, device=dev)
        _cabi.check(lib.lb200_features(C.byref(fc), _cabi.ptr(self.window), None, None, 0, _cabi.ptr(node_feat), None, st))
        fe = self._feature_cfg(n_loc, 1)
        _cabi.check(lib.lb200_features(C.byref(fe), _cabi.ptr(pos_loc), None, _cabi.ptr(self.idx), e_cap, None,
                                       _cabi.ptr(self.edge_feat), st))
        self._mark("features")
        # ---- forward with the per-MP-step halo exchange of P
        scratch = self._gns_buffers(n_loc, e_cap)
        cfg = gns_cfg(self.packed, n_loc, e_cap, fc.node_stride, fc.node_stride)
        if self.world > 1:
            cfg.n_owned = n_own
            cfg.halo_fn = C.cast(self._halo_cb, C.c_void_p).value
        out = torch.empty((n_own, self.dim), dtype=torch.float32, device=dev)
        _cabi.check(lib.lb200_gns_forward(C.byref(cfg), _cabi.ptr(self.packed.blob), _cabi.ptr(node_feat),
                                          _cabi.ptr(self.edge_feat), _cabi.ptr(self.ptype), _cabi.ptr(self.rowptr),
                                          _cabi.ptr(self.perm), _cabi.ptr(self.snd), _cabi.ptr(self.rcv), _cabi.ptr(out),
                                          _cabi.ptr(scratch), scratch.numel(), st))
        if self._halo_error is not None:
            err, self._halo_error = self._halo_error, None
            raise err
        self._mark("forward (10 x [halo of P, message, node])")
        self.edges_last = n_edges
        # ---- integrate the owned particles (periodic shift along every axis), then migrate
        ic = _cabi.IntegrateCfg()
        ic.n, ic.dim, ic.t_window = n_own, self.dim, self.isl
        ic.pos_f64, ic.periodic, ic.out_mode = int(self.tdtype == torch.float64), (1 << self.dim) - 1, 0
        ic.box = _cabi.vec3(self.box, 1.0)
        ic.mean = _cabi.vec3(self.stats["acceleration"]["mean"])
        ic.std = _cabi.vec3(self.stats["acceleration"]["std"], 1.0)
        _cabi.check(lib.lb200_integrate(C.byref(ic), _cabi.ptr(out), _cabi.ptr(self.window), _cabi.ptr(self.ptype), None,
                                        None, None, st))
        self._mark("integrate")
        # particles that left the slab move to their new owner at the start of the next step (or in
        # settle()): its single collective carries the migration counts together with the halo counts

    def settle(self):
        """Hand particles that left their slab during the last step to their new owner (a step does
        this lazily at its start)."""
        if self.world > 1:
            n = self.window.shape[0]
            w2, pt, gid = migrate(self.domain, self.window[:, -1, self.axis],
                                  [self.window.view(n, -1), self.ptype, self.gid], self.group)
            self.window = w2.view(-1, self.isl, self.dim).contiguous()
            self.ptype, self.gid = pt.contiguous(), gid.contiguous()

    def _feature_cfg(self, n, t_window):
        fc = _cabi.FeatureCfg()
        fc.n, fc.dim, fc.t_window = n, self.dim, t_window
        # velocities of the owned window wrap along every axis; edge displacements of the local
        # cloud are open along the cut axis (ghosts were shifted next to the slab)
        fc.pos_f64 = int(self.tdtype == torch.float64)
        fc.periodic = (1 << self.dim) - 1 if t_window > 1 else self.pmask
        fc.box = _cabi.vec3(self.box, 1.0)
        fc.r_cutoff = self.radius
        fc.vel_mean = _cabi.vec3(self.stats["velocity"]["mean"])
        fc.vel_std = _cabi.vec3(self.stats["velocity"]["std"], 1.0)
        fc.magnitude_features = 0
        fc.bound_features = 0
        if self.force is not None:
            f = self.force
            fc.force_mode, fc.force_axis, fc.force_threshold = 1, f.axis, min(f.threshold, 1e300)
            fc.force_lo, fc.force_hi = _cabi.vec3(f.lo), _cabi.vec3(f.hi)
        fc.node_stride = 0
        fc.node_stride = self.lib.lb200_node_feature_width(C.byref(fc)) if t_window > 1 else 0
        return fc

    # ------------------------------------------------------------------ results
    def gather_positions(self, n_total):
        """All ranks' most recent positions ordered by global particle id (for checks)."""
        dev = self.window.device
        pos = torch.zeros((n_total, self.dim), dtype=self.tdtype, device=dev)
        pos[self.gid] = self.window[:, -1]
        if self.world > 1:
            dist.all_reduce(pos, group=self.group)
        return pos
