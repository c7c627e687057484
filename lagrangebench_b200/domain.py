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
* per message-passing step the sender projections of the boundary rows (512 B per ghost) are
  stored by the node kernel's epilogue straight into the neighbours' peer-mapped arrays
  (``csrc/gns_tc.cu``), followed by one signal/wait kernel (``csrc/peer.cu``) -- no collective
  library and no host on the data path, so whole steps replay from a CUDA graph;
* ghost sets are chosen with a margin and stay fixed until a particle has drifted half the margin
  along the cut axis (device-side check, OR-ed over all ranks together with the neighbor-list
  overflow bits before a step's integrate takes effect); then particles that left their slab
  migrate to their new owner and the sets are chosen again.

The host logic (ownership, halo selection, migration, exchange order) is device-agnostic and is
covered on CPU with the gloo backend (``tests/test_domain_cpu.py``).
"""

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .case_setup import get_dataset_stats, periodic_mask
from .models import gns_cfg, pack_params


class SlabDomain:
    """Ownership and halo geometry of a 1-D slab decomposition of a periodic box."""

    def __init__(self, box, axis, world, rank, halo, periodic=True):
        self.box = [float(b) for b in box]
        self.axis, self.world, self.rank = int(axis), int(world), int(rank)
        self.periodic = bool(periodic)  # the cut axis wraps: rank 0 and rank world-1 are neighbours
        self.length = self.box[self.axis]
        self.width = self.length / self.world
        self.lo, self.hi = self.rank * self.width, (self.rank + 1) * self.width
        self.halo = float(halo)
        if self.world > 1 and self.width < 2 * self.halo:
            raise ValueError("slabs thinner than two halos: use fewer ranks")
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world
        self.has_left = self.world > 1 and (self.periodic or self.rank > 0)
        self.has_right = self.world > 1 and (self.periodic or self.rank < self.world - 1)

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


class _DeviceView:
    """``__cuda_array_interface__`` wrapper: lets torch alias raw device memory (a peer heap)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _is_gloo(group=None):
    return dist.is_initialized() and dist.get_backend(group) == "gloo"


def _p2p_rows(domain, to_left, to_right, n_from_left, n_from_right, group=None):
    """Neighbour exchange of rows at ghost-selection time (rare, host-orchestrated): NCCL on the
    device, or staged through the host under gloo."""
    dev = to_left.device
    if _is_gloo(group) and to_left.is_cuda:
        fl, fr = _p2p_rows(domain, to_left.cpu(), to_right.cpu(), n_from_left, n_from_right, group)
        return fl.to(dev), fr.to(dev)
    tail = tuple(to_left.shape[1:])
    from_left = to_left.new_empty((n_from_left,) + tail)
    from_right = to_left.new_empty((n_from_right,) + tail)
    ops = []
    if domain.has_left:
        ops.append(dist.P2POp(dist.isend, to_left.contiguous(), domain.left, group))
    if domain.has_right:
        ops.append(dist.P2POp(dist.isend, to_right.contiguous(), domain.right, group))
    if domain.has_right:
        ops.append(dist.P2POp(dist.irecv, from_right, domain.right, group))
    if domain.has_left:
        ops.append(dist.P2POp(dist.irecv, from_left, domain.left, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return from_left, from_right


def select_ghosts(domain, coord, counts_all=None, group=None):
    """Ghost sets of one selection: ``coord`` = cut-axis coordinate of the owned rows.  Returns a dict
    with the ascending row indices the two neighbours hold as ghosts (``send_left`` /
    ``send_right``), this rank's ghost counts and the first row of its blocks inside the
    neighbours' clouds.  One collective (all ranks' ``[n_owned, n_send_left, n_send_right]``)."""
    n_own = int(coord.shape[0])
    m_left, m_right = domain.halo_masks(coord)
    send_left = m_left.nonzero().squeeze(1) if domain.has_left else coord.new_empty(0, dtype=torch.int64)
    send_right = m_right.nonzero().squeeze(1) if domain.has_right else coord.new_empty(0, dtype=torch.int64)
    mine = [n_own, int(send_left.numel()), int(send_right.numel())]
    if counts_all is None:
        if domain.world > 1:
            counts_all = [None] * domain.world
            dist.all_gather_object(counts_all, mine, group=group)
        else:
            counts_all = [mine]
    own = [c[0] for c in counts_all]
    n_sl = [c[1] for c in counts_all]
    n_sr = [c[2] for c in counts_all]
    w = domain.world

    def has_left(r):
        return w > 1 and (domain.periodic or r > 0)

    def has_right(r):
        return w > 1 and (domain.periodic or r < w - 1)

    def ghosts_left(r):  # rows rank r receives from its left neighbour: that neighbour's right-going set
        return n_sr[(r - 1) % w] if has_left(r) else 0

    def ghosts_right(r):
        return n_sl[(r + 1) % w] if has_right(r) else 0

    n_loc_all = [own[r] + ghosts_left(r) + ghosts_right(r) for r in range(w)]
    return {
        "send_left": send_left, "send_right": send_right,
        "n_ghost_left": ghosts_left(domain.rank), "n_ghost_right": ghosts_right(domain.rank),
        # my left-going rows are the left neighbour's FROM-RIGHT block, behind its own rows and its from-left block
        "dst_row_left": own[domain.left] + ghosts_left(domain.left) if domain.has_left else 0,
        "dst_row_right": own[domain.right] if domain.has_right else 0,
        "n_loc_max": max(n_loc_all), "counts_all": counts_all,
    }


class DistributedRollout:
    """Rollout of ONE cloud sharded over the ranks of ``group`` (GNS; periodic or walled boxes,
    kinematic particles follow their owner) -- the device-resident decomposed step loop of
    ``lb200_rollout_steps`` with an ``lb200_shard``.

    ``halo_margin``: the ghost zone is ``(1 + halo_margin) * cutoff`` wide; ghost sets are kept
    until a particle has moved ``halo_margin * cutoff / 2`` along the cut axis."""

    def __init__(self, box, metadata, params, num_mp_steps, force=None, axis=None, dtype=torch.float32,
                 multiplier=1.25, input_seq_length=6, group=None, noise_std=3.0e-4, timing=False,
                 halo_margin=0.25, steps_per_sync=32):
        _cabi.require_cuda()
        self.lib = _cabi.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _cabi.MAX_RANKS:
            raise ValueError(f"at most {_cabi.MAX_RANKS} ranks")
        self.box = [float(b) for b in box]
        self.dim = len(self.box)
        self.metadata = metadata
        pbc = list(metadata["periodic_boundary_conditions"])
        self.periodic = bool(np.array(pbc).any())  # case.py:104: all directions or none
        self.bound_features = not any(pbc)         # features.py:87
        self.axis = int(np.argmax(self.box)) if axis is None else int(axis)
        self.tdtype = dtype
        npd = np.float64 if dtype == torch.float64 else np.float32
        self.npd = npd
        self.radius = float(npd(metadata["default_connectivity_radius"]))
        self.halo_margin = float(halo_margin)
        self.domain = SlabDomain(self.box, self.axis, self.world, self.rank, self.radius * (1.0 + self.halo_margin),
                                 periodic=self.periodic)
        self.stats = get_dataset_stats(metadata, False, noise_std, npd)
        self.multiplier = float(multiplier)
        self.isl = int(input_seq_length)
        self.force = force
        self.packed = pack_params(params, num_mp_steps, self.dim)
        self.num_mp_steps = num_mp_steps
        # the local cloud is open along the cut axis (ghosts sit next to the slab); a single rank keeps the box
        full = periodic_mask(self.periodic, self.dim)
        self.pmask = full if self.world == 1 else full & ~(1 << self.axis)
        self.full_mask = full
        self.steps_per_sync = int(steps_per_sync)
        self.n_reallocations = 0
        self.n_migrations = 0
        self.n_selections = 0
        self.edges_last = 0
        self.steps_done = 0
        self.halo_bytes = 0
        self.timing = bool(timing)
        self.future = None
        self._need_select = True
        self._heap = None       # (n_cap, bytes, local pointer, {rank: mapped pointer})
        self._caps = (0, 0)     # (e_cap, cell_cap)
        self._scratch = None
        self._status = None
        self._stream = None
        self._keep = {}

    # ------------------------------------------------------------------ state
    def scatter(self, positions, particle_type):
        """Keep this rank's share of a globally known trajectory ``(N, T, d)``: the first
        ``input_seq_length`` frames are the window, later frames the positions kinematic
        particles are overridden with (``evaluate/rollout.py:64-69``)."""
        pos = torch.as_tensor(positions)
        owner = self.domain.owner(pos[:, self.isl - 1, self.axis].to(torch.float64))
        mine = (owner == self.rank).nonzero().squeeze(1)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.window = pos[mine, :self.isl].to(dev, self.tdtype).contiguous()
        self.ptype = torch.as_tensor(particle_type)[mine].to(dev, torch.int32).contiguous()
        self.gid = mine.to(dev)
        self.future = pos[mine, self.isl:].to(dev, self.tdtype).contiguous() if pos.shape[1] > self.isl else None
        self._need_select = True
        self.steps_done = 0
        return self

    # ------------------------------------------------------------------ peer heap
    def _ensure_heap(self, n_loc_max):
        """Same-size heap on every rank, mapped into every process by CUDA IPC (collective)."""
        lib = self.lib
        if self._heap is not None and n_loc_max <= self._heap[0]:
            return
        n_cap = int(n_loc_max * 1.1) + 1024
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody stores into a heap that is about to go away
        self._free_heap()
        nbytes = lib.lb200_peer_heap_layout(n_cap, None, None, None)
        local = C.c_void_p()
        handle = C.create_string_buffer(64)
        _cabi.check(lib.lb200_peer_heap_create(nbytes, C.byref(local), handle))
        mapped = {}
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=self.group)
            for r, h in enumerate(handles):
                if r != self.rank:
                    ptr = C.c_void_p()
                    _cabi.check(lib.lb200_peer_heap_open(h, C.byref(ptr)))
                    mapped[r] = ptr.value
            dist.barrier(group=self.group)
        self._heap = (n_cap, nbytes, local.value, mapped)

    def _free_heap(self):
        if self._heap is None:
            return
        _, _, local, mapped = self._heap
        for ptr in mapped.values():
            self.lib.lb200_peer_heap_close(C.c_void_p(ptr))
        self.lib.lb200_peer_heap_destroy(C.c_void_p(local))
        self._heap = None

    def close(self):
        """Release the peer heap (collective when world > 1)."""
        if self._heap is not None:
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)
            self._free_heap()

    def _pos_local(self, n_rows):
        """This rank's cloud (owned + ghost positions) as the last step left it in the heap."""
        n_cap, _, local, _ = self._heap
        off = C.c_int64()
        self.lib.lb200_peer_heap_layout(n_cap, C.byref(off), None, None)
        ts = "<f8" if self.tdtype == torch.float64 else "<f4"
        return torch.as_tensor(_DeviceView(local + off.value, (n_rows, self.dim), ts), device=self.window.device)

    # ------------------------------------------------------------------ ghost selection (host, rare)
    def _select(self):
        dom, dev = self.domain, self.window.device
        # particles that left their slab go to their new owner (one collective; P2P only if somebody moves)
        if self.world > 1:
            n = self.window.shape[0]
            tensors = [self.window.view(n, -1), self.ptype, self.gid]
            if self.future is not None:
                tensors.append(self.future.view(n, -1))
            coord = self.window[:, -1, self.axis]
            if _is_gloo(self.group):
                staged = [t.cpu() for t in tensors]
                moved = migrate(dom, coord.cpu(), staged, self.group)
                changed = moved is not staged
                moved = [t.to(dev) for t in moved] if changed else tensors
            else:
                moved = migrate(dom, coord, tensors, self.group)
                changed = moved is not tensors
            self.n_migrations += int(changed)
            self.window = moved[0].view(-1, self.isl, self.dim).contiguous()
            self.ptype, self.gid = moved[1].contiguous(), moved[2].contiguous()
            if self.future is not None:
                self.future = moved[3].view(self.window.shape[0], -1, self.dim).contiguous()
        n_own = self.window.shape[0]
        coord = self.window[:, -1, self.axis].contiguous()
        sel = select_ghosts(dom, coord, group=self.group)
        self._ensure_heap(sel["n_loc_max"])
        i32 = dict(dtype=torch.int32, device=dev)
        push_left = torch.full((n_own,), -1, **i32)
        push_right = torch.full((n_own,), -1, **i32)
        sl, sr = sel["send_left"], sel["send_right"]
        push_left[sl] = torch.arange(sl.numel(), **i32)
        push_right[sr] = torch.arange(sr.numel(), **i32)
        n_fl, n_fr = sel["n_ghost_left"], sel["n_ghost_right"]
        n_loc = n_own + n_fl + n_fr
        # the first local cloud comes over the host-orchestrated path (sizes the neighbor capacities)
        pos_own = self.window[:, -1].contiguous()
        pos_loc = torch.empty((n_loc, self.dim), dtype=self.tdtype, device=dev)
        pos_loc[:n_own] = pos_own
        length = self.box[self.axis]
        shift_left = length if (self.periodic and self.rank == 0) else 0.0               # added to rows sent left
        shift_right = -length if (self.periodic and self.rank == self.world - 1) else 0.0
        if self.world > 1:
            to_left, to_right = pos_own.index_select(0, sl).clone(), pos_own.index_select(0, sr).clone()
            to_left[:, self.axis] += shift_left
            to_right[:, self.axis] += shift_right
            fl, fr = _p2p_rows(dom, to_left, to_right, n_fl, n_fr, self.group)
            pos_loc[n_own:n_own + n_fl] = fl
            pos_loc[n_own + n_fl:] = fr
        sh = _cabi.Shard()
        sh.rank, sh.world = self.rank, self.world
        sh.has_left, sh.has_right = int(dom.has_left), int(dom.has_right)
        sh.n_owned, sh.n_ghost_left, sh.n_ghost_right = n_own, n_fl, n_fr
        sh.n_send_left, sh.n_send_right = int(sl.numel()), int(sr.numel())
        sh.push_left, sh.push_right = push_left.data_ptr(), push_right.data_ptr()
        sh.dst_row_left, sh.dst_row_right = sel["dst_row_left"], sel["dst_row_right"]
        sh.axis, sh.n_cap = self.axis, self._heap[0]
        sh.shift_left, sh.shift_right = shift_left, shift_right
        sh.axis_length = length if self.periodic else 0.0
        sh.ref_coord = coord.data_ptr()
        sh.drift_limit = 0.5 * self.halo_margin * self.radius if self.world > 1 else 1e300
        _, _, local, mapped = self._heap
        sh.heap = local
        sh.heap_left = mapped.get(dom.left) if dom.has_left else None
        sh.heap_right = mapped.get(dom.right) if dom.has_right else None
        for r in range(self.world):
            sh.heap_all[r] = local if r == self.rank else mapped[r]
        self.shard = sh
        self._keep = {"push_left": push_left, "push_right": push_right, "coord": coord}  # the struct holds raw pointers
        self.n_own, self.n_ghost_left, self.n_ghost_right = n_own, n_fl, n_fr
        self.halo_rows = int(sl.numel() + sr.numel())
        self._size(pos_loc)
        if self.future is not None:
            self.targets = self.future.permute(1, 0, 2).contiguous()  # (T, n_own, d)
        else:
            self.targets = None
        self._need_select = False
        self.n_selections += 1

    def _size(self, pos_loc):
        """Neighbor capacities from a count-only search of the local cloud (host read, like ``allocate``)."""
        lib, dev = self.lib, self.window.device
        n_loc = pos_loc.shape[0]
        g = _cabi.Grid()
        box = (C.c_double * 3)(*(self.box + [1.0] * (3 - self.dim)))
        _cabi.check(lib.lb200_grid_init(C.byref(g), n_loc, self.dim, int(self.tdtype == torch.float64), self.pmask,
                                        box, self.radius))
        if not g.use_cells:
            raise NotImplementedError("decomposed clouds use the cell list")
        scratch = torch.empty(lib.lb200_nbr_scratch_bytes(C.byref(g)), dtype=torch.uint8, device=dev)
        stats = torch.zeros(4, dtype=torch.int32, device=dev)
        _cabi.check(lib.lb200_nbr_build(C.byref(g), _cabi.ptr(pos_loc.contiguous()), 0, None, 0, _cabi.ptr(stats),
                                        _cabi.ptr(scratch), scratch.numel(), _cabi.stream()))
        n_edges, max_occ = stats[:2].tolist()
        e_cap = max(self._caps[0], max(1, int(n_edges * self.multiplier)))
        cell_cap = max(self._caps[1], max(1, int(max_occ * self.multiplier)))
        self._caps = (e_cap, cell_cap)
        self.grid = g

    # ------------------------------------------------------------------ configuration of one call
    def _feature_cfg(self, n):
        fc = _cabi.FeatureCfg()
        fc.n, fc.dim, fc.t_window = n, self.dim, self.isl
        fc.pos_f64 = int(self.tdtype == torch.float64)
        fc.periodic = self.full_mask  # velocities of the owned window wrap along every periodic axis
        fc.box = _cabi.vec3(self.box, 1.0)
        fc.r_cutoff = self.radius
        fc.vel_mean = _cabi.vec3(self.stats["velocity"]["mean"])
        fc.vel_std = _cabi.vec3(self.stats["velocity"]["std"], 1.0)
        fc.magnitude_features = 0
        fc.bound_features = int(self.bound_features)
        bounds = np.asarray(self.metadata["bounds"], dtype=self.npd)
        fc.bounds_lo, fc.bounds_hi = _cabi.vec3(bounds[:, 0]), _cabi.vec3(bounds[:, 1])
        if self.force is not None:
            f = self.force
            fc.force_mode, fc.force_axis, fc.force_threshold = 1, f.axis, min(f.threshold, 1e300)
            fc.force_lo, fc.force_hi = _cabi.vec3(f.lo), _cabi.vec3(f.hi)
        fc.node_stride = 0
        fc.node_stride = self.lib.lb200_node_feature_width(C.byref(fc))
        return fc

    def _configure(self):
        lib = self.lib
        n_own = self.n_own
        n_loc = n_own + self.n_ghost_left + self.n_ghost_right
        e_cap, cell_cap = self._caps
        cfg = _cabi.RolloutCfg()
        cfg.grid = self.grid
        cfg.grid.n = cfg.grid.n_valid = n_loc
        cfg.feat = self._feature_cfg(n_own)
        cfg.gns = gns_cfg(self.packed, n_loc, e_cap, cfg.feat.node_stride, cfg.feat.node_stride)
        cfg.gns.n_owned = n_own
        ic = cfg.integ
        ic.n, ic.dim, ic.t_window = n_own, self.dim, self.isl
        ic.pos_f64, ic.periodic, ic.out_mode = int(self.tdtype == torch.float64), self.full_mask, 0
        ic.box = _cabi.vec3(self.box, 1.0)
        ic.mean = _cabi.vec3(self.stats["acceleration"]["mean"])
        ic.std = _cabi.vec3(self.stats["acceleration"]["std"], 1.0)
        cfg.cell_capacity, cfg.e_cap = cell_cap, e_cap
        shard_ptr = C.cast(C.pointer(self.shard), C.c_void_p).value
        cfg.shard = shard_ptr
        cfg.gns.shard = shard_ptr
        nbytes = lib.lb200_rollout_scratch_bytes(C.byref(cfg))
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(int(nbytes * 1.05), dtype=torch.uint8, device=self.window.device)
        self._cfg = cfg

    # ------------------------------------------------------------------ the rollout
    def run(self, n_steps):
        """Advance the cloud by ``n_steps``: chunks of ``steps_per_sync`` device-resident steps, one host
        synchronisation per chunk (the status words every rank agrees on).  The most recent positions
        are ``window[:, -1]`` (rows follow ``gid``)."""
        lib, dev = self.lib, self.window.device
        if self._status is None:
            self._status = torch.zeros(4, dtype=torch.int32, device=dev)
            self._stream = torch.cuda.Stream(device=dev)
        done = 0
        while done < n_steps:
            if self._need_select:
                self._select()
                self._configure()
            chunk = min(self.steps_per_sync, n_steps - done)
            frame0 = 0
            tgt = None
            if self.targets is not None:
                frame0 = min(self.steps_done, self.targets.shape[0] - 1)
                chunk = max(1, min(chunk, self.targets.shape[0] - frame0))
                tgt = self.targets
            caller = torch.cuda.current_stream(dev)
            self._stream.wait_stream(caller)
            with torch.cuda.stream(self._stream):
                _cabi.check(lib.lb200_rollout_steps(
                    C.byref(self._cfg), chunk, _cabi.ptr(self.packed.blob), _cabi.ptr(self.window), _cabi.ptr(self.ptype),
                    None, _cabi.ptr(tgt), None, frame0, None, _cabi.ptr(self._status), _cabi.ptr(self._scratch), self._scratch.numel(),
                    _cabi.stream()))
            caller.wait_stream(self._stream)
            completed, bits, n_edges, _ = self._status.tolist()  # the one host synchronisation of the chunk
            done += completed
            self.steps_done += completed
            self.edges_last = n_edges
            self.halo_bytes = self.halo_rows * 512 * self.num_mp_steps
            if bits & _cabi.ERR_NONFINITE:
                raise FloatingPointError("rollout produced NaN / Inf accelerations (fp16 split out of range)")
            if bits & _cabi.OVF_PEER_TIMEOUT:
                raise RuntimeError("a neighbouring rank stopped answering (peer signal timeout)")
            if bits & (_cabi.OVF_NEIGHBOR_LIST | _cabi.OVF_CELL_LIST):
                # rollout.py:135-151: re-allocate from the current state and retry the step.  The bits are
                # OR-ed over the ranks, so every rank comes here; each one re-sizes from its own cloud.
                self.n_reallocations += 1
                self._size(self._pos_local(self.n_own + self.n_ghost_left + self.n_ghost_right).clone())
                self._configure()
            if bits & _cabi.OVF_DRIFT:
                self._need_select = True
        return self

    def step(self):
        return self.run(1)

    def settle(self):
        """Hand particles that left their slab to their new owner now (a run does it lazily)."""
        if self.world > 1:
            self._select()
            self._configure()

    # ------------------------------------------------------------------ results
    def gather_positions(self, n_total):
        """All ranks' most recent positions ordered by global particle id (for checks)."""
        dev = self.window.device
        pos = torch.zeros((n_total, self.dim), dtype=self.tdtype, device=dev)
        pos[self.gid] = self.window[:, -1]
        if self.world > 1:
            if _is_gloo(self.group):
                host = pos.cpu()
                dist.all_reduce(host, group=self.group)
                pos = host.to(dev)
            else:
                dist.all_reduce(pos, group=self.group)
        return pos
