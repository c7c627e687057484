"""Oracle: cell-list radius neighbor search (test infrastructure only).

Restates ``jax_sph.jax_md.partition.neighbor_list`` (jax-sph 0.0.3, fork of jax-md
``partition.py``; third-party, NOT vendored in the reference) for the one configuration
the reference uses (``lagrangebench/case_setup/case.py:120-130``):
``format=Sparse``, ``mask_self=False``, ``dr_threshold=0``, backend ``jaxmd_vmap``.
Call sites: ``.allocate`` ``case.py:184-186``, ``.update`` ``case.py:188-190``,
``receivers, senders = nbrs.idx`` ``case_setup/features.py:110``.

Published algorithm restated (function names are jax-md's):
  1. ``neighbor_list``: ``box = f32(box)``; ``cutoff = r_cutoff``; the cell list is used
     iff ``all(cutoff < box / 3)``; otherwise every particle's candidates are
     ``0..N-1`` (all pairs).
  2. ``_cell_dimensions``: ``cells_per_side = floor(box / cutoff)`` (f32),
     ``cell_size = box / cells_per_side`` (f32).
  3. ``cell_list.build_cells``: ``hash = sum(int32(pos / cell_size) * mult)`` with
     ``mult = (1, nx, nx*ny)``; particles are STABLE-argsorted by hash; the particle of
     sorted rank ``k`` goes to slot ``hash*cap + (k mod cap)``; empty slots hold ``N``;
     ``cap = int(max_occupancy * multiplier)`` fixed at allocate; overflow if any cell
     holds more than ``cap``.
  4. ``cell_list_candidate_fn``: buffer viewed as ``(nz, ny, nx, cap)``; candidates of a
     particle = own cell's slots, then for every ``dindex`` of
     ``np.ndindex(3,..,3) - 1`` except 0 the slots of cell ``c - dindex`` (wrap-around,
     ``shift_array``), axis 0 of the buffer being the slowest spatial axis.
  5. ``prune_neighbor_list_sparse``: flatten row-major over (center i, candidate slot);
     keep candidate j iff ``sum(disp(pos[i], pos[j])**2) < cutoff**2`` and ``j < N``;
     stable compaction; ``idx = stack(j, i)``; pad with ``N``.
  6. capacity: at allocate ``E_cap = int(E * multiplier)`` clipped to the candidate-table
     size and ``N**2``; at update the list is truncated to ``E_cap`` and
     ``did_buffer_overflow = (E > E_cap) | cell overflow``.

Pinned by the reference's only neighbor-list golden vector,
``tests/case_test.py:77-82`` (``idx == [[0,1,2,2,1,3],[0,1,1,2,2,3]]``), replayed in
``tests/test_oracle_golden.py``.  NOT pinned by any reference test (and therefore
"parity unpinned"): the slot rotation of rule 3, the offset order / axis
mapping of rule 4 and the tie behaviour at ``d^2 == r^2``.

Documented deviation: when ``E_cap`` equals the full candidate-table size while some
candidates are masked out, jax-md's compaction scatters all masked-out entries onto the
last slot (duplicate scatter indices: update order is XLA-implementation-defined).  The
oracle pads that slot with ``N`` instead.
"""

import itertools

import numpy as np

from . import space


class NeighborList:
    """Mirror of the fields of jax-md's ``NeighborList`` that lagrangebench touches."""

    def __init__(self, idx, reference_position, did_buffer_overflow, cell_list_capacity,
                 max_occupancy, n_edges, update_fn):
        self.idx = idx  # (2, E_cap) int32; row 0 = receivers (candidate j), row 1 = senders (center i)
        self.reference_position = reference_position
        self.did_buffer_overflow = did_buffer_overflow
        self.cell_list_capacity = cell_list_capacity
        self.max_occupancy = max_occupancy  # E_cap
        self.n_edges = n_edges  # un-truncated true edge count (oracle extra)
        self.update_fn = update_fn

    def update(self, position, num_particles=None, **kwargs):
        return self.update_fn(position, self, num_particles=num_particles)


class NeighborListFns:
    def __init__(self, allocate, update):
        self.allocate = allocate
        self.update = update


def cell_dimensions(box, cutoff):
    """``_cell_dimensions`` in f32 (``box = f32(box)`` in ``neighbor_list``)."""
    box32 = np.asarray(box, dtype=np.float32)
    cut32 = np.float32(cutoff)
    cells_per_side_f = np.floor(box32 / cut32)
    cell_size = box32 / cells_per_side_f  # f32
    cells_per_side = cells_per_side_f.astype(np.int32)
    return cell_size, cells_per_side


def use_cell_list(box, cutoff):
    box32 = np.asarray(box, dtype=np.float32)
    return bool(np.all(np.float32(cutoff) < box32 / np.float32(3.0)))


def neighbor_offsets(dim):
    """Candidate-cell order: own cell, then ``ndindex(3,..)-1`` skipping zero."""
    offs = [np.zeros(dim, dtype=np.int64)]
    for dindex in itertools.product((-1, 0, 1), repeat=dim):
        if all(v == 0 for v in dindex):
            continue
        offs.append(np.array(dindex, dtype=np.int64))
    return np.stack(offs)  # (3^dim, dim); component 0 acts on the slowest buffer axis


def cell_hashes(position, cell_size, cells_per_side):
    dim = position.shape[1]
    cs = cell_size.astype(position.dtype)
    indices = (position / cs).astype(np.int32)  # truncation toward zero
    mult = np.concatenate(([1], np.cumprod(cells_per_side[:-1]))).astype(np.int32)
    assert mult.shape[0] == dim
    return (indices * mult).sum(axis=1).astype(np.int32)


def neighbor_list(displacement_fn, box, r_cutoff, capacity_multiplier=1.25, dtype=np.float32):
    """Returns ``NeighborListFns(allocate, update)``.

    ``dtype`` is the effective position dtype (float64 only when JAX x64 is enabled in
    the reference, ``lagrangebench/runner.py:35-36``)."""
    dtype = np.dtype(dtype)
    box = np.asarray(box, dtype=np.float64)
    dim = box.shape[0]
    cutoff = dtype.type(r_cutoff)
    cutoff_sq = cutoff * cutoff
    with_cells = use_cell_list(box, r_cutoff)
    if with_cells:
        cell_size, cells_per_side = cell_dimensions(box, r_cutoff)
        if np.any(cells_per_side < 3):
            raise ValueError("Box must be at least 3x the size of the grid spacing in each dimension.")
        cell_count = int(np.prod(cells_per_side))
        offsets = neighbor_offsets(dim)
        # neighbour-cell table (cell_count, 3^dim) in candidate order
        dims_rev = tuple(int(x) for x in cells_per_side[::-1])  # (nz, ny, nx)
        coords = np.stack(np.unravel_index(np.arange(cell_count), dims_rev), axis=1)  # slowest axis first
        nbr_cells = np.empty((cell_count, offsets.shape[0]), dtype=np.int64)
        for o, off in enumerate(offsets):
            c = (coords - off[None, :]) % np.array(dims_rev)[None, :]
            nbr_cells[:, o] = np.ravel_multi_index(tuple(c.T), dims_rev)

    def _build(position, nbrs, num_particles=None):
        position = np.asarray(position, dtype=dtype)
        n_rows = position.shape[0]  # the pad value of the list
        # ``num_particles`` (case.py:182-190): the first rows are real particles, the rest padding
        # (``data.py:183-197``, particle type -1) that stays out of the search -- what the reference's
        # matscipy backend, the only one that pads, does with it (jax-sph 0.0.3, unvendored)
        n = n_rows if num_particles is None else int(num_particles)
        position = position[:n]
        cell_overflow = False
        if with_cells:
            hashes = cell_hashes(position, cell_size, cells_per_side)
            occupancy = np.bincount(hashes, minlength=cell_count)
            if nbrs is None:
                cap = int(occupancy.max() * capacity_multiplier)
            else:
                cap = nbrs.cell_list_capacity
            cell_overflow = bool(occupancy.max() > cap)
            order = np.argsort(hashes, kind="stable")
            sorted_hash = hashes[order]
            slot = sorted_hash.astype(np.int64) * cap + (np.arange(n) % cap)
            buf = np.full(cell_count * cap, n, dtype=np.int32)
            buf[slot] = order.astype(np.int32)
            buf = buf.reshape(cell_count, cap)
            n_cand = offsets.shape[0] * cap
        else:
            cap = None
            n_cand = n
        recv_parts, send_parts = [], []
        chunk = max(1, min(n, (1 << 22) // max(1, n_cand)))
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            ids = np.arange(a, b)
            if with_cells:
                cand = buf[nbr_cells[hashes[a:b]]].reshape(b - a, n_cand)
            else:
                cand = np.broadcast_to(np.arange(n, dtype=np.int32)[None, :], (b - a, n))
            valid = cand < n
            j = np.where(valid, cand, n - 1)
            d_r = displacement_fn(position[a:b][:, None, :], position[j])
            d2 = space.sum_sq(d_r)
            mask = (d2 < cutoff_sq) & valid
            recv_parts.append(cand[mask].astype(np.int32))
            send_parts.append(np.broadcast_to(ids[:, None], cand.shape)[mask].astype(np.int32))
        recv = np.concatenate(recv_parts)
        send = np.concatenate(send_parts)
        n_edges = int(recv.shape[0])
        if nbrs is None:
            e_cap = int(n_edges * capacity_multiplier)
            e_cap = min(e_cap, n_rows * n_cand if with_cells else n_rows * n_rows, n_rows * n_rows)
        else:
            e_cap = nbrs.max_occupancy
        idx = np.full((2, e_cap), n_rows, dtype=np.int32)
        m = min(e_cap, n_edges)
        idx[0, :m] = recv[:m]
        idx[1, :m] = send[:m]
        overflow = bool(n_edges > e_cap) or cell_overflow
        if nbrs is not None:
            overflow = overflow or bool(nbrs.did_buffer_overflow)  # jax-md error codes are sticky
        out = NeighborList(idx, position, overflow, cap, e_cap, n_edges, update)
        out.cell_overflow = cell_overflow  # oracle extra: list contents are unspecified when set
        return out

    def allocate(position, num_particles=None, **kwargs):
        return _build(position, None, num_particles)

    def update(position, nbrs, num_particles=None, **kwargs):
        return _build(position, nbrs, num_particles)

    return NeighborListFns(allocate, update)
