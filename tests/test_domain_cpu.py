"""Host logic of the slab decomposition on CPU: world_size-2 (and 3) gloo process groups."""

import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lagrangebench_b200.domain import SlabDomain, _p2p_rows, migrate, select_ghosts


def test_slab_geometry():
    d = SlabDomain([1.0, 2.0, 0.5], axis=1, world=4, rank=0, halo=0.072)
    assert (d.lo, d.hi, d.left, d.right) == (0.0, 0.5, 3, 1)
    y = torch.tensor([0.01, 0.49, 0.5, 1.99, 2.0])
    assert d.owner(y).tolist() == [0, 0, 1, 3, 3]
    left, right = d.halo_masks(torch.tensor([0.01, 0.2, 0.45]))
    assert left.tolist() == [True, False, False] and right.tolist() == [False, False, True]
    assert d.ghost_shift(True) == -2.0 and d.ghost_shift(False) == 0.0
    assert SlabDomain([1.0, 2.0, 0.5], 1, 4, 3, 0.072).ghost_shift(False) == 2.0
    with pytest.raises(ValueError):
        SlabDomain([1.0, 1.0], 0, 8, 0, 0.1)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        box = [1.0, float(world), 1.0]
        dom = SlabDomain(box, axis=1, world=world, rank=rank, halo=0.1)
        g = torch.Generator().manual_seed(100 + rank)
        n = 50 + 7 * rank
        pos = torch.rand((n, 3), generator=g)
        pos[:, 1] = dom.lo + pos[:, 1] * dom.width * 0.999
        gid = torch.arange(n) + 1000 * rank
        # --- ghost selection (one collective) + the row exchange of selection time: every ghost really comes from
        #     the right neighbour and region, counts and destination rows agree between the two sides
        m_l, m_r = dom.halo_masks(pos[:, 1])
        payload = torch.cat([pos, gid[:, None].to(pos.dtype)], dim=1)
        sel = select_ghosts(dom, pos[:, 1])
        ok = torch.equal(sel["send_left"], m_l.nonzero().squeeze(1)) and torch.equal(sel["send_right"], m_r.nonzero().squeeze(1))
        counts = sel["counts_all"]
        ok &= counts[rank] == [n, int(m_l.sum()), int(m_r.sum())]
        ok &= (sel["n_ghost_left"], sel["n_ghost_right"]) == (counts[dom.left][2], counts[dom.right][1])
        # my left-going rows land behind the left neighbour's own rows and its from-left block; right-going ones
        # directly behind the right neighbour's own rows
        ok &= sel["dst_row_left"] == counts[dom.left][0] + counts[(dom.left - 1) % world][2]
        ok &= sel["dst_row_right"] == counts[dom.right][0]
        ok &= sel["n_loc_max"] == max(counts[r][0] + counts[(r - 1) % world][2] + counts[(r + 1) % world][1]
                                      for r in range(world))
        from_left, from_right = _p2p_rows(dom, payload.index_select(0, sel["send_left"]),
                                          payload.index_select(0, sel["send_right"]), sel["n_ghost_left"],
                                          sel["n_ghost_right"])
        ok &= bool(((from_left[:, 3] // 1000).long() == dom.left).all())
        ok &= bool(((from_right[:, 3] // 1000).long() == dom.right).all())
        left_dom = SlabDomain(box, 1, world, dom.left, 0.1)
        right_dom = SlabDomain(box, 1, world, dom.right, 0.1)
        ok &= bool((from_left[:, 1] >= left_dom.hi - 0.1).all())   # the left neighbour's right face
        ok &= bool((from_right[:, 1] < right_dom.lo + 0.1).all())  # the right neighbour's left face
        # --- migration: move everything by +0.6 slab widths (periodic), rows are conserved
        moved = pos.clone()
        moved[:, 1] = torch.remainder(moved[:, 1] + 0.6 * dom.width, box[1])
        ok &= not bool((dom.owner(moved[:, 1]) == rank).all())  # somebody has to leave
        new_pos, new_gid = migrate(dom, moved[:, 1], [moved, gid])
        ok &= bool((dom.owner(new_pos[:, 1]) == rank).all())
        stay = dom.owner(moved[:, 1]) == rank  # stayers first, in their original order
        ok &= torch.equal(new_gid[:int(stay.sum())], gid[stay])
        ok &= new_pos.shape[0] == new_gid.shape[0]
        total = torch.tensor([new_gid.shape[0]])
        dist.all_reduce(total)
        gsum = torch.tensor([int(new_gid.sum())])
        dist.all_reduce(gsum)
        expect_n = sum(50 + 7 * r for r in range(world))
        expect_sum = sum(int((torch.arange(50 + 7 * r) + 1000 * r).sum()) for r in range(world))
        ok &= int(total) == expect_n and int(gsum) == expect_sum
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29611), (3, 29612)])
def test_exchange_and_migrate_gloo(world, port):
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert all(results[r] for r in range(world)), dict(results)
