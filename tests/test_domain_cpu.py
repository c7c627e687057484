"""Host logic of the slab decomposition on CPU: world_size-2 (and 3) gloo process groups."""

import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lagrangebench_b200.domain import SlabDomain, exchange_rows, exchange_rows_sized, halo_sets, migrate, step_counts


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
        # --- halo exchange: every ghost really comes from the right neighbour and region
        m_l, m_r = dom.halo_masks(pos[:, 1])
        payload = torch.cat([pos, gid[:, None].to(pos.dtype)], dim=1)
        from_left, from_right = exchange_rows(dom, payload[m_l], payload[m_r])
        ok = True
        if world > 1:
            ok &= bool(((from_left[:, 3] // 1000).long() == dom.left).all())
            ok &= bool(((from_right[:, 3] // 1000).long() == dom.right).all())
            left_dom = SlabDomain(box, 1, world, dom.left, 0.1)
            right_dom = SlabDomain(box, 1, world, dom.right, 0.1)
            ok &= bool((from_left[:, 1] >= left_dom.hi - 0.1).all())   # the left neighbour's right face
            ok &= bool((from_right[:, 1] < right_dom.lo + 0.1).all())  # the right neighbour's left face
        # --- the single-synchronisation variant used by the rollout: same sets, same order, known counts
        s_l, s_r, n_fl, n_fr = halo_sets(dom, pos[:, 1])
        ok &= torch.equal(s_l, m_l.nonzero().squeeze(1)) and torch.equal(s_r, m_r.nonzero().squeeze(1))
        if world > 1:
            fl2, fr2 = exchange_rows_sized(dom, payload.index_select(0, s_l), payload.index_select(0, s_r), n_fl, n_fr)
            ok &= torch.equal(fl2, from_left) and torch.equal(fr2, from_right)
        else:
            ok &= n_fl == 0 and n_fr == 0
        # --- the merged collective of a rollout step: migration matrix + every rank's halo counts
        matrix, halo_counts, o_l, o_r = step_counts(dom, pos[:, 1])
        ok &= int(matrix.sum()) == int(matrix.diagonal().sum())  # everybody is at home
        ok &= int(matrix[rank, rank]) == n
        ok &= torch.equal(o_l[:int(m_l.sum())], s_l) and torch.equal(o_r[:int(m_r.sum())], s_r)
        ok &= (int(halo_counts[rank, 0]), int(halo_counts[rank, 1])) == (int(m_l.sum()), int(m_r.sum()))
        ok &= (int(halo_counts[dom.left, 1]), int(halo_counts[dom.right, 0])) == (n_fl, n_fr) or world == 1
        # --- migration: move everything by +0.6 slab widths (periodic), rows are conserved
        moved = pos.clone()
        moved[:, 1] = torch.remainder(moved[:, 1] + 0.6 * dom.width, box[1])
        matrix2 = step_counts(dom, moved[:, 1])[0]
        ok &= world == 1 or int(matrix2.sum()) != int(matrix2.diagonal().sum())
        new_pos, new_gid = migrate(dom, moved[:, 1], [moved, gid], matrix=matrix2 if world > 1 else None)
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
