"""Host logic of the table-sharded mode that needs neither a GPU nor a process group."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ceo-recommender_b200"))


def _loads(towers, pieces, slices, owner, world):
    load = [0] * world
    for (t, k, c), r in zip(slices, owner):
        load[r] += towers[t][1] // pieces[t]
    return load


def test_every_slice_has_exactly_one_owner_and_plans_are_deterministic():
    from ceo_firm_matching.distributed import plan_table_slices
    for towers in ([(4, 48), (7, 8)], [(4, 8), (7, 8)], [(1, 32)], [(3, 6), (2, 10)], [(16, 64)]):
        for world in (1, 2, 3, 4, 8):
            pieces, slices, owner = plan_table_slices(towers, world)
            assert (pieces, slices, owner) == plan_table_slices(towers, world)          # same plan on every rank
            assert len(slices) == len(set(slices)) == sum(K * pieces[t] for t, (K, E) in enumerate(towers))
            assert all(0 <= r < world for r in owner)
            for t, (K, E) in enumerate(towers):
                assert E % pieces[t] == 0
                assert pieces[t] == 1 or (E // pieces[t]) % 4 == 0                       # slices stay float4 rows
            assert sum(_loads(towers, pieces, slices, owner, world)) == sum(K * E for K, E in towers)


def test_more_ranks_than_slices_leaves_ranks_without_tables():
    from ceo_firm_matching.distributed import plan_table_slices
    pieces, slices, owner = plan_table_slices([(1, 6)], 8)          # width 6 cannot be cut into float4 slices
    assert pieces == [1] and len(slices) == 1 and owner == [0]


def test_shard_bounds_cover_the_range_without_overlap():
    from ceo_firm_matching.distributed import shard_bounds
    for n in (0, 1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
