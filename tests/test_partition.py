"""CPU tests of the multi-GPU host logic: the FLOP-weighted partition of (exploration set x grid tile) and, with two
gloo processes, the all-gather of the per-set bests followed by the deterministic combine rule."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from cbo_with_oop_b200.partition import SetSize, partition


def covered_once(sizes, world, tile=128):
    sl = partition(sizes, world, tile)
    assert len(sl) == world and all(len(r) == len(sizes) for r in sl)
    for s, sz in enumerate(sizes):
        pos = 0
        for r in range(world):
            b, c = sl[r][s]
            if c:
                assert b == pos, (s, r, sl[r][s], pos)      # contiguous and in rank order
                assert b % tile == 0                           # cuts fall on tile boundaries
                pos += c
        assert pos == sz.g_total
    return sl


def test_config5_partitions_into_whole_sets():
    sizes = [SetSize(10 ** 6, 10 ** 4, 32)] * 16
    for world in (1, 2, 4, 8):
        sl = covered_once(sizes, world)
        for r in range(world):
            owned = [s for s in range(16) if sl[r][s][1]]
            assert len(owned) == 16 // world and all(sl[r][s] == (0, 10 ** 6) for s in owned)


def test_mixed_set_sizes_are_balanced_and_exact():
    sizes = [SetSize(100, 100, 10)] * 5 + [SetSize(10 ** 4, 100, 10)] * 10 + [SetSize(10 ** 6, 100, 10)] * 10   # config 2 shape
    for world in (2, 3, 8):
        sl = covered_once(sizes, world)
        load = [sum(c * sizes[s].weight for s, (b, c) in enumerate(r)) for r in sl]
        assert max(load) / (sum(load) / world) < 1.10


def test_degenerate_cases():
    covered_once([SetSize(50, 10, 5)], 4)            # fewer tiles than ranks: some ranks get nothing
    covered_once([SetSize(7, 0, 3), SetSize(129, 0, 3)], 2)
    with pytest.raises(ValueError):
        partition([SetSize(10, 1, 1)], 0)


def combine_reference(tables):
    """NumPy statement of csrc/sweep.cu combine_kernel: tables[rank][set] = (value, index, n_nan)."""
    R, S = len(tables), len(tables[0])
    per_set = []
    for s in range(S):
        best = (-np.inf, np.iinfo(np.int64).max, 0)
        nn = 0
        for r in range(R):
            v, i, n = tables[r][s]
            ii = np.iinfo(np.int64).max if i < 0 else i
            nn += n
            if v > best[0] or (v == best[0] and ii < best[1]):
                best = (v, ii, 0)
        per_set.append((best[0], -1 if best[1] == np.iinfo(np.int64).max else best[1], nn))
    sel = (-np.inf, -1, -1)
    for s, (v, i, n) in enumerate(per_set):
        if i >= 0 and (sel[2] < 0 or v > sel[0]):
            sel = (v, i, s)
    return per_set, sel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, out_dir):
    import torch
    import torch.distributed as dist
    from cbo_with_oop_b200._lib import SetBest
    from cbo_with_oop_b200.dist import gather_set_bests
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sizes = [SetSize(1000 + 300 * s, 50, 8) for s in range(S)]
    mine = partition(sizes, world)[rank]
    rng = np.random.default_rng(100 + rank)
    table = (SetBest * S)()
    for s in range(S):
        b, c = mine[s]
        if c:
            table[s].value, table[s].index, table[s].n_nan = float(np.round(rng.normal(), 1)), int(b + rng.integers(c)), int(rank + 1)
        else:
            table[s].value, table[s].index, table[s].n_nan = -np.inf, -1, 0
    local = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
    gathered = torch.empty(world * S * C.sizeof(SetBest), dtype=torch.uint8)
    gather_set_bests(local, gathered, world)
    np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), gathered.numpy())
    np.save(os.path.join(out_dir, f"local_{rank}.npy"), local.numpy())
    dist.destroy_process_group()


def test_all_gather_and_combine_with_two_gloo_ranks(tmp_path):
    import torch.multiprocessing as mp
    from cbo_with_oop_b200._lib import SetBest
    world, S = 2, 5
    mp.spawn(_worker, args=(world, _free_port(), S, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "gathered_0.npy"), np.load(tmp_path / "gathered_1.npy")
    np.testing.assert_array_equal(g0, g1)                                  # every rank sees the same table
    for r in range(world):                                                 # rank-major layout
        np.testing.assert_array_equal(g0.reshape(world, -1)[r], np.load(tmp_path / f"local_{r}.npy"))
    rows = (SetBest * (world * S)).from_buffer_copy(g0.tobytes())
    tables = [[(rows[r * S + s].value, rows[r * S + s].index, rows[r * S + s].n_nan) for s in range(S)] for r in range(world)]
    per_set, sel = combine_reference(tables)
    sizes = [SetSize(1000 + 300 * s, 50, 8) for s in range(S)]
    sl = partition(sizes, world)
    for s in range(S):
        owners = [r for r in range(world) if sl[r][s][1]]
        assert per_set[s][1] >= 0 and per_set[s][2] == sum(r + 1 for r in owners)
        assert per_set[s][0] == max(tables[r][s][0] for r in owners)
    assert sel[2] == int(np.argmax([p[0] for p in per_set]))               # first set attaining the maximum
