"""CPU suite, part 3: the multi-GPU sharding logic (row bands of one large image, round-robin frame ownership).
The N>1 path is exercised with a world_size-2 gloo job: every rank plans its own shards, the plans are all-gathered and
must tile the work exactly with no overlap -- the same host logic bench.py / the band API run per GPU."""
import os
import socket
import sys

import numpy as np
import pytest

import anime4kcpp_b200 as A


def test_band_plan_tiles_the_output_exactly():
    for h in (1, 7, 64, 1080, 8192):
        for factor in (2.0, 4.0):
            for n in (1, 2, 3, 8):
                power = int(np.log2(factor))
                prev = 0
                for b in range(n):
                    sy0, sy1, oy0, oy1 = A.band_plan(h, factor, 9, n, b)
                    assert oy0 == prev and oy1 >= oy0 and 0 <= sy0 <= sy1 <= h
                    prev = oy1
                    if oy1 > oy0:
                        # the band's source window covers its own rows plus the full receptive field (or the true border)
                        need = 9 + 3 if power == 1 else 2 * 9 + 3
                        assert sy0 == max(0, (oy0 >> power) - need) and sy1 == min(h, (oy1 >> power) + need)
                assert prev == h << power


def test_band_plan_rejects_bad_requests():
    with pytest.raises(A.Acb200Error):
        A.band_plan(100, 3.0, 9, 2, 0)      # not a power of two
    with pytest.raises(A.Acb200Error):
        A.band_plan(100, 2.0, 9, 2, 2)      # band index out of range


def test_model_halo_is_the_layer_count():
    assert A.Model("acnet-legacy-hdn0").halo() == 9 and A.Model("acnet-f8b4").halo() == 6
    assert A.Model("acnet-f8b18").halo() == 20 and A.Model("arnet-f8b64").halo() == 130


def test_frame_owner_round_robin():
    assert [A.frame_owner(i, 3) for i in range(7)] == [0, 1, 2, 0, 1, 2, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, h, factor, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank plans ITS shard exactly as it would before touching its GPU
    sy0, sy1, oy0, oy1 = A.band_plan(h, factor, 9, world, rank)
    mine = torch.tensor([sy0, sy1, oy0, oy1, sum(1 for f in range(n_frames) if A.frame_owner(f, world) == rank)], dtype=torch.int64)
    allp = [torch.zeros(5, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allp, mine)
    # the "host-side gather": concatenate the bands the ranks own, in rank order
    rows = torch.zeros(int(h * factor), dtype=torch.int64)
    rows[oy0:oy1] = rank + 1
    dist.all_reduce(rows)
    if rank == 0:
        q.put(([p.tolist() for p in allp], rows.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_job_shards_bands_and_frames():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    h, factor, n_frames, world = 1081, 2.0, 17, 2
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, h, factor, n_frames, q)) for r in range(world)]
    [p.start() for p in procs]
    plans, rows = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert plans[0][2] == 0 and plans[0][3] == plans[1][2] and plans[1][3] == int(h * factor)      # bands abut, nothing missing
    assert plans[0][1] > plans[0][3] // 2 and plans[1][0] < plans[1][2] // 2                       # halo overlap on both sides of the cut
    assert plans[0][4] + plans[1][4] == n_frames and abs(plans[0][4] - plans[1][4]) <= 1            # frames dealt evenly
    assert set(rows) == {1, 2} and rows == sorted(rows)                                              # every output row owned exactly once
