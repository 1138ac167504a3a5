"""world_size-2 test of the ONE exchange step of the multi-GPU path: timesteps dealt
round-robin, each rank builds its share of the bank, a single all-gather distributes it.  Runs on
CPU with the gloo backend; the same code path runs over NCCL in bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from magicdance_b200 import parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layout = P.BankLayout([(16, 8), (4, 16)])  # two fake attention layers
    built = []

    def build_fn(chunk, slots):
        assert slots.shape == (len(chunk), layout.numel) and len(chunk) <= 3
        for index, flat in zip(chunk, slots):
            built.append(index)
            views = layout.views(flat, [16, 4], 1)
            for li, (k, vt, n, b) in enumerate(views):
                k.fill_(index + 0.25 * li)
                vt.fill_(-(index + 0.25 * li))

    indices = list(range(9, -1, -1))  # 10 timesteps over 2 ranks
    table = P.build_and_gather_bank(indices, layout, build_fn, "cpu", world, rank, chunk=3)
    # one asynchronous all-gather per slot row, in consumption order: a step waits for its own row only
    for ix in indices:
        table.wait(ix)
    ok = sorted(table) == list(range(10)) and built == P.shard_timesteps(indices, world, rank)
    for ix, flat in table.items():
        for li, (k, vt, n, b) in enumerate(layout.views(flat, [16, 4], 1)):
            ok &= bool((k == ix + 0.25 * li).all()) and bool((vt == -(ix + 0.25 * li)).all())
            ok &= k.shape == (layout.layer_shapes[li][0], layout.layer_shapes[li][1])
    # frames: each rank gets a disjoint contiguous block
    mine = list(P.shard_frames(7, world, rank))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok &= sorted(sum(gathered, [])) == list(range(7))
    q.put((rank, ok, len(built)))
    dist.destroy_process_group()


def _run(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    return res


def test_bank_allgather_world2_gloo():
    res = _run(2)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [5, 5]  # each rank built exactly its half of the timesteps


def test_bank_allgather_uneven_world3_gloo():
    """10 timesteps over 3 ranks -> 4, 3, 3: the short ranks' last slot is padding that nobody reads — the shape of
    the 8-GPU split of the 50 DDIM steps (7, 7, 6, 6, 6, 6, 6, 6)."""
    res = _run(3)
    assert [r[1] for r in res] == [True, True, True]
    assert [r[2] for r in res] == [4, 3, 3]


def test_bank_allgather_world4_gloo():
    """the 4-GPU leg of the scaling run: 10 timesteps -> 3, 3, 2, 2"""
    res = _run(4)
    assert [r[1] for r in res] == [True] * 4
    assert [r[2] for r in res] == [3, 3, 2, 2]
