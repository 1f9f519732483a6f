"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard utterances by rank and all-gather fixed-width hypothesis records."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from espnet_b200.sharding import gather_hypotheses, pack_hypotheses, shard_indices, unpack_hypotheses


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_utts, rank, world)
    # fake per-utterance n-best: tokens derived from the global utterance index
    local = [[(list(range(1, 2 + (i % 5))), -float(i) - 0.5)] for i in mine]
    rec = pack_hypotheses(local, nbest=1, max_tokens=8)
    allrec = gather_hypotheses(rec, world)
    if rank == 0:
        q.put([unpack_hypotheses(allrec[r], nbest=1) for r in range(world)])
    dist.destroy_process_group()


def _worker_results(rank, world, port, q):
    """all_gather_results over Speech2Text-shaped result lists (with scores); rank 1 holds one utterance fewer and one empty n-best."""
    from collections import namedtuple

    from espnet_b200.sharding import all_gather_results

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = namedtuple("H", "score")
    if rank == 0:
        local = [[("a", ["x"], [3, 4, 5], H(-1.25)), ("b", ["y"], [3, 4], H(-2.5))], [("c", ["z"], [9], H(-0.125))]]
    else:
        local = [[]]
    out = all_gather_results(local, nbest=2, max_tokens=6, world=world)
    q.put((rank, out))
    dist.destroy_process_group()


def test_all_gather_results_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_results, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):      # every rank sees the whole job
        out = got[r]
        assert out[0] == [[([3, 4, 5], -1.25), ([3, 4], -2.5)], [([9], -0.125)]]
        assert out[1] == [[]]


def test_shard_and_gather_two_ranks():
    world, n_utts = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_utts, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reassemble in global order
    out = {}
    for r in range(world):
        for j, i in enumerate(shard_indices(n_utts, r, world)):
            out[i] = got[r][j]
    assert sorted(out) == list(range(n_utts))
    for i in range(n_utts):
        toks, score = out[i][0]
        assert toks == list(range(1, 2 + (i % 5)))
        assert abs(score - (-float(i) - 0.5)) < 1e-6


def test_empty_nbest_keeps_its_row():
    """A real utterance whose search ended nothing (empty n-best) must not be confused with a padding row."""
    rec = pack_hypotheses([[([3, 4], -1.0)], [], [([5], -2.0)]], nbest=2, max_tokens=4, rows=5)
    out = unpack_hypotheses(rec, nbest=2)
    assert len(out) == 3 and out[1] == [] and out[2][0][0] == [5] and out[0][0][0] == [3, 4]


def test_shard_indices_cover_everything():
    for n in (1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            allidx = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert allidx == list(range(n))
