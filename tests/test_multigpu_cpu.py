"""N>1 path on CPU: two gloo ranks, each owning a key-range shard answered by the CPU oracle; fixed-size per-query records
are all-gathered and merged; the result equals the unsharded oracle. No GPU needed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, nq, limit, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import loader
    from shard_ref import RECORD_DTYPE, merge_records, pack_records, shard_bounds
    from stringsearchlib_b200 import synth
    from util import spec_keys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = shard_bounds(n, world)
    keys = synth.ascii_keys(b[rank + 1] - b[rank], first=b[rank])
    # a duplicate of a key that lives in the other shard: the merge must collapse it
    S = loader.SpecIndex(keys, 1, None, 3, False, row_base=b[rank])
    q = synth.ascii_queries(nq, n)
    counts, ext, sc = S.search_batch(q, 0.0, limit, threads=2)
    t = S.tables()
    rec = pack_records(counts, ext, sc, spec_keys(t), t["firstRowExt"], limit)
    mine = torch.from_numpy(rec.view(np.uint8).reshape(nq, limit, RECORD_DTYPE.itemsize).copy())
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    g = np.stack([x.numpy() for x in gathered]).reshape(world, nq, limit * RECORD_DTYPE.itemsize).view(RECORD_DTYPE)
    merged, mcounts = merge_records(g.reshape(world, nq, limit), limit)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    np.save(os.path.join(out_dir, f"counts_{rank}.npy"), mcounts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_unsharded(oracle_built, tmp_path):
    import torch.multiprocessing as mp
    from shard_ref import key_hash
    from stringsearchlib_b200 import synth
    from util import spec_keys
    n, nq, limit, world = 20000, 64, 10, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, nq, limit, str(tmp_path)), nprocs=world, join=True)
    keys = synth.ascii_keys(n)
    S = oracle_built.SpecIndex(keys, 1)
    t = S.tables()
    names = spec_keys(t)
    counts, ext, sc = S.search_batch(synth.ascii_queries(nq, n), 0.0, limit)
    m0 = np.load(tmp_path / "merged_0.npy")
    m1 = np.load(tmp_path / "merged_1.npy")
    assert np.array_equal(m0, m1)                                # every rank ends with the same merged lists
    assert np.array_equal(np.load(tmp_path / "counts_0.npy"), counts)
    for i in range(nq):
        c = int(counts[i])
        assert np.array_equal(m0["scoreBits"][i, :c], sc[i, :c].view(np.uint32))
        assert np.array_equal(m0["row"][i, :c], t["firstRowExt"][ext[i, :c]])
        assert [int(h) for h in m0["hash"][i, :c]] == [key_hash(names[e]) for e in ext[i, :c]]


def test_merge_collapses_equal_keys_across_shards():
    from shard_ref import EMPTY_ROW, RECORD_DTYPE, merge_records
    f = lambda x: np.float32(x).view(np.uint32)
    g = np.zeros((2, 1, 3), RECORD_DTYPE)
    g["row"] = EMPTY_ROW
    g[0, 0, 0] = (f(0.5), 6, 10, 111)      # key A seen by shard 0 with score 0.5
    g[0, 0, 1] = (f(0.25), 5, 11, 222)
    g[1, 0, 0] = (f(0.75), 6, 90, 111)     # key A again in shard 1 with a better score -> one entry, score 0.75
    g[1, 0, 1] = (f(0.25), 5, 95, 333)     # ties on (score, length): smaller row first
    m, c = merge_records(g, 3)
    assert c[0] == 3
    assert [int(x) for x in m["hash"][0]] == [111, 222, 333]
    assert m["scoreBits"][0, 0] == f(0.75)


def _sizes_worker(rank, world, port, rows, rs, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import loader
    from shard_ref import count_distinct_part, digests_of, shard_bounds, spec_terms
    from util import random_corpus
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    words, w = random_corpus(np.random.default_rng(5), rows, rs, weighted=True, max_len=14)
    b = shard_bounds(rows, world)
    lo, hi = b[rank] * rs, b[rank + 1] * rs
    S = loader.SpecIndex(words[lo:hi], rs, w[lo:hi], 3, False, row_base=b[rank])
    t = S.tables()
    res = []
    for mine in (digests_of(spec_terms(t)), digests_of(list(t["gramUnits"]))):
        # all-gather of the counts, then of the padded digests (what api.global_sizes does with NCCL)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int64))
        counts = [int(c) for c in counts]
        cap = max(max(counts), 1)
        pad = np.zeros((cap, 16), np.uint8)
        pad[:len(mine)] = np.frombuffer(mine.tobytes(), np.uint8).reshape(-1, 16)
        gathered = [torch.zeros((cap, 16), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(pad))
        flat = np.concatenate([g.numpy()[:c] for g, c in zip(gathered, counts)])
        flat = np.frombuffer(flat.tobytes(), dtype="V16")
        part = torch.tensor([count_distinct_part(flat, rank, world)], dtype=torch.int64)
        dist.all_reduce(part)
        res.append((int(part), sum(counts)))
    np.save(os.path.join(out_dir, f"sizes_{rank}.npy"), np.array(res, np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_global_sizes_equal_unsharded(oracle_built, tmp_path):
    """getSize / getLibSize over two key-range shards = the unsharded oracle's (SURVEY.md §8e "Global counters")."""
    import torch.multiprocessing as mp
    from util import random_corpus
    rows, rs, world = 3000, 2, 2
    mp.spawn(_sizes_worker, args=(world, _free_port(), rows, rs, str(tmp_path)), nprocs=world, join=True)
    words, w = random_corpus(np.random.default_rng(5), rows, rs, weighted=True, max_len=14)
    S = oracle_built.SpecIndex(words, rs, w)
    r0, r1 = np.load(tmp_path / "sizes_0.npy"), np.load(tmp_path / "sizes_1.npy")
    assert np.array_equal(r0, r1)
    assert (int(r0[0, 0]), int(r0[1, 0])) == (S.size(), S.lib_size())
    assert r0[0, 1] > r0[0, 0] and r0[1, 1] > r0[1, 0]          # terms and grams are shared between the shards: sums overcount
