"""Host logic of the multi-GPU path, exercised with world_size-2 gloo process groups on CPU."""

import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from windgnn_b200.shard import aggregate_throughput, max_over_ranks, shard_range


def test_shard_range_covers_batch_exactly():
    for n in (0, 1, 7, 4096, 1 << 20, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        slow = max_over_ranks(0.5 + rank)              # rank r pretends to take 0.5 + r seconds
        thr = aggregate_throughput(100, 0.5 + rank)
        # shards of an elementwise-independent op reassemble to the unsharded result
        x = torch.arange(n, dtype=torch.float32)
        parts = [None] * world
        dist.all_gather_object(parts, (x[lo:hi] * 2 + 1).tolist())
        q.put((rank, spans, slow, thr, sum(parts, [])))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_reduction():
    world, n = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, spans, slow, thr, whole in results:
        assert spans == [(0, 6), (6, 11)]
        assert slow == 1.5                              # max over ranks, not the local time
        assert thr == pytest.approx(2 * 100 / 1.5)      # all ranks' units / slowest rank
        assert whole == [2.0 * i + 1 for i in range(n)]


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """`bench.py --impl reference` launched like the driver does for N > 1: rank 0 alone works and
    prints the JSON line, the other rank exits 0 silently."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--ref-seqs", "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    # the unmodified reference modules when oracle/_ref has been built (oracle/build_ref.py), else the port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "step6_gcn_gru_combined_model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["same_config"] is False   # 16 of 4096 here


# ---- data-parallel training: one flat gradient bucket, one all-reduce, mean folded into the optimiser ----
def _train_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np

        from oracle.train_oracle import adam_step
        from windgnn_b200.train import allreduce_mean_, split_flat

        shapes = [(3, 2), (2,), (4, 3), (4,)]
        n = sum(int(np.prod(s)) for s in shapes)
        # each rank's "local gradient": deterministic, different per rank
        local = torch.arange(n, dtype=torch.float32) * (rank + 1) - 3.0
        bucket = local.clone()
        scale = allreduce_mean_(bucket)
        views = split_flat(bucket, shapes)
        # what the optimiser then does (the CUDA Adam kernel applies `scale` itself): oracle Adam on CPU
        p0 = np.linspace(-1, 1, n)
        p1, _, _ = adam_step(p0, bucket.numpy().astype(np.float64) * scale, np.zeros(n), np.zeros(n), 1)
        q.put((rank, scale, bucket.tolist(), [tuple(v.shape) for v in views], p1.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_bucket_allreduce():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 6 + 2 + 12 + 4
    expect_sum = [(i * 1 - 3.0) + (i * 2 - 3.0) for i in range(n)]
    (r0, s0, b0, sh0, p0), (r1, s1, b1, sh1, p1) = results
    assert s0 == s1 == 0.5
    assert b0 == b1 == expect_sum                      # both ranks hold the summed bucket
    assert sh0 == [(3, 2), (2,), (4, 3), (4,)]
    assert p0 == p1                                    # replicas stay bit-identical after the step


def test_allreduce_mean_is_identity_without_a_process_group():
    from windgnn_b200.train import allreduce_mean_, split_flat

    t = torch.arange(5.0)
    assert allreduce_mean_(t) == 1.0 and t.tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(RuntimeError):
        split_flat(t, [(2,), (2,)])


def test_uneven_shards_weighting_reproduces_the_global_mean():
    from windgnn_b200.train import shard_weight

    assert shard_weight(512, 512) == 1.0                        # single process
    # two ranks with 3 and 5 windows: mean of the weighted local means == the global mean
    import numpy as np

    rng = np.random.default_rng(0)
    a, b = rng.random(3), rng.random(5)
    w_a, w_b = 3 * 2 / 8, 5 * 2 / 8                             # what shard_weight returns under world = 2
    assert (w_a * a.mean() + w_b * b.mean()) / 2 == pytest.approx(np.concatenate([a, b]).mean())
    with pytest.raises(ValueError):
        shard_weight(4, 0)
