"""CPU tests of the multi-GPU host logic (boda_b200/shard.py) with world_size = 2 over gloo: the same code path bench.py runs over NCCL.

The per-rank compute stand-in here is the CPU oracle's whole-net forward (tests may use the oracle as a checker; the product's forward
needs a GPU): what is under test is the partitioning, the single weight broadcast and the logits gather -- i.e. that sharding the batch,
running each shard with broadcast weights and gathering in rank order reproduces the unsharded forward bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_batch():
    from boda_b200 import shard
    for B in (0, 1, 7, 32, 33, 256):
        for G in (1, 2, 3, 4, 8):
            rs = [shard.shard_range(B, G, r) for r in range(G)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            for (b0, e0), (b1, e1) in zip(rs, rs[1:]):
                assert e0 == b1 and b0 <= e0
            assert max(e - b for b, e in rs) == -(-B // G) or B == 0
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)
    x = np.arange(10 * 2).reshape(10, 2)
    parts = shard.split_inputs(x, 4)
    assert [p.shape[0] for p in parts] == [3, 3, 3, 1] and np.array_equal(np.concatenate(parts), x)


def test_param_layout_is_deterministic():
    from boda_b200 import shard
    lay = shard.param_layout({"b_filts": (2, 3), "a_biases": (5,), "c": (1, 1, 2)})
    assert [l[0] for l in lay] == ["a_biases", "b_filts", "c"]
    assert [(l[1], l[2]) for l in lay] == [(0, 5), (5, 6), (11, 2)]


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        from boda_b200 import nets, shard
        from oracle import net_oracle
        dist.init_process_group("gloo", rank=rank, world_size=world)
        B = 4  # global batch, 2 images per rank
        b0, b1 = shard.shard_range(B, world, rank)
        txt_local, in_node, out_node = nets.tiny_net(b1 - b0)
        shapes = nets.conv_param_shapes(txt_local)
        params0 = nets.synth_params(txt_local) if rank == 0 else None  # only rank 0 synthesises; everyone else receives
        params = shard.broadcast_params(dist, shapes, params0)
        x_global = nets.synth_input((B, 3, 31, 29))
        x_local = shard.split_inputs(x_global, world)[rank]
        local = net_oracle.run_pipe(txt_local, {in_node: x_local}, params)[out_node]
        gathered = shard.gather_logits(dist, torch.from_numpy(np.ascontiguousarray(local))).numpy()
        # the one-step-late gather pipeline bench.py runs: step i's logits (here: local * (i + 1)) come out of step i+1's end_step()
        gp = shard.GatherPipeline(dist, torch.from_numpy(np.ascontiguousarray(local)))
        pipe_ok, outs = True, []
        for i in range(3):
            gp.begin_step(i)
            gp.stage(i, torch.from_numpy(np.ascontiguousarray(local * np.float32(i + 1))))
            o = gp.end_step()
            pipe_ok &= (o is None) == (i == 0)
            if o is not None:
                outs.append(o.numpy().copy())
        outs.append(gp.drain(3).numpy().copy())
        pipe_ok &= len(outs) == 3 and all(np.array_equal(o, gathered * np.float32(i + 1)) for i, o in enumerate(outs))
        t = shard.max_over_ranks(dist, float(rank + 1))
        res = {"rank": rank, "ok_t": t == float(world), "pipe_ok": bool(pipe_ok), "params_sum": float(sum(float(np.abs(v).sum()) for v in params.values())), "gathered": gathered}
        if rank == 0:
            txt_full, i2, o2 = nets.tiny_net(B)
            full = net_oracle.run_pipe(txt_full, {i2: x_global}, nets.synth_params(txt_full))[o2]
            res["match_full"] = bool(np.array_equal(full, gathered))
        q.put(res)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent
        q.put({"rank": rank, "error": repr(e)})


def test_two_rank_broadcast_shard_gather_matches_unsharded_forward(oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    res.sort(key=lambda r: r["rank"])
    assert all("error" not in r for r in res), res
    assert res[0]["params_sum"] == res[1]["params_sum"] and res[0]["params_sum"] > 0  # rank 1 got rank 0's weights
    assert np.array_equal(res[0]["gathered"], res[1]["gathered"])
    assert res[0]["gathered"].shape[0] == 4
    assert res[0]["match_full"]
    assert all(r["ok_t"] for r in res)
    assert all(r["pipe_ok"] for r in res)


def test_equal_shard_size_rejects_uneven_batches():
    """The fixed-size gathers need equal shards on every rank: the check that guards them names the padded batch."""
    from boda_b200 import shard
    assert shard.equal_shard_size(256, 8) == 32 and shard.equal_shard_size(0, 4) == 0
    with pytest.raises(ValueError, match="pad it to 36"):
        shard.equal_shard_size(33, 4)
    assert [shard.shard_range(33, 4, r) for r in range(4)] == [(0, 9), (9, 18), (18, 27), (27, 33)]
