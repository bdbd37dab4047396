"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path -- sample-index sharding
plus one all-reduce of the float4 accumulators -- with the CPU oracle standing in for the
per-rank renderer (it takes the same first/count/stride triple as agpt_render)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partitions_samples(agpt):
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 5, 16, 17, 255):
            seen = []
            for rank in range(world):
                first, count, stride = agpt.multigpu.shard(10, total, rank, world)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(10, 10 + total)), (world, total)
    with pytest.raises(ValueError):
        agpt.multigpu.shard(0, 4, 2, 2)


def _worker(rank, world, port_file, out_file, cfg, level, W, H, total, md, da):
    sys.path.insert(0, ROOT)
    from tests.conftest import load_agpt
    agpt = load_agpt()
    from oracle import port_binding as port
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    hs = agpt.HostScene(cfg, level); ps = port.PortScene(hs)
    first, count, stride = agpt.multigpu.shard(0, total, rank, world)
    acc, _ = ps.render(W, H, first, count, md, da, threads=2, sample_stride=stride)
    t = torch.from_numpy(acc)
    agpt.multigpu.allreduce_accumulator(t)
    if rank == 0:
        np.save(out_file, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_split_equals_single(agpt, tmp_path):
    from oracle import port_binding as port
    if not port.available():
        pytest.fail("oracle/libagpt_oracle.so not built")
    cfg, level, W, H, total = 6, 2, 48, 27, 6
    d = agpt.config_defaults(cfg)
    rendezvous = str(tmp_path / "rdzv"); out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, rendezvous, out, cfg, level, W, H, total, d["max_depth"], d["depth_arg"]), nprocs=2, join=True)
    got = np.load(out)
    hs = agpt.HostScene(cfg, level); ps = port.PortScene(hs)
    want, _ = ps.render(W, H, 0, total, d["max_depth"], d["depth_arg"], threads=2)
    # only the fp32 summation order differs (<= 1e-5 relative)
    denom = np.maximum(np.abs(want), 1e-3)
    assert np.max(np.abs(got - want) / denom) <= 1e-5
