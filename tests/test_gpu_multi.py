"""GPU (needs >= 2 devices, skipped otherwise): sample-index sharding on real hardware.

* G contexts in one process: agpt_render_multi + agpt_reduce_accum (this library's peer-memory kernels,
  rank-order sums) == the sum of the per-GPU stride renders bit for bit, and within 1e-5 of the one-GPU
  render (only the fp32 summation order differs);
* the NCCL route (AGPT_REDUCE=nccl) gives the same film;
* agpt_reduce_resolve (fused sum + CopyToSurface) == the reference's lin2rgb / rgb2uint on that sum;
* CudaPathTracer over a device list (host mirror) == the C-ABI calls;
* G processes, one rank per GPU: accumulators exchanged through CUDA IPC handles, same kernels.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG, LEVEL, W, H, SPP = 3, 3, 320, 180, 8


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def world(agpt):
    n = agpt.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return min(n, 4)


@pytest.fixture(scope="module")
def shards(agpt, world):
    """Per-rank films rendered one after the other on GPU 0 with the stride a rank would use, and the whole render."""
    d = agpt.config_defaults(CFG)
    hs = agpt.HostScene(CFG, LEVEL)
    ctx = agpt.Context(0)
    hs.upload(ctx); ctx.set_film(W, H)
    films = []
    for g in range(world):
        ctx.clear()
        ctx.render(g, (SPP - g + world - 1) // world, d["max_depth"], d["depth_arg"], sample_stride=world)
        films.append(ctx.read_accum())
    ctx.clear(); ctx.render(0, SPP, d["max_depth"], d["depth_arg"])
    whole = ctx.read_accum()
    ctx.close()
    total = films[0].copy()
    for f in films[1:]:
        total = total + f                      # rank order, fp32
    return hs, d, films, total, whole


def make_group(agpt, hs, world):
    ctxs = [agpt.Context(g) for g in range(world)]
    for c in ctxs:
        hs.upload(c); c.set_film(W, H); c.clear()
    return agpt.Group(ctxs)


@pytest.mark.parametrize("route", ["p2p", "nccl"])
def test_sharded_render_and_reduce(agpt, world, shards, route):
    hs, d, films, total, whole = shards
    group = make_group(agpt, hs, world)
    old = os.environ.get("AGPT_REDUCE")
    os.environ["AGPT_REDUCE"] = route
    try:
        group.render(0, SPP, d["max_depth"], d["depth_arg"])
        for g, c in enumerate(group.contexts):
            assert np.array_equal(bits(c.read_accum()), bits(films[g])), f"rank {g}: sharded render differs from the stride render"
        group.reduce()
        got = [c.read_accum() for c in group.contexts]
    finally:
        if old is None:
            os.environ.pop("AGPT_REDUCE", None)
        else:
            os.environ["AGPT_REDUCE"] = old
    st = group.contexts[0].stats()
    assert st.reduce_path == (1 if route == "p2p" else 2)
    for g in range(1, world):
        assert np.array_equal(bits(got[0]), bits(got[g])), "ranks disagree after the all-reduce"
    if route == "p2p" or world == 2:
        assert np.array_equal(bits(got[0]), bits(total)), "rank-order sum"
    denom = np.maximum(np.abs(whole), 1e-3)
    rel = float(np.max(np.abs(got[0] - whole) / denom))
    print(f"{route}: {world}-GPU film vs 1-GPU film: max rel diff {rel:.2e}, reduce {st.ms_reduce:.3f} ms")
    assert rel <= 1e-5
    for c in group.contexts:
        c.close()


def test_reduce_to_root_and_fused_resolve(agpt, ref, world, shards):
    hs, d, films, total, whole = shards
    group = make_group(agpt, hs, world)
    group.render(0, SPP, d["max_depth"], d["depth_arg"])
    rgb = group.reduce_resolve(SPP, keep_sum=True)
    assert np.array_equal(rgb, ref.resolve(total, SPP)), "fused reduce + CopyToSurface bytes"
    assert np.array_equal(bits(group.contexts[0].read_accum()), bits(total)), "keep_sum leaves the summed film on GPU 0"
    assert np.array_equal(bits(group.contexts[1].read_accum()), bits(films[1])), "the other accumulators are untouched"
    # root-only reduce from fresh shards
    for g, c in enumerate(group.contexts):
        c.write_accum(films[g])
    group.reduce(root=0)
    assert np.array_equal(bits(group.contexts[0].read_accum()), bits(total))
    assert np.array_equal(bits(group.contexts[1].read_accum()), bits(films[1]))
    for c in group.contexts:
        c.close()


def test_host_mirror_over_a_device_list(agpt, ref, world, shards):
    hs, d, films, total, whole = shards
    tr = agpt.HostTracer(d["max_depth"], devices=list(range(world)))
    acc = np.zeros((H, W, 4), np.float32)
    tr.render(hs, W, H, acc, 0, SPP, d["depth_arg"])
    assert np.array_equal(bits(acc), bits(total))
    # successive renders accumulate like successive Ticks; the fused resolve shows the film so far
    acc2 = np.zeros((H, W, 4), np.float32)
    tr.render(hs, W, H, acc2, 0, SPP // 2, d["depth_arg"])
    rgb = tr.render_resolve(hs, W, H, acc2, SPP // 2, SPP // 2, SPP - SPP // 2, d["depth_arg"])
    assert np.array_equal(rgb, ref.resolve(acc2, SPP))
    denom = np.maximum(np.abs(whole), 1e-3)
    assert float(np.max(np.abs(acc2 - whole) / denom)) <= 1e-5
    tr.close()


def _ipc_rank(rank, world, conns, barrier, out):
    import sys
    sys.path.insert(0, ROOT)
    from tests.conftest import load_agpt
    agpt = load_agpt()
    d = agpt.config_defaults(CFG)
    hs = agpt.HostScene(CFG, LEVEL)
    ctx = agpt.Context(rank)
    hs.upload(ctx); ctx.set_film(W, H); ctx.clear()
    ctx.render(rank, (SPP - rank + world - 1) // world, d["max_depth"], d["depth_arg"], sample_stride=world)
    mine = ctx.read_accum()
    # exchange the 64-byte handles (any transport will do: here pipes through the parent)
    conns[rank].send(ctx.accum_ipc_handle())
    handles = conns[rank].recv()
    ctx.open_peer_accums(rank, handles)
    barrier.wait()                                   # every rank has rendered
    rgb = ctx.reduce_resolve_peers(SPP) if rank == 0 else None
    barrier.wait()                                   # the root has read everybody
    ctx.allreduce_accum_peers()
    barrier.wait()                                   # every slice has been written everywhere
    summed = ctx.read_accum()
    barrier.wait()
    ctx.close_peer_accums()
    out.put((rank, mine, summed, rgb))
    barrier.wait()
    ctx.close()


def test_ranks_in_separate_processes_over_cuda_ipc(agpt, ref, world, shards):
    hs, d, films, total, whole = shards
    n = 2
    mpc = mp.get_context("spawn")
    parent, child = zip(*[mpc.Pipe() for _ in range(n)])
    barrier = mpc.Barrier(n)
    out = mpc.Queue()
    procs = [mpc.Process(target=_ipc_rank, args=(r, n, child, barrier, out)) for r in range(n)]
    for p in procs:
        p.start()
    handles = [parent[r].recv() for r in range(n)]
    for r in range(n):
        parent[r].send(handles)
    results = {}
    for _ in range(n):
        r, mine, summed, rgb = out.get(timeout=300)
        results[r] = (mine, summed, rgb)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = results[0][0] + results[1][0]
    for r in range(n):
        assert np.array_equal(bits(results[r][1]), bits(want)), f"rank {r}: all-reduce over IPC peers"
    assert np.array_equal(results[0][2], ref.resolve(want, SPP)), "root's fused reduce + resolve over IPC peers"
