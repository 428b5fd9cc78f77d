"""Host-side logic of the ray-sharded render (world_size 2, gloo, CPU): shard bounds + all-gather assembly.
The per-shard renderer is a deterministic stand-in (the CUDA path cannot run here); what is under test
is that the gathered image equals the single-process result for even and ragged ray counts."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diner_b200.multi_gpu import render_sharded, shard_bounds


def _fake_render(r):
    rgb = torch.stack((r[..., 0] * 2 + r[..., 3], r[..., 1] - r[..., 4], r[..., 2] * r[..., 5]), -1)
    return rgb, r[..., 6] + r[..., 7]


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    rays = torch.randn(2, n_rays, 8, generator=g)
    rgb, depth = render_sharded(_fake_render, rays)
    ref_rgb, ref_depth = _fake_render(rays)
    ok = bool(torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth))

    def packed(r, out, ray_offset):                       # the zero-copy contract: write (SB,n,4) into the gather slice
        c, d = _fake_render(r)
        out[..., :3] = c
        out[..., 3] = d
    for rr in (rays, rays[:1].contiguous()):              # SB = 2 and the SB = 1 fast path (views of the gather buffer)
        rgb_p, depth_p = render_sharded(packed, rr, packed=True)
        c, d = _fake_render(rr)
        ok = ok and bool(torch.equal(rgb_p, c) and torch.equal(depth_p, d))
        for w in ([3.0, 1.0], [1.0, 1e6]):                # weighted shards (speed-balanced sharding), incl. an almost empty one
            rgb_w, depth_w = render_sharded(packed, rr, packed=True, weights=w, align=4 if n_rays % 4 == 0 else 1)
            ok = ok and bool(torch.equal(rgb_w, c) and torch.equal(depth_w, d))
    q.put((rank, ok, tuple(rgb.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [64, 37, 1])
def test_sharded_render_matches_single(n_rays):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (2, n_rays, 3) for _, _, shape in res)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 262144):
        for world in (1, 2, 3, 8):
            for weights, align in ((None, 1), ([1.0 + 0.05 * r for r in range(world)], 1), ([1.0 + 0.05 * r for r in range(world)], 512)):
                spans = [shard_bounds(n, world, r, weights, align) for r in range(world)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
                assert all(hi - lo <= per for lo, hi, per in spans) and len({per for _, _, per in spans}) == 1
                if weights is not None and align > 1:
                    assert all(lo % align == 0 for lo, _, _ in spans)
    lo, hi, _ = shard_bounds(262144, 8, 7, [1.0] * 7 + [1.1], 512)
    assert hi - lo > 262144 // 8                             # the faster rank gets more rays


def test_bench_sample_indices_in_range():
    """bench.py's bounded CPU samples (cpu_baseline / --impl reference) must index inside the 512x512 ray list."""
    import bench
    for n in (64, 1024, 4096, 12288, bench.H * bench.W):
        p = bench.strided_pick(n)
        assert p.numel() == n and int(p.min()) >= 0 and int(p.max()) < bench.H * bench.W
        assert p.unique().numel() == n


def test_bench_traffic_summary_matches_workload():
    """roofline.traffic comes from the ncu summary captured on the SAME workload shape (profiles/*_traffic.json carry it)."""
    import bench
    d = bench.latest_traffic("dtu512")
    s = bench.latest_traffic("stress1024")
    assert d and d.get("workload", "dtu512") == "dtu512" and d.get("samples", 524288) == 524288
    assert s and s["workload"] == "stress1024" and s["samples"] == 16384 * 256
    assert bench.latest_traffic("no-such-workload") is None
