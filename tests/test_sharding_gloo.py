"""N > 1 path on CPU: two processes (gloo) shard a frame sequence the way bench.py's ranks do --
contiguous shards, per-shard LFSR state by jump-ahead, no data-path collective -- and the gathered
output must equal one continuous run. The grain itself is computed by the oracle here (there is no
GPU in this container); what is under test is the shard arithmetic and the shim's host-side
skip_frames bookkeeping, which needs no device."""
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


def _worker(rank, world, port, total, w, h, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tests.util import Oracle, load_golden, program_case, synth_frames
    from versatilefilmgrain_b200 import VfgsHw
    from versatilefilmgrain_b200.sharding import position_shard, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    G = load_golden()
    case = "fgs_sei_ff_test6.cfg|d10|420|g100"
    frames = synth_frames(total, w, h, "420", 10, seed=77)
    per = frames.size // total
    first, count = shard_range(total, rank, world)

    o = Oracle()
    st = program_case(o, G, case)
    epoch = [int(v) for v in st["lfsr"]]
    position_shard(o, epoch, first, w, h)
    # the product's host-side bookkeeping must land on the same registers (no GPU involved)
    hw = VfgsHw()
    hw.reset()
    program_case(hw, G, case)
    position_shard(hw, epoch, first, w, h)
    assert hw.get_lfsr()[2:] == o.get_lfsr()[2:], (rank, hw.get_lfsr(), o.get_lfsr())

    mine = o.add_grain_frames(frames[first * per:(first + count) * per].copy(), count, w, h, 0)
    # host gather in frame order (all_gather of padded shards; plumbing only)
    maxc = (total + world - 1) // world
    buf = torch.zeros(maxc * per, dtype=torch.int32)
    buf[: mine.size] = torch.from_numpy(mine.astype(np.int32))
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    if rank == 0:
        out = np.concatenate([parts[r].numpy()[: shard_range(total, r, world)[1] * per] for r in range(world)]).astype(np.uint16)
        ref = Oracle()
        program_case(ref, G, case)
        want = ref.add_grain_frames(frames, total, w, h, 0)
        q.put(bool(np.array_equal(out, want)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8])
def test_two_rank_shards_equal_continuous_run(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 256, 88, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_weighted_shards_partition_exactly():
    from versatilefilmgrain_b200.sharding import weighted_shard_ranges
    for total in (0, 1, 7, 768, 2400):
        for weights in ([1], [1, 1], [8.4, 8.4, 8.4, 8.4, 11.4, 11.4, 11.4, 11.4], [0.0, 3.0, 1.0]):
            spans = weighted_shard_ranges(total, weights)
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
    spans = weighted_shard_ranges(768, [8.4] * 4 + [11.4] * 4)
    assert [c for _, c in spans] == [81, 82, 81, 82, 110, 111, 110, 111] or sum(c for _, c in spans[4:]) > sum(c for _, c in spans[:4])
    assert weighted_shard_ranges(10, [0.0, 3.0, 1.0])[0] == (0, 0)


def test_shard_range_partitions_exactly():
    from versatilefilmgrain_b200.sharding import shard_range, steps_per_frame
    for total in (0, 1, 7, 64, 2400):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
    assert steps_per_frame(3840, 2160) == 32160 and steps_per_frame(1920, 1080) == 8040 and steps_per_frame(7680, 4320) == 129120
