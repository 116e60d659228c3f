"""Frame sharding on REAL GPUs: one process per GPU, the CUDA kernels on every rank, each rank on a contiguous
shard that starts at a non-zero frame (positioned by GF(2) jump-ahead, vfgs_b200_skip_frames), no collective on
the data path, host gather in frame order. The gathered output must equal, byte for byte (SHA-256 and array
equality), (a) one continuous single-GPU run over all frames and (b) the oracle's continuous run.

This is the reference's state carry across frames (vfgs_hw.c:288-312: the LFSR registers run on from frame to
frame) reproduced without processing the frames in between. Needs >= 2 visible CUDA devices (gpurun --gpus 2);
skipped otherwise. The CPU-only counterpart of the shard arithmetic is tests/test_sharding_gloo.py."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _rank_worker(rank, world, case, total, w, h, od, outdir):
    """One process per GPU: bind the library to cuda:<rank>, program the golden hardware state, jump to the
    shard's first frame, run the shard through the device entry point, write the result to a file."""
    sys.path.insert(0, ROOT)
    import torch
    from tests.fixtures import load_golden, program_case, synth_frames
    from versatilefilmgrain_b200 import VfgsHw
    from versatilefilmgrain_b200.sharding import position_shard, shard_range

    torch.cuda.set_device(rank)
    G = load_golden()
    meta = G.cases[case]
    hw = VfgsHw(device=rank)
    hw.reset()
    st = program_case(hw, G, case)
    epoch = [int(v) for v in st["lfsr"]]
    frames = synth_frames(total, w, h, meta["fmt"], meta["depth"], seed=41)
    per = frames.size // total
    first, count = shard_range(total, rank, world)
    position_shard(hw, epoch, first, w, h)
    mine = frames[first * per:(first + count) * per]
    src = torch.from_numpy(mine.view(np.int16) if mine.dtype == np.uint16 else mine).cuda()
    out8 = od == 8 or meta["depth"] == 8
    dst = torch.zeros(mine.size, dtype=torch.uint8 if out8 else torch.int16, device="cuda")
    launches = hw.launch_count()
    hw.add_grain_frames_device(src, dst, count, w, h, od)
    torch.cuda.synchronize()
    assert hw.launch_count() > launches and torch.cuda.current_device() == rank
    got = dst.cpu().numpy()
    np.save(os.path.join(outdir, f"rank{rank}.npy"), got if out8 else got.view(np.uint16))
    np.save(os.path.join(outdir, f"lfsr{rank}.npy"), np.array(hw.get_lfsr(), dtype=np.uint64))


CASES = [
    # golden case, out depth: fast kernel; gather kernel (8 luma patterns) + fast chroma; chroma gather; fused 10->8
    ("fgs_afgs1_test1.cfg|d10|420|g100", 0),
    ("fgs_sei.cfg|d10|420|g100", 0),
    ("fgs_sei_ff_test5.cfg|d10|420|g100", 8),
    ("fgs_sei_ff_test4.cfg|d10|444|g150", 0),
]


@pytest.mark.parametrize("case,od", CASES)
def test_sharded_over_gpus_equals_continuous_run(tmp_path, case, od):
    ndev = _device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 CUDA devices (gpurun --gpus 2)")
    import torch
    import torch.multiprocessing as mp
    from tests.util import Oracle, load_golden, program_case, sha, synth_frames
    from versatilefilmgrain_b200 import VfgsHw
    from versatilefilmgrain_b200.sharding import shard_range

    G = load_golden()
    meta = G.cases[case]
    w, h = 704, 200  # H % 16 != 0: the (R - 1) frame advance and the short last stripe are both in play
    for world in sorted({2, min(ndev, 8)}):
        total = 8 * world + 3  # every shard has >= 8 frames; the first three ranks get one more
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_rank_worker, args=(r, world, case, total, w, h, od, str(tmp_path))) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
            assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
        gathered = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])

        frames = synth_frames(total, w, h, meta["fmt"], meta["depth"], seed=41)
        o = Oracle(); program_case(o, G, case)
        want = o.add_grain_frames(frames, total, w, h, od)
        assert np.array_equal(gathered, want), (case, world)

        # continuous single-GPU run of the same sequence (this process, cuda:0)
        hw = VfgsHw(device=0)
        hw.reset(); program_case(hw, G, case)
        src = torch.from_numpy(frames.view(np.int16) if frames.dtype == np.uint16 else frames).cuda()
        out8 = od == 8 or meta["depth"] == 8
        dst = torch.zeros(frames.size, dtype=torch.uint8 if out8 else torch.int16, device="cuda")
        hw.add_grain_frames_device(src, dst, total, w, h, od)
        torch.cuda.synchronize()
        one = dst.cpu().numpy()
        one = one if out8 else one.view(np.uint16)
        assert sha(gathered) == sha(one) == sha(want), (case, world)
        # the last rank ends on the registers of the continuous run; every rank ends where the next one started
        last = [int(v) for v in np.load(tmp_path / f"lfsr{world - 1}.npy")]
        assert last == hw.get_lfsr() == o.get_lfsr()
        for r in range(world - 1):
            first_next = shard_range(total, r + 1, world)[0]
            o2 = Oracle(); st = program_case(o2, G, case)
            o2.skip_frames(first_next, w, h)
            assert [int(v) for v in np.load(tmp_path / f"lfsr{r}.npy")][2:] == o2.get_lfsr()[2:]
