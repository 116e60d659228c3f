"""Frame sharding across GPUs (one process per GPU, no data-path collective).

Frames are independent once the LFSR state at a frame's first block-row is known, and that state is
a closed-form function of the frame index: the register at frame f is
the epoch register advanced f * (R - 1) * nb steps, R = ceil(H/16), nb = ceil(W/16). A rank that owns
frames [first, first + count) of a sequence therefore calls ``vfgs_b200_skip_frames(first, W, H)`` on
the epoch state and processes its shard; outputs are gathered by the host in frame order.
"""
from __future__ import annotations


def shard_range(total_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard of ``total_frames`` for ``rank``: (first frame, frame count). The first
    ``total_frames % world`` ranks get one extra frame."""
    if world < 1 or not (0 <= rank < world) or total_frames < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_frames, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def weighted_shard_ranges(total_frames: int, weights) -> list[tuple[int, int]]:
    """Contiguous shards whose sizes follow ``weights`` (e.g. the host<->device copy bandwidth each rank measured:
    the GPUs of one box do not all sit behind equally fast links, profiles/r02_pcie_matrix.json). Largest-remainder
    rounding; every rank's (first frame, frame count), in rank order, covering [0, total_frames) exactly."""
    w = [max(float(x), 0.0) for x in weights]
    if not w or total_frames < 0 or sum(w) <= 0:
        raise ValueError("bad shard request")
    exact = [total_frames * x / sum(w) for x in w]
    counts = [int(e) for e in exact]
    for i in sorted(range(len(w)), key=lambda i: exact[i] - counts[i], reverse=True)[: total_frames - sum(counts)]:
        counts[i] += 1
    out, first = [], 0
    for c in counts:
        out.append((first, c))
        first += c
    return out


def steps_per_frame(width: int, height: int) -> int:
    """LFSR steps between the first block-rows of two consecutive frames: (R - 1) * nb. The first
    block-row of a frame re-uses the state of the previous frame's last block-row (vfgs_hw.c:291-298
    latches only on y != 0), hence R - 1 and not R."""
    nb = (width + 15) // 16
    rows = (height + 15) // 16
    return (rows - 1) * nb


def position_shard(hw, epoch_regs, first_frame: int, width: int, height: int) -> None:
    """Put ``hw`` (anything with set_lfsr/skip_frames) at the state of frame ``first_frame`` of the
    sequence whose frame 0 starts from ``epoch_regs``."""
    hw.set_lfsr(epoch_regs)
    if first_frame:
        hw.skip_frames(first_frame, width, height)
