"""Frame-wise multi-GPU partition: frames are independent, so rank r of R encodes frames
r, r+R, ... in its own process/GPU with replicated weights; there is NO collective on the hot path.  The only
communication is the final gather of the per-frame (bytes, points) records."""
from typing import List, Sequence

import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Static round-robin: frame counts differ by at most one between ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_frames, world))


def gather_frame_records(local_ids: Sequence[int], local_bytes: Sequence[int], local_points: Sequence[int], n_frames: int,
                         device="cpu"):
    """All ranks end up with int64 [n_frames, 2] = (stream bytes, points) for every frame.  Works on gloo (CPU) and
    nccl (device='cuda'); a single all_reduce over a zero-initialised table, since each frame has exactly one owner."""
    table = torch.zeros((n_frames, 2), dtype=torch.int64, device=device)
    for i, b, p in zip(local_ids, local_bytes, local_points):
        table[i, 0] = int(b)
        table[i, 1] = int(p)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table


def gather_frame_table(local_ids: Sequence[int], local_rows, n_frames: int, n_cols: int, device="cpu"):
    """float64 [n_frames, n_cols] table of per-frame report values (bits, points, seconds, chamfer, psnr ...), every frame
    filled by its one owner rank and summed over ranks (zeros elsewhere), like ``gather_frame_records``."""
    table = torch.zeros((n_frames, n_cols), dtype=torch.float64, device=device)
    for i, row in zip(local_ids, local_rows):
        table[i] = torch.as_tensor([float(x) for x in row], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table


def init_from_env():
    """(rank, world, device index) of a ``torchrun`` launch (RANK / WORLD_SIZE / LOCAL_RANK), initialising the process group
    when world > 1: NCCL with one GPU per rank, gloo when the ranks have to share GPUs (fewer devices than ranks; the only
    traffic is the final report table, which then stays on the host)."""
    import os
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = torch.cuda.device_count() if torch.cuda.is_available() else 0
    dev = local % ndev if ndev else -1
    if dev >= 0:
        torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        if ndev >= world:
            dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo")
    return rank, world, dev


def table_device():
    return "cuda" if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl" else "cpu"


def mean_bpp(table: torch.Tensor) -> float:
    """encode.py:294: mean over frames of 8*bytes/points."""
    return float((8.0 * table[:, 0].double() / table[:, 1].double()).mean())
