"""N>1 host logic (frame partition + result gather) with world_size 2 on the gloo backend (CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scp_b200 import partition


def test_round_robin_covers_all_frames_once():
    for n, w in ((10, 1), (10, 2), (7, 4), (3, 8), (1000, 8)):
        seen = sorted(i for r in range(w) for i in partition.frames_for_rank(n, r, w))
        assert seen == list(range(n))
        sizes = [len(partition.frames_for_rank(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = partition.frames_for_rank(n_frames, rank, world)
    table = partition.gather_frame_records(ids, [1000 + 10 * i for i in ids], [100 + i for i in ids], n_frames)
    if rank == 0:
        torch.save(table, out)
    dist.destroy_process_group()


def test_gather_over_gloo_world2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "t.pt")
    mp.spawn(_worker, args=(2, port, 9, out), nprocs=2, join=True)
    table = torch.load(out)
    assert table[:, 0].tolist() == [1000 + 10 * i for i in range(9)]
    assert table[:, 1].tolist() == [100 + i for i in range(9)]
    single = partition.gather_frame_records(range(9), [1000 + 10 * i for i in range(9)], [100 + i for i in range(9)], 9)
    assert torch.equal(single, table)                 # union of per-rank results == single-process result
    assert abs(partition.mean_bpp(table) - float((8.0 * table[:, 0].double() / table[:, 1].double()).mean())) < 1e-12
