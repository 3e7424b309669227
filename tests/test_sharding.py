"""Multi-GPU host logic on CPU: the batch partition and the gather of the per-rank (offset, size) tables, with two real
processes over the gloo backend.  Each rank "encodes" its frames with the oracle (test infrastructure, allowed here) so
that rank 0 can check the reassembled batch byte for byte against a single-process encode."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sh = importlib.import_module("jpeg-encoder-decoder_b200.sharding")


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 2, 7, 8, 1023, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                first, count = sh.shard_range(n, r, world)
                assert count in (n // world, n // world + 1)
                covered.extend(range(first, first + count))
                assert all(sh.owner_of(f, n, world) == r for f in range(first, first + count))
            assert covered == list(range(n))


def test_local_offsets_and_single_rank_table():
    sizes = np.array([5, 0, 7, 3])
    assert sh.local_offsets(sizes).tolist() == [0, 5, 5, 12]
    t = sh.gather_tables(sizes, 4, 0, 1)
    assert t.tolist() == [[0, 0, 5], [0, 5, 0], [0, 5, 7], [0, 12, 3]]


def _worker(rank, world, port, n_frames, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_checkers
    fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        oracle = cpu_checkers.Oracle()
        first, count = sh.shard_range(n_frames, rank, world)
        streams = [oracle.encode(fr.noise_frame(first + i, 32, 16))["jpg"].tobytes() for i in range(count)]
        blob = b"".join(streams)                                        # the rank's compacted output
        table = sh.gather_tables(np.array([len(s) for s in streams]), n_frames, rank, world)
        with open(os.path.join(out_dir, f"rank{rank}.bin"), "wb") as f:
            f.write(blob)
        dist.barrier()
        if rank == 0:
            np.save(os.path.join(out_dir, "table.npy"), table)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_gloo_reassemble_the_batch(tmp_path, oracle, frames):
    import torch.multiprocessing as mp
    world, n_frames = 2, 7                                               # ragged: rank 0 owns 4 frames, rank 1 owns 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    table = np.load(tmp_path / "table.npy")
    blobs = [open(tmp_path / f"rank{r}.bin", "rb").read() for r in range(world)]
    assert table[:, 0].tolist() == [0, 0, 0, 0, 1, 1, 1]
    for f in range(n_frames):
        r, off, size = (int(v) for v in table[f])
        assert blobs[r][off:off + size] == oracle.encode(frames.noise_frame(f, 32, 16))["jpg"].tobytes(), f
    assert sum(len(b) for b in blobs) == int(table[:, 2].sum())
