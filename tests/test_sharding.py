"""N > 1 plumbing on CPU: world_size-2 gloo run of the shard -> process -> gather-digit-strings path (the GPU
stage is replaced by the CPU oracle here; the collective and the shard arithmetic are what is under test)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from util import ROOT, deck_frames, load_pkg


def test_shard_ranges_are_session_aligned_and_cover():
    sh = __import__("importlib").import_module
    pkg = load_pkg()
    import importlib.util
    spec = importlib.util.spec_from_file_location("cardio_dmz_b200.sharding", os.path.join(ROOT, "card.io-dmz_b200", "sharding.py"))
    s = importlib.util.module_from_spec(spec); spec.loader.exec_module(s)
    for n in (1, 7, 8, 100, 1000, 100000):
        for world in (1, 2, 3, 8):
            ranges = [s.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a % 8 == 0


def _worker(rank, world, port, n, out_path):
    import importlib.util
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(ROOT, "card.io-dmz_b200", "sharding.py"))
    s = importlib.util.module_from_spec(spec); spec.loader.exec_module(s)
    from oracle.binding import Oracle
    lo, hi = s.shard_range(n, rank, world)
    recs = Oracle("port").process_frames(deck_frames(lo, hi - lo)) if hi > lo else np.zeros(0, Oracle("port").process_frames(deck_frames(0, 1)).dtype)
    got = s.gather_digit_records(s.digit_records(recs), n, dist, rank, world)
    if rank == 0:
        np.save(out_path, got)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(tmp_path, oracle):
    n = 20  # uneven: 3 sessions -> ranks get 8 and 12 frames
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(ROOT, "card.io-dmz_b200", "sharding.py"))
    s = importlib.util.module_from_spec(spec); spec.loader.exec_module(s)
    want = s.digit_records(oracle.process_frames(deck_frames(0, n)))
    assert np.array_equal(np.load(out), want)
