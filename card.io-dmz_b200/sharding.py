"""Multi-GPU plumbing: frames are independent units, so the batch is sharded across ranks with NO collective on
the hot path; only the fixed-size per-frame digit strings are gathered to rank 0 at the end.
Sessions (8 consecutive frames) stay on one rank because scanner_add_frame's EMA couples them sequentially."""
import numpy as np

SESSION = 8
DIGIT_RECORD_BYTES = 32  # 16 digits, n_numbers, usable, upside_down, all_found, 12 bytes padding


def shard_range(n_frames, rank, world, session=SESSION):
    """Contiguous, session-aligned [lo, hi) of rank `rank`."""
    n_sessions = (n_frames + session - 1) // session
    lo_s = n_sessions * rank // world
    hi_s = n_sessions * (rank + 1) // world
    return min(lo_s * session, n_frames), min(hi_s * session, n_frames)


def digit_records(records):
    """(n, 32) uint8 digit-string records from a RECORD_DTYPE array."""
    n = records.shape[0]
    out = np.zeros((n, DIGIT_RECORD_BYTES), np.uint8)
    out[:, :16] = records["scores"].reshape(n, 16, 10).argmax(axis=2).astype(np.uint8)
    out[:, 16] = records["h_n_offsets"]
    out[:, 17] = records["usable"]
    out[:, 18] = records["upside_down"]
    out[:, 19] = records["all_found"]
    return out


def gather_digit_records(local, n_total, dist, rank, world, device=None):
    """Gather the per-rank (n_local, 32) arrays to rank 0 in global frame order (uneven shards allowed)."""
    import torch
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    mx = max(counts)
    buf = torch.zeros((mx, DIGIT_RECORD_BYTES), dtype=torch.uint8, device=device)
    buf[: local.shape[0]] = torch.as_tensor(local, device=device)
    outs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)])


# ---- the same records, built on the device from raw b200_frame_record bytes (bench.py's device-resident path)
REC_SCORES, REC_N_OFFSETS, REC_USABLE, REC_UPSIDE_DOWN, REC_ALL_FOUND = 84, 84 + 640, 84 + 640 + 48 + 28, 84 + 640 + 48 + 29, 80


def digit_records_torch(rec_bytes):
    """(n, 32) uint8 digit-string records from an (n, 808) uint8 tensor of b200_frame_record bytes (any device)."""
    import torch
    n = rec_bytes.shape[0]
    scores = rec_bytes[:, REC_SCORES:REC_SCORES + 640].contiguous().view(torch.float32).view(n, 16, 10)
    out = torch.zeros((n, DIGIT_RECORD_BYTES), dtype=torch.uint8, device=rec_bytes.device)
    out[:, :16] = scores.argmax(dim=2).to(torch.uint8)
    out[:, 16] = rec_bytes[:, REC_N_OFFSETS]
    out[:, 17] = rec_bytes[:, REC_USABLE]
    out[:, 18] = rec_bytes[:, REC_UPSIDE_DOWN]
    out[:, 19] = rec_bytes[:, REC_ALL_FOUND]
    return out


def gather_equal_shards_torch(local, dist, rank, world):
    """Equal-size shards already on the device: one gather to rank 0 (the only collective of the path).  Returns the
    (world * n, 32) tensor on rank 0, None elsewhere."""
    import torch
    outs = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, outs, dst=0)
    return torch.cat(outs) if rank == 0 else None
