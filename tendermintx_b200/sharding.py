"""Multi-GPU plumbing: proofs are independent, so ranks shard the list of statements and exchange nothing on the data
path.  The only collective is the start-up broadcast of the circuit artefact (the `build` output /
preprocessed data) from rank 0; timing is reduced with MAX over ranks.  Works with NCCL (GPU) and gloo (CPU tests).
"""
import numpy as np


def assign_proofs(n_proofs, rank, world):
    """Proof j runs on rank j mod world."""
    return list(range(rank, n_proofs, world))


def broadcast_bytes(data, src=0, device="cpu"):
    """Broadcast a bytes object from `src` to every rank (length first, then the payload)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(data)
    rank = dist.get_rank()
    n = torch.tensor([len(data) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=device)
    if rank == src:
        buf.copy_(torch.from_numpy(np.frombuffer(bytes(data), dtype=np.uint8).copy()))
    dist.broadcast(buf, src)
    return buf.cpu().numpy().tobytes()


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (device timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]
