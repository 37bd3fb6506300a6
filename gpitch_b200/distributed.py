"""Window sharding across GPUs (one process per GPU, torch.distributed).  Windows are independent GPs with their own
data, hyper-parameters and variational parameters (gpitch/separation.py:289-313), so the only exchange is an
all-gather of per-window results (ELBOs, LAPACK status, predictions); there is no gradient reduction."""
import torch
import torch.distributed as dist


def shard_windows(num_windows, world_size, rank):
    """Contiguous block of ceil(W / G) windows per rank (keeps overlap-add neighbours together)."""
    per = (num_windows + world_size - 1) // world_size
    lo = min(num_windows, rank * per)
    return lo, min(num_windows, lo + per)


def all_gather_windows(local, num_windows, group=None):
    """Gather per-window tensors [W_local, ...] from every rank into [W, ...] on every rank (ragged last shard is
    padded to the common block size for the collective and trimmed afterwards)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    per = (num_windows + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out, 0)[:num_windows]
