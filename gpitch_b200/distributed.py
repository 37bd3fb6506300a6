"""Window sharding across GPUs (one process per GPU, torch.distributed).  Windows are independent GPs with their own
data, hyper-parameters and variational parameters (gpitch/separation.py:289-313), so the only exchange is an
all-gather of per-window results (ELBOs, LAPACK status, predictions); there is no gradient reduction."""
import os

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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs lookup by PCI address), so that pinned
    host buffers allocated afterwards are first-touched on that node and H2D/D2H traffic of the host-buffer path
    (BatchedPdgp.elbo_host) does not cross the socket interconnect when 8 ranks stream at once.
    Returns the node id, or None when the topology is not exposed (VMs report -1) -- then nothing is changed."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        addr = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % addr) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, AttributeError, ValueError):
        return None
