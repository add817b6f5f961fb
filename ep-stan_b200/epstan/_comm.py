"""Site-shard communicator: one process per GPU, torch.distributed underneath.

The only cross-GPU data dependency of the EP loop is the sum of the site
natural parameters (reference method.py:1073-1074) plus the all-sites
positive-definiteness flag (method.py:1145); see SURVEY 8e.  Everything else
here is bookkeeping (gathering the per-site host mirrors after a run).
"""

import numpy as np


class Comm(object):
    """Single-process communicator (one GPU holds every site)."""
    rank = 0
    size = 1

    def shard(self, K):
        """Contiguous block of sites owned by this rank: [k_begin, k_end)."""
        base, rem = divmod(K, self.size)
        k0 = self.rank * base + min(self.rank, rem)
        return k0, k0 + base + (1 if self.rank < rem else 0)

    def allreduce_sum_(self, tensor):
        return tensor

    def allreduce_scalar(self, value, op='sum'):
        return value

    def allgather_sites(self, local, K, axis):
        return local

    def allreduce_array(self, arr):
        return arr

    def barrier(self):
        pass

    def sync_device(self):
        pass


class TorchComm(Comm):
    """torch.distributed process group (NCCL over NVLink on the GPU box, gloo in
    the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def _device(self):
        import torch
        if self.backend == 'nccl':
            return torch.device('cuda', torch.cuda.current_device())
        return torch.device('cpu')

    def allreduce_sum_(self, tensor):
        """In-place sum of a (device or host) tensor over the ranks -- the single
        data-path collective per update attempt."""
        self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.group)
        return tensor

    def allreduce_scalar(self, value, op='sum'):
        import torch
        ops = {'sum': self._dist.ReduceOp.SUM, 'min': self._dist.ReduceOp.MIN, 'max': self._dist.ReduceOp.MAX}
        t = torch.tensor([float(value)], dtype=torch.float64, device=self._device())
        self._dist.all_reduce(t, op=ops[op], group=self.group)
        return float(t.item())

    def allreduce_array(self, arr):
        """Sum of a small host array over the ranks."""
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self._device())
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allgather_sites(self, local, K, axis):
        """Concatenate per-rank site blocks (split along `axis`) into the full array."""
        import torch
        dev = self._device()
        base, rem = divmod(K, self.size)
        cmax = base + (1 if rem else 0)
        counts = [base + (1 if rk < rem else 0) for rk in range(self.size)]
        loc = np.moveaxis(np.asarray(local, dtype=np.float64), axis, 0)
        padded = np.zeros((cmax,) + loc.shape[1:])
        padded[:loc.shape[0]] = loc
        mine = torch.from_numpy(padded).to(dev)
        pieces = [torch.empty_like(mine) for _ in range(self.size)]   # equal sizes (NCCL needs that)
        self._dist.all_gather(pieces, mine, group=self.group)
        full = np.concatenate([p.cpu().numpy()[:c] for p, c in zip(pieces, counts)], axis=0)
        return np.moveaxis(full, 0, axis)

    def barrier(self):
        self._dist.barrier(group=self.group)

    def sync_device(self):
        """Wait until the collectives issued so far (torch's current stream) have completed."""
        import torch
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()


def default_comm():
    """TorchComm when a multi-rank process group is initialised, else Comm."""
    try:
        import sys
        if 'torch' not in sys.modules:
            return Comm()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except Exception:
        pass
    return Comm()
