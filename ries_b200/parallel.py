"""Multi-GPU host logic.  Searches for different targets are independent units (BASELINE config 5): targets
are dealt round-robin to ranks, no data-path collective is needed, and a job's throughput is the sum of
units over ranks divided by the slowest rank's time."""


def shard_targets(targets, rank, world):
    return [t for i, t in enumerate(targets) if i % world == rank]


def aggregate(dist, units, seconds, device="cuda"):
    import torch
    u = torch.tensor([units], dtype=torch.float64, device=device)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(u)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return u.item(), t.item()
