"""Multi-GPU inference: frames are independent in eval mode, so the batch is
sharded contiguously across ranks -- one process per GPU, full weight replica, no data-path
collective -- and the only communication is one final gather of the (B_local, num_joints, 3) poses
(192 B/frame) over NCCL/NVLink.  The same code runs over gloo on CPU tensors for the host-logic tests."""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous near-equal shards: the first ``total % world`` ranks get one extra frame."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_poses(local_pose, total, group=None):
    """All ranks receive the (total, num_joints, 3) poses in global frame order.  Ragged shards are
    padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_pose
    if total % world == 0:            # equal shards (the benchmarked case): one collective, no padding or re-assembly
        assert local_pose.shape[0] == total // world
        out = local_pose.new_empty((total,) + tuple(local_pose.shape[1:]))
        dist.all_gather_into_tensor(out, local_pose.contiguous(), group=group)
        return out
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = local_pose.new_zeros((biggest,) + tuple(local_pose.shape[1:]))
    pad[: local_pose.shape[0]] = local_pose
    out = local_pose.new_empty((world * biggest,) + tuple(local_pose.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((world, biggest) + tuple(local_pose.shape[1:]))
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def gather_job_poses(local_poses, group=None):
    """The final gather of a sharded JOB: every rank has run its own K batches of B_local frames with no collective in
    between and holds ``(K, B_local, num_joints, 3)`` poses; all ranks receive ``(world, K, B_local, num_joints, 3)``
    (rank-major, equal shards).  One collective per job instead of one per batch: ranks never wait for each other while
    they compute, so per-step rank skew (power capping) does not accumulate into the job time."""
    world = dist.get_world_size(group)
    local_poses = local_poses.contiguous()
    out = local_poses.new_empty((world * local_poses.shape[0],) + tuple(local_poses.shape[1:]))
    dist.all_gather_into_tensor(out, local_poses, group=group)         # rank-major concatenation along dim 0
    return out.view((world,) + tuple(local_poses.shape))


class ShardedLifter:
    """predict_pose over a global batch: every rank passes the SAME global batch (or only its own shard
    with ``presharded=True``) and gets all poses back."""

    def __init__(self, net, group=None):
        self.net, self.group = net, group

    def predict_pose(self, frames, total=None, presharded=False):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if presharded:
            assert total is not None
            local = frames
        else:
            total = frames.shape[0]
            lo, hi = shard_bounds(total, world, rank)
            local = frames[lo:hi]
        pose = self.net.predict_pose(local) if local.shape[0] > 0 else \
            frames.new_zeros((0, self.net.num_joints, 3), dtype=torch.float32)
        return gather_poses(pose, total, self.group)
