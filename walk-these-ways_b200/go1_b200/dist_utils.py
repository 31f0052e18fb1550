"""Host-side conventions of the multi-GPU learner (one process per GPU, envs sharded).

The CUDA path applies exactly these formulas on the device (normalize_adv_kernel, ppo_loss_kernel with
inv_count = 1/(M*world), adaptive_lr_kernel); they are stated once here so that the world_size-2 gloo tests can check
them against a single-process computation without a GPU."""
import torch
import torch.distributed as dist


def global_advantage_stats(local_adv, group=None):
    """(mean, unbiased std) of the advantages of ALL ranks from all-reduced (sum, sum of squares, count) in fp64
    (rollout_storage.py:87-88 semantics on the global batch)."""
    s = torch.tensor([local_adv.double().sum(), (local_adv.double() ** 2).sum(), float(local_adv.numel())], dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(s, group=group)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return float(mean), float(var.clamp(min=0).sqrt())


def allreduce_flat_grads(params, group=None):
    """SUM all-reduce of one flat gradient bucket; each rank scaled its loss by 1/(global batch), so the sum is the
    global-batch gradient (what clip_grad_norm_ then sees, ppo.py:157)."""
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if dist.is_initialized():
        dist.all_reduce(flat, group=group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
    return flat


def adaptive_lr(lr, kl, desired_kl=0.01, lo=1e-5, hi=1e-2):
    """ppo.py:124-132."""
    if kl > desired_kl * 2.0:
        return max(lo, lr / 1.5)
    if kl < desired_kl / 2.0 and kl > 0.0:
        return min(hi, lr * 1.5)
    return lr
