"""Multi-GPU plumbing around the hot path (one process per GPU, torch.distributed).

The path shards without any data-path collective: the loss by batch element (train.py:204-209),
the evaluation by image pair.  The only exchanges are the scalar loss reduction — ONE all-reduce
of a 4-float vector instead of the reference's four 4-byte all-reduces (train.py:92-96,
common.py:105-113) — and one gather of the per-pair metric rows to rank 0, which then applies
eval.py's aggregation (eval.py:231-266) so the table equals a single-process run.
"""
import numpy as np
import torch
import torch.distributed as dist

METRIC_NAMES = ('sd', 'ag', 'sf', 'mse', 'psnr', 'cc', 'scd', 'en', 'ce', 'mi',
                'qabf', 'nabf', 'labf', 'ssim', 'msssim', 'viff')   # eval.py:52-68


def reduce_loss_scalars(total, loss1, loss2, loss3, world_size, group=None):
    """reduce_value(x, world) of common.py:105-113 for the four logged scalars in one collective.
    Returns the four averaged 0-dim tensors (same values the reference's four calls produce)."""
    vec = torch.stack([total.detach().reshape(()), loss1.detach().reshape(()), loss2.detach().reshape(()),
                       loss3.detach().reshape(())])
    if world_size > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        vec = vec / world_size
    return vec[0], vec[1], vec[2], vec[3]


def reduce_loss_vector(vec, world_size, group=None):
    """The same reduction IN PLACE on the 4-float vector the fused kernel wrote (core.loss.last_loss_vector():
    [l_ssim, l_pixel, l_grad, total]): one all-reduce with the averaging done by the collective itself (NCCL AVG; SUM and
    one division on backends without it), no staging copy — the step is kernel + one 16-byte collective."""
    if world_size > 1:
        if dist.get_backend(group) == 'nccl':
            dist.all_reduce(vec, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
            vec.div_(world_size)
    return vec


def shard_indices(n_items, rank, world_size):
    """Pair i is evaluated by rank i % world_size (round-robin keeps shards within one item)."""
    return list(range(rank, n_items, world_size))


def gather_rows(local_rows, local_idx, n_items, rank, world_size, device=None, group=None):
    """local_rows: (n_local, K) float64 tensor, local_idx: the pair indices they belong to.
    Returns the (n_items, K) table in pair order on rank 0 (None elsewhere)."""
    local_rows = torch.as_tensor(local_rows, dtype=torch.float64)
    K = local_rows.shape[1] if local_rows.dim() == 2 and local_rows.numel() else len(METRIC_NAMES)
    if world_size == 1:
        table = torch.empty(n_items, K, dtype=torch.float64)
        table[torch.as_tensor(local_idx, dtype=torch.long)] = local_rows.cpu()
        return table
    dev = device if device is not None else local_rows.device
    cap = (n_items + world_size - 1) // world_size
    buf = torch.full((cap, K + 1), -1.0, dtype=torch.float64, device=dev)
    if len(local_idx):
        buf[:len(local_idx), 0] = torch.as_tensor(local_idx, dtype=torch.float64, device=dev)
        buf[:len(local_idx), 1:] = local_rows.to(dev)
    out = [torch.empty_like(buf) for _ in range(world_size)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    table = torch.empty(n_items, K, dtype=torch.float64)
    seen = 0
    for part in out:
        part = part.cpu()
        for row in part:
            i = int(row[0].item())
            if i >= 0:
                table[i] = row[1:]
                seen += 1
    assert seen == n_items, f'gathered {seen} of {n_items} rows'
    return table


def aggregate_columns(table, names=METRIC_NAMES):
    """eval.py:231-266 verbatim semantics: per metric [mean, std, v0, v1, ...] where the std is taken
    over the list that already has the mean inserted at its front (a quirk of the reference)."""
    cols = {}
    table = np.asarray(table, dtype=np.float64)
    for k, name in enumerate(names):
        vals = [float(v) for v in table[:, k]]
        vals.insert(0, np.mean(vals))
        vals.insert(1, np.std(vals))
        cols[name] = vals
    return cols


def evaluate_sharded(load_pair, n_pairs, rank=0, world_size=1, compute_rows=None, device=None, group=None):
    """Sharded eval.py loop (eval.py:176-225): `load_pair(i)` -> (img1, img2, imgf) as (1,1,H,W)
    float32 tensors; pairs of equal shape are batched into one launch per kernel family.
    `compute_rows(a, b, f)` -> (n, 16) float64 (defaults to the B200 suite; injectable for tests).
    Returns the aggregated columns on rank 0, None on the other ranks."""
    if compute_rows is None:
        from .core.metric import eval_metrics_batch

        def compute_rows(a, b, f):
            dev = device if device is not None else torch.device('cuda', torch.cuda.current_device())
            return eval_metrics_batch(a.to(dev), b.to(dev), f.to(dev))
    mine = shard_indices(n_pairs, rank, world_size)
    by_shape = {}
    for i in mine:
        a, b, f = load_pair(i)
        by_shape.setdefault(tuple(a.shape[-2:]), []).append((i, a, b, f))
    idx, rows = [], []
    for items in by_shape.values():
        a = torch.cat([it[1] for it in items]); b = torch.cat([it[2] for it in items]); f = torch.cat([it[3] for it in items])
        r = compute_rows(a, b, f)
        idx += [it[0] for it in items]
        rows.append(torch.as_tensor(r, dtype=torch.float64).cpu())
    local = torch.cat(rows) if rows else torch.empty(0, len(METRIC_NAMES), dtype=torch.float64)
    table = gather_rows(local, idx, n_pairs, rank, world_size, device=device, group=group)
    return aggregate_columns(table) if rank == 0 else None
