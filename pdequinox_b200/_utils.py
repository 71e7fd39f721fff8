"""Host-side helpers mirrored from pdequinox/_utils.py: count_parameters (:58-63), dataloader (:66-122),
cycling_dataloader (:125-187), sum_receptive_fields (:240-267)."""
import math

from . import random as jr
from ._module import count_parameters  # noqa: F401


def sum_receptive_fields(receptive_fields):
    num_spatial_dims = len(receptive_fields[0])
    return tuple(
        (sum(r[i][0] for r in receptive_fields), sum(r[i][1] for r in receptive_fields))
        for i in range(num_spatial_dims))


# ---- minibatch iteration (host side; the arrays may be numpy arrays or torch tensors, on the host or the device) ----
def _leaves(tree):
    if isinstance(tree, dict):
        return [leaf for k in sorted(tree) for leaf in _leaves(tree[k])]
    if isinstance(tree, (tuple, list)):
        return [leaf for t in tree for leaf in _leaves(t)]
    return [] if tree is None else [tree]


def _map(fn, tree):
    if isinstance(tree, dict):
        return {k: _map(fn, v) for k, v in tree.items()}
    if isinstance(tree, tuple):
        return tuple(_map(fn, t) for t in tree)
    if isinstance(tree, list):
        return [_map(fn, t) for t in tree]
    return None if tree is None else fn(tree)


def _take(indices):
    def take(a):
        try:
            import torch
            if isinstance(a, torch.Tensor):
                return a[torch.as_tensor(indices, device=a.device, dtype=torch.long)]
        except ImportError:  # pragma: no cover
            pass
        return a[indices]
    return take


def dataloader(data, *, batch_size: int, key):
    """One epoch over `data` (an array or a tuple / list / dict of arrays sharing their leading sample axis) in shuffled
    minibatches: ceil(n_samples / batch_size) batches, the last one possibly smaller. Same contract as the reference's
    generator (pdequinox/_utils.py:66-122); the shuffle comes from this package's host PRNG, so the ORDER of the samples
    differs from a JAX run with the same seed."""
    sizes = [int(a.shape[0]) for a in _leaves(data)]
    if not sizes:
        raise ValueError("dataloader: `data` holds no array")
    if any(n != sizes[0] for n in sizes):
        raise ValueError("All arrays / PyTree leaves must have the same number of samples. (Leading array axis)")
    n_samples = sizes[0]
    order = jr.permutation(key, n_samples)
    for b in range(math.ceil(n_samples / batch_size)):
        yield _map(_take(order[b * batch_size:min((b + 1) * batch_size, n_samples)]), data)


def cycling_dataloader(data, *, batch_size: int, num_steps: int, key, return_info: bool = False):
    """`num_steps` minibatches, running as many freshly shuffled epochs of `dataloader` as that takes; a batch never mixes
    two epochs (pdequinox/_utils.py:125-187). With `return_info` every item is (batch, epoch index, batch index)."""
    step, epoch = 0, 0
    while True:
        key, subkey = jr.split(key)
        for batch_id, batch in enumerate(dataloader(data, batch_size=batch_size, key=subkey)):
            if step == num_steps:
                return
            yield (batch, epoch, batch_id) if return_info else batch
            step += 1
        epoch += 1
