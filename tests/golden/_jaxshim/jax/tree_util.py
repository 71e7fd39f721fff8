import numpy as _np


def tree_leaves(tree):
    """Array leaves of nested Modules / lists / tuples (attribute assignment order)."""
    if isinstance(tree, _np.ndarray):
        return [tree]
    if isinstance(tree, (list, tuple)):
        out = []
        for t in tree:
            out.extend(tree_leaves(t))
        return out
    if hasattr(tree, "__dict__") and not callable(getattr(tree, "__wrapped_fn__", None)):
        out = []
        for v in vars(tree).values():
            out.extend(tree_leaves(v))
        return out
    return []


def tree_map(f, tree, *rest):
    raise NotImplementedError("not needed for the hot path")
