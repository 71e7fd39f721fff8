"""Host-side module protocol mirroring the reference's `eqx.Module` conventions.

* a module is constructed as `Cls(num_spatial_dims, in_channels, out_channels, *, ..., key)` and called
  on ONE un-batched sample `x: (C_in, *spatial)` (README.md:74-76 of the reference); batching is
  `vmap(model)(x_batched)`;
* parameters are the array leaves in dataclass-field order (`tree_leaves`), same names and shapes
  as the reference pytree;
* the reverse pass is explicit (the reference relies on JAX autodiff; the FFI integration uses
  `jax.custom_vjp`): `_fwd(x, tape)` pushes what `_bwd(gy, tape)` needs.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_PRECISION = {"fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32}
_precision = "fp32"


def set_precision(mode: str):
    """'fp32' (parity mode, rel-L2 <= 1e-5) or 'tf32' (single-pass TF32 on tcgen05, rel-L2 <= 2e-3)."""
    global _precision
    if mode not in _PRECISION:
        raise ValueError(f"precision must be one of {list(_PRECISION)}, got {mode}")
    _precision = mode  # read when a call builds its descriptor / plan: the library itself holds no precision state


def get_precision() -> str:
    return _precision


def precision_code() -> int:
    return _PRECISION[_precision]


def default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def to_param(array: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(array, dtype=np.float32)).to(default_device())


class Tape:
    """Saved state of one forward pass + a pool of persistent device buffers (no allocation after
    the first step, so a train step can be captured in a CUDA graph)."""

    def __init__(self):
        self.stack = []
        self.pool = {}

    def buf(self, owner, tag, shape, zero=False):
        key = (id(owner), tag, tuple(int(s) for s in shape))
        t = self.pool.get(key)
        if t is None:
            t = torch.empty(key[2], dtype=torch.float32, device=default_device())
            self.pool[key] = t
            if zero:
                t.zero_()
        return t

    def push(self, item):
        self.stack.append(item)

    def pop(self):
        return self.stack.pop()


_CONV_WORKSPACE_OWNER = object()  # every conv call of one tape (= one stream) shares one scratch buffer


def conv_workspace(tape, nbytes):
    """Scratch for the tcgen05 conv passes (pdeqx_conv_workspace_bytes): one persistent buffer per tape (a tape serves one
    stream at a time, and the passes keep nothing in it between launches); a fresh one per call outside a tape."""
    if nbytes == 0:
        return None
    return new_buf(tape, _CONV_WORKSPACE_OWNER, "conv_ws", ((nbytes + 3) // 4,))


def new_buf(tape, owner, tag, shape):
    if tape is None:
        return torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device=default_device())
    return tape.buf(owner, tag, shape)


class Module:
    """Base class. Subclasses list their pytree fields (in reference order) in `_fields`;
    `_param_fields` names the fields that hold parameter tensors (or None)."""

    _fields: tuple = ()
    _param_fields: tuple = ()

    # ---- pytree -------------------------------------------------------------------------
    def named_leaves(self, prefix=""):
        """(path, owner, field) for every parameter tensor, in the reference's leaf order."""
        out = []
        for f in self._fields:
            v = getattr(self, f)
            out.extend(_named_leaves(v, prefix + f, self, f))
        return out

    def grad(self, field):
        g = getattr(self, "_grads", {}).get(field)
        if g is None:
            raise _lib.PdeqxError(f"no gradient buffer bound for {type(self).__name__}.{field}; use pdequinox_b200.train")
        return g

    # ---- call conventions ---------------------------------------------------------------
    num_spatial_dims: int

    def __call__(self, x, *, key=None):
        """One un-batched sample (C, *spatial) -> (C_out, *spatial_out)."""
        d = _spatial_dims_of(self)
        if x.ndim != d + 1:
            raise ValueError(f"Input to `{type(self).__name__}` needs to have rank {d + 1}, but input has shape {tuple(x.shape)}.")
        return self._fwd(_as_device(x)[None], None)[0]

    def _fwd(self, x, tape):  # x: (B, C, *S)
        raise NotImplementedError

    def _bwd(self, gy, tape, need_gx=True):
        raise NotImplementedError


def _spatial_dims_of(m):
    d = getattr(m, "num_spatial_dims", None)
    if d is None:
        raise _lib.PdeqxError(f"{type(m).__name__} does not define num_spatial_dims")
    return int(d)


def _named_leaves(v, path, owner, field):
    if isinstance(v, torch.Tensor):
        return [(path, owner, field)]
    if isinstance(v, Module):
        return v.named_leaves(path + ".")
    if isinstance(v, (list, tuple)):
        out = []
        for i, e in enumerate(v):
            if isinstance(e, Module):
                out.extend(e.named_leaves(f"{path}.{i}."))
        return out
    return []


def _as_device(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected a torch tensor or numpy array, got {type(x)}")
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.PdeqxError("pdequinox_b200 needs a CUDA device: there is no CPU fallback")
        x = x.cuda()
    return x.to(torch.float32).contiguous()


def vmap(module):
    """`jax.vmap(model)` of the reference's call convention: maps over the leading axis, which the
    kernels handle natively (the batch is a kernel dimension, weights are shared)."""
    d = _spatial_dims_of(module)

    def batched(x):
        if x.ndim != d + 2:
            raise ValueError(f"vmap({type(module).__name__}) expects rank {d + 2} input (batch, C, *spatial), got {tuple(x.shape)}")
        return module._fwd(_as_device(x), None)

    return batched


def tree_leaves(module):
    """Parameter tensors in the reference's pytree order (eqx.filter(model, eqx.is_array))."""
    return [getattr(o, f) for _, o, f in module.named_leaves()]


def tree_paths(module):
    return [p for p, _, _ in module.named_leaves()]


def count_parameters(module) -> int:
    """pdequinox/_utils.py:58-63."""
    return int(sum(t.numel() for t in tree_leaves(module)))
