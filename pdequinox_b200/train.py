"""Batch-sharded training step: MSE loss, explicit reverse pass, gradient all-reduce, fused Adam.

The reference has no training code of its own; users write (README.md:74-86)
    loss = mean((vmap(model)(x) - y)**2); eqx.filter_value_and_grad; optax.adam(3e-4); eqx.apply_updates
and no multi-device code exists at all. This module is that step for one rank of a
data-parallel job: parameters, gradients and Adam moments live in flat fp32 arenas in the reference's
pytree leaf order; each rank runs the forward and reverse pass on its shard of the batch, the gradient
arena is summed across ranks with ONE NCCL all-reduce over NVLink (torch.distributed is the plumbing),
and the 1/world_size scaling is fused into the Adam kernel.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._module import Tape, _as_device, default_device


class _DeviceBuffer:
    """A float32 buffer from pdeqx_dp_alloc (a cudaMalloc allocation of its own, which CUDA IPC can export -- a slice of
    torch's caching allocator cannot), wrapped as a torch tensor through __cuda_array_interface__."""

    def __init__(self, n_floats):
        import ctypes as C
        self.ptr = C.c_void_p()
        _lib.check(_lib.lib().pdeqx_dp_alloc(max(int(n_floats), 1) * 4, C.byref(self.ptr)))
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (self.ptr.value, False), "version": 3}
        self.tensor = torch.as_tensor(self, device=default_device())

    def handle(self):
        import ctypes as C
        h = (C.c_char * 64)()
        _lib.check(_lib.lib().pdeqx_dp_export(self.ptr, h))
        return bytes(h.raw)


class ParameterArena:
    """Re-homes every parameter of `model` into one flat buffer (16-byte aligned slices, pytree order)
    and gives every parameter a gradient view in a second flat buffer. `exportable`: hold both arenas in allocations that
    CUDA IPC can export to the other ranks of the node (the peer-memory exchange of `PeerExchange`)."""

    def __init__(self, model, exportable=False, group=None):
        self.model = model
        self.symmetric = None  # (params handle, grads handle) of torch symmetric memory
        self.entries = []  # (path, owner, field, offset, numel, shape)
        off = 0
        for path, owner, field in model.named_leaves():
            t = getattr(owner, field)
            self.entries.append((path, owner, field, off, t.numel(), tuple(t.shape)))
            off += (t.numel() + 3) // 4 * 4
        self.size = off
        dev = default_device()
        self.buffers = None
        # NVSwitch multicast (multimem.ld_reduce / multimem.st in the exchange kernel; torch's symmetric memory provides the
        # mapping). Measured per exchange: 222 vs 145 us at 2 ranks (nothing for the switch to reduce), 208 vs 196 us at 4,
        # 203 vs 253 us at 8 -> on from 8 ranks (PDEQX_DP_MULTICAST=0 / 1 forces it off / on)
        mc = os.environ.get("PDEQX_DP_MULTICAST", "auto")
        world = torch.distributed.get_world_size(group) if exportable else 1
        sym = _symmetric_arenas(off, group) if (exportable and (mc == "1" or (mc == "auto" and world >= 8))) else None
        if sym is not None:
            self.params, self.grads = sym[0], sym[1]
            self.symmetric = tuple(sym[2])
        elif exportable:
            self.buffers = (_DeviceBuffer(off), _DeviceBuffer(off))
            self.params, self.grads = self.buffers[0].tensor, self.buffers[1].tensor
        else:
            self.params = torch.zeros(off, dtype=torch.float32, device=dev)
            self.grads = torch.zeros(off, dtype=torch.float32, device=dev)
        for path, owner, field, o, n, shape in self.entries:
            view = self.params[o:o + n].view(shape)
            view.copy_(getattr(owner, field))
            setattr(owner, field, view)
            if not hasattr(owner, "_grads"):
                owner._grads = {}
            owner._grads[field] = self.grads[o:o + n].view(shape)

    def named_grads(self):
        return {p: self.grads[o:o + n].view(shape) for p, _, _, o, n, shape in self.entries}

    def named_params(self):
        return {p: self.params[o:o + n].view(shape) for p, _, _, o, n, shape in self.entries}

    def load(self, named):
        """Overwrite parameters from a {path: array} mapping (tests: identical weights as the oracle)."""
        import numpy as np
        for p, _, _, o, n, shape in self.entries:
            if p in named:
                a = torch.from_numpy(np.ascontiguousarray(named[p], dtype=np.float32)).reshape(shape)
                self.params[o:o + n].view(shape).copy_(a)


def allreduce_gradients_(grads, loss, world, group=None):
    """The one exchange step of the data-parallel path: SUM the flat gradient arena (and the scalar loss) over
    the ranks. The 1/world scaling of the gradients is left to the Adam kernel (grad_scale); the loss is
    averaged here. Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
    if world > 1:
        torch.distributed.all_reduce(grads, group=group)
        torch.distributed.all_reduce(loss, group=group)
        loss.div_(world)
    return grads, loss


def _symmetric_arenas(n_floats, group):
    """(params, grads, handles) as torch symmetric-memory tensors (peer-mapped AND, on an NVSwitch fabric, multicast-mapped
    on every rank), or None -- on EVERY rank -- if this torch build / fabric cannot provide them on some rank. Plumbing only:
    the kernels just get pointers."""
    dist = torch.distributed
    group = group if group is not None else dist.group.WORLD
    dev = default_device()
    tensors, why = None, ""
    try:
        import torch.distributed._symmetric_memory as symm
        tensors = [symm.empty(int(n_floats), dtype=torch.float32, device=dev) for _ in range(2)]
        for t in tensors:
            t.zero_()
    except Exception as exc:  # noqa: BLE001 - any failure means "not available": the IPC path needs nothing of this
        why = f"{type(exc).__name__}: {exc}"
    flag = torch.tensor([1 if tensors is not None else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # the rendezvous below is collective: all ranks or none
    if not bool(flag.item()):
        if why:
            import warnings
            warnings.warn(f"symmetric memory unavailable ({why}); the peer exchange uses CUDA IPC without multicast")
        return None
    handles = [symm.rendezvous(t, group) for t in tensors]
    return tensors[0], tensors[1], handles


class PeerExchange:
    """This rank's view of every rank's parameter arena, gradient arena and signal pad: what pdeqx_dp_adam_step
    (reduce-scatter + Adam on the rank's shard + all-gather, one kernel over NVLink peer memory) needs. The arenas are
    either torch symmetric-memory tensors (`ParameterArena(exportable="symmetric")`: peer pointers and, on NVSwitch, the
    multicast addresses that let the switch do the reduction and the broadcast) or pdeqx_dp_alloc buffers exchanged as CUDA
    IPC handles through torch.distributed. The signal pads always travel as IPC handles. `ok` is False -- on EVERY rank --
    if any rank could not map a peer's memory (ranks on different nodes, no peer access); the caller then keeps the NCCL
    all-reduce."""

    def __init__(self, arena, group=None):
        import ctypes as C
        import socket
        dist = torch.distributed
        L = _lib.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.pad = _DeviceBuffer(L.pdeqx_dp_pad_bytes() // 4)
        self.imported = []
        symmetric = arena.symmetric is not None
        mine = (socket.gethostname(), None if symmetric else arena.buffers[0].handle(),
                None if symmetric else arena.buffers[1].handle(), self.pad.handle())
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.peers = _lib.DpPeers()
        self.peers.world, self.peers.rank = self.world, self.rank
        ok = self.world <= _lib.MAX_RANKS and all(e[0] == mine[0] for e in everyone)
        self.multicast = False
        if ok:
            for r, (_, hp, hg, hs) in enumerate(everyone):
                handles = (hs,) if symmetric else (hp, hg, hs)
                if r == self.rank:
                    ptrs = [self.pad.ptr.value] if symmetric else [arena.buffers[0].ptr.value, arena.buffers[1].ptr.value, self.pad.ptr.value]
                else:
                    ptrs = []
                    for h in handles:
                        out = C.c_void_p()
                        if ok and L.pdeqx_dp_import((C.c_char * 64).from_buffer_copy(h), C.byref(out)) == 0:
                            self.imported.append(out)
                            ptrs.append(out.value)
                        else:
                            ok = False
                            ptrs.append(None)
                if symmetric:
                    hp_, hg_ = arena.symmetric
                    self.peers.params[r], self.peers.grads[r], self.peers.pads[r] = hp_.buffer_ptrs[r], hg_.buffer_ptrs[r], ptrs[0]
                else:
                    self.peers.params[r], self.peers.grads[r], self.peers.pads[r] = ptrs
            if symmetric:
                hp_, hg_ = arena.symmetric
                if hp_.multicast_ptr and hg_.multicast_ptr:
                    self.peers.mc_params, self.peers.mc_grads = hp_.multicast_ptr, hg_.multicast_ptr
                    self.multicast = True
        flag = torch.tensor([1 if ok else 0, 1 if self.multicast else 0], dtype=torch.int32, device=default_device())
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(flag[0].item())
        if self.multicast and not bool(flag[1].item()):  # all ranks or none
            self.peers.mc_params = self.peers.mc_grads = None
            self.multicast = False
        self.shard_floats = int(L.pdeqx_dp_shard_floats(arena.size, self.world))

    def check(self):
        """Raises if a peer rank did not arrive at one of the exchange's flag barriers (synchronises the stream)."""
        import ctypes as C
        _lib.check(_lib.lib().pdeqx_dp_check(C.byref(self.peers), _lib.stream()))

    def close(self):
        for p in self.imported:
            _lib.lib().pdeqx_dp_unimport(p)
        self.imported = []


class TrainStep:
    """One Adam step on mse(vmap(model)(x), y); `process_group` = a torch.distributed process group (or None = the default
    one when initialised). exchange: how the ranks' gradients meet -- "peer": pdeqx_dp_adam_step over NVLink peer memory
    (needs NCCL ranks of one node, one GPU each); "allreduce": one NCCL / gloo all-reduce of the gradient arena + a replicated
    Adam pass; "auto" (default): "peer" where it is available."""

    def __init__(self, model, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8, process_group=None, distributed=None, overlap=False,
                 exchange="auto"):
        self.model = model
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        world = torch.distributed.get_world_size(process_group) if distributed else 1
        want_peer = (distributed and world > 1 and exchange in ("auto", "peer") and not overlap and torch.cuda.is_available()
                     and torch.distributed.get_backend(process_group) == "nccl")
        if exchange == "peer" and not want_peer:
            raise _lib.PdeqxError("exchange='peer' needs an initialised NCCL process group with more than one rank")
        self.arena = ParameterArena(model, exportable=want_peer, group=process_group)
        self.peer = None
        if want_peer:
            self.peer = PeerExchange(self.arena, process_group)
            if not self.peer.ok:
                if exchange == "peer":
                    raise _lib.PdeqxError("exchange='peer': a rank could not map a peer's memory (different nodes / no peer access)")
                self.peer.close()
                self.peer = None
        dev = default_device()
        n_moments = self.peer.shard_floats if self.peer is not None else self.arena.size
        self.mu = torch.zeros(n_moments, dtype=torch.float32, device=dev)  # peer exchange: the rank's shard only
        self.nu = torch.zeros(n_moments, dtype=torch.float32, device=dev)
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.count = 0
        self.adam_state = torch.zeros(4, dtype=torch.float32, device=dev)  # device-side step counter + bias corrections
        self.graph = None
        self.tape = Tape()
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.partials = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.pg = process_group
        self.distributed = distributed
        self.world = world
        # Bucketed exchange (overlap=True): a model that reports finished pytree prefixes during its reverse pass
        # (Sequential) gets each bucket's all-reduce enqueued on a side stream right away. Off by default: measured on
        # 2 x B200 it is SLOWER (6.14 vs 5.94 ms per step) -- the NCCL kernel occupies a few SMs, and the persistent
        # one-CTA-per-SM tcgen05 kernels (static unit striding) then wait for their last CTAs to get an SM.
        self._reduced = []  # (offset, numel) ranges already handed to the communicator in this step
        self.comm_stream = None
        if overlap and self.world > 1 and self.arena.params.is_cuda:
            self.comm_stream = torch.cuda.Stream()
            model._grad_ready_hook = self._reduce_prefix

    def _prefix_range(self, prefix):
        lo, hi = None, None
        for p, _, _, o, n, _ in self.arena.entries:
            if p.startswith(prefix):
                lo = o if lo is None else min(lo, o)
                hi = o + (n + 3) // 4 * 4 if hi is None else max(hi, o + (n + 3) // 4 * 4)
        return lo, hi

    def _reduce_prefix(self, prefix):
        lo, hi = self._prefix_range(prefix)
        if lo is None:
            return
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self.comm_stream.wait_event(ready)
        with torch.cuda.stream(self.comm_stream):
            torch.distributed.all_reduce(self.arena.grads[lo:hi], group=self.pg)
        self._reduced.append((lo, hi))

    def loss_and_grad(self, x, y):
        """Local-shard loss (device scalar) and gradients into the arena (not yet reduced)."""
        x, y = _as_device(x), _as_device(y)
        tape = self.tape
        tape.stack.clear()
        tape.fold_lifting = True  # no gradient w.r.t. the network input is taken: a pointwise lifting may be folded away
        self._reduced = []
        pred = self.model._fwd(x, tape)
        if pred.shape != y.shape:
            raise ValueError(f"prediction {tuple(pred.shape)} and target {tuple(y.shape)} differ in shape")
        gpred = tape.buf(self, "gpred", pred.shape)
        _lib.check(_lib.lib().pdeqx_mse_loss(pred.numel(), _lib.ptr(pred), _lib.ptr(y), 1.0, _lib.ptr(self.loss),
                                             _lib.ptr(gpred), _lib.ptr(self.partials), _lib.stream()))
        self.model._bwd(gpred, tape, False)
        return self.loss

    def apply_gradients(self):
        """All-reduce (sum) the gradient arena over the data-parallel group, then Adam with grad/world -- or, with the
        peer-memory exchange, both in one kernel (reduce-scatter + Adam on this rank's shard + all-gather)."""
        if self.peer is not None:
            import ctypes as C
            self.count += 1
            _lib.check(_lib.lib().pdeqx_dp_adam_step(C.byref(self.peer.peers), self.arena.size, _lib.ptr(self.mu), _lib.ptr(self.nu),
                                                     self.lr, self.b1, self.b2, self.eps, _lib.ptr(self.adam_state),
                                                     _lib.ptr(self.loss), _lib.stream()))
            return
        if self.distributed:
            covered = sum(hi - lo for lo, hi in self._reduced)
            if self.comm_stream is not None and covered == self.arena.size:
                # every bucket is already in flight on the side stream: join it, then only the scalar loss is left
                torch.cuda.current_stream().wait_stream(self.comm_stream)
                torch.distributed.all_reduce(self.loss, group=self.pg)
                self.loss.div_(self.world)
            elif covered == 0:
                allreduce_gradients_(self.arena.grads, self.loss, self.world, self.pg)
            else:
                raise _lib.PdeqxError("gradient buckets cover only part of the arena")
            self._reduced = []
        self.count += 1
        # the update count lives on the device (incremented by the call): the same launch sequence can be replayed
        # from a captured CUDA graph
        _lib.check(_lib.lib().pdeqx_adam_step_dev(self.arena.size, _lib.ptr(self.arena.params), _lib.ptr(self.arena.grads),
                                                  _lib.ptr(self.mu), _lib.ptr(self.nu), self.lr, self.b1, self.b2,
                                                  self.eps, _lib.ptr(self.adam_state), 1.0 / self.world, _lib.stream()))

    def __call__(self, x, y):
        """x, y: this rank's shard (B_local, C, *S). Returns the (global mean) loss as a device scalar."""
        self.loss_and_grad(x, y)
        self.apply_gradients()
        return self.loss


class GraphedTrainStep:
    """The whole train step (forward, reverse pass, gradient all-reduce, Adam) captured ONCE into a CUDA graph and
    replayed: ~100 kernel launches and their host-side descriptor work collapse into one graph launch, which is what
    keeps small per-GPU batches (8 samples per GPU at 8 GPUs) from being launch-bound. Inputs are copied into static
    device buffers; every intermediate lives in the step's persistent tape buffers, so replays allocate nothing.

    Construction is free of side effects on the optimisation: the `warmup` eager steps that allocate the tape buffers
    before the capture run on the example batch, and the parameters, Adam moments, the device-side update counter and
    `step.count` are rewound to their values at entry afterwards (capturing itself executes nothing)."""

    def __init__(self, step: "TrainStep", x_example, y_example, warmup=3):
        self.step = step
        self.x = _as_device(x_example).clone()
        self.y = _as_device(y_example).clone()
        entry = (step.arena.params.clone(), step.mu.clone(), step.nu.clone(), step.adam_state.clone(), step.count)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream: allocates every tape buffer, sets kernel attributes
            for _ in range(max(warmup, 1)):
                step(self.x, self.y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            step(self.x, self.y)
        step.arena.params.copy_(entry[0])
        step.mu.copy_(entry[1])
        step.nu.copy_(entry[2])
        step.adam_state.copy_(entry[3])
        step.count = entry[4]
        self.loss = step.loss
        # input pipeline: the NEXT batch is copied host -> device on a side stream while the current step runs
        self._copy_stream = torch.cuda.Stream()
        self._staged = None
        self._x_stage = torch.empty_like(self.x)
        self._y_stage = torch.empty_like(self.y)
        self._copied = torch.cuda.Event()
        self._consumed = torch.cuda.Event()

    def prefetch(self, x_host, y_host):
        """Start copying the next batch (pinned host tensors) into the staging buffers; returns immediately."""
        cs = self._copy_stream
        cs.wait_event(self._consumed)  # the previous staged batch has been moved into the graph's input buffers
        with torch.cuda.stream(cs):
            self._x_stage.copy_(x_host, non_blocking=True)
            self._y_stage.copy_(y_host, non_blocking=True)
            self._copied.record(cs)
        self._staged = True

    def __call__(self, x=None, y=None):
        """x, y: this rank's shard (device or pinned host tensors); None: use the batch staged by `prefetch`, or (if
        nothing is staged) the resident inputs."""
        cur = torch.cuda.current_stream()
        if x is not None:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y, non_blocking=True)
        elif self._staged:
            cur.wait_event(self._copied)
            self.x.copy_(self._x_stage, non_blocking=True)
            self.y.copy_(self._y_stage, non_blocking=True)
            self._consumed.record(cur)
            self._staged = None
        self.graph.replay()
        self.step.count += 1
        return self.loss


def shard_batch(x, rank, world):
    """Rank r gets samples [r*B/n, (r+1)*B/n); B must divide evenly so that the mean of
    the local gradients equals the full-batch gradient."""
    B = x.shape[0]
    if B % world != 0:
        raise ValueError(f"global batch {B} is not divisible by the number of ranks {world}")
    per = B // world
    return x[rank * per:(rank + 1) * per]
