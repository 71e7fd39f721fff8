"""Host-side minibatch generators: the contract of pdequinox/_utils.py:66-187 (dataloader, cycling_dataloader) --
ceil(n / batch_size) shuffled batches per epoch, equal leading axes required, a batch never mixes two epochs."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import random as jr


def test_one_epoch_covers_every_sample_once():
    x = np.arange(103 * 3, dtype=np.float32).reshape(103, 3)
    y = np.arange(103, dtype=np.float32)
    batches = list(pdeqx.dataloader((x, y), batch_size=32, key=jr.PRNGKey(0)))
    assert [len(b[1]) for b in batches] == [32, 32, 32, 7]          # ceil(103 / 32) batches, the last one ragged
    seen = np.concatenate([b[1] for b in batches])
    assert sorted(seen.tolist()) == list(range(103))                 # a permutation of the samples
    assert not np.array_equal(seen, y)                               # ... that is shuffled
    for bx, by in batches:                                           # the leaves stay aligned
        assert np.array_equal(bx[:, 0], by * 3)


def test_same_key_same_order_and_other_key_other_order():
    x = np.arange(50)
    a = np.concatenate(list(pdeqx.dataloader(x, batch_size=8, key=jr.PRNGKey(3))))
    b = np.concatenate(list(pdeqx.dataloader(x, batch_size=8, key=jr.PRNGKey(3))))
    c = np.concatenate(list(pdeqx.dataloader(x, batch_size=8, key=jr.PRNGKey(4))))
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_batch_larger_than_data_and_exact_division():
    x = np.arange(10)
    assert [len(b) for b in pdeqx.dataloader(x, batch_size=64, key=0)] == [10]
    assert [len(b) for b in pdeqx.dataloader(x, batch_size=5, key=0)] == [5, 5]


def test_leaves_must_agree_on_the_sample_axis():
    with pytest.raises(ValueError, match="same number of samples"):
        next(pdeqx.dataloader((np.zeros((4, 2)), np.zeros((5, 2))), batch_size=2, key=0))


def test_dict_and_torch_leaves():
    data = {"u": torch.arange(12).reshape(12, 1), "f": np.arange(12) * 2}
    got = list(pdeqx.dataloader(data, batch_size=5, key=1))
    assert [b["u"].shape[0] for b in got] == [5, 5, 2]
    for b in got:
        assert isinstance(b["u"], torch.Tensor)
        assert np.array_equal(b["u"][:, 0].numpy() * 2, b["f"])


def test_cycling_dataloader_steps_epochs_and_info():
    x = np.arange(10)
    items = list(pdeqx.cycling_dataloader(x, batch_size=4, num_steps=8, key=jr.PRNGKey(7), return_info=True))
    assert len(items) == 8
    assert [(e, b) for _, e, b in items] == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (2, 0), (2, 1)]
    assert [len(d) for d, _, _ in items] == [4, 4, 2, 4, 4, 2, 4, 4]   # a batch never mixes two epochs
    for epoch in (0, 1):
        seen = np.concatenate([d for d, e, _ in items if e == epoch])
        assert sorted(seen.tolist()) == list(range(10))
    first = np.concatenate([d for d, e, _ in items if e == 0])
    second = np.concatenate([d for d, e, _ in items if e == 1])
    assert not np.array_equal(first, second)                           # every epoch is shuffled afresh
    plain = list(pdeqx.cycling_dataloader(x, batch_size=4, num_steps=3, key=jr.PRNGKey(7)))
    assert all(np.array_equal(p, d) for p, (d, _, _) in zip(plain, items))


def test_cycling_dataloader_zero_steps():
    assert list(pdeqx.cycling_dataloader(np.arange(4), batch_size=2, num_steps=0, key=0)) == []
