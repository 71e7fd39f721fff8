"""The reference's README quickstart (README.md:55-100 of Ceyron/pdequinox: train a ClassicUNet as a 1-D Poisson solver)
written against pdequinox_b200: same data, architecture call, loss, Adam(3e-4), batch size and dataloader; the
`jax.vmap(model)` / `eqx.filter_value_and_grad` / `optax.adam` triple is one `TrainStep` here.

    python examples/quickstart_poisson.py [--arch unet|fno] [--epochs 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import random as jr
from pdequinox_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="unet", choices=["unet", "fno"])
ap.add_argument("--epochs", type=int, default=20)
args = ap.parse_args()

force_fields, displacement_fields = pdeqx.sample_data.poisson_1d_dirichlet(key=jr.PRNGKey(0))
force_train, force_test = force_fields[:800], force_fields[800:]
disp_train, disp_test = displacement_fields[:800], displacement_fields[800:]

if args.arch == "unet":
    model = pdeqx.arch.ClassicUNet(1, 1, 1, boundary_mode="dirichlet", key=jr.PRNGKey(1))
else:
    model = pdeqx.arch.ClassicFNO(1, 1, 1, key=jr.PRNGKey(1))
print(f"{type(model).__name__}: {pdeqx.count_parameters(model)} parameters")

step = TrainStep(model, lr=3e-4)  # loss = mean((vmap(model)(x) - y)^2), reverse pass, Adam: one call


def test_error():
    pred = pdeqx.vmap(model)(torch.from_numpy(force_test).cuda()).cpu().numpy()
    return float(np.mean((pred - disp_test) ** 2))


print(f"test MSE before training: {test_error():.3e}")
shuffle_key = jr.PRNGKey(151)
history = []
for epoch in range(args.epochs):
    shuffle_key, subkey = jr.split(shuffle_key)
    for x, y in pdeqx.dataloader((force_train, disp_train), batch_size=32, key=subkey):
        history.append(step(x, y).item())
    if epoch % 5 == 4 or epoch == args.epochs - 1:
        print(f"epoch {epoch + 1:3d}: train loss {np.mean(history[-25:]):.3e}")
final = test_error()
print(f"test MSE after {args.epochs} epochs: {final:.3e}")
assert np.isfinite(final) and np.mean(history[-25:]) < 0.2 * np.mean(history[:25]), "the loss did not go down"
