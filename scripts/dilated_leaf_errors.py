"""Per-leaf gradient error of DilatedResNet (configs[2] geometry, one sample, TF32) against the float64 oracle, folded
(default) or with PDEQX_GN_FOLD=0: the diagnostic behind the tolerance of tests/test_gpu_full_shapes.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep
from oracle import nn as onn
from test_gpu_full_shapes import relu_masks_of_forward, oracle_params, dev, host
from conftest import rel_l2

pdeqx.set_precision("tf32")
for seed in (43, 44, 45):
    rng = np.random.default_rng(seed)
    model = pdeqx.DilatedResNet(2, 1, 1, hidden_channels=64, boundary_mode="periodic", key=2)
    step = TrainStep(model)
    with torch.no_grad():
        for k, v in step.arena.named_params().items():
            if "norm_layers" in k:
                v.add_(dev(0.2 * rng.standard_normal(tuple(v.shape))))
    p = oracle_params(step)
    x, y = rng.standard_normal((1, 1, 128, 128)), rng.standard_normal((1, 1, 128, 128))
    pred = host(pdeqx.vmap(model)(dev(x)))
    e_pred = rel_l2(pred, onn.dilated_resnet_forward(p, x, 2, "periodic"))
    loss = step.loss_and_grad(x, y).item()
    masks = relu_masks_of_forward(model, x)
    loss_ref, g_ref = onn.dilated_resnet_loss_and_grad(p, x, y, 2, "periodic", relu_masks=masks)
    g = {k: host(v) for k, v in step.arena.named_grads().items()}
    errs = sorted(((rel_l2(g[k], g_ref[k]), k) for k in g_ref), reverse=True)
    print(f"seed {seed} fold={os.environ.get('PDEQX_GN_FOLD', '1')} pred {e_pred:.2e} loss {abs(loss - loss_ref) / abs(loss_ref):.2e} rms-leaf {np.sqrt(np.mean([e * e for e, _ in errs])):.2e}")
    for e, k in errs[:6]:
        print(f"   {e:.2e} {k}")
