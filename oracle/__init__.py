"""CPU oracle for the pdequinox hot path (TEST INFRASTRUCTURE, not product code).

This package is a NumPy restatement of the reference's algorithms for
`SpectralConv`, `PhysicsConv` / `PointwiseLinearConv`, the FNO / ResNet blocks
built from them, the MSE + Adam training step, and their analytic VJPs.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
    `--impl reference` legs may import anything from here;
  * `pdequinox_b200/` never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * pinned against the reference's own known-answer tests
    (tests/test_spectral_conv.py, tests/test_boundary_handling.py,
    tests/test_unet_translation.py, tests/test_count_parameters.py of the
    reference), re-stated in tests/test_oracle_reference_known_answers.py;
  * pinned against golden vectors produced by executing the reference's OWN
    source files (pdequinox/conv/*.py, blocks, arch) in this container on top
    of a NumPy stand-in for the absent jax/equinox packages
    (tests/golden/make_golden.py + tests/golden/_jaxshim);
  * third-party numerics that live in jax/equinox/optax themselves
    (gelu variant, GroupNorm, Adam, lax.conv_general_dilated) are restated from
    their documented definitions: "parity unpinned" for those pieces, because
    neither package can be installed here.
"""

from . import conv, nn, spectral  # noqa: F401
