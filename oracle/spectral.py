"""Oracle: FNO spectral convolution (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/pdequinox/conv/_spectral_conv.py:
  * generate_modes_slices  <- :73-104
  * spectral_conv_nd       <- :107-136 (rfftn -> per-corner einsum -> irfftn)
numpy.fft is pocketfft, the same c2r convention the JAX CPU backend uses.

Also holds the dense truncated-DFT formulation that the
CUDA kernels implement, cross-checked against the FFT formulation in
tests/test_oracle_*.py, and the analytic VJP (the reference relies on JAX
autodiff; checked against finite differences in float64).

Every function takes arrays with arbitrary leading batch axes: x is
(..., C_i, *S) with D = len(modes) trailing spatial axes.
"""
from __future__ import annotations

from itertools import product

import numpy as np


def generate_modes_slices(modes):
    """Corner-block slices, in the reference's order (_spectral_conv.py:73-104).

    Returns a list of 2**(D-1) lists, each [slice(None) (channels), *D slices].
    Last axis is always [:m_D]; every other axis is [:m] or [-m:]; bit a of the
    block index g selects the high slice on axis a.
    """
    *ms, ml = modes
    slices_ = [[slice(None, ml)]]
    slices_ += [[slice(None, mode), slice(-mode, None)] for mode in reversed(ms)]
    return [[slice(None)] + list(reversed(i)) for i in product(*slices_)]


def spectral_conv_nd(x, weight_r, weight_i, modes):
    """Reference formulation (_spectral_conv.py:107-136), batched over leading axes.

    x: (..., C_i, *S); weight_*: (G, C_o, C_i, *M) -- the CALL-TIME interpretation
    (`_, o, *_ = weight.shape`, einsum "i...,oi...->o...").
    """
    modes = tuple(int(m) for m in modes)
    D = len(modes)
    *lead, ci = x.shape[: x.ndim - D]
    spatial = x.shape[x.ndim - D:]
    *si, sl = spatial
    weight = weight_r + 1j * weight_i
    o = weight.shape[1]
    axes = tuple(range(x.ndim - D, x.ndim))
    x_fft = np.fft.rfftn(x, s=spatial, axes=axes)
    cdtype = np.complex64 if x.dtype == np.float32 else np.complex128
    out = np.zeros((*lead, o, *si, sl // 2 + 1), dtype=cdtype)
    for g, slice_g in enumerate(generate_modes_slices(modes)):
        idx = (Ellipsis, *slice_g)
        blk = x_fft[idx]  # (..., i, *m)
        sub = "xyz"[:D]
        mixed = np.einsum(f"...i{sub},oi{sub}->...o{sub}", blk, weight[g])
        out[idx] = mixed.astype(cdtype)
    y = np.fft.irfftn(out, s=spatial, axes=axes)
    return y.astype(x.dtype)


# ---------------------------------------------------------------------------
# Dense truncated-DFT formulation (what the CUDA kernels compute)
# ---------------------------------------------------------------------------

def retained_wavenumbers(s, m, last):
    """Signed->unsigned bin index kappa(j) of retained column j on one axis.

    Last axis: j in [0,m) -> j. Other axes: j in [0,2m): j<m -> j (low slice),
    else s-2m+j (high slice == negative frequencies -m..-1).
    """
    if last:
        return np.arange(m)
    j = np.arange(2 * m)
    return np.where(j < m, j, s - 2 * m + j)


def overwrite_mask(s, m):
    """1 for retained columns that survive `.at[].set` when slices overlap.

    If 2m > s the low and high slices of an axis overlap and the LATER block
    (high slice) overwrites (_spectral_conv.py:135); per axis this means a
    low-slice column whose bin is also covered by the high slice contributes
    nothing to the inverse transform.
    """
    kap = retained_wavenumbers(s, m, last=False)
    keep = np.ones(2 * m)
    hi = set(kap[m:].tolist())
    for j in range(m):
        if int(kap[j]) in hi:
            keep[j] = 0.0
    return keep


def dft_tables(spatial, modes, dtype=np.float64):
    """Twiddle tables of the dense formulation, all generated in float64.

    Returns dict with
      fwd_last  (m_D, s_D) complex : exp(-2 pi i k n / s_D)
      fwd[a]    (2m_a, s_a) complex: exp(-2 pi i kappa_a(j) n / s_a), a < D-1
      inv[a]    (s_a, 2m_a) complex: keep_a(j) * exp(+2 pi i kappa_a(j) n / s_a) / s_a
      inv_last  (m_D, s_D) complex : c(k)/s_D * exp(+2 pi i k n / s_D)
    with c(k)=1 for k=0 and the Nyquist bin (s_D even), else 2
    (pocketfft c2r: imaginary parts of those two bins are dropped).
    """
    spatial = tuple(int(s) for s in spatial)
    modes = tuple(int(m) for m in modes)
    D = len(spatial)
    sD, mD = spatial[-1], modes[-1]
    if mD > sD // 2 + 1:
        raise ValueError("num_modes on the last axis exceeds s//2+1 (reference einsum shape-errors)")
    k = np.arange(mD)[:, None]
    n = np.arange(sD)[None, :]
    ang = 2.0 * np.pi * ((k * n) % sD) / sD
    c = np.full(mD, 2.0)
    c[0] = 1.0
    if sD % 2 == 0 and mD - 1 >= sD // 2:
        c[sD // 2] = 1.0
    cdt = np.complex128 if dtype == np.float64 else np.complex64
    t = {
        "fwd_last": np.exp(-1j * ang).astype(cdt),
        "inv_last": ((c[:, None] / sD) * np.exp(1j * ang)).astype(cdt),
        "fwd": [],
        "inv": [],
    }
    for a in range(D - 1):
        s, m = spatial[a], modes[a]
        if m > s:
            raise ValueError("num_modes exceeds the axis length")
        kap = retained_wavenumbers(s, m, last=False)[:, None]
        nn = np.arange(s)[None, :]
        ang = 2.0 * np.pi * ((kap * nn) % s) / s
        keep = overwrite_mask(s, m)
        t["fwd"].append(np.exp(-1j * ang).astype(cdt))  # (2m, s)
        t["inv"].append(((keep[:, None] * np.exp(1j * ang)) / s).T.copy().astype(cdt))  # (s, 2m)
    return t


def _corner_weights_to_modes(weight, modes):
    """(G, o, i, *M) complex -> (o, i, 2m_1, .., 2m_{D-1}, m_D): corner g placed at its j-range."""
    D = len(modes)
    G, o, i = weight.shape[:3]
    full = np.zeros((o, i, *[2 * m for m in modes[:-1]], modes[-1]), dtype=weight.dtype)
    for g in range(G):
        idx = [slice(None), slice(None)]
        for a in range(D - 1):
            m = modes[a]
            hi = (g >> a) & 1
            idx.append(slice(m, 2 * m) if hi else slice(0, m))
        idx.append(slice(None))
        full[tuple(idx)] = weight[g]
    return full


def _modes_to_corner_weights(full, modes):
    D = len(modes)
    G = 2 ** (D - 1)
    out = []
    for g in range(G):
        idx = [slice(None), slice(None)]
        for a in range(D - 1):
            m = modes[a]
            hi = (g >> a) & 1
            idx.append(slice(m, 2 * m) if hi else slice(0, m))
        idx.append(slice(None))
        out.append(full[tuple(idx)])
    return np.stack(out)


def forward_modes(x, modes, tables=None):
    """x (..., C, *S) real -> x_hat (..., C, 2m_1, .., 2m_{D-1}, m_D) complex."""
    D = len(modes)
    spatial = x.shape[x.ndim - D:]
    t = tables or dft_tables(spatial, modes, x.dtype)
    z = np.tensordot(x, t["fwd_last"], axes=([x.ndim - 1], [1]))  # contracted axis -> last
    for a in reversed(range(D - 1)):
        ax = z.ndim - D + a
        z = np.moveaxis(np.tensordot(z, t["fwd"][a], axes=([ax], [1])), -1, ax)
    return z


def inverse_modes(y_hat, spatial, modes, tables=None, real_dtype=np.float64):
    """y_hat (..., C, 2m_1.., m_D) complex -> y (..., C, *S) real (irfftn semantics)."""
    D = len(modes)
    t = tables or dft_tables(spatial, modes, real_dtype)
    z = y_hat
    for a in range(D - 1):
        ax = z.ndim - D + a
        z = np.moveaxis(np.tensordot(z, t["inv"][a], axes=([ax], [1])), -1, ax)
    y = np.tensordot(z, t["inv_last"], axes=([z.ndim - 1], [0]))
    return np.real(y)


def mix_modes(x_hat, weight_full):
    """(..., i, *J) x (o, i, *J) -> (..., o, *J)."""
    D = weight_full.ndim - 2
    sub = "xyz"[:D]
    return np.einsum(f"...i{sub},oi{sub}->...o{sub}", x_hat, weight_full)


def spectral_conv_dense(x, weight_r, weight_i, modes):
    """Dense truncated-DFT formulation of spectral_conv_nd."""
    modes = tuple(int(m) for m in modes)
    D = len(modes)
    spatial = x.shape[x.ndim - D:]
    t = dft_tables(spatial, modes, x.dtype)
    w = _corner_weights_to_modes(weight_r + 1j * weight_i, modes)
    x_hat = forward_modes(x, modes, t)
    y_hat = mix_modes(x_hat, w)
    y = inverse_modes(y_hat, spatial, modes, t, x.dtype)
    return y.astype(x.dtype)


def spectral_conv_vjp(x, weight_r, weight_i, modes, gy):
    """Analytic VJP of spectral_conv_nd.

    Returns (gx, gw_r, gw_i, x_hat); weight grads are summed over all leading
    (batch) axes, which is what grad(mean-loss(vmap(model))) produces.
    Complex gradient convention: g = dL/dRe + i dL/dIm, so a complex-linear map
    z = A u back-propagates as g_u = A^H g_z.
    """
    modes = tuple(int(m) for m in modes)
    D = len(modes)
    spatial = x.shape[x.ndim - D:]
    t = dft_tables(spatial, modes, x.dtype)
    w = _corner_weights_to_modes(weight_r + 1j * weight_i, modes)
    x_hat = forward_modes(x, modes, t)

    # adjoint of the c2r last-axis stage: y = Re(z . inv_last)  =>  g_z = gy . conj(inv_last)^T
    gz = np.tensordot(gy, np.conj(t["inv_last"]), axes=([gy.ndim - 1], [1]))
    # adjoint of the leading-axis inverse stages
    for a in reversed(range(D - 1)):
        ax = gz.ndim - D + a
        gz = np.moveaxis(np.tensordot(gz, np.conj(t["inv"][a]), axes=([ax], [0])), -1, ax)
    g_yhat = gz  # (..., o, *J)

    sub = "xyz"[:D]
    g_xhat = np.einsum(f"...o{sub},oi{sub}->...i{sub}", g_yhat, np.conj(w))
    lead = x_hat.ndim - (D + 1)
    xh = x_hat.reshape((-1,) + x_hat.shape[lead:])
    gh = g_yhat.reshape((-1,) + g_yhat.shape[lead:])
    gw_full = np.einsum(f"bo{sub},bi{sub}->oi{sub}", gh, np.conj(xh))
    gw = _modes_to_corner_weights(gw_full, modes)

    # adjoint of the forward stages
    g = g_xhat
    for a in range(D - 1):
        ax = g.ndim - D + a
        g = np.moveaxis(np.tensordot(g, np.conj(t["fwd"][a]), axes=([ax], [0])), -1, ax)
    gx = np.real(np.tensordot(g, np.conj(t["fwd_last"]), axes=([g.ndim - 1], [0])))
    rd = x.dtype
    return gx.astype(rd), np.real(gw).astype(rd), np.imag(gw).astype(rd), x_hat


def init_spectral_weights(rng, num_spatial_dims, in_channels, out_channels, num_modes, dtype=np.float32):
    """Same distribution and ALLOCATION shape as SpectralConv.__init__ (_spectral_conv.py:45-63):
    scale * N(0,1), scale = 1/(in*out), shape (2**(D-1), in, out, *modes). (jax.random bit
    streams are not reproducible without jax; only the distribution is.)"""
    if isinstance(num_modes, int):
        num_modes = (num_modes,) * num_spatial_dims
    if len(num_modes) != num_spatial_dims:
        raise ValueError("num_modes must have the same length as num_spatial_dims")
    shape = (2 ** (num_spatial_dims - 1), in_channels, out_channels) + tuple(num_modes)
    scale = 1.0 / (in_channels * out_channels)
    wr = (scale * rng.standard_normal(shape)).astype(dtype)
    wi = (scale * rng.standard_normal(shape)).astype(dtype)
    return wr, wi
