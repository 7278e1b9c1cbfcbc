"""Synthetic workloads of SURVEY.md §8(d): counter-based, so the host (numpy) and the device (torch)
generate bit-identical inputs without a copy.

    x = splitmix64_mix(seed * 0x9E3779B97F4A7C15 + 4 * i + k)      (mod 2^64)
    u(seed, i, k) = float32(x >> 40) * 2^-24   in [0, 1)

i = item id (0-based), k = component 0..3. Pure input generation — no BVH arithmetic here.
"""
from __future__ import annotations

import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB


def _mix_np(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform_np(seed: int, n: int, k: int, start: int = 0) -> np.ndarray:
    """u(seed, i, k) for i in [start, start+n) as float32."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed * _GOLD) & 0xFFFFFFFFFFFFFFFF)
        i = np.arange(start, start + n, dtype=np.uint64)
        z = base + np.uint64(4) * i + np.uint64(k)
    x = _mix_np(z)
    return (x >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def _mix_torch(z):
    import torch
    # int64 arithmetic wraps mod 2^64; logical right shift emulated by masking the sign extension
    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)
    m1 = torch.tensor(_M1 - (1 << 64), dtype=torch.int64, device=z.device)
    m2 = torch.tensor(_M2 - (1 << 64), dtype=torch.int64, device=z.device)
    z = (z ^ lsr(z, 30)) * m1
    z = (z ^ lsr(z, 27)) * m2
    z = z ^ lsr(z, 31)
    return z


def uniform_torch(seed: int, n: int, k: int, device, start: int = 0):
    import torch
    base = (seed * _GOLD) & 0xFFFFFFFFFFFFFFFF
    if base >= 1 << 63:
        base -= 1 << 64
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    z = i * 4 + (base + k)          # wraps mod 2^64 in int64
    x = _mix_torch(z)
    hi = (x >> 40) & ((1 << 24) - 1)
    return hi.to(torch.float32) * (2.0 ** -24)


def sphere_radius_scale(n: int) -> float:
    """s = 0.8124 * N^(-1/3): about 8 neighbours per leaf, C ~ 4 N contacts."""
    return 0.8124 * float(n) ** (-1.0 / 3.0)


def random_spheres_np(n: int, seed: int = 42, scale: float | None = None) -> np.ndarray:
    """Config 1/2/3/5 leaves: centres U[0,1)^3, radius s * (0.5 + 0.5 u). Returns BSphere{Float32}[n]."""
    s = np.float32(sphere_radius_scale(n) if scale is None else scale)
    out = np.zeros(n, np.dtype([("x", np.float32, 3), ("r", np.float32)]))
    for k in range(3):
        out["x"][:, k] = uniform_np(seed, n, k)
    out["r"] = s * (np.float32(0.5) + np.float32(0.5) * uniform_np(seed, n, 3))
    return out


def random_spheres_torch(n: int, device, seed: int = 42, scale: float | None = None):
    """Same bits as random_spheres_np, generated on the device: float32 tensor (n, 4) = BSphere{Float32}[n]."""
    import torch
    s = np.float32(sphere_radius_scale(n) if scale is None else scale)
    out = torch.empty((n, 4), dtype=torch.float32, device=device)
    for k in range(3):
        out[:, k] = uniform_torch(seed, n, k, device)
    out[:, 3] = float(s) * (0.5 + 0.5 * uniform_torch(seed, n, 3, device))
    return out


def shell_spheres_np(n_theta: int = 1000, n_phi: int = 1000) -> np.ndarray:
    """Config 4 leaves: spheres on an n_theta x n_phi (theta, phi) grid over the unit sphere surface,
    radius = 0.75 x local grid spacing ("mesh-like": every leaf touches its grid neighbours)."""
    th = (np.arange(n_theta, dtype=np.float64) + 0.5) * (np.pi / n_theta)
    ph = np.arange(n_phi, dtype=np.float64) * (2.0 * np.pi / n_phi)
    T, P = np.meshgrid(th, ph, indexing="ij")
    out = np.zeros(n_theta * n_phi, np.dtype([("x", np.float32, 3), ("r", np.float32)]))
    out["x"][:, 0] = (np.sin(T) * np.cos(P)).ravel()
    out["x"][:, 1] = (np.sin(T) * np.sin(P)).ravel()
    out["x"][:, 2] = np.cos(T).ravel()
    spacing = np.maximum(np.pi / n_theta, np.sin(T) * (2.0 * np.pi / n_phi))
    out["r"] = (0.75 * spacing).ravel()
    return out


def random_rays_np(n: int, seed: int = 7, start: int = 0):
    """Config 4 rays: origins U[-1.5, 1.5)^3, directions uniform on S^2. Returns (points, directions) as
    float32 arrays of shape (n, 3) (== the reference's column-major 3 x n)."""
    p = np.empty((n, 3), np.float32)
    for k in range(3):
        p[:, k] = np.float32(3.0) * uniform_np(seed, n, k, start) - np.float32(1.5)
    z = np.float32(2.0) * uniform_np(seed + 1000, n, 0, start) - np.float32(1.0)
    phi = np.float32(2.0 * np.pi) * uniform_np(seed + 1000, n, 1, start)
    r = np.sqrt(np.maximum(np.float32(0.0), np.float32(1.0) - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    return p, d


def random_rays_torch(n: int, device, seed: int = 7, start: int = 0):
    """Device twin of random_rays_np (origins bit-identical; directions agree to float32 rounding of
    cos/sin and are only used for throughput runs, parity tests use the numpy version on both sides)."""
    import torch
    p = torch.empty((n, 3), dtype=torch.float32, device=device)
    for k in range(3):
        p[:, k] = 3.0 * uniform_torch(seed, n, k, device, start) - 1.5
    z = 2.0 * uniform_torch(seed + 1000, n, 0, device, start) - 1.0
    phi = (2.0 * np.pi) * uniform_torch(seed + 1000, n, 1, device, start)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    d = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()
    return p, d
