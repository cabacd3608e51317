"""CPU oracle for the spectral encoding front end -- TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement of the reference's algorithm for the hot path
(points -> 16x360 min-range image -> hole interpolation -> row-wise 360-pt rFFT
magnitude -> 50-bin exponential histogram per row -> L1-normalised 800-D
descriptor). It exists to CHECK the CUDA path. Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; nothing under ``neural_spectral_codec_b200/``
does, and the product path raises if the CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against the reference itself, imported
read-only in the build container by ``tests/golden/make_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` requires bit-identical intermediates
(range image, interpolated image, freq->bin LUT) and a bit-identical descriptor
on the generating platform (allclose elsewhere: FFT/atan2 backends differ).

Every function cites the reference lines (relative to /root/reference) it
follows. The arithmetic types are the reference's under NumPy >= 2:
float32 for range / azimuth / elevation / column, float64 for the row index
(``f32_array - np.float64`` promotes), float64 ``np.interp`` for hole filling,
complex64 FFT, float32 histogram.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch


@dataclass
class OracleConfig:
    """Constructor surface of the reference encoder (src/encoding/spectral_encoder.py:35-47)
    plus the projector defaults it never forwards (src/encoding/range_image.py:102-109)."""
    n_elevation: int = 16
    n_azimuth: int = 360
    n_bins: int = 50
    alpha: float = 2.0
    epsilon: float = 1e-8
    target_elevation_bins: int = 16
    interpolate_empty: bool = True
    elevation_range: Tuple[float, float] = (-24.8, 2.0)
    min_range: float = 1.0
    max_range: float = 80.0
    # np.deg2rad returns np.float64 scalars (range_image.py:126-127); keeping the
    # numpy scalar type matters because it drives the float64 promotion of the row index.
    el_min: np.float64 = field(init=False)
    el_max: np.float64 = field(init=False)

    def __post_init__(self):
        self.el_min = np.deg2rad(self.elevation_range[0])
        self.el_max = np.deg2rad(self.elevation_range[1])

    @property
    def n_freqs(self) -> int:  # spectral_encoder.py:88
        return self.n_azimuth // 2 + 1

    @property
    def output_dim(self) -> int:  # spectral_encoder.py:91
        return self.target_elevation_bins * self.n_bins


# --------------------------------------------------------------------------- projection
def spherical(points: np.ndarray, cfg: OracleConfig) -> Dict[str, np.ndarray]:
    """Steps 1-8 of ``RangeImageProjector.project`` (range_image.py:146-198) with every
    intermediate kept. ``kept`` indexes the ORIGINAL rows of ``points`` that survive both
    the finite filter (:151-155) and the range filter (:174-177)."""
    p = np.asarray(points)
    cx, cy, cz = p[:, 0], p[:, 1], p[:, 2]
    finite = np.isfinite(cx) & np.isfinite(cy) & np.isfinite(cz)          # :151
    cx, cy, cz = cx[finite], cy[finite], cz[finite]                        # :152-154
    sq = [np.clip(c ** 2, 0, 1e10) for c in (cx, cy, cz)]                  # :159-161
    rng = np.sqrt(sq[0] + sq[1] + sq[2])                                   # :162  (x2+y2)+z2
    az = np.arctan2(cy, cx)                                                # :166
    az = (az + np.pi) % (2 * np.pi)                                        # :167
    el = np.arctan2(cz, np.sqrt(sq[0] + sq[1]))                            # :170-171
    ok = (rng >= cfg.min_range) & (rng <= cfg.max_range) & np.isfinite(rng)  # :174
    rng, az, el = rng[ok], az[ok], el[ok]                                  # :175-177
    frac = (el - cfg.el_min) / (cfg.el_max - cfg.el_min)                   # :186 (float64)
    row = np.clip(np.floor(frac * cfg.n_elevation).astype(int), 0, cfg.n_elevation - 1)  # :187-191
    col = np.clip(np.floor(az / (2 * np.pi) * cfg.n_azimuth).astype(int), 0, cfg.n_azimuth - 1)  # :194-198
    kept = np.flatnonzero(finite)[ok]
    return {"range": rng, "azimuth": az, "elevation": el, "row": row, "col": col, "kept": kept}


def project(points: np.ndarray, cfg: OracleConfig) -> np.ndarray:
    """``RangeImageProjector.project(points, keep_intensity=False)[0]`` (range_image.py:129-214):
    per-pixel MIN range (``np.minimum.at``, :208), empty pixels -> 0 (:214)."""
    s = spherical(points, cfg)
    flat = np.full(cfg.n_elevation * cfg.n_azimuth, np.inf, dtype=np.float32)   # :205
    np.minimum.at(flat, s["row"] * cfg.n_azimuth + s["col"], s["range"])       # :202,208
    img = flat.reshape(cfg.n_elevation, cfg.n_azimuth)                           # :211
    img[img == np.inf] = 0.0                                                     # :214
    return img


def project_with_intensity(points: np.ndarray, cfg: OracleConfig):
    """``RangeImageProjector.project(points, keep_intensity=True)`` (range_image.py:129-232): the
    range image and, for 4-column input, the intensity image -- per pixel the largest intensity
    among the points whose range equals the pixel's minimum, floored at the initial 0
    (``np.maximum.at`` on zeros, :220-226). 3-column input -> ``(image, None)``."""
    img = project(points, cfg)
    p = np.asarray(points)
    if p.shape[1] != 4:                                                          # :179-182
        return img, None
    s = spherical(points, cfg)
    lin = s["row"] * cfg.n_azimuth + s["col"]
    inten = p[s["kept"], 3]
    flat = np.zeros(cfg.n_elevation * cfg.n_azimuth, dtype=np.float32)         # :220
    closest = s["range"] == img.reshape(-1)[lin]                                 # :223
    with np.errstate(invalid="ignore"):                                          # NaN intensities propagate
        np.maximum.at(flat, lin[closest], inten[closest])                        # :226
    return img, flat.reshape(cfg.n_elevation, cfg.n_azimuth)                     # :228


# ------------------------------------------------------------------------ interpolation
def interpolate_range_image(img: np.ndarray, method: str = "linear") -> np.ndarray:
    """``interpolate_range_image(img, method)`` (range_image.py:15-89); the encoder uses 'linear'.
    'nearest' (:66-75): a hole takes the valid pixel at the smallest circular distance, the first
    one in ascending column order on a tie (``np.argmin``).

    Pass 1 (:33-64): in each row that has some but not all pixels > 0, every empty pixel
    is linearly interpolated along azimuth between its nearest valid neighbours, with the
    row treated as circular by tiling the valid samples at -W / 0 / +W; ``np.interp``
    evaluates in float64 and the store back into the float32 row rounds to nearest.
    Pass 2 (:77-87): rows still empty copy the nearest non-empty row, scanning rows in
    increasing order and mutating in place, testing ``row - k`` before ``row + k``.
    """
    out = img.copy()                                                             # :30
    H, W = out.shape
    for r in range(H):
        valid = out[r] > 0                                                       # :35
        n_valid = int(valid.sum())
        if n_valid == 0 or n_valid == W:                                         # :37-43
            continue
        vi = np.flatnonzero(valid)                                               # :46
        if method == "nearest":
            for x in np.flatnonzero(~valid):                                     # :68-75
                d = np.minimum(np.abs(vi - x), W - np.abs(vi - x))
                out[r, x] = out[r, vi[np.argmin(d)]]
            continue
        xp = np.concatenate([vi - W, vi, vi + W])                                # :55-59
        fp = np.tile(out[r][valid], 3)                                           # :47,60
        holes = np.flatnonzero(~valid)                                           # :50
        out[r, holes] = np.interp(holes, xp, fp)                                 # :63-64
    for r in range(H):                                                           # :78
        if np.any(out[r] > 0):
            continue
        for k in range(1, H):                                                    # :81
            if r - k >= 0 and np.any(out[r - k] > 0):                            # :82-84
                out[r] = out[r - k]
                break
            if r + k < H and np.any(out[r + k] > 0):                             # :85-87
                out[r] = out[r + k]
                break
    return out


# ----------------------------------------------------------------------------- spectrum
def bin_edges(cfg: OracleConfig, alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``SpectralEncoder._compute_bin_edges`` (spectral_encoder.py:93-116), float32 torch ops."""
    a = torch.tensor(cfg.alpha, dtype=torch.float32) if alpha is None else alpha
    t = torch.linspace(0, 1, cfg.n_bins + 1, device=a.device)                    # :107
    e = (torch.exp(a * t) - 1) / (torch.exp(a) - 1 + cfg.epsilon)                # :111
    return e * cfg.n_freqs                                                       # :114


def freq_to_bin(cfg: OracleConfig, alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Frequency index -> histogram bin (spectral_encoder.py:136-145): int64 ``(n_freqs,)``."""
    k = torch.arange(cfg.n_freqs, dtype=torch.float32)                           # :136-140
    b = torch.searchsorted(bin_edges(cfg, alpha), k, right=True) - 1             # :144
    return torch.clamp(b, 0, cfg.n_bins - 1)                                     # :145


def pool_rows(img: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """Row pooling when the image height differs from ``target_elevation_bins``
    (spectral_encoder.py:171-176)."""
    if img.shape[0] == cfg.target_elevation_bins:
        return img
    return torch.nn.functional.adaptive_avg_pool2d(
        img.unsqueeze(0).unsqueeze(0), (cfg.target_elevation_bins, img.shape[1])).squeeze()


def fft_magnitudes(img: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """|rfft(norm='ortho')| * sqrt(W) per row (spectral_encoder.py:180-186) -> ``(rows, n_freqs)`` f32."""
    spec = torch.fft.rfft(img, dim=1, norm="ortho")                              # :180
    return torch.abs(spec) * np.sqrt(cfg.n_azimuth)                              # :183,186


def histogram_rows(mag: torch.Tensor, cfg: OracleConfig,
                   alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``_bin_fft_magnitudes`` (spectral_encoder.py:118-158): per-row scatter-add of the
    magnitudes into their bins, in ascending-frequency order; un-normalised ``(rows*n_bins,)``."""
    lut = freq_to_bin(cfg, alpha).long()
    hist = torch.zeros(mag.shape[0], cfg.n_bins)                                 # :149
    for r in range(mag.shape[0]):                                                # :152
        hist[r].scatter_add_(0, lut, mag[r])                                     # :153-155
    return hist.flatten()                                                        # :158


def encode_range_image(img: torch.Tensor, cfg: OracleConfig,
                       alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``SpectralEncoder.encode_range_image`` (spectral_encoder.py:160-204)."""
    h = histogram_rows(fft_magnitudes(pool_rows(img, cfg), cfg), cfg, alpha)
    total = h.sum()                                                              # :197
    if total > cfg.epsilon:                                                      # :198
        return h / (total + cfg.epsilon)                                         # :199
    return torch.ones_like(h) / h.numel()                                        # :202


def encode_points(points: np.ndarray, cfg: OracleConfig) -> torch.Tensor:
    """``SpectralEncoder.encode_points`` (spectral_encoder.py:206-229): project ->
    (optional) interpolate -> encode_range_image. float32 ``(target_rows*n_bins,)``."""
    img = project(points, cfg)                                                   # :217
    if cfg.interpolate_empty:                                                    # :220
        img = interpolate_range_image(img)                                       # :221
    return encode_range_image(torch.from_numpy(img).float(), cfg)                # :224,227


def encode_batch(images: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """``SpectralEncoder.forward`` / ``encode_batch`` (spectral_encoder.py:231-261): range
    images in, no projection and NO interpolation."""
    return torch.stack([encode_range_image(images[i], cfg) for i in range(images.shape[0])], 0)


def stages(points: np.ndarray, cfg: OracleConfig) -> Dict[str, np.ndarray]:
    """Every intermediate of ``encode_points`` for per-stage parity tests."""
    img = project(points, cfg)
    filled = interpolate_range_image(img) if cfg.interpolate_empty else img
    t = pool_rows(torch.from_numpy(filled).float(), cfg)
    mag = fft_magnitudes(t, cfg)
    return {
        "range_image": img,
        "interpolated": filled,
        "magnitudes": mag.numpy(),
        "freq_to_bin": freq_to_bin(cfg).numpy(),
        "descriptor": encode_range_image(torch.from_numpy(filled).float(), cfg).numpy(),
    }


# ------------------------------------------------------- float64 evaluation (for tests)
def edge_distance(points: np.ndarray, cfg: OracleConfig) -> Tuple[np.ndarray, np.ndarray]:
    """Distance in radians (float64, from the float32 inputs) of every point's azimuth and
    elevation to the nearest column / row edge -- the 'within 1e-5 of a bin edge' test of
    ``BASELINE.json:north_star``. Non-finite rows get distance 0."""
    p = np.asarray(points, dtype=np.float64)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    with np.errstate(invalid="ignore"):
        az = np.arctan2(y, x) + np.pi
        el = np.arctan2(z, np.sqrt(x * x + y * y))
    cw = 2 * np.pi / cfg.n_azimuth
    d_az = np.abs(az - np.round(az / cw) * cw)
    rw = (float(cfg.el_max) - float(cfg.el_min)) / cfg.n_elevation
    q = np.clip(np.round((el - float(cfg.el_min)) / rw), 1, cfg.n_elevation - 1)  # outer edges clamp
    d_el = np.abs(el - (float(cfg.el_min) + q * rw))
    bad = ~np.isfinite(d_az) | ~np.isfinite(d_el)
    d_az[bad] = 0.0
    d_el[bad] = 0.0
    return d_az, d_el


def range_edge_mask(points: np.ndarray, cfg: OracleConfig, rel: float = 1e-6) -> np.ndarray:
    """True for points whose float64 range is within ``rel`` (relative) of min/max range."""
    p = np.asarray(points, dtype=np.float64)
    r = np.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2 + p[:, 2] ** 2)
    with np.errstate(invalid="ignore"):
        m = (np.abs(r - cfg.min_range) <= rel * cfg.min_range) | (np.abs(r - cfg.max_range) <= rel * cfg.max_range)
    return m


def strip_ambiguous(points: np.ndarray, cfg: OracleConfig, tol: float = 1e-5) -> np.ndarray:
    """Remove the points the north star excuses (within ``tol`` rad of a bin edge); on the
    remainder the range image must be bit-identical between the CUDA path and the oracle."""
    d_az, d_el = edge_distance(points, cfg)
    finite = np.isfinite(np.asarray(points)[:, :3]).all(axis=1)
    keep = (~finite) | ((d_az > tol) & (d_el > tol))
    return np.ascontiguousarray(np.asarray(points)[keep])


def descriptor_f64(points: np.ndarray, cfg: OracleConfig) -> np.ndarray:
    """Same pipeline with a float64 spectrum on the float32 interpolated image: the 'truth'
    against which both the reference's float32 result and the GPU result are compared."""
    img = project(points, cfg)
    if cfg.interpolate_empty:
        img = interpolate_range_image(img)
    t = pool_rows(torch.from_numpy(img).float(), cfg).numpy().astype(np.float64)
    mag = np.abs(np.fft.rfft(t, axis=1))
    lut = freq_to_bin(cfg).numpy()
    hist = np.zeros((t.shape[0], cfg.n_bins))
    for k in range(cfg.n_freqs):
        hist[:, lut[k]] += mag[:, k]
    h = hist.reshape(-1)
    s = h.sum()
    return h / (s + cfg.epsilon) if s > cfg.epsilon else np.full_like(h, 1.0 / h.size)
