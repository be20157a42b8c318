"""Synthetic stereo pairs for the headless benchmark and the parity tests.

The reference's only sample pair (src/utils/left.tif, right.tif) is a missing blob
(.MISSING_LARGE_BLOBS), so every config in BASELINE.json runs on synthetic input:

  left   RGB u8, multi-octave value noise (octave periods 2,4,...,256 px, channels
         independent, counter-based SplitMix64 hash of (seed, channel, octave, ix, iy)),
         scaled into [16, 240] so that no 5x5 window is all zero (no 0/0 in NCC^2,
         MatchLib.cu:684);
  field  dH = A_h sin(2 pi x / P_x) cos(2 pi y / P_y) + B_h,  dV = A_v sin(2 pi y / P_y),
         |dH| <= 0.004 W, |dV| <= 2 px;
  right  bilinear resample of left at (x - dH, y - dV), rounded to u8, so that the pixel
         matching left (x, y) is right (x + dH, y + dV) -- the reference's sign convention
         (getPointCloud.cpp:912-913).

The ground-truth field is only a sanity check (end-point error); parity is always
CUDA-vs-oracle on the same bytes.
"""
import numpy as np

SEED_BASE = 0x5EED0000
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _lattice(seed, ch, octave, nx, ny):
    """Uniform [0,1) values on an (ny, nx) lattice, hashed from the coordinates."""
    ix = np.arange(nx, dtype=np.uint64)[None, :]
    iy = np.arange(ny, dtype=np.uint64)[:, None]
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0x100000001B3)
               + np.uint64(ch) * np.uint64(0x9E3779B1)
               + np.uint64(octave) * np.uint64(0x85EBCA77)) & _M64
        h = _splitmix64(key ^ _splitmix64(ix * np.uint64(0xD6E8FEB86659FD93) + iy * np.uint64(0xC2B2AE3D27D4EB4F)))
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))


def value_noise(seed, ch, width, height, periods=(2, 4, 8, 16, 32, 64, 128, 256)):
    acc = np.zeros((height, width), np.float32)
    x = np.arange(width, dtype=np.float32)
    y = np.arange(height, dtype=np.float32)
    for o, p in enumerate(periods):
        nx, ny = width // p + 2, height // p + 2
        lat = _lattice(seed, ch, o, nx, ny)
        gx, gy = x / p, y / p
        x0, y0 = np.floor(gx).astype(np.int64), np.floor(gy).astype(np.int64)
        fx, fy = (gx - x0).astype(np.float32), (gy - y0).astype(np.float32)
        # smoothstep weights keep the octave band-limited
        fx = fx * fx * (3 - 2 * fx)
        fy = fy * fy * (3 - 2 * fy)
        top = lat[y0][:, x0] * (1 - fx)[None, :] + lat[y0][:, x0 + 1] * fx[None, :]
        bot = lat[y0 + 1][:, x0] * (1 - fx)[None, :] + lat[y0 + 1][:, x0 + 1] * fx[None, :]
        acc += top * (1 - fy)[:, None] + bot * fy[:, None]
    acc /= len(periods)
    return acc


def disparity_field(width, height, k=0):
    """Ground-truth (dH, dV), float32 (height, width)."""
    x = np.arange(width, dtype=np.float32)[None, :]
    y = np.arange(height, dtype=np.float32)[:, None]
    Px, Py = 0.61 * width, 0.83 * height
    Ah = 0.0025 * width
    Bh = 0.0015 * width * (1.0 if k % 2 == 0 else -1.0)
    Av = min(2.0, 0.004 * height)
    dH = (Ah * np.sin(2 * np.pi * x / Px) * np.cos(2 * np.pi * y / Py) + Bh).astype(np.float32)
    dV = np.broadcast_to((Av * np.sin(2 * np.pi * y / Py)).astype(np.float32), (height, width)).copy()
    return dH, dV


def _bilinear(img, xs, ys):
    h, w = img.shape
    xs = np.clip(xs, 0, w - 1)
    ys = np.clip(ys, 0, h - 1)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    fx = (xs - x0).astype(np.float32)
    fy = (ys - y0).astype(np.float32)
    a = img[y0, x0] * (1 - fx) + img[y0, x1] * fx
    b = img[y1, x0] * (1 - fx) + img[y1, x1] * fx
    return a * (1 - fy) + b * fy


def make_pair(width, height, k=0):
    """-> (left_u8 (h,w,3), right_u8 (h,w,3), dH, dV) for pair index k."""
    seed = SEED_BASE + k
    dH, dV = disparity_field(width, height, k)
    xx = np.arange(width, dtype=np.float32)[None, :] - dH
    yy = np.arange(height, dtype=np.float32)[:, None] - dV
    left = np.empty((height, width, 3), np.uint8)
    right = np.empty((height, width, 3), np.uint8)
    for ch in range(3):
        v = value_noise(seed, ch, width, height)
        v = np.clip(0.5 + (v - 0.5) * 2.6, 0.0, 1.0)
        lf = 16.0 + 224.0 * v
        left[:, :, ch] = np.rint(lf).astype(np.uint8)
        right[:, :, ch] = np.clip(np.rint(_bilinear(lf, xx, yy)), 16, 240).astype(np.uint8)
    return left, right, dH, dV
