"""Seeded synthetic luma content: low-passed noise texture + gradient, translated by
non-integer velocities and slowly rotated, with moving hard-edged rectangles and a little per-frame noise, so that
motion vectors are non-trivial and sub-pel refinement matters.  Pure numpy; used by tests, bench.py and to write
planar YUV files for the reference encoder."""
import numpy as np


def _base_texture(rng, h, w):
    t = rng.integers(0, 256, size=(h + 2, w + 2)).astype(np.float32)
    k = (t[:-2, :-2] + t[:-2, 1:-1] + t[:-2, 2:] + t[1:-1, :-2] + t[1:-1, 1:-1] + t[1:-1, 2:] +
         t[2:, :-2] + t[2:, 1:-1] + t[2:, 2:]) / 9.0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    grad = 40.0 * np.sin(xx / 97.0) + 30.0 * np.cos(yy / 61.0)
    out = (k - 128.0) * 1.9 + 128.0 + grad
    return out


def _bilinear(img, ys, xs):
    h, w = img.shape
    ys = np.clip(ys, 0, h - 1.001)
    xs = np.clip(xs, 0, w - 1.001)
    y0 = np.floor(ys).astype(np.int32)
    x0 = np.floor(xs).astype(np.int32)
    fy = (ys - y0).astype(np.float32)
    fx = (xs - x0).astype(np.float32)
    a = img[y0, x0] * (1 - fx) + img[y0, x0 + 1] * fx
    b = img[y0 + 1, x0] * (1 - fx) + img[y0 + 1, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def luma_frame(width, height, t, seed=1234, bit_depth=8, vx=1.75, vy=-0.6, theta_deg=0.2, n_rect=6):
    """Frame t of the clip, uint8 (bit_depth 8) or uint16 (bit_depth 10: value << 2 plus two random LSBs)."""
    rng0 = np.random.default_rng(seed)
    pad = 96
    base = _base_texture(rng0, height + 2 * pad, width + 2 * pad)
    rects = [(rng0.integers(0, width), rng0.integers(0, height), rng0.integers(24, 120), rng0.integers(16, 90),
              rng0.uniform(-3.5, 3.5), rng0.uniform(-2.5, 2.5), int(rng0.integers(20, 236))) for _ in range(n_rect)]
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    cy, cx = height / 2.0, width / 2.0
    th = np.deg2rad(theta_deg * t)
    c, s = np.cos(th), np.sin(th)
    xs = c * (xx - cx) - s * (yy - cy) + cx + pad + vx * t
    ys = s * (xx - cx) + c * (yy - cy) + cy + pad + vy * t
    img = _bilinear(base, ys, xs)
    for (rx, ry, rw, rh, rvx, rvy, level) in rects:
        x0 = int(round(rx + rvx * t)) % width
        y0 = int(round(ry + rvy * t)) % height
        img[y0:min(height, y0 + rh), x0:min(width, x0 + rw)] = level
    rngt = np.random.default_rng(seed * 7919 + t)
    img = img + rngt.integers(-2, 3, size=img.shape)
    img8 = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if bit_depth == 8:
        return img8
    lsb = rngt.integers(0, 1 << (bit_depth - 8), size=img.shape).astype(np.uint16)
    return (img8.astype(np.uint16) << (bit_depth - 8)) | lsb


def pad_plane(samples, margin_x=80, margin_y=80):
    """TComPicYuv-style padded Pel (int16) buffer with replicated borders (TLibCommon/TComPicYuv.cpp:197-242)."""
    return np.ascontiguousarray(np.pad(samples.astype(np.int16), ((margin_y, margin_y), (margin_x, margin_x)), mode="edge"))


def write_yuv420(path, frames, bit_depth=8):
    """Planar 4:2:0 file as TVideoIOYuv reads it (8-bit bytes or 16-bit little endian); chroma is mid-grey."""
    with open(path, "wb") as f:
        for y in frames:
            h, w = y.shape
            if bit_depth == 8:
                f.write(y.astype(np.uint8).tobytes())
                f.write(np.full((h // 2) * (w // 2) * 2, 128, dtype=np.uint8).tobytes())
            else:
                f.write(y.astype("<u2").tobytes())
                f.write(np.full((h // 2) * (w // 2) * 2, 1 << (bit_depth - 1), dtype="<u2").tobytes())
