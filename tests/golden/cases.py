"""Deterministic input generators shared by ``make_golden.py`` and the tests.

Pure numpy (``default_rng`` / PCG64 is stable across numpy versions and hosts),
so the GPU box regenerates exactly the inputs the golden outputs were computed
on.  ``crops.npz`` holds a few small uint8 crops of natural images (the only
inputs that cannot be regenerated from a seed).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def uniform01(seed, shape):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def u8(seed, shape):
    return np.random.default_rng(seed).integers(0, 256, size=shape).astype(np.float32)


def frac255(seed, shape):
    return (np.random.default_rng(seed).random(shape, dtype=np.float32) * np.float32(255.0)).astype(np.float32)


def natural(seed, shape, flats=True):
    """Smooth random field + edges + flat patches, quantised to 0..255 (uint8-like float32)."""
    rng = np.random.default_rng(seed)
    h, w = shape[-2:]
    lead = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    out = np.empty((lead, h, w), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    for n in range(lead):
        img = np.zeros((h, w))
        for _ in range(6):
            fy, fx = rng.uniform(0.002, 0.06, 2)
            ph = rng.uniform(0, 2 * np.pi, 2)
            img += rng.uniform(10, 40) * np.sin(2 * np.pi * fy * yy + ph[0]) * np.cos(2 * np.pi * fx * xx + ph[1])
        img += 128 + rng.normal(0, 3.0, (h, w))
        # a few step edges
        for _ in range(3):
            y0, x0 = rng.integers(0, h), rng.integers(0, w)
            img[y0:, x0:] += rng.uniform(-40, 40)
        if flats:
            for _ in range(3):
                y0, x0 = rng.integers(0, max(1, h - 8)), rng.integers(0, max(1, w - 8))
                hh, ww = rng.integers(8, max(9, h // 3)), rng.integers(8, max(9, w // 3))
                img[y0:y0 + hh, x0:x0 + ww] = rng.integers(0, 256)
        out[n] = np.clip(np.rint(img), 0, 255)
    return out.reshape(shape)


def crops():
    z = np.load(os.path.join(HERE, 'crops.npz'))
    return {k: z[k] for k in z.files}


def fuse(a, b, kind, seed=0):
    """Synthetic fused image from two sources (SURVEY.md §8(c) fixture kinds)."""
    if kind == 'max':
        return np.maximum(a, b)
    if kind == 'avg_floor':
        return np.floor((a + b) * np.float32(0.5)).astype(np.float32)
    if kind == 'avg':
        return ((a + b) * np.float32(0.5)).astype(np.float32)
    if kind == 'avg_noise':
        rng = np.random.default_rng(seed + 7919)
        scale = np.float32(4.0) if a.max() > 2 else np.float32(4.0 / 255.0)
        return ((a + b) * np.float32(0.5) + rng.normal(0, 1, a.shape).astype(np.float32) * scale).astype(np.float32)
    if kind == 'rand01':
        return uniform01(seed + 104729, a.shape)
    if kind == 'randu8':
        return u8(seed + 104729, a.shape)
    if kind == 'randfrac':
        return frac255(seed + 104729, a.shape)
    raise ValueError(kind)


def _c(name):
    return crops()[name].astype(np.float32)


# ---- loss cases: (name, builder) -> (img1, img2, imgf) each (B,1,H,W) float32 in ~[0,1] ----
def loss_case(name):
    if name == 'rand_2x40x37':
        a, b = uniform01(1, (2, 1, 40, 37)), uniform01(2, (2, 1, 40, 37))
        return a, b, fuse(a, b, 'rand01', 3)
    if name == 'rand_1x11x11':
        a, b = uniform01(4, (1, 1, 11, 11)), uniform01(5, (1, 1, 11, 11))
        return a, b, fuse(a, b, 'rand01', 6)
    if name == 'rand_3x64x96':
        a, b = uniform01(7, (3, 1, 64, 96)), uniform01(8, (3, 1, 64, 96))
        return a, b, fuse(a, b, 'rand01', 9)
    if name == 'rand_1x12x300':
        a, b = uniform01(10, (1, 1, 12, 300)), uniform01(11, (1, 1, 12, 300))
        return a, b, fuse(a, b, 'rand01', 12)
    if name == 'rand_1x300x13':
        a, b = uniform01(13, (1, 1, 300, 13)), uniform01(14, (1, 1, 300, 13))
        return a, b, fuse(a, b, 'rand01', 15)
    if name == 'rand_2x150x260':
        a, b = uniform01(16, (2, 1, 150, 260)), uniform01(17, (2, 1, 150, 260))
        return a, b, fuse(a, b, 'rand01', 18)
    if name == 'natural_2x96x128':
        a = natural(20, (2, 1, 96, 128)) / np.float32(255.0)
        b = natural(21, (2, 1, 96, 128)) / np.float32(255.0)
        return a, b, fuse(a, b, 'avg_noise', 22)
    if name == 'polar_crop_avg':
        a = _c('polar_vis')[None, None] / np.float32(255.0)
        b = _c('polar_po')[None, None] / np.float32(255.0)
        return a, b, fuse(a, b, 'avg_noise', 23)
    if name == 'ir_crop_max':
        a = _c('ir_vis')[None, None] / np.float32(255.0)
        b = _c('ir_ir')[None, None] / np.float32(255.0)
        return a, b, fuse(a, b, 'max')
    if name == 'constant':
        a = np.full((1, 1, 24, 40), 0.3, np.float32)
        b = np.full((1, 1, 24, 40), 0.6, np.float32)
        return a, b, np.full((1, 1, 24, 40), 0.5, np.float32)
    if name == 'unbounded_f':
        a, b = uniform01(30, (2, 1, 48, 56)), uniform01(31, (2, 1, 48, 56))
        f = (np.random.default_rng(32).normal(0.5, 0.6, a.shape)).astype(np.float32)
        return a, b, f
    raise KeyError(name)


LOSS_CASES = ['rand_2x40x37', 'rand_1x11x11', 'rand_3x64x96', 'rand_1x12x300', 'rand_1x300x13',
              'rand_2x150x260', 'natural_2x96x128', 'polar_crop_avg', 'ir_crop_max', 'constant',
              'unbounded_f']
# cases whose gradient has exact ties (sign(0)) and is excluded from the gradient gate
LOSS_GRAD_TIE_CASES = {'ir_crop_max', 'constant'}


# ---- metric cases -> (img1, img2, imgf) each (1,1,H,W) float32, nominally 0..255 ----
def metric_case(name):
    if name == 'u8_64x80':
        a, b = u8(40, (1, 1, 64, 80)), u8(41, (1, 1, 64, 80))
        return a, b, fuse(a, b, 'avg_floor')
    if name == 'u8_rand_f_97x131':
        a, b = u8(42, (1, 1, 97, 131)), u8(43, (1, 1, 97, 131))
        return a, b, fuse(a, b, 'randu8', 44)
    if name == 'frac_72x90':
        a, b = frac255(45, (1, 1, 72, 90)), frac255(46, (1, 1, 72, 90))
        return a, b, fuse(a, b, 'randfrac', 47)
    if name == 'natural_flat_97x131':
        a, b = natural(48, (1, 1, 97, 131)), natural(49, (1, 1, 97, 131))
        return a, b, fuse(a, b, 'avg_floor')
    if name == 'natural_max_120x152':
        a, b = natural(50, (1, 1, 120, 152)), natural(51, (1, 1, 120, 152))
        return a, b, fuse(a, b, 'max')
    if name == 'polar_crop':
        a, b = _c('polar_vis')[None, None], _c('polar_po')[None, None]
        return a, b, fuse(a, b, 'avg_floor')
    if name == 'ir_crop':
        a, b = _c('ir_vis')[None, None], _c('ir_ir')[None, None]
        return a, b, fuse(a, b, 'max')
    if name == 'polar_odd_153':
        # 128x153: exercises the odd-width reflect pad of the MS-SSIM pyramid (metric.py:389-392)
        a, b = natural(52, (1, 1, 128, 153), flats=False), natural(53, (1, 1, 128, 153), flats=False)
        return a, b, fuse(a, b, 'avg_floor')
    if name == 'u8_200x216':
        a, b = u8(54, (1, 1, 200, 216)), u8(55, (1, 1, 200, 216))
        return a, b, fuse(a, b, 'avg_floor')
    raise KeyError(name)


METRIC_CASES = ['u8_64x80', 'u8_rand_f_97x131', 'frac_72x90', 'natural_flat_97x131',
                'natural_max_120x152', 'polar_crop', 'ir_crop', 'polar_odd_153', 'u8_200x216']


def hist_edge_vector():
    """Histogram rule edge values."""
    v = np.array([0.0, -0.0, 255.0, 255.5, 255.99998, 256.0, 256.00003, -1e-7, -1.0, 300.0, np.nan,
                  1.0, 0.99999994, 127.5, 128.0, np.inf, -np.inf, 42.0, 42.0, 7.25], dtype=np.float32)
    w = np.array([5.0, 256.0, 255.0, 0.0, 13.0, 256.0, 1.0, 2.0, 3.0, 4.0, 5.0,
                  np.nan, 6.0, 255.99998, -0.0, 9.0, 10.0, 42.0, 43.0, 300.0], dtype=np.float32)
    return v.reshape(1, 1, 4, 5), w.reshape(1, 1, 4, 5)
