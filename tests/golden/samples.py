"""The reference's 21 shipped sample pairs (data/samples/**) as committed uint8 fixtures, and the
synthetic fused images the parity tests pair them with.  ``samples.npz`` is written by ``make_golden_samples.py`` (which
reads the images from /root/reference with cv2, exactly as eval.py:176-181 does); the GPU box only reads the .npz files."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = ('max', 'avg_round', 'avg_noise')          # max(a,b), round((a+b)/2), (a+b)/2 + noise
_cache = {}


def _npz():
    if 'z' not in _cache:
        _cache['z'] = np.load(os.path.join(HERE, 'samples.npz'))
    return _cache['z']


def names():
    """['polar/1.jpg', ..., 'infrared/36.png'] — the 21 pairs in a fixed order."""
    z = _npz()
    return [str(n) for n in z['names']]


def pair(name):
    """(img1, img2) uint8 (H, W): polar = (vis, po), infrared = (vis, ir) — the eval.py / test.py source order."""
    z = _npz()
    return z[name + '/a'], z[name + '/b']


def fused(a, b, kind, seed):
    """Synthetic fused image on the 0..255 scale, float32 (H, W).  'max' and 'avg_round' are integer-valued (what an 8-bit
    result file gives eval.py); 'avg_noise' is fractional and leaves [0, 255] here and there (histogram edge rule)."""
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    if kind == 'max':
        return np.maximum(a32, b32)
    if kind == 'avg_round':
        return np.floor((a32 + b32) * np.float32(0.5) + np.float32(0.5)).astype(np.float32)
    if kind == 'avg_noise':
        rng = np.random.default_rng(1000003 + seed)
        return ((a32 + b32) * np.float32(0.5) + rng.normal(0, 1, a.shape).astype(np.float32) * np.float32(4.0)).astype(np.float32)
    raise ValueError(kind)


def case(name, kind):
    """-> img1, img2, imgf float32 (1,1,H,W) on the 0..255 scale (the metric convention, eval.py:176-194).  The loss
    convention is the same three divided by float32 255 (data/dataset.py feeds uint8 / 255)."""
    a, b = pair(name)
    f = fused(a, b, kind, names().index(name))
    return a.astype(np.float32)[None, None], b.astype(np.float32)[None, None], f[None, None]


def unit(x):
    return (x / np.float32(255.0)).astype(np.float32)


def grad_probe(shape, seed, n=4096, keep=512):
    """Flat indices of the gradient elements the golden file keeps (a fingerprint of the full-size fp64 gradient): the
    first `keep` of `n` seeded draws (the file was first written with all 4096 and then trimmed)."""
    size = int(np.prod(shape))
    return np.random.default_rng(77 + seed).integers(0, size, size=min(n, size))[:keep]


def joint_checksum(j):
    """Two 64-bit weighted sums of a 256x256 integer joint histogram + its number of non-empty bins."""
    j = np.asarray(j, dtype=np.int64).reshape(-1)
    idx = np.arange(j.size, dtype=np.int64)
    return np.array([int((j * (idx + 1)).sum()), int((j * ((idx * 2654435761) % 4294967296)).sum()), int((j != 0).sum())], dtype=np.int64)
