"""Synthetic rotating-digit sequences.  TEST / BENCH INPUT GENERATOR (no MNIST offline: the reference's ``rot-mnist.mat`` lives on
Google Drive, README.md:19, and torchvision cannot download here).

Mirrors the reference's data path for the rotating-MNIST task (paths relative to /root/reference/experiments):
  * ``data/mnist.py:149-160`` rotate_img: the un-rotated frame followed by ``scipy.ndimage.rotate(img, a, axes=(1, 2),
    reshape=False)`` for every angle a;
  * ``data/mnist.py:174-175``: angles = rad2deg(linspace(0, 2 pi, n_angles)[1:])  -> T = n_angles frames per sequence;
  * ``data/utils.py:13-14`` Dataset.__getitem__: float32, view (T,1,28,28), (x - 0.1307) / 0.3081.
The digit itself is procedural: a "3"-like glyph made of two stacked arcs, rasterised and blurred, with per-sample jitter of the arc
centres / radii / stroke width drawn from ``np.random.RandomState(seed)``.  Deterministic given
(N, T, seed): the GPU box regenerates exactly the array the golden ELBO vectors were made from (checked through a checksum).
"""
import numpy as np
from scipy.ndimage import gaussian_filter, rotate

MNIST_MEAN, MNIST_STD = 0.1307, 0.3081      # data/utils.py:7-8


def _arc(img, cy, cx, r, a0, a1, width, n=160):
    """rasterise an arc of radius r around (cy, cx) from angle a0 to a1 (radians, image coordinates) with a soft stroke"""
    yy, xx = np.mgrid[0:28, 0:28].astype(np.float64)
    for a in np.linspace(a0, a1, n):
        py, px = cy + r * np.sin(a), cx + r * np.cos(a)
        img += np.exp(-((yy - py) ** 2 + (xx - px) ** 2) / (2.0 * width ** 2))
    return img


def glyph(rs):
    """one 28 x 28 "3"-like stroke image in [0, 1] with jitter from the RandomState rs"""
    img = np.zeros((28, 28), dtype=np.float64)
    cx = 13.0 + rs.uniform(-1.0, 1.0)
    cy = 14.0 + rs.uniform(-1.0, 1.0)
    r_top, r_bot = 4.3 + rs.uniform(-0.4, 0.4), 5.0 + rs.uniform(-0.4, 0.4)
    width = 0.9 + rs.uniform(-0.15, 0.25)
    lean = rs.uniform(-0.25, 0.25)
    # upper bowl: open to the left; lower bowl: open to the left, slightly larger (angles measured clockwise from +x in image coords)
    _arc(img, cy - r_top, cx, r_top, -0.80 * np.pi + lean, 0.50 * np.pi + lean, width)
    _arc(img, cy + r_bot, cx, r_bot, -0.50 * np.pi + lean, 0.80 * np.pi + lean, width)
    img = gaussian_filter(img, 0.6)
    img = img / img.max()
    return np.clip(img, 0.0, 1.0)


def rotate_img(img, angles):
    """data/mnist.py:149-160: (n,28,28) -> (n, 1 + len(angles), 28, 28), frame 0 un-rotated"""
    frames = [np.array(img).reshape((-1, 1, 28, 28))]
    for a in angles:
        frames.append(rotate(img, a, axes=(1, 2), reshape=False).reshape((-1, 1, 28, 28)))
    return np.concatenate(frames, axis=1)


def rotating_sequences(N, T=16, seed=121, normalise=True):
    """(N, T, 1, 28, 28) float32: N jittered glyphs, each rotated through T - 1 uniformly spaced angles of a full turn
    (data/mnist.py:174-175), scaled to [0, 1], then normalised like data/utils.py:13-14."""
    rs = np.random.RandomState(seed)
    imgs = np.stack([glyph(rs) for _ in range(N)])                          # (N,28,28)
    angles = np.rad2deg(np.linspace(0, 2 * np.pi, T)[1:])
    seq = rotate_img(imgs, angles)                                          # (N,T,28,28)
    seq = np.clip(seq, 0.0, 1.0).astype(np.float32).reshape(N, T, 1, 28, 28)
    if normalise:
        seq = (seq - np.float32(MNIST_MEAN)) / np.float32(MNIST_STD)
    return seq


def checksum(x):
    """order-sensitive fp64 checksum of an array (regenerated inputs must match the golden file's)"""
    v = np.asarray(x, dtype=np.float64).ravel()
    return float(np.dot(v, np.cos(np.arange(v.size, dtype=np.float64) * 0.37)))
