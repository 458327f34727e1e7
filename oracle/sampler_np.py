"""numpy restatement of the sampling index arithmetic the reference inherits from PyTorch
(TEST INFRASTRUCTURE; see oracle/eamm_oracle.py header).  Pins, independently of torch:

  * F.grid_sample(bilinear, padding zeros, align_corners=False)   dense_motion.py:77, generator.py:57
        pixel = ((g + 1) * size - 1) / 2 ; taps floor(pixel), floor(pixel)+1 ; out-of-range taps read 0
  * F.interpolate(mode='bilinear', align_corners=False)           generator.py:55
        src = max((dst + 0.5) * in/out - 0.5, 0) ; upper tap clamped to in-1
  * F.interpolate(scale_factor=2) nearest                         util.py:896
        src = floor(dst / 2)
"""
import numpy as np


def unnormalize(g, size):
    return ((g + 1.0) * size - 1.0) / 2.0


def grid_sample_taps(gx, gy, W, H):
    """Integer taps and weights for one grid point: [(y, x, weight, in_bounds)] * 4 (nw, ne, sw, se)."""
    ix, iy = unnormalize(np.float32(gx), W), unnormalize(np.float32(gy), H)
    x0, y0 = int(np.floor(ix)), int(np.floor(iy))
    wx1, wy1 = np.float32(ix - x0), np.float32(iy - y0)
    taps = []
    for dy, wy in ((0, np.float32(1) - wy1), (1, wy1)):
        for dx, wx in ((0, np.float32(1) - wx1), (1, wx1)):
            y, x = y0 + dy, x0 + dx
            taps.append((y, x, np.float32(wy * wx), 0 <= x < W and 0 <= y < H))
    return taps


def grid_sample(img, grid):
    """img [C,H,W], grid [h,w,2] (x,y) -> [C,h,w]; pure-python loops, for small cases only."""
    C, H, W = img.shape
    h, w, _ = grid.shape
    out = np.zeros((C, h, w), dtype=np.float32)
    for i in range(h):
        for j in range(w):
            for (y, x, wt, ok) in grid_sample_taps(grid[i, j, 0], grid[i, j, 1], W, H):
                if ok:
                    out[:, i, j] += img[:, y, x] * wt
    return out


def bilinear_upsample_index(dst, in_size, out_size):
    src = max((dst + 0.5) * (in_size / out_size) - 0.5, 0.0)
    i0 = int(np.floor(src))
    i1 = min(i0 + 1, in_size - 1)
    return i0, i1, np.float32(src - i0)


def nearest_up2_index(dst):
    return dst // 2
