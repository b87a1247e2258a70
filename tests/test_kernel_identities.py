"""Arithmetic identities two kernels rely on, checked exhaustively in plain Python (no GPU, no library):
they restate the table builder of the occupancy-grid kernel (prepost.cu `occ_axis`) and the validity columns of
the uint8 initial block (umma_initial.cu `u8_k_valid`) next to the straightforward forms they replace.  The GPU
parity tests check the kernels themselves; these say WHY the shortcuts are exact."""
import itertools

import numpy as np


def _four_tap(img, sx, sy, ax, ay):
    """cv::remap INTER_LINEAR, BORDER_CONSTANT 0, on (label + 1) & 255 with 5-bit weights: the form K9 used to compute"""
    rows, cols = img.shape
    acc = 512
    for dy, wy in ((0, 32 - ay), (1, ay)):
        for dx, wx in ((0, 32 - ax), (1, ax)):
            y, x = sy + dy, sx + dx
            if 0 <= y < rows and 0 <= x < cols:
                acc += ((int(img[y, x]) + 1) & 255) * wy * wx
    return acc >> 10


def _occ_axis(s, a, n):
    """prepost.cu occ_axis: first index of a 2-pixel block that lies inside [0, n) and the weights of its two pixels"""
    if 0 <= s and s + 1 < n:
        return s, 32 - a, a
    if s == -1:
        return 0, a, 0
    if s == n - 1:
        return n - 2, 0, 32 - a
    return 0, 0, 0


def _fixed_block(img, sx, sy, ax, ay):
    cols = img.shape[1]
    bx, wx0, wx1 = _occ_axis(sx, ax, cols)
    by, wy0, wy1 = _occ_axis(sy, ay, img.shape[0])
    p = lambda y, x: (int(img[y, x]) + 1) & 255
    top = p(by, bx) * wx0 + p(by, bx + 1) * wx1            # dp2a, row 0
    bot = p(by + 1, bx) * wx0 + p(by + 1, bx + 1) * wx1    # dp2a, row 1
    return (top * wy0 + bot * wy1 + 512) >> 10             # dp2a over the two rows


def test_fixed_2x2_block_equals_clamped_four_tap_blend():
    rng = np.random.default_rng(0)
    for rows, cols in ((2, 2), (2, 5), (3, 2), (4, 7)):
        img = rng.integers(0, 256, (rows, cols)).astype(np.uint8)      # 255 wraps to 0 (np.add(segmap, 1), bev.py:177)
        img[0, 0] = 255
        for sy, sx in itertools.product(range(-3, rows + 2), range(-3, cols + 2)):
            for ay, ax in ((0, 0), (0, 31), (31, 0), (13, 7), (31, 31), (16, 16)):
                assert _fixed_block(img, sx, sy, ax, ay) == _four_tap(img, sx, sy, ax, ay), (rows, cols, sy, sx, ay, ax)


def test_initial_block_validity_columns():
    """sum over the 9 taps of valid(tap) * T(tap) == the four grouped columns, for every (first row?, first column?)"""
    rng = np.random.default_rng(1)
    T = rng.normal(size=(3, 3))
    col = {27: 0.0, 28: 0.0, 29: 0.0, 30: 0.0}
    k_valid = lambda ky, kx: (27 if kx == 0 else 28) if ky == 0 else (29 if kx == 0 else 30)    # umma_initial.cu u8_k_valid
    for ky in range(3):
        for kx in range(3):
            col[k_valid(ky, kx)] += T[ky, kx]
    for oy_pos, ox_pos in itertools.product((False, True), repeat=2):
        # 3x3 stride-2 pad-1 window of output pixel (oy, ox): row ky = 0 is padding iff oy == 0, column kx = 0 iff ox == 0
        want = sum(T[ky, kx] for ky in range(3) for kx in range(3) if (ky > 0 or oy_pos) and (kx > 0 or ox_pos))
        flags = {27: float(oy_pos and ox_pos), 28: float(oy_pos), 29: float(ox_pos), 30: 1.0}
        got = sum(flags[k] * col[k] for k in col)
        assert abs(got - want) < 1e-12
