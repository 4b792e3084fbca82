"""Overlapping-window split / overlap-add merge of wide latents -- mirror of diff_collage/w_img.py:8-48.

The reference goes through F.unfold / F.fold; windows here are plain slices (same values, no im2col copy of the
whole tensor).  Tiles are ordered batch-major, tile-minor: row b*n + i is window i of sample b.
"""
import torch as th

BASE_LEN = 128  # the reference hard-codes the latent length of one window (w_img.py:13)


def split_wimg(wimg, n_img, rtn_overlap=True):
    if wimg.ndim == 3:
        wimg = wimg[None]
    b, c, h, w = wimg.shape
    overlap_size = (n_img * BASE_LEN - w) // (n_img - 1)
    assert n_img * BASE_LEN - overlap_size * (n_img - 1) == w
    stride = BASE_LEN - overlap_size
    tiles = th.stack([wimg[:, :, :, i * stride: i * stride + BASE_LEN] for i in range(n_img)], dim=1)
    tiles = tiles.reshape(b * n_img, c, h, BASE_LEN)
    if rtn_overlap:
        return tiles, overlap_size
    return tiles


def avg_merge_wimg(imgs, overlap_size, n=None, is_avg=True):
    bn, c, h, w = imgs.shape
    if n is None:
        n = bn
    b = bn // n
    stride = w - overlap_size
    total = n * w - (n - 1) * overlap_size
    tiles = imgs.reshape(b, n, c, h, w)
    out = imgs.new_zeros(b, c, h, total)
    for i in range(n):
        out[:, :, :, i * stride: i * stride + w] += tiles[:, i]
    if is_avg:
        counter = imgs.new_zeros(total)
        for i in range(n):
            counter[i * stride: i * stride + w] += 1
        return out / counter
    return out
