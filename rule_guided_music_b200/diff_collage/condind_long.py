"""CondIndSimple: eps of a long latent = sum of eps over overlapping 128-wide windows minus eps of the 64-wide
overlaps (conditional-independence factorisation) -- mirror of diff_collage/condind_long.py:8-51.

Two denoiser calls per evaluation, both batched over (sample, window): full windows (256 tokens) and trailing
half windows (128 tokens).  The reference also evaluates the last window's half tile and then zeroes it (:43); here
it is simply not evaluated (identical result, 1/(2n) fewer denoiser rows).
"""
import torch as th

from .generic_sampler import SimpleWork
from .w_img import avg_merge_wimg, split_wimg


def _window_eps(eps_scalar_t_fn, xs, scalar_t, y, num_img, overlap_size):
    """xs ((b n), c, h, w) windows -> per-window eps with the overlap correction applied, same layout."""
    bn, c, h, w = xs.shape
    b = bn // num_img
    y_rep = None if y is None else y.repeat_interleave(num_img)
    full_eps = eps_scalar_t_fn(xs, scalar_t.repeat_interleave(num_img), y=y_rep).reshape(b, num_img, c, h, w).clone()
    if num_img > 1:
        # every window but the last of a sample, as static slices (a boolean mask would need a device->host sync for
        # its size, which is also illegal inside a CUDA-graph capture); same row order as xs[keep]
        half_in = xs.reshape(b, num_img, c, h, w)[:, :-1, :, :, -overlap_size:].reshape(b * (num_img - 1), c, h,
                                                                                        overlap_size).contiguous()
        y_half = None if y is None else y.repeat_interleave(num_img - 1)
        half_eps = eps_scalar_t_fn(half_in, scalar_t.repeat_interleave(num_img - 1), y=y_half)
        full_eps[:, :-1, :, :, -overlap_size:] -= half_eps.reshape(b, num_img - 1, c, h, overlap_size)
    return full_eps.reshape(bn, c, h, w)


class CondIndSimple(SimpleWork):
    def __init__(self, shape, eps_scalar_t_fn, num_img, overlap_size=32):
        c, h, w = shape
        assert overlap_size == w // 2
        self.overlap_size = overlap_size
        self.num_img = num_img
        final_img_w = w * num_img - self.overlap_size * (num_img - 1)
        super().__init__((c, h, final_img_w), self.get_eps_t_fn(eps_scalar_t_fn))

    def loss(self, x):
        x1, x2 = x[:-1], x[1:]
        return th.sum((th.abs(x1[:, :, :, -self.overlap_size:] - x2[:, :, :, : self.overlap_size])) ** 2, dim=(1, 2, 3))

    def get_eps_t_fn(self, eps_scalar_t_fn):
        def eps_t_fn(long_x, scalar_t, y=None):
            xs = split_wimg(long_x, self.num_img, rtn_overlap=False)
            whole = _window_eps(eps_scalar_t_fn, xs, scalar_t, y, self.num_img, self.overlap_size)
            return avg_merge_wimg(whole, self.overlap_size, n=self.num_img, is_avg=False)

        return eps_t_fn
