"""Parity metrics between two renders of the same frame:
  depth     bit-exact (implies coverage and every winner's z)
  coverage  depth < 1e10 mask, bit-exact
  RGB       compared after savePPM's 8-bit quantiser: |delta| <= 1 LSB everywhere except at most
            0.1 % of pixels (float rounding of pow/normalisation at a quantisation boundary)
"""
import numpy as np

import pyoracle

RGB_LSB_TOL = 1          # per channel, in 8-bit units
RGB_OUTLIER_FRACTION = 1e-3


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def compare(got_image, got_depth, want_image, want_depth):
    d_got, d_want = bits(got_depth), bits(want_depth)
    depth_mismatch = int((d_got != d_want).sum())
    cov_mismatch = int(((got_depth < 1e10) != (want_depth < 1e10)).sum())
    q_got = pyoracle.quantize_rgb8(got_image).astype(np.int16)
    q_want = pyoracle.quantize_rgb8(want_image).astype(np.int16)
    diff = np.abs(q_got - q_want).max(axis=-1)
    with np.errstate(invalid="ignore"):
        fdiff = np.abs(got_image.astype(np.float64) - want_image.astype(np.float64))
    return dict(depth_mismatch=depth_mismatch, coverage_mismatch=cov_mismatch,
                rgb_over_1lsb=int((diff > RGB_LSB_TOL).sum()), rgb_any_diff=int((diff > 0).sum()),
                rgb_max_lsb=int(diff.max()) if diff.size else 0,
                float_rgb_bit_mismatch=int((bits(got_image) != bits(want_image)).any(axis=-1).sum()),
                float_rgb_max_abs=float(np.nanmax(fdiff)) if fdiff.size else 0.0,
                pixels=int(diff.size), covered=int((want_depth < 1e10).sum()))


def assert_parity(rep, what=""):
    assert rep["depth_mismatch"] == 0, "%s depth differs in %d pixels: %r" % (what, rep["depth_mismatch"], rep)
    assert rep["coverage_mismatch"] == 0, "%s coverage differs: %r" % (what, rep)
    assert rep["rgb_over_1lsb"] <= RGB_OUTLIER_FRACTION * rep["pixels"], "%s RGB beyond 1 LSB: %r" % (what, rep)
