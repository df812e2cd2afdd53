"""Restatement of ``skimage.draw.disk`` -- test infrastructure, not product.

The reference rasterises every ship / laser with
``draw.disk((self.y, self.x), radius=self.radius, shape=grid.shape)``
(/root/reference/ofighters/lib/form.py:222-228).  scikit-image is an un-pinned
third-party dependency (requirements.txt:5) that is not installed here, so its
published algorithm (``disk -> ellipse -> _ellipse_in_shape``, rotation 0) is
restated:

    upper_left  = ceil(center - R)   (as int), clipped to >= 0
    lower_right = floor(center + R)  (as int), clipped to <= shape-1
    shifted     = center - upper_left                       (fp64)
    pixel (i, j) of the bounding box is set iff
        ((i - shifted_r)/R)**2 + ((j - shifted_c)/R)**2 < 1  (strict, fp64)

Known-answer vectors from the upstream docstrings are checked in
tests/test_oracle.py.
"""
import numpy as np


def disk(center, radius, *, shape=None):
    r, c = center
    center = np.array([r, c], dtype=np.float64)
    radii = np.array([radius, radius], dtype=np.float64)
    upper_left = np.ceil(center - radii).astype(int)
    lower_right = np.floor(center + radii).astype(int)
    if shape is not None:
        upper_left = np.maximum(upper_left, np.array([0, 0]))
        lower_right = np.minimum(lower_right, np.array(shape[:2]) - 1)
    shifted_center = center - upper_left
    bounding_shape = lower_right - upper_left + 1
    if bounding_shape[0] <= 0 or bounding_shape[1] <= 0:
        e = np.zeros((0,), dtype=np.intp)
        return e, e.copy()
    r_lim, c_lim = np.ogrid[0:float(bounding_shape[0]), 0:float(bounding_shape[1])]
    rr = (r_lim - shifted_center[0])
    cc = (c_lim - shifted_center[1])
    distances = (rr / radii[0]) ** 2 + (cc / radii[1]) ** 2
    ri, ci = np.nonzero(distances < 1)
    ri = ri + upper_left[0]
    ci = ci + upper_left[1]
    return ri, ci
