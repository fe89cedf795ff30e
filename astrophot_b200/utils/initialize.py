"""Windows and start values from a segmentation map (reference: `utils/initialize/segmentation_map.py:33-342`; SURVEY.md
§8f-4, input prep).  Host numpy, vectorised over the segments (one labelled pass per quantity instead of one boolean
mask per segment).  Maps and images are arrays or ``.npy`` paths; FITS needs astropy, which this package does not use.

Every function returns a dict keyed by segment id, like the reference's."""
from copy import deepcopy

import numpy as np
import torch
from scipy import ndimage

__all__ = ("auto_variance", "centroids_from_segmentation_map", "PA_from_segmentation_map", "q_from_segmentation_map",
           "windows_from_segmentation_map", "scale_windows", "filter_windows", "transfer_windows")


def _array(a, hdul_index=0):
    if isinstance(a, str):
        if a.endswith(".npy"):
            return np.load(a)
        if a.endswith(".fits"):
            raise ValueError("reading FITS needs astropy, which astrophot_b200 does not depend on; pass an array or .npy")
        raise ValueError(f"unrecognized file type, should be one of: fits, npy\n{a}")
    return np.asarray(a)


def _labels(seg_map, skip_index):
    """(ids kept, dense label image 1..K with 0 = skipped)"""
    ids, inv = np.unique(seg_map, return_inverse=True)
    keep = np.array([i is not None and i not in skip_index for i in ids])
    dense = np.where(keep, np.cumsum(keep), 0)[inv.reshape(seg_map.shape)]
    return ids[keep], dense


def _segment_sums(dense, n, values):
    return np.bincount(dense.ravel(), weights=np.ravel(values), minlength=n + 1)[1:]


def centroids_from_segmentation_map(seg_map, image, hdul_index_seg=0, hdul_index_img=0, skip_index=(0,)):
    """Flux-weighted centroid (x, y) of every segment, pixel coordinates."""
    seg_map, image = _array(seg_map, hdul_index_seg), _array(image, hdul_index_img)
    ids, dense = _labels(seg_map, skip_index)
    yy, xx = np.indices(seg_map.shape)
    tot = _segment_sums(dense, len(ids), image)
    cx = _segment_sums(dense, len(ids), xx * image) / tot
    cy = _segment_sums(dense, len(ids), yy * image) / tot
    return {i: [x, y] for i, x, y in zip(ids, cx, cy)}


def _offsets(seg_map, ids, dense, centroids):
    yy, xx = np.indices(seg_map.shape)
    cx = np.concatenate(([0.0], [centroids[i][0] for i in ids]))
    cy = np.concatenate(([0.0], [centroids[i][1] for i in ids]))
    return xx - cx[dense], yy - cy[dense]


def PA_from_segmentation_map(seg_map, image, centroids=None, hdul_index_seg=0, hdul_index_img=0, skip_index=(0,),
                             north=np.pi / 2):
    """Position angle of every segment: second angular moment of its light about its centroid, plus ``north``."""
    seg_map, image = _array(seg_map, hdul_index_seg), _array(image, hdul_index_img)
    if centroids is None:
        centroids = centroids_from_segmentation_map(seg_map, image, skip_index=skip_index)
    ids, dense = _labels(seg_map, skip_index)
    dx, dy = _offsets(seg_map, ids, dense, centroids)
    th2 = 2 * np.arctan2(dy, dx)
    tot = _segment_sums(dense, len(ids), image)
    c = _segment_sums(dense, len(ids), image * np.cos(th2)) / tot
    s = _segment_sums(dense, len(ids), image * np.sin(th2)) / tot
    return {i: pa for i, pa in zip(ids, np.arctan2(s, c) / 2 % np.pi + north)}


def q_from_segmentation_map(seg_map, image, centroids=None, PAs=None, hdul_index_seg=0, hdul_index_img=0,
                            skip_index=(0,), north=np.pi / 2):
    """Axis ratio of every segment from the light-weighted sin^2 / cos^2 of the angle to its major axis."""
    seg_map, image = _array(seg_map, hdul_index_seg), _array(image, hdul_index_img)
    if centroids is None:
        centroids = centroids_from_segmentation_map(seg_map, image, skip_index=skip_index)
    if PAs is None:
        PAs = PA_from_segmentation_map(seg_map, image, centroids=centroids, skip_index=skip_index)
    ids, dense = _labels(seg_map, skip_index)
    dx, dy = _offsets(seg_map, ids, dense, centroids)
    pa = np.concatenate(([0.0], [PAs[i] + north for i in ids]))
    th = np.arctan2(dy, dx) - pa[dense]
    tot = _segment_sums(dense, len(ids), image)
    c2 = _segment_sums(dense, len(ids), image * np.cos(th) ** 2) / tot
    s2 = _segment_sums(dense, len(ids), image * np.sin(th) ** 2) / tot
    return {i: q for i, q in zip(ids, s2 / np.maximum(s2, c2))}


def windows_from_segmentation_map(seg_map, hdul_index=0, skip_index=(0,)):
    """Bounding box [[xmin, xmax], [ymin, ymax]] (inclusive pixel indices) of every segment."""
    seg_map = _array(seg_map, hdul_index)
    ids, dense = _labels(seg_map, skip_index)
    out = {}
    for i, box in zip(ids, ndimage.find_objects(dense, max_label=len(ids))):
        ys, xs = box
        out[i] = [[np.int64(xs.start), np.int64(xs.stop - 1)], [np.int64(ys.start), np.int64(ys.stop - 1)]]
    return out


def scale_windows(windows, image_shape=None, expand_scale=1.0, expand_border=0.0):
    """Grow every window about its centre by ``expand_scale`` and then by ``expand_border`` pixels on each side
    (truncated to integers), clipped to the image when ``image_shape`` is given."""
    out = {}
    for key, win in windows.items():
        new = []
        for axis, (lo, hi) in enumerate(deepcopy(win)):
            mid, half = (lo + hi) / 2, expand_scale * (hi - lo) / 2 + expand_border
            a, b = int(mid - half), int(mid + half)
            if image_shape is not None:
                a, b = max(0, a), min(image_shape[1 - axis], b)
            new.append([a, b])
        out[key] = new
    return out


def filter_windows(windows, min_size=None, max_size=None, min_area=None, max_area=None, min_flux=None, max_flux=None,
                   image=None):
    """Keep the windows whose smaller side, larger side, area and enclosed flux are inside the given bounds."""
    out = {}
    for key, ((x0, x1), (y0, y1)) in windows.items():
        w, h = x1 - x0, y1 - y0
        flux = np.sum(image[y0:y1, x0:x1]) if (min_flux is not None or max_flux is not None) else None
        tests = ((min_size, min(w, h), 1), (max_size, max(w, h), -1), (min_area, w * h, 1), (max_area, w * h, -1),
                 (min_flux, flux, 1), (max_flux, flux, -1))
        if all(bound is None or sign * (val - bound) >= 0 for bound, val, sign in tests):
            out[key] = windows[key]
    return out


def transfer_windows(windows, base_image, new_image):
    """The same sky regions as pixel windows of another image (different origin, pixel scale or rotation)."""
    top = np.array([float(v) for v in new_image.shape]) - 1
    out = {}
    for key, ((x0, x1), (y0, y1)) in windows.items():
        def to_new(x, y):
            return new_image.plane_to_pixel(base_image.pixel_to_plane(torch.tensor([x, y]))).detach().cpu().numpy()
        lo = np.clip(np.floor(to_new(x0, y0)), a_min=0, a_max=top)
        hi = np.clip(np.ceil(to_new(x1, y1)), a_min=0, a_max=top)
        out[key] = [[lo[0], hi[0]], [lo[1], hi[1]]]
    return out


def auto_variance(data, mask=None):
    """Variance map estimated from the image itself (reference: `utils/initialize/variance.py:12-55`): the scatter of
    the pixels about a lightly smoothed copy of the image, binned by flux, is fitted with a straight line in flux
    (read noise + Poisson term) and evaluated at every pixel; masked pixels get infinite variance.  Images too small
    or too flat for that get a constant."""
    from scipy.ndimage import gaussian_filter
    from scipy.stats import binned_statistic

    from ..errors import InvalidData

    if isinstance(data, torch.Tensor):
        data = data.detach().cpu().numpy()
    if isinstance(mask, torch.Tensor):
        mask = mask.detach().cpu().numpy()
    if mask is None:
        mask = np.zeros(data.shape, dtype=int)
    good = np.logical_not(mask)
    flat = np.var(data[good])
    if not np.isfinite(flat) or flat == 0:
        return np.ones_like(data)
    if min(data.shape) < 20:
        return np.ones_like(data) * flat
    inner = (slice(4, -4), slice(4, -4))                       # away from the filter's edge effects
    clean = gaussian_filter(mask, 1.1)[inner] == 0             # pixels whose smoothing kernel saw no masked pixel
    flux = data[inner][clean]
    resid = (data[inner] - gaussian_filter(data, 1.1)[inner])[clean]
    lo, hi = np.quantile(data[good], 0.01), np.quantile(data[good], 0.99)
    std, edges, _ = binned_statistic(flux.flatten(), resid.flatten(), statistic="std", bins=np.linspace(lo, hi, 11))
    left = edges[:-1]
    empty = ~np.isfinite(std)
    if np.any(empty):
        std[empty] = np.sqrt(np.interp(left[empty], left[~empty], std[~empty] ** 2))
    slope, offset = np.polyfit(left[:-2], std[:-2] ** 2, 1)   # the two brightest bins are left out
    if slope < 0:
        raise InvalidData("Variance appears to be decreasing with flux! Cannot accurately estimate variance.")
    variance = np.clip(slope * data + offset, np.min(std) ** 2, None)
    variance[~good] = np.inf
    return variance
