"""Angle-binned footprint masks: the batched replacement for `get_pixel_footprint`
(reference utilities/path_tools.py:122-162, hook `get_pixel_footprint_impl` :101-103).

The reference rasterises the robot polygon for every pose: rotate footprint/res by theta, round the
vertices half-to-even, cv2.fillPoly.  The mask is a pure function of the *rounded vertex tuple*, and
that tuple only changes at the angles where some rotated coordinate crosses k + 1/2.  So the table
is exact, not an approximation:

  * `edges`   all such crossing angles in [-pi, pi), computed analytically and sorted;
  * one bin per interval: the rounded tuple at the bin midpoint and its filled mask, stored as bit
    rows relative to the mask's bounding box (diff-drive masks have non-contiguous rows, so spans
    would not do);
  * on device the bin found by binary search is *verified* against the tuple recomputed from the
    actual angle (and neighbours are probed on a miss), so edge round-off cannot pick a wrong mask.

Built once per (robot, resolution, footprint scale) on the host with the same cv2.fillPoly call the
reference makes; 3960 bins for the tricycle at 0.03 m.
"""
import numpy as np

try:
    import cv2
except ImportError:  # pragma: no cover
    cv2 = None


def rounded_vertices(angle, fp_pix):
    """round_half_even(R(angle) * fp_pix) with the reference's evaluation order
    (np.dot(footprint / res, m[2,2,1]), utilities/path_tools.py:142-150)."""
    c, s = np.cos(angle), np.sin(angle)
    m = np.array([[c, -s], [s, c]], dtype=np.float64).reshape(2, 2, 1)
    return np.round(np.dot(fp_pix, m)[:, :, 0]).astype(np.int32)


def _crossing_angles(fp_pix):
    """Angles in [-pi, pi) where some rotated vertex coordinate equals k + 1/2."""
    out = [np.array([-np.pi])]
    for a, b in fp_pix:
        r = np.hypot(a, b)
        if r < 0.5:
            continue
        phi = np.arctan2(b, a)
        k = np.arange(-np.floor(r + 0.5) - 1, np.floor(r + 0.5) + 2)
        q = (k + 0.5) / r
        q = q[np.abs(q) <= 1.0]
        ac = np.arccos(q)
        asn = np.arcsin(q)
        out += [ac - phi, -ac - phi, asn - phi, np.pi - asn - phi]   # x' = r cos(t+phi), y' = r sin(t+phi)
    ang = np.concatenate(out)
    ang = (ang + np.pi) % (2 * np.pi) - np.pi
    return np.unique(ang)


def fill_polygon_mask(int_verts):
    """Filled mask of an integer polygon: (mask uint8 [h, w], xmin, ymin) in polygon coordinates,
    rasterised by the reference's own raster routine (cv2.fillPoly, utilities/path_tools.py:159)."""
    if cv2 is None:
        raise RuntimeError("building the footprint table needs OpenCV (cv2.fillPoly), like the reference")
    half = np.abs(int_verts).max(axis=0) + 1
    canvas = np.zeros((2 * half[1] + 1, 2 * half[0] + 1), dtype=np.uint8)
    cv2.fillPoly(canvas, [np.ascontiguousarray(int_verts + half, dtype=np.int32)], (255, 255, 255))
    ys, xs = np.nonzero(canvas)
    y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
    return canvas[y0:y1 + 1, x0:x1 + 1] != 0, int(x0 - half[0]), int(y0 - half[1])


class FootprintLut(object):
    """Host-side table; `.arrays()` gives the flat numpy buffers the C-ABI's BcgFootprintLut wants."""

    def __init__(self, footprint, resolution):
        footprint = np.asarray(footprint, dtype=np.float64)
        assert footprint.ndim == 2 and footprint.shape[1] == 2
        if len(footprint) > 32:
            raise ValueError("footprints of more than 32 vertices are not supported")
        self.fp_pix = np.ascontiguousarray(footprint / resolution)
        self.resolution = float(resolution)
        edges = _crossing_angles(self.fp_pix)
        mids = 0.5 * (edges + np.append(edges[1:], np.pi))
        tuples = np.stack([rounded_vertices(t, self.fp_pix) for t in mids])        # [B, nv, 2]
        # merge neighbouring bins that round to the same tuple (duplicated crossings)
        keep = np.ones(len(edges), dtype=bool)
        keep[1:] = np.any(tuples[1:].reshape(len(edges) - 1, -1) != tuples[:-1].reshape(len(edges) - 1, -1), axis=1)
        self.edges = np.ascontiguousarray(np.append(edges[keep], np.pi))
        self.verts = np.ascontiguousarray(tuples[keep].astype(np.int16).reshape(keep.sum(), -1))
        self.n_bins = int(keep.sum())
        self.n_verts = len(footprint)
        masks, cache = [], {}
        for t in tuples[keep]:
            key = t.tobytes()
            if key not in cache:
                cache[key] = fill_polygon_mask(t)
            masks.append(cache[key])
        self.max_rows = (max(m.shape[0] for m, _, _ in masks) + 1) // 2 * 2     # even: the device loads rows in 16-byte pairs
        width = max(m.shape[1] for m, _, _ in masks)
        self.wpr = (width + 63) // 64
        self.header = np.zeros((self.n_bins, 4), dtype=np.int16)
        self.rows = np.zeros((self.n_bins, self.max_rows, self.wpr), dtype=np.uint64)
        self.pixels = np.zeros(self.n_bins, dtype=np.int32)
        weights = (np.uint64(1) << np.arange(64, dtype=np.uint64))
        for k, (m, x0, y0) in enumerate(masks):
            h, w = m.shape
            self.header[k] = (x0, y0, h, w)
            padded = np.zeros((h, self.wpr * 64), dtype=bool)
            padded[:, :w] = m
            self.rows[k, :h] = (padded.reshape(h, self.wpr, 64) * weights).sum(axis=2, dtype=np.uint64)
            self.pixels[k] = int(m.sum())

        # uniform bucket table: first candidate bin for an angle without a binary search
        self.n_buckets = 16384
        self.bucket_scale = self.n_buckets / (2 * np.pi)
        lefts = -np.pi + np.arange(self.n_buckets) / self.bucket_scale
        self.bucket_first = np.clip(np.searchsorted(self.edges, lefts, side="right") - 1, 0,
                                    self.n_bins - 1).astype(np.int32)

    def bin_of(self, angle):
        """Host twin of the device lookup (used by tests and by roofline accounting)."""
        t = angle if -np.pi <= angle < np.pi else float((angle + np.pi) % (2 * np.pi) - np.pi)
        k = int(np.searchsorted(self.edges, t, side="right")) - 1
        k = min(max(k, 0), self.n_bins - 1)
        want = rounded_vertices(angle, self.fp_pix).astype(np.int16).reshape(-1)
        for d in (0, 1, -1, 2, -2, 3, -3, 4, -4):
            j = (k + d) % self.n_bins
            if np.array_equal(self.verts[j], want):
                return j
        raise LookupError("angle %r falls in no table bin" % angle)

    def mask(self, k):
        """(bool mask [h, w], xmin, ymin) of bin k, decoded from the bit rows."""
        x0, y0, h, w = [int(v) for v in self.header[k]]
        bits = (self.rows[k, :h, :, None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)
        return bits.reshape(h, -1)[:, :w].astype(bool), x0, y0

    def canvas(self, angle):
        """The reference's own return format (utilities/path_tools.py:122-162, fill=True): uint8 image
        (2 hy + 1, 2 hx + 1) with the footprint in 255 about the centre pixel, half sizes = ceil of the largest
        rotated coordinate.  Array-for-array equal to get_pixel_footprint (reference test_path_tools.py:453-462)."""
        c, s = np.cos(angle), np.sin(angle)
        m = np.array([[c, -s], [s, c]], dtype=np.float64).reshape(2, 2, 1)
        rot = np.dot(self.fp_pix, m)[:, :, 0]
        half = np.ceil(np.maximum(rot.max(axis=0), -rot.min(axis=0))).astype(np.int32)
        mask, x0, y0 = self.mask(self.bin_of(angle))
        out = np.zeros((2 * half[1] + 1, 2 * half[0] + 1), dtype=np.uint8)
        h, w = mask.shape
        out[y0 + half[1]:y0 + half[1] + h, x0 + half[0]:x0 + half[0] + w] = mask.astype(np.uint8) * 255
        return out

    def arrays(self):
        # a bin's vertex tuple and header side by side, rows padded to 16 bytes: what the device lookup fetches per candidate
        stride = (2 * self.n_verts + 4 + 7) // 8 * 8
        bins = np.zeros((self.n_bins, stride), dtype=np.int16)
        bins[:, :2 * self.n_verts] = self.verts
        bins[:, 2 * self.n_verts:2 * self.n_verts + 4] = self.header
        return dict(edges=self.edges, verts=self.verts, header=self.header, rows=self.rows,
                    fp_pix=self.fp_pix.reshape(-1), bucket_first=self.bucket_first, bins=bins)
