"""CPU: the angle-binned footprint table reproduces get_pixel_footprint (reference
utilities/path_tools.py:122-162; its own consistency test is test_path_tools.py:453-462) mask for mask."""
import numpy as np
import pytest

from bc_gym_planning_env_b200.footprint_lut import FootprintLut, rounded_vertices
from oracle import plan_env_oracle as O

RECT = np.array([[-0.77, -0.385], [-0.77, 0.385], [0.67, 0.385], [0.67, -0.385]])


@pytest.mark.parametrize("name,footprint", [("rect", RECT), ("tricycle", O.TRICYCLE_FOOTPRINT), ("diffdrive", O.DIFFDRIVE_FOOTPRINT)])
@pytest.mark.parametrize("resolution", [0.03, 0.05])
def test_table_masks_equal_reference_masks(name, footprint, resolution):
    lut = FootprintLut(footprint, resolution)
    rng = np.random.RandomState(1)
    angles = np.concatenate([rng.uniform(-np.pi, np.pi, 400), [0., np.pi / 2, -np.pi / 2, -np.pi, np.pi - 1e-12],
                             rng.uniform(-20, 20, 20)])
    for a in angles:
        k = lut.bin_of(a)
        mask, x0, y0 = lut.mask(k)
        ref = O.pixel_footprint(a, footprint, resolution)
        hy, hx = ref.shape[0] // 2, ref.shape[1] // 2
        ys, xs = np.nonzero(ref)
        want = np.zeros_like(mask)
        want[ys - hy - y0, xs - hx - x0] = True
        assert np.array_equal(mask, want), a
        assert lut.pixels[k] == int((ref != 0).sum())


def test_table_shape_facts():
    lut = FootprintLut(O.TRICYCLE_FOOTPRINT, 0.03)
    assert 3900 < lut.n_bins < 4000 and lut.wpr == 1 and lut.max_rows <= 64      # 3960 bins
    assert lut.edges[0] == -np.pi and lut.edges[-1] == np.pi and np.all(np.diff(lut.edges) > 0)
    assert 1480 <= lut.pixels.min() and lut.pixels.max() <= 1600                 # 1493-1580 px
    # every bucket's first bin really contains the bucket's left end
    lefts = -np.pi + np.arange(lut.n_buckets) / lut.bucket_scale
    k = lut.bucket_first
    assert np.all(lut.edges[k] <= lefts + 1e-15) and np.all(lefts < lut.edges[k + 1])


def test_bin_edges_sit_where_a_rounded_vertex_flips():
    lut = FootprintLut(O.DIFFDRIVE_FOOTPRINT, 0.05)
    for k in range(1, lut.n_bins, 37):
        below = rounded_vertices(lut.edges[k] - 1e-9, lut.fp_pix)
        above = rounded_vertices(lut.edges[k] + 1e-9, lut.fp_pix)
        assert not np.array_equal(below, above)


def test_canvas_equals_the_reference_return_format():
    """FootprintLut.canvas == get_pixel_footprint's array (the native-vs-Python equality of reference test_path_tools.py:453-462)."""
    rng = np.random.RandomState(4)
    for footprint, res in ((O.TRICYCLE_FOOTPRINT, 0.03), (RECT, 0.05), (O.DIFFDRIVE_FOOTPRINT, 0.03)):
        lut = FootprintLut(footprint, res)
        for a in np.concatenate([rng.uniform(-np.pi, np.pi, 120), [0., np.pi / 2, -np.pi, 7.3]]):
            got, want = lut.canvas(a), O.pixel_footprint(a, footprint, res)
            assert got.dtype == np.uint8 and got.shape == want.shape and np.array_equal(got, want), a
