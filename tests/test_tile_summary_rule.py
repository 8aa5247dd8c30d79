"""CPU: the rule by which the device generators keep the tile summary (one bit per 32 x 16 tile of the occupancy
plane) in step while drawing or erasing a wall -- csrc/bcg_generate.cuh draw_wall: only the pixel with which the line
ENTERS a tile touches the summary (the line's first pixel, a pixel in column 0 of a tile, in row 0 of a tile when the
line goes up in y / row 15 when it goes down, or in the map's last row).  Checked here on the oracle's restatement of
cv2.line (oracle/aisle_oracle.py line_pixels, pinned against cv2.line in tests/test_oracle_aisle.py): for random
segments, clipped to random maps the way draw_wall skips out-of-map pixels, the tiles flagged by the rule are exactly
the tiles the line has pixels in."""
import numpy as np

from oracle.aisle_oracle import line_pixels


def _tiles_by_rule(px, rows, pitch, sy):
    j = np.arange(len(px))
    x, y = px[:, 0], px[:, 1]
    inside = (x >= 0) & (y >= 0) & (x < pitch) & (y < rows)
    enters = (j == 0) | ((x & 31) == 0) | ((y & 15) == (0 if sy > 0 else 15)) | (y == rows - 1)
    sel = inside & enters
    return set(zip((x[sel] >> 5).tolist(), (y[sel] >> 4).tolist())), set(zip((x[inside] >> 5).tolist(), (y[inside] >> 4).tolist()))


def test_tile_entry_pixels_cover_every_tile_a_wall_touches():
    rng = np.random.RandomState(12)
    checked = 0
    for trial in range(4000):
        rows, pitch = int(rng.randint(20, 700)), 32 * int(rng.randint(1, 25))
        lo, hi = (-60, 60) if trial % 3 == 0 else (0, 0)          # a third of the segments start or end outside the map
        x0, x1 = rng.randint(lo, pitch + hi, size=2)
        y0, y1 = rng.randint(lo, rows + hi, size=2)
        if trial % 10 == 0:
            y1 = y0                                               # horizontal
        if trial % 10 == 1:
            x1 = x0                                               # vertical
        px = line_pixels(int(x0), int(y0), int(x1), int(y1))
        if x1 < x0:
            y0, y1 = y1, y0
        sy = 1 if y1 - y0 >= 0 else -1
        flagged, touched = _tiles_by_rule(px, rows, pitch, sy)
        assert flagged == touched, (trial, (x0, y0, x1, y1), rows, pitch)
        checked += len(touched)
    assert checked > 20000
