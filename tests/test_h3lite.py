"""H3-compatible indexing: count KATs held by the reference's own tests + self-consistency."""
import numpy as np
import pytest

from graph_weather_b200 import h3lite as h3


def test_cell_counts():
    # tests/test_model.py:30-31,88 (5882 = h3.get_num_cells(2)); tests/test_stretched_mesh.py:19 uses 49 = 7^2 children
    assert h3.get_num_cells(2) == 5882
    for res, n in ((0, 122), (1, 842), (2, 5882), (3, 41162)):
        t = h3.table(res)
        assert t.num == n == h3.get_num_cells(res)
        assert int(t.is_pentagon.sum()) == 12
    cells = h3.uncompact_cells(h3.get_res0_cells(), 2)
    assert len(cells) == len(set(cells)) == 5882
    assert all(len(c) == 15 and c.startswith("82") for c in cells)


def test_latent_edge_count():
    # tests/test_model.py:31 -- 41162 directed latent edges (7 per hexagon incl. self loop, 6 per pentagon)
    cells = sorted(h3.uncompact_cells(h3.get_res0_cells(), 2))
    n = sum(len(h3.grid_disk(c, 1)) for c in cells)
    assert n == 41162 == 7 * 5882 - 12


def test_uk_box_kat():
    # tests/test_dynamic_graph_builder.py:11-65: 25 points -> 5 unique res-2 cells, 175 decoder edges, 19 latent edges, 5 self loops
    ll = [(float(a), float(b)) for a in range(50, 55) for b in range(-2, 3)]
    cells = [h3.latlng_to_cell(a, b, 2) for a, b in ll]
    u = sorted(set(cells))
    assert len(u) == 5
    assert sum(len(h3.grid_disk(c, 1)) for c in cells) == 175
    edges = [(c, h) for c in u for h in h3.grid_disk(c, 1) if h in u]
    assert len(edges) == 19 and sum(a == b for a, b in edges) == 5


def test_known_h3_values():
    # H3's documentation example cell 8928308280fffff (San Francisco) has the res-2 ancestor 822837fffffffff
    assert h3.latlng_to_cell(37.7752702151959, -122.418307270836, 2) == "822837fffffffff"
    # H3's pentagon base cells
    t0 = h3.table(0)
    assert sorted(int(b) for b in t0.base_cell[t0.is_pentagon]) == [4, 14, 24, 38, 49, 58, 63, 72, 83, 97, 107, 117]


def test_roundtrip_and_symmetry():
    t = h3.table(2)
    # centre of every cell maps back to the cell
    back = t.locate(t.lat, t.lng)
    assert np.array_equal(back, np.arange(t.num))
    # neighbour relation is symmetric
    for c in range(0, t.num, 37):
        for n in t.nbr[c]:
            if n >= 0:
                assert c in t.nbr[n]
    # poles and antimeridian are accepted (tests/test_dynamic_graph_builder.py:99-100)
    for lat, lon in ((90.0, 0.0), (-90.0, 0.0), (0.0, 180.0), (0.0, -180.0), (45.0, 359.75)):
        assert len(h3.latlng_to_cell(lat, lon, 2)) == 15


def test_great_circle_distance():
    assert h3.great_circle_distance((0, 0), (0, 90), unit="rads") == pytest.approx(np.pi / 2, abs=1e-15)
    assert h3.great_circle_distance((10, 20), (10, 20), unit="rads") == 0.0


def test_parent_child():
    t2 = h3.table(2)
    counts = np.bincount(t2.parent, minlength=h3.table(1).num)
    assert set(counts.tolist()) == {6, 7} and int((counts == 6).sum()) == 12
