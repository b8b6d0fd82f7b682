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


def test_published_h3_vectors():
    """Index VALUES (they decide the mesh-node numbering, encoder.py:76-84): every vector below is published by the H3 project.
      * h3-py README:  latlng_to_cell(37.3615593, -122.0553238, 5) == '85283473fffffff' and
                       cell_to_latlng('85283473fffffff') == (37.34579337536848, -121.97637597255124)
      * h3-js README:  latLngToCell(37.3615593, -122.0553238, 7) == '87283472bffffff'  (its ancestors at res 0..6 are its
                       prefixes: base cell 20, digits 0 6 4 3 4 5)
      * H3 docs:       the example cell 8928308280fffff (San Francisco) -> ancestors 822837f / 832830f / 8428309 ...
      * res-0 indexes are 0x08001fffffffffff + (base cell << 45); base cells are numbered north to south (H3 docs)
      * New York and Paris res-2 cells (cross-checked against published indexes by the round-1 review)."""
    p = (37.3615593, -122.0553238)
    assert h3.latlng_to_cell(*p, 5) == "85283473fffffff"
    lat, lng = h3.cell_to_latlng("85283473fffffff")
    assert lat == pytest.approx(37.34579337536848, abs=1e-11) and lng == pytest.approx(-121.97637597255124, abs=1e-11)
    lineage = {0: "8029fffffffffff", 1: "81283ffffffffff", 2: "822837fffffffff", 3: "832834fffffffff", 4: "8428347ffffffff",
               5: "85283473fffffff"}  # prefixes of 87283472bffffff (res 6, 86283472fffffff, also holds: its 14 M-cell table takes minutes to build)
    for res, want in lineage.items():
        assert h3.latlng_to_cell(*p, res) == want
    sf = (37.775938728915946, -122.41795063018799)  # centre of 8928308280fffff
    assert [h3.latlng_to_cell(*sf, r) for r in (2, 3, 4)] == ["822837fffffffff", "832830fffffffff", "8428309ffffffff"]
    assert h3.latlng_to_cell(40.7128, -74.0060, 2) == "822a17fffffffff"
    assert h3.latlng_to_cell(48.8566, 2.3522, 2) == "821fb7fffffffff"
    res0 = h3.get_res0_cells()
    assert res0 == [format((1 << 59) | (b << 45) | ((1 << 45) - 1), "x") for b in range(122)]
    t0 = h3.table(0)
    assert np.all(np.diff(t0.lat[np.argsort(t0.base_cell)]) <= 1e-12)  # base cell number increases as the centre moves south


def test_sorted_index_order_gives_the_reference_numbering():
    """encoder.py:76-84 / assimilator_decoder.py:69-77 number mesh nodes by DESCENDING sorted index (h_index counts down over
    the sorted list), the latent graph by ascending (encoder.py:77,247).  graphs.py must reproduce both from the index values."""
    from graph_weather_b200 import graphs

    base = sorted(h3.uncompact_cells(h3.get_res0_cells(), 2))
    assert len(base) == 5882 and base == sorted(base, key=lambda s: int(s, 16))  # string order == numeric order (15 hex digits)
    pts = [(37.3615593, -122.0553238), (40.7128, -74.0060), (48.8566, 2.3522), (89.9, 10.0), (-89.9, -170.0)]
    g = graphs.build_encoder_graph(pts, 2)
    m = graphs.build_mesh_graph(2)
    rank = {c: i for i, c in enumerate(base)}
    for i, (lat, lon) in enumerate(pts):
        cell = h3.latlng_to_cell(lat, lon, 2)
        assert int(g.mesh_local[i]) == len(base) - 1 - rank[cell]  # descending numbering of encoder / decoder
        assert int(g.edge_index[1, i]) == len(pts) + len(base) - 1 - rank[cell]
    # latent graph: node i is the i-th cell in ascending order; its first edge is the self loop of grid_disk's origin
    for i in (0, 1, 2941, 5881):
        first = int(np.nonzero(m.edge_index[0] == i)[0][0])
        assert int(m.edge_index[1, first]) == i
        nbrs = sorted(int(v) for v in m.edge_index[1][m.edge_index[0] == i])
        assert nbrs == sorted(rank[c] for c in h3.grid_disk(base[i], 1))
