"""DynamicGraphBuilder against the reference's own tests (/root/reference/tests/test_dynamic_graph_builder.py:11-107, same
assertions) and against the loop-for-loop construction with the h3 API restatement."""
import numpy as np
import pytest
import torch

from graph_weather_b200 import h3lite as h3
from graph_weather_b200.dynamic_graph_builder import DynamicGraphBuilder
from graph_weather_b200.graphs import validate_lat_lons


def _small_region():
    return [(float(lat), float(lon)) for lat in range(50, 55) for lon in range(-2, 3)]


def test_encoder_graph():
    builder = DynamicGraphBuilder(resolution=2)
    graph, h3_indices = builder.build_encoder_graph(_small_region())
    assert graph.edge_index.shape == (2, 25) and graph.edge_index.dtype == torch.long
    assert graph.edge_index[0].tolist() == list(range(25))
    assert graph.edge_index[1].min().item() >= 25
    assert graph.edge_attr.shape == (25, 2) and graph.edge_attr.abs().max() <= 1.0
    assert all(0 <= idx < h3.get_num_cells(2) for idx in h3_indices)


def test_decoder_and_latent_graph_counts():
    builder = DynamicGraphBuilder(resolution=2)
    lat_lons = _small_region()
    graph = builder.build_decoder_graph(lat_lons)
    assert graph.edge_index.shape == (2, 175) and graph.edge_index.dtype == torch.long
    assert graph.edge_attr.shape == (175, 2) and graph.edge_attr.abs().max() <= 1.0
    unique_cells = sorted(set(h3.latlng_to_cell(lat, lon, 2) for lat, lon in lat_lons))
    assert len(unique_cells) == 5
    latent = builder.build_latent_graph(unique_cells)
    assert latent.edge_index.shape == (2, 19) and latent.edge_attr.shape == (19, 2)
    assert int((latent.edge_index[0] == latent.edge_index[1]).sum()) == 5


def test_builder_caching_and_validation():
    builder = DynamicGraphBuilder(resolution=2)
    lat_lons = _small_region()
    res1, res2 = builder(lat_lons), builder(lat_lons)
    assert all(a is b for a, b in zip(res1, res2))
    assert res1[0] is not builder([(0.0, 0.0), (1.0, 1.0)])[0]
    with pytest.raises(ValueError, match="must not be empty"):
        builder([])
    with pytest.raises(ValueError, match="latitude"):
        builder([(91.0, 0.0)])
    with pytest.raises(ValueError, match="latitude"):
        builder([(-91.0, 0.0)])
    assert builder([(-90.0, 0.0), (90.0, 180.0)])[0].edge_index.shape == (2, 2)
    with pytest.raises(ValueError, match="must not be empty"):
        validate_lat_lons([])
    validate_lat_lons([(0.0, 0.0), (45.0, 90.0)])


def test_against_the_reference_loops():
    """dynamic_graph_builder.py:40-128 restated with the h3 API (dict / sorted / loops), on scattered points incl. a pentagon's
    neighbourhood: same node numbering, same edge sets, same attributes."""
    rng = np.random.Generator(np.random.PCG64(4))
    lat_lons = [(float(a), float(b)) for a, b in zip(rng.uniform(-90, 90, 120), rng.uniform(-180, 180, 120))] + [(50.0 + 0.1 * i, 0.0) for i in range(20)]
    b = DynamicGraphBuilder(2)
    cells = [h3.latlng_to_cell(lat, lon, 2) for lat, lon in lat_lons]
    unique = sorted(set(cells))
    idx = {c: i for i, c in enumerate(unique)}
    enc, h3_indices = b.build_encoder_graph(lat_lons)
    assert enc.edge_index[1].tolist() == [len(lat_lons) + idx[c] for c in cells]
    assert h3_indices == [b.global_h3_map[c] for c in unique]
    d = [h3.great_circle_distance(p, h3.cell_to_latlng(c), unit="rads") for p, c in zip(lat_lons, cells)]
    assert np.abs(enc.edge_attr.numpy() - np.array([[np.sin(x), np.cos(x)] for x in d], dtype=np.float32)).max() <= 2e-7
    hood = sorted(set(h for c in unique for h in h3.grid_disk(c, 1)))
    hidx = {c: i for i, c in enumerate(hood)}
    dec = b.build_decoder_graph(lat_lons)
    want = sorted((hidx[h], len(hood) + i) for i, c in enumerate(cells) for h in h3.grid_disk(c, 1))
    assert sorted(zip(dec.edge_index[0].tolist(), dec.edge_index[1].tolist())) == want
    lat = b.build_latent_graph(unique)
    want = sorted((idx[c], idx[h]) for c in unique for h in h3.grid_disk(c, 1) if h in idx)
    assert sorted(zip(lat.edge_index[0].tolist(), lat.edge_index[1].tolist())) == want
