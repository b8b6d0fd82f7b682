"""RegionalForecaster (graph_weather/models/regional_forecast.py): parameter contract, graphs, oracle and the CUDA path against a
fixture produced by the reference's own, unmodified file (tests/golden/make_golden.py::run_regional)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate, weights

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-4


@pytest.fixture(scope="module")
def fx():
    z = np.load(os.path.join(HERE, "golden", "regional_europe_b2.npz"))
    cfg = json.loads(str(z["config"]))
    lat_lons = [(float(a), float(b)) for a, b in z["lat_lons"]]
    shapes = {k: tuple(s) for k, s in zip(cfg["keys"], cfg["shapes"])}
    return dict(z=z, cfg=cfg, lat_lons=lat_lons, shapes=shapes, sd=weights.make_state_dict(shapes, cfg["seed"]),
                x=weights.make_features(cfg["batch"], len(lat_lons), 102, cfg["seed"]),
                gc=weights.make_features(cfg["batch"], len(lat_lons), 78, cfg["seed"] + 1))  # fmt: skip


def _model(**kw):
    from graph_weather_b200.regional import RegionalForecasterConfig

    return RegionalForecasterConfig(enable_nudging=True, **kw).build()


def test_state_dict_contract_matches_the_reference(fx):
    m = _model()
    sd = m.state_dict()
    assert list(sd.keys()) == fx["cfg"]["keys"]
    assert [list(v.shape) for v in sd.values()] == fx["cfg"]["shapes"]
    m.load_state_dict(fx["sd"])  # strict


def test_region_graphs_match_the_reference(fx):
    from graph_weather_b200.regional import _RegionGraphs

    m = _model()
    g = _RegionGraphs(m.graph_builder, fx["lat_lons"])
    z = fx["z"]
    assert g.h3_indices.tolist() == z["h3_indices"].tolist()
    assert g.n_mesh == 80 and g.n_lat_edges == z["lat_edge_index"].shape[1] == 478
    n = len(fx["lat_lons"])
    assert np.array_equal(g.mesh_local, z["enc_edge_index"][1] - n)
    # target-sorted latent edges are a permutation of the reference's edge list, attributes travelling with their edges
    ref = {(int(s), int(d)): a for s, d, a in zip(z["lat_edge_index"][0], z["lat_edge_index"][1], z["lat_edge_attr"])}
    assert len(ref) == 478 and np.all(np.diff(g.lat_dst) >= 0)
    for s, d, a in zip(g.lat_src, g.lat_dst, g.lat_attr):
        assert np.allclose(ref[(int(s), int(d))], a, atol=1e-6)
    assert g.lat_ptr[-1] == 478 and g.dec_ptr.tolist() == list(range(n + 1))


def test_oracle_matches_the_reference_fixture(fx):
    g = restate.regional_graphs(fx["lat_lons"])
    z = fx["z"]
    assert g["h3_indices"] == z["h3_indices"].tolist()
    assert np.array_equal(g["enc_edge_index"].numpy(), z["enc_edge_index"])
    out = restate.regional_forward(fx["sd"], g, fx["x"])
    assert float((out - torch.from_numpy(z["out"])).abs().max()) < 1e-5
    out_n = restate.regional_forward(fx["sd"], g, fx["x"], global_context=fx["gc"], lat_lons=fx["lat_lons"])
    assert float((out_n - torch.from_numpy(z["out_nudged"])).abs().max()) < 1e-5


def test_boundary_nudging_arithmetic(fx):
    """BoundaryNudgingLayer is plain tensor arithmetic outside the GNN: given the reference's un-nudged output it must reproduce
    the reference's nudged output on any device."""
    m = _model()
    m.load_state_dict(fx["sd"])
    z = fx["z"]
    got = m.nudging(torch.from_numpy(z["out"]), fx["gc"], fx["lat_lons"])
    assert float((got - torch.from_numpy(z["out_nudged"])).abs().max()) < 1e-5


def test_no_host_path():
    m = _model()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 2, 102), [(0.0, 0.0), (1.0, 1.0)])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["auto", "fp32_simt"])
def test_regional_forecaster_matches_reference_fixture(fx, precision):
    m = _model(precision=precision).cuda().eval()
    m.load_state_dict(fx["sd"])
    z = fx["z"]
    x, gc = fx["x"].cuda(), fx["gc"].cuda()
    out = m(x, fx["lat_lons"])
    assert out.shape == (2, len(fx["lat_lons"]), 78)
    assert float((out.cpu() - torch.from_numpy(z["out"])).abs().max()) < TOL
    out_n = m(x, fx["lat_lons"], global_context=gc)
    assert float((out_n.cpu() - torch.from_numpy(z["out_nudged"])).abs().max()) < TOL
    # a second region of another size gets its own plan; the first one still answers from its cached graphs
    small = fx["lat_lons"][:300]
    o2 = m(x[:, :300].contiguous(), small)
    g2 = restate.regional_graphs(small)
    ref2 = restate.regional_forward(fx["sd"], g2, fx["x"][:, :300])
    assert float((o2.cpu() - ref2).abs().max()) < TOL
    assert torch.equal(m(x, fx["lat_lons"]), out)
