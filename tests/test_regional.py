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


# ---- the reference's own test-suite for this model (/root/reference/tests/test_regional_forecast.py:11-199), same small config ----
def _small_config(**kw):
    from graph_weather_b200.regional import RegionalForecasterConfig

    return RegionalForecasterConfig(feature_dim=12, aux_dim=4, node_dim=32, edge_dim=32, num_blocks=2, hidden_dim_processor_node=32,
                                    hidden_dim_processor_edge=32, hidden_dim_decoder=32, **kw)  # fmt: skip


_UK = [(51.5, -0.1), (52.0, 0.5), (53.0, -1.0), (54.0, -2.0), (50.0, -3.0)]
_DE = [(52.5, 13.4), (48.1, 11.6), (50.9, 6.9)]


def test_config_build_and_relaxation_weights():
    """:35-41 and :178-184."""
    from graph_weather_b200.regional import BoundaryNudgingLayer

    model = _small_config().build()
    assert hasattr(model, "forward") and hasattr(model, "graph_builder") and hasattr(model, "h3_embeddings")
    assert model.nudging is None
    w = BoundaryNudgingLayer._compute_relaxation_weights(_UK, torch.device("cpu"))
    assert w.shape == (5, 1) and w.min() >= 0.0 and w.max() <= 1.0 and torch.isclose(w.max(), torch.tensor(1.0))


@pytest.mark.gpu
def test_reference_suite_small_config():
    """:44-175: shapes, NaN-free, another region / another length on the same model, output_dim override, the residual with zero
    weights, nudging off / without context / with context.  (Hidden size 32: the exact-fp32 CUDA-core path.)"""
    torch.manual_seed(0)
    model = _small_config().build().cuda().eval()
    out = model(torch.randn(2, 5, 16, device="cuda"), _UK)
    assert out.shape == (2, 5, 12) and not torch.isnan(out).any()
    assert model(torch.randn(1, 3, 16, device="cuda"), _DE).shape == (1, 3, 12)
    assert model(torch.randn(1, 5, 16, device="cuda"), _UK).shape == (1, 5, 12)
    m6 = _small_config(output_dim=6).build().cuda().eval()
    assert m6(torch.randn(1, 5, 16, device="cuda"), _UK).shape == (1, 5, 6)
    x = torch.randn(1, 5, 16, device="cuda")
    with torch.no_grad():
        for q in model.parameters():
            q.zero_()
    assert torch.allclose(model(x, _UK), x[..., :12], atol=1e-5)  # :113-125
    ctx = torch.randn(1, 5, 12, device="cuda") * 10.0
    assert torch.allclose(model(x, _UK, global_context=ctx), model(x, _UK))  # nudging disabled: the context is ignored (:144-154)
    mn = _small_config(enable_nudging=True, nudging_hidden_dim=16).build().cuda().eval()
    o0 = mn(x, _UK, global_context=None)
    assert o0.shape == (1, 5, 12) and not torch.isnan(o0).any()
    assert not torch.allclose(o0, mn(x, _UK, global_context=ctx))  # :167-175
    mt = _small_config().build().cuda()  # train mode + grad: the backward of this model is not built -- it says so (:87-99 is out of scope)
    with torch.enable_grad(), pytest.raises(NotImplementedError):
        mt(torch.randn(1, 5, 16, device="cuda"), _UK)
