"""CPU: the grid <-> graph mapping of GraphWeatherForecaster (forecast.py:178-213) against outputs of the reference's own loops
(tests/golden/forecaster_constraints_10deg_b2.npz, made by tests/golden/make_golden.py), constructor behaviour of the constraint
layer, and the HF-hub save / load round trip the reference gets from PyTorchModelHubMixin (forecast.py:61)."""
import json
import os

import numpy as np
import pytest
import torch


def _irregular(golden_dir):
    z = np.load(os.path.join(golden_dir, "forecaster_constraints_10deg_b2.npz"))
    cfg = json.loads(str(z["config"]))
    ll = [(a, b) for a in cfg["lats"] for b in cfg["lons"]]
    return z, ll


def test_grid_mapping_reproduces_the_reference_loops(golden_dir):
    from graph_weather_b200 import GraphWeatherForecaster

    z, ll = _irregular(golden_dir)
    m = GraphWeatherForecaster(ll, constraint_type="additive", feature_dim=4, aux_dim=0, output_dim=4)
    assert m.grid_shape == (4, 6)
    assert np.array_equal(np.array(m.node_to_grid), z["node_to_grid"])  # truncation quirks included: shared and empty cells
    g = torch.from_numpy(z["map_in"])
    grid = m.graph_to_grid(g)
    assert np.array_equal(grid.numpy(), z["map_grid"])  # last node wins a shared cell, untouched cells are zero
    assert np.array_equal(m.grid_to_graph(grid).numpy(), z["map_back"])


def test_regular_grid_mapping_is_the_identity():
    from graph_weather_b200.constraint import GridMapping

    ll = [(float(a), float(b)) for a in range(-90, 90, 10) for b in range(0, 360, 10)]
    gm = GridMapping(ll)
    assert gm.grid_shape == (18, 36)
    x = torch.randn(2, len(ll), 5)
    assert torch.equal(gm.grid_to_graph(gm.graph_to_grid(x)), x)
    assert torch.equal(gm.graph_to_grid(x), x.permute(0, 2, 1).reshape(2, 5, 18, 36))  # == rearrange "b (h w) c -> b c h w"


def test_constraint_layer_constructor_and_errors():
    from graph_weather_b200 import GraphWeatherForecaster, PhysicalConstraintLayer

    ll = [(a, b) for a in np.linspace(-90, 90, 2) for b in np.linspace(-90, 90, 2)]
    m = GraphWeatherForecaster(ll, constraint_type="additive", feature_dim=2, aux_dim=0, output_dim=2)
    assert isinstance(m.constraint, PhysicalConstraintLayer) and m.constraint.upsampling_factor == 1
    assert not any(k.startswith("constraint") for k in m.state_dict())  # no parameters, no cycle through the back-reference
    plain = GraphWeatherForecaster(ll, feature_dim=2, aux_dim=0, output_dim=2)
    assert list(plain.state_dict().keys()) == list(m.state_dict().keys())
    assert not hasattr(plain, "constraint")
    with pytest.raises(RuntimeError):  # CUDA only: no CPU fallback
        m.constraint(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))
    with pytest.raises(NotImplementedError):
        PhysicalConstraintLayer(m, (2, 2), 2, "additive")


def test_hub_save_and_load_round_trip(tmp_path):
    """save_pretrained / from_pretrained (PyTorchModelHubMixin, forecast.py:61; analysis.py:52): config.json carries the
    constructor arguments, model.safetensors the reference-named parameters."""
    from graph_weather_b200 import GraphWeatherAssimilator, GraphWeatherForecaster

    ll = [(float(a), float(b)) for a in range(-90, 90, 30) for b in range(0, 360, 30)]
    torch.manual_seed(3)
    m = GraphWeatherForecaster(ll, num_blocks=2, constraint_type="additive", feature_dim=6, aux_dim=2)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.01)  # h3_nodes / LayerNorm parameters off their defaults
    m.save_pretrained(tmp_path / "fc")
    cfg = json.load(open(tmp_path / "fc" / "config.json"))
    assert cfg["num_blocks"] == 2 and cfg["constraint_type"] == "additive" and cfg["feature_dim"] == 6
    m2 = GraphWeatherForecaster.from_pretrained(tmp_path / "fc")
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd1.keys()) == list(sd2.keys()) and all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    assert m2.constraint_type == "additive" and m2.grid_shape == m.grid_shape
    a = GraphWeatherAssimilator(output_lat_lons=ll, analysis_dim=5, num_blocks=1)
    a.save_pretrained(tmp_path / "as")
    a2 = GraphWeatherAssimilator.from_pretrained(tmp_path / "as")
    assert all(torch.equal(v, a2.state_dict()[k]) for k, v in a.state_dict().items())
