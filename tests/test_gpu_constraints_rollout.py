"""-m gpu: PhysicalConstraintLayer on device and the zero-copy rollout, against outputs of the reference's own code
(tests/golden/forecaster_constraints_10deg_b2.npz) and the reference's conservation tests (tests/test_model.py:374-465)."""
import json
import os

import numpy as np
import pytest
import torch

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _built():
    ge.build()


def _grid(step):
    return [(float(lat), float(lon)) for lat in range(-90, 90, step) for lon in range(0, 360, step)]


@pytest.mark.parametrize("ctype", ["additive", "softmax", "multiplicative"])
def test_constrained_forecaster_matches_reference_fixture(golden_dir, ctype):
    from graph_weather_b200 import GraphWeatherForecaster
    from oracle import weights

    z = np.load(os.path.join(golden_dir, "forecaster_constraints_10deg_b2.npz"))
    cfg = json.loads(str(z["config"]))
    ll = _grid(cfg["step"])
    sd = weights.make_state_dict(weights.forecaster_shapes(), cfg["seed"])
    x = weights.make_features(cfg["batch"], len(ll), 102, cfg["seed"]).cuda()
    model = GraphWeatherForecaster(ll, constraint_type=ctype).cuda().eval()
    model.load_state_dict(sd)
    out = model(x)
    ref = torch.from_numpy(z[ctype])
    if ctype == "multiplicative":
        # y = hr * mean(lr) / (mean(hr) + 1e-8): with random weights some channel means are ~1e-5, so the reference's own
        # formula amplifies 1e-6 differences of hr by 1e4.  The layer itself is pinned at 1e-5 below; here: consistency with
        # the unconstrained forward through the same layer, and the fixture wherever the ratio is well conditioned.
        plain = GraphWeatherForecaster(ll).cuda().eval()
        plain.load_state_dict(sd)
        hr = plain(x)
        assert torch.equal(out, model.constraint.apply_rows(hr, x, model._grid_mapping.tensors(x.device)[0].to(torch.int32), 78))
        mean_hr = hr.mean(dim=1, keepdim=True).abs().cpu()
        good = (mean_hr > 1e-2).expand_as(ref)
        assert good.any()
        rel = ((out.cpu() - ref).abs() / ref.abs().clamp(min=1.0))[good]
        assert float(rel.max()) < 1e-3
    else:
        err = float((out.cpu() - ref).abs().max())
        print(f"{ctype}: max|gpu - reference| = {err:.3e}")
        assert err < TOL


@pytest.mark.parametrize("ctype", ["additive", "multiplicative", "softmax"])
def test_constraint_layer_on_an_irregular_grid(golden_dir, ctype):
    """The layer alone, graph (3D) and grid (4D) inputs, on a grid whose mapping has shared and empty cells."""
    from graph_weather_b200 import GraphWeatherForecaster

    z = np.load(os.path.join(golden_dir, "forecaster_constraints_10deg_b2.npz"))
    cfg = json.loads(str(z["config"]))
    ll = [(a, b) for a in cfg["lats"] for b in cfg["lons"]]
    m = GraphWeatherForecaster(ll, constraint_type=ctype, feature_dim=4, aux_dim=0, output_dim=4).cuda()
    hr, lr = torch.from_numpy(z["hr"]).cuda(), torch.from_numpy(z["lr"]).cuda()
    got_graph = m.constraint(hr, lr).cpu().numpy()
    got_grid = m.constraint(m.graph_to_grid(hr), m.graph_to_grid(lr)).cpu().numpy()
    for got, key in ((got_graph, ctype + "_graph"), (got_grid, ctype + "_grid")):
        ref = z[key]
        assert got.shape == ref.shape
        assert np.all(np.abs(got - ref) <= 1e-5 * np.maximum(1.0, np.abs(ref))), key


@pytest.mark.parametrize("ctype", ["additive", "multiplicative", "softmax"])
def test_conservation_like_the_reference_tests(ctype):
    """tests/test_model.py:374-465: on a 2 x 2 grid the grid mean of the output equals the grid mean of the input."""
    from graph_weather_b200 import GraphWeatherForecaster

    lats = np.linspace(-90, 90, 2)
    lons = np.linspace(-90, 90, 2)
    lat_lons = [(lat, lon) for lat in lats for lon in lons]
    torch.manual_seed(0)
    model = GraphWeatherForecaster(lat_lons, constraint_type=ctype, feature_dim=2, aux_dim=0, output_dim=2).cuda()
    inp = torch.randn(1, len(lat_lons), 2)
    output = model(inp.cuda()).cpu()
    lr_input_avg = model.graph_to_grid(inp[..., :2]).mean(dim=(-2, -1))
    lr_output_avg = model.graph_to_grid(output).mean(dim=(-2, -1))
    assert torch.allclose(lr_input_avg, lr_output_avg, atol=0.0001), f"Conservation failed: {lr_input_avg} vs {lr_output_avg}"


def test_rollout_equals_the_manual_loop(golden_dir):
    """model.rollout(features, steps) == the loop a user of the reference writes: feed the forecast back as the first 78
    features, keep (or replace) the auxiliary columns.  Bit for bit: the same kernels run, only the output row stride differs."""
    from graph_weather_b200 import GraphWeatherForecaster
    from oracle import weights

    ll = _grid(10)
    sd = weights.make_state_dict(weights.forecaster_shapes(), 16)
    x = weights.make_features(2, len(ll), 102, 16).cuda()
    model = GraphWeatherForecaster(ll).cuda().eval()
    model.load_state_dict(sd)
    steps = 3
    aux = torch.randn(steps, 2, len(ll), 24, generator=torch.Generator().manual_seed(1)).cuda()
    for a in (None, aux):
        got = model.rollout(x, steps, aux=a)
        cur = x.clone()
        for t in range(steps):
            if a is not None:
                cur[..., 78:] = a[t]
            y = model(cur)
            assert torch.equal(got[t], y), (t, a is None)
            cur = torch.cat([y, cur[..., 78:]], dim=-1)
        assert torch.equal(model.rollout(x, steps, aux=a, return_all=False), got[-1])
    assert torch.equal(x, weights.make_features(2, len(ll), 102, 16).cuda())  # the caller's tensor is not written
    constrained = GraphWeatherForecaster(ll, constraint_type="additive").cuda().eval()
    constrained.load_state_dict(sd)
    got = constrained.rollout(x, 2)
    y0 = constrained(x)
    y1 = constrained(torch.cat([y0, x[..., 78:]], dim=-1))
    assert torch.equal(got[0], y0) and torch.equal(got[1], y1)
