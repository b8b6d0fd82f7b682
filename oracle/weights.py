"""TEST INFRASTRUCTURE ONLY.  Seeded, reproducible weights for parity runs.

Fixtures under tests/golden/ store only (config, seed, outputs); weights and inputs are regenerated from the seed
with numpy's PCG64 (bit-stable across numpy versions and machines), loaded into the reference model with
`load_state_dict` when the fixture is made and into the restatement / the CUDA path when it is checked.  Unlike the
reference's default init (LayerNorm weight=1, bias=0, h3_nodes=0) every parameter is non-trivial so that a kernel
which forgets a gamma/beta/bias or the h3_nodes rows cannot pass.
"""

from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


def _mlp_shapes(prefix, in_dim, out_dim, hidden, hidden_layers, norm=True):
    s = OrderedDict()
    d = in_dim
    for i in range(hidden_layers):
        s[f"{prefix}.model.{2 * i}.weight"] = (hidden, d)
        s[f"{prefix}.model.{2 * i}.bias"] = (hidden,)
        d = hidden
    j = 2 * hidden_layers
    s[f"{prefix}.model.{j}.weight"] = (out_dim, d)
    s[f"{prefix}.model.{j}.bias"] = (out_dim,)
    if norm:
        s[f"{prefix}.model.{j + 1}.weight"] = (out_dim,)
        s[f"{prefix}.model.{j + 1}.bias"] = (out_dim,)
    return s


def _gp_shapes(prefix, blocks, nd, ed, hn, he, hln, hle):
    s = OrderedDict()
    for b in range(blocks):
        s.update(_mlp_shapes(f"{prefix}.blocks.{b}.edge_model.edge_mlp", 2 * nd + ed, ed, he, hle))
        s.update(_mlp_shapes(f"{prefix}.blocks.{b}.node_model.node_mlp", nd + ed, nd, hn, hln))
    return s


def forecaster_shapes(num_h3=5882, feature_dim=78, aux_dim=24, output_dim=None, node_dim=256, edge_dim=256, num_blocks=9,
                      hidden_dim_processor_node=256, hidden_dim_processor_edge=256, hidden_layers_processor_node=2,
                      hidden_layers_processor_edge=2, hidden_dim_decoder=128, hidden_layers_decoder=2,
                      assimilator=False, observation_dim=2):  # fmt: skip
    """state_dict keys/shapes of GraphWeatherForecaster (forecast.py:129-170) or GraphWeatherAssimilator (analysis.py:96-134)."""
    out_dim = feature_dim if output_dim is None else output_dim
    in_dim = observation_dim if assimilator else feature_dim + aux_dim
    hn, he = hidden_dim_processor_node, hidden_dim_processor_edge
    hln, hle = hidden_layers_processor_node, hidden_layers_processor_edge
    s = OrderedDict()
    if not assimilator:
        s["encoder.h3_nodes"] = (num_h3, in_dim)
    s.update(_mlp_shapes("encoder.node_encoder", in_dim, node_dim, hn, hln))
    s.update(_mlp_shapes("encoder.edge_encoder", 3 if assimilator else 2, edge_dim, he, hle))
    s.update(_mlp_shapes("encoder.latent_edge_encoder", 2, edge_dim, he, hle))
    s.update(_gp_shapes("encoder.graph_processor", 1, node_dim, edge_dim, hn, he, hln, hle))
    s.update(_gp_shapes("processor.graph_processor", num_blocks, node_dim, edge_dim, hn, he, hln, hle))
    s.update(_mlp_shapes("decoder.edge_encoder", 2, edge_dim, he, 2))
    s.update(_gp_shapes("decoder.graph_processor", 1, node_dim, edge_dim, hn, he, hln, hle))
    s.update(_mlp_shapes("decoder.node_decoder", node_dim, out_dim, hidden_dim_decoder, hidden_layers_decoder, norm=False))
    return s


def make_state_dict(shapes, seed: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("h3_nodes"):
            v = rng.standard_normal(shape) * 0.5
        elif len(shape) == 2:  # nn.Linear weight [out, in]: Kaiming-uniform bound 1/sqrt(fan_in) like the default
            v = rng.uniform(-1.0, 1.0, shape) / np.sqrt(shape[1])
        elif name.split(".")[-2].isdigit() and _is_layernorm(name, shapes):
            v = 1.0 + 0.2 * rng.standard_normal(shape) if name.endswith("weight") else 0.1 * rng.standard_normal(shape)
        else:  # Linear bias
            fan_in = shapes[name[: -len("bias")] + "weight"][1]
            v = rng.uniform(-1.0, 1.0, shape) / np.sqrt(fan_in)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def _is_layernorm(name, shapes):
    return name.endswith("weight") and len(shapes[name]) == 1 or (
        name.endswith("bias") and len(shapes.get(name[: -len("bias")] + "weight", (0, 0))) == 1
    )


def make_features(batch, n, dim, seed):
    rng = np.random.Generator(np.random.PCG64(seed + 1000003))
    return torch.from_numpy(rng.standard_normal((batch, n, dim)).astype(np.float32))
