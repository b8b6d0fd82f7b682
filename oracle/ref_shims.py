"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (graph_weather_b200/).

Imports the UNMODIFIED reference hot-path sources from /root/reference by installing `sys.modules` shims for the
third-party packages that are absent from this image:

  torch_geometric.nn.MetaLayer   semantics restated from PyG (call site graph_net_block.py:221-228,299):
                                 row,col = edge_index; e = edge_model(x[row], x[col], e, u, None);
                                 x = node_model(x, edge_index, e, u, None); return x, e, u
  torch_geometric.data.Data      attribute bag with .to()   (encoder.py:107,268; assimilator_decoder.py:106)
  torch_scatter.scatter_sum      zeros(dim_size,F).scatter_add_(0, index, src)   (graph_net_block.py:188)
  h3                             graph_weather_b200.h3lite (H3-compatible restatement; see its docstring)

and stub parent packages so that `graph_weather/__init__.py` and `graph_weather/models/__init__.py` (which pull
xarray / natten / fengwu ...) are not executed.  /root/reference exists only in the build container; on the GPU
box `available()` is False and the committed fixtures under tests/golden/ (made with this module by
tests/golden/make_golden.py) carry the reference's outputs instead.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GW_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "graph_weather", "models", "layers"))


def _install_third_party_shims():
    import torch

    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_data = types.ModuleType("torch_geometric.data")

        class MetaLayer(torch.nn.Module):
            def __init__(self, edge_model=None, node_model=None, global_model=None):
                super().__init__()
                self.edge_model = edge_model
                self.node_model = node_model
                self.global_model = global_model

            def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
                row, col = edge_index[0], edge_index[1]
                if self.edge_model is not None:
                    edge_attr = self.edge_model(x[row], x[col], edge_attr, u, batch if batch is None else batch[row])
                if self.node_model is not None:
                    x = self.node_model(x, edge_index, edge_attr, u, batch)
                if self.global_model is not None:
                    u = self.global_model(x, edge_index, edge_attr, u, batch)
                return x, edge_attr, u

        class Data:
            def __init__(self, **kw):
                self.__dict__.update(kw)

            def to(self, device):
                for k, v in list(self.__dict__.items()):
                    if torch.is_tensor(v):
                        setattr(self, k, v.to(device))
                return self

        tg_nn.MetaLayer = MetaLayer
        tg_data.Data = Data
        tg.nn, tg.data = tg_nn, tg_data
        sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.data": tg_data})

    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")

        def scatter_sum(src, index, dim=0, out=None, dim_size=None):
            assert dim == 0 and out is None
            if dim_size is None:
                dim_size = int(index.max()) + 1
            res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
            return res.scatter_add_(0, index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), src)

        ts.scatter_sum = scatter_sum
        sys.modules["torch_scatter"] = ts

    if "h3" not in sys.modules:
        repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if repo not in sys.path:
            sys.path.insert(0, repo)
        from graph_weather_b200 import h3lite

        sys.modules["h3"] = h3lite


def load_reference_losses():
    """graph_weather/models/losses.py imported unmodified; its top-level `import torch_harmonics` (used only by the spectral
    AMSE loss, not by NormalizedMSELoss) is satisfied by a stub module."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT} (only present in the build container)")
    import importlib.util

    if "torch_harmonics" not in sys.modules:
        th = types.ModuleType("torch_harmonics")
        th.RealSHT = type("RealSHT", (), {})  # only named in a type annotation of the AMSE loss (losses.py:132)
        sys.modules["torch_harmonics"] = th
    spec = importlib.util.spec_from_file_location("_gw_reference_losses", os.path.join(REFERENCE_ROOT, "graph_weather", "models", "losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_loaded = None


def load_reference():
    """Returns a namespace with the reference classes imported from their own files, unmodified."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT} (only present in the build container)")
    _install_third_party_shims()
    base = os.path.join(REFERENCE_ROOT, "graph_weather")
    for name, path in (
        ("graph_weather", base),
        ("graph_weather.models", os.path.join(base, "models")),
        ("graph_weather.models.layers", os.path.join(base, "models", "layers")),
    ):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    gnb = importlib.import_module("graph_weather.models.layers.graph_net_block")
    # thermalizer/constraint layers are imported at module top by processor.py / forecast.py; they are plain torch
    proc = importlib.import_module("graph_weather.models.layers.processor")
    enc = importlib.import_module("graph_weather.models.layers.encoder")
    adec = importlib.import_module("graph_weather.models.layers.assimilator_decoder")
    dec = importlib.import_module("graph_weather.models.layers.decoder")
    aenc = importlib.import_module("graph_weather.models.layers.assimilator_encoder")
    models = sys.modules["graph_weather.models"]
    models.Encoder, models.Processor, models.Decoder = enc.Encoder, proc.Processor, dec.Decoder
    models.AssimilatorEncoder, models.AssimilatorDecoder = aenc.AssimilatorEncoder, adec.AssimilatorDecoder
    fc = importlib.import_module("graph_weather.models.forecast")
    an = importlib.import_module("graph_weather.models.analysis")
    if "graph_weather.models.graphcast" not in sys.modules:
        gcp = types.ModuleType("graph_weather.models.graphcast")
        gcp.__path__ = [os.path.join(base, "models", "graphcast")]
        sys.modules["graph_weather.models.graphcast"] = gcp
    gc = importlib.import_module("graph_weather.models.graphcast.model")
    dgb = importlib.import_module("graph_weather.models.layers.dynamic_graph_builder")  # needs graph_weather.utils (plain python)
    reg = importlib.import_module("graph_weather.models.regional_forecast")
    ns = types.SimpleNamespace(
        DynamicGraphBuilder=dgb.DynamicGraphBuilder,
        RegionalForecaster=reg.RegionalForecaster,
        RegionalForecasterConfig=reg.RegionalForecasterConfig,
        BoundaryNudgingLayer=reg.BoundaryNudgingLayer,
        MLP=gnb.MLP,
        GraphProcessor=gnb.GraphProcessor,
        Encoder=enc.Encoder,
        Processor=proc.Processor,
        Decoder=dec.Decoder,
        AssimilatorEncoder=aenc.AssimilatorEncoder,
        AssimilatorDecoder=adec.AssimilatorDecoder,
        GraphWeatherForecaster=fc.GraphWeatherForecaster,
        GraphWeatherAssimilator=an.GraphWeatherAssimilator,
        GraphCast=gc.GraphCast,
        GraphCastConfig=gc.GraphCastConfig,
    )
    _loaded = ns
    return ns
