"""DynamicGraphBuilder (graph_weather/models/layers/dynamic_graph_builder.py:13-155): encoder / decoder / latent graphs for an
arbitrary (regional) set of coordinates, numbered LOCALLY over the cells the coordinates touch.

Same class, methods, return values and caching rule as the reference (a repeated call with the *same list object* returns the
cached graphs, :131-139).  The reference walks the coordinates in Python and calls h3 per point / per neighbour; here every
graph is a handful of vectorised numpy operations on the cell tables of graph_weather_b200.h3lite.  Edge order, node numbering
(sorted unique cells; neighbourhood cells for the decoder) and edge attributes follow the reference loop for loop, except that a
cell's neighbours come in increasing-index order where h3.grid_disk returns ring order (edge order only, not the graph)."""

from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import graphs, h3lite


class Data:
    """The attribute bag the reference gets from torch_geometric.data.Data (edge_index, edge_attr, .to())."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class DynamicGraphBuilder:
    def __init__(self, resolution: int = 2):
        self.resolution = resolution
        t = h3lite.table(resolution)
        self._t = t
        self.all_h3 = [format(int(v), "x") for v in t.index[t.order]]  # sorted(h3.uncompact_cells(res0, resolution)), :22
        self.global_h3_map = {cell: i for i, cell in enumerate(self.all_h3)}
        self._disk = graphs._disk_table(t)  # [cells, 7]: self, then neighbours by index; -1 pads pentagons
        self._prev_lat_lons: Optional[List[Tuple[float, float]]] = None
        self._cached_encoder_graph: Optional[Data] = None
        self._cached_decoder_graph: Optional[Data] = None
        self._cached_latent_graph: Optional[Data] = None
        self._cached_h3_indices: Optional[List[int]] = None

    # cell id (table order) of every coordinate, and the sorted unique cells as global ranks
    def _cells(self, lat_lons):
        ll = np.asarray([(float(a), float(b)) for a, b in lat_lons], dtype=np.float64).reshape(-1, 2)
        cell = self._t.locate(np.radians(ll[:, 0]), np.radians(ll[:, 1]))
        rank = self._t.rank[cell]  # position in the sorted index list == the order of Python's sorted() on the hex strings
        uniq = np.unique(rank)
        return ll, cell, rank, uniq

    def _assign_h3_cells(self, lat_lons):
        """(:31-38) cells per coordinate, sorted unique cells, cell -> local index."""
        _, _, rank, uniq = self._cells(lat_lons)
        h3_cells = [self.all_h3[r] for r in rank]
        unique_cells = [self.all_h3[r] for r in uniq]
        return h3_cells, unique_cells, {c: i for i, c in enumerate(unique_cells)}

    def build_encoder_graph(self, lat_lons) -> Tuple[Data, List[int]]:
        """(:40-66) one edge per coordinate -> its cell (local index offset by the number of coordinates)."""
        ll, cell, rank, uniq = self._cells(lat_lons)
        n = ll.shape[0]
        local = np.searchsorted(uniq, rank)
        attr = graphs._sincos_attr(ll[:, 0], ll[:, 1], np.degrees(self._t.lat[cell]), np.degrees(self._t.lng[cell]))
        ei = np.stack([np.arange(n, dtype=np.int64), n + local.astype(np.int64)])
        return Data(edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(attr)), [int(r) for r in uniq]

    def build_decoder_graph(self, lat_lons) -> Data:
        """(:68-98) every cell of grid_disk(cell(p), 1) -> coordinate p; sources are numbered over the sorted neighbourhood cells."""
        ll, cell, rank, uniq = self._cells(lat_lons)
        n = ll.shape[0]
        disk = self._disk[cell]  # [n, 7]
        valid = disk >= 0
        p_of_edge = np.broadcast_to(np.arange(n)[:, None], disk.shape)[valid]
        h_cell = disk[valid]
        hood = np.unique(self._t.rank[self._disk[self._t.order[uniq]][self._disk[self._t.order[uniq]] >= 0]])  # sorted neighbourhood
        src = np.searchsorted(hood, self._t.rank[h_cell])
        attr = graphs._sincos_attr(ll[p_of_edge, 0], ll[p_of_edge, 1], np.degrees(self._t.lat[h_cell]), np.degrees(self._t.lng[h_cell]))
        ei = np.stack([src.astype(np.int64), hood.size + p_of_edge.astype(np.int64)])
        return Data(edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(attr))

    def build_latent_graph(self, unique_cells: List[str]) -> Data:
        """(:100-128) neighbour edges among the given cells only (self loops included)."""
        ranks = np.array([self.global_h3_map[c] for c in unique_cells], dtype=np.int64)
        pos = {int(r): i for i, r in enumerate(ranks)}  # the reference numbers cells in the order given
        cells = self._t.order[ranks]
        disk = self._disk[cells]
        valid = disk >= 0
        s_local = np.broadcast_to(np.arange(ranks.size)[:, None], disk.shape)[valid]
        d_cell = disk[valid]
        d_rank = self._t.rank[d_cell]
        keep = np.isin(d_rank, ranks)
        s_local, d_cell, d_rank = s_local[keep], d_cell[keep], d_rank[keep]
        d_local = np.array([pos[int(r)] for r in d_rank], dtype=np.int64)
        s_cell = cells[s_local]
        attr = graphs._sincos_attr(np.degrees(self._t.lat[s_cell]), np.degrees(self._t.lng[s_cell]), np.degrees(self._t.lat[d_cell]),
                                   np.degrees(self._t.lng[d_cell]))  # fmt: skip
        ei = np.stack([s_local.astype(np.int64), d_local])
        return Data(edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(attr.reshape(-1, 2)))

    def __call__(self, lat_lons):
        """(:130-155) (encoder_graph, decoder_graph, latent_graph, h3_indices), cached per list object."""
        if lat_lons is self._prev_lat_lons:
            return (self._cached_encoder_graph, self._cached_decoder_graph, self._cached_latent_graph, self._cached_h3_indices)
        graphs.validate_lat_lons(lat_lons)
        encoder_graph, h3_indices = self.build_encoder_graph(lat_lons)
        _, unique_cells, _ = self._assign_h3_cells(lat_lons)
        decoder_graph = self.build_decoder_graph(lat_lons)
        latent_graph = self.build_latent_graph(unique_cells)
        self._prev_lat_lons = lat_lons
        self._cached_encoder_graph, self._cached_decoder_graph = encoder_graph, decoder_graph
        self._cached_latent_graph, self._cached_h3_indices = latent_graph, h3_indices
        return encoder_graph, decoder_graph, latent_graph, h3_indices
