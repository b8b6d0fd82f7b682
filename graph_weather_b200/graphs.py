"""Host-side graph construction for the encode-process-decode path (vectorised numpy, one-off per lat_lons).

Mirrors, array for array, what the reference builds with Python loops over `h3`:
  * encoder bipartite graph      encoder.py:76-109            (one edge per lat/lon point -> its mesh cell)
  * latent mesh graph            encoder.py:244-268           (each cell -> every cell of grid_disk(cell, 1))
  * decoder bipartite graph      assimilator_decoder.py:69-106 (every cell of grid_disk(cell(p),1) -> point p)
including the reference's numbering quirk: encoder and decoder number mesh nodes in DESCENDING sorted-index
order (encoder.py:80-84, assimilator_decoder.py:72-77) while the latent graph numbers them ascending
(encoder.py:77).  Outputs are the reference's COO arrays (int64 edge_index, float32 [sin d, cos d]) plus the
target-sorted int32 forms the CUDA kernels consume.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import h3lite


def _as_latlon_array(lat_lons) -> np.ndarray:
    ll = np.asarray(lat_lons, dtype=np.float64)
    if ll.ndim != 2 or ll.shape[1] < 2:
        raise ValueError("lat_lons must be a sequence of (lat, lon) pairs")
    return ll[:, :2]


def validate_lat_lons(lat_lons) -> None:
    """Same checks and messages as graph_weather/utils.py:6-12: non-empty, every latitude inside [-90, 90]."""
    if len(lat_lons) == 0:
        raise ValueError("lat_lons must not be empty.")
    for index, (lat, _lon) in enumerate(lat_lons):
        if not (-90.0 <= lat <= 90.0):
            raise ValueError(f"Coordinate {index}: latitude {lat} is outside [-90, 90].")


def _sincos_attr(lat1_deg, lng1_deg, lat2_deg, lng2_deg) -> np.ndarray:
    d = h3lite.haversine_rads(np.radians(lat1_deg), np.radians(lng1_deg), np.radians(lat2_deg), np.radians(lng2_deg))
    return np.stack([np.sin(d), np.cos(d)], axis=1).astype(np.float32)


def _disk_table(t) -> np.ndarray:
    """[num_cells, 7] cell ids of grid_disk(cell, 1) in h3lite.grid_disk order (self, then neighbours by index);
    -1 pads the 12 pentagons."""
    nbr = t.nbr
    key = np.where(nbr >= 0, t.index[np.maximum(nbr, 0)], np.iinfo(np.uint64).max)
    o = np.argsort(key, axis=1, kind="stable")
    nbr_sorted = np.take_along_axis(nbr, o, axis=1)
    return np.concatenate([np.arange(t.num)[:, None], nbr_sorted], axis=1)


@dataclass
class MeshGraph:
    """Latent graph (encoder.py:244-268): nodes = cells in ascending index order."""

    num_h3: int
    edge_index: np.ndarray  # [2, El] int64, reference order (by source, then disk order)
    edge_attr: np.ndarray  # [El, 2] float32
    # target-sorted view for the kernels: perm[j] = reference edge id of sorted edge j
    perm: np.ndarray = field(default=None)
    src: np.ndarray = field(default=None)  # [El] int32, sorted order
    dst: np.ndarray = field(default=None)  # [El] int32, sorted order (non-decreasing)
    ptr: np.ndarray = field(default=None)  # [H+1] int32 CSR over dst


def _finish_target_sorted(src, dst, num_targets):
    perm = np.lexsort((np.arange(dst.size), dst))  # by target, ties in reference edge order
    s, d = src[perm].astype(np.int32), dst[perm].astype(np.int32)
    ptr = np.zeros(num_targets + 1, dtype=np.int32)
    np.cumsum(np.bincount(d, minlength=num_targets), out=ptr[1:])
    return perm.astype(np.int64), s, d, ptr


def build_mesh_graph(resolution: int = 2) -> MeshGraph:
    t = h3lite.table(resolution)
    H = t.num
    disk = _disk_table(t)[t.order]  # row i = disk of the cell with rank i
    valid = disk >= 0
    src_rank = np.broadcast_to(np.arange(H)[:, None], disk.shape)[valid]
    dst_cell = disk[valid]
    dst_rank = t.rank[dst_cell]
    src_cell = t.order[src_rank]
    attr = _sincos_attr(
        np.degrees(t.lat[src_cell]), np.degrees(t.lng[src_cell]), np.degrees(t.lat[dst_cell]), np.degrees(t.lng[dst_cell])
    )
    g = MeshGraph(H, np.stack([src_rank, dst_rank]).astype(np.int64), attr)
    g.perm, g.src, g.dst, g.ptr = _finish_target_sorted(src_rank, dst_rank, H)
    return g


@dataclass
class EncoderGraph:
    """Bipartite lat/lon -> mesh graph (encoder.py:76-109). Node ids: points 0..N-1, mesh N + (H-1-rank)."""

    num_latlons: int
    num_h3: int
    edge_index: np.ndarray  # [2, N] int64
    edge_attr: np.ndarray  # [N, F] float32 ([sin d, cos d] or [sin d, cos d, height])
    mesh_local: np.ndarray  # [N] int32: H-1-rank(cell(p)), the mesh-node slot each point feeds
    perm: np.ndarray = field(default=None)  # [N] target-sorted order of points
    ptr: np.ndarray = field(default=None)  # [H+1] CSR over mesh slots into perm


def build_encoder_graph(lat_lons, resolution: int = 2, heights=None) -> EncoderGraph:
    ll = _as_latlon_array(lat_lons)
    t = h3lite.table(resolution)
    N, H = ll.shape[0], t.num
    cell = t.locate(np.radians(ll[:, 0]), np.radians(ll[:, 1]))
    mesh_local = (H - 1 - t.rank[cell]).astype(np.int64)
    attr = _sincos_attr(ll[:, 0], ll[:, 1], np.degrees(t.lat[cell]), np.degrees(t.lng[cell]))
    if heights is not None:  # assimilator_encoder.py:198-202
        attr = np.concatenate([attr, np.asarray(heights, dtype=np.float32).reshape(-1, 1)], axis=1)
    ei = np.stack([np.arange(N, dtype=np.int64), mesh_local + N])
    g = EncoderGraph(N, H, ei, attr, mesh_local.astype(np.int32))
    perm, _, _, ptr = _finish_target_sorted(np.arange(N), mesh_local, H)
    g.perm, g.ptr = perm.astype(np.int32), ptr
    return g


@dataclass
class DecoderGraph:
    """Bipartite mesh -> lat/lon graph (assimilator_decoder.py:69-106). Node ids: mesh H-1-rank, points H+p.
    Edges are already target-sorted by construction (all edges of point p are contiguous)."""

    num_latlons: int
    num_h3: int
    edge_index: np.ndarray  # [2, Ed] int64
    edge_attr: np.ndarray  # [Ed, 2] float32
    src: np.ndarray  # [Ed] int32 mesh slot
    ptr: np.ndarray  # [N+1] int32 CSR over points


def build_decoder_graph(lat_lons, resolution: int = 2) -> DecoderGraph:
    ll = _as_latlon_array(lat_lons)
    t = h3lite.table(resolution)
    N, H = ll.shape[0], t.num
    cell = t.locate(np.radians(ll[:, 0]), np.radians(ll[:, 1]))
    disk = _disk_table(t)[cell]  # [N, 7]
    valid = disk >= 0
    p_of_edge = np.broadcast_to(np.arange(N)[:, None], disk.shape)[valid]
    h_cell = disk[valid]
    src = (H - 1 - t.rank[h_cell]).astype(np.int64)
    attr = _sincos_attr(ll[p_of_edge, 0], ll[p_of_edge, 1], np.degrees(t.lat[h_cell]), np.degrees(t.lng[h_cell]))
    ptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(valid.sum(axis=1), out=ptr[1:])
    ei = np.stack([src, p_of_edge.astype(np.int64) + H])
    return DecoderGraph(N, H, ei, attr, src.astype(np.int32), ptr)


def replicate_edge_index(edge_index: np.ndarray, batch: int) -> np.ndarray:
    """The reference's batch replication `ei + i*max(ei) + i` (encoder.py:212-218, assimilator_decoder.py:180-186)."""
    m = int(edge_index.max())
    return np.concatenate([edge_index + i * m + i for i in range(batch)], axis=1)
