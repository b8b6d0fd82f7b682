"""H3-compatible hexagonal indexing, restated from the published H3 algorithm (no `h3` wheel exists
in this image and there is no network).

The reference builds its three graphs with the `h3` C library (pinned h3==4.3.1, pyproject.toml:74):
    encoder.py:76-109,244-268; assimilator_decoder.py:69-106; assimilator_encoder.py:74,170-242.
The seven functions it calls are restated here with the same names and argument conventions:
    get_res0_cells, uncompact_cells, latlng_to_cell, cell_to_latlng, grid_disk, great_circle_distance,
    get_num_cells.

What is exact with respect to H3 (geometry):
  * the icosahedron orientation: the 20 face centres and the Class II i-axis azimuths are H3's
    `faceCenterGeo` / `faceAxesAzRadsCII` constants (checked here: the centres form an exact
    icosahedron and every i-axis points at a face vertex to 1e-16);
  * point -> cell membership: nearest face, gnomonic projection, `RES0_U_GNOMONIC` scaling, sqrt(7)
    per resolution, Class III rotation asin(sqrt(3/28)), hexagonal rounding in the face plane;
  * cell centres (inverse gnomonic of the lattice point on the owning face), the aperture-7
    parent/child relation, neighbour sets (grid_disk k=1), the pentagon set (12 icosahedron vertices),
    the cell counts 2+120*7^res, and the haversine `great_circle_distance`.
What is best-effort (cannot be pinned without the h3 wheel): the 64-bit index *values* (base-cell
numbers by centre latitude north->south, digits in the home face's ijk frame).  Index values only
determine the order of `sorted(cells)`, i.e. the mesh-node numbering; graph topology and edge
attributes do not depend on them.  The count KATs of the reference tests (5882 cells / 41162 latent
edges at res 2, tests/test_model.py:30-31) hold and are asserted in tests/test_h3lite.py.

Indices are returned as 15-hex-digit strings like h3-py v4, so `sorted()` on them is numeric order.
Everything is vectorised numpy + one scipy cKDTree per resolution; tables are cached per resolution.
"""

from __future__ import annotations

import functools
import math

import numpy as np
from scipy.spatial import cKDTree

# --- H3 constants (h3/src/h3lib/lib/faceijk.c, constants.h) -------------------------------------
_FACE_CENTER_GEO = np.array(
    [
        [0.803582649718989942, 1.248397419617396099],
        [1.307747883455638156, 2.536945009877921159],
        [1.054751253523952054, -1.347517358900396623],
        [0.600191595538186799, -0.450603909469755746],
        [0.491715428198773866, 0.401988202911306943],
        [0.172745327415618701, 1.678146885280433686],
        [0.605929321571350690, 2.953923329812411617],
        [0.427370518328979641, -1.888876200336285401],
        [-0.079066118549212831, -0.733429513380867741],
        [-0.230961644455383637, 0.506495587332349035],
        [0.079066118549212831, 2.408163140208925497],
        [0.230961644455383637, -2.635097066257444203],
        [-0.172745327415618701, -1.463445768309359553],
        [-0.605929321571350690, -0.187669323777381622],
        [-0.427370518328979641, 1.252716453253507838],
        [-0.600191595538186799, 2.690988744120037492],
        [-0.491715428198773866, -2.739604450678486295],
        [-0.803582649718989942, -1.893195233972397139],
        [-1.307747883455638156, -0.604647643711872080],
        [-1.054751253523952054, 1.794075294689396615],
    ]
)
# azimuth (clockwise from north) of the Class II i-axis at each face centre
_FACE_AXES_AZ_I = np.array(
    [
        5.619958268523939882, 5.760339081714187279, 0.780213654393430055, 0.430469363979999913,
        6.130269123335111400, 2.692877706530642877, 2.982963003477243874, 3.532912002790141181,
        3.494305004259568154, 3.003214169499538391, 5.930472956509811562, 0.138378484090254847,
        0.448714947059150361, 0.158629650112549365, 5.891865957979238535, 2.711123289609793325,
        3.294508837434268316, 3.804819692245439833, 3.664438879055192436, 2.361378999196363184,
    ]
)
_RES0_U_GNOMONIC = 0.38196601125010500003
_M_AP7_ROT_RADS = 0.333473172251832115336090755351601070065900389
_SQRT7 = math.sqrt(7.0)
_SIN60 = math.sqrt(3.0) / 2.0
MAX_RES = 6  # table sizes grow as 7^res; the reference uses res 2 (default) everywhere on this path

# H3 digit -> unit ijk vector; planar directions: i at 0 deg, j at 120 deg, k at 240 deg (CCW)
_DIGIT_ANGLE_DEG = {4: 0.0, 6: 60.0, 2: 120.0, 3: 180.0, 1: 240.0, 5: 300.0}
_CCW_DIGITS = [4, 6, 2, 3, 1, 5]  # digits met going counter-clockwise from the i axis


def _geo_to_vec(lat, lng):
    cl = np.cos(lat)
    return np.stack([cl * np.cos(lng), cl * np.sin(lng), np.sin(lat)], axis=-1)


def _vec_to_geo(v):
    lat = np.arcsin(np.clip(v[..., 2], -1.0, 1.0))
    lng = np.arctan2(v[..., 1], v[..., 0])
    return lat, lng


@functools.lru_cache(maxsize=None)
def _faces():
    """Face frames: centre c, in-plane unit vectors ex (Class II i-axis) and ey (90 deg CCW seen from outside)."""
    lat, lng = _FACE_CENTER_GEO[:, 0], _FACE_CENTER_GEO[:, 1]
    c = _geo_to_vec(lat, lng)
    north = np.stack([-np.sin(lat) * np.cos(lng), -np.sin(lat) * np.sin(lng), np.cos(lat)], axis=-1)
    east = np.stack([-np.sin(lng), np.cos(lng), np.zeros_like(lng)], axis=-1)
    az = _FACE_AXES_AZ_I
    ex = np.cos(az)[:, None] * north + np.sin(az)[:, None] * east
    ey = np.cross(c, ex)
    return c, ex, ey


def _class3(res: int) -> bool:
    return res % 2 == 1


def _lattice_to_plane(a, b, res):
    """Lattice point a*i + b*j (i at 0 deg, j at 120 deg) -> gnomonic-plane (x, y) on the face."""
    x = a - 0.5 * b
    y = _SIN60 * b
    if _class3(res):  # the Class III lattice is the Class II frame rotated CCW by asin(sqrt(3/28))
        cr, sr = math.cos(_M_AP7_ROT_RADS), math.sin(_M_AP7_ROT_RADS)
        x, y = cr * x - sr * y, sr * x + cr * y
    s = _RES0_U_GNOMONIC / (_SQRT7**res)
    return x * s, y * s


def _plane_to_lattice(x, y, res):
    """Nearest lattice point to plane coordinates (hexagonal rounding == H3's _hex2dToCoordIJK cell)."""
    s = (_SQRT7**res) / _RES0_U_GNOMONIC
    x, y = x * s, y * s
    if _class3(res):
        cr, sr = math.cos(_M_AP7_ROT_RADS), math.sin(_M_AP7_ROT_RADS)
        x, y = cr * x + sr * y, -sr * x + cr * y
    b = y / _SIN60
    a = x + 0.5 * b
    # 60-degree axial coordinates (q along i, r along i+j): p = (a-b)*i + b*(i+j)
    q, r = a - b, b
    s3 = -q - r
    rq, rr, rs = np.rint(q), np.rint(r), np.rint(s3)
    dq, dr, ds = np.abs(rq - q), np.abs(rr - r), np.abs(rs - s3)
    fix_q = (dq > dr) & (dq > ds)
    fix_r = (~fix_q) & (dr > ds)
    rq = np.where(fix_q, -rr - rs, rq)
    rr = np.where(fix_r, -rq - rs, rr)
    return (rq + rr).astype(np.int64), rr.astype(np.int64)


def _face_plane_to_vec(face, x, y):
    c, ex, ey = _faces()
    v = c[face] + x[..., None] * ex[face] + y[..., None] * ey[face]
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


class _ResTable:
    """All cells of one resolution: unit-vector centres, kd-tree, neighbour lists, parent links, indices."""

    def __init__(self, res: int):
        if not 0 <= res <= MAX_RES:
            raise ValueError(f"h3lite supports resolutions 0..{MAX_RES}, got {res}")
        self.res = res
        c, ex, ey = _faces()
        rad = 2.0 * _RES0_U_GNOMONIC  # in-plane distance face centre -> vertex
        n = int(math.ceil(2.0 * (_SQRT7**res) * 2.0 / math.sqrt(3.0))) + 2  # axial coords reach 2/sqrt(3) x the radius
        aa, bb = np.meshgrid(np.arange(-n, n + 1), np.arange(-n, n + 1), indexing="ij")
        aa, bb = aa.ravel(), bb.ravel()
        x, y = _lattice_to_plane(aa, bb, res)
        # triangle with vertices at angles 0, 120, 240 deg and circumradius rad: three half-planes
        eps = 1e-9
        inside = np.ones_like(x, dtype=bool)
        for ang in (60.0, 180.0, 300.0):  # outward edge normals
            nx, ny = math.cos(math.radians(ang)), math.sin(math.radians(ang))
            inside &= (x * nx + y * ny) <= rad * 0.5 + eps
        x, y = x[inside], y[inside]
        vecs, faces = [], []
        for f in range(20):
            vecs.append(_face_plane_to_vec(f, x, y))
            faces.append(np.full(x.shape, f))
        vecs = np.concatenate(vecs)
        faces = np.concatenate(faces)
        # dedupe cells shared by several faces (edge / vertex centred); keep the lowest face as owner
        tree = cKDTree(vecs)
        spacing = _RES0_U_GNOMONIC / (_SQRT7**res)
        groups = tree.query_ball_point(vecs, r=spacing * 1e-3)
        keep = np.array([min(g) == i for i, g in enumerate(groups)])
        self.center = vecs[keep]
        self.owner_face = faces[keep]
        self.n_faces_sharing = np.array([len(g) for g in groups])[keep]
        self.num = self.center.shape[0]
        assert self.num == 2 + 120 * 7**res, (self.num, res)
        self.tree = cKDTree(self.center)
        self.is_pentagon = self.n_faces_sharing == 5
        assert int(self.is_pentagon.sum()) == 12
        # neighbours: the 6 (5 at pentagons) nearest other centres
        _, nn = self.tree.query(self.center, k=7)
        assert np.all(nn[:, 0] == np.arange(self.num))
        self.nbr = nn[:, 1:].copy()
        self.nbr[self.is_pentagon, 5] = -1
        self.lat, self.lng = _vec_to_geo(self.center)
        self.parent = None
        self.index = None  # uint64 H3-style index per cell
        self.order = None  # permutation: rank -> cell id, ascending index

    # -- point location ---------------------------------------------------------------------
    def locate(self, lat_rad, lng_rad):
        """H3 _geoToFaceIjk: nearest face, gnomonic projection, hex rounding -> canonical cell id."""
        c, ex, ey = _faces()
        v = _geo_to_vec(np.asarray(lat_rad, dtype=np.float64), np.asarray(lng_rad, dtype=np.float64))
        face = np.argmax(v @ c.T, axis=-1)
        cf = c[face]
        q = v / np.sum(v * cf, axis=-1, keepdims=True) - cf
        x = np.sum(q * ex[face], axis=-1)
        y = np.sum(q * ey[face], axis=-1)
        a, b = _plane_to_lattice(x, y, self.res)
        px, py = _lattice_to_plane(a, b, self.res)
        approx = _face_plane_to_vec(face, px, py)  # exact on the face, <3% of a spacing off just beyond an edge
        dist, cell = self.tree.query(approx, k=1)
        spacing = _RES0_U_GNOMONIC / (_SQRT7**self.res)
        if np.any(dist > 0.25 * spacing):
            raise AssertionError("h3lite: ambiguous lattice point resolution")
        return cell


@functools.lru_cache(maxsize=None)
def _table(res: int) -> _ResTable:
    t = _ResTable(res)
    _assign_indices(t)
    return t


def _tangent_angle(face, origin_vec, target_vec):
    """CCW angle (deg, seen from outside) of origin->target measured from the face's i-axis, in the face plane."""
    c, ex, ey = _faces()

    def proj(v):
        q = v / np.sum(v * c[face], axis=-1, keepdims=True) - c[face]
        return np.sum(q * ex[face], axis=-1), np.sum(q * ey[face], axis=-1)

    ox, oy = proj(origin_vec)
    tx, ty = proj(target_vec)
    return np.degrees(np.arctan2(ty - oy, tx - ox)) % 360.0


def _assign_indices(t: _ResTable):
    """Base cell + digits.  Base cells: numbered by centre latitude, north to south (H3 docs).  Digits:
    direction of each cell from its parent, in the base cell's home-face ijk frame (see module docstring)."""
    res = t.res
    if res == 0:
        order = np.argsort(-t.lat, kind="stable")
        bc = np.empty(t.num, dtype=np.int64)
        bc[order] = np.arange(t.num)
        t.base_cell = bc
        t.home_face = t.owner_face.copy()
        t.digits = np.zeros((t.num, 0), dtype=np.int64)
    else:
        p = _table(res - 1)
        _, par = p.tree.query(t.center, k=1)
        t.parent = par
        t.base_cell = p.base_cell[par]
        t.home_face = p.home_face[par]
        digit = np.zeros(t.num, dtype=np.int64)
        rot = math.degrees(_M_AP7_ROT_RADS) if _class3(res) else 0.0
        is_child_center = np.linalg.norm(t.center - p.center[par], axis=-1) < 1e-9
        ang = _tangent_angle(t.home_face, p.center[par], t.center)
        # hexagon parents: nearest of the six lattice directions (Class III children sit on the rotated lattice)
        # The six children are taken in CCW order and the whole ring is anchored at the child that sits
        # closest to a lattice direction, so gnomonic distortion beyond a face edge cannot make two collide.
        rel = ((ang - rot) % 360.0) / 60.0
        hexkid = np.nonzero(~is_child_center & ~p.is_pentagon[par])[0]
        o = np.lexsort((rel[hexkid], par[hexkid]))
        hk = hexkid[o].reshape(-1, 6)  # rows: one hexagon parent, children CCW by angle
        assert np.all(par[hk] == par[hk[:, :1]])
        r6 = rel[hk]
        dev = np.abs(r6 - np.rint(r6))
        anchor = np.argmin(dev, axis=1)
        k_anchor = np.rint(r6[np.arange(hk.shape[0]), anchor]).astype(np.int64) % 6
        pos = (np.arange(6)[None, :] - anchor[:, None]) % 6  # CCW steps from the anchor child
        digit[hk] = np.array(_CCW_DIGITS)[(k_anchor[:, None] + pos) % 6]
        # pentagon parents: 5 children in CCW order take the CCW digit sequence with the K axis deleted,
        # anchored at the child lying inside the home face
        for pc in np.nonzero(p.is_pentagon)[0]:
            kids = np.nonzero((par == pc) & ~is_child_center)[0]
            assert kids.size == 5, kids.size
            c, ex, ey = _faces()
            hf = p.home_face[pc]
            to_face = _tangent_angle(hf, p.center[pc], c[hf][None, :])[0]
            a = ang[kids]
            start = int(np.argmin(np.minimum((a - to_face) % 360.0, (to_face - a) % 360.0)))
            ccw = kids[np.argsort((a - a[start]) % 360.0)]
            k0 = int(np.rint(((to_face - rot) % 360.0) / 60.0)) % 6
            seq = [d for d in (_CCW_DIGITS[k0:] + _CCW_DIGITS[:k0]) if d != 1]
            if _CCW_DIGITS[k0] == 1:  # home-face direction coincides with the deleted axis: start from the next
                seq = [d for d in (_CCW_DIGITS[k0 + 1 :] + _CCW_DIGITS[: k0 + 1]) if d != 1]
            digit[ccw] = np.array(seq)
        digit[is_child_center] = 0
        t.digits = np.concatenate([p.digits[par], digit[:, None]], axis=1)
        # the 7 (6) children of a parent must carry distinct digits
        key = par * 8 + digit
        assert np.unique(key).size == t.num, "h3lite: digit collision"
    idx = np.full(t.num, (1 << 59) | (res << 52), dtype=np.uint64)
    idx |= t.base_cell.astype(np.uint64) << np.uint64(45)
    for r in range(1, 16):
        d = t.digits[:, r - 1].astype(np.uint64) if r <= res else np.full(t.num, 7, dtype=np.uint64)
        idx |= d << np.uint64(3 * (15 - r))
    t.index = idx
    t.order = np.argsort(idx, kind="stable")
    t.rank = np.empty(t.num, dtype=np.int64)
    t.rank[t.order] = np.arange(t.num)
    t.index_to_cell = {int(v): i for i, v in enumerate(idx)}


# --- public API mirroring h3-py v4 ---------------------------------------------------------------
def _to_str(idx: int) -> str:
    return format(int(idx), "x")


def _to_int(h) -> int:
    return int(h, 16) if isinstance(h, str) else int(h)


def get_resolution(h) -> int:
    """h3.get_resolution"""
    return (_to_int(h) >> 52) & 0xF


def _cell_id(h):
    v = _to_int(h)
    t = _table((v >> 52) & 0xF)
    try:
        return t, t.index_to_cell[v]
    except KeyError as e:  # same exception class h3-py raises for bad cells (H3CellInvalidError is a ValueError)
        raise ValueError(f"invalid H3 cell {h!r}") from e


def get_num_cells(res: int) -> int:
    """h3.get_num_cells (used at encoder.py:113, assimilator_encoder.py:80)."""
    return 2 + 120 * 7**res


def get_res0_cells():
    """h3.get_res0_cells (encoder.py:76)."""
    t = _table(0)
    return [_to_str(v) for v in t.index[t.order]]


def uncompact_cells(cells, res: int):
    """h3.uncompact_cells (encoder.py:76): all descendants at `res` of the given cells."""
    out = []
    by_res = {}
    for h in cells:
        by_res.setdefault(get_resolution(h), []).append(h)
    tt = _table(res)
    for r0, hs in by_res.items():
        if r0 > res:
            raise ValueError("cannot uncompact to a coarser resolution")
        anc = np.arange(tt.num)
        for r in range(res, r0, -1):
            anc = _table(r).parent[anc]
        t0 = _table(r0)
        want = np.zeros(t0.num, dtype=bool)
        for h in hs:
            want[_cell_id(h)[1]] = True
        sel = np.nonzero(want[anc])[0]
        out.extend(_to_str(v) for v in tt.index[sel])
    return out


def latlng_to_cell(lat: float, lng: float, res: int) -> str:
    """h3.latlng_to_cell, degrees in (encoder.py:78)."""
    t = _table(res)
    cell = t.locate(np.radians([lat]), np.radians([lng]))[0]
    return _to_str(t.index[cell])


def cell_to_latlng(h):
    """h3.cell_to_latlng -> (lat, lng) degrees (encoder.py:90)."""
    t, c = _cell_id(h)
    return (math.degrees(t.lat[c]), math.degrees(t.lng[c]))


def grid_disk(h, k: int = 1):
    """h3.grid_disk(h, 1): the cell itself then its neighbours (encoder.py:256, assimilator_decoder.py:94).
    Order: origin first, then neighbours by increasing index (h3's own ring order is not reproduced; it only
    affects the order in which edges are listed, not the graph)."""
    if k != 1:
        raise NotImplementedError("h3lite.grid_disk supports k=1 (all the reference uses on this path)")
    t, c = _cell_id(h)
    nb = t.nbr[c]
    nb = nb[nb >= 0]
    nb = nb[np.argsort(t.index[nb])]
    return [_to_str(t.index[c])] + [_to_str(v) for v in t.index[nb]]


def is_pentagon(h) -> bool:
    """h3.is_pentagon"""
    t, c = _cell_id(h)
    return bool(t.is_pentagon[c])


def great_circle_distance(a, b, unit: str = "km") -> float:
    """h3.great_circle_distance: haversine exactly as H3's greatCircleDistanceRads (latLng.c); degrees in."""
    lat1, lng1 = math.radians(a[0]), math.radians(a[1])
    lat2, lng2 = math.radians(b[0]), math.radians(b[1])
    s_lat = math.sin((lat2 - lat1) * 0.5)
    s_lng = math.sin((lng2 - lng1) * 0.5)
    aa = s_lat * s_lat + math.cos(lat1) * math.cos(lat2) * s_lng * s_lng
    d = 2.0 * math.atan2(math.sqrt(aa), math.sqrt(1.0 - aa))
    if unit == "rads":
        return d
    if unit == "km":
        return d * 6371.007180918475
    if unit == "m":
        return d * 6371007.180918475
    raise ValueError(unit)


# --- vectorised helpers used by graph_weather_b200.graphs (not part of the h3 API) ------------------
def haversine_rads(lat1, lng1, lat2, lng2):
    """Vectorised great_circle_distance(unit='rads'); radians in."""
    s_lat = np.sin((lat2 - lat1) * 0.5)
    s_lng = np.sin((lng2 - lng1) * 0.5)
    aa = s_lat * s_lat + np.cos(lat1) * np.cos(lat2) * s_lng * s_lng
    return 2.0 * np.arctan2(np.sqrt(aa), np.sqrt(1.0 - aa))


def table(res: int) -> _ResTable:
    """Cached per-resolution cell table (centres, neighbours, index order)."""
    return _table(res)


def device_tables(res: int):
    """The arrays gw_plan_set_h3_tables uploads for the device-side point location (csrc/gw_graph.cu mirrors `_ResTable.locate`):
    face frames [20, 9] (centre, i-axis, j-axis), the per-face lattice table cell_of [20, 2n+1, 2n+1] (canonical cell of lattice
    point (a, b) on the face, -1 where no cell lies), cell centres as `graphs._sincos_attr` sees them (radians taken through
    degrees and back), and the plane -> lattice transform (scale, rotation cos / sin)."""
    t = _table(res)
    c, ex, ey = _faces()
    frames = np.concatenate([c, ex, ey], axis=1).astype(np.float64)
    n = int(math.ceil(2.0 * (_SQRT7**res) * 2.0 / math.sqrt(3.0))) + 2
    aa, bb = np.meshgrid(np.arange(-n, n + 1), np.arange(-n, n + 1), indexing="ij")
    px, py = _lattice_to_plane(aa.ravel(), bb.ravel(), res)
    spacing = _RES0_U_GNOMONIC / (_SQRT7**res)
    cell_of = np.full((20, (2 * n + 1) ** 2), -1, dtype=np.int32)
    for f in range(20):
        vec = _face_plane_to_vec(np.full(px.shape, f), px, py)
        dist, cell = t.tree.query(vec, k=1)
        ok = dist <= 0.25 * spacing  # the same acceptance test as `locate`
        cell_of[f, ok] = cell[ok]
    rot = _class3(res)
    return dict(
        res=res, n_cells=t.num, lattice_n=n, frames=frames, cell_of=cell_of.reshape(20, 2 * n + 1, 2 * n + 1),
        cell_lat=np.radians(np.degrees(t.lat)), cell_lng=np.radians(np.degrees(t.lng)),
        scale=(_SQRT7**res) / _RES0_U_GNOMONIC, rot_cos=math.cos(_M_AP7_ROT_RADS) if rot else 1.0, rot_sin=math.sin(_M_AP7_ROT_RADS) if rot else 0.0,
        rank=t.rank.astype(np.int64),
    )  # fmt: skip
