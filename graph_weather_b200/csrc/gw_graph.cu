// gw_graph.cu -- device-side construction of the per-call observation graph of the assimilator
// (AssimilatorEncoder.create_input_graph, graph_weather/models/layers/assimilator_encoder.py:170-216).
//
// The reference loops over the observations in Python for every forward: h3.latlng_to_cell(lat, lon, res), the haversine
// distance to the cell centre, edge_attr = [sin d, cos d, height], one edge observation -> mesh node.  Here one kernel
// does the point location and the edge attributes for all observations in double precision with the same formulas as the
// host restatement (graph_weather_b200/h3lite.py: nearest icosahedron face, gnomonic projection, hexagonal rounding on the
// face lattice, canonical cell through a per-face lattice table), a stable radix sort groups the observations by mesh slot
// in observation order (the order scatter_sum adds them in, graph_net_block.py:188), and a binary search per slot gives the
// CSR.  Nothing is copied to the host and nothing synchronises.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "gw_internal.h"

namespace gw {

constexpr double kSin60 = 0.86602540378443864676;
constexpr double kPi180 = 0.017453292519943295769;  // numpy.radians: x * (pi / 180)

// lat / lng in degrees -> canonical cell id at the table's resolution, or -1
__device__ int h3_locate(const H3Tables& t, double lat_deg, double lng_deg) {
  const double lat = lat_deg * kPi180, lng = lng_deg * kPi180;
  const double cl = cos(lat);
  const double vx = cl * cos(lng), vy = cl * sin(lng), vz = sin(lat);
  int face = 0;
  double best = -2.0;
  for (int f = 0; f < 20; ++f) {  // nearest face centre (first maximum, like numpy.argmax)
    const double* c = t.frames + 9 * f;
    const double d = vx * c[0] + vy * c[1] + vz * c[2];
    if (d > best) best = d, face = f;
  }
  const double* c = t.frames + 9 * face;
  const double dn = vx * c[0] + vy * c[1] + vz * c[2];  // (this file is compiled with -fmad=false: products and sums round separately, like numpy)
  const double qx = vx / dn - c[0], qy = vy / dn - c[1], qz = vz / dn - c[2];  // gnomonic projection onto the face plane
  double x = (qx * c[3] + qy * c[4] + qz * c[5]) * t.scale, y = (qx * c[6] + qy * c[7] + qz * c[8]) * t.scale;
  const double xr = t.cr * x + t.sr * y, yr = -t.sr * x + t.cr * y;  // Class III lattices are rotated
  x = xr, y = yr;
  const double b = y / kSin60, a = x + 0.5 * b;
  const double q = a - b, r = b, s = -q - r;  // cube coordinates; round, then repair the largest rounding error
  double rq = rint(q), rr = rint(r);
  const double rs = rint(s);
  const double dq = fabs(rq - q), dr = fabs(rr - r), ds = fabs(rs - s);
  const bool fix_q = dq > dr && dq > ds, fix_r = !fix_q && dr > ds;
  if (fix_q) rq = -rr - rs;
  if (fix_r) rr = -rq - rs;
  const int ai = (int)(rq + rr), bi = (int)rr;
  const int w = 2 * t.lat_n + 1;
  if (ai < -t.lat_n || ai > t.lat_n || bi < -t.lat_n || bi > t.lat_n) return -1;
  return __ldg(t.cell_of + (size_t)face * w * w + (size_t)(ai + t.lat_n) * w + (bi + t.lat_n));
}

// per observation: mesh slot, edge attributes [sin d, cos d, height], sort key / value
__global__ void gw_obs_locate_kernel(H3Tables t, const float* __restrict__ llh, int n, int32_t* __restrict__ slot, float* __restrict__ attr,
                                     uint32_t* __restrict__ key, int32_t* __restrict__ val, int32_t* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lat = (double)llh[3 * i], lng = (double)llh[3 * i + 1];
  const int cell = h3_locate(t, lat, lng);
  if (cell < 0) {  // cannot happen for finite coordinates; recorded instead of writing out of bounds
    if (status) atomicOr(status, 16);
    slot[i] = 0, key[i] = 0u, val[i] = i;
    attr[3 * i] = attr[3 * i + 1] = attr[3 * i + 2] = 0.f;
    return;
  }
  const int s = __ldg(t.cell_slot + cell);
  // haversine exactly as H3's greatCircleDistanceRads / h3lite.haversine_rads: point first, cell centre second
  const double lat1 = lat * kPi180, lng1 = lng * kPi180, lat2 = __ldg(t.cell_lat + cell), lng2 = __ldg(t.cell_lng + cell);
  const double s_lat = sin((lat2 - lat1) * 0.5), s_lng = sin((lng2 - lng1) * 0.5);
  const double aa = s_lat * s_lat + cos(lat1) * cos(lat2) * s_lng * s_lng;
  const double d = 2.0 * atan2(sqrt(aa), sqrt(1.0 - aa));
  slot[i] = s, key[i] = (uint32_t)s, val[i] = i;
  attr[3 * i] = (float)sin(d), attr[3 * i + 1] = (float)cos(d), attr[3 * i + 2] = llh[3 * i + 2];
}

// ptr[s] = first position of key >= s in the sorted keys (s = 0 .. n_slots)
__global__ void gw_csr_from_sorted_kernel(const uint32_t* __restrict__ keys, int n, int n_slots, int32_t* __restrict__ ptr) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_slots) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < (uint32_t)s) lo = mid + 1; else hi = mid;
  }
  ptr[s] = lo;
}

size_t obs_graph_workspace_bytes(int n) {
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, n);
  return ((sort_bytes + 255) / 256) * 256 + 3 * (((size_t)n * 4 + 255) / 256) * 256;
}

// slot / perm / ptr / attr are the plan's encoder-graph arrays (gw_plan_set_encoder_graph layout); ws >= obs_graph_workspace_bytes(n)
cudaError_t launch_obs_graph(const H3Tables& t, const float* llh, int n, int n_slots, int32_t* slot, int32_t* perm, int32_t* ptr, float* attr,
                             void* ws, size_t ws_bytes, int32_t* status, cudaStream_t st) {
  if (n <= 0) return cudaErrorInvalidValue;
  const size_t nb = (((size_t)n * 4 + 255) / 256) * 256;
  uint8_t* w = static_cast<uint8_t*>(ws);
  uint32_t* key = reinterpret_cast<uint32_t*>(w);
  uint32_t* key_sorted = reinterpret_cast<uint32_t*>(w + nb);
  int32_t* val = reinterpret_cast<int32_t*>(w + 2 * nb);
  void* sort_ws = w + 3 * nb;
  size_t sort_bytes = ws_bytes - 3 * nb;
  gw_obs_locate_kernel<<<(n + 127) / 128, 128, 0, st>>>(t, llh, n, slot, attr, key, val, status);
  int bits = 1;
  while ((1 << bits) < n_slots) ++bits;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(sort_ws, sort_bytes, key, key_sorted, val, perm, n, 0, bits, st);  // stable: ties stay in observation order
  if (e != cudaSuccess) return e;
  gw_csr_from_sorted_kernel<<<(n_slots + 1 + 255) / 256, 256, 0, st>>>(key_sorted, n, n_slots, ptr);
  count_launch(3);
  return cudaGetLastError();
}

// ---- CSR of a graph's edges by SOURCE (backward of the x[src] gathers: a per-source sum of edge gradients) ---------------------
__global__ void gw_iota_keys_kernel(const int32_t* __restrict__ src, int n, uint32_t* __restrict__ key, int32_t* __restrict__ val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (uint32_t)src[i], val[i] = i;
}
size_t sort_csr_workspace_bytes(int n) { return obs_graph_workspace_bytes(n); }
// perm[j] = edge ids ordered by src (ties in edge order), ptr[s] = first position of source s; ws >= sort_csr_workspace_bytes(n)
cudaError_t launch_sort_csr(const int32_t* src, int n, int n_slots, int32_t* perm, int32_t* ptr, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const size_t nb = (((size_t)n * 4 + 255) / 256) * 256;
  uint8_t* w = static_cast<uint8_t*>(ws);
  uint32_t* key = reinterpret_cast<uint32_t*>(w);
  uint32_t* key_sorted = reinterpret_cast<uint32_t*>(w + nb);
  int32_t* val = reinterpret_cast<int32_t*>(w + 2 * nb);
  void* sort_ws = w + 3 * nb;
  size_t sort_bytes = ws_bytes - 3 * nb;
  gw_iota_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, n, key, val);
  int bits = 1;
  while ((1 << bits) < n_slots) ++bits;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(sort_ws, sort_bytes, key, key_sorted, val, perm, n, 0, bits, st);
  if (e != cudaSuccess) return e;
  gw_csr_from_sorted_kernel<<<(n_slots + 1 + 255) / 256, 256, 0, st>>>(key_sorted, n, n_slots, ptr);
  count_launch(3);
  return cudaGetLastError();
}

}  // namespace gw
