// gw_internal.h -- declarations shared by the translation units of libgwb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "gw_ops.h"

namespace gw {

void count_launch(int n = 1);
void set_error(const std::string& msg);

// exact-fp32 CUDA-core execution of one row op (gw_simt.cu)
cudaError_t launch_rowop_simt(const GemmOp& op, cudaStream_t stream);

cudaError_t launch_pad_rows(const float* src, int ld_src, int width, float* dst, int ld_dst, long long rows, float* amax, cudaStream_t stream);
cudaError_t launch_absmax_flat(const float* p, long long n, float* amax, cudaStream_t stream);
cudaError_t launch_csr_expand(const int32_t* ptr, int n, int32_t* dst, int* stats, cudaStream_t stream);
// backward primitives (gw_simt.cu)
cudaError_t launch_wgrad(const float* dY, int ldy, int N, const RowSrc& a, int K, int rows_per_sample, int batch, float* dW, int ldw, float* db,
                         cudaStream_t st);
cudaError_t launch_ln_bwd(const float* dy, int ld_dy, const float* z, int ld_z, int N, const float* gamma, long long R, float* dz, int ld_dz,
                          float* dgamma, float* dbeta, cudaStream_t st);
cudaError_t launch_batch_reduce(const float* in, int ld_in, long long rows, int width, int batch, float* out, int ld_out, bool accumulate,
                                cudaStream_t st);
cudaError_t launch_gather_rows(const float* in, int ld_in, int src_rows, const int32_t* idx, long long rows, int width, int batch, float* out,
                               int ld_out, bool accumulate, cudaStream_t st);
cudaError_t launch_strided_add(const float* src, int ld_src, float* dst, int ld_dst, long long rows, int width, cudaStream_t st);
cudaError_t launch_transpose(const float* W, int rows, int cols, float* WT, cudaStream_t st);
int seg_chunk_bound(int n_seg, int n_rows);
cudaError_t launch_seg_chunks(const int32_t* ptr, int n_seg, int32_t* chunk_seg, int32_t* chunk_j0, int32_t* seg_chunk0, cudaStream_t st);
cudaError_t launch_segsum_chunked(const float* base, int ld, const int32_t* ptr, const int32_t* perm, int src_rows, int rows, int batch,
                                  const int32_t* chunk_seg, const int32_t* chunk_j0, const int32_t* seg_chunk0, int max_chunks, float* partial,
                                  float* out, int ldo, cudaStream_t st);
cudaError_t launch_seg_carry(const float* carry, const int32_t* seg_dst, int rows, int seg_rows, int batch, float* out, int ldo,
                             cudaStream_t stream);
cudaError_t launch_segsum(const float* base, int ld, int width, const int32_t* ptr, const int32_t* perm, int src_rows,
                          int rows, int batch, float* out, int ldo, cudaStream_t stream);

// tcgen05 chain kernel (gw_tc3.cu)
cudaError_t launch_chain_tc3(const TcChain& ch, cudaStream_t stream);
bool tc3_chain_is_lean(const TcChain& ch);  // would the launch take the lean (perm32, 256-bit access) path?
// Packs W[n, k] (n < N_src rows of stride ldw, k < K_src) into the UMMA operand image the chain kernel streams with
// cp.async.bulk; `parts` = 2 (fp16 hi, lo) or 1 (bf16).  dst must hold tc_packed_bytes(K_src, N_src, parts, perm).
size_t tc_packed_bytes(int K_src, int N_src, int parts, int perm);
int tc_packed_rows(int N_src, int perm);  // rows of the packed image: N padded to 16 (perm16) or 64 (perm32)
cudaError_t launch_pack_weights(const float* W, int ldw, int K_src, int N_src, float wscale, int parts, int perm, void* dst,
                                cudaStream_t stream);  // perm: 1 = perm16 (general path), 2 = perm32 (lean path) feature order of gw_tc3.cu (gw_pack.cu)
cudaError_t launch_absmax(const float* W, int ldw, int K_src, int N_src, float* out_max, cudaStream_t stream);

// device-side observation graph of the assimilator (gw_graph.cu)
struct H3Tables {
  int res = -1, n_cells = 0, lat_n = 0;  // lattice table covers a, b in [-lat_n, lat_n]
  const double* frames = nullptr;        // [20][9]: face centre c, in-plane unit vectors ex, ey
  const int32_t* cell_of = nullptr;      // [20][(2 lat_n + 1)^2]: canonical cell of lattice point (a, b) on the face, -1 if none
  const int32_t* cell_slot = nullptr;    // [n_cells]: mesh-node slot of the cell in the encoder's numbering (H - 1 - rank)
  const double* cell_lat = nullptr;      // [n_cells] radians (as the host path sees them: through degrees and back)
  const double* cell_lng = nullptr;
  double scale = 0.0, cr = 1.0, sr = 0.0;  // plane -> lattice scale; Class III rotation (cos, sin), identity for even res
};

size_t obs_graph_workspace_bytes(int n);
size_t sort_csr_workspace_bytes(int n);
cudaError_t launch_sort_csr(const int32_t* src, int n, int n_slots, int32_t* perm, int32_t* ptr, void* ws, size_t ws_bytes, cudaStream_t st);
cudaError_t launch_obs_graph(const H3Tables& t, const float* llh, int n, int n_slots, int32_t* slot, int32_t* perm, int32_t* ptr, float* attr,
                             void* ws, size_t ws_bytes, int32_t* status, cudaStream_t st);

}  // namespace gw
