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

}  // namespace gw
