// Internal interface of fcfv_prep.cu (space-filling-curve element order); see that file.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace fcfv {
cudaError_t hilbert_element_order(cudaStream_t st, int n, const double* d_x, const double* d_y, long long* d_old_of_new,
                                  void* tmp, size_t tmp_bytes, size_t* need);
}
