// fp64 instantiation of the reference-order shallow-water kernel, in its own translation unit so
// that it can be compiled with -fmad=false (see swm_kernels.cuh and somax_b200/_lib.py).
#include "swm_kernels.cuh"

namespace sb {

int swm_launch_reference_order_f64(const SwmArgs<double>& A, const Stage<double>& st, dim3 grid, dim3 block,
                                   cudaStream_t s) {
  swm_rhs_kernel<double><<<grid, block, 0, s>>>(A, st);
  return 0;
}

}  // namespace sb
