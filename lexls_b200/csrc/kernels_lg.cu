// kernels_lg.cu — the kernels that keep the work matrix in global memory (L2) with a shared-memory view of the level being
// factorized.  Their own translation unit because they are compiled WITH the optimizer's run-time loop unrolling: their time
// goes into long dynamic loops (C4: 40 x 200 levels), while the shared-memory kernels are instruction-fetch bound and are
// compiled without it (build.py).
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lse_lg_kernel(int W) { return W == 8 ? (kernel_fn)lse_kernel<8, true> : (kernel_fn)lse_kernel<16, true>; }
kernel_fn pick_lsi_lg_kernel(int W) { return W == 8 ? (kernel_fn)lsi_kernel<8, false, true> : (kernel_fn)lsi_kernel<16, false, true>; }
} // namespace lexls_dev
