// kernels_lse_reg.cu — LexLSE kernels with the regularization routines compiled in (lexlse.h:274-411); the default kernels of
// kernels_lse.cu carry none of that code.  Regularized batches run on the shared-memory routines (two or more warps per problem).
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lse_reg_kernel(int W, bool LG)
{
    if (LG) return W == 8 ? (kernel_fn)lse_kernel<8, true, true> : (kernel_fn)lse_kernel<16, true, true>;
    switch (W)
    {
    case 2: return (kernel_fn)lse_kernel<2, false, true>;
    case 4: return (kernel_fn)lse_kernel<4, false, true>;
    default: return (kernel_fn)lse_kernel<8, false, true>;
    }
}
} // namespace lexls_dev
