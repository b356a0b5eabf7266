// kernels_lse.cu — instantiations of the LexLSE kernel (one translation unit per kernel family: parallel compilation)
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lse_pair_kernel()
{
#ifdef LEXLS_B200_WITH_PAIR
    return (kernel_fn)lse_pair_kernel<0>;
#else
    return nullptr; // experiment not compiled into the product library (build with -DLEXLS_B200_WITH_PAIR)
#endif
}
kernel_fn pick_lse_lg_kernel(int W); // kernels_lg.cu
kernel_fn pick_lse_kernel(int W, bool LG)
{
    if (LG) return pick_lse_lg_kernel(W);
    switch (W)
    {
    case 1: return (kernel_fn)lse_kernel<1, false>;
    case 2: return (kernel_fn)lse_kernel<2, false>;
    case 4: return (kernel_fn)lse_kernel<4, false>;
    default: return (kernel_fn)lse_kernel<8, false>;
    }
}
} // namespace lexls_dev
