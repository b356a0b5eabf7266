// kernels_lsi_wn.cu — multi-warp LexLSI kernels with the work matrix in shared memory (the global-memory variant: kernels_lg.cu)
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lsi_lg_kernel(int W); // kernels_lg.cu
kernel_fn pick_lsi_wn_kernel(int W, bool DS, bool LG)
{
    if (LG) return pick_lsi_lg_kernel(W);
    if (DS) switch (W)
        {
        case 2: return (kernel_fn)lsi_kernel<2, true, false>;
        case 4: return (kernel_fn)lsi_kernel<4, true, false>;
        default: return (kernel_fn)lsi_kernel<8, true, false>;
        }
    switch (W)
    {
    case 2: return (kernel_fn)lsi_kernel<2, false, false>;
    case 4: return (kernel_fn)lsi_kernel<4, false, false>;
    default: return (kernel_fn)lsi_kernel<8, false, false>;
    }
}
} // namespace lexls_dev
