// kernels_lsi_wn.cu — multi-warp LexLSI kernels (larger n) and the "work matrix in global memory" variant
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lsi_wn_kernel(int W, bool DS, bool LG)
{
    if (LG) return W == 8 ? (kernel_fn)lsi_kernel<8, false, true> : (kernel_fn)lsi_kernel<16, false, true>;
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
