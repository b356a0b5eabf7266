// kernels_lsi_reg.cu — LexLSI kernels whose inner LexLSE is regularized (lexlsi.h:331-333, :497; objective.h:489-493)
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lsi_reg_kernel(int W, bool LG)
{
    if (LG) return W == 8 ? (kernel_fn)lsi_kernel<8, false, true, true> : (kernel_fn)lsi_kernel<16, false, true, true>;
    switch (W)
    {
    case 2: return (kernel_fn)lsi_kernel<2, false, false, true>;
    case 4: return (kernel_fn)lsi_kernel<4, false, false, true>;
    default: return (kernel_fn)lsi_kernel<8, false, false, true>;
    }
}
} // namespace lexls_dev
