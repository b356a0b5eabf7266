// kernels_lsi_w1.cu — warp-per-problem LexLSI kernels (register-panel l-QR)
#include "lexls_device.cuh"
namespace lexls_dev
{
kernel_fn pick_lsi_w1_kernel(bool DS)
{
    return DS ? (kernel_fn)lsi_kernel<1, true, false> : (kernel_fn)lsi_kernel<1, false, false>;
}
} // namespace lexls_dev
