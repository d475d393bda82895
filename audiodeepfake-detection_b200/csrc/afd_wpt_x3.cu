// Wavelet-packet kernel instantiations with the extended epilogue (afd_wpt_forward_ex), filter lengths 50 .. 64.
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_xgroup3, 50, true)
}  // namespace afd
