// Wavelet-packet kernel instantiations with the extended epilogue (afd_wpt_forward_ex), filter lengths 2 .. 16.
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_xgroup0, 2, true)
}  // namespace afd
