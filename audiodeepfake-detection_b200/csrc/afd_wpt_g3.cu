// Wavelet-packet kernel instantiations for filter lengths 50 .. 64 (see afd_wpt_kernel.cuh).
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_group3, 50, false)
}  // namespace afd
