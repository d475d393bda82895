// Wavelet-packet kernel instantiations for filter lengths 2 .. 16 (see afd_wpt_kernel.cuh).
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_group0, 2, false)
}  // namespace afd
