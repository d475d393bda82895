// Wavelet-packet kernel instantiations for filter lengths 34 .. 48 (see afd_wpt_kernel.cuh).
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_group2, 34, false)
}  // namespace afd
