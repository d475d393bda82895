// Wavelet-packet kernel instantiations for filter lengths 18 .. 32 (see afd_wpt_kernel.cuh).
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_group1, 18, false)
}  // namespace afd
