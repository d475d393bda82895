// Wavelet-packet kernel instantiations with the extended epilogue (afd_wpt_forward_ex), filter lengths 34 .. 48.
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_xgroup2, 34, true)
}  // namespace afd
