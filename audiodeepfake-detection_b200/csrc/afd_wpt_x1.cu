// Wavelet-packet kernel instantiations with the extended epilogue (afd_wpt_forward_ex), filter lengths 18 .. 32.
#include "afd_wpt_kernel.cuh"

namespace afd {
AFD_WPT_GROUP(wpt_xgroup1, 18, true)
}  // namespace afd
