// placeholder until the tcgen05 path lands
#include "conv.h"
namespace stp {
bool tc_conv_supported(const ConvP&) { return false; }
int launch_tc_conv(const ConvP&, cudaStream_t) { set_error("tc conv not built"); return STP_E_UNSUPPORTED; }
bool tc_wgrad_supported(const WgradP&) { return false; }
int launch_tc_wgrad(const WgradP&, float*, void*, size_t, cudaStream_t) { set_error("tc wgrad not built"); return STP_E_UNSUPPORTED; }
size_t tc_wgrad_workspace(int64_t, int, int, int, int) { return 0; }
}  // namespace stp
