// placeholder until K4 lands
#include "kernels.hpp"
namespace iqgpu {
bool fftfilt_supported(unsigned) { return false; }
cudaError_t launch_fftfilt(const float2*, size_t, unsigned, const float2*, const float2*, float2*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_fft_forward(const float2*, unsigned, const float2*, float2*, cudaStream_t) { return cudaErrorNotSupported; }
}
