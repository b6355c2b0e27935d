// stubs.cu — entry points not implemented yet (replaced as the files land).
#include "api_internal.h"
namespace rnnwf {
#define NOTIMPL(name) do { set_error(name " is not implemented in this build"); return -2; } while (0)
template <typename T> size_t mdrnn_workspace_bytes_t(const rnnwf_model&, int, int64_t, int) { return 0; }
template size_t mdrnn_workspace_bytes_t<float>(const rnnwf_model&, int, int64_t, int);
template size_t mdrnn_workspace_bytes_t<double>(const rnnwf_model&, int, int64_t, int);
template <typename T> int mdrnn_sample_t(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t) { NOTIMPL("mdrnn_sample"); }
template int mdrnn_sample_t<float>(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t);
template int mdrnn_sample_t<double>(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t);
template <typename T> int mdrnn_logpsi_t(const rnnwf_model&, const void*, const uint8_t*, int64_t, double*, void*, size_t, cudaStream_t) { NOTIMPL("mdrnn_logpsi"); }
template int mdrnn_logpsi_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, double*, void*, size_t, cudaStream_t);
template int mdrnn_logpsi_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, double*, void*, size_t, cudaStream_t);
template <typename T> int mdrnn_tfim_eloc_t(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, double*, double*, void*, size_t, cudaStream_t) { NOTIMPL("mdrnn_tfim_eloc"); }
template int mdrnn_tfim_eloc_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, double*, double*, void*, size_t, cudaStream_t);
template int mdrnn_tfim_eloc_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, double*, double*, void*, size_t, cudaStream_t);
template <typename T> int mdrnn_vmc_grad_t(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double*, void*, size_t, cudaStream_t) { NOTIMPL("mdrnn_vmc_grad"); }
template int mdrnn_vmc_grad_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double*, void*, size_t, cudaStream_t);
template int mdrnn_vmc_grad_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double*, void*, size_t, cudaStream_t);
}  // namespace rnnwf
