// Small ABI utilities of libocrf_raster.so.
#include "common.cuh"

extern "C" int ocrf_abi_version(void) { return OCRF_ABI_VERSION; }

extern "C" const char* ocrf_error_string(int code) {
  if (code == 0) return "success";
  if (code == OCRF_EINVAL) return "ocrf: invalid argument";
  if (code == OCRF_ECAPACITY) return "ocrf: workspace capacity exceeded";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "ocrf: unknown error";
}

extern "C" int ocrf_debug_sync(void* stream) {
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}
