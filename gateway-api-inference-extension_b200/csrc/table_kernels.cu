// table_kernels.cu — scatter of the host mirror's dirty slots / row words into the device prefix table.
#include "kernels.cuh"

namespace eppscore {

// ---------------------------------------------------------------------------------------------
// prefix-table maintenance: the host mirror is authoritative; these scatter its dirty words/slots.
// ---------------------------------------------------------------------------------------------
__global__ void scatter_u32_kernel(uint32_t* dst, const uint32_t* idx, const uint32_t* val, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}
__global__ void scatter_slots_kernel(Slot* dst, const uint32_t* idx, const Slot* val, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}
int launch_scatter_u32(uint32_t* dst, const uint32_t* idx, const uint32_t* val, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  scatter_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, idx, val, n);
  return 1;
}
int launch_scatter_slots(Slot* dst, const uint32_t* idx, const Slot* val, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  scatter_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, idx, val, n);
  return 1;
}

}  // namespace eppscore
