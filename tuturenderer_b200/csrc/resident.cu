// resident.cu — translation unit of the register-resident path-tracing kernel (resident.cuh).  Kept apart
// from tutu_b200.cu: ptxas 12.9 crashes on a unit in which two kernels inline shade_vertex.
#define TUTU_RESIDENT_IMPL
#include "tutu_internal.hpp"
#include "resident.cuh"

namespace tutu {

cudaError_t pt_resident_grid(int sm_count, int* grid) {
  int per_sm = 0;
  const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pt_resident, TUTU_RESIDENT_BLOCK, 0);
  *grid = sm_count * (per_sm < 1 ? 1 : per_sm);
  return e;
}

cudaError_t pt_resident_launch(int grid, cudaStream_t s, const DevScene& sc, const SmallScene& ss,
                               const ResidentArgs& a) {
  pt_resident<<<grid, TUTU_RESIDENT_BLOCK, 0, s>>>(sc, ss, a);
  return cudaGetLastError();
}

}  // namespace tutu
