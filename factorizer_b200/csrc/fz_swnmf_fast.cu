// Specialised fused kernels (placeholder until the TMA/register path lands).
#include "fz_internal.cuh"

namespace fz {
bool fast_supported(const DevGeom&, const fz_solver&) { return false; }
size_t fast_saved_bytes(const DevGeom&, const fz_solver&) { return 0; }
size_t fast_workspace_bytes(const DevGeom&, const fz_solver&) { return 0; }
int fast_forward(const float*, const float*, const float*, float*, void*, void*, const DevGeom&,
                 const fz_solver&, int, cudaStream_t) {
    return fail(FZ_ERR_UNSUPPORTED, "fast path not built");
}
int fast_backward(const float*, const float*, const float*, const float*, const void*, float*, void*,
                  const DevGeom&, const fz_solver&, int, int, cudaStream_t) {
    return fail(FZ_ERR_UNSUPPORTED, "fast path not built");
}
}  // namespace fz
