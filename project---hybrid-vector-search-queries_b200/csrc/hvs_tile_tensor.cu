// hvs_tile_tensor.cu -- K3: tcgen05 candidate pass for large shared slices (placeholder until the
// tensor-core kernel lands: the planner never schedules tensor items while this returns false).
#include "hvs_engine.h"

namespace hvs {

bool tensor_path_available() { return false; }

void build_bf16_image(hvs_engine *, int) {}

cudaError_t launch_tile_tensor(hvs_engine *e, const float *, const QSlice *, const TileItem *, uint32_t, uint32_t n_items,
                               const uint32_t *, uint64_t *, uint32_t *, uint32_t *)
{
    if (!n_items) return cudaSuccess;
    e->err = "tensor path not built";
    return cudaErrorNotSupported;
}

}  // namespace hvs
