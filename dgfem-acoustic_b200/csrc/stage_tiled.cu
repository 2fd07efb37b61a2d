// Tiled DMMA stage kernel for the FP64-bound flagship orders (placeholder until the tuned kernel lands).
#include "dgb_internal.h"

namespace dgb {
StageKernel selectTiledKernel(int dim, int order) {
    (void)dim; (void)order;
    return StageKernel{};
}
}  // namespace dgb
