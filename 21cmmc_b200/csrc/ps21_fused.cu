// Fused two-pass path for power-of-two shapes (placeholder until the kernels land).
#include "ps21_internal.cuh"
#include "../../include/ps21.h"

namespace ps21 {

int fused_try_setup(ps21_plan*, ShapePlan&) {
    set_error("fused path not built");
    return -1;
}
void fused_destroy(ShapePlan&) {}
int fused_build_lut(ps21_plan*, int, const double*, const double*, const double*, const double*, cudaStream_t) {
    return PS21_ERR_INVALID;
}
int fused_forward(ps21_plan*, ShapePlan&, const float*, long long, long long, long long, cudaStream_t) {
    return PS21_ERR_INVALID;
}
long long fused_scratch_elems(const ShapePlan&, int, int) { return 0; }

}  // namespace ps21
