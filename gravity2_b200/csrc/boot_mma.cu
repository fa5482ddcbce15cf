// Batched bootstrap on the 5th-generation tensor cores (tcgen05): placeholder until the kernel lands.
#include "common.cuh"

int boot_mma_available() { return 0; }
int boot_mma_launch(gv2_ctx* ctx, const float*, int64_t, int64_t, int64_t, const float*, int64_t, double*) {
    return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "tcgen05 bootstrap kernel not built");
}
