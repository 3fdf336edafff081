// ps_broadphase.cu — sort-based uniform-grid broad phase for one large world (BASELINE config 5).
// (first-light stub: the entry point exists and fails loudly; the kernels land later in this round)
#include <string>
#include "../../include/ps_b200.h"

extern "C" int ps_broadphase_pairs(int32_t n_bodies, const int32_t* shape, const double* dims, const double* mass,
                                   const double* pos, const double* rot, int32_t aabb_mode, int64_t* n_pairs,
                                   int32_t* pair_ids, int64_t pair_capacity, double* kernel_ms)
{
    (void)n_bodies; (void)shape; (void)dims; (void)mass; (void)pos; (void)rot; (void)aabb_mode; (void)n_pairs;
    (void)pair_ids; (void)pair_capacity; (void)kernel_ms;
    return PS_ERR_INVALID_OPERATION;
}
