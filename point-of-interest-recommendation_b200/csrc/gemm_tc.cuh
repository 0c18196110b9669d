// gemm_tc.cuh -- tcgen05 (5th-gen tensor core) GEMM path, gemm modes 1 (3xTF32) and 2 (1xTF32).
// Placeholder until the UMMA kernel lands: reports "unsupported" so gemm_tn() falls back to the
// fp32 FMA kernel in gemm_simt.cuh.
#pragma once
#include "common.cuh"

static inline bool tc_gemm_supported(int64_t, int, int, int, int) { return false; }

template <class Epi>
static int launch_gemm_tn_tc(poi_engine* e, const float*, int, const float*, int, int64_t, int, int,
                             const Epi&, bool) {
    POI_FAIL(e, "tcgen05 GEMM path not built");
}
