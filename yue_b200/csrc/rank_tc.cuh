// K3 (tcgen05 flavour) -- placeholder until the tensor-core kernel lands; the exact kernel
// serves every request meanwhile (rank_tc_supported() == false).
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

namespace yue {
struct RankTcState { bool q_dirty = true; };
inline bool rank_tc_supported(int, int) { return false; }
inline void rank_tc_release(RankTcState&) {}
template <class Fallback>
inline int rank_tc_run(RankTcState&, cudaStream_t, int, const float*, const float*, int, int, int,
                       const int32_t*, int64_t, int, const int64_t*, const int32_t*, int32_t*, float*,
                       std::string& err, int64_t&, Fallback) {
    err = "tcgen05 ranking not built";
    return 6;
}
}  // namespace yue
