// tcgen05 tensor-core CNN path (impl = 0) — placeholder interface; see nc_cnn_tc.cuh history.
#pragma once
#include "nc_common.cuh"

namespace nc {

struct TcModel { bool ready = false; };

inline int tc_model_prepare(cudaStream_t, TcModel& T, int, const float*, size_t, std::string*) { T.ready = false; return NC_OK; }
inline void tc_model_release(TcModel&) {}
inline int tc_forward(cudaStream_t, TcModel& T, int, const void*, int64_t, int64_t, const NcSiteMeta*, const float*, const float*,
                      const double*, const float*, float*, float*, int, uint64_t*, std::string*) {
    return NC_ESTATE;
}

}  // namespace nc
