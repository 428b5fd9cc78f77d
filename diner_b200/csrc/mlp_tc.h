// tcgen05 (5th-gen tensor core) path of the per-sample network: packed weights + launch entry points.
#pragma once
#include <cuda_runtime.h>
#include "diner_internal.h"

struct TcState {
    bool ready = false;
    char why[160] = "diner_set_mlp not called";
    void* wpack = nullptr;        // bf16 hi/lo weight tiles in UMMA K-major SWIZZLE_128B layout
    size_t wpack_bytes = 0;
    float* bias = nullptr;        // per-phase cumulative bias vectors
    void* scratch = nullptr;      // combined activations x_c between the pre- and post-combine kernels
    size_t scratch_bytes = 0;
    int* err_flag = nullptr;      // device-side watchdog flag (mbarrier timeouts)
};

cudaError_t tc_pack_weights(TcState& t, const MlpDev& m, cudaStream_t st);
cudaError_t tc_query(TcState& t, const SceneDev& s, const MlpDev& m, const QueryArgs& q, bool parity,
                     int num_sms, cudaStream_t st);
void tc_release(TcState& t);
