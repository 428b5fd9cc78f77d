#include "mlp_tc.h"
#include <string.h>
#include <stdio.h>
cudaError_t tc_pack_weights(TcState& t, const MlpDev& m, cudaStream_t st) {
    t.ready = false;
    snprintf(t.why, sizeof(t.why), "tcgen05 path not built yet");
    return cudaSuccess;
}
cudaError_t tc_query(TcState& t, const SceneDev& s, const MlpDev& m, const QueryArgs& q, bool parity,
                     int num_sms, cudaStream_t st) { return cudaErrorNotSupported; }
void tc_release(TcState& t) {}
