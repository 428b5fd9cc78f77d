// Single translation unit of libdiner_b200.so (keeps __constant__ data and inlined helpers in one module).
#include "scene.cu"
#include "sampler.cu"
#include "mlp_simt.cu"
#include "backward_simt.cu"
#include "composite.cu"
#include "mlp_tc.cu"
#include "mlp_tc2.cu"
#include "capi.cu"
