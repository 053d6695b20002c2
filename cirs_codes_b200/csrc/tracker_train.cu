// K6 placeholder (replaced by the real training pass in the next commit).
#include "common.cuh"
#include "../../include/cirs_b200.h"
extern "C" int64_t cirs_tracker_train_workspace_bytes(const cirs_tracker_weights*, int32_t, int32_t) { return 256; }
extern "C" int cirs_tracker_train(const cirs_tracker_weights*, const cirs_tracker_weights*, int32_t, int32_t,
                                  const int32_t*, const int32_t*, const float*, const int32_t*, const float*,
                                  const float*, const float*, float*, void*, int64_t, void*) {
  cirs_set_error("cirs_tracker_train: not built yet");
  return CIRS_ERR_ARG;
}
