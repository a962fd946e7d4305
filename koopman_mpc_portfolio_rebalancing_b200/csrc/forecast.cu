// placeholder until the forecast path lands (next commit)
#include "../../include/kmpc.h"
extern "C" {
int kmpc_model_load(kmpc_handle*, const kmpc_model_desc*, kmpc_model**) { return KMPC_E_UNSUPPORTED; }
int kmpc_model_free(kmpc_model*) { return KMPC_OK; }
int kmpc_forecast(kmpc_handle*, const kmpc_model*, const float*, int, const double*, const double*, int, int, int, int, int, int, int, float*, void*) { return KMPC_E_UNSUPPORTED; }
int kmpc_encode(kmpc_handle*, const kmpc_model*, const float*, int, float*, void*) { return KMPC_E_UNSUPPORTED; }
int kmpc_rollout(kmpc_handle*, const kmpc_model*, const float*, int, int, int, float*, void*) { return KMPC_E_UNSUPPORTED; }
}
