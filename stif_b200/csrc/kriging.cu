// TEMPORARY stub -- replaced by the real implementation.
#include "common.cuh"
using namespace stif;
extern "C" {
int stif_kriging_set_obs(stif_ctx *, const double *, const double *, const double *, const double *, int64_t, int) { set_error("kriging not built yet"); return STIF_ESTATE; }
int stif_kriging_set_model(stif_ctx *, int, int, int, int, const double *, int) { set_error("kriging not built yet"); return STIF_ESTATE; }
int stif_variogram_model_eval(stif_ctx *, const double *, const double *, int64_t, double *) { set_error("kriging not built yet"); return STIF_ESTATE; }
int stif_kriging_predict_host(stif_ctx *, const double *, const double *, const double *, const int64_t *, int64_t, int, int, double, double, int, double *, double *, float *, float *, uint32_t *, int32_t *) { set_error("kriging not built yet"); return STIF_ESTATE; }
int stif_kriging_predict_dev(stif_ctx *, const double *, const double *, const double *, const int64_t *, int64_t, int, int, double, double, int, double *, double *, float *, float *, uint32_t *, int32_t *, void *) { set_error("kriging not built yet"); return STIF_ESTATE; }
int stif_kriging_stats(stif_ctx *, uint64_t *) { set_error("kriging not built yet"); return STIF_ESTATE; }
}
