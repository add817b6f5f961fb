#include "epg_internal.h"
void epg_sites_free(epg_ctx*) {}
extern "C" {
int epg_cv_moments(epg_ctx* c, int, int, int, const double*, const double*, const double*, const double*, int, double, double, double, double*, double*, int32_t*) { return epg_fail_msg(c, "not implemented"); }
int epg_upload_sites(epg_ctx* c, int, int, const int64_t*, const double*, const int64_t*, const int32_t*, const int32_t*) { return epg_fail_msg(c, "not implemented"); }
int epg_tilted_sample(epg_ctx* c, int, int, const uint32_t*, const epg_sampler_opts*, double*, double*, int64_t*, double*) { return epg_fail_msg(c, "not implemented"); }
int epg_num_params(epg_ctx* c, int) { return -1; }
int epg_logdensity(epg_ctx* c, int, int, const double*, double*, double*) { return epg_fail_msg(c, "not implemented"); }
}
