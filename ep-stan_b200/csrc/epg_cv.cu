#include "epg_internal.h"
extern "C" int epg_cv_moments(epg_ctx* c, int, int, int, const double*, const double*, const double*, const double*, int, double, double, double, double*, double*, int32_t*) { return epg_fail_msg(c, "not implemented"); }
