// placeholder — replaced by the scan-to-map implementation
#include "ll_ctx.h"
int ll_map_alloc(ll_ctx*) { return LL_OK; }
void ll_map_free(ll_ctx*) {}
int ll_launch_mapping(ll_ctx*, int) { return LL_OK; }
int ll_map_insert_impl(ll_ctx*, const float*, int, const float*, int) { return LL_E_INVAL; }
extern "C" int ll_mapping_step(ll_ctx*, ll_cloud_view, ll_cloud_view, const double*, const double*, double*, double*) { return LL_E_INVAL; }
extern "C" int ll_map_insert(ll_ctx*, ll_cloud_view, ll_cloud_view) { return LL_E_INVAL; }
