// pw_extra.cu — entry points that are filled in incrementally (each returns PW_ERR_UNSUPPORTED until built)
#include "pw_engine.h"
using namespace pw;

extern "C" {
int pw_b200_filter(const PwPredicate*, int32_t, const struct ArrowArray* const*, const struct ArrowSchema* const*, size_t,
                   struct ArrowArray*, struct ArrowSchema*) { return fail(PW_ERR_UNSUPPORTED, "pw_b200_filter: not built yet"); }
int pw_b200_frame_filter_select(const PwPredicate*, int32_t, const PwFrame*, struct ArrowArray*, struct ArrowSchema*, int64_t*) {
  return fail(PW_ERR_UNSUPPORTED, "pw_b200_frame_filter_select: not built yet"); }
int pw_b200_frame_group_tuples(const PwFrame*, const int32_t*, int32_t, int32_t, struct ArrowArray*, struct ArrowArray*,
                               struct ArrowArray*, struct ArrowSchema*) { return fail(PW_ERR_UNSUPPORTED, "pw_b200_frame_group_tuples: not built yet"); }
uint32_t _polars_plugin_get_version(void) { return (0u << 16) | 1u; }
const char* _polars_plugin_get_last_error_message(void) { return pw_b200_last_error(); }
void _polars_plugin_filter_groupby_agg(const SeriesExport*, size_t, const uint8_t*, size_t, SeriesExport* rv, const CallerContext*) {
  if (rv) { rv->private_data = nullptr; rv->release = nullptr; }
  fail(PW_ERR_UNSUPPORTED, "plugin shim: not built yet");
}
void _polars_plugin_field_filter_groupby_agg(const struct ArrowSchema*, size_t, struct ArrowSchema* out, const uint8_t*, size_t) {
  if (out) out->release = nullptr;
}
}
