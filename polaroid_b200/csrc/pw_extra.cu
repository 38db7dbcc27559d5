// pw_extra.cu — entry points that are filled in incrementally (each returns PW_ERR_UNSUPPORTED until built)
#include "pw_engine.h"
using namespace pw;

extern "C" {
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
