// placeholder translation unit: solver entry points are filled in incrementally
#include "ms_common.cuh"
using namespace ms;
extern "C" {
#define NOT_YET(name) return fail(MS_ERR_STATE, name ": not implemented yet")
int ms_solver_create(const ms_state*, uint32_t, ms_solver**) { NOT_YET("ms_solver_create"); }
void ms_solver_destroy(ms_solver*) {}
int ms_solver_reset(ms_solver*, void*) { NOT_YET("ms_solver_reset"); }
int ms_solver_counts(const ms_solver*, int32_t*, int32_t*, int32_t*) { NOT_YET("ms_solver_counts"); }
int ms_solver_export_tree(const ms_solver*, ms_state*, int32_t*, uint8_t*, int32_t*, int32_t*, uint8_t*) { NOT_YET("ms_solver_export_tree"); }
int ms_solver_export_table(const ms_solver*, uint64_t*, uint8_t*, uint8_t*, double*, double*, uint8_t*, void*) { NOT_YET("ms_solver_export_table"); }
int ms_solver_import_table(ms_solver*, const double*, const double*, void*) { NOT_YET("ms_solver_import_table"); }
int ms_solver_device_ptrs(ms_solver*, double**, double**, double**, double**, size_t*) { NOT_YET("ms_solver_device_ptrs"); }
int ms_cfr_iterate(ms_solver*, int32_t, void*) { NOT_YET("ms_cfr_iterate"); }
int ms_cfr_traverse(ms_solver*, int32_t, double, double, double*, void*) { NOT_YET("ms_cfr_traverse"); }
int ms_mccfr_inplace(ms_solver*, int64_t, uint64_t, uint64_t, void*) { NOT_YET("ms_mccfr_inplace"); }
int ms_mccfr_batch(ms_solver*, int32_t, int64_t, uint64_t, uint64_t, void*) { NOT_YET("ms_mccfr_batch"); }
int ms_mccfr_apply(ms_solver*, void*) { NOT_YET("ms_mccfr_apply"); }
int ms_solver_counters(ms_solver*, uint64_t*, int, void*) { NOT_YET("ms_solver_counters"); }
int ms_best_response(ms_solver*, int32_t, double*, void*) { NOT_YET("ms_best_response"); }
}
