#include "../../salmon_b200/csrc/common.cuh"
#include <stdarg.h>
namespace sb { static thread_local char g_err[512]; void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);} }
extern "C" const char* sb_last_error(void) { return sb::g_err; }
// stubs for the device-side entry points pipeline.cu references
extern "C" void sb_map_default_params(sb_map_params*) {}
extern "C" void sb_em_default_params(sb_em_params*) {}
extern "C" int sb_index_get_meta(const sb_index*, uint32_t*, uint32_t*, uint32_t*, const char* const**, const uint32_t**) { return -1; }
extern "C" sb_map_ctx* sb_map_create(sb_index*, const sb_map_params*, int, uint32_t, uint32_t) { return nullptr; }
extern "C" void sb_map_destroy(sb_map_ctx*) {}
extern "C" int sb_map_batch(sb_map_ctx*, const uint8_t*, const uint8_t*, uint32_t, uint32_t, sb_map_batch_stats*) { return -1; }
extern "C" int sb_map_finish(sb_map_ctx*, sb_map_result*) { return -1; }
extern "C" sb_em_ctx* sb_em_create(int) { return nullptr; }
extern "C" void sb_em_destroy(sb_em_ctx*) {}
extern "C" int sb_em_optimize(sb_em_ctx*, const sb_eq_csr*, const sb_em_params*, const double*, const double*, const uint64_t*, double*, sb_em_stats*) { return -1; }
extern "C" int sb_bootstrap(sb_em_ctx*, const sb_em_params*, double, uint32_t, uint64_t, sb_sample_cb, void*) { return -1; }
extern "C" int sb_gibbs(sb_em_ctx*, const double*, int, int, double, uint32_t, uint32_t, int, double, uint64_t, sb_sample_cb, void*) { return -1; }
extern "C" int sb_write_quant_sf(const char*, uint32_t, const char* const*, const uint32_t*, const double*, const double*, double, int) { return -1; }
extern "C" int sb_write_eq_classes(const char*, uint32_t, const char* const*, uint64_t, const uint64_t*, const uint32_t*, const double*, const uint64_t*) { return -1; }
extern "C" int sb_index_host_arrays(const sb_index*, const uint64_t**, const uint8_t**, const void**, uint64_t*, const void**, uint64_t*) { return -1; }
// (round 2 entry points)
extern "C" int sb_version(void) { return 0; }
extern "C" int sb_map_partial_get(sb_map_ctx*, sb_map_partial*) { return -1; }
extern "C" int sb_map_project_global(sb_map_ctx*, const sb_map_partial*, uint32_t, const uint32_t*, sb_map_result*) { return -1; }
extern "C" int sb_map_online_state(sb_map_ctx*, double*, double*, double*, uint64_t*) { return -1; }
extern "C" sb_comm* sb_comm_create(int, int, const void*, int) { return nullptr; }
extern "C" void sb_comm_destroy(sb_comm*) {}
extern "C" int sb_comm_rank(const sb_comm*) { return 0; }
extern "C" int sb_comm_size(const sb_comm*) { return 1; }
extern "C" int sb_comm_allreduce(sb_comm*, void*, size_t, int, int) { return -1; }
extern "C" int sb_comm_allgather(sb_comm*, const void*, void*, size_t) { return -1; }
extern "C" int sb_em_peer_setup(sb_em_ctx*, sb_comm*, uint32_t) { return -1; }
extern "C" int sb_em_set_option(sb_em_ctx*, const char*, int64_t) { return -1; }
