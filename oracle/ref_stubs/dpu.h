/* TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the UPMEM SDK's <dpu.h> so that the reference's host
 * translation units (which mix the scalar host oracles spmm_host_* with the
 * DPU transfer code) can be compiled in place from /root/reference by
 * oracle/build_ref.sh.  Only the host oracles and the partitioners are ever
 * called from the resulting objects; every dpu_* entry point below aborts.
 * Nothing here is derived from UPMEM sources: it declares just enough names
 * for the reference files to parse.
 */
#ifndef PYGIM_B200_STUB_DPU_H
#define PYGIM_B200_STUB_DPU_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

struct dpu_t;
struct dpu_set_t { struct dpu_t *dpu; int kind; };
typedef int dpu_error_t;
typedef int dpu_xfer_t;
typedef int dpu_xfer_flags_t;
typedef int dpu_launch_policy_t;

#define DPU_OK 0
#define DPU_XFER_TO_DPU 0
#define DPU_XFER_FROM_DPU 1
#define DPU_XFER_DEFAULT 0
#define DPU_XFER_ASYNC 2
#define DPU_SYNCHRONOUS 0
#define DPU_ASYNCHRONOUS 1
#define DPU_MRAM_HEAP_POINTER_NAME "__sys_used_mram_end"
#define DPU_ALLOCATE_ALL 0

#define DPU_ASSERT(stmt) do { if ((stmt) != DPU_OK) abort(); } while (0)
/* zero-trip loops: the stubbed runtime owns no DPUs */
#define PYGIM_STUB_PICK(_1, _2, _3, NAME, ...) NAME
#define PYGIM_STUB_FOREACH2(set, one) for ((one) = (set); 0;)
#define PYGIM_STUB_FOREACH3(set, one, i) for ((one) = (set), (i) = 0; 0;)
#define DPU_FOREACH(...) PYGIM_STUB_PICK(__VA_ARGS__, PYGIM_STUB_FOREACH3, PYGIM_STUB_FOREACH2, 0)(__VA_ARGS__)
#define DPU_RANK_FOREACH(...) PYGIM_STUB_PICK(__VA_ARGS__, PYGIM_STUB_FOREACH3, PYGIM_STUB_FOREACH2, 0)(__VA_ARGS__)

dpu_error_t dpu_alloc(uint32_t nr, const char *profile, struct dpu_set_t *set);
dpu_error_t dpu_alloc_ranks(uint32_t nr, const char *profile, struct dpu_set_t *set);
dpu_error_t dpu_free(struct dpu_set_t set);
dpu_error_t dpu_load(struct dpu_set_t set, const char *path, void *program);
dpu_error_t dpu_get_nr_dpus(struct dpu_set_t set, uint32_t *nr);
dpu_error_t dpu_get_nr_ranks(struct dpu_set_t set, uint32_t *nr);
dpu_error_t dpu_prepare_xfer(struct dpu_set_t set, void *buffer);
dpu_error_t dpu_push_xfer(struct dpu_set_t set, dpu_xfer_t xfer, const char *symbol, uint64_t offset,
                          uint64_t length, dpu_xfer_flags_t flags);
dpu_error_t dpu_broadcast_to(struct dpu_set_t set, const char *symbol, uint64_t offset, const void *src,
                             uint64_t length, dpu_xfer_flags_t flags);
dpu_error_t dpu_launch(struct dpu_set_t set, dpu_launch_policy_t policy);
dpu_error_t dpu_sync(struct dpu_set_t set);
dpu_error_t dpu_log_read(struct dpu_set_t set, FILE *stream);
dpu_error_t dpulog_read_for_dpu(struct dpu_t *dpu, FILE *stream);
#endif
