/* TEST INFRASTRUCTURE ONLY: aborting definitions for the stubbed UPMEM
 * runtime declared in dpu.h (ctypes dlopens with RTLD_NOW, so every symbol
 * must resolve).  A call into any of these means a test reached DPU code. */
#include "dpu.h"
#define STUB(sig) dpu_error_t sig { fprintf(stderr, "pygim_b200 oracle/_ref: UPMEM runtime is stubbed\n"); abort(); return 1; }
STUB(dpu_alloc(uint32_t nr, const char *profile, struct dpu_set_t *set))
STUB(dpu_alloc_ranks(uint32_t nr, const char *profile, struct dpu_set_t *set))
STUB(dpu_free(struct dpu_set_t set))
STUB(dpu_load(struct dpu_set_t set, const char *path, void *program))
STUB(dpu_get_nr_dpus(struct dpu_set_t set, uint32_t *nr))
STUB(dpu_get_nr_ranks(struct dpu_set_t set, uint32_t *nr))
STUB(dpu_prepare_xfer(struct dpu_set_t set, void *buffer))
STUB(dpu_push_xfer(struct dpu_set_t set, dpu_xfer_t xfer, const char *symbol, uint64_t offset, uint64_t length, dpu_xfer_flags_t flags))
STUB(dpu_broadcast_to(struct dpu_set_t set, const char *symbol, uint64_t offset, const void *src, uint64_t length, dpu_xfer_flags_t flags))
STUB(dpu_launch(struct dpu_set_t set, dpu_launch_policy_t policy))
STUB(dpu_sync(struct dpu_set_t set))
STUB(dpu_log_read(struct dpu_set_t set, FILE *stream))
STUB(dpulog_read_for_dpu(struct dpu_t *dpu, FILE *stream))
