/*
 * acg_b200_probe.h -- profiling / hardware-behaviour probes.  NOT part of the product ABI: these symbols exist only in
 * libacg_b200_probe.so (the same sources compiled with -DACG_PROBES), which scripts/ and the probe tests load
 * explicitly (action_conditioned_gans_b200._lib.use_probe_library()).  In that build the environment variable
 * ACG_DBG_SKIP switches pipeline stages of the conv kernels off (timing experiments; results are then garbage).
 */
#ifndef ACG_B200_PROBE_H_
#define ACG_B200_PROBE_H_
#include "acg_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Probe (tests only): D[128][N] = A_window * B^T where A_window's logical row m is shared-memory row
 * shift + (m/8)*pitch + (m%8) of a 128-byte-swizzled K-major [n_rows][64] bf16 tile (start not 1024 B aligned, 8-row
 * groups spaced by `pitch` rows).  base_offset_mode 1 sets the descriptor's base-offset field to (addr>>7)&7. */
/* Probe (profiling experiments, ACG_DBG_SKIP=8): out[8] = {setup ns, main-loop ns, epilogue ns, CTAs, K blocks,
 * MMA-thread wait for halo ns, MMA-thread wait for weights ns, 0} of the conv_tc kernels summed over CTAs since the
 * previous call (synchronises the device). */
int acg_debug_phase_times(unsigned long long* out8);
int acg_debug_umma_shift(const void* a_rows, int n_rows, const void* b_rows, int N, int shift, int pitch,
                         int base_offset_mode, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACG_B200_PROBE_H_ */
