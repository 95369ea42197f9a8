// swd_osd.cuh — osd_window post-BP + bit-packed GF(2) OSD (placeholder until the kernels land).
#pragma once
#include "swd_kernels.cuh"
struct OsdSmem { int total; };
struct OsdWork { u8 *bp_dec = nullptr, *osd0 = nullptr, *osdw = nullptr; double *lpr = nullptr; int *bp_iter = nullptr; long long out_cap = 0; };
static inline size_t osd_bytes_per_shot(int m, int n) { return 0; }
static inline void osd_bind(OsdWork *, unsigned char *, long long, int, int) {}
static inline int osd_reserve_outputs(OsdWork *, long long, int) { return -2; }
static inline int osd_setup(int, int, int, int, int, int, int, OsdSmem *, int *, int *) { return -2; }
static inline int osd_launch(const GraphDev &, const u8 *, const Workspace &, const SubLayout &, const PathSmem &, const GdgDev &,
                             const OsdSmem &, const OsdWork &, int, int, int, size_t, int, int, int, int, int, u8 *, u8 *, double *,
                             long long, cudaStream_t, uint64_t *) { return -2; }
