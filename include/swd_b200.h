/*
 * swd_b200.h — C-ABI of the B200-native batched window decoder (libswd_b200.so).
 *
 * Drop-in boundary for the hot path of gongaa/SlidingWindowDecoder: one handle
 * replaces one Cython decoder object of the reference and decodes a BATCH of
 * syndromes of the same window on one B200.  Plain pointers and sizes only; no
 * torch / numpy types.  Every entry point returns an int status (0 = OK, <0 =
 * error, see swd_strerror) and never aborts the process (the reference's C
 * layer may exit(1)/abort(), mod2sparse.c:94-97, mod2sparse_extra.cpp:139).
 *
 * Reference interfaces replaced (file:line under the reference's src/):
 *   swd_create            <- bp_history_decoder.__cinit__  bp_guessing_decoder.pyx:6-46
 *                            bpgdg_decoder.__cinit__       bp_guessing_decoder.pyx:161-219
 *                            bpgd_decoder.__cinit__        bp_guessing_decoder.pyx:474-499
 *                            osd_window.__cinit__          osd_window.pyx:8-126
 *                            (numpy2mod2sparse / spmatrix2mod2sparse, mod2sparse.pyx:6-32)
 *   swd_decode_batch_*    <- bpgdg_decoder.decode          bp_guessing_decoder.pyx:221-236
 *                            bpgd_decoder.decode           bp_guessing_decoder.pyx:501-514
 *                            osd_window.decode             osd_window.pyx:158-199
 *                            which in turn replace BPGD_main_thread::do_work (bpgd.cpp:591-688),
 *                            BPGD::{reset,min_sum_log,select_vn,peel,vn_set_value,get_pm}
 *                            (bpgd.cpp:13-351), index_sort (bpgd.cpp:384-389),
 *                            mod2sparse_decomp_osd + LU_forward_backward_solve
 *                            (mod2sparse_extra.cpp:78-376)
 *   swd_destroy           <- __dealloc__                   bp_guessing_decoder.pyx:142-154,444-469
 *   swd_window_*          <- the per-window bookkeeping of guessing.py:141-227 / osd.py:134-179
 *                            (commit first F rounds, syndrome update, flagged / logical counters)
 */
#ifndef SWD_B200_H
#define SWD_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWD_OK               0
#define SWD_ERR_INVALID     -1   /* bad argument (NULL, negative size, length mismatch)            */
#define SWD_ERR_UNSUPPORTED -2   /* graph exceeds a kernel limit (see DESIGN.md "limits")           */
#define SWD_ERR_CUDA        -3   /* CUDA runtime error; text via swd_last_error()                   */
#define SWD_ERR_NOMEM       -4

/* decoder kinds == the reference's three decoder classes */
#define SWD_KIND_BPGDG       0   /* bpgdg_decoder  (GDG)                    */
#define SWD_KIND_BPGD        1   /* bpgd_decoder   (plain guided decimation) */
#define SWD_KIND_OSD_WINDOW  2   /* osd_window     (BP + shortening + OSD)   */

#define SWD_OSD_0  0
#define SWD_OSD_E  1
#define SWD_OSD_CS 2

/* kwargs of the reference constructors, same meaning and defaults
 * (bp_guessing_decoder.pyx:7-9,162-171,475-478; osd_window.pyx:10-16). */
typedef struct swd_config {
    int    kind;
    int    device;                /* CUDA device ordinal                                         */
    /* BP on the full window ("max_iter" / "pre_max_iter") */
    int    max_iter;
    double ms_scaling_factor;
    /* GDG / GD */
    int    max_iter_per_step;
    int    max_step;
    int    max_tree_depth;
    int    max_side_depth;
    int    max_tree_branch_step;
    int    max_side_branch_step;
    double gdg_factor;            /* gdg_factor, or gd_factor for SWD_KIND_BPGD                   */
    int    new_n;                 /* <= 0 : min(n, 2m)                                            */
    int    multi_thread;          /* 1: bpgd.cpp branch tree; 0: single-thread schedule (pyx:254) */
    int    low_error_mode;
    /* osd_window */
    int    post_max_iter;
    int    osd_method;            /* SWD_OSD_*                                                    */
    int    osd_order;
    /* BP flavour of the full-window BP: SWD_BP_MIN_SUM (the reference's own decoders) or SWD_BP_PRODUCT_SUM (what the
     * drivers can ask of ldpc.BpOsdDecoder via bp_method="product_sum", osd.py:142-150; third-party, parity unpinned). */
    int    bp_method;
} swd_config;
#define SWD_BP_MIN_SUM      0
#define SWD_BP_PRODUCT_SUM  1

typedef struct swd_decoder swd_decoder;

/* Work counters accumulated since creation / last reset (for roofline accounting). */
typedef struct swd_counters {
    uint64_t shots;               /* syndromes decoded                                           */
    uint64_t pre_bp_edge_iters;   /* edges x iterations executed by the full-window BP kernel    */
    uint64_t path_edge_iters;     /* active edges x iterations executed inside GDG/GD/post-BP    */
    uint64_t gdg_shots;           /* shots that went past the pre-BP                             */
    uint64_t osd_shots;           /* shots that reached OSD                                      */
    uint64_t kernel_launches;     /* kernels launched by this library                            */
    uint64_t paths_run;           /* branch paths that executed at least one BP call             */
    uint64_t bp_calls;            /* min_sum_log-equivalent calls over all branch paths          */
    uint64_t path_vn_iters;       /* active variable nodes x iterations inside GDG/GD/post-BP    */
    uint64_t path_cn_iters;       /* active check nodes x iterations inside GDG/GD/post-BP       */
    uint64_t path_slot_iters;     /* message slots (live, dead, pad) scanned by those check passes */
    uint64_t osd_cols_scanned;    /* columns gathered against T by osd_kernel's elimination scan */
    uint64_t osd_pivots;          /* pivots found (= rank x OSD shots)                           */
} swd_counters;

/* Kernel classes for swd_get_kernel_times */
#define SWD_K_PRE_BP     0
#define SWD_K_SORT_RESET 1
#define SWD_K_PATH_MAIN  2   /* path_kernel phase 0: main + tree branches (or the single GD / post-BP path) */
#define SWD_K_PATH_SIDE  3   /* path_kernel phase 1: side branches                                         */
#define SWD_K_SELECT     4
#define SWD_K_OSD        5   /* osd_window: osd_kernel + finish / path-metric kernels */
#define SWD_K_PATH_TRUNK 6   /* path_kernel on shared-prefix nodes (decimation depths < max_tree_depth)                 */
#define SWD_K_POST_BP    7   /* osd_window: masked min-sum on the shortened graph (post_bp_kernel)                               */
#define SWD_K_COUNT      8

/* pcm as CSC: colptr[n+1], rowidx[nnz] (any order inside a column; sorted internally, as
 * mod2sparse_insert keeps them).  channel_llr[n] = log((1-p)/p) computed by the caller with libm
 * (bp_guessing_decoder.pyx:46).  All arrays are HOST pointers and are copied. */
int  swd_create(const swd_config *cfg, int m, int n, const int32_t *colptr, const int32_t *rowidx,
                const double *channel_llr, swd_decoder **out);
void swd_destroy(swd_decoder *d);

/* Decode B syndromes. synd[B*m], corr[B*n] one byte per bit (0/1), converge[B];
 * min_pm[B] optional (NULL ok).  _host: host pointers, H2D/D2H copies inside the call,
 * synchronous.  _device: device pointers, asynchronous on `stream` (a cudaStream_t). */
int  swd_decode_batch_host(swd_decoder *d, const uint8_t *synd, int64_t B,
                           uint8_t *corr, uint8_t *converge, double *min_pm);
int  swd_decode_batch_device(swd_decoder *d, const uint8_t *d_synd, int64_t B,
                             uint8_t *d_corr, uint8_t *d_converge, double *d_min_pm, void *stream);

/* Bit-packed shot I/O (8x fewer bytes over PCIe / HBM at the boundary): row b of a packed array holds ceil(nbits / 64)
 * uint64 words, bit j of the row = (word[j >> 6] >> (j & 63)) & 1  (numpy: packbits(bitorder="little").view(uint64)).
 * synd_packed [B, ceil(m/64)], corr_packed [B, ceil(n/64)]; converge / min_pm as above.  Same decode(syndrome) semantics
 * as the byte entry points (bp_guessing_decoder.pyx:221-252, osd_window.pyx:158-199).  The _device variant unpacks into /
 * packs from the decoder's own staging buffers on `stream`: calls on one decoder must be stream-ordered. */
int  swd_decode_batch_host_packed(swd_decoder *d, const uint64_t *synd_packed, int64_t B,
                                  uint64_t *corr_packed, uint8_t *converge, double *min_pm);
int  swd_decode_batch_device_packed(swd_decoder *d, const uint64_t *d_synd_packed, int64_t B,
                                    uint64_t *d_corr_packed, uint8_t *d_converge, double *d_min_pm, void *stream);
/* the two conversions on their own (device pointers), for callers that keep shot data packed (window driver, samplers) */
int  swd_pack_bits(int device, const uint8_t *d_bytes, int64_t B, int nbits, uint64_t *d_packed, void *stream);
int  swd_unpack_bits(int device, const uint64_t *d_packed, int64_t B, int nbits, uint8_t *d_bytes, void *stream);
/* 1 if the full-window BP of this decoder streams its messages from HBM (graph beyond one SM's shared memory), else 0 */
int  swd_is_streamed(swd_decoder *d);
/* Host-only diagnostic (no CUDA call): the shared-memory message layout swd_create chooses for the full-window BP kernel of
 * this graph (static per decoder: row starts on distinct banks, slot order inside a row chosen so that both passes of
 * bp_decode_llr, bp_guessing_decoder.pyx:62-127, are bank-conflict free).  slot_of_entry [nnz] (per CSC entry, rows ascending
 * inside a column), row_start [m + 1], stats[3] = {slots incl. pads, extra variable-pass wavefronts of the plain CSR order,
 * of the chosen layout}; any pointer may be NULL.  optimize = 0: the plain CSR order. */
int  swd_pre_bp_layout(int m, int n, const int32_t *colptr, const int32_t *rowidx, int optimize,
                       int32_t *slot_of_entry, int32_t *row_start, int64_t *stats);

/* osd_window read-only properties of the LAST batch (osd_window.pyx:487-517), host copies.
 * Any pointer may be NULL.  bp_dec/osd0/osdw: [B*n]; log_prob_ratios: [B*n*4]; bp_iteration: [B]. */
int  swd_osd_last_outputs(swd_decoder *d, int64_t B, uint8_t *bp_dec, uint8_t *osd0, uint8_t *osdw,
                          double *log_prob_ratios, int32_t *bp_iteration);

/* Per-kernel device timing.  When enabled, every kernel launch is bracketed by a pair of CUDA events
 * on the launching stream; swd_get_kernel_times synchronises, sums the elapsed times per kernel class
 * into ms[SWD_K_COUNT] / launches[SWD_K_COUNT] (accumulated since the last call) and recycles the events.
 * The work counters of the branch-path / post-BP min-sum calls (swd_counters: path_edge_iters, path_vn_iters, path_cn_iters,
 * path_slot_iters) accumulate only while profiling is enabled: counting them costs 1.7 % of the decoded shots/s. */
int  swd_set_profiling(swd_decoder *d, int enable);
int  swd_get_kernel_times(swd_decoder *d, double *ms, uint64_t *launches);

int  swd_get_counters(swd_decoder *d, swd_counters *out);
int  swd_reset_counters(swd_decoder *d);
int  swd_rank(swd_decoder *d);          /* GF(2) rank of the pcm (mod2sparse_rank), -1 if n/a      */
int  swd_new_n(swd_decoder *d);

/* ---- sliding-window bookkeeping on the device (guessing.py:204-227, osd.py:170-179) ------------
 * A "window plan" holds the full detector matrix chk (num_det x num_col, CSC) and the observable
 * matrix obs (num_obs x num_col, CSC).  All shot data is one byte per bit, row-major, on device. */
typedef struct swd_window swd_window;
int  swd_window_create(int device, int num_det, int num_col, const int32_t *chk_colptr, const int32_t *chk_rowidx,
                       int num_obs, const int32_t *obs_colptr, const int32_t *obs_rowidx, swd_window **out);
void swd_window_destroy(swd_window *w);
/* copy syndrome columns [row0,row0+m) of d_det[B,num_det] into d_synd[B,m]  (detector_win = new_det_data[:, a0:b0]) */
int  swd_window_extract(swd_window *w, const uint8_t *d_det, int64_t B, int row0, int m, uint8_t *d_synd, void *stream);
/* commit: for every shot XOR chk[:, col0+j] into d_det and obs[:, col0+j] into d_obs for each j < ncommit with
 * d_corr[b, j] == 1 (d_corr has row stride n_win); i.e. new_det = det + e_hat @ chk.T restricted to the commit. */
int  swd_window_commit(swd_window *w, const uint8_t *d_corr, int64_t B, int n_win, int col0, int ncommit,
                       uint8_t *d_det, uint8_t *d_obs, void *stream);
/* counts over shots: flagged = any(det residual), logical = flagged or any(obs residual). out[0]=flagged, out[1]=failed */
int  swd_window_count_failures(swd_window *w, const uint8_t *d_det, const uint8_t *d_obs, int64_t B,
                               unsigned long long *d_out2, void *stream);

/* DEM sampling on the device - what CompiledDemSampler.sample draws in the reference's drivers (guessing.py:129-130):
 * one independent Bernoulli(priors[c]) per DEM column, det = chk . e, obs = obs_mat . e (mod 2).  Philox4x32-10 keyed by
 * `seed`, counter = (shot_offset + b, column block): shot b of a call with offset o equals shot 0 of a call with offset o + b.
 * d_err (optional, [B, num_col]) receives the sampled error vectors. */
int  swd_window_set_priors(swd_window *w, const double *priors /* host, [num_col] */);
int  swd_window_sample(swd_window *w, uint64_t seed, int64_t shot_offset, int64_t B, uint8_t *d_det, uint8_t *d_obs,
                       uint8_t *d_err, void *stream);

/* ---- bp4_osd (src/bp4_osd.pyx:6-684): quaternary min-sum BP over the pair (Hx, Hz) of a CSS code under depolarizing
 * noise, then one OSD per basis for shots whose BP did not converge.  Replaces bp4_osd.__cinit__ / decode (pyx:8-221);
 * swd_bp4_camel_decode_batch_host replaces camel_decode (pyx:223-248).  The two graphs have no column- / row-weight limit
 * (CAMEL codes tie every check to the last qubit).  llr_*: log((1-px-py-pz)/p_*) per qubit, prior_llr_x / _z:
 * log((1-(px+py))/(px+py)) and log((1-(pz+py))/(pz+py)) (pyx:123-133), computed by the caller with libm.
 * decode: synd_x[B*mx] (syndrome of Hx, i.e. of the Z part), synd_z[B*mz]; dec[B*2n] = x part then z part per shot
 * (stackchar2numpy); optional bp_dec / osd0 [B*2n], log_prob_ratios [B*n*3] (x, y, z), bp_iteration [B]. */
typedef struct swd_bp4 swd_bp4;
int  swd_bp4_create(int device, int mx, int mz, int n, const int32_t *hx_colptr, const int32_t *hx_rowidx,
                    const int32_t *hz_colptr, const int32_t *hz_rowidx, const double *llr_x, const double *llr_y,
                    const double *llr_z, const double *prior_llr_x, const double *prior_llr_z, int max_iter,
                    double ms_scaling_factor, int osd_method, int osd_order, swd_bp4 **out);
void swd_bp4_destroy(swd_bp4 *b);
int  swd_bp4_rank(swd_bp4 *b, int which /* 0: Hx, 1: Hz */);
int  swd_bp4_decode_batch_host(swd_bp4 *b, const uint8_t *synd_x, const uint8_t *synd_z, int64_t B, uint8_t *dec, uint8_t *converge,
                               uint8_t *bp_dec, uint8_t *osd0, double *log_prob_ratios, int32_t *bp_iteration);
/* camel_decode (pyx:223-248): four BP runs per shot with the last qubit pinned to I / X / Z / Y (vn_set_value, pyx:389-423), the
 * converged run with the smallest path metric (cal_pm, pyx:250-259) wins, ties to the earlier value.  dec[B*2n], converge[B];
 * optional min_pm[B] (10000.0 and dec = 0 when no run converged - the reference returns an earlier call's buffer there),
 * log_prob_ratios [B*n*3] and bp_iteration [B] of the last run, as the reference's properties hold after the call. */
int  swd_bp4_camel_decode_batch_host(swd_bp4 *b, const uint8_t *synd_x, const uint8_t *synd_z, int64_t B, uint8_t *dec,
                                     uint8_t *converge, double *min_pm, double *log_prob_ratios, int32_t *bp_iteration);
/* The same two calls with device pointers (e.g. torch tensors) on the caller's stream, without host synchronisation; the
 * optional outputs may be NULL.  Results pass through the decoder's own buffers: calls on one swd_bp4 must be stream-ordered. */
int  swd_bp4_decode_batch_device(swd_bp4 *b, const uint8_t *d_synd_x, const uint8_t *d_synd_z, int64_t B, uint8_t *d_dec,
                                 uint8_t *d_converge, uint8_t *d_bp_dec, uint8_t *d_osd0, double *d_log_prob_ratios,
                                 int32_t *d_bp_iteration, void *stream);
int  swd_bp4_camel_decode_batch_device(swd_bp4 *b, const uint8_t *d_synd_x, const uint8_t *d_synd_z, int64_t B, uint8_t *d_dec,
                                       uint8_t *d_converge, double *d_min_pm, double *d_log_prob_ratios, int32_t *d_bp_iteration,
                                       void *stream);

const char *swd_strerror(int status);
const char *swd_last_error(void);
const char *swd_version(void);

#ifdef __cplusplus
}
#endif
#endif
