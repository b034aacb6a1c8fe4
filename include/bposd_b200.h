/*
 * bposd_b200.h -- C ABI of the B200-native BP+OSD decoder (libbposd_b200.so).
 *
 * This is the drop-in boundary for the decode hot path of quantumgizmos/bp_osd.  In the
 * reference the boundary is a Python class, `bposd.bposd_decoder` == `ldpc.bposd_decoder`
 * (/root/reference/src/bposd/__init__.py:1), also used as `ldpc.BpOsdDecoder`
 * (src/bposd/css_decode_sim.py:6).  Each entry point below names the reference call site it
 * replaces.  Plain pointers and sizes only; no torch or Python types.  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative BPOSD_E* code; the message
 * is available from bposd_last_error().  A handle is bound to one CUDA device, is not
 * thread-safe, and owns copies of everything passed to bposd_create (the reference's
 * harness keeps and re-uses its matrices and marks its probability arrays read-only,
 * css_decode_sim.py:141-142,432-434).  Pointers named d_* are device pointers on the
 * handle's device, h_* are host pointers.  `stream` is a cudaStream_t passed as void*.
 */
#ifndef BPOSD_B200_H
#define BPOSD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bposd_handle bposd_t;

/* bp_method: README.md:183, css_decode_sim.py:35,70,448 ("ms"/"minimum_sum", "ps"/"product_sum") */
#define BPOSD_BP_PRODUCT_SUM 0
#define BPOSD_BP_MINIMUM_SUM 1
/* osd_method: README.md:185, css_decode_sim.py:41,450.  OFF = BP only (no reference analogue) */
#define BPOSD_OSD_0 0
#define BPOSD_OSD_E 1
#define BPOSD_OSD_CS 2
#define BPOSD_OSD_OFF 3
/* precision: 64 = bit-exact mode (IEEE double, reference accumulation order), 32 = fast mode */
#define BPOSD_FP64 64
#define BPOSD_FP32 32

#define BPOSD_OK 0
#define BPOSD_EINVAL (-1)  /* bad argument (maps to ValueError in the Python layer)          */
#define BPOSD_ECUDA (-2)   /* CUDA runtime error; bposd_last_error() holds the CUDA message   */
#define BPOSD_ENOMEM (-3)  /* host or device allocation failed                                */
#define BPOSD_EUNSUP (-4)  /* configuration not supported by the kernels built into this lib  */

/* Device output block of bposd_decode_batch.  Any pointer may be NULL (that output is not
 * produced).  Layout row-major [B, n].  Mirrors the result attributes the reference reads
 * after decode(): osdw_decoding / osd0_decoding / bp_decoding (css_decode_sim.py:257,294,338;
 * README.md:202), converge (css_decode_sim.py:331-336), log_prob_ratios and iter (ldpc
 * attributes required by BASELINE.json's north_star). */
typedef struct {
    uint8_t *d_osdw;    /* [B, n] 0/1                                                  */
    uint8_t *d_osd0;    /* [B, n] 0/1                                                  */
    uint8_t *d_bp;      /* [B, n] 0/1, hard decision of the last BP iteration          */
    void *d_llr;        /* [B, n] double (precision 64) or float (precision 32)        */
    uint8_t *d_converge;/* [B] 1 if BP reproduced the syndrome                         */
    int32_t *d_iter;    /* [B] BP iterations executed                                  */
} bposd_out_t;

typedef struct {
    int32_t m, n, nnz, rank, k;     /* k = n - rank                                      */
    int32_t max_iter, bp_method, osd_method, osd_order, precision, device;
    int32_t bp_kernel;              /* 0 generic/global, 1 generic/smem, 2 in-place smem, 3 cluster DSMEM */
    int32_t bp_threads, bp_ctas_per_sm, bp_smem_bytes;
    int32_t osd_threads, osd_smem_bytes, sm_count;
    int32_t osd_variant;            /* 3 register kernel, 1 shared-memory kernel, 4 / 2 HBM-resident OSD-0 kernels (cluster / single CTA), 0 OSD unsupported */
    int32_t bp_layout_excess;       /* kernel 2: shared-memory wavefronts per bit sweep above the conflict-free count;
                                       kernel 3: edges whose bit and check live in different CTAs, per mille */
    int32_t bp_cluster_size;        /* CTAs per cluster of kernel 3, else 1 */
    int32_t reserved;
    double ms_scaling_factor;
} bposd_info_t;

/* Per-launch statistics of the most recent bposd_decode_batch / bposd_decode_host call
 * (read back with a stream synchronise). */
typedef struct {
    int64_t shots;
    int64_t bp_converged;
    int64_t osd_invocations;
    int64_t bp_iterations;   /* sum over shots of iterations executed */
    float ms_bp;             /* CUDA-event time of the BP kernels, summed over chunks  */
    float ms_osd;            /* CUDA-event time of the OSD kernels, summed over chunks */
    int32_t launches;        /* kernels launched */
    int32_t chunks;
} bposd_stats_t;

/* Constructor.  Replaces `bposd_decoder(H, error_rate=..., channel_probs=..., max_iter=...,
 * bp_method=..., ms_scaling_factor=..., osd_method=..., osd_order=...)`
 * (README.md:178-187; css_decode_sim.py:444-463).  H is CSR over GF(2) with ascending
 * column indices inside each row.  max_iter == 0 means n (css_decode_sim.py:72).
 * ms_scaling_factor == 0 selects the variable factor 1 - 2^-iteration (README.md:184). */
int bposd_create(const int32_t *h_indptr, const int32_t *h_indices, int32_t m, int32_t n,
                 const double *h_channel_probs, int32_t max_iter, int32_t bp_method,
                 double ms_scaling_factor, int32_t osd_method, int32_t osd_order,
                 int32_t precision, int32_t device, bposd_t **out);

/* Replaces `decoder.update_channel_probs(probs)` (css_decode_sim.py:229,248). */
int bposd_update_channel_probs(bposd_t *h, const double *h_channel_probs);

/* Batched form of `decoder.decode(syndrome)` (README.md:197; css_decode_sim.py:174-202):
 * d_syndromes is [B, m] uint8 (0/1) on the device.  Asynchronous on `stream` except for a
 * final stream synchronise that collects bposd_stats_t.
 * d_priors_per_shot: NULL, or [B, n] prior LLRs log((1-p)/p) in the handle's precision, one
 * row per shot (the per-shot channel update of css_decode_sim.py:207-248).
 * d_weights_per_shot: NULL, or [B, n] OSD weights log(1/p) (double), one row per shot; the reference's
 * update_channel_probs changes the BP priors and the OSD weights together (row a16). */
int bposd_decode_batch(bposd_t *h, const uint8_t *d_syndromes, int64_t B, const bposd_out_t *out,
                       const void *d_priors_per_shot, const double *d_weights_per_shot, void *stream);

/* Same, with host buffers: copies h_syndromes [B, m] to the device, decodes, copies the
 * requested outputs back (NULL = not wanted).  This is what the reference-facing
 * `decode()` / `decode_batch(numpy)` call goes through.  Large batches are cut into chunks that are
 * double-buffered over two internal streams, so the copies overlap the kernels; pass pinned host
 * memory for the copies to be asynchronous.
 * Batches of at most one shot per SM (single-shot decode() above all) take a latency path: one kernel
 * launch and one synchronisation, syndromes read from and results written to pinned host memory that the
 * device addresses directly, one SM per shot with few bits per thread; the OSD kernel is launched only if
 * the converge flags the host reads back say a shot needs it.  bposd_stats_t then carries no event times
 * (ms_bp = ms_osd = 0). */
int bposd_decode_host(bposd_t *h, const uint8_t *h_syndromes, int64_t B, uint8_t *h_osdw,
                      uint8_t *h_osd0, uint8_t *h_bp, void *h_llr, uint8_t *h_converge,
                      int32_t *h_iter);

/* Bit-packed forms of the two decode entry points (BASELINE.json north_star: "bit-packed H e mod 2 syndrome kernel").
 * A syndrome row is ceil(m/8) bytes, a decoding row ceil(n/8) bytes, bit i%8 of byte i/8 = entry i -- the layout of
 * numpy.packbits(x, axis=1, bitorder="little"); the reference forms both as byte arrays (`hz @ error % 2`,
 * css_decode_sim.py:173-201; `osdw_decoding`, README.md:202), which is 8x the PCIe traffic the path needs.
 * d_osdw / d_osd0 / d_bp of `out` (resp. the h_*_bits arguments) are packed, LLRs, converge flags and iteration
 * counts are as in the byte forms.  The BP kernels read the packed syndromes directly; the decodings are packed by a
 * streaming kernel before they leave the device. */
int bposd_decode_batch_packed(bposd_t *h, const uint8_t *d_syndrome_bits, int64_t B, const bposd_out_t *out,
                              const void *d_priors_per_shot, const double *d_weights_per_shot, void *stream);
int bposd_decode_host_packed(bposd_t *h, const uint8_t *h_syndrome_bits, int64_t B, uint8_t *h_osdw_bits,
                             uint8_t *h_osd0_bits, uint8_t *h_bp_bits, void *h_llr, uint8_t *h_converge,
                             int32_t *h_iter);

/* Device-side restatement of the harness step `_generate_error` + `H @ e % 2`
 * (css_decode_sim.py:465-498,173-201): Philox4x32-10, counter (j/4, 0, shot_lo, shot_hi),
 * key = seed, one 32-bit uniform r per qubit; r < t1 -> Z, t1 <= r < t2 -> X, t2 <= r < t3 -> Y
 * (h_t1..h_t3 are per-qubit cumulative uint32 thresholds).  Writes the sector error this
 * decoder corrects (`sector` 0: X component, 1: Z component) to d_errors [B, n] (may be
 * NULL) and its syndrome H e mod 2 to d_syndromes [B, m]. */
int bposd_set_channel_thresholds(bposd_t *h, const uint32_t *h_t1, const uint32_t *h_t2,
                                 const uint32_t *h_t3);
int bposd_sample_syndromes(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector,
                           uint8_t *d_errors, uint8_t *d_syndromes, void *stream);
/* Same, syndromes written bit-packed ([B, ceil(m/8)], layout as above); d_errors stays one byte per qubit. */
int bposd_sample_syndromes_packed(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector,
                                  uint8_t *d_errors, uint8_t *d_syndrome_bits, void *stream);

/* Logical operators for the failure check (css_decode_sim.py:257-272): CSR, K rows. */
int bposd_set_logicals(bposd_t *h, const int32_t *h_indptr, const int32_t *h_indices, int32_t K);

/* fail[b] = ((L @ (e ^ d)) % 2).any() for B shots; also accumulates *d_fail_count (int64,
 * may be NULL) and the minimum weight of a failing residual into *d_min_weight (int32, may
 * be NULL; css_decode_sim.py:261-264). */
int bposd_logical_check(bposd_t *h, const uint8_t *d_errors, const uint8_t *d_decodings,
                        int64_t B, uint8_t *d_fail, int64_t *d_fail_count, int32_t *d_min_weight,
                        int32_t *d_resid_weight /* [B] weight of e ^ d per shot, may be NULL */, void *stream);

/* Per-shot channel update between the two sectors of a CSS decode (`_channel_update`,
 * css_decode_sim.py:207-248): the second decoder's probability of qubit j is h_probs_if1[j] where the
 * first sector's decoding d_first_decoding[b, j] is 1 and h_probs_if0[j] where it is 0.  Writes the
 * BP priors log((1-p)/p) ([B, n], handle precision) and the OSD weights log(1/p) ([B, n] double, may be
 * NULL) that bposd_decode_batch takes as d_priors_per_shot / d_weights_per_shot. */
int bposd_channel_update(bposd_t *h, const uint8_t *d_first_decoding, int64_t B, const double *h_probs_if0,
                         const double *h_probs_if1, void *d_priors_out, double *d_weights_out, void *stream);

/* Counters of the two-sector simulation (`_encoded_error_rates`, css_decode_sim.py:250-365) from the
 * per-shot flags of bposd_logical_check.  X failures are looked at first, Z only where X passed (the
 * reference's `elif`).  Accumulates into h_counters[8] (int64): [0] shots, [1] bp_converge_x,
 * [2] bp_converge_z, [3] bp_success (both sectors converged and no logical error), [4] osd0_success,
 * [5] osdw_success, [6] minimum weight of a failing osdw/osd0 residual (min-combined, 0 = none yet). */
typedef struct {
    const uint8_t *d_fail_x, *d_fail_z;     /* [B] */
    const int32_t *d_weight_x, *d_weight_z; /* [B] residual weights, may be NULL */
} bposd_css_sector_t;
int bposd_css_counters(bposd_t *h, int64_t B, const bposd_css_sector_t *osdw, const bposd_css_sector_t *osd0,
                       const bposd_css_sector_t *bp, const uint8_t *d_converge_x, const uint8_t *d_converge_z,
                       int64_t *h_counters, void *stream);

/* One Monte-Carlo step of one sector, all on the device (css_decode_sim.py:163-205 for a
 * single sector): sample B errors starting at global shot index shot0, syndromes, decode,
 * residual check against the logicals, accumulate into h_counters[8] (int64):
 * [0] shots, [1] bp_converged, [2] bp_success, [3] osd0_success, [4] osdw_success,
 * [5] osd_invocations, [6] bp_iterations, [7] min_logical_weight (min-combined). */
int bposd_sample_and_decode(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector,
                            int64_t *h_counters, void *stream);

int bposd_get_info(const bposd_t *h, bposd_info_t *info);
int bposd_get_stats(const bposd_t *h, bposd_stats_t *stats);
/* Tuning knobs (0 = keep automatic): BP kernel variant (+1 of bposd_info_t.bp_kernel),
 * threads per CTA, workspace bytes for the failed-shot LLR buffer. */
int bposd_set_tuning(bposd_t *h, int32_t bp_kernel_plus1, int32_t bp_threads, int64_t workspace_bytes);
/* Measured INT32 logic-op (LOP3) rate of the device in ops/s: the denominator of the OSD roofline. */
int bposd_int32_peak(bposd_t *h, double *ops_per_s);
/* Measured shared-memory bandwidth of the device in bytes/s (conflict-free 16-byte LDS + STS in equal parts, the access
 * mix of the in-place BP kernels): the denominator of the BP roofline. */
int bposd_smem_peak(bposd_t *h, double *bytes_per_s);
/* Measured fp64 fused-multiply-add rate of the device in DFMA/s (thread-level operations): the denominator of the
 * product-sum roofline (that update is bound by the fp64 pipe: tanh, log and three divisions per edge). */
int bposd_fp64_peak(bposd_t *h, double *fma_per_s);
/* Test hook: the device side of include/bposd_math.h evaluated element-wise on device arrays of `count` doubles, so that
 * the tests can compare it bit for bit with the host side (the oracle).  fn: 0 out = a / b (the in-range division
 * sequence), 1 tanh(a), 2 log(a), 3 (1 + a) / (1 - a) as the product-sum update forms it; b is read for fn = 0 only. */
int bposd_math_probe(bposd_t *h, int32_t fn, const double *a, const double *b, double *out, int64_t count);
/* BP schedule (ldpc.BpOsdDecoder's `schedule` / `serial_schedule_order`; /root/reference never passes them, SURVEY row
 * f4).  schedule 0: parallel (flooding), the default.  1: serial -- inside an iteration the bits are visited one after the
 * other in `order` (host array, a permutation of 0 .. n-1; NULL = 0, 1, ..., n-1), each recomputing the check-to-bit
 * messages of its edges from the current bit-to-check messages.  One thread per shot (bp_serial_kernel.cuh). */
int bposd_set_schedule(bposd_t *h, int32_t schedule, const int32_t *order);
/* Thread-block-cluster size of BP kernel variant 3 (messages split over the shared memory of 2, 4, 8 or
 * 16 CTAs, reached through distributed shared memory); 0 = smallest size that fits. */
int bposd_set_cluster_size(bposd_t *h, int32_t cluster_size);
/* OSD kernel variant: 0 automatic; 3 register kernel (the m x m row-operation matrix in registers, warp-specialised
 * resolver / updater pipeline with one barrier per 16 sorted columns; OSD-0/E/CS; the default for m <= 1024);
 * 1 shared-memory kernel (the same matrix in shared memory, one sweep and two barriers per pivot; OSD-0/E/CS; the
 * fall-back for larger m while the matrix fits); 4 HBM-resident left-looking cluster kernel (OSD-0 only; one thread-block
 * cluster per failed shot, checks split over its CTAs, 64-column panels, row-operation masks streamed from HBM with
 * cp.async.bulk + mbarrier, pivot rows exchanged through distributed shared memory; chosen automatically when nothing
 * fits in shared memory, BASELINE config 5); 2 the single-CTA form of it (32-column panels, one CTA per failed shot).
 * workspace_bytes > 0 caps the HBM workspace of variants 4 / 2 (ceil(n/64) * m * 8 bytes per concurrently processed
 * failed shot). */
int bposd_set_osd_variant(bposd_t *h, int32_t variant, int64_t workspace_bytes);
const char *bposd_last_error(const bposd_t *h);
const char *bposd_version(void);
void bposd_destroy(bposd_t *h);

#ifdef __cplusplus
}
#endif
#endif /* BPOSD_B200_H */
