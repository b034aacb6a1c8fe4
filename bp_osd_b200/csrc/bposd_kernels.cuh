// bposd_kernels.cuh -- sm_100a device code of the BP+OSD decode path.
//
// Kernels (one per hot-path function of SURVEY.md section 8a):
//   bp_generic_kernel   rows a3-a8   flooding min-sum / product-sum BP, any degrees, one shot per
//                                    CTA at a time, persistent CTAs pulling shots from an atomic
//                                    queue (iteration counts are heavy tailed), messages in
//                                    shared memory when they fit, else in an L2-resident
//                                    per-CTA scratch in HBM.
//   osd_kernel          rows a9-a14  per failed shot: stable LLR rank sort, GF(2) elimination
//                                    kept as the m x m row-operation matrix T in shared memory
//                                    (column-major bit-packed), OSD-0 read-out, OSD-E / OSD-CS
//                                    candidate search (popcount or ordered soft weights).
//   sample_syndrome_kernel rows a17-a18  Philox4x32-10 error sampler + H e mod 2.
//   logical_check_kernel   row a19   residual against the logical operators + counters.
//
// fp64 mode must reproduce the reference arithmetic bit for bit: this file is compiled with
// -fmad=false and every floating-point accumulation follows the edge order documented in
// SURVEY.md section 8a (ascending column inside a check, ascending row inside a bit).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/bposd_math.h"

namespace bposd {

struct GraphDev {
    int m, n, nnz;
    const int *row_ptr, *col_idx;            // CSR: ascending column inside a row
    const int *col_ptr, *row_idx, *csc_slot; // CSC: ascending row inside a column; CSR slot of each entry
};

template <typename real>
struct BpArgs {
    GraphDev g;
    int max_iter;
    int method;   // 0 product-sum, 1 min-sum
    real alpha0;  // 0 => 1 - 2^-it
    int safe_it;  // no message can have overflowed before this pass (see overflow_safe_iterations); 0: always check
    const real *prior;
    long long prior_stride; // 0: one prior vector for all shots, n: per-shot rows
    int uniform_prior;      // 1: every bit has the same prior (a scalar in registers)
    const uint8_t *synd;
    int synd_packed;        // 0: [B, m] bytes (0/1); 1: [B, ceil(m/8)] bytes, bit i%8 of byte i/8 = check i
    long long B;
    uint8_t *bp, *osd0, *osdw;
    real *llr;
    uint8_t *converge;
    int *iter;
    int *fail_count;
    int *fail_list;
    real *fail_llr; // [capacity, n], used when llr == nullptr
    int osd_off;
    unsigned long long *queue;
    unsigned long long *stat; // [0] converged shots, [1] iterations
    real *g_scratch;          // global mode: per-CTA [2*nnz + n] reals
    uint8_t *g_dec;           // global mode: per-CTA [n]
};

template <typename real> __device__ __forceinline__ real real_max();
template <> __device__ __forceinline__ double real_max<double>() { return DBL_MAX; }
template <> __device__ __forceinline__ float real_max<float>() { return FLT_MAX; }

// Overflow guard of the prefix / suffix check update (fast_check_compute).  The reference's running minimum starts from the
// largest finite value (`temp = numeric_limits<double>::max(); if (abs(b2c) < temp) temp = abs(b2c)`), so a check message
// never exceeds it even when every incoming message has overflowed to +-inf; the prefix / suffix form starts from the
// row's own entries and would pass the infinity on.  Messages of shots that do not converge grow geometrically (cfg 3
// reaches 6e234 after 1 922 passes), so this matters for max_iter beyond a few thousand passes -- config 5 runs 40 000.
// An infinite bit-to-check message and +-max are the same thing to the reference's check update (neither is smaller than
// the starting value; `<= 0` sees the same sign), so: from the pass where an overflow is conceivable (safe_it, host side)
// every thread tests its bits' LLRs after the bit sweep; a sum that is not finite and below 1e291 raises a CTA-wide flag, and
// the next check sweep first rewrites +-inf in its rows as +-max.  Nothing is added to passes before safe_it.
template <typename real> __device__ __forceinline__ bool llr_near_overflow(real t);
// Threshold: below half an ulp of the largest finite value (1e292), so that a bit whose LLR passes the test cannot have
// produced an infinite bit-to-check message either: such a message is the LLR minus one finite check message.
template <> __device__ __forceinline__ bool llr_near_overflow<double>(double t) { return !(fabs(t) <= 1e291); }
template <> __device__ __forceinline__ bool llr_near_overflow<float>(float t) { return !(fabsf(t) <= 1e30f); }
template <typename real> __device__ __forceinline__ real clamp_inf(real v);
template <> __device__ __forceinline__ double clamp_inf<double>(double v) { return (fabs(v) > DBL_MAX) ? copysign(DBL_MAX, v) : v; }
template <> __device__ __forceinline__ float clamp_inf<float>(float v) { return (fabsf(v) > FLT_MAX) ? copysignf(FLT_MAX, v) : v; }

// syndrome bit of check i of a shot, from the byte-per-check or the bit-packed layout
__device__ __forceinline__ unsigned synd_bit(const uint8_t *synd, long long shot, int m, int i, int packed) {
    return packed ? ((unsigned)(synd[shot * (long long)((m + 7) >> 3) + (i >> 3)] >> (i & 7)) & 1u)
                  : ((unsigned)synd[shot * (long long)m + i] & 1u);
}

__device__ __forceinline__ double ms_alpha(double alpha0, int it) {
    return alpha0 == 0.0 ? 1.0 - ldexp(1.0, -it) : alpha0;
}
__device__ __forceinline__ float ms_alpha(float alpha0, int it) {
    return alpha0 == 0.0f ? 1.0f - ldexpf(1.0f, -it) : alpha0;
}
// Product-sum in fp32 (fast mode only): a product of tanh values rounds to exactly +-1 once every incoming
// message exceeds ~17, and log((1+x)/(1-x)) then returns +-inf, which turns the next bit sums into NaN.
// The fast mode keeps x one ulp inside (-1, 1), i.e. caps a check message at log(2^25) ~ 17.3.  The fp64
// mode applies no clipping, exactly like the reference (row a5).
__device__ __forceinline__ double ps_clamp(double x) { return x; }
__device__ __forceinline__ float ps_clamp(float x) { return fminf(fmaxf(x, -0.99999994f), 0.99999994f); }
// fp64: the portable FMA-free functions the oracle also compiles (include/bposd_math.h) -- same bits on both sides
// (1 + x) / (1 - x) of the product-sum update (the reference's one division per edge).  fp64: the in-range IEEE sequence
// of include/bposd_math.h (1 +- x lie in {0} U [2^-53, 2]); a saturated product x = 1 divides by zero: +inf, as on the host.
__device__ __forceinline__ double ps_ratio(double x) {
    const double a = 1 + x, b = 1 - x, q = bpm_div(a, b);
    return b == 0 ? __longlong_as_double(0x7ff0000000000000ll) : q;
}
__device__ __forceinline__ float ps_ratio(float x) { return (1 + x) / (1 - x); }
__device__ __forceinline__ double r_tanh(double x) { return bpm_tanh(x); }
__device__ __forceinline__ float r_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double r_log(double x) { return bpm_log(x); }
__device__ __forceinline__ float r_log(float x) { return logf(x); }
__device__ __forceinline__ double r_abs(double x) { return fabs(x); }
__device__ __forceinline__ float r_abs(float x) { return fabsf(x); }

// ---------------------------------------------------------------------------------------------
// Shared epilogue: write one shot's results and queue it for OSD if BP failed.
// ---------------------------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ void bp_write_shot(const BpArgs<real> &a, long long shot, bool conv, int iters,
                                              const real *llr_s, const uint8_t *dec_s, int *sh_slot) {
    const int n = a.g.n;
    const int tid = threadIdx.x, T = blockDim.x;
    const bool final_here = conv || a.osd_off;
    if (!final_here) {
        if (tid == 0) {
            int slot = atomicAdd(a.fail_count, 1);
            a.fail_list[slot] = (int)shot;
            *sh_slot = slot;
        }
        __syncthreads();
    }
    const long long base = shot * (long long)n;
    for (int j = tid; j < n; j += T) {
        uint8_t d = dec_s[j];
        if (a.bp) a.bp[base + j] = d;
        if (final_here) {
            if (a.osd0) a.osd0[base + j] = d;
            if (a.osdw) a.osdw[base + j] = d;
        }
        if (a.llr) a.llr[base + j] = llr_s[j];
        else if (!final_here) a.fail_llr[(long long)(*sh_slot) * n + j] = llr_s[j];
    }
    if (tid == 0) {
        if (a.converge) a.converge[shot] = conv ? 1 : 0;
        if (a.iter) a.iter[shot] = iters;
    }
}

#ifndef BPOSD_KERNELS_COMMON_ONLY // the specialised-BP translation units need only the declarations above
// ---------------------------------------------------------------------------------------------
// Generic BP kernel (rows a3-a8).  Literal flooding schedule with separate bit->check and
// check->bit arrays, any row/column degree.  SMEM=true: state carved from dynamic shared
// memory; SMEM=false: per-CTA scratch in global memory (stays in the 126 MB L2 for the
// concurrently active shots when it can).
// Convergence of iteration `it` is tested at the start of pass it+1, folded into the check
// sweep (each check XORs the hard decisions of its neighbours), so an iteration costs two
// block barriers; the vote rides on __syncthreads_and.
// ---------------------------------------------------------------------------------------------
template <typename real, bool SMEM>
__global__ void __launch_bounds__(1024) bp_generic_kernel(BpArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n, E = g.nnz;
    const int tid = threadIdx.x, T = blockDim.x;

    real *b2c, *c2b, *llr_s;
    uint8_t *dec_s, *synd_s;
    __shared__ long long sh_shot;
    __shared__ int sh_slot;
    if (SMEM) {
        b2c = reinterpret_cast<real *>(smem_raw);
        c2b = b2c + E;
        llr_s = c2b + E;
        dec_s = reinterpret_cast<uint8_t *>(llr_s + n);
        synd_s = dec_s + n;
    } else {
        real *base = a.g_scratch + (size_t)blockIdx.x * (2 * (size_t)E + n);
        b2c = base;
        c2b = base + E;
        llr_s = base + 2 * (size_t)E;
        dec_s = a.g_dec + (size_t)blockIdx.x * n;
        synd_s = smem_raw;
    }
    unsigned long long n_conv = 0, n_iter = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh_shot = (long long)atomicAdd(a.queue, 1ull);
        __syncthreads();
        const long long shot = sh_shot;
        if (shot >= a.B) break;
        const real *prior = a.prior + shot * a.prior_stride;

        for (int i = tid; i < m; i += T) synd_s[i] = (uint8_t)synd_bit(a.synd, shot, m, i, a.synd_packed);
        for (int j = tid; j < n; j += T) { // a3
            real p = prior[j];
            llr_s[j] = p;
            dec_s[j] = 0;
            for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; q++) b2c[g.csc_slot[q]] = p;
        }
        __syncthreads();

        bool conv = false;
        int iters = 0;
        for (int it = 1;; it++) {
            // ---- check sweep of pass `it` (+ convergence vote for pass it-1) ----
            bool ok = true;
            const bool last = it > a.max_iter;
            if (a.method == 1) {
                const real alpha = ms_alpha(a.alpha0, it);
                for (int i = tid; i < m; i += T) { // a4
                    const int beg = g.row_ptr[i], end = g.row_ptr[i + 1];
                    real min1 = real_max<real>(), min2 = real_max<real>();
                    int arg = -1, tot = synd_s[i], par = 0;
                    for (int e = beg; e < end; e++) {
                        real v = b2c[e];
                        tot += (v <= 0) ? 1 : 0;
                        real av = r_abs(v);
                        if (av < min1) { min2 = min1; min1 = av; arg = e; }
                        else if (av < min2) min2 = av;
                        par ^= dec_s[g.col_idx[e]];
                    }
                    if (it > 1 && par != synd_s[i]) ok = false;
                    if (!last)
                        for (int e = beg; e < end; e++) {
                            real v = b2c[e];
                            int sg = tot + ((v <= 0) ? 1 : 0);
                            real mag = (e == arg) ? min2 : min1;
                            c2b[e] = mag * ((sg & 1) ? -alpha : alpha);
                        }
                }
            } else {
                for (int i = tid; i < m; i += T) { // a5
                    const int beg = g.row_ptr[i], end = g.row_ptr[i + 1];
                    int par = 0;
                    real t = 1;
                    for (int e = beg; e < end; e++) {
                        par ^= dec_s[g.col_idx[e]];
                        if (!last) {
                            // tanh is evaluated once per edge: the value is parked in b2c[e], which the bit
                            // sweep overwrites before anything reads it again
                            const real th = r_tanh(b2c[e] / 2);
                            b2c[e] = th;
                            c2b[e] = t;
                            t *= th;
                        }
                    }
                    if (it > 1 && par != synd_s[i]) ok = false;
                    if (!last) {
                        t = 1;
                        const real sgn = synd_s[i] ? (real)-1 : (real)1;
                        for (int e = end - 1; e >= beg; e--) {
                            real x = ps_clamp(c2b[e] * t);
                            c2b[e] = sgn * r_log(ps_ratio(x));
                            t *= b2c[e];
                        }
                    }
                }
            }
            const int all_ok = __syncthreads_and(ok ? 1 : 0);
            if (it > 1 && all_ok) { conv = true; iters = it - 1; break; } // a7
            if (last) { iters = a.max_iter; break; }
            // ---- bit sweep of pass `it` (a6 then a8) ----
            for (int j = tid; j < n; j += T) {
                const int beg = g.col_ptr[j], end = g.col_ptr[j + 1];
                real t = prior[j];
                for (int q = beg; q < end; q++) {
                    int e = g.csc_slot[q];
                    b2c[e] = t;
                    t += c2b[e];
                }
                llr_s[j] = t;
                dec_s[j] = (t <= 0) ? 1 : 0;
                real s = 0;
                for (int q = end - 1; q >= beg; q--) {
                    int e = g.csc_slot[q];
                    b2c[e] += s;
                    s += c2b[e];
                }
            }
            __syncthreads();
        }
        if (a.max_iter <= 0) { conv = false; iters = 0; }
        bp_write_shot<real>(a, shot, conv, iters, llr_s, dec_s, &sh_slot);
        if (tid == 0) { n_conv += conv ? 1 : 0; n_iter += (unsigned long long)iters; }
    }
    if (tid == 0 && a.stat) {
        atomicAdd(&a.stat[0], n_conv);
        atomicAdd(&a.stat[1], n_iter);
    }
}

// ---------------------------------------------------------------------------------------------
// OSD kernel (rows a9-a14): one CTA per failed shot.
// ---------------------------------------------------------------------------------------------
template <typename real>
struct OsdArgs {
    GraphDev g;
    int S;        // 32-bit words per column of T (ceil(m/32))
    int St;       // stride of a T column in words (S, padded to odd)
    int method;   // 0 osd0, 1 osd_e, 2 osd_cs
    int order;    // search depth w
    int uniform;  // 1: all channel probabilities equal and in (0,1): weight = popcount
    const double *weight; // [n] log(1/p_j), or [B, n] per shot when weight_stride == n
    long long weight_stride;
    const uint8_t *synd;
    int synd_packed;      // as BpArgs::synd_packed
    const real *llr;      // [B, n] if llr_by_shot else [capacity, n] indexed by fail slot
    int llr_by_shot;
    const int *fail_count;
    const int *fail_list;
    uint8_t *osd0, *osdw;
    unsigned long long *stat; // [2] osd invocations
    int maxrank;  // rank(H): no pivot can follow the maxrank-th (osd_reg_kernel stops its scan there)
    int maxdeg;   // largest column degree of H
    int np2;      // n rounded up to a power of two (osd_reg_kernel: size of its bitonic sort)
};

__device__ __forceinline__ unsigned long long sort_key(double x) {
    if (x == 0.0) x = 0.0; // -0 and +0 compare equal in the reference's comparator
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long sort_key(float x) {
    if (x == 0.0f) x = 0.0f;
    unsigned int b = __float_as_uint(x);
    return (unsigned long long)((b >> 31) ? ~b : (b | 0x80000000u));
}

#define OSD_NONE 0xFFFFu

// Column view of H for the OSD kernels: the CSC arrays in global memory, or a 16-bit copy of them in shared memory
// (osd_reg_kernel stages one per CTA: the candidate search walks a column per candidate, and two dependent L2 round
// trips per candidate were most of its time).
struct CscView {
    const int *ptr32, *row32;
    const uint16_t *ptr16, *row16;
    __device__ __forceinline__ int beg(int c) const { return ptr16 ? (int)ptr16[c] : ptr32[c]; }
    __device__ __forceinline__ int end(int c) const { return ptr16 ? (int)ptr16[c + 1] : ptr32[c + 1]; }
    __device__ __forceinline__ int row(int q) const { return row16 ? (int)row16[q] : row32[q]; }
};
__device__ __forceinline__ CscView csc_global(const GraphDev &g) { return CscView{g.col_ptr, g.row_idx, nullptr, nullptr}; }

// XOR of the T columns selected by H column c, word w
__device__ __forceinline__ uint32_t reduced_col_word(const CscView &cv, const uint32_t *Tc, int St, int c, int w) {
    uint32_t v = 0;
    for (int q = cv.beg(c), qe = cv.end(c); q < qe; q++) v ^= Tc[cv.row(q) * St + w];
    return v;
}
__device__ __forceinline__ uint32_t reduced_col_word(const GraphDev &g, const uint32_t *Tc, int St, int c, int w) {
    return reduced_col_word(csc_global(g), Tc, St, c, w);
}

// a11 read-out + a12-a14 candidate search + result write of one failed shot, shared by the OSD kernels.  On entry the
// elimination is finished: Tc holds the m x m row-operation matrix (column r at Tc + r*St, S words), `used` the pivot
// rows, `sprime` = (T s) & used, order/prow/np/nnp the column order, pivot row of every pivot column and the non-pivot
// positions in sorted order; the caller has synchronised the CTA.  wscr: nwarps * (S + 64) words of scratch.
template <typename real>
__device__ __forceinline__ void osd_readout_and_search(const OsdArgs<real> &a, long long shot, const double *weight, const uint32_t *Tc,
                                                       int St, const uint32_t *used, const uint32_t *sprime, uint32_t *wscr, double *red_w,
                                                       int *red_c, const uint16_t *order, const uint16_t *prow, const uint16_t *np, int nnp,
                                                       int *sh_best, int *sh_found, const CscView g) {
    const int n = a.g.n, S = a.S;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const long long base = shot * (long long)n;
    for (int j = tid; j < n; j += T) {
        const unsigned pr = prow[j];
        uint8_t x = (pr != OSD_NONE) ? (uint8_t)((sprime[pr >> 5] >> (pr & 31)) & 1u) : 0;
        if (a.osd0) a.osd0[base + j] = x;
        if (a.osdw) a.osdw[base + j] = x; // overwritten below if a candidate wins
    }
    if (tid == 0 && a.stat) atomicAdd(&a.stat[2], 1ull);

    const int wd = a.order;
    if (a.method == 0 || wd <= 0 || !a.osdw) return;

    // ---- a12-a14: candidate search.  Candidate -1 is OSD-0 itself. ----
    long long ncand = (a.method == 1) ? ((1ll << wd) - 1) : ((long long)nnp + (long long)wd * (wd - 1) / 2);
    uint32_t *s2 = wscr + (size_t)warp * (S + 64);
    int *selcol = reinterpret_cast<int *>(s2 + S);
    double bestW = 0;
    long long bestC = -2; // nothing yet
    for (long long c = -1 + warp; c < ncand; c += nwarps) {
        // decode the selection of non-pivot positions
        int nsel = 0;
        __syncwarp();
        if (c >= 0) {
            if (a.method == 1) {
                long long v = c + 1;
                for (int b = 0; b < wd; b++)
                    if ((v >> b) & 1) { if (lane == 0) selcol[nsel] = order[np[b]]; nsel++; }
            } else if (c < nnp) {
                if (lane == 0) selcol[0] = order[np[c]];
                nsel = 1;
            } else {
                int idx = (int)(c - nnp), i = 0;
                while (idx >= wd - 1 - i) { idx -= wd - 1 - i; i++; }
                if (lane == 0) { selcol[0] = order[np[i]]; selcol[1] = order[np[i + 1 + idx]]; }
                nsel = 2;
            }
        }
        __syncwarp();
        // s'' = s' + reduced images of the selected columns
        int pc = 0;
        for (int w = lane; w < S; w += 32) {
            uint32_t v = sprime[w];
            for (int q = 0; q < nsel; q++) v ^= reduced_col_word(g, Tc, St, selcol[q], w);
            v &= used[w];
            s2[w] = v;
            pc += __popc(v);
        }
        double W;
        if (a.uniform) {
            for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
            W = (double)(pc + nsel);
        } else {
            __syncwarp();
            W = 0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                int x = 0;
                if (j < n) {
                    const unsigned pr = prow[j];
                    if (pr != OSD_NONE) x = (s2[pr >> 5] >> (pr & 31)) & 1u;
                    else
                        for (int q = 0; q < nsel; q++) x |= (selcol[q] == j) ? 1 : 0;
                }
                unsigned mask = __ballot_sync(0xffffffffu, x);
                while (mask) { // ascending j, sequential fp64 accumulation (row a14)
                    const int b = __ffs(mask) - 1;
                    W += weight[j0 + b];
                    mask &= mask - 1;
                }
            }
        }
        if (bestC == -2 || W < bestW) { bestW = W; bestC = c; }
    }
    if (lane == 0) { red_w[warp] = bestW; red_c[warp] = (int)bestC; }
    __syncthreads();
    if (tid == 0) {
        double bw = 0; int bc = -2;
        for (int k = 0; k < nwarps; k++) {
            if (red_c[k] == -2) continue;
            if (bc == -2 || red_w[k] < bw || (red_w[k] == bw && red_c[k] < bc)) { bw = red_w[k]; bc = red_c[k]; }
        }
        *sh_best = bc;
    }
    __syncthreads();
    const int bc = *sh_best;
    if (bc < 0) return; // OSD-0 stands (strict '<' in the reference: ties keep the earlier)
    // rebuild the winner and write it
    if (warp == 0) {
        int nsel = 0;
        if (a.method == 1) {
            long long v = (long long)bc + 1;
            for (int b = 0; b < wd; b++)
                if ((v >> b) & 1) { if (lane == 0) selcol[nsel] = order[np[b]]; nsel++; }
        } else if (bc < nnp) {
            if (lane == 0) selcol[0] = order[np[bc]];
            nsel = 1;
        } else {
            int idx = bc - nnp, i = 0;
            while (idx >= wd - 1 - i) { idx -= wd - 1 - i; i++; }
            if (lane == 0) { selcol[0] = order[np[i]]; selcol[1] = order[np[i + 1 + idx]]; }
            nsel = 2;
        }
        __syncwarp();
        for (int w = lane; w < S; w += 32) {
            uint32_t v = sprime[w];
            for (int q = 0; q < nsel; q++) v ^= reduced_col_word(g, Tc, St, selcol[q], w);
            s2[w] = v & used[w];
        }
        if (lane == 0) *sh_found = nsel;
    }
    __syncthreads();
    {
        const int nsel = *sh_found;
        const uint32_t *w0 = wscr;
        const int *sel0 = reinterpret_cast<const int *>(w0 + S);
        for (int j = tid; j < n; j += T) {
            const unsigned pr = prow[j];
            int x = 0;
            if (pr != OSD_NONE) x = (w0[pr >> 5] >> (pr & 31)) & 1u;
            else
                for (int q = 0; q < nsel; q++) x |= (sel0[q] == j) ? 1 : 0;
            a.osdw[base + j] = (uint8_t)x;
        }
    }
}

template <typename real>
__global__ void __launch_bounds__(1024, 1) osd_kernel(OsdArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n, S = a.S, St = a.St;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int maxrank = m < n ? m : n;

    // shared-memory carve-up
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);            // n
    double *red_w = reinterpret_cast<double *>(keys + n);                                   // 32
    uint32_t *Tc = reinterpret_cast<uint32_t *>(red_w + 32);                                // m*St
    uint32_t *vmask = Tc + (size_t)m * St;                                                  // S
    uint32_t *used = vmask + S;                                                             // S
    uint32_t *sprime = used + S;                                                            // S
    uint32_t *wscr = sprime + S;                                                            // nwarps*(S+64)
    int *red_c = reinterpret_cast<int *>(wscr + (size_t)nwarps * (S + 64));                 // 32
    uint16_t *order = reinterpret_cast<uint16_t *>(red_c + 32);                             // n
    uint16_t *prow = order + n;                                                             // n
    uint16_t *np = prow + n;                                                                // n
    __shared__ int sh_found, sh_t, sh_rank, sh_nnp, sh_best;

    const int nfail = *a.fail_count;
    for (int f = blockIdx.x; f < nfail; f += gridDim.x) {
        const long long shot = a.fail_list[f];
        const real *llr = a.llr + (a.llr_by_shot ? shot : (long long)f) * n;
        const double *weight = a.weight + shot * a.weight_stride;
        __syncthreads();

        // ---- a9: stable ascending rank sort on (llr, index) ----
        for (int j = tid; j < n; j += T) { keys[j] = sort_key(llr[j]); prow[j] = OSD_NONE; }
        for (int r = tid; r < m; r += T) {
            for (int w = 0; w < St; w++) Tc[r * St + w] = 0;
            Tc[r * St + (r >> 5)] = 1u << (r & 31);
        }
        for (int w = tid; w < S; w += T) used[w] = 0;
        if (tid == 0) { sh_t = 0; sh_rank = 0; sh_nnp = 0; }
        __syncthreads();
        for (int j = tid; j < n; j += T) {
            const unsigned long long kj = keys[j];
            int rank = 0;
            for (int i = 0; i < n; i++) {
                const unsigned long long ki = keys[i];
                rank += (ki < kj || (ki == kj && i < j)) ? 1 : 0;
            }
            order[rank] = (uint16_t)j;
        }
        __syncthreads();

        // ---- a10: elimination, columns in sorted order.  T (m x m) is kept column-major:
        // Tc[r] = column r of T as an m-bit vector.  The reduced image of H column c is the XOR
        // of the T columns named by the rows of c; a pivot row p is the lowest unused row set in
        // it; the row operation "rows i in v\{p} += row p" is, column by column of T,
        // "if bit p of Tc[r] then Tc[r] ^= v\{p}".
        // Pivot search is speculative and parallel: every warp reduces one of the next nwarps columns against the
        // current T; the first of them (in sorted order) that has a 1 in an unused row is the next pivot, the ones
        // before it are dependent columns for good (they were tested against the same T), the ones after it are
        // simply tested again after the row operation.
        for (;;) {
            const int t0 = sh_t, rank0 = sh_rank, nnp0 = sh_nnp; // uniform: written before the last barrier
            if (t0 >= n || rank0 >= maxrank) break;
            const int lim = min(nwarps, n - t0);
            {
                int best = 0x7fffffff;
                uint32_t *mine = wscr + (size_t)warp * (S + 64);
                if (warp < lim) {
                    const int c = order[t0 + warp];
                    for (int w = lane; w < S; w += 32) {
                        const uint32_t v = reduced_col_word(g, Tc, St, c, w);
                        mine[w] = v;
                        const uint32_t cand = v & ~used[w];
                        if (cand && best == 0x7fffffff) best = w * 32 + __ffs(cand) - 1;
                    }
                    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                }
                if (lane == 0) red_c[warp] = best;
            }
            __syncthreads();
            int wi = 0;
            while (wi < lim && red_c[wi] == 0x7fffffff) wi++;
            if (tid < wi) np[nnp0 + tid] = (uint16_t)(t0 + tid);
            if (wi == lim) { // no pivot among these columns
                if (tid == 0) { sh_t = t0 + lim; sh_nnp = nnp0 + lim; }
                __syncthreads();
                continue;
            }
            const int p = red_c[wi], pw = p >> 5, pb = p & 31;
            const uint32_t *vm = wscr + (size_t)wi * (S + 64);
            if (tid == 0) { // nobody reads these again before the barrier below
                prow[order[t0 + wi]] = (uint16_t)p;
                used[pw] |= 1u << pb;
                sh_t = t0 + wi + 1; sh_rank = rank0 + 1; sh_nnp = nnp0 + wi;
            }
            const uint32_t keep = ~(1u << pb);
            for (int r = tid; r < m; r += T) {
                uint32_t *col = Tc + (size_t)r * St;
                if ((col[pw] >> pb) & 1u) {
                    int w = 0;
                    for (; w + 4 <= S; w += 4) { // four independent read-modify-writes in flight
                        const uint32_t c0 = col[w], c1 = col[w + 1], c2 = col[w + 2], c3 = col[w + 3];
                        const uint32_t v0 = vm[w], v1 = vm[w + 1], v2 = vm[w + 2], v3 = vm[w + 3];
                        col[w] = c0 ^ (w == pw ? v0 & keep : v0);
                        col[w + 1] = c1 ^ (w + 1 == pw ? v1 & keep : v1);
                        col[w + 2] = c2 ^ (w + 2 == pw ? v2 & keep : v2);
                        col[w + 3] = c3 ^ (w + 3 == pw ? v3 & keep : v3);
                    }
                    for (; w < S; w++) col[w] ^= (w == pw ? vm[w] & keep : vm[w]);
                }
            }
            __syncthreads();
        }
        // positions never examined (rank reached min(m,n)) are non-pivots, in order
        {
            const int t0 = sh_t, nnp0 = sh_nnp;
            for (int t = t0 + tid; t < n; t += T) np[nnp0 + (t - t0)] = (uint16_t)t;
        }
        const int nnp = sh_nnp + (n - sh_t);
        __syncthreads();

        // ---- a11: s' = T s, then OSD-0 read-out ----
        {
            uint32_t *mine = wscr + (size_t)warp * (S + 64);
            for (int w = lane; w < S; w += 32) {
                uint32_t acc = 0;
                for (int r = warp; r < m; r += nwarps)
                    if (synd_bit(a.synd, shot, m, r, a.synd_packed)) acc ^= Tc[(size_t)r * St + w];
                mine[w] = acc;
            }
            __syncthreads();
            for (int w = tid; w < S; w += T) {
                uint32_t acc = 0;
                for (int k = 0; k < nwarps; k++) acc ^= wscr[(size_t)k * (S + 64) + w];
                sprime[w] = acc & used[w];
            }
            __syncthreads();
        }
        osd_readout_and_search<real>(a, shot, weight, Tc, St, used, sprime, wscr, red_w, red_c, order, prow, np, nnp, &sh_best, &sh_found, csc_global(g));
    }
}

// ---------------------------------------------------------------------------------------------
// Large-H OSD-0 (rows a9-a11 when the m x m row-operation matrix does not fit in shared memory,
// BASELINE config 5: m = 19 200, n = 40 000).  One CTA per failed shot, left-looking panel
// Gauss-Jordan over GF(2):
//   * the sorted columns are consumed in panels of 32; the panel's bits, one 32-bit word per
//     check, are built in shared memory straight from the sparse H (the dense permuted matrix is
//     never materialised);
//   * before a panel is factorised, every earlier panel's row operations are replayed on it.  The
//     operations of panel q are stored in HBM as one 32-bit multiplier mask per check (bit k: "add
//     pivot row k of panel q to this check"), streamed back coalesced; the <= 32 pivot-row words
//     are resolved sequentially by one warp (shuffles), folded into four 256-entry XOR tables
//     (method of four Russians) and applied to all checks with four shared-memory look-ups each;
//   * factorising the panel scans its 32 columns in order: a column is a pivot iff an unused check
//     has a 1 in it (lowest such check is taken; the OSD-0 result does not depend on that choice,
//     row a10); the pivot word is added to every other check that has the bit (Jordan form, so the
//     transformed syndrome is the solution and no back-substitution is needed);
//   * the scan stops as soon as rank(H) pivots are found, so columns to the right of the last
//     pivot are never touched.
// Work per shot ~ (panels^2 / 2) * m mask words streamed from HBM/L2; memory per CTA
// panels * m * 4 bytes (96 MB for config 5).
// ---------------------------------------------------------------------------------------------
template <typename real>
struct OsdLargeArgs {
    GraphDev g;
    const uint8_t *synd;
    int synd_packed;
    const real *llr;
    int llr_by_shot;
    const int *fail_count;
    const int *fail_list;
    uint8_t *osd0, *osdw;
    unsigned long long *stat;
    int maxrank;       // rank(H): stop once this many pivots are found
    int npanels;       // ceil(n / 32)
    uint32_t *ws_mask; // [grid][npanels][m]
    int *ws_order;     // [grid][n]
    int *ws_piv_row;   // [grid][min(m,n)]
    int *ws_piv_pos;   // [grid][min(m,n)]
    int *ws_pstart;    // [grid][npanels + 1]
};

#define OSDL_UNUSED 0xFFFFu

template <typename real>
__global__ void __launch_bounds__(1024) osd0_large_kernel(OsdLargeArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    uint32_t *P = reinterpret_cast<uint32_t *>(smem_raw);                  // [m] panel word of every check
    uint32_t *tab = P + m;                                                 // [4][256] XOR tables
    uint32_t *Rk = tab + 1024;                                             // [32] resolved pivot-row words
    int *red = reinterpret_cast<int *>(Rk + 32);                           // [32] block-reduce scratch
    uint32_t *Mloc = reinterpret_cast<uint32_t *>(red + 32);               // [m] multiplier masks of the panel being factorised
    int *pstart_s = reinterpret_cast<int *>(Mloc + m);                     // [npanels + 1] first pivot of every panel
    uint16_t *rowpanel = reinterpret_cast<uint16_t *>(pstart_s + a.npanels + 1 + ((a.npanels + 1) & 1)); // [m] panel in which the check became a pivot
    uint8_t *s8 = reinterpret_cast<uint8_t *>(rowpanel + ((m + 1) & ~1));  // [m] transformed syndrome
    __shared__ int sh_p, sh_rank, sh_cnt;

    uint32_t *maskbase = a.ws_mask + (size_t)blockIdx.x * a.npanels * m;
    int *order = a.ws_order + (size_t)blockIdx.x * n;
    const int minmn = m < n ? m : n;
    int *piv_row = a.ws_piv_row + (size_t)blockIdx.x * minmn;
    int *piv_pos = a.ws_piv_pos + (size_t)blockIdx.x * minmn;
    int *pstart = a.ws_pstart + (size_t)blockIdx.x * (a.npanels + 1);

    const int nfail = *a.fail_count;
    for (int f = blockIdx.x; f < nfail; f += gridDim.x) {
        const long long shot = a.fail_list[f];
        const real *llr = a.llr + (a.llr_by_shot ? shot : (long long)f) * n;
        __syncthreads();

        // ---- a9: stable ascending rank sort on (llr, index); keys streamed from global (broadcast loads)
        for (int j0 = tid; j0 < n; j0 += 8 * T) {
            unsigned long long kj[8];
            int rk[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int j = j0 + u * T;
                kj[u] = (j < n) ? sort_key(llr[j]) : 0ull;
                rk[u] = 0;
            }
            for (int i = 0; i < n; i++) {
                const unsigned long long ki = sort_key(llr[i]);
#pragma unroll
                for (int u = 0; u < 8; u++) rk[u] += (ki < kj[u] || (ki == kj[u] && i < j0 + u * T)) ? 1 : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (j0 + u * T < n) order[rk[u]] = j0 + u * T;
        }
        for (int i = tid; i < m; i += T) { rowpanel[i] = OSDL_UNUSED; s8[i] = (uint8_t)synd_bit(a.synd, shot, m, i, a.synd_packed); }
        if (tid == 0) { sh_rank = 0; pstart_s[0] = 0; }
        __syncthreads();

        // ---- a10: elimination, one panel of 32 sorted columns at a time
        int npan_done = 0;
        for (int w = 0; w < a.npanels; w++) {
            if (sh_rank >= a.maxrank) break;
            uint32_t *Mw = maskbase + (size_t)w * m;
            for (int i = tid; i < m; i += T) { P[i] = 0; Mloc[i] = 0; }
            __syncthreads();
            if (tid < 32) {
                const int t = w * 32 + tid;
                if (t < n) {
                    const int j = order[t];
                    for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; q++) atomicOr(&P[g.row_idx[q]], 1u << tid);
                }
            }
            __syncthreads();
            // replay the row operations of every earlier panel that holds pivots, in order.  The pivot rows and
            // their masks of the next such panel are fetched by warp 0 one step ahead (two dependent HBM/L2 loads).
            int q = 0;
            while (q < w && pstart_s[q + 1] == pstart_s[q]) q++;
            int pf_pr = 0;
            uint32_t pf_mj = 0;
            if (warp == 0 && q < w) {
                const int ps0 = pstart_s[q], cnt0 = pstart_s[q + 1] - ps0;
                if (lane < cnt0) { pf_pr = piv_row[ps0 + lane]; pf_mj = (maskbase + (size_t)q * m)[pf_pr]; }
            }
            while (q < w) {
                const int ps = pstart_s[q], cnt = pstart_s[q + 1] - ps;
                int qn = q + 1;
                while (qn < w && pstart_s[qn + 1] == pstart_s[qn]) qn++;
                const uint32_t *Mq = maskbase + (size_t)q * m;
                if (warp == 0) {
                    const int pr = pf_pr;
                    const uint32_t mj = pf_mj;
                    uint32_t cur = lane < cnt ? P[pr] : 0u, R = 0;
                    for (int k = 0; k < cnt; k++) {
                        const uint32_t rk = __shfl_sync(0xffffffffu, cur, k);
                        if (lane == k) R = cur;
                        else if ((mj >> k) & 1u) cur ^= rk;
                    }
                    Rk[lane] = (lane < cnt) ? R : 0u;
                    __syncwarp();
                    // the parallel step below skips pivot rows of panel q, so their final words go in now
                    if (lane < cnt) P[pr] = cur;
                    pf_pr = 0; pf_mj = 0;
                    if (qn < w) {
                        const int psn = pstart_s[qn], cntn = pstart_s[qn + 1] - psn;
                        if (lane < cntn) { pf_pr = piv_row[psn + lane]; pf_mj = (maskbase + (size_t)qn * m)[pf_pr]; }
                    }
                }
                __syncthreads();
                for (int e = tid; e < 1024; e += T) {
                    uint32_t v = 0;
#pragma unroll
                    for (int b = 0; b < 8; b++) v ^= (((e & 255) >> b) & 1) ? Rk[(e >> 8) * 8 + b] : 0u;
                    tab[e] = v;
                }
                __syncthreads();
                // masks are streamed eight rows per thread at a time so the loads overlap
                for (int i0 = tid; i0 < m; i0 += 8 * T) {
                    uint32_t mi[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) { const int i = i0 + u * T; mi[u] = i < m ? __ldcs(Mq + i) : 0u; }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = i0 + u * T;
                        if (mi[u] && rowpanel[i] != (uint16_t)q)
                            P[i] ^= tab[mi[u] & 255u] ^ tab[256 + ((mi[u] >> 8) & 255u)] ^ tab[512 + ((mi[u] >> 16) & 255u)] ^ tab[768 + (mi[u] >> 24)];
                    }
                }
                __syncthreads();
                q = qn;
            }
            // factorise this panel
            if (tid == 0) sh_cnt = 0;
            __syncthreads();
            for (int c = 0; c < 32; c++) {
                const int t = w * 32 + c;
                if (t >= n || sh_rank >= a.maxrank) break;
                int best = 0x7fffffff;
                for (int i = tid; i < m; i += T)
                    if (((P[i] >> c) & 1u) && rowpanel[i] == OSDL_UNUSED) { best = i; break; }
                for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (lane == 0) red[warp] = best;
                __syncthreads();
                if (warp == 0) {
                    int b = (lane < nwarps) ? red[lane] : 0x7fffffff;
                    for (int o = 16; o > 0; o >>= 1) b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
                    if (lane == 0) sh_p = b;
                }
                __syncthreads();
                const int p = sh_p;
                if (p == 0x7fffffff) continue; // dependent column: not a pivot (uniform branch)
                const uint32_t Pp = P[p];
                const uint8_t sp = s8[p];
                const int k = sh_cnt;
                __syncthreads(); // everyone has read sh_p / sh_cnt / P[p] before they change
                for (int i = tid; i < m; i += T)
                    if (i != p && ((P[i] >> c) & 1u)) { P[i] ^= Pp; s8[i] ^= sp; Mloc[i] |= 1u << k; }
                if (tid == 0) {
                    const int r = sh_rank;
                    piv_row[r] = p; piv_pos[r] = t; rowpanel[p] = (uint16_t)w;
                    sh_rank = r + 1; sh_cnt = k + 1;
                }
                __syncthreads();
            }
            if (tid == 0) pstart_s[w + 1] = sh_rank;
            if (sh_cnt > 0) // this panel's masks go to HBM once, coalesced
                for (int i = tid; i < m; i += T) Mw[i] = Mloc[i];
            npan_done = w + 1;
            __syncthreads();
        }
        (void)npan_done;

        // ---- a11: OSD-0 read-out.  Jordan form: x[pivot column of check p] = transformed syndrome bit of p
        const long long base = shot * (long long)n;
        for (int j = tid; j < n; j += T) {
            if (a.osd0) a.osd0[base + j] = 0;
            if (a.osdw) a.osdw[base + j] = 0;
        }
        __syncthreads();
        const int rank = sh_rank;
        for (int r = tid; r < rank; r += T) {
            const uint8_t x = s8[piv_row[r]];
            if (x) {
                const int j = order[piv_pos[r]];
                if (a.osd0) a.osd0[base + j] = 1;
                if (a.osdw) a.osdw[base + j] = 1;
            }
        }
        if (tid == 0 && a.stat) atomicAdd(&a.stat[2], 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// Harness step kernels (rows a17-a19)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct SampleArgs {
    GraphDev g;
    const uint32_t *t1, *t2, *t3;
    unsigned long long seed, shot0;
    long long B;
    int sector; // 0: X component, 1: Z component
    uint8_t *errors;
    uint8_t *synd;
    int synd_packed; // 1: write ceil(m/8) bytes per shot (bit i%8 of byte i/8 = check i) instead of m
};

// one CTA per shot (grid-stride): errors staged in shared memory, one thread per check for H e
__global__ void __launch_bounds__(1024) sample_syndrome_kernel(SampleArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *e_s = smem_raw;
    const int n = a.g.n, m = a.g.m, tid = threadIdx.x, T = blockDim.x;
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        const unsigned long long gshot = a.shot0 + (unsigned long long)b;
        __syncthreads();
        for (int q = tid; q * 4 < n; q += T) {
            uint32_t c[4] = {(uint32_t)q, 0u, (uint32_t)gshot, (uint32_t)(gshot >> 32)};
            philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const int j = q * 4 + l;
                if (j < n) {
                    const uint32_t r = c[l];
                    const uint32_t u1 = a.t1[j], u2 = a.t2[j], u3 = a.t3[j];
                    const int z = (r < u1) || (r >= u2 && r < u3);
                    const int x = (r >= u1 && r < u3);
                    const uint8_t e = (uint8_t)(a.sector ? z : x);
                    e_s[j] = e;
                    if (a.errors) a.errors[b * n + j] = e;
                }
            }
        }
        __syncthreads();
        if (!a.synd_packed) {
            for (int i = tid; i < m; i += T) {
                int acc = 0;
                for (int p = a.g.row_ptr[i]; p < a.g.row_ptr[i + 1]; p++) acc ^= e_s[a.g.col_idx[p]];
                a.synd[b * m + i] = (uint8_t)acc;
            }
        } else {
            // bit-packed H e mod 2: a warp ballots 32 checks at a time, every eighth lane stores one byte
            const long long mb = (m + 7) >> 3;
            for (int i0 = 0; i0 < m; i0 += T) {
                const int i = i0 + tid;
                int acc = 0;
                if (i < m)
                    for (int p = a.g.row_ptr[i]; p < a.g.row_ptr[i + 1]; p++) acc ^= e_s[a.g.col_idx[p]];
                const unsigned bal = __ballot_sync(0xffffffffu, acc & 1);
                if ((tid & 7) == 0 && i < m) a.synd[b * mb + (i >> 3)] = (uint8_t)(bal >> (tid & 31));
            }
        }
    }
}

struct LogicalArgs {
    int n, K;
    const int *l_ptr, *l_idx;
    const uint8_t *errors, *dec;
    long long B;
    uint8_t *fail;
    unsigned long long *fail_count;
    int *min_weight;
    int *resid_weight; // [B] Hamming weight of e ^ d per shot (may be NULL)
};

// one warp per shot: lanes split the logical rows; a shot fails if any row has odd overlap
__global__ void __launch_bounds__(256) logical_check_kernel(LogicalArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long nfail = 0;
    int minw = 0x7fffffff;
    for (long long b = warp0; b < a.B; b += nw) {
        const uint8_t *e = a.errors + b * a.n, *d = a.dec + b * a.n;
        int any = 0;
        for (int r = lane; r < a.K; r += 32) {
            int acc = 0;
            for (int p = a.l_ptr[r]; p < a.l_ptr[r + 1]; p++) {
                const int j = a.l_idx[p];
                acc ^= (e[j] ^ d[j]) & 1;
            }
            any |= acc;
        }
        any = __any_sync(0xffffffffu, any);
        if ((any && a.min_weight) || a.resid_weight) {
            int wsum = 0;
            for (int j = lane; j < a.n; j += 32) wsum += (e[j] ^ d[j]) & 1;
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            if (any) minw = min(minw, wsum);
            if (a.resid_weight && lane == 0) a.resid_weight[b] = wsum;
        }
        if (lane == 0) {
            if (a.fail) a.fail[b] = (uint8_t)any;
            nfail += any ? 1 : 0;
        }
    }
    if (lane == 0) {
        if (a.fail_count && nfail) atomicAdd(a.fail_count, nfail);
        if (a.min_weight && minw != 0x7fffffff) atomicMin(a.min_weight, minw);
    }
}

// ---------------------------------------------------------------------------------------------
// Harness kernels for the two-sector CSS simulation (css_decode_sim.py:207-248, 250-365)
// ---------------------------------------------------------------------------------------------
// Per-shot channel update: the second sector's probability of qubit j is one of two host-computed
// values, selected by the first sector's decoding bit.  Emits BP priors log((1-p)/p) and OSD
// weights log(1/p); the logarithms are taken on the host (glibc) so the values are the ones
// update_channel_probs would have produced.
template <typename real>
__global__ void __launch_bounds__(256) channel_update_kernel(const uint8_t *__restrict__ first, long long total, int n,
                                                             const double *__restrict__ prior0, const double *__restrict__ prior1,
                                                             const double *__restrict__ w0, const double *__restrict__ w1,
                                                             real *__restrict__ priors, double *__restrict__ weights) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % n);
        const bool one = first[i] != 0;
        priors[i] = (real)(one ? prior1[j] : prior0[j]);
        if (weights) weights[i] = one ? w1[j] : w0[j];
    }
}

// [B, n] bytes (0/1) -> [B, ceil(n/8)] bytes, bit j%8 of byte j/8 = entry j (the layout of numpy.packbits(bitorder="little"))
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t *__restrict__ src, long long B, int n, uint8_t *__restrict__ dst) {
    const long long nb = (n + 7) >> 3, total = B * nb;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / nb;
        const int jb = (int)(idx - b * nb);
        const uint8_t *row = src + b * n + jb * 8;
        const int cnt = min(8, n - jb * 8);
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < cnt) v |= (unsigned)(row[k] & 1) << k;
        dst[idx] = (uint8_t)v;
    }
}

// INT32 logic-op calibration for the OSD roofline ("measure a LOP3 microbenchmark peak in the
// same run"): eight independent LOP3 chains per thread, 8 * iters logic ops per thread.
__global__ void __launch_bounds__(256) lop3_peak_kernel(uint32_t *out, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = a[j] ^ (a[(j + 1) & 7] & a[(j + 3) & 7]); // one LOP3 each
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= a[j];
    if (r == 0x12345678u) out[0] = r; // keeps the chains alive
}

// fp64 calibration for the product-sum roofline: eight independent DFMA chains per thread, 8 * iters fused multiply-adds
// per thread (product-sum spends its instructions on tanh / log / divisions, i.e. on the fp64 pipe, not on memory).
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = 1.0 + 1e-9 * (threadIdx.x + 8 * j + blockIdx.x);
    const double m = 0.99999999, c = 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = __fma_rn(a[j], m, c); // one DFMA each
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r += a[j];
    if (r == 0.123456789) out[0] = r; // keeps the chains alive
}

// Shared-memory bandwidth calibration for the BP roofline: the message traffic of the in-place BP kernels is 16-byte
// LDS / STS in equal parts, so the peak is measured the same way -- every thread loads, modifies and stores 16-byte
// words of its own (lane-consecutive, i.e. conflict free: four 128-byte wavefronts per warp instruction), four
// independent words per trip so that the loop is bandwidth and not latency bound.
__global__ void __launch_bounds__(1024) smem_peak_kernel(float *out, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned base = (unsigned)__cvta_generic_to_shared(smem_raw) + threadIdx.x * 16u;
    const unsigned stride = blockDim.x * 16u;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = make_float4(threadIdx.x, k, 1.f, 2.f);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + k * stride), "f"(v[k].x), "f"(v[k].y), "f"(v[k].z), "f"(v[k].w) : "memory");
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "r"(base + k * stride) : "memory");
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k].x += 1.f;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + k * stride), "f"(v[k].x), "f"(v[k].y), "f"(v[k].z), "f"(v[k].w) : "memory");
        }
    }
    if (v[0].x + v[1].x + v[2].x + v[3].x == -1.f) out[0] = v[0].y;
}

struct CssSector {
    const uint8_t *fail_x, *fail_z; // [B] logical X / Z failure flags of one decoding (osdw, osd0 or bp)
    const int *weight_x, *weight_z; // [B] residual weights (may be NULL: no min-weight tracking)
};

// counters: [0] shots, [1] bp_converge_x, [2] bp_converge_z, [3] bp_success, [4] osd0_success, [5] osdw_success,
// [6] minimum weight of a failing residual (X checked first, Z only if X passed: the reference's `elif`).
__global__ void __launch_bounds__(256) css_counters_kernel(long long B, CssSector osdw, CssSector osd0, CssSector bp,
                                                           const uint8_t *conv_x, const uint8_t *conv_z,
                                                           unsigned long long *counters, int *min_weight) {
    unsigned long long c[5] = {0, 0, 0, 0, 0};
    int minw = 0x7fffffff;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const bool cx = conv_x[b] != 0, cz = conv_z[b] != 0;
        c[0] += cx; c[1] += cz;
        if (cx && cz && !bp.fail_x[b] && !bp.fail_z[b]) c[2]++;
        if (osd0.fail_x[b]) { if (osd0.weight_x) minw = min(minw, osd0.weight_x[b]); }
        else if (osd0.fail_z[b]) { if (osd0.weight_z) minw = min(minw, osd0.weight_z[b]); }
        else c[3]++;
        if (osdw.fail_x[b]) { if (osdw.weight_x) minw = min(minw, osdw.weight_x[b]); }
        else if (osdw.fail_z[b]) { if (osdw.weight_z) minw = min(minw, osdw.weight_z[b]); }
        else c[4]++;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
        if ((threadIdx.x & 31) == 0 && c[k]) atomicAdd(&counters[1 + k], c[k]);
    }
    for (int o = 16; o > 0; o >>= 1) minw = min(minw, __shfl_xor_sync(0xffffffffu, minw, o));
    if ((threadIdx.x & 31) == 0 && minw != 0x7fffffff) atomicMin(min_weight, minw);
}

#endif // BPOSD_KERNELS_COMMON_ONLY

} // namespace bposd
