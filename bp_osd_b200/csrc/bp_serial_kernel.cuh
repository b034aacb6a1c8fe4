// bp_serial_kernel.cuh -- the serial BP schedule of ldpc (`schedule="serial"`, an option of
// ldpc.BpOsdDecoder that /root/reference never passes; the semantics are those of upstream ldpc's bp_decode_serial).
//
// The serial schedule visits the bits one after the other inside an iteration: bit j recomputes the check-to-bit message
// of each of its edges from the CURRENT bit-to-check messages of the check's other edges, sums them into its LLR, takes
// its hard decision and refreshes its own bit-to-check messages before bit j+1 starts.  A shot is therefore one long
// dependency chain (n steps per iteration) -- there is nothing to spread over a CTA.  The parallel axis is the shot:
//
//   * ONE THREAD PER SHOT.  Every shot walks the same Tanner graph in the same order, so the 32 shots of a warp run in
//     lockstep with no divergence (a lane that has converged idles until its warp is done);
//   * messages live in an HBM workspace laid out EDGE-MAJOR per warp -- element e of lane l at [e * 32 + l] -- so every
//     message access of a warp is one fully used 256-byte (fp64) line, and the lines a bit step touches (its checks'
//     rows) are re-used from L1 / L2 by the neighbouring bit steps;
//   * convergence is tracked incrementally: one mismatch bit per check and a count of unsatisfied checks, updated when a
//     hard decision flips; tested at the end of every iteration like the reference does;
//   * results, failed-shot list and LLR workspace follow the protocol of the other BP kernels, so OSD runs unchanged.
//
// This is the compatibility path for an option the reference cannot reach, not a tuned kernel: any degrees, any size,
// min-sum and product-sum, fp64 bit-exact with the oracle.
#pragma once
#include "bposd_kernels.cuh"

namespace bposd {

// bytes of workspace one warp needs: b2c, c2b [E] and llr [n] reals, decision [n] and check state [m] bytes, x 32 lanes
template <typename real>
static inline size_t serial_warp_bytes(int m, int n, int E) {
    return (size_t)32 * ((2 * (size_t)E + (size_t)n) * sizeof(real) + (size_t)n + (size_t)m);
}

template <typename real>
__global__ void __launch_bounds__(128) bp_serial_kernel(BpArgs<real> a, const int *__restrict__ order, unsigned char *__restrict__ ws,
                                                        size_t warp_bytes) {
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n, E = g.nnz;
    const int lane = threadIdx.x & 31;
    const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned char *base = ws + gw * warp_bytes;
    real *b2c = reinterpret_cast<real *>(base) + lane;   // element e at b2c[e * 32]
    real *c2b = b2c + (size_t)E * 32;
    real *llr_s = c2b + (size_t)E * 32;
    uint8_t *dec = reinterpret_cast<uint8_t *>(llr_s - lane + (size_t)n * 32) + lane; // decision of bit j at dec[j * 32]
    uint8_t *chk = dec + (size_t)n * 32;                 // check i: bit0 parity mismatch, bit1 syndrome
    unsigned long long n_conv = 0, n_iter = 0;

    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(a.queue, 32ull);
        first = __shfl_sync(0xffffffffu, first, 0);
        if ((long long)first >= a.B) break;
        const long long shot = (long long)first + lane;
        const bool live = shot < a.B;
        const real *prior = a.prior + (live ? shot : 0) * a.prior_stride;

        int nbad = 0;
        for (int i = 0; i < m; i++) {
            const unsigned s = live ? synd_bit(a.synd, shot, m, i, a.synd_packed) : 0u;
            chk[(size_t)i * 32] = (uint8_t)(s | (s << 1));
            nbad += (int)s;
        }
        for (int j = 0; j < n; j++) { // a3
            const real p = prior[j];
            dec[(size_t)j * 32] = 0;
            llr_s[(size_t)j * 32] = p;
            for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; q++) b2c[(size_t)g.csc_slot[q] * 32] = p;
        }

        bool conv = false, done = !live;
        int iters = 0;
        for (int it = 1; it <= a.max_iter; it++) {
            if (__all_sync(0xffffffffu, done)) break;
            const real alpha = ms_alpha(a.alpha0, it);
            if (!done) {
                for (int k = 0; k < n; k++) {
                    const int j = order ? order[k] : k;
                    const int qb = g.col_ptr[j], qe = g.col_ptr[j + 1];
                    real llr = prior[j];
                    for (int q = qb; q < qe; q++) {
                        const int e = g.csc_slot[q], i = g.row_idx[q];
                        const unsigned syn = (unsigned)(chk[(size_t)i * 32] >> 1) & 1u;
                        real c;
                        if (a.method == 1) {
                            int sgn = (int)syn;
                            real t = real_max<real>();
                            for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; x++) {
                                if (x == e) continue;
                                const real v = b2c[(size_t)x * 32];
                                const real av = r_abs(v);
                                if (av < t) t = av;
                                sgn += (v <= 0) ? 1 : 0;
                            }
                            c = alpha * t;
                            c = (sgn & 1) ? -c : c;
                        } else {
                            real pr = 1;
                            for (int x = g.row_ptr[i]; x < g.row_ptr[i + 1]; x++) {
                                if (x == e) continue;
                                pr *= r_tanh(b2c[(size_t)x * 32] / 2);
                            }
                            pr = ps_clamp(pr);
                            c = r_log(ps_ratio(pr));
                            c = syn ? -c : c;
                        }
                        c2b[(size_t)e * 32] = c;
                        b2c[(size_t)e * 32] = llr;
                        llr += c;
                    }
                    llr_s[(size_t)j * 32] = llr;
                    const uint8_t d = (llr <= 0) ? 1 : 0;
                    if (d != dec[(size_t)j * 32]) {
                        dec[(size_t)j * 32] = d;
                        for (int q = qb; q < qe; q++) {
                            uint8_t &w = chk[(size_t)g.row_idx[q] * 32];
                            w ^= 1;
                            nbad += (w & 1) ? 1 : -1;
                        }
                    }
                    real t = 0;
                    for (int q = qe - 1; q >= qb; q--) {
                        const int e = g.csc_slot[q];
                        b2c[(size_t)e * 32] += t;
                        t += c2b[(size_t)e * 32];
                    }
                }
                iters = it;
                if (nbad == 0) { conv = true; done = true; }
            }
        }

        if (live) {
            const bool final_here = conv || a.osd_off;
            int slot = 0;
            if (!final_here) {
                slot = atomicAdd(a.fail_count, 1);
                a.fail_list[slot] = (int)shot;
            }
            const long long ob = shot * (long long)n;
            for (int j = 0; j < n; j++) {
                const uint8_t d = dec[(size_t)j * 32];
                const real l = llr_s[(size_t)j * 32];
                if (a.bp) a.bp[ob + j] = d;
                if (final_here) {
                    if (a.osd0) a.osd0[ob + j] = d;
                    if (a.osdw) a.osdw[ob + j] = d;
                }
                if (a.llr) a.llr[ob + j] = l;
                else if (!final_here) a.fail_llr[(long long)slot * n + j] = l;
            }
            if (a.converge) a.converge[shot] = conv ? 1 : 0;
            if (a.iter) a.iter[shot] = iters;
            n_conv += conv ? 1 : 0;
            n_iter += (unsigned long long)iters;
        }
    }
    if (a.stat && (n_conv | n_iter)) {
        atomicAdd(&a.stat[0], n_conv);
        atomicAdd(&a.stat[1], n_iter);
    }
}

} // namespace bposd
