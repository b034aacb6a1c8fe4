// osd_panel_kernel.cuh -- OSD-0 / OSD-E / OSD-CS (rows a9-a14) by left-looking panel elimination in shared memory.
//
// One CTA per BP-failed shot.  The row-operation history of the GF(2) Gauss-Jordan elimination is kept as
// *pivot blocks*: every 32 consecutive pivots share one 32-bit multiplier mask per check ("add pivot row k
// of the block to this check"), ceil(rank/32) * m words in all (m^2/8 bytes: 119 KB for the [[1922,50,16]]
// code).  Sorted columns are consumed in panels of 32: the panel's bits (one word per check) are built from
// the sparse H, every pivot block is replayed on it -- the <= 32 pivot-row words are resolved with the
// block's inverse unit-triangular matrix (32 independent shuffles instead of a 32-step dependent chain),
// folded into four 256-entry XOR tables (method of four Russians) and applied with four look-ups per check
// -- and the panel is then factorised column by column (warp-reduced, one atomicMin, one barrier per
// column).  The scan stops at rank(H) pivots.  Compared with keeping the m x m transformation matrix and
// sweeping it once per pivot (osd_kernel) this moves ~20x fewer shared-memory words per shot.
//
// Candidates (rows a12-a14) are evaluated 32 at a time: a panel of non-pivot columns is reduced by the same
// replay, after which bit c of check i is the reduced column c at row i; weights are vertical popcounts
// (uniform channel) or ordered fp64 sums (non-uniform channel), the first minimum wins (strict '<').
#pragma once
#include "bposd_kernels.cuh"

namespace bposd {

template <typename real>
struct OsdPanelArgs {
    GraphDev g;
    int S;        // 32-bit words per m-bit vector
    int nb;       // pivot blocks: ceil(min(m, n) / 32)
    int maxrank;  // rank(H)
    int method;   // 0 osd0, 1 osd_e, 2 osd_cs
    int order;    // search depth w
    int uniform;  // 1: all channel probabilities equal and in (0,1): weight = popcount
    const double *weight;
    long long weight_stride;
    const uint8_t *synd;
    const real *llr;
    int llr_by_shot;
    const int *fail_count;
    const int *fail_list;
    uint8_t *osd0, *osdw;
    unsigned long long *stat;
};

static inline size_t osd_panel_smem_bytes(int m, int n, int threads) {
    const size_t S = (m + 31) / 32, nb = (std::min(m, n) + 31) / 32, nw = threads / 32;
    size_t b = std::max((size_t)nb * m * 4, (size_t)n * 8); // masks (sort keys alias them)
    b += 3 * (size_t)m * 4;                                // P, P0, P1
    b += 1024 * 4 + nb * 32 * 4;                           // tables, Linv
    b += nw * (S + 64) * 4 + 2 * S * 4;                    // per-warp scratch, sp, usedw
    b += 64 * 4 + 32 * 8 + 32 * 4 + 64;                    // cntW, red_w, red_c, pair lists
    b += 3 * (size_t)n * 2 + 2 * nb * 32 * 2;              // order, np, prow, piv_row, piv_pos
    b += 2 * (size_t)m + nb + 64;                          // s8, rowblk, linv_cnt
    return b + 64;
}

#define OSDP_NONE8 0xFFu

template <typename real>
__global__ void __launch_bounds__(1024) osd_panel_kernel(OsdPanelArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n, S = a.S, nb = a.nb;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    size_t mask_bytes = (size_t)nb * m * 4;
    if (mask_bytes < (size_t)n * 8) mask_bytes = (size_t)n * 8;
    mask_bytes = (mask_bytes + 15) / 16 * 16;
    uint32_t *M = reinterpret_cast<uint32_t *>(smem_raw);                                   // [nb][m]
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);            // [n], alias of M
    uint32_t *P = reinterpret_cast<uint32_t *>(smem_raw + mask_bytes);                      // [m]
    uint32_t *P0 = P + m, *P1 = P0 + m;                                                     // saved candidate panels
    uint32_t *tab = P1 + m;                                                                 // [4][256]
    uint32_t *Linv = tab + 1024;                                                            // [nb][32]
    uint32_t *wscr = Linv + (size_t)nb * 32;                                                // [nwarps][S + 64]
    uint32_t *sp = wscr + (size_t)nwarps * (S + 64);                                        // [S] s' on used rows
    uint32_t *usedw = sp + S;                                                               // [S]
    int *cntW = reinterpret_cast<int *>(usedw + S);                                         // [64]
    double *red_w = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(cntW + 64) + 7) & ~(uintptr_t)7); // [32]
    int *red_c = reinterpret_cast<int *>(red_w + 32);                                       // [32]
    uint8_t *pair_i = reinterpret_cast<uint8_t *>(red_c + 32), *pair_j = pair_i + 32;       // [32] each
    uint16_t *order = reinterpret_cast<uint16_t *>(pair_j + 32);                            // [n]
    uint16_t *np = order + n;                                                               // [n] non-pivot positions, sorted order
    uint16_t *prow = np + n;                                                                // [n] pivot row of a column
    uint16_t *piv_row = prow + n;                                                           // [nb*32]
    uint16_t *piv_pos = piv_row + (size_t)nb * 32;                                          // [nb*32]
    uint8_t *s8 = reinterpret_cast<uint8_t *>(piv_pos + (size_t)nb * 32);                   // [m] transformed syndrome
    uint8_t *rowblk = s8 + m;                                                               // [m] block in which the check became a pivot
    uint8_t *linv_cnt = rowblk + m;                                                         // [nb] pivots Linv[b] was built for
    __shared__ int sh_t, sh_rank, sh_nnp, sh_nblk, sh_cnt, sh_min[3], sh_best;
    __shared__ uint32_t Rk[32];

    // ---- replay every pivot block on the panel in P (block-wide; ends with a barrier) ----
    auto replay_all = [&]() {
        const int nblk = sh_nblk, lastcnt = sh_cnt;
        const int total = nblk + (lastcnt > 0 ? 1 : 0);
        for (int q = 0; q < total; q++) {
            const int cnt = q < nblk ? 32 : lastcnt;
            const uint32_t *Mq = M + (size_t)q * m;
            if (warp == 0) {
                const int pr = lane < cnt ? piv_row[q * 32 + lane] : 0;
                const uint32_t mj = lane < cnt ? Mq[pr] : 0u;
                if (linv_cnt[q] != cnt) { // block grew since its inverse was built: one dependent pass on the identity
                    uint32_t cur = lane < cnt ? (1u << lane) : 0u, R = 0;
                    for (int k = 0; k < cnt; k++) {
                        const uint32_t rk = __shfl_sync(0xffffffffu, cur, k);
                        if (lane == k) R = cur;
                        else if ((mj >> k) & 1u) cur ^= rk;
                    }
                    Linv[q * 32 + lane] = lane < cnt ? R : 0u;
                    if (lane == 0) linv_cnt[q] = (uint8_t)cnt;
                }
                const uint32_t li = lane < cnt ? Linv[q * 32 + lane] : 0u;
                const uint32_t praw = lane < cnt ? P[pr] : 0u;
                uint32_t R = 0;
                for (int j = 0; j < cnt; j++) { // R = Linv * Praw: independent shuffles
                    const uint32_t v = __shfl_sync(0xffffffffu, praw, j);
                    R ^= ((li >> j) & 1u) ? v : 0u;
                }
                uint32_t cur = praw;
                for (int k = 0; k < cnt; k++) { // final words of the block's own pivot rows
                    const uint32_t rk = __shfl_sync(0xffffffffu, R, k);
                    cur ^= (k != lane && ((mj >> k) & 1u)) ? rk : 0u;
                }
                Rk[lane] = lane < cnt ? R : 0u;
                if (lane < cnt) P[pr] = cur;
            }
            __syncthreads();
            for (int e = tid; e < 1024; e += T) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) v ^= (((e & 255) >> b) & 1) ? Rk[(e >> 8) * 8 + b] : 0u;
                tab[e] = v;
            }
            __syncthreads();
            for (int i = tid; i < m; i += T) {
                const uint32_t mi = Mq[i];
                if (mi && rowblk[i] != (uint8_t)q)
                    P[i] ^= tab[mi & 255u] ^ tab[256 + ((mi >> 8) & 255u)] ^ tab[512 + ((mi >> 16) & 255u)] ^ tab[768 + (mi >> 24)];
            }
            __syncthreads();
        }
    };
    // ---- panel of <= 32 sorted positions pos[0..ncol) -> P (block-wide; ends with a barrier) ----
    auto build_panel = [&](const uint16_t *pos_list, int first_pos, int ncol) {
        for (int i = tid; i < m; i += T) P[i] = 0;
        __syncthreads();
        if (tid < ncol) {
            const int t = pos_list ? pos_list[tid] : first_pos + tid;
            const int j = order[t];
            for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; q++) atomicOr(&P[g.row_idx[q]], 1u << tid);
        }
        __syncthreads();
    };

    const int nfail = *a.fail_count;
    for (int f = blockIdx.x; f < nfail; f += gridDim.x) {
        const long long shot = a.fail_list[f];
        const real *llr = a.llr + (a.llr_by_shot ? shot : (long long)f) * n;
        const uint8_t *synd = a.synd + shot * m;
        const double *weight = a.weight + shot * a.weight_stride;
        __syncthreads();

        // ---- a9: stable ascending rank sort on (llr, index) ----
        for (int j = tid; j < n; j += T) { keys[j] = sort_key(llr[j]); prow[j] = OSD_NONE; }
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
        for (size_t e = tid; e < (size_t)nb * m; e += T) M[e] = 0;
        for (int i = tid; i < m; i += T) { rowblk[i] = OSDP_NONE8; s8[i] = synd[i] & 1; }
        for (int b = tid; b < nb; b += T) linv_cnt[b] = 0;
        if (tid == 0) { sh_t = 0; sh_rank = 0; sh_nnp = 0; sh_nblk = 0; sh_cnt = 0; sh_min[0] = sh_min[1] = sh_min[2] = 0x7fffffff; }
        __syncthreads();

        // ---- a10: elimination, one panel of 32 sorted columns at a time ----
        while (sh_t < n && sh_rank < a.maxrank) {
            const int t0 = sh_t, ncol = min(32, n - t0);
            build_panel(nullptr, t0, ncol);
            replay_all();
            int c = 0;
            for (; c < ncol; c++) {
                if (sh_rank >= a.maxrank) break;
                int best = 0x7fffffff;
                for (int i = tid; i < m; i += T)
                    if (((P[i] >> c) & 1u) && rowblk[i] == OSDP_NONE8) { best = i; break; }
                best = __reduce_min_sync(0xffffffffu, best);
                int *slot = &sh_min[c % 3];
                if (lane == 0 && best != 0x7fffffff) atomicMin(slot, best);
                __syncthreads();
                const int p = *slot;
                if (tid == 0) sh_min[(c + 2) % 3] = 0x7fffffff; // next used two columns from now, two barriers away
                if (p == 0x7fffffff) { // dependent column
                    if (tid == 0) { np[sh_nnp] = (uint16_t)(t0 + c); sh_nnp++; }
                    continue;
                }
                const uint32_t Pp = P[p];
                const uint8_t spv = s8[p];
                const int b = sh_nblk, k = sh_cnt;
                uint32_t *Mb = M + (size_t)b * m;
                __syncthreads(); // everyone holds p, Pp, b, k before they change
                for (int i = tid; i < m; i += T)
                    if (i != p && ((P[i] >> c) & 1u)) { P[i] ^= Pp; s8[i] ^= spv; Mb[i] |= 1u << k; }
                if (tid == 0) {
                    prow[order[t0 + c]] = (uint16_t)p;
                    rowblk[p] = (uint8_t)b;
                    piv_row[b * 32 + k] = (uint16_t)p;
                    piv_pos[b * 32 + k] = (uint16_t)(t0 + c);
                    sh_rank++;
                    if (k == 31) { sh_nblk = b + 1; sh_cnt = 0; } else sh_cnt = k + 1;
                }
                __syncthreads();
            }
            __syncthreads();
            if (tid == 0) { sh_t = t0 + c; sh_min[0] = sh_min[1] = sh_min[2] = 0x7fffffff; }
            __syncthreads();
        }
        // positions never examined are non-pivots, in order
        {
            const int t1 = sh_t, nnp1 = sh_nnp;
            for (int t = t1 + tid; t < n; t += T) np[nnp1 + (t - t1)] = (uint16_t)t;
        }
        const int nnp = sh_nnp + (n - sh_t);
        __syncthreads();

        // ---- a11: OSD-0 read-out (Jordan form: the transformed syndrome is the solution on the pivot rows) ----
        for (int w = warp; w < S; w += nwarps) {
            const int i = w * 32 + lane;
            const bool u = i < m && rowblk[i] != OSDP_NONE8;
            const unsigned ub = __ballot_sync(0xffffffffu, u);
            const unsigned sb = __ballot_sync(0xffffffffu, u && s8[i]);
            if (lane == 0) { usedw[w] = ub; sp[w] = sb; }
        }
        __syncthreads();
        const long long base = shot * (long long)n;
        for (int j = tid; j < n; j += T) {
            const unsigned pr = prow[j];
            const uint8_t x = (pr != OSD_NONE) ? s8[pr] : 0;
            if (a.osd0) a.osd0[base + j] = x;
            if (a.osdw) a.osdw[base + j] = x; // overwritten below if a candidate wins
        }
        if (tid == 0 && a.stat) atomicAdd(&a.stat[2], 1ull);

        const int wd = min(a.order, nnp);
        if (a.method == 0 || wd <= 0 || !a.osdw) continue;

        // ---- a12-a14: candidate search ----
        // A candidate is a set of selected non-pivot positions; its solution on check i is
        // x_i = s'_i ^ (parity of the selected reduced columns at row i), and it costs |x| + |selection| (uniform
        // channel) or the ordered sum of log(1/p_j) over its support (row a14).  Candidate -1 is OSD-0.
        uint32_t *s2 = wscr + (size_t)warp * (S + 64);
        int *selcol = reinterpret_cast<int *>(s2 + S);
        // ordered soft weight of the candidate whose row solution is packed in s2 and whose nsel selected columns are in selcol
        auto soft_weight = [&](int nsel) -> double {
            double W = 0;
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
                while (mask) { // ascending j, sequential fp64 accumulation
                    const int b = __ffs(mask) - 1;
                    W += weight[j0 + b];
                    mask &= mask - 1;
                }
            }
            return W;
        };
        double bestW = 0;        // per-warp running best (uniform: integer valued)
        long long bestC = -2;    // -2: none yet, -1: OSD-0
        if (warp == 0) { // candidate -1
            if (a.uniform) { int pc = 0; for (int w = lane; w < S; w += 32) pc += __popc(sp[w]); pc = __reduce_add_sync(0xffffffffu, pc); bestW = pc; }
            else { for (int w = lane; w < S; w += 32) s2[w] = sp[w]; __syncwarp(); bestW = soft_weight(0); }
            bestC = -1;
        }
        const long long npairs = (long long)wd * (wd - 1) / 2;
        if (a.method == 2) {
            // weight-1 sweep over all non-pivot columns, 32 per panel; the first two panels are kept for the pairs
            for (int gidx = 0; gidx * 32 < nnp; gidx++) {
                const int ncol = min(32, nnp - gidx * 32);
                build_panel(np + gidx * 32, 0, ncol);
                replay_all();
                if (gidx < 2) {
                    uint32_t *keep = gidx == 0 ? P0 : P1;
                    for (int i = tid; i < m; i += T) keep[i] = P[i];
                }
                if (a.uniform) {
                    if (tid < 32) cntW[tid] = 0;
                    __syncthreads();
                    int acc = 0;
                    for (int w = warp; w < S; w += nwarps) {
                        const int i = w * 32 + lane;
                        const uint32_t q = (i < m && rowblk[i] != OSDP_NONE8) ? (P[i] ^ (s8[i] ? 0xffffffffu : 0u)) : 0u;
                        for (int c = 0; c < ncol; c++) {
                            const unsigned bal = __ballot_sync(0xffffffffu, (q >> c) & 1u);
                            if (lane == c) acc += __popc(bal);
                        }
                    }
                    if (lane < ncol && acc) atomicAdd(&cntW[lane], acc);
                    __syncthreads();
                    if (warp == 0) { // ascending candidate order, strict '<'
                        for (int c = 0; c < ncol; c++) {
                            const double W = (double)(cntW[c] + 1);
                            if (bestC == -2 || W < bestW) { bestW = W; bestC = (long long)gidx * 32 + c; }
                        }
                    }
                    __syncthreads();
                } else {
                    __syncthreads();
                    for (int c = warp; c < ncol; c += nwarps) {
                        for (int w = lane; w < S; w += 32) s2[w] = 0; // lanes write disjoint words below
                        __syncwarp();
                        for (int w = 0; w < S; w++) {
                            const int i = w * 32 + lane;
                            const bool x = i < m && rowblk[i] != OSDP_NONE8 && (((P[i] >> c) & 1u) ^ (s8[i] & 1u));
                            const unsigned bal = __ballot_sync(0xffffffffu, x);
                            if (lane == 0) s2[w] = bal;
                        }
                        if (lane == 0) selcol[0] = order[np[gidx * 32 + c]];
                        __syncwarp();
                        const double W = soft_weight(1);
                        const long long idx = (long long)gidx * 32 + c;
                        if (bestC == -2 || W < bestW || (W == bestW && idx < bestC)) { bestW = W; bestC = idx; }
                        __syncwarp();
                    }
                    __syncthreads();
                }
            }
        } else {
            build_panel(np, 0, wd);   // OSD-E: the first wd non-pivot columns (wd <= 30)
            replay_all();
            for (int i = tid; i < m; i += T) P0[i] = P[i];
            __syncthreads();
        }
        // combinations out of the kept panels: pairs (OSD-CS) or all subsets (OSD-E), one candidate per warp
        {
            __syncthreads();
            const long long ncomb = a.method == 2 ? npairs : ((1ll << wd) - 1);
            const long long idx0 = a.method == 2 ? (long long)nnp : 0; // candidate index of the first combination
            for (long long cc = warp; cc < ncomb; cc += nwarps) {
                int nsel = 0;
                uint32_t v0 = 0, v1 = 0; // selection masks over P0 / P1 columns
                if (a.method == 2) {
                    long long idx = cc; int i = 0;
                    while (idx >= wd - 1 - i) { idx -= wd - 1 - i; i++; }
                    const int j = i + 1 + (int)idx;
                    if (i < 32) v0 |= 1u << i; else v1 |= 1u << (i - 32);
                    if (j < 32) v0 |= 1u << j; else v1 |= 1u << (j - 32);
                    if (lane == 0) { selcol[0] = order[np[i]]; selcol[1] = order[np[j]]; }
                    nsel = 2;
                } else {
                    v0 = (uint32_t)(cc + 1);
                    for (int b = 0; b < wd; b++)
                        if ((v0 >> b) & 1u) { if (lane == 0) selcol[nsel] = order[np[b]]; nsel++; }
                }
                __syncwarp();
                int pc = 0;
                for (int w = 0; w < S; w++) {
                    const int i = w * 32 + lane;
                    bool x = false;
                    if (i < m && rowblk[i] != OSDP_NONE8) {
                        uint32_t par = P0[i] & v0;
                        if (v1) par ^= P1[i] & v1;
                        x = ((__popc(par) ^ s8[i]) & 1) != 0;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, x);
                    pc += __popc(bal);
                    if (!a.uniform && lane == 0) s2[w] = bal;
                }
                double W;
                if (a.uniform) W = (double)(pc + nsel);
                else { __syncwarp(); W = soft_weight(nsel); }
                const long long idx = idx0 + cc;
                if (bestC == -2 || W < bestW || (W == bestW && idx < bestC)) { bestW = W; bestC = idx; }
                __syncwarp();
            }
        }
        if (lane == 0) { red_w[warp] = bestW; red_c[warp] = (int)bestC; }
        __syncthreads();
        if (tid == 0) {
            double bw = 0; int bc = -2;
            for (int k = 0; k < nwarps; k++) {
                if (red_c[k] == -2) continue;
                if (bc == -2 || red_w[k] < bw || (red_w[k] == bw && red_c[k] < bc)) { bw = red_w[k]; bc = red_c[k]; }
            }
            sh_best = bc;
        }
        __syncthreads();
        const int bc = sh_best;
        if (bc < 0) continue; // OSD-0 stands (ties keep the earlier candidate)
        // rebuild the winner: selection masks over a panel, then x_i and the selected columns
        uint32_t v0 = 0, v1 = 0;
        int sel_a = -1, sel_b = -1;      // selected non-pivot positions (index into np) for CS
        const uint32_t *src0 = P0;
        if (a.method == 2 && bc < nnp) {
            const int gidx = bc / 32;
            if (gidx >= 2) { // not one of the kept panels: reduce it again
                build_panel(np + gidx * 32, 0, min(32, nnp - gidx * 32));
                replay_all();
                src0 = P;
            } else if (gidx == 1) src0 = P1;
            v0 = 1u << (bc & 31);
            sel_a = bc;
        } else if (a.method == 2) {
            int idx = bc - nnp, i = 0;
            while (idx >= wd - 1 - i) { idx -= wd - 1 - i; i++; }
            const int j = i + 1 + idx;
            if (i < 32) v0 |= 1u << i; else v1 |= 1u << (i - 32);
            if (j < 32) v0 |= 1u << j; else v1 |= 1u << (j - 32);
            sel_a = i; sel_b = j;
        } else v0 = (uint32_t)(bc + 1);
        for (int j = tid; j < n; j += T) {
            const unsigned pr = prow[j];
            int x = 0;
            if (pr != OSD_NONE) {
                uint32_t par = src0[pr] & v0;
                if (v1) par ^= P1[pr] & v1;
                x = (__popc(par) ^ s8[pr]) & 1;
            }
            a.osdw[base + j] = (uint8_t)x;
        }
        __syncthreads();
        if (a.method == 2) {
            if (tid == 0) {
                a.osdw[base + order[np[sel_a]]] = 1;
                if (sel_b >= 0) a.osdw[base + order[np[sel_b]]] = 1;
            }
        } else {
            if (tid < wd && ((v0 >> tid) & 1u)) a.osdw[base + order[np[tid]]] = 1;
        }
    }
}

} // namespace bposd
