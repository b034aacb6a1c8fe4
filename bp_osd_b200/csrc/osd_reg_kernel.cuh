// osd_reg_kernel.cuh -- OSD (rows a9-a14) with the row-operation matrix T held in REGISTERS, for m <= 1024.
//
// One CTA per failed shot.  The elimination keeps, as osd_kernel does, the m x m matrix T of row operations (the
// reduced image of H column c is the XOR of the T columns named by the rows of c, so the dense m x n matrix is never
// formed), but a thread OWNS its CPT columns of T in registers (W 32-bit words each) for the whole elimination, and the
// sorted columns are consumed in rounds of G candidates with two block barriers per ROUND instead of two per pivot:
//
//   phase U (all threads)  apply the <= G pivots of the previous round to the owned columns: for pivot (p, v), a column
//                          with bit p set becomes column ^ v.  v is read with broadcast 16-byte LDS; the word that
//                          holds bit p is picked from the register array through a CTA-uniform jump table.  A warp
//                          none of whose columns has the bit skips the pivot.  Then the (few) columns that this round's
//                          candidates name are copied to a shared-memory mirror (Tpub).
//   phase S (warp 0)       lane l holds word l of the reduced image of each of the G candidates (XOR of <= maxdeg
//                          mirrored columns).  The candidates are resolved in sorted order without leaving the warp: a
//                          candidate with a 1 in an unused row is a pivot (lowest such row; the OSD result does not
//                          depend on that choice, row a10); its image v, with bit p cleared, is recorded for the next
//                          phase U and applied on the spot to the images of the later candidates of the round and to
//                          the transformed syndrome s' (one shuffle + one LOP3 each).  Candidates without such a row
//                          are dependent for good.
//   phase P (other warps)  meanwhile: fetch the rows of the NEXT round's candidate columns from the CSC arrays (L2
//                          latency hidden behind phase S) and mark the T columns they will need.
//
// The scan stops when rank(H) pivots are found (later columns cannot be pivots).  After the last round T is mirrored
// completely and the read-out / candidate search of osd_kernel runs on the mirror (osd_readout_and_search).
//
// Cost per shot, cfg 3 (m = 961, W = 32, 936 pivots): ~ 936 * 16 warps * ~90 instructions in phase U and ~120 rounds
// of ~1 300 serial cycles in phase S, against ~1 900 pivot steps of two barriers + a shared-memory read-modify-write
// sweep of T each in osd_kernel.
#pragma once
#include "bposd_kernels.cuh"

namespace bposd {

constexpr int kOsdRegG = 16;   // candidates per round
constexpr int kOsdRegCPT = 2;  // columns of T per thread
#define OSDR_NONE 0xFFFFu

__host__ __device__ constexpr int osd_reg_ws(int W) { return W + 4; } // mirror column stride: 16-byte stores of a warp spread over all banks

// word `w` (CTA-uniform) of each owned column: a jump table, not W predicated selects
template <int W, int CPT>
__device__ __forceinline__ void osd_reg_pick(const uint32_t (&col)[CPT][W], int w, uint32_t (&sel)[CPT]) {
#define OSDR_CASE(K)                                                        \
    case K:                                                                 \
        if constexpr (K < W) {                                              \
            _Pragma("unroll") for (int c = 0; c < CPT; c++) sel[c] = col[c][K]; \
        }                                                                   \
        break;
    switch (w) {
        OSDR_CASE(0) OSDR_CASE(1) OSDR_CASE(2) OSDR_CASE(3) OSDR_CASE(4) OSDR_CASE(5) OSDR_CASE(6) OSDR_CASE(7)
        OSDR_CASE(8) OSDR_CASE(9) OSDR_CASE(10) OSDR_CASE(11) OSDR_CASE(12) OSDR_CASE(13) OSDR_CASE(14) OSDR_CASE(15)
        OSDR_CASE(16) OSDR_CASE(17) OSDR_CASE(18) OSDR_CASE(19) OSDR_CASE(20) OSDR_CASE(21) OSDR_CASE(22) OSDR_CASE(23)
        OSDR_CASE(24) OSDR_CASE(25) OSDR_CASE(26) OSDR_CASE(27) OSDR_CASE(28) OSDR_CASE(29) OSDR_CASE(30) OSDR_CASE(31)
        default: break;
    }
#undef OSDR_CASE
}

// apply the recorded pivots of one round to the owned columns
template <int W, int CPT>
__device__ __forceinline__ void osd_reg_apply(uint32_t (&col)[CPT][W], int g, const int *piv_p, const uint32_t *piv_v) {
    for (int j = 0; j < g; j++) {
        const int p = piv_p[j];
        uint32_t sel[CPT];
#pragma unroll
        for (int c = 0; c < CPT; c++) sel[c] = 0;
        osd_reg_pick<W, CPT>(col, p >> 5, sel);
        uint32_t mask[CPT], any = 0;
#pragma unroll
        for (int c = 0; c < CPT; c++) { mask[c] = 0u - ((sel[c] >> (p & 31)) & 1u); any |= mask[c]; }
        if (__any_sync(0xffffffffu, any != 0)) {
            const uint4 *v4 = reinterpret_cast<const uint4 *>(piv_v + (size_t)j * W);
#pragma unroll
            for (int w4 = 0; w4 < W / 4; w4++) {
                const uint4 v = v4[w4];
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    col[c][4 * w4 + 0] ^= v.x & mask[c];
                    col[c][4 * w4 + 1] ^= v.y & mask[c];
                    col[c][4 * w4 + 2] ^= v.z & mask[c];
                    col[c][4 * w4 + 3] ^= v.w & mask[c];
                }
            }
        }
    }
}

template <int W>
__device__ __forceinline__ void osd_reg_publish(const uint32_t (&c)[W], uint32_t *dst) {
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int w4 = 0; w4 < W / 4; w4++) d4[w4] = make_uint4(c[4 * w4], c[4 * w4 + 1], c[4 * w4 + 2], c[4 * w4 + 3]);
}

template <int W>
static inline size_t osd_reg_smem_bytes(int m, int n, int threads, int maxdeg) {
    const int S = (m + 31) / 32, nw = threads / 32, G = kOsdRegG;
    size_t b = 0;
    b += std::max((size_t)m * osd_reg_ws(W) * 4, (size_t)n * 8); // Tpub, aliased by the sort keys
    b += 32 * 8;                                                   // red_w
    b += (size_t)2 * W * 4 + (size_t)2 * S * 4;                    // needed[2][W], used[S], sprime[S]
    b += (size_t)nw * (S + 64) * 4;                                // wscr
    b += 32 * 4;                                                   // red_c
    b += (size_t)G * W * 4 + (size_t)G * 4;                        // piv_v, piv_p
    b += (size_t)3 * n * 2;                                        // order, prow, np
    b += (size_t)2 * G * maxdeg * 2;                               // cand_rows
    return b + 64;                                                 // alignment slack
}

template <typename real, int W>
__global__ void __launch_bounds__(W * 32 / kOsdRegCPT < 64 ? 64 : W * 32 / kOsdRegCPT) osd_reg_kernel(OsdArgs<real> a) {
    constexpr int CPT = kOsdRegCPT, G = kOsdRegG, WS = osd_reg_ws(W);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n, S = a.S, maxdeg = a.maxdeg;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    // shared-memory carve-up (every section a multiple of 16 bytes where it has to be)
    const size_t tpub_bytes = ((size_t)m * WS * 4 > (size_t)n * 8 ? (size_t)m * WS * 4 : (size_t)n * 8);
    uint32_t *Tpub = reinterpret_cast<uint32_t *>(smem_raw);                                  // m * WS
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);              // n (sort only)
    double *red_w = reinterpret_cast<double *>(smem_raw + (tpub_bytes + 15) / 16 * 16);       // 32
    uint32_t *piv_v = reinterpret_cast<uint32_t *>(red_w + 32);                               // G * W (16-byte aligned)
    uint32_t *needed = piv_v + G * W;                                                         // 2 * W
    uint32_t *used = needed + 2 * W;                                                          // S
    uint32_t *sprime = used + S;                                                              // S
    uint32_t *wscr = sprime + S;                                                              // nwarps * (S + 64)
    int *red_c = reinterpret_cast<int *>(wscr + (size_t)nwarps * (S + 64));                   // 32
    int *piv_p = red_c + 32;                                                                  // G
    uint16_t *order = reinterpret_cast<uint16_t *>(piv_p + G);                                // n
    uint16_t *prow = order + n;                                                               // n
    uint16_t *np = prow + n;                                                                  // n
    uint16_t *cand_rows = np + n;                                                             // 2 * G * maxdeg
    __shared__ int sh_found, sh_best, sh_rank, sh_nnp, sh_g, sh_done;

    const int nfail = *a.fail_count;
    for (int f = blockIdx.x; f < nfail; f += gridDim.x) {
        const long long shot = a.fail_list[f];
        const real *llr = a.llr + (a.llr_by_shot ? shot : (long long)f) * n;
        const uint8_t *synd = a.synd + shot * m;
        const double *weight = a.weight + shot * a.weight_stride;
        __syncthreads();

        // ---- a9: stable ascending rank sort on (llr, index) ----
        for (int j = tid; j < n; j += T) { keys[j] = sort_key(llr[j]); prow[j] = OSDR_NONE; }
        if (tid < 2 * W) needed[tid] = 0;
        if (tid == 0) { sh_rank = 0; sh_nnp = 0; sh_g = 0; sh_done = 0; }
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
        // T = identity, in registers: thread tid owns columns tid, tid + T, ...
        uint32_t col[CPT][W];
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const int r = tid + c * T;
#pragma unroll
            for (int w = 0; w < W; w++) col[c][w] = (r < m && (r >> 5) == w) ? (1u << (r & 31)) : 0u;
        }
        __syncthreads(); // order complete, keys dead (Tpub may be written from here on)
        // rows of the candidates of round 0 (phase P of a round "-1")
        for (int e = tid; e < G * maxdeg; e += T) {
            const int i = e / maxdeg, k = e - i * maxdeg, t = i;
            unsigned q = OSDR_NONE;
            if (t < n) {
                const int c = order[t], beg = g.col_ptr[c];
                if (k < g.col_ptr[c + 1] - beg) q = (unsigned)g.row_idx[beg + k];
            }
            cand_rows[e] = (uint16_t)q;
            if (q != OSDR_NONE) atomicOr(&needed[q >> 5], 1u << (q & 31));
        }
        // the resolver warp keeps the used-row mask and the transformed syndrome, one word per lane
        uint32_t r_used = 0, r_sp = 0;
        if (warp == 0 && lane < S) {
#pragma unroll 4
            for (int b = 0; b < 32; b++) {
                const int i = lane * 32 + b;
                if (i < m) r_sp |= (uint32_t)(synd[i] & 1) << b;
            }
        }
        __syncthreads();

        // ---- a10: elimination in rounds of G sorted columns ----
        for (int R = 0;; R++) {
            const int buf = R & 1;
            // phase U
            osd_reg_apply<W, CPT>(col, sh_g, piv_p, piv_v);
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const int r = tid + c * T;
                if (r < m && ((needed[buf * W + (r >> 5)] >> (r & 31)) & 1u)) osd_reg_publish<W>(col[c], Tpub + (size_t)r * WS);
            }
            if (tid < W) needed[(buf ^ 1) * W + tid] = 0;
            __syncthreads(); // A
            if (warp == 0) {
                // phase S
                const int t0 = R * G;
                const uint16_t *rows = cand_rows + (size_t)buf * G * maxdeg;
                uint32_t v[G];
#pragma unroll
                for (int i = 0; i < G; i++) {
                    v[i] = 0;
                    if (t0 + i < n)
                        for (int k = 0; k < maxdeg; k++) {
                            const unsigned q = rows[i * maxdeg + k];
                            if (q == OSDR_NONE) break;
                            if (lane < W) v[i] ^= Tpub[(size_t)q * WS + lane];
                        }
                }
                int rank = sh_rank, nnp = sh_nnp, gcount = 0;
#pragma unroll
                for (int i = 0; i < G; i++) {
                    if (t0 + i < n) {
                        const uint32_t cand = v[i] & ~r_used;
                        const unsigned bal = (rank < a.maxrank) ? __ballot_sync(0xffffffffu, cand != 0) : 0u;
                        if (bal == 0) {
                            if (lane == 0) np[nnp] = (uint16_t)(t0 + i);
                            nnp++;
                        } else {
                            const int pw = __ffs(bal) - 1;
                            const int pb = __ffs(__shfl_sync(0xffffffffu, cand, pw)) - 1;
                            uint32_t vclr = v[i];
                            if (lane == pw) { vclr &= ~(1u << pb); r_used |= 1u << pb; }
                            if (lane == 0) { piv_p[gcount] = pw * 32 + pb; prow[order[t0 + i]] = (uint16_t)(pw * 32 + pb); }
                            if (lane < W) piv_v[gcount * W + lane] = vclr;
#pragma unroll
                            for (int c2 = i + 1; c2 < G; c2++)
                                if ((__shfl_sync(0xffffffffu, v[c2], pw) >> pb) & 1u) v[c2] ^= vclr;
                            if ((__shfl_sync(0xffffffffu, r_sp, pw) >> pb) & 1u) r_sp ^= vclr;
                            gcount++; rank++;
                        }
                    }
                }
                if (lane == 0) {
                    sh_rank = rank; sh_nnp = nnp; sh_g = gcount;
                    sh_done = (rank >= a.maxrank || t0 + G >= n) ? 1 : 0;
                }
            } else {
                // phase P: rows of the next round's candidates
                const int t1 = (R + 1) * G;
                uint16_t *rows = cand_rows + (size_t)(buf ^ 1) * G * maxdeg;
                for (int e = tid - 32; e < G * maxdeg; e += T - 32) {
                    const int i = e / maxdeg, k = e - i * maxdeg, t = t1 + i;
                    unsigned q = OSDR_NONE;
                    if (t < n) {
                        const int c = order[t], beg = g.col_ptr[c];
                        if (k < g.col_ptr[c + 1] - beg) q = (unsigned)g.row_idx[beg + k];
                    }
                    rows[e] = (uint16_t)q;
                    if (q != OSDR_NONE) atomicOr(&needed[(buf ^ 1) * W + (q >> 5)], 1u << (q & 31));
                }
            }
            __syncthreads(); // B
            if (sh_done) {
                // positions never examined are non-pivots, in order (row a10: the scan stops at rank(H) pivots)
                const int t_end = min(n, (R + 1) * G), nnp0 = sh_nnp;
                for (int t = t_end + tid; t < n; t += T) np[nnp0 + (t - t_end)] = (uint16_t)t;
                break;
            }
        }
        // the last round's pivots still have to reach T before the candidate search reads it
        const bool need_T = !(a.method == 0 || a.order <= 0 || !a.osdw);
        if (need_T) {
            osd_reg_apply<W, CPT>(col, sh_g, piv_p, piv_v);
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const int r = tid + c * T;
                if (r < m) osd_reg_publish<W>(col[c], Tpub + (size_t)r * WS);
            }
        }
        if (warp == 0 && lane < S) { used[lane] = r_used; sprime[lane] = r_sp & r_used; }
        __syncthreads();
        osd_readout_and_search<real>(a, shot, weight, Tpub, WS, used, sprime, wscr, red_w, red_c, order, prow, np, a.g.n - sh_rank,
                                     &sh_best, &sh_found);
    }
}

} // namespace bposd
