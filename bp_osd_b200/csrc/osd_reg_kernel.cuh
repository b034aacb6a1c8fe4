// osd_reg_kernel.cuh -- OSD (rows a9-a14) with the row-operation matrix T held in REGISTERS, for m <= 1024.
//
// One CTA per failed shot.  As in osd_kernel the elimination keeps the m x m matrix T of row operations (the reduced
// image of H column c is the XOR of the T columns named by the rows of c, so the dense m x n matrix is never formed),
// but here
//   * an UPDATER thread owns CPT columns of T in registers (W 32-bit words each) for the whole elimination;
//   * the sorted columns are consumed in rounds of G candidates, ONE block barrier per round (osd_kernel: two per
//     pivot, ~1 900 pivot steps on the bench code);
//   * the CTA is warp specialised and software pipelined: in round R
//       one warp, the RESOLVER takes the reduced images of the G candidates of round R (lane l holds word l of each;
//                             they were mirrored by the updaters one round earlier, so the resolver first applies the
//                             pivots of round R-1, which it still holds in registers), walks them in sorted order -- a
//                             candidate with a 1 in an unused row is a pivot (lowest such row; the OSD result does not
//                             depend on that choice, row a10), its image v with the pivot bit cleared is recorded and
//                             applied on the spot to the later candidates of the round and to the transformed syndrome
//                             s' (one shuffle + one LOP3 each); a candidate without such a row is dependent for good;
//       the UPDATER warps     meanwhile apply the pivots of round R-1 to their columns (a column with bit p set becomes
//                             column ^ v; v is read with broadcast 16-byte LDS, the word that holds bit p is picked from
//                             the register array through a CTA-uniform branch tree, a warp none of whose columns has
//                             the bit skips the pivot), copy the few columns that the candidates of round R+1 name into
//                             image slots in shared memory, and look up the rows of the candidates of round R+2.
//   * the scan stops when rank(H) pivots are found; T is then mirrored to shared memory once and the read-out /
//     candidate search shared with osd_kernel (osd_readout_and_search) runs on the mirror, with the CSC arrays of H in
//     a 16-bit shared-memory copy (the search walks one column per candidate);
//   * the column order comes from a bitonic sort of (key, index) pairs in shared memory (the pairs are distinct, so the
//     result is the stable ascending order of row a9) instead of the O(n^2) rank sort.
//
// Bound: the T update is LOP3 work -- 64 lanes per clock per SM on sm_100a -- of ~2 x 32 LOP3 per (thread, pivot) for
// every warp that holds at least one column with the pivot bit set.
#pragma once
#include "bposd_kernels.cuh"

namespace bposd {

constexpr int kOsdRegCPT = 2; // columns of T per updater thread
#ifndef BPOSD_OSDR_G
#define BPOSD_OSDR_G 16
#endif
constexpr int kOsdRegG = BPOSD_OSDR_G; // candidates per round
#define OSDR_NONE16 0xFFFFu
#define OSDR_NONE32 0xFFFFFFFFu
#define OSDR_NONE8 0xFFu

// word `w` (CTA-uniform) of each owned column: a branch tree on a uniform value, not W predicated selects
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

// barrier over the whole CTA that the two warp roles reach from different code paths (named barrier 1, explicit count)
__device__ __forceinline__ void osd_reg_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__host__ __device__ constexpr int osd_reg_ws(int W) { return W + 4; } // mirror column stride: 16-byte stores of a warp spread over all banks
static inline int osd_reg_np2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
static inline int osd_reg_threads(int m) { return 32 + std::max(64, ((m + kOsdRegCPT - 1) / kOsdRegCPT + 31) / 32 * 32); }

// Shared-memory layout, the same arithmetic on host and device.
struct OsdRegLayout {
    size_t o_mirror, o_redw, o_pivv, o_img, o_slotof, o_used, o_sprime, o_wscr, o_redc, o_pivp, o_order, o_prow, o_np, o_cptr,
        o_crow, o_eslot, total;
};
__host__ __device__ inline OsdRegLayout osd_reg_layout(int m, int n, int E, int W, int KD, int G, int threads, int np2) {
    OsdRegLayout L;
    const size_t S = (size_t)(m + 31) / 32, nw = (size_t)threads / 32;
    auto al = [](size_t x) { return (x + 15) / 16 * 16; };
    size_t o = 0;
    const size_t mirror = (size_t)m * osd_reg_ws(W) * 4, sortb = (size_t)np2 * 8 + (size_t)np2 * 2;
    L.o_mirror = o; o = al(o + (mirror > sortb ? mirror : sortb)); // T mirror (after the elimination) / sort keys + indices (before)
    L.o_redw = o; o = al(o + 32 * 8);
    L.o_pivv = o; o = al(o + (size_t)2 * G * W * 4);                // pivot vectors, double buffered
    L.o_img = o; o = al(o + (size_t)2 * G * KD * W * 4);            // candidate image slots, double buffered
    L.o_slotof = o; o = al(o + (size_t)2 * m * 4);                  // column -> image slot, double buffered
    L.o_used = o; o = al(o + S * 4);
    L.o_sprime = o; o = al(o + S * 4);
    L.o_wscr = o; o = al(o + nw * (S + 64) * 4);
    L.o_redc = o; o = al(o + 32 * 4);
    L.o_pivp = o; o = al(o + (size_t)2 * G * 4);
    L.o_order = o; o = al(o + (size_t)n * 2);
    L.o_prow = o; o = al(o + (size_t)n * 2);
    L.o_np = o; o = al(o + (size_t)n * 2);
    L.o_cptr = o; o = al(o + ((size_t)n + 1) * 2);
    L.o_crow = o; o = al(o + (size_t)(E > 0 ? E : 1) * 2);
    L.o_eslot = o; o = al(o + (size_t)3 * G * KD);
    L.total = o;
    return L;
}

template <typename real, int W, int KD>
__global__ void __launch_bounds__(32 + (W * 32 / kOsdRegCPT < 64 ? 64 : W * 32 / kOsdRegCPT)) osd_reg_kernel(OsdArgs<real> a) {
    constexpr int CPT = kOsdRegCPT, G = kOsdRegG, WS = osd_reg_ws(W), NE = G * KD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n, E = g.nnz, S = a.S;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    // the resolver is the LAST warp: the sub-partition schedulers of sm_100a favour the highest warp id, and the resolver's
    // serial chain, not the updaters' bulk work, sets the length of a round (ncu: with the resolver as warp 0 the
    // updaters spent 47 % of their time at the round barrier, profiles/r2c_osd_ncu_summary.md)
    const int TU = T - 32, tu = tid; // updater threads are 0 .. TU-1
    const bool resolver = warp == (T >> 5) - 1;
    const int NP = a.np2;
    const OsdRegLayout L = osd_reg_layout(m, n, E, W, KD, G, T, NP);
    uint32_t *Tmir = reinterpret_cast<uint32_t *>(smem_raw + L.o_mirror);
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(smem_raw + L.o_mirror);
    uint16_t *sidx = reinterpret_cast<uint16_t *>(skey + NP);
    double *red_w = reinterpret_cast<double *>(smem_raw + L.o_redw);
    uint32_t *piv_v = reinterpret_cast<uint32_t *>(smem_raw + L.o_pivv);   // [2][G][W]
    uint32_t *img = reinterpret_cast<uint32_t *>(smem_raw + L.o_img);      // [2][NE][W]
    uint32_t *slot_of = reinterpret_cast<uint32_t *>(smem_raw + L.o_slotof); // [2][m]
    uint32_t *used = reinterpret_cast<uint32_t *>(smem_raw + L.o_used);
    uint32_t *sprime = reinterpret_cast<uint32_t *>(smem_raw + L.o_sprime);
    uint32_t *wscr = reinterpret_cast<uint32_t *>(smem_raw + L.o_wscr);
    int *red_c = reinterpret_cast<int *>(smem_raw + L.o_redc);
    int *piv_p = reinterpret_cast<int *>(smem_raw + L.o_pivp);             // [2][G]
    uint16_t *order = reinterpret_cast<uint16_t *>(smem_raw + L.o_order);
    uint16_t *prow = reinterpret_cast<uint16_t *>(smem_raw + L.o_prow);
    uint16_t *np = reinterpret_cast<uint16_t *>(smem_raw + L.o_np);
    uint16_t *cptr = reinterpret_cast<uint16_t *>(smem_raw + L.o_cptr);
    uint16_t *crow = reinterpret_cast<uint16_t *>(smem_raw + L.o_crow);
    uint8_t *eslot = smem_raw + L.o_eslot;                                  // [3][NE]
    __shared__ int sh_found, sh_best, sh_rank, sh_nnp, sh_rounds, sh_g[2], sh_done[2];

    // CSC of H, 16-bit, once per CTA
    for (int j = tid; j <= n; j += T) cptr[j] = (uint16_t)g.col_ptr[j];
    for (int e = tid; e < E; e += T) crow[e] = (uint16_t)g.row_idx[e];
    const CscView cv{nullptr, nullptr, cptr, crow};

    // rows of the candidates of round `round` -> image slot of every (candidate, edge) entry; a T column named twice
    // gets one slot (the first claimant's), which its owner fills when it sees slot_of[column] set
    auto prepare = [&](int round, int first, int stride) {
        uint32_t *so = slot_of + (size_t)(round & 1) * m;
        uint8_t *es = eslot + (size_t)(round % 3) * NE;
        for (int e = first; e < NE; e += stride) {
            const int i = e / KD, k = e - i * KD, t = round * G + i;
            unsigned slot = OSDR_NONE8;
            if (t < n) {
                const int c = order[t], beg = cptr[c];
                if (k < (int)cptr[c + 1] - beg) {
                    const unsigned q = crow[beg + k];
                    const unsigned old = atomicCAS(&so[q], OSDR_NONE32, (unsigned)e);
                    slot = (old == OSDR_NONE32) ? (unsigned)e : old;
                }
            }
            es[e] = (uint8_t)slot;
        }
    };

    const int nfail = *a.fail_count;
    for (int f = blockIdx.x; f < nfail; f += gridDim.x) {
        const long long shot = a.fail_list[f];
        const real *llr = a.llr + (a.llr_by_shot ? shot : (long long)f) * n;
        const double *weight = a.weight + shot * a.weight_stride;
        __syncthreads();

        // ---- a9: ascending order on (llr, index): bitonic sort of the distinct pairs (key, index) ----
        for (int j = tid; j < NP; j += T) {
            skey[j] = (j < n) ? sort_key(llr[j]) : ~0ull;
            sidx[j] = (j < n) ? (uint16_t)j : (uint16_t)0xFFFFu;
        }
        for (int j = tid; j < n; j += T) prow[j] = OSDR_NONE16;
        for (int j = tid; j < 2 * m; j += T) slot_of[j] = OSDR_NONE32;
        if (tid == 0) { sh_g[0] = sh_g[1] = 0; sh_done[0] = sh_done[1] = 0; sh_rank = 0; sh_nnp = 0; }
        __syncthreads();
        for (int k = 2; k <= NP; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (NP >> 1); t += T) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const unsigned long long ka = skey[i], kb = skey[l];
                    const unsigned ia = sidx[i], ib = sidx[l];
                    const bool gt = ka > kb || (ka == kb && ia > ib);
                    if (gt == ((i & k) == 0)) { skey[i] = kb; skey[l] = ka; sidx[i] = (uint16_t)ib; sidx[l] = (uint16_t)ia; }
                }
                __syncthreads();
            }
        for (int j = tid; j < n; j += T) order[j] = sidx[j];
        __syncthreads(); // order complete

        // pipeline prologue, part 1: image slots of round 0
        prepare(0, tid, T);
        __syncthreads();
        const bool need_T = !(a.method == 0 || a.order <= 0 || !a.osdw); // the candidate search reads T
        int R = 0;
        if (resolver) {
            // =================== resolver warp ===================
            prepare(1, lane, 32);
            // used-row mask and transformed syndrome, one word per lane
            uint32_t r_used = 0, r_sp = 0;
            int gprev = 0, r_rank = 0, r_nnp = 0;
            if (lane < S) {
#pragma unroll 4
                for (int b = 0; b < 32; b++) {
                    const int i = lane * 32 + b;
                    if (i < m) r_sp |= synd_bit(a.synd, shot, m, i, a.synd_packed) << b;
                }
            }
            osd_reg_bar(T);
            for (;; R++) {
                const int par = R & 1, t0 = R * G;
                const uint32_t *im = img + (size_t)par * NE * W;
                const uint8_t *es = eslot + (size_t)(R % 3) * NE;
                uint32_t v[G];
#pragma unroll
                for (int i = 0; i < G; i++) {
                    v[i] = 0;
#pragma unroll
                    for (int k = 0; k < KD; k++) {
                        const unsigned s = es[i * KD + k];
                        if (s != OSDR_NONE8 && lane < W) v[i] ^= im[(size_t)s * W + lane];
                    }
                }
                // the images were copied before the pivots of round R-1 reached T: apply those here (they are still in
                // the other half of the pivot buffers, which the updaters are reading, not writing)
                {
                    const int *pq = piv_p + (par ^ 1) * G;
                    const uint32_t *pvq = piv_v + (size_t)(par ^ 1) * G * W;
                    for (int j = 0; j < gprev; j++) {
                        const int p = pq[j], pw = p >> 5, pb = p & 31;
                        const uint32_t vj = (lane < W) ? pvq[j * W + lane] : 0u;
#pragma unroll
                        for (int i = 0; i < G; i++)
                            if ((__shfl_sync(0xffffffffu, v[i], pw) >> pb) & 1u) v[i] ^= vj;
                    }
                }
                int gcount = 0;
                int *pp = piv_p + par * G;
                uint32_t *pvv = piv_v + (size_t)par * G * W;
#pragma unroll
                for (int i = 0; i < G; i++) {
                    if (t0 + i < n) {
                        const uint32_t cand = v[i] & ~r_used;
                        const unsigned bal = (r_rank < a.maxrank) ? __ballot_sync(0xffffffffu, cand != 0) : 0u;
                        if (bal == 0) {
                            if (lane == 0) np[r_nnp] = (uint16_t)(t0 + i);
                            r_nnp++;
                        } else {
                            const int pw = __ffs(bal) - 1;
                            const int pb = __ffs(__shfl_sync(0xffffffffu, cand, pw)) - 1;
                            uint32_t vclr = v[i];
                            if (lane == pw) { vclr &= ~(1u << pb); r_used |= 1u << pb; }
                            if (lane == 0) { pp[gcount] = pw * 32 + pb; prow[order[t0 + i]] = (uint16_t)(pw * 32 + pb); }
                            if (lane < W) pvv[gcount * W + lane] = vclr;
#pragma unroll
                            for (int c2 = i + 1; c2 < G; c2++)
                                if ((__shfl_sync(0xffffffffu, v[c2], pw) >> pb) & 1u) v[c2] ^= vclr;
                            if ((__shfl_sync(0xffffffffu, r_sp, pw) >> pb) & 1u) r_sp ^= vclr;
                            gcount++; r_rank++;
                        }
                    }
                }
                gprev = gcount;
                const int done = (r_rank >= a.maxrank || t0 + G >= n) ? 1 : 0;
                if (lane == 0) { sh_g[par] = gcount; sh_done[par] = done; }
                osd_reg_bar(T);
                if (done) break;
            }
            if (lane == 0) { sh_rank = r_rank; sh_nnp = r_nnp; sh_rounds = R; }
            if (lane < S) { used[lane] = r_used; sprime[lane] = r_sp & r_used; }
        } else {
            // =================== updater warps ===================
            // T = identity, in registers: updater tu owns columns tu, tu + TU
            uint32_t col[CPT][W];
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const int r = tu + c * TU;
#pragma unroll
                for (int w = 0; w < W; w++) col[c][w] = (r < m && (r >> 5) == w) ? (1u << (r & 31)) : 0u;
            }
            // pipeline prologue, part 2: images of round 0 (T is still the identity); the resolver prepares round 1
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const int r = tu + c * TU;
                if (r < m) {
                    const unsigned s = slot_of[r];
                    if (s != OSDR_NONE32) { osd_reg_publish<W>(col[c], img + (size_t)s * W); slot_of[r] = OSDR_NONE32; }
                }
            }
            osd_reg_bar(T);
            for (;; R++) {
                const int par = R & 1;
                // pivots of round R-1 into T, then the images of round R+1, then the slots of round R+2
                osd_reg_apply<W, CPT>(col, sh_g[par ^ 1], piv_p + (par ^ 1) * G, piv_v + (size_t)(par ^ 1) * G * W);
                uint32_t *so = slot_of + (size_t)(par ^ 1) * m;
                uint32_t *im = img + (size_t)(par ^ 1) * NE * W;
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    const int r = tu + c * TU;
                    if (r < m) {
                        const unsigned s = so[r];
                        if (s != OSDR_NONE32) { osd_reg_publish<W>(col[c], im + (size_t)s * W); so[r] = OSDR_NONE32; }
                    }
                }
                prepare(R + 2, tu, TU);
                osd_reg_bar(T);
                if (sh_done[par]) break;
            }
            // the last round's pivots still have to reach T before the candidate search reads it
            if (need_T) {
                const int par = R & 1;
                osd_reg_apply<W, CPT>(col, sh_g[par], piv_p + par * G, piv_v + (size_t)par * G * W);
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    const int r = tu + c * TU;
                    if (r < m) osd_reg_publish<W>(col[c], Tmir + (size_t)r * WS);
                }
            }
        }
        __syncthreads();
        R = sh_rounds;
        {
            // positions never examined are non-pivots, in order (row a10: the scan stops at rank(H) pivots)
            const int t_end = min(n, (R + 1) * G), nnp0 = sh_nnp;
            for (int t = t_end + tid; t < n; t += T) np[nnp0 + (t - t_end)] = (uint16_t)t;
        }
        __syncthreads();
        osd_readout_and_search<real>(a, shot, weight, Tmir, WS, used, sprime, wscr, red_w, red_c, order, prow, np, n - sh_rank,
                                     &sh_best, &sh_found, cv);
    }
}

} // namespace bposd
