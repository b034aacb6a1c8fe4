// osd_cluster_kernel.cuh -- OSD-0 (rows a9-a11) for parity-check matrices whose m x m row-operation matrix does not fit
// in shared memory (BASELINE config 5: m = 19 200, n = 40 000), one THREAD-BLOCK CLUSTER per failed shot.
//
// Left-looking panel Gauss-Jordan over GF(2), as osd0_large_kernel, but
//   * the checks (rows) are split over the CL CTAs of a cluster (rows_per_cta each): a panel of W = 64 sorted columns is
//     one 64-bit word per check, so a CTA's slice of the panel is a few KB of shared memory and the row work of every
//     step is divided by CL;
//   * the row operations of a finished panel are stored in HBM as one 64-bit multiplier mask per check (bit k: "add
//     pivot row k of that panel to this check"); replaying panel q on the current panel streams the CTA's slice of
//     those masks with cp.async.bulk (TMA, 1-D) into a shared-memory ring, completion on an mbarrier, three replays
//     ahead of their use;
//   * what every CTA needs from the others for a replay is the current words of panel q's <= 64 pivot rows: their owners
//     push them into every CTA's gather buffer through distributed shared memory, one barrier.cluster later each CTA
//     forms the resolved pivot words R = C_q g and the final pivot rows F_q g locally (C_q, F_q: 64 x 64 bit matrices
//     that depend on panel q alone and are computed once, when q is factorised -- the sequential 64-step pivot chain of
//     osd0_large_kernel is gone from the replay), folds R into eight 256-entry XOR tables (method of four Russians)
//     and applies them to its rows;
//   * factorising the panel scans its 64 columns in order: every CTA proposes its lowest unused check with a 1 in the
//     column, the proposals (row, panel word, syndrome bit) are exchanged through DSMEM and the lowest CTA's wins (the
//     OSD-0 result does not depend on that choice, row a10); multipliers are recorded as ballots (bit planes) and
//     transposed to per-row masks once per panel;
//   * the column order comes from a cluster-wide bitonic sort of the distinct pairs (key, index) in HBM/L2 (the result is
//     the stable ascending order of row a9) instead of the O(n^2) rank sort;
//   * the scan stops as soon as rank(H) pivots are found.
// Work per shot ~ (panels^2 / 2) replays, each one barrier.cluster + a few hundred instructions per thread; memory per
// cluster panels * m * 8 bytes (96 MB for config 5).
#pragma once
#include "bp_cluster_kernel.cuh" // DSMEM primitives

namespace bposd {

struct OsdcPivot { unsigned long long C, F; int row, pos; }; // per pivot: rows of C_q and F_q, the check, the sorted position

template <typename real>
struct OsdClusterArgs {
    GraphDev g;
    const uint8_t *synd;
    int synd_packed;
    const real *llr;
    int llr_by_shot;
    const int *fail_count;
    const int *fail_list;
    uint8_t *osd0, *osdw;
    unsigned long long *stat;
    int maxrank;  // rank(H)
    int npanels;  // ceil(n / 64)
    int CL, rpc;  // cluster size, rows per CTA (even)
    int np2;      // n rounded up to a power of two
    int stages;   // TMA ring depth of this launch (2 .. kOsdcMaxStages)
    int nfail_lo, nfail_hi; // this launch handles the chunk only if its failed-shot count lies in [lo, hi]: the host enqueues one
                            // launch per cluster size (big clusters for a few failed shots: latency; small ones for many:
                            // throughput) and the others return at once, so no host synchronisation is needed to choose
    // per-cluster workspaces
    unsigned long long *ws_mask; // [nclusters][npanels][rpc * CL]
    unsigned long long *ws_key;  // [nclusters][np2]
    unsigned *ws_idx;            // [nclusters][np2]  (after the sort: the column order)
    OsdcPivot *ws_piv;           // [nclusters][min(m, n)]
};

constexpr int kOsdcMaxStages = 8; // TMA ring depth: as many replays ahead as shared memory allows, at most 8 (masks come from HBM once
                                  // several clusters run: ~2 us away, a replay takes ~1 us)
constexpr int kOsdcThreads = 512;

struct OsdcLayout { size_t o_ring, o_tab, o_gath, o_rk, o_cf, o_candp, o_candr, o_cands, o_plane, o_used, o_sbit, o_red, o_mbar, o_plist, o_prow, o_mjs, o_cs, total; };
__host__ __device__ inline OsdcLayout osdc_layout(int rpc, int npanels, int stages) {
    OsdcLayout L;
    auto al = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t pw = (size_t)(rpc + 31) / 32;
    size_t o = al((size_t)rpc * 8);                                   // P: the panel word of every local check
    L.o_ring = o; o = al(o + (size_t)stages * rpc * 8);               // multiplier masks of replayed panels (TMA destinations)
    L.o_tab = o; o = al(o + 8 * 256 * 8);                             // XOR tables
    L.o_gath = o; o = al(o + 2 * 64 * 8);                             // gathered pivot-row words, double buffered
    L.o_rk = o; o = al(o + 64 * 8);                                   // resolved pivot words
    L.o_cf = o; o = al(o + 64 * (8 + 8 + 4));                         // C / F rows and checks of the replayed panel's pivots
    L.o_candp = o; o = al(o + 2 * 16 * 8);                            // pivot proposals: panel word ...
    L.o_candr = o; o = al(o + 2 * 16 * 4);                            // ... check ...
    L.o_cands = o; o = al(o + 2 * 16 * 4);                            // ... syndrome bit
    L.o_plane = o; o = al(o + 64 * pw * 4);                           // multiplier bit planes of the panel being factorised
    L.o_used = o; o = al(o + pw * 4);
    L.o_sbit = o; o = al(o + pw * 4);
    L.o_red = o; o = al(o + 32 * 4);
    L.o_mbar = o; o = al(o + kOsdcMaxStages * 8);
    L.o_plist = o; o = al(o + ((size_t)npanels + 1) * 4 * 2);         // panels that hold pivots; first pivot of each
    L.o_prow = o; o = al(o + 64 * 4);                                 // checks of this panel's pivots
    L.o_mjs = o; o = al(o + 64 * 8);                                  // their multiplier masks
    L.o_cs = o; o = al(o + 2 * 64 * 8);                               // C and F of the panel being closed
    L.total = o + 16;
    return L;
}

__device__ __forceinline__ void osdc_mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void osdc_mbar_expect(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void osdc_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void osdc_mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ unsigned long long osdc_shfl_xor64(unsigned long long v, int o) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, o), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), o);
    return ((unsigned long long)hi << 32) | lo;
}

template <typename real>
__global__ void __launch_bounds__(kOsdcThreads, 1) osd0_cluster_kernel(OsdClusterArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = a.g;
    const int m = g.m, n = g.n;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int CL = a.CL, rpc = a.rpc, pw = (rpc + 31) / 32;
    const int rank = (int)cluster_ctarank();
    const int cid = (int)cluster_id_x(), nclusters = (int)gridDim.x / CL;
    const int base = rank * rpc, nloc = max(0, min(rpc, m - base));
    const long long gt = (long long)rank * T + tid, GT = (long long)CL * T; // cluster-wide thread index
    const int mpad = rpc * CL;
    const int kOsdcStages = a.stages;
    const OsdcLayout L = osdc_layout(rpc, a.npanels, kOsdcStages);
    unsigned long long *P = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *ring = reinterpret_cast<unsigned long long *>(smem_raw + L.o_ring);
    unsigned long long *tab = reinterpret_cast<unsigned long long *>(smem_raw + L.o_tab);
    unsigned long long *gath = reinterpret_cast<unsigned long long *>(smem_raw + L.o_gath);
    unsigned long long *Rk = reinterpret_cast<unsigned long long *>(smem_raw + L.o_rk);
    unsigned long long *cfC = reinterpret_cast<unsigned long long *>(smem_raw + L.o_cf), *cfF = cfC + 64;
    int *cfRow = reinterpret_cast<int *>(cfF + 64);
    unsigned long long *candp = reinterpret_cast<unsigned long long *>(smem_raw + L.o_candp);
    unsigned *candr = reinterpret_cast<unsigned *>(smem_raw + L.o_candr);
    unsigned *cands = reinterpret_cast<unsigned *>(smem_raw + L.o_cands);
    unsigned *plane = reinterpret_cast<unsigned *>(smem_raw + L.o_plane); // [64][pw]
    unsigned *used = reinterpret_cast<unsigned *>(smem_raw + L.o_used);
    unsigned *sbit = reinterpret_cast<unsigned *>(smem_raw + L.o_sbit);
    int *red = reinterpret_cast<int *>(smem_raw + L.o_red);
    const uint32_t mbar0 = smem_u32(smem_raw + L.o_mbar);
    int *plist = reinterpret_cast<int *>(smem_raw + L.o_plist);          // [npanels + 1] panels with pivots
    int *pfirst = plist + a.npanels + 1;                                  // [npanels + 1] ordinal of the first pivot of plist[s]
    int *prow_s = reinterpret_cast<int *>(smem_raw + L.o_prow);
    unsigned long long *mjs = reinterpret_cast<unsigned long long *>(smem_raw + L.o_mjs);
    unsigned long long *Cs = reinterpret_cast<unsigned long long *>(smem_raw + L.o_cs), *Fs = Cs + 64;

    unsigned long long *maskbase = a.ws_mask + (size_t)cid * a.npanels * mpad;
    unsigned long long *key = a.ws_key + (size_t)cid * a.np2;
    unsigned *order = a.ws_idx + (size_t)cid * a.np2;
    OsdcPivot *piv = a.ws_piv + (size_t)cid * (m < n ? m : n);
    const uint32_t ring_bytes = (uint32_t)rpc * 8u;

    if (tid == 0)
        for (int s = 0; s < kOsdcStages; s++) osdc_mbar_init(mbar0 + 8u * s, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    // ring bookkeeping: stage of the next bulk copy to issue / to wait for, and the mbarrier phase parity of the latter
    uint32_t ist = 0, cst = 0, cpar = 0;

    // push one value into the same shared-memory location of every CTA of the cluster
    auto push_all_u64 = [&](void *local, unsigned long long v) {
        const uint32_t a0 = smem_u32(local);
        for (int c = 0; c < CL; c++) st_dsmem_u64(mapa_u32(a0, (uint32_t)c), v);
    };
    auto push_all_u32 = [&](void *local, unsigned v) {
        const uint32_t a0 = smem_u32(local);
        for (int c = 0; c < CL; c++) st_dsmem_u32(mapa_u32(a0, (uint32_t)c), v);
    };

    const int nfail = *a.fail_count;
    if (nfail < a.nfail_lo || nfail > a.nfail_hi) return; // another cluster size takes this chunk (uniform over the grid)
    for (int f = cid; f < nfail; f += nclusters) {
        const long long shot = a.fail_list[f];
        const real *llr = a.llr + (a.llr_by_shot ? shot : (long long)f) * n;
        cluster_sync_all();

        // ---- a9: ascending order on (llr, index): cluster-wide bitonic sort of the distinct pairs (key, index) ----
        for (long long j = gt; j < a.np2; j += GT) {
            key[j] = (j < n) ? sort_key(llr[j]) : ~0ull;
            order[j] = (j < n) ? (unsigned)j : 0xFFFFFFFFu;
        }
        for (int w = tid; w < pw; w += T) {
            unsigned sb = 0;
            for (int b = 0; b < 32; b++) {
                const int i = w * 32 + b;
                if (i < nloc) sb |= synd_bit(a.synd, shot, m, base + i, a.synd_packed) << b;
            }
            sbit[w] = sb;
            used[w] = 0;
        }
        cluster_sync_all();
        for (int k = 2; k <= a.np2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (long long t = gt; t < (a.np2 >> 1); t += GT) {
                    const long long i = ((t & ~(long long)(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const unsigned long long ka = key[i], kb = key[l];
                    const unsigned ia = order[i], ib = order[l];
                    const bool gtr = ka > kb || (ka == kb && ia > ib);
                    if (gtr == ((i & k) == 0)) { key[i] = kb; key[l] = ka; order[i] = ib; order[l] = ia; }
                }
                cluster_sync_all();
            }

        // ---- a10: elimination, one panel of 64 sorted columns at a time ----
        int rank_tot = 0, np_cnt = 0, colstep = 0, pairstep = 0;
        for (int w = 0; w < a.npanels && rank_tot < a.maxrank; w++) {
            unsigned long long *Mw = maskbase + (size_t)w * mpad;
            for (int i = tid; i < rpc; i += T) P[i] = 0;
            for (int i = tid; i < 64 * pw; i += T) plane[i] = 0;
            __syncthreads();
            if (tid < 64) {
                const int t = w * 64 + tid;
                if (t < n) {
                    const int j = (int)order[t];
                    for (int q = g.col_ptr[j]; q < g.col_ptr[j + 1]; q++) {
                        const int i = g.row_idx[q] - base;
                        if (i >= 0 && i < nloc) atomicOr(&P[i], 1ull << tid);
                    }
                }
            }
            // replay the row operations of every earlier panel that holds pivots, in order
            const int npairs = np_cnt;
            if (tid == 0) // first replays' masks into the ring
                for (int s = 0; s < npairs && s < kOsdcStages; s++) {
                    osdc_mbar_expect(mbar0 + 8u * ist, ring_bytes);
                    osdc_bulk_g2s(smem_u32(ring + (size_t)ist * rpc), maskbase + (size_t)plist[s] * mpad + base, ring_bytes, mbar0 + 8u * ist);
                    ist = ist + 1 == (uint32_t)kOsdcStages ? 0u : ist + 1;
                }
            OsdcPivot rec{0ull, 0ull, -1, 0};
            if (npairs > 0 && tid < 64) {
                const int ps = pfirst[0], cnt = pfirst[1] - ps;
                if (tid < cnt) rec = piv[ps + tid];
            }
            __syncthreads();
            for (int s = 0; s < npairs; s++) {
                const int ps = pfirst[s], cnt = pfirst[s + 1] - ps;
                const int par = pairstep & 1;
                pairstep++;
                // owners hand the current words of panel q's pivot rows to every CTA
                if (tid < 64) {
                    cfC[tid] = rec.C; cfF[tid] = rec.F; cfRow[tid] = rec.row;
                    if (tid < cnt) {
                        const int li = rec.row - base;
                        if (li >= 0 && li < nloc) push_all_u64(&gath[par * 64 + tid], P[li]);
                    }
                    // the next replay's pivots (global, L2): in flight during the rest of this one
                    rec = OsdcPivot{0ull, 0ull, -1, 0};
                    if (s + 1 < npairs) {
                        const int ps2 = pfirst[s + 1], cnt2 = pfirst[s + 2] - ps2;
                        if (tid < cnt2) rec = piv[ps2 + tid];
                    }
                }
                cluster_sync_all();
                {
                    // R = C g and the final pivot rows F g: pivot k = tid / 8, eight threads (lanes of one warp) share its 64 terms
                    const int k = tid >> 3, sl = tid & 7;
                    unsigned long long R = 0, fin = 0;
                    if (k < cnt && 8 * sl < cnt) {
                        const unsigned cb = (unsigned)(cfC[k] >> (8 * sl)) & 255u, fb = (unsigned)(cfF[k] >> (8 * sl)) & 255u;
                        const unsigned long long *gp = gath + par * 64 + 8 * sl;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const unsigned long long gv = gp[j];
                            R ^= ((cb >> j) & 1u) ? gv : 0ull;
                            fin ^= ((fb >> j) & 1u) ? gv : 0ull;
                        }
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) { R ^= osdc_shfl_xor64(R, o); fin ^= osdc_shfl_xor64(fin, o); }
                    if (sl == 0 && k < 64) {
                        Rk[k] = (k < cnt) ? R : 0ull;
                        if (k < cnt) {
                            const int li = cfRow[k] - base;
                            if (li >= 0 && li < nloc) P[li] = fin; // the apply below skips pivot rows of q (their stored mask is 0)
                        }
                    }
                }
                __syncthreads();
                const int ntab = (cnt + 7) >> 3;
                for (int e = tid; e < ntab * 256; e += T) {
                    unsigned long long v = 0;
                    const int b0 = (e >> 8) * 8;
#pragma unroll
                    for (int b = 0; b < 8; b++) v ^= ((e >> b) & 1) ? Rk[b0 + b] : 0ull;
                    tab[e] = v;
                }
                osdc_mbar_wait(mbar0 + 8u * cst, cpar);
                __syncthreads();
                const unsigned long long *Mq = ring + (size_t)cst * rpc;
                for (int i = tid; i < nloc; i += T) {
                    const unsigned long long mk = Mq[i];
                    if (mk) {
                        unsigned long long v = 0;
#pragma unroll
                        for (int b = 0; b < 8; b++)
                            if (b < ntab) v ^= tab[b * 256 + (int)((mk >> (8 * b)) & 255ull)];
                        P[i] ^= v;
                    }
                }
                cst++;
                if (cst == (uint32_t)kOsdcStages) { cst = 0; cpar ^= 1u; }
                __syncthreads();
                if (tid == 0 && s + kOsdcStages < npairs) { // the ring stage is free again: next replay's masks
                    osdc_mbar_expect(mbar0 + 8u * ist, ring_bytes);
                    osdc_bulk_g2s(smem_u32(ring + (size_t)ist * rpc), maskbase + (size_t)plist[s + kOsdcStages] * mpad + base, ring_bytes, mbar0 + 8u * ist);
                    ist = ist + 1 == (uint32_t)kOsdcStages ? 0u : ist + 1;
                }
            }
            __syncthreads();

            // factorise this panel
            int cnt = 0;
            const int ps_w = rank_tot;
            for (int c = 0; c < 64; c++) {
                const int t = w * 64 + c;
                if (t >= n || rank_tot >= a.maxrank) break;
                const int par = colstep & 1;
                colstep++;
                int best = 0x7fffffff;
                for (int i = tid; i < nloc; i += T)
                    if (((P[i] >> c) & 1ull) && !((used[i >> 5] >> (i & 31)) & 1u)) { best = i; break; }
                for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (lane == 0) red[warp] = best;
                __syncthreads();
                if (warp == 0) {
                    int b = (lane < nwarps) ? red[lane] : 0x7fffffff;
                    for (int o = 16; o > 0; o >>= 1) b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
                    if (lane == 0) { // this CTA's proposal to every CTA
                        const bool have = b != 0x7fffffff;
                        push_all_u32(&candr[par * 16 + rank], have ? (unsigned)(base + b) : 0xFFFFFFFFu);
                        push_all_u64(&candp[par * 16 + rank], have ? P[b] : 0ull);
                        push_all_u32(&cands[par * 16 + rank], have ? ((sbit[b >> 5] >> (b & 31)) & 1u) : 0u);
                    }
                }
                cluster_sync_all();
                int win = -1;
                for (int r2 = 0; r2 < CL; r2++)
                    if (candr[par * 16 + r2] != 0xFFFFFFFFu) { win = r2; break; }
                if (win < 0) continue; // dependent column: not a pivot (uniform over the cluster)
                const int p = (int)candr[par * 16 + win];
                const unsigned long long Pp = candp[par * 16 + win];
                const unsigned sp = cands[par * 16 + win];
                const int k = cnt;
                const int lp = p - base; // local index of the pivot if it lives here
                for (int i0 = warp * 32; i0 < nloc; i0 += T) {
                    const int i = i0 + lane;
                    const bool hit = i < nloc && i != lp && ((P[i] >> c) & 1ull);
                    if (hit) P[i] ^= Pp;
                    const unsigned bal = __ballot_sync(0xffffffffu, hit);
                    if (lane == 0 && bal) {
                        plane[k * pw + (i0 >> 5)] = bal;
                        if (sp) sbit[i0 >> 5] ^= bal;
                    }
                }
                if (tid == 0) {
                    prow_s[k] = p;
                    if (lp >= 0 && lp < nloc) used[lp >> 5] |= 1u << (lp & 31);
                    if (rank == 0) { piv[rank_tot].row = p; piv[rank_tot].pos = t; }
                }
                cnt++; rank_tot++;
                __syncthreads();
            }
            if (cnt > 0) {
                // multiplier masks of this panel's pivot rows (kept apart: the replay resolves those rows through C / F),
                // taken out of the bit planes so that their stored masks are 0
                if (tid < cnt) {
                    const int li = prow_s[tid] - base;
                    if (li >= 0 && li < nloc) {
                        unsigned long long mj = 0;
                        for (int k2 = 0; k2 < cnt; k2++) {
                            const unsigned wv = plane[k2 * pw + (li >> 5)];
                            if ((wv >> (li & 31)) & 1u) { mj |= 1ull << k2; atomicAnd(&plane[k2 * pw + (li >> 5)], ~(1u << (li & 31))); }
                        }
                        push_all_u64(&mjs[tid], mj);
                    }
                }
                cluster_sync_all();
                // C = (I + L)^-1 with L the strictly lower part of the multiplier matrix (pivot k takes pivot k' < k): row k of C
                // is e_k + sum of the rows k' < k that k takes; F = (I + U) C with U the strictly upper part
                if (warp == 0) {
                    for (int k = 0; k < cnt; k++) {
                        const unsigned long long mj = mjs[k];
                        unsigned long long part = 0;
                        if (lane < k && ((mj >> lane) & 1ull)) part ^= Cs[lane];
                        if (lane + 32 < k && ((mj >> (lane + 32)) & 1ull)) part ^= Cs[lane + 32];
                        for (int o = 16; o > 0; o >>= 1) part ^= osdc_shfl_xor64(part, o);
                        if (lane == 0) Cs[k] = part ^ (1ull << k);
                        __syncwarp();
                    }
                    for (int k = lane; k < cnt; k += 32) {
                        const unsigned long long mj = mjs[k];
                        unsigned long long fv = Cs[k];
                        for (int k2 = k + 1; k2 < cnt; k2++)
                            if ((mj >> k2) & 1ull) fv ^= Cs[k2];
                        Fs[k] = fv;
                    }
                    __syncwarp();
                    if (rank == 0)
                        for (int k = lane; k < cnt; k += 32) { piv[ps_w + k].C = Cs[k]; piv[ps_w + k].F = Fs[k]; }
                }
                __syncthreads();
                // bit planes -> one 64-bit mask per check, to HBM (read back by this CTA only, through the async proxy)
                for (int i = tid; i < rpc; i += T) {
                    unsigned long long mk = 0;
                    if (i < nloc)
                        for (int k2 = 0; k2 < cnt; k2++) mk |= (unsigned long long)((plane[k2 * pw + (i >> 5)] >> (i & 31)) & 1u) << k2;
                    Mw[base + i] = mk;
                }
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                if (tid == 0) { plist[np_cnt] = w; pfirst[np_cnt] = ps_w; pfirst[np_cnt + 1] = rank_tot; }
                np_cnt++;
            }
            cluster_sync_all(); // pivot records of this panel (written by CTA 0) are visible to every CTA's next replay
        }

        // ---- a11: OSD-0 read-out.  Jordan form: x[pivot column of check p] = transformed syndrome bit of p
        const long long obase = shot * (long long)n;
        for (long long j = gt; j < n; j += GT) {
            if (a.osd0) a.osd0[obase + j] = 0;
            if (a.osdw) a.osdw[obase + j] = 0;
        }
        cluster_sync_all();
        for (int r = tid; r < rank_tot; r += T) {
            const int li = piv[r].row - base;
            if (li >= 0 && li < nloc && ((sbit[li >> 5] >> (li & 31)) & 1u)) {
                const int j = (int)order[piv[r].pos];
                if (a.osd0) a.osd0[obase + j] = 1;
                if (a.osdw) a.osdw[obase + j] = 1;
            }
        }
        if (rank == 0 && tid == 0 && a.stat) atomicAdd(&a.stat[2], 1ull);
    }
    cluster_sync_all(); // no CTA exits while another may still address its shared memory
}

} // namespace bposd
