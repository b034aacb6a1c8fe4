// bp_cluster_kernel.cuh -- min-sum BP for parity-check matrices whose messages exceed one SM's shared
// memory (BASELINE config 5: 134 400 edges, 1.07 MB of fp64 messages): the in-place message array of
// bp_fast_kernel is split over the CTAs of a thread-block cluster and the bit sweep reaches the other
// CTAs' slices through distributed shared memory (ld/st/atom.shared::cluster).
//
//   * physical rows (checks) are partitioned over the CL CTAs of a cluster, rows_per_cta each; the check
//     sweep is entirely local (the same fast_check_row as the single-CTA kernel, bit-exact in fp64);
//   * bits are partitioned too (bits_per_cta each); the thread that owns a bit gathers / scatters its
//     <= DV slots with 32-bit cluster-window addresses resolved once per CTA with `mapa`; the host
//     partition pass (cluster_build) places bits and checks to keep as many edges CTA-local as it can;
//   * one parity-mismatch bit per check, toggled with atom.shared::cluster.xor when a hard decision flips;
//     every CTA votes with __syncthreads_and, the votes are exchanged through DSMEM and ride on the cluster
//     barrier that separates the sweeps (two barrier.cluster per iteration);
//   * persistent clusters pull shots from the same atomic queue as the other BP kernels.
// A shot that needs all max_iter = n iterations costs tens of microseconds per iteration here instead
// of the millisecond of the HBM-scratch kernel, which is what bounds config 5's tail.
#pragma once
#include "bp_fast_kernel.cuh"

namespace bposd {

#define BPC_NONE 0xFFFFFFFFu

struct ClusterTables {
    int DC = 0, DV = 0, regular = 0;
    int CL = 0, rows_per_cta = 0, bits_per_cta = 0, elem_bytes = 0;
    uint32_t *d_vslot = nullptr;  // [CL*bits_per_cta, DV] physical slot of the k-th edge of the bit at position q
    uint8_t *d_cdeg = nullptr;    // [CL*rows_per_cta] degree of the check in physical row p (0: absent)
    uint32_t *d_row_of = nullptr; // [CL*rows_per_cta] original check of physical row p, BPC_NONE: absent
    uint32_t *d_bit_of = nullptr; // [CL*bits_per_cta] original bit at position q, BPC_NONE: absent
    long long remote_edges = 0, total_edges = 0;
};

struct ClusterDev {
    const uint32_t *vslot;
    const uint8_t *cdeg;
    const uint32_t *row_of;
    const uint32_t *bit_of;
    int rows_per_cta, bits_per_cta, CL;
    int flip_table; // 1: per-edge parity-flip descriptors live in shared memory (bits_per_cta * DV words)
};

static inline void cluster_free(ClusterTables &t) {
    cudaFree(t.d_vslot); cudaFree(t.d_cdeg); cudaFree(t.d_row_of); cudaFree(t.d_bit_of);
    t.d_vslot = nullptr; t.d_cdeg = nullptr; t.d_row_of = nullptr; t.d_bit_of = nullptr;
    t.CL = 0;
}

template <typename real>
static inline size_t cluster_smem_bytes(int DC, int rows_per_cta, int bits_per_cta, bool with_priors, int DV = 0) {
    size_t msgs = ((size_t)rows_per_cta * fast_row_stride(DC, (int)sizeof(real)) * sizeof(real) + 15) / 16 * 16;
    size_t meta = ((size_t)rows_per_cta + 15) / 16 * 16;
    size_t prior = with_priors ? ((size_t)bits_per_cta * sizeof(real) + 15) / 16 * 16 : 0;
    size_t flip = (size_t)bits_per_cta * DV * 4; // DV > 0: with the parity-flip descriptor table
    return msgs + meta + prior + flip + 16;
}

// Host partition pass: alternate "every bit goes to the CTA that holds most of its checks" and "every check
// goes to the CTA that holds most of its bits" under the capacity of a CTA, starting from contiguous blocks
// of checks; keep the assignment with the fewest remote edges.  Inside a CTA, checks and bits stay in
// ascending index order, so the result writes remain mostly coalesced.
static inline cudaError_t cluster_build(ClusterTables &t, int CL, int m, int n, const std::vector<int> &row_ptr,
                                        const std::vector<int> &col_idx, const std::vector<int> &col_ptr,
                                        const std::vector<int> &row_idx, const std::vector<int> &csc_slot, int elem_bytes) {
    int mr = 0, mc = 0, minr = 1 << 30, minc = 1 << 30;
    for (int i = 0; i < m; i++) { int d = row_ptr[i + 1] - row_ptr[i]; mr = std::max(mr, d); minr = std::min(minr, d); }
    for (int j = 0; j < n; j++) { int d = col_ptr[j + 1] - col_ptr[j]; mc = std::max(mc, d); minc = std::min(minc, d); }
    cluster_free(t);
    fast_class(mc, mr, &t.DC, &t.DV);
    const int rpc = (m + CL - 1) / CL, bpc = (n + CL - 1) / CL;
    t.CL = CL; t.rows_per_cta = rpc; t.bits_per_cta = bpc;
    t.regular = (m > 0 && minr == t.DC && mr == t.DC && minc == t.DV && mc == t.DV && m % CL == 0) ? 1 : 0;
    const int E = row_ptr[m];
    const int RS = fast_row_stride(t.DC, elem_bytes);
    t.elem_bytes = elem_bytes;
    std::vector<int> row_cta(m), bit_cta(n, 0), best_row, best_bit;
    for (int i = 0; i < m; i++) row_cta[i] = i / rpc;
    long long best_cut = -1;
    std::vector<int> tally(CL), load(CL);
    for (int round = 0; round < 8; round++) {
        std::fill(load.begin(), load.end(), 0);
        for (int j = 0; j < n; j++) {
            std::fill(tally.begin(), tally.end(), 0);
            for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) tally[row_cta[row_idx[q]]]++;
            int best = -1;
            for (int c = 0; c < CL; c++)
                if (load[c] < bpc && (best < 0 || tally[c] > tally[best] || (tally[c] == tally[best] && load[c] < load[best]))) best = c;
            bit_cta[j] = best; load[best]++;
        }
        long long cut = 0;
        for (int i = 0; i < m; i++)
            for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) cut += (bit_cta[col_idx[e]] != row_cta[i]) ? 1 : 0;
        if (best_cut < 0 || cut < best_cut) { best_cut = cut; best_row = row_cta; best_bit = bit_cta; }
        std::fill(load.begin(), load.end(), 0);
        for (int i = 0; i < m; i++) {
            std::fill(tally.begin(), tally.end(), 0);
            for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) tally[bit_cta[col_idx[e]]]++;
            int best = -1;
            for (int c = 0; c < CL; c++)
                if (load[c] < rpc && (best < 0 || tally[c] > tally[best] || (tally[c] == tally[best] && load[c] < load[best]))) best = c;
            row_cta[i] = best; load[best]++;
        }
    }
    t.remote_edges = best_cut; t.total_edges = E;
    // physical rows and bit positions
    std::vector<int> prow(m), pos(n), cnt(CL, 0);
    for (int i = 0; i < m; i++) prow[i] = best_row[i] * rpc + cnt[best_row[i]]++;
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int j = 0; j < n; j++) pos[j] = best_bit[j] * bpc + cnt[best_bit[j]]++;
    std::vector<uint32_t> vs((size_t)CL * bpc * t.DV, BPC_NONE), rowof((size_t)CL * rpc, BPC_NONE), bitof((size_t)CL * bpc, BPC_NONE);
    std::vector<uint8_t> cd((size_t)CL * rpc, 0);
    for (int i = 0; i < m; i++) { rowof[prow[i]] = (uint32_t)i; cd[prow[i]] = (uint8_t)(row_ptr[i + 1] - row_ptr[i]); }
    for (int j = 0; j < n; j++) {
        bitof[pos[j]] = (uint32_t)j;
        for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) {
            const int i = row_idx[q], e = csc_slot[q];
            vs[(size_t)pos[j] * t.DV + (q - col_ptr[j])] = (uint32_t)(prow[i] * RS + (e - row_ptr[i]));
        }
    }
    cudaError_t e = cudaMalloc((void **)&t.d_vslot, vs.size() * 4);
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_cdeg, cd.size());
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_row_of, rowof.size() * 4);
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_bit_of, bitof.size() * 4);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_vslot, vs.data(), vs.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_row_of, rowof.data(), rowof.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_bit_of, bitof.data(), bitof.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(t.d_cdeg, cd.data(), cd.size(), cudaMemcpyHostToDevice);
}

// ---- distributed-shared-memory primitives -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem(uint32_t a, double) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float ld_dsmem(uint32_t a, float) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void st_dsmem(uint32_t a, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void st_dsmem(uint32_t a, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void st_dsmem_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void st_dsmem_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void xor_dsmem_u32(uint32_t a, uint32_t v) { asm volatile("red.shared::cluster.xor.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Register-frugal form of the check update (row a4) for CTAs of up to 1024 threads (64 registers each):
// pass 1 streams the row once for (smallest, second smallest, position of the smallest, sign parity), pass 2
// streams it again and overwrites every slot with "minimum over the other edges" = second smallest at the
// position of the smallest, smallest elsewhere -- the same values as the prefix/suffix form of
// fast_check_row (min is exact; on a tie both forms give the tied value).  "<= 0" counts zero as negative.
template <typename real, int DC, bool REG>
__device__ __forceinline__ void cluster_check_row(real *row, unsigned mt, real alpha) {
    constexpr int CH = (sizeof(real) == 8) ? 2 : ((DC % 4 == 0) ? 4 : 2); // elements per vector access
    real min1 = real_max<real>(), min2 = real_max<real>();
    int arg = -1;
    unsigned par = mt >> 7;
    const int deg = REG ? DC : (int)((mt >> 1) & 0x1f);
#pragma unroll
    for (int k0 = 0; k0 < DC; k0 += CH) {
        real v[CH];
        RowIO<real, CH>::load(row + k0, v);
#pragma unroll
        for (int u = 0; u < CH; u++) {
            const real av = abs_bits(v[u]);
            par ^= (sign_word(v[u]) >> 31) | (av == (real)0 ? 1u : 0u);
            if (av < min1) { min2 = min1; min1 = av; arg = k0 + u; }
            else if (av < min2) min2 = av;
        }
    }
#pragma unroll
    for (int k0 = 0; k0 < DC; k0 += CH) {
        real v[CH], out[CH];
        RowIO<real, CH>::load(row + k0, v);
#pragma unroll
        for (int u = 0; u < CH; u++) {
            const unsigned neg = (sign_word(v[u]) >> 31) | (abs_bits(v[u]) == (real)0 ? 1u : 0u);
            const real mag = (k0 + u == arg) ? min2 : min1;
            out[u] = mag * (((par ^ neg) & 1u) ? -alpha : alpha);
            if (!REG && k0 + u >= deg) out[u] = real_max<real>();
        }
        RowIO<real, CH>::store(row + k0, out);
    }
}

// MAXT (a multiple of 128 >= blockDim.x) sets the register budget: 65536 / MAXT per thread, so that the per-thread
// slot addresses never spill -- a spill costs an L2 round trip per iteration here, because barrier.cluster
// invalidates L1.
template <typename real, int DC, int DV, int VPT, int MAXT, bool REG>
__global__ void __launch_bounds__(MAXT, 1) bp_cluster_kernel(BpArgs<real> a, ClusterDev t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.g.m, n = a.g.n;
    const int tid = threadIdx.x, T = blockDim.x;
    const int rpc = t.rows_per_cta, bpc = t.bits_per_cta, CL = t.CL;
    const uint32_t rank = cluster_ctarank();
    constexpr int RS = fast_row_stride(DC, (int)sizeof(real));                                 // row stride in elements
    real *msg = reinterpret_cast<real *>(smem_raw);                                            // [rpc * RS] this CTA's rows
    uint8_t *meta = smem_raw + ((size_t)rpc * RS * sizeof(real) + 15) / 16 * 16;               // bit0 mismatch, bits1-5 degree, bit7 syndrome
    real *prior_s = reinterpret_cast<real *>(meta + ((size_t)rpc + 15) / 16 * 16);             // [bpc] priors by position (non-uniform only)
    uint32_t *flip_desc = reinterpret_cast<uint32_t *>(prior_s + ((size_t)bpc * sizeof(real) + 15) / 16 * 16 / sizeof(real)); // [bpc * DV]
    __shared__ long long sh_shot;
    __shared__ int sh_slot;
    __shared__ unsigned sh_vote[2][16];
    const uint32_t msg_s = smem_u32(msg), meta_s = smem_u32(meta);
    const unsigned slots_per_cta = (unsigned)rpc * RS;
    // A hard-decision flip toggles the parity-mismatch bit of every neighbouring check, wherever it lives.  The
    // cluster address of that check's meta word (4-byte aligned) and its byte lane (low two bits) are resolved once
    // per CTA; doing it at flip time costs a global load per edge that always misses L1 (barrier.cluster invalidates
    // it), and flips are frequent on the shots that matter for the tail (BP that oscillates for thousands of passes).
    if (t.flip_table)
        for (int e = tid; e < bpc * DV; e += T) {
            const uint32_t s = t.vslot[(size_t)rank * bpc * DV + e];
            uint32_t d = 0;
            if (s != BPC_NONE) {
                const uint32_t prow = s / (unsigned)RS, lp = prow % (unsigned)rpc;
                d = mapa_u32(meta_s + (lp & ~3u), prow / (unsigned)rpc) | (lp & 3u);
            }
            flip_desc[e] = d;
        }
    __syncthreads();

    // cluster-window addresses of the slots of this thread's bits (shot independent).  Local and remote slots are
    // addressed alike: splitting them (plain ld/st.shared for CTA-local slots) was measured and made no difference,
    // the extra predicates and selects cost as much as the narrower path saves (profiles/r01k_cluster_probe.log).
    uint32_t off[VPT][DV];
    int dj[VPT];
    unsigned valid = 0;
#pragma unroll
    for (int r = 0; r < VPT; r++) {
        const int lq = tid + r * T;
        const size_t q = (size_t)rank * bpc + lq;
        dj[r] = 0;
        const bool have = lq < bpc && t.bit_of[q] != BPC_NONE;
        valid |= have ? (1u << r) : 0u;
#pragma unroll
        for (int k = 0; k < DV; k++) {
            const uint32_t s = have ? t.vslot[q * DV + k] : BPC_NONE;
            off[r][k] = 0;
            if (s != BPC_NONE) {
                off[r][k] = mapa_u32(msg_s + (s % slots_per_cta) * (unsigned)sizeof(real), s / slots_per_cta);
                dj[r]++;
            }
        }
    }
    unsigned long long n_conv = 0, n_iter = 0;
    const bool uniform = a.uniform_prior != 0;
    const real prior_u = a.prior[0];
    (void)m;

    for (;;) {
        if (rank == 0 && tid == 0) {
            const unsigned long long s = atomicAdd(a.queue, 1ull);
            for (int c = 0; c < CL; c++) st_dsmem_u64(mapa_u32(smem_u32(&sh_shot), c), s);
        }
        cluster_sync_all();
        const long long shot = sh_shot;
        if (shot >= a.B) break;
        const real *prior = a.prior + shot * a.prior_stride;

        for (int p = tid; p < rpc; p += T) {
            const uint32_t orig = t.row_of[(size_t)rank * rpc + p];
            const unsigned s = (orig != BPC_NONE) ? synd_bit(a.synd, shot, a.g.m, (int)orig, a.synd_packed) : 0u;
            const unsigned deg = t.cdeg[(size_t)rank * rpc + p];
            meta[p] = (uint8_t)(s | (deg << 1) | (s << 7));
            if (!REG)
                for (int k = (int)deg; k < DC; k++) msg[p * RS + k] = real_max<real>();
        }
        real llr[VPT];
        unsigned dprev = 0;
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            llr[r] = 0;
            if ((valid >> r) & 1u) {
                const int lq = tid + r * T;
                const real pj = uniform ? prior_u : prior[t.bit_of[(size_t)rank * bpc + lq]];
                if (!uniform) prior_s[lq] = pj;
                llr[r] = pj;
#pragma unroll
                for (int k = 0; k < DV; k++)
                    if (REG || k < dj[r]) st_dsmem(off[r][k], pj);
            }
        }
        cluster_sync_all();

        bool conv = false;
        int iters = 0;
        real pow2 = 1;
        for (int it = 1;; it++) {
            const bool last = it > a.max_iter;
            pow2 *= (real)0.5;
            const real alpha = (a.alpha0 == (real)0) ? (real)1 - pow2 : a.alpha0;
            bool ok = true;
            // ---- check sweep over this CTA's rows (a4) + local convergence vote for the previous pass (a7)
            for (int p = tid; p < rpc; p += T) {
                const unsigned mt = meta[p];
                if (mt & 1u) ok = false;
                if (last) continue;
                cluster_check_row<real, DC, REG>(msg + (size_t)p * RS, mt, alpha);
            }
            const int cta_ok = __syncthreads_and(ok ? 1 : 0);
            if (tid < CL) st_dsmem_u32(mapa_u32(smem_u32(&sh_vote[it & 1][rank]), tid), (uint32_t)cta_ok);
            cluster_sync_all();
            int all_ok = 1;
            for (int c = 0; c < CL; c++) all_ok &= (int)sh_vote[it & 1][c];
            if (it > 1 && all_ok) { conv = true; iters = it - 1; break; }
            if (last) { iters = a.max_iter; break; }
            // ---- bit sweep (a6 + a8) through distributed shared memory
            // remote loads take ~200 cycles: issue the gathers of four bits back to back before using any
            unsigned dnow = 0;
            constexpr int BATCH = VPT < 4 ? VPT : 4;
#pragma unroll
            for (int r0 = 0; r0 < VPT; r0 += BATCH) {
                real c[BATCH][DV];
#pragma unroll
                for (int u = 0; u < BATCH; u++) {
                    const int r = r0 + u;
#pragma unroll
                    for (int k = 0; k < DV; k++)
                        c[u][k] = (r < VPT && ((valid >> r) & 1u) && (REG || k < dj[r])) ? ld_dsmem(off[r][k], (real)0) : (real)0;
                }
#pragma unroll
                for (int u = 0; u < BATCH; u++) {
                    const int r = r0 + u;
                    if (r < VPT && ((valid >> r) & 1u)) {
                        real pre[DV];
                        real tt = uniform ? prior_u : prior_s[tid + r * T];
#pragma unroll
                        for (int k = 0; k < DV; k++)
                            if (REG || k < dj[r]) { pre[k] = tt; tt += c[u][k]; }
                        llr[r] = tt;
                        dnow |= ((tt <= 0) ? 1u : 0u) << r;
                        real sfx = 0;
#pragma unroll
                        for (int k = DV - 1; k >= 0; k--)
                            if (REG || k < dj[r]) {
                                st_dsmem(off[r][k], (REG && k == DV - 1) ? pre[k] : pre[k] + sfx);
                                sfx = (REG && k == DV - 1) ? c[u][k] : sfx + c[u][k];
                            }
                    }
                }
            }
            if (dnow != dprev) {
                // a hard decision flipped (rare): toggle the parity-mismatch bit of every neighbouring check
                unsigned flip = dnow ^ dprev;
                dprev = dnow;
#pragma unroll
                for (int r = 0; r < VPT; r++)
                    if ((flip >> r) & 1u) {
                        const size_t q = (size_t)rank * bpc + tid + r * T;
                        for (int k = 0; k < dj[r]; k++) {
                            if (t.flip_table) {
                                const uint32_t d = flip_desc[(tid + r * T) * DV + k];
                                xor_dsmem_u32(d & ~3u, 1u << ((d & 3u) * 8u));
                            } else {
                                const uint32_t prow = t.vslot[q * DV + k] / (unsigned)RS;
                                const uint32_t lp = prow % (unsigned)rpc;
                                xor_dsmem_u32(mapa_u32(meta_s + (lp & ~3u), prow / (unsigned)rpc), 1u << ((lp & 3u) * 8u));
                            }
                        }
                    }
            }
            cluster_sync_all();
        }

        // ---- results ----
        const bool final_here = conv || a.osd_off;
        if (!final_here) {
            if (rank == 0 && tid == 0) {
                const int slot = atomicAdd(a.fail_count, 1);
                a.fail_list[slot] = (int)shot;
                for (int c = 0; c < CL; c++) st_dsmem_u32(mapa_u32(smem_u32(&sh_slot), c), (uint32_t)slot);
            }
            cluster_sync_all();
        }
        const long long base = shot * (long long)n;
#pragma unroll
        for (int r = 0; r < VPT; r++)
            if ((valid >> r) & 1u) {
                const uint32_t j = t.bit_of[(size_t)rank * bpc + tid + r * T];
                const uint8_t d = (llr[r] <= 0) ? 1 : 0;
                if (a.bp) a.bp[base + j] = d;
                if (final_here) {
                    if (a.osd0) a.osd0[base + j] = d;
                    if (a.osdw) a.osdw[base + j] = d;
                }
                if (a.llr) a.llr[base + j] = llr[r];
                else if (!final_here) a.fail_llr[(long long)sh_slot * n + j] = llr[r];
            }
        if (rank == 0 && tid == 0) {
            if (a.converge) a.converge[shot] = conv ? 1 : 0;
            if (a.iter) a.iter[shot] = iters;
            n_conv += conv ? 1 : 0;
            n_iter += (unsigned long long)iters;
        }
    }
    if (rank == 0 && tid == 0 && a.stat) {
        atomicAdd(&a.stat[0], n_conv);
        atomicAdd(&a.stat[1], n_iter);
    }
}

// ---- dispatch ------------------------------------------------------------------------------------
// bits per thread: the smallest of {1, 2, 3, 4, 6, 8} that lets a CTA of <= 1024 threads cover its bits
static inline int cluster_vpt(int bits_per_cta) {
    for (int v : {1, 2, 3, 4, 6, 8}) // prefer CTAs of <= 640 threads: 96+ registers per thread, no spills in fp64
        if ((bits_per_cta + v - 1) / v <= 640) return v;
    for (int v : {6, 8})
        if ((bits_per_cta + v - 1) / v <= 1024) return v;
    return 0;
}
static inline int cluster_threads(int bits_per_cta) {
    const int v = cluster_vpt(bits_per_cta);
    return v ? std::max(32, ((bits_per_cta + v - 1) / v + 31) / 32 * 32) : 0;
}

// Instantiated per (precision, degree class) in its own translation unit (bp_cluster_inst.cu); see FastInst.
template <typename real, int DC, int DV>
struct ClusterInst {
    static cudaError_t prepare(const ClusterTables &t, int threads, size_t smem, int *max_clusters);
    static cudaError_t launch(const ClusterTables &t, const BpArgs<real> &a, const ClusterDev &d, int nclusters, int threads, size_t smem, cudaStream_t st);
};

#ifdef BPOSD_CLUSTER_INSTANTIATE
#define BPOSD_CL_REG2(EXPR) do { if (reg__) { constexpr bool REG = true; EXPR; } else { constexpr bool REG = false; EXPR; } } while (0)
// (bits per thread, CTA size class) pairs that cluster_vpt / cluster_threads can produce: up to 640 threads for every
// VPT, more only for VPT 6 (896, 1024) and 8 (768, 896, 1024)
#define BPOSD_CL_REG(EXPR)                                                                       \
    do {                                                                                         \
        if (maxt__ <= 512) { constexpr int MAXT = 512; BPOSD_CL_REG2(EXPR); }                    \
        else { constexpr int MAXT = 640; BPOSD_CL_REG2(EXPR); }                                  \
    } while (0)
#define BPOSD_CL_REG_BIG(EXPR)                                                                   \
    do {                                                                                         \
        if (maxt__ <= 512) { constexpr int MAXT = 512; BPOSD_CL_REG2(EXPR); }                    \
        else if (maxt__ <= 640) { constexpr int MAXT = 640; BPOSD_CL_REG2(EXPR); }               \
        else if (maxt__ <= 768) { constexpr int MAXT = 768; BPOSD_CL_REG2(EXPR); }               \
        else if (maxt__ <= 896) { constexpr int MAXT = 896; BPOSD_CL_REG2(EXPR); }               \
        else { constexpr int MAXT = 1024; BPOSD_CL_REG2(EXPR); }                                 \
    } while (0)
#define BPOSD_CL_GEOM(t, EXPR)                                                                   \
    do {                                                                                         \
        const int vpt__ = cluster_vpt(t.bits_per_cta);                                           \
        const int maxt__ = cluster_threads(t.bits_per_cta);                                      \
        const bool reg__ = t.regular != 0;                                                       \
        if (vpt__ == 1) { constexpr int VPT = 1; BPOSD_CL_REG(EXPR); }                           \
        else if (vpt__ == 2) { constexpr int VPT = 2; BPOSD_CL_REG(EXPR); }                      \
        else if (vpt__ == 3) { constexpr int VPT = 3; BPOSD_CL_REG(EXPR); }                      \
        else if (vpt__ == 4) { constexpr int VPT = 4; BPOSD_CL_REG(EXPR); }                      \
        else if (vpt__ == 6) { constexpr int VPT = 6; BPOSD_CL_REG_BIG(EXPR); }                  \
        else { constexpr int VPT = 8; BPOSD_CL_REG_BIG(EXPR); }                                  \
    } while (0)

template <typename real, int DC, int DV>
cudaError_t ClusterInst<real, DC, DV>::prepare(const ClusterTables &t, int threads, size_t smem, int *max_clusters) {
    cudaError_t e = cudaSuccess;
    BPOSD_CL_GEOM(t, {
        auto kern = (bp_cluster_kernel<real, DC, DV, VPT, MAXT, REG>);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && t.CL > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(t.CL, 1, 1); cfg.blockDim = dim3(threads, 1, 1); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = t.CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            e = cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg);
        }
    });
    return e;
}

template <typename real, int DC, int DV>
cudaError_t ClusterInst<real, DC, DV>::launch(const ClusterTables &t, const BpArgs<real> &a, const ClusterDev &d, int nclusters, int threads,
                                              size_t smem, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    BPOSD_CL_GEOM(t, {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nclusters * t.CL, 1, 1); cfg.blockDim = dim3(threads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = t.CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, bp_cluster_kernel<real, DC, DV, VPT, MAXT, REG>, a, d);
    });
    return e;
}
#endif // BPOSD_CLUSTER_INSTANTIATE

template <typename real>
static inline cudaError_t cluster_prepare(const ClusterTables &t, int threads, size_t smem, int *max_clusters) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_CLASS(t, (e = ClusterInst<real, DC, DV>::prepare(t, threads, smem, max_clusters)));
    return e;
}

template <typename real>
static inline cudaError_t cluster_launch(const ClusterTables &t, const BpArgs<real> &a, int nclusters, int threads, size_t smem,
                                         int flip_table, cudaStream_t st) {
    ClusterDev d;
    d.vslot = t.d_vslot; d.cdeg = t.d_cdeg; d.row_of = t.d_row_of; d.bit_of = t.d_bit_of;
    d.rows_per_cta = t.rows_per_cta; d.bits_per_cta = t.bits_per_cta; d.CL = t.CL;
    d.flip_table = flip_table;
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_CLASS(t, (e = ClusterInst<real, DC, DV>::launch(t, a, d, nclusters, threads, smem, st)));
    return e;
}

} // namespace bposd
