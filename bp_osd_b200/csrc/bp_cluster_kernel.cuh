// bp_cluster_kernel.cuh -- min-sum BP for parity-check matrices whose messages exceed one SM's shared
// memory (BASELINE config 5: 134 400 edges, 1.07 MB of fp64 messages): the in-place message array of
// bp_fast_kernel is split over the CTAs of a thread-block cluster, and the edges that cross CTAs are
// exchanged through distributed shared memory in COALESCED blocks ("halo exchange").
//
//   * physical rows (checks) are partitioned over the CL CTAs of a cluster, rows_per_cta each, and so are the
//     bits (bits_per_cta each); the host partition pass (cluster_build) places bits and checks to keep as many
//     edges CTA-local as it can (config 5: 71 %);
//   * an edge whose check lives in CTA A and whose bit lives in CTA B has TWO homes: its slot in the row of the
//     check (in A) and an entry of B's mailbox.  The mailbox of B is ordered by (source CTA, source order), so
//     the entries that A exchanges with B are contiguous in B;
//   * the check sweep (fast_check_row, the single-CTA kernel's) and the bit sweep (fast_bit_sweep, likewise) only
//     touch the CTA's own shared memory with plain LDS / STS; in between, A PUSHES the new check-to-bit values of
//     its remote edges into the mailboxes (consecutive threads write consecutive remote addresses) and, after the
//     bit sweep, PULLS the new bit-to-check values back into its rows (consecutive remote reads);
//   * one parity-mismatch bit per check, toggled with red.shared::cluster.xor when a hard decision flips;
//     every CTA votes with __syncthreads_and, the votes are exchanged through DSMEM and ride on the cluster
//     barrier that separates the sweeps (two barrier.cluster per iteration);
//   * persistent clusters pull shots from the same atomic queue as the other BP kernels.
//
// Why the exchange is blocked: ld/st.shared::cluster moves 32-byte sectors over the SM-to-SM crossbar (~20 B/clk
// per SM).  The first version of this kernel let every bit gather / scatter its slots directly with 8-byte
// ld/st.shared::cluster -- local slots included -- and ncu showed 28 sectors per warp request and 43 % of all
// stall samples on the first use of a gathered value (profiles/r2f_cluster_hot_lines.txt): 1.3 TB/s of sector
// traffic for 0.33 TB/s of payload.  Blocked, the crossbar carries each remote value once per direction in
// fully used sectors, and 71 % of the edges never leave the SM.
#pragma once
#include <cstdlib>
#include "bp_fast_kernel.cuh"

namespace bposd {

#define BPC_NONE 0xFFFFFFFFu

struct ClusterTables {
    int DC = 0, DV = 0, regular = 0;
    int CL = 0, rows_per_cta = 0, bits_per_cta = 0, elem_bytes = 0;
    int nbox_max = 0, nout_max = 0; // mailbox entries / exchange-list entries of the fullest CTA
    uint32_t *d_vslot = nullptr;  // [CL*bits_per_cta, DV] slot (element index inside the owner CTA: row slots, then the mailbox) of the k-th edge of the bit at position q
    uint32_t *d_flip = nullptr;   // [CL*bits_per_cta, DV] (CTA << 24 | local row) of that edge's check
    uint32_t *d_xloc = nullptr;   // [CL, nout_max] exchange list: element index of the row slot in this CTA ...
    uint32_t *d_xrem = nullptr;   // [CL, nout_max] ... and (CTA << 24 | element index) of its mailbox entry in the bit's CTA
    int *d_nout = nullptr;        // [CL] entries of each CTA's exchange list
    uint8_t *d_cdeg = nullptr;    // [CL*rows_per_cta] degree of the check in physical row p (0: absent)
    uint32_t *d_row_of = nullptr; // [CL*rows_per_cta] original check of physical row p, BPC_NONE: absent
    uint32_t *d_bit_of = nullptr; // [CL*bits_per_cta] original bit at position q, BPC_NONE: absent
    long long remote_edges = 0, total_edges = 0;
};

struct ClusterDev {
    const uint32_t *vslot, *flip, *xloc, *xrem;
    const int *nout;
    const uint8_t *cdeg;
    const uint32_t *row_of;
    const uint32_t *bit_of;
    int rows_per_cta, bits_per_cta, CL, nbox_max, nout_max;
    int flip_table; // parity-flip descriptors of the REMOTE edges in shared memory, one per mailbox entry: 1 32-bit cluster
                    // addresses, 2 16-bit (CTA << 12 | row) words (rows_per_cta <= 4096), 0 none (global table at flip time)
};

static inline void cluster_free(ClusterTables &t) {
    cudaFree(t.d_vslot); cudaFree(t.d_flip); cudaFree(t.d_xloc); cudaFree(t.d_xrem); cudaFree(t.d_nout);
    cudaFree(t.d_cdeg); cudaFree(t.d_row_of); cudaFree(t.d_bit_of);
    t.d_vslot = t.d_flip = t.d_xloc = t.d_xrem = nullptr; t.d_nout = nullptr;
    t.d_cdeg = nullptr; t.d_row_of = nullptr; t.d_bit_of = nullptr;
    t.CL = 0;
}

// Shared-memory layout of one CTA, the same arithmetic on host and device.
struct ClusterLayout { size_t o_meta, o_prior, o_xl, o_flip, total; };
__host__ __device__ inline ClusterLayout cluster_layout(int RS, int elem, int rpc, int bpc, int nbox_max, int nout_max, int DV, int flip_table,
                                                        int prior_table = 1) {
    ClusterLayout L;
    auto al = [](size_t x) { return (x + 15) / 16 * 16; };
    size_t o = al(((size_t)rpc * RS + nbox_max + kFastDummySlots) * elem); // row slots | mailbox | dummy / zero slots
    L.o_meta = o; o = al(o + (size_t)rpc);                                 // one byte per check
    L.o_prior = o; o = al(o + (prior_table ? (size_t)bpc * elem : 0));     // priors by position (non-uniform channels only)
    L.o_xl = o; o = al(o + (size_t)(nout_max > 0 ? nout_max : 1) * 8);     // exchange list (byte offset, cluster address)
    L.o_flip = o; o = al(o + (flip_table == 1 ? (size_t)nbox_max * 4 : flip_table == 2 ? (size_t)nbox_max * 2 : 0)); // flip descriptors of the mailbox entries
    L.total = o + 16;
    return L;
}
template <typename real>
static inline size_t cluster_smem_need(const ClusterTables &t, int flip_table, int prior_table = 1) {
    return cluster_layout(fast_row_stride(t.DC, (int)sizeof(real)), (int)sizeof(real), t.rows_per_cta, t.bits_per_cta, t.nbox_max, t.nout_max, t.DV, flip_table,
                          prior_table).total;
}
// lower bound before the partition is known (no mailbox, no exchange list): filters cluster sizes that cannot fit
template <typename real>
static inline size_t cluster_smem_min(int DC, int rows_per_cta, int bits_per_cta, int prior_table = 1) {
    return cluster_layout(fast_row_stride(DC, (int)sizeof(real)), (int)sizeof(real), rows_per_cta, bits_per_cta, 0, 0, 0, 0, prior_table).total;
}

// Degree class of the cluster kernel: the single-CTA classes, plus rows of exactly 7 slots when no check has more than 7
// edges (hypergraph products of (3,4)-regular codes, BASELINE config 5): 56-byte rows moved element by element instead
// of 80-byte padded rows moved in 16-byte pieces -- 30 % less shared memory per shot (a cluster of 8 CTAs holds config 5
// where the padded rows need 16: 18 shots in flight on 144 SMs instead of 7 on 112) and 14 instead of 20 wavefronts per
// warp and row access.
static inline void cluster_class(int max_col_deg, int max_row_deg, int *DC, int *DV) {
    fast_class(max_col_deg, max_row_deg, DC, DV);
    if (*DC == 8 && max_row_deg == 7) *DC = 7;
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
    cluster_class(mc, mr, &t.DC, &t.DV);
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
    // exchange lists: the remote edges of every CTA's rows, ordered by (CTA of the bit, row slot); the mailbox of a CTA
    // is filled source CTA by source CTA in that order, so what two CTAs exchange is contiguous on the mailbox side
    struct Rem { int B, loc, e; };
    std::vector<std::vector<Rem>> out(CL);
    for (int i = 0; i < m; i++)
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
            const int A = best_row[i], B = best_bit[col_idx[e]];
            if (A != B) out[A].push_back(Rem{B, (prow[i] % rpc) * RS + (e - row_ptr[i]), e});
        }
    std::vector<int> box_cnt(CL, 0), box_of_edge(std::max(E, 1), -1);
    int nout_max = 0;
    for (int A = 0; A < CL; A++) {
        std::sort(out[A].begin(), out[A].end(), [](const Rem &x, const Rem &y) { return x.B != y.B ? x.B < y.B : x.loc < y.loc; });
        for (const Rem &r : out[A]) box_of_edge[r.e] = box_cnt[r.B]++;
        nout_max = std::max(nout_max, (int)out[A].size());
    }
    int nbox_max = 0;
    for (int c = 0; c < CL; c++) nbox_max = std::max(nbox_max, box_cnt[c]);
    nbox_max = (nbox_max + 1) & ~1; // keeps the areas behind the mailbox 16-byte aligned in fp64
    t.nbox_max = nbox_max; t.nout_max = nout_max;
    if ((size_t)rpc * RS + nbox_max + kFastDummySlots >= (1u << 24)) return cudaErrorInvalidValue;
    std::vector<uint32_t> vs((size_t)CL * bpc * t.DV, BPC_NONE), fl((size_t)CL * bpc * t.DV, BPC_NONE), rowof((size_t)CL * rpc, BPC_NONE),
        bitof((size_t)CL * bpc, BPC_NONE), xloc((size_t)CL * std::max(nout_max, 1), BPC_NONE), xrem((size_t)CL * std::max(nout_max, 1), BPC_NONE);
    std::vector<int> nout(CL, 0);
    std::vector<uint8_t> cd((size_t)CL * rpc, 0);
    for (int i = 0; i < m; i++) { rowof[prow[i]] = (uint32_t)i; cd[prow[i]] = (uint8_t)(row_ptr[i + 1] - row_ptr[i]); }
    for (int j = 0; j < n; j++) {
        bitof[pos[j]] = (uint32_t)j;
        for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) {
            const int i = row_idx[q], e = csc_slot[q];
            const size_t at = (size_t)pos[j] * t.DV + (q - col_ptr[j]);
            vs[at] = (best_row[i] == best_bit[j]) ? (uint32_t)((prow[i] % rpc) * RS + (e - row_ptr[i])) : (uint32_t)(rpc * RS + box_of_edge[e]);
            fl[at] = ((uint32_t)best_row[i] << 24) | (uint32_t)(prow[i] % rpc);
        }
    }
    for (int A = 0; A < CL; A++) {
        nout[A] = (int)out[A].size();
        for (size_t x = 0; x < out[A].size(); x++) {
            xloc[(size_t)A * nout_max + x] = (uint32_t)out[A][x].loc;
            xrem[(size_t)A * nout_max + x] = ((uint32_t)out[A][x].B << 24) | (uint32_t)(rpc * RS + box_of_edge[out[A][x].e]);
        }
    }
    auto up = [](auto **dst, const auto &src) -> cudaError_t {
        using T = typename std::remove_reference<decltype(src)>::type::value_type;
        cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(src.size(), 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        return src.empty() ? cudaSuccess : cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    };
    cudaError_t e;
    if ((e = up(&t.d_vslot, vs)) != cudaSuccess) return e;
    if ((e = up(&t.d_flip, fl)) != cudaSuccess) return e;
    if ((e = up(&t.d_xloc, xloc)) != cudaSuccess) return e;
    if ((e = up(&t.d_xrem, xrem)) != cudaSuccess) return e;
    if ((e = up(&t.d_nout, nout)) != cudaSuccess) return e;
    if ((e = up(&t.d_cdeg, cd)) != cudaSuccess) return e;
    if ((e = up(&t.d_row_of, rowof)) != cudaSuccess) return e;
    return up(&t.d_bit_of, bitof);
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

// Register-frugal form of the check update (row a4) for CTAs of more than 640 threads (64 registers each):
// pass 1 streams the row once for (smallest, second smallest, position of the smallest, sign parity), pass 2
// streams it again and overwrites every slot with "minimum over the other edges" = second smallest at the
// position of the smallest, smallest elsewhere -- the same values as the prefix/suffix form of
// fast_check_row (min is exact; on a tie both forms give the tied value).  "<= 0" counts zero as negative.
template <typename real, int DC, bool REG>
__device__ __forceinline__ void cluster_check_row(real *row, unsigned mt, real alpha) {
    constexpr int CH = (DC % 2) ? 1 : (sizeof(real) == 8) ? 2 : ((DC % 4 == 0) ? 4 : 2); // elements per vector access
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

// MAXT (a multiple of 128 >= blockDim.x) sets the register budget: 65536 / MAXT per thread.  Up to 640 threads (96+
// registers) the check update is the single-CTA kernel's prefix/suffix form (fewer instructions, a row in registers);
// bigger CTAs fall back to the two-pass form above.
#ifndef BPOSD_CLUSTER_FASTROW_MAXT
#define BPOSD_CLUSTER_FASTROW_MAXT 640
#endif
template <typename real, int DC, int DV, int VPT, int MAXT, bool REG>
__global__ void __launch_bounds__(MAXT, 1) bp_cluster_kernel(BpArgs<real> a, ClusterDev t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.g.n;
    const int tid = threadIdx.x;
    int T;
    asm volatile("mov.u32 %0, %%ntid.x;" : "=r"(T));
    const int rpc = t.rows_per_cta, bpc = t.bits_per_cta, CL = t.CL;
    const uint32_t rank = cluster_ctarank();
    constexpr int RS = fast_row_stride(DC, (int)sizeof(real)); // row stride in elements
    const ClusterLayout L = cluster_layout(RS, (int)sizeof(real), rpc, bpc, t.nbox_max, t.nout_max, DV, t.flip_table, a.uniform_prior ? 0 : 1);
    const unsigned nslot = (unsigned)rpc * RS + (unsigned)t.nbox_max;      // row slots, then the mailbox; dummy / zero slots follow
    real *msg = reinterpret_cast<real *>(smem_raw);
    uint8_t *meta = smem_raw + L.o_meta;                                   // bit0 mismatch, bits1-5 degree, bit7 syndrome
    real *prior_s = reinterpret_cast<real *>(smem_raw + L.o_prior);        // [bpc] priors by position (non-uniform only)
    uint2 *xl = reinterpret_cast<uint2 *>(smem_raw + L.o_xl);              // [nout] (byte offset of the row slot, cluster address of the mailbox entry)
    uint32_t *flip_desc = reinterpret_cast<uint32_t *>(smem_raw + L.o_flip); // [nbox_max] flip descriptors of the mailbox entries
    __shared__ long long sh_shot;
    __shared__ int sh_slot;
    __shared__ unsigned sh_vote[2][16];
    __shared__ unsigned sh_big[2][16];
    const uint32_t msg_s = smem_u32(msg), meta_s = smem_u32(meta);
    const int nout = t.nout[rank];
    for (int x = tid; x < nout; x += T) {
        const uint32_t loc = t.xloc[(size_t)rank * t.nout_max + x], rem = t.xrem[(size_t)rank * t.nout_max + x];
        xl[x] = make_uint2(loc * (unsigned)sizeof(real), mapa_u32(msg_s + (rem & 0xFFFFFFu) * (unsigned)sizeof(real), rem >> 24));
    }
    // A hard-decision flip toggles the parity-mismatch bit of every neighbouring check, wherever it lives.  A check of this
    // CTA (three edges in four) is found by arithmetic on the slot offset and toggled with a local shared-memory atomic.
    // A remote check needs its cluster address: one descriptor per MAILBOX ENTRY, resolved once per CTA (32-bit cluster
    // address | byte lane, or 16 bits (CTA << 12 | row) when shared memory is short); without a table it costs a global
    // load per edge at flip time that always misses L1 (barrier.cluster invalidates it), with every other thread of the
    // CTA waiting at the next barrier -- and some bit flips in most passes.
    auto flip_descriptor = [&](uint32_t f) -> uint32_t {
        if (f == BPC_NONE) return 0u;
        const uint32_t lp = f & 0xFFFFFFu;
        return mapa_u32(meta_s + (lp & ~3u), f >> 24) | (lp & 3u);
    };
    uint16_t *flip_desc16 = reinterpret_cast<uint16_t *>(flip_desc);
    if (t.flip_table)
        for (int e = tid; e < bpc * DV; e += T) {
            const uint32_t sl = t.vslot[(size_t)rank * bpc * DV + e];
            if (sl == BPC_NONE || sl < (uint32_t)rpc * RS) continue;
            const uint32_t f = t.flip[(size_t)rank * bpc * DV + e];
            if (t.flip_table == 1) flip_desc[sl - (uint32_t)rpc * RS] = flip_descriptor(f);
            else flip_desc16[sl - (uint32_t)rpc * RS] = (uint16_t)(((f >> 24) << 12) | (f & 0xFFFu));
        }

    // byte offsets (inside this CTA's shared memory) of the slots of this thread's bits: a row slot for an edge whose
    // check lives here, a mailbox entry otherwise (shot independent)
    unsigned off[VPT][DV];
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
            off[r][k] = s * (unsigned)sizeof(real);
            dj[r] += (s != BPC_NONE) ? 1 : 0;
            if (REG && !have) off[r][k] = (nslot + (unsigned)k) * (unsigned)sizeof(real);               // dummy slots
            if (!REG && s == BPC_NONE) off[r][k] = (nslot + (unsigned)kFastZeroSlot) * (unsigned)sizeof(real); // the constant-zero slot
        }
    }
    const unsigned zoff = (nslot + (unsigned)kFastZeroSlot) * (unsigned)sizeof(real); // offset of an absent edge (irregular codes)
    unsigned long long n_conv = 0, n_iter = 0;
    const bool uniform = a.uniform_prior != 0;
    const real prior_u = a.prior[0];
    __syncthreads();

    // remote halves of the exchange: rows -> mailboxes (push, after a check sweep), mailboxes -> rows (pull, after a bit sweep)
    auto push_rows = [&]() {
        for (int x = tid; x < nout; x += T) {
            const uint2 d = xl[x];
            st_dsmem(d.y, *reinterpret_cast<const real *>(smem_raw + d.x));
        }
    };
    auto pull_rows = [&]() {
        constexpr int U = 4; // independent remote loads in flight per thread
        for (int x0 = tid; x0 < nout; x0 += U * T) {
            real v[U];
            unsigned o[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int x = x0 + u * T;
                const uint2 d = x < nout ? xl[x] : make_uint2(0xFFFFFFFFu, 0u);
                o[u] = d.x;
                v[u] = (x < nout) ? ld_dsmem(d.y, (real)0) : (real)0;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (o[u] != 0xFFFFFFFFu) *reinterpret_cast<real *>(smem_raw + o[u]) = v[u];
        }
    };

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
        if (!REG && tid == 0) msg[nslot + kFastZeroSlot] = (real)0;
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
                    if (REG || off[r][k] != zoff) *reinterpret_cast<real *>(smem_raw + off[r][k]) = pj;
            }
        }
        cluster_sync_all(); // every mailbox holds the priors of its bits
        pull_rows();
        __syncthreads();

        bool conv = false;
        int iters = 0;
        int danger = 0; // overflow guard (llr_near_overflow): some LLR of the cluster came near overflow in the last bit sweep
        real pow2 = 1;
        for (int it = 1;; it++) {
            const bool last = it > a.max_iter;
            pow2 *= (real)0.5;
            const real alpha = (a.alpha0 == (real)0) ? (real)1 - pow2 : a.alpha0;
            const uint32_t alpha_w = sign_word(alpha);
            bool ok = true;
            if (danger && !last) { // rare: +-inf in the rows (local and pulled values alike) become +-max
                for (int p = tid; p < rpc; p += T)
                    for (int k = 0; k < DC; k++) msg[(size_t)p * RS + k] = clamp_inf<real>(msg[(size_t)p * RS + k]);
            }
            // ---- check sweep over this CTA's rows (a4) + local convergence vote for the previous pass (a7)
            for (int p = tid; p < rpc; p += T) {
                const unsigned mt = meta[p];
                if (mt & 1u) ok = false;
                if (last) continue;
                if constexpr (MAXT <= BPOSD_CLUSTER_FASTROW_MAXT) fast_check_row<real, DC, REG>(msg + (size_t)p * RS, mt, alpha, alpha_w);
                else cluster_check_row<real, DC, REG>(msg + (size_t)p * RS, mt, alpha);
            }
            const int cta_ok = __syncthreads_and(ok ? 1 : 0);
            if (!last) push_rows(); // new check-to-bit values of the remote edges into the mailboxes of their bits
            if (tid < CL) st_dsmem_u32(mapa_u32(smem_u32(&sh_vote[it & 1][rank]), tid), (uint32_t)cta_ok);
            cluster_sync_all();
            int all_ok = 1;
            for (int c = 0; c < CL; c++) all_ok &= (int)sh_vote[it & 1][c];
            if (it > 1 && all_ok) { conv = true; iters = it - 1; break; }
            if (last) { iters = a.max_iter; break; }
            // ---- bit sweep (a6 + a8): rows and mailbox of this CTA only
            const unsigned dnow = valid & (uniform ? fast_bit_sweep<real, DV, VPT, REG, true, true>(smem_raw, off, dj, llr, prior_u, prior_s, tid, T, bpc, zoff)
                                                   : fast_bit_sweep<real, DV, VPT, REG, false, true>(smem_raw, off, dj, llr, prior_u, prior_s, tid, T, bpc, zoff));
            if (dnow != dprev) {
                // a hard decision flipped (rare): toggle the parity-mismatch bit of every neighbouring check
                unsigned flip = dnow ^ dprev;
                dprev = dnow;
#pragma unroll
                for (int r = 0; r < VPT; r++)
                    if ((flip >> r) & 1u) {
                        const int lq = tid + r * T;
                        // the slots come from the global table here (a flip is rare): indexing `off` with a run-time k would
                        // move the whole array to local memory, and its LDLs into every bit position of every sweep
#pragma unroll 1
                        for (int k = 0; k < DV; k++) {
                            const uint32_t sl = t.vslot[((size_t)rank * bpc + lq) * DV + k];
                            if (sl == BPC_NONE) continue; // absent edge
                            if (sl < (uint32_t)rpc * RS) { // the check lives here
                                const unsigned p = sl / (unsigned)RS;
                                atomicXor(reinterpret_cast<unsigned *>(meta + (p & ~3u)), 1u << ((p & 3u) * 8u));
                                continue;
                            }
                            const unsigned b = sl - (unsigned)rpc * RS;
                            uint32_t d;
                            if (t.flip_table == 1) d = flip_desc[b];
                            else if (t.flip_table == 2) { const uint32_t w = flip_desc16[b]; d = mapa_u32(meta_s + (w & 0xFFCu), w >> 12) | (w & 3u); }
                            else d = flip_descriptor(t.flip[((size_t)rank * bpc + lq) * DV + k]);
                            xor_dsmem_u32(d & ~3u, 1u << ((d & 3u) * 8u));
                        }
                    }
            }
            const bool guard = it >= a.safe_it; // the same in every CTA of the cluster
            if (guard) {
                bool big = false;
#pragma unroll
                for (int r = 0; r < VPT; r++) big |= llr_near_overflow<real>(llr[r]) && ((valid >> r) & 1u);
                const int cta_big = __syncthreads_or(big ? 1 : 0);
                if (tid < CL) st_dsmem_u32(mapa_u32(smem_u32(&sh_big[it & 1][rank]), tid), (uint32_t)cta_big);
            }
            cluster_sync_all();
            danger = 0;
            if (guard)
                for (int c = 0; c < CL; c++) danger |= (int)sh_big[it & 1][c];
            pull_rows(); // new bit-to-check values of the remote edges back into the rows
            __syncthreads();
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
    if (const char *f = std::getenv("BPOSD_CLUSTER_VPT")) { // tuning knob (A/B of the CTA size): 6 or 8 bits per thread
        const int v = std::atoi(f);
        if ((v == 6 || v == 8) && (bits_per_cta + v - 1) / v <= 1024) return v;
    }
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
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kern);
        // only ever raised: two plans (with / without a prior array) may share an instantiation
        if (e == cudaSuccess && (int)smem > fa.maxDynamicSharedSizeBytes) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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

#ifdef BPOSD_DEV_CLASS_ONLY
#define BPOSD_CLUSTER_CLASS(t, EXPR) do { constexpr int DC = BPOSD_DEV_DC, DV = (BPOSD_DEV_DC + 1) / 2; EXPR; } while (0)
#else
#define BPOSD_CLUSTER_CLASS(t, EXPR)                                                             \
    do {                                                                                         \
        if (t.DC == 7) { constexpr int DC = 7, DV = 4; EXPR; }                                   \
        else BPOSD_FAST_CLASS(t, EXPR);                                                          \
    } while (0)
#endif

template <typename real>
static inline cudaError_t cluster_prepare(const ClusterTables &t, int threads, size_t smem, int *max_clusters) {
    cudaError_t e = cudaSuccess;
    BPOSD_CLUSTER_CLASS(t, (e = ClusterInst<real, DC, DV>::prepare(t, threads, smem, max_clusters)));
    return e;
}

template <typename real>
static inline cudaError_t cluster_launch(const ClusterTables &t, const BpArgs<real> &a, int nclusters, int threads, size_t smem,
                                         int flip_table, cudaStream_t st) {
    ClusterDev d;
    d.vslot = t.d_vslot; d.flip = t.d_flip; d.xloc = t.d_xloc; d.xrem = t.d_xrem; d.nout = t.d_nout;
    d.cdeg = t.d_cdeg; d.row_of = t.d_row_of; d.bit_of = t.d_bit_of;
    d.rows_per_cta = t.rows_per_cta; d.bits_per_cta = t.bits_per_cta; d.CL = t.CL;
    d.nbox_max = t.nbox_max; d.nout_max = t.nout_max;
    d.flip_table = flip_table;
    cudaError_t e = cudaSuccess;
    BPOSD_CLUSTER_CLASS(t, (e = ClusterInst<real, DC, DV>::launch(t, a, d, nclusters, threads, smem, st)));
    return e;
}

} // namespace bposd
