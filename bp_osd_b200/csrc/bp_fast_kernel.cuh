// bp_fast_kernel.cuh -- in-place shared-memory min-sum BP for sm_100a (rows a3-a8, bit-exact in fp64).
//
// One message array in shared memory, check-major with a fixed row stride RS >= DC: slot = prow*RS + kappa,
// where prow is the *physical* row the host layout pass assigned to a check and kappa its edge's
// physical position in that row (min is order independent, so both are free; the host pass picks
// them to make the bit-side gathers/scatters bank-conflict free, see fast_build).  A pass is two
// sweeps over the array:
//   check sweep: the thread that owns a physical row pulls it with 16-byte LDS, forms for every
//                edge the minimum over the *other* edges with prefix/suffix running minima (the
//                reference's own formulation: 3d-4 compare-selects, no argmin bookkeeping), applies
//                sign and scaling with integer sign-bit arithmetic (exact slow path when a message
//                is +-0, where "<= 0" and the sign bit disagree), and overwrites the row in place;
//   bit sweep:   the thread that owns a bit gathers its <= DV messages, forms prefix/suffix sums in
//                the reference's order (prior + c_1 + ... left to right; suffix from the last edge
//                backwards) and overwrites them with the new bit->check messages.
// Convergence (H*decoding == syndrome) is tracked incrementally: every check keeps one parity-mismatch
// bit (in a 32-bit flag word of its own), initialised to its syndrome bit; a bit whose hard decision flips
// toggles the mismatch bits of its checks with a shared-memory atomicXor (frequent in the first passes,
// rare later), and the next check sweep only votes on that bit.
// Per-bit state that must survive to the end of the shot (LLR of the last executed pass, edge
// offsets, last hard decisions) lives in registers; nothing but the syndrome and the results
// touches HBM.  Persistent CTAs pull shots from an atomic queue (iteration counts are heavy tailed).
#pragma once
#include "bposd_kernels.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

namespace bposd {

struct FastTables {
    int DC = 0, DV = 0;            // degree class (upper bounds, compile-time in the kernel)
    int regular = 0;               // every row has exactly DC entries and every column exactly DV
    int ps = 0;                    // product-sum check update (rows keep their edges in ascending-column order: products are order dependent)
    int elem_bytes = 8;
    uint16_t *d_vslot = nullptr;   // [n, DV] message slot of the k-th edge (ascending row) of the bit at position q, 0xFFFF = none
    uint8_t *d_cdeg = nullptr;     // [m] degree of the check stored in physical row p
    uint16_t *d_row_of = nullptr;  // [m] original check index of physical row p
    uint16_t *d_bit_of = nullptr;  // [n] bit handled at thread position q (the layout pass permutes bits too)
    long long conflicts_before = 0, conflicts_after = 0, wavefronts_ideal = 0;
};

static inline bool fast_supported(int max_col_deg, int max_row_deg, int method) {
    if (method == 0) // product-sum: instantiated for the degree classes up to (8, 4) (a row's tanh / log chain lives in registers)
        return max_col_deg >= 1 && max_col_deg <= 4 && max_row_deg >= 1 && max_row_deg <= 8;
    return method == 1 && max_col_deg >= 1 && max_col_deg <= 8 && max_row_deg >= 1 && max_row_deg <= 16;
}

static inline void fast_class(int max_col_deg, int max_row_deg, int *DC, int *DV) {
    if (max_row_deg <= 4 && max_col_deg <= 2) { *DC = 4; *DV = 2; }
    else if (max_row_deg <= 6 && max_col_deg <= 3) { *DC = 6; *DV = 3; }
    else if (max_row_deg <= 8 && max_col_deg <= 4) { *DC = 8; *DV = 4; }
    else { *DC = 16; *DV = 8; }
}

// Row stride of the message array in elements.  A check's row is moved with 16-byte vector accesses by
// the thread that owns it, so 8 consecutive rows must start in 8 different 16-byte bank groups: the
// stride is padded to an odd number of 16-byte chunks (64-byte rows would be 4-way conflicted).
// fp32 rows of 6 use 8-byte accesses and a 24-byte stride, which is conflict free as it is.
__host__ __device__ constexpr int fast_row_stride(int DC, int elem_bytes) {
    return (DC * elem_bytes) % 16 != 0 ? DC
           : (((DC * elem_bytes) / 16) % 2 == 1 ? DC : DC + 16 / elem_bytes);
}

static inline void fast_free(FastTables &t) {
    cudaFree(t.d_vslot); cudaFree(t.d_cdeg); cudaFree(t.d_row_of); cudaFree(t.d_bit_of);
    t.d_vslot = nullptr; t.d_cdeg = nullptr; t.d_row_of = nullptr; t.d_bit_of = nullptr;
}

// bits per thread and the CTA size cap: (2, 1024) | (4, 512) | (8, 512) | (8, 1024)
// Geometry: bits per thread (VPT) and CTA size cap (MAXT) by code size.  Measured on B200 (profiles/):
// fp64 wants no spills (register cap 128, few big-ILP warps), fp32 wants occupancy (cap 64).
//   n <= 256: (2, 128)   n <= 2048: (8, 256)   n <= 4096: (8, 512)   n <= 8192: (8, 1024)
#ifndef BPOSD_CHECK_UNROLL
#define BPOSD_CHECK_UNROLL 1
#endif
constexpr int kCheckUnroll = BPOSD_CHECK_UNROLL; // rows of the check sweep interleaved per thread
#ifndef BPOSD_REGCAP64
#define BPOSD_REGCAP64 128
#endif
#ifndef BPOSD_REGCAP32
#define BPOSD_REGCAP32 64
#endif
// the 256 < n <= 2048 class (the bench code) can be re-tuned at build time: bits per thread, CTA size cap
#ifndef BPOSD_MID_VPT
#define BPOSD_MID_VPT 8
#endif
#ifndef BPOSD_MID_MAXT
#define BPOSD_MID_MAXT 256
#endif
// latency geometry of the mid class: the whole CTA works on one shot, so what matters when only a few shots are in
// flight (single-shot decode()) is the length of a pass, not shots per SM: few bits per thread, a full-size CTA
// Measured on B200 (profiles/r03a_ab_probe.log, single-shot decode() p50 on the bench code): fp64 (4, 512) 95.6 us,
// (3, 704) 102 us, (2, 1024) 104 us -- 1024 threads cap the kernel at 64 registers and the check update spills;
// fp32 (2, 1024) 80 us, (4, 512) 90 us.
#ifndef BPOSD_LAT_VPT64
#define BPOSD_LAT_VPT64 4
#endif
#ifndef BPOSD_LAT_MAXT64
#define BPOSD_LAT_MAXT64 512
#endif
#ifndef BPOSD_LAT_VPT32
#define BPOSD_LAT_VPT32 2
#endif
#ifndef BPOSD_LAT_MAXT32
#define BPOSD_LAT_MAXT32 1024
#endif
// geometry ids: 0 (2, 128) | 1 mid (throughput) | 2 (8, 512) | 3 (8, 1024) | 4 mid (latency)
static inline int fast_geom(int n, bool latency) { return n <= 256 ? 0 : (n <= 2048 ? (latency ? 4 : 1) : (n <= 4096 ? 2 : 3)); }
static inline int fast_vpt_g(int geom, int elem_bytes = 8) {
    return geom == 0 ? 2 : (geom == 1 ? BPOSD_MID_VPT : (geom == 4 ? (elem_bytes == 8 ? BPOSD_LAT_VPT64 : BPOSD_LAT_VPT32) : 8));
}
static inline int fast_maxt_g(int geom, int elem_bytes = 8) {
    return geom == 0 ? 128 : (geom == 1 ? BPOSD_MID_MAXT : (geom == 2 ? 512 : (geom == 4 ? (elem_bytes == 8 ? BPOSD_LAT_MAXT64 : BPOSD_LAT_MAXT32) : 1024)));
}
static inline int fast_vpt(int n) { return fast_vpt_g(fast_geom(n, false)); }
static inline int fast_maxt(int n) { return fast_maxt_g(fast_geom(n, false)); }
// fp64, throughput geometry of the mid class, rows of <= 6 / bits of <= 3 edges (the bench code): 80 registers, i.e.
// 3 CTAs of 256 threads per SM instead of 2 -- possible once the parity-flip arithmetic stays inside its branch
// (kFlipInBranch).  Measured on B200 (profiles/r03l_ab_probe.log): 136.3 vs 129.7 M shot-iterations/s.
#ifndef BPOSD_REGCAP64_SMALL
#define BPOSD_REGCAP64_SMALL 80
#endif
// product-sum in fp64: the kernel is bound by instruction issue (tanh / log / three IEEE divisions per edge), so resident
// warps matter more than spills of the six interleaved chains
#ifndef BPOSD_REGCAP64_PS
#define BPOSD_REGCAP64_PS 64 // measured on B200 (profiles/r2n_ps_ab.log, cfg 4): 128 -> 36.6, 80 -> 46.1, 64 -> 52.6 M shot-iterations/s
#endif
template <typename real, int MAXT, int DC, int DV, int VPT, bool REG, bool PS = false>
constexpr bool kFastSmallClass = sizeof(real) == 8 && MAXT == BPOSD_MID_MAXT && VPT == BPOSD_MID_VPT && MAXT == 256 && DC <= 6 && DV <= 3 && REG && !PS &&
                                 BPOSD_REGCAP64_SMALL < BPOSD_REGCAP64; // regular codes only: the irregular form spills twice as much (unmeasured)
template <typename real, int MAXT, int DC = 16, int DV = 8, int VPT = 0, bool REG = false, bool PS = false> constexpr int fast_minb() {
    constexpr int cap = sizeof(real) == 8 ? (PS ? BPOSD_REGCAP64_PS : (kFastSmallClass<real, MAXT, DC, DV, VPT, REG, PS> ? BPOSD_REGCAP64_SMALL : BPOSD_REGCAP64))
                                          : BPOSD_REGCAP32;
    return (65536 / (MAXT * cap)) < 1 ? 1 : (65536 / (MAXT * cap));
}

static inline int fast_default_threads_g(int n, int geom, int elem_bytes) {
    const int vpt = fast_vpt_g(geom, elem_bytes);
    int t = ((n + vpt - 1) / vpt + 31) / 32 * 32;
    return std::min(fast_maxt_g(geom, elem_bytes), std::max(32, t));
}
static inline int fast_default_threads(int n, int m) {
    (void)m;
    return fast_default_threads_g(n, fast_geom(n, false), 8);
}

// ---------------------------------------------------------------------------------------------
// Host layout pass ("the Tanner graph compiled once per parity-check matrix").
// The bit sweep's k-th access of a warp touches slot[j][k] for its 32 bits j; shared memory serves
// one 128-byte wavefront per cycle, so 8-byte accesses go out as two half-warps of 16 lanes over
// 16 eight-byte banks and 4-byte accesses as 32 lanes over 32 banks.  Local search over the row
// permutation and the in-row edge positions minimises the total number of wavefronts.
// ---------------------------------------------------------------------------------------------
struct LayoutOpt {
    int m, n, DC, DV, RS, nbanks, group;
    std::vector<int> prow;             // check -> physical row
    std::vector<int> kappa;            // CSR edge -> position inside its physical row
    std::vector<int> pos;              // bit -> position (thread tid = pos % T handles it in round pos / T)
    std::vector<int> edge_k;           // CSR edge -> index of the edge inside its bit (ascending row)
    std::vector<int> edge_row;         // CSR edge -> check
    std::vector<int> edge_col;         // CSR edge -> bit
    std::vector<int> hist;             // [group][bank] number of lanes of the group that hit the bank
    int group_of(int e) const { return (pos[edge_col[e]] / group) * DV + edge_k[e]; }
    int bank_of(int e) const { return (prow[edge_row[e]] * RS + kappa[e]) % nbanks; }
    int &cell(int e) { return hist[(size_t)group_of(e) * nbanks + bank_of(e)]; }
    // pair cost: sum over groups and banks of C(count, 2); moving one edge changes it by count differences
    long long remove(int e) { int &c = cell(e); c--; return -(long long)c; }
    long long add(int e) { int &c = cell(e); long long d = c; c++; return d; }
    long long wavefronts() const { // what the hardware pays: max bank multiplicity per group
        long long w = 0;
        for (size_t g = 0; g * nbanks < hist.size(); g++) {
            int mx = 0;
            for (int b = 0; b < nbanks; b++) mx = std::max(mx, hist[g * nbanks + b]);
            w += mx;
        }
        return w;
    }
};

struct LayoutResult {
    std::vector<int> prow, kappa, pos;
    long long before = 0, after = 0, ideal = 0;
};

#ifndef BPOSD_LAYOUT_MS
#define BPOSD_LAYOUT_MS 1500 // wall-clock budget of the layout search per (matrix, precision)
#endif

// Local search over three kinds of moves, always anchored at an edge that currently conflicts:
// swap the physical rows of two checks, swap two positions inside a check's row, swap the thread
// positions of two bits.  Downhill and sideways moves are taken (the plateaus are wide).
static inline LayoutResult fast_layout_search(int m, int n, int DC, int DV, int elem_bytes, const std::vector<int> &row_ptr,
                                              const std::vector<int> &col_idx, const std::vector<int> &col_ptr,
                                              const std::vector<int> &csc_slot, bool keep_row_order = false) {
    const int E = row_ptr[m];
    LayoutOpt L;
    L.m = m; L.n = n; L.DC = DC; L.DV = DV; L.RS = fast_row_stride(DC, elem_bytes);
    L.group = elem_bytes == 8 ? 16 : 32; // lanes served together by one wavefront
    L.nbanks = L.group;                  // banks in units of the element size
    L.prow.resize(m); L.kappa.resize(E); L.edge_row.resize(E); L.edge_col.resize(E); L.edge_k.assign(E, 0); L.pos.resize(n);
    for (int i = 0; i < m; i++) {
        L.prow[i] = i;
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) { L.kappa[e] = e - row_ptr[i]; L.edge_row[e] = i; L.edge_col[e] = col_idx[e]; }
    }
    for (int j = 0; j < n; j++) {
        L.pos[j] = j;
        for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) L.edge_k[csc_slot[q]] = q - col_ptr[j];
    }
    const int ngroups = ((n + L.group - 1) / L.group) * DV;
    L.hist.assign((size_t)ngroups * L.nbanks, 0);
    long long cur = 0, ideal = 0;
    for (int e = 0; e < E; e++) cur += L.add(e);
    {
        std::vector<char> used(ngroups, 0);
        for (int e = 0; e < E; e++) used[L.group_of(e)] = 1;
        for (char u : used) ideal += u;
    }
    LayoutResult res;
    res.ideal = ideal;
    res.before = L.wavefronts() - ideal;
    std::mt19937 rng(12345u);
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(BPOSD_LAYOUT_MS);
    const long long budget = 40000ll * std::max(E, 1);
    auto rem_bit = [&](int j) { long long d = 0; for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) d += L.remove(csc_slot[q]); return d; };
    auto add_bit = [&](int j) { long long d = 0; for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) d += L.add(csc_slot[q]); return d; };
    auto rem_row = [&](int i) { long long d = 0; for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) d += L.remove(e); return d; };
    auto add_row = [&](int i) { long long d = 0; for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) d += L.add(e); return d; };
    for (long long iter = 0; iter < budget && cur > 0 && E > 0; iter++) {
        if ((iter & 0xFFFF) == 0xFFFF && std::chrono::steady_clock::now() > t_end) break;
        const int e0 = (int)(rng() % (unsigned)E);
        if (L.cell(e0) <= 1) continue; // anchor moves at a conflicting edge
        unsigned mv = rng() % 3u;
        if (keep_row_order && mv == 1) mv = (rng() & 1u) ? 0u : 2u; // product-sum: the edges of a row stay in ascending-column order
        if (mv == 0) {
            const int i1 = L.edge_row[e0], i2 = (int)(rng() % (unsigned)m);
            if (i1 == i2) continue;
            long long d = rem_row(i1) + rem_row(i2);
            std::swap(L.prow[i1], L.prow[i2]);
            d += add_row(i1) + add_row(i2);
            if (d <= 0) { cur += d; continue; }
            rem_row(i1); rem_row(i2);
            std::swap(L.prow[i1], L.prow[i2]);
            add_row(i1); add_row(i2);
        } else if (mv == 1) {
            // exchange two occupied positions of this row (pad positions of short rows stay at the end:
            // the kernel keeps +max in positions >= degree)
            const int i = L.edge_row[e0], dg = row_ptr[i + 1] - row_ptr[i];
            if (dg < 2) continue;
            const int e1 = e0, e2 = row_ptr[i] + (int)(rng() % (unsigned)dg);
            if (e1 == e2) continue;
            long long d = L.remove(e1) + L.remove(e2);
            std::swap(L.kappa[e1], L.kappa[e2]);
            d += L.add(e1) + L.add(e2);
            if (d <= 0) { cur += d; continue; }
            L.remove(e1); L.remove(e2);
            std::swap(L.kappa[e1], L.kappa[e2]);
            L.add(e1); L.add(e2);
        } else {
            const int j1 = L.edge_col[e0], j2 = (int)(rng() % (unsigned)n);
            if (j1 == j2 || L.pos[j1] / L.group == L.pos[j2] / L.group) continue;
            long long d = rem_bit(j1) + rem_bit(j2);
            std::swap(L.pos[j1], L.pos[j2]);
            d += add_bit(j1) + add_bit(j2);
            if (d <= 0) { cur += d; continue; }
            rem_bit(j1); rem_bit(j2);
            std::swap(L.pos[j1], L.pos[j2]);
            add_bit(j1); add_bit(j2);
        }
    }
    res.after = L.wavefronts() - ideal;
    res.prow = L.prow; res.kappa = L.kappa; res.pos = L.pos;
    return res;
}

static inline cudaError_t fast_build(FastTables &t, int m, int n, const std::vector<int> &row_ptr,
                                     const std::vector<int> &col_idx, const std::vector<int> &col_ptr,
                                     const std::vector<int> &row_idx, const std::vector<int> &csc_slot,
                                     int elem_bytes, int method = 1) {
    int mr = 0, mc = 0, minr = 1 << 30, minc = 1 << 30;
    for (int i = 0; i < m; i++) { int d = row_ptr[i + 1] - row_ptr[i]; mr = std::max(mr, d); minr = std::min(minr, d); }
    for (int j = 0; j < n; j++) { int d = col_ptr[j + 1] - col_ptr[j]; mc = std::max(mc, d); minc = std::min(minc, d); }
    fast_class(mc, mr, &t.DC, &t.DV);
    t.elem_bytes = elem_bytes;
    t.ps = method == 0 ? 1 : 0;
    t.regular = (m > 0 && minr == t.DC && mr == t.DC && minc == t.DV && mc == t.DV) ? 1 : 0;
    const int RS = fast_row_stride(t.DC, elem_bytes);
    if ((long long)m * RS >= 0xFFFF || m == 0 || n >= 0xFFFF) { t.DC = 0; return cudaSuccess; } // slots and bits must fit in 16 bits
    // the search is deterministic; results are cached per (matrix, precision) for the life of the process
    static std::mutex mu;
    static std::map<std::string, LayoutResult> cache;
    std::string key;
    {
        const int head[6] = {m, n, t.DC, t.DV, elem_bytes, t.ps};
        key.append(reinterpret_cast<const char *>(head), sizeof(head));
        key.append(reinterpret_cast<const char *>(row_ptr.data()), row_ptr.size() * sizeof(int));
        key.append(reinterpret_cast<const char *>(col_idx.data()), col_idx.size() * sizeof(int));
    }
    LayoutResult L;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it == cache.end()) {
            if (cache.size() > 64) cache.clear();
            it = cache.emplace(key, fast_layout_search(m, n, t.DC, t.DV, elem_bytes, row_ptr, col_idx, col_ptr, csc_slot, t.ps != 0)).first;
        }
        L = it->second;
    }
    t.wavefronts_ideal = L.ideal;
    t.conflicts_before = L.before;
    t.conflicts_after = L.after;
    std::vector<int> row_at(m);
    for (int i = 0; i < m; i++) row_at[L.prow[i]] = i;

    // tables are indexed by thread position, not by bit: position q is served by thread q % T in round q / T
    std::vector<uint16_t> vs((size_t)n * t.DV, 0xFFFF), rowof(m), bitof(n);
    std::vector<uint8_t> cd(m, 0);
    for (int p = 0; p < m; p++) { rowof[p] = (uint16_t)row_at[p]; cd[p] = (uint8_t)(row_ptr[row_at[p] + 1] - row_ptr[row_at[p]]); }
    for (int j = 0; j < n; j++) {
        bitof[L.pos[j]] = (uint16_t)j;
        for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) {
            const int i = row_idx[q], e = csc_slot[q];
            vs[(size_t)L.pos[j] * t.DV + (q - col_ptr[j])] = (uint16_t)(L.prow[i] * RS + L.kappa[e]);
        }
    }
    fast_free(t);
    cudaError_t e = cudaMalloc((void **)&t.d_vslot, vs.size() * sizeof(uint16_t));
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_cdeg, cd.size());
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_row_of, rowof.size() * sizeof(uint16_t));
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_bit_of, bitof.size() * sizeof(uint16_t));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_vslot, vs.data(), vs.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_row_of, rowof.data(), rowof.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_bit_of, bitof.data(), bitof.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(t.d_cdeg, cd.data(), cd.size(), cudaMemcpyHostToDevice);
}

// slots after the last row that thread positions beyond the last bit read and write (regular codes: the bit sweep
// then needs no "is this position a bit" guard); never read by a check
constexpr int kFastDummySlots = 8;
constexpr int kFastZeroSlot = kFastDummySlots - 1; // irregular codes: the slot absent edges read, always 0.0

template <typename real>
static inline size_t fast_smem_bytes(const FastTables &t, int n, int m, bool with_priors = true) {
    if (t.DC == 0) return (size_t)1 << 40;
    // the message array doubles as the staging area of the per-shot results ([n] reals + [n] bytes)
    size_t msgs = (std::max(((size_t)m * fast_row_stride(t.DC, (int)sizeof(real)) + kFastDummySlots) * sizeof(real), (size_t)n * (sizeof(real) + 1)) + 15) / 16 * 16;
    size_t meta = ((size_t)m * 4 + 15) / 16 * 16; // one 32-bit word per check (see the kernel)
    // copy of the priors by thread position when they are not uniform (a uniform prior is one register: launches with
    // uniform priors carve no such array, which is what lets a fourth CTA of the bench code fit on an SM)
    size_t prior = with_priors ? ((size_t)n * sizeof(real) + 15) / 16 * 16 : 0;
    return msgs + meta + prior + 16;
}

// ---- vector row load/store helpers -----------------------------------------------------------
template <typename real, int DC> struct RowIO;
// Even DC: 16-byte accesses (rows are 16-byte aligned, see fast_row_stride).  Odd DC (the cluster kernel's rows of 7): one
// element per access -- an odd stride in elements is conflict free as it is, and a row of 7 doubles moves 14 wavefronts
// per warp where the padded row of 8 + 2 moves 20.
template <int DC> struct RowIO<double, DC> {
    __device__ static __forceinline__ void load(const double *p, double (&v)[DC]) {
        if constexpr (DC % 2 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 2) { double2 t = *reinterpret_cast<const double2 *>(p + k); v[k] = t.x; v[k + 1] = t.y; }
        } else {
#pragma unroll
            for (int k = 0; k < DC; k++) v[k] = p[k];
        }
    }
    __device__ static __forceinline__ void store(double *p, const double (&v)[DC]) {
        if constexpr (DC % 2 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 2) *reinterpret_cast<double2 *>(p + k) = make_double2(v[k], v[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < DC; k++) p[k] = v[k];
        }
    }
};
template <int DC> struct RowIO<float, DC> {
    __device__ static __forceinline__ void load(const float *p, float (&v)[DC]) {
        if constexpr (DC % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 4) { float4 t = *reinterpret_cast<const float4 *>(p + k); v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w; }
        } else if constexpr (DC % 2 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 2) { float2 t = *reinterpret_cast<const float2 *>(p + k); v[k] = t.x; v[k + 1] = t.y; }
        } else {
#pragma unroll
            for (int k = 0; k < DC; k++) v[k] = p[k];
        }
    }
    __device__ static __forceinline__ void store(float *p, const float (&v)[DC]) {
        if constexpr (DC % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 4) *reinterpret_cast<float4 *>(p + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        } else if constexpr (DC % 2 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 2) *reinterpret_cast<float2 *>(p + k) = make_float2(v[k], v[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < DC; k++) p[k] = v[k];
        }
    }
};

// bit-level helpers: word that carries the IEEE sign bit in bit 31, |x| on the integer pipe, and
// "x with the high word replaced" (used to build +-alpha without a select)
__device__ __forceinline__ uint32_t sign_word(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ uint32_t sign_word(float x) { return __float_as_uint(x); }
__device__ __forceinline__ double abs_bits(double x) { return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x)); }
__device__ __forceinline__ float abs_bits(float x) { return __uint_as_float(__float_as_uint(x) & 0x7fffffffu); }
__device__ __forceinline__ double with_sign_word(double x, uint32_t w) { return __hiloint2double((int)w, __double2loint(x)); }
__device__ __forceinline__ float with_sign_word(float, uint32_t w) { return __uint_as_float(w); }
template <typename real> __device__ __forceinline__ real lt_min(real a, real b) { return (a < b) ? a : b; } // `if (a < t) t = a`
// the same on magnitudes, carrying the raw value: `if (|a| < |t|) t = a`
template <typename real> __device__ __forceinline__ real mag_min(real a, real b) { return (fabs(a) < fabs(b)) ? a : b; }

// One check of the min-sum check sweep (row a4), in place on its DC-slot row.  `mt` is the check's meta
// byte (bit 7 syndrome, bits 1-5 degree).  Prefix/suffix running minima in the reference's own compare order
// (`if (a < t) t = a`), which also reproduces its NaN propagation on shots whose sums overflowed.
// (A shallower tree of compares for all-finite shots and two rows interleaved per thread were both built, verified
// bit-exact and measured not faster: profiles/r03a_ab_probe.log, r03d_ab_probe.log.)
// fp32 fast mode: sm_100a has a min-sum instruction -- min.xorsign.abs.f32 returns min(|a|, |b|) carrying sign(a) ^ sign(b)
// (FMNMX with the xorsign modifier), so one instruction per prefix / suffix step yields the magnitude AND the sign product of
// "all the other edges"; the fp64 form below needs DSETP + two selects per step plus separate sign-word arithmetic.
#ifndef BPOSD_F32_XORSIGN
#define BPOSD_F32_XORSIGN 1
#endif
__device__ __forceinline__ float min_xorsign_abs(float a, float b) {
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

template <typename real, int DC, bool REG>
__device__ __forceinline__ void fast_check_compute(real (&v)[DC], real (&out)[DC], unsigned mt, real alpha, uint32_t alpha_w) {
    if constexpr (sizeof(real) == 4 && BPOSD_F32_XORSIGN != 0 && DC > 1) {
        float suf[DC];
        suf[DC - 1] = v[DC - 1];
#pragma unroll
        for (int k = DC - 2; k >= 1; k--) suf[k] = min_xorsign_abs(v[k], suf[k + 1]);
        float run = v[0];
        out[0] = suf[1];
#pragma unroll
        for (int k = 1; k < DC - 1; k++) { out[k] = min_xorsign_abs(suf[k + 1], run); run = min_xorsign_abs(v[k], run); }
        out[DC - 1] = run;
        const float all_min = min_xorsign_abs(v[DC - 1], run);
        if (fabsf(all_min) == 0.0f) {
            // some message is +-0: "<= 0" counts +0 as negative, the sign bit does not -> exact path (as in the fp64 form)
            int tot = (int)(mt >> 7);
#pragma unroll
            for (int k = 0; k < DC; k++) tot += ((sign_word(v[k]) >> 31) | (fabsf(v[k]) == 0.0f ? 1u : 0u)) ? 1 : 0;
#pragma unroll
            for (int k = 0; k < DC; k++) {
                const int sg = tot + (((sign_word(v[k]) >> 31) | (fabsf(v[k]) == 0.0f ? 1u : 0u)) ? 1 : 0);
                out[k] = fabsf(out[k]) * ((sg & 1) ? -alpha : alpha);
            }
        } else {
            const float salpha = (mt & 0x80u) ? -alpha : alpha; // the syndrome bit flips every outgoing sign
#pragma unroll
            for (int k = 0; k < DC; k++) out[k] = out[k] * salpha;
        }
        if (!REG) {
            const int deg = (mt >> 1) & 0x1f;
#pragma unroll
            for (int k = 0; k < DC; k++) out[k] = (k < deg) ? out[k] : real_max<real>();
        }
        (void)alpha_w;
        return;
    }
    // Magnitudes are never materialised: the compares take |a| < |b| and select the raw values, the final multiply
    // takes |min|, and sm_100a folds both into operand modifiers of DSETP / DMUL (FSETP / FMUL) -- 12 instructions
    // fewer per row of 6 than clearing the sign bits first.  Compare order and tie behaviour are the reference's.
    real suf[DC];
    uint32_t X = (mt & 0x80u) << 24;
#pragma unroll
    for (int k = 0; k < DC; k++) X ^= sign_word(v[k]);
    // suffix minima first, then one running prefix minimum: out[k] = min(prefix before k, suffix after k)
    suf[DC - 1] = v[DC - 1];
#pragma unroll
    for (int k = DC - 2; k >= 1; k--) suf[k] = mag_min(v[k], suf[k + 1]);
    real run = v[0];
    out[0] = (DC > 1) ? suf[DC > 1 ? 1 : 0] : real_max<real>();
#pragma unroll
    for (int k = 1; k < DC - 1; k++) { out[k] = mag_min(suf[k + 1], run); run = mag_min(v[k], run); }
    if (DC > 1) out[DC - 1] = run;
    const real all_min = (DC > 1) ? mag_min(v[DC - 1], run) : v[0];
    if (fabs(all_min) == (real)0) {
        // some message is +-0: "<= 0" counts +0 as negative, the sign bit does not -> exact path
        int tot = (int)(mt >> 7);
#pragma unroll
        for (int k = 0; k < DC; k++) tot += ((sign_word(v[k]) >> 31) | (fabs(v[k]) == (real)0 ? 1u : 0u)) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < DC; k++) {
            const int sg = tot + (((sign_word(v[k]) >> 31) | (fabs(v[k]) == (real)0 ? 1u : 0u)) ? 1 : 0);
            out[k] = fabs(out[k]) * ((sg & 1) ? -alpha : alpha);
        }
    } else {
        // sign of edge k = syndrome ^ (parity of all sign bits) ^ own sign bit; fold it into alpha
        const uint32_t XA = (X & 0x80000000u) ^ alpha_w;
#pragma unroll
        for (int k = 0; k < DC; k++) out[k] = fabs(out[k]) * with_sign_word(alpha, XA ^ (sign_word(v[k]) & 0x80000000u));
    }
    if (!REG) {
        const int deg = (mt >> 1) & 0x1f;
#pragma unroll
        for (int k = 0; k < DC; k++) out[k] = (k < deg) ? out[k] : real_max<real>();
    }
}

// Product-sum form of the check update (row a5), the reference's own sequence: tanh of every incoming message once;
// forward pass c2b[e_t] = prod_{u<t} tanh (from 1.0, ascending column); reverse pass x = c2b[e_t] * prod_{u>t} tanh
// (accumulated from the last edge backwards), c2b[e_t] = (syndrome ? -1 : +1) * log((1 + x) / (1 - x)).  The layout pass
// keeps a row's edges in ascending-column order for product-sum (products round differently in another order); the pads
// of short rows hold +max, whose tanh is exactly 1.
template <typename real, int DC, bool REG>
__device__ __forceinline__ void fast_check_compute_ps(real (&v)[DC], real (&out)[DC], unsigned mt) {
    real th[DC];
#pragma unroll
    for (int k = 0; k < DC; k++) th[k] = r_tanh(v[k] / 2);
    real t = 1;
#pragma unroll
    for (int k = 0; k < DC; k++) { out[k] = t; t *= th[k]; }
    t = 1;
    const real sgn = (mt & 0x80u) ? (real)-1 : (real)1;
#pragma unroll
    for (int k = DC - 1; k >= 0; k--) {
        const real x = ps_clamp(out[k] * t);
        out[k] = sgn * r_log(ps_ratio(x));
        t *= th[k];
    }
    if (!REG) {
        const int deg = (mt >> 1) & 0x1f;
#pragma unroll
        for (int k = 0; k < DC; k++) out[k] = (k < deg) ? out[k] : real_max<real>();
    }
}

template <typename real, int DC, bool REG, bool PS = false>
__device__ __forceinline__ void fast_check_row(real *row, unsigned mt, real alpha, uint32_t alpha_w) {
    if constexpr (DC % 2 == 1 && !PS) {
        // rows of an odd number of slots (the cluster kernel's rows of 7): computed as the next even class with a constant
        // +max in the last position -- the pad slot the padded layout keeps in memory -- so that both layouts give the same
        // bits in every case, overflowed fp32 shots included (the pad caps a check message at the largest finite value,
        // like the reference's running minimum, which starts from it)
        real v[DC + 1], out[DC + 1];
        real (&vr)[DC] = reinterpret_cast<real (&)[DC]>(v);
        RowIO<real, DC>::load(row, vr);
        v[DC] = real_max<real>();
        fast_check_compute<real, DC + 1, REG>(v, out, mt, alpha, alpha_w);
        RowIO<real, DC>::store(row, reinterpret_cast<real (&)[DC]>(out));
    } else {
        real v[DC], out[DC];
        RowIO<real, DC>::load(row, v);
        if constexpr (PS) fast_check_compute_ps<real, DC, REG>(v, out, mt);
        else fast_check_compute<real, DC, REG>(v, out, mt, alpha, alpha_w);
        RowIO<real, DC>::store(row, out);
    }
}

// Bit sweep (rows a6 + a8) of one thread: VPT positions, each a bit with <= DV edges.  Sums run in the reference's
// order (prior + c_1 + ... left to right for the LLR and the prefix; the suffix from the last edge backwards).
// Returns the hard decisions, bit r = position tid + r*T (garbage for positions past the last bit: the caller masks).
// UNI: every bit has the same prior (a register), else priors come from shared memory by position.
// Regular codes (REG) run without any per-position branch: positions past the last bit own dummy slots.
#ifndef BPOSD_BIT_GUARDS
#define BPOSD_BIT_GUARDS 0 // 1: the older guarded form for every code (A/B)
#endif
#ifndef BPOSD_FLIP_NOHOIST
#define BPOSD_FLIP_NOHOIST 1
#endif
#ifndef BPOSD_T_REG
#define BPOSD_T_REG 1 // profiles/r03h_ab_probe.log: fp64 128.4 -> 129.7, fp32 192.0 -> 195.0 M shot-iterations/s
#endif
#ifndef BPOSD_BIT_GROUP
#define BPOSD_BIT_GROUP 1 // positions whose messages are loaded together before the first store (regular codes); measured on
                          // B200 (profiles/r03h_ab_probe.log): 1 -> 128.4, 2 -> 128.1, 4 -> 127.4, 8 -> 127.1 M shot-iterations/s
#endif
// ZOFF (cluster kernel, irregular codes): an edge is present iff its offset is not the constant-zero slot's (`zoff`), so the
// per-position degrees `dj` need no registers of their own (8 positions x 4 edges + 8 LLRs + 8 degrees do not fit the 96
// registers of a 640-thread CTA: the offsets went to local memory and every position's first LDS waited for an LDL that
// misses L1 after a cluster barrier -- 20 % of the stall samples, profiles/r2ag_cluster_hot_lines.txt).
template <typename real, int DV, int VPT, bool REG, bool UNI, bool ZOFF = false>
__device__ __forceinline__ unsigned fast_bit_sweep(unsigned char *smem_raw, const unsigned (&off)[VPT][DV], const int (&dj)[VPT],
                                                   real (&llr)[VPT], real prior_u, const real *prior_s, int tid, int T, int n,
                                                   unsigned zoff = 0) {
    unsigned dnow = 0;
    if constexpr (BPOSD_BIT_GUARDS == 0 && REG) {
        // regular code: every position has DV slots (real ones or its dummies), nothing is conditional.  The messages of
        // G positions can be loaded before any of them is stored (the compiler cannot move a shared-memory load above a
        // store on its own: it does not know that slots never alias).  The ncu source view shows the DADD after each
        // load triple holding 14 % of the stall samples, but batching the loads buys nothing (see BPOSD_BIT_GROUP): the
        // other warps already cover that latency and the shared-memory pipe, not the wait, sets the pace.
        constexpr int G = (VPT % BPOSD_BIT_GROUP == 0) ? BPOSD_BIT_GROUP : 1;
#pragma unroll
        for (int r0 = 0; r0 < VPT; r0 += G) {
            real c[G][DV];
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int k = 0; k < DV; k++) c[g][k] = *reinterpret_cast<const real *>(smem_raw + off[r0 + g][k]);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int r = r0 + g;
                real pre[DV];
                real t = UNI ? prior_u : prior_s[min(tid + r * T, n - 1)];
#pragma unroll
                for (int k = 0; k < DV; k++) { pre[k] = t; t += c[g][k]; }
                llr[r] = t;
                dnow |= ((t <= 0) ? 1u : 0u) << r;
                real sfx = 0;
#pragma unroll
                for (int k = DV - 1; k >= 0; k--) {
                    // the last edge gets pre + 0, which is pre itself (sign of zero is immaterial downstream)
                    *reinterpret_cast<real *>(smem_raw + off[r][k]) = (k == DV - 1) ? pre[k] : pre[k] + sfx;
                    sfx = (k == DV - 1) ? c[g][k] : sfx + c[g][k];
                }
            }
        }
    } else if constexpr (BPOSD_BIT_GUARDS == 0) {
        // irregular code: an absent edge (k >= degree, or a position past the last bit) reads the constant-zero slot, so
        // the sums need no guard (x + 0 is x, and 0 + 0 is the +0 the running suffix starts from); only the stores are
        // predicated.  Same values as the guarded form below.
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            real c[DV], pre[DV];
#pragma unroll
            for (int k = 0; k < DV; k++) c[k] = *reinterpret_cast<const real *>(smem_raw + off[r][k]);
            real t = UNI ? prior_u : prior_s[min(tid + r * T, n - 1)];
#pragma unroll
            for (int k = 0; k < DV; k++) { pre[k] = t; t += c[k]; }
            llr[r] = t;
            dnow |= ((t <= 0) ? 1u : 0u) << r;
            real sfx = 0;
#pragma unroll
            for (int k = DV - 1; k >= 0; k--) {
                if (ZOFF ? (off[r][k] != zoff) : (k < dj[r])) *reinterpret_cast<real *>(smem_raw + off[r][k]) = pre[k] + sfx;
                sfx = sfx + c[k];
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            const int j = tid + r * T;
            if (j < n) {
                real c[DV], pre[DV];
#pragma unroll
                for (int k = 0; k < DV; k++) c[k] = (REG || k < dj[r]) ? *reinterpret_cast<const real *>(smem_raw + off[r][k]) : (real)0;
                real t = UNI ? prior_u : prior_s[j];
#pragma unroll
                for (int k = 0; k < DV; k++)
                    if (REG || k < dj[r]) { pre[k] = t; t += c[k]; }
                llr[r] = t;
                dnow |= ((t <= 0) ? 1u : 0u) << r;
                real sfx = 0;
#pragma unroll
                for (int k = DV - 1; k >= 0; k--)
                    if (REG || k < dj[r]) {
                        *reinterpret_cast<real *>(smem_raw + off[r][k]) = (REG && k == DV - 1) ? pre[k] : pre[k] + sfx;
                        sfx = (REG && k == DV - 1) ? c[k] : sfx + c[k];
                    }
            }
        }
    }
    return dnow;
}

template <typename real, int DC, int DV, int VPT, int MAXT, bool REG, bool PS = false>
__global__ void __launch_bounds__(MAXT, (fast_minb<real, MAXT, DC, DV, VPT, REG, PS>())) bp_fast_kernel(BpArgs<real> a, const uint16_t *__restrict__ vslot_tab,
                                                       const uint8_t *__restrict__ cdeg_tab,
                                                       const uint16_t *__restrict__ row_of_tab,
                                                       const uint16_t *__restrict__ bit_of_tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.g.m, n = a.g.n;
    const int tid = threadIdx.x;
#if BPOSD_T_REG
    // the CTA size pinned in a register: left to itself the compiler re-reads it from the constant bank at the end of
    // every check row and the next row's address waits for that load (ncu source view: 5 % of the stall samples)
    int T;
    asm volatile("mov.u32 %0, %%ntid.x;" : "=r"(T));
#else
    const int T = blockDim.x;
#endif
    real *msg = reinterpret_cast<real *>(smem_raw);
    constexpr int RS = fast_row_stride(DC, (int)sizeof(real)); // row stride in elements (>= DC, see fast_row_stride)
    const size_t msg_bytes = (((size_t)m * RS + kFastDummySlots) * sizeof(real) > (size_t)n * (sizeof(real) + 1) ? ((size_t)m * RS + kFastDummySlots) * sizeof(real)
                                                                                             : (size_t)n * (sizeof(real) + 1));
    // one word per check: bit0 parity mismatch, bits1-5 degree, bit7 syndrome.  A word, not a byte, so that toggling the
    // mismatch bit is `atomicXor(word, 1)`: one address per edge, where byte-packed flags needed an address and a shifted
    // mask per edge (the compiler keeps those loop invariants in registers: 48 of them for 8 bits x 3 edges)
    unsigned *meta = reinterpret_cast<unsigned *>(smem_raw + (msg_bytes + 15) / 16 * 16);
    real *prior_s = reinterpret_cast<real *>(reinterpret_cast<unsigned char *>(meta) + ((size_t)m * 4 + 15) / 16 * 16); // [n] priors by position when they are not uniform
    real *st_llr = msg;                                         // per-shot result staging (after the last pass)
    uint8_t *st_dec = reinterpret_cast<uint8_t *>(st_llr + n);
    __shared__ long long sh_shot;
    __shared__ int sh_slot;

    // per-position registers: byte offsets of the edges' slots (shot independent).  Position q = tid + r*T
    // holds bit bit_of_tab[q]; only the prior look-up and the result staging need the bit index.
    unsigned off[VPT][DV];
    int dj[VPT];
    unsigned valid = 0; // bit r: position tid + r*T holds a bit
#pragma unroll
    for (int r = 0; r < VPT; r++) {
        const int j = tid + r * T;
        dj[r] = 0;
#pragma unroll
        for (int k = 0; k < DV; k++) {
            const unsigned s = (j < n) ? vslot_tab[(size_t)j * DV + k] : 0xFFFFu;
            off[r][k] = s * (unsigned)sizeof(real);
            dj[r] += (s != 0xFFFFu) ? 1 : 0;
            // regular codes: a position past the last bit works on dummy slots instead of being branched around;
            // irregular codes: every absent edge reads the constant-zero slot (and is never stored)
            if (BPOSD_BIT_GUARDS == 0 && REG && j >= n) off[r][k] = (unsigned)((m * RS + k) * (int)sizeof(real));
            if (BPOSD_BIT_GUARDS == 0 && !REG && s == 0xFFFFu) off[r][k] = (unsigned)((m * RS + kFastZeroSlot) * (int)sizeof(real));
        }
        valid |= (j < n ? 1u : 0u) << r;
    }
    unsigned long long n_conv = 0, n_iter = 0;
    long long next_static = blockIdx.x; // a.queue == nullptr (latency path): CTA b takes shots b, b + grid, ...
    const bool uniform = a.uniform_prior != 0;
    const real prior_u = a.prior[0];

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            if (a.queue) sh_shot = (long long)atomicAdd(a.queue, 1ull);
            else { sh_shot = next_static; next_static += gridDim.x; }
        }
        __syncthreads();
        const long long shot = sh_shot;
        if (shot >= a.B) break;
        const real *prior = a.prior + shot * a.prior_stride;

        for (int p = tid; p < m; p += T) {
            const unsigned s = synd_bit(a.synd, shot, m, row_of_tab[p], a.synd_packed);
            const unsigned deg = cdeg_tab[p];
            meta[p] = s | (deg << 1) | (s << 7);
            if (!REG) // absent slots of short rows hold +max: neutral for min and sign (the result staging
                      // of the previous shot overwrote the array, so they are set again for every shot)
                for (int k = (int)deg; k < DC; k++) msg[p * RS + k] = real_max<real>();
        }
        if (!REG && tid == 0) msg[m * RS + kFastZeroSlot] = (real)0; // (the previous shot's result staging may have covered it)
        real llr[VPT];
        unsigned dprev = 0;
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            const int j = tid + r * T;
            llr[r] = 0;
            if (j < n) {
                const real pj = uniform ? prior_u : prior[bit_of_tab[j]];
                if (!uniform) prior_s[j] = pj;
                llr[r] = pj;
#pragma unroll
                for (int k = 0; k < DV; k++)
                    if (REG || k < dj[r]) *reinterpret_cast<real *>(smem_raw + off[r][k]) = pj;
            }
        }
        __syncthreads();

        bool conv = false;
        int iters = 0;
        int danger = 0; // some LLR came near overflow in the previous pass: rows may hold +-inf (see llr_near_overflow)
        real pow2 = 1; // 2^-it, exact
        for (int it = 1;; it++) {
            const bool last = it > a.max_iter;
            pow2 *= (real)0.5;
            const real alpha = (a.alpha0 == (real)0) ? (real)1 - pow2 : a.alpha0;
            const uint32_t alpha_w = sign_word(alpha);
            bool ok = true;
            if (!PS && danger && !last) { // rare: rewrite +-inf as +-max in this thread's rows (the reference's check update cannot tell them apart)
                for (int p = tid; p < m; p += T)
                    for (int k = 0; k < DC; k++) msg[(size_t)p * RS + k] = clamp_inf<real>(msg[(size_t)p * RS + k]);
            }
            // ---- check sweep (a4) + convergence vote for the previous pass (a7) ----
            // (prefetching the thread's next row into a second register set before updating the current one was measured:
            // 111 vs 128 M shot-iterations/s -- the second set spills at the 128-register cap; profiles/r03i_ab_probe.log)
#pragma unroll kCheckUnroll
            for (int p = tid; p < m; p += T) {
                const unsigned mt = meta[p];
                if (mt & 1u) ok = false;
                if (last) continue;
                fast_check_row<real, DC, REG, PS>(msg + (size_t)p * RS, mt, alpha, alpha_w);
            }
            const int all_ok = __syncthreads_and(ok ? 1 : 0);
            if (it > 1 && all_ok) { conv = true; iters = it - 1; break; }
            if (last) { iters = a.max_iter; break; }
            // ---- bit sweep (a6 + a8) ----
            const unsigned dnow = valid & (uniform ? fast_bit_sweep<real, DV, VPT, REG, true>(smem_raw, off, dj, llr, prior_u, prior_s, tid, T, n)
                                                   : fast_bit_sweep<real, DV, VPT, REG, false>(smem_raw, off, dj, llr, prior_u, prior_s, tid, T, n));
            if (dnow != dprev) {
                // a hard decision flipped (rare): toggle the parity-mismatch bit of every neighbouring check
                unsigned flip = dnow ^ dprev;
                dprev = dnow;
#pragma unroll
                for (int r = 0; r < VPT; r++)
                    if ((flip >> r) & 1u) {
#pragma unroll
                        for (int k = 0; k < DV; k++)
                            if (REG || k < dj[r]) {
                                unsigned o = off[r][k];
                                // register-capped class: keep the slot -> row arithmetic inside this branch (left alone, the
                                // compiler hoists it out of the pass loop and parks one address per edge in a register)
                                if constexpr (kFastSmallClass<real, MAXT, DC, DV, VPT, REG, PS> && BPOSD_FLIP_NOHOIST != 0) asm volatile("" : "+r"(o));
                                const unsigned p = o / (unsigned)(RS * sizeof(real));
                                atomicXor(&meta[p], 1u);
                            }
                    }
            }
            if (!PS && it >= a.safe_it) {
                bool big = false;
#pragma unroll
                for (int r = 0; r < VPT; r++) big |= llr_near_overflow<real>(llr[r]) && ((valid >> r) & 1u);
                danger = __syncthreads_or(big ? 1 : 0);
            } else __syncthreads();
        }

        // ---- results ----
        const bool final_here = conv || a.osd_off;
        if (!final_here && a.fail_count) { // latency path: no list, the host reads the converge flags (LLRs go to a.llr)
            if (tid == 0) {
                const int slot = atomicAdd(a.fail_count, 1);
                a.fail_list[slot] = (int)shot;
                sh_slot = slot;
            }
            __syncthreads();
        }
        // stage the results by bit index in the (now idle) message array, then write them out coalesced
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            const int j = tid + r * T;
            if (j < n) {
                const int b = bit_of_tab[j];
                st_llr[b] = llr[r];
                st_dec[b] = (llr[r] <= 0) ? 1 : 0;
            }
        }
        __syncthreads();
        const long long base = shot * (long long)n;
        for (int j = tid; j < n; j += T) {
            const uint8_t d = st_dec[j];
            if (a.bp) a.bp[base + j] = d;
            if (final_here) {
                if (a.osd0) a.osd0[base + j] = d;
                if (a.osdw) a.osdw[base + j] = d;
            }
            if (a.llr) a.llr[base + j] = st_llr[j];
            else if (!final_here) a.fail_llr[(long long)sh_slot * n + j] = st_llr[j];
        }
        if (tid == 0) {
            if (a.converge) a.converge[shot] = conv ? 1 : 0;
            if (a.iter) a.iter[shot] = iters;
            n_conv += conv ? 1 : 0;
            n_iter += (unsigned long long)iters;
        }
    }
    if (tid == 0 && a.stat) {
        atomicAdd(&a.stat[0], n_conv);
        atomicAdd(&a.stat[1], n_iter);
    }
}

// ---- dispatch over the degree classes ---------------------------------------------------------
// Every (precision, degree class) pair is its own translation unit (bp_fast_inst.cu, compiled once per pair by
// bp_osd_b200/build.py with -DBPOSD_INST_REAL / _DC / _DV): FastInst<real, DC, DV> is declared here for every
// includer and defined (which is what instantiates the kernels) only where BPOSD_FAST_INSTANTIATE is set.
template <typename real, int DC, int DV>
struct FastInst {
    static cudaError_t set_smem(const FastTables &t, int geom, size_t smem);
    static cudaError_t occupancy(const FastTables &t, int geom, int threads, size_t smem, int *occ);
    static void launch(const FastTables &t, int geom, const BpArgs<real> &a, int grid, int threads, int smem, cudaStream_t st);
};

#ifdef BPOSD_FAST_INSTANTIATE
#define BPOSD_FAST_REG1(EXPR) do { if (reg__) { constexpr bool REG = true; EXPR; } else { constexpr bool REG = false; EXPR; } } while (0)
// product-sum kernels exist for the throughput geometries of the degree classes up to (8, 4) (fast_supported)
#define BPOSD_FAST_REG(EXPR)                                                                     \
    do {                                                                                         \
        if constexpr (DC <= 8) {                                                                 \
            if (ps__) { constexpr bool PS = true; BPOSD_FAST_REG1(EXPR); break; }                \
        }                                                                                        \
        { constexpr bool PS = false; BPOSD_FAST_REG1(EXPR); }                                    \
    } while (0)
#define BPOSD_FAST_REG_MS(EXPR) do { constexpr bool PS = false; BPOSD_FAST_REG1(EXPR); } while (0)
#define BPOSD_FAST_GEOM(t, geom, EXPR)                                                           \
    do {                                                                                         \
        const int geom__ = (geom);                                                               \
        const bool reg__ = t.regular != 0;                                                       \
        const bool ps__ = t.ps != 0;                                                             \
        (void)ps__;                                                                              \
        if (geom__ == 0) { constexpr int VPT = 2, MAXT = 128; BPOSD_FAST_REG(EXPR); }            \
        else if (geom__ == 1) { constexpr int VPT = BPOSD_MID_VPT, MAXT = BPOSD_MID_MAXT; BPOSD_FAST_REG(EXPR); } \
        else if (geom__ == 4) { constexpr int VPT = sizeof(real) == 8 ? BPOSD_LAT_VPT64 : BPOSD_LAT_VPT32,               \
                                              MAXT = sizeof(real) == 8 ? BPOSD_LAT_MAXT64 : BPOSD_LAT_MAXT32; BPOSD_FAST_REG_MS(EXPR); } \
        else if (geom__ == 2) { constexpr int VPT = 8, MAXT = 512; BPOSD_FAST_REG(EXPR); }       \
        else { constexpr int VPT = 8, MAXT = 1024; BPOSD_FAST_REG(EXPR); }                       \
    } while (0)

template <typename real, int DC, int DV>
cudaError_t FastInst<real, DC, DV>::set_smem(const FastTables &t, int geom, size_t smem) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_GEOM(t, geom, e = cudaFuncSetAttribute(bp_fast_kernel<real, DC, DV, VPT, MAXT, REG, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return e;
}
template <typename real, int DC, int DV>
cudaError_t FastInst<real, DC, DV>::occupancy(const FastTables &t, int geom, int threads, size_t smem, int *occ) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_GEOM(t, geom, e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, bp_fast_kernel<real, DC, DV, VPT, MAXT, REG, PS>, threads, smem));
    return e;
}
template <typename real, int DC, int DV>
void FastInst<real, DC, DV>::launch(const FastTables &t, int geom, const BpArgs<real> &a, int grid, int threads, int smem, cudaStream_t st) {
    BPOSD_FAST_GEOM(t, geom, (bp_fast_kernel<real, DC, DV, VPT, MAXT, REG, PS><<<grid, threads, smem, st>>>(a, t.d_vslot, t.d_cdeg, t.d_row_of, t.d_bit_of)));
}
#endif // BPOSD_FAST_INSTANTIATE

#ifdef BPOSD_DEV_CLASS_ONLY // tuning builds (BPOSD_DEV_CLASS=6,3 python -m bp_osd_b200.build): one degree class alone
#define BPOSD_FAST_CLASS(t, EXPR) do { constexpr int DC = BPOSD_DEV_DC, DV = BPOSD_DEV_DC / 2; EXPR; } while (0)
#else
#define BPOSD_FAST_CLASS(t, EXPR)                                                                \
    do {                                                                                         \
        if (t.DC == 4) { constexpr int DC = 4, DV = 2; EXPR; }                                   \
        else if (t.DC == 6) { constexpr int DC = 6, DV = 3; EXPR; }                              \
        else if (t.DC == 8) { constexpr int DC = 8, DV = 4; EXPR; }                              \
        else { constexpr int DC = 16, DV = 8; EXPR; }                                            \
    } while (0)
#endif

template <typename real>
static inline cudaError_t fast_set_smem_t(const FastTables &t, int geom, size_t smem) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_CLASS(t, (e = FastInst<real, DC, DV>::set_smem(t, geom, smem)));
    return e;
}
template <typename real>
static inline cudaError_t fast_occupancy_t(const FastTables &t, int geom, int threads, size_t smem, int *occ) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_CLASS(t, (e = FastInst<real, DC, DV>::occupancy(t, geom, threads, smem, occ)));
    return e;
}
template <typename real>
static inline void fast_launch(const FastTables &t, int geom, const BpArgs<real> &a, int grid, int threads, int smem, cudaStream_t st) {
    BPOSD_FAST_CLASS(t, (FastInst<real, DC, DV>::launch(t, geom, a, grid, threads, smem, st)));
}

} // namespace bposd
