// bp_fast_kernel.cuh -- in-place shared-memory min-sum BP for sm_100a (rows a3-a8, bit-exact in fp64).
//
// One message array, check-major with a fixed row stride DC (slot = check*DC + k, k-th edge of
// the check in ascending column order).  A pass is two sweeps over it:
//   check sweep: the thread that owns a check pulls its whole row with 16-byte LDS, reduces
//                min1/min2/argmin/sign parity in registers, and overwrites the row with the
//                check->bit messages (16-byte STS);
//   bit sweep:   the thread that owns a bit gathers its <= DV messages, forms the prefix/suffix
//                sums in the reference's order (prior + c_1 + ... left to right; suffix from the
//                last edge backwards), overwrites them with the new bit->check messages and drops
//                its hard decision next to each edge (one byte per edge, row stride DS) so the next
//                check sweep can test H*decoding == syndrome while it reads the row anyway.
// Per-bit state that must survive to the end of the shot (the LLR of the last executed pass and
// the edge slots) lives in registers; nothing but the syndrome and the results touches HBM.
// Shots are pulled from an atomic queue by persistent CTAs, one shot per CTA at a time.
#pragma once
#include "bposd_kernels.cuh"
#include <vector>

namespace bposd {

struct FastTables {
    int DC = 0, DV = 0;            // degree class (upper bounds, compile-time in the kernel)
    uint16_t *d_vslot = nullptr;   // [n, DV] message slot of the k-th edge of bit j (ascending row), 0xFFFF = none
    uint8_t *d_cdeg = nullptr;     // [m] row degrees
};

static inline bool fast_supported(int max_col_deg, int max_row_deg, int method) {
    return method == 1 && max_col_deg >= 1 && max_col_deg <= 8 && max_row_deg >= 1 && max_row_deg <= 16;
}

static inline void fast_class(int max_col_deg, int max_row_deg, int *DC, int *DV) {
    if (max_row_deg <= 4 && max_col_deg <= 2) { *DC = 4; *DV = 2; }
    else if (max_row_deg <= 6 && max_col_deg <= 3) { *DC = 6; *DV = 3; }
    else if (max_row_deg <= 8 && max_col_deg <= 4) { *DC = 8; *DV = 4; }
    else { *DC = 16; *DV = 8; }
}

static inline int fast_dstride(int DC) { return (DC + 7) / 8 * 8; }

static inline void fast_free(FastTables &t) {
    cudaFree(t.d_vslot); cudaFree(t.d_cdeg);
    t.d_vslot = nullptr; t.d_cdeg = nullptr;
}

static inline cudaError_t fast_build(FastTables &t, int m, int n, const std::vector<int> &row_ptr,
                                     const std::vector<int> &col_idx, const std::vector<int> &col_ptr,
                                     const std::vector<int> &row_idx, const std::vector<int> &csc_slot) {
    int mr = 0, mc = 0;
    for (int i = 0; i < m; i++) mr = std::max(mr, row_ptr[i + 1] - row_ptr[i]);
    for (int j = 0; j < n; j++) mc = std::max(mc, col_ptr[j + 1] - col_ptr[j]);
    fast_class(mc, mr, &t.DC, &t.DV);
    if ((long long)m * t.DC >= 0xFFFF) { t.DC = 0; return cudaSuccess; } // slots must fit in 16 bits
    std::vector<uint16_t> vs((size_t)n * t.DV, 0xFFFF);
    std::vector<uint8_t> cd(std::max(m, 1), 0);
    for (int i = 0; i < m; i++) cd[i] = (uint8_t)(row_ptr[i + 1] - row_ptr[i]);
    for (int j = 0; j < n; j++)
        for (int q = col_ptr[j]; q < col_ptr[j + 1]; q++) {
            const int i = row_idx[q], k = csc_slot[q] - row_ptr[i];
            vs[(size_t)j * t.DV + (q - col_ptr[j])] = (uint16_t)(i * t.DC + k);
        }
    (void)col_idx;
    cudaError_t e = cudaMalloc((void **)&t.d_vslot, vs.size() * sizeof(uint16_t));
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void **)&t.d_cdeg, cd.size());
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t.d_vslot, vs.data(), vs.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(t.d_cdeg, cd.data(), cd.size(), cudaMemcpyHostToDevice);
}

template <typename real>
static inline size_t fast_smem_bytes(const FastTables &t, int n, int m) {
    (void)n;
    if (t.DC == 0) return (size_t)1 << 40;
    size_t msgs = ((size_t)m * t.DC * sizeof(real) + 15) / 16 * 16;
    size_t dbits = ((size_t)m * fast_dstride(t.DC) + 15) / 16 * 16;
    size_t meta = ((size_t)m + 15) / 16 * 16;
    return msgs + dbits + meta + 16;
}

static inline int fast_vpt(int n) { return n <= 4096 ? 4 : 8; }

static inline int fast_default_threads(int n, int m) {
    (void)m;
    const int vpt = fast_vpt(n);
    int t = ((n + vpt - 1) / vpt + 31) / 32 * 32;
    return std::min(1024, std::max(32, t));
}

// ---- vector row load/store helpers -----------------------------------------------------------
template <typename real, int DC> struct RowIO;
template <int DC> struct RowIO<double, DC> {
    static_assert(DC % 2 == 0, "row stride must keep 16-byte alignment");
    __device__ static __forceinline__ void load(const double *p, double (&v)[DC]) {
#pragma unroll
        for (int k = 0; k < DC; k += 2) { double2 t = *reinterpret_cast<const double2 *>(p + k); v[k] = t.x; v[k + 1] = t.y; }
    }
    __device__ static __forceinline__ void store(double *p, const double (&v)[DC]) {
#pragma unroll
        for (int k = 0; k < DC; k += 2) *reinterpret_cast<double2 *>(p + k) = make_double2(v[k], v[k + 1]);
    }
};
template <int DC> struct RowIO<float, DC> {
    static_assert(DC % 2 == 0, "row stride must keep 8-byte alignment");
    __device__ static __forceinline__ void load(const float *p, float (&v)[DC]) {
        if (DC % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 4) { float4 t = *reinterpret_cast<const float4 *>(p + k); v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w; }
        } else {
#pragma unroll
            for (int k = 0; k < DC; k += 2) { float2 t = *reinterpret_cast<const float2 *>(p + k); v[k] = t.x; v[k + 1] = t.y; }
        }
    }
    __device__ static __forceinline__ void store(float *p, const float (&v)[DC]) {
        if (DC % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DC; k += 4) *reinterpret_cast<float4 *>(p + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < DC; k += 2) *reinterpret_cast<float2 *>(p + k) = make_float2(v[k], v[k + 1]);
        }
    }
};

template <typename real, int DC, int DV, int VPT>
__global__ void __launch_bounds__(1024) bp_fast_kernel(BpArgs<real> a, const uint16_t *__restrict__ vslot_tab,
                                                       const uint8_t *__restrict__ cdeg_tab) {
    constexpr int DS = (DC + 7) / 8 * 8; // hard-decision bytes per check row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.g.m, n = a.g.n;
    const int tid = threadIdx.x, T = blockDim.x;
    real *msg = reinterpret_cast<real *>(smem_raw);
    uint8_t *dbit = smem_raw + ((size_t)m * DC * sizeof(real) + 15) / 16 * 16;
    uint8_t *meta = dbit + ((size_t)m * DS + 15) / 16 * 16;
    __shared__ long long sh_shot;
    __shared__ int sh_slot;

    // per-bit registers: edge slots and degree (shot independent), LLR (per shot)
    uint16_t sl[VPT][DV];
    int dj[VPT];
#pragma unroll
    for (int r = 0; r < VPT; r++) {
        const int j = tid + r * T;
        dj[r] = 0;
#pragma unroll
        for (int k = 0; k < DV; k++) {
            sl[r][k] = (j < n) ? vslot_tab[(size_t)j * DV + k] : (uint16_t)0xFFFF;
            dj[r] += (sl[r][k] != 0xFFFF) ? 1 : 0;
        }
    }
    // absent slots of short rows hold +max forever: neutral for min and sign
    for (int i = tid; i < m; i += T) {
        const int d = cdeg_tab[i];
        for (int k = d; k < DC; k++) msg[i * DC + k] = real_max<real>();
        for (int k = 0; k < DS; k++) dbit[i * DS + k] = 0;
    }
    unsigned long long n_conv = 0, n_iter = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh_shot = (long long)atomicAdd(a.queue, 1ull);
        __syncthreads();
        const long long shot = sh_shot;
        if (shot >= a.B) break;
        const real *prior = a.prior + shot * a.prior_stride;

        for (int i = tid; i < m; i += T) meta[i] = (uint8_t)(cdeg_tab[i] | ((a.synd[shot * m + i] & 1) << 7));
        real llr[VPT];
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            const int j = tid + r * T;
            llr[r] = 0;
            if (j < n) {
                const real p = prior[j];
                llr[r] = p;
#pragma unroll
                for (int k = 0; k < DV; k++)
                    if (k < dj[r]) msg[sl[r][k]] = p;
            }
        }
        __syncthreads();

        bool conv = false;
        int iters = 0;
        for (int it = 1;; it++) {
            const bool last = it > a.max_iter;
            const real alpha = ms_alpha(a.alpha0, it);
            bool ok = true;
            // ---- check sweep (a4) + convergence vote for the previous pass (a7) ----
            for (int i = tid; i < m; i += T) {
                real v[DC];
                RowIO<real, DC>::load(msg + (size_t)i * DC, v);
                unsigned par = 0;
#pragma unroll
                for (int k = 0; k < DS; k += 8) {
                    const uint2 d = *reinterpret_cast<const uint2 *>(dbit + (size_t)i * DS + k);
                    par ^= d.x ^ d.y;
                }
                par ^= par >> 16; par ^= par >> 8;
                const unsigned mt = meta[i];
                const int deg = mt & 0x7f, s = mt >> 7;
                if (it > 1 && (int)(par & 1u) != s) ok = false;
                if (!last) {
                    real min1 = real_max<real>(), min2 = real_max<real>();
                    int arg = -1, tot = s;
#pragma unroll
                    for (int k = 0; k < DC; k++) {
                        const real av = r_abs(v[k]);
                        tot += (v[k] <= 0) ? 1 : 0;
                        if (av < min1) { min2 = min1; min1 = av; arg = k; }
                        else if (av < min2) min2 = av;
                    }
#pragma unroll
                    for (int k = 0; k < DC; k++) {
                        const int sg = tot + ((v[k] <= 0) ? 1 : 0);
                        const real mag = (k == arg) ? min2 : min1;
                        const real out = mag * ((sg & 1) ? -alpha : alpha);
                        v[k] = (k < deg) ? out : real_max<real>();
                    }
                    RowIO<real, DC>::store(msg + (size_t)i * DC, v);
                }
            }
            const int all_ok = __syncthreads_and(ok ? 1 : 0);
            if (it > 1 && all_ok) { conv = true; iters = it - 1; break; }
            if (last) { iters = a.max_iter; break; }
            // ---- bit sweep (a6 + a8) ----
#pragma unroll
            for (int r = 0; r < VPT; r++) {
                const int j = tid + r * T;
                if (j < n) {
                    real c[DV], pre[DV];
#pragma unroll
                    for (int k = 0; k < DV; k++) c[k] = (k < dj[r]) ? msg[sl[r][k]] : (real)0;
                    real t = prior[j];
#pragma unroll
                    for (int k = 0; k < DV; k++)
                        if (k < dj[r]) { pre[k] = t; t += c[k]; }
                    llr[r] = t;
                    const uint8_t d = (t <= 0) ? 1 : 0;
                    real sfx = 0;
#pragma unroll
                    for (int k = DV - 1; k >= 0; k--)
                        if (k < dj[r]) {
                            const unsigned s16 = sl[r][k];
                            msg[s16] = pre[k] + sfx;
                            sfx += c[k];
                            dbit[s16 + (s16 / DC) * (DS - DC)] = d;
                        }
                }
            }
            __syncthreads();
        }

        // ---- results ----
        const bool final_here = conv || a.osd_off;
        if (!final_here) {
            if (tid == 0) {
                const int slot = atomicAdd(a.fail_count, 1);
                a.fail_list[slot] = (int)shot;
                sh_slot = slot;
            }
            __syncthreads();
        }
        const long long base = shot * (long long)n;
#pragma unroll
        for (int r = 0; r < VPT; r++) {
            const int j = tid + r * T;
            if (j < n) {
                const uint8_t d = (llr[r] <= 0) ? 1 : 0;
                if (a.bp) a.bp[base + j] = d;
                if (final_here) {
                    if (a.osd0) a.osd0[base + j] = d;
                    if (a.osdw) a.osdw[base + j] = d;
                }
                if (a.llr) a.llr[base + j] = llr[r];
                else if (!final_here) a.fail_llr[(long long)sh_slot * n + j] = llr[r];
            }
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
#define BPOSD_FAST_DISPATCH(real, t, n, EXPR)                                                    \
    do {                                                                                         \
        const int vpt__ = fast_vpt(n);                                                           \
        if (t.DC == 4 && vpt__ == 4) { constexpr int DC = 4, DV = 2, VPT = 4; EXPR; }            \
        else if (t.DC == 4) { constexpr int DC = 4, DV = 2, VPT = 8; EXPR; }                     \
        else if (t.DC == 6 && vpt__ == 4) { constexpr int DC = 6, DV = 3, VPT = 4; EXPR; }       \
        else if (t.DC == 6) { constexpr int DC = 6, DV = 3, VPT = 8; EXPR; }                     \
        else if (t.DC == 8 && vpt__ == 4) { constexpr int DC = 8, DV = 4, VPT = 4; EXPR; }       \
        else if (t.DC == 8) { constexpr int DC = 8, DV = 4, VPT = 8; EXPR; }                     \
        else if (vpt__ == 4) { constexpr int DC = 16, DV = 8, VPT = 4; EXPR; }                   \
        else { constexpr int DC = 16, DV = 8, VPT = 8; EXPR; }                                   \
    } while (0)

template <typename real>
static inline cudaError_t fast_set_smem_t(const FastTables &t, int n, size_t smem) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_DISPATCH(real, t, n, e = cudaFuncSetAttribute(bp_fast_kernel<real, DC, DV, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return e;
}
template <typename real>
static inline cudaError_t fast_occupancy_t(const FastTables &t, int n, int threads, size_t smem, int *occ) {
    cudaError_t e = cudaSuccess;
    BPOSD_FAST_DISPATCH(real, t, n, e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, bp_fast_kernel<real, DC, DV, VPT>, threads, smem));
    return e;
}
template <typename real>
static inline void fast_launch(const FastTables &t, const BpArgs<real> &a, int grid, int threads, int smem, cudaStream_t st) {
    BPOSD_FAST_DISPATCH(real, t, a.g.n, (bp_fast_kernel<real, DC, DV, VPT><<<grid, threads, smem, st>>>(a, t.d_vslot, t.d_cdeg)));
}

} // namespace bposd
