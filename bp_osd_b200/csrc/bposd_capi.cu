// bposd_capi.cu -- host side of libbposd_b200.so: the C ABI declared in include/bposd_b200.h.
//
// Compiles the Tanner graph once per parity-check matrix (CSR + CSC + slot maps on the device),
// picks a BP kernel variant and launch geometry from the code size, owns the workspaces, and
// sequences BP -> OSD (-> logical check) on the caller's stream.  There is no CPU fallback:
// every entry point either runs the CUDA kernels or returns an error.
#include "../../include/bposd_b200.h"
#include "bposd_kernels.cuh"
#include "bp_fast_kernel.cuh"
#include "bp_cluster_kernel.cuh"
#include "osd_reg_kernel.cuh"
#include "osd_cluster_kernel.cuh"
#include "bp_serial_kernel.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace bposd;

struct bposd_handle {
    int device = 0, sm_count = 0, smem_optin = 0;
    int m = 0, n = 0, nnz = 0, rank = 0, k = 0;
    int max_iter = 0, bp_method = 1, osd_method = 0, osd_order = 0, precision = 64;
    double alpha = 1.0;
    std::vector<int> row_ptr, col_idx, col_ptr, row_idx, csc_slot;
    std::vector<double> probs;
    int max_row_deg = 0, max_col_deg = 0;
    // device copies
    int *d_row_ptr = nullptr, *d_col_idx = nullptr, *d_col_ptr = nullptr, *d_row_idx = nullptr, *d_csc_slot = nullptr;
    double *d_prior64 = nullptr, *d_weight = nullptr;
    float *d_prior32 = nullptr;
    int uniform = 0, uniform_prior = 0;
    int safe_it = 0; // overflow guard: first pass at which a message could have overflowed with the static channel (fp64)
    // fast-kernel tables
    FastTables fast;
    // cluster-kernel tables (built on demand, when the messages exceed one SM's shared memory)
    ClusterTables clus;
    int clus_nclusters = 0, clus_flip_table = 0;
    // the same for launches whose bits all share one prior: no prior array in shared memory, so a smaller cluster may hold
    // the code (more shots in flight, fewer remote edges); CL == 0: no separate plan, such launches use `clus`
    ClusterTables clus_u;
    int clus_u_nclusters = 0, clus_u_flip_table = 0, clus_u_threads = 0, clus_u_smem = 0;
    // Two decode slots: each owns its control words, failed-shot workspace, staging buffers, events and
    // (for the host-buffer pipeline) a stream, so chunk i+1 can be copied in while chunk i decodes.
    struct Slot {
        // control words: [0] queue, [1] converged, [2] iterations, [3] osd invocations, [4] (int) fail_count
        unsigned long long *d_ctrl = nullptr;
        unsigned char *d_serial_ws = nullptr; // serial-schedule BP: per-warp message workspace (bp_serial_kernel.cuh)
        size_t serial_ws_bytes = 0;
        unsigned long long *h_ctrl = nullptr; // pinned copy of words 0..3
        int *d_fail_list = nullptr;
        void *d_fail_llr = nullptr;
        long long fail_list_cap = 0, fail_llr_cap = 0;
        uint8_t *b_synd = nullptr, *b_err = nullptr, *b_osdw = nullptr, *b_osd0 = nullptr, *b_bp = nullptr, *b_conv = nullptr;
        uint8_t *b_pk_osdw = nullptr, *b_pk_osd0 = nullptr, *b_pk_bp = nullptr; // bit-packed copies of the decodings ([cap, ceil(n/8)])
        void *b_llr = nullptr;
        int32_t *b_iter = nullptr;
        long long b_cap = 0;
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // BP start, BP end, OSD end, everything enqueued for the chunk
        cudaStream_t stream = nullptr;
        bool pending = false;
        long long pending_shots = 0;
        int pending_launches = 0;
    } slot[2];
    long long fail_cap = 0; // shots per chunk the failed-shot workspace can hold
    long long workspace_bytes = 2ll << 30;
    bool workspace_user_set = false;
    long long host_chunk_min = 16384, host_chunk_max = 131072; // shots per chunk of the host-buffer pipeline
    void *d_scratch = nullptr; // global-mode BP scratch
    uint8_t *d_scratch_dec = nullptr;
    size_t scratch_bytes = 0;
    // launch geometry
    int bp_kernel = 1, bp_threads = 256, bp_ctas_per_sm = 1, bp_smem = 0, bp_grid = 0;
    int bp_smem_uni = 0, bp_grid_uni = 0, bp_ctas_uni = 0; // fast kernel, launches with a uniform prior (no prior array in shared memory)
    int bp_geom = 1;                    // fast kernel geometry id (bp_fast_kernel.cuh: fast_geom)
    int lat_geom = -1, lat_threads = 0; // latency geometry of the fast kernel (-1: none; small host batches use it)
    long long lat_max_shots = 0;        // host batches up to this size take the latency path (0: disabled)
    int osd_threads = 256, osd_smem = 0, osd_ctas_per_sm = 1, osd_S = 0, osd_St = 0;
    bool osd_supported = true;
    // large-H OSD-0 (T does not fit in shared memory): HBM workspace, one CTA per failed shot
    bool osd_large = false;
    int osd_variant = 0; // 0 auto, 1 T-matrix shared-memory kernel, 2 HBM-resident kernel, 3 T-matrix register kernel
    bool osd_reg = false; // the register kernel is selected (m <= 1024)
    int osdr_W = 0, osdr_KD = 4, osdr_threads = 0, osdr_smem = 0;
    int osdl_smem = 0, osdl_grid = 0, osdl_npanels = 0;
    long long osdl_ws_cap = 16ll << 30;
    uint32_t *d_osdl_mask = nullptr;
    int *d_osdl_order = nullptr, *d_osdl_piv_row = nullptr, *d_osdl_piv_pos = nullptr, *d_osdl_pstart = nullptr;
    int osdl_alloc_grid = 0;
    // large-H OSD-0, one thread-block cluster per failed shot (osd_cluster_kernel.cuh): rows split over the cluster, masks TMA-streamed
    bool osd_clus = false;
    struct OsdcCfg { int CL, rpc, smem, ncl, stages; };
    std::vector<OsdcCfg> osdc; // cluster sizes, largest first: a few failed shots take the big clusters, many the small ones
    int osdc_npanels = 0;
    size_t osdc_alloc_mask = 0, osdc_alloc_shots = 0;
    unsigned long long *d_osdc_mask = nullptr, *d_osdc_key = nullptr;
    unsigned *d_osdc_idx = nullptr;
    OsdcPivot *d_osdc_piv = nullptr;
    int force_kernel = 0, force_threads = 0, force_cluster = 0;
    int schedule = 0;       // 0 parallel (flooding), 1 serial (bit after bit)
    int *d_order = nullptr; // serial schedule: bit order (nullptr: 0 .. n-1)
    bool geometry_ready = false;
    // harness
    uint32_t *d_t1 = nullptr, *d_t2 = nullptr, *d_t3 = nullptr;
    int *d_l_ptr = nullptr, *d_l_idx = nullptr;
    int K = 0;
    // internal buffers for decode_host / sample_and_decode
    unsigned long long *d_counters = nullptr; // 8 words
    int *d_minw = nullptr;
    double *d_cu_tab = nullptr; // channel-update tables [4, n]
    uint8_t *h_stage = nullptr; // pinned, device-mapped staging of the small-batch host paths (latency)
    uint8_t *d_stage = nullptr; // the same block as the device addresses it
    size_t stage_bytes = 4 << 20;
    bposd_stats_t stats{};
    std::string err;
};

#define CU_TRY(h, call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
            return BPOSD_ECUDA;                                                                  \
        }                                                                                        \
    } while (0)

static int fail(bposd_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

static thread_local std::string g_create_err;

template <typename T>
static cudaError_t upload(T **dst, const std::vector<T> &src) {
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void **)dst, bytes);
    if (e != cudaSuccess) return e;
    if (!src.empty()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static int host_rank(const bposd_handle *h) {
    // GF(2) rank of H by packed row elimination; once per code (row a1: k = n - rank)
    const int m = h->m, n = h->n, W = (n + 63) / 64;
    std::vector<uint64_t> a((size_t)m * W, 0);
    for (int i = 0; i < m; i++)
        for (int e = h->row_ptr[i]; e < h->row_ptr[i + 1]; e++)
            a[(size_t)i * W + h->col_idx[e] / 64] ^= 1ull << (h->col_idx[e] % 64);
    int r = 0;
    for (int c = 0; c < n && r < m; c++) {
        const int w = c / 64;
        const uint64_t bit = 1ull << (c % 64);
        int p = -1;
        for (int i = r; i < m; i++)
            if (a[(size_t)i * W + w] & bit) { p = i; break; }
        if (p < 0) continue;
        if (p != r) std::swap_ranges(a.begin() + (size_t)p * W, a.begin() + (size_t)(p + 1) * W, a.begin() + (size_t)r * W);
        for (int i = r + 1; i < m; i++)
            if (a[(size_t)i * W + w] & bit)
                for (int x = w; x < W; x++) a[(size_t)i * W + x] ^= a[(size_t)r * W + x];
        r++;
    }
    return r;
}

// First pass at which a message could have overflowed (overflow guard of the prefix / suffix check update, see
// llr_near_overflow in bposd_kernels.cuh).  With P the largest |prior| and d the largest column degree, the largest
// magnitude of any message or LLR after pass t is at most P (d + 1)^t: a check message never exceeds the largest
// bit-to-check message, and a bit sums its prior and at most d check messages.  Guarding starts while that bound is
// still below the threshold the kernels test against (1e291).  Per-shot priors are not known on the host: 0.
static int overflow_safe_iterations(const bposd_handle *h, bool per_shot_priors, bool fp32) {
    // fp32 fast mode: not guarded.  It promises no bit-exactness, a shot whose messages overflow is a shot that does not
    // converge (OSD takes it either way), and fp32 would reach the guarded range at pass ~48, inside the bulk of the
    // iterations: 247.9 -> 232.9 M shot-iterations/s on the bench workload (profiles/r2y_bench_fp32.json).
    if (fp32) return 0x7fffffff;
    if (per_shot_priors) return 0;
    double P = 1.0;
    for (int j = 0; j < h->n; j++) {
        const double p = h->probs[j];
        const double l = std::fabs(std::log((1.0 - p) / p));
        if (!(l <= 1e30)) return 0; // p = 0 or 1: infinite priors from the start
        P = std::max(P, l);
    }
    const double lim = std::log(fp32 ? 1e30 : 1e291) - std::log(P) - std::log(4.0);
    const double per = std::log((double)std::max(h->max_col_deg, 1) + 1.0);
    const double t = lim / per;
    return t < 1.0 ? 0 : (t > 1e9 ? 1000000000 : (int)t);
}

static int upload_probs(bposd_handle *h) {
    const int n = h->n;
    std::vector<double> prior(n), weight(n);
    std::vector<float> prior32(n);
    bool uni = n > 0;
    for (int j = 0; j < n; j++) {
        const double p = h->probs[j];
        prior[j] = std::log((1.0 - p) / p); // row a3
        weight[j] = std::log(1 / p);        // row a14
        prior32[j] = (float)prior[j];
        if (!(p == h->probs[0])) uni = false;
    }
    h->uniform_prior = uni ? 1 : 0; // all priors bit-identical (any value, including +-inf)
    h->safe_it = overflow_safe_iterations(h, false, false);
    if (uni && !(h->probs[0] > 0.0 && h->probs[0] < 1.0)) uni = false;
    h->uniform = uni ? 1 : 0;       // OSD weights: popcount ordering is exact only for 0 < p < 1
    CU_TRY(h, cudaMemcpy(h->d_prior64, prior.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_prior32, prior32.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_weight, weight.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    return BPOSD_OK;
}

static GraphDev graph_of(const bposd_handle *h) {
    GraphDev g;
    g.m = h->m; g.n = h->n; g.nnz = h->nnz;
    g.row_ptr = h->d_row_ptr; g.col_idx = h->d_col_idx;
    g.col_ptr = h->d_col_ptr; g.row_idx = h->d_row_idx; g.csc_slot = h->d_csc_slot;
    return g;
}

template <typename real>
static int plan_geometry_t(bposd_handle *h) {
    const size_t rs = sizeof(real);
    const int n = h->n, m = h->m, E = h->nnz;
    int threads = h->force_threads > 0 ? h->force_threads : std::min(1024, std::max(32, ((n + 1) / 2 + 31) / 32 * 32));
    threads = std::min(1024, (threads + 31) / 32 * 32);
    // candidate kernels, best first
    int kernel = -1;
    size_t smem = 0;
    const bool fast_ok = fast_supported(h->max_col_deg, h->max_row_deg, h->bp_method) && h->fast.DC > 0 && n <= 8192;
    const size_t smem_fast = fast_ok ? fast_smem_bytes<real>(h->fast, n, m) : ((size_t)1 << 40);
    const size_t smem_gen = (2 * (size_t)E + n) * rs + n + m + 16;
    const int want = h->force_kernel - 1;
    if ((want < 0 || want == 2) && fast_ok && smem_fast <= (size_t)h->smem_optin) {
        kernel = 2; smem = smem_fast;
        if (h->force_threads <= 0) threads = fast_default_threads(n, m);
    } else if ((want < 0 || want == 1 || want == 2) && smem_gen <= (size_t)h->smem_optin) {
        kernel = 1; smem = smem_gen;
    } else {
        kernel = 0; smem = (size_t)m + 16;
    }
    if (want == 0) { kernel = 0; smem = (size_t)m + 16; }
    // cluster kernel: min-sum, supported degrees, and either forced or nothing smem-resident fits
    if ((want == 3 || (want < 0 && kernel == 0)) && h->bp_method == 1 && fast_supported(h->max_col_deg, h->max_row_deg, h->bp_method) && m > 0) {
        int DCc = 0, DVc = 0;
        cluster_class(h->max_col_deg, h->max_row_deg, &DCc, &DVc);
        // candidate cluster sizes: the smallest cluster whose per-CTA slice fits in shared memory first (fewer remote
        // edges, more shots in flight), larger ones as fall-backs if the device cannot schedule it
        struct Plan { int CL = 0, threads = 0, smem = 0, ncl = 0, flip = 0; };
        auto plan_cluster = [&](ClusterTables &tab, int prior_table, Plan &out) -> int {
            out = Plan();
            for (int CL : {2, 4, 8, 16}) {
                if (h->force_cluster > 0 && CL != h->force_cluster) continue;
                const int rpc = (m + CL - 1) / CL, bpc = (n + CL - 1) / CL;
                if (cluster_smem_min<real>(DCc, rpc, bpc, prior_table) > (size_t)h->smem_optin || cluster_vpt(bpc) == 0) continue;
                if (tab.CL != CL || tab.elem_bytes != (int)rs) {
                    cudaError_t e = cluster_build(tab, CL, m, n, h->row_ptr, h->col_idx, h->col_ptr, h->row_idx, h->csc_slot, (int)rs);
                    if (e != cudaSuccess) return fail(h, BPOSD_ECUDA, std::string("cluster_build: ") + cudaGetErrorString(e));
                }
                const int ct = cluster_threads(tab.bits_per_cta);
                // parity-flip descriptors of the remote edges in shared memory when they fit: 32-bit, else 16-bit, else none
                size_t csmem = cluster_smem_need<real>(tab, 1, prior_table);
                int flip = 1;
                if (csmem > (size_t)h->smem_optin && tab.rows_per_cta <= 4096 && CL <= 16) { csmem = cluster_smem_need<real>(tab, 2, prior_table); flip = 2; }
                if (csmem > (size_t)h->smem_optin) { csmem = cluster_smem_need<real>(tab, 0, prior_table); flip = 0; }
                if (csmem > (size_t)h->smem_optin) continue; // rows + mailbox + exchange list of the fullest CTA do not fit: next size
                int ncl = 0;
                cudaError_t e = (ct > 0 && ct <= 1024) ? cluster_prepare<real>(tab, ct, csmem, &ncl) : cudaErrorInvalidConfiguration;
                if (e == cudaSuccess && ncl >= 1) {
                    out.CL = CL; out.threads = ct; out.smem = (int)csmem; out.ncl = ncl; out.flip = flip;
                    return BPOSD_OK;
                }
                cudaGetLastError();
            }
            return BPOSD_OK;
        };
        Plan pg, pu;
        int prc = plan_cluster(h->clus, 1, pg);
        if (prc != BPOSD_OK) return prc;
        bool done = false;
        if (pg.CL) {
            h->bp_kernel = 3; h->bp_threads = pg.threads; h->bp_smem = pg.smem; h->bp_ctas_per_sm = 1;
            h->clus_nclusters = pg.ncl; h->bp_grid = pg.ncl * pg.CL; h->clus_flip_table = pg.flip;
            kernel = 3;
            done = true;
            // uniform-prior launches: kept only if a smaller cluster does (the tables of a plan are a few MB of HBM)
            prc = plan_cluster(h->clus_u, 0, pu);
            if (prc != BPOSD_OK) return prc;
            if (pu.CL && pu.CL < pg.CL) {
                h->clus_u_threads = pu.threads; h->clus_u_smem = pu.smem; h->clus_u_nclusters = pu.ncl; h->clus_u_flip_table = pu.flip;
            } else {
                cluster_free(h->clus_u);
                h->clus_u_nclusters = 0;
            }
        }
        if (!done && (want == 3 || h->force_cluster > 0))
            return fail(h, BPOSD_EUNSUP, "the cluster BP kernel cannot be launched for this matrix / cluster size");
    } else if (want == 3) return fail(h, BPOSD_EUNSUP, "the cluster BP kernel needs min-sum and row/column degrees up to 16/8");
    int occ = 0;
    h->lat_geom = -1; h->lat_max_shots = 0;
    if (kernel == 3) {
        occ = 1;
    } else if (kernel == 2) {
        h->bp_geom = fast_geom(n, false);
        threads = std::min(threads, fast_maxt(n));
        if (threads * fast_vpt(n) < n) threads = fast_default_threads(n, m);
        CU_TRY(h, fast_set_smem_t<real>(h->fast, h->bp_geom, smem));
        CU_TRY(h, fast_occupancy_t<real>(h->fast, h->bp_geom, threads, smem, &occ));
        {
            const size_t smem_u = fast_smem_bytes<real>(h->fast, n, m, false);
            int occ_u = 0;
            CU_TRY(h, fast_occupancy_t<real>(h->fast, h->bp_geom, threads, smem_u, &occ_u));
            h->bp_smem_uni = (int)smem_u; h->bp_ctas_uni = std::max(occ_u, 1); h->bp_grid_uni = h->bp_ctas_uni * h->sm_count;
        }
        // latency geometry: one shot per SM, as many threads on it as the code has work for
        h->lat_geom = fast_geom(n, true);
        h->lat_threads = fast_default_threads_g(n, h->lat_geom, (int)rs);
        int occ_l = 0;
        if (fast_set_smem_t<real>(h->fast, h->lat_geom, smem) != cudaSuccess ||
            fast_occupancy_t<real>(h->fast, h->lat_geom, h->lat_threads, smem, &occ_l) != cudaSuccess || occ_l < 1) {
            cudaGetLastError();
            h->lat_geom = h->bp_geom; h->lat_threads = threads;
        }
        h->lat_max_shots = h->sm_count;
        if (h->bp_method != 1) { h->lat_geom = -1; h->lat_max_shots = 0; } // the latency geometry is instantiated for min-sum only
    } else if (kernel == 1) {
        CU_TRY(h, cudaFuncSetAttribute(bp_generic_kernel<real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bp_generic_kernel<real, true>, threads, smem));
    } else {
        CU_TRY(h, cudaFuncSetAttribute(bp_generic_kernel<real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bp_generic_kernel<real, false>, threads, smem));
    }
    if (occ < 1) return fail(h, BPOSD_EUNSUP, "BP kernel does not fit on an SM");
    if (kernel != 3) {
        h->bp_kernel = kernel;
        h->bp_threads = threads;
        h->bp_smem = (int)smem;
        h->bp_ctas_per_sm = occ;
        h->bp_grid = occ * h->sm_count;
    }
    if (kernel == 0) {
        const size_t per_cta = (2 * (size_t)E + n) * rs;
        const size_t need = per_cta * h->bp_grid;
        if (need > h->scratch_bytes) {
            cudaFree(h->d_scratch); cudaFree(h->d_scratch_dec);
            h->d_scratch = nullptr; h->d_scratch_dec = nullptr;
            CU_TRY(h, cudaMalloc(&h->d_scratch, need));
            CU_TRY(h, cudaMalloc((void **)&h->d_scratch_dec, (size_t)n * h->bp_grid));
            h->scratch_bytes = need;
        }
    }
    // OSD kernel
    h->osd_S = (m + 31) / 32;
    h->osd_St = h->osd_S | 1;
    // about one thread per two checks, at most 512: the elimination is a chain of short barrier-separated steps, so
    // cheap barriers matter more than lanes (measured on B200, profiles/r01y_osd_threads.log: m = 192 -> 128 threads
    // is 4.9x faster than 992 and 20 % faster than 192; m = 961 -> 512 threads is 5 % faster than 992)
    h->osd_threads = std::min(512, std::max(64, (m / 2 + 31) / 32 * 32));
    if (const char *ev = std::getenv("BPOSD_OSD_THREADS")) { // tuning experiments only
        const int v = std::atoi(ev);
        if (v >= 64 && v <= 1024 && v % 32 == 0) h->osd_threads = v;
    }
    const int nw = h->osd_threads / 32;
    const size_t osd_smem = (size_t)n * 8 + 256 + ((size_t)m * h->osd_St + 3 * (size_t)h->osd_S + (size_t)nw * (h->osd_S + 64)) * 4 + 128 + 3 * (size_t)n * 2 + 16;
    h->osd_smem = (int)osd_smem;
    h->osd_supported = osd_smem <= (size_t)h->smem_optin && n < 65535 && m < 65535;
    // HBM-resident OSD-0: used when T does not fit (or when forced), only for search depth 0
    h->osdl_npanels = (n + 31) / 32;
    h->osdl_smem = (int)(8 * (size_t)m + 4096 + 256 + 4 * ((size_t)h->osdl_npanels + 2) + 2 * ((size_t)m + 2) + (size_t)m + 16);
    const bool large_ok = h->osd_order == 0 && (size_t)h->osdl_smem <= (size_t)h->smem_optin && m < 65535;
    h->osd_large = large_ok && (h->osd_variant == 2 || (h->osd_variant == 0 && !h->osd_supported));
    if (h->osd_variant == 2 && !large_ok)
        return fail(h, BPOSD_EUNSUP, "the HBM-resident OSD kernel handles OSD-0 (osd_order 0) with m < 65535 only");
    if (h->osd_large) {
        CU_TRY(h, cudaFuncSetAttribute(osd0_large_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->osdl_smem));
        const size_t per_cta = (size_t)h->osdl_npanels * m * 4 + (size_t)n * 4 + 2 * (size_t)std::min(m, n) * 4 + ((size_t)h->osdl_npanels + 1) * 4;
        h->osdl_grid = (int)std::max<long long>(1, std::min<long long>(h->sm_count, h->osdl_ws_cap / (long long)std::max<size_t>(per_cta, 1)));
        h->osd_supported = true;
    }
    // cluster OSD-0 (variant 4): preferred over the single-CTA HBM kernel whenever T does not fit in shared memory
    h->osd_clus = false;
    h->osdc.clear();
    if (h->osd_order == 0 && m >= 2 && (h->osd_variant == 4 || (h->osd_variant == 0 && (h->osd_large || !h->osd_supported)))) {
        const int CL0 = m >= 4096 ? 16 : (m >= 1024 ? 8 : (m >= 256 ? 4 : 2));
        h->osdc_npanels = (n + 63) / 64;
        auto kern = osd0_cluster_kernel<real>;
        int smem_max = 0;
        for (int CL = CL0; CL >= 2 && CL >= CL0 / 4; CL >>= 1) {
            bposd_handle::OsdcCfg c;
            c.CL = CL;
            c.rpc = (((m + CL - 1) / CL) + 1) & ~1;
            c.ncl = 0;
            c.stages = kOsdcMaxStages; // as deep a TMA ring as fits
            while (c.stages > 2 && osdc_layout(c.rpc, h->osdc_npanels, c.stages).total > (size_t)h->smem_optin) c.stages--;
            c.smem = (int)osdc_layout(c.rpc, h->osdc_npanels, c.stages).total;
            if ((size_t)c.smem > (size_t)h->smem_optin) continue;
            smem_max = std::max(smem_max, c.smem);
            h->osdc.push_back(c);
        }
        if (!h->osdc.empty()) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
            if (e == cudaSuccess && CL0 > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) { cudaGetLastError(); h->osdc.clear(); }
        }
        for (size_t k = 0; k < h->osdc.size();) {
            bposd_handle::OsdcCfg &c = h->osdc[k];
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(c.CL, 1, 1); cfg.blockDim = dim3(kOsdcThreads, 1, 1); cfg.dynamicSmemBytes = c.smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c.CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
            const size_t per_cluster = (size_t)h->osdc_npanels * c.rpc * c.CL * 8 + (size_t)osd_reg_np2(n) * 12 + (size_t)std::min(m, n) * sizeof(OsdcPivot);
            if (e != cudaSuccess || ncl < 1) { cudaGetLastError(); h->osdc.erase(h->osdc.begin() + k); continue; }
            c.ncl = (int)std::max<long long>(1, std::min<long long>(ncl, h->osdl_ws_cap / (long long)std::max<size_t>(per_cluster, 1)));
            k++;
        }
        if (!h->osdc.empty()) { h->osd_clus = true; h->osd_large = false; h->osd_supported = true; }
        else if (h->osd_variant == 4) return fail(h, BPOSD_EUNSUP, "the cluster OSD-0 kernel cannot be launched for this matrix");
    } else if (h->osd_variant == 4)
        return fail(h, BPOSD_EUNSUP, "the cluster OSD kernel handles OSD-0 (osd_order 0) only");
    // register kernel (default whenever it applies): T in registers, one barrier per 16 sorted columns
    {
        const int Sw = (m + 31) / 32;
        h->osdr_W = Sw <= 4 ? 4 : (Sw <= 8 ? 8 : (Sw <= 16 ? 16 : 32));
        h->osdr_KD = h->max_col_deg <= 4 ? 4 : 8;
        h->osdr_threads = osd_reg_threads(m);
        const size_t rs_ = osd_reg_layout(m, n, E, h->osdr_W, h->osdr_KD, kOsdRegG, h->osdr_threads, osd_reg_np2(n)).total;
        h->osdr_smem = (int)std::min<size_t>(rs_, (size_t)1 << 30);
        const bool reg_ok = m >= 1 && m <= 1024 && n < 65535 && E < 65536 && h->max_col_deg <= 8 && rs_ <= (size_t)h->smem_optin;
        if (h->osd_variant == 3 && !reg_ok)
            return fail(h, BPOSD_EUNSUP, "the register OSD kernel needs m <= 1024, n < 65535, fewer than 65536 edges and column degrees up to 8");
        h->osd_reg = reg_ok && (h->osd_variant == 3 || h->osd_variant == 0);
        if (h->osd_reg) {
            h->osd_large = false; h->osd_clus = false;
            h->osd_supported = true;
            int occ4 = 0;
#define BPOSD_OSDR_SETUP(Wv, KDv)                                                                                                         \
    do {                                                                                                                                  \
        CU_TRY(h, cudaFuncSetAttribute(osd_reg_kernel<real, Wv, KDv>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->osdr_smem));         \
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, osd_reg_kernel<real, Wv, KDv>, h->osdr_threads, h->osdr_smem));      \
    } while (0)
#define BPOSD_OSDR_W(KDv)                                                                        \
    switch (h->osdr_W) {                                                                         \
        case 4: BPOSD_OSDR_SETUP(4, KDv); break;                                                 \
        case 8: BPOSD_OSDR_SETUP(8, KDv); break;                                                 \
        case 16: BPOSD_OSDR_SETUP(16, KDv); break;                                               \
        default: BPOSD_OSDR_SETUP(32, KDv); break;                                               \
    }
            if (h->osdr_KD == 4) { BPOSD_OSDR_W(4) } else { BPOSD_OSDR_W(8) }
#undef BPOSD_OSDR_W
#undef BPOSD_OSDR_SETUP
            h->osd_ctas_per_sm = std::max(1, occ4);
        }
    }
    if (h->osd_supported && !h->osd_large && !h->osd_reg && !h->osd_clus) {
        CU_TRY(h, cudaFuncSetAttribute(osd_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osd_smem));
        int occ2 = 0;
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, osd_kernel<real>, h->osd_threads, osd_smem));
        h->osd_ctas_per_sm = std::max(1, occ2);
    }
    // failed-shot LLR workspace capacity (shots per chunk).  Large codes (config 5: 320 KB of LLRs per shot) get a bigger
    // workspace unless the caller set one: their OSD launch is latency bound (~0.3 s for up to a cluster-load of failed
    // shots), so it should see the failed shots of as many decoded shots as memory allows
    const long long per_shot = (long long)n * (long long)rs;
    if (!h->workspace_user_set && per_shot >= (64 << 10)) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            h->workspace_bytes = std::max<long long>(2ll << 30, std::min<long long>(16ll << 30, (long long)(free_b / 4)));
        else cudaGetLastError();
    }
    h->fail_cap = std::max<long long>(1024, h->workspace_bytes / std::max<long long>(per_shot, 1));
    h->geometry_ready = true;
    return BPOSD_OK;
}

static int plan_geometry(bposd_handle *h) {
    return h->precision == 64 ? plan_geometry_t<double>(h) : plan_geometry_t<float>(h);
}

extern "C" const char *bposd_version(void) { return "bposd_b200 0.1 (sm_100a)"; }

extern "C" const char *bposd_last_error(const bposd_t *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" void bposd_destroy(bposd_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_row_ptr); cudaFree(h->d_col_idx); cudaFree(h->d_col_ptr); cudaFree(h->d_row_idx); cudaFree(h->d_csc_slot);
    cudaFree(h->d_prior64); cudaFree(h->d_prior32); cudaFree(h->d_weight);
    fast_free(h->fast);
    cluster_free(h->clus);
    cluster_free(h->clus_u);
    cudaFree(h->d_order);
    for (auto &sl : h->slot) {
        cudaFree(sl.d_ctrl); cudaFree(sl.d_fail_list); cudaFree(sl.d_fail_llr); cudaFree(sl.d_serial_ws);
        if (sl.h_ctrl) cudaFreeHost(sl.h_ctrl);
        cudaFree(sl.b_synd); cudaFree(sl.b_err); cudaFree(sl.b_osdw); cudaFree(sl.b_osd0); cudaFree(sl.b_bp);
        cudaFree(sl.b_conv); cudaFree(sl.b_llr); cudaFree(sl.b_iter);
        cudaFree(sl.b_pk_osdw); cudaFree(sl.b_pk_osd0); cudaFree(sl.b_pk_bp);
        for (auto &e : sl.ev) if (e) cudaEventDestroy(e);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    cudaFree(h->d_scratch); cudaFree(h->d_scratch_dec);
    cudaFree(h->d_osdl_mask); cudaFree(h->d_osdl_order); cudaFree(h->d_osdl_piv_row); cudaFree(h->d_osdl_piv_pos); cudaFree(h->d_osdl_pstart);
    cudaFree(h->d_osdc_mask); cudaFree(h->d_osdc_key); cudaFree(h->d_osdc_idx); cudaFree(h->d_osdc_piv);
    cudaFree(h->d_t1); cudaFree(h->d_t2); cudaFree(h->d_t3); cudaFree(h->d_l_ptr); cudaFree(h->d_l_idx);
    cudaFree(h->d_counters); cudaFree(h->d_minw); cudaFree(h->d_cu_tab);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    delete h;
}

extern "C" int bposd_create(const int32_t *indptr, const int32_t *indices, int32_t m, int32_t n,
                            const double *probs, int32_t max_iter, int32_t bp_method, double alpha,
                            int32_t osd_method, int32_t osd_order, int32_t precision, int32_t device,
                            bposd_t **out) {
    g_create_err.clear();
    auto bad = [&](int code, const std::string &msg) { g_create_err = msg; return code; };
    if (!out) return bad(BPOSD_EINVAL, "out is NULL");
    *out = nullptr;
    if (!indptr || (!indices && m > 0 && indptr[m] > 0) || !probs) return bad(BPOSD_EINVAL, "NULL input pointer");
    if (m < 0 || n <= 0) return bad(BPOSD_EINVAL, "parity-check matrix must have at least one column");
    if (bp_method != BPOSD_BP_PRODUCT_SUM && bp_method != BPOSD_BP_MINIMUM_SUM) return bad(BPOSD_EINVAL, "unknown bp_method");
    if (osd_method < BPOSD_OSD_0 || osd_method > BPOSD_OSD_OFF) return bad(BPOSD_EINVAL, "unknown osd_method");
    if (precision != 64 && precision != 32) return bad(BPOSD_EINVAL, "precision must be 64 or 32");
    if (max_iter < 0) return bad(BPOSD_EINVAL, "max_iter must be non-negative");
    if (osd_order < 0) return bad(BPOSD_EINVAL, "osd_order must be non-negative");
    for (int i = 0; i < m; i++) {
        if (indptr[i + 1] < indptr[i]) return bad(BPOSD_EINVAL, "indptr must be non-decreasing");
        for (int e = indptr[i]; e < indptr[i + 1]; e++) {
            if (indices[e] < 0 || indices[e] >= n) return bad(BPOSD_EINVAL, "column index out of range");
            if (e > indptr[i] && indices[e] <= indices[e - 1]) return bad(BPOSD_EINVAL, "column indices must be strictly ascending inside a row");
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return bad(BPOSD_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return bad(BPOSD_EINVAL, "device index out of range");

    bposd_handle *h = new (std::nothrow) bposd_handle();
    if (!h) return bad(BPOSD_ENOMEM, "out of host memory");
    h->device = device;
    auto die = [&](int code) { g_create_err = h->err; bposd_destroy(h); return code; };
#define CR_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e__); return die(BPOSD_ECUDA); } } while (0)
    CR_TRY(cudaSetDevice(device));
    CR_TRY(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    CR_TRY(cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    h->m = m; h->n = n; h->nnz = indptr[m];
    h->row_ptr.assign(indptr, indptr + m + 1);
    h->col_idx.assign(indices, indices + h->nnz);
    h->col_ptr.assign(n + 1, 0);
    for (int e = 0; e < h->nnz; e++) h->col_ptr[h->col_idx[e] + 1]++;
    for (int j = 0; j < n; j++) { h->max_col_deg = std::max(h->max_col_deg, h->col_ptr[j + 1]); h->col_ptr[j + 1] += h->col_ptr[j]; }
    h->row_idx.resize(h->nnz); h->csc_slot.resize(h->nnz);
    {
        std::vector<int> fill(n, 0);
        for (int i = 0; i < m; i++) {
            h->max_row_deg = std::max(h->max_row_deg, h->row_ptr[i + 1] - h->row_ptr[i]);
            for (int e = h->row_ptr[i]; e < h->row_ptr[i + 1]; e++) {
                const int j = h->col_idx[e], p = h->col_ptr[j] + fill[j]++;
                h->row_idx[p] = i;
                h->csc_slot[p] = e;
            }
        }
    }
    h->probs.assign(probs, probs + n);
    h->max_iter = max_iter > 0 ? max_iter : n;
    h->bp_method = bp_method;
    h->alpha = alpha;
    h->osd_method = osd_method;
    h->osd_order = (osd_method == BPOSD_OSD_0 || osd_method == BPOSD_OSD_OFF) ? 0 : osd_order;
    h->precision = precision;
    h->rank = host_rank(h);
    h->k = n - h->rank;
    if (h->osd_order > h->k) { h->err = "osd_order must not exceed n - rank(H)"; return die(BPOSD_EINVAL); }
    if (osd_method == BPOSD_OSD_E && h->osd_order > 30) { h->err = "osd_e order above 30 is not supported"; return die(BPOSD_EINVAL); }
    if (h->osd_order > 63) { h->err = "osd_order above 63 is not supported"; return die(BPOSD_EINVAL); }

    CR_TRY(upload(&h->d_row_ptr, h->row_ptr));
    CR_TRY(upload(&h->d_col_idx, h->col_idx));
    CR_TRY(upload(&h->d_col_ptr, h->col_ptr));
    CR_TRY(upload(&h->d_row_idx, h->row_idx));
    CR_TRY(upload(&h->d_csc_slot, h->csc_slot));
    CR_TRY(cudaMalloc((void **)&h->d_prior64, n * sizeof(double)));
    CR_TRY(cudaMalloc((void **)&h->d_prior32, n * sizeof(float)));
    CR_TRY(cudaMalloc((void **)&h->d_weight, n * sizeof(double)));
    if (upload_probs(h) != BPOSD_OK) return die(BPOSD_ECUDA);
    for (auto &sl : h->slot) {
        CR_TRY(cudaMalloc((void **)&sl.d_ctrl, 8 * sizeof(unsigned long long)));
        CR_TRY(cudaMemset(sl.d_ctrl, 0, 8 * sizeof(unsigned long long)));
        CR_TRY(cudaMallocHost((void **)&sl.h_ctrl, 8 * sizeof(unsigned long long)));
        for (auto &e : sl.ev) CR_TRY(cudaEventCreate(&e));
        CR_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    }
    CR_TRY(cudaMalloc((void **)&h->d_counters, 8 * sizeof(unsigned long long)));
    CR_TRY(cudaMalloc((void **)&h->d_minw, sizeof(int)));
    if (fast_supported(h->max_col_deg, h->max_row_deg, bp_method)) {
        cudaError_t e = fast_build(h->fast, m, n, h->row_ptr, h->col_idx, h->col_ptr, h->row_idx, h->csc_slot, precision / 8, bp_method);
        if (e != cudaSuccess) { h->err = std::string("fast_build: ") + cudaGetErrorString(e); return die(BPOSD_ECUDA); }
    }
    int rc = plan_geometry(h);
    if (rc != BPOSD_OK) return die(rc);
#undef CR_TRY
    *out = h;
    return BPOSD_OK;
}

extern "C" int bposd_update_channel_probs(bposd_t *h, const double *probs) {
    if (!h || !probs) return fail(h, BPOSD_EINVAL, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    h->probs.assign(probs, probs + h->n);
    return upload_probs(h);
}

extern "C" int bposd_set_tuning(bposd_t *h, int32_t kernel_plus1, int32_t threads, int64_t workspace_bytes) {
    if (!h) return BPOSD_EINVAL;
    if (kernel_plus1 < 0 || kernel_plus1 > 4) return fail(h, BPOSD_EINVAL, "bp kernel selector out of range");
    if (threads < 0 || threads > 1024 || (threads % 32)) return fail(h, BPOSD_EINVAL, "threads must be a multiple of 32 up to 1024");
    CU_TRY(h, cudaSetDevice(h->device));
    h->force_kernel = kernel_plus1;
    h->force_threads = threads;
    if (workspace_bytes > 0) { h->workspace_bytes = workspace_bytes; h->workspace_user_set = true; }
    return plan_geometry(h);
}

extern "C" int bposd_int32_peak(bposd_t *h, double *ops_per_s) {
    if (!h || !ops_per_s) return BPOSD_EINVAL;
    CU_TRY(h, cudaSetDevice(h->device));
    uint32_t *d_out = nullptr;
    CU_TRY(h, cudaMalloc((void **)&d_out, 4));
    const int iters = 1 << 16, grid = h->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    CU_TRY(h, cudaEventCreate(&e0));
    CU_TRY(h, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) { // first repetition warms up
        CU_TRY(h, cudaEventRecord(e0, nullptr));
        lop3_peak_kernel<<<grid, threads>>>(d_out, iters);
        CU_TRY(h, cudaEventRecord(e1, nullptr));
        CU_TRY(h, cudaEventSynchronize(e1));
        float ms = 0;
        CU_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *ops_per_s = 8.0 * iters * (double)grid * threads / (best * 1e-3);
    return BPOSD_OK;
}

extern "C" int bposd_fp64_peak(bposd_t *h, double *fma_per_s) {
    if (!h || !fma_per_s) return BPOSD_EINVAL;
    CU_TRY(h, cudaSetDevice(h->device));
    double *d_out = nullptr;
    CU_TRY(h, cudaMalloc((void **)&d_out, 8));
    const int iters = 1 << 14, grid = h->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    CU_TRY(h, cudaEventCreate(&e0));
    CU_TRY(h, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) { // first repetition warms up
        CU_TRY(h, cudaEventRecord(e0, nullptr));
        dfma_peak_kernel<<<grid, threads>>>(d_out, iters);
        CU_TRY(h, cudaEventRecord(e1, nullptr));
        CU_TRY(h, cudaEventSynchronize(e1));
        float ms = 0;
        CU_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *fma_per_s = 8.0 * iters * (double)grid * threads / (best * 1e-3);
    return BPOSD_OK;
}

extern "C" int bposd_smem_peak(bposd_t *h, double *bytes_per_s) {
    if (!h || !bytes_per_s) return BPOSD_EINVAL;
    CU_TRY(h, cudaSetDevice(h->device));
    float *d_out = nullptr;
    CU_TRY(h, cudaMalloc((void **)&d_out, 4));
    const int iters = 1 << 14, threads = 512, per_sm = 2, grid = h->sm_count * per_sm;
    const size_t smem = (size_t)threads * 16 * 4;
    CU_TRY(h, cudaFuncSetAttribute(smem_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CU_TRY(h, cudaEventCreate(&e0));
    CU_TRY(h, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) { // first repetition warms up
        CU_TRY(h, cudaEventRecord(e0, nullptr));
        smem_peak_kernel<<<grid, threads, smem>>>(d_out, iters);
        CU_TRY(h, cudaEventRecord(e1, nullptr));
        CU_TRY(h, cudaEventSynchronize(e1));
        float ms = 0;
        CU_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    // per trip and thread: four 16-byte loads + four 16-byte stores
    *bytes_per_s = 128.0 * iters * (double)grid * threads / (best * 1e-3);
    return BPOSD_OK;
}

// fn: 0 a / b by bpm_div, 1 tanh(a), 2 log(a), 3 (1 + a) / (1 - a) by ps_ratio -- the device side of include/bposd_math.h,
// evaluated element-wise so that tests can compare it bit for bit with the host side of the same header.
__global__ void math_probe_kernel(int fn, const double *a, const double *b, double *out, long long count) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const double x = a[i];
        out[i] = fn == 0 ? bpm_div(x, b[i]) : fn == 1 ? r_tanh(x) : fn == 2 ? r_log(x) : ps_ratio(x);
    }
}

extern "C" int bposd_math_probe(bposd_t *h, int32_t fn, const double *a, const double *b, double *out, int64_t count) {
    if (!h || !a || !out || count < 0 || fn < 0 || fn > 3 || (fn == 0 && !b)) return BPOSD_EINVAL;
    if (count == 0) return BPOSD_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    math_probe_kernel<<<h->sm_count * 8, 256>>>(fn, a, b, out, (long long)count);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaDeviceSynchronize());
    return BPOSD_OK;
}

extern "C" int bposd_set_schedule(bposd_t *h, int32_t schedule, const int32_t *order) {
    if (!h) return BPOSD_EINVAL;
    if (schedule != 0 && schedule != 1) return fail(h, BPOSD_EINVAL, "schedule must be 0 (parallel) or 1 (serial)");
    CU_TRY(h, cudaSetDevice(h->device));
    if (order) {
        std::vector<char> seen((size_t)h->n, 0);
        for (int k = 0; k < h->n; k++) {
            if (order[k] < 0 || order[k] >= h->n || seen[order[k]]) return fail(h, BPOSD_EINVAL, "serial_schedule_order must be a permutation of 0 .. n-1");
            seen[order[k]] = 1;
        }
    }
    CU_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->d_order);
    h->d_order = nullptr;
    if (schedule == 1 && order && h->n > 0) {
        CU_TRY(h, cudaMalloc((void **)&h->d_order, (size_t)h->n * sizeof(int)));
        CU_TRY(h, cudaMemcpy(h->d_order, order, (size_t)h->n * sizeof(int), cudaMemcpyHostToDevice));
    }
    h->schedule = schedule;
    return BPOSD_OK;
}

extern "C" int bposd_set_cluster_size(bposd_t *h, int32_t cluster_size) {
    if (!h) return BPOSD_EINVAL;
    if (cluster_size != 0 && cluster_size != 2 && cluster_size != 4 && cluster_size != 8 && cluster_size != 16)
        return fail(h, BPOSD_EINVAL, "cluster size must be 0 (automatic), 2, 4, 8 or 16");
    CU_TRY(h, cudaSetDevice(h->device));
    h->force_cluster = cluster_size;
    return plan_geometry(h);
}

extern "C" int bposd_set_osd_variant(bposd_t *h, int32_t variant, int64_t workspace_bytes) {
    if (!h) return BPOSD_EINVAL;
    if (variant < 0 || variant > 4) return fail(h, BPOSD_EINVAL, "osd variant must be 0 (auto), 1 (T matrix in shared memory), 2 (HBM resident, one CTA per shot), 3 (T matrix in registers) or 4 (HBM resident, one cluster per shot)");
    CU_TRY(h, cudaSetDevice(h->device));
    const int old = h->osd_variant;
    h->osd_variant = variant;
    if (workspace_bytes > 0) h->osdl_ws_cap = workspace_bytes;
    int rc = plan_geometry(h);
    if (rc != BPOSD_OK) { h->osd_variant = old; std::string keep = h->err; plan_geometry(h); h->err = keep; }
    return rc;
}

extern "C" int bposd_get_info(const bposd_t *h, bposd_info_t *info) {
    if (!h || !info) return BPOSD_EINVAL;
    info->m = h->m; info->n = h->n; info->nnz = h->nnz; info->rank = h->rank; info->k = h->k;
    info->max_iter = h->max_iter; info->bp_method = h->bp_method; info->osd_method = h->osd_method;
    info->osd_order = h->osd_order; info->precision = h->precision; info->device = h->device;
    info->bp_kernel = h->bp_kernel; info->bp_threads = h->bp_threads; info->bp_ctas_per_sm = h->bp_ctas_per_sm;
    info->bp_smem_bytes = h->bp_smem;
    if (h->bp_kernel == 2 && h->uniform_prior) { info->bp_ctas_per_sm = h->bp_ctas_uni; info->bp_smem_bytes = h->bp_smem_uni; } info->osd_threads = h->osd_threads; info->osd_smem_bytes = h->osd_smem;
    info->sm_count = h->sm_count; info->ms_scaling_factor = h->alpha;
    info->osd_variant = !h->osd_supported ? 0 : (h->osd_clus ? 4 : (h->osd_large ? 2 : (h->osd_reg ? 3 : 1)));
    if (h->osd_clus) { info->osd_threads = kOsdcThreads; info->osd_smem_bytes = h->osdc[0].smem; }
    if (h->osd_reg) { info->osd_threads = h->osdr_threads; info->osd_smem_bytes = h->osdr_smem; }
    const bool plan_u = h->bp_kernel == 3 && h->uniform_prior && h->clus_u.CL > 0; // what a launch with the static channel uses
    const ClusterTables &ct = plan_u ? h->clus_u : h->clus;
    if (plan_u) { info->bp_threads = h->clus_u_threads; info->bp_smem_bytes = h->clus_u_smem; }
    info->bp_layout_excess = h->bp_kernel == 3 ? (int32_t)(1000 * ct.remote_edges / std::max<long long>(ct.total_edges, 1))
                                               : (int32_t)h->fast.conflicts_after;
    info->bp_cluster_size = h->bp_kernel == 3 ? ct.CL : 1;
    return BPOSD_OK;
}

extern "C" int bposd_get_stats(const bposd_t *h, bposd_stats_t *stats) {
    if (!h || !stats) return BPOSD_EINVAL;
    *stats = h->stats;
    return BPOSD_OK;
}

// OSD on the shots of a chunk that BP did not settle: `d_fail_list[0 .. *d_fail_count)` are their indices into the
// chunk; `llr` is indexed by shot (llr_by_shot) or by position in the list.  All pointers are device-visible.
template <typename real>
static int launch_osd(bposd_handle *h, cudaStream_t st, const GraphDev &g, const uint8_t *d_synd, int synd_packed, const real *llr, int llr_by_shot,
                      const int *d_fail_count, const int *d_fail_list, uint8_t *d_osd0, uint8_t *d_osdw,
                      unsigned long long *d_stat, long long Bc, bool per_shot_priors, const double *d_weights, int *launches) {
    const int n = h->n, m = h->m;
    if (h->osd_clus) {
        const int np2 = osd_reg_np2(n);
        size_t need_mask = 0, need_shots = 0;
        for (const auto &c : h->osdc) {
            const size_t ncl = (size_t)c.ncl; // for every resident cluster, once (no re-allocation when the batch size changes)
            need_mask = std::max(need_mask, ncl * h->osdc_npanels * c.rpc * c.CL);
            need_shots = std::max(need_shots, ncl);
        }
        if (h->osdc_alloc_mask < need_mask || h->osdc_alloc_shots < need_shots) {
            cudaFree(h->d_osdc_mask); cudaFree(h->d_osdc_key); cudaFree(h->d_osdc_idx); cudaFree(h->d_osdc_piv);
            h->d_osdc_mask = h->d_osdc_key = nullptr; h->d_osdc_idx = nullptr; h->d_osdc_piv = nullptr;
            h->osdc_alloc_mask = h->osdc_alloc_shots = 0;
            CU_TRY(h, cudaMalloc((void **)&h->d_osdc_mask, need_mask * 8));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdc_key, need_shots * np2 * 8));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdc_idx, need_shots * np2 * 4));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdc_piv, need_shots * std::max<size_t>(std::min(m, n), 1) * sizeof(OsdcPivot)));
            h->osdc_alloc_mask = need_mask; h->osdc_alloc_shots = need_shots;
        }
        // one launch per cluster size; each takes the chunk only if the failed-shot count (known on the device) is in its range
        int lo = 0;
        for (size_t k = 0; k < h->osdc.size(); k++) {
            const auto &c = h->osdc[k];
            const int ncl = (int)std::min<long long>(Bc, c.ncl);
            const bool lastcfg = k + 1 == h->osdc.size();
            const int hi = lastcfg ? 0x7fffffff : c.ncl; // up to one shot per resident cluster: the big clusters; beyond: smaller ones
            if (!lastcfg && lo > hi) continue;
            OsdClusterArgs<real> o;
            o.g = g;
            o.synd = d_synd; o.synd_packed = synd_packed;
            o.llr = llr; o.llr_by_shot = llr_by_shot;
            o.fail_count = d_fail_count; o.fail_list = d_fail_list;
            o.osd0 = d_osd0; o.osdw = d_osdw;
            o.stat = d_stat;
            o.maxrank = h->rank; o.npanels = h->osdc_npanels; o.CL = c.CL; o.rpc = c.rpc; o.np2 = np2;
            o.nfail_lo = lo; o.nfail_hi = hi; o.stages = c.stages;
            o.ws_mask = h->d_osdc_mask; o.ws_key = h->d_osdc_key; o.ws_idx = h->d_osdc_idx; o.ws_piv = h->d_osdc_piv;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ncl * c.CL, 1, 1); cfg.blockDim = dim3(kOsdcThreads, 1, 1); cfg.dynamicSmemBytes = c.smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c.CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CU_TRY(h, cudaLaunchKernelEx(&cfg, osd0_cluster_kernel<real>, o));
            (*launches)++;
            lo = hi == 0x7fffffff ? hi : hi + 1;
        }
        return BPOSD_OK;
    }
    if (h->osd_large) {
        const int ogrid = (int)std::min<long long>(Bc, h->osdl_grid);
        if (h->osdl_alloc_grid < ogrid) {
            cudaFree(h->d_osdl_mask); cudaFree(h->d_osdl_order); cudaFree(h->d_osdl_piv_row); cudaFree(h->d_osdl_piv_pos); cudaFree(h->d_osdl_pstart);
            h->d_osdl_mask = nullptr; h->d_osdl_order = h->d_osdl_piv_row = h->d_osdl_piv_pos = h->d_osdl_pstart = nullptr;
            h->osdl_alloc_grid = 0;
            const size_t gsz = (size_t)ogrid, mn = (size_t)std::min(m, n);
            CU_TRY(h, cudaMalloc((void **)&h->d_osdl_mask, gsz * h->osdl_npanels * m * 4));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdl_order, gsz * n * 4));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdl_piv_row, gsz * std::max<size_t>(mn, 1) * 4));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdl_piv_pos, gsz * std::max<size_t>(mn, 1) * 4));
            CU_TRY(h, cudaMalloc((void **)&h->d_osdl_pstart, gsz * ((size_t)h->osdl_npanels + 1) * 4));
            h->osdl_alloc_grid = ogrid;
        }
        OsdLargeArgs<real> o;
        o.g = g;
        o.synd = d_synd;
        o.synd_packed = synd_packed;
        o.llr = llr;
        o.llr_by_shot = llr_by_shot;
        o.fail_count = d_fail_count;
        o.fail_list = d_fail_list;
        o.osd0 = d_osd0; o.osdw = d_osdw;
        o.stat = d_stat;
        o.maxrank = h->rank;
        o.npanels = h->osdl_npanels;
        o.ws_mask = h->d_osdl_mask; o.ws_order = h->d_osdl_order;
        o.ws_piv_row = h->d_osdl_piv_row; o.ws_piv_pos = h->d_osdl_piv_pos; o.ws_pstart = h->d_osdl_pstart;
        osd0_large_kernel<real><<<ogrid, 1024, h->osdl_smem, st>>>(o);
        CU_TRY(h, cudaGetLastError());
        (*launches)++;
    } else {
        OsdArgs<real> o;
        o.g = g;
        o.S = h->osd_S; o.St = h->osd_St;
        o.maxrank = h->rank; o.maxdeg = std::max(h->max_col_deg, 1); o.np2 = osd_reg_np2(n);
        o.method = h->osd_method; o.order = h->osd_order;
        o.uniform = (per_shot_priors || d_weights) ? 0 : h->uniform;
        o.weight = d_weights ? d_weights : h->d_weight;
        o.weight_stride = d_weights ? n : 0;
        o.synd = d_synd;
        o.synd_packed = synd_packed;
        o.llr = llr;
        o.llr_by_shot = llr_by_shot;
        o.fail_count = d_fail_count;
        o.fail_list = d_fail_list;
        o.osd0 = d_osd0; o.osdw = d_osdw;
        o.stat = d_stat;
        const int ogrid = (int)std::min<long long>(Bc, (long long)h->osd_ctas_per_sm * h->sm_count);
        if (h->osd_reg) {
#define BPOSD_OSDR_LAUNCH(Wv, KDv) osd_reg_kernel<real, Wv, KDv><<<ogrid, h->osdr_threads, h->osdr_smem, st>>>(o)
#define BPOSD_OSDR_W(KDv)                                                                        \
    switch (h->osdr_W) {                                                                         \
        case 4: BPOSD_OSDR_LAUNCH(4, KDv); break;                                                \
        case 8: BPOSD_OSDR_LAUNCH(8, KDv); break;                                                \
        case 16: BPOSD_OSDR_LAUNCH(16, KDv); break;                                              \
        default: BPOSD_OSDR_LAUNCH(32, KDv); break;                                              \
    }
            if (h->osdr_KD == 4) { BPOSD_OSDR_W(4) } else { BPOSD_OSDR_W(8) }
#undef BPOSD_OSDR_W
#undef BPOSD_OSDR_LAUNCH
        } else
        osd_kernel<real><<<ogrid, h->osd_threads, h->osd_smem, st>>>(o);
        CU_TRY(h, cudaGetLastError());
        (*launches)++;
    }
    return BPOSD_OK;
}

// Enqueue one chunk (BP -> OSD -> control-word read-back) on `st`.  No host synchronisation.
template <typename real>
static int launch_chunk(bposd_handle *h, bposd_handle::Slot &sl, cudaStream_t st, const uint8_t *d_synd, long long Bc,
                        const bposd_out_t &out, const void *d_priors, const double *d_weights, int synd_packed = 0) {
    const int n = h->n, m = h->m;
    const bool osd_on = h->osd_method != BPOSD_OSD_OFF;
    real *llr_out = static_cast<real *>(out.d_llr);
    const bool need_ws = osd_on && !llr_out;
    if (need_ws && (!sl.d_fail_llr || Bc > sl.fail_llr_cap)) {
        cudaFree(sl.d_fail_llr);
        sl.d_fail_llr = nullptr; sl.fail_llr_cap = 0;
        const long long cap = std::max(Bc, (long long)1024);
        CU_TRY(h, cudaMalloc(&sl.d_fail_llr, (size_t)cap * n * sizeof(real)));
        sl.fail_llr_cap = cap;
    }
    if (!sl.d_fail_list || Bc > sl.fail_list_cap) {
        cudaFree(sl.d_fail_list);
        sl.d_fail_list = nullptr; sl.fail_list_cap = 0;
        CU_TRY(h, cudaMalloc((void **)&sl.d_fail_list, (size_t)Bc * sizeof(int)));
        sl.fail_list_cap = Bc;
    }
    int *d_fail_count = reinterpret_cast<int *>(sl.d_ctrl + 4);
    CU_TRY(h, cudaMemsetAsync(sl.d_ctrl, 0, 8 * sizeof(unsigned long long), st));
    int launches = 0;
    BpArgs<real> a;
    a.g = graph_of(h);
    a.max_iter = h->max_iter;
    a.method = h->bp_method;
    a.alpha0 = (real)h->alpha;
    a.safe_it = sizeof(real) == 4 ? 0x7fffffff : (d_priors ? 0 : h->safe_it);
    a.uniform_prior = d_priors ? 0 : h->uniform_prior;
    if (d_priors) { a.prior = static_cast<const real *>(d_priors); a.prior_stride = n; }
    else { a.prior = (sizeof(real) == 8) ? (const real *)h->d_prior64 : (const real *)h->d_prior32; a.prior_stride = 0; }
    a.synd = d_synd;
    a.synd_packed = synd_packed;
    a.B = Bc;
    a.bp = out.d_bp; a.osd0 = out.d_osd0; a.osdw = out.d_osdw;
    a.llr = llr_out;
    a.converge = out.d_converge; a.iter = out.d_iter;
    a.fail_count = d_fail_count;
    a.fail_list = sl.d_fail_list;
    a.fail_llr = static_cast<real *>(sl.d_fail_llr);
    a.osd_off = osd_on ? 0 : 1;
    a.queue = sl.d_ctrl;
    a.stat = sl.d_ctrl + 1;
    a.g_scratch = static_cast<real *>(h->d_scratch);
    a.g_dec = h->d_scratch_dec;
    const int grid = (int)std::min<long long>(Bc, h->bp_grid);
    CU_TRY(h, cudaEventRecord(sl.ev[0], st));
    if (h->schedule == 1) {
        // serial schedule: one thread per shot, a warp's 32 shots in lockstep, messages in an HBM workspace per warp
        const size_t wb = serial_warp_bytes<real>(m, n, h->nnz);
        const long long want_warps = std::min<long long>((Bc + 31) / 32, (long long)h->sm_count * 16);
        const long long fit_warps = std::max<long long>(1, (long long)(((size_t)4 << 30) / wb)); // at most 4 GiB of workspace
        const int wpb = 4; // warps per CTA; the grid is whole CTAs and every warp of it owns a workspace
        const int warps = ((int)std::min(want_warps, fit_warps) + wpb - 1) / wpb * wpb;
        if ((size_t)warps * wb > sl.serial_ws_bytes) {
            cudaFree(sl.d_serial_ws);
            sl.d_serial_ws = nullptr; sl.serial_ws_bytes = 0;
            CU_TRY(h, cudaMalloc((void **)&sl.d_serial_ws, (size_t)warps * wb));
            sl.serial_ws_bytes = (size_t)warps * wb;
        }
        bp_serial_kernel<real><<<warps / wpb, wpb * 32, 0, st>>>(a, h->d_order, sl.d_serial_ws, wb);
    } else if (h->bp_kernel == 3) {
        if (a.uniform_prior && h->clus_u.CL > 0) {
            const int ncl = (int)std::min<long long>(Bc, h->clus_u_nclusters);
            CU_TRY(h, cluster_launch<real>(h->clus_u, a, ncl, h->clus_u_threads, (size_t)h->clus_u_smem, h->clus_u_flip_table, st));
        } else {
            const int ncl = (int)std::min<long long>(Bc, h->clus_nclusters);
            CU_TRY(h, cluster_launch<real>(h->clus, a, ncl, h->bp_threads, (size_t)h->bp_smem, h->clus_flip_table, st));
        }
    } else if (h->bp_kernel == 2) {
        if (a.uniform_prior) // no prior array in shared memory: smaller footprint, possibly one more CTA per SM
            fast_launch<real>(h->fast, h->bp_geom, a, (int)std::min<long long>(Bc, h->bp_grid_uni), h->bp_threads, h->bp_smem_uni, st);
        else fast_launch<real>(h->fast, h->bp_geom, a, grid, h->bp_threads, h->bp_smem, st);
    }
    else if (h->bp_kernel == 1) bp_generic_kernel<real, true><<<grid, h->bp_threads, h->bp_smem, st>>>(a);
    else bp_generic_kernel<real, false><<<grid, h->bp_threads, h->bp_smem, st>>>(a);
    CU_TRY(h, cudaGetLastError());
    launches++;
    CU_TRY(h, cudaEventRecord(sl.ev[1], st));
    if (osd_on) {
        int nl = 0;
        const int rc = launch_osd<real>(h, st, a.g, a.synd, synd_packed, llr_out ? llr_out : static_cast<const real *>(sl.d_fail_llr), llr_out ? 1 : 0,
                                        d_fail_count, sl.d_fail_list, a.osd0, a.osdw, sl.d_ctrl + 1, Bc, d_priors != nullptr, d_weights, &nl);
        if (rc) return rc;
        launches += nl;
    }
    CU_TRY(h, cudaEventRecord(sl.ev[2], st));
    CU_TRY(h, cudaMemcpyAsync(sl.h_ctrl, sl.d_ctrl, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    sl.pending = true;
    sl.pending_shots = Bc;
    sl.pending_launches = launches;
    return BPOSD_OK;
}

// Wait for a slot's chunk (ev[3] must have been recorded after everything enqueued for it) and fold its
// counters and CUDA-event times into the handle's statistics.
static int collect_chunk(bposd_handle *h, bposd_handle::Slot &sl) {
    if (!sl.pending) return BPOSD_OK;
    sl.pending = false;
    CU_TRY(h, cudaEventSynchronize(sl.ev[3]));
    float t = 0;
    CU_TRY(h, cudaEventElapsedTime(&t, sl.ev[0], sl.ev[1])); h->stats.ms_bp += t;
    CU_TRY(h, cudaEventElapsedTime(&t, sl.ev[1], sl.ev[2])); h->stats.ms_osd += t;
    h->stats.shots += sl.pending_shots;
    h->stats.bp_converged += (int64_t)sl.h_ctrl[1];
    h->stats.bp_iterations += (int64_t)sl.h_ctrl[2];
    h->stats.osd_invocations += (int64_t)sl.h_ctrl[3];
    h->stats.launches += sl.pending_launches;
    h->stats.chunks += 1;
    return BPOSD_OK;
}

static int check_osd_supported(bposd_handle *h) {
    if (h->osd_method != BPOSD_OSD_OFF && !h->osd_supported)
        return fail(h, BPOSD_EUNSUP, "OSD for this matrix size needs more shared memory than an SM has (the HBM-resident "
                                     "kernel handles OSD-0 only); use osd_method osd0 or off");
    return BPOSD_OK;
}

static int ensure_buffers(bposd_handle *h, bposd_handle::Slot &sl, long long B, bool want_llr, bool want_err, bool want_packed);
static int pack_launch(bposd_handle *h, const uint8_t *src, long long Bc, uint8_t *dst, cudaStream_t st);

template <typename real>
static int decode_batch_t(bposd_handle *h, const uint8_t *d_synd, long long B, const bposd_out_t *out,
                          const void *d_priors, const double *d_weights, cudaStream_t st, int packed = 0) {
    const int n = h->n, m = h->m;
    int rc = check_osd_supported(h);
    if (rc) return rc;
    const bool need_ws = h->osd_method != BPOSD_OSD_OFF && !out->d_llr;
    long long chunk = std::min<long long>(need_ws ? std::min(B, h->fail_cap) : B, 1ll << 30);
    // bit-packed form: the kernels write one byte per bit into the slot's buffers, a pack kernel hands the bits to the caller
    const long long sb = packed ? (m + 7) / 8 : m, db = packed ? (n + 7) / 8 : n;
    if (packed) chunk = std::min<long long>(chunk, 262144);
    h->stats = bposd_stats_t{};
    bposd_handle::Slot &sl = h->slot[0];
    if (packed) {
        rc = collect_chunk(h, sl);
        if (rc) return rc;
        rc = ensure_buffers(h, sl, std::min(chunk, B), false, false, false);
        if (rc) return rc;
    }
    for (long long c0 = 0; c0 < B; c0 += chunk) {
        const long long Bc = std::min(chunk, B - c0);
        bposd_out_t o{};
        if (!packed) {
            o.d_bp = out->d_bp ? out->d_bp + c0 * n : nullptr;
            o.d_osd0 = out->d_osd0 ? out->d_osd0 + c0 * n : nullptr;
            o.d_osdw = out->d_osdw ? out->d_osdw + c0 * n : nullptr;
        } else {
            o.d_bp = out->d_bp ? sl.b_bp : nullptr;
            o.d_osd0 = out->d_osd0 ? sl.b_osd0 : nullptr;
            o.d_osdw = out->d_osdw ? sl.b_osdw : nullptr;
        }
        o.d_llr = out->d_llr ? static_cast<void *>(static_cast<real *>(out->d_llr) + c0 * n) : nullptr;
        o.d_converge = out->d_converge ? out->d_converge + c0 : nullptr;
        o.d_iter = out->d_iter ? out->d_iter + c0 : nullptr;
        rc = launch_chunk<real>(h, sl, st, d_synd + c0 * sb, Bc, o,
                                d_priors ? static_cast<const void *>(static_cast<const real *>(d_priors) + c0 * n) : nullptr,
                                d_weights ? d_weights + c0 * n : nullptr, packed);
        if (rc) return rc;
        if (packed) {
            if (out->d_osdw) { rc = pack_launch(h, sl.b_osdw, Bc, out->d_osdw + c0 * db, st); if (rc) return rc; }
            if (out->d_osd0) { rc = pack_launch(h, sl.b_osd0, Bc, out->d_osd0 + c0 * db, st); if (rc) return rc; }
            if (out->d_bp) { rc = pack_launch(h, sl.b_bp, Bc, out->d_bp + c0 * db, st); if (rc) return rc; }
            sl.pending_launches += (out->d_osdw ? 1 : 0) + (out->d_osd0 ? 1 : 0) + (out->d_bp ? 1 : 0);
        }
        CU_TRY(h, cudaEventRecord(sl.ev[3], st));
        rc = collect_chunk(h, sl); // the control words and events are re-used by the next chunk
        if (rc) return rc;
    }
    return BPOSD_OK;
}

static int decode_batch_entry(bposd_t *h, const uint8_t *d_synd, int64_t B, const bposd_out_t *out, const void *d_priors,
                              const double *d_weights, void *stream, int packed) {
    if (!h) return BPOSD_EINVAL;
    if (!out || (!d_synd && B > 0 && h->m > 0)) return fail(h, BPOSD_EINVAL, "NULL argument");
    if (B < 0) return fail(h, BPOSD_EINVAL, "negative batch size");
    CU_TRY(h, cudaSetDevice(h->device));
    if (B == 0) { h->stats = bposd_stats_t{}; return BPOSD_OK; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return h->precision == 64 ? decode_batch_t<double>(h, d_synd, B, out, d_priors, d_weights, st, packed)
                              : decode_batch_t<float>(h, d_synd, B, out, d_priors, d_weights, st, packed);
}

extern "C" int bposd_decode_batch(bposd_t *h, const uint8_t *d_synd, int64_t B, const bposd_out_t *out,
                                  const void *d_priors, const double *d_weights, void *stream) {
    return decode_batch_entry(h, d_synd, B, out, d_priors, d_weights, stream, 0);
}

extern "C" int bposd_decode_batch_packed(bposd_t *h, const uint8_t *d_synd_bits, int64_t B, const bposd_out_t *out,
                                         const void *d_priors, const double *d_weights, void *stream) {
    return decode_batch_entry(h, d_synd_bits, B, out, d_priors, d_weights, stream, 1);
}

static int ensure_buffers(bposd_handle *h, bposd_handle::Slot &sl, long long B, bool want_llr, bool want_err, bool want_packed) {
    const size_t rs = h->precision == 64 ? 8 : 4;
    if (B > sl.b_cap) {
        cudaFree(sl.b_synd); cudaFree(sl.b_osdw); cudaFree(sl.b_osd0); cudaFree(sl.b_bp); cudaFree(sl.b_conv);
        cudaFree(sl.b_iter); cudaFree(sl.b_err); cudaFree(sl.b_llr);
        cudaFree(sl.b_pk_osdw); cudaFree(sl.b_pk_osd0); cudaFree(sl.b_pk_bp);
        sl.b_pk_osdw = sl.b_pk_osd0 = sl.b_pk_bp = nullptr;
        sl.b_synd = sl.b_osdw = sl.b_osd0 = sl.b_bp = sl.b_conv = sl.b_err = nullptr;
        sl.b_iter = nullptr; sl.b_llr = nullptr;
        sl.b_cap = 0;
        CU_TRY(h, cudaMalloc((void **)&sl.b_synd, (size_t)B * std::max(h->m, 1)));
        CU_TRY(h, cudaMalloc((void **)&sl.b_osdw, (size_t)B * h->n));
        CU_TRY(h, cudaMalloc((void **)&sl.b_osd0, (size_t)B * h->n));
        CU_TRY(h, cudaMalloc((void **)&sl.b_bp, (size_t)B * h->n));
        CU_TRY(h, cudaMalloc((void **)&sl.b_conv, (size_t)B));
        CU_TRY(h, cudaMalloc((void **)&sl.b_iter, (size_t)B * 4));
        sl.b_cap = B;
    }
    if (want_llr && !sl.b_llr) CU_TRY(h, cudaMalloc(&sl.b_llr, (size_t)sl.b_cap * h->n * rs));
    if (want_err && !sl.b_err) CU_TRY(h, cudaMalloc((void **)&sl.b_err, (size_t)sl.b_cap * h->n));
    if (want_packed && !sl.b_pk_osdw) {
        const size_t nb = ((size_t)h->n + 7) / 8;
        CU_TRY(h, cudaMalloc((void **)&sl.b_pk_osdw, (size_t)sl.b_cap * nb));
        CU_TRY(h, cudaMalloc((void **)&sl.b_pk_osd0, (size_t)sl.b_cap * nb));
        CU_TRY(h, cudaMalloc((void **)&sl.b_pk_bp, (size_t)sl.b_cap * nb));
    }
    return BPOSD_OK;
}

// [Bc, n] bytes -> [Bc, ceil(n/8)] bytes on `st` (the BP / OSD kernels write one byte per bit; the packed entry points
// hand out bits)
static int pack_launch(bposd_handle *h, const uint8_t *src, long long Bc, uint8_t *dst, cudaStream_t st) {
    const long long total = Bc * (((long long)h->n + 7) / 8);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16));
    pack_bits_kernel<<<grid, 256, 0, st>>>(src, Bc, h->n, dst);
    CU_TRY(h, cudaGetLastError());
    return BPOSD_OK;
}

static int ensure_stage(bposd_handle *h) {
    if (h->h_stage) return BPOSD_OK;
    CU_TRY(h, cudaHostAlloc((void **)&h->h_stage, h->stage_bytes, cudaHostAllocMapped));
    CU_TRY(h, cudaHostGetDevicePointer((void **)&h->d_stage, h->h_stage, 0));
    return BPOSD_OK;
}

// Latency path of the host-buffer decode: a handful of shots, single-shot decode() above all.  One kernel launch and
// one synchronisation per call: the kernel reads the syndromes from, and writes every result to, pinned host memory
// that the device addresses directly (no copy calls, no control words, no events), every shot has an SM to itself
// (grid = B <= SM count, static assignment) and the CTA uses the latency geometry (bp_fast_kernel.cuh: few bits per
// thread, so a pass is as short as one SM can make it).  The host reads the converge flags; only if a shot did not
// converge (0.14 % on the bench code) does it hand the list to the OSD kernel and synchronise a second time.
template <typename real>
static int decode_latency_t(bposd_handle *h, const uint8_t *h_synd, long long B, uint8_t *h_osdw, uint8_t *h_osd0,
                            uint8_t *h_bp, void *h_llr, uint8_t *h_conv, int32_t *h_iter, bool *taken) {
    const int n = h->n, m = h->m;
    *taken = false;
    auto al = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t o_synd = 0, o_osdw = o_synd + al((size_t)B * m), o_osd0 = o_osdw + al((size_t)B * n), o_bp = o_osd0 + al((size_t)B * n),
                 o_llr = o_bp + al((size_t)B * n), o_conv = o_llr + (h_llr ? al((size_t)B * n * sizeof(real)) : 0),
                 o_iter = o_conv + al((size_t)B), o_list = o_iter + al((size_t)B * 4), o_cnt = o_list + al((size_t)B * 4),
                 total = o_cnt + 16;
    if (total > h->stage_bytes) return BPOSD_OK;
    *taken = true;
    int rc = ensure_stage(h);
    if (rc) return rc;
    for (auto &s2 : h->slot) {
        rc = collect_chunk(h, s2);
        if (rc) return rc;
    }
    bposd_handle::Slot &sl = h->slot[0];
    const bool osd_on = h->osd_method != BPOSD_OSD_OFF;
    if (!h_llr) {
        rc = ensure_buffers(h, sl, B, true, false, false);
        if (rc) return rc;
    }
    cudaStream_t st = sl.stream;
    uint8_t *sg = h->h_stage, *dg = h->d_stage;
    std::memcpy(sg + o_synd, h_synd, (size_t)B * m);
    BpArgs<real> a;
    a.g = graph_of(h);
    a.max_iter = h->max_iter;
    a.method = h->bp_method;
    a.alpha0 = (real)h->alpha;
    a.safe_it = sizeof(real) == 4 ? 0x7fffffff : h->safe_it; // computed once per channel (a loop of n logarithms: 10 us on this path)
    a.uniform_prior = h->uniform_prior;
    a.prior = (sizeof(real) == 8) ? (const real *)h->d_prior64 : (const real *)h->d_prior32;
    a.prior_stride = 0;
    a.synd = dg + o_synd;
    a.synd_packed = 0;
    a.B = B;
    a.bp = h_bp ? dg + o_bp : nullptr;
    a.osd0 = h_osd0 ? dg + o_osd0 : nullptr;
    a.osdw = h_osdw ? dg + o_osdw : nullptr;
    a.llr = h_llr ? reinterpret_cast<real *>(dg + o_llr) : static_cast<real *>(sl.b_llr);
    a.converge = dg + o_conv;
    a.iter = reinterpret_cast<int *>(dg + o_iter);
    a.fail_count = nullptr; a.fail_list = nullptr; a.fail_llr = nullptr;
    a.osd_off = osd_on ? 0 : 1;
    a.queue = nullptr; a.stat = nullptr;
    a.g_scratch = nullptr; a.g_dec = nullptr;
    fast_launch<real>(h->fast, h->lat_geom, a, (int)B, h->lat_threads, h->bp_smem, st);
    CU_TRY(h, cudaGetLastError());
    // (polling a completion flag in pinned memory instead was measured: no gain over this, profiles/r03b_lat_probe.log)
    CU_TRY(h, cudaStreamSynchronize(st));
    int launches = 1;
    const uint8_t *conv = sg + o_conv;
    const int32_t *iters = reinterpret_cast<const int32_t *>(sg + o_iter);
    int *list = reinterpret_cast<int *>(sg + o_list), *cnt = reinterpret_cast<int *>(sg + o_cnt);
    long long nconv = 0, niter = 0;
    int nfail = 0;
    for (long long b = 0; b < B; b++) {
        nconv += conv[b] ? 1 : 0;
        niter += iters[b];
        if (!conv[b]) list[nfail++] = (int)b;
    }
    if (osd_on && nfail > 0 && (a.osd0 || a.osdw)) {
        *cnt = nfail;
        int nl = 0;
        rc = launch_osd<real>(h, st, a.g, a.synd, 0, a.llr, 1, reinterpret_cast<const int *>(dg + o_cnt),
                              reinterpret_cast<const int *>(dg + o_list), a.osd0, a.osdw, nullptr, nfail, false, nullptr, &nl);
        if (rc) return rc;
        CU_TRY(h, cudaStreamSynchronize(st));
        launches += nl;
    }
    if (h_osdw) std::memcpy(h_osdw, sg + o_osdw, (size_t)B * n);
    if (h_osd0) std::memcpy(h_osd0, sg + o_osd0, (size_t)B * n);
    if (h_bp) std::memcpy(h_bp, sg + o_bp, (size_t)B * n);
    if (h_llr) std::memcpy(h_llr, sg + o_llr, (size_t)B * n * sizeof(real));
    if (h_conv) std::memcpy(h_conv, conv, (size_t)B);
    if (h_iter) std::memcpy(h_iter, iters, (size_t)B * 4);
    h->stats = bposd_stats_t{};
    h->stats.shots = B;
    h->stats.bp_converged = nconv;
    h->stats.bp_iterations = niter;
    h->stats.osd_invocations = osd_on ? nfail : 0;
    h->stats.launches = launches;
    h->stats.chunks = 1;
    return BPOSD_OK;
}

// Host-buffer decode: the batch is cut into chunks that alternate between the two slots, each on its
// own stream (H2D -> BP -> OSD -> D2H in stream order), so the copies of one chunk overlap the
// kernels of the other.  Kernels that share a per-handle scratch (HBM-scratch BP, HBM OSD) run
// single-slot.
template <typename real>
static int decode_host_t(bposd_handle *h, const uint8_t *h_synd, long long B, uint8_t *h_osdw, uint8_t *h_osd0,
                         uint8_t *h_bp, void *h_llr, uint8_t *h_conv, int32_t *h_iter, int packed = 0) {
    const int n = h->n, m = h->m;
    int rc = check_osd_supported(h);
    if (rc) return rc;
    // bytes per shot on the host side: one per bit, or bit-packed (syndromes ceil(m/8), decodings ceil(n/8))
    const size_t sb = packed ? ((size_t)m + 7) / 8 : (size_t)m, db = packed ? ((size_t)n + 7) / 8 : (size_t)n;
    if (h->bp_kernel == 2 && h->lat_geom >= 0 && B <= h->lat_max_shots && h->schedule == 0) {
        bool taken = false;
        if (!packed) {
            rc = decode_latency_t<real>(h, h_synd, B, h_osdw, h_osd0, h_bp, h_llr, h_conv, h_iter, &taken);
        } else {
            // a handful of shots: the bits are unpacked / packed on the host around the byte-per-bit latency path
            std::vector<uint8_t> us((size_t)B * m), uw(h_osdw ? (size_t)B * n : 0), u0(h_osd0 ? (size_t)B * n : 0), ub(h_bp ? (size_t)B * n : 0);
            for (long long b = 0; b < B; b++)
                for (int i = 0; i < m; i++) us[(size_t)b * m + i] = (h_synd[(size_t)b * sb + (i >> 3)] >> (i & 7)) & 1;
            rc = decode_latency_t<real>(h, us.data(), B, h_osdw ? uw.data() : nullptr, h_osd0 ? u0.data() : nullptr,
                                        h_bp ? ub.data() : nullptr, h_llr, h_conv, h_iter, &taken);
            if (!rc && taken) {
                auto pack = [&](const std::vector<uint8_t> &u, uint8_t *dst) {
                    std::memset(dst, 0, (size_t)B * db);
                    for (long long b = 0; b < B; b++)
                        for (int j = 0; j < n; j++) dst[(size_t)b * db + (j >> 3)] |= (uint8_t)((u[(size_t)b * n + j] & 1) << (j & 7));
                };
                if (h_osdw) pack(uw, h_osdw);
                if (h_osd0) pack(u0, h_osd0);
                if (h_bp) pack(ub, h_bp);
            }
        }
        if (rc || taken) return rc;
    }
    const bool need_ws = h->osd_method != BPOSD_OSD_OFF && !h_llr;
    const bool pipelined = h->bp_kernel != 0 && !h->osd_large && !h->osd_clus && B >= 2 * h->host_chunk_min;
    long long chunk = B;
    if (pipelined) chunk = std::min<long long>(std::max<long long>((B + 15) / 16, h->host_chunk_min), h->host_chunk_max);
    if (need_ws) chunk = std::min(chunk, h->fail_cap);
    const int nslots = pipelined ? 2 : 1;
    h->stats = bposd_stats_t{};
    // device-side sources of the three decodings of a slot: the kernels' byte-per-bit buffers, or their packed copies
    auto pack_outputs = [&](bposd_handle::Slot &sl, long long Bc, cudaStream_t st) -> int {
        if (!packed) return BPOSD_OK;
        int r = BPOSD_OK;
        if (h_osdw && !r) r = pack_launch(h, sl.b_osdw, Bc, sl.b_pk_osdw, st);
        if (h_osd0 && !r) r = pack_launch(h, sl.b_osd0, Bc, sl.b_pk_osd0, st);
        if (h_bp && !r) r = pack_launch(h, sl.b_bp, Bc, sl.b_pk_bp, st);
        sl.pending_launches += (h_osdw ? 1 : 0) + (h_osd0 ? 1 : 0) + (h_bp ? 1 : 0);
        return r;
    };
    // Small batches (single-shot decode() above all): pageable host buffers would make every copy a blocking
    // call, so inputs and outputs go through one pinned staging block and the copies are truly asynchronous.
    {
        auto al = [](size_t x) { return (x + 15) / 16 * 16; };
        const size_t o_synd = 0, o_osdw = o_synd + al((size_t)B * sb), o_osd0 = o_osdw + (h_osdw ? al((size_t)B * db) : 0),
                     o_bp = o_osd0 + (h_osd0 ? al((size_t)B * db) : 0), o_llr = o_bp + (h_bp ? al((size_t)B * db) : 0),
                     o_conv = o_llr + (h_llr ? al((size_t)B * n * sizeof(real)) : 0), o_iter = o_conv + (h_conv ? al((size_t)B) : 0),
                     total = o_iter + (h_iter ? al((size_t)B * 4) : 0);
        if (!pipelined && chunk == B && total <= h->stage_bytes) {
            rc = ensure_stage(h);
            if (rc) return rc;
            bposd_handle::Slot &sl = h->slot[0];
            rc = collect_chunk(h, sl);
            if (rc) return rc;
            rc = ensure_buffers(h, sl, B, h_llr != nullptr, false, packed != 0);
            if (rc) return rc;
            cudaStream_t st = sl.stream;
            uint8_t *sg = h->h_stage;
            std::memcpy(sg + o_synd, h_synd, (size_t)B * sb);
            CU_TRY(h, cudaMemcpyAsync(sl.b_synd, sg + o_synd, (size_t)B * sb, cudaMemcpyHostToDevice, st));
            bposd_out_t o{};
            o.d_osdw = h_osdw ? sl.b_osdw : nullptr;
            o.d_osd0 = h_osd0 ? sl.b_osd0 : nullptr;
            o.d_bp = h_bp ? sl.b_bp : nullptr;
            o.d_llr = h_llr ? sl.b_llr : nullptr;
            o.d_converge = h_conv ? sl.b_conv : nullptr;
            o.d_iter = h_iter ? sl.b_iter : nullptr;
            rc = launch_chunk<real>(h, sl, st, sl.b_synd, B, o, nullptr, nullptr, packed);
            if (rc) return rc;
            rc = pack_outputs(sl, B, st);
            if (rc) return rc;
            if (h_osdw) CU_TRY(h, cudaMemcpyAsync(sg + o_osdw, packed ? sl.b_pk_osdw : sl.b_osdw, (size_t)B * db, cudaMemcpyDeviceToHost, st));
            if (h_osd0) CU_TRY(h, cudaMemcpyAsync(sg + o_osd0, packed ? sl.b_pk_osd0 : sl.b_osd0, (size_t)B * db, cudaMemcpyDeviceToHost, st));
            if (h_bp) CU_TRY(h, cudaMemcpyAsync(sg + o_bp, packed ? sl.b_pk_bp : sl.b_bp, (size_t)B * db, cudaMemcpyDeviceToHost, st));
            if (h_llr) CU_TRY(h, cudaMemcpyAsync(sg + o_llr, sl.b_llr, (size_t)B * n * sizeof(real), cudaMemcpyDeviceToHost, st));
            if (h_conv) CU_TRY(h, cudaMemcpyAsync(sg + o_conv, sl.b_conv, (size_t)B, cudaMemcpyDeviceToHost, st));
            if (h_iter) CU_TRY(h, cudaMemcpyAsync(sg + o_iter, sl.b_iter, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(h, cudaEventRecord(sl.ev[3], st));
            rc = collect_chunk(h, sl);
            if (rc) return rc;
            if (h_osdw) std::memcpy(h_osdw, sg + o_osdw, (size_t)B * db);
            if (h_osd0) std::memcpy(h_osd0, sg + o_osd0, (size_t)B * db);
            if (h_bp) std::memcpy(h_bp, sg + o_bp, (size_t)B * db);
            if (h_llr) std::memcpy(h_llr, sg + o_llr, (size_t)B * n * sizeof(real));
            if (h_conv) std::memcpy(h_conv, sg + o_conv, (size_t)B);
            if (h_iter) std::memcpy(h_iter, sg + o_iter, (size_t)B * 4);
            return BPOSD_OK;
        }
    }
    long long idx = 0;
    for (long long c0 = 0; c0 < B; c0 += chunk, idx++) {
        const long long Bc = std::min(chunk, B - c0);
        bposd_handle::Slot &sl = h->slot[idx % nslots];
        rc = collect_chunk(h, sl); // the slot's previous chunk must have left its buffers
        if (rc) return rc;
        rc = ensure_buffers(h, sl, std::min(chunk, B), h_llr != nullptr, false, packed != 0);
        if (rc) return rc;
        cudaStream_t st = sl.stream;
        CU_TRY(h, cudaMemcpyAsync(sl.b_synd, h_synd + c0 * sb, (size_t)Bc * sb, cudaMemcpyHostToDevice, st));
        bposd_out_t o{};
        o.d_osdw = h_osdw ? sl.b_osdw : nullptr;
        o.d_osd0 = h_osd0 ? sl.b_osd0 : nullptr;
        o.d_bp = h_bp ? sl.b_bp : nullptr;
        o.d_llr = h_llr ? sl.b_llr : nullptr;
        o.d_converge = h_conv ? sl.b_conv : nullptr;
        o.d_iter = h_iter ? sl.b_iter : nullptr;
        rc = launch_chunk<real>(h, sl, st, sl.b_synd, Bc, o, nullptr, nullptr, packed);
        if (rc) return rc;
        rc = pack_outputs(sl, Bc, st);
        if (rc) return rc;
        if (h_osdw) CU_TRY(h, cudaMemcpyAsync(h_osdw + c0 * db, packed ? sl.b_pk_osdw : sl.b_osdw, (size_t)Bc * db, cudaMemcpyDeviceToHost, st));
        if (h_osd0) CU_TRY(h, cudaMemcpyAsync(h_osd0 + c0 * db, packed ? sl.b_pk_osd0 : sl.b_osd0, (size_t)Bc * db, cudaMemcpyDeviceToHost, st));
        if (h_bp) CU_TRY(h, cudaMemcpyAsync(h_bp + c0 * db, packed ? sl.b_pk_bp : sl.b_bp, (size_t)Bc * db, cudaMemcpyDeviceToHost, st));
        if (h_llr) CU_TRY(h, cudaMemcpyAsync(static_cast<real *>(h_llr) + c0 * n, sl.b_llr, (size_t)Bc * n * sizeof(real), cudaMemcpyDeviceToHost, st));
        if (h_conv) CU_TRY(h, cudaMemcpyAsync(h_conv + c0, sl.b_conv, (size_t)Bc, cudaMemcpyDeviceToHost, st));
        if (h_iter) CU_TRY(h, cudaMemcpyAsync(h_iter + c0, sl.b_iter, (size_t)Bc * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaEventRecord(sl.ev[3], st));
    }
    for (int k = 0; k < 2; k++) {
        rc = collect_chunk(h, h->slot[k]);
        if (rc) return rc;
    }
    return BPOSD_OK;
}

static int decode_host_entry(bposd_t *h, const uint8_t *h_synd, int64_t B, uint8_t *h_osdw, uint8_t *h_osd0, uint8_t *h_bp,
                             void *h_llr, uint8_t *h_conv, int32_t *h_iter, int packed) {
    if (!h) return BPOSD_EINVAL;
    if (B < 0 || (!h_synd && B > 0 && h->m > 0)) return fail(h, BPOSD_EINVAL, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    if (B == 0) { h->stats = bposd_stats_t{}; return BPOSD_OK; }
    return h->precision == 64 ? decode_host_t<double>(h, h_synd, B, h_osdw, h_osd0, h_bp, h_llr, h_conv, h_iter, packed)
                              : decode_host_t<float>(h, h_synd, B, h_osdw, h_osd0, h_bp, h_llr, h_conv, h_iter, packed);
}

extern "C" int bposd_decode_host(bposd_t *h, const uint8_t *h_synd, int64_t B, uint8_t *h_osdw, uint8_t *h_osd0,
                                 uint8_t *h_bp, void *h_llr, uint8_t *h_conv, int32_t *h_iter) {
    return decode_host_entry(h, h_synd, B, h_osdw, h_osd0, h_bp, h_llr, h_conv, h_iter, 0);
}

extern "C" int bposd_decode_host_packed(bposd_t *h, const uint8_t *h_synd_bits, int64_t B, uint8_t *h_osdw_bits, uint8_t *h_osd0_bits,
                                        uint8_t *h_bp_bits, void *h_llr, uint8_t *h_conv, int32_t *h_iter) {
    return decode_host_entry(h, h_synd_bits, B, h_osdw_bits, h_osd0_bits, h_bp_bits, h_llr, h_conv, h_iter, 1);
}

extern "C" int bposd_set_channel_thresholds(bposd_t *h, const uint32_t *t1, const uint32_t *t2, const uint32_t *t3) {
    if (!h || !t1 || !t2 || !t3) return fail(h, BPOSD_EINVAL, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->n * sizeof(uint32_t);
    for (int j = 0; j < h->n; j++)
        if (t1[j] > t2[j] || t2[j] > t3[j]) return fail(h, BPOSD_EINVAL, "thresholds must be cumulative (t1 <= t2 <= t3)");
    if (!h->d_t1) {
        CU_TRY(h, cudaMalloc((void **)&h->d_t1, bytes));
        CU_TRY(h, cudaMalloc((void **)&h->d_t2, bytes));
        CU_TRY(h, cudaMalloc((void **)&h->d_t3, bytes));
    }
    CU_TRY(h, cudaMemcpy(h->d_t1, t1, bytes, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_t2, t2, bytes, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_t3, t3, bytes, cudaMemcpyHostToDevice));
    return BPOSD_OK;
}

static int sample_entry(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector, uint8_t *d_errors, uint8_t *d_synd,
                        void *stream, int packed) {
    if (!h) return BPOSD_EINVAL;
    if (!h->d_t1) return fail(h, BPOSD_EINVAL, "call bposd_set_channel_thresholds first");
    if (B < 0 || (!d_synd && B > 0) || (sector != 0 && sector != 1)) return fail(h, BPOSD_EINVAL, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    if (B == 0) return BPOSD_OK;
    SampleArgs a;
    a.g = graph_of(h);
    a.t1 = h->d_t1; a.t2 = h->d_t2; a.t3 = h->d_t3;
    a.seed = seed; a.shot0 = shot0; a.B = B; a.sector = sector;
    a.errors = d_errors; a.synd = d_synd; a.synd_packed = packed;
    const int threads = std::min(1024, std::max(32, ((h->n + 3) / 4 + 31) / 32 * 32));
    const size_t smem = (size_t)h->n + 16;
    if (smem > 48 * 1024) CU_TRY(h, cudaFuncSetAttribute(sample_syndrome_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<long long>(B, (long long)h->sm_count * 8);
    sample_syndrome_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    CU_TRY(h, cudaGetLastError());
    return BPOSD_OK;
}

extern "C" int bposd_sample_syndromes(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector,
                                      uint8_t *d_errors, uint8_t *d_synd, void *stream) {
    return sample_entry(h, seed, shot0, B, sector, d_errors, d_synd, stream, 0);
}

extern "C" int bposd_sample_syndromes_packed(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector,
                                             uint8_t *d_errors, uint8_t *d_synd_bits, void *stream) {
    return sample_entry(h, seed, shot0, B, sector, d_errors, d_synd_bits, stream, 1);
}

extern "C" int bposd_set_logicals(bposd_t *h, const int32_t *indptr, const int32_t *indices, int32_t K) {
    if (!h || !indptr || K < 0) return fail(h, BPOSD_EINVAL, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    for (int e = 0; e < indptr[K]; e++)
        if (indices[e] < 0 || indices[e] >= h->n) return fail(h, BPOSD_EINVAL, "logical operator column out of range");
    cudaFree(h->d_l_ptr); cudaFree(h->d_l_idx);
    h->d_l_ptr = h->d_l_idx = nullptr;
    std::vector<int> p(indptr, indptr + K + 1), x(indices, indices + indptr[K]);
    CU_TRY(h, upload(&h->d_l_ptr, p));
    CU_TRY(h, upload(&h->d_l_idx, x));
    h->K = K;
    return BPOSD_OK;
}

static int logical_launch(bposd_handle *h, const uint8_t *d_err, const uint8_t *d_dec, long long B, uint8_t *d_fail,
                          unsigned long long *d_count, int *d_minw, cudaStream_t st, int *d_resid_weight = nullptr) {
    LogicalArgs a;
    a.n = h->n; a.K = h->K; a.l_ptr = h->d_l_ptr; a.l_idx = h->d_l_idx;
    a.errors = d_err; a.dec = d_dec; a.B = B; a.fail = d_fail; a.fail_count = d_count; a.min_weight = d_minw;
    a.resid_weight = d_resid_weight;
    const int grid = (int)std::min<long long>((B + 7) / 8, (long long)h->sm_count * 8);
    logical_check_kernel<<<std::max(grid, 1), 256, 0, st>>>(a);
    CU_TRY(h, cudaGetLastError());
    return BPOSD_OK;
}

extern "C" int bposd_logical_check(bposd_t *h, const uint8_t *d_err, const uint8_t *d_dec, int64_t B, uint8_t *d_fail,
                                   int64_t *d_fail_count, int32_t *d_min_weight, int32_t *d_resid_weight, void *stream) {
    if (!h) return BPOSD_EINVAL;
    if (!h->d_l_ptr) return fail(h, BPOSD_EINVAL, "call bposd_set_logicals first");
    if (B < 0 || ((!d_err || !d_dec) && B > 0)) return fail(h, BPOSD_EINVAL, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    if (B == 0) return BPOSD_OK;
    return logical_launch(h, d_err, d_dec, B, d_fail, reinterpret_cast<unsigned long long *>(d_fail_count), d_min_weight,
                          static_cast<cudaStream_t>(stream), d_resid_weight);
}

extern "C" int bposd_channel_update(bposd_t *h, const uint8_t *d_first, int64_t B, const double *h_p0, const double *h_p1,
                                    void *d_priors, double *d_weights, void *stream) {
    if (!h) return BPOSD_EINVAL;
    if (B < 0 || !h_p0 || !h_p1 || ((!d_first || !d_priors) && B > 0)) return fail(h, BPOSD_EINVAL, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const int n = h->n;
    for (int j = 0; j < n; j++)
        if (!(h_p0[j] >= 0.0 && h_p0[j] <= 1.0) || !(h_p1[j] >= 0.0 && h_p1[j] <= 1.0))
            return fail(h, BPOSD_EINVAL, "channel probabilities must lie in [0, 1]");
    // tables [prior0 | prior1 | weight0 | weight1], the same expressions as upload_probs (rows a3, a14)
    std::vector<double> tab(4 * (size_t)n);
    for (int j = 0; j < n; j++) {
        tab[j] = std::log((1.0 - h_p0[j]) / h_p0[j]);
        tab[n + j] = std::log((1.0 - h_p1[j]) / h_p1[j]);
        tab[2 * (size_t)n + j] = std::log(1 / h_p0[j]);
        tab[3 * (size_t)n + j] = std::log(1 / h_p1[j]);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!h->d_cu_tab) CU_TRY(h, cudaMalloc((void **)&h->d_cu_tab, 4 * (size_t)n * sizeof(double)));
    // stream-ordered with the kernel below; the staging vector is pageable, so the copy returns after staging
    CU_TRY(h, cudaMemcpyAsync(h->d_cu_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    if (B == 0) return BPOSD_OK;
    const long long total = (long long)B * n;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
    const double *t = h->d_cu_tab;
    if (h->precision == 64)
        channel_update_kernel<double><<<grid, 256, 0, st>>>(d_first, total, n, t, t + n, t + 2 * (size_t)n, t + 3 * (size_t)n,
                                                            static_cast<double *>(d_priors), d_weights);
    else
        channel_update_kernel<float><<<grid, 256, 0, st>>>(d_first, total, n, t, t + n, t + 2 * (size_t)n, t + 3 * (size_t)n,
                                                           static_cast<float *>(d_priors), d_weights);
    CU_TRY(h, cudaGetLastError());
    return BPOSD_OK;
}

extern "C" int bposd_css_counters(bposd_t *h, int64_t B, const bposd_css_sector_t *osdw, const bposd_css_sector_t *osd0,
                                  const bposd_css_sector_t *bp, const uint8_t *d_conv_x, const uint8_t *d_conv_z,
                                  int64_t *h_counters, void *stream) {
    if (!h) return BPOSD_EINVAL;
    if (B < 0 || !osdw || !osd0 || !bp || !h_counters || ((!d_conv_x || !d_conv_z) && B > 0)) return fail(h, BPOSD_EINVAL, "bad argument");
    for (const bposd_css_sector_t *s : {osdw, osd0, bp})
        if ((!s->d_fail_x || !s->d_fail_z) && B > 0) return fail(h, BPOSD_EINVAL, "failure flags missing");
    CU_TRY(h, cudaSetDevice(h->device));
    if (B == 0) return BPOSD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU_TRY(h, cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(unsigned long long), st));
    const int big = 0x7fffffff;
    CU_TRY(h, cudaMemcpyAsync(h->d_minw, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    auto conv = [](const bposd_css_sector_t *s) { CssSector c; c.fail_x = s->d_fail_x; c.fail_z = s->d_fail_z; c.weight_x = s->d_weight_x; c.weight_z = s->d_weight_z; return c; };
    const int grid = (int)std::max<long long>(1, std::min<long long>((B + 255) / 256, (long long)h->sm_count * 4));
    css_counters_kernel<<<grid, 256, 0, st>>>(B, conv(osdw), conv(osd0), conv(bp), d_conv_x, d_conv_z, h->d_counters, h->d_minw);
    CU_TRY(h, cudaGetLastError());
    unsigned long long c[8];
    int minw = 0;
    CU_TRY(h, cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaMemcpyAsync(&minw, h->d_minw, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    h_counters[0] += B;
    for (int k = 1; k <= 5; k++) h_counters[k] += (int64_t)c[k];
    if (minw != big && (h_counters[6] <= 0 || minw < h_counters[6])) h_counters[6] = minw;
    return BPOSD_OK;
}

__global__ void bp_success_kernel(const uint8_t *fail, const uint8_t *conv, long long B, unsigned long long *count) {
    // BP counts as a success only where it converged and the residual is trivial (css_decode_sim.py:336-349)
    unsigned long long c = 0;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x)
        c += (conv[b] && !fail[b]) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

extern "C" int bposd_sample_and_decode(bposd_t *h, uint64_t seed, uint64_t shot0, int64_t B, int32_t sector,
                                       int64_t *h_counters, void *stream) {
    if (!h || !h_counters) return fail(h, BPOSD_EINVAL, "NULL argument");
    if (!h->d_t1) return fail(h, BPOSD_EINVAL, "call bposd_set_channel_thresholds first");
    if (!h->d_l_ptr) return fail(h, BPOSD_EINVAL, "call bposd_set_logicals first");
    if (B < 0) return fail(h, BPOSD_EINVAL, "negative batch size");
    CU_TRY(h, cudaSetDevice(h->device));
    if (B == 0) return BPOSD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bposd_handle::Slot &sl = h->slot[0];
    int rc = ensure_buffers(h, sl, B, false, true, false);
    if (rc) return rc;
    rc = bposd_sample_syndromes(h, seed, shot0, B, sector, sl.b_err, sl.b_synd, stream);
    if (rc) return rc;
    bposd_out_t o{};
    o.d_osdw = sl.b_osdw; o.d_osd0 = sl.b_osd0; o.d_bp = sl.b_bp; o.d_converge = sl.b_conv;
    rc = bposd_decode_batch(h, sl.b_synd, B, &o, nullptr, nullptr, stream);
    if (rc) return rc;
    // d_counters: [0] osdw failures, [1] osd0 failures, [2] bp successes; b_iter reused as fail flags for bp
    CU_TRY(h, cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(unsigned long long), st));
    const int big = 0x7fffffff;
    CU_TRY(h, cudaMemcpyAsync(h->d_minw, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    rc = logical_launch(h, sl.b_err, sl.b_osdw, B, nullptr, h->d_counters + 0, h->d_minw, st);
    if (rc) return rc;
    rc = logical_launch(h, sl.b_err, sl.b_osd0, B, nullptr, h->d_counters + 1, h->d_minw, st);
    if (rc) return rc;
    uint8_t *bpfail = reinterpret_cast<uint8_t *>(sl.b_iter);
    rc = logical_launch(h, sl.b_err, sl.b_bp, B, bpfail, nullptr, nullptr, st);
    if (rc) return rc;
    bp_success_kernel<<<std::max(1, std::min(h->sm_count * 4, (int)((B + 255) / 256))), 256, 0, st>>>(bpfail, sl.b_conv, B, h->d_counters + 2);
    CU_TRY(h, cudaGetLastError());
    unsigned long long c[8];
    int minw = 0;
    CU_TRY(h, cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaMemcpyAsync(&minw, h->d_minw, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    h->stats.launches += 5;
    h_counters[0] += B;
    h_counters[1] += h->stats.bp_converged;
    h_counters[2] += (int64_t)c[2];
    h_counters[3] += B - (int64_t)c[1];
    h_counters[4] += B - (int64_t)c[0];
    h_counters[5] += h->stats.osd_invocations;
    h_counters[6] += h->stats.bp_iterations;
    if (minw != big && (h_counters[7] <= 0 || minw < h_counters[7])) h_counters[7] = minw;
    return BPOSD_OK;
}
