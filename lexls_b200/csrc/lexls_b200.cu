// lexls_b200.cu — C-ABI of the batched LexLSE / LexLSI solver (include/lexls_b200.h) and its kernel launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/lexls_b200.h"
#include "lexls_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

using namespace lexls_dev;

// ---------------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                                        \
    do                                                                                                                  \
    {                                                                                                                   \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess)                                                                                          \
            return fail(LEXLS_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                          \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------------
struct lexls_b200_batch;
struct Shard
{
    lexls_b200_batch *h;
    uint64_t first, count;
};

struct lexls_b200_batch
{
    // a multi-device batch (lexls_b200_create_batch_multi) owns one single-device handle per shard and nothing else: every
    // entry point fans out over the shards with the caller's arrays offset by the shard's first problem
    std::vector<Shard> shards;
    bool multi() const { return !shards.empty(); }

    int kind = 0, device = 0;
    uint64_t batch = 0;
    Shape sh{};
    Params pr{};
    lexls_b200_params_t user_params{};
    SmemMap sm{};
    int W = 1;
    bool DS = false, LG = false, pair = false;
    unsigned grid = 0, smem = 0, per_sm = 0;
    std::string kname;
    bool planned = false;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_switch = nullptr;
    bool timed = false;

    // device buffers
    double *rec = nullptr, *obj = nullptr, *sb = nullptr, *fval = nullptr, *x0 = nullptr, *v0 = nullptr;
    unsigned *sbidx = nullptr, *fidx = nullptr;
    unsigned char *ftype = nullptr, *guess = nullptr;
    unsigned char *rownz = nullptr; // LSI: [B][M] 1 = the row of A is not all zero (set_data)
    uint64_t rec_stride = 0, obj_stride = 0, sbidx_stride = 0;
    bool have_x0 = false, have_guess = false;
    std::vector<char> have_v0;
    std::vector<char> data_set;
    double *x = nullptr, *v = nullptr, *viol = nullptr, *lod_scratch = nullptr;
    unsigned char *act = nullptr;
    unsigned *order = nullptr, *counters = nullptr, *ranks = nullptr, *colperm = nullptr;
    double *lexqr = nullptr, *hhout = nullptr, *lambda = nullptr; // optional outputs (lexls_b200_request_outputs)
    double *xstar = nullptr, *lexdata = nullptr, *lexqr_star = nullptr, *hh_star = nullptr; // LEXLS_B200_OUT_XSTAR
    double reg_factor_user[LX_MAX_LEVELS] = {0};                   // setRegularizationFactor per user objective
    double *vec_scratch = nullptr;                                 // SmemMap::vecg: iteration vectors of the resident problems
    size_t vec_bytes = 0;
    double *xln = nullptr, *ln_scratch = nullptr;                  // LEXLS_B200_OUT_LEAST_NORM
    size_t ln_bytes = 0;
    uint64_t ln_stride = 0;
    double *reg_scratch = nullptr;                                 // regularization work arrays, one slice per resident CTA
    size_t reg_bytes = 0;
    uint64_t reg_stride = 0;
    unsigned *wsl = nullptr;                                       // chronological activation list (deactivate_first_wrong_sign)
    unsigned *glist = nullptr;                                     // api_activate calls in call order (set_active_guess_ordered)
    unsigned gcap = 0;
    bool have_glist = false;
    double *wslog = nullptr;                                      // working-set log (params.log_working_set_enabled)
    unsigned *wslog_count = nullptr;
    unsigned wslog_cap = 0;
    unsigned want = 0;
    int *status = nullptr, *flag = nullptr;
    int *hflag = nullptr;          // pinned host copy of `flag`, refreshed by every solve_batch (no extra synchronisation)
    bool async_copies = false;     // lexls_b200_set_async_copies: host set_* calls return without waiting for their copy
    bool flag_pending = false;     // a solve has been enqueued whose flag copy has not been inspected yet
    cudaEvent_t ev_copy = nullptr; // completion of the last host -> staging copy (the caller's buffer is then free)
    uint64_t nfix_cap = 0;         // elements allocated in fidx / fval / ftype
    std::vector<std::pair<void *, size_t>> guards; // LEXLS_B200_CANARY=1: (buffer, bytes) of every allocation, each followed by a guard band
    unsigned long long *queue = nullptr;
    uint64_t lod_stride = 0;
    size_t lod_bytes = 0;
    // staging for host inputs.  Default: one slot, every set_* call copies into it, waits for the copy and launches its repack
    // kernel.  With lexls_b200_set_async_copies the slot becomes an ARENA: every call takes its own slice, the copy is enqueued
    // and the repack launch is DEFERRED (pending) to the next call that needs the device records (solve, validate, a result
    // read, a *_dev setter ...).  A round of host inputs then crosses the link back to back — in stream order a copy would
    // otherwise wait for the repack kernel of the previous level, and that kernel for SM slots beside another handle's
    // persistent solve (the pipelined end-to-end path) — and the caller can wait for the copies alone (lexls_b200_wait_copies).
    void *stage = nullptr;
    size_t stage_bytes = 0, stage_cursor = 0, round_bytes = 0, round_peak = 0;
    std::vector<std::function<int()>> pending;
};


// LEXLS_B200_CANARY=1 (debug aid; compute-sanitizer is closed on the GPU pool): every device buffer of a handle is followed by a
// 256-byte guard band filled with 0xA5; lexls_b200_check_guard_bands counts the bands a kernel wrote into.
#define LX_GUARD_BYTES 256
static bool canary_enabled()
{
    const char *e = getenv("LEXLS_B200_CANARY");
    return e && atoi(e) != 0;
}
static cudaError_t guarded_malloc(lexls_b200_batch *h, void **p, size_t bytes)
{
    if (!canary_enabled()) return cudaMalloc(p, bytes);
    cudaError_t e = cudaMalloc(p, bytes + LX_GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaMemset((char *)*p + bytes, 0xA5, LX_GUARD_BYTES);
    if (e == cudaSuccess)
    {
        for (auto &g : h->guards)
            if (g.first == *p) g.second = (size_t)-1; // address reused after a free: the old record is stale
        h->guards.push_back(std::make_pair(*p, bytes));
    }
    return e;
}

static void guarded_free(lexls_b200_batch *h, void *p)
{
    for (auto &g : h->guards)
        if (g.first == p) g.second = (size_t)-1;
    cudaFree(p);
}

static int ensure_stage(lexls_b200_batch *h, size_t bytes)
{
    if (h->stage_bytes >= bytes) return 0;
    if (h->stage) CK(cudaStreamSynchronize(h->stream)); // an earlier (asynchronous) copy or repack may still use the old buffer
    if (h->stage) guarded_free(h, h->stage);
    h->stage       = nullptr;
    h->stage_bytes = 0;
    CK(guarded_malloc(h, (void **)&h->stage, bytes));
    h->stage_bytes = bytes;
    return 0;
}

// enqueue the deferred repack launches (stream order: after the copies they read)
static int flush_pending(lexls_b200_batch *h)
{
    if (h->pending.empty()) return 0;
    std::vector<std::function<int()>> todo;
    todo.swap(h->pending);
    int rc = 0;
    for (auto &f : todo)
        if (!rc) rc = f();
    return rc;
}
// launch now, or with the next flush (asynchronous-copies mode)
static int run_or_defer(lexls_b200_batch *h, std::function<int()> launch)
{
    if (!h->async_copies) return launch();
    h->pending.push_back(std::move(launch));
    return 0;
}
// staging memory for one host input: the single slot, or a fresh slice of the arena
static int stage_take(lexls_b200_batch *h, size_t bytes, void **p)
{
    if (!h->async_copies)
    {
        if (int rc = ensure_stage(h, bytes)) return rc;
        *p = h->stage;
        return 0;
    }
    bytes = (bytes + 255) & ~(size_t)255;
    if (h->stage_cursor + bytes > h->stage_bytes)
    {
        // wrap: the repacks that read the arena are enqueued first, so the copies that follow are ordered behind them by the
        // stream; the arena grows towards one full round of inputs (a few warm-up rounds), after which nothing wraps
        if (int rc = flush_pending(h)) return rc;
        h->stage_cursor = 0;
        // what this round has taken so far, or what the largest finished round took: with 1/8 of headroom the arena holds a
        // full round after the second one and never wraps again (a reallocation synchronises the whole device)
        const size_t want = std::max(h->round_bytes + bytes, h->round_peak);
        if (want > h->stage_bytes)
            if (int rc = ensure_stage(h, want + want / 8 + (1u << 20))) return rc;
    }
    *p = (char *)h->stage + h->stage_cursor;
    h->stage_cursor += bytes;
    h->round_bytes += bytes;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// repack kernels: caller layout -> device records
// ---------------------------------------------------------------------------------------------------------------
// dst[p*dstride + c*dld + roff + i] = src[p*sstride + c*sld + i], i < m, c < cols.  flag |= 1 when col lbcol > col ubcol.
__global__ void repack_level_kernel(double *dst, unsigned long long dstride, int dld, int roff, const double *src,
                                    unsigned long long sstride, int sld, int m, int cols, unsigned long long batch, int lbcol,
                                    int ubcol, int *flag, int bit = 1)
{
    const unsigned long long per = (unsigned long long)m * cols;
    const unsigned long long tot = per * batch;
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < tot;
         e += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long p = e / per;
        const unsigned r           = (unsigned)(e - p * per);
        const int c = r / m, i = r - c * m;
        const double val = src[p * sstride + (unsigned long long)c * sld + i];
        dst[p * dstride + (unsigned long long)c * dld + roff + i] = val;
        if (lbcol >= 0 && c == ubcol)
        {
            const double lb = src[p * sstride + (unsigned long long)lbcol * sld + i];
            if (lb > val) atomicOr(flag, bit);
        }
    }
}

// LexLSI: nz[p * M + roff + i] = (squared norm of row i of the level's A > 0): the "don't activate 0 * x = b" test of setData
// (lexlsi.h:373-377) belongs to set-up, not to the solve; same (S) chain over the columns as the oracle
__global__ void row_nonzero_kernel(const double *obj, unsigned long long ostride, int ldo, int roff, int m, int n, unsigned long long batch,
                                   unsigned char *nz, int M)
{
    const unsigned long long tot = (unsigned long long)m * batch;
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < tot;
         e += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long p = e / m;
        const int i                = (int)(e - p * m);
        const double *row          = obj + p * ostride + roff + i;
        double s                   = 0.0;
        for (int col = 0; col < n; col++)
        {
            const double a = row[(unsigned long long)ldo * col];
            s              = fma(a, a, s);
        }
        nz[p * M + roff + i] = s > 0 ? 1 : 0;
    }
}

__global__ void check_unique_kernel(const unsigned *idx, unsigned long long stride, int m, unsigned nVar, unsigned long long count,
                                    int *flag)
{
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < count;
         p += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned *v = idx + p * stride;
        for (int a = 0; a < m; a++)
        {
            if (v[a] >= nVar) atomicOr(flag, 4);
            for (int b = a + 1; b < m; b++)
                if (v[a] == v[b]) atomicOr(flag, 2);
        }
    }
}

// gather one level of a [B][Mall] result array into a dense [B][m] array
template <typename T>
__global__ void slice_level_kernel(T *dst, const T *src, int Mall, int off, int m, unsigned long long batch)
{
    const unsigned long long tot = (unsigned long long)m * batch;
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < tot;
         e += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long p = e / m;
        const int i                = (int)(e - p * m);
        dst[e]                     = src[p * Mall + off + i];
    }
}
template <typename T>
__global__ void scatter_level_kernel(T *dst, const T *src, int Mall, int off, int m, unsigned long long batch)
{
    const unsigned long long tot = (unsigned long long)m * batch;
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < tot;
         e += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long p = e / m;
        const int i                = (int)(e - p * m);
        dst[p * Mall + off + i]    = src[e];
    }
}

static unsigned blocks_for(unsigned long long work)
{
    unsigned long long b = (work + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148ull * 16) b = 148ull * 16;
    return (unsigned)b;
}

// ---------------------------------------------------------------------------------------------------------------
// planning: shared-memory map and kernel variant
// ---------------------------------------------------------------------------------------------------------------
static unsigned align_up(unsigned v, unsigned a)
{
    return (v + a - 1) / a * a;
}

static SmemMap build_map(const Shape &sh, bool lsi, bool DS, bool LG, bool panel_only = false, bool vecg = false)
{
    SmemMap m{};
    unsigned o = 0;
    auto take  = [&](unsigned bytes, unsigned al) {
        o           = align_up(o, al);
        unsigned at = o;
        o += bytes;
        return at;
    };
    const unsigned n = sh.n, M = sh.M, Mall = sh.Mall, m0 = sh.m0, G = sh.G, P = sh.P;
    // LOD first: it is the target of a TMA bulk copy (LSE) -> 128-byte alignment; size rounded to 16 bytes, followed by
    // 32 bytes of zeros (the register-window Gauss step reads row pairs that may end up to 3 rows past the matrix)
    // (the "work matrix in global memory" variants keep a view of the level being factorized here instead: row pairs of the
    //  largest level + one pair of slack, leading dimension 2 * odd so that lanes over columns spread over the banks)
    unsigned maxdim = 0;
    for (int g = 0; g < sh.G; g++) maxdim = std::max(maxdim, (unsigned)sh.dim[g + sh.off]);
    const unsigned vpairs = ((maxdim + 3) / 2) | 1u;
    m.ldv = (int)(2 * vpairs);
    m.lod = take(LG ? 2 * vpairs * (n + 1) * 8 : align_up((unsigned)sh.ld * (n + 1) * 8, 16) + 32, 128);
    m.obj = take((lsi && DS) ? align_up((unsigned)sh.ldo * (n + 2) * 8, 16) : 16, (lsi && DS) ? 128 : 16);
    m.sb  = take((lsi && DS) ? 2 * m0 * 8 + 16 : 16, 16);
    // x / dx: LexLSI iterates; colnorm / vv: scratch of the shared-memory QR (not needed when every level runs in the
    // register-window panel)
    m.x = take(lsi ? n * 8 : 8, 8);
    m.dx = take(lsi ? n * 8 : 8, 8);
    m.lx = take(n * 8, 8);
    const bool alias_tmp = !lsi && panel_only && n <= LX_PANEL_TMP_DOUBLES && M <= LX_PANEL_TMP_DOUBLES;
    m.lxo = alias_tmp ? 0 : take(n * 8, 8);
    // LexLSI: the column norms of the QR live in the multiplier workspace (wsp is dead while a level is factorized).  The
    // 300 bytes matter: with them C3 fills all 228 KB of an SM (8 x (28,160 + 1,024)) and the repack kernels of the next
    // sub-batch (PipelinedBatch) cannot become resident beside the solve — no copy/solve overlap
    const bool alias_cn = lsi && !panel_only;
    m.colnorm = alias_cn ? 0 : take(panel_only ? 8 : n * 8, 8);
    m.invd = take(n * 8, 8);
    m.hh = take((M + 1) * 8, 8);
    m.wsp = alias_tmp ? 0 : take((2 * m0 + M + n + 2) * 8, 8);
    if (alias_cn) m.colnorm = m.wsp;
    const unsigned vecs = (lsi && !vecg) ? Mall + 1 : 1; // vecg: the four iteration vectors live in a global scratch slice
    m.v = take(vecs * 8, 8);
    m.dv = take(vecs * 8, 8);
    m.Ax = take(vecs * 8, 8);
    m.Adx = take(vecs * 8, 8);
    m.vv = take(panel_only ? 16 : ((unsigned)sh.ld + 2) * 8, 16);
    m.ptmp = take(panel_only ? LX_PANEL_TMP_DOUBLES * 8 : 16, 16); // staging area of the register-window panel only
    if (alias_tmp) m.lxo = m.wsp = m.ptmp; // LexLSE: un-permute buffer (solve) and residuals (get_v) reuse the panel's staging area
    m.active = take(lsi ? (Mall + 1) * 2 : 4, 4);
    m.inactive = take(lsi ? (Mall + 1) * 2 : 4, 4);
    m.fidx = take((std::max(m0, (unsigned)sh.nfix) + 1) * 2, 4);
    m.colperm = take((n + 1) * 2, 4);
    m.varofcol = take((n + 1) * 2, 4);
    m.sbidx = take((m0 + 1) * 2, 4);
    m.atype = take(lsi ? Mall + 1 : 4, 4);
    m.alltype = take(lsi ? Mall + 1 : 4, 4);
    m.ctype = take(lsi ? M + 1 : 4, 4);
    m.ftype = take(std::max(m0, (unsigned)sh.nfix) + 1, 4);
    m.ldim = take((G + 1) * 4, 4);
    m.frow = take((G + 1) * 4, 4);
    m.fcol = take((G + 1) * 4, 4);
    m.rank = take((G + 1) * 4, 4);
    m.nact = take((P + 1) * 4, 4);
    m.ninact = take((P + 1) * 4, 4);
    m.scal = take(8 * 4, 8);
    m.mbar = take(8, 8);
    m.total = align_up(o, 128);
    m.vecg  = (lsi && vecg) ? 1 : 0;
    return m;
}

template <typename K> static int occupancy_of(K kernel, int threads, unsigned smem, int *per_sm)
{
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kernel, threads, smem));
    return 0;
}


// the solver kernels are instantiated in their own translation units (kernels_*.cu), compiled in parallel
namespace lexls_dev
{
kernel_fn pick_lse_kernel(int W, bool LG);
kernel_fn pick_lse_pair_kernel();
kernel_fn pick_lsi_w1_kernel(bool DS);
kernel_fn pick_lsi_wn_kernel(int W, bool DS, bool LG);
kernel_fn pick_lse_reg_kernel(int W, bool LG);
kernel_fn pick_lsi_reg_kernel(int W, bool LG);
} // namespace lexls_dev

static kernel_fn pick_kernel(int kind, int W, bool DS, bool LG, std::string &name, bool reg = false)
{
    char buf[64];
    if (reg) // separate instantiations with the regularization routines (W >= 2, objective data in L2)
    {
        snprintf(buf, sizeof buf, "%s_w%d%s_reg", kind == LEXLS_B200_LSE ? "lse" : "lsi", W, LG ? "_lodglobal" : (kind == LEXLS_B200_LSE ? "" : "_l2data"));
        name = buf;
        return kind == LEXLS_B200_LSE ? pick_lse_reg_kernel(W, LG) : pick_lsi_reg_kernel(W, LG);
    }
    if (kind == LEXLS_B200_LSE)
    {
        snprintf(buf, sizeof buf, "lse_w%d%s", W, LG ? "_lodglobal" : "");
        name = buf;
        return pick_lse_kernel(W, LG);
    }
    snprintf(buf, sizeof buf, "lsi_w%d%s", W, LG ? "_lodglobal" : (DS ? "_smemdata" : "_l2data"));
    name = buf;
    if (!LG && W == 1) return pick_lsi_w1_kernel(DS);
    return pick_lsi_wn_kernel(W, DS, LG);
}

// cudaGetDeviceProperties costs ~1 ms: cached per device (plan() runs again whenever the shape of a handle changes)
static int device_props(int device, const cudaDeviceProp **out)
{
    static cudaDeviceProp props[64];
    static bool have[64] = {false};
    if (device < 0 || device >= 64) return fail(LEXLS_B200_EINVAL, "bad device ordinal");
    if (!have[device])
    {
        CK(cudaGetDeviceProperties(&props[device], device));
        have[device] = true;
    }
    *out = &props[device];
    return 0;
}

static int plan(lexls_b200_batch *h)
{
    if (h->planned) return 0;
    const cudaDeviceProp *propp = nullptr;
    if (int rc = device_props(h->device, &propp)) return rc;
    const cudaDeviceProp &prop = *propp;
    const unsigned max_smem = (unsigned)prop.sharedMemPerBlockOptin;
    const bool lsi          = h->kind == LEXLS_B200_LSI;
    const int cols          = h->sh.n + 1;

    // LexLSE: one warp per problem (register-window panel).  LexLSI: the active-set loop is latency bound and its code
    // has to stay small (32 KB instruction cache), so two warps per problem on the compact shared-memory routines with
    // the objective data read through L2 keep more warps resident (measured on C3: 18.5 ms vs 29.8 ms, profiles/).
    // Wider problems: the reflector is applied with one thread per column, so the group gets a thread for every column
    // (residency is set by the work matrix in shared memory, not by the warps: test_01.dat, 89 columns, 84 -> 76 ms with 4)
    int W = cols <= 64 ? (lsi ? 2 : 1) : (cols <= 128 ? 4 : 8);
    if (const char *e = getenv("LEXLS_B200_WARPS"))
    {
        int w = atoi(e);
        if (w == 1 || w == 2 || w == 4 || w == 8) W = w;
        if (h->user_params.regularization_type != 0 && W == 1) W = 2;
    }
    bool DS = false, LG = false;
    if (const char *e = getenv("LEXLS_B200_DATA"))
    {
        if (!strcmp(e, "smem")) DS = lsi && h->user_params.regularization_type == 0; // the regularized kernels read the data through L2
    }
    // every level handled by the register-window panel (lexls_panel.cuh): <= 48 columns, <= 40 rows per level
    if (h->user_params.regularization_type != 0 && W == 1) W = 2; // the regularization routine lives in the shared-memory kernels
    bool panel_only = (W == 1) && cols <= 48;
    for (int k = 0; k < h->sh.P; k++)
        if (h->sh.dim[k] > 40 && !(lsi && h->sh.off && k == 0)) panel_only = false;
    if (W == 1 && !panel_only) W = 2; // the W = 1 kernels have no shared-memory QR (see lse_factorize)
    SmemMap m = build_map(h->sh, lsi, DS, LG, panel_only);
    // keep the objective data in shared memory only while at least 3 problems stay resident per SM
    if (lsi && DS && m.total * 3 > max_smem)
    {
        DS = false;
        m  = build_map(h->sh, lsi, DS, LG, panel_only);
    }
    // Residency decides: every variant is latency bound (one serial pivot chain per problem), so the warps per problem are
    // chosen to keep the most problems resident per SM — shared memory (the work matrix) and registers (<= 128 per thread
    // for the LexLSE kernels, 96 for the LexLSI kernels) bound that; on a tie more warps win (two threads per column in the
    // reflector apply), up to two threads per column.  Measured with scripts/shape_perf.py, ms per batch:
    //   LexLSE 60/60/60 x 41 (3 per SM either way): 2 warps 4.44, 4 warps 4.04;   20 x 101 (8 vs 4 per SM): 0.20 vs 0.32;
    //   LexLSE 30 x 151: 2 warps (5 per SM) 0.58, 4 warps (4) 0.64, 8 warps (2) 1.02;   40/40/40 x 201 (1 per SM): 8 warps 4.2, 4 warps 5.1;
    //   LexLSI 60/60/60 x 41: 68 vs 62;   LexLSI 30 x 141 (5 per SM either way): 4.37 vs 4.28, 8 warps (2 per SM) 7.2.
    const unsigned sm_smem = (unsigned)prop.sharedMemPerMultiprocessor;
    auto fit = [&](const SmemMap &mm) { return sm_smem / (mm.total + 1024u); };
    const bool w_forced = getenv("LEXLS_B200_WARPS") != nullptr;
    if (!w_forced && !panel_only && m.total <= max_smem)
    {
        const unsigned by_smem = fit(m);
        int wmax = 2;
        while (wmax < 8 && 32 * wmax < 2 * cols) wmax *= 2; // two threads per column are the most the apply can use
        unsigned best = 0;
        for (int w = 2; w <= wmax; w *= 2)
        {
            const unsigned by_regs = lsi ? (w == 2 ? 10u : (w == 4 ? 5u : 2u)) : (w == 2 ? 8u : (w == 4 ? 4u : 2u));
            const unsigned r       = by_smem < by_regs ? by_smem : by_regs;
            if (r >= best)
            {
                best = r;
                W    = w;
            }
        }
    }
    if (lsi && !LG && getenv("LEXLS_B200_VECG"))
    {
        // experiment knob (measured: no gain on C3, see lexls_device.cuh at LX_LSI_W2_MINB): moving v, dv, Ax, Adx out of shared
        // memory when it lets one more problem become resident per SM
        const SmemMap mg = build_map(h->sh, lsi, DS, LG, panel_only, true);
        if (m.total <= max_smem && fit(mg) > fit(m)) m = mg;
    }
    bool auto_lg = false;
    // (a batch that fits one wave of the shared-memory kernel stays there: 2.7 vs 3.0 ms for one test_01.dat problem)
    if (m.total <= max_smem && fit(m) <= 1 && h->batch > (uint64_t)prop.multiProcessorCount && !getenv("LEXLS_B200_NO_AUTO_LODGLOBAL"))
    {
        const SmemMap mlg = build_map(h->sh, lsi, false, true);
        auto_lg           = mlg.total <= max_smem && fit(mlg) >= 2;
    }
    if (m.total > max_smem || auto_lg || getenv("LEXLS_B200_FORCE_LODGLOBAL"))
    {
        LG = true;
        DS = false;
        W  = 16;
        m  = build_map(h->sh, lsi, DS, LG);
        if (m.total > max_smem) return fail(LEXLS_B200_ECAPACITY, "problem shape exceeds the shared memory of every kernel variant");
        // the pivot chain is latency bound: two 8-warp groups per SM (two problems in flight) beat one 16-warp group
        // whenever two copies of the level view fit (1 KB per CTA is reserved by the system)
        if (2ull * (m.total + 1024) <= (unsigned long long)prop.sharedMemPerMultiprocessor) W = 8;
        if (const char *e = getenv("LEXLS_B200_LG_WARPS"))
        {
            const int w = atoi(e);
            if (w == 8 || w == 16) W = w;
        }
    }
    h->W  = W;
    h->DS = DS;
    h->LG = LG;
    h->sm = m;
    h->smem = m.total;
    // two LexLSE problems per warp (lexls_pair.cuh): no fixed variables, register-window shapes with <= 20 rows per level
    h->pair = false;
    if (!lsi && !LG && W == 1 && panel_only && h->sh.nfix == 0 && getenv("LEXLS_B200_PAIR") && pick_lse_pair_kernel())
    {
        h->pair = true;
        for (int k = 0; k < h->sh.P; k++)
            if (h->sh.dim[k] > 20) h->pair = false;
    }
    kernel_fn k = pick_kernel(h->kind, W, DS, LG, h->kname, h->user_params.regularization_type != 0);
    if (h->pair)
    {
        k        = pick_lse_pair_kernel();
        h->kname = "lse_pair";
        h->smem  = 2 * m.total;
    }
    int per_sm  = 0;
    if (int rc = occupancy_of(k, 32 * W, h->smem, &per_sm)) return rc;
    if (per_sm < 1) return fail(LEXLS_B200_ECAPACITY, "kernel does not fit on an SM");
    if (const char *e = getenv("LEXLS_B200_PER_SM"))
    {
        // experiment knob (profiles/r2_experiments.json): throughput against the number of resident problems per SM — the
        // dynamic shared memory is padded until the occupancy drops to the requested value
        const int cap = atoi(e);
        while (cap >= 1 && per_sm > cap && h->smem + 1024 <= max_smem)
        {
            h->smem += 1024;
            if (int rc = occupancy_of(k, 32 * W, h->smem, &per_sm)) return rc;
        }
    }
    h->per_sm = per_sm;
    unsigned long long g = (unsigned long long)prop.multiProcessorCount * per_sm;
    if (g > h->batch) g = h->batch;
    if (LG)
        if (const char *e = getenv("LEXLS_B200_LG_GRID"))
        {
            // experiment knob (profiles/r2_c4_traffic.json): fewer problems in flight so that their work matrices and objective
            // records stay L2-resident (1.3 MB per C4 problem, 126 MB of L2) — less DRAM traffic, less latency hiding
            const unsigned long long cap = strtoull(e, nullptr, 10);
            if (cap >= 1 && cap < g) g = cap;
        }
    if (g < 1) g = 1;
    h->grid = (unsigned)g;
    if (LG)
    {
        h->lod_stride = (uint64_t)h->sh.ld * (h->sh.n + 1);
        const size_t need = sizeof(double) * h->lod_stride * h->grid;
        if (need > h->lod_bytes)
        {
            if (h->lod_scratch) guarded_free(h, h->lod_scratch);
            h->lod_scratch = nullptr;
            h->lod_bytes   = 0;
            CK(guarded_malloc(h, (void **)&h->lod_scratch, need));
            h->lod_bytes = need;
        }
    }
    if (h->user_params.regularization_type != 0)
    {
        const uint64_t n = (uint64_t)h->sh.n;
        h->reg_stride    = 2 * n * (n + 1) + 12 * n + 8; // Z, D, d, two work vectors, nine for the conjugate-gradient types
        const size_t need = sizeof(double) * h->reg_stride * h->grid;
        if (need > h->reg_bytes)
        {
            if (h->reg_scratch) guarded_free(h, h->reg_scratch);
            h->reg_scratch = nullptr;
            h->reg_bytes   = 0;
            CK(guarded_malloc(h, (void **)&h->reg_scratch, need));
            h->reg_bytes = need;
        }
    }
    h->planned = true;
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// multi-device batches: contiguous shards, one handle + stream per device, no inter-device traffic.
// Blocking calls (host copies in, results out) run one host thread per device, so the copies of all devices are in
// flight together and land in disjoint slices of the caller's (ideally pinned) buffer.
// ---------------------------------------------------------------------------------------------------------------
template <class F> static int fan_out(lexls_b200_batch *m, bool threaded, F f)
{
    const size_t n = m->shards.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    auto run = [&](size_t i) {
        rc[i] = f(m->shards[i]);
        if (rc[i]) msg[i] = g_err; // thread-local message of the worker
    };
    if (threaded && n > 1)
    {
        std::vector<std::thread> t;
        for (size_t i = 0; i < n; i++) t.emplace_back(run, i);
        for (auto &x : t) x.join();
    }
    else
        for (size_t i = 0; i < n; i++) run(i);
    for (size_t i = 0; i < n; i++)
        if (rc[i]) return fail(rc[i], "device " + std::to_string(m->shards[i].h->device) + ": " + msg[i]);
    return 0;
}
#define MULTI(h, threaded, expr)                                                                                        \
    if ((h) && (h)->multi()) return fan_out((h), (threaded), [&](const Shard &s) -> int { return (expr); })
#define MULTI_REFUSE(h, what)                                                                                           \
    if ((h) && (h)->multi()) return fail(LEXLS_B200_EINVAL, what " is a single-device call: use lexls_b200_get_shard")
template <typename T> static const T *off(const T *p, uint64_t n)
{
    return p ? p + n : p;
}
template <typename T> static T *off(T *p, uint64_t n)
{
    return p ? p + n : p;
}

// ---------------------------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int lexls_b200_abi_version(void)
{
    return LEXLS_B200_ABI_VERSION;
}
const char *lexls_b200_last_error(void)
{
    return g_err.c_str();
}
void lexls_b200_default_params(lexls_b200_params_t *p)
{
    p->max_number_of_factorizations = 200;
    p->tol_linear_dependence        = 1e-12;
    p->tol_wrong_sign_lambda        = 1e-08;
    p->tol_correct_sign_lambda      = 1e-12;
    p->tol_feasibility              = 1e-13;
    p->cycling_handling_enabled     = 0;
    p->cycling_max_counter          = 50;
    p->cycling_relax_step           = 1e-08;
    p->regularization_type          = 0;
    p->modify_x_guess_enabled       = 0;
    p->modify_type_active_enabled   = 0;
    p->modify_type_inactive_enabled = 0;
    p->set_min_init_ctr_violation   = 1;
    p->use_phase1_v0                = 0;
    p->log_working_set_enabled      = 0;
    p->deactivate_first_wrong_sign  = 0;
    p->variable_regularization_factor = 0.0;
    p->max_number_of_cg_iterations    = 10; // typedefs.h:121, :282
}
int lexls_b200_device_count(int *count)
{
    int c         = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess)
    {
        *count = 0;
        return fail(LEXLS_B200_ECUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = c;
    return 0;
}

static void to_dev_params(const lexls_b200_params_t &u, Params &p)
{
    p.max_fact     = u.max_number_of_factorizations;
    p.tol_lindep   = u.tol_linear_dependence;
    p.tol_wrong    = u.tol_wrong_sign_lambda;
    p.tol_correct  = u.tol_correct_sign_lambda;
    p.tol_feas     = u.tol_feasibility;
    p.cycling      = u.cycling_handling_enabled;
    p.cyc_max      = u.cycling_max_counter;
    p.cyc_step     = u.cycling_relax_step;
    p.mod_x        = u.modify_x_guess_enabled;
    p.mod_active   = u.modify_type_active_enabled;
    p.mod_inactive = u.modify_type_inactive_enabled;
    p.min_viol     = u.set_min_init_ctr_violation;
    p.phase1_v0    = u.use_phase1_v0;
    p.dfirst       = u.deactivate_first_wrong_sign;
    p.reg_type     = u.regularization_type;
    p.reg_var      = u.variable_regularization_factor;
    p.reg_cg_iter  = u.max_number_of_cg_iterations;
    // Gauss step on the FP64 tensor cores (mma.sync m8n8k4, bit-identical to the DFMA chain: profiles/r2_dmma_order.json);
    // LEXLS_B200_DMMA=0 selects the DFMA version for A/B measurements
    const char *e = getenv("LEXLS_B200_DMMA");
    p.dmma        = e ? (atoi(e) != 0) : 1; // only the global-work-matrix kernels contain the tensor-core tiles
}

int lexls_b200_create_batch(int kind, uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType, uint64_t batch,
                            int device, lexls_b200_batch_t **out)
{
    if (!out) return fail(LEXLS_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (kind != LEXLS_B200_LSE && kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "unknown solver kind");
    if (nVar == 0 || nObj == 0 || !ObjDim) return fail(LEXLS_B200_EINVAL, "nVar, nObj and ObjDim are required");
    if (nObj > LX_MAX_LEVELS) return fail(LEXLS_B200_ECAPACITY, "more than 16 levels");
    if (batch == 0) return fail(LEXLS_B200_EINVAL, "empty batch");
    if (nVar > 60000) return fail(LEXLS_B200_ECAPACITY, "nVar too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LEXLS_B200_ECUDA, "no CUDA device available: lexls_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(LEXLS_B200_EINVAL, "bad device ordinal");
    for (uint32_t k = 0; k < nObj; k++)
    {
        int t = ObjType ? ObjType[k] : 0;
        if (t != 0 && t != 1) return fail(LEXLS_B200_EINVAL, "Unknown objective type");
        if (t == 1 && k != 0) return fail(LEXLS_B200_EINVAL, "only the first objective may be SIMPLE_BOUNDS");
        if (t == 1 && kind == LEXLS_B200_LSE) return fail(LEXLS_B200_EINVAL, "LexLSE has no simple-bounds objective; use set_fixed_variables");
    }
    CK(cudaSetDevice(device));
    lexls_b200_batch *h = new lexls_b200_batch();
    h->kind   = kind;
    h->device = device;
    h->batch  = batch;
    Shape &s  = h->sh;
    s.kind    = kind;
    s.n       = (int)nVar;
    s.P       = (int)nObj;
    s.off     = (ObjType && ObjType[0] == 1) ? 1 : 0;
    s.G       = s.P - s.off;
    s.m0      = s.off ? (int)ObjDim[0] : 0;
    s.M       = 0;
    s.uoff[0] = 0;
    for (int k = 0; k < s.P; k++)
    {
        s.dim[k]      = (int)ObjDim[k];
        s.uoff[k + 1] = s.uoff[k] + s.dim[k];
    }
    s.goff[0] = 0;
    for (int g = 0; g < s.G; g++) s.goff[g + 1] = s.goff[g] + s.dim[g + s.off];
    s.M    = s.goff[s.G];
    s.Mall = s.m0 + s.M;
    if (s.Mall > 60000) { delete h; return fail(LEXLS_B200_ECAPACITY, "too many constraints"); }
    s.ld   = (std::max(s.M, 2) + 1) & ~1; // even: every column starts 16-byte aligned (128-bit row-pair accesses)
    {
        // the shared-memory QR also wants ld/2 odd (lanes over columns conflict-free); the register-window LexLSE path
        // (<= 48 columns, <= 40 rows per level) does not read LOD that way and prefers the smaller footprint
        bool panel = kind == LEXLS_B200_LSE && nVar + 1 <= 48;
        for (int k = 0; k < s.P; k++)
            if (s.dim[k] > 40) panel = false;
        if (!panel && ((s.ld >> 1) & 1) == 0) s.ld += 2;
    }
    s.nfix = 0;
    if (kind == LEXLS_B200_LSE)
    {
        s.ldo         = s.ld;
        h->rec_stride = ((uint64_t)s.ld * (nVar + 1) + 1) & ~1ull; // whole doubles, 16-byte multiple
    }
    else
    {
        s.ldo         = std::max(s.M, 1);
        h->obj_stride = ((uint64_t)s.ldo * (nVar + 2) + 1) & ~1ull;
    }
    lexls_b200_default_params(&h->user_params);
    to_dev_params(h->user_params, h->pr);
    h->have_v0.assign(s.P, 0);
    h->data_set.assign(s.P, 0);

    auto bail = [&](int rc) {
        lexls_b200_destroy(h);
        return rc;
    };
#define CKB(call)                                                                                                       \
    do                                                                                                                  \
    {                                                                                                                   \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) return bail(fail(LEXLS_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
    } while (0)
    CKB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CKB(cudaEventCreate(&h->ev0));
    CKB(cudaEventCreate(&h->ev1));
    CKB(cudaEventCreateWithFlags(&h->ev_switch, cudaEventDisableTiming));
    CKB(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    CKB(cudaMallocHost(&h->hflag, sizeof(int)));
    *h->hflag = 0;
    const uint64_t B = batch;
    if (kind == LEXLS_B200_LSE)
    {
        CKB(guarded_malloc(h, (void **)&h->rec, sizeof(double) * h->rec_stride * B));
        CKB(cudaMemsetAsync(h->rec, 0, sizeof(double) * h->rec_stride * B, h->stream));
        CKB(guarded_malloc(h, (void **)&h->v, sizeof(double) * std::max(s.M, 1) * B));
    }
    else
    {
        CKB(guarded_malloc(h, (void **)&h->obj, sizeof(double) * h->obj_stride * B));
        CKB(cudaMemsetAsync(h->obj, 0, sizeof(double) * h->obj_stride * B, h->stream));
        CKB(guarded_malloc(h, (void **)&h->rownz, (size_t)std::max(s.M, 1) * B));
        CKB(cudaMemsetAsync(h->rownz, 0, (size_t)std::max(s.M, 1) * B, h->stream));
        if (s.m0)
        {
            CKB(guarded_malloc(h, (void **)&h->sb, sizeof(double) * 2 * s.m0 * B));
            CKB(guarded_malloc(h, (void **)&h->sbidx, sizeof(unsigned) * s.m0 * B));
        }
        CKB(guarded_malloc(h, (void **)&h->v, sizeof(double) * s.Mall * B));
        CKB(guarded_malloc(h, (void **)&h->viol, sizeof(double) * s.Mall * B));
        CKB(guarded_malloc(h, (void **)&h->act, s.Mall * B));
        CKB(guarded_malloc(h, (void **)&h->order, sizeof(unsigned) * s.Mall * B));
    }
    CKB(guarded_malloc(h, (void **)&h->x, sizeof(double) * nVar * B));
    CKB(guarded_malloc(h, (void **)&h->status, sizeof(int) * B));
    CKB(guarded_malloc(h, (void **)&h->counters, sizeof(unsigned) * 8 * B));
    CKB(guarded_malloc(h, (void **)&h->ranks, sizeof(unsigned) * std::max(s.G, 1) * B));
    CKB(guarded_malloc(h, (void **)&h->colperm, sizeof(unsigned) * nVar * B));
    CKB(guarded_malloc(h, (void **)&h->queue, sizeof(unsigned long long)));
    CKB(guarded_malloc(h, (void **)&h->flag, sizeof(int)));
    CKB(cudaMemsetAsync(h->flag, 0, sizeof(int), h->stream));
    CKB(cudaMemsetAsync(h->status, 0xff, sizeof(int) * B, h->stream));
    CKB(cudaMemsetAsync(h->counters, 0, sizeof(unsigned) * 8 * B, h->stream));
    CKB(cudaMemsetAsync(h->x, 0, sizeof(double) * nVar * B, h->stream));
#undef CKB
    *out = h;
    return 0;
}

int lexls_b200_destroy(lexls_b200_batch_t *h)
{
    if (h && h->multi())
    {
        for (auto &s : h->shards) lexls_b200_destroy(s.h);
        delete h;
        return 0;
    }
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    void *ptrs[] = {h->rownz, h->rec, h->obj, h->sb, h->fval, h->x0, h->v0, h->sbidx, h->fidx, h->ftype, h->guess, h->x, h->v, h->viol,
                    h->lod_scratch, h->act, h->order, h->counters, h->ranks, h->colperm, h->status, h->flag, h->queue, h->stage,
                    h->lexqr, h->hhout, h->lambda, h->wslog, h->wslog_count, h->xstar, h->lexdata, h->lexqr_star, h->hh_star, h->glist, h->wsl, h->reg_scratch, h->xln, h->ln_scratch, h->vec_scratch};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_switch) cudaEventDestroy(h->ev_switch);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->hflag) cudaFreeHost(h->hflag);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

int lexls_b200_reset(lexls_b200_batch_t *h)
{
    MULTI(h, false, lexls_b200_reset(s.h));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    if (int rc = flush_pending(h)) return rc;
    h->have_x0    = false;
    h->have_guess = false;
    h->have_glist = false;
    std::fill(h->have_v0.begin(), h->have_v0.end(), 0);
    if (h->guess) CK(cudaMemsetAsync(h->guess, 0, (size_t)h->sh.Mall * h->batch, h->stream));
    CK(cudaMemsetAsync(h->status, 0xff, sizeof(int) * h->batch, h->stream));
    CK(cudaMemsetAsync(h->counters, 0, sizeof(unsigned) * 8 * h->batch, h->stream));
    CK(cudaMemsetAsync(h->x, 0, sizeof(double) * h->sh.n * h->batch, h->stream));
    return 0;
}

int lexls_b200_set_params(lexls_b200_batch_t *h, const lexls_b200_params_t *p)
{
    MULTI(h, false, lexls_b200_set_params(s.h, p));
    if (!h || !p) return fail(LEXLS_B200_EINVAL, "NULL argument");
    {
        const int t = p->regularization_type; // typedefs.h:32-44
        if (t < 0 || t > 9) return fail(LEXLS_B200_EINVAL, "unknown regularization type");
        if ((h->user_params.regularization_type != 0) != (t != 0)) h->planned = false; // regularized solves run on the shared-memory routines
    }
    if (p->deactivate_first_wrong_sign && h->kind == LEXLS_B200_LSI && h->sh.off)
        return fail(LEXLS_B200_EUNSUPPORTED,
                    "deactivate_first_wrong_sign with a simple-bounds objective: the reference tests the wrong multipliers and indexes out "
                    "of bounds there (lexlse.h:596-599, lexlsi.h:1092), so there is no behaviour to reproduce");
    if (p->max_number_of_factorizations == 0) return fail(LEXLS_B200_EINVAL, "max_number_of_factorizations must be positive");
    h->user_params = *p;
    to_dev_params(*p, h->pr);
    return 0;
}

int lexls_b200_set_stream(lexls_b200_batch_t *h, void *cuda_stream)
{
    MULTI_REFUSE(h, "set_stream");
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    if (next != h->stream)
    {
        if (int rc = flush_pending(h)) return rc; // deferred repacks belong behind their copies, on the old stream
        // work already enqueued for this handle (the zero fill of create_batch, earlier set_* calls or solves) must be ordered
        // before whatever the caller enqueues on the new stream
        CK(cudaSetDevice(h->device));
        CK(cudaEventRecord(h->ev_switch, h->stream));
        CK(cudaStreamWaitEvent(next, h->ev_switch, 0));
        h->stream = next;
    }
    return 0;
}

// Input validation is done on the device by the repack kernels, which OR bits into h->flag: 1 general lb > ub, 8 simple
// lb > ub, 2 duplicate VarIndex, 4 VarIndex >= nVar.  The set_* calls do not wait for it.  The flag is inspected
//   * by lexls_b200_validate (synchronises; what the batch-of-1 facades call to throw where the reference throws), and
//   * after a solve: solve_batch enqueues a copy of the flag into pinned host memory in front of the kernel (which returns
//     at once, status UNKNOWN, when the flag is set); the first synchronising call afterwards reports it.
// So a batch costs ONE synchronisation (the result read), not one per set_* call.
static int flag_error(lexls_b200_batch *h, int flag)
{
    if (!flag) return 0;
    std::fill(h->data_set.begin(), h->data_set.end(), 0); // the offending data must be set again
    if (flag & 1) return fail(LEXLS_B200_EBOUNDS, "(general) Lower bound is greater than upper bound.");
    if (flag & 8) return fail(LEXLS_B200_EBOUNDS, "(simple) Lower bound is greater than upper bound.");
    if (flag & 2) return fail(LEXLS_B200_EINVAL, "Elements of VarIndex are not unique.");
    return fail(LEXLS_B200_EINVAL, "VarIndex >= nVar");
}
// after the stream has been synchronised: report (once) what the flag copy of the last solve says
static int check_pending_flag(lexls_b200_batch *h)
{
    if (!h->flag_pending) return 0;
    h->flag_pending = false;
    const int flag  = *h->hflag;
    if (!flag) return 0;
    *h->hflag = 0;
    cudaMemsetAsync(h->flag, 0, sizeof(int), h->stream);
    return flag_error(h, flag);
}
// the source is a temporary of the library: its copy must have left it in every copy mode
static int wait_copy_of_temporary(lexls_b200_batch *h)
{
    CK(cudaEventRecord(h->ev_copy, h->stream));
    CK(cudaEventSynchronize(h->ev_copy));
    return 0;
}
// host inputs: returns once the caller's buffer has been consumed (copied into the staging buffer)
static int wait_copy(lexls_b200_batch *h)
{
    if (h->async_copies) return 0; // the caller keeps its buffers valid until the next synchronisation
    CK(cudaEventRecord(h->ev_copy, h->stream));
    CK(cudaEventSynchronize(h->ev_copy));
    return 0;
}

static int set_data_impl(lexls_b200_batch *h, uint32_t level, const double *data, uint32_t ld, uint64_t stride, bool on_device)
{
    if (!h || !data) return fail(LEXLS_B200_EINVAL, "NULL argument");
    const Shape &s = h->sh;
    if ((int)level >= s.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    if (s.off && level == 0) return fail(LEXLS_B200_EINVAL, "ObjType = GENERAL_OBJECTIVE is assumed (use set_simple_bounds)");
    const int m = s.dim[level], g = (int)level - s.off;
    const bool lsi = h->kind == LEXLS_B200_LSI;
    const int cols = s.n + (lsi ? 2 : 1);
    if (m == 0)
    {
        h->data_set[level] = 1;
        return 0;
    }
    if ((int)ld < m) return fail(LEXLS_B200_EINVAL, "Incorrect number of equations (ld < ObjDim)");
    if (h->batch > 1 && stride < (uint64_t)ld * (cols - 1) + m) return fail(LEXLS_B200_EINVAL, "problem_stride too small");
    CK(cudaSetDevice(h->device));
    const double *src = data;
    if (!on_device)
    {
        const size_t bytes = sizeof(double) * (stride * (h->batch - 1) + (uint64_t)ld * (cols - 1) + m);
        void *slot = nullptr;
        if (int rc = stage_take(h, bytes, &slot)) return rc;
        CK(cudaMemcpyAsync(slot, data, bytes, cudaMemcpyHostToDevice, h->stream));
        if (int rc = wait_copy(h)) return rc;
        src = (const double *)slot;
    }
    else if (int rc = flush_pending(h))
        return rc;
    h->data_set[level] = 1; // lb > ub is reported by lexls_b200_validate / after the solve (see flag_error)
    const unsigned long long work = (unsigned long long)m * cols * h->batch;
    return run_or_defer(h, [=]() -> int {
        if (lsi)
        {
            repack_level_kernel<<<blocks_for(work), 256, 0, h->stream>>>(h->obj, h->obj_stride, h->sh.ldo, h->sh.goff[g], src, stride,
                                                                         (int)ld, m, cols, h->batch, h->sh.n, h->sh.n + 1, h->flag);
            row_nonzero_kernel<<<blocks_for((unsigned long long)m * h->batch), 256, 0, h->stream>>>(h->obj, h->obj_stride, h->sh.ldo, h->sh.goff[g], m,
                                                                                                  h->sh.n, h->batch, h->rownz, h->sh.M);
        }
        else
            repack_level_kernel<<<blocks_for(work), 256, 0, h->stream>>>(h->rec, h->rec_stride, h->sh.ld, h->sh.goff[g], src, stride,
                                                                         (int)ld, m, cols, h->batch, -1, -1, h->flag);
        CK(cudaGetLastError());
        return 0;
    });
}

int lexls_b200_set_data(lexls_b200_batch_t *h, uint32_t level, const double *data, uint32_t ld, uint64_t problem_stride)
{
    MULTI(h, true, lexls_b200_set_data(s.h, level, off(data, s.first * problem_stride), ld, problem_stride));
    return set_data_impl(h, level, data, ld, problem_stride, false);
}
int lexls_b200_set_data_dev(lexls_b200_batch_t *h, uint32_t level, const double *data, uint32_t ld, uint64_t problem_stride)
{
    MULTI_REFUSE(h, "set_data_dev");
    return set_data_impl(h, level, data, ld, problem_stride, true);
}

static int set_sb_impl(lexls_b200_batch *h, const unsigned *var_index, uint64_t index_stride, const double *lbub, uint32_t ld,
                       uint64_t stride, bool on_device)
{
    if (!h || !var_index || !lbub) return fail(LEXLS_B200_EINVAL, "NULL argument");
    const Shape &s = h->sh;
    if (h->kind != LEXLS_B200_LSI || !s.off) return fail(LEXLS_B200_EINVAL, "ObjType = SIMPLE_BOUNDS_OBJECTIVE is assumed");
    const int m = s.m0;
    if (m == 0)
    {
        h->data_set[0] = 1;
        return 0;
    }
    if ((int)ld < m) return fail(LEXLS_B200_EINVAL, "Incorrect number of equations (ld < ObjDim)");
    if (index_stride != 0 && index_stride < (uint64_t)m) return fail(LEXLS_B200_EINVAL, "index_stride too small");
    CK(cudaSetDevice(h->device));
    const uint64_t nidx = index_stride ? (index_stride * (h->batch - 1) + m) : m;
    const double *src   = lbub;
    const unsigned *isrc = var_index;
    if (!on_device)
    {
        const size_t dbytes = sizeof(double) * (stride * (h->batch - 1) + (uint64_t)ld + m);
        const size_t ibytes = sizeof(unsigned) * nidx;
        void *slot = nullptr;
        if (int rc = stage_take(h, dbytes + ibytes + 16, &slot)) return rc;
        CK(cudaMemcpyAsync(slot, lbub, dbytes, cudaMemcpyHostToDevice, h->stream));
        char *ip = (char *)slot + ((dbytes + 15) & ~(size_t)15);
        CK(cudaMemcpyAsync(ip, var_index, ibytes, cudaMemcpyHostToDevice, h->stream));
        if (int rc = wait_copy(h)) return rc;
        src  = (const double *)slot;
        isrc = (const unsigned *)ip;
    }
    else if (int rc = flush_pending(h))
        return rc;
    h->sbidx_stride = index_stride ? m : 0;
    h->data_set[0]  = 1; // lb > ub / duplicate indices are reported by lexls_b200_validate / after the solve (see flag_error)
    const unsigned nvar = (unsigned)s.n;
    return run_or_defer(h, [=]() -> int {
        repack_level_kernel<<<blocks_for((unsigned long long)m * 2 * h->batch), 256, 0, h->stream>>>(h->sb, 2ull * m, m, 0, src, stride,
                                                                                                     (int)ld, m, 2, h->batch, 0, 1, h->flag, 8);
        CK(cudaGetLastError());
        if (index_stride)
            CK(cudaMemcpy2DAsync(h->sbidx, sizeof(unsigned) * m, isrc, sizeof(unsigned) * index_stride, sizeof(unsigned) * m, h->batch,
                                 cudaMemcpyDeviceToDevice, h->stream));
        else
            CK(cudaMemcpyAsync(h->sbidx, isrc, sizeof(unsigned) * m, cudaMemcpyDeviceToDevice, h->stream));
        const unsigned long long cnt = index_stride ? h->batch : 1;
        check_unique_kernel<<<blocks_for(cnt), 256, 0, h->stream>>>(h->sbidx, h->sbidx_stride, m, nvar, cnt, h->flag);
        CK(cudaGetLastError());
        return 0;
    });
}
int lexls_b200_set_simple_bounds(lexls_b200_batch_t *h, const uint32_t *var_index, uint64_t index_stride, const double *lbub,
                                 uint32_t ld, uint64_t problem_stride)
{
    MULTI(h, true, lexls_b200_set_simple_bounds(s.h, off(var_index, s.first * index_stride), index_stride, off(lbub, s.first * problem_stride), ld, problem_stride));
    return set_sb_impl(h, var_index, index_stride, lbub, ld, problem_stride, false);
}
int lexls_b200_set_simple_bounds_dev(lexls_b200_batch_t *h, const uint32_t *var_index, uint64_t index_stride, const double *lbub,
                                     uint32_t ld, uint64_t problem_stride)
{
    MULTI_REFUSE(h, "set_simple_bounds_dev");
    return set_sb_impl(h, var_index, index_stride, lbub, ld, problem_stride, true);
}

int lexls_b200_set_fixed_variables(lexls_b200_batch_t *h, uint32_t count, const uint32_t *var_index, const double *value,
                                   const uint8_t *type)
{
    MULTI(h, true, lexls_b200_set_fixed_variables(s.h, count, off(var_index, s.first * count), off(value, s.first * count), off(type, s.first * count)));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (h->kind != LEXLS_B200_LSE) return fail(LEXLS_B200_EINVAL, "fixed variables belong to LexLSE; LexLSI uses a simple-bounds objective");
    if (count > (uint32_t)h->sh.n) return fail(LEXLS_B200_EINVAL, "Cannot fix more than nVar variables");
    CK(cudaSetDevice(h->device));
    // the plan (kernel variant, shared-memory map, scratch) depends on the COUNT only: a batch-of-1 facade that re-sends the
    // same number of fixed variables before every factorize() keeps its plan and its device buffers
    if (h->sh.nfix != (int)count) h->planned = false;
    h->sh.nfix = (int)count;
    if (count == 0) return 0;
    if (!var_index || !value) return fail(LEXLS_B200_EINVAL, "NULL argument");
    const uint64_t N = (uint64_t)count * h->batch;
    {
        std::vector<unsigned char> seen((size_t)h->sh.n);
        for (uint64_t p = 0; p < h->batch; p++)
        {
            std::fill(seen.begin(), seen.end(), 0);
            for (uint32_t a = 0; a < count; a++)
            {
                const uint32_t v = var_index[p * count + a];
                if (v >= (uint32_t)h->sh.n) return fail(LEXLS_B200_EINVAL, "VarIndex >= nVar");
                if (seen[v]) return fail(LEXLS_B200_EINVAL, "Elements of VarIndex are not unique.");
                seen[v] = 1;
            }
        }
    }
    if (N > h->nfix_cap)
    {
        if (h->fidx) guarded_free(h, h->fidx);
        if (h->fval) guarded_free(h, h->fval);
        if (h->ftype) guarded_free(h, h->ftype);
        h->fidx = nullptr; h->fval = nullptr; h->ftype = nullptr;
        h->nfix_cap = 0;
        CK(guarded_malloc(h, (void **)&h->fidx, sizeof(unsigned) * N));
        CK(guarded_malloc(h, (void **)&h->fval, sizeof(double) * N));
        CK(guarded_malloc(h, (void **)&h->ftype, N));
        h->nfix_cap = N;
    }
    CK(cudaMemcpyAsync(h->fidx, var_index, sizeof(unsigned) * N, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->fval, value, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
    if (type)
        CK(cudaMemcpyAsync(h->ftype, type, N, cudaMemcpyHostToDevice, h->stream));
    else
        CK(cudaMemsetAsync(h->ftype, LEXLS_B200_CTR_ACTIVE_UB, N, h->stream)); // the reference's default, lexlse.h:1381
    return wait_copy(h);
}

} // extern "C"
template <typename T>
static int scatter_level(lexls_b200_batch *h, T **dstbuf, uint32_t level, const T *src, bool on_device)
{
    const Shape &s = h->sh;
    if ((int)level >= s.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    CK(cudaSetDevice(h->device));
    if (!*dstbuf)
    {
        CK(cudaMalloc(dstbuf, sizeof(T) * s.Mall * h->batch));
        CK(cudaMemsetAsync(*dstbuf, 0, sizeof(T) * s.Mall * h->batch, h->stream));
    }
    const int m = s.dim[level];
    if (m == 0) return 0;
    const T *dsrc = src;
    if (!on_device)
    {
        const size_t bytes = sizeof(T) * m * h->batch;
        void *slot = nullptr;
        if (int rc = stage_take(h, bytes, &slot)) return rc;
        CK(cudaMemcpyAsync(slot, src, bytes, cudaMemcpyHostToDevice, h->stream));
        if (int rc = wait_copy(h)) return rc; // no-op in asynchronous-copies mode
        dsrc = (const T *)slot;
    }
    else if (int rc = flush_pending(h))
        return rc;
    T *const dst = *dstbuf;
    const int Mall = s.Mall, uoff = s.uoff[level];
    return run_or_defer(h, [=]() -> int {
        scatter_level_kernel<T><<<blocks_for((unsigned long long)m * h->batch), 256, 0, h->stream>>>(dst, dsrc, Mall, uoff, m, h->batch);
        CK(cudaGetLastError());
        return 0;
    });
}

extern "C" {
static int set_guess_impl(lexls_b200_batch *h, uint32_t level, const uint8_t *types, bool on_device)
{
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (h->kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "active-set guesses belong to LexLSI");
    if ((int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    if (!types)
    {
        if (h->guess && h->sh.dim[level])
        {
            std::vector<uint8_t> z((size_t)h->sh.dim[level] * h->batch, 0);
            if (int rc = scatter_level<unsigned char>(h, &h->guess, level, z.data(), false)) return rc;
            return wait_copy_of_temporary(h);
        }
        return 0;
    }
    h->have_guess = true;
    h->have_glist = false; // the two ways of giving a guess are alternatives
    return scatter_level<unsigned char>(h, &h->guess, level, types, on_device);
}
int lexls_b200_set_active_guess(lexls_b200_batch_t *h, uint32_t level, const uint8_t *types)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_set_active_guess(s.h, level, off(types, s.first * h->sh.dim[level])));
    return set_guess_impl(h, level, types, false);
}
int lexls_b200_set_active_guess_dev(lexls_b200_batch_t *h, uint32_t level, const uint8_t *types)
{
    MULTI_REFUSE(h, "set_active_guess_dev");
    return set_guess_impl(h, level, types, true);
}

static int set_x0_impl(lexls_b200_batch *h, const double *x0, bool on_device)
{
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (h->kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "x0 belongs to LexLSI");
    if (!x0)
    {
        h->have_x0 = false;
        return 0;
    }
    CK(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * h->sh.n * h->batch;
    if (!h->x0) CK(guarded_malloc(h, (void **)&h->x0, bytes));
    CK(cudaMemcpyAsync(h->x0, x0, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    if (!on_device) CK(cudaStreamSynchronize(h->stream));
    h->have_x0 = true;
    return 0;
}
int lexls_b200_set_x0(lexls_b200_batch_t *h, const double *x0)
{
    MULTI(h, true, lexls_b200_set_x0(s.h, off(x0, s.first * h->sh.n)));
    return set_x0_impl(h, x0, false);
}
int lexls_b200_set_x0_dev(lexls_b200_batch_t *h, const double *x0)
{
    MULTI_REFUSE(h, "set_x0_dev");
    return set_x0_impl(h, x0, true);
}
int lexls_b200_set_v0(lexls_b200_batch_t *h, uint32_t level, const double *v0)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_set_v0(s.h, level, off(v0, s.first * h->sh.dim[level])));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (h->kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "v0 belongs to LexLSI");
    if ((int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    if (!v0)
    {
        h->have_v0[level] = 0;
        return 0;
    }
    h->have_v0[level] = 1;
    return scatter_level<double>(h, &h->v0, level, v0, false);
}

// ---- solve ------------------------------------------------------------------------------------------------------
int lexls_b200_solve_batch(lexls_b200_batch_t *h)
{
    MULTI(h, false, lexls_b200_solve_batch(s.h)); // asynchronous per device: all devices run concurrently
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    const Shape &s = h->sh;
    for (int k = 0; k < s.P; k++)
        if (!h->data_set[k] && s.dim[k] > 0) return fail(LEXLS_B200_EINVAL, "setData has not been called for every objective");
    if (int rc = flush_pending(h)) return rc; // the deferred repacks of this round (asynchronous-copies mode)
    h->stage_cursor = 0;                      // the next round's copies are ordered behind them by the stream
    h->round_peak   = std::max(h->round_peak, h->round_bytes);
    h->round_bytes  = 0;
    if (int rc = plan(h)) return rc;

    // hot_start_related_tests, lexlsi.h:758-793: v0 is honoured only if given for ALL levels AND x0 is given
    bool use_v0 = h->v0 != nullptr && h->have_x0;
    for (int k = 0; k < s.P && use_v0; k++)
        if (!h->have_v0[k]) use_v0 = false;
    if (h->kind == LEXLS_B200_LSI && h->pr.phase1_v0 && !h->have_x0)
        return fail(LEXLS_B200_EINVAL, "when use_phase1_v0 = true, x_guess has to be specified");

    Buffers b{};
    b.rec          = h->rec;
    b.obj          = h->obj;
    b.rownz        = h->rownz;
    b.sb           = h->sb;
    b.sbidx        = h->sbidx;
    b.sbidx_stride = h->sbidx_stride;
    b.fidx         = h->fidx;
    b.fval         = h->fval;
    b.ftype        = h->ftype;
    b.guess        = (h->have_guess && !h->have_glist) ? h->guess : nullptr;
    b.x0           = h->have_x0 ? h->x0 : nullptr;
    b.v0           = use_v0 ? h->v0 : nullptr;
    b.rec_stride   = h->rec_stride;
    b.obj_stride   = h->obj_stride;
    b.x            = h->x;
    b.v            = h->v;
    b.viol         = h->viol;
    b.act          = h->act;
    b.order        = h->order;
    b.status       = h->status;
    b.counters     = h->counters;
    b.ranks        = h->ranks;
    b.colperm      = (h->kind == LEXLS_B200_LSE) ? h->colperm : nullptr;
    b.lexqr        = (h->want & LEXLS_B200_OUT_LEXQR) ? h->lexqr : nullptr;
    b.hhout        = h->hhout;
    b.lambda       = (h->want & LEXLS_B200_OUT_LAMBDA) ? h->lambda : nullptr;
    if (h->want & LEXLS_B200_OUT_XSTAR)
    {
        b.xstar      = h->xstar;
        b.lexdata    = h->lexdata;
        b.lexqr_star = h->lexqr_star;
        b.hh_star    = h->hh_star;
    }
    b.glist = h->have_glist ? h->glist : nullptr;
    if (h->sm.vecg)
    {
        b.vec_stride      = (uint64_t)((s.Mall + 2) | 1); // odd: the four vectors of a problem do not share cache-line phases
        const size_t need = sizeof(double) * 4 * b.vec_stride * h->grid;
        if (need > h->vec_bytes)
        {
            if (h->vec_scratch) guarded_free(h, h->vec_scratch);
            h->vec_scratch = nullptr;
            h->vec_bytes   = 0;
            CK(guarded_malloc(h, (void **)&h->vec_scratch, need));
            h->vec_bytes = need;
        }
        b.vec_scratch = h->vec_scratch;
    }
    if ((h->want & LEXLS_B200_OUT_LEAST_NORM) && h->kind == LEXLS_B200_LSE)
    {
        const uint64_t n = (uint64_t)s.n;
        h->ln_stride     = 2 * n * (n + 1) + 2 * n + 8;
        const size_t need = sizeof(double) * h->ln_stride * h->grid;
        if (need > h->ln_bytes)
        {
            if (h->ln_scratch) guarded_free(h, h->ln_scratch);
            h->ln_scratch = nullptr;
            h->ln_bytes   = 0;
            CK(guarded_malloc(h, (void **)&h->ln_scratch, need));
            h->ln_bytes = need;
        }
        b.xln        = h->xln;
        b.ln_scratch = h->ln_scratch;
        b.ln_stride  = h->ln_stride;
    }
    for (int g = 0; g < LX_MAX_LEVELS; g++) h->pr.reg_factor[g] = (g + s.off < s.P) ? h->reg_factor_user[g + s.off] : 0.0;
    if (h->pr.reg_type != 0)
    {
        b.reg_scratch = h->reg_scratch;
        b.reg_stride  = h->reg_stride;
    }
    if (h->kind == LEXLS_B200_LSI && h->pr.dfirst)
    {
        if (!h->wsl) CK(guarded_malloc(h, (void **)&h->wsl, sizeof(unsigned) * (size_t)std::max(s.Mall, 1) * h->batch));
        b.wsl = h->wsl;
    }
    b.gcap  = h->gcap;
    if (h->kind == LEXLS_B200_LSI && h->user_params.log_working_set_enabled)
    {
        // at most one entry per iteration, and every iteration but the first costs a factorization
        const unsigned cap = h->user_params.max_number_of_factorizations + 1;
        if (!h->wslog || h->wslog_cap < cap)
        {
            if (h->wslog) guarded_free(h, h->wslog);
            if (h->wslog_count) guarded_free(h, h->wslog_count);
            h->wslog = nullptr;
            h->wslog_count = nullptr;
            CK(guarded_malloc(h, (void **)&h->wslog, sizeof(double) * 6 * (size_t)cap * h->batch));
            CK(guarded_malloc(h, (void **)&h->wslog_count, sizeof(unsigned) * h->batch));
            h->wslog_cap = cap;
        }
        b.wslog       = h->wslog;
        b.wslog_count = h->wslog_count;
        b.wslog_cap   = h->wslog_cap;
    }
    b.queue        = h->queue;
    b.batch        = h->batch;
    b.lod_scratch  = h->lod_scratch;
    b.lod_stride   = h->lod_stride;

    std::string name;
    kernel_fn k = h->pair ? pick_lse_pair_kernel() : pick_kernel(h->kind, h->W, h->DS, h->LG, name, h->user_params.regularization_type != 0);
    b.flag         = h->flag;
    CK(cudaMemsetAsync(h->queue, 0, sizeof(unsigned long long), h->stream));
    CK(cudaMemcpyAsync(h->hflag, h->flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream)); // inspected after the next synchronisation
    h->flag_pending = true;
    CK(cudaEventRecord(h->ev0, h->stream));
    void *args[] = {(void *)&h->sh, (void *)&h->pr, (void *)&h->sm, (void *)&b};
    if (h->LG && getenv("LEXLS_B200_LG_PERSIST"))
    {
        // experiment knob: ask L2 to keep the work-matrix scratch (read and written many times per factorization) and let the
        // objective records (read twice per iteration) stream
        const cudaDeviceProp *pp = nullptr;
        if (int rc = device_props(h->device, &pp)) return rc;
        const size_t want = std::min(h->lod_bytes, (size_t)pp->persistingL2CacheMaxSize);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr  = h->lod_scratch;
        av.accessPolicyWindow.num_bytes = std::min(h->lod_bytes, (size_t)pp->accessPolicyMaxWindowSize);
        av.accessPolicyWindow.hitRatio  = (float)std::min(1.0, (double)want / (double)std::max<size_t>(av.accessPolicyWindow.num_bytes, 1));
        av.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    CK(cudaLaunchKernel((const void *)k, dim3(h->grid), dim3(32 * h->W), args, h->smem, h->stream));
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    return 0;
}

int lexls_b200_synchronize(lexls_b200_batch_t *h)
{
    MULTI(h, false, lexls_b200_synchronize(s.h));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    if (int rc = flush_pending(h)) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return check_pending_flag(h);
}
int lexls_b200_wait_copies(lexls_b200_batch_t *h)
{
    MULTI(h, false, lexls_b200_wait_copies(s.h));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev_copy, h->stream)); // the deferred repacks are not enqueued yet: this is the end of the copies
    CK(cudaEventSynchronize(h->ev_copy));
    return 0;
}
int lexls_b200_validate(lexls_b200_batch_t *h)
{
    MULTI(h, false, lexls_b200_validate(s.h));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    if (int rc = flush_pending(h)) return rc;
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, h->flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (flag) CK(cudaMemsetAsync(h->flag, 0, sizeof(int), h->stream));
    h->flag_pending = false;
    return flag_error(h, flag);
}
int lexls_b200_solve_batch_sync(lexls_b200_batch_t *h)
{
    if (int rc = lexls_b200_solve_batch(h)) return rc;
    return lexls_b200_synchronize(h);
}

// ---- results ----------------------------------------------------------------------------------------------------
static int d2h(lexls_b200_batch *h, void *dst, const void *src, size_t bytes)
{
    CK(cudaSetDevice(h->device));
    if (int rc = flush_pending(h)) return rc;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return check_pending_flag(h);
}
} // extern "C"
template <typename T> static int get_level(lexls_b200_batch *h, uint32_t level, const T *src, T *dst)
{
    if (!h || !dst) return fail(LEXLS_B200_EINVAL, "NULL argument");
    const Shape &s = h->sh;
    if (!src) return fail(LEXLS_B200_EINVAL, "this result does not exist for this solver kind");
    const bool lsi = h->kind == LEXLS_B200_LSI;
    int off, m, width;
    if (lsi)
    {
        if ((int)level >= s.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
        off = s.uoff[level]; m = s.dim[level]; width = s.Mall;
    }
    else
    {
        if ((int)level >= s.G) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
        off = s.goff[level]; m = s.dim[level]; width = s.M;
    }
    if (m == 0) return 0;
    CK(cudaSetDevice(h->device));
    const size_t bytes = sizeof(T) * m * h->batch;
    if (int rc = flush_pending(h)) return rc; // the slot is also the arena of the deferred host inputs
    if (int rc = ensure_stage(h, bytes)) return rc;
    slice_level_kernel<T><<<blocks_for((unsigned long long)m * h->batch), 256, 0, h->stream>>>((T *)h->stage, src, width, off, m, h->batch);
    CK(cudaGetLastError());
    return d2h(h, dst, h->stage, bytes);
}

extern "C" {
int lexls_b200_get_x(lexls_b200_batch_t *h, double *x)
{
    MULTI(h, true, lexls_b200_get_x(s.h, off(x, s.first * h->sh.n)));
    if (!h || !x) return fail(LEXLS_B200_EINVAL, "NULL argument");
    return d2h(h, x, h->x, sizeof(double) * h->sh.n * h->batch);
}
int lexls_b200_get_v(lexls_b200_batch_t *h, uint32_t level, double *v)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_get_v(s.h, level, off(v, s.first * h->sh.dim[level])));
    return get_level<double>(h, level, h ? h->v : nullptr, v);
}
int lexls_b200_get_constraint_violation(lexls_b200_batch_t *h, uint32_t level, double *viol)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_get_constraint_violation(s.h, level, off(viol, s.first * h->sh.dim[level])));
    return get_level<double>(h, level, h ? h->viol : nullptr, viol);
}
int lexls_b200_get_active_set(lexls_b200_batch_t *h, uint32_t level, uint8_t *types)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_get_active_set(s.h, level, off(types, s.first * h->sh.dim[level])));
    return get_level<unsigned char>(h, level, h ? h->act : nullptr, types);
}
int lexls_b200_get_active_order(lexls_b200_batch_t *h, uint32_t level, uint32_t *order)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_get_active_order(s.h, level, off(order, s.first * h->sh.dim[level])));
    return get_level<unsigned>(h, level, h ? h->order : nullptr, order);
}
int lexls_b200_get_status(lexls_b200_batch_t *h, int32_t *status)
{
    MULTI(h, true, lexls_b200_get_status(s.h, off(status, s.first)));
    if (!h || !status) return fail(LEXLS_B200_EINVAL, "NULL argument");
    return d2h(h, status, h->status, sizeof(int) * h->batch);
}
int lexls_b200_get_counters(lexls_b200_batch_t *h, uint32_t *counters)
{
    MULTI(h, true, lexls_b200_get_counters(s.h, off(counters, s.first * 8)));
    if (!h || !counters) return fail(LEXLS_B200_EINVAL, "NULL argument");
    return d2h(h, counters, h->counters, sizeof(unsigned) * 8 * h->batch);
}
int lexls_b200_get_ranks(lexls_b200_batch_t *h, uint32_t *ranks)
{
    MULTI(h, true, lexls_b200_get_ranks(s.h, off(ranks, s.first * h->sh.G)));
    if (!h || !ranks) return fail(LEXLS_B200_EINVAL, "NULL argument");
    return d2h(h, ranks, h->ranks, sizeof(unsigned) * h->sh.G * h->batch);
}
int lexls_b200_request_outputs(lexls_b200_batch_t *h, uint32_t mask)
{
    MULTI(h, false, lexls_b200_request_outputs(s.h, mask));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (mask & ~(uint32_t)(LEXLS_B200_OUT_LEXQR | LEXLS_B200_OUT_LAMBDA | LEXLS_B200_OUT_XSTAR | LEXLS_B200_OUT_LEAST_NORM))
        return fail(LEXLS_B200_EINVAL, "unknown output bit");
    if ((mask & LEXLS_B200_OUT_LEAST_NORM) && h->kind != LEXLS_B200_LSE) return fail(LEXLS_B200_EINVAL, "solveLeastNorm exists for LexLSE only");
    if ((mask & LEXLS_B200_OUT_LAMBDA) && h->kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "Lagrange multipliers exist for LexLSI only");
    if ((mask & LEXLS_B200_OUT_XSTAR) && h->kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "get_xStar exists for LexLSI only");
    CK(cudaSetDevice(h->device));
    const Shape &s = h->sh;
    const int M    = std::max(s.M, 1);
    if ((mask & LEXLS_B200_OUT_LEXQR) && !h->lexqr)
    {
        CK(guarded_malloc(h, (void **)&h->lexqr, sizeof(double) * (size_t)M * (s.n + 1) * h->batch));
        CK(guarded_malloc(h, (void **)&h->hhout, sizeof(double) * (size_t)M * h->batch));
    }
    if ((mask & LEXLS_B200_OUT_LAMBDA) && !h->lambda)
        CK(guarded_malloc(h, (void **)&h->lambda, sizeof(double) * (size_t)s.P * std::max(s.Mall, 1) * h->batch));
    if ((mask & LEXLS_B200_OUT_LEAST_NORM) && !h->xln) CK(guarded_malloc(h, (void **)&h->xln, sizeof(double) * (size_t)s.n * h->batch));
    if ((mask & LEXLS_B200_OUT_XSTAR) && !h->xstar)
    {
        CK(guarded_malloc(h, (void **)&h->xstar, sizeof(double) * (size_t)s.n * h->batch));
        CK(guarded_malloc(h, (void **)&h->lexdata, sizeof(double) * (size_t)M * (s.n + 1) * h->batch));
        CK(guarded_malloc(h, (void **)&h->lexqr_star, sizeof(double) * (size_t)M * (s.n + 1) * h->batch));
        CK(guarded_malloc(h, (void **)&h->hh_star, sizeof(double) * (size_t)M * h->batch));
    }
    h->want = mask;
    return 0;
}
int lexls_b200_get_lexqr(lexls_b200_batch_t *h, double *lexqr, double *hh)
{
    MULTI(h, true, lexls_b200_get_lexqr(s.h, off(lexqr, s.first * (uint64_t)h->sh.M * (h->sh.n + 1)), off(hh, s.first * (uint64_t)h->sh.M)));
    if (!h || !lexqr) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (!(h->want & LEXLS_B200_OUT_LEXQR)) return fail(LEXLS_B200_EINVAL, "request LEXLS_B200_OUT_LEXQR before the solve");
    const Shape &s = h->sh;
    if (int rc = d2h(h, lexqr, h->lexqr, sizeof(double) * (size_t)s.M * (s.n + 1) * h->batch)) return rc;
    if (hh) return d2h(h, hh, h->hhout, sizeof(double) * (size_t)s.M * h->batch);
    return 0;
}
int lexls_b200_get_lambda(lexls_b200_batch_t *h, uint32_t level, double *lambda)
{
    if (h && h->multi() && (int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    MULTI(h, true, lexls_b200_get_lambda(s.h, level, off(lambda, s.first * (uint64_t)h->sh.P * h->sh.dim[level])));
    if (!h || !lambda) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (!(h->want & LEXLS_B200_OUT_LAMBDA)) return fail(LEXLS_B200_EINVAL, "request LEXLS_B200_OUT_LAMBDA before the solve");
    const Shape &s = h->sh;
    if ((int)level >= s.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    const int m = s.dim[level];
    if (m == 0) return 0;
    CK(cudaSetDevice(h->device));
    const unsigned long long cols = (unsigned long long)h->batch * s.P; // one slice per (problem, objective column)
    const size_t bytes            = sizeof(double) * m * cols;
    if (int rc = flush_pending(h)) return rc; // the slot is also the arena of the deferred host inputs
    if (int rc = ensure_stage(h, bytes)) return rc;
    slice_level_kernel<double><<<blocks_for((unsigned long long)m * cols), 256, 0, h->stream>>>((double *)h->stage, h->lambda, s.Mall,
                                                                                               s.uoff[level], m, cols);
    CK(cudaGetLastError());
    return d2h(h, lambda, h->stage, bytes);
}
int lexls_b200_get_working_set_log(lexls_b200_batch_t *h, uint32_t capacity, double *entries, uint32_t *count)
{
    MULTI(h, true, lexls_b200_get_working_set_log(s.h, capacity, off(entries, s.first * (uint64_t)capacity * 6), off(count, s.first)));
    if (!h || !entries || !count) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (!h->wslog) return fail(LEXLS_B200_EINVAL, "set log_working_set_enabled before the solve");
    CK(cudaSetDevice(h->device));
    if (int rc = d2h(h, count, h->wslog_count, sizeof(unsigned) * h->batch)) return rc;
    const uint32_t take = std::min(capacity, h->wslog_cap);
    for (uint64_t p = 0; p < h->batch; p++) count[p] = std::min(count[p], take); // = valid rows of entries[p]
    if (take == 0) return 0;
    CK(cudaMemcpy2DAsync(entries, sizeof(double) * 6 * capacity, h->wslog, sizeof(double) * 6 * h->wslog_cap, sizeof(double) * 6 * take,
                         h->batch, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int lexls_b200_get_column_permutations(lexls_b200_batch_t *h, uint32_t *colperm)
{
    MULTI(h, true, lexls_b200_get_column_permutations(s.h, off(colperm, s.first * h->sh.n)));
    if (!h || !colperm) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (h->kind != LEXLS_B200_LSE) return fail(LEXLS_B200_EINVAL, "column permutations are reported for LexLSE batches");
    return d2h(h, colperm, h->colperm, sizeof(unsigned) * h->sh.n * h->batch);
}
int lexls_b200_device_results(lexls_b200_batch_t *h, const double **x, const int32_t **status, const uint32_t **counters)
{
    MULTI_REFUSE(h, "device_results");
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    if (int rc = flush_pending(h)) return rc;
    if (x) *x = h->x;
    if (status) *status = h->status;
    if (counters) *counters = h->counters;
    return 0;
}

int lexls_b200_kernel_info(lexls_b200_batch_t *h, lexls_b200_kernel_info_t *info)
{
    if (h && h->multi()) return lexls_b200_kernel_info(h->shards[0].h, info); // same shape, same plan on every device
    if (!h || !info) return fail(LEXLS_B200_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    if (int rc = plan(h)) return rc;
    memset(info, 0, sizeof(*info));
    snprintf(info->name, sizeof(info->name), "%s", h->kname.c_str());
    info->grid               = h->grid;
    info->block              = 32 * h->W;
    info->smem_bytes         = h->smem;
    info->problems_per_sm    = h->pair ? 2 * h->per_sm : h->per_sm;
    info->launches_per_solve = 1;
    info->warps_per_problem  = h->W;
    return 0;
}

int lexls_b200_last_solve_ms(lexls_b200_batch_t *h, float *ms)
{
    if (h && h->multi())
    {
        if (!ms) return fail(LEXLS_B200_EINVAL, "NULL argument");
        float worst = 0.f;
        for (auto &s : h->shards)
        {
            float t = 0.f;
            if (int rc = lexls_b200_last_solve_ms(s.h, &t)) return rc;
            worst = std::max(worst, t);
        }
        *ms = worst; // the devices run concurrently: the batch takes as long as the slowest shard
        return 0;
    }
    if (!h || !ms) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (!h->timed) return fail(LEXLS_B200_EINVAL, "no solve has been issued");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

int lexls_b200_get_xstar(lexls_b200_batch_t *h, double *xstar, double *data, double *lexqr, double *hh)
{
    MULTI(h, true, lexls_b200_get_xstar(s.h, off(xstar, s.first * h->sh.n), off(data, s.first * (uint64_t)h->sh.M * (h->sh.n + 1)),
                                        off(lexqr, s.first * (uint64_t)h->sh.M * (h->sh.n + 1)), off(hh, s.first * (uint64_t)h->sh.M)));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (!(h->want & LEXLS_B200_OUT_XSTAR)) return fail(LEXLS_B200_EINVAL, "request LEXLS_B200_OUT_XSTAR before the solve");
    const Shape &s   = h->sh;
    const size_t mat = sizeof(double) * (size_t)s.M * (s.n + 1) * h->batch;
    if (xstar)
        if (int rc = d2h(h, xstar, h->xstar, sizeof(double) * (size_t)s.n * h->batch)) return rc;
    if (data && s.M)
        if (int rc = d2h(h, data, h->lexdata, mat)) return rc;
    if (lexqr && s.M)
        if (int rc = d2h(h, lexqr, h->lexqr_star, mat)) return rc;
    if (hh && s.M)
        if (int rc = d2h(h, hh, h->hh_star, sizeof(double) * (size_t)s.M * h->batch)) return rc;
    return 0;
}

int lexls_b200_set_active_guess_ordered(lexls_b200_batch_t *h, uint32_t count, const uint32_t *obj, const uint32_t *ctr, const uint8_t *type,
                                        uint64_t problem_stride)
{
    MULTI(h, true, lexls_b200_set_active_guess_ordered(s.h, count, off(obj, s.first * problem_stride), off(ctr, s.first * problem_stride),
                                                       off(type, s.first * problem_stride), problem_stride));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if (h->kind != LEXLS_B200_LSI) return fail(LEXLS_B200_EINVAL, "active-set guesses belong to LexLSI");
    if (count == 0)
    {
        h->have_glist = false;
        return 0;
    }
    if (!obj || !ctr || !type) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (h->batch > 1 && problem_stride < count) return fail(LEXLS_B200_EINVAL, "problem_stride too small");
    const Shape &s = h->sh;
    const unsigned cap = count + 1; // the terminator
    std::vector<unsigned> packed((size_t)cap * h->batch, 0xFFFFFFFFu);
    for (uint64_t p = 0; p < h->batch; p++)
    {
        unsigned w = 0;
        for (uint32_t e = 0; e < count; e++)
        {
            const uint64_t at = p * problem_stride + e;
            const unsigned t  = type[at];
            if (t == LEXLS_B200_CTR_INACTIVE) continue; // padding of a ragged list
            if (obj[at] >= (uint32_t)s.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
            if (ctr[at] >= (uint32_t)s.dim[obj[at]]) return fail(LEXLS_B200_EINVAL, "CtrIndex >= nCtr");
            if (t != LEXLS_B200_CTR_ACTIVE_LB && t != LEXLS_B200_CTR_ACTIVE_UB) continue; // the reference only warns (lexlsi.h:130-134)
            packed[p * cap + w++] = (obj[at] << 24) | (t << 16) | ctr[at];
        }
    }
    CK(cudaSetDevice(h->device));
    if (h->gcap < cap)
    {
        if (h->glist) guarded_free(h, h->glist);
        h->glist = nullptr;
        h->gcap  = 0;
        CK(guarded_malloc(h, (void **)&h->glist, sizeof(unsigned) * (size_t)cap * h->batch));
        h->gcap = cap;
    }
    if (h->gcap != cap) // a shorter list than the buffer was sized for: keep the buffer's row stride
    {
        std::vector<unsigned> wide((size_t)h->gcap * h->batch, 0xFFFFFFFFu);
        for (uint64_t p = 0; p < h->batch; p++) std::copy(packed.begin() + p * cap, packed.begin() + (p + 1) * cap, wide.begin() + p * h->gcap);
        packed.swap(wide);
    }
    CK(cudaMemcpyAsync(h->glist, packed.data(), sizeof(unsigned) * packed.size(), cudaMemcpyHostToDevice, h->stream));
    if (int rc = wait_copy_of_temporary(h)) return rc;
    h->have_glist = true;
    return 0;
}

int lexls_b200_set_async_copies(lexls_b200_batch_t *h, int enabled)
{
    MULTI(h, false, lexls_b200_set_async_copies(s.h, enabled));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    if (int rc = flush_pending(h)) return rc;
    if (h->async_copies && !enabled) CK(cudaStreamSynchronize(h->stream)); // back to one slot: nothing may still read the arena
    h->stage_cursor = 0;
    h->round_bytes  = 0;
    h->async_copies = enabled != 0;
    return 0;
}

int lexls_b200_get_x_least_norm(lexls_b200_batch_t *h, double *x)
{
    MULTI(h, true, lexls_b200_get_x_least_norm(s.h, off(x, s.first * h->sh.n)));
    if (!h || !x) return fail(LEXLS_B200_EINVAL, "NULL argument");
    if (!(h->want & LEXLS_B200_OUT_LEAST_NORM)) return fail(LEXLS_B200_EINVAL, "request LEXLS_B200_OUT_LEAST_NORM before the solve");
    return d2h(h, x, h->xln, sizeof(double) * h->sh.n * h->batch);
}

int lexls_b200_set_regularization_factor(lexls_b200_batch_t *h, uint32_t level, double factor)
{
    MULTI(h, false, lexls_b200_set_regularization_factor(s.h, level, factor));
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    if ((int)level >= h->sh.P) return fail(LEXLS_B200_EINVAL, "ObjIndex >= nObj");
    h->reg_factor_user[level] = factor; // a simple-bounds objective is never regularized (objective.h:489-493): its factor is ignored
    return 0;
}

int lexls_b200_check_guard_bands(lexls_b200_batch_t *h, uint32_t *violations)
{
    if (!h || !violations) return fail(LEXLS_B200_EINVAL, "NULL argument");
    *violations = 0;
    if (h->multi())
    {
        for (auto &s : h->shards)
        {
            uint32_t v = 0;
            if (int rc = lexls_b200_check_guard_bands(s.h, &v)) return rc;
            *violations += v;
        }
        return 0;
    }
    if (!canary_enabled()) return fail(LEXLS_B200_EINVAL, "guard bands exist only for handles created with LEXLS_B200_CANARY=1");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    unsigned char band[LX_GUARD_BYTES];
    for (auto &g : h->guards)
    {
        if (g.second == (size_t)-1) continue;
        CK(cudaMemcpy(band, (char *)g.first + g.second, LX_GUARD_BYTES, cudaMemcpyDeviceToHost));
        for (int i = 0; i < LX_GUARD_BYTES; i++)
            if (band[i] != 0xA5)
            {
                (*violations)++;
                break;
            }
    }
    return 0;
}

int lexls_b200_create_batch_multi(int kind, uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType, uint64_t batch,
                                  const int *devices, int ndev, lexls_b200_batch_t **out)
{
    if (!out) return fail(LEXLS_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (!devices || ndev < 1) return fail(LEXLS_B200_EINVAL, "a device list is required");
    if (batch == 0) return fail(LEXLS_B200_EINVAL, "empty batch");
    for (int a = 0; a < ndev; a++)
        for (int b = a + 1; b < ndev; b++)
            if (devices[a] == devices[b]) return fail(LEXLS_B200_EINVAL, "a device appears twice in the device list");
    lexls_b200_batch *m = new lexls_b200_batch();
    // contiguous ranges; the first (batch % ndev) devices take one problem more
    const uint64_t base = batch / (uint64_t)ndev, extra = batch % (uint64_t)ndev;
    uint64_t first = 0;
    for (int i = 0; i < ndev; i++)
    {
        const uint64_t count = base + ((uint64_t)i < extra ? 1 : 0);
        if (count == 0) continue;
        lexls_b200_batch *c = nullptr;
        if (int rc = lexls_b200_create_batch(kind, nVar, nObj, ObjDim, ObjType, count, devices[i], &c))
        {
            const std::string msg = g_err;
            for (auto &s : m->shards) lexls_b200_destroy(s.h);
            delete m;
            return fail(rc, "device " + std::to_string(devices[i]) + ": " + msg);
        }
        m->shards.push_back(Shard{c, first, count});
        first += count;
    }
    m->kind   = kind;
    m->batch  = batch;
    m->device = m->shards[0].h->device;
    m->sh     = m->shards[0].h->sh; // the shape (dimensions, offsets) is shared by every shard
    lexls_b200_default_params(&m->user_params);
    *out = m;
    return 0;
}
int lexls_b200_shard_count(lexls_b200_batch_t *h, int *count)
{
    if (!h || !count) return fail(LEXLS_B200_EINVAL, "NULL argument");
    *count = h->multi() ? (int)h->shards.size() : 1;
    return 0;
}
int lexls_b200_get_shard(lexls_b200_batch_t *h, int index, lexls_b200_batch_t **shard, uint64_t *first, uint64_t *count, int *device)
{
    if (!h) return fail(LEXLS_B200_EINVAL, "NULL handle");
    const int n = h->multi() ? (int)h->shards.size() : 1;
    if (index < 0 || index >= n) return fail(LEXLS_B200_EINVAL, "shard index out of range");
    lexls_b200_batch *c = h->multi() ? h->shards[index].h : h;
    if (shard) *shard = c;
    if (first) *first = h->multi() ? h->shards[index].first : 0;
    if (count) *count = c->batch;
    if (device) *device = c->device;
    return 0;
}

} // extern "C"

