// lexls_device.cuh — device side of the batched LexLSE / LexLSI solver (sm_100a, FP64).
//
// One "group" of W warps (one CTA) owns one problem at a time; its l-QR work matrix (the reference's LOD,
// include/lexls/lexlse.h:2827) lives in shared memory for the whole solve, column-major with an ODD leading
// dimension so that both "lanes over rows" and "lanes over columns" accesses are bank-conflict free.
// The whole LexLSI active-set loop (lexlsi.h:205-246) runs on the device: no host round trip per iteration.
//
// Arithmetic contract (same as oracle/lexls_oracle.hpp, which restates the reference on the CPU):
//   (S) sequential fused multiply-add, one accumulator, ascending index   -> every per-column / per-row sum
//   (T) tree32: term i -> lane i%32 (fma, ascending), then xor-butterfly   -> every single-scalar reduction
// The file is compiled with -fmad=false; every contraction is an explicit fma().
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace lexls_dev
{

#define LX_MAX_LEVELS 16
#ifndef LX_LSE_COMPACT_QR
#define LX_LSE_COMPACT_QR false
#endif
#ifndef LX_LSE_COMPACT_GS
#define LX_LSE_COMPACT_GS false
#endif

enum
{
    CTR_INACTIVE           = 0,
    CTR_ACTIVE_LB          = 1,
    CTR_ACTIVE_UB          = 2,
    CTR_ACTIVE_EQ          = 3,
    CORRECT_SIGN_OF_LAMBDA = 4
};
enum
{
    OP_UNDEFINED = 0,
    OP_ADD       = 1,
    OP_REMOVE    = 2
};
enum
{
    ST_UNKNOWN        = -1,
    ST_SOLVED         = 0,
    ST_SOLVED_CYCLING = 1,
    ST_MAX_FACT       = 2
};

// ---------------------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------------------
struct Shape
{
    int kind;  // 0 LSE, 1 LSI
    int n;     // nVar
    int P;     // user levels (nObj)
    int off;   // nObjOffset: 1 when level 0 is SIMPLE_BOUNDS (lexlsi.h:925-929)
    int G;     // LexLSE levels = P - off
    int m0;    // rows of the simple-bounds level (0 if none)
    int M;     // rows of all general levels
    int Mall;  // m0 + M
    int ld;    // leading dimension of LOD in shared memory (even with ld/2 odd, >= M): see lse_factorize
    int ldo;   // leading dimension of the stacked objective record [M x (n+2)] (LSI) / of the LSE record (= ld)
    int nfix;  // LSE: number of fixed variables
    int dim[LX_MAX_LEVELS];       // user ObjDim
    int uoff[LX_MAX_LEVELS + 1];  // first row of user level k in the "all rows" numbering (simple bounds first)
    int goff[LX_MAX_LEVELS + 1];  // first row of LexLSE level g in the stacked general rows
};

struct Params
{
    unsigned max_fact;
    double tol_lindep, tol_wrong, tol_correct, tol_feas;
    int cycling;
    unsigned cyc_max;
    double cyc_step;
    int mod_x, mod_active, mod_inactive, min_viol, phase1_v0;
    int reg_type;       // RegularizationType, typedefs.h:32-44 (0 = none; 1 TIKHONOV, 2 TIKHONOV_CG, 3 R, 4 R_NO_Z, 5 RT_NO_Z, 6 RT_NO_Z_CG,
                        // 7 TIKHONOV_1, 8 TIKHONOV_2, 9 TEST)
    unsigned reg_cg_iter; // max_number_of_CG_iterations, typedefs.h:91 (types 2 and 6)
    double reg_var;     // variable_regularization_factor, typedefs.h:170-176
    double reg_factor[LX_MAX_LEVELS]; // setRegularizationFactor per LexLSE level (lexlse.h:1478, objective.h:492)
    int dfirst; // deactivate_first_wrong_sign (lexlsi.h:1048-1105): drop the wrong-sign constraint that was activated first
    int dmma; // Gauss step: Schur update (and the off-diagonal part of Left * inv(R)) on the FP64 tensor cores
};

// byte offsets into the dynamic shared memory of one group
struct SmemMap
{
    unsigned lod, obj, sb, x, dx, lx, lxo, colnorm, invd, hh, wsp, v, dv, Ax, Adx, vv, ptmp;
    unsigned active, inactive, atype, alltype, ctype, fidx, ftype, colperm, varofcol, sbidx;
    unsigned ldim, frow, fcol, rank, nact, ninact, scal, mbar, total;
    int ldv; // leading dimension of the level view of the global-memory variants
    int vecg; // LexLSI: v, dv, Ax, Adx live in a global scratch slice (vec_scratch) instead of shared memory: one more problem per SM
};

struct Buffers
{
    // inputs
    const double *rec;      // LSE: [B][rec_stride] smem image of [A|b], column-major ld = shape.ld
    double *obj;            // LSI: [B][obj_stride] stacked general levels [M x (n+2)] column-major ld = ldo (bounds may be relaxed)
    const unsigned char *rownz; // LSI: [B][M] 1 = the row of A is not all zero (row_nonzero_kernel at set_data, lexlsi.h:373-377)
    double *sb;             // LSI: [B][2*m0] lb | ub of the simple-bounds level
    const unsigned *sbidx;  // LSI: [B or 1][m0] variable indices
    unsigned long long sbidx_stride;
    const unsigned *fidx;   // LSE: [B][nfix]
    const double *fval;     // LSE: [B][nfix]
    const unsigned char *ftype;  // LSE: [B][nfix] or null
    const unsigned char *guess;  // LSI: [B][Mall] or null
    const double *x0;            // LSI: [B][n] or null
    const double *v0;            // LSI: [B][Mall] or null (only used when given for every level and x0 is given)
    unsigned long long rec_stride, obj_stride;
    // outputs
    double *x;                 // [B][n]
    double *v;                 // [B][Mall]  (LSE: [B][M])
    double *viol;              // [B][Mall]
    unsigned char *act;        // [B][Mall]
    unsigned *order;           // [B][Mall]
    int *status;               // [B]
    unsigned *counters;        // [B][8]
    unsigned *ranks;           // [B][G]
    unsigned *colperm;         // [B][n] or null (debug)
    double *lexqr;             // [B][(n+1) x M] factorized work matrix (column-major, ld = M) or null   (get_lexqr)
    double *hhout;             // [B][M] Householder scalars, with lexqr
    double *lambda;            // LSI: [B][P][Mall] Lagrange multipliers (getLambda, lexlsi.h:552-605) or null
    double *xstar;             // LSI: [B][n] get_xStar (lexlsi.h:519): x of the LexLSE problem of the FINAL working set, or null
    double *lexdata;           // LSI, with xstar: [B][(n+1) x M] that problem's constraints as given (LexLSE::get_data, lexlse.h:119, :1631)
    double *lexqr_star;        // LSI, with xstar: [B][(n+1) x M] its factorization (what get_lexqr returns after get_xStar)
    double *hh_star;           // LSI, with xstar: [B][M]
    double *vec_scratch;       // LSI, SmemMap::vecg: [grid][4 * vec_stride] v | dv | Ax | Adx of the resident problem
    unsigned long long vec_stride;
    double *xln;               // LSE: [B][n] least-norm solution (solveLeastNorm, lexlse.h:1046-1262) or null
    double *ln_scratch;        // LSE, with xln: [grid][ln_stride] work arrays (staircase copy n x (n+1), D n x n, 2n vectors)
    unsigned long long ln_stride;
    double *reg_scratch;       // regularization: [grid][reg_stride] = null-space basis Z (n x (n+1)) | D (n x (n+1)) | 3n work doubles
    unsigned long long reg_stride;
    unsigned *wsl;             // LSI, deactivate_first_wrong_sign: [B][Mall] chronological list of activations (obj << 16 | ctr), lexlsi.h:158
    const unsigned *glist;     // LSI: [B][gcap] api_activate calls in CALL ORDER: obj << 24 | type << 16 | ctr, 0xFFFFFFFF ends the list; or null
    unsigned gcap;
    double *wslog;             // LSI: [B][wslog_cap][6] working-set log {obj, ctr, type, alpha_or_lambda, rank, cycling} or null
    unsigned *wslog_count;     // LSI: [B] entries written (capped by wslog_cap)
    unsigned wslog_cap;
    unsigned long long *queue; // work counter
    const int *flag;           // input-validation flag of the set_* calls: non-zero -> the solve is refused (status UNKNOWN)
    unsigned long long batch;
    double *lod_scratch;       // "LOD in global memory" variants: [grid][lod_stride] work matrices (L2 resident)
    unsigned long long lod_stride;
};

// ---------------------------------------------------------------------------------------------------------------
// group primitives
// ---------------------------------------------------------------------------------------------------------------
template <int W> __device__ __forceinline__ void gsync()
{
    if constexpr (W == 1)
        __syncwarp();
    else
        __syncthreads();
}

__device__ __forceinline__ double bfly_sum(double a)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
}

struct ArgMin
{
    double v;
    int pos;
};
// lexicographic (value, position) minimum over the warp
__device__ __forceinline__ ArgMin bfly_argmin(ArgMin a)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
    {
        double ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        int op    = __shfl_xor_sync(0xffffffffu, a.pos, o);
        if (ov < a.v || (ov == a.v && op < a.pos))
        {
            a.v   = ov;
            a.pos = op;
        }
    }
    return a;
}
// (value, position) maximum, ties -> smaller position (Eigen maxCoeff: first maximum)
__device__ __forceinline__ ArgMin bfly_argmax(ArgMin a)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
    {
        double ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        int op    = __shfl_xor_sync(0xffffffffu, a.pos, o);
        if (ov > a.v || (ov == a.v && op < a.pos))
        {
            a.v   = ov;
            a.pos = op;
        }
    }
    return a;
}

// ---------------------------------------------------------------------------------------------------------------
// per-problem view of shared memory
// ---------------------------------------------------------------------------------------------------------------
struct Ctx
{
    double *LOD, *LVIEW, *obj, *sb, *x, *dx, *lx, *lxo, *colnorm, *invd, *hh, *wsp, *v, *dv, *Ax, *Adx, *vv, *ptmp;
    unsigned short *active, *inactive, *fidx, *colperm, *varofcol, *sbidx;
    unsigned char *atype, *alltype, *ctype, *ftype;
    int *ldim, *frow, *fcol, *rank, *nact, *ninact, *scal;
    unsigned long long *mbar;
};

__device__ __forceinline__ Ctx make_ctx(unsigned char *base, const SmemMap &m)
{
    Ctx c;
    c.LOD      = (double *)(base + m.lod);
    c.LVIEW    = c.LOD; // the global-memory variants re-point LOD; this stays the shared-memory level view
    c.obj      = (double *)(base + m.obj);
    c.sb       = (double *)(base + m.sb);
    c.x        = (double *)(base + m.x);
    c.dx       = (double *)(base + m.dx);
    c.lx       = (double *)(base + m.lx);
    c.lxo      = (double *)(base + m.lxo);
    c.colnorm  = (double *)(base + m.colnorm);
    c.invd     = (double *)(base + m.invd);
    c.hh       = (double *)(base + m.hh);
    c.wsp      = (double *)(base + m.wsp);
    c.v        = (double *)(base + m.v);
    c.dv       = (double *)(base + m.dv);
    c.Ax       = (double *)(base + m.Ax);
    c.Adx      = (double *)(base + m.Adx);
    c.vv       = (double *)(base + m.vv);
    c.ptmp     = (double *)(base + m.ptmp);
    c.active   = (unsigned short *)(base + m.active);
    c.inactive = (unsigned short *)(base + m.inactive);
    c.atype    = (unsigned char *)(base + m.atype);
    c.alltype  = (unsigned char *)(base + m.alltype);
    c.ctype    = (unsigned char *)(base + m.ctype);
    c.fidx     = (unsigned short *)(base + m.fidx);
    c.ftype    = (unsigned char *)(base + m.ftype);
    c.colperm  = (unsigned short *)(base + m.colperm);
    c.varofcol = (unsigned short *)(base + m.varofcol);
    c.sbidx    = (unsigned short *)(base + m.sbidx);
    c.ldim     = (int *)(base + m.ldim);
    c.frow     = (int *)(base + m.frow);
    c.fcol     = (int *)(base + m.fcol);
    c.rank     = (int *)(base + m.rank);
    c.nact     = (int *)(base + m.nact);
    c.ninact   = (int *)(base + m.ninact);
    c.scal     = (int *)(base + m.scal);
    c.mbar     = (unsigned long long *)(base + m.mbar);
    return c;
}

// ---------------------------------------------------------------------------------------------------------------
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (cp.async.bulk; SASS: UBLKCP)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// bulk copies are limited to < 2^20 bytes per instruction and need 16-byte granularity; issue in chunks
__device__ __forceinline__ void tma_load_record(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    mbar_expect_tx(bar, bytes);
    unsigned done = 0;
    while (done < bytes)
    {
        unsigned chunk = bytes - done;
        if (chunk > 65536u) chunk = 65536u;
        tma_bulk_g2s((char *)dst + done, (const char *)src + done, chunk, bar);
        done += chunk;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// flop accounting: the algorithmic count of BASELINE.md / SURVEY.md 8(d), evaluated with the ACTUAL dims and ranks
// ---------------------------------------------------------------------------------------------------------------
// algorithmic flop counter.  A double: every term is an integer far below 2^53, so the arithmetic is exact, and the
// FP64 pipe is idle where the 64-bit integer multiplies of the closed forms were ~100 instructions per level.
struct FlopCount
{
    double f;
};
// flops of the column-pivoted QR of one level: m rows, cc open columns, rank r
__device__ __forceinline__ double qr_level_flops(const int dim, const int ncols, const int rank)
{
    const int rpi  = (rank < dim) ? rank : dim - 1;
    const double m = dim, cc = ncols, r = rank, rp = rpi;
    const double R1 = r * (r - 1.0) * 0.5;
    double f        = 2.0 * m * cc + 2.0 * (r * m - R1) + 2.0 * (r * (cc - 1.0) - R1);
    if (rpi > 0)
    {
        const double T1 = rp * (rp - 1.0) * 0.5;
        const double T2 = (double)((unsigned long long)((unsigned)(rpi - 1) * (unsigned)rpi) * (unsigned)(2 * rpi - 1) / 6ull);
        f += 3.0 * (rp * m - T1) + 4.0 * rp + 4.0 * (rp * m * cc - (m + cc) * T1 + T2);
    }
    return f;
}

// aligned 2-double accesses (rows 2t, 2t+1 of one column): LDS.128 / STS.128 in shared memory
__device__ __forceinline__ double2 ld2(const double *p)
{
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ void st2(double *p, double2 v)
{
    *reinterpret_cast<double2 *>(p) = v;
}

// order-preserving map double -> u64 (so that integer max == floating max)
__device__ __forceinline__ unsigned long long dkey(double v)
{
    const long long b = __double_as_longlong(v);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 tensor cores.  tcgen05 has no FP64 kind, so on sm_100a FP64 tensor work is mma.sync (SASS: DMMA).
// D = A * B + C for one 8 x 8 output tile and 4 pivots.  Measured on B200 (scripts/dmma_order.cu ->
// profiles/r2_dmma_order.json, 262,144 outputs with heavy cancellation): every output is bit-identical to the chain
// c = fma(a[k], b[k], c), k = 0, 1, 2, 3 — the (S) order of the arithmetic contract — so a k-loop of these
// instructions reproduces the DFMA Schur update of the oracle bit for bit while issuing 8x fewer instructions.
// Fragments (lane = 4 g + t): A(8 x 4) a = A[g][t];  B(4 x 8) b = B[t][g];  C(8 x 8) c0, c1 = C[g][2t], C[g][2t + 1].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_884(double &c0, double &c1, const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C(rows, cols) -= L(rows, 0:kmax) * U(0:kmax, cols) in 8 x 8 tiles, one warp per row tile (lexlse.h:464-469 and the
// finished-column part of :450-461).  rows [r_lo, r_hi) of `LOD` (leading dimension ld; shared memory, or global memory for
// the global-work-matrix kernels), L(row, k) = LOD[row + ld * (lc + k)], U(k, col) = RB[k + rl * col] (shared memory),
// columns [c_lo, c_hi).  Per entry the updates arrive in ascending k.  Padding pivots / rows / columns carry zeros in A.
template <int W>
__device__ __noinline__ void schur_dmma(double *const LOD, const int ld, const double *const RB, const int rl, const int r_lo,
                                        const int r_hi, const int lc, const int kmax, const int c_lo, const int c_hi)
{
    // (a real call: the tensor-core variant must not disturb the register allocation of the DFMA code around it)
    constexpr int KF = 10; // A fragments held in registers per pass: 40 pivots
    constexpr int CT = 4;  // column tiles in flight per warp: independent accumulators hide the DMMA and load latency
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int nrt = (r_hi - r_lo + 7) >> 3, nct = (c_hi - c_lo + 7) >> 3;
    for (int rt = warp; rt < nrt; rt += W)
    {
        const int r      = r_lo + 8 * rt + g;
        const bool rok   = r < r_hi;
        double *const rp = LOD + (rok ? r : r_hi - 1);
        for (int k0 = 0; k0 < kmax; k0 += 4 * KF)
        {
            double af[KF];
#pragma unroll
            for (int q = 0; q < KF; q++)
            {
                const int k = k0 + 4 * q + t4;
                af[q]       = (rok && k < kmax) ? -rp[ld * (lc + k)] : 0.0;
            }
            for (int ct0 = 0; ct0 < nct; ct0 += CT)
            {
                double d0[CT], d1[CT];
                const double *bp[CT];
                bool s0[CT], s1[CT], bok[CT];
#pragma unroll
                for (int j = 0; j < CT; j++)
                {
                    const int cc  = c_lo + 8 * (ct0 + j); // beyond the last tile: every predicate below is false
                    const int c0i = cc + 2 * t4, c1i = c0i + 1, cb = cc + g;
                    s0[j]  = rok && c0i < c_hi;
                    s1[j]  = rok && c1i < c_hi;
                    bok[j] = cb < c_hi;
                    d0[j]  = s0[j] ? rp[ld * c0i] : 0.0;
                    d1[j]  = s1[j] ? rp[ld * c1i] : 0.0;
                    bp[j]  = RB + rl * (bok[j] ? cb : c_lo) + k0 + t4;
                }
#pragma unroll
                for (int q = 0; q < KF; q++)
                    if (k0 + 4 * q < kmax) // warp-uniform
                    {
                        const bool kok = k0 + 4 * q + t4 < kmax;
#pragma unroll
                        for (int j = 0; j < CT; j++)
                        {
                            const double b = (bok[j] && kok) ? bp[j][4 * q] : 0.0;
                            dmma_884(d0[j], d1[j], af[q], b);
                        }
                    }
#pragma unroll
                for (int j = 0; j < CT; j++)
                {
                    const int c0i = c_lo + 8 * (ct0 + j) + 2 * t4;
                    if (s0[j]) rp[ld * c0i] = d0[j];
                    if (s1[j]) rp[ld * (c0i + 1)] = d1[j];
                }
            }
        }
    }
}

} // namespace lexls_dev
#include "lexls_panel.cuh"
#ifdef LEXLS_B200_WITH_PAIR // the two-problems-per-warp experiment (measured slower): not in the product build
#include "lexls_pair.cuh"
#endif
namespace lexls_dev
{

// ---------------------------------------------------------------------------------------------------------------
// column-pivoted Householder QR of one level in shared memory, lexlse.h:182-272 (any shape, W warps)
// ---------------------------------------------------------------------------------------------------------------
// LGV (the "work matrix in global memory" kernels): the rows of the level being factorized — the aligned row pairs
// [fr & ~1, fr + dim rounded up) of columns ColIndex .. n — are staged in a shared-memory view `lv` (leading dimension ldv)
// for the duration of the level, so that only the column swaps of the other rows go to global memory (L2).
template <int W, bool LGV, bool REG = false>
__device__ __forceinline__ void qr_level_smem(const Ctx &c, const int n, const int ld, const int fr, const int dim, int &ColIndex,
                                              const double tol, double *const lv = nullptr, const int ldv = 0, double *const Zs = nullptr)
{
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x, lane = tid & 31;
    double *const LOD = c.LOD;
    const int hp      = ld >> 1; // row pairs per column
    const int vt0 = fr >> 1, vt1 = (fr + dim + 1) >> 1; // row pairs of the level
    // pair t of column col / element of level row r
    auto P = [&](int col, int t) -> double * {
        if (LGV && t >= vt0 && t < vt1) return lv + ldv * col + 2 * (t - vt0);
        return LOD + ld * col + 2 * t;
    };
    auto E = [&](int col, int r) -> double * {
        if (LGV) return lv + ldv * col + (r - 2 * vt0);
        return LOD + ld * col + r;
    };
    const int fc0 = ColIndex;
    if constexpr (LGV)
    {
        const int npv = vt1 - vt0, ncv = n + 1 - fc0;
        for (int e = tid; e < npv * ncv; e += NT)
        {
            const int t = e % npv, col = fc0 + e / npv;
            st2(lv + ldv * col + 2 * t, ld2(LOD + ld * col + 2 * (vt0 + t)));
        }
        gsync<W>();
    }
    // initial squared column norms over this level's rows, lexlse.h:193-196   (S2)
    for (int k = ColIndex + tid; k < n; k += NT)
    {
        const double *col = E(k, 0);
        double acc[2]     = {0.0, 0.0};
        for (int r = fr; r < fr + dim; r++) acc[r & 1] = fma(col[r], col[r], acc[r & 1]);
        c.colnorm[k] = acc[0] + acc[1];
    }
    gsync<W>();

    // LGV: the column swaps of the row pairs outside the level go to global memory (L2).  Each thread owns one such pair
    // (from the top of the group: the warps without a staged pair) and stores the swap of a step during the NEXT step,
    // so that no barrier waits for the round trip; same thread, same addresses: program order keeps the swaps in sequence
    const int npv_ = vt1 - vt0, nout = LGV ? hp - npv_ : 0;
    const int og = NT - 1 - tid, tg = (og < vt0) ? og : og + npv_;
    const bool outside = LGV && og < nout;
    double2 pend_c = make_double2(0.0, 0.0), pend_p = make_double2(0.0, 0.0);
    int pendp = -1, pendc = 0;

    for (int counter = 0; counter < dim; counter++)
    {
        const int row = fr + counter, rem = dim - counter;

        // pivot = first maximum of the down-dated norms, lexlse.h:205-206 (every warp computes it redundantly):
        // integer max over order-preserving keys (REDUX), then the smallest column index holding the maximum
        int p;
        {
            double best = 0.0;
            int col     = 0x7fffffff;
            for (int k = ColIndex + lane; k < n; k += 32)
            {
                const double v = c.colnorm[k];
                if (col == 0x7fffffff || v > best)
                {
                    best = v;
                    col  = k;
                }
            }
            const unsigned long long key = (col == 0x7fffffff) ? 0ull : dkey(best);
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
            const bool mine   = (hi == mh) && ((unsigned)key == ml) && (col != 0x7fffffff);
            p                 = (int)__reduce_min_sync(0xffffffffu, mine ? (unsigned)col : 0x7fffffffu);
            if (p >= n) p = ColIndex;
        }

        // ||tail||^2 by (W8): row i > counter -> slot (i - wb) % 8 with wb = counter & ~3, butterfly 4, 2, 1;
        // ||col||^2 = c0^2 + ||tail||^2 (lexlse.h:210 and makeHouseholderInPlace)
        const double *colp = E(p, fr);
        double acc2        = 0.0;
        if (lane < 8)
        {
            int i = (counter & ~3) + lane;
            if (i <= counter) i += 8; // counter - (counter & ~3) < 4: one bump reaches the first row below the pivot
            for (; i < dim; i += 8)
            {
                const double a = colp[i];
                acc2           = fma(a, a, acc2);
            }
        }
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, 4);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, 2);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, 1);
        acc2 = __shfl_sync(0xffffffffu, acc2, 0);
        const double c0     = colp[counter];
        const double maxval = fma(c0, c0, acc2);
        if (maxval < tol) break; // rank found, lexlse.h:214-217

        // Householder scalars (makeHouseholderInPlace), branch-free: the square root and the two reciprocals overlap
        const bool hhon      = (rem > 1) && !(acc2 <= DBL_MIN); // a non-trivial reflector exists
        const bool zero_tail = (rem > 1) && !hhon;
        const double sq      = sqrt(maxval);
        const double bb      = (c0 >= 0.0) ? -sq : sq;
        const double beta    = hhon ? bb : c0;
        const double inv     = 1.0 / beta; // reciprocal of the diagonal entry R(c,c): Gauss step, solve, and tau
        const double rdn     = 1.0 / (c0 - bb);
        const double rden    = hhon ? rdn : 0.0;
        const double tau     = hhon ? (bb - c0) * inv : 0.0;

        if constexpr (W > 1) gsync<W>(); // all warps finished reading colnorm / column p

        if (tid == 0)
        {
            const double t      = c.colnorm[ColIndex];
            c.colnorm[ColIndex] = maxval;
            c.colperm[ColIndex] = (unsigned short)p;
            c.invd[ColIndex]    = inv;
            if (p != ColIndex)
            {
                c.colnorm[p]         = t;
                const unsigned short s = c.varofcol[ColIndex];
                c.varofcol[ColIndex] = c.varofcol[p];
                c.varofcol[p]        = s;
                if constexpr (REG) // regularization: the accumulated null-space basis follows the column swap, lexlse.h:228-231
                    for (int q = 0; q < fc0; q++)
                    {
                        const double z      = Zs[q + n * ColIndex];
                        Zs[q + n * ColIndex] = Zs[q + n * p];
                        Zs[q + n * p]        = z;
                    }
            }
            if (rem > 1) c.hh[row] = tau;
        }
        // column swap over ALL rows (lexlse.h:222-227) fused with the Householder scaling (:241-242); the reflector
        // v = [1; essential], zero outside its rows, is also written to c.vv for the broadcast reads of the apply
        const int pb = row >> 1, pe = (row + rem + 1) >> 1; // pairs touched by the reflector
        if constexpr (LGV)
        {
            if (outside)
            {
                if (pendp >= 0)
                {
                    st2(LOD + ld * pendp + 2 * tg, pend_c);
                    st2(LOD + ld * pendc + 2 * tg, pend_p);
                }
                pendp = -1;
                if (p != ColIndex)
                {
                    pend_p = ld2(LOD + ld * p + 2 * tg);
                    pend_c = ld2(LOD + ld * ColIndex + 2 * tg);
                    pendp  = p;
                    pendc  = ColIndex;
                }
            }
            if (p != ColIndex)
                for (int o = NT + tid; o < nout; o += NT) // more outside pairs than threads: swapped at once
                {
                    const int t     = (o < vt0) ? o : o + npv_;
                    const double2 a = ld2(LOD + ld * p + 2 * t);
                    st2(LOD + ld * p + 2 * t, ld2(LOD + ld * ColIndex + 2 * t));
                    st2(LOD + ld * ColIndex + 2 * t, a);
                }
        }
        for (int t = (LGV ? vt0 : 0) + tid; t < (LGV ? vt1 : hp); t += NT)
        {
            double *const pp = P(p, t), *const pc = P(ColIndex, t);
            double2 a        = ld2(pp);
            if (p != ColIndex) st2(pp, ld2(pc));
            if (rem > 1 && t >= pb && t < pe)
            {
                const int r0 = 2 * t, r1 = r0 + 1;
                double2 v;
                if (r0 == row)
                {
                    a.x = beta;
                    v.x = 1.0;
                }
                else if (r0 > row && r0 < row + rem)
                {
                    a.x = zero_tail ? 0.0 : a.x * rden;
                    v.x = a.x;
                }
                else
                    v.x = 0.0;
                if (r1 == row)
                {
                    a.y = beta;
                    v.y = 1.0;
                }
                else if (r1 > row && r1 < row + rem)
                {
                    a.y = zero_tail ? 0.0 : a.y * rden;
                    v.y = a.y;
                }
                else
                    v.y = 0.0;
                st2(c.vv + 2 * t, v);
            }
            st2(pc, a);
        }
        gsync<W>();

        // apply H to columns ColIndex+1 .. n (RHS included), lexlse.h:243-246, and down-date the norms, :262-266
        {
            const bool apply  = tau != 0.0;
            const double *vp  = c.vv + 2 * pb;
            const int np      = pe - pb;
            // one thread per column: both parity chains of (S2) in one thread
            auto apply_thread_per_column = [&]() {
                for (int j = ColIndex + 1 + tid; j <= n; j += NT)
                {
                    double *cp = P(j, pb);
                    if (apply)
                    {
                        double ae = 0.0, ao = 0.0;
#pragma unroll 2
                        for (int q = 0; q < np; q++)
                        {
                            const double2 v = ld2(vp + 2 * q), a = ld2(cp + 2 * q);
                            ae = fma(v.x, a.x, ae);
                            ao = fma(v.y, a.y, ao);
                        }
                        const double w = tau * (ae + ao);
#pragma unroll 2
                        for (int q = 0; q < np; q++)
                        {
                            const double2 v = ld2(vp + 2 * q);
                            double2 a       = ld2(cp + 2 * q);
                            a.x             = fma(-v.x, w, a.x);
                            a.y             = fma(-v.y, w, a.y);
                            st2(cp + 2 * q, a);
                        }
                    }
                    if (j < n)
                    {
                        const double a0 = *E(j, row);
                        c.colnorm[j]    = fma(-a0, a0, c.colnorm[j]);
                    }
                }
            };
            // two threads per column, one per row parity: each runs one of the two (S2) chains, the pair adds them
            auto apply_thread_pair_per_column = [&]() {
                const int par = tid & 1;
                for (int jb = ColIndex + 1 + ((tid & ~31) >> 1); jb <= n; jb += NT / 2) // warp-uniform trip count
                {
                    const int j      = jb + (lane >> 1);
                    const bool valid = j <= n;
                    double *cp       = P(valid ? j : n, pb) + par;
                    const double *vq = vp + par;
                    double acc       = 0.0;
                    if (apply && valid)
                    {
#pragma unroll 4
                        for (int q = 0; q < np; q++) acc = fma(vq[2 * q], cp[2 * q], acc);
                    }
                    const double w = tau * (acc + __shfl_xor_sync(0xffffffffu, acc, 1));
                    if (apply && valid)
                    {
#pragma unroll 4
                        for (int q = 0; q < np; q++) cp[2 * q] = fma(-vq[2 * q], w, cp[2 * q]);
                    }
                    if (valid && j < n && par == (row & 1))
                    {
                        const double a0 = *E(j, row);
                        c.colnorm[j]    = fma(-a0, a0, c.colnorm[j]);
                    }
                }
            };
            if constexpr (LGV || W >= 4)
            {
                // more open columns than thread pairs: one pass with one thread per column instead of two passes with two
                // (same sums: the pair adds its two chains, the single thread adds the same two chains)
                if (n - ColIndex > NT / 2)
                    apply_thread_per_column();
                else
                    apply_thread_pair_per_column();
            }
            else
                apply_thread_per_column();
        }
        ColIndex++;
        gsync<W>();
        if (ColIndex == n) break; // lexlse.h:255-258
    }
    if constexpr (LGV)
    {
        if (outside && pendp >= 0)
        {
            st2(LOD + ld * pendp + 2 * tg, pend_c);
            st2(LOD + ld * pendc + 2 * tg, pend_p);
        }
        // the level's rows (R, the reflectors, the not eliminated rest) go back to the work matrix
        gsync<W>();
        const int npv = vt1 - vt0, ncv = n + 1 - fc0;
        for (int e = tid; e < npv * ncv; e += NT)
        {
            const int t = e % npv, col = fc0 + e / npv;
            st2(LOD + ld * col + 2 * (vt0 + t), ld2(lv + ldv * col + 2 * t));
        }
        gsync<W>();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Gauss step of one level in shared memory, lexlse.h:431-471: Left := Left * inv(R), Trailing -= Left * Up
// ---------------------------------------------------------------------------------------------------------------
template <int W, bool LGV = (W == 16)>
__device__ __forceinline__ void gauss_step_smem(const Ctx &c, const int n, const int ld, const int nCtr, const int fr, const int dim,
                                                const int fc, const int rank, const int ColIndex, const int ldv = 0,
                                                const int dmma = 0)
{
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x;
    double *const LOD = c.LOD;
    // R(k, col) / Up(k, col) = RB[k + rl * col].  Global-memory variants (LGV): the level's rows are still in the
    // shared-memory view that qr_level_smem staged (columns >= fc), so only the lower rows stream from global memory.
    const double *const RB = LGV ? c.LVIEW + (fr - 2 * (fr >> 1)) : LOD + fr;
    const int rl           = LGV ? ldv : ld;
    const int frn = fr + dim, Rows = nCtr - frn;
    const int t0 = frn >> 1, t1 = (nCtr + 1) >> 1; // row pairs of the lower levels (the first may straddle)
    // Left := Left * inv(R): one thread per row pair, 4 columns at a time (left-looking across blocks of 4 columns,
    // then the 4x4 triangle); per entry the update order is k = 0,1,2,... as in the column-by-column solve
    if constexpr (LGV)
    {
        // global work matrix: one thread per ROW and 16 columns at a time — every re-read of an L entry is a round trip to L2,
        // and blocks of 16 need 48 of them per row of a 40-row level where blocks of 4 need 180.  Same updates per entry,
        // in the same order (k ascending, then the triangle of the block, then the scaling).
        constexpr int TB = 16; // measured: 8 -> 930 ms, 16 -> 878 ms, 24 -> 948 ms on C4 (registers)
        for (int r = frn + tid; r < nCtr; r += NT)
        {
            double *rp = LOD + r;
            for (int cb = 0; cb < rank; cb += TB)
            {
                int col[TB];
                double acc[TB];
#pragma unroll
                for (int q = 0; q < TB; q++)
                {
                    col[q] = fc + ((cb + q < rank) ? cb + q : rank - 1);
                    acc[q] = rp[ld * col[q]];
                }
                for (int k = 0; k < cb; k++)
                {
                    const double x   = rp[ld * (fc + k)];
                    const double *Rk = RB + k;
#pragma unroll
                    for (int q = 0; q < TB; q++) acc[q] = fma(-x, Rk[rl * col[q]], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < TB; q++)
                    if (cb + q < rank)
                    {
#pragma unroll
                        for (int q2 = 0; q2 < q; q2++) acc[q] = fma(-acc[q2], RB[cb + q2 + rl * col[q]], acc[q]);
                        acc[q] *= c.invd[col[q]];
                        rp[ld * col[q]] = acc[q];
                    }
            }
        }
    }
    else
    for (int t = t0 + tid; t < t1; t += NT)
    {
        double *rp       = LOD + 2 * t;
        const bool full  = (2 * t >= frn);
        for (int cb = 0; cb < rank; cb += 4)
        {
            int col[4];
            double2 acc[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                col[q] = fc + ((cb + q < rank) ? cb + q : rank - 1);
                acc[q] = ld2(rp + ld * col[q]);
            }
            for (int k = 0; k < cb; k++)
            {
                const double2 x  = ld2(rp + ld * (fc + k));
                const double *Rk = RB + k;
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const double rr = Rk[rl * col[q]];
                    acc[q].x        = fma(-x.x, rr, acc[q].x);
                    acc[q].y        = fma(-x.y, rr, acc[q].y);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                if (cb + q < rank)
                {
#pragma unroll
                    for (int q2 = 0; q2 < q; q2++)
                    {
                        const double rr = RB[cb + q2 + rl * col[q]];
                        acc[q].x        = fma(-acc[q2].x, rr, acc[q].x);
                        acc[q].y        = fma(-acc[q2].y, rr, acc[q].y);
                    }
                    const double iv = c.invd[col[q]];
                    acc[q].x *= iv;
                    acc[q].y *= iv;
                    if (full)
                        st2(rp + ld * col[q], acc[q]);
                    else
                        rp[ld * col[q] + 1] = acc[q].y;
                }
            }
        }
    }
    gsync<W>();
    // Trailing -= Left * Up over columns ColIndex .. n.  dmma: 8 x 8 tiles on the FP64 tensor cores (same chain per entry);
    // otherwise work item = (row pair, strip of SW columns); the global-work-matrix
    // kernels use strips of 8: every L entry is an L2 round trip there, and wider strips halve them
    // Measured (profiles/r2_dmma_ab.json): the tensor-core tiles pay off where the blocks are large — the global-work-matrix
    // kernels (C4: 40 pivots x up to 200 columns x up to 360 rows per level, 880 -> 853 ms); for the small shared-memory shapes
    // they are slower (C2 3.40 -> 3.53 ms, C3 unchanged) and the extra routine costs the two-warp kernel its occupancy, so
    // only the LGV kernels carry it.
    if (LGV && dmma)
        schur_dmma<W>(LOD, ld, RB, rl, frn, nCtr, fc, rank, ColIndex, n + 1);
    else
    {
        constexpr int SW = LGV ? 8 : 4;
        const int ncols = n - ColIndex + 1, nstrips = (ncols + SW - 1) / SW, npairs = t1 - t0;
        int s = 0, tt = tid;
        while (tt >= npairs && s < nstrips)
        {
            tt -= npairs;
            s++;
        }
        while (s < nstrips)
        {
            const int t = t0 + tt, j0 = ColIndex + SW * s;
            double *rp      = LOD + 2 * t;
            const bool full = (2 * t >= frn);
            int col[SW];
            double2 acc[SW];
#pragma unroll
            for (int q = 0; q < SW; q++)
            {
                col[q] = (j0 + q <= n) ? j0 + q : n;
                acc[q] = ld2(rp + ld * col[q]);
            }
            for (int k = 0; k < rank; k++)
            {
                const double2 l  = ld2(rp + ld * (fc + k));
                const double *up = RB + k;
#pragma unroll
                for (int q = 0; q < SW; q++)
                {
                    const double u = up[rl * col[q]];
                    acc[q].x       = fma(-l.x, u, acc[q].x);
                    acc[q].y       = fma(-l.y, u, acc[q].y);
                }
            }
#pragma unroll
            for (int q = 0; q < SW; q++)
                if (j0 + q <= n)
                {
                    if (full)
                        st2(rp + ld * col[q], acc[q]);
                    else
                        rp[ld * col[q] + 1] = acc[q].y;
                }
            tt += NT;
            while (tt >= npairs && s < nstrips)
            {
                tt -= npairs;
                s++;
            }
        }
    }
    gsync<W>();
}

// ---------------------------------------------------------------------------------------------------------------
// Regularization of one level, lexlse.h:274-411 (dispatch), :1700-1760 (tikhonov_1), :2078-2138 (tikhonov_2), :2143-2170 (R),
// :2175-2202 (R_NO_Z), :2207-2242 (RT_NO_Z), :2247-2251 (TEST), :2592-2625 (accumulate_nullspace_basis).  A rarely used,
// "not mature" (typedefs.h:178) family: one warp, the work arrays Z (accumulated null-space basis, n x (n+1)) and D (n x (n+1))
// in a global scratch buffer.  Same arithmetic as the oracle: every entry of a product is one sequential fma chain in ascending
// summation index, so entries may be computed by different lanes; the Cholesky factorization is column oriented.
//   R(i,j) = LOD(fr+i, fc+j) for i <= j;  T(i,c) = LOD(fr+i, fc+r+c);  b(i) = LOD(fr+i, n);  q = fc - nFixed rows of Z are live
// ---------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void regularize_level(double *const LOD, const int ld, const int n, const int fr, const int fc, const int r,
                                              const int RC, const int nFixed, const int type, const double factor, const double var_eps,
                                              const double *const invd, double *const scratch, const unsigned cg_iterations = 10)
{
    const int lane = threadIdx.x & 31;
    double *const Z = scratch, *const D = scratch + (size_t)n * (n + 1), *const d = D + (size_t)n * n, *const w0 = D + (size_t)n * (n + 1),
                 *const w1 = w0 + n, *const cgw = w1 + n; // cgw: 9 n doubles for the conjugate-gradient types
    auto L  = [&](int row, int col) -> double & { return LOD[row + (size_t)ld * col]; };
    auto Rm = [&](int i, int j) -> double { return LOD[fr + i + (size_t)ld * (fc + j)]; };
    auto Tm = [&](int i, int c) -> double { return LOD[fr + i + (size_t)ld * (fc + r + c)]; };
    auto Zn = [&](int p, int c) -> double & { return Z[p + (size_t)n * c]; };
    auto Dm = [&](int i, int j) -> double & { return D[i + (size_t)n * j]; };
    const int q = fc - nFixed;

    // the factor, lexlse.h:277-311
    double a = factor;
    if (var_eps != 0.0)
    {
        a = 0.0;
        if (r > 0)
        {
            double n0 = 0.0, n1 = 0.0;
            if (lane == 0)
            {
                for (int i = 0; i < r; i++)
                {
                    w0[i] = L(fr + i, n);
                    n0    = fma(w0[i], w0[i], n0);
                }
                for (int i = r; i--;)
                {
                    double acc = w0[i];
                    for (int k = i + 1; k < r; k++) acc = fma(-Rm(i, k), w0[k], acc);
                    w0[i] = acc / Rm(i, i);
                }
                for (int i = 0; i < r; i++) n1 = fma(w0[i], w0[i], n1);
            }
            n0 = __shfl_sync(0xffffffffu, n0, 0);
            n1 = __shfl_sync(0xffffffffu, n1, 0);
            const double ce = n0 / n1;
            if (ce < var_eps) a = sqrt(1 - (ce * ce) / (var_eps * var_eps)) * factor;
        }
    }
    const bool on   = !(fabs(a - 0.0) < 1e-15);
    const double mu = a * a;

    // Eigen::LLT on the leading N x N block of D (lower triangle), then solveInPlace(d)
    auto llt_solve = [&](int N) {
        for (int j = 0; j < N; j++)
        {
            double acc = Dm(j, j);
            for (int k = 0; k < j; k++) acc = fma(-Dm(j, k), Dm(j, k), acc);
            const double ljj = sqrt(acc);
            __syncwarp();
            if (lane == 0) Dm(j, j) = ljj;
            for (int i = j + 1 + lane; i < N; i += 32)
            {
                double v = Dm(i, j);
                for (int k = 0; k < j; k++) v = fma(-Dm(i, k), Dm(j, k), v);
                Dm(i, j) = v / ljj;
            }
            __syncwarp();
        }
        if (lane == 0)
        {
            for (int i = 0; i < N; i++)
            {
                double v = d[i];
                for (int k = 0; k < i; k++) v = fma(-Dm(i, k), d[k], v);
                d[i] = v / Dm(i, i);
            }
            for (int i = N; i--;)
            {
                double v = d[i];
                for (int k = i + 1; k < N; k++) v = fma(-Dm(k, i), d[k], v);
                d[i] = v / Dm(i, i);
            }
        }
        __syncwarp();
    };
    // b(0:r) = triu(R) d(0:r)
    auto apply_R = [&]() {
        for (int i = lane; i < r; i += 32)
        {
            double acc = 0.0;
            for (int k = i; k < r; k++) acc = fma(Rm(i, k), d[k], acc);
            w0[i] = acc;
        }
        __syncwarp();
        for (int i = lane; i < r; i += 32) L(fr + i, n) = w0[i];
        __syncwarp();
    };
    // column j of the stacked [triu(R) T], row k
    auto colRT = [&](int k, int j) -> double { return j < r ? (k <= j ? Rm(k, j) : 0.0) : Tm(k, j - r); };

    auto tikhonov_1 = [&]() {
        const int N = RC + r;
        for (int e = lane; e < N * N; e += 32)
        {
            const int i = e / N, j = e - i * N;
            if (j > i) continue;
            double acc    = 0.0;
            const int top = (j < r) ? j + 1 : r;
            for (int k = 0; k < top; k++) acc = fma(colRT(k, i), colRT(k, j), acc);
            double up = 0.0;
            for (int p = 0; p < q; p++) up = fma(Zn(p, fc + i), Zn(p, fc + j), up);
            acc      = fma(mu, up, acc);
            Dm(i, j) = (i == j) ? acc + mu : acc;
        }
        for (int i = lane; i < N; i += 32)
        {
            double acc    = 0.0;
            const int top = (i < r) ? i + 1 : r;
            for (int k = 0; k < top; k++) acc = fma(colRT(k, i), L(fr + k, n), acc);
            double up = 0.0;
            for (int p = 0; p < q; p++) up = fma(Zn(p, fc + i), Zn(p, n), up);
            d[i] = fma(mu, up, acc);
        }
        __syncwarp();
        llt_solve(N);
        for (int i = lane; i < r; i += 32)
        {
            double acc = 0.0;
            for (int k = i; k < r; k++) acc = fma(Rm(i, k), d[k], acc);
            for (int c = 0; c < RC; c++) acc = fma(Tm(i, c), d[r + c], acc);
            w0[i] = acc;
        }
        __syncwarp();
        for (int i = lane; i < r; i += 32) L(fr + i, n) = w0[i];
        __syncwarp();
    };
    auto tikhonov_2 = [&]() {
        const int N = q + r, Wd = RC + r;
        for (int e = lane; e < N * N; e += 32)
        {
            const int i = e / N, j = e - i * N;
            if (j > i) continue;
            double acc = 0.0;
            if (i < r) // then j < r as well
            {
                for (int k = i; k < r; k++) acc = fma(Rm(i, k), Rm(j, k), acc);
                for (int c = 0; c < RC; c++) acc = fma(Tm(i, c), Tm(j, c), acc);
            }
            else if (j >= r)
            {
                for (int c = 0; c < Wd; c++) acc = fma(Zn(i - r, fc + c), Zn(j - r, fc + c), acc);
                acc = mu * acc;
            }
            else
            {
                for (int k = j; k < r; k++) acc = fma(Zn(i - r, fc + k), Rm(j, k), acc);
                for (int c = 0; c < RC; c++) acc = fma(Zn(i - r, fc + r + c), Tm(j, c), acc);
                acc = a * acc;
            }
            if (i == j)
            {
                w1[i]    = acc;
                Dm(i, i) = acc + mu;
            }
            else
            {
                Dm(i, j) = acc;
                Dm(j, i) = acc;
            }
        }
        for (int i = lane; i < r; i += 32) d[i] = L(fr + i, n);
        for (int p = lane; p < q; p += 32) d[r + p] = a * Zn(p, n);
        __syncwarp();
        llt_solve(N);
        for (int i = lane; i < r; i += 32)
        {
            double acc = 0.0;
            for (int j = 0; j < N; j++)
            {
                const double e = (i == j) ? ((w1[i] + mu) - mu) : (i > j ? Dm(j, i) : Dm(i, j));
                acc            = fma(e, d[j], acc);
            }
            w0[i] = acc;
        }
        __syncwarp();
        for (int i = lane; i < r; i += 32) L(fr + i, n) = w0[i];
        __syncwarp();
    };
    auto accumulate = [&]() {
        const int rows = q + r;
        for (int e = lane; e < r * r; e += 32) Zn(q + e / r, fc + e % r) = (e / r == e % r) ? 1.0 : 0.0;
        __syncwarp();
        for (int p = lane; p < rows; p += 32) // each lane owns rows of Z: Left := Left * inv(triu(R)), then Trailing -= Left * Up
        {
            for (int c = 0; c < r; c++)
            {
                double acc = Zn(p, fc + c);
                for (int k = 0; k < c; k++) acc = fma(-Zn(p, fc + k), Rm(k, c), acc);
                Zn(p, fc + c) = acc * invd[fc + c];
            }
            for (int j = fc + r; j <= n; j++)
            {
                double acc = Zn(p, j);
                for (int k = 0; k < r; k++) acc = fma(-Zn(p, fc + k), L(fr + k, j), acc);
                Zn(p, j) = acc;
            }
        }
        __syncwarp();
    };

    // CGLS on  min || [triu(R) T; a S; a I] x - [y; a s; 0] ||  from x = 0 (cg_tikhonov lexlse.h:2367-2461: S = Z(:, fc..fc+N),
    // s = Z(:, n); cg_RT :2472-2550: no null-space rows), at most cg_iterations steps or until ||s|| <= 1e-12, then
    // y := [triu(R) T] x.  One lane per output entry ((S) chains from 0, partial products added afterwards), squared norms
    // lane-strided + butterfly = the oracle's tree32 order.
    auto cgls = [&](const int qz) {
        const int N = r + RC, NR = r + qz + N;
        double *const rv = cgw, *const qv = cgw + 3 * n, *const sv = cgw + 6 * n, *const pv = cgw + 7 * n, *const xv = cgw + 8 * n;
        for (int i = lane; i < NR; i += 32) rv[i] = (i < r) ? L(fr + i, n) : (i < r + qz ? Zn(i - r, n) * a : 0.0);
        for (int c = lane; c < N; c += 32) xv[c] = 0.0;
        __syncwarp();
        auto form_s = [&]() {
            for (int c = lane; c < N; c += 32)
            {
                double t = 0.0;
                for (int p = 0; p < qz; p++) t = fma(Zn(p, fc + c), rv[r + p], t);
                t        = (t + rv[r + qz + c]) * a;
                double u = 0.0;
                if (c < r)
                    for (int k = 0; k <= c; k++) u = fma(Rm(k, c), rv[k], u);
                else
                    for (int k = 0; k < r; k++) u = fma(Tm(k, c - r), rv[k], u);
                sv[c] = t + u;
            }
            __syncwarp();
        };
        auto sqnorm = [&](const double *v, const int len) {
            double acc = 0.0;
            for (int i = lane; i < len; i += 32) acc = fma(v[i], v[i], acc);
            return bfly_sum(acc);
        };
        form_s();
        for (int c = lane; c < N; c += 32) pv[c] = sv[c];
        __syncwarp();
        double gamma  = sqnorm(sv, N);
        unsigned iter = 0;
        while (sqrt(gamma) > 1e-12 && iter < cg_iterations)
        {
            for (int i = lane; i < NR; i += 32)
            {
                double v;
                if (i < r)
                {
                    double t = 0.0, u = 0.0;
                    for (int c = 0; c < RC; c++) t = fma(Tm(i, c), pv[r + c], t);
                    for (int k = i; k < r; k++) u = fma(Rm(i, k), pv[k], u);
                    v = t + u;
                }
                else if (i < r + qz)
                {
                    double t = 0.0;
                    for (int c = 0; c < N; c++) t = fma(Zn(i - r, fc + c), pv[c], t);
                    v = t * a;
                }
                else
                    v = a * pv[i - r - qz];
                qv[i] = v;
            }
            __syncwarp();
            const double alpha = gamma / sqnorm(qv, NR);
            for (int c = lane; c < N; c += 32) xv[c] = fma(alpha, pv[c], xv[c]);
            for (int i = lane; i < NR; i += 32) rv[i] = fma(-alpha, qv[i], rv[i]);
            __syncwarp();
            form_s();
            const double gamma_previous = gamma;
            gamma                       = sqnorm(sv, N);
            const double beta           = gamma / gamma_previous;
            for (int c = lane; c < N; c += 32) pv[c] = fma(beta, pv[c], sv[c]);
            __syncwarp();
            iter++;
        }
        for (int i = lane; i < r; i += 32)
        {
            double t = 0.0, u = 0.0;
            for (int k = i; k < r; k++) t = fma(Rm(i, k), xv[k], t);
            for (int c = 0; c < RC; c++) u = fma(Tm(i, c), xv[r + c], u);
            w0[i] = t + u;
        }
        __syncwarp();
        for (int i = lane; i < r; i += 32) L(fr + i, n) = w0[i];
        __syncwarp();
    };

    switch (type)
    {
    case 7: // REGULARIZATION_TIKHONOV_1: regularize_tikhonov_1_test = tikhonov_1 on the solution path (lexlse.h:378-387; its X_mu /
        //     residual_mu by-products are diagnostics nothing on the path reads)
        if (on) tikhonov_1();
        accumulate();
        break;
    case 2: // REGULARIZATION_TIKHONOV_CG
        if (on) cgls(q);
        accumulate();
        break;
    case 6: // REGULARIZATION_RT_NO_Z_CG
        if (on) cgls(0);
        break;
    case 1: // REGULARIZATION_TIKHONOV
        if (on)
        {
            if (fc + r <= RC)
                tikhonov_2();
            else
                tikhonov_1();
        }
        accumulate();
        break;
    case 8: // REGULARIZATION_TIKHONOV_2
        if (on) tikhonov_2();
        accumulate();
        break;
    case 3: // REGULARIZATION_R
        if (on)
        {
            for (int e = lane; e < r * r; e += 32)
            {
                const int i = e / r, j = e - i * r;
                if (j > i) continue;
                double acc = 0.0;
                for (int k = 0; k <= j; k++) acc = fma(Rm(k, i), Rm(k, j), acc);
                double up = 0.0;
                for (int p = 0; p < q; p++) up = fma(Zn(p, fc + i), Zn(p, fc + j), up);
                acc      = fma(mu, up, acc);
                Dm(i, j) = (i == j) ? acc + mu : acc;
            }
            for (int i = lane; i < r; i += 32)
            {
                double up = 0.0;
                for (int p = 0; p < q; p++) up = fma(Zn(p, fc + i), Zn(p, n), up);
                double acc = up * mu;
                for (int k = 0; k <= i; k++) acc = fma(Rm(k, i), L(fr + k, n), acc);
                d[i] = acc;
            }
            __syncwarp();
            llt_solve(r);
            apply_R();
        }
        accumulate();
        break;
    case 4: // REGULARIZATION_R_NO_Z
        if (on)
        {
            for (int e = lane; e < r * r; e += 32)
            {
                const int i = e / r, j = e - i * r;
                if (j > i) continue;
                double acc = 0.0;
                for (int k = 0; k <= j; k++) acc = fma(Rm(k, i), Rm(k, j), acc);
                Dm(i, j) = (i == j) ? acc + mu : acc;
            }
            for (int i = lane; i < r; i += 32)
            {
                double acc = 0.0;
                for (int k = 0; k <= i; k++) acc = fma(Rm(k, i), L(fr + k, n), acc);
                d[i] = acc;
            }
            __syncwarp();
            llt_solve(r);
            apply_R();
        }
        break;
    case 5: // REGULARIZATION_RT_NO_Z
        if (on)
        {
            for (int e = lane; e < r * r; e += 32)
            {
                const int i = e / r, j = e - i * r;
                if (j > i) continue;
                double acc = 0.0;
                for (int k = i; k < r; k++) acc = fma(Rm(i, k), Rm(j, k), acc);
                for (int c = 0; c < RC; c++) acc = fma(Tm(i, c), Tm(j, c), acc);
                if (i == j)
                {
                    w1[i]    = acc;
                    Dm(i, i) = acc + mu;
                }
                else
                {
                    Dm(i, j) = acc;
                    Dm(j, i) = acc;
                }
            }
            for (int i = lane; i < r; i += 32) d[i] = L(fr + i, n);
            __syncwarp();
            llt_solve(r);
            for (int i = lane; i < r; i += 32)
            {
                double acc = 0.0;
                for (int j = 0; j < r; j++)
                {
                    const double e = (i == j) ? ((w1[i] + mu) - mu) : (i > j ? Dm(j, i) : Dm(i, j));
                    acc            = fma(e, d[j], acc);
                }
                w0[i] = acc;
            }
            __syncwarp();
            for (int i = lane; i < r; i += 32) L(fr + i, n) = w0[i];
            __syncwarp();
        }
        break;
    case 9: // REGULARIZATION_TEST
        if (on)
            for (int i = lane; i < r; i += 32) L(fr + i, n) = L(fr + i, n) * a;
        __syncwarp();
        break;
    default: break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// l-QR factorization, lexlse.h:117-506 (REGULARIZATION_NONE)
//   G levels with c.ldim[g] rows starting at c.frow[g]; nCtr = total rows; nFixed fixed variables whose values are in
//   c.lx[0..nFixed) and indices in c.fidx.  Writes c.rank/c.fcol/c.hh/c.invd/c.colperm/c.varofcol; returns TotalRank.
//   LOD is column-major with an EVEN leading dimension ld (ld/2 odd): every column starts 16-byte aligned, so rows are
//   read and written in aligned pairs (2t, 2t+1) with 128-bit accesses, conflict-free both for lanes over row pairs and
//   for lanes over columns.  Rows [nCtr, ld) are zero padding.
// ---------------------------------------------------------------------------------------------------------------
template <int W, bool CQ = false, bool CG = CQ, bool LGV = (W == 16), bool REG = false> __device__ int lse_factorize(const Ctx &c, const SmemMap &smap, const int n, const int ld, const int G,
                                              const int nCtr, const int nFixed, const double tol, FlopCount &fl, const int dmma = 0,
                                              const Params *const reg = nullptr, double *const reg_scratch = nullptr)
{
    const PanelArgs pa = make_panel_args(smap, n, ld, nCtr, tol, dmma);
    // regularization (planned onto the shared-memory routines only, W >= 2): the null-space basis starts from zero, lexlse.h:1685
    // (REG kernels are separate instantiations — kernels_*_reg.cu — so that the default kernels carry none of this)
    constexpr bool regon = REG && (W > 1);
    if constexpr (regon)
    {
        for (int e = threadIdx.x; e < n * (n + 1); e += 32 * W) reg_scratch[e] = 0.0;
        gsync<W>();
    }
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x, lane = tid & 31;
    double *const LOD = c.LOD;
    const int hp      = ld >> 1; // row pairs per column

#pragma unroll 1
    for (int k = tid; k < n; k += NT) c.varofcol[k] = (unsigned short)k;
#pragma unroll 1
    for (int g = tid; g < G; g += NT)
    {
        c.rank[g] = 0;
        c.fcol[g] = 0;
    }
    gsync<W>();

    // ---- fixed variables, lexlse.h:132-156 ----------------------------------------------------------------
    if (nFixed > 0)
    {
        for (int k = 0; k < nFixed; k++)
        {
            const int coeff = c.fidx[k];
            if (k != coeff)
#pragma unroll 1
                for (int t = tid; t < hp; t += NT)
                {
                    const double2 a = ld2(LOD + ld * k + 2 * t);
                    st2(LOD + ld * k + 2 * t, ld2(LOD + ld * coeff + 2 * t));
                    st2(LOD + ld * coeff + 2 * t, a);
                }
            gsync<W>(); // everyone has read fidx[k] before it may be patched
            if (tid == 0)
            {
                c.colperm[k] = (unsigned short)coeff;
                if (k != coeff)
                {
                    unsigned short t  = c.varofcol[k];
                    c.varofcol[k]     = c.varofcol[coeff];
                    c.varofcol[coeff] = t;
                }
                for (int i = k + 1; i < nFixed; i++)
                    if (c.fidx[i] == k)
                    {
                        c.fidx[i] = (unsigned short)coeff;
                        break;
                    }
            }
            gsync<W>();
        }
        // RHS -= LOD[:, 0:nFixed] * x[0:nFixed]   (S)
#pragma unroll 1
        for (int r = tid; r < nCtr; r += NT)
        {
            double acc = LOD[r + ld * n];
            for (int f = 0; f < nFixed; f++) acc = fma(-LOD[r + ld * f], c.lx[f], acc);
            LOD[r + ld * n] = acc;
        }
        gsync<W>();
    }

    int ColIndex = nFixed;
    if (ColIndex >= n) return nFixed; // lexlse.h:164-175

    int sumRank = 0;
    for (int g = 0; g < G; g++)
    {
        const int fr = c.frow[g], dim = c.ldim[g], fc = ColIndex;
        if (tid == 0) c.fcol[g] = fc;

        // warp-per-problem kernels are only planned for shapes that the register-window panel covers (plan() in
        // lexls_b200.cu), so they carry no shared-memory fallback code
        constexpr bool PANEL = (W == 1);
        if constexpr (PANEL)
            qr_level_panel_dispatch<CQ>(pa, fr, dim, ColIndex);
        else
            qr_level_smem<W, LGV, regon>(c, n, ld, fr, dim, ColIndex, tol, c.LVIEW, smap.ldv, regon ? reg_scratch : nullptr);

        const int rank = ColIndex - fc;
        sumRank += rank;
        if (tid == 0) c.rank[g] = rank;
        fl.f += qr_level_flops(dim, n - fc, rank);

        if constexpr (regon)
            {
                // lexlse.h:274-411: between the QR of the level and its Gauss step (the modified right-hand side is what the
                // lower levels see).
                gsync<W>();
                if (threadIdx.x < 32)
                {
                    regularize_level(LOD, ld, n, fr, fc, rank, n - ColIndex, nFixed, reg->reg_type, reg->reg_factor[g], reg->reg_var,
                                     c.invd, reg_scratch, reg->reg_cg_iter);
                    if constexpr (LGV) // the Gauss step reads the level's rows from the shared-memory view: refresh its right-hand side
#pragma unroll 1
                        for (int i = threadIdx.x; i < rank; i += 32) c.LVIEW[(fr - 2 * (fr >> 1)) + i + smap.ldv * n] = LOD[fr + i + ld * n];
                }
                gsync<W>();
            }

        // Gauss step, lexlse.h:431-471
        if (g + 1 < G && rank > 0)
        {
            const int Rows = nCtr - (fr + dim);
            if constexpr (PANEL)
                gauss_step_panel_dispatch<CG>(pa, fr, dim, fc, rank, ColIndex);
            else
                gauss_step_smem<W, LGV>(c, n, ld, nCtr, fr, dim, fc, rank, ColIndex, smap.ldv, dmma);
            const double rr = (double)Rows * (double)rank;
            fl.f += rr * (double)rank + 2.0 * rr * (double)(n - ColIndex + 1);
        }

        if (ColIndex == n) // lexlse.h:475-490
        {
            if (tid == 0)
                for (int k = g + 1; k < G; k++) c.fcol[k] = c.fcol[k - 1] + c.rank[k - 1];
            break;
        }
    }
    gsync<W>();
    fl.f += (double)sumRank * (double)sumRank;
    return nFixed + sumRank;
}

// ---------------------------------------------------------------------------------------------------------------
// back-substitution, lexlse.h:1015-1045; result (already un-permuted, x = P*x) in c.lx and c.lxo
// ---------------------------------------------------------------------------------------------------------------
template <int W, bool LGV = (W == 16)> __device__ void lse_solve(const Ctx &c, const int n, const int ld, const int G, const int ldv = 0)
{
    constexpr int NT  = 32 * W;
    // LGV: work matrix in global memory: each level's R rows are staged in the shared-memory view
    const int tid = threadIdx.x, lane = tid & 31;
    const double *LOD = c.LOD;
    int accr = 0;
    for (int g = G; g--;)
    {
        const int r = c.rank[g];
        if (r == 0) continue;
        const int fr = c.frow[g], fc = c.fcol[g];
        const int fc1 = (accr > 0) ? c.fcol[g + 1] : 0;
        // element (i, col) of the level's first r rows = B[ro + i + lb * col]
        const double *B = LOD;
        int lb = ld, ro = fr;
        if constexpr (LGV)
        {
            const int ncv = n + 1 - fc;
            for (int e = tid; e < r * ncv; e += NT)
            {
                const int i = e % r, col = fc + e / r;
                c.LVIEW[i + ldv * col] = LOD[fr + i + ld * col];
            }
            gsync<W>();
            B  = c.LVIEW;
            lb = ldv;
            ro = 0;
        }
        if (tid < 32)
        {
            if (r <= 32)
            {
                double xi = 0.0, inv = 0.0;
                if (lane < r)
                {
                    double acc       = B[ro + lane + lb * n];
                    const double *bp = B + ro + lane + lb * fc1;
                    const double *xp = c.lx + fc1;
                    int j            = 0;
                    if constexpr (W == 1) // (measured: helps the LexLSE kernel, hurts the two-warp LexLSI kernel)
                    for (; j + 4 <= accr; j += 4) // loads first, then the sequential chain (S)
                    {
                        double b4[4], x4[4];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                        {
                            b4[u] = bp[lb * (j + u)];
                            x4[u] = xp[j + u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) acc = fma(-b4[u], x4[u], acc);
                    }
                    for (; j < accr; j++) acc = fma(-bp[lb * j], xp[j], acc);
                    xi  = acc;
                    inv = c.invd[fc + lane];
                }
                // column-oriented back-substitution: the entry of the next column is loaded one step ahead, the update is
                // computed by every lane and kept by the lanes above the pivot (no divergent branch in the chain)
                const double *Bl = B + ro + ((lane < r) ? lane : 0) + lb * fc;
                double anext     = Bl[lb * (r - 1)];
                for (int j = r; j--;)
                {
                    const double a  = anext;
                    if (j > 0) anext = Bl[lb * (j - 1)];
                    const double t  = xi * inv;
                    const double xj = __shfl_sync(0xffffffffu, t, j);
                    const double u  = fma(-a, xj, xi);
                    xi              = (lane == j) ? xj : ((lane < j) ? u : xi);
                }
                if (lane < r) c.lx[fc + lane] = xi;
            }
            else
            {
                for (int i = lane; i < r; i += 32)
                {
                    double acc = B[ro + i + lb * n];
                    for (int j = 0; j < accr; j++) acc = fma(-B[ro + i + lb * (fc1 + j)], c.lx[fc1 + j], acc);
                    c.lx[fc + i] = acc;
                }
                __syncwarp();
                for (int j = r; j--;)
                {
                    const double xj = c.lx[fc + j] * c.invd[fc + j];
                    __syncwarp();
                    if (lane == 0) c.lx[fc + j] = xj;
                    for (int i = lane; i < j; i += 32) c.lx[fc + i] = fma(-B[ro + i + lb * (fc + j)], xj, c.lx[fc + i]);
                    __syncwarp();
                }
            }
            __syncwarp();
        }
        accr += r;
        if constexpr (LGV) gsync<W>(); // the view is reused by the next level
    }
    gsync<W>();
    // x = P*x (lexlse.h:1044): column c of the factorization belongs to variable varofcol[c]
    for (int k = tid; k < n; k += NT) c.lxo[c.varofcol[k]] = c.lx[k];
    gsync<W>();
    for (int k = tid; k < n; k += NT) c.lx[k] = c.lxo[k];
    gsync<W>();
}

// Q_g * y for LexLSE level g (householderSequence::applyOnTheLeft); y has ldim[g] entries; one warp, all lanes.
// Dots in the (T) order: level row r -> slot (lane) r % 32.  Levels of up to 64 rows keep y in registers (rows lane and
// lane + 32), so one reflector costs two loads, two FMAs, one butterfly and one broadcast.
__device__ __forceinline__ void apply_Q(const Ctx &c, const int ld, const int g, double *y, const int lane)
{
    const int fr = c.frow[g], fc = c.fcol[g], dim = c.ldim[g], rank = c.rank[g];
    if (rank == 0) return;
    if (dim <= 64)
    {
        const int r0 = lane, r1 = lane + 32;
        double y0 = (r0 < dim) ? y[r0] : 0.0, y1 = (r1 < dim) ? y[r1] : 0.0;
        for (int i = rank; i--;)
        {
            const double tau = c.hh[fr + i];
            if (dim - i == 1 || tau == 0.0) continue;
            const double *col = c.LOD + fr + ld * (fc + i);
            const double e0 = (r0 > i && r0 < dim) ? col[r0] : 0.0;
            const double e1 = (r1 > i && r1 < dim) ? col[r1] : 0.0;
            double acc      = fma(e0, y0, 0.0);
            acc             = fma(e1, y1, acc);
            acc             = bfly_sum(acc);
            const double yi = __shfl_sync(0xffffffffu, (i < 32) ? y0 : y1, i & 31);
            const double w  = tau * (acc + yi);
            y0              = fma(-e0, w, y0);
            y1              = fma(-e1, w, y1);
            if (r0 == i) y0 = yi - w;
            if (r1 == i) y1 = yi - w;
        }
        if (r0 < dim) y[r0] = y0;
        if (r1 < dim) y[r1] = y1;
        __syncwarp();
        return;
    }
    for (int i = rank; i--;)
    {
        const int len    = dim - i;
        const double tau = c.hh[fr + i];
        if (len == 1 || tau == 0.0) continue;
        const double *col = c.LOD + fr + ld * (fc + i);
        double acc        = 0.0;
#pragma unroll 1
        for (int r = i + 1 + ((lane - (i + 1)) & 31); r < dim; r += 32) acc = fma(col[r], y[r], acc);
        acc             = bfly_sum(acc);
        const double yi = y[i];
        const double w  = tau * (acc + yi);
        __syncwarp();
        if (lane == 0) y[i] = yi - w;
#pragma unroll 1
        for (int r = i + 1 + lane; r < dim; r += 32) y[r] = fma(-col[r], w, y[r]);
        __syncwarp();
    }
}

// residuals of the factorization, lexlse.h:1560-1582, into c.wsp[0..nCtr)   (warp 0)
template <int W> __device__ void lse_get_v(const Ctx &c, const int n, const int ld, const int G)
{
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32)
    {
        for (int g = 0; g < G; g++)
        {
            const int fr = c.frow[g], dim = c.ldim[g], rank = c.rank[g];
#pragma unroll 1
            for (int i = lane; i < dim; i += 32) c.wsp[fr + i] = (i < rank) ? 0.0 : -c.LOD[fr + i + ld * n];
            __syncwarp();
            if (rank < dim) apply_Q(c, ld, g, c.wsp + fr, lane); // full row rank: Q * 0 = 0 exactly
        }
    }
    gsync<W>();
}

// get_lexqr (lexlse.h:1626): the factorized work matrix and the Householder scalars of problem p go to global memory
template <int NT>
__device__ __forceinline__ void dump_lod(const Ctx &c, double *const qbase, double *const hbase, const unsigned long long p, const int n,
                                         const int ld, const int nCtr, const int M)
{
    double *dst = qbase + p * (unsigned long long)M * (n + 1);
#pragma unroll 1
    for (int e = threadIdx.x; e < M * (n + 1); e += NT)
    {
        const int r = e % M, col = e / M;
        dst[e]      = (r < nCtr) ? c.LOD[r + ld * col] : 0.0;
    }
    if (hbase)
#pragma unroll 1
        for (int k = threadIdx.x; k < M; k += NT) hbase[p * M + k] = (k < nCtr) ? c.hh[k] : 0.0;
}
template <int NT>
__device__ __forceinline__ void dump_lexqr(const Ctx &c, const Buffers &bf, const unsigned long long p, const int n, const int ld,
                                           const int nCtr, const int M)
{
    dump_lod<NT>(c, bf.lexqr, bf.hhout, p, n, ld, nCtr, M);
}

// ---------------------------------------------------------------------------------------------------------------
// Least-norm solution of a factorized hierarchy, lexlse.h:1131-1205 (solveLeastNorm_2; _1 :1051-1124 and _3 :1213-1262 reach the
// same minimiser by Givens rotations / the Tikhonov basis): the free variables z minimise ||x||^2, (I + S'S) z = S's with
// [S s] = inv(R) [T b] over the staircase of all levels.  One warp, work arrays in a global scratch buffer, the oracle's arithmetic
// (sequential fma chains, the triangular solves divide).  xperm receives x in factorization order: [fixed | determined | free].
// ---------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void lse_least_norm(const double *const LOD, const int ld, const int n, const int G, const int *const frow,
                                                   const int *const fcol, const int *const rank, const int nFixed, double *const scratch,
                                                   double *const xperm)
{
    const int lane = threadIdx.x & 31;
    int nVarRank   = 0;
    for (int g = 0; g < G; g++) nVarRank += rank[g];
    const int nVarFree = n - (nVarRank + nFixed), Wc = nVarFree + 1;
    double *const A = scratch, *const D = scratch + (size_t)n * (n + 1), *const z = D + (size_t)n * n;
    auto Am = [&](int i, int j) -> double & { return A[i + (size_t)nVarRank * j]; };
    auto Dd = [&](int i, int j) -> double & { return D[i + (size_t)nVarFree * j]; };
    int counter = 0;
    for (int g = 0; g < G; g++)
    {
        const int fr = frow[g], fc = fcol[g], r = rank[g], cols = n + 1 - fc;
        for (int e = lane; e < r * cols; e += 32)
        {
            const int i = e % r, j = e / r;
            if (j >= i) Am(counter + i, counter + j) = LOD[fr + i + (size_t)ld * (fc + j)];
        }
        counter += r;
    }
    __syncwarp();
    for (int c = lane; c < Wc; c += 32) // [S s] = inv(R) [T b]: one column per lane
        for (int i = nVarRank; i--;)
        {
            double acc = Am(i, nVarRank + c);
            for (int k = i + 1; k < nVarRank; k++) acc = fma(-Am(i, k), Am(k, nVarRank + c), acc);
            Am(i, nVarRank + c) = acc / Am(i, i);
        }
    __syncwarp();
    for (int e = lane; e < nVarFree * nVarFree; e += 32)
    {
        const int i = e / nVarFree, j = e - i * nVarFree;
        if (j > i) continue;
        double acc = 0.0;
        for (int k = 0; k < nVarRank; k++) acc = fma(Am(k, nVarRank + i), Am(k, nVarRank + j), acc);
        Dd(i, j) = (i == j) ? acc + 1.0 : acc;
    }
    for (int i = lane; i < nVarFree; i += 32)
    {
        double acc = 0.0;
        for (int k = 0; k < nVarRank; k++) acc = fma(Am(k, nVarRank + i), Am(k, nVarRank + nVarFree), acc);
        z[i] = acc;
    }
    __syncwarp();
    for (int j = 0; j < nVarFree; j++) // Eigen::LLT, column oriented
    {
        double acc = Dd(j, j);
        for (int k = 0; k < j; k++) acc = fma(-Dd(j, k), Dd(j, k), acc);
        const double ljj = sqrt(acc);
        __syncwarp();
        if (lane == 0) Dd(j, j) = ljj;
        for (int i = j + 1 + lane; i < nVarFree; i += 32)
        {
            double v = Dd(i, j);
            for (int k = 0; k < j; k++) v = fma(-Dd(i, k), Dd(j, k), v);
            Dd(i, j) = v / ljj;
        }
        __syncwarp();
    }
    if (lane == 0)
    {
        for (int i = 0; i < nVarFree; i++)
        {
            double v = z[i];
            for (int k = 0; k < i; k++) v = fma(-Dd(i, k), z[k], v);
            z[i] = v / Dd(i, i);
        }
        for (int i = nVarFree; i--;)
        {
            double v = z[i];
            for (int k = i + 1; k < nVarFree; k++) v = fma(-Dd(k, i), z[k], v);
            z[i] = v / Dd(i, i);
        }
    }
    __syncwarp();
    for (int c = lane; c < nVarFree; c += 32) xperm[nFixed + nVarRank + c] = z[c];
    counter = 0;
    for (int g = 0; g < G; g++)
    {
        const int fr = frow[g], r = rank[g];
        for (int i = lane; i < r; i += 32)
        {
            double acc = LOD[fr + i + (size_t)ld * n];
            for (int c = 0; c < nVarFree; c++) acc = fma(-LOD[fr + i + (size_t)ld * (nVarRank + nFixed + c)], z[c], acc);
            xperm[nFixed + counter + i] = acc;
        }
        counter += r;
    }
    __syncwarp();
    if (lane == 0)
        for (int i = nVarRank; i--;)
        {
            double acc = xperm[nFixed + i];
            for (int k = i + 1; k < nVarRank; k++) acc = fma(-Am(i, k), xperm[nFixed + k], acc);
            xperm[nFixed + i] = acc / Am(i, i);
        }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// LexLSE kernel: one problem per group, record staged by one TMA bulk copy
// ---------------------------------------------------------------------------------------------------------------
#ifndef LX_LSE_W2_MINB
#define LX_LSE_W2_MINB 8
#endif
template <int W, bool LG, bool REG = false>
__global__ void __launch_bounds__(32 * W, W == 1 ? 10 : (W == 8 ? 2 : (W == 4 ? 4 : (W == 2 ? LX_LSE_W2_MINB : 1)))) lse_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Params pr,
                                                      const __grid_constant__ SmemMap sm, const __grid_constant__ Buffers bf)
{
    constexpr int NT = 32 * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Ctx c = make_ctx(smem_raw, sm);
    if constexpr (LG) c.LOD = bf.lod_scratch + blockIdx.x * bf.lod_stride;
    const int tid = threadIdx.x;
    const int n = sh.n, ld = sh.ld, G = sh.G, nCtr = sh.M;
    // the set_* calls found invalid input (lb > ub, duplicate VarIndex): the reference throws in setData; here the solve is
    // refused — every status stays TERMINATION_STATUS_UNKNOWN and the host reports the error at its next synchronisation
    if (bf.flag && *bf.flag)
    {
        for (unsigned long long q = blockIdx.x * (unsigned long long)NT + threadIdx.x; q < bf.batch; q += (unsigned long long)gridDim.x * NT)
            bf.status[q] = ST_UNKNOWN;
        return;
    }

    if (tid == 0)
    {
        mbar_init(c.mbar, 1);
        fence_mbar_init();
    }
    for (int g = tid; g < G; g += NT)
    {
        c.ldim[g] = sh.dim[g];
        c.frow[g] = sh.goff[g];
    }
    if constexpr (!LG)
        if (tid < 4) c.LOD[ld * (n + 1) + tid] = 0.0; // zero pad behind LOD (see build_map)
    gsync<W>();

    // The record of the NEXT problem is requested (TMA bulk copy) as soon as the work matrix of the current one is dead, so
    // the copy overlaps the output stores and the queue round trip.
    auto fetch_and_load = [&]() {
        if (tid == 0)
        {
            const unsigned long long q       = atomicAdd(bf.queue, 1ull);
            *(unsigned long long *)c.scal = q;
            if constexpr (!LG)
                if (q < bf.batch)
                {
                    fence_proxy_async();
                    tma_load_record(c.LOD, bf.rec + q * bf.rec_stride, (unsigned)(bf.rec_stride * sizeof(double)), c.mbar);
                }
        }
    };
    unsigned parity = 0;
    fetch_and_load();
    gsync<W>();
    unsigned long long p = *(unsigned long long *)c.scal;
    while (p < bf.batch)
    {
        if constexpr (LG)
        {
            const double *src = bf.rec + p * bf.rec_stride;
            for (unsigned long long e = tid; e < bf.rec_stride; e += NT) c.LOD[e] = src[e];
        }
        for (int k = tid; k < nCtr; k += NT) c.hh[k] = 0.0;
        for (int k = tid; k < n; k += NT) c.lx[k] = 0.0;
        gsync<W>();
        for (int f = tid; f < sh.nfix; f += NT)
        {
            c.fidx[f] = (unsigned short)bf.fidx[p * sh.nfix + f];
            c.lx[f]   = bf.fval[p * sh.nfix + f];
        }
        if constexpr (!LG)
        {
            mbar_wait(c.mbar, parity);
            parity ^= 1;
        }
        gsync<W>();

        FlopCount fl;
        fl.f                = 0;
        const int totalRank = lse_factorize<W, LX_LSE_COMPACT_QR, LX_LSE_COMPACT_GS, LG, REG>(c, sm, n, ld, G, nCtr, sh.nfix, pr.tol_lindep, fl, pr.dmma, &pr,
                                                                                 bf.reg_scratch ? bf.reg_scratch + blockIdx.x * bf.reg_stride : nullptr);
        lse_solve<W, LG>(c, n, ld, G, sm.ldv);
        for (int k = tid; k < n; k += NT) bf.x[p * n + k] = c.lx[k];
        if (bf.v) lse_get_v<W>(c, n, ld, G);
        if (bf.lexqr) dump_lexqr<NT>(c, bf, p, n, ld, nCtr, nCtr);
        if (bf.xln)
        {
            // solveLeastNorm (lexlse.h:1046-1262) next to the basic solution: warp 0, x in factorization order in the scratch, then
            // the same un-permutation as solve()
            gsync<W>();
            double *const sc = bf.ln_scratch + blockIdx.x * bf.ln_stride, *const xperm = sc + (size_t)2 * n * (n + 1) + n;
            if (tid < 32)
            {
                for (int f = tid; f < sh.nfix; f += 32) xperm[f] = bf.fval[p * sh.nfix + f];
                lse_least_norm(c.LOD, ld, n, G, c.frow, c.fcol, c.rank, sh.nfix, sc, xperm);
                for (int k = tid; k < n; k += 32) bf.xln[p * n + c.varofcol[k]] = xperm[k];
            }
        }
        gsync<W>(); // the work matrix is dead: every thread is done with it
        fetch_and_load();
        if (bf.v)
            for (int k = tid; k < nCtr; k += NT) bf.v[p * nCtr + k] = c.wsp[k];
        for (int g = tid; g < G; g += NT) bf.ranks[p * G + g] = c.rank[g];
        if (bf.colperm)
            for (int k = tid; k < n; k += NT) bf.colperm[p * n + k] = (k < totalRank) ? c.colperm[k] : 0u;
        if (tid == 0)
        {
            bf.status[p]    = ST_SOLVED;
            unsigned *cnt   = bf.counters + p * 8;
            cnt[0]          = 1;
            cnt[1] = cnt[2] = cnt[3] = cnt[4] = 0;
            cnt[5]                            = (unsigned)totalRank;
            cnt[6]                            = (unsigned)nCtr;
            cnt[7]                            = (unsigned)(((unsigned long long)fl.f + 500ull) / 1000ull);
        }
        gsync<W>();
        p = *(unsigned long long *)c.scal;
    }
}

#ifdef LEXLS_B200_WITH_PAIR
// ---------------------------------------------------------------------------------------------------------------
// LexLSE, two problems per warp (lexls_pair.cuh): each half warp factorizes its own problem in its own shared-memory
// slice; solve / residuals / outputs then run per problem on the full warp with the ordinary routines.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lse_factorize_pair(unsigned char *smem, const SmemMap &smap, const unsigned slice, const int n,
                                                  const int ld, const int G, const int nCtr, const double tol, const bool on,
                                                  FlopCount &fl)
{
    PairArgs pa;
    pa.lod = smap.lod; pa.ptmp = smap.ptmp; pa.invd = smap.invd; pa.hh = smap.hh; pa.colperm = smap.colperm;
    pa.varofcol = smap.varofcol; pa.slice = slice; pa.n = n; pa.ld = ld; pa.nCtr = nCtr; pa.tol = tol;
    const int lane = threadIdx.x & 31, hb = lane >> 4, cl = lane & 15;
    const Ctx c = make_ctx(smem + hb * slice, smap);
    for (int k = cl; k < n; k += 16) c.varofcol[k] = (unsigned short)k;
    for (int g = cl; g < G; g += 16)
    {
        c.rank[g] = 0;
        c.fcol[g] = 0;
    }
    __syncwarp();
    int ColIndex = 0, sumRank = 0;
    for (int g = 0; g < G; g++)
    {
        const int fr = c.frow[g], dim = c.ldim[g], fc = ColIndex;
        if (cl == 0) c.fcol[g] = fc;
        const bool go = on && ColIndex < n && dim > 0;
        if (__any_sync(0xffffffffu, go))
        {
            const int ncmax = (int)__reduce_max_sync(0xffffffffu, go ? (unsigned)(n + 1 - ColIndex) : 0u);
            const int S     = (ncmax + 15) >> 4;
#define LX_PQR(NGv)                                                                                                     \
    {                                                                                                                   \
        if (S == 1)                                                                                                     \
            ColIndex = qr_level_pair<NGv, 1>(pa, fr, dim, ColIndex, go);                                                \
        else if (S == 2)                                                                                                \
            ColIndex = qr_level_pair<NGv, 2>(pa, fr, dim, ColIndex, go);                                                \
        else                                                                                                            \
            ColIndex = qr_level_pair<NGv, 3>(pa, fr, dim, ColIndex, go);                                                \
    }
            if (dim <= 12) LX_PQR(3) else LX_PQR(5)
#undef LX_PQR
        }
        const int rank = ColIndex - fc;
        sumRank += rank;
        if (cl == 0) c.rank[g] = rank;
        {
            // algorithmic flops of this level's QR in closed form
            const long long m = dim, cc = n - fc, r = rank, rp = (r < m) ? r : m - 1;
            const long long T1 = rp * (rp - 1) / 2, T2 = (rp - 1) * rp * (2 * rp - 1) / 6, R1 = r * (r - 1) / 2;
            long long f = 2 * m * cc + 2 * (r * m - R1) + 2 * (r * (cc - 1) - R1);
            if (rp > 0) f += 3 * (rp * m - T1) + 4 * rp + 4 * (rp * m * cc - (m + cc) * T1 + T2);
            if (go) fl.f += (unsigned long long)f;
        }
        if (g + 1 < G && __any_sync(0xffffffffu, on && rank > 0) && fr + dim < nCtr)
        {
            const int rk    = on ? rank : 0;
            const int rkmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)rk);
            const bool odd  = fr & 1;
#define LX_PGS(NBv)                                                                                                     \
    {                                                                                                                   \
        if (odd)                                                                                                        \
            gauss_rows_pair<NBv, true>(pa, fr + dim, fr, fc, rk, ColIndex);                                             \
        else                                                                                                            \
            gauss_rows_pair<NBv, false>(pa, fr + dim, fr, fc, rk, ColIndex);                                            \
    }
            if (rkmax <= 12) LX_PGS(3) else LX_PGS(5)
#undef LX_PGS
            const int Rows = nCtr - (fr + dim);
            if (on && rank > 0) fl.f += (unsigned long long)Rows * rank * rank + 2ull * Rows * rank * (n - ColIndex + 1);
        }
    }
    __syncwarp();
    fl.f += (unsigned long long)sumRank * sumRank;
    return sumRank;
}

template <int UNUSED>
__global__ void __launch_bounds__(32, 5) lse_pair_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Params pr,
                                                         const __grid_constant__ SmemMap sm, const __grid_constant__ Buffers bf)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned slice = sm.total;
    const int tid = threadIdx.x, hb = tid >> 4, cl = tid & 15;
    const int n = sh.n, ld = sh.ld, G = sh.G, nCtr = sh.M;
    const Ctx c0 = make_ctx(smem_raw, sm);
    {
        const Ctx ch = make_ctx(smem_raw + hb * slice, sm);
        if (tid == 0)
        {
            mbar_init(c0.mbar, 1);
            fence_mbar_init();
        }
        for (int g = cl; g < G; g += 16)
        {
            ch.ldim[g] = sh.dim[g];
            ch.frow[g] = sh.goff[g];
        }
        if (cl < 4) ch.LOD[ld * (n + 1) + cl] = 0.0; // zero pad behind LOD (see build_map)
        for (int k = cl; k < LX_PANEL_TMP_DOUBLES; k += 16) ch.ptmp[k] = 0.0; // an idle half multiplies stale staging data by 0
    }
    __syncwarp();

    unsigned parity = 0;
    for (;;)
    {
        if (tid == 0) *(unsigned long long *)c0.scal = atomicAdd(bf.queue, 2ull);
        __syncwarp();
        const unsigned long long q = *(unsigned long long *)c0.scal;
        if (q >= bf.batch) break;
        const bool two             = q + 1 < bf.batch;
        const unsigned long long p = q + hb;
        const bool on              = hb == 0 || two;
        const Ctx ch               = make_ctx(smem_raw + hb * slice, sm);
        if (tid == 0)
        {
            const unsigned bytes = (unsigned)(bf.rec_stride * sizeof(double));
            fence_proxy_async();
            mbar_expect_tx(c0.mbar, two ? 2 * bytes : bytes);
            tma_bulk_g2s(c0.LOD, bf.rec + q * bf.rec_stride, bytes, c0.mbar);
            if (two) tma_bulk_g2s(smem_raw + slice + sm.lod, bf.rec + (q + 1) * bf.rec_stride, bytes, c0.mbar);
        }
        for (int k = cl; k < nCtr; k += 16) ch.hh[k] = 0.0;
        for (int k = cl; k < n; k += 16) ch.lx[k] = 0.0;
        __syncwarp();
        mbar_wait(c0.mbar, parity);
        parity ^= 1;
        __syncwarp();

        FlopCount fl;
        fl.f               = 0;
        const int totalRank = lse_factorize_pair(smem_raw, sm, slice, n, ld, G, nCtr, pr.tol_lindep, on, fl);
        if (cl == 0)
        {
            ch.scal[4]                        = totalRank;
            *(unsigned long long *)(ch.scal + 6) = (unsigned long long)fl.f;
        }
        __syncwarp();
        (void)p;
        // back-substitution, residuals and outputs: one problem at a time on the whole warp
        for (int h = 0; h < (two ? 2 : 1); h++)
        {
            const Ctx c                = make_ctx(smem_raw + h * slice, sm);
            const unsigned long long pp = q + h;
            lse_solve<1>(c, n, ld, G);
            for (int k = tid; k < n; k += 32) bf.x[pp * n + k] = c.lx[k];
            if (bf.v)
            {
                lse_get_v<1>(c, n, ld, G);
                for (int k = tid; k < nCtr; k += 32) bf.v[pp * nCtr + k] = c.wsp[k];
            }
            for (int g = tid; g < G; g += 32) bf.ranks[pp * G + g] = c.rank[g];
            const int tr = c.scal[4];
            if (bf.colperm)
                for (int k = tid; k < n; k += 32) bf.colperm[pp * n + k] = (k < tr) ? c.colperm[k] : 0u;
            if (bf.lexqr) dump_lexqr<32>(c, bf, pp, n, ld, nCtr, nCtr);
            if (tid == 0)
            {
                bf.status[pp]   = ST_SOLVED;
                unsigned *cnt   = bf.counters + pp * 8;
                cnt[0]          = 1;
                cnt[1] = cnt[2] = cnt[3] = cnt[4] = 0;
                cnt[5]                            = (unsigned)tr;
                cnt[6]                            = (unsigned)nCtr;
                cnt[7]                            = (unsigned)((*(unsigned long long *)(c.scal + 6) + 500) / 1000);
            }
            __syncwarp();
        }
    }
}

#endif // LEXLS_B200_WITH_PAIR

// ===============================================================================================================
// LexLSI
// ===============================================================================================================

// working-set operations, workingset.h:57-115 (one thread)
__device__ __forceinline__ void ws_activate(const Ctx &c, const Shape &sh, int k, int i, int type)
{
    const int o = sh.uoff[k];
    int ni      = c.ninact[k];
    // the position is unique; the two guesses cover the sequences the initial working set produces (ascending
    // activations with swap-with-last: an element is still at its own index, or it was the last one moved forward)
    int q = 0;
    if (i < ni && c.inactive[o + i] == i)
        q = i;
    else if (ni > 0 && c.inactive[o + ni - 1] == i)
        q = ni - 1;
    else
        while (q < ni && c.inactive[o + q] != i) q++;
    c.inactive[o + q] = c.inactive[o + ni - 1];
    c.ninact[k]       = ni - 1;
    c.alltype[o + i]  = (unsigned char)type;
    int na            = c.nact[k];
    c.active[o + na]  = (unsigned short)i;
    c.atype[o + na]   = (unsigned char)type;
    c.nact[k]         = na + 1;
}
__device__ __forceinline__ void ws_deactivate(const Ctx &c, const Shape &sh, int k, int a)
{
    const int o  = sh.uoff[k];
    const int i  = c.active[o + a];
    const int na = c.nact[k];
    for (int t = a; t + 1 < na; t++)
    {
        c.active[o + t] = c.active[o + t + 1];
        c.atype[o + t]  = c.atype[o + t + 1];
    }
    c.nact[k]                   = na - 1;
    c.alltype[o + i]            = CTR_INACTIVE;
    c.inactive[o + c.ninact[k]] = (unsigned short)i;
    c.ninact[k]++;
}

// bounds of row i of user level k
template <bool DS> struct ObjView
{
    const Ctx &c;
    const Shape &sh;
    double *obj; // general rows [M x (n+2)], ld = ldo
    double *sb;  // [2*m0]
    __device__ __forceinline__ double A(int grow, int col) const { return obj[grow + sh.ldo * col]; }
    __device__ __forceinline__ double &lb(int k, int i) const
    {
        if (sh.off && k == 0) return sb[i];
        return obj[sh.goff[k - sh.off] + i + sh.ldo * sh.n];
    }
    __device__ __forceinline__ double &ub(int k, int i) const
    {
        if (sh.off && k == 0) return sb[sh.m0 + i];
        return obj[sh.goff[k - sh.off] + i + sh.ldo * (sh.n + 1)];
    }
};

// out[all rows] = A * vec (general levels, (S) per row) or a gather (simple bounds); objective.h:242-273
template <int W, bool DS>
__device__ __forceinline__ void obj_multiply(const Ctx &c, const Shape &sh, const ObjView<DS> &ov, const double *vec, double *out)
{
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x;
    for (int i = tid; i < sh.m0; i += NT) out[i] = vec[c.sbidx[i]];
    for (int r = tid; r < sh.M; r += NT)
    {
        double acc = 0.0;
        int col    = 0;
        for (; col + 8 <= sh.n; col += 8) // loads first (L2 latency), then the sequential chain (S)
        {
            double t8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) t8[u] = ov.A(r, col + u);
#pragma unroll
            for (int u = 0; u < 8; u++) acc = fma(t8[u], vec[col + u], acc);
        }
        for (; col < sh.n; col++) acc = fma(ov.A(r, col), vec[col], acc);
        out[sh.m0 + r] = acc;
    }
}

// Objective::formStep, objective.h:288-338 (Adx must be current)
template <int W, bool DS> __device__ __forceinline__ void obj_form_dv(const Ctx &c, const Shape &sh, const ObjView<DS> &ov)
{
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x;
    for (int k = 0; k < sh.P; k++)
    {
        const int o = sh.uoff[k];
#pragma unroll 1
        for (int i = tid; i < sh.dim[k]; i += NT)
        {
            const int t = c.alltype[o + i];
            double d    = -c.v[o + i];
            if (t != CTR_INACTIVE)
            {
                const double rhs = (t == CTR_ACTIVE_LB) ? ov.lb(k, i) : ov.ub(k, i);
                d                = d + ((c.Ax[o + i] + c.Adx[o + i]) - rhs);
            }
            c.dv[o + i] = d;
        }
    }
}

// Real calls for what the LexLSI kernel needs at several places (inlined copies cost instruction-cache misses, see build.py):
// the context is rebuilt from the shared-memory map; vectors that may live outside shared memory (SmemMap::vecg) are passed in.
static __device__ __noinline__ void ws_activate_call(const Shape &sh, const SmemMap &sm, int k, int i, int type)
{
    extern __shared__ __align__(128) unsigned char smem_ws[];
    const Ctx c = make_ctx(smem_ws, sm);
    ws_activate(c, sh, k, i, type);
}
static __device__ __noinline__ void ws_deactivate_call(const Shape &sh, const SmemMap &sm, int k, int a)
{
    extern __shared__ __align__(128) unsigned char smem_ws[];
    const Ctx c = make_ctx(smem_ws, sm);
    ws_deactivate(c, sh, k, a);
}
template <int W, bool DS>
__device__ __noinline__ void obj_multiply_call(const Shape &sh, const SmemMap &sm, double *objp, double *sbp, const unsigned vec_off, double *out)
{
    extern __shared__ __align__(128) unsigned char smem_om[];
    const Ctx c = make_ctx(smem_om, sm);
    const ObjView<DS> ov{c, sh, objp, sbp};
    obj_multiply<W, DS>(c, sh, ov, (const double *)(smem_om + vec_off), out);
}
// Adx = A dx and dv (formStep, objective.h:288-338) of the current direction dx
template <int W, bool DS>
__device__ __noinline__ void lsi_direction(const Shape &sh, const SmemMap &sm, double *objp, double *sbp, double *v, double *dv, double *Ax,
                                           double *Adx)
{
    extern __shared__ __align__(128) unsigned char smem_om[];
    Ctx c = make_ctx(smem_om, sm);
    c.v   = v;
    c.dv  = dv;
    c.Ax  = Ax;
    c.Adx = Adx;
    const ObjView<DS> ov{c, sh, objp, sbp};
    obj_multiply<W, DS>(c, sh, ov, c.dx, c.Adx);
    gsync<W>();
    obj_form_dv<W, DS>(c, sh, ov);
    gsync<W>();
}

// LexLSI::formLexLSE + Objective::formLexLSE, lexlsi.h:968-982, objective.h:434-494; returns nCtr, sets nFixed
template <int W, bool DS>
__device__ __forceinline__ int lsi_form_lexlse(const Ctx &c, const Shape &sh, const ObjView<DS> &ov, int &nFixed, const int nFixedPrev)
{
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x;
    const int n = sh.n, ld = sh.ld;
    // setObjDim -> initialize(): hh = 0, x.tail(nVar - OLD nVarFixed) = 0 (quirk kept, see oracle)
    int nCtr = 0;
    for (int g = 0; g < sh.G; g++)
    {
        const int d = c.nact[g + sh.off];
        if (tid == 0)
        {
            c.ldim[g] = d;
            c.frow[g] = nCtr;
        }
        nCtr += d;
    }
#pragma unroll 1
    for (int k = tid; k < nCtr; k += NT) c.hh[k] = 0.0;
#pragma unroll 1
    for (int k = nFixedPrev + tid; k < n; k += NT) c.lx[k] = 0.0;
    // rows are processed in aligned pairs: the partner of the last row must be a zero row, not a stale one
    if (nCtr & 1)
#pragma unroll 1
        for (int col = tid; col <= n; col += NT) c.LOD[nCtr + ld * col] = 0.0;
    gsync<W>();
    nFixed = 0;
    if (sh.off)
    {
        nFixed = c.nact[0];
#pragma unroll 1
        for (int a = tid; a < nFixed; a += NT)
        {
            const int i = c.active[a], t = c.atype[a];
            c.fidx[a]  = c.sbidx[i];
            c.ftype[a] = (unsigned char)t;
            c.lx[a]    = (t == CTR_ACTIVE_LB) ? ov.sb[i] : ov.sb[sh.m0 + i];
        }
    }
    // source row (in the stacked objective record) and type of every LexLSE row; the multiplier workspace is free here
    int *const srow = (int *)c.wsp;
    for (int g = 0; g < sh.G; g++)
    {
        const int k = g + sh.off, o = sh.uoff[k], go = sh.goff[g], fr = c.frow[g], d = c.ldim[g];
#pragma unroll 1
        for (int a = tid; a < d; a += NT)
        {
            c.ctype[fr + a] = c.atype[o + a];
            srow[fr + a]    = go + c.active[o + a];
        }
    }
    gsync<W>();
    // the gather over ALL rows at once (a loop per level keeps dim[g] of the NT threads busy, one level after the other:
    // 10 % of the C4 factorization time): consecutive threads take consecutive LexLSE rows — coalesced on both sides, the
    // sources of a level are mostly consecutive too — and one item is eight columns, eight loads in flight (the objective
    // record may live in L2); the right-hand side column is the active bound
    const int nblk = (n + 8) >> 3;
#pragma unroll 1
    for (int e = tid; e < nCtr * nblk; e += NT)
    {
        const int r = e % nCtr, c0 = (e / nCtr) << 3;
        const int sr = srow[r];
        double *dst  = c.LOD + r + ld * c0;
        if (c0 + 8 <= n)
        {
            double t8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) t8[u] = ov.A(sr, c0 + u);
#pragma unroll
            for (int u = 0; u < 8; u++) dst[ld * u] = t8[u];
        }
        else
        {
#pragma unroll 1
            for (int col = c0; col < n; col++) c.LOD[r + ld * col] = ov.A(sr, col);
            c.LOD[r + ld * n] = (c.ctype[r] == CTR_ACTIVE_LB) ? ov.A(sr, n) : ov.A(sr, n + 1);
        }
    }
    gsync<W>();
    return nCtr;
}

// Lagrange multipliers of LexLSE objective j and the sign test, lexlse.h:611-762 + :935-987   (warp 0 only)
// returns true when a constraint to drop was found; obj2rem = LexLSE level (or -1 for the fixed variables)
__device__ __forceinline__ bool lse_sensitivity(const Ctx &c, const int n, const int ld, const int j, const int nFixed,
                                                const double tol_wrong, const double tol_correct, int &ctr2rem, int &obj2rem,
                                                const int lane, double *lam2rem = nullptr)
{
    int nLambda = 0, nRank = 0;
    for (int k = 0; k < j; k++)
    {
        nLambda += c.ldim[k];
        nRank += c.rank[k];
    }
    nLambda += c.ldim[j];
    const int total = nFixed + nLambda + nRank + nFixed;
    for (int i = lane; i < total; i += 32) c.wsp[i] = 0.0;
    double *LambdaFixed = c.wsp;
    double *Lambda      = c.wsp + nFixed;
    double *rhs         = c.wsp + nFixed + nLambda;
    __syncwarp();
    // an objective with full row rank has a zero residual: its multipliers are Q * 0 = 0 for every constraint, nothing has a
    // wrong sign and no type is re-marked (the workspace, all zeros, is what getLambda reads)
    if (c.rank[j] == c.ldim[j]) return false;

    ArgMin best;
    best.v   = 0.0;        // maxAbsValue starts at 0 (lexlse.h:618); candidates are strictly below -tol_wrong
    best.pos = 0x7fffffff; // scan position: (group order << 16) | k

    // One loop over the groups in scan order — objective j (order 0), the objectives above it (order j - k), the fixed
    // variables (order j + 1, "k = -1") — so that the Householder sequence, the transposed product and the sign test exist
    // ONCE in the code (the kernel is bound by instruction fetch); the two products of a group are independent of its sign test.
    for (int k = j; k >= -1; k--)
    {
        int first, dim;
        const double *lam;
        unsigned char *types;
        if (k >= 0)
        {
            const int fr = c.frow[k], fc = c.fcol[k], rank = c.rank[k];
            dim = c.ldim[k];
            if (k == j) // residual of objective j below its rank: Lambda_j = Q [0; -b2], lexlse.h:640-648
            {
#pragma unroll 1
                for (int i = rank + lane; i < dim; i += 32) Lambda[fr + i] = -c.LOD[fr + i + ld * n];
            }
            else
            {
#pragma unroll 1
                for (int i = lane; i < rank; i += 32) Lambda[fr + i] = rhs[fc + i];
            }
            __syncwarp();
            apply_Q(c, ld, k, Lambda + fr, lane);
            if (j > 0)
            {
                // rhs[0:fc] -= LOD(fr.., 0:fc)^T * Lambda[fr..]   (S) per column
#pragma unroll 1
                for (int col = lane; col < fc; col += 32)
                {
                    const double *colp = c.LOD + ld * col + fr;
                    double acc         = rhs[col];
#pragma unroll 4
                    for (int i = 0; i < dim; i++) acc = fma(-colp[i], Lambda[fr + i], acc);
                    rhs[col] = acc;
                }
            }
            first = fr;
            lam   = Lambda;
            types = c.ctype;
        }
        else
        {
            if (nFixed <= 0) break;
#pragma unroll 1
            for (int f = lane; f < nFixed; f += 32)
            {
                const double *colp = c.LOD + ld * f;
                double acc         = 0.0;
#pragma unroll 4
                for (int i = 0; i < nLambda; i++) acc = fma(colp[i], Lambda[i], acc);
                LambdaFixed[f] = -acc;
            }
            __syncwarp();
            first = 0;
            dim   = nFixed;
            lam   = LambdaFixed;
            types = c.ftype;
        }
        // findDescentDirection, lexlse.h:935-987
        const int order = j - k;
#pragma unroll 1
        for (int q = lane; q < dim; q += 32)
        {
            const int t = types[first + q];
            if (t != CTR_ACTIVE_EQ && t != CORRECT_SIGN_OF_LAMBDA)
            {
                double a = lam[first + q];
                if (t == CTR_ACTIVE_LB) a = -a;
                if (a > tol_correct)
                    types[first + q] = CORRECT_SIGN_OF_LAMBDA;
                else if (a < -tol_wrong)
                {
                    const int pos = (order << 16) | q;
                    if (a < best.v || (a == best.v && pos < best.pos))
                    {
                        best.v   = a;
                        best.pos = pos;
                    }
                }
            }
        }
        __syncwarp();
    }
    best = bfly_argmin(best);
    __syncwarp();
    if (best.pos == 0x7fffffff) return false;
    if (lam2rem) *lam2rem = best.v; // lambda_wrong_sign of the working-set log
    const int order = best.pos >> 16;
    ctr2rem         = best.pos & 0xffff;
    obj2rem         = (order == j + 1) ? -1 : (j - order);
    return true;
}

// formLexLSE + factorize + solve of one active-set iteration as ONE real call: the LexLSI kernel needs it at three places
// (phase 1, the first step from a user guess, the loop), and three inlined copies do not fit the instruction cache.
// The context is rebuilt from the shared-memory map (kernel parameter, __grid_constant__: its address may be taken).
template <int W, bool DS, bool LG, bool REG = false>
__device__ __noinline__ void lsi_factor_and_solve(const Shape &sh, const SmemMap &sm, const double tol, double *lod, double *objp,
                                                  double *sbp, int &nFixed, int &nFixedPrev, int &totalRank, FlopCount &fl,
                                                  const int dmma, double *const data_dump = nullptr, const Params *const reg = nullptr,
                                                  double *const reg_scratch = nullptr)
{
    extern __shared__ __align__(128) unsigned char smem_fs[];
    Ctx c = make_ctx(smem_fs, sm);
    c.LOD = lod;
    ObjView<DS> ov{c, sh, objp, sbp};
    const int nCtr = lsi_form_lexlse<W, DS>(c, sh, ov, nFixed, nFixedPrev);
    nFixedPrev     = nFixed;
    if (data_dump) // PROBLEM_DATA = LOD on entry of factorize(), lexlse.h:119 (data_dump already points at this problem)
    {
        dump_lod<32 * W>(c, data_dump, nullptr, 0ull, sh.n, sh.ld, nCtr, sh.M);
        gsync<W>();
    }
    totalRank      = lse_factorize<W, true, true, LG, REG>(c, sm, sh.n, sh.ld, sh.G, nCtr, nFixed, tol, fl, dmma, reg, reg_scratch);
    lse_solve<W, LG>(c, sh.n, sh.ld, sh.G, sm.ldv);
}

// two-warp LexLSI kernel (C3): shared memory keeps eight problems per SM, and eight is also the minimum block count the kernel
// is compiled for (ptxas takes 126 registers).  History, because it was measured both ways: while every host input of the
// pipelined end-to-end path was repacked right behind its copy, a cap of 96 registers (minimum block count 9) paid — the free
// quarter of the register file let the NEXT chunk's repack kernels run beside the persistent solve CTAs (end to end 1.65 ->
// 1.89 M problems/s on C3 for 2.5 % of kernel time).  Since the repacks are deferred to the chunk's own solve
// (lexls_b200_wait_copies) the end-to-end number no longer depends on it (2.16 M either way) and the uncapped kernel is 3 %
// faster (cold 11.06 -> 10.68 ms, warm 2.55 -> 2.48 ms, fewer spills).  A ninth resident problem
// (iteration vectors in the global scratch, SmemMap::vecg, LEXLS_B200_VECG=1) needs the 96-register build.
#ifndef LX_LSI_W2_MINB
#define LX_LSI_W2_MINB 8
#endif
template <int W, bool DS, bool LG, bool REG = false>
__global__ void __launch_bounds__(32 * W, (LG && W == 8) ? 2 : ((!LG && !DS && W == 2) ? LX_LSI_W2_MINB : 0)) lsi_kernel(const __grid_constant__ Shape sh, const __grid_constant__ Params pr,
                                                      const __grid_constant__ SmemMap sm, const __grid_constant__ Buffers bf)
{
    constexpr int NT = 32 * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Ctx c = make_ctx(smem_raw, sm);
    if constexpr (LG) c.LOD = bf.lod_scratch + blockIdx.x * bf.lod_stride;
    if (sm.vecg)
    {
        // the four iteration vectors are touched a few times per active-set iteration only: from an L2-resident slice they cost
        // nothing measurable, and the 3.2 KB they free let a ninth C3 problem become resident per SM
        double *const vs = bf.vec_scratch + (size_t)blockIdx.x * 4 * bf.vec_stride;
        c.v   = vs;
        c.dv  = vs + bf.vec_stride;
        c.Ax  = vs + 2 * bf.vec_stride;
        c.Adx = vs + 3 * bf.vec_stride;
    }
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = sh.n, ld = sh.ld, G = sh.G, P = sh.P;
    // the set_* calls found invalid input (lb > ub, duplicate VarIndex): the reference throws in setData; here the solve is
    // refused — every status stays TERMINATION_STATUS_UNKNOWN and the host reports the error at its next synchronisation
    if (bf.flag && *bf.flag)
    {
        for (unsigned long long q = blockIdx.x * (unsigned long long)NT + threadIdx.x; q < bf.batch; q += (unsigned long long)gridDim.x * NT)
            bf.status[q] = ST_UNKNOWN;
        return;
    }

    if (tid == 0)
    {
        mbar_init(c.mbar, 1);
        fence_mbar_init();
    }
    if constexpr (!LG)
        if (tid < 4) c.LOD[ld * (n + 1) + tid] = 0.0; // zero pad behind LOD (see build_map)
    gsync<W>();
    unsigned parity = 0;

    for (;;)
    {
        if (tid == 0) *(unsigned long long *)c.scal = atomicAdd(bf.queue, 1ull);
        gsync<W>();
        const unsigned long long p = *(unsigned long long *)c.scal;
        if (p >= bf.batch) break;

        double *gobj = bf.obj + p * bf.obj_stride;
        double *gsb  = bf.sb ? bf.sb + p * 2ull * sh.m0 : nullptr;
        ObjView<DS> ov{c, sh, DS ? c.obj : gobj, DS ? c.sb : gsb};

        // ---- stage the objective record -----------------------------------------------------------------------
        if constexpr (DS)
        {
            if (tid == 0)
            {
                fence_proxy_async();
                tma_load_record(c.obj, gobj, (unsigned)(bf.obj_stride * sizeof(double)), c.mbar);
            }
#pragma unroll 1
            for (int i = tid; i < 2 * sh.m0; i += NT) c.sb[i] = gsb[i];
        }
#pragma unroll 1
        for (int i = tid; i < sh.m0; i += NT) c.sbidx[i] = (unsigned short)bf.sbidx[p * bf.sbidx_stride + i];
        const bool have_x0 = bf.x0 != nullptr;
        const bool have_v0 = bf.v0 != nullptr;
#pragma unroll 1
        for (int k = tid; k < n; k += NT)
        {
            c.x[k]  = have_x0 ? bf.x0[p * n + k] : 0.0;
            c.dx[k] = 0.0;
            c.lx[k] = 0.0;
        }
#pragma unroll 1
        for (int i = tid; i < sh.Mall; i += NT)
        {
            c.v[i]       = have_v0 ? bf.v0[p * sh.Mall + i] : 0.0;
            c.dv[i]      = 0.0;
            c.alltype[i] = CTR_INACTIVE;
        }
        for (int k = 0; k < P; k++)
        {
            const int o = sh.uoff[k];
#pragma unroll 1
            for (int i = tid; i < sh.dim[k]; i += NT) c.inactive[o + i] = (unsigned short)i;
            if (tid == 0)
            {
                c.nact[k]   = 0;
                c.ninact[k] = sh.dim[k];
            }
        }
        if constexpr (DS)
        {
            mbar_wait(c.mbar, parity);
            parity ^= 1;
        }
        gsync<W>();

        // ---- initial working set: equalities in setData order (lexlsi.h:369-385, :463-476), then the user guess ---
        // The per-row decisions are taken in parallel into byte flags (scratch: the not-yet-used dv / Adx arrays);
        // one thread then replays them in the reference's order, which fixes the order of active[] / inactive[].
        {
            unsigned char *eqflag = (unsigned char *)c.Adx, *gs = (unsigned char *)c.dv;
            for (int k = 0; k < P; k++)
            {
                const int o = sh.uoff[k];
#pragma unroll 1
                for (int i = tid; i < sh.dim[k]; i += NT)
                {
                    const double bl = ov.lb(k, i), bu = ov.ub(k, i);
                    bool eq = fabs(bl - bu) < 1e-15; // isEqual, utility.h:48-51
                    // don't activate 0*x = b rows, lexlsi.h:377: the row norms were tested when the data was set
                    if (eq && !(sh.off && k == 0)) eq = bf.rownz[p * sh.M + sh.goff[k - sh.off] + i] != 0;
                    eqflag[o + i] = eq;
                    gs[o + i]     = bf.guess ? bf.guess[p * sh.Mall + o + i] : 0;
                }
            }
            gsync<W>();
            // The working sets of different levels do not interact: one warp per level replays its level's part of the
            // reference's sequence (all equalities in ascending order, then the guess).  A level of equalities only — the
            // usual case — ends with active = 0, 1, 2, ... and nothing inactive: filled in parallel.  Mixed levels replay.
#pragma unroll 1
            for (int k = tid >> 5; k < P; k += W)
            {
                const int o = sh.uoff[k], m = sh.dim[k];
                bool alleq = true, anyeq = false, anyguess = false;
#pragma unroll 1
                for (int i = lane; i < m; i += 32)
                {
                    const bool e = eqflag[o + i] != 0;
                    alleq        = alleq && e;
                    anyeq        = anyeq || e;
                    const int t  = gs[o + i];
                    anyguess     = anyguess || t == CTR_ACTIVE_LB || t == CTR_ACTIVE_UB;
                }
                alleq    = __all_sync(0xffffffffu, alleq);
                anyeq    = __any_sync(0xffffffffu, anyeq);
                anyguess = (__any_sync(0xffffffffu, anyguess) && bf.guess != nullptr) || bf.glist != nullptr;
                if (alleq && m > 0)
                {
#pragma unroll 1
                    for (int i = lane; i < m; i += 32)
                    {
                        c.active[o + i]  = (unsigned short)i;
                        c.atype[o + i]   = CTR_ACTIVE_EQ;
                        c.alltype[o + i] = CTR_ACTIVE_EQ;
                    }
                    if (lane == 0)
                    {
                        c.nact[k]   = m;
                        c.ninact[k] = 0;
                    }
                }
                else if (lane == 0)
                {
                    if (anyeq)
                        for (int i = 0; i < m; i++)
                            if (eqflag[o + i]) ws_activate_call(sh, sm, k, i, CTR_ACTIVE_EQ);
                    if (bf.glist)
                    {
                        // api_activate in the caller's CALL order (lexlsi.h:120-136 -> workingset.h:67-77): the order fixes
                        // active[] and with it the row order of the LexLSE problem.  Levels do not interact, so this
                        // level replays its own entries of the list.
                        const unsigned *gl = bf.glist + p * bf.gcap;
                        for (unsigned e = 0; e < bf.gcap; e++)
                        {
                            const unsigned w = gl[e];
                            if (w == 0xFFFFFFFFu) break;
                            const int ko = (int)(w >> 24), t = (int)((w >> 16) & 0xffu), i = (int)(w & 0xffffu);
                            if (ko == k && i < m && (t == CTR_ACTIVE_LB || t == CTR_ACTIVE_UB) && c.alltype[o + i] == CTR_INACTIVE)
                                ws_activate_call(sh, sm, k, i, t);
                        }
                    }
                    else if (anyguess)
                        for (int i = 0; i < m; i++)
                        {
                            const int t = gs[o + i];
                            if ((t == CTR_ACTIVE_LB || t == CTR_ACTIVE_UB) && c.alltype[o + i] == CTR_INACTIVE)
                                ws_activate_call(sh, sm, k, i, t); // api_activate, lexlsi.h:120-136 (level-major, row ascending)
                        }
                }
                __syncwarp();
            }
            gsync<W>();
#pragma unroll 1
            for (int i = tid; i < sh.Mall; i += NT) c.dv[i] = 0.0;
            gsync<W>();
        }

        // WS of the reference (lexlsi.h:158, :190-192): every activation in chronological order, across objectives.  Only
        // deactivate_first_wrong_sign reads it.  The initial working set was built level by level above; its order in time is:
        // all equalities (setData order: level-major, rows ascending — the EQ prefix of each active[] list), then the guess
        // (call order of the list, or level-major / row ascending).  Thread 0 owns the list from here on.
        unsigned *const wsl = bf.wsl ? bf.wsl + p * (unsigned long long)sh.Mall : nullptr;
        int wsn             = 0;
        if (wsl && tid == 0)
        {
            for (int k = 0; k < P; k++)
                for (int a = 0; a < c.nact[k] && c.atype[sh.uoff[k] + a] == CTR_ACTIVE_EQ; a++)
                    wsl[wsn++] = ((unsigned)k << 16) | c.active[sh.uoff[k] + a];
            const int neq = wsn;
            if (bf.glist)
            {
                const unsigned *gl = bf.glist + p * bf.gcap;
                for (unsigned e = 0; e < bf.gcap && gl[e] != 0xFFFFFFFFu; e++)
                {
                    const unsigned w  = gl[e];
                    const int ko = (int)(w >> 24), t = (int)((w >> 16) & 0xffu), i = (int)(w & 0xffffu);
                    if (i >= sh.dim[ko] || c.alltype[sh.uoff[ko] + i] != t) continue; // not activated by this call
                    const unsigned key = ((unsigned)ko << 16) | (unsigned)i;
                    bool seen          = false;
                    for (int q = neq; q < wsn; q++) seen = seen || wsl[q] == key;
                    if (!seen) wsl[wsn++] = key;
                }
            }
            else
                for (int k = 0; k < P; k++)
                    for (int a = 0; a < c.nact[k]; a++)
                        if (c.atype[sh.uoff[k] + a] != CTR_ACTIVE_EQ) wsl[wsn++] = ((unsigned)k << 16) | c.active[sh.uoff[k] + a];
        }
        auto track_add = [&](int k, int i) {
            if (wsl) wsl[wsn++] = ((unsigned)k << 16) | (unsigned)i;
        };
        auto track_del = [&](int k, int i) { // order-preserving erase, as std::vector::erase
            if (!wsl) return;
            const unsigned key = ((unsigned)k << 16) | (unsigned)i;
            int q              = 0;
            while (q < wsn && wsl[q] != key) q++;
            for (; q + 1 < wsn; q++) wsl[q] = wsl[q + 1];
            wsn--;
        };

        double *const reg_scratch = REG ? bf.reg_scratch + blockIdx.x * bf.reg_stride : nullptr;
        unsigned nIter = 0, nFact = 0, nAct = 0, nDeact = 0, cycCounter = 0;
        int status = ST_UNKNOWN, totalRank = 0, nFixed = 0, nFixedPrev = 0;
        int cycPrevOp = OP_UNDEFINED, cycPrevObj = 0, cycPrevCtr = 0, cycPrevType = CTR_INACTIVE;
        FlopCount fl;
        fl.f = 0;

        // working-set log, lexlsi.h:1187-1222 (typedefs.h:380-432): one entry per ADD / REMOVE
        unsigned nLog  = 0;
        auto log_entry = [&](int obj, int ctr, int type, double alpha_or_lambda) {
            if (bf.wslog)
            {
                if (tid == 0 && nLog < bf.wslog_cap)
                {
                    double *e = bf.wslog + (p * bf.wslog_cap + nLog) * 6;
                    e[0]      = (double)obj;
                    e[1]      = (double)ctr;
                    e[2]      = (double)type;
                    e[3]      = alpha_or_lambda;
                    e[4]      = (double)totalRank;
                    e[5]      = 0.0;
                }
                nLog++;
            }
        };
        auto factor_and_solve = [&]() {
            lsi_factor_and_solve<W, DS, LG, REG>(sh, sm, pr.tol_lindep, c.LOD, ov.obj, ov.sb, nFixed, nFixedPrev, totalRank, fl, pr.dmma, nullptr,
                                                 REG ? &pr : nullptr, reg_scratch);
        };

        // ---- phase 1, lexlsi.h:816-869 (phase1_v0: :880-915) ---------------------------------------------------
        if (!pr.phase1_v0 && !have_x0)
        {
            factor_and_solve();
#pragma unroll 1
            for (int k = tid; k < n; k += NT) c.x[k] = c.lx[k];
            gsync<W>();
        }
        // Objective::phase1 for every level, objective.h:353-374
        obj_multiply_call<W, DS>(sh, sm, ov.obj, ov.sb, sm.x, c.Ax);
        gsync<W>();
        if (!have_v0)
        {
            if (have_x0 && (pr.mod_active || pr.mod_inactive || pr.mod_x))
            {
                // formInitialWorkingSet, objective.h:115-172 — sequential semantics, level by level
                for (int k = 0; k < P; k++)
                {
                    if (tid == 0 && (pr.mod_active || pr.mod_inactive))
                    {
                        const int o = sh.uoff[k];
                        for (int i = 0; i < sh.dim[k]; i++)
                        {
                            const int t     = c.alltype[o + i];
                            const double ax = c.Ax[o + i];
                            if (t == CTR_INACTIVE && pr.mod_inactive)
                            {
                                if (ax <= ov.lb(k, i))
                                {
                                    ws_activate_call(sh, sm, k, i, CTR_ACTIVE_LB);
                                    track_add(k, i);
                                }
                                else if (ax >= ov.ub(k, i))
                                {
                                    ws_activate_call(sh, sm, k, i, CTR_ACTIVE_UB);
                                    track_add(k, i);
                                }
                            }
                            else if (t == CTR_ACTIVE_LB && pr.mod_active)
                            {
                                if (ax > ov.lb(k, i))
                                {
                                    int a = 0;
                                    while (c.active[o + a] != i) a++;
                                    ws_deactivate_call(sh, sm, k, a);
                                    track_del(k, i);
                                    if (ax >= ov.ub(k, i))
                                    {
                                        ws_activate_call(sh, sm, k, i, CTR_ACTIVE_UB);
                                        track_add(k, i);
                                    }
                                }
                            }
                            else if (t == CTR_ACTIVE_UB && pr.mod_active)
                            {
                                if (ax < ov.ub(k, i))
                                {
                                    int a = 0;
                                    while (c.active[o + a] != i) a++;
                                    ws_deactivate_call(sh, sm, k, a);
                                    track_del(k, i);
                                    if (ax <= ov.lb(k, i))
                                    {
                                        ws_activate_call(sh, sm, k, i, CTR_ACTIVE_LB);
                                        track_add(k, i);
                                    }
                                }
                            }
                        }
                    }
                    gsync<W>();
                    if (k == 0 && sh.off && pr.mod_x)
                    {
                        // ensureZeroCtrViolationForSimpleBounds, objective.h:73-103; then every Ax is rebuilt from the new x
#pragma unroll 1
                        for (int i = tid; i < sh.m0; i += NT)
                        {
                            const int t = c.alltype[i];
                            double val;
                            if (t == CTR_INACTIVE)
                                val = 0.5 * (ov.sb[i] + ov.sb[sh.m0 + i]);
                            else if (t == CTR_ACTIVE_LB)
                                val = ov.sb[i];
                            else
                                val = ov.sb[sh.m0 + i];
                            c.x[c.sbidx[i]] = val;
                        }
                        gsync<W>();
                        obj_multiply_call<W, DS>(sh, sm, ov.obj, ov.sb, sm.x, c.Ax);
                        gsync<W>();
                    }
                }
            }
            // initialize_v0, objective.h:183-237
            for (int k = 0; k < P; k++)
            {
                const int o = sh.uoff[k];
#pragma unroll 1
                for (int i = tid; i < sh.dim[k]; i += NT)
                {
                    const int t     = c.alltype[o + i];
                    const double ax = c.Ax[o + i], bl = ov.lb(k, i), bu = ov.ub(k, i);
                    double val = ax - 0.5 * (bl + bu);
                    if (t == CTR_ACTIVE_LB)
                        val = ax - bl;
                    else if (t == CTR_ACTIVE_UB)
                        val = ax - bu;
                    else if (t == CTR_INACTIVE)
                    {
                        if (pr.min_viol)
                        {
                            if (ax <= bl)
                                val = ax - bl;
                            else if (ax >= bu)
                                val = ax - bu;
                            else
                                val = 0.0;
                        }
                        else if (ax >= (bl - pr.tol_feas) && ax <= (bu + pr.tol_feas))
                            val = 0.0;
                    }
                    c.v[o + i] = val;
                }
            }
            gsync<W>();
        }
        if (!pr.phase1_v0)
        {
            if (have_x0)
            {
                factor_and_solve();
#pragma unroll 1
                for (int k = tid; k < n; k += NT) c.dx[k] = c.lx[k] - c.x[k];
                gsync<W>();
            }
            nFact++;
        }
        lsi_direction<W, DS>(sh, sm, ov.obj, ov.sb, c.v, c.dv, c.Ax, c.Adx);

        // ---- active-set loop, lexlsi.h:223-244; body = verifyWorkingSet :1144-1265 ------------------------------
        for (;;)
        {
            bool normalIteration = true;
            if (nIter != 0)
            {
                factor_and_solve();
#pragma unroll 1
                for (int k = tid; k < n; k += NT) c.dx[k] = c.lx[k] - c.x[k];
                gsync<W>();
                lsi_direction<W, DS>(sh, sm, ov.obj, ov.sb, c.v, c.dv, c.Ax, c.Adx);
                nFact++;
            }
            else if (pr.phase1_v0)
                normalIteration = false;

            // ratio test, lexlsi.h:1006-1029 + objective.h:521-578: first strict minimum in scan order
            ArgMin bl;
            bl.v   = 1.0;
            bl.pos = 0x7fffffff;
            for (int k = 0; k < P; k++)
            {
                const int o = sh.uoff[k], ni = c.ninact[k];
#pragma unroll 1
                for (int q = lane; q < ni; q += 32)
                {
                    const int i      = c.inactive[o + q];
                    const double den = c.Adx[o + i] - c.dv[o + i];
                    double rhs;
                    int t;
                    if (den < -pr.tol_feas)
                    {
                        t   = CTR_ACTIVE_LB;
                        rhs = ov.lb(k, i);
                    }
                    else if (den > pr.tol_feas)
                    {
                        t   = CTR_ACTIVE_UB;
                        rhs = ov.ub(k, i);
                    }
                    else
                        continue;
                    const double num = (rhs - c.Ax[o + i]) + c.v[o + i];
                    double ratio     = num / den;
                    if (ratio < 0) ratio = 0;
                    const int pos = ((o + q) << 1) | (t == CTR_ACTIVE_UB);
                    if (ratio < bl.v)
                    {
                        bl.v   = ratio;
                        bl.pos = pos;
                    }
                }
            }
            bl                 = bfly_argmin(bl);
            const double alpha = bl.v;
            int operation      = OP_UNDEFINED;
            int idObj = 0, idCtr = 0, idType = CTR_INACTIVE;

            if (alpha < 1.0)
            {
                // blocking constraint -> activate, lexlsi.h:1181-1198
                const int gq = bl.pos >> 1;
                int k        = 0;
                while (gq >= sh.uoff[k + 1]) k++;
                idObj  = k;
                idCtr  = c.inactive[gq];
                idType = (bl.pos & 1) ? CTR_ACTIVE_UB : CTR_ACTIVE_LB;
                gsync<W>();
                if (tid == 0)
                {
                    ws_activate_call(sh, sm, idObj, idCtr, idType);
                    track_add(idObj, idCtr);
                }
                log_entry(idObj, idCtr, idType, alpha);
                nAct++;
                operation = OP_ADD;
            }
            else if (normalIteration)
            {
                // findActiveCtr2Remove_largest, lexlsi.h:1115-1139
                gsync<W>();
                if (tid < 32)
                {
                    int found = 0, ctr2rem = 0, obj2rem = 0;
                    double lam2rem = 0.0;
                    int jfound = 0;
                    for (int j = 0; j < G; j++)
                        if (lse_sensitivity(c, n, ld, j, nFixed, pr.tol_wrong, pr.tol_correct, ctr2rem, obj2rem, lane, &lam2rem))
                        {
                            found  = 1;
                            jfound = j;
                            break;
                        }
                    if (found && pr.dfirst && lane == 0)
                    {
                        // findActiveCtr2Remove_first, lexlsi.h:1063-1105: among the constraints of groups jfound, .., 0 whose
                        // multiplier has the wrong sign (the workspace still holds them), the earliest entry of WS.  Hierarchies
                        // with a simple-bounds objective are refused by set_params (the reference reads out of bounds there).
                        lam2rem = 0.0; // lexlsi.h:1069
                        for (int e = 0; e < wsn; e++)
                        {
                            const int k = (int)(wsl[e] >> 16), i = (int)(wsl[e] & 0xffffu), g = k - sh.off;
                            if (g < 0 || g > jfound) continue;
                            int a = 0;
                            while (c.active[sh.uoff[k] + a] != i) a++;
                            const int row = c.frow[g] + a, t = c.ctype[row];
                            if (t == CTR_ACTIVE_EQ || t == CORRECT_SIGN_OF_LAMBDA) continue;
                            double lam = c.wsp[nFixed + row];
                            if (t == CTR_ACTIVE_LB) lam = -lam;
                            if (lam < -pr.tol_wrong)
                            {
                                ctr2rem = a;
                                obj2rem = g;
                                break;
                            }
                        }
                    }
                    if (lane == 0)
                    {
                        c.scal[2] = found;
                        c.scal[3] = ctr2rem;
                        c.scal[4] = obj2rem + sh.off;
                        *(double *)(c.scal + 6) = lam2rem;
                    }
                }
                gsync<W>();
                if (c.scal[2])
                {
                    idObj  = c.scal[4];
                    idCtr  = c.active[sh.uoff[idObj] + c.scal[3]];
                    idType = c.atype[sh.uoff[idObj] + c.scal[3]];
                    log_entry(idObj, idCtr, CTR_INACTIVE, *(const double *)(c.scal + 6));
                    gsync<W>();
                    if (tid == 0)
                    {
                        ws_deactivate_call(sh, sm, idObj, c.scal[3]);
                        track_del(idObj, idCtr);
                    }
                    nDeact++;
                    operation = OP_REMOVE;
                }
                else
                    status = ST_SOLVED;
            }

            // step, lexlsi.h:1243-1250, objective.h:585-589
            if (alpha > 0)
            {
#pragma unroll 1
                for (int k = tid; k < n; k += NT) c.x[k] = fma(alpha, c.dx[k], c.x[k]);
#pragma unroll 1
                for (int i = tid; i < sh.Mall; i += NT)
                {
                    c.v[i]  = fma(alpha, c.dv[i], c.v[i]);
                    c.Ax[i] = fma(alpha, c.Adx[i], c.Ax[i]);
                }
            }

            // cycling handler, cycling.h:32-65
            if (pr.cycling && operation != OP_UNDEFINED)
            {
                int ret    = ST_UNKNOWN;
                bool early = false;
                if (operation == OP_ADD && cycPrevOp == OP_REMOVE && idObj == cycPrevObj && idCtr == cycPrevCtr &&
                    idType == cycPrevType)
                {
                    if (cycCounter >= pr.cyc_max)
                    {
                        ret   = ST_SOLVED_CYCLING;
                        early = true;
                    }
                    else
                    {
                        if (tid == 0)
                        {
                            if (cycPrevType == CTR_ACTIVE_LB)
                                ov.lb(cycPrevObj, cycPrevCtr) -= pr.cyc_step;
                            else
                                ov.ub(cycPrevObj, cycPrevCtr) += pr.cyc_step;
                        }
                        cycCounter++;
                        if (bf.wslog && tid == 0 && nLog >= 1 && nLog <= bf.wslog_cap)
                            bf.wslog[(p * bf.wslog_cap + (nLog - 1)) * 6 + 5] = 1.0; // cycling_detected of this entry
                    }
                }
                if (!early)
                {
                    cycPrevOp   = operation;
                    cycPrevObj  = idObj;
                    cycPrevCtr  = idCtr;
                    cycPrevType = idType;
                }
                status = ret;
            }
            nIter++;
            gsync<W>();

            if (status == ST_SOLVED || status == ST_SOLVED_CYCLING) break;
            if (nFact >= pr.max_fact)
            {
                status = ST_MAX_FACT;
                break;
            }
        }

        // ---- optional outputs: get_lexqr (the last factorization) and getLambda, lexlsi.h:552-605 ------------------
        if (bf.lexqr)
        {
            int nc = 0;
            for (int g = 0; g < G; g++) nc += c.ldim[g];
            dump_lexqr<NT>(c, bf, p, n, ld, nc, sh.M);
            gsync<W>();
        }
        if (bf.xstar)
        {
            // get_xStar, lexlsi.h:519-526: formLexLSE + factorize + solve with the FINAL working set; get_data() and
            // get_lexqr() called after it return this problem's constraints and factorization (tests/test_numerical_error.cpp:83-86)
            lsi_factor_and_solve<W, DS, LG, REG>(sh, sm, pr.tol_lindep, c.LOD, ov.obj, ov.sb, nFixed, nFixedPrev, totalRank, fl, pr.dmma,
                                                 bf.lexdata + p * (unsigned long long)sh.M * (n + 1), REG ? &pr : nullptr, reg_scratch);
#pragma unroll 1
            for (int k = tid; k < n; k += NT) bf.xstar[p * n + k] = c.lx[k];
            int nc = 0;
            for (int g = 0; g < G; g++) nc += c.ldim[g];
            dump_lod<NT>(c, bf.lexqr_star, bf.hh_star, p, n, ld, nc, sh.M);
            gsync<W>();
        }
        if (bf.lambda)
        {
            double *lam = bf.lambda + p * (unsigned long long)P * sh.Mall;
#pragma unroll 1
            for (int e = tid; e < P * sh.Mall; e += NT) lam[e] = 0.0;
            // lexlsi.h:566-571: the working set changed after the last factorization (unless get_xStar just redid it)
            if (status != ST_SOLVED && !bf.xstar) factor_and_solve();
            gsync<W>();
            for (int j = 0; j < G; j++)
            {
                if (tid < 32)
                {
                    int ctr2rem = 0, obj2rem = 0;
                    lse_sensitivity(c, n, ld, j, nFixed, pr.tol_wrong, pr.tol_correct, ctr2rem, obj2rem, lane);
                }
                gsync<W>();
                // workspace = [Lambda of the fixed variables | Lambda of LexLSE rows 0 .. frow[j] + ldim[j])
                double *col = lam + (unsigned long long)(sh.off + j) * sh.Mall;
#pragma unroll 1
                for (int a = tid; a < nFixed; a += NT) col[c.active[a]] = c.wsp[a];
                for (int g = 0; g <= j; g++)
                {
                    const int o = sh.uoff[g + sh.off], fr = c.frow[g];
#pragma unroll 1
                    for (int a = tid; a < c.ldim[g]; a += NT) col[o + c.active[o + a]] = c.wsp[nFixed + fr + a];
                }
                gsync<W>();
            }
        }

        // ---- results ------------------------------------------------------------------------------------------
#pragma unroll 1
        for (int k = tid; k < n; k += NT) bf.x[p * n + k] = c.x[k];
        unsigned nActive = 0;
        for (int k = 0; k < P; k++)
        {
            const int o = sh.uoff[k];
            nActive += c.nact[k];
#pragma unroll 1
            for (int i = tid; i < sh.dim[k]; i += NT)
            {
                const unsigned long long gi = p * sh.Mall + o + i;
                if (bf.v) bf.v[gi] = c.v[o + i];
                if (bf.act) bf.act[gi] = c.alltype[o + i];
                if (bf.order) bf.order[gi] = (i < c.nact[k]) ? (unsigned)c.active[o + i] : 0xFFFFFFFFu;
                if (bf.viol)
                {
                    // getConstraintViolation, objective.h:611-630
                    const double ax = c.Ax[o + i], bl2 = ov.lb(k, i), bu2 = ov.ub(k, i);
                    bf.viol[gi] = (ax <= bl2) ? (ax - bl2) : ((ax >= bu2) ? (ax - bu2) : 0.0);
                }
            }
        }
#pragma unroll 1
        for (int g = tid; g < G; g += NT) bf.ranks[p * G + g] = c.rank[g];
        if constexpr (DS)
        {
            // relaxed bounds (cycling handling) belong to the solver-owned copy of the data, objective.h:774-788
            if (pr.cycling && cycCounter > 0)
            {
#pragma unroll 1
                for (int r = tid; r < sh.M; r += NT)
                {
                    gobj[r + sh.ldo * n]       = c.obj[r + sh.ldo * n];
                    gobj[r + sh.ldo * (n + 1)] = c.obj[r + sh.ldo * (n + 1)];
                }
#pragma unroll 1
                for (int i = tid; i < 2 * sh.m0; i += NT) gsb[i] = c.sb[i];
            }
        }
        if (tid == 0)
        {
            if (bf.wslog_count) bf.wslog_count[p] = nLog < bf.wslog_cap ? nLog : bf.wslog_cap;
            bf.status[p]  = status;
            unsigned *cnt = bf.counters + p * 8;
            cnt[0]        = nFact;
            cnt[1]        = nAct;
            cnt[2]        = nDeact;
            cnt[3]        = nIter;
            cnt[4]        = cycCounter;
            cnt[5]        = (unsigned)totalRank;
            cnt[6]        = nActive;
            cnt[7]        = (unsigned)(((unsigned long long)fl.f + 500ull) / 1000ull);
        }
        gsync<W>();
    }
}

typedef void (*kernel_fn)(const Shape, const Params, const SmemMap, const Buffers);

} // namespace lexls_dev

