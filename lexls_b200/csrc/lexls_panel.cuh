// lexls_panel.cuh — register-window l-QR for the warp-per-problem kernels (included by lexls_device.cuh).
//
// The shared-memory factorization (qr_level_smem / gauss_step_smem in lexls_device.cuh) needs two LDS per DFMA, so it is
// bound by the LSU and by load-use latency (ncu, profiles/r1_lse_c2_smem_w1_full.txt: fp64 pipe 9.7 % busy, 30.8 k warp
// instructions per C2 problem).  Here the block being transformed lives in REGISTERS and the inner loops are pure DFMA.
//
// QR of one level (lexlse.h:182-272), qr_level_win<NG,S>:
//   lane = (par, cl),  par = lane >> 4 (row parity),  cl = lane & 15 (column lane);  S column slots per lane
//   a[q][s] = level row (wb + 2q + par), PHYSICAL column slot (cl, s)          q < 2 NG  (NG groups of 4 rows)
//   * sliding window: the pivot row is always inside the first 4-row group of the window: register Q0 (static, 0 or 1)
//     of the lanes of parity h (dynamic), so "rows above the pivot" never need masks and every register index is
//     static; after four steps the window slides down by one group (register moves).
//     Rows beyond the level's dimension are zero and whole groups beyond it are skipped by uniform branches.
//   * columns never move inside the window: each slot carries its LOGICAL column position pos[s] (what the reference's
//     swaps would have produced, lexlse.h:222-232) and the variable it belongs to.  Rows that are NOT in the window
//     (other levels, finished pivot rows) live in the shared-memory LOD in logical order and are swapped there.
//   * per step the pivot column is published once through shared memory (double buffered); every lane rebuilds the
//     reflector for its rows, the Householder application is straight register FMA code: per parity one sequential
//     chain per column — the (S2) order of the arithmetic contract — closed by one xor-16 shuffle.
//   * the finished reflector column and the finished pivot row (R entries) go to LOD as they are produced.
// Gauss step (lexlse.h:431-471), gauss_rows_dyn: one lower ROW per lane; R and Up are read as warp-uniform (broadcast)
//   128-bit loads, two pivots per load.  Per entry the updates arrive in ascending k.
// Shapes not covered by a compiled instance fall back to the shared-memory code, level by level (same arithmetic).
#pragma once

namespace lexls_dev
{

#define LX_PANEL_MAXRH 20
// the level routines are real calls; -DLX_PANEL_CALL=__forceinline__ inlines them (experiment: no callee-saved register traffic)
#ifndef LX_PANEL_CALL
#define LX_PANEL_CALL __noinline__
#endif
// experiment switch (profiles/r2_experiments.json): one column-slot count for every level (less code, idle slots)
#ifndef LX_PANEL_ONE_S
#define LX_PANEL_ONE_S 0
#endif
// Schur update of the warp-per-problem Gauss step on the FP64 tensor cores: measured SLOWER than the row-per-lane DFMA code
// on C2 (3.40 -> 3.53 ms: 15 small tiles per level do not amortise the fragment loads), so it is compiled in on request only
#ifndef LX_PANEL_DMMA
#define LX_PANEL_DMMA 0
#endif
// staging buffer (doubles): 2 phases x [2 parities x RH rows of the pivot column]
#define LX_PANEL_TMP_PHASE (2 * LX_PANEL_MAXRH)
#define LX_PANEL_TMP_DOUBLES (2 * LX_PANEL_TMP_PHASE)

// The panel routines are real calls (not inlined).  They rebuild their pointers from the dynamic shared-memory base and
// byte offsets, so that the compiler emits LDS/STS instead of generic loads and stores.
struct PanelArgs
{
    unsigned lod, ptmp, invd, hh, colperm, varofcol; // byte offsets into the group's shared memory
    int n, ld, nCtr;
    double tol;
    int dmma;
};
__device__ __forceinline__ PanelArgs make_panel_args(const SmemMap &m, int n, int ld, int nCtr, double tol, int dmma = 0)
{
    PanelArgs a;
    a.lod      = m.lod;
    a.ptmp     = m.ptmp;
    a.invd     = m.invd;
    a.hh       = m.hh;
    a.colperm  = m.colperm;
    a.varofcol = m.varofcol;
    a.n        = n;
    a.ld       = ld;
    a.nCtr     = nCtr;
    a.tol      = tol;
    a.dmma     = dmma;
    return a;
}
#define LX_PANEL_PTRS                                                                                                   \
    extern __shared__ __align__(128) unsigned char lx_panel_smem[];                                                     \
    double *const LOD              = (double *)(lx_panel_smem + pa.lod);                                                \
    double *const ptmp             = (double *)(lx_panel_smem + pa.ptmp);                                               \
    double *const invd             = (double *)(lx_panel_smem + pa.invd);                                               \
    double *const hh               = (double *)(lx_panel_smem + pa.hh);                                                 \
    unsigned short *const colperm  = (unsigned short *)(lx_panel_smem + pa.colperm);                                    \
    unsigned short *const varofcol = (unsigned short *)(lx_panel_smem + pa.varofcol);                                   \
    const int n = pa.n, ld = pa.ld;

// ---------------------------------------------------------------------------------------------------------------
// one pivot step.  The pivot row is window row 2 Q0 + h: register Q0 (static) of the lanes of parity h (dynamic), so the
// code of the steps (wb + 0, wb + 1) and of (wb + 2, wb + 3) is shared.  Returns false when the level is finished.
// ---------------------------------------------------------------------------------------------------------------
template <int NG, int S, int Q0>
__device__ __forceinline__ bool qr_win_step(double (&a)[2 * NG][S], double (&cn)[S], int (&pos)[S], const int (&var)[S],
                                            int &ColIndex, const int wb, const int h, const int ng, const int dim,
                                            const int fr, const PanelArgs &pa)
{
    constexpr int RH = 2 * NG;
    LX_PANEL_PTRS
    const int lane = threadIdx.x & 31, par = lane >> 4, cl = lane & 15;
    const int counter = wb + 2 * Q0 + h;
    if (counter >= dim) return false;
    const int rem    = dim - counter;
    double *const tp = ptmp + h * LX_PANEL_TMP_PHASE; // h == counter & 1

    // pivot = first maximum of the down-dated norms over logical columns ColIndex .. n-1, lexlse.h:205-206
    int p;
    {
        // closed slots (finished columns, the RHS, padding) carry a norm of -inf and never win
        double best = cn[0];
        int bpos    = pos[0];
#pragma unroll
        for (int s = 1; s < S; s++)
        {
            const bool take = cn[s] > best || (cn[s] == best && (unsigned)pos[s] < (unsigned)bpos);
            best            = take ? cn[s] : best;
            bpos            = take ? pos[s] : bpos;
        }
        const unsigned long long key = dkey(best);
        const unsigned hi = (unsigned)(key >> 32);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
        const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
        const bool mine   = (hi == mh) && ((unsigned)key == ml);
        p                 = (int)__reduce_min_sync(0xffffffffu, mine ? (unsigned)bpos : 0xffffffffu);
    }

    // the owner lanes (one per parity) publish the pivot column's window rows (rows past the level are zeros)
#pragma unroll
    for (int s = 0; s < S; s++)
        if (pos[s] == p)
        {
#pragma unroll
            for (int g = 0; g < NG; g++) st2(tp + par * RH + 2 * g, make_double2(a[2 * g][s], a[2 * g + 1][s]));
        }
    __syncwarp();
    double x[RH];
#pragma unroll
    for (int g = 0; g < NG; g++)
    {
        const double2 t = ld2(tp + par * RH + 2 * g);
        x[2 * g]        = t.x;
        x[2 * g + 1]    = t.y;
    }
    const double c0 = tp[h * RH + Q0];

    // ||tail||^2 by (W8): slot of window row j = 2q + par is j % 8 = 2 (q % 4) + par; ||col||^2 = c0^2 + ||tail||^2
    const bool below = par > h; // register Q0 of this lane is below the pivot row
    double acc2;
    {
        const double xq = below ? x[Q0] : 0.0;
        double t[4];
        t[0] = (Q0 == 0) ? xq * xq : 0.0;
        t[1] = (Q0 == 0) ? x[1] * x[1] : xq * xq;
        t[2] = x[2] * x[2];
        t[3] = x[3] * x[3];
#pragma unroll
        for (int g = 2; g < NG; g++) // rows past the level were published as zeros: no guard needed
        {
            t[(2 * g) & 3]     = fma(x[2 * g], x[2 * g], t[(2 * g) & 3]);
            t[(2 * g + 1) & 3] = fma(x[2 * g + 1], x[2 * g + 1], t[(2 * g + 1) & 3]);
        }
        const double P = (t[0] + t[2]) + (t[1] + t[3]);
        acc2           = P + __shfl_xor_sync(0xffffffffu, P, 16);
    }
    const double maxval = fma(c0, c0, acc2);
    if (maxval < pa.tol) return false; // rank found, lexlse.h:214-217 (nothing has been modified)

    // column swap p <-> ColIndex (lexlse.h:222-232): relabel the slots, swap the rows that live in LOD
    {
        int pvar = -1;
#pragma unroll
        for (int s = 0; s < S; s++)
        {
            const bool isp = pos[s] == p, isc = pos[s] == ColIndex;
            pvar           = isp ? var[s] : pvar;
            pos[s]         = isp ? ColIndex : (isc ? p : pos[s]);
            cn[s]          = isp ? -CUDART_INF : cn[s];
        }
        if (pvar >= 0 && par == 0) varofcol[ColIndex] = (unsigned short)pvar;
    }
    if (p != ColIndex)
    {
        const int hp = ld >> 1;
#pragma unroll 1
        for (int t = lane; t < hp; t += 32)
        {
            const double2 u = ld2(LOD + ld * p + 2 * t), w = ld2(LOD + ld * ColIndex + 2 * t);
            st2(LOD + ld * p + 2 * t, w);
            st2(LOD + ld * ColIndex + 2 * t, u);
        }
        __syncwarp();
    }

    // Householder scalars (makeHouseholderInPlace), branch-free: the square root and the two reciprocals overlap
    const bool hhon      = (rem > 1) && !(acc2 <= DBL_MIN); // a non-trivial reflector exists
    const bool zero_tail = (rem > 1) && !hhon;
    const double sq      = sqrt(maxval);
    const double bb      = (c0 >= 0.0) ? -sq : sq;
    const double beta    = hhon ? bb : c0;
    const double inv     = 1.0 / beta;
    const double rdn     = 1.0 / (c0 - bb);
    const double rden    = hhon ? rdn : 0.0;
    const double tau     = hhon ? (bb - c0) * inv : 0.0;

    // the finished column: beta on the diagonal, essential part of the reflector below (lexlse.h:241-242)
    {
        double *const dst = LOD + fr + wb + ld * ColIndex;
#pragma unroll 1
        for (int l = lane; l < dim - wb; l += 32)
            if (l >= 2 * Q0 + h)
            {
                const double xr = tp[(l & 1) * RH + (l >> 1)];
                dst[l]          = (l == 2 * Q0 + h) ? beta : (zero_tail ? 0.0 : xr * rden);
            }
        if (lane == 0)
        {
            colperm[ColIndex] = (unsigned short)p;
            invd[ColIndex]    = inv;
            if (rem > 1) hh[fr + counter] = tau;
        }
    }

    // apply H = I - tau v v' to the other columns (RHS included), lexlse.h:243-246   (S2); x becomes the reflector v
    if (tau != 0.0)
    {
        x[Q0] = below ? x[Q0] * rden : ((par == h) ? 1.0 : 0.0);
        if (Q0 == 0) x[1] *= rden;
        double e[S];
#pragma unroll
        for (int s = 0; s < S; s++)
        {
            e[s] = 0.0;
#pragma unroll
            for (int q = Q0; q < 2; q++) e[s] = fma(x[q], a[q][s], e[s]);
        }
#pragma unroll
        for (int g = 1; g < NG; g++)
            if (g < ng)
            {
                x[2 * g] *= rden;
                x[2 * g + 1] *= rden;
#pragma unroll
                for (int s = 0; s < S; s++)
                {
                    e[s] = fma(x[2 * g], a[2 * g][s], e[s]);
                    e[s] = fma(x[2 * g + 1], a[2 * g + 1][s], e[s]);
                }
            }
#pragma unroll
        for (int s = 0; s < S; s++)
        {
            e[s] = tau * (e[s] + __shfl_xor_sync(0xffffffffu, e[s], 16));
#pragma unroll
            for (int q = Q0; q < 2; q++) a[q][s] = fma(-x[q], e[s], a[q][s]);
        }
#pragma unroll
        for (int g = 1; g < NG; g++)
            if (g < ng)
            {
#pragma unroll
                for (int s = 0; s < S; s++)
                {
                    a[2 * g][s]     = fma(-x[2 * g], e[s], a[2 * g][s]);
                    a[2 * g + 1][s] = fma(-x[2 * g + 1], e[s], a[2 * g + 1][s]);
                }
            }
    }
    // the pivot row is final: R entries of the columns that are still open (and of the RHS)
    {
        double *const dst = LOD + fr + counter;
#pragma unroll
        for (int s = 0; s < S; s++)
            if (par == h && pos[s] > ColIndex) dst[ld * pos[s]] = a[Q0][s];
    }
    ColIndex++;
    if (ColIndex == n) return false; // lexlse.h:255-258

    // down-date the norms with the pivot row, lexlse.h:262-266
#pragma unroll
    for (int s = 0; s < S; s++)
    {
        const double a0 = __shfl_sync(0xffffffffu, a[Q0][s], (h << 4) | cl);
        cn[s]           = fma(-a0, a0, cn[s]);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// QR of one level in registers.  cb = first not yet eliminated column on entry; returns ColIndex (rank = result - cb).
// ---------------------------------------------------------------------------------------------------------------
template <int NG, int S, bool COMPACT>
__device__ LX_PANEL_CALL int qr_level_win(const PanelArgs pa, const int fr, const int dim, const int cb)
{
    constexpr int RH = 2 * NG;
    LX_PANEL_PTRS
    (void)hh; (void)invd; (void)colperm; (void)ptmp;
    const int lane = threadIdx.x & 31, par = lane >> 4, cl = lane & 15;
    int ColIndex   = cb;

    double a[RH][S], cn[S];
    int pos[S], var[S];
#pragma unroll
    for (int s = 0; s < S; s++)
    {
        const int col = cb + cl + 16 * s;
        pos[s]        = (col <= n) ? col : -1;
        var[s]        = (col < n) ? varofcol[col] : 0;
        const double *src = LOD + fr + par + ld * ((col <= n) ? col : n);
#pragma unroll
        for (int q = 0; q < RH; q++) a[q][s] = (2 * q + par < dim && col <= n) ? src[2 * q] : 0.0;
    }
    // initial squared column norms over the level's rows, lexlse.h:193-196   (S2)
#pragma unroll
    for (int s = 0; s < S; s++)
    {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < RH; q++) acc = fma(a[q][s], a[q][s], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        cn[s] = (pos[s] >= 0 && pos[s] < n) ? acc : -CUDART_INF; // closed: the RHS and padding slots
    }

    int wb = 0;
    for (;;)
    {
        const int ng = (dim - wb + 3) >> 2; // groups of 4 rows left in the window (>= 1)
        bool go      = true;
#pragma unroll 1
        for (int h = 0; h < 2 && go; h++) go = qr_win_step<NG, S, 0>(a, cn, pos, var, ColIndex, wb, h, ng, dim, fr, pa);
        if (!go) break;
        if constexpr (COMPACT)
        {
            // one step body only (instruction-cache footprint): slide by one register (two rows) every two steps.
            // (W8) still holds: the accumulator a row feeds rotates with the window, the final sum is symmetric.
#pragma unroll
            for (int s = 0; s < S; s++)
            {
#pragma unroll
                for (int q = 0; q + 1 < RH; q++) a[q][s] = a[q + 1][s];
                a[RH - 1][s] = 0.0;
            }
            wb += 2;
        }
        else
        {
#pragma unroll 1
            for (int h = 0; h < 2 && go; h++) go = qr_win_step<NG, S, 1>(a, cn, pos, var, ColIndex, wb, h, ng, dim, fr, pa);
            if (!go) break;
            // slide the window down by one group
#pragma unroll
            for (int s = 0; s < S; s++)
            {
#pragma unroll
                for (int q = 0; q + 2 < RH; q++) a[q][s] = a[q + 2][s];
                a[RH - 2][s] = 0.0;
                a[RH - 1][s] = 0.0;
            }
            wb += 4;
        }
        if (wb >= dim) break;
    }

    // the open columns (the not yet eliminated rest, and the RHS) go back to LOD at their logical positions
    const int nrem = dim - wb;
#pragma unroll
    for (int s = 0; s < S; s++)
        if (pos[s] >= ColIndex)
        {
            if (nrem > 0)
            {
                double *const dst = LOD + fr + wb + par + ld * pos[s];
#pragma unroll
                for (int q = 0; q < RH; q++)
                    if (2 * q + par < nrem) dst[2 * q] = a[q][s];
            }
            if (par == 0 && pos[s] < n) varofcol[pos[s]] = (unsigned short)var[s];
        }
    __syncwarp();
    return ColIndex;
}

// ---------------------------------------------------------------------------------------------------------------
// Gauss step of level (frg, fc, rank) for ALL lower rows [r0, nCtr): one row per lane, 32 rows per pass.
//   Left(row, 0:rank) := Left * inv(R)   (left-looking over blocks of 4 columns), then
//   Trailing(row, ColIndex..n) -= Left * Up (four columns at a time).
// Left(row, k) is re-read from the lane's own LOD row; R(k, c) and Up(k, c) sit in LOD rows frg + k at warp-uniform
// addresses and are read as aligned row pairs (LDS.128 broadcast, two pivots per load; a first odd pivot is peeled).
// Plain dynamic loops: one small routine for every shape (the instruction cache holds 32 KB).  Per entry the updates
// arrive in ascending k, as in the oracle.
// ---------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void gauss_rows_dyn(const PanelArgs pa, const int r0, const int frg, const int fc, const int rank,
                                            const int ColIndex)
{
    LX_PANEL_PTRS
    (void)hh; (void)colperm; (void)varofcol; (void)ptmp;
    const int lane  = threadIdx.x & 31;
    const int kpeel = frg & 1; // rows frg + k with k >= kpeel, k - kpeel even, start an aligned pair
    for (int rb = r0; rb < pa.nCtr; rb += 32)
    {
        const bool act     = rb + lane < pa.nCtr;
        double *const rowp = LOD + (act ? rb + lane : rb);
        const double *Rp   = LOD + frg;
        const double *lp   = rowp + ld * fc; // Left(row, k) = lp[ld * k]

        for (int cb = 0; cb < rank; cb += 4)
        {
            const double *Rc[4];
            double acc[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int cq = fc + ((cb + q < rank) ? cb + q : rank - 1);
                Rc[q]        = Rp + ld * cq;
                acc[q]       = rowp[ld * cq];
            }
            // finished columns k < cb (cb is a multiple of 4), ascending k
            int k = 0;
            if (kpeel && cb > 0)
            {
                const double l0 = lp[0];
#pragma unroll
                for (int q = 0; q < 4; q++) acc[q] = fma(-l0, Rc[q][0], acc[q]);
                k = 1;
            }
#pragma unroll 2
            for (; k + 1 < cb; k += 2)
            {
                const double l0 = lp[ld * k], l1 = lp[ld * (k + 1)];
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const double2 r = ld2(Rc[q] + k);
                    acc[q]          = fma(-l0, r.x, acc[q]);
                    acc[q]          = fma(-l1, r.y, acc[q]);
                }
            }
            if (k < cb)
            {
                const double l0 = lp[ld * k];
#pragma unroll
                for (int q = 0; q < 4; q++) acc[q] = fma(-l0, Rc[q][k], acc[q]);
            }
            // the 4 x 4 triangle
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
#pragma unroll
                for (int q2 = 0; q2 < q; q2++) acc[q] = fma(-acc[q2], Rc[q][cb + q2], acc[q]);
                acc[q] *= invd[fc + ((cb + q < rank) ? cb + q : rank - 1)];
                if (act && cb + q < rank) rowp[ld * (fc + cb + q)] = acc[q];
            }
            __syncwarp(); // lanes re-read their own stores only, but keep the passes in step
        }

        // Trailing -= Left * Up over columns ColIndex .. n   (pa.dmma: done by schur_dmma in the dispatch routine)
        for (int c = (LX_PANEL_DMMA && pa.dmma) ? n + 1 : ColIndex; c <= n; c += 4)
        {
            const double *Uc[4];
            double t[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int cq = (c + q <= n) ? c + q : n;
                Uc[q]        = Rp + ld * cq;
                t[q]         = rowp[ld * cq];
            }
            int k = 0;
            if (kpeel)
            {
                const double l0 = lp[0];
#pragma unroll
                for (int q = 0; q < 4; q++) t[q] = fma(-l0, Uc[q][0], t[q]);
                k = 1;
            }
#pragma unroll 2
            for (; k + 1 < rank; k += 2)
            {
                const double l0 = lp[ld * k], l1 = lp[ld * (k + 1)];
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const double2 u = ld2(Uc[q] + k);
                    t[q]            = fma(-l0, u.x, t[q]);
                    t[q]            = fma(-l1, u.y, t[q]);
                }
            }
            if (k < rank)
            {
                const double l0 = lp[ld * k];
#pragma unroll
                for (int q = 0; q < 4; q++) t[q] = fma(-l0, Uc[q][k], t[q]);
            }
            if (act)
            {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (c + q <= n) rowp[ld * (c + q)] = t[q];
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// (register variant, used by the LexLSE kernel) Gauss step of level (frg, fc, rank) for the lower rows [r0, r0 + rows), rows <= 32: one row per lane.
//   Left(row, 0:rank) := Left * inv(R)   (left-looking over blocks of 4 columns), kept in registers l[];
//   Trailing(row, ColIndex..n) -= l * Up (four columns at a time).
// R(k, c) and Up(k, c) sit in LOD rows frg + k: warp-uniform addresses, read in aligned row pairs (LDS.128 broadcast);
// ODD = frg is odd, the pairs are then (1,2), (3,4), ...  Reads may touch up to 3 rows past the level (finite LOD data
// or the zero pad behind LOD), always multiplied by l = 0.
// ---------------------------------------------------------------------------------------------------------------
template <int NB, bool ODD>
__device__ LX_PANEL_CALL void gauss_rows_win(const PanelArgs pa, const int r0, const int rows, const int frg, const int fc,
                                            const int rank, const int ColIndex)
{
    constexpr int NK = 4 * NB;
    LX_PANEL_PTRS
    (void)hh; (void)colperm; (void)varofcol; (void)ptmp;
    const int lane     = threadIdx.x & 31;
    const bool act     = lane < rows;
    double *const rowp = LOD + r0 + (act ? lane : 0);
    const double *Rp   = LOD + frg;
    double l[NK + 1];
    l[NK] = 0.0;

#pragma unroll
    for (int b = 0; b < NB; b++)
    {
        if (4 * b >= rank)
        {
#pragma unroll
            for (int q = 0; q < 4; q++) l[4 * b + q] = 0.0;
            continue;
        }
        int col[4];
        double acc[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            col[q] = fc + ((4 * b + q < rank) ? 4 * b + q : rank - 1);
            acc[q] = rowp[ld * col[q]];
        }
        // contributions of the finished columns k < 4b, ascending k
        if constexpr (!ODD)
        {
#pragma unroll
            for (int k = 0; k < 4 * b; k += 2)
            {
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const double2 r = ld2(Rp + k + ld * col[q]);
                    acc[q]          = fma(-l[k], r.x, acc[q]);
                    acc[q]          = fma(-l[k + 1], r.y, acc[q]);
                }
            }
        }
        else if (b > 0)
        {
#pragma unroll
            for (int q = 0; q < 4; q++) acc[q] = fma(-l[0], Rp[ld * col[q]], acc[q]);
#pragma unroll
            for (int k = 1; k + 1 < 4 * b; k += 2)
            {
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const double2 r = ld2(Rp + k + ld * col[q]);
                    acc[q]          = fma(-l[k], r.x, acc[q]);
                    acc[q]          = fma(-l[k + 1], r.y, acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) acc[q] = fma(-l[4 * b - 1], Rp[4 * b - 1 + ld * col[q]], acc[q]);
        }
        // the 4 x 4 triangle
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
#pragma unroll
            for (int q2 = 0; q2 < q; q2++) acc[q] = fma(-acc[q2], Rp[4 * b + q2 + ld * col[q]], acc[q]);
            acc[q] *= invd[col[q]];
            if (4 * b + q < rank)
            {
                l[4 * b + q] = acc[q];
                if (act) rowp[ld * col[q]] = acc[q];
            }
            else
                l[4 * b + q] = 0.0;
        }
    }

    // Trailing -= Left * Up over columns ColIndex .. n   (pa.dmma: done by schur_dmma in the dispatch routine)
    for (int c = (LX_PANEL_DMMA && pa.dmma) ? n + 1 : ColIndex; c <= n; c += 4)
    {
        int cc[4];
        double t[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            cc[q] = (c + q <= n) ? c + q : n;
            t[q]  = rowp[ld * cc[q]];
        }
        if constexpr (!ODD)
        {
#pragma unroll
            for (int b = 0; b < NB; b++)
            {
                if (4 * b >= rank) break;
#pragma unroll
                for (int k = 4 * b; k < 4 * b + 4; k += 2)
                {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                    {
                        const double2 u = ld2(Rp + k + ld * cc[q]);
                        t[q]            = fma(-l[k], u.x, t[q]);
                        t[q]            = fma(-l[k + 1], u.y, t[q]);
                    }
                }
            }
        }
        else
        {
#pragma unroll
            for (int q = 0; q < 4; q++) t[q] = fma(-l[0], Rp[ld * cc[q]], t[q]);
#pragma unroll
            for (int b = 0; b < NB; b++)
            {
                if (4 * b + 1 >= rank) break;
#pragma unroll
                for (int k = 4 * b + 1; k < 4 * b + 5; k += 2)
                {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                    {
                        const double2 u = ld2(Rp + k + ld * cc[q]);
                        t[q]            = fma(-l[k], u.x, t[q]);
                        t[q]            = fma(-l[k + 1], u.y, t[q]);
                    }
                }
            }
        }
        if (act)
        {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (c + q <= n) rowp[ld * cc[q]] = t[q];
        }
    }
    __syncwarp();
}

// dispatch on (rows, columns); false = shape not covered by a compiled instance
template <bool COMPACT>
__device__ __forceinline__ bool qr_level_panel_dispatch(const PanelArgs &pa, const int fr, const int dim, int &ColIndex)
{
    const int ncols = pa.n + 1 - ColIndex;
    const int S     = (ncols + 15) >> 4;
    if (dim < 1) return true; // empty level: rank 0
    if (ColIndex >= pa.n) return true; // every variable is eliminated: rank 0, nothing moves (the level keeps its RHS column)
    if (S > 3 || dim > 40) return false;
#define LX_QR(NGv)                                                                                                      \
    {                                                                                                                   \
        if (LX_PANEL_ONE_S)                                                                                             \
            ColIndex = qr_level_win<NGv, 3, COMPACT>(pa, fr, dim, ColIndex);                                            \
        else if (S == 1)                                                                                                     \
            ColIndex = qr_level_win<NGv, 1, COMPACT>(pa, fr, dim, ColIndex);                                            \
        else if (S == 2)                                                                                                \
            ColIndex = qr_level_win<NGv, 2, COMPACT>(pa, fr, dim, ColIndex);                                            \
        else                                                                                                            \
            ColIndex = qr_level_win<NGv, 3, COMPACT>(pa, fr, dim, ColIndex);                                            \
        return true;                                                                                                    \
    }
    if (dim <= 20) LX_QR(5)
    LX_QR(10)
#undef LX_QR
}

// Gauss step over all lower rows: COMPACT -> the small dynamic-loop routine; otherwise Left in registers, 32 rows at a time
template <bool COMPACT>
__device__ __forceinline__ bool gauss_step_panel_dispatch(const PanelArgs &pa, const int fr, const int dim, const int fc,
                                                          const int rank, const int ColIndex)
{
    auto trailing_dmma = [&]() {
        // Schur update on the FP64 tensor cores: 8 x 8 tiles, same ascending-k chain per entry (lexls_device.cuh: schur_dmma)
        LX_PANEL_PTRS
        (void)hh; (void)colperm; (void)varofcol; (void)ptmp; (void)invd;
        __syncwarp();
        schur_dmma<1>(LOD, ld, LOD + fr, ld, fr + dim, pa.nCtr, fc, rank, ColIndex, n + 1);
        __syncwarp();
    };
    if (COMPACT || rank > 40)
    {
        gauss_rows_dyn(pa, fr + dim, fr, fc, rank, ColIndex);
        if (LX_PANEL_DMMA && pa.dmma) trailing_dmma();
        return true;
    }
    const int frn  = fr + dim;
    const bool odd = fr & 1;
    for (int r0 = frn; r0 < pa.nCtr; r0 += 32)
    {
        const int rows = min(32, pa.nCtr - r0);
#define LX_GS(NBv)                                                                                                      \
    {                                                                                                                   \
        if (odd)                                                                                                        \
            gauss_rows_win<NBv, true>(pa, r0, rows, fr, fc, rank, ColIndex);                                            \
        else                                                                                                            \
            gauss_rows_win<NBv, false>(pa, r0, rows, fr, fc, rank, ColIndex);                                           \
    }
        if (rank <= 12) LX_GS(3)
        else if (rank <= 20) LX_GS(5)
        else LX_GS(10)
#undef LX_GS
    }
    if (LX_PANEL_DMMA && pa.dmma) trailing_dmma();
    return true;
}

} // namespace lexls_dev
