// lexls_pair.cuh — two LexLSE problems per warp (included by lexls_device.cuh after lexls_panel.cuh).
//
// The warp-per-problem register window (lexls_panel.cuh) spends ~240 of its ~300 instructions per pivot step on bookkeeping
// that does not depend on the lane count (arg-max, norm, square root, reciprocals, relabelling, publishing the column).
// Here each HALF warp owns one problem, so one instruction stream does that bookkeeping for two problems:
//
//   lane = (hb, cl): hb = lane >> 4 = which problem, cl = lane & 15 = column lane; S column slots per lane.
//   a[q][s] = level row q (ALL rows of the level, no parity split), physical column slot (cl, s)       q < 4 NG
//
// * no window sliding: rows stay in their registers; the 4-row groups above the pivot's group are skipped by uniform
//   branches (the pivot step index is the same for both problems), the pivot's own group is handled through four
//   temporaries, rows above the pivot inside it get v = 0 (an exact no-op in the FMAs).
// * no cross-lane traffic in the application: the two (S2) chains (even / odd rows) run in the same lane; the pivot row
//   is local, so the norm down-date needs no shuffle; REDUX / shuffles / votes use the half-warp masks.
// * the two problems may leave a level at different steps (rank deficiency): a half that is done keeps executing with
//   tau = 0, v = 0 and all its stores predicated off.
// Arithmetic orders are those of the contract (S2), (W8): results are bit-identical to the other kernels and the oracle.
// Scope: LexLSE without fixed variables, <= 48 columns, <= 20 rows per level (the C2 shape); other shapes use lse_kernel.
#pragma once

namespace lexls_dev
{

struct PairArgs
{
    unsigned lod, ptmp, invd, hh, colperm, varofcol; // byte offsets inside one problem's shared-memory slice
    unsigned slice;                                  // bytes per slice (two slices per warp)
    int n, ld, nCtr;
    double tol;
};

#define LX_PAIR_PTRS                                                                                                    \
    extern __shared__ __align__(128) unsigned char lx_pair_smem[];                                                      \
    const int lane = threadIdx.x & 31, hb = lane >> 4, cl = lane & 15;                                                  \
    const unsigned hmask = 0xFFFFu << (16 * hb);                                                                        \
    unsigned char *const hbase     = lx_pair_smem + hb * pa.slice;                                                      \
    double *const LOD              = (double *)(hbase + pa.lod);                                                        \
    double *const ptmp             = (double *)(hbase + pa.ptmp);                                                       \
    double *const invd             = (double *)(hbase + pa.invd);                                                       \
    double *const hh               = (double *)(hbase + pa.hh);                                                         \
    unsigned short *const colperm  = (unsigned short *)(hbase + pa.colperm);                                            \
    unsigned short *const varofcol = (unsigned short *)(hbase + pa.varofcol);                                           \
    const int n = pa.n, ld = pa.ld;

// ---------------------------------------------------------------------------------------------------------------
// QR of one level for both problems.  cb = first open column of this half's problem; on = this half takes part.
// Returns this half's ColIndex after the level.
// ---------------------------------------------------------------------------------------------------------------
template <int NG, int S>
__device__ __noinline__ int qr_level_pair(const PairArgs pa, const int fr, const int dim, const int cb, const bool on)
{
    constexpr int RH = 4 * NG;
    LX_PAIR_PTRS
    (void)hmask;
    int ColIndex = cb;
    bool act     = on && cb < n;

    double a[RH][S], cn[S];
    int pos[S], var[S];
#pragma unroll
    for (int s = 0; s < S; s++)
    {
        const int col = cb + cl + 16 * s;
        pos[s]        = (col <= n) ? col : -1;
        var[s]        = (col < n) ? varofcol[col] : 0;
        const double *src = LOD + fr + ld * ((col <= n) ? col : n);
#pragma unroll
        for (int q = 0; q < RH; q++) a[q][s] = (q < dim && col <= n) ? src[q] : 0.0;
    }
    // initial squared column norms, lexlse.h:193-196   (S2): one chain per row parity
#pragma unroll
    for (int s = 0; s < S; s++)
    {
        double e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int q = 0; q < RH; q += 2)
        {
            e0 = fma(a[q][s], a[q][s], e0);
            e1 = fma(a[q + 1][s], a[q + 1][s], e1);
        }
        cn[s] = (pos[s] >= 0 && pos[s] < n) ? e0 + e1 : -CUDART_INF;
    }

    for (int counter = 0; counter < dim; counter++)
    {
        if (!__any_sync(0xffffffffu, act)) break;
        const int g0 = counter >> 2, r = counter & 3, rem = dim - counter;
        double *const tp = ptmp + (counter & 1) * LX_PANEL_TMP_PHASE;

        // pivot = first maximum of the down-dated norms over the open columns, lexlse.h:205-206
        int p;
        {
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
            const unsigned mh = __reduce_max_sync(hmask, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(hmask, lo);
            const bool mine   = (hi == mh) && ((unsigned)key == ml);
            p                 = (int)__reduce_min_sync(hmask, mine ? (unsigned)bpos : 0xffffffffu);
        }

        // the owner lane publishes the pivot column's rows from the pivot's group on
#pragma unroll
        for (int s = 0; s < S; s++)
            if (act && pos[s] == p)
            {
#pragma unroll
                for (int g = 0; g < NG; g++)
                    if (g >= g0)
                    {
                        st2(tp + 4 * g, make_double2(a[4 * g][s], a[4 * g + 1][s]));
                        st2(tp + 4 * g + 2, make_double2(a[4 * g + 2][s], a[4 * g + 3][s]));
                    }
            }
        __syncwarp();
        double x[RH];
#pragma unroll
        for (int g = 0; g < NG; g++)
        {
            if (g >= g0)
            {
                const double2 u = ld2(tp + 4 * g), w = ld2(tp + 4 * g + 2);
                x[4 * g]     = u.x;
                x[4 * g + 1] = u.y;
                x[4 * g + 2] = w.x;
                x[4 * g + 3] = w.y;
            }
            else
            {
                x[4 * g] = x[4 * g + 1] = x[4 * g + 2] = x[4 * g + 3] = 0.0;
            }
        }
        // the pivot's group through temporaries: rows above the pivot -> 0, the pivot entry -> c0
        double xg[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int g = 0; g < NG; g++)
            if (g == g0)
            {
#pragma unroll
                for (int i = 0; i < 4; i++) xg[i] = x[4 * g + i];
            }
        const double c0 = (r == 0) ? xg[0] : ((r == 1) ? xg[1] : ((r == 2) ? xg[2] : xg[3]));
#pragma unroll
        for (int i = 0; i < 4; i++) xg[i] = (i > r) ? xg[i] : 0.0; // rows below the pivot only

        // ||tail||^2 by (W8): row q -> slot q % 8 (the two halves of the first butterfly stage are symmetric, so the slot
        // of a row does not depend on which group holds the pivot); rows past the level were loaded as zeros
        double acc2;
        {
            double t[8];
#pragma unroll
            for (int k = 0; k < 8; k++) t[k] = 0.0;
#pragma unroll
            for (int g = 0; g < NG; g++)
            {
                if (g > g0)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++) t[(4 * g + i) & 7] = fma(x[4 * g + i], x[4 * g + i], t[(4 * g + i) & 7]);
                }
                else if (g == g0)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++) t[(4 * g + i) & 7] = fma(xg[i], xg[i], t[(4 * g + i) & 7]);
                }
            }
            acc2 = ((t[0] + t[4]) + (t[2] + t[6])) + ((t[1] + t[5]) + (t[3] + t[7]));
        }
        const double maxval = fma(c0, c0, acc2);
        if (act && maxval < pa.tol) act = false; // rank found, lexlse.h:214-217: this half leaves the level untouched

        // Householder scalars (makeHouseholderInPlace), branch-free; an idle half gets tau = 0, v = 0
        const bool hhon      = act && (rem > 1) && !(acc2 <= DBL_MIN);
        const bool zero_tail = act && (rem > 1) && !hhon;
        const double sq      = sqrt(act ? maxval : 1.0);
        const double bb      = (c0 >= 0.0) ? -sq : sq;
        const double beta    = hhon ? bb : c0;
        const double inv     = 1.0 / (act ? beta : 1.0);
        const double rdn     = 1.0 / (hhon ? (c0 - bb) : 1.0);
        const double rden    = hhon ? rdn : 0.0;
        const double tau     = hhon ? (bb - c0) * inv : 0.0;

        // column swap p <-> ColIndex (lexlse.h:222-232): relabel the slots, swap the rows that live in LOD
        if (act)
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
            if (pvar >= 0) varofcol[ColIndex] = (unsigned short)pvar;
            if (p != ColIndex)
            {
                const int hp = ld >> 1;
#pragma unroll 1
                for (int t = cl; t < hp; t += 16)
                {
                    const double2 u = ld2(LOD + ld * p + 2 * t), w = ld2(LOD + ld * ColIndex + 2 * t);
                    st2(LOD + ld * p + 2 * t, w);
                    st2(LOD + ld * ColIndex + 2 * t, u);
                }
            }
        }
        __syncwarp();
        if (act)
        {
            // the finished column: beta on the diagonal, essential part of the reflector below (lexlse.h:241-242)
            double *const dst = LOD + fr + ld * ColIndex;
#pragma unroll 1
            for (int l = counter + cl; l < dim; l += 16)
            {
                const double xr = tp[l];
                dst[l]          = (l == counter) ? beta : (zero_tail ? 0.0 : xr * rden);
            }
            if (cl == 0)
            {
                colperm[ColIndex] = (unsigned short)p;
                invd[ColIndex]    = inv;
                if (rem > 1) hh[fr + counter] = tau;
            }
        }

        // reflector v = [0 above; 1; essential] and its application (S2): even rows and odd rows chain separately
        if (__any_sync(0xffffffffu, tau != 0.0))
        {
#pragma unroll
            for (int i = 0; i < 4; i++) xg[i] = (i > r) ? xg[i] * rden : ((i == r && hhon) ? 1.0 : 0.0);
#pragma unroll
            for (int g = 0; g < NG; g++)
            {
                if (g == g0)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++) x[4 * g + i] = xg[i];
                }
                else if (g > g0)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++) x[4 * g + i] *= rden;
                }
            }
            double e0[S], e1[S];
#pragma unroll
            for (int s = 0; s < S; s++) e0[s] = e1[s] = 0.0;
#pragma unroll
            for (int g = 0; g < NG; g++)
                if (g >= g0)
                {
#pragma unroll
                    for (int s = 0; s < S; s++)
                    {
                        e0[s] = fma(x[4 * g], a[4 * g][s], e0[s]);
                        e1[s] = fma(x[4 * g + 1], a[4 * g + 1][s], e1[s]);
                        e0[s] = fma(x[4 * g + 2], a[4 * g + 2][s], e0[s]);
                        e1[s] = fma(x[4 * g + 3], a[4 * g + 3][s], e1[s]);
                    }
                }
#pragma unroll
            for (int s = 0; s < S; s++) e0[s] = tau * (e0[s] + e1[s]);
#pragma unroll
            for (int g = 0; g < NG; g++)
                if (g >= g0)
                {
#pragma unroll
                    for (int s = 0; s < S; s++)
                    {
#pragma unroll
                        for (int i = 0; i < 4; i++) a[4 * g + i][s] = fma(-x[4 * g + i], e0[s], a[4 * g + i][s]);
                    }
                }
        }

        // the pivot row is final: R entries of the open columns (and of the RHS) go to LOD, norms are down-dated
        {
            double pr[S];
#pragma unroll
            for (int s = 0; s < S; s++) pr[s] = 0.0;
#pragma unroll
            for (int g = 0; g < NG; g++)
                if (g == g0)
                {
#pragma unroll
                    for (int s = 0; s < S; s++)
                        pr[s] = (r == 0) ? a[4 * g][s] : ((r == 1) ? a[4 * g + 1][s] : ((r == 2) ? a[4 * g + 2][s] : a[4 * g + 3][s]));
                }
            if (act)
            {
                double *const dst = LOD + fr + counter;
#pragma unroll
                for (int s = 0; s < S; s++)
                    if (pos[s] > ColIndex) dst[ld * pos[s]] = pr[s];
                ColIndex++;
                if (ColIndex == n) act = false; // lexlse.h:255-258
#pragma unroll
                for (int s = 0; s < S; s++) cn[s] = fma(-pr[s], pr[s], cn[s]); // lexlse.h:262-266 (closed slots stay -inf)
            }
        }
    }

    // the open columns (the not yet eliminated rest, and the RHS) go back to LOD at their logical positions
    if (on)
    {
#pragma unroll
        for (int s = 0; s < S; s++)
            if (pos[s] >= ColIndex)
            {
                double *const dst = LOD + fr + ld * pos[s];
#pragma unroll
                for (int q = 0; q < RH; q++)
                    if (q < dim) dst[q] = a[q][s];
                if (pos[s] < n) varofcol[pos[s]] = (unsigned short)var[s];
            }
    }
    __syncwarp();
    return ColIndex;
}

// ---------------------------------------------------------------------------------------------------------------
// Gauss step of level (frg, fc, rank) for the lower rows of both problems: one row per lane (16 rows of each problem per
// pass), Left(row, :) in registers.  fc / rank / ColIndex are per half; the loops run to the larger rank with Left = 0
// beyond a half's own rank (exact no-ops).  Same arithmetic as gauss_rows_win.
// ---------------------------------------------------------------------------------------------------------------
template <int NB, bool ODD>
__device__ __noinline__ void gauss_rows_pair(const PairArgs pa, const int frn, const int frg, const int fc, const int rank,
                                             const int ColIndex)
{
    constexpr int NK = 4 * NB;
    LX_PAIR_PTRS
    (void)hh; (void)colperm; (void)varofcol; (void)ptmp;
    const int rmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)rank);
    const int cmin = (int)__reduce_min_sync(0xffffffffu, (unsigned)ColIndex);
    const double *Rp = LOD + frg;
    for (int r0 = frn; r0 < pa.nCtr; r0 += 16)
    {
        const bool act     = r0 + cl < pa.nCtr;
        double *const rowp = LOD + (act ? r0 + cl : r0);
        double l[NK + 1];
        l[NK] = 0.0;
#pragma unroll
        for (int b = 0; b < NB; b++)
        {
            if (4 * b >= rmax)
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
                col[q] = fc + ((4 * b + q < rank) ? 4 * b + q : (rank > 0 ? rank - 1 : 0));
                acc[q] = rowp[ld * col[q]];
            }
            if constexpr (!ODD)
            {
#pragma unroll
                for (int k = 0; k < 4 * b; k += 2)
                {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                    {
                        const double2 rr = ld2(Rp + k + ld * col[q]);
                        acc[q]           = fma(-l[k], rr.x, acc[q]);
                        acc[q]           = fma(-l[k + 1], rr.y, acc[q]);
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
                        const double2 rr = ld2(Rp + k + ld * col[q]);
                        acc[q]           = fma(-l[k], rr.x, acc[q]);
                        acc[q]           = fma(-l[k + 1], rr.y, acc[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) acc[q] = fma(-l[4 * b - 1], Rp[4 * b - 1 + ld * col[q]], acc[q]);
            }
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
        // Trailing -= Left * Up over this half's columns ColIndex .. n (the loop starts at the smaller ColIndex)
        for (int c = cmin; c <= n; c += 4)
        {
            int cc[4];
            double t[4];
            bool mine[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                cc[q]   = (c + q <= n) ? c + q : n;
                mine[q] = (c + q <= n) && (c + q >= ColIndex) && rank > 0;
                t[q]    = rowp[ld * cc[q]];
            }
            if constexpr (!ODD)
            {
#pragma unroll
                for (int b = 0; b < NB; b++)
                {
                    if (4 * b >= rmax) break;
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
                    if (4 * b + 1 >= rmax) break;
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
                    if (mine[q]) rowp[ld * cc[q]] = t[q];
            }
        }
    }
    __syncwarp();
}

} // namespace lexls_dev
