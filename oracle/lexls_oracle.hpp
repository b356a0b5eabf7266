// lexls_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain C++17, Eigen-free restatement of the reference's LexLSE / LexLSI algorithm
// (jrl-umi3218/lexls v1.0.0, header-only; cited below as <file>:<line> relative to the
// reference's include/lexls/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may include, link or execute this file; the product (lexls_b200/) never
// does.
//
// The reference cannot be built in this environment (its only dependency, Eigen >= 3.2, is neither
// installed nor vendored), so this is a restatement; it is pinned against the reference's own
// fixtures in tests/test_oracle_*.py (tests/test_01.dat #Solution, the Escande multiplier KAT,
// the pinv identity, fixed-variables == equality re-posing, LexLSI-final-set == LexLSE).
//
// Arithmetic contract ("canonical order").  Eigen's summation order is packet/alignment dependent
// (the reference says so itself, tests/test_numerical_error.cpp:5-22), so no order is "the"
// reference order.  This oracle fixes one, chosen so that a warp-per-problem GPU kernel can
// reproduce it bit for bit:
//   (S) "sequential": one accumulator, explicit fused multiply-add, ascending index.  Used whenever
//       many independent outputs exist and the sum runs over columns / pivots (the Gauss / Schur
//       update, gemv rows, back-substitution).
//   (S2) "sequential by row parity": a per-column sum over ROWS of the work matrix keeps two
//       accumulators, one for even and one for odd absolute row indices (each sequential, fma,
//       ascending), combined as even + odd.  Used for the initial column norms and for the dot
//       product of a Householder application (the kernel reads rows in aligned pairs).
//   (T) "tree32": a dot product over level rows i+1 .. dim-1 (Householder-sequence application): the
//       term of level row r goes to slot r % 32 (fma, ascending r), the 32 slots are combined by an
//       xor-butterfly 16, 8, 4, 2, 1.  (A warp keeps level row r in lane r % 32.)
//   (W8) "window tree8": the squared norm of the pivot column's tail at pivot step `counter` of a
//       level.  With wb = counter & ~3 (the base of the 4-row group holding the pivot row), the
//       term of in-level row i > counter goes to slot (i - wb) % 8 (fma, ascending i); the 8 slots
//       are combined by an xor-butterfly with offsets 4, 2, 1.  (The register-window kernel keeps
//       row wb + 2q + par in register q of the lanes of parity par: slot = 2 (q % 4) + par.)
// No other floating-point contraction is allowed: compile with -ffp-contract=off.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace lexls_oracle
{
typedef unsigned int Index;

// typedefs.h:49-55
enum TerminationStatus
{
    TERMINATION_STATUS_UNKNOWN            = -1,
    PROBLEM_SOLVED                        = 0,
    PROBLEM_SOLVED_CYCLING_HANDLING       = 1,
    MAX_NUMBER_OF_FACTORIZATIONS_EXCEEDED = 2
};
// typedefs.h:60-64
enum ObjectiveType
{
    GENERAL_OBJECTIVE       = 0,
    SIMPLE_BOUNDS_OBJECTIVE = 1
};
// typedefs.h:69-76
enum ConstraintActivationType
{
    CTR_INACTIVE           = 0,
    CTR_ACTIVE_LB          = 1,
    CTR_ACTIVE_UB          = 2,
    CTR_ACTIVE_EQ          = 3,
    CORRECT_SIGN_OF_LAMBDA = 4
};
// typedefs.h:537-542
enum OperationType
{
    OPERATION_UNDEFINED = 0,
    OPERATION_ADD       = 1,
    OPERATION_REMOVE    = 2
};

// typedefs.h:127-295 (defaults :268-294).  Regularization is out of scope.
struct Parameters
{
    Index max_number_of_factorizations  = 200;
    double tol_linear_dependence        = 1e-12;
    double tol_wrong_sign_lambda        = 1e-08;
    double tol_correct_sign_lambda      = 1e-12;
    double tol_feasibility              = 1e-13;
    bool cycling_handling_enabled       = false;
    Index cycling_max_counter           = 50;
    double cycling_relax_step           = 1e-08;
    bool modify_x_guess_enabled         = false;
    bool modify_type_active_enabled     = false;
    bool modify_type_inactive_enabled   = false;
    bool set_min_init_ctr_violation     = true;
    bool use_phase1_v0                  = false;
    bool log_working_set_enabled        = false;
    bool deactivate_first_wrong_sign    = false;
    int regularization_type             = 0;   // typedefs.h:32-44 (REGULARIZATION_NONE)
    double variable_regularization_factor = 0.0; // typedefs.h:170-176
    Index max_number_of_CG_iterations     = 10;  // typedefs.h:91, :167 (the conjugate-gradient regularization types)
};

// typedefs.h:32-44
enum RegularizationType
{
    REGULARIZATION_NONE        = 0,
    REGULARIZATION_TIKHONOV    = 1,
    REGULARIZATION_TIKHONOV_CG = 2, // not restated (iterative, "not mature")
    REGULARIZATION_R           = 3,
    REGULARIZATION_R_NO_Z      = 4,
    REGULARIZATION_RT_NO_Z     = 5,
    REGULARIZATION_RT_NO_Z_CG  = 6, // not restated
    REGULARIZATION_TIKHONOV_1  = 7, // not restated (the reference's own "WARNING: TESTING" variant)
    REGULARIZATION_TIKHONOV_2  = 8,
    REGULARIZATION_TEST        = 9
};

struct Exception : public std::runtime_error
{
    explicit Exception(const char *m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------------------------------------
// canonical arithmetic
// ---------------------------------------------------------------------------------------------
static inline double fmad(double a, double b, double c)
{
    return __builtin_fma(a, b, c);
}

// (T) tree32 dot product of a[i*sa] and b[i*sb], i = 0..count-1; term i belongs to level row first + i
static inline double tree32_dot(Index count, const double *a, long sa, const double *b, long sb, Index first)
{
    double p[32];
    for (int l = 0; l < 32; l++) p[l] = 0.0;
    for (Index i = 0; i < count; i++) p[(first + i) & 31] = fmad(a[(long)i * sa], b[(long)i * sb], p[(first + i) & 31]);
    for (Index off = 16; off >= 1; off >>= 1)
    {
        double q[32];
        for (Index l = 0; l < 32; l++) q[l] = p[l] + p[l ^ off];
        for (Index l = 0; l < 32; l++) p[l] = q[l];
    }
    return p[0];
}

// (W8) squared norm of col[counter+1 .. dim-1] (col = the level's rows of the pivot column, contiguous)
static inline double tail_sqnorm_w8(const double *col, Index counter, Index dim)
{
    const Index wb = counter & ~3u;
    double p[8];
    for (int l = 0; l < 8; l++) p[l] = 0.0;
    for (Index i = counter + 1; i < dim; i++) p[(i - wb) & 7] = fmad(col[i], col[i], p[(i - wb) & 7]);
    for (Index off = 4; off >= 1; off >>= 1)
    {
        double q[8];
        for (Index l = 0; l < 8; l++) q[l] = p[l] + p[l ^ off];
        for (Index l = 0; l < 8; l++) p[l] = q[l];
    }
    return p[0];
}

// utility.h:48-51
static inline bool isEqual(double a, double b, double tol = 1e-15)
{
    return std::fabs(a - b) < tol;
}

// typedefs.h:621-670
struct ObjectiveInfo
{
    Index dim = 0, rank = 0, first_row_index = 0, first_col_index = 0;
    double regularization_factor = 0.0; // typedefs.h:661
};

// typedefs.h:380-432
struct WorkingSetLogEntry
{
    Index obj_index;
    Index ctr_index;
    int ctr_type;
    double alpha_or_lambda;
    Index rank;
    bool cycling_detected;
};

// =============================================================================================
// LexLSE (lexlse.h)
// =============================================================================================
class LexLSE
{
public:
    LexLSE() {}
    LexLSE(Index nVar_, Index nObj_, const Index *ObjDim_)
    {
        resize(nVar_, nObj_, ObjDim_);
        setObjDim(ObjDim_);
    }

    // lexlse.h:67-103
    void resize(Index nVar_, Index nObj_, const Index *maxObjDim)
    {
        nVar = nVar_;
        nObj = nObj_;
        obj_info.assign(nObj, ObjectiveInfo());
        x.assign(nVar, 0.0); // the reference leaves x uninitialised here; initialize() zeroes the tail
        maxCtr = 0;
        for (Index k = 0; k < nObj; k++) maxCtr += maxObjDim[k];
        ld = maxCtr > 0 ? maxCtr : 1;
        hh_scalars.assign(maxCtr, 0.0);
        ctr_type.assign(maxCtr, CTR_INACTIVE);
        LOD.assign((size_t)ld * (nVar + 1), 0.0);
        PROBLEM_DATA.assign((size_t)ld * (nVar + 1), 0.0);
        Index dim = std::max(maxCtr, nVar);
        dWorkspace.assign(2 * dim + nVar + 1 + 2 * nVar, 0.0);
        column_permutations.assign(nVar, 0);
        ColNorms.assign(nVar, 0.0);
        invdiag.assign(nVar, 0.0);
        null_space.assign((size_t)nVar * (nVar + 1), 0.0); // lexlse.h:93
        array.assign((size_t)nVar * (nVar + 1), 0.0);      // lexlse.h:94
        nVarFixed = nVarFixedInit = 0;
        fixed_var_index.clear();
        fixed_var_type.clear();
        nCtr = 0;
    }

    // lexlse.h:1426-1442 (+ initialize :1672-1693)
    void setObjDim(const Index *ObjDim_)
    {
        nCtr = 0;
        for (Index k = 0; k < nObj; k++)
        {
            nCtr += ObjDim_[k];
            obj_info[k].dim = ObjDim_[k];
            if (k > 0) obj_info[k].first_row_index = obj_info[k - 1].first_row_index + obj_info[k - 1].dim;
        }
        if (nCtr > maxCtr) throw Exception("setObjDim: more constraints than allocated");
        initialize();
    }

    // lexlse.h:1449-1462
    void setFixedVariablesCount(Index nVarFixed_)
    {
        if (nVarFixed_ > nVar) throw Exception("Cannot fix more than nVar variables");
        nVarFixed = nVarFixed_;
        fixed_var_index.resize(nVarFixed);
        fixed_var_type.resize(nVarFixed);
    }

    // lexlse.h:1381-1388
    void fixVariable(Index VarIndex, double VarValue, int type = CTR_ACTIVE_UB)
    {
        fixed_var_index[nVarFixedInit] = VarIndex;
        x[nVarFixedInit]               = VarValue;
        fixed_var_type[nVarFixedInit]  = type;
        nVarFixedInit++;
    }

    void setTolLinearDependence(double t)
    {
        tol_linear_dependence = t;
    }
    // lexlse.h:1467-1470 (the regularization part of ParametersLexLSE)
    void setRegularization(int type, double variable_factor, Index max_cg_iterations = 10)
    {
        regularization_type            = type;
        variable_regularization_factor = variable_factor;
        max_number_of_CG_iterations    = max_cg_iterations;
    }
    // lexlse.h:1478-1482
    void setRegularizationFactor(Index ObjIndex, double factor)
    {
        obj_info[ObjIndex].regularization_factor = factor;
    }

    // lexlse.h:1522-1530; data is column-major dim x (nVar+1) with leading dimension ldd
    void setData(Index ObjIndex, const double *data, Index ldd)
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        Index r0 = obj_info[ObjIndex].first_row_index, m = obj_info[ObjIndex].dim;
        for (Index c = 0; c <= nVar; c++)
            for (Index i = 0; i < m; i++) L(r0 + i, c) = data[(size_t)c * ldd + i];
    }

    // lexlse.h:1539-1543; row has nVar entries spaced `stride` apart
    void setCtr(Index CtrIndex, const double *row, size_t stride, double rhs)
    {
        for (Index c = 0; c < nVar; c++) L(CtrIndex, c) = row[c * stride];
        L(CtrIndex, nVar) = rhs;
    }

    // lexlse.h:1548-1552
    void setCtrType(Index ObjIndex, Index CtrIndex, int type)
    {
        ctr_type[obj_info[ObjIndex].first_row_index + CtrIndex] = type;
    }

    // -----------------------------------------------------------------------------------------
    // lexlse.h:117-506 with REGULARIZATION_NONE
    // -----------------------------------------------------------------------------------------
    void factorize()
    {
        PROBLEM_DATA = LOD; // lexlse.h:119 (kept: get_data() is part of the surface)

        // fixed variables, lexlse.h:132-156
        if (nVarFixed > 0)
        {
            for (Index k = 0; k < nVarFixed; k++)
            {
                Index coeff            = fixed_var_index[k];
                column_permutations[k] = coeff;
                if (k != coeff)
                    for (Index r = 0; r < nCtr; r++) std::swap(L(r, k), L(r, coeff));
                for (Index i = k + 1; i < nVarFixed; i++)
                    if (fixed_var_index[i] == k)
                    {
                        fixed_var_index[i] = coeff;
                        break;
                    }
            }
            // RHS -= LOD[:,0:nFixed] * x[0:nFixed]   (S): per row, ascending fixed index
            for (Index r = 0; r < nCtr; r++)
            {
                double acc = L(r, nVar);
                for (Index f = 0; f < nVarFixed; f++) acc = fmad(-L(r, f), x[f], acc);
                L(r, nVar) = acc;
            }
        }

        Index ColIndex         = nVarFixed;
        Index RemainingColumns = nVar - nVarFixed;
        if (ColIndex >= nVar) // lexlse.h:164-175
        {
            TotalRank = nVarFixed;
            return;
        }

        for (Index ObjIndex = 0; ObjIndex < nObj; ObjIndex++)
        {
            const Index FirstRowIndex = obj_info[ObjIndex].first_row_index;
            const Index FirstColIndex = obj_info[ObjIndex].first_col_index = ColIndex;
            const Index ObjDim                                            = obj_info[ObjIndex].dim;

            // initial squared column norms over this level's rows, lexlse.h:193-196  (S)
            for (Index k = ColIndex; k < nVar; k++)
            {
                double acc[2] = {0.0, 0.0};
                for (Index i = 0; i < ObjDim; i++)
                {
                    const Index r = FirstRowIndex + i;
                    acc[r & 1]    = fmad(L(r, k), L(r, k), acc[r & 1]);
                }
                ColNorms[k] = acc[0] + acc[1];
            }

            for (Index counter = 0; counter < ObjDim; counter++) // lexlse.h:199
            {
                const Index RowIndex      = FirstRowIndex + counter;
                const Index RemainingRows = ObjDim - counter;

                // pivot: first maximum of the (down-dated) norms, lexlse.h:205-206
                Index maxColNormIndex = ColIndex;
                double best = ColNorms[ColIndex], second = -1.0;
                for (Index k = ColIndex + 1; k < nVar; k++)
                {
                    if (ColNorms[k] > best)
                    {
                        second          = best;
                        best            = ColNorms[k];
                        maxColNormIndex = k;
                    }
                    else if (ColNorms[k] > second)
                        second = ColNorms[k];
                }
                if (second >= 0.0 && best > 0.0) min_pivot_gap = std::min(min_pivot_gap, (best - second) / best);

                // exact recomputation of the chosen column's norm, lexlse.h:210-211: ||col||^2 = c0^2 + ||tail||^2 with the
                // tail norm by (W8) — the same two quantities Eigen's makeHouseholder forms (beta^2 = |c0|^2 + tailSqNorm)
                const double *pc        = &L(RowIndex, maxColNormIndex);
                const double tailSqNorm = (RemainingRows > 1) ? tail_sqnorm_w8(pc - counter, counter, ObjDim) : 0.0;
                const double c0         = pc[0];
                double maxColNormValue  = fmad(c0, c0, tailSqNorm);
                ColNorms[maxColNormIndex] = maxColNormValue;

                if (maxColNormValue > 0.0)
                    min_rank_margin = std::min(
                        min_rank_margin, std::fabs(std::log10(maxColNormValue) - std::log10(tol_linear_dependence)));

                if (maxColNormValue < tol_linear_dependence) break; // lexlse.h:214-217

                // column swap over ALL nCtr rows, lexlse.h:222-232
                column_permutations[ColIndex] = maxColNormIndex;
                if (ColIndex != maxColNormIndex)
                {
                    for (Index r = 0; r < nCtr; r++) std::swap(L(r, ColIndex), L(r, maxColNormIndex));
                    std::swap(ColNorms[ColIndex], ColNorms[maxColNormIndex]);
                    if (regularization_type != REGULARIZATION_NONE) // lexlse.h:228-231
                        for (Index p = 0; p < FirstColIndex; p++) std::swap(Zn(p, ColIndex), Zn(p, maxColNormIndex));
                }

                // Householder, lexlse.h:239-248 (Eigen primitives restated from
                // interfaces/matlab-octave/tests/implementation/lexqr/eigen_like_syntax/
                //   makeHouseholderInPlace.m:33-58, applyHouseholderOnTheLeft.m:11-20)
                double beta = c0;
                if (RemainingRows > 1)
                {
                    double *col = &L(RowIndex, ColIndex);
                    double tau;
                    if (tailSqNorm <= DBL_MIN) // Eigen >= 3.3 (the .m file, Eigen 3.2 style, has == 0)
                    {
                        tau = 0.0;
                        for (Index i = 1; i < RemainingRows; i++) col[i] = 0.0;
                    }
                    else
                    {
                        beta = std::sqrt(maxColNormValue);
                        if (c0 >= 0.0) beta = -beta;
                        const double rden = 1.0 / (c0 - beta); // essential = tail * (1/(c0-beta)) (Eigen 3.2 form of tail/(c0-beta))
                        for (Index i = 1; i < RemainingRows; i++) col[i] = col[i] * rden;
                        tau = (beta - c0) * (1.0 / beta);
                    }
                    col[0] = beta;
                    if (tau != 0.0)
                    {
                        // block rows RowIndex.., columns ColIndex+1 .. nVar (RHS included); v = [1; essential]   (S2)
                        for (Index j = ColIndex + 1; j <= nVar; j++)
                        {
                            double *a     = &L(RowIndex, j);
                            double acc[2] = {0.0, 0.0};
                            acc[RowIndex & 1] = fmad(1.0, a[0], acc[RowIndex & 1]);
                            for (Index i = 1; i < RemainingRows; i++)
                                acc[(RowIndex + i) & 1] = fmad(col[i], a[i], acc[(RowIndex + i) & 1]);
                            const double w = tau * (acc[0] + acc[1]);
                            a[0]           = fmad(-1.0, w, a[0]);
                            for (Index i = 1; i < RemainingRows; i++) a[i] = fmad(-col[i], w, a[i]);
                        }
                    }
                    hh_scalars[FirstRowIndex + counter] = tau;
                }
                // reciprocal of the diagonal entry R(c,c): the Gauss step and solve() multiply by it (Eigen's
                // matrix triangular solver does the same; its vector solver divides — last-bit difference only)
                invdiag[ColIndex] = 1.0 / beta;

                ColIndex++;
                RemainingColumns = nVar - ColIndex;
                if (RemainingColumns == 0) break; // lexlse.h:255-258

                // norm down-date, lexlse.h:262-266
                for (Index j = ColIndex; j < nVar; j++) ColNorms[j] = fmad(-L(RowIndex, j), L(RowIndex, j), ColNorms[j]);
            }

            const Index ObjRank = obj_info[ObjIndex].rank = ColIndex - FirstColIndex; // lexlse.h:272

            if (regularization_type != REGULARIZATION_NONE) regularize(ObjIndex, FirstRowIndex, FirstColIndex, ObjRank, RemainingColumns);

            // Gauss step, lexlse.h:431-471
            if (ObjIndex + 1 < nObj && ObjRank > 0)
            {
                const Index FirstRowNext = FirstRowIndex + ObjDim;
                const Index Rows         = nCtr - FirstRowNext;
                // Left := Left * inv(R)  (R = upper-triangular ObjRank x ObjRank), column by column.
                for (Index c = 0; c < ObjRank; c++)
                {
                    const double inv = invdiag[FirstColIndex + c];
                    for (Index r = 0; r < Rows; r++)
                    {
                        double acc = L(FirstRowNext + r, FirstColIndex + c);
                        for (Index k = 0; k < c; k++)
                            acc = fmad(-L(FirstRowNext + r, FirstColIndex + k), L(FirstRowIndex + k, FirstColIndex + c), acc);
                        L(FirstRowNext + r, FirstColIndex + c) = acc * inv;
                    }
                }
                // Trailing -= Left * Up  (columns ColIndex .. nVar, RHS included)
                for (Index j = ColIndex; j <= nVar; j++)
                    for (Index r = 0; r < Rows; r++)
                    {
                        double acc = L(FirstRowNext + r, j);
                        for (Index k = 0; k < ObjRank; k++)
                            acc = fmad(-L(FirstRowNext + r, FirstColIndex + k), L(FirstRowIndex + k, j), acc);
                        L(FirstRowNext + r, j) = acc;
                    }
            }

            if (RemainingColumns == 0) // lexlse.h:475-490
            {
                for (Index k = ObjIndex + 1; k < nObj; k++)
                    obj_info[k].first_col_index = obj_info[k - 1].first_col_index + obj_info[k - 1].rank;
                break;
            }
        }

        TotalRank = nVarFixed;
        for (Index k = 0; k < nObj; k++) TotalRank += obj_info[k].rank; // lexlse.h:494-498
    }

    // -----------------------------------------------------------------------------------------
    // lexlse.h:1015-1045
    // -----------------------------------------------------------------------------------------
    void solve()
    {
        Index AccumulatedRanks = 0;
        for (Index k = nObj; k--;)
        {
            const Index ObjRank = obj_info[k].rank;
            if (ObjRank == 0) continue;
            const Index fr = obj_info[k].first_row_index, fc = obj_info[k].first_col_index;
            double *xk = &x[fc];
            for (Index i = 0; i < ObjRank; i++)
            {
                double acc = L(fr + i, nVar);
                if (AccumulatedRanks > 0)
                {
                    const Index fc1 = obj_info[k + 1].first_col_index;
                    for (Index j = 0; j < AccumulatedRanks; j++) acc = fmad(-L(fr + i, fc1 + j), x[fc1 + j], acc); // (S)
                }
                xk[i] = acc;
            }
            // upper-triangular solve, column oriented
            for (Index j = ObjRank; j--;)
            {
                xk[j] = xk[j] * invdiag[fc + j];
                for (Index i = 0; i < j; i++) xk[i] = fmad(-L(fr + i, fc + j), xk[j], xk[i]);
            }
            AccumulatedRanks += ObjRank;
        }
        // x = P*x with P = T_0 T_1 ... T_{TotalRank-1}  (lexlse.h:501-504, :1044, equivalence :1869-1884)
        for (Index k = TotalRank; k--;) std::swap(x[k], x[column_permutations[k]]);
    }

    // -----------------------------------------------------------------------------------------
    // lexlse.h:1131-1205 solveLeastNorm_2: of all solutions of the factorized hierarchy (the free variables are arbitrary),
    // the one of least norm, by the normal equations (I + S'S) z = S's with [S s] = inv(R) [T b] over the staircase of all
    // levels.  (solveLeastNorm_1, Givens rotations :1051-1124, and _3, the Tikhonov basis :1213-1262, compute the same
    // minimiser with other roundings; this restatement stands for all three.)  Arithmetic: sequential fma chains, ascending
    // index; the triangular solves divide.
    // -----------------------------------------------------------------------------------------
    void solveLeastNorm()
    {
        Index nVarRank = 0;
        for (Index k = 0; k < nObj; k++) nVarRank += obj_info[k].rank;
        const Index nVarFree = nVar - (nVarRank + nVarFixed), W = nVarFree + 1;
        // staircase copy: RT(counter + i, counter + j) = LOD(fr + i, fc + j), j >= i  (upper triangle of R, all of T and b)
        std::vector<double> RT((size_t)nVarRank * (nVarRank + W), 0.0);
        auto A = [&](Index i, Index j) -> double & { return RT[(size_t)j * nVarRank + i]; };
        Index counter = 0;
        for (Index k = 0; k < nObj; k++)
        {
            const Index fr = obj_info[k].first_row_index, fc = obj_info[k].first_col_index, r = obj_info[k].rank;
            for (Index i = 0; i < r; i++)
                for (Index j = i; fc + j <= nVar; j++) A(counter + i, counter + j) = L(fr + i, fc + j);
            counter += r;
        }
        // [S s] = inv(R) [T b], column by column (back substitution)
        for (Index c = 0; c < W; c++)
            for (Index i = nVarRank; i--;)
            {
                double acc = A(i, nVarRank + c);
                for (Index k = i + 1; k < nVarRank; k++) acc = fmad(-A(i, k), A(k, nVarRank + c), acc);
                A(i, nVarRank + c) = acc / A(i, i);
            }
        // D = S'S + I (lower), d = S's
        std::vector<double> Dn((size_t)nVarFree * nVarFree, 0.0), z(nVarFree, 0.0);
        for (Index i = 0; i < nVarFree; i++)
        {
            for (Index j = 0; j <= i; j++)
            {
                double acc = 0.0;
                for (Index k = 0; k < nVarRank; k++) acc = fmad(A(k, nVarRank + i), A(k, nVarRank + j), acc);
                Dn[(size_t)j * nVarFree + i] = (i == j) ? acc + 1.0 : acc;
            }
            double acc = 0.0;
            for (Index k = 0; k < nVarRank; k++) acc = fmad(A(k, nVarRank + i), A(k, nVarRank + nVarFree), acc);
            z[i] = acc;
        }
        auto Dd = [&](Index i, Index j) -> double & { return Dn[(size_t)j * nVarFree + i]; };
        for (Index j = 0; j < nVarFree; j++) // Eigen::LLT
        {
            double acc = Dd(j, j);
            for (Index k = 0; k < j; k++) acc = fmad(-Dd(j, k), Dd(j, k), acc);
            const double ljj = std::sqrt(acc);
            Dd(j, j)         = ljj;
            for (Index i = j + 1; i < nVarFree; i++)
            {
                double a = Dd(i, j);
                for (Index k = 0; k < j; k++) a = fmad(-Dd(i, k), Dd(j, k), a);
                Dd(i, j) = a / ljj;
            }
        }
        for (Index i = 0; i < nVarFree; i++)
        {
            double a = z[i];
            for (Index k = 0; k < i; k++) a = fmad(-Dd(i, k), z[k], a);
            z[i] = a / Dd(i, i);
        }
        for (Index i = nVarFree; i--;)
        {
            double a = z[i];
            for (Index k = i + 1; k < nVarFree; k++) a = fmad(-Dd(k, i), z[k], a);
            z[i] = a / Dd(i, i);
        }
        for (Index c = 0; c < nVarFree; c++) x[nVarFixed + nVarRank + c] = z[c];
        // the determined variables: R y = b - T z over the staircase
        counter = 0;
        for (Index k = 0; k < nObj; k++)
        {
            const Index fr = obj_info[k].first_row_index, r = obj_info[k].rank;
            for (Index i = 0; i < r; i++)
            {
                double acc = L(fr + i, nVar);
                for (Index c = 0; c < nVarFree; c++) acc = fmad(-L(fr + i, nVarRank + nVarFixed + c), z[c], acc);
                x[nVarFixed + counter + i] = acc;
            }
            counter += r;
        }
        for (Index i = nVarRank; i--;)
        {
            double acc = x[nVarFixed + i];
            for (Index k = i + 1; k < nVarRank; k++) acc = fmad(-A(i, k), x[nVarFixed + k], acc);
            x[nVarFixed + i] = acc / A(i, i);
        }
        for (Index k = TotalRank; k--;) std::swap(x[k], x[column_permutations[k]]); // x = P*x
    }

    // Q_k * y for level k; y has obj_info[k].dim entries.  householderSequence(...).applyOnTheLeft,
    // eigen_like_syntax/applyOnTheLeft.m:11-14, householderSequence.m:11-15.
    void applyQ(Index k, double *y) const
    {
        const Index fr = obj_info[k].first_row_index, fc = obj_info[k].first_col_index;
        const Index dim = obj_info[k].dim, rank = obj_info[k].rank;
        for (Index i = rank; i--;)
        {
            const Index len = dim - i;
            const double tau = hh_scalars[fr + i];
            if (len == 1 || tau == 0.0) continue; // 1-row block scales by (1 - tau) with tau == 0
            const double *ess = &Lc(fr + i + 1, fc + i);
            double acc = tree32_dot(len - 1, ess, 1, y + i + 1, 1, i + 1); // (T)
            double tmp = acc + y[i];
            double w   = tau * tmp;
            y[i]       = y[i] - w;
            for (Index r = 1; r < len; r++) y[i + r] = fmad(-ess[r - 1], w, y[i + r]);
        }
    }

    // lexlse.h:1560-1582; returns a pointer to nCtr residuals (v = A x - b restricted to the factorization)
    const double *get_v()
    {
        double *v = &dWorkspace[0];
        for (Index k = 0; k < nObj; k++)
        {
            const Index fr = obj_info[k].first_row_index, dim = obj_info[k].dim, rank = obj_info[k].rank;
            for (Index i = 0; i < rank; i++) v[fr + i] = 0.0;
            for (Index i = rank; i < dim; i++) v[fr + i] = -Lc(fr + i, nVar);
            applyQ(k, v + fr);
        }
        return v;
    }

    // lexlse.h:935-987
    bool findDescentDirection(int FirstRowIndex,
                              Index ObjDim,
                              double &maxAbsValue,
                              Index &CtrIndex,
                              const double *lambda,
                              double tol_wrong_sign_lambda,
                              double tol_correct_sign_lambda)
    {
        bool found = false;
        for (Index k = 0; k < ObjDim; k++)
        {
            Index ind;
            int *aCtrType;
            if (FirstRowIndex < 0)
            {
                ind      = k;
                aCtrType = &fixed_var_type[ind];
            }
            else
            {
                ind      = FirstRowIndex + k;
                aCtrType = &ctr_type[ind];
            }
            if (*aCtrType != CTR_ACTIVE_EQ && *aCtrType != CORRECT_SIGN_OF_LAMBDA)
            {
                double aLambda = lambda[ind];
                if (*aCtrType == CTR_ACTIVE_LB) aLambda = -aLambda;
                if (aLambda > tol_correct_sign_lambda)
                    *aCtrType = CORRECT_SIGN_OF_LAMBDA;
                else if (aLambda < -tol_wrong_sign_lambda)
                {
                    if (aLambda < maxAbsValue)
                    {
                        if (maxAbsValue < 0.0) min_lambda_gap = std::min(min_lambda_gap, maxAbsValue - aLambda);
                        found       = true;
                        maxAbsValue = aLambda;
                        CtrIndex    = k;
                    }
                    else
                        min_lambda_gap = std::min(min_lambda_gap, aLambda - maxAbsValue);
                }
            }
        }
        return found;
    }

    // lexlse.h:611-762.  On exit dWorkspace = [LambdaFixed (nVarFixed) | Lambda (nLambda) | rhs].
    bool ObjectiveSensitivity(Index ObjIndex,
                              Index &CtrIndex2Remove,
                              int &ObjIndex2Remove,
                              double tol_wrong_sign_lambda,
                              double tol_correct_sign_lambda,
                              double &maxAbsValue)
    {
        maxAbsValue = 0.0;
        bool found  = false;
        Index nLambda = 0, nRank = 0;
        for (Index k = 0; k < ObjIndex; k++)
        {
            nLambda += obj_info[k].dim;
            nRank += obj_info[k].rank;
        }
        nLambda += obj_info[ObjIndex].dim;

        const Index total = nVarFixed + nLambda + nRank + nVarFixed;
        for (Index i = 0; i < total; i++) dWorkspace[i] = 0.0;
        double *LambdaFixed = &dWorkspace[0];
        double *Lambda      = &dWorkspace[nVarFixed];
        double *rhs         = &dWorkspace[nVarFixed + nLambda];

        Index fr = obj_info[ObjIndex].first_row_index, fc = obj_info[ObjIndex].first_col_index;
        Index dim = obj_info[ObjIndex].dim, rank = obj_info[ObjIndex].rank;

        for (Index i = rank; i < dim; i++) Lambda[fr + i] = -L(fr + i, nVar); // lexlse.h:658-659
        applyQ(ObjIndex, Lambda + fr);                                         // :662-664

        found = findDescentDirection((int)fr, dim, maxAbsValue, CtrIndex2Remove, Lambda, tol_wrong_sign_lambda,
                                     tol_correct_sign_lambda);
        if (found) ObjIndex2Remove = (int)ObjIndex;

        if (ObjIndex > 0)
        {
            subtractTransposeProduct(rhs, fc, fr, dim, Lambda); // :706-707
            for (Index k = ObjIndex; k--;)
            {
                fr   = obj_info[k].first_row_index;
                fc   = obj_info[k].first_col_index;
                dim  = obj_info[k].dim;
                rank = obj_info[k].rank;
                for (Index i = 0; i < rank; i++) Lambda[fr + i] = rhs[fc + i]; // :719
                applyQ(k, Lambda + fr);                                          // :722-725
                subtractTransposeProduct(rhs, fc, fr, dim, Lambda);              // :727-728
                bool t = findDescentDirection((int)fr, dim, maxAbsValue, CtrIndex2Remove, Lambda,
                                              tol_wrong_sign_lambda, tol_correct_sign_lambda);
                if (t) ObjIndex2Remove = (int)k;
                found = found || t;
            }
        }

        if (nVarFixed > 0) // :742-758
        {
            for (Index f = 0; f < nVarFixed; f++)
            {
                double acc = 0.0;
                for (Index i = 0; i < nLambda; i++) acc = fmad(L(i, f), Lambda[i], acc); // (S)
                LambdaFixed[f] = -acc;
            }
            bool t = findDescentDirection(-1, nVarFixed, maxAbsValue, CtrIndex2Remove, LambdaFixed,
                                          tol_wrong_sign_lambda, tol_correct_sign_lambda);
            if (t) ObjIndex2Remove = -1;
            found = found || t;
        }
        return found;
    }

    // lexlse.h:866-910: every constraint of the group whose multiplier has the wrong sign (same re-typing side effect)
    void findDescentDirectionList(Index ObjIndex, int FirstRowIndex, Index ObjDim, const double *lambda, double tol_wrong_sign_lambda,
                                  double tol_correct_sign_lambda, std::vector<std::pair<int, int>> &ctr_wrong_sign)
    {
        for (Index k = 0; k < ObjDim; k++)
        {
            const Index ind = (Index)FirstRowIndex + k;
            int *aCtrType   = &ctr_type[ind];
            if (*aCtrType != CTR_ACTIVE_EQ && *aCtrType != CORRECT_SIGN_OF_LAMBDA)
            {
                double aLambda = lambda[ind];
                if (*aCtrType == CTR_ACTIVE_LB) aLambda = -aLambda;
                if (aLambda > tol_correct_sign_lambda)
                    *aCtrType = CORRECT_SIGN_OF_LAMBDA;
                else if (aLambda < -tol_wrong_sign_lambda)
                    ctr_wrong_sign.push_back(std::make_pair((int)ObjIndex, (int)k));
            }
        }
    }

    // lexlse.h:511-602 (the variant behind deactivate_first_wrong_sign): all constraints with a wrong-sign multiplier of
    // objective ObjIndex.  Same arithmetic as ObjectiveSensitivity above.  Restated for nVarFixed == 0 ONLY: with fixed
    // variables the reference (lexlse.h:596-599) tests Lambda — the multipliers of the general constraints — instead of
    // LambdaFixed, over obj_info[0].dim entries with the fixed variables' types, and LexLSI then indexes the active list
    // of the simple-bounds objective with those positions (lexlsi.h:1092), out of bounds whenever the position exceeds the
    // number of active bounds: undefined behaviour that cannot be a parity target.
    void ObjectiveSensitivityList(Index ObjIndex, double tol_wrong_sign_lambda, double tol_correct_sign_lambda,
                                  std::vector<std::pair<int, int>> &ctr_wrong_sign)
    {
        if (nVarFixed > 0)
            throw Exception("deactivate_first_wrong_sign with fixed variables / a simple-bounds objective is undefined in the reference (lexlse.h:596-599)");
        Index nLambda = 0, nRank = 0;
        for (Index k = 0; k < ObjIndex; k++)
        {
            nLambda += obj_info[k].dim;
            nRank += obj_info[k].rank;
        }
        nLambda += obj_info[ObjIndex].dim;
        const Index total = nLambda + nRank;
        for (Index i = 0; i < total; i++) dWorkspace[i] = 0.0;
        double *Lambda = &dWorkspace[0];
        double *rhs    = &dWorkspace[nLambda];
        Index fr = obj_info[ObjIndex].first_row_index, fc = obj_info[ObjIndex].first_col_index;
        Index dim = obj_info[ObjIndex].dim, rank = obj_info[ObjIndex].rank;
        for (Index i = rank; i < dim; i++) Lambda[fr + i] = -L(fr + i, nVar);
        applyQ(ObjIndex, Lambda + fr);
        findDescentDirectionList(ObjIndex, (int)fr, dim, Lambda, tol_wrong_sign_lambda, tol_correct_sign_lambda, ctr_wrong_sign);
        if (ObjIndex > 0)
        {
            subtractTransposeProduct(rhs, fc, fr, dim, Lambda);
            for (Index k = ObjIndex; k--;)
            {
                fr   = obj_info[k].first_row_index;
                fc   = obj_info[k].first_col_index;
                dim  = obj_info[k].dim;
                rank = obj_info[k].rank;
                for (Index i = 0; i < rank; i++) Lambda[fr + i] = rhs[fc + i];
                applyQ(k, Lambda + fr);
                subtractTransposeProduct(rhs, fc, fr, dim, Lambda);
                findDescentDirectionList(k, (int)fr, dim, Lambda, tol_wrong_sign_lambda, tol_correct_sign_lambda, ctr_wrong_sign);
            }
        }
    }

    // getters ------------------------------------------------------------------------------
    const std::vector<double> &get_x() const { return x; }
    Index getDim(Index k) const { return obj_info[k].dim; }
    Index getRank(Index k) const { return obj_info[k].rank; }
    Index getTotalRank() const { return TotalRank; }
    Index getFixedVariablesCount() const { return nVarFixed; }
    Index get_nCtr() const { return nCtr; }
    Index get_ld() const { return ld; }
    const std::vector<double> &get_lexqr() const { return LOD; }
    const std::vector<double> &get_data() const { return PROBLEM_DATA; }
    const std::vector<double> &getWorkspace() const { return dWorkspace; }
    const std::vector<Index> &getColumnPermutations() const { return column_permutations; }
    const std::vector<double> &getHouseholderScalars() const { return hh_scalars; }
    const ObjectiveInfo &getObjInfo(Index k) const { return obj_info[k]; }

    // near-tie diagnostics (not part of the reference): smallest relative gap between the two largest
    // candidate pivot norms, smallest |log10(norm/tol)| at a rank decision, smallest gap between
    // competing wrong-sign multipliers.
    double min_pivot_gap   = 1e300;
    double min_rank_margin = 1e300;
    double min_lambda_gap  = 1e300;

private:
    double &L(Index r, Index c) { return LOD[(size_t)c * ld + r]; }
    const double &Lc(Index r, Index c) const { return LOD[(size_t)c * ld + r]; }

    // rhs[0:ColDim] -= LOD(rows fr..fr+dim, 0:ColDim)^T * Lambda[fr..fr+dim]   (S) per column
    void subtractTransposeProduct(double *rhs, Index ColDim, Index fr, Index dim, const double *Lambda)
    {
        for (Index c = 0; c < ColDim; c++)
        {
            double acc = rhs[c];
            for (Index i = 0; i < dim; i++) acc = fmad(-L(fr + i, c), Lambda[fr + i], acc);
            rhs[c] = acc;
        }
    }

    // lexlse.h:1672-1693
    // =========================================================================================
    // Regularization, lexlse.h:274-411 (dispatch), :1700-1760, :2078-2242 (the direct variants), :2560-2625 (null-space
    // basis).  Restated types: TIKHONOV (1), R (3), R_NO_Z (4), RT_NO_Z (5), TIKHONOV_2 (8), TEST (9).
    // Arithmetic: every entry of a product is ONE sequential fma chain in ascending summation index (S); the Cholesky
    // factorization (Eigen::LLT) is the column-oriented textbook form with the same chains; divisions are divisions.
    //   R(i,j) = L(fr+i, fc+j) (only i <= j is read: the part below the diagonal holds Householder vectors)
    //   T(i,c) = L(fr+i, fc+r+c), c < RC;   b(i) = L(fr+i, nVar)
    //   Z(p,c) = null_space(p, c) (leading dimension nVar), s(p) = Z(p, nVar);   D = `array` (leading dimension nVar)
    // =========================================================================================
    double &Zn(Index p, Index c) { return null_space[(size_t)c * nVar + p]; }
    double &Dm(Index i, Index j) { return array[(size_t)j * nVar + i]; }

    // Eigen::LLT<>(D) on the leading N x N block (lower triangle), then solveInPlace(d)
    void llt_solve(Index N, double *d)
    {
        for (Index j = 0; j < N; j++)
        {
            double acc = Dm(j, j);
            for (Index k = 0; k < j; k++) acc = fmad(-Dm(j, k), Dm(j, k), acc);
            const double ljj = std::sqrt(acc);
            Dm(j, j)         = ljj;
            for (Index i = j + 1; i < N; i++)
            {
                double a = Dm(i, j);
                for (Index k = 0; k < j; k++) a = fmad(-Dm(i, k), Dm(j, k), a);
                Dm(i, j) = a / ljj;
            }
        }
        for (Index i = 0; i < N; i++) // L y = d
        {
            double a = d[i];
            for (Index k = 0; k < i; k++) a = fmad(-Dm(i, k), d[k], a);
            d[i] = a / Dm(i, i);
        }
        for (Index i = N; i--;) // L' z = y
        {
            double a = d[i];
            for (Index k = i + 1; k < N; k++) a = fmad(-Dm(k, i), d[k], a);
            d[i] = a / Dm(i, i);
        }
    }

    void regularize(Index ObjIndex, Index fr, Index fc, Index r, Index RC)
    {
        // the factor, lexlse.h:277-311
        double a = obj_info[ObjIndex].regularization_factor;
        if (variable_regularization_factor != 0.0)
        {
            a = 0.0;
            if (r > 0)
            {
                double *t = &dWorkspace[0]; // rhs_tmp
                double n0 = 0.0;
                for (Index i = 0; i < r; i++)
                {
                    t[i] = L(fr + i, nVar);
                    n0   = fmad(t[i], t[i], n0);
                }
                for (Index i = r; i--;) // R y = rhs_tmp (vector solve: divides)
                {
                    double acc = t[i];
                    for (Index k = i + 1; k < r; k++) acc = fmad(-L(fr + i, fc + k), t[k], acc);
                    t[i] = acc / L(fr + i, fc + i);
                }
                double n1 = 0.0;
                for (Index i = 0; i < r; i++) n1 = fmad(t[i], t[i], n1);
                const double ce = n0 / n1, eps = variable_regularization_factor;
                if (ce < eps) a = std::sqrt(1 - (ce * ce) / (eps * eps)) * obj_info[ObjIndex].regularization_factor;
            }
        }
        aRegularizationFactor = a;
        const bool on         = !(std::fabs(a - 0.0) < 1e-15); // !isEqual(a, 0), utility.h:48-51
        const double mu       = a * a;
        const Index q         = fc - nVarFixed; // rows of the accumulated null-space basis
        double *d             = &array[(size_t)nVar * nVar];
        auto R = [&](Index i, Index j) -> double { return L(fr + i, fc + j); };
        auto T = [&](Index i, Index c) -> double { return L(fr + i, fc + r + c); };
        auto b = [&](Index i) -> double & { return L(fr + i, nVar); };

        switch (regularization_type)
        {
        case REGULARIZATION_TIKHONOV: // lexlse.h:317-331
            if (on)
            {
                if (fc + r <= RC)
                    tikhonov_2(fr, fc, r, RC, a, mu, q, d);
                else
                    tikhonov_1(fr, fc, r, RC, mu, q, d);
            }
            accumulate_nullspace_basis(fr, fc, r, RC, q);
            break;
        case REGULARIZATION_TIKHONOV_2: // :389-396
            if (on) tikhonov_2(fr, fc, r, RC, a, mu, q, d);
            accumulate_nullspace_basis(fr, fc, r, RC, q);
            break;
        case REGULARIZATION_TIKHONOV_1: // :378-387: regularize_tikhonov_1_test = tikhonov_1 on the solution path (its X_mu /
            //                             residual_mu by-products are diagnostics that nothing on the path reads)
            if (on) tikhonov_1(fr, fc, r, RC, mu, q, d);
            accumulate_nullspace_basis(fr, fc, r, RC, q);
            break;
        case REGULARIZATION_TIKHONOV_CG: // :335-342, :2256-2279 (start from x = 0), cg_tikhonov :2367-2461
            if (on) cgls(fr, fc, r, RC, a, q, true);
            accumulate_nullspace_basis(fr, fc, r, RC, q);
            break;
        case REGULARIZATION_RT_NO_Z_CG: // :370-376, :2331-2357 (start from x = 0), cg_RT :2472-2550
            if (on) cgls(fr, fc, r, RC, a, 0, false);
            break;
        case REGULARIZATION_R: // :344-351, :2143-2170
            if (on)
            {
                for (Index i = 0; i < r; i++)
                    for (Index j = 0; j <= i; j++)
                    {
                        double acc = 0.0;
                        for (Index k = 0; k <= j; k++) acc = fmad(R(k, i), R(k, j), acc); // R'R
                        double up = 0.0;
                        for (Index p = 0; p < q; p++) up = fmad(Zn(p, fc + i), Zn(p, fc + j), up);
                        acc      = fmad(mu, up, acc);
                        Dm(i, j) = (i == j) ? acc + mu : acc;
                    }
                for (Index i = 0; i < r; i++)
                {
                    double up = 0.0;
                    for (Index p = 0; p < q; p++) up = fmad(Zn(p, fc + i), Zn(p, nVar), up);
                    double acc = up * mu;
                    for (Index k = 0; k <= i; k++) acc = fmad(R(k, i), b(k), acc);
                    d[i] = acc;
                }
                llt_solve(r, d);
                apply_R(fr, fc, r, d);
            }
            accumulate_nullspace_basis(fr, fc, r, RC, q);
            break;
        case REGULARIZATION_R_NO_Z: // :353-360, :2175-2202
            if (on)
            {
                for (Index i = 0; i < r; i++)
                    for (Index j = 0; j <= i; j++)
                    {
                        double acc = 0.0;
                        for (Index k = 0; k <= j; k++) acc = fmad(R(k, i), R(k, j), acc);
                        Dm(i, j) = (i == j) ? acc + mu : acc;
                    }
                for (Index i = 0; i < r; i++)
                {
                    double acc = 0.0;
                    for (Index k = 0; k <= i; k++) acc = fmad(R(k, i), b(k), acc);
                    d[i] = acc;
                }
                llt_solve(r, d);
                apply_R(fr, fc, r, d);
            }
            break;
        case REGULARIZATION_RT_NO_Z: // :362-368, :2207-2242
            if (on)
            {
                for (Index i = 0; i < r; i++)
                    for (Index j = 0; j <= i; j++)
                    {
                        double acc = 0.0;
                        for (Index k = i; k < r; k++) acc = fmad(R(i, k), R(j, k), acc); // R R'
                        for (Index c = 0; c < RC; c++) acc = fmad(T(i, c), T(j, c), acc); // + T T'
                        Dm(i, j) = (i == j) ? acc + mu : acc;
                        if (i != j) Dm(j, i) = acc; // kept for the symmetric product below (the factorization overwrites the lower part)
                        else array_diag(i) = acc;
                    }
                for (Index i = 0; i < r; i++) d[i] = b(i);
                llt_solve(r, d);
                // b = (R R' + T T') d  (the reference subtracts mu from the diagonal again: (x + mu) - mu, replicated)
                for (Index i = 0; i < r; i++)
                {
                    double acc = 0.0;
                    for (Index j = 0; j < r; j++)
                    {
                        const double e = (i == j) ? ((array_diag(i) + mu) - mu) : (i > j ? Dm(j, i) : Dm(i, j));
                        acc            = fmad(e, d[j], acc);
                    }
                    dWorkspace[i] = acc;
                }
                for (Index i = 0; i < r; i++) b(i) = dWorkspace[i];
            }
            break;
        case REGULARIZATION_TEST: // :398-404, :2247-2251
            if (on)
                for (Index i = 0; i < r; i++) b(i) = b(i) * a;
            break;
        default:
            break;
        }
    }

    // CGLS on  min || [triu(R) T; a S; a I] x - [y; a s; 0] ||  from x = 0 (lexlse.h:2367-2461 with the null-space rows S =
    // Z(:, fc .. fc+N), s = Z(:, nVar); :2472-2550 without them), at most max_number_of_CG_iterations steps or until
    // ||s|| <= 1e-12, then y := [triu(R) T] x (:2276-2278).  Fixed orders: every matrix-vector entry is one (S) chain from 0,
    // two partial products are added afterwards (as Eigen's "+=" of a product does); squared norms are (T) tree32 over the index.
    void cgls(Index fr, Index fc, Index r, Index RC, double a, Index qz, bool with_null_space)
    {
        const Index N = r + RC, NR = r + qz + N;
        auto R = [&](Index i, Index j) -> double { return L(fr + i, fc + j); };
        auto T = [&](Index i, Index c) -> double { return L(fr + i, fc + r + c); };
        std::vector<double> rv(NR, 0.0), qv(NR, 0.0), sv(N, 0.0), pv(N, 0.0), xv(N, 0.0);
        (void)with_null_space; // qz = 0 drops the null-space rows
        for (Index i = 0; i < r; i++) rv[i] = L(fr + i, nVar);              // y - [R T] 0
        for (Index p = 0; p < qz; p++) rv[r + p] = Zn(p, nVar) * a;          // a (s - S 0)
        auto form_s = [&]() {
            for (Index c = 0; c < N; c++)
            {
                double t = 0.0;
                for (Index p = 0; p < qz; p++) t = fmad(Zn(p, fc + c), rv[r + p], t); // S' r2
                t = (t + rv[r + qz + c]) * a;                                          // (+ r3) * a
                double u = 0.0;
                if (c < r)
                    for (Index k = 0; k <= c; k++) u = fmad(R(k, c), rv[k], u); // triu(R)' r1
                else
                    for (Index k = 0; k < r; k++) u = fmad(T(k, c - r), rv[k], u); // T' r1
                sv[c] = t + u;
            }
        };
        form_s();
        for (Index c = 0; c < N; c++) pv[c] = sv[c];
        double gamma = tree32_dot(N, sv.data(), 1, sv.data(), 1, 0);
        Index iter   = 0;
        while (std::sqrt(gamma) > 1e-12 && iter < max_number_of_CG_iterations)
        {
            for (Index i = 0; i < r; i++)
            {
                double t = 0.0, u = 0.0;
                for (Index c = 0; c < RC; c++) t = fmad(T(i, c), pv[r + c], t);
                for (Index k = i; k < r; k++) u = fmad(R(i, k), pv[k], u);
                qv[i] = t + u;
            }
            for (Index p = 0; p < qz; p++)
            {
                double t = 0.0;
                for (Index c = 0; c < N; c++) t = fmad(Zn(p, fc + c), pv[c], t);
                qv[r + p] = t * a;
            }
            for (Index c = 0; c < N; c++) qv[r + qz + c] = a * pv[c];
            const double alpha = gamma / tree32_dot(NR, qv.data(), 1, qv.data(), 1, 0);
            for (Index c = 0; c < N; c++) xv[c] = fmad(alpha, pv[c], xv[c]);
            for (Index i = 0; i < NR; i++) rv[i] = fmad(-alpha, qv[i], rv[i]);
            form_s();
            const double gamma_previous = gamma;
            gamma                       = tree32_dot(N, sv.data(), 1, sv.data(), 1, 0);
            const double beta           = gamma / gamma_previous;
            for (Index c = 0; c < N; c++) pv[c] = fmad(beta, pv[c], sv[c]);
            iter++;
        }
        for (Index i = 0; i < r; i++)
        {
            double t = 0.0, u = 0.0;
            for (Index k = i; k < r; k++) t = fmad(R(i, k), xv[k], t);
            for (Index c = 0; c < RC; c++) u = fmad(T(i, c), xv[r + c], u);
            dWorkspace[i] = t + u;
        }
        for (Index i = 0; i < r; i++) L(fr + i, nVar) = dWorkspace[i];
    }

    // scratch for the untouched diagonal of a symmetric D while its lower triangle is being factorized
    double &array_diag(Index i) { return dWorkspace[nVar + 1 + i]; }

    // b(0:r) = triu(R) * d(0:r)
    void apply_R(Index fr, Index fc, Index r, const double *d)
    {
        for (Index i = 0; i < r; i++)
        {
            double acc = 0.0;
            for (Index k = i; k < r; k++) acc = fmad(L(fr + i, fc + k), d[k], acc);
            dWorkspace[i] = acc;
        }
        for (Index i = 0; i < r; i++) L(fr + i, nVar) = dWorkspace[i];
    }

    // lexlse.h:1700-1760: normal equations over the remaining variables, N = RC + r unknowns
    void tikhonov_1(Index fr, Index fc, Index r, Index RC, double mu, Index q, double *d)
    {
        auto R = [&](Index i, Index j) -> double { return L(fr + i, fc + j); };
        auto T = [&](Index i, Index c) -> double { return L(fr + i, fc + r + c); };
        const Index N = RC + r;
        // G(i,j) = column i of [R T] dot column j of [R T] (upper-triangular R), i >= j
        auto col = [&](Index k, Index j) -> double { return j < r ? (k <= j ? R(k, j) : 0.0) : T(k, j - r); };
        for (Index i = 0; i < N; i++)
            for (Index j = 0; j <= i; j++)
            {
                double acc      = 0.0;
                const Index top = (j < r) ? j + 1 : r; // rows where column j is non-zero
                for (Index k = 0; k < top; k++) acc = fmad(col(k, i), col(k, j), acc);
                double up = 0.0;
                for (Index p = 0; p < q; p++) up = fmad(Zn(p, fc + i), Zn(p, fc + j), up);
                acc      = fmad(mu, up, acc);
                Dm(i, j) = (i == j) ? acc + mu : acc;
            }
        for (Index i = 0; i < N; i++)
        {
            double acc      = 0.0;
            const Index top = (i < r) ? i + 1 : r;
            for (Index k = 0; k < top; k++) acc = fmad(col(k, i), L(fr + k, nVar), acc);
            double up = 0.0;
            for (Index p = 0; p < q; p++) up = fmad(Zn(p, fc + i), Zn(p, nVar), up);
            d[i] = fmad(mu, up, acc);
        }
        llt_solve(N, d);
        for (Index i = 0; i < r; i++) // b = triu(R) d(0:r) + T d(r:N)
        {
            double acc = 0.0;
            for (Index k = i; k < r; k++) acc = fmad(R(i, k), d[k], acc);
            for (Index c = 0; c < RC; c++) acc = fmad(T(i, c), d[r + c], acc);
            dWorkspace[i] = acc;
        }
        for (Index i = 0; i < r; i++) L(fr + i, nVar) = dWorkspace[i];
    }

    // lexlse.h:2078-2138: the dual form, N = q + r unknowns ([R T; a Z] [R T; a Z]' + mu I)
    void tikhonov_2(Index fr, Index fc, Index r, Index RC, double a, double mu, Index q, double *d)
    {
        auto R = [&](Index i, Index j) -> double { return L(fr + i, fc + j); };
        auto T = [&](Index i, Index c) -> double { return L(fr + i, fc + r + c); };
        const Index N = q + r, W = RC + r; // W = columns of the stacked matrix
        // row i of the stacked matrix M = [triu(R) T; a * Z(:, fc .. fc + W)]: entries M(i, c), c < W
        auto rowdot = [&](Index i, Index j) -> double { // M(i,:) . M(j,:) without the factors a (applied by the caller)
            double acc = 0.0;
            if (i < r && j < r)
            {
                const Index lo = std::max(i, j);
                for (Index k = lo; k < r; k++) acc = fmad(R(i, k), R(j, k), acc);
                for (Index c = 0; c < RC; c++) acc = fmad(T(i, c), T(j, c), acc);
            }
            else if (i >= r && j >= r)
                for (Index c = 0; c < W; c++) acc = fmad(Zn(i - r, fc + c), Zn(j - r, fc + c), acc);
            else
            {
                const Index zi = (i >= r) ? i - r : j - r, ri = (i >= r) ? j : i;
                for (Index k = ri; k < r; k++) acc = fmad(Zn(zi, fc + k), R(ri, k), acc);
                for (Index c = 0; c < RC; c++) acc = fmad(Zn(zi, fc + r + c), T(ri, c), acc);
            }
            return acc;
        };
        double *diag = &dWorkspace[nVar + 1]; // the diagonal before "+ mu", and the upper copy of the off-diagonal part
        for (Index i = 0; i < N; i++)
            for (Index j = 0; j <= i; j++)
            {
                double v = rowdot(i, j);
                if (i >= r && j >= r)
                    v = mu * v;
                else if (i >= r)
                    v = a * v;
                if (i == j)
                {
                    diag[i]  = v;
                    Dm(i, i) = v + mu;
                }
                else
                {
                    Dm(i, j) = v;
                    Dm(j, i) = v; // the upper copy survives the factorization of the lower triangle
                }
            }
        for (Index i = 0; i < r; i++) d[i] = L(fr + i, nVar);
        for (Index p = 0; p < q; p++) d[r + p] = a * Zn(p, nVar);
        llt_solve(N, d);
        for (Index i = 0; i < r; i++) // b = first r entries of (D - mu I) d, the diagonal as (x + mu) - mu
        {
            double acc = 0.0;
            for (Index j = 0; j < N; j++)
            {
                const double e = (i == j) ? ((diag[i] + mu) - mu) : (i > j ? Dm(j, i) : Dm(i, j));
                acc            = fmad(e, d[j], acc);
            }
            ColNorms[i] = acc; // free between levels
        }
        for (Index i = 0; i < r; i++) L(fr + i, nVar) = ColNorms[i];
    }

    // lexlse.h:2592-2625: Z := Z * [ -inv(R) T ; I ] accumulated in place, the right-hand-side column included
    void accumulate_nullspace_basis(Index fr, Index fc, Index r, Index RC, Index q)
    {
        const Index rows = q + r;
        for (Index i = 0; i < r; i++)
            for (Index j = 0; j < r; j++) Zn(q + i, fc + j) = (i == j) ? 1.0 : 0.0;
        // Left := Left * inv(triu(R)), column by column (the Gauss step's routine: multiply by the reciprocal diagonal)
        for (Index c = 0; c < r; c++)
        {
            const double inv = invdiag[fc + c];
            for (Index p = 0; p < rows; p++)
            {
                double acc = Zn(p, fc + c);
                for (Index k = 0; k < c; k++) acc = fmad(-Zn(p, fc + k), L(fr + k, fc + c), acc);
                Zn(p, fc + c) = acc * inv;
            }
        }
        // Trailing -= Left * Up, columns fc + r .. nVar
        for (Index j = fc + r; j <= nVar; j++)
            for (Index p = 0; p < rows; p++)
            {
                double acc = Zn(p, j);
                for (Index k = 0; k < r; k++) acc = fmad(-Zn(p, fc + k), L(fr + k, j), acc);
                Zn(p, j) = acc;
            }
    }
    double aRegularizationFactor = 0.0;

    void initialize()
    {
        nVarFixedInit = 0;
        TotalRank     = 0;
        for (Index k = 0; k < nObj; k++)
        {
            obj_info[k].rank            = 0;
            obj_info[k].first_col_index = 0;
        }
        std::fill(hh_scalars.begin(), hh_scalars.end(), 0.0);
        std::fill(null_space.begin(), null_space.end(), 0.0); // lexlse.h:1685
        // NB (reference quirk, replicated): nVarFixed is still the PREVIOUS problem's count here, so
        // entries [new nVarFixed, old nVarFixed) keep whatever the last solve left in them.
        for (Index i = nVarFixed; i < nVar; i++) x[i] = 0.0;
    }

    Index nVar = 0, nObj = 0, nCtr = 0, maxCtr = 0, ld = 1;
    Index nVarFixed = 0, nVarFixedInit = 0, TotalRank = 0;
    double tol_linear_dependence = 1e-12;
    std::vector<ObjectiveInfo> obj_info;
    std::vector<double> x, hh_scalars, LOD, PROBLEM_DATA, dWorkspace, ColNorms, invdiag, null_space, array;
    int regularization_type = 0;
    double variable_regularization_factor = 0.0;
    Index max_number_of_CG_iterations = 10;
    std::vector<int> ctr_type, fixed_var_type;
    std::vector<Index> fixed_var_index, column_permutations;
};

// =============================================================================================
// WorkingSet (workingset.h)
// =============================================================================================
class WorkingSet
{
public:
    void resize(Index dim) // :26-33
    {
        all_type.resize(dim);
        reset();
    }
    void reset() // :38-47
    {
        std::fill(all_type.begin(), all_type.end(), (int)CTR_INACTIVE);
        active.clear();
        active_ctr_type.clear();
        inactive.resize(all_type.size());
        for (Index i = 0; i < inactive.size(); i++) inactive[i] = i;
    }
    void activate(Index CtrIndex, int type) // :57-78
    {
        if (all_type[CtrIndex] != CTR_INACTIVE) throw Exception("Cannot activate an active constraint");
        Index ind     = getCtrIndex(CtrIndex);
        inactive[ind] = inactive.back();
        inactive.pop_back();
        all_type[CtrIndex] = type;
        active.push_back(CtrIndex);
        active_ctr_type.push_back(type);
    }
    void deactivate(Index CtrIndexActive) // :91-115
    {
        Index CtrIndex = active[CtrIndexActive];
        if (all_type[CtrIndex] == CTR_INACTIVE) throw Exception("Cannot deactivate an inactive constraint");
        active.erase(active.begin() + CtrIndexActive);
        active_ctr_type.erase(active_ctr_type.begin() + CtrIndexActive);
        all_type[CtrIndex] = CTR_INACTIVE;
        inactive.push_back(CtrIndex);
    }
    Index getCtrIndex(Index k) const // :154-168
    {
        const std::vector<Index> &v = isActive(k) ? active : inactive;
        return (Index)(std::find(v.begin(), v.end(), k) - v.begin());
    }
    bool isActive(Index k) const { return all_type[k] != CTR_INACTIVE; }

    std::vector<Index> active, inactive;
    std::vector<int> active_ctr_type, all_type;
};

// =============================================================================================
// Objective (objective.h)
// =============================================================================================
class Objective
{
public:
    void resize(Index nCtr_, Index nVar_, int obj_type_) // :36-71
    {
        obj_type = obj_type_;
        nCtr     = nCtr_;
        nVar     = nVar_;
        ws.resize(nCtr);
        v.assign(nCtr, 0.0);
        dv.assign(nCtr, 0.0);
        Ax.assign(nCtr, 0.0);
        Adx.assign(nCtr, 0.0);
        if (obj_type == GENERAL_OBJECTIVE)
        {
            data.assign((size_t)nCtr * (nVar + 2), 0.0);
            lb_index = nVar;
            ub_index = nVar + 1;
        }
        else if (obj_type == SIMPLE_BOUNDS_OBJECTIVE)
        {
            data.assign((size_t)nCtr * 2, 0.0);
            var_index.assign(nCtr, 0);
            lb_index = 0;
            ub_index = 1;
        }
        else
            throw Exception("Unknown objective type");
        v0_is_specified = false;
    }

    double &D(Index r, Index c) { return data[(size_t)c * nCtr + r]; }
    const double &D(Index r, Index c) const { return data[(size_t)c * nCtr + r]; }

    // :73-103
    void ensureZeroCtrViolationForSimpleBounds(std::vector<double> &x)
    {
        if (obj_type != SIMPLE_BOUNDS_OBJECTIVE) return;
        for (Index i = 0; i < nCtr; i++)
        {
            Index VarIndex = var_index[i];
            int t          = ws.all_type[i];
            if (t == CTR_INACTIVE)
                x[VarIndex] = 0.5 * (D(i, lb_index) + D(i, ub_index));
            else if (t == CTR_ACTIVE_EQ || t == CTR_ACTIVE_UB)
                x[VarIndex] = D(i, ub_index);
            else if (t == CTR_ACTIVE_LB)
                x[VarIndex] = D(i, lb_index);
        }
    }

    // :115-172
    void formInitialWorkingSet(std::vector<double> &x, bool modify_type_active, bool modify_type_inactive, bool modify_x_guess)
    {
        if (modify_type_active || modify_type_inactive)
        {
            for (Index i = 0; i < nCtr; i++)
            {
                if (!ws.isActive(i) && modify_type_inactive)
                {
                    if (Ax[i] <= D(i, lb_index))
                        ws.activate(i, CTR_ACTIVE_LB);
                    else if (Ax[i] >= D(i, ub_index))
                        ws.activate(i, CTR_ACTIVE_UB);
                }
                else if (ws.all_type[i] == CTR_ACTIVE_LB && modify_type_active)
                {
                    if (Ax[i] > D(i, lb_index))
                    {
                        ws.deactivate(ws.getCtrIndex(i));
                        if (Ax[i] >= D(i, ub_index)) ws.activate(i, CTR_ACTIVE_UB);
                    }
                }
                else if (ws.all_type[i] == CTR_ACTIVE_UB && modify_type_active)
                {
                    if (Ax[i] < D(i, ub_index))
                    {
                        ws.deactivate(ws.getCtrIndex(i));
                        if (Ax[i] <= D(i, lb_index)) ws.activate(i, CTR_ACTIVE_LB);
                    }
                }
            }
        }
        if (obj_type == SIMPLE_BOUNDS_OBJECTIVE && modify_x_guess)
        {
            ensureZeroCtrViolationForSimpleBounds(x);
            initialize_Ax(x);
        }
    }

    // :183-237
    void initialize_v0(double tol_feasibility, bool set_min_init_ctr_violation)
    {
        for (Index i = 0; i < nCtr; i++) v[i] = Ax[i] - 0.5 * (D(i, lb_index) + D(i, ub_index));
        for (Index a = 0; a < ws.active.size(); a++)
        {
            Index i = ws.active[a];
            int t   = ws.active_ctr_type[a];
            if (t == CTR_ACTIVE_LB)
                v[i] = Ax[i] - D(i, lb_index);
            else if (t == CTR_ACTIVE_UB)
                v[i] = Ax[i] - D(i, ub_index);
        }
        for (Index i = 0; i < nCtr; i++)
        {
            if (ws.isActive(i)) continue;
            if (set_min_init_ctr_violation)
            {
                if (Ax[i] <= D(i, lb_index))
                    v[i] = Ax[i] - D(i, lb_index);
                else if (Ax[i] >= D(i, ub_index))
                    v[i] = Ax[i] - D(i, ub_index);
                else
                    v[i] = 0.0;
            }
            else if (Ax[i] >= (D(i, lb_index) - tol_feasibility) && Ax[i] <= (D(i, ub_index) + tol_feasibility))
                v[i] = 0.0;
        }
    }

    // out = A * vec (general) or a gather (simple bounds); :242-273   (S) per row, ascending column
    void multiply(const std::vector<double> &vec, std::vector<double> &out) const
    {
        if (obj_type == GENERAL_OBJECTIVE)
            for (Index i = 0; i < nCtr; i++)
            {
                double acc = 0.0;
                for (Index c = 0; c < nVar; c++) acc = fmad(D(i, c), vec[c], acc);
                out[i] = acc;
            }
        else
            for (Index i = 0; i < nCtr; i++) out[i] = vec[var_index[i]];
    }
    void initialize_Ax(const std::vector<double> &x) { multiply(x, Ax); }

    // :288-338
    void formStep(const std::vector<double> &dx)
    {
        multiply(dx, Adx);
        for (Index i = 0; i < nCtr; i++) dv[i] = -v[i];
        for (Index a = 0; a < ws.active.size(); a++)
        {
            Index i    = ws.active[a];
            int t      = ws.active_ctr_type[a];
            double rhs = (t == CTR_ACTIVE_LB) ? D(i, lb_index) : D(i, ub_index);
            dv[i]      = dv[i] + ((Ax[i] + Adx[i]) - rhs);
        }
    }

    // :353-374
    void phase1(std::vector<double> &x,
                bool x_guess_is_specified,
                bool modify_type_active,
                bool modify_type_inactive,
                bool modify_x_guess,
                bool set_min_init_ctr_violation,
                double tol_feasibility)
    {
        initialize_Ax(x);
        if (!v0_is_specified)
        {
            if (x_guess_is_specified) formInitialWorkingSet(x, modify_type_active, modify_type_inactive, modify_x_guess);
            initialize_v0(tol_feasibility, set_min_init_ctr_violation);
        }
    }

    // :434-494
    void formLexLSE(LexLSE &lexlse, Index &counter, Index ObjIndex)
    {
        if (obj_type == SIMPLE_BOUNDS_OBJECTIVE)
        {
            lexlse.setFixedVariablesCount((Index)ws.active.size());
            for (Index a = 0; a < ws.active.size(); a++)
            {
                Index i = ws.active[a];
                int t   = ws.active_ctr_type[a];
                if (t == CTR_ACTIVE_LB)
                    lexlse.fixVariable(var_index[i], D(i, 0), CTR_ACTIVE_LB);
                else if (t == CTR_ACTIVE_UB)
                    lexlse.fixVariable(var_index[i], D(i, 1), CTR_ACTIVE_UB);
                else if (t == CTR_ACTIVE_EQ)
                    lexlse.fixVariable(var_index[i], D(i, 1), CTR_ACTIVE_EQ);
            }
        }
        else
        {
            for (Index a = 0; a < ws.active.size(); a++)
            {
                Index i    = ws.active[a];
                int t      = ws.active_ctr_type[a];
                double rhs = (t == CTR_ACTIVE_LB) ? D(i, nVar) : D(i, nVar + 1);
                lexlse.setCtrType(ObjIndex, a, t);
                lexlse.setCtr(counter, &data[i], nCtr, rhs);
                counter++;
            }
            lexlse.setRegularizationFactor(ObjIndex, regularization_factor); // objective.h:492
        }
    }
    double regularization_factor = 0.0; // objective.h:828, :973

    // :521-578
    bool checkBlockingConstraints(Index &CtrIndexBlocking, int &CtrTypeBlocking, double &alpha, double tol_feasibility,
                                  double &min_alpha_gap) const
    {
        const double alpha_input = alpha;
        for (Index q = 0; q < ws.inactive.size(); q++)
        {
            Index i    = ws.inactive[q];
            double den = Adx[i] - dv[i];
            int t;
            double rhs;
            if (den < -tol_feasibility)
            {
                t   = CTR_ACTIVE_LB;
                rhs = D(i, lb_index);
            }
            else if (den > tol_feasibility)
            {
                t   = CTR_ACTIVE_UB;
                rhs = D(i, ub_index);
            }
            else
                continue;
            double num   = (rhs - Ax[i]) + v[i];
            double ratio = num / den;
            if (ratio < 0) ratio = 0;
            if (ratio < alpha)
            {
                if (alpha < 1.0) min_alpha_gap = std::min(min_alpha_gap, alpha - ratio);
                alpha            = ratio;
                CtrIndexBlocking = i;
                CtrTypeBlocking  = t;
            }
            else if (ratio < 1.0)
                min_alpha_gap = std::min(min_alpha_gap, ratio - alpha);
        }
        return alpha < alpha_input;
    }

    // :585-589
    void step(double alpha)
    {
        for (Index i = 0; i < nCtr; i++)
        {
            v[i]  = fmad(alpha, dv[i], v[i]);
            Ax[i] = fmad(alpha, Adx[i], Ax[i]);
        }
    }

    // :611-630
    void getConstraintViolation(std::vector<double> &out) const
    {
        out.resize(nCtr);
        for (Index i = 0; i < nCtr; i++)
        {
            if (Ax[i] <= D(i, lb_index))
                out[i] = Ax[i] - D(i, lb_index);
            else if (Ax[i] >= D(i, ub_index))
                out[i] = Ax[i] - D(i, ub_index);
            else
                out[i] = 0.0;
        }
    }

    // :774-788
    void relax_bounds(Index CtrIndex, int CtrType, double p)
    {
        if (CtrType == CTR_ACTIVE_LB)
            D(CtrIndex, lb_index) -= p;
        else if (CtrType == CTR_ACTIVE_UB)
            D(CtrIndex, ub_index) += p;
        else
            throw Exception("Should not be here");
    }

    // :845-857
    bool isZeroNormal(Index CtrIndex) const
    {
        if (obj_type != GENERAL_OBJECTIVE) return false;
        double s = 0.0;
        for (Index c = 0; c < nVar; c++) s = fmad(D(CtrIndex, c), D(CtrIndex, c), s);
        return s == 0;
    }

    Index nVar = 0, nCtr = 0, lb_index = 0, ub_index = 0;
    int obj_type = GENERAL_OBJECTIVE;
    std::vector<Index> var_index;
    std::vector<double> data; // column-major nCtr x (nVar+2)  or  nCtr x 2
    WorkingSet ws;
    std::vector<double> v, dv, Ax, Adx;
    bool v0_is_specified = false;
};

// =============================================================================================
// LexLSI (lexlsi.h) + CyclingHandler (cycling.h)
// =============================================================================================
class LexLSI
{
public:
    LexLSI() {}
    LexLSI(Index nVar_, Index nObj_, const Index *ObjDim_, const int *ObjType_) { resize(nVar_, nObj_, ObjDim_, ObjType_); }

    // lexlsi.h:77-83, :923-946
    void resize(Index nVar_, Index nObj_, const Index *ObjDim_, const int *ObjType_)
    {
        nVar       = nVar_;
        nObj       = nObj_;
        nObjOffset = (ObjType_[0] == SIMPLE_BOUNDS_OBJECTIVE) ? 1 : 0;
        lexlse.resize(nVar, nObj - nObjOffset, ObjDim_ + nObjOffset);
        lexlse.setTolLinearDependence(parameters.tol_linear_dependence);
        nActive.assign(nObj, 0);
        objectives.assign(nObj, Objective());
        for (Index k = 0; k < nObj; k++) objectives[k].resize(ObjDim_[k], nVar, ObjType_[k]);
        x.assign(nVar, 0.0);
        dx.assign(nVar, 0.0);
        nIterations = nActivations = nDeactivations = nFactorizations = lexlse_rank = 0;
        x_guess_is_specified = false;
        status               = TERMINATION_STATUS_UNKNOWN;
        cyc_counter          = 0;
        cyc_prev_op          = OPERATION_UNDEFINED;
        cyc_prev_obj = cyc_prev_ctr = 0;
        cyc_prev_type               = CTR_INACTIVE;
        working_set_log.clear();
        WS.clear();
    }

    // lexlsi.h:325-342
    void setParameters(const Parameters &p)
    {
        parameters = p;
        lexlse.setTolLinearDependence(p.tol_linear_dependence);
        lexlse.setRegularization(p.regularization_type, p.variable_regularization_factor, p.max_number_of_CG_iterations); // lexlsi.h:331-333
    }
    // lexlsi.h:497-502 (a simple-bounds objective is never regularized, objective.h:489-493)
    void setRegularizationFactor(Index ObjIndex, double factor)
    {
        objectives[ObjIndex].regularization_factor = factor;
    }

    // lexlsi.h:350-435; data column-major dim x (nVar+2) with leading dimension ldd
    void setData(Index ObjIndex, const double *data, Index ldd)
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        Objective &o = objectives[ObjIndex];
        if (o.obj_type != GENERAL_OBJECTIVE) throw Exception("ObjType = GENERAL_OBJECTIVE is assumed");
        for (Index i = 0; i < o.nCtr; i++)
        {
            double bl = data[(size_t)nVar * ldd + i], bu = data[(size_t)(nVar + 1) * ldd + i];
            if (isEqual(bl, bu))
            {
                double s = 0.0;
                for (Index c = 0; c < nVar; c++) s = fmad(data[(size_t)c * ldd + i], data[(size_t)c * ldd + i], s);
                if (s > 0) activate(ObjIndex, i, CTR_ACTIVE_EQ, false);
            }
            else if (bl > bu)
                throw Exception("(general) Lower bound is greater than upper bound.");
        }
        for (Index c = 0; c < nVar + 2; c++)
            for (Index i = 0; i < o.nCtr; i++) o.D(i, c) = data[(size_t)c * ldd + i];
    }

    // lexlsi.h:444-491; data column-major dim x 2
    void setData(Index ObjIndex, const Index *VarIndex, const double *data, Index ldd)
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        Objective &o = objectives[ObjIndex];
        if (o.obj_type != SIMPLE_BOUNDS_OBJECTIVE) throw Exception("ObjType = SIMPLE_BOUNDS_OBJECTIVE is assumed");
        for (Index i = 0; i < o.nCtr; i++)
        {
            double bl = data[i], bu = data[(size_t)ldd + i];
            if (isEqual(bl, bu))
                activate(ObjIndex, i, CTR_ACTIVE_EQ, false);
            else if (bl > bu)
                throw Exception("(simple) Lower bound is greater than upper bound.");
        }
        for (Index k = 0; k < o.nCtr; k++)
            for (Index j = 0; j < o.nCtr; j++)
                if (VarIndex[k] == VarIndex[j] && j != k) throw Exception("Elements of VarIndex are not unique.");
        for (Index i = 0; i < o.nCtr; i++)
        {
            o.var_index[i] = VarIndex[i];
            o.D(i, 0)      = data[i];
            o.D(i, 1)      = data[(size_t)ldd + i];
        }
    }

    // lexlsi.h:120-136
    void api_activate(Index ObjIndex, Index CtrIndex, int type)
    {
        if (!objectives[ObjIndex].ws.isActive(CtrIndex))
            if (type == CTR_ACTIVE_LB || type == CTR_ACTIVE_UB) activate(ObjIndex, CtrIndex, type, false);
    }

    // lexlsi.h:306-320
    void set_x0(const double *x0)
    {
        for (Index i = 0; i < nVar; i++) x[i] = x0[i];
        x_guess_is_specified = true;
    }
    void set_v0(Index ObjIndex, const double *v0)
    {
        Objective &o = objectives[ObjIndex];
        for (Index i = 0; i < o.nCtr; i++) o.v[i] = v0[i];
        o.v0_is_specified = true;
    }

    // lexlsi.h:205-246
    TerminationStatus solve()
    {
        if (parameters.use_phase1_v0)
            phase1_v0();
        else
            phase1();
        while (1)
        {
            verifyWorkingSet();
            if (status == PROBLEM_SOLVED || status == PROBLEM_SOLVED_CYCLING_HANDLING) break;
            if (nFactorizations >= parameters.max_number_of_factorizations)
            {
                status = MAX_NUMBER_OF_FACTORIZATIONS_EXCEEDED;
                break;
            }
        }
        return status;
    }

    // getters -------------------------------------------------------------------------------
    const std::vector<double> &get_x() const { return x; }
    const std::vector<double> &get_v(Index k) const { return objectives[k].v; }
    void getConstraintViolation(Index k, std::vector<double> &out) const { objectives[k].getConstraintViolation(out); }
    void getActiveCtr(Index k, std::vector<int> &ctr_type) const // lexlsi.h:688-698
    {
        const Objective &o = objectives[k];
        ctr_type.assign(o.nCtr, CTR_INACTIVE);
        for (Index a = 0; a < o.ws.active.size(); a++) ctr_type[o.ws.active[a]] = o.ws.active_ctr_type[a];
    }
    const WorkingSet &getWorkingSet(Index k) const { return objectives[k].ws; }
    Index getFactorizationsCount() const { return nFactorizations; }
    Index getActivationsCount() const { return nActivations; }
    Index getDeactivationsCount() const { return nDeactivations; }
    Index getIterationsCount() const { return nIterations; }
    Index getCyclingCounter() const { return cyc_counter; }
    Index getActiveCtrCount(Index k) const { return (Index)objectives[k].ws.active.size(); }
    Index getTotalRank() const { return lexlse.getTotalRank(); }
    TerminationStatus getStatus() const { return status; }
    LexLSE &getLexLSE() { return lexlse; }
    Objective &getObjective(Index k) { return objectives[k]; }
    const std::vector<WorkingSetLogEntry> &getWorkingSetLog() const { return working_set_log; }

    // lexlsi.h:519-526
    const std::vector<double> &get_xStar()
    {
        formLexLSE();
        lexlse.factorize();
        lexlse.solve();
        return lexlse.get_x();
    }

    // lexlsi.h:552-605: vec_lambda[k] is dim_k x nObj, column-major
    void getLambda(std::vector<std::vector<double>> &vec_lambda)
    {
        Index nActiveCtr = 0;
        vec_lambda.resize(nObj);
        for (Index k = 0; k < nObj; k++)
        {
            nActiveCtr += getActiveCtrCount(k);
            vec_lambda[k].assign((size_t)objectives[k].nCtr * nObj, 0.0);
        }
        if (status != PROBLEM_SOLVED)
        {
            formLexLSE();
            lexlse.factorize();
        }
        std::vector<double> L_active((size_t)nActiveCtr * nObj, 0.0);
        Index nMeaningful = lexlse.getFixedVariablesCount();
        Index ctr;
        int obj;
        double maxAbs;
        for (Index k = 0; k < nObj - nObjOffset; k++)
        {
            lexlse.ObjectiveSensitivity(k, ctr, obj, parameters.tol_wrong_sign_lambda, parameters.tol_correct_sign_lambda, maxAbs);
            nMeaningful += lexlse.getDim(k);
            for (Index i = 0; i < nMeaningful; i++) L_active[(size_t)(nObjOffset + k) * nActiveCtr + i] = lexlse.getWorkspace()[i];
        }
        Index acc = 0;
        for (Index k = 0; k < nObj; k++)
        {
            for (Index a = 0; a < getActiveCtrCount(k); a++)
            {
                Index ind = objectives[k].ws.active[a];
                for (Index j = 0; j < nObj; j++)
                    vec_lambda[k][(size_t)j * objectives[k].nCtr + ind] = L_active[(size_t)j * nActiveCtr + acc + a];
            }
            acc += getActiveCtrCount(k);
        }
    }

    double min_alpha_gap = 1e300; // near-tie diagnostic for the ratio test (not in the reference)

private:
    // lexlsi.h:148-173
    void activate(Index ObjIndex, Index CtrIndex, int type, bool CountActivation = true)
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (CtrIndex >= objectives[ObjIndex].nCtr) throw Exception("CtrIndex >= nCtr");
        WS.push_back(std::make_pair(ObjIndex, CtrIndex));
        objectives[ObjIndex].ws.activate(CtrIndex, type);
        if (CountActivation) nActivations++;
    }
    // lexlsi.h:181-198
    void deactivate(Index ObjIndex, Index CtrIndexActive)
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        std::pair<Index, Index> key(ObjIndex, objectives[ObjIndex].ws.active[CtrIndexActive]);
        WS.erase(std::find(WS.begin(), WS.end(), key));
        objectives[ObjIndex].ws.deactivate(CtrIndexActive);
        nDeactivations++;
    }

    // lexlsi.h:758-793
    void hot_start_related_tests()
    {
        bool partial   = false;
        bool attempted = objectives[0].v0_is_specified;
        for (Index k = 1; k < nObj; k++)
            if (objectives[k].v0_is_specified != attempted)
            {
                attempted = true;
                partial   = true;
                break;
            }
        bool forgot_x = (!x_guess_is_specified) && attempted;
        if (partial || forgot_x)
            for (Index k = 0; k < nObj; k++) objectives[k].v0_is_specified = false;
    }

    // lexlsi.h:816-869
    void phase1()
    {
        hot_start_related_tests();
        if (!x_guess_is_specified)
        {
            formLexLSE();
            lexlse.factorize();
            lexlse.solve();
            lexlse_rank = lexlse.getTotalRank();
            x           = lexlse.get_x();
        }
        for (Index k = 0; k < nObj; k++)
            objectives[k].phase1(x, x_guess_is_specified, parameters.modify_type_active_enabled,
                                 parameters.modify_type_inactive_enabled, parameters.modify_x_guess_enabled,
                                 parameters.set_min_init_ctr_violation, parameters.tol_feasibility);
        if (x_guess_is_specified)
        {
            formLexLSE();
            lexlse.factorize();
            lexlse.solve();
            lexlse_rank = lexlse.getTotalRank();
            for (Index i = 0; i < nVar; i++) dx[i] = lexlse.get_x()[i] - x[i];
        }
        for (Index k = 0; k < nObj; k++) objectives[k].formStep(dx);
        nFactorizations++;
    }

    // lexlsi.h:880-915
    void phase1_v0()
    {
        if (!x_guess_is_specified) throw Exception("when use_phase1_v0 = true, x_guess has to be specified");
        hot_start_related_tests();
        for (Index k = 0; k < nObj; k++)
            objectives[k].phase1(x, x_guess_is_specified, parameters.modify_type_active_enabled,
                                 parameters.modify_type_inactive_enabled, parameters.modify_x_guess_enabled,
                                 parameters.set_min_init_ctr_violation, parameters.tol_feasibility);
        for (Index k = 0; k < nObj; k++) objectives[k].formStep(dx);
    }

    // lexlsi.h:968-982
    void formLexLSE()
    {
        for (Index k = 0; k < nObj; k++) nActive[k] = getActiveCtrCount(k);
        lexlse.setObjDim(&nActive[0] + nObjOffset);
        Index counter = 0;
        for (Index k = 0; k < nObj; k++) objectives[k].formLexLSE(lexlse, counter, k - nObjOffset);
    }

    // lexlsi.h:1115-1139
    bool findActiveCtr2Remove_largest(Index &ObjIndex2Remove, Index &CtrIndex2Remove, double &lambda_wrong_sign)
    {
        bool found  = false;
        int obj_int = 0;
        for (Index k = 0; k < nObj - nObjOffset; k++)
        {
            found = lexlse.ObjectiveSensitivity(k, CtrIndex2Remove, obj_int, parameters.tol_wrong_sign_lambda,
                                                parameters.tol_correct_sign_lambda, lambda_wrong_sign);
            if (found) break;
        }
        ObjIndex2Remove = (Index)(obj_int + (int)nObjOffset);
        return found;
    }

    // lexlsi.h:1063-1105 (+ findFirstCtrWrongSign :1034-1046): of the constraints whose multiplier has the wrong sign for the
    // first objective that has any, remove the one that entered the working set FIRST (WS = chronological list of activations)
    bool findActiveCtr2Remove_first(Index &ObjIndex2Remove, Index &CtrIndex2Remove, double &lambda_wrong_sign)
    {
        if (nObjOffset > 0)
            throw Exception("deactivate_first_wrong_sign with a simple-bounds objective is undefined in the reference (lexlse.h:596-599)");
        std::vector<std::pair<int, int>> wrong;
        lambda_wrong_sign = 0; // lexlsi.h:1069
        bool found        = false;
        for (Index k = 0; k < nObj - nObjOffset; k++)
        {
            lexlse.ObjectiveSensitivityList(k, parameters.tol_wrong_sign_lambda, parameters.tol_correct_sign_lambda, wrong);
            if (!wrong.empty())
            {
                found = true;
                break;
            }
        }
        if (!found) return false;
        for (auto &w : wrong) // positions in the active lists -> constraint indices, lexlsi.h:1088-1095
        {
            w.first += (int)nObjOffset;
            w.second = (int)objectives[w.first].ws.active[w.second];
        }
        Index k = 0;
        for (;; k++)
            if (std::find(wrong.begin(), wrong.end(), std::make_pair((int)WS[k].first, (int)WS[k].second)) != wrong.end()) break;
        ObjIndex2Remove = WS[k].first;
        const std::vector<Index> &act = objectives[ObjIndex2Remove].ws.active;
        CtrIndex2Remove = (Index)(std::find(act.begin(), act.end(), WS[k].second) - act.begin()); // getCtrIndex, :1101
        return true;
    }

    // lexlsi.h:1144-1265
    OperationType verifyWorkingSet()
    {
        Index ObjIndex2Manipulate = 0, CtrIndex2Manipulate = 0;
        int CtrType2Manipulate  = CTR_INACTIVE;
        bool normalIteration    = true;
        OperationType operation = OPERATION_UNDEFINED;
        Index id_obj = 0, id_ctr = 0;
        int id_type = CTR_INACTIVE;
        double alpha;

        if (nIterations != 0)
        {
            formLexLSE();
            lexlse.factorize();
            lexlse.solve();
            lexlse_rank = lexlse.getTotalRank();
            for (Index i = 0; i < nVar; i++) dx[i] = lexlse.get_x()[i] - x[i]; // formStep :987-994
            for (Index k = 0; k < nObj; k++) objectives[k].formStep(dx);
            nFactorizations++;
        }
        else if (parameters.use_phase1_v0)
            normalIteration = false;

        // checkBlockingConstraints, lexlsi.h:1006-1029
        alpha         = 1;
        bool blocking = false;
        for (Index k = 0; k < nObj; k++)
            if (objectives[k].checkBlockingConstraints(CtrIndex2Manipulate, CtrType2Manipulate, alpha,
                                                       parameters.tol_feasibility, min_alpha_gap))
                ObjIndex2Manipulate = k;
        blocking = alpha < 1;

        if (blocking)
        {
            id_obj  = ObjIndex2Manipulate;
            id_ctr  = CtrIndex2Manipulate;
            id_type = CtrType2Manipulate;
            if (parameters.log_working_set_enabled)
                working_set_log.push_back({ObjIndex2Manipulate, CtrIndex2Manipulate, CtrType2Manipulate, alpha, lexlse_rank, false});
            operation = OPERATION_ADD;
            activate(ObjIndex2Manipulate, CtrIndex2Manipulate, CtrType2Manipulate);
        }
        else if (normalIteration)
        {
            double lambda_wrong_sign;
            const bool found = parameters.deactivate_first_wrong_sign // lexlsi.h:1048-1060
                                   ? findActiveCtr2Remove_first(ObjIndex2Manipulate, CtrIndex2Manipulate, lambda_wrong_sign)
                                   : findActiveCtr2Remove_largest(ObjIndex2Manipulate, CtrIndex2Manipulate, lambda_wrong_sign);
            if (found)
            {
                const WorkingSet &w = objectives[ObjIndex2Manipulate].ws;
                id_obj              = ObjIndex2Manipulate;
                id_ctr              = w.active[CtrIndex2Manipulate];
                id_type             = w.active_ctr_type[CtrIndex2Manipulate];
                if (parameters.log_working_set_enabled)
                    working_set_log.push_back({ObjIndex2Manipulate, id_ctr, CTR_INACTIVE, lambda_wrong_sign, lexlse_rank, false});
                operation = OPERATION_REMOVE;
                deactivate(ObjIndex2Manipulate, CtrIndex2Manipulate);
            }
            else
                status = PROBLEM_SOLVED;
        }

        if (alpha > 0)
        {
            for (Index i = 0; i < nVar; i++) x[i] = fmad(alpha, dx[i], x[i]);
            for (Index k = 0; k < nObj; k++) objectives[k].step(alpha);
        }

        // cycling.h:32-65
        if (parameters.cycling_handling_enabled && operation != OPERATION_UNDEFINED)
        {
            bool cycling_detected = false;
            TerminationStatus ret = TERMINATION_STATUS_UNKNOWN;
            bool early            = false;
            if (operation == OPERATION_ADD && cyc_prev_op == OPERATION_REMOVE && id_obj == cyc_prev_obj &&
                id_ctr == cyc_prev_ctr && id_type == cyc_prev_type)
            {
                if (cyc_counter >= parameters.cycling_max_counter)
                {
                    ret   = PROBLEM_SOLVED_CYCLING_HANDLING;
                    early = true;
                }
                else
                {
                    objectives[cyc_prev_obj].relax_bounds(cyc_prev_ctr, cyc_prev_type, parameters.cycling_relax_step);
                    cyc_counter++;
                    cycling_detected = true;
                }
            }
            if (!early)
            {
                cyc_prev_op   = operation;
                cyc_prev_obj  = id_obj;
                cyc_prev_ctr  = id_ctr;
                cyc_prev_type = id_type;
            }
            status = ret;
            if (parameters.log_working_set_enabled) working_set_log.back().cycling_detected = cycling_detected;
        }

        nIterations++;
        return operation;
    }

    Index nVar = 0, nObj = 0, nObjOffset = 0;
    Index nActivations = 0, nDeactivations = 0, nFactorizations = 0, nIterations = 0, lexlse_rank = 0;
    bool x_guess_is_specified = false;
    std::vector<double> x, dx;
    std::vector<Index> nActive;
    TerminationStatus status = TERMINATION_STATUS_UNKNOWN;
    LexLSE lexlse;
    std::vector<Objective> objectives;
    std::vector<WorkingSetLogEntry> working_set_log;
    Parameters parameters;
    std::vector<std::pair<Index, Index>> WS;
    // CyclingHandler state (cycling.h:85-109)
    Index cyc_counter = 0;
    int cyc_prev_op   = OPERATION_UNDEFINED;
    Index cyc_prev_obj = 0, cyc_prev_ctr = 0;
    int cyc_prev_type = CTR_INACTIVE;
};

} // namespace lexls_oracle
