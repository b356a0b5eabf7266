// lexls/lexlse.h — drop-in LexLS::internal::LexLSE (reference include/lexls/lexlse.h; member signatures :43-103, :117, :1015,
// :1381-1631) implemented as a batch-of-1 facade over the C-ABI in include/lexls_b200.h.  All arithmetic runs in the CUDA
// library; there is no CPU path.  factorize() launches the device l-QR + back-substitution; solve() only fetches.
#pragma once

#include <lexls/typedefs.h>
#include <lexls_b200.h>

namespace LexLS
{
namespace internal
{
inline void lexls_b200_check(int rc)
{
    if (rc != 0) throw Exception(lexls_b200_last_error());
}

class LexLSE
{
public:
    LexLSE() : h_(0), nVar(0), nObj(0), nCtr(0), nVarFixed(0), nVarFixedInit(0), TotalRank(0) {}
    LexLSE(Index nVar_, Index nObj_, Index *ObjDim_) : h_(0), nVarFixed(0), nVarFixedInit(0), TotalRank(0)
    {
        resize(nVar_, nObj_, ObjDim_);
        setObjDim(ObjDim_);
    }
    ~LexLSE() { release(); }
    LexLSE(const LexLSE &) = delete;
    LexLSE &operator=(const LexLSE &) = delete;

    void resize(Index nVar_, Index nObj_, Index *maxObjDim) // lexlse.h:67
    {
        nVar = nVar_;
        nObj = nObj_;
        maxDim.assign(maxObjDim, maxObjDim + nObj);
        dim = maxDim;
        x.resize(nVar);
    }
    void setObjDim(Index *ObjDim_) // lexlse.h:1426 — (re)creates the device batch for these dimensions
    {
        dim.assign(ObjDim_, ObjDim_ + nObj);
        nCtr = 0;
        first_row.assign(nObj, 0);
        for (Index k = 0; k < nObj; k++)
        {
            first_row[k] = nCtr;
            nCtr += dim[k];
        }
        release();
        lexls_b200_check(lexls_b200_create_batch(LEXLS_B200_LSE, nVar, nObj, dim.data(), 0, 1, 0, &h_));
        lexls_b200_check(lexls_b200_request_outputs(h_, LEXLS_B200_OUT_LEXQR)); // get_lexqr() is part of the surface
        applyParameters();
        data_.resize(nCtr > 0 ? nCtr : 1, nVar + 1);
        lexqr_.resize(nCtr > 0 ? nCtr : 1, nVar + 1);
        hh_.resize(nCtr > 0 ? nCtr : 1);
        ctr_type_.assign(nCtr, static_cast<uint8_t>(CTR_INACTIVE));
        dirty_ = true;
        nVarFixedInit = 0;
        TotalRank     = 0;
        ranks.assign(nObj, 0);
        v.resize(nCtr > 0 ? nCtr : 1);
    }
    void setParameters(const ParametersLexLSE &p) // lexlse.h:1467
    {
        parameters = p;
        if (h_) applyParameters();
    }
    void setData(Index ObjIndex, const dMatrixConstRef &data) // lexlse.h:1522 (copies: the caller's buffer may go away)
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (static_cast<Index>(data.rows()) != dim[ObjIndex] || static_cast<Index>(data.cols()) != nVar + 1)
            throw Exception("Incorrect number of equations");
        for (Index c = 0; c <= nVar; c++)
            for (Index i = 0; i < dim[ObjIndex]; i++) data_(first_row[ObjIndex] + i, c) = data(i, c);
        dirty_ = true;
    }
    void setCtr(Index CtrIndex, const dRowVectorType &row, RealScalar rhs) // lexlse.h:1539
    {
        if (CtrIndex >= nCtr) throw Exception("CtrIndex >= nCtr");
        for (Index c = 0; c < nVar; c++) data_(CtrIndex, c) = row(c);
        data_(CtrIndex, nVar) = rhs;
        dirty_ = true;
    }
    void setCtrType(Index ObjIndex, Index CtrIndex, ConstraintActivationType type) // lexlse.h:1548 (used by the multipliers only)
    {
        if (ObjIndex >= nObj || CtrIndex >= dim[ObjIndex]) throw Exception("CtrIndex >= nCtr");
        ctr_type_[first_row[ObjIndex] + CtrIndex] = static_cast<uint8_t>(type);
    }
    void setRegularizationFactor(Index ObjIndex, RealScalar factor) // lexlse.h:1478
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (reg_factor.size() < nObj) reg_factor.resize(nObj, 0.0);
        reg_factor[ObjIndex] = factor;
        if (h_) lexls_b200_check(lexls_b200_set_regularization_factor(h_, ObjIndex, factor));
    }
    void setFixedVariablesCount(Index nVarFixed_) // lexlse.h:1449
    {
        if (nVarFixed_ > nVar) throw Exception("Cannot fix more than nVar variables");
        nVarFixed = nVarFixed_;
        fixed_index.assign(nVarFixed, 0);
        fixed_value.assign(nVarFixed, 0.0);
        fixed_type.assign(nVarFixed, static_cast<uint8_t>(CTR_ACTIVE_UB));
        nVarFixedInit = 0;
    }
    void fixVariable(Index VarIndex, RealScalar VarValue, ConstraintActivationType type = CTR_ACTIVE_UB) // lexlse.h:1381
    {
        if (nVarFixedInit >= nVarFixed) throw Exception("more fixVariable calls than setFixedVariablesCount announced");
        fixed_index[nVarFixedInit] = VarIndex;
        fixed_value[nVarFixedInit] = VarValue;
        fixed_type[nVarFixedInit]  = static_cast<uint8_t>(type);
        nVarFixedInit++;
    }
    void fixVariables(Index nVarFixed_, Index *VarIndex, RealScalar *VarValue, ConstraintActivationType *type) // lexlse.h:1398
    {
        setFixedVariablesCount(nVarFixed_);
        for (Index i = 0; i < nVarFixed_; i++) fixVariable(VarIndex[i], VarValue[i], type[i]);
    }
    void factorize() // lexlse.h:117 (+ :1015: the device kernel also back-substitutes)
    {
        if (dirty_)
        {
            for (Index k = 0; k < nObj; k++)
                if (dim[k])
                    lexls_b200_check(lexls_b200_set_data(h_, k, data_.data() + first_row[k], data_.rows(),
                                                         static_cast<uint64_t>(data_.rows()) * (nVar + 1)));
            dirty_ = false;
        }
        lexls_b200_check(lexls_b200_set_fixed_variables(h_, nVarFixed, fixed_index.data(), fixed_value.data(), fixed_type.data()));
        lexls_b200_check(lexls_b200_solve_batch_sync(h_));
        lexls_b200_check(lexls_b200_get_ranks(h_, ranks.data()));
        uint32_t cnt[LEXLS_B200_NCOUNTERS];
        lexls_b200_check(lexls_b200_get_counters(h_, cnt));
        TotalRank = cnt[LEXLS_B200_CNT_TOTAL_RANK];
        if (nCtr) lexls_b200_check(lexls_b200_get_lexqr(h_, lexqr_.data(), hh_.data()));
    }
    void solve() { lexls_b200_check(lexls_b200_get_x(h_, x.data())); } // lexlse.h:1015
    // lexlse.h:1046-1262: the minimiser of least norm instead of the basic solution of solve().  The three variants of the
    // reference (Givens rotations, normal equations, the Tikhonov basis) compute the same vector with different roundings; all
    // three are served by the kernel's normal-equation routine (solveLeastNorm_2).  The factorization is launched once more with
    // the extra output switched on (data and fixed variables are still on the device).
    void solveLeastNorm_2()
    {
        lexls_b200_check(lexls_b200_request_outputs(h_, LEXLS_B200_OUT_LEXQR | LEXLS_B200_OUT_LEAST_NORM));
        lexls_b200_check(lexls_b200_solve_batch_sync(h_));
        lexls_b200_check(lexls_b200_get_x_least_norm(h_, x.data()));
    }
    void solveLeastNorm_1() { solveLeastNorm_2(); }
    void solveLeastNorm_3() { solveLeastNorm_2(); }
    const dVectorType &get_x() const { return x; }                      // lexlse.h:1587
    dVectorType &get_v()                                                // lexlse.h:1560 (use the first nCtr entries)
    {
        for (Index k = 0; k < nObj; k++)
            if (dim[k]) lexls_b200_check(lexls_b200_get_v(h_, k, v.data() + first_row[k]));
        return v;
    }
    const dMatrixType &get_lexqr() const { return lexqr_; } // lexlse.h:1626: nCtr x (nVar+1), R / Householder vectors / L
    const dMatrixType &get_data() const { return data_; }   // lexlse.h:1631: the constraints as they were given
    Index getDim(Index k) const { return dim[k]; }
    Index getRank(Index k) const { return ranks[k]; }
    Index getTotalRank() const { return TotalRank; }
    Index getFixedVariablesCount() const { return nVarFixed; }
    Index get_nObj() const { return nObj; }
    Index get_nVar() const { return nVar; }

private:
    void release()
    {
        if (h_) lexls_b200_destroy(h_);
        h_ = 0;
    }
    void applyParameters()
    {
        lexls_b200_params_t p;
        lexls_b200_default_params(&p);
        p.tol_linear_dependence          = parameters.tol_linear_dependence;
        p.regularization_type            = static_cast<int32_t>(parameters.regularization_type);
        p.variable_regularization_factor = parameters.variable_regularization_factor;
        p.max_number_of_cg_iterations    = static_cast<uint32_t>(parameters.max_number_of_CG_iterations);
        lexls_b200_check(lexls_b200_set_params(h_, &p)); // the conjugate-gradient variants and TIKHONOV_1 are refused here
        for (Index k = 0; k < nObj && k < reg_factor.size(); k++) lexls_b200_check(lexls_b200_set_regularization_factor(h_, k, reg_factor[k]));
    }
    lexls_b200_batch_t *h_;
    Index nVar, nObj, nCtr, nVarFixed, nVarFixedInit, TotalRank;
    std::vector<Index> maxDim, dim, first_row, ranks, fixed_index;
    std::vector<RealScalar> fixed_value, reg_factor;
    std::vector<uint8_t> fixed_type;
    ParametersLexLSE parameters;
    dVectorType x, v, hh_;
    dMatrixType data_, lexqr_;
    std::vector<uint8_t> ctr_type_;
    bool dirty_;
};
} // namespace internal
} // namespace LexLS
