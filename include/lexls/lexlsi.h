// lexls/lexlsi.h — drop-in LexLS::internal::LexLSI (reference include/lexls/lexlsi.h; member signatures :51-104, :120, :205,
// :306-545, :635-750) as a batch-of-1 facade over the C-ABI in include/lexls_b200.h.  solve() runs the whole active-set loop
// on the GPU.  api_activate calls are replayed by the kernel in CALL ORDER (lexls_b200_set_active_guess_ordered), as the
// reference applies them (lexlsi.h:120-136 -> workingset.h:67-77).  get_xStar / get_data / get_lexqr follow the reference's
// state semantics: after get_xStar(), get_data() and get_lexqr() describe the LexLSE problem of the final working set.
#pragma once

#include <lexls/lexlse.h>

namespace LexLS
{
namespace internal
{
class LexLSI
{
public:
    LexLSI() : h_(0), nVar(0), nObj(0), status(TERMINATION_STATUS_UNKNOWN) { std::memset(cnt, 0, sizeof cnt); }
    LexLSI(Index nVar_, Index nObj_, Index *ObjDim_, ObjectiveType *ObjType_) : h_(0), status(TERMINATION_STATUS_UNKNOWN)
    {
        std::memset(cnt, 0, sizeof cnt);
        resize(nVar_, nObj_, ObjDim_, ObjType_);
    }
    ~LexLSI() { release(); }
    LexLSI(const LexLSI &) = delete;
    LexLSI &operator=(const LexLSI &) = delete;

    void resize(Index nVar_, Index nObj_, Index *ObjDim_, ObjectiveType *ObjType_) // lexlsi.h:77
    {
        nVar = nVar_;
        nObj = nObj_;
        dim.assign(ObjDim_, ObjDim_ + nObj);
        std::vector<int32_t> types(nObj);
        for (Index k = 0; k < nObj; k++) types[k] = static_cast<int32_t>(ObjType_[k]);
        release();
        lexls_b200_check(lexls_b200_create_batch(LEXLS_B200_LSI, nVar, nObj, dim.data(), types.data(), 1, 0, &h_));
        // getLambda() / get_lexqr() / get_xStar() are part of the surface: a batch of one always asks the kernel for them
        lexls_b200_check(lexls_b200_request_outputs(h_, LEXLS_B200_OUT_LEXQR | LEXLS_B200_OUT_LAMBDA | LEXLS_B200_OUT_XSTAR));
        applyParameters();
        nGeneral = 0;
        for (Index k = 0; k < nObj; k++)
            if (ObjType_[k] != SIMPLE_BOUNDS_OBJECTIVE) nGeneral += dim[k];
        lexqr_.resize(nGeneral > 0 ? nGeneral : 1, nVar + 1);
        lexqr_star_.resize(nGeneral > 0 ? nGeneral : 1, nVar + 1);
        data_star_.resize(nGeneral > 0 ? nGeneral : 1, nVar + 1);
        xstar_.resize(nVar);
        star_called = false;
        g_obj.clear();
        g_ctr.clear();
        g_type.clear();
        guess_dirty = false;
        x.resize(nVar);
        v.assign(nObj, dVectorType());
        for (Index k = 0; k < nObj; k++) v[k].resize(dim[k]);
        status = TERMINATION_STATUS_UNKNOWN;
    }
    void reset() // lexlsi.h:88
    {
        lexls_b200_check(lexls_b200_reset(h_));
        g_obj.clear();
        g_ctr.clear();
        g_type.clear();
        guess_dirty = false;
        star_called = false;
        status      = TERMINATION_STATUS_UNKNOWN;
    }
    void setParameters(const ParametersLexLSI &p) // lexlsi.h:325
    {
        parameters = p;
        if (h_) applyParameters();
    }
    void setData(Index ObjIndex, const dMatrixConstRef &data) // lexlsi.h:350
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (static_cast<Index>(data.rows()) != dim[ObjIndex]) throw Exception("Incorrect number of equations");
        lexls_b200_check(lexls_b200_set_data(h_, ObjIndex, detail::data_of(data), detail::ld_of(data),
                                             static_cast<uint64_t>(detail::ld_of(data)) * (nVar + 2)));
        lexls_b200_check(lexls_b200_validate(h_)); // throws here, as lexlsi.h:430 does, when lb > ub
    }
    void setData(Index ObjIndex, Index *VarIndex, const dMatrixConstRef &data) // lexlsi.h:444
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (static_cast<Index>(data.rows()) != dim[ObjIndex]) throw Exception("Incorrect number of equations");
        lexls_b200_check(lexls_b200_set_simple_bounds(h_, VarIndex, 0, detail::data_of(data), detail::ld_of(data),
                                                      static_cast<uint64_t>(detail::ld_of(data)) * 2));
        lexls_b200_check(lexls_b200_validate(h_)); // lexlsi.h:474 (lb > ub), :485 (VarIndex not unique)
    }
    void api_activate(Index ObjIndex, Index CtrIndex, ConstraintActivationType type) // lexlsi.h:120
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (CtrIndex >= dim[ObjIndex]) throw Exception("CtrIndex >= nCtr");
        if (type == CTR_ACTIVE_LB || type == CTR_ACTIVE_UB)
        {
            // recorded in call order; the kernel skips a constraint that is already active (lexlsi.h:124)
            g_obj.push_back(ObjIndex);
            g_ctr.push_back(CtrIndex);
            g_type.push_back(static_cast<uint8_t>(type));
            guess_dirty = true;
        }
        else
            std::printf("WARNING: the user cannot define explicitly which constraints are of type CTR_ACTIVE_EQ \n\n");
    }
    void setRegularizationFactor(Index ObjIndex, RealScalar factor) // lexlsi.h:497
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        lexls_b200_check(lexls_b200_set_regularization_factor(h_, ObjIndex, factor));
    }
    void set_x0(const dVectorType &x0) { lexls_b200_check(lexls_b200_set_x0(h_, x0.data())); }                  // lexlsi.h:306
    void set_v0(Index ObjIndex, dVectorType &v0_) { lexls_b200_check(lexls_b200_set_v0(h_, ObjIndex, v0_.data())); } // :317

    TerminationStatus solve() // lexlsi.h:205
    {
        if (guess_dirty)
        {
            lexls_b200_check(lexls_b200_set_active_guess_ordered(h_, static_cast<uint32_t>(g_obj.size()), g_obj.data(), g_ctr.data(),
                                                                 g_type.data(), g_obj.size()));
            guess_dirty = false;
        }
        lexls_b200_check(lexls_b200_solve_batch_sync(h_));
        star_called = false;
        int32_t st = 0;
        lexls_b200_check(lexls_b200_get_status(h_, &st));
        status = static_cast<TerminationStatus>(st);
        lexls_b200_check(lexls_b200_get_counters(h_, cnt));
        lexls_b200_check(lexls_b200_get_x(h_, x.data()));
        for (Index k = 0; k < nObj; k++)
            if (dim[k]) lexls_b200_check(lexls_b200_get_v(h_, k, v[k].data()));
        if (nGeneral) lexls_b200_check(lexls_b200_get_lexqr(h_, lexqr_.data(), 0));
        lexls_b200_check(lexls_b200_get_xstar(h_, xstar_.data(), nGeneral ? data_star_.data() : 0, nGeneral ? lexqr_star_.data() : 0, 0));
        working_set_log.clear();
        if (parameters.log_working_set_enabled)
        {
            const uint32_t cap = parameters.max_number_of_factorizations + 1;
            std::vector<double> e(static_cast<size_t>(cap) * 6, 0.0);
            uint32_t count = 0;
            lexls_b200_check(lexls_b200_get_working_set_log(h_, cap, e.data(), &count));
            for (uint32_t i = 0; i < count && i < cap; i++)
            {
                WorkingSetLogEntry w(static_cast<Index>(e[6 * i]), static_cast<Index>(e[6 * i + 1]),
                                     static_cast<ConstraintActivationType>(static_cast<int>(e[6 * i + 2])), e[6 * i + 3],
                                     static_cast<Index>(e[6 * i + 4]));
                w.cycling_detected = e[6 * i + 5] != 0.0;
                working_set_log.push_back(w);
            }
        }
        return status;
    }
    std::vector<WorkingSetLogEntry> &getWorkingSetLog() { return working_set_log; } // lexlsi.h:739
    // lexlsi.h:552-605: vec_lambda[k] is ObjDim[k] x nObj, rows in the order of the problem definition
    void getLambda(std::vector<dMatrixType> &vec_lambda)
    {
        vec_lambda.resize(nObj);
        for (Index k = 0; k < nObj; k++)
        {
            vec_lambda[k].resize(dim[k], nObj);
            if (dim[k]) lexls_b200_check(lexls_b200_get_lambda(h_, k, vec_lambda[k].data()));
        }
    }
    // lexlsi.h:519-526: the solution of the LexLSE problem of the last active set (formed, factorized and solved once more —
    // by the solve kernel, at its end).  As in the reference, get_data() and get_lexqr() afterwards describe THAT problem.
    const dVectorType &get_xStar()
    {
        star_called = true;
        return xstar_;
    }
    // lexlsi.h:607: the last factorization (active rows in working-set order)
    const dMatrixType &get_lexqr() const { return star_called ? lexqr_star_ : lexqr_; }
    // lexlsi.h:612 -> lexlse.h:119, :1631: the constraints [A | rhs] of the inner LexLSE problem as they were given to its
    // last factorize().  Without a get_xStar() call that is the problem of the last in-loop factorization, which differs from
    // the one of the final working set only when the solve did not end PROBLEM_SOLVED.
    const dMatrixType &get_data()
    {
        if (!star_called && status != PROBLEM_SOLVED && status != TERMINATION_STATUS_UNKNOWN)
            throw Exception("get_data(): call get_xStar() first (the final working set was never factorized by solve())");
        return data_star_;
    }
    const dVectorType &get_x() const { return x; }                  // lexlsi.h:508
    const dVectorType &get_v(Index ObjIndex) const { return v[ObjIndex]; } // :528
    void getConstraintViolation(Index ObjIndex, dVectorType &out) const // :542
    {
        out.resize(dim[ObjIndex]);
        if (dim[ObjIndex]) lexls_b200_check(lexls_b200_get_constraint_violation(h_, ObjIndex, out.data()));
    }
    void getActiveCtr(Index ObjIndex, std::vector<ConstraintActivationType> &ctr_type) const // :688
    {
        std::vector<uint8_t> t(dim[ObjIndex], 0);
        if (dim[ObjIndex]) lexls_b200_check(lexls_b200_get_active_set(h_, ObjIndex, t.data()));
        ctr_type.resize(dim[ObjIndex]);
        for (Index i = 0; i < dim[ObjIndex]; i++) ctr_type[i] = static_cast<ConstraintActivationType>(t[i]);
    }
    void getActiveCtr_order(std::vector<ConstraintIdentifier> &ctr) const // :703
    {
        for (Index k = 0; k < nObj; k++)
        {
            if (!dim[k]) continue;
            std::vector<uint32_t> order(dim[k]);
            std::vector<uint8_t> t(dim[k]);
            lexls_b200_check(lexls_b200_get_active_order(h_, k, order.data()));
            lexls_b200_check(lexls_b200_get_active_set(h_, k, t.data()));
            for (Index a = 0; a < dim[k] && order[a] != 0xFFFFFFFFu; a++)
                ctr.push_back(ConstraintIdentifier(k, order[a], static_cast<ConstraintActivationType>(t[order[a]])));
        }
    }
    Index getFactorizationsCount() const { return cnt[LEXLS_B200_CNT_FACTORIZATIONS]; } // :643
    Index getActivationsCount() const { return cnt[LEXLS_B200_CNT_ACTIVATIONS]; }       // :652
    Index getDeactivationsCount() const { return cnt[LEXLS_B200_CNT_DEACTIVATIONS]; }   // :661
    Index getCyclingCounter() const { return cnt[LEXLS_B200_CNT_CYCLING]; }             // :635
    Index getActiveCtrCount() const { return cnt[LEXLS_B200_CNT_ACTIVE]; }              // :674
    Index getActiveCtrCount(Index ObjIndex) const                                        // :666
    {
        std::vector<ConstraintActivationType> t;
        getActiveCtr(ObjIndex, t);
        Index n = 0;
        for (size_t i = 0; i < t.size(); i++) n += (t[i] != CTR_INACTIVE);
        return n;
    }
    Index getTotalRank() { return cnt[LEXLS_B200_CNT_TOTAL_RANK]; } // :747
    Index getObjectivesCount() const { return nObj; }
    Index getObjDim(Index k) const { return dim[k]; }

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
        p.max_number_of_factorizations = parameters.max_number_of_factorizations;
        p.tol_linear_dependence        = parameters.tol_linear_dependence;
        p.tol_wrong_sign_lambda        = parameters.tol_wrong_sign_lambda;
        p.tol_correct_sign_lambda      = parameters.tol_correct_sign_lambda;
        p.tol_feasibility              = parameters.tol_feasibility;
        p.cycling_handling_enabled     = parameters.cycling_handling_enabled;
        p.cycling_max_counter          = parameters.cycling_max_counter;
        p.cycling_relax_step           = parameters.cycling_relax_step;
        p.regularization_type          = static_cast<int32_t>(parameters.regularization_type);
        p.modify_x_guess_enabled       = parameters.modify_x_guess_enabled;
        p.modify_type_active_enabled   = parameters.modify_type_active_enabled;
        p.modify_type_inactive_enabled = parameters.modify_type_inactive_enabled;
        p.set_min_init_ctr_violation   = parameters.set_min_init_ctr_violation;
        p.use_phase1_v0                = parameters.use_phase1_v0;
        p.log_working_set_enabled      = parameters.log_working_set_enabled;
        p.deactivate_first_wrong_sign  = parameters.deactivate_first_wrong_sign;
        p.variable_regularization_factor = parameters.variable_regularization_factor;
        p.max_number_of_cg_iterations    = static_cast<uint32_t>(parameters.max_number_of_CG_iterations);
        lexls_b200_check(lexls_b200_set_params(h_, &p));
    }
    lexls_b200_batch_t *h_;
    Index nVar, nObj;
    std::vector<Index> dim;
    std::vector<uint32_t> g_obj, g_ctr; // api_activate calls in call order
    std::vector<uint8_t> g_type;
    bool guess_dirty = false, star_called = false;
    dVectorType xstar_;
    dMatrixType lexqr_star_, data_star_;
    ParametersLexLSI parameters;
    TerminationStatus status;
    uint32_t cnt[LEXLS_B200_NCOUNTERS];
    dVectorType x;
    std::vector<dVectorType> v;
    Index nGeneral;
    dMatrixType lexqr_;
    std::vector<WorkingSetLogEntry> working_set_log;
};
} // namespace internal
} // namespace LexLS
