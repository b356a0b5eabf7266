// lexls/batch.h — the batch entry point in C++: LexLS::LexLSIBatch / LexLS::LexLSEBatch (RAII over the C-ABI handle of
// include/lexls_b200.h) and the free functions LexLS::solveBatch.  One object owns `batch` problems of ONE shape
// (nVar, nObj, ObjDim, ObjType) and solves them with one kernel launch; the members mirror the single-problem classes
// (reference include/lexls/lexlsi.h:51, :325, :350, :444, :120, :306, :205, :508, :688 and lexlse.h:50, :1522, :1381,
// :117, :1015, :1587) with a leading problem index where a result of one problem is read.  The reference has no such
// entry point: this is the header a maintainer would add next to lexlsi.h.
#pragma once

#include <algorithm>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

#include <lexls/lexlsi.h>

namespace LexLS
{
namespace internal
{
// [batch][m x cols] column-major image of per-problem matrices (what lexls_b200_set_data takes)
inline void pack_level(const std::vector<dMatrixType> &per_problem, Index rows, Index cols, std::vector<RealScalar> &packed)
{
    packed.resize(per_problem.size() * static_cast<size_t>(rows) * cols);
    for (size_t q = 0; q < per_problem.size(); q++)
    {
        const dMatrixType &m = per_problem[q];
        if (static_cast<Index>(m.rows()) != rows || static_cast<Index>(m.cols()) != cols)
            throw Exception("solveBatch: a problem's matrix has the wrong size");
        std::copy(m.data(), m.data() + static_cast<size_t>(rows) * cols, packed.begin() + q * static_cast<size_t>(rows) * cols);
    }
}
} // namespace internal

namespace internal
{
// A set_* call whose source is a temporary (a packed scratch vector) must not return before the copy has left it: a no-op in
// the default mode (the C-ABI has waited already), one event wait in asynchronous-copies mode.
inline void lexls_b200_release_temporary(lexls_b200_batch_t *h) { lexls_b200_check(lexls_b200_wait_copies(h)); }
} // namespace internal

class LexLSIBatch
{
public:
    LexLSIBatch(Index nVar_, Index nObj_, Index *ObjDim_, ObjectiveType *ObjType_, uint64_t batch_, int device = 0)
    : h_(0), nVar(nVar_), nObj(nObj_), batch(batch_), dim(ObjDim_, ObjDim_ + nObj_), solved(false)
    {
        type.resize(nObj);
        for (Index k = 0; k < nObj; k++) type[k] = static_cast<int32_t>(ObjType_[k]);
        internal::lexls_b200_check(lexls_b200_create_batch(LEXLS_B200_LSI, nVar, nObj, dim.data(), type.data(), batch, device, &h_));
    }
    // the batch sharded over several GPUs of the box (contiguous ranges, one stream + host thread per device inside the library)
    LexLSIBatch(Index nVar_, Index nObj_, Index *ObjDim_, ObjectiveType *ObjType_, uint64_t batch_, const std::vector<int> &devices)
    : h_(0), nVar(nVar_), nObj(nObj_), batch(batch_), dim(ObjDim_, ObjDim_ + nObj_), solved(false)
    {
        type.resize(nObj);
        for (Index k = 0; k < nObj; k++) type[k] = static_cast<int32_t>(ObjType_[k]);
        internal::lexls_b200_check(lexls_b200_create_batch_multi(LEXLS_B200_LSI, nVar, nObj, dim.data(), type.data(), batch, devices.data(),
                                                                 static_cast<int>(devices.size()), &h_));
    }
    ~LexLSIBatch()
    {
        if (h_) lexls_b200_destroy(h_);
    }
    LexLSIBatch(const LexLSIBatch &) = delete;
    LexLSIBatch &operator=(const LexLSIBatch &) = delete;

    uint64_t size() const { return batch; }
    lexls_b200_batch_t *handle() { return h_; } // for the *_dev entry points and lexls_b200_set_stream
    void setRegularizationFactor(Index ObjIndex, RealScalar factor) // lexlsi.h:497 (one factor per objective for the whole batch)
    {
        internal::lexls_b200_check(lexls_b200_set_regularization_factor(h_, ObjIndex, factor));
    }

    void setParameters(const ParametersLexLSI &par) // lexlsi.h:325
    {
        lexls_b200_params_t p;
        lexls_b200_default_params(&p);
        p.regularization_type            = static_cast<int32_t>(par.regularization_type);
        p.variable_regularization_factor = par.variable_regularization_factor;
        p.max_number_of_cg_iterations    = static_cast<uint32_t>(par.max_number_of_CG_iterations);
        p.max_number_of_factorizations = par.max_number_of_factorizations;
        p.tol_linear_dependence        = par.tol_linear_dependence;
        p.tol_wrong_sign_lambda        = par.tol_wrong_sign_lambda;
        p.tol_correct_sign_lambda      = par.tol_correct_sign_lambda;
        p.tol_feasibility              = par.tol_feasibility;
        p.cycling_handling_enabled     = par.cycling_handling_enabled;
        p.cycling_max_counter          = par.cycling_max_counter;
        p.cycling_relax_step           = par.cycling_relax_step;
        p.modify_x_guess_enabled       = par.modify_x_guess_enabled;
        p.modify_type_active_enabled   = par.modify_type_active_enabled;
        p.modify_type_inactive_enabled = par.modify_type_inactive_enabled;
        p.set_min_init_ctr_violation   = par.set_min_init_ctr_violation;
        p.use_phase1_v0                = par.use_phase1_v0;
        p.log_working_set_enabled      = par.log_working_set_enabled;
        p.deactivate_first_wrong_sign  = par.deactivate_first_wrong_sign;
        internal::lexls_b200_check(lexls_b200_set_params(h_, &p));
    }

    // general level: [A | lb | ub] of every problem, m_k x (nVar + 2) each (lexlsi.h:350)
    void setData(Index ObjIndex, const std::vector<dMatrixType> &per_problem)
    {
        check_level(ObjIndex, per_problem.size());
        if (type[ObjIndex] == static_cast<int32_t>(SIMPLE_BOUNDS_OBJECTIVE)) throw Exception("setData: use the VarIndex overload for simple bounds");
        internal::pack_level(per_problem, dim[ObjIndex], nVar + 2, scratch);
        setDataPacked(ObjIndex, scratch.data(), dim[ObjIndex], static_cast<uint64_t>(dim[ObjIndex]) * (nVar + 2));
        internal::lexls_b200_release_temporary(h_);
    }
    // already packed host data: problem q at data + q * problem_stride, column-major with leading dimension ld
    void setDataPacked(Index ObjIndex, const RealScalar *data, Index ld, uint64_t problem_stride)
    {
        internal::lexls_b200_check(lexls_b200_set_data(h_, ObjIndex, data, ld > 0 ? ld : 1, problem_stride));
        solved = false;
    }
    // simple bounds (level 0): one VarIndex (0-based) shared by the batch, [lb | ub] m_0 x 2 per problem (lexlsi.h:444)
    void setData(Index ObjIndex, Index *VarIndex, const std::vector<dMatrixType> &per_problem)
    {
        check_level(ObjIndex, per_problem.size());
        if (ObjIndex != 0 || type[0] != static_cast<int32_t>(SIMPLE_BOUNDS_OBJECTIVE)) throw Exception("setData: level 0 is not SIMPLE_BOUNDS_OBJECTIVE");
        internal::pack_level(per_problem, dim[0], 2, scratch);
        internal::lexls_b200_check(lexls_b200_set_simple_bounds(h_, VarIndex, 0, scratch.data(), dim[0] > 0 ? dim[0] : 1,
                                                                static_cast<uint64_t>(dim[0]) * 2));
        internal::lexls_b200_release_temporary(h_);
        solved = false;
    }
    // warm start: the api_activate calls of problem q as one type per row (CTR_INACTIVE = no call), lexlsi.h:120
    void setActiveGuess(Index ObjIndex, const std::vector<std::vector<ConstraintActivationType>> &per_problem)
    {
        check_level(ObjIndex, per_problem.size());
        std::vector<uint8_t> t(batch * static_cast<size_t>(dim[ObjIndex]));
        for (uint64_t q = 0; q < batch; q++)
        {
            if (per_problem[q].size() != dim[ObjIndex]) throw Exception("setActiveGuess: wrong number of constraints");
            for (Index i = 0; i < dim[ObjIndex]; i++) t[q * dim[ObjIndex] + i] = static_cast<uint8_t>(per_problem[q][i]);
        }
        internal::lexls_b200_check(lexls_b200_set_active_guess(h_, ObjIndex, t.data()));
        internal::lexls_b200_release_temporary(h_);
        solved = false;
    }
    // warm start in CALL ORDER: calls[q] = the api_activate(obj, ctr, type) sequence of problem q (lexlsi.h:120-136)
    void setActiveGuessOrdered(const std::vector<std::vector<ConstraintIdentifier> > &calls)
    {
        if (calls.size() != batch) throw Exception("setActiveGuessOrdered: one call list per problem");
        size_t cnt = 0;
        for (uint64_t q = 0; q < batch; q++) cnt = std::max(cnt, calls[q].size());
        if (cnt == 0)
        {
            internal::lexls_b200_check(lexls_b200_set_active_guess_ordered(h_, 0, 0, 0, 0, 0));
            return;
        }
        std::vector<uint32_t> o(batch * cnt, 0), c(batch * cnt, 0);
        std::vector<uint8_t> t(batch * cnt, 0); // CTR_INACTIVE pads the shorter lists
        for (uint64_t q = 0; q < batch; q++)
            for (size_t e = 0; e < calls[q].size(); e++)
            {
                o[q * cnt + e] = calls[q][e].obj_index;
                c[q * cnt + e] = calls[q][e].ctr_index;
                t[q * cnt + e] = static_cast<uint8_t>(calls[q][e].ctr_type);
            }
        internal::lexls_b200_check(lexls_b200_set_active_guess_ordered(h_, static_cast<uint32_t>(cnt), o.data(), c.data(), t.data(), cnt));
        internal::lexls_b200_release_temporary(h_);
        solved = false;
    }
    void set_x0(const std::vector<dVectorType> &x0) // lexlsi.h:306
    {
        if (x0.size() != batch) throw Exception("set_x0: one vector per problem");
        std::vector<RealScalar> p(batch * static_cast<size_t>(nVar));
        for (uint64_t q = 0; q < batch; q++)
        {
            if (static_cast<Index>(x0[q].size()) != nVar) throw Exception("set_x0: wrong size");
            std::copy(x0[q].data(), x0[q].data() + nVar, p.begin() + q * nVar);
        }
        internal::lexls_b200_check(lexls_b200_set_x0(h_, p.data()));
        internal::lexls_b200_release_temporary(h_);
        solved = false;
    }
    void reset() // lexlsi.h:88 (keeps data and parameters)
    {
        internal::lexls_b200_check(lexls_b200_reset(h_));
        solved = false;
    }

    // one launch: the active-set loop of every problem runs on the device (lexlsi.h:205)
    void solve()
    {
        internal::lexls_b200_check(lexls_b200_solve_batch_sync(h_));
        x.resize(batch * static_cast<size_t>(nVar));
        status.resize(batch);
        counters.resize(batch * static_cast<size_t>(LEXLS_B200_NCOUNTERS));
        internal::lexls_b200_check(lexls_b200_get_x(h_, x.data()));
        internal::lexls_b200_check(lexls_b200_get_status(h_, status.data()));
        internal::lexls_b200_check(lexls_b200_get_counters(h_, counters.data()));
        active.assign(nObj, std::vector<uint8_t>());
        solved = true;
    }

    TerminationStatus getStatus(uint64_t q) const { need(q); return static_cast<TerminationStatus>(status[q]); }
    dVectorType get_x(uint64_t q) const // lexlsi.h:508
    {
        need(q);
        dVectorType r(nVar);
        std::copy(x.begin() + q * nVar, x.begin() + (q + 1) * nVar, r.data());
        return r;
    }
    const RealScalar *get_x_packed() const { need(0); return x.data(); } // [batch][nVar]
    Index getFactorizationsCount(uint64_t q) const { need(q); return counters[q * LEXLS_B200_NCOUNTERS + LEXLS_B200_CNT_FACTORIZATIONS]; }
    Index getActivationsCount(uint64_t q) const { need(q); return counters[q * LEXLS_B200_NCOUNTERS + LEXLS_B200_CNT_ACTIVATIONS]; }
    Index getDeactivationsCount(uint64_t q) const { need(q); return counters[q * LEXLS_B200_NCOUNTERS + LEXLS_B200_CNT_DEACTIVATIONS]; }
    Index getTotalRank(uint64_t q) const { need(q); return counters[q * LEXLS_B200_NCOUNTERS + LEXLS_B200_CNT_TOTAL_RANK]; }
    dVectorType get_v(uint64_t q, Index ObjIndex) // lexlsi.h:528
    {
        need(q);
        std::vector<RealScalar> t(batch * static_cast<size_t>(dim[ObjIndex]));
        internal::lexls_b200_check(lexls_b200_get_v(h_, ObjIndex, t.data()));
        dVectorType r(dim[ObjIndex]);
        std::copy(t.begin() + q * dim[ObjIndex], t.begin() + (q + 1) * dim[ObjIndex], r.data());
        return r;
    }
    void getActiveCtr(uint64_t q, Index ObjIndex, std::vector<ConstraintActivationType> &out) // lexlsi.h:688
    {
        need(q);
        if (active[ObjIndex].empty() && dim[ObjIndex])
        {
            active[ObjIndex].resize(batch * static_cast<size_t>(dim[ObjIndex]));
            internal::lexls_b200_check(lexls_b200_get_active_set(h_, ObjIndex, active[ObjIndex].data()));
        }
        out.resize(dim[ObjIndex]);
        for (Index i = 0; i < dim[ObjIndex]; i++) out[i] = static_cast<ConstraintActivationType>(active[ObjIndex][q * dim[ObjIndex] + i]);
    }

private:
    void check_level(Index ObjIndex, size_t count) const
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (count != batch) throw Exception("one matrix per problem of the batch is expected");
    }
    void need(uint64_t q) const
    {
        if (!solved) throw Exception("solve() has not been called");
        if (q >= batch) throw Exception("problem index >= batch");
    }
    lexls_b200_batch_t *h_;
    Index nVar, nObj;
    uint64_t batch;
    std::vector<Index> dim;
    std::vector<int32_t> type;
    std::vector<RealScalar> scratch, x;
    std::vector<int32_t> status;
    std::vector<uint32_t> counters;
    std::vector<std::vector<uint8_t>> active;
    bool solved;
};

class LexLSEBatch
{
public:
    LexLSEBatch(Index nVar_, Index nObj_, Index *ObjDim_, uint64_t batch_, int device = 0)
    : h_(0), nVar(nVar_), nObj(nObj_), batch(batch_), dim(ObjDim_, ObjDim_ + nObj_), solved(false)
    {
        internal::lexls_b200_check(lexls_b200_create_batch(LEXLS_B200_LSE, nVar, nObj, dim.data(), 0, batch, device, &h_));
    }
    LexLSEBatch(Index nVar_, Index nObj_, Index *ObjDim_, uint64_t batch_, const std::vector<int> &devices)
    : h_(0), nVar(nVar_), nObj(nObj_), batch(batch_), dim(ObjDim_, ObjDim_ + nObj_), solved(false)
    {
        internal::lexls_b200_check(lexls_b200_create_batch_multi(LEXLS_B200_LSE, nVar, nObj, dim.data(), 0, batch, devices.data(),
                                                                 static_cast<int>(devices.size()), &h_));
    }
    ~LexLSEBatch()
    {
        if (h_) lexls_b200_destroy(h_);
    }
    LexLSEBatch(const LexLSEBatch &) = delete;
    LexLSEBatch &operator=(const LexLSEBatch &) = delete;

    uint64_t size() const { return batch; }
    lexls_b200_batch_t *handle() { return h_; }
    void setRegularizationFactor(Index ObjIndex, RealScalar factor) // lexlse.h:1478
    {
        internal::lexls_b200_check(lexls_b200_set_regularization_factor(h_, ObjIndex, factor));
    }
    void setParameters(const ParametersLexLSE &par) // lexlse.h:1467
    {
        lexls_b200_params_t p;
        lexls_b200_default_params(&p);
        p.regularization_type            = static_cast<int32_t>(par.regularization_type);
        p.variable_regularization_factor = par.variable_regularization_factor;
        p.max_number_of_cg_iterations    = static_cast<uint32_t>(par.max_number_of_CG_iterations);
        p.tol_linear_dependence = par.tol_linear_dependence;
        internal::lexls_b200_check(lexls_b200_set_params(h_, &p));
    }
    void setData(Index ObjIndex, const std::vector<dMatrixType> &per_problem) // [A | b], m_k x (nVar + 1) each, lexlse.h:1522
    {
        if (ObjIndex >= nObj) throw Exception("ObjIndex >= nObj");
        if (per_problem.size() != batch) throw Exception("one matrix per problem of the batch is expected");
        internal::pack_level(per_problem, dim[ObjIndex], nVar + 1, scratch);
        setDataPacked(ObjIndex, scratch.data(), dim[ObjIndex], static_cast<uint64_t>(dim[ObjIndex]) * (nVar + 1));
        internal::lexls_b200_release_temporary(h_);
    }
    void setDataPacked(Index ObjIndex, const RealScalar *data, Index ld, uint64_t problem_stride)
    {
        internal::lexls_b200_check(lexls_b200_set_data(h_, ObjIndex, data, ld > 0 ? ld : 1, problem_stride));
        solved = false;
    }
    // the same `count` variables fixed in every problem slot order: var_index / value are [batch][count] (lexlse.h:1381, :1449)
    void fixVariables(Index count, const Index *var_index, const RealScalar *value)
    {
        internal::lexls_b200_check(lexls_b200_set_fixed_variables(h_, count, var_index, value, 0));
        solved = false;
    }
    void solve() // factorize + solve of every problem, one launch (lexlse.h:117, :1015)
    {
        internal::lexls_b200_check(lexls_b200_solve_batch_sync(h_));
        x.resize(batch * static_cast<size_t>(nVar));
        ranks.resize(batch * static_cast<size_t>(nObj));
        internal::lexls_b200_check(lexls_b200_get_x(h_, x.data()));
        internal::lexls_b200_check(lexls_b200_get_ranks(h_, ranks.data()));
        solved = true;
    }
    dVectorType get_x(uint64_t q) const // lexlse.h:1587
    {
        need(q);
        dVectorType r(nVar);
        std::copy(x.begin() + q * nVar, x.begin() + (q + 1) * nVar, r.data());
        return r;
    }
    const RealScalar *get_x_packed() const { need(0); return x.data(); }
    Index getRank(uint64_t q, Index ObjIndex) const { need(q); return ranks[q * nObj + ObjIndex]; } // lexlse.h:1603

private:
    void need(uint64_t q) const
    {
        if (!solved) throw Exception("solve() has not been called");
        if (q >= batch) throw Exception("problem index >= batch");
    }
    lexls_b200_batch_t *h_;
    Index nVar, nObj;
    uint64_t batch;
    std::vector<Index> dim;
    std::vector<RealScalar> scratch, x;
    std::vector<uint32_t> ranks;
    bool solved;
};

struct BatchResult
{
    std::vector<dVectorType> x;
    std::vector<TerminationStatus> status; // LexLSE: PROBLEM_SOLVED for every problem
};

// LexLSI: data[q][k] = [A | lb | ub] of level k of problem q (level 0 of a SIMPLE_BOUNDS hierarchy: [lb | ub], with VarIndex)
namespace internal
{
template <class Batch> inline BatchResult solve_lsi_batch(Batch &b, Index nObj, ObjectiveType *ObjType, const ParametersLexLSI &par,
                                                          const std::vector<std::vector<dMatrixType> > &data, Index *VarIndex);
template <class Batch> inline BatchResult solve_lse_batch(Batch &b, Index nObj, const ParametersLexLSE &par,
                                                          const std::vector<std::vector<dMatrixType> > &data);
} // namespace internal

// the same entry points over several GPUs of one box: `devices` = CUDA ordinals, the batch is split into contiguous shards
inline BatchResult solveBatch(Index nVar, Index nObj, Index *ObjDim, ObjectiveType *ObjType, const ParametersLexLSI &par,
                              const std::vector<std::vector<dMatrixType> > &data, Index *VarIndex, const std::vector<int> &devices)
{
    LexLSIBatch b(nVar, nObj, ObjDim, ObjType, data.size(), devices);
    return internal::solve_lsi_batch(b, nObj, ObjType, par, data, VarIndex);
}
inline BatchResult solveBatch(Index nVar, Index nObj, Index *ObjDim, const ParametersLexLSE &par,
                              const std::vector<std::vector<dMatrixType> > &data, const std::vector<int> &devices)
{
    LexLSEBatch b(nVar, nObj, ObjDim, data.size(), devices);
    return internal::solve_lse_batch(b, nObj, par, data);
}

inline BatchResult solveBatch(Index nVar, Index nObj, Index *ObjDim, ObjectiveType *ObjType, const ParametersLexLSI &par,
                              const std::vector<std::vector<dMatrixType>> &data, Index *VarIndex = 0, int device = 0)
{
    LexLSIBatch b(nVar, nObj, ObjDim, ObjType, data.size(), device);
    return internal::solve_lsi_batch(b, nObj, ObjType, par, data, VarIndex);
}
namespace internal
{
template <class Batch> inline BatchResult solve_lsi_batch(Batch &b, Index nObj, ObjectiveType *ObjType, const ParametersLexLSI &par,
                                                          const std::vector<std::vector<dMatrixType> > &data, Index *VarIndex)
{
    b.setParameters(par);
    std::vector<dMatrixType> level(data.size());
    for (Index k = 0; k < nObj; k++)
    {
        for (size_t q = 0; q < data.size(); q++)
        {
            if (data[q].size() != nObj) throw Exception("solveBatch: every problem needs nObj matrices");
            level[q] = data[q][k];
        }
        if (ObjType[k] == SIMPLE_BOUNDS_OBJECTIVE)
            b.setData(k, VarIndex, level);
        else
            b.setData(k, level);
    }
    b.solve();
    BatchResult r;
    for (uint64_t q = 0; q < b.size(); q++)
    {
        r.x.push_back(b.get_x(q));
        r.status.push_back(b.getStatus(q));
    }
    return r;
}
} // namespace internal

// LexLSE: data[q][k] = [A | b] of level k of problem q
inline BatchResult solveBatch(Index nVar, Index nObj, Index *ObjDim, const ParametersLexLSE &par,
                              const std::vector<std::vector<dMatrixType>> &data, int device = 0)
{
    LexLSEBatch b(nVar, nObj, ObjDim, data.size(), device);
    return internal::solve_lse_batch(b, nObj, par, data);
}
namespace internal
{
template <class Batch> inline BatchResult solve_lse_batch(Batch &b, Index nObj, const ParametersLexLSE &par,
                                                          const std::vector<std::vector<dMatrixType> > &data)
{
    b.setParameters(par);
    std::vector<dMatrixType> level(data.size());
    for (Index k = 0; k < nObj; k++)
    {
        for (size_t q = 0; q < data.size(); q++)
        {
            if (data[q].size() != nObj) throw Exception("solveBatch: every problem needs nObj matrices");
            level[q] = data[q][k];
        }
        b.setData(k, level);
    }
    b.solve();
    BatchResult r;
    for (uint64_t q = 0; q < b.size(); q++)
    {
        r.x.push_back(b.get_x(q));
        r.status.push_back(PROBLEM_SOLVED);
    }
    return r;
}
} // namespace internal
// ---------------------------------------------------------------------------------------------------------------
// Host-resident inputs, pipelined: the batch is split into `chunks` sub-batches, each with its own handle (its own CUDA
// stream, asynchronous-copies mode: lexls_b200_set_async_copies) and its own host thread.  The host->device copies of one
// chunk overlap the solve and the result read of the previous ones; the copy phases are serialised (the host link is the
// shared resource) and each chunk waits ONCE, for its copies alone (lexls_b200_wait_copies) — the device-side repacks are
// launched by the chunk's solve.  BatchT = LexLSEBatch or LexLSIBatch.  The caller's input buffers must stay valid until
// solve() returns; pinned memory gives the full link rate.
//
//   PipelinedBatch<LexLSEBatch> pb(batch, 4, [&](uint64_t count) { return new LexLSEBatch(nVar, nObj, ObjDim, count); });
//   pb.solve([&](LexLSEBatch &c, uint64_t first, uint64_t) {                       // set* calls of problems [first, first + count)
//       for (Index k = 0; k < nObj; k++) c.setDataPacked(k, level[k] + first * stride[k], ObjDim[k], stride[k]); },
//            x_out, nVar);
// ---------------------------------------------------------------------------------------------------------------
template <class BatchT> class PipelinedBatch
{
public:
    template <class Make> PipelinedBatch(uint64_t batch, int chunks, Make make)
    {
        if (chunks < 1) chunks = 1;
        if (static_cast<uint64_t>(chunks) > batch) chunks = static_cast<int>(batch);
        // equal chunks, but the last two are a half and a quarter: what remains after the last copy is the last chunk's solve
        std::vector<double> w(chunks, 1.0);
        if (chunks >= 3 && batch >= 8192) // smaller batches: the small chunks would not fill the GPU
        {
            w[chunks - 2] = 0.5;
            w[chunks - 1] = 0.25;
        }
        double tot = 0.0, acc = 0.0;
        for (int c = 0; c < chunks; c++) tot += w[c];
        std::vector<uint64_t> cut(chunks + 1, 0);
        for (int c = 0; c < chunks; c++)
        {
            acc += w[c];
            cut[c + 1] = (c + 1 == chunks) ? batch : static_cast<uint64_t>(batch * (acc / tot));
        }
        for (int c = 0; c < chunks; c++)
        {
            const uint64_t lo = cut[c], hi = cut[c + 1];
            first_.push_back(lo);
            part_.push_back(std::unique_ptr<BatchT>(make(hi - lo)));
            internal::lexls_b200_check(lexls_b200_set_async_copies(part_.back()->handle(), 1));
        }
    }
    size_t chunks() const { return part_.size(); }
    BatchT &chunk(size_t c) { return *part_[c]; }
    uint64_t first(size_t c) const { return first_[c]; }

    // feed(chunk, first, count) issues the host set* calls of the chunk's problems; x_out receives [batch][nVar]
    template <class Feed> void solve(Feed feed, RealScalar *x_out, Index nVar)
    {
        std::vector<std::string> errors(part_.size());
        std::vector<std::thread> workers;
        for (size_t c = 0; c < part_.size(); c++)
            workers.push_back(std::thread([this, c, &feed, x_out, nVar, &errors]() {
                try
                {
                    BatchT &b = *part_[c];
                    {
                        std::lock_guard<std::mutex> link(copy_phase_);
                        feed(b, first_[c], b.size());
                        internal::lexls_b200_check(lexls_b200_wait_copies(b.handle()));
                    }
                    internal::lexls_b200_check(lexls_b200_solve_batch(b.handle()));
                    internal::lexls_b200_check(lexls_b200_get_x(b.handle(), x_out + first_[c] * nVar)); // synchronises; reports bad inputs
                }
                catch (std::exception &e)
                {
                    errors[c] = e.what();
                    if (errors[c].empty()) errors[c] = "error";
                }
            }));
        for (size_t c = 0; c < workers.size(); c++) workers[c].join();
        for (size_t c = 0; c < errors.size(); c++)
            if (!errors[c].empty()) throw Exception(errors[c].c_str());
    }

private:
    std::vector<uint64_t> first_;
    std::vector<std::unique_ptr<BatchT> > part_;
    std::mutex copy_phase_;
};

} // namespace LexLS
