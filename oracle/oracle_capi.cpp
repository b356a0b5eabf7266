// oracle_capi.cpp — C entry points over the CPU ORACLE (test infrastructure, NOT product code).
//
// Batch layouts are the same "caller layouts" the product's C-ABI (include/lexls_b200.h) accepts:
// for level k a dense array [batch][m_k x cols_k] with every problem column-major and contiguous
// (cols = nVar+1 for LexLSE [A|b], nVar+2 for LexLSI [A|lb|ub], 2 for a simple-bounds level).
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may call this.
#include "lexls_oracle.hpp"

#include <atomic>
#include <cstring>
#include <thread>

using namespace lexls_oracle;

extern "C" {


struct oracle_params
{
    uint32_t max_number_of_factorizations;
    double tol_linear_dependence;
    double tol_wrong_sign_lambda;
    double tol_correct_sign_lambda;
    double tol_feasibility;
    int32_t cycling_handling_enabled;
    uint32_t cycling_max_counter;
    double cycling_relax_step;
    int32_t modify_x_guess_enabled;
    int32_t modify_type_active_enabled;
    int32_t modify_type_inactive_enabled;
    int32_t set_min_init_ctr_violation;
    int32_t use_phase1_v0;
    int32_t log_working_set_enabled;
    int32_t deactivate_first_wrong_sign;
};

// Regularization settings for the following oracle_*_solve_batch calls (test infrastructure: a process-wide setting keeps
// the batch signatures unchanged).  type = typedefs.h:32-44; factors[k] = setRegularizationFactor(k, .) per objective of
// the hierarchy as the caller numbers them (LexLSI: a simple-bounds objective 0 is counted and ignored, objective.h:489).
static int g_reg_type = 0;
static unsigned g_reg_cg_iterations = 10;
static double g_reg_var = 0.0;
static std::vector<double> g_reg_factors;
static int g_least_norm = 0; // LexLSE batches: solveLeastNorm instead of solve()

extern "C" int oracle_set_least_norm(int on)
{
    g_least_norm = on;
    return 0;
}

extern "C" int oracle_set_regularization(int type, double variable_factor, const double *factors, uint32_t count)
{
    g_reg_type = type;
    g_reg_var  = variable_factor;
    g_reg_factors.assign(factors ? factors : nullptr, factors ? factors + count : nullptr);
    return 0;
}
extern "C" int oracle_set_cg_iterations(uint32_t max_iterations)
{
    g_reg_cg_iterations = max_iterations;
    return 0;
}

static Parameters to_params(const oracle_params *p)
{
    Parameters q;
    q.regularization_type            = g_reg_type;
    q.variable_regularization_factor = g_reg_var;
    q.max_number_of_CG_iterations    = g_reg_cg_iterations;
    if (!p) return q;
    q.max_number_of_factorizations = p->max_number_of_factorizations;
    q.tol_linear_dependence        = p->tol_linear_dependence;
    q.tol_wrong_sign_lambda        = p->tol_wrong_sign_lambda;
    q.tol_correct_sign_lambda      = p->tol_correct_sign_lambda;
    q.tol_feasibility              = p->tol_feasibility;
    q.cycling_handling_enabled     = p->cycling_handling_enabled != 0;
    q.cycling_max_counter          = p->cycling_max_counter;
    q.cycling_relax_step           = p->cycling_relax_step;
    q.modify_x_guess_enabled       = p->modify_x_guess_enabled != 0;
    q.modify_type_active_enabled   = p->modify_type_active_enabled != 0;
    q.modify_type_inactive_enabled = p->modify_type_inactive_enabled != 0;
    q.set_min_init_ctr_violation   = p->set_min_init_ctr_violation != 0;
    q.use_phase1_v0                = p->use_phase1_v0 != 0;
    q.log_working_set_enabled      = p->log_working_set_enabled != 0;
    q.deactivate_first_wrong_sign  = p->deactivate_first_wrong_sign != 0;
    return q;
}

void oracle_default_params(oracle_params *p)
{
    Parameters q;
    p->max_number_of_factorizations = q.max_number_of_factorizations;
    p->tol_linear_dependence        = q.tol_linear_dependence;
    p->tol_wrong_sign_lambda        = q.tol_wrong_sign_lambda;
    p->tol_correct_sign_lambda      = q.tol_correct_sign_lambda;
    p->tol_feasibility              = q.tol_feasibility;
    p->cycling_handling_enabled     = q.cycling_handling_enabled;
    p->cycling_max_counter          = q.cycling_max_counter;
    p->cycling_relax_step           = q.cycling_relax_step;
    p->modify_x_guess_enabled       = q.modify_x_guess_enabled;
    p->modify_type_active_enabled   = q.modify_type_active_enabled;
    p->modify_type_inactive_enabled = q.modify_type_inactive_enabled;
    p->set_min_init_ctr_violation   = q.set_min_init_ctr_violation;
    p->use_phase1_v0                = q.use_phase1_v0;
    p->log_working_set_enabled      = q.log_working_set_enabled;
    p->deactivate_first_wrong_sign  = q.deactivate_first_wrong_sign;
}

static thread_local std::string g_err;
const char *oracle_last_error()
{
    return g_err.c_str();
}

} // extern "C"

template <class F> static int run_parallel(uint64_t batch, int threads, F body)
{
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > batch) threads = (int)std::max<uint64_t>(batch, 1);
    std::atomic<uint64_t> next(0);
    std::atomic<int> failed(0);
    std::string err;
    auto worker = [&]() {
        try
        {
            body(next, failed);
        }
        catch (std::exception &e)
        {
            if (!failed.exchange(1)) err = e.what();
        }
    };
    if (threads == 1)
        worker();
    else
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    if (failed.load())
    {
        g_err = err;
        return 1;
    }
    return 0;
}

extern "C" {

// ---------------------------------------------------------------------------------------------
// LexLSE batch: setData per level, optional fixed variables, factorize, solve, get_v.
//   data[k]      : [batch][m_k x (nVar+1)]
//   fixed_index  : [batch][nFixed] (or NULL), fixed_value: [batch][nFixed], fixed_type: [batch][nFixed] or NULL
// outputs (any may be NULL): x [batch][nVar], v [batch][nCtr], ranks [batch][nObj],
//   lexqr [batch][nCtr x (nVar+1)] (factorized LOD, column-major ld = nCtr), hh [batch][nCtr],
//   colperm [batch][nVar], total_rank [batch], margins [batch][2] = {min_pivot_gap, min_rank_margin}
// ---------------------------------------------------------------------------------------------
int oracle_lse_solve_batch(uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, uint64_t batch, const double *const *data,
                           uint32_t nFixed, const uint32_t *fixed_index, const double *fixed_value, const uint8_t *fixed_type,
                           double tol_linear_dependence, double *x, double *v, uint32_t *ranks, double *lexqr, double *hh,
                           uint32_t *colperm, uint32_t *total_rank, double *margins, int threads)
{
    uint32_t nCtr = 0;
    for (uint32_t k = 0; k < nObj; k++) nCtr += ObjDim[k];
    return run_parallel(batch, threads, [&](std::atomic<uint64_t> &next, std::atomic<int> &failed) {
        LexLSE lse;
        lse.resize(nVar, nObj, ObjDim);
        for (;;)
        {
            uint64_t p = next.fetch_add(1);
            if (p >= batch || failed.load()) break;
            lse.setTolLinearDependence(tol_linear_dependence);
            lse.setRegularization(g_reg_type, g_reg_var, g_reg_cg_iterations);
            lse.setObjDim(ObjDim);
            for (uint32_t k = 0; k < nObj; k++) lse.setRegularizationFactor(k, k < g_reg_factors.size() ? g_reg_factors[k] : 0.0);
            lse.setFixedVariablesCount(nFixed);
            for (uint32_t f = 0; f < nFixed; f++)
                lse.fixVariable(fixed_index[p * nFixed + f], fixed_value[p * nFixed + f],
                                fixed_type ? fixed_type[p * nFixed + f] : (int)CTR_ACTIVE_UB);
            for (uint32_t k = 0; k < nObj; k++)
                lse.setData(k, data[k] + p * (uint64_t)ObjDim[k] * (nVar + 1), ObjDim[k]);
            lse.min_pivot_gap = lse.min_rank_margin = 1e300;
            lse.factorize();
            if (g_least_norm)
                lse.solveLeastNorm();
            else
                lse.solve();
            if (x) std::memcpy(x + p * nVar, lse.get_x().data(), sizeof(double) * nVar);
            if (v) std::memcpy(v + p * nCtr, lse.get_v(), sizeof(double) * nCtr);
            if (ranks)
                for (uint32_t k = 0; k < nObj; k++) ranks[p * nObj + k] = lse.getRank(k);
            if (lexqr)
                for (uint32_t c = 0; c <= nVar; c++)
                    for (uint32_t r = 0; r < nCtr; r++)
                        lexqr[p * (uint64_t)nCtr * (nVar + 1) + (uint64_t)c * nCtr + r] = lse.get_lexqr()[(size_t)c * lse.get_ld() + r];
            if (hh) std::memcpy(hh + p * nCtr, lse.getHouseholderScalars().data(), sizeof(double) * nCtr);
            if (colperm) std::memcpy(colperm + p * nVar, lse.getColumnPermutations().data(), sizeof(uint32_t) * nVar);
            if (total_rank) total_rank[p] = lse.getTotalRank();
            if (margins)
            {
                margins[p * 2 + 0] = lse.min_pivot_gap;
                margins[p * 2 + 1] = lse.min_rank_margin;
            }
        }
    });
}

// ---------------------------------------------------------------------------------------------
// LexLSI batch.  Call order per problem follows the MEX caller (interfaces/matlab-octave/lexlsi.cpp:527-625):
// setParameters, setData level 0.., set_x0, set_v0, api_activate (level-major, row ascending), solve.
//   data[k]   : general [batch][m_k x (nVar+2)], simple bounds [batch][m_0 x 2]
//   var_index : [batch][m_0] when ObjType[0] is SIMPLE_BOUNDS (0-based), else NULL
//   guess[k]  : [batch][m_k] activation types (0 = none) or NULL; guess itself may be NULL
//   x0        : [batch][nVar] or NULL;  v0[k]: [batch][m_k] or NULL (v0 itself may be NULL)
// outputs (any may be NULL): x [batch][nVar]; v[k] [batch][m_k] (Objective::v); viol[k] [batch][m_k]
//   (getConstraintViolation); active[k] [batch][m_k] (getActiveCtr types); status [batch];
//   counters [batch][8] = {nFactorizations, nActivations, nDeactivations, nIterations, cyclingCounter,
//   totalRank, nActiveTotal, 0}; margins [batch][4] = {min_pivot_gap, min_rank_margin, min_alpha_gap, min_lambda_gap};
//   active_order [batch][sum m_k] = concatenated active[] lists per level (padded with 0xFFFFFFFF)
// ---------------------------------------------------------------------------------------------
// Extras of the _ex entry: api_activate calls in CALL ORDER (lexlsi.h:120-136) as per-problem lists
//   (g_obj, g_ctr, g_type)[p * g_count + e], type 0 = padding — used instead of guess[] when g_count > 0;
//   xstar [batch][nVar] = get_xStar() (lexlsi.h:519), called after solve(); then lexdata / lexqr_star
//   [batch][(nVar+1) x M] (column-major, leading dimension M = rows of all general levels, rows >= nCtr zero) =
//   get_data() / get_lexqr() as tests/test_numerical_error.cpp:83-86 reads them.
static int lsi_batch_impl(uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType, uint64_t batch,
                          const double *const *data, const uint32_t *var_index, const uint8_t *const *guess, const double *x0,
                          const double *const *v0, const oracle_params *params, double *x, double *const *v,
                          double *const *viol, uint8_t *const *active, int32_t *status, uint32_t *counters, double *margins,
                          uint32_t *active_order, int threads, uint32_t g_count, const uint32_t *g_obj, const uint32_t *g_ctr,
                          const uint8_t *g_type, double *xstar, double *lexdata, double *lexqr_star);

int oracle_lsi_solve_batch(uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType, uint64_t batch,
                           const double *const *data, const uint32_t *var_index, const uint8_t *const *guess, const double *x0,
                           const double *const *v0, const oracle_params *params, double *x, double *const *v,
                           double *const *viol, uint8_t *const *active, int32_t *status, uint32_t *counters, double *margins,
                           uint32_t *active_order, int threads)
{
    return lsi_batch_impl(nVar, nObj, ObjDim, ObjType, batch, data, var_index, guess, x0, v0, params, x, v, viol, active, status,
                          counters, margins, active_order, threads, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int oracle_lsi_solve_batch_ex(uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType, uint64_t batch,
                              const double *const *data, const uint32_t *var_index, const uint8_t *const *guess, const double *x0,
                              const double *const *v0, const oracle_params *params, double *x, double *const *v,
                              double *const *viol, uint8_t *const *active, int32_t *status, uint32_t *counters, double *margins,
                              uint32_t *active_order, int threads, uint32_t g_count, const uint32_t *g_obj, const uint32_t *g_ctr,
                              const uint8_t *g_type, double *xstar, double *lexdata, double *lexqr_star)
{
    return lsi_batch_impl(nVar, nObj, ObjDim, ObjType, batch, data, var_index, guess, x0, v0, params, x, v, viol, active, status,
                          counters, margins, active_order, threads, g_count, g_obj, g_ctr, g_type, xstar, lexdata, lexqr_star);
}

} // extern "C"

static int lsi_batch_impl(uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType, uint64_t batch,
                          const double *const *data, const uint32_t *var_index, const uint8_t *const *guess, const double *x0,
                          const double *const *v0, const oracle_params *params, double *x, double *const *v,
                          double *const *viol, uint8_t *const *active, int32_t *status, uint32_t *counters, double *margins,
                          uint32_t *active_order, int threads, uint32_t g_count, const uint32_t *g_obj, const uint32_t *g_ctr,
                          const uint8_t *g_type, double *xstar, double *lexdata, double *lexqr_star)
{
    Parameters par = to_params(params);
    std::vector<int> types(nObj);
    uint32_t mtot = 0;
    for (uint32_t k = 0; k < nObj; k++)
    {
        types[k] = ObjType ? ObjType[k] : 0;
        mtot += ObjDim[k];
    }
    return run_parallel(batch, threads, [&](std::atomic<uint64_t> &next, std::atomic<int> &failed) {
        LexLSI lsi;
        std::vector<double> tmp;
        std::vector<int> at;
        for (;;)
        {
            uint64_t p = next.fetch_add(1);
            if (p >= batch || failed.load()) break;
            lsi.resize(nVar, nObj, ObjDim, types.data());
            lsi.setParameters(par);
            for (uint32_t k = 0; k < nObj; k++) lsi.setRegularizationFactor(k, k < g_reg_factors.size() ? g_reg_factors[k] : 0.0);
            lsi.min_alpha_gap = 1e300;
            lsi.getLexLSE().min_pivot_gap = lsi.getLexLSE().min_rank_margin = lsi.getLexLSE().min_lambda_gap = 1e300;
            for (uint32_t k = 0; k < nObj; k++)
            {
                if (types[k] == SIMPLE_BOUNDS_OBJECTIVE)
                    lsi.setData(k, var_index + p * ObjDim[k], data[k] + p * (uint64_t)ObjDim[k] * 2, ObjDim[k]);
                else
                    lsi.setData(k, data[k] + p * (uint64_t)ObjDim[k] * (nVar + 2), ObjDim[k]);
            }
            if (x0) lsi.set_x0(x0 + p * nVar);
            if (v0)
                for (uint32_t k = 0; k < nObj; k++)
                    if (v0[k]) lsi.set_v0(k, v0[k] + p * ObjDim[k]);
            if (g_count > 0)
            {
                for (uint32_t e = 0; e < g_count; e++)
                {
                    const uint64_t at = p * (uint64_t)g_count + e;
                    if (g_type[at] != CTR_INACTIVE) lsi.api_activate(g_obj[at], g_ctr[at], g_type[at]);
                }
            }
            else if (guess)
                for (uint32_t k = 0; k < nObj; k++)
                    if (guess[k])
                        for (uint32_t i = 0; i < ObjDim[k]; i++)
                        {
                            int t = guess[k][p * ObjDim[k] + i];
                            if (t != CTR_INACTIVE) lsi.api_activate(k, i, t);
                        }
            lsi.solve();
            if (x) std::memcpy(x + p * nVar, lsi.get_x().data(), sizeof(double) * nVar);
            uint32_t nact = 0, off = 0;
            for (uint32_t k = 0; k < nObj; k++)
            {
                if (v && v[k]) std::memcpy(v[k] + p * ObjDim[k], lsi.get_v(k).data(), sizeof(double) * ObjDim[k]);
                if (viol && viol[k])
                {
                    lsi.getConstraintViolation(k, tmp);
                    std::memcpy(viol[k] + p * ObjDim[k], tmp.data(), sizeof(double) * ObjDim[k]);
                }
                if (active && active[k])
                {
                    lsi.getActiveCtr(k, at);
                    for (uint32_t i = 0; i < ObjDim[k]; i++) active[k][p * ObjDim[k] + i] = (uint8_t)at[i];
                }
                if (active_order)
                {
                    const WorkingSet &w = lsi.getWorkingSet(k);
                    for (uint32_t i = 0; i < ObjDim[k]; i++)
                        active_order[p * mtot + off + i] = i < w.active.size() ? w.active[i] : 0xFFFFFFFFu;
                }
                off += ObjDim[k];
                nact += lsi.getActiveCtrCount(k);
            }
            if (status) status[p] = (int32_t)lsi.getStatus();
            if (counters)
            {
                uint32_t *c = counters + p * 8;
                c[0]        = lsi.getFactorizationsCount();
                c[1]        = lsi.getActivationsCount();
                c[2]        = lsi.getDeactivationsCount();
                c[3]        = lsi.getIterationsCount();
                c[4]        = lsi.getCyclingCounter();
                c[5]        = lsi.getTotalRank();
                c[6]        = nact;
                c[7]        = 0;
            }
            if (margins)
            {
                margins[p * 4 + 0] = lsi.getLexLSE().min_pivot_gap;
                margins[p * 4 + 1] = lsi.getLexLSE().min_rank_margin;
                margins[p * 4 + 2] = lsi.min_alpha_gap;
                margins[p * 4 + 3] = lsi.getLexLSE().min_lambda_gap;
            }
            if (xstar || lexdata || lexqr_star)
            {
                // the call sequence of tests/test_numerical_error.cpp:83-86: get_xStar(), then get_data() and get_lexqr()
                const std::vector<double> &xs = lsi.get_xStar();
                if (xstar) std::memcpy(xstar + p * nVar, xs.data(), sizeof(double) * nVar);
                uint32_t M = 0;
                for (uint32_t k = 0; k < nObj; k++)
                    if (types[k] != SIMPLE_BOUNDS_OBJECTIVE) M += ObjDim[k];
                const LexLSE &e     = lsi.getLexLSE();
                const uint32_t nCtr = e.get_nCtr(), ld = e.get_ld();
                for (int which = 0; which < 2; which++)
                {
                    double *dst = which ? lexqr_star : lexdata;
                    if (!dst) continue;
                    const std::vector<double> &src = which ? e.get_lexqr() : e.get_data();
                    dst += p * (uint64_t)M * (nVar + 1);
                    for (uint32_t c = 0; c <= nVar; c++)
                        for (uint32_t r = 0; r < M; r++) dst[(uint64_t)c * M + r] = r < nCtr ? src[(uint64_t)c * ld + r] : 0.0;
                }
            }
        }
    });
}

extern "C" {

// ---------------------------------------------------------------------------------------------
// Single LexLSI problem with debugging outputs: multipliers (getLambda, lexlsi.h:552-605) and the
// working-set log (typedefs.h:380-432).  lambda[k] is m_k x nObj column-major.  log entries are
// rows of 6 doubles {obj, ctr, type, alpha_or_lambda, rank, cycling_detected}; returns their count
// (at most log_capacity are written), or -1 on error.
// ---------------------------------------------------------------------------------------------
int oracle_lsi_solve_detail(uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType,
                            const double *const *data, const uint32_t *var_index, const uint8_t *const *guess, const double *x0,
                            const oracle_params *params, double *x, int32_t *status, uint32_t *counters, double *const *lambda,
                            double *log, int log_capacity)
{
    try
    {
        Parameters par = to_params(params);
        std::vector<int> types(nObj);
        for (uint32_t k = 0; k < nObj; k++) types[k] = ObjType ? ObjType[k] : 0;
        LexLSI lsi(nVar, nObj, ObjDim, types.data());
        lsi.setParameters(par);
        for (uint32_t k = 0; k < nObj; k++)
        {
            if (types[k] == SIMPLE_BOUNDS_OBJECTIVE)
                lsi.setData(k, var_index, data[k], ObjDim[k]);
            else
                lsi.setData(k, data[k], ObjDim[k]);
        }
        if (x0) lsi.set_x0(x0);
        if (guess)
            for (uint32_t k = 0; k < nObj; k++)
                if (guess[k])
                    for (uint32_t i = 0; i < ObjDim[k]; i++)
                        if (guess[k][i] != CTR_INACTIVE) lsi.api_activate(k, i, guess[k][i]);
        lsi.solve();
        if (x) std::memcpy(x, lsi.get_x().data(), sizeof(double) * nVar);
        if (status) *status = (int32_t)lsi.getStatus();
        if (counters)
        {
            counters[0] = lsi.getFactorizationsCount();
            counters[1] = lsi.getActivationsCount();
            counters[2] = lsi.getDeactivationsCount();
            counters[3] = lsi.getIterationsCount();
            counters[4] = lsi.getCyclingCounter();
            counters[5] = lsi.getTotalRank();
        }
        int nlog = (int)lsi.getWorkingSetLog().size();
        if (log)
            for (int i = 0; i < nlog && i < log_capacity; i++)
            {
                const WorkingSetLogEntry &e = lsi.getWorkingSetLog()[i];
                double *row                 = log + 6 * i;
                row[0]                      = e.obj_index;
                row[1]                      = e.ctr_index;
                row[2]                      = e.ctr_type;
                row[3]                      = e.alpha_or_lambda;
                row[4]                      = e.rank;
                row[5]                      = e.cycling_detected;
            }
        if (lambda)
        {
            std::vector<std::vector<double>> vl;
            lsi.getLambda(vl);
            for (uint32_t k = 0; k < nObj; k++)
                if (lambda[k]) std::memcpy(lambda[k], vl[k].data(), sizeof(double) * vl[k].size());
        }
        return nlog;
    }
    catch (std::exception &e)
    {
        g_err = e.what();
        return -1;
    }
}

} // extern "C"
