// lexls/typedefs.h — public types of the drop-in C++ API (source-compatible with the reference's header of the same name:
// include/lexls/typedefs.h:16-30 typedefs, :32-76 enums, :78-295 parameter structs, :300-314 Exception, :380-527 log/identifier
// structs).  With Eigen available the matrix types ARE the Eigen types the reference uses; without Eigen (as in this build
// image) a minimal column-major stand-in with the same accessors is provided so that callers and tests still compile.
#pragma once

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#if !defined(LEXLS_B200_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define LEXLS_B200_HAVE_EIGEN 1
#endif
#endif

namespace LexLS
{
typedef unsigned int Index;
typedef double RealScalar;

#ifdef LEXLS_B200_HAVE_EIGEN
typedef Eigen::Matrix<RealScalar, Eigen::Dynamic, Eigen::Dynamic, Eigen::ColMajor> dMatrixType;
typedef Eigen::Matrix<RealScalar, Eigen::Dynamic, 1> dVectorType;
typedef Eigen::Matrix<RealScalar, 1, Eigen::Dynamic> dRowVectorType;
typedef Eigen::Matrix<Index, Eigen::Dynamic, 1> iVectorType;
typedef Eigen::Ref<dMatrixType> dMatrixRef;
typedef Eigen::Ref<const dMatrixType> dMatrixConstRef;
namespace detail
{
inline const RealScalar *data_of(const dMatrixConstRef &m) { return m.data(); }
inline Index ld_of(const dMatrixConstRef &m) { return static_cast<Index>(m.outerStride()); }
} // namespace detail
#else
// ---- Eigen-free stand-ins (column-major, dynamic) --------------------------------------------------------------
template <typename T> class ColMajorMatrix
{
public:
    ColMajorMatrix() : r_(0), c_(0) {}
    ColMajorMatrix(Index r, Index c) : r_(r), c_(c), d_(static_cast<size_t>(r) * c, T()) {}
    void resize(Index r, Index c)
    {
        r_ = r;
        c_ = c;
        d_.assign(static_cast<size_t>(r) * c, T());
    }
    void setZero() { std::fill(d_.begin(), d_.end(), T()); }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index outerStride() const { return r_; }
    T *data() { return d_.data(); }
    const T *data() const { return d_.data(); }
    T &operator()(Index i, Index j) { return d_[static_cast<size_t>(j) * r_ + i]; }
    const T &operator()(Index i, Index j) const { return d_[static_cast<size_t>(j) * r_ + i]; }
    T &coeffRef(Index i, Index j) { return (*this)(i, j); }
    const T &coeff(Index i, Index j) const { return (*this)(i, j); }

private:
    Index r_, c_;
    std::vector<T> d_;
};
template <typename T> class DenseVector
{
public:
    DenseVector() {}
    explicit DenseVector(Index n) : d_(n, T()) {}
    void resize(Index n) { d_.assign(n, T()); }
    void setZero() { std::fill(d_.begin(), d_.end(), T()); }
    Index size() const { return static_cast<Index>(d_.size()); }
    Index rows() const { return size(); }
    T *data() { return d_.data(); }
    const T *data() const { return d_.data(); }
    T &operator()(Index i) { return d_[i]; }
    const T &operator()(Index i) const { return d_[i]; }
    T &operator[](Index i) { return d_[i]; }
    const T &operator[](Index i) const { return d_[i]; }
    T &coeffRef(Index i) { return d_[i]; }
    const T &coeff(Index i) const { return d_[i]; }

private:
    std::vector<T> d_;
};
typedef ColMajorMatrix<RealScalar> dMatrixType;
typedef DenseVector<RealScalar> dVectorType;
typedef DenseVector<RealScalar> dRowVectorType;
typedef DenseVector<Index> iVectorType;
// non-owning view with an outer stride, the role Eigen::Ref<const dMatrixType> plays in the reference (typedefs.h:30)
class dMatrixConstRef
{
public:
    dMatrixConstRef(const dMatrixType &m) : p_(m.data()), r_(m.rows()), c_(m.cols()), ld_(m.rows()) {}
    dMatrixConstRef(const RealScalar *p, Index r, Index c, Index ld) : p_(p), r_(r), c_(c), ld_(ld) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index outerStride() const { return ld_; }
    const RealScalar *data() const { return p_; }
    const RealScalar &operator()(Index i, Index j) const { return p_[static_cast<size_t>(j) * ld_ + i]; }
    const RealScalar &coeff(Index i, Index j) const { return (*this)(i, j); }

private:
    const RealScalar *p_;
    Index r_, c_, ld_;
};
typedef dMatrixConstRef dMatrixRef;
namespace detail
{
inline const RealScalar *data_of(const dMatrixConstRef &m) { return m.data(); }
inline Index ld_of(const dMatrixConstRef &m) { return m.outerStride(); }
} // namespace detail
#endif

enum RegularizationType
{
    REGULARIZATION_NONE = 0,
    REGULARIZATION_TIKHONOV,
    REGULARIZATION_TIKHONOV_CG,
    REGULARIZATION_R,
    REGULARIZATION_R_NO_Z,
    REGULARIZATION_RT_NO_Z,
    REGULARIZATION_RT_NO_Z_CG,
    REGULARIZATION_TIKHONOV_1,
    REGULARIZATION_TIKHONOV_2,
    REGULARIZATION_TEST
};
enum TerminationStatus
{
    TERMINATION_STATUS_UNKNOWN = -1,
    PROBLEM_SOLVED,
    PROBLEM_SOLVED_CYCLING_HANDLING,
    MAX_NUMBER_OF_FACTORIZATIONS_EXCEEDED
};
enum ObjectiveType
{
    GENERAL_OBJECTIVE = 0,
    SIMPLE_BOUNDS_OBJECTIVE
};
enum ConstraintActivationType
{
    CTR_INACTIVE = 0,
    CTR_ACTIVE_LB,
    CTR_ACTIVE_UB,
    CTR_ACTIVE_EQ,
    CORRECT_SIGN_OF_LAMBDA
};

class ParametersLexLSE
{
public:
    RealScalar tol_linear_dependence;
    Index max_number_of_CG_iterations;
    RegularizationType regularization_type;
    RealScalar variable_regularization_factor;
    ParametersLexLSE() { setDefaults(); }
    void setDefaults()
    {
        tol_linear_dependence          = 1e-12;
        max_number_of_CG_iterations    = 10;
        regularization_type            = REGULARIZATION_NONE;
        variable_regularization_factor = 0.0;
    }
};

class ParametersLexLSI
{
public:
    Index max_number_of_factorizations;
    RealScalar tol_linear_dependence, tol_wrong_sign_lambda, tol_correct_sign_lambda, tol_feasibility;
    RegularizationType regularization_type;
    Index max_number_of_CG_iterations;
    RealScalar variable_regularization_factor;
    bool cycling_handling_enabled;
    Index cycling_max_counter;
    RealScalar cycling_relax_step;
    std::string output_file_name;
    bool modify_x_guess_enabled, modify_type_active_enabled, modify_type_inactive_enabled, set_min_init_ctr_violation;
    bool use_phase1_v0, log_working_set_enabled, deactivate_first_wrong_sign;
    ParametersLexLSI() { setDefaults(); }
    void setDefaults()
    {
        max_number_of_factorizations   = 200;
        tol_linear_dependence          = 1e-12;
        tol_wrong_sign_lambda          = 1e-08;
        tol_correct_sign_lambda        = 1e-12;
        tol_feasibility                = 1e-13;
        cycling_handling_enabled       = false;
        cycling_max_counter            = 50;
        cycling_relax_step             = 1e-08;
        regularization_type            = REGULARIZATION_NONE;
        max_number_of_CG_iterations    = 10;
        variable_regularization_factor = 0.0;
        modify_x_guess_enabled         = false;
        modify_type_active_enabled     = false;
        modify_type_inactive_enabled   = false;
        set_min_init_ctr_violation     = true;
        use_phase1_v0                  = false;
        log_working_set_enabled        = false;
        deactivate_first_wrong_sign    = false;
    }
};

class Exception : public std::exception
{
public:
    explicit Exception(const char *message) : msg_(message) {}
    ~Exception() throw() {}
    const char *what() const throw() { return msg_.c_str(); }

private:
    std::string msg_;
};

// one step of the active-set method (reference typedefs.h:380-432): the constraint that was added (ctr_type = its
// activation type, alpha_or_lambda = step length) or removed (ctr_type = CTR_INACTIVE, alpha_or_lambda = the multiplier
// with the wrong sign), the rank of the factorization at that step and whether the cycling handler fired
class WorkingSetLogEntry
{
public:
    WorkingSetLogEntry() : obj_index(0), ctr_index(0), ctr_type(CTR_INACTIVE), alpha_or_lambda(0), rank(0), cycling_detected(false) {}
    WorkingSetLogEntry(Index o, Index c, ConstraintActivationType t, RealScalar a, Index r)
    : obj_index(o), ctr_index(c), ctr_type(t), alpha_or_lambda(a), rank(r), cycling_detected(false)
    {
    }
    Index obj_index, ctr_index;
    ConstraintActivationType ctr_type;
    RealScalar alpha_or_lambda;
    Index rank;
    bool cycling_detected;
};

class ConstraintIdentifier
{
public:
    ConstraintIdentifier() : obj_index(0), ctr_index(0), ctr_type(CTR_INACTIVE), alpha_or_lambda(0), cycling_detected(false) {}
    ConstraintIdentifier(Index o, Index c, ConstraintActivationType t)
    : obj_index(o), ctr_index(c), ctr_type(t), alpha_or_lambda(0), cycling_detected(false)
    {
    }
    void set(Index o, Index c, ConstraintActivationType t)
    {
        obj_index = o;
        ctr_index = c;
        ctr_type  = t;
    }
    bool operator==(const ConstraintIdentifier &ci) const
    {
        return obj_index == ci.obj_index && ctr_index == ci.ctr_index && ctr_type == ci.ctr_type;
    }
    Index obj_index, ctr_index;
    ConstraintActivationType ctr_type;
    RealScalar alpha_or_lambda;
    bool cycling_detected;
};

} // namespace LexLS
