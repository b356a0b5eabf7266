// lexls/tools.h — LexLS::tools::HierarchyFileProcessor: reader of the reference's text hierarchy format (".dat"), the
// importer the reference's own C++ callers use (reference include/lexls/tools.h:273-453; used by tests/test_01.cpp:40 and
// tests/test_numerical_error.cpp:52).  Same signature, same field names, same error messages; written as a line tokenizer.
//
//   #HierType  100 equalities [A b] | 200 inequalities [A lb ub] | 210 = 200 + a trailing activation type per row
//   #nVar n    #nObj P    #nCtr m_0 .. m_{P-1}    #ObjType 100 (simple bounds, first objective only) | 200 (general)
//   #OBJECTIVE k   followed by m_k rows; a simple-bounds row is  var_index(1-based) {b | lb ub}
//   #SolGuess / #Solution   n numbers each (optional)
// The five header fields may come in any order; surplus rows / columns are ignored (tools.h:203).
#pragma once

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <lexls/typedefs.h>

namespace LexLS
{
namespace tools
{
enum HierarchyType
{
    HIERARCHY_TYPE_NONE       = 0,
    HIERARCHY_TYPE_EQUALITY   = 1,
    HIERARCHY_TYPE_INEQUALITY = 2
};

class HierarchyFileProcessor
{
public:
    // reference tools.h:287-297; with Eigen dMatrixType / dVectorType ARE Eigen::MatrixXd / Eigen::VectorXd
    void import(const std::string &file_name, HierarchyType &type_of_hierarchy, unsigned int &number_of_variables,
                unsigned int &number_of_objectives, std::vector<unsigned int> &number_of_constraints,
                std::vector<LexLS::ObjectiveType> &types_of_objectives, std::vector<dMatrixType> &objectives,
                std::vector<std::vector<ConstraintActivationType> > &active_set_guess, dVectorType &solution_guess,
                dVectorType &solution)
    {
        type_of_hierarchy    = HIERARCHY_TYPE_NONE;
        number_of_variables  = 0;
        number_of_objectives = 0;
        number_of_constraints.clear();
        types_of_objectives.clear();
        objectives.clear();
        active_set_guess.clear();
        solution_guess.resize(0);
        solution.resize(0);

        std::ifstream in(file_name.c_str());
        if (!in) throw std::runtime_error("Cannot open file for reading");
        std::vector<std::string> lines;
        for (std::string l; std::getline(in, l);) lines.push_back(l);
        size_t at = 0;

        // ---- header: five keyed fields, any order -------------------------------------------------------------
        enum { K_NVAR = 1, K_NOBJ = 2, K_NCTR = 4, K_HIER = 8, K_TYPE = 16, K_ALL = 31 };
        unsigned seen = 0, hier_code = 0;
        while (at < lines.size() && seen != K_ALL)
        {
            const std::string key = squeeze(lines[at++]);
            unsigned bit          = 0;
            if (key == "#nVar") bit = K_NVAR;
            else if (key == "#nObj") bit = K_NOBJ;
            else if (key == "#nCtr") bit = K_NCTR;
            else if (key == "#HierType") bit = K_HIER;
            else if (key == "#ObjType") bit = K_TYPE;
            if (!bit) continue;
            if (seen & bit) throw std::runtime_error("Duplicate header field.");
            if (at >= lines.size()) break;
            const std::vector<double> val = numbers(lines[at++]);
            if (bit == K_NCTR)
            {
                for (size_t i = 0; i < val.size(); i++) number_of_constraints.push_back(static_cast<unsigned int>(val[i]));
                seen |= bit;
                continue;
            }
            if (bit == K_TYPE)
            {
                for (size_t i = 0; i < val.size(); i++)
                {
                    const unsigned t = static_cast<unsigned>(val[i]);
                    if (t == 100) types_of_objectives.push_back(LexLS::SIMPLE_BOUNDS_OBJECTIVE);
                    else if (t == 200) types_of_objectives.push_back(LexLS::GENERAL_OBJECTIVE);
                    else throw std::runtime_error("Unsupported type of objective.");
                }
                seen |= bit;
                continue;
            }
            if (val.empty()) continue; // the field stays unset, as a failed stream read leaves it (tools.h:60-66)
            if (bit == K_NVAR) number_of_variables = static_cast<unsigned int>(val[0]);
            if (bit == K_NOBJ) number_of_objectives = static_cast<unsigned int>(val[0]);
            if (bit == K_HIER)
            {
                hier_code = static_cast<unsigned>(val[0]);
                if (hier_code == 100) type_of_hierarchy = HIERARCHY_TYPE_EQUALITY;
                else if (hier_code == 200 || hier_code == 210) type_of_hierarchy = HIERARCHY_TYPE_INEQUALITY;
                else throw std::runtime_error("Unsupported type of hierarchy.");
            }
            seen |= bit;
        }
        if (seen != K_ALL) throw std::runtime_error("At least one required parameters is not set.");
        if (types_of_objectives.size() != number_of_objectives || number_of_constraints.size() != number_of_objectives)
            throw std::runtime_error("Wrong number of objectives.");

        // ---- objectives, in ascending order after the header ----------------------------------------------------
        const unsigned bounds = (hier_code == 100) ? 1u : 2u;
        const bool with_guess = hier_code == 210;
        objectives.resize(number_of_objectives);
        if (with_guess) active_set_guess.resize(number_of_objectives);
        unsigned k = 0;
        while (at < lines.size() && k < number_of_objectives)
        {
            if (squeeze(lines[at++]).compare(0, 10, "#OBJECTIVE") != 0) continue;
            unsigned cols;
            if (types_of_objectives[k] == LexLS::SIMPLE_BOUNDS_OBJECTIVE)
            {
                if (k != 0) throw std::runtime_error("Simple constraints are supported only in the first objective.");
                cols = 1 + bounds;
            }
            else
                cols = number_of_variables + bounds;
            const unsigned rows = number_of_constraints[k];
            objectives[k].resize(rows, cols);
            if (with_guess) active_set_guess[k].assign(rows, CTR_INACTIVE);
            for (unsigned r = 0; r < rows; r++)
            {
                if (at >= lines.size()) throw std::runtime_error("Not enough data.");
                const std::vector<double> val = numbers(lines[at++]);
                if (val.size() < cols) throw std::runtime_error("Not enough data.");
                for (unsigned c = 0; c < cols; c++) objectives[k](r, c) = val[c];
                if (with_guess && val.size() > cols)
                {
                    const unsigned t = static_cast<unsigned>(val[cols]);
                    if (t > static_cast<unsigned>(CTR_ACTIVE_EQ)) throw std::runtime_error("Unsupported constraint activation type.");
                    active_set_guess[k][r] = static_cast<ConstraintActivationType>(t);
                }
            }
            k++;
        }
        if (k != number_of_objectives) throw std::runtime_error("The number of objectives is lower than expected.");

        // ---- optional vectors -----------------------------------------------------------------------------------
        while (at < lines.size())
        {
            const std::string key = lines[at++];
            if (key == "#SolGuess") vector_after(lines, at, number_of_variables, solution_guess);
            else if (key == "#Solution") vector_after(lines, at, number_of_variables, solution);
        }
    }

private:
    static std::string squeeze(const std::string &s)
    {
        std::string r;
        for (size_t i = 0; i < s.size(); i++)
            if (!std::isspace(static_cast<unsigned char>(s[i]))) r.push_back(s[i]);
        return r;
    }
    static std::vector<double> numbers(const std::string &s)
    {
        std::vector<double> v;
        std::istringstream ss(s);
        for (double x; ss >> x;) v.push_back(x);
        return v;
    }
    static void vector_after(const std::vector<std::string> &lines, size_t &at, unsigned n, dVectorType &out)
    {
        std::vector<double> v;
        while (at < lines.size() && v.size() < n)
        {
            const std::vector<double> t = numbers(lines[at++]);
            v.insert(v.end(), t.begin(), t.end());
        }
        if (v.size() < n) throw std::runtime_error("Could not read a solution vector.");
        out.resize(n);
        for (unsigned i = 0; i < n; i++) out(i) = v[i];
    }
};
} // namespace tools
} // namespace LexLS
