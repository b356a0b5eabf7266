// Replays the call sequence of the reference's own C++ caller tests/test_numerical_error.cpp:52-137 through the drop-in
// headers on the reference's fixture tests/test_01.dat:
//   import the hierarchy (lexls/tools.h), LexLSI::setParameters / setData / solve, getActiveCtrCount, get_xStar, get_data,
//   get_lexqr, getActiveCtr_order, then a LexLSE with the inequality problem's dimensions (INITIALIZE_INEQUALITY branch:
//   LexLSE(nVar, nObj, maxObjDim) + setObjDim) fed with get_data() blocks and the constraint types, factorize, solve, and the
//   three error norms the reference prints.
// The reference documents (test_numerical_error.cpp:5-22) that with Eigen's vectorised reductions those errors are NOT zero.
// Here both solvers run the same fixed summation orders, so the last equality-constrained problem of LexLSI is reproduced
// by LexLSE bit for bit: all three errors must be exactly 0.
// One deviation the fixture forces: its first objective is SIMPLE_BOUNDS (rows = var_index lb ub), which the reference's
// sequence (written for general objectives) would reject at lexlsi.h:358; it goes through setData(ObjIndex, VarIndex, data),
// and its active bounds become LexLSE::fixVariable calls — exactly what LexLSI::formLexLSE does (objective.h:434-462).
#include <cmath>
#include <cstdio>
#include <lexls/lexlsi.h>
#include <lexls/tools.h>

using namespace LexLS;

static double block_diff_norm(const dMatrixType &a, const dMatrixType &b, Index r0, Index rows, Index cols)
{
    double s = 0.0;
    for (Index c = 0; c < cols; c++)
        for (Index r = 0; r < rows; r++)
        {
            const double d = a(r0 + r, c) - b(r0 + r, c);
            s += d * d;
        }
    return std::sqrt(s);
}

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        std::printf("Usage: %s hierarchy.dat\n", argv[0]);
        return 1;
    }
    try
    {
        tools::HierarchyType type_of_hierarchy;
        Index number_of_variables, number_of_objectives;
        std::vector<Index> number_of_constraints;
        std::vector<ObjectiveType> types_of_objectives;
        std::vector<dMatrixType> objectives;
        std::vector<std::vector<ConstraintActivationType> > active_set_guess;
        dVectorType solution_guess, solution;

        tools::HierarchyFileProcessor fprocessor;
        fprocessor.import(argv[1], type_of_hierarchy, number_of_variables, number_of_objectives, number_of_constraints,
                          types_of_objectives, objectives, active_set_guess, solution_guess, solution);
        if (type_of_hierarchy != tools::HIERARCHY_TYPE_INEQUALITY || number_of_objectives < 2) { std::printf("FAIL import\n"); return 1; }
        const Index n = number_of_variables, P = number_of_objectives;

        internal::LexLSI lsi(n, P, &number_of_constraints[0], &types_of_objectives[0]);
        ParametersLexLSI parameters;
        parameters.max_number_of_factorizations = 113;
        lsi.setParameters(parameters);

        const bool simple0 = types_of_objectives[0] == SIMPLE_BOUNDS_OBJECTIVE;
        std::vector<Index> var_index;
        dMatrixType bounds0;
        for (Index i = 0; i < P; ++i)
        {
            if (i == 0 && simple0)
            {
                const Index m = number_of_constraints[0];
                var_index.resize(m);
                bounds0.resize(m, 2);
                for (Index r = 0; r < m; r++)
                {
                    var_index[r]  = static_cast<Index>(objectives[0](r, 0)) - 1; // 1-based in the file (lexlsi.cpp:412)
                    bounds0(r, 0) = objectives[0](r, 1);
                    bounds0(r, 1) = objectives[0](r, 2);
                }
                lsi.setData(0, &var_index[0], bounds0);
            }
            else
                lsi.setData(i, objectives[i]);
        }
        const TerminationStatus st = lsi.solve();
        if (st != PROBLEM_SOLVED) { std::printf("FAIL status %d\n", int(st)); return 1; }
        if (solution.size() == n)
        {
            double e = 0.0, s = 0.0;
            for (Index k = 0; k < n; k++)
            {
                e += (lsi.get_x()(k) - solution(k)) * (lsi.get_x()(k) - solution(k));
                s += solution(k) * solution(k);
            }
            std::printf("x vs #Solution: %.3e (relative)\n", std::sqrt(e / s));
            if (std::sqrt(e / s) > 1e-9) { std::printf("FAIL solution\n"); return 1; }
        }

        Index nCtr = 0;
        std::vector<Index> eq_obj_dim(P);
        for (Index k = 0; k < P; k++)
        {
            eq_obj_dim[k] = lsi.getActiveCtrCount(k);
            if (!(k == 0 && simple0)) nCtr += eq_obj_dim[k];
            std::printf("%u ", eq_obj_dim[k]);
        }
        std::printf("\n");

        // ==============================================================
        const dVectorType xStar = lsi.get_xStar();
        const dMatrixType data  = lsi.get_data(); // call after lsi.get_xStar()
        const dMatrixType lexqr = lsi.get_lexqr();
        std::vector<ConstraintIdentifier> ctr;
        lsi.getActiveCtr_order(ctr);

        // ==============================================================
        const Index off = simple0 ? 1 : 0, G = P - off;
        internal::LexLSE lse(n, G, &number_of_constraints[off]); // INITIALIZE_INEQUALITY: sized for the inequality problem ...
        lse.setObjDim(&eq_obj_dim[off]);                         // ... and used with the active dimensions
        size_t k = 0;
        if (simple0)
        {
            lse.setFixedVariablesCount(eq_obj_dim[0]);
            for (Index j = 0; j < eq_obj_dim[0]; ++j, ++k)
            {
                if (ctr[k].obj_index != 0) { std::printf("FAIL order\n"); return 1; }
                const Index c = ctr[k].ctr_index;
                lse.fixVariable(var_index[c], ctr[k].ctr_type == CTR_ACTIVE_LB ? bounds0(c, 0) : bounds0(c, 1), ctr[k].ctr_type);
            }
        }
        Index row_ind = 0;
        for (Index i = 0; i < G; ++i)
        {
#ifdef LEXLS_B200_HAVE_EIGEN
            lse.setData(i, data.block(row_ind, 0, eq_obj_dim[off + i], n + 1));
#else
            lse.setData(i, dMatrixConstRef(data.data() + row_ind, eq_obj_dim[off + i], n + 1, data.rows()));
#endif
            for (Index j = 0; j < eq_obj_dim[off + i]; ++j, ++k) lse.setCtrType(i, j, ctr[k].ctr_type);
            row_ind += eq_obj_dim[off + i];
        }
        if (k != ctr.size()) { std::printf("FAIL working-set size %zu vs %zu\n", k, ctr.size()); return 1; }
        lse.factorize();
        lse.solve();

        // ==============================================================
        const dMatrixType &data1  = lse.get_data();
        const dMatrixType &lexqr1 = lse.get_lexqr();
        double ex = 0.0;
        for (Index j = 0; j < n; j++) ex += (lse.get_x()(j) - xStar(j)) * (lse.get_x()(j) - xStar(j));
        const double e1 = block_diff_norm(data, data1, 0, nCtr, n + 1), e2 = block_diff_norm(lexqr, lexqr1, 0, nCtr, n + 1);
        std::printf("nCtr = %u\n", nCtr);
        std::printf("error(x)     = %1.18e \n", std::sqrt(ex));
        std::printf("error(data)  = %1.18e \n", e1);
        std::printf("error(lexqr) = %1.18e \n", e2);
        row_ind = 0;
        for (Index i = 0; i < G; ++i)
        {
            std::printf("   obj(%u): %1.18e \n", i + off, block_diff_norm(lexqr, lexqr1, row_ind, eq_obj_dim[off + i], n + 1));
            row_ind += eq_obj_dim[off + i];
        }
        double dx = 0.0;
        for (Index j = 0; j < n; j++) dx += (lsi.get_x()(j) - xStar(j)) * (lsi.get_x()(j) - xStar(j));
        std::printf("|x - xStar|  = %1.3e \n", std::sqrt(dx));
        if (ex != 0.0 || e1 != 0.0 || e2 != 0.0) { std::printf("FAIL: LexLSE does not reproduce the last LexLSI problem bit for bit\n"); return 1; }
        if (std::sqrt(dx) > 1e-9) { std::printf("FAIL: a solved problem must end on xStar\n"); return 1; }
        std::printf("OK\n");
        return 0;
    }
    catch (std::exception &e)
    {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
}
