// Compiles against the drop-in headers exactly as a user of the reference would (#include <lexls/lexlsi.h>) and, on a GPU,
// checks two analytic known answers: the Escande hierarchy (x_k = 1/k; interfaces/matlab-octave/tests/lexlsi/lambda_test.m)
// and LexLSE {A1 x = b1; I x = 0} == minimum-norm solution (examples/example_lexlse.m:17-29) on a case with a closed form.
#include <cmath>
#include <cstdio>
#include <lexls/lexls.h>

using namespace LexLS;

int main()
{
    try
    {
        const Index n = 6;
        std::vector<Index> dims(n + 1, 1);
        std::vector<ObjectiveType> types(n + 1, GENERAL_OBJECTIVE);
        internal::LexLSI lsi(n, n + 1, dims.data(), types.data());
        ParametersLexLSI par;
        par.max_number_of_factorizations = 50;
        lsi.setParameters(par);
        for (Index k = 0; k <= n; k++)
        {
            dMatrixType row(1, n + 2);
            for (Index j = 0; j < n; j++) row(0, j) = (k == n) ? 1.0 : (j == k ? double(k + 1) : 0.0);
            row(0, n) = row(0, n + 1) = 1.0;
            lsi.setData(k, row);
        }
        TerminationStatus st = lsi.solve();
        if (st != PROBLEM_SOLVED) { std::printf("FAIL status %d\n", int(st)); return 1; }
        for (Index k = 0; k < n; k++)
            if (std::fabs(lsi.get_x()(k) - 1.0 / (k + 1)) > 1e-12) { std::printf("FAIL x[%u]\n", k); return 1; }
        if (lsi.getFactorizationsCount() != 1 || lsi.getTotalRank() != n) { std::printf("FAIL counters\n"); return 1; }
        std::vector<ConstraintActivationType> t;
        lsi.getActiveCtr(n, t);
        if (t[0] != CTR_ACTIVE_EQ) { std::printf("FAIL active type\n"); return 1; }
        // multipliers of the last objective (sum(x) = 1): lambda_test.m — objective k < n carries
        // -(1 - sum_j 1/j) / k in column n, the last objective's own entry is its residual sum_j 1/j - 1
        {
            std::vector<dMatrixType> lam;
            lsi.getLambda(lam);
            double hsum = 0.0;
            for (Index k = 1; k <= n; k++) hsum += 1.0 / k;
            const double r = hsum - 1.0;
            if (lam.size() != n + 1 || lam[0].rows() != 1 || lam[0].cols() != n + 1) { std::printf("FAIL lambda shape\n"); return 1; }
            if (std::fabs(lam[n](0, n) - r) > 1e-12) { std::printf("FAIL lambda own %g vs %g\n", lam[n](0, n), r); return 1; }
            for (Index k = 0; k < n; k++)
                if (std::fabs(std::fabs(lam[k](0, n)) - r / (k + 1)) > 1e-12) { std::printf("FAIL lambda[%u] %g\n", k, lam[k](0, n)); return 1; }
            if (lsi.get_lexqr().rows() != n + 1 || lsi.get_lexqr().cols() != n + 1) { std::printf("FAIL lexqr shape\n"); return 1; }
        }

        // LexLSE: x1 + x2 + x3 = 3 ; I x = 0  ->  minimum-norm solution (1,1,1); with x3 fixed to 2: (0.5, 0.5, 2)
        Index d2[2] = {1, 3};
        internal::LexLSE lse(3, 2, d2);
        dMatrixType a(1, 4), e(3, 4);
        a(0, 0) = a(0, 1) = a(0, 2) = 1.0; a(0, 3) = 3.0;
        e.setZero();
        e(0, 0) = e(1, 1) = e(2, 2) = 1.0;
        lse.setData(0, a);
        lse.setData(1, e);
        lse.factorize();
        lse.solve();
        for (Index k = 0; k < 3; k++)
            if (std::fabs(lse.get_x()(k) - 1.0) > 1e-12) { std::printf("FAIL lse x\n"); return 1; }
        if (lse.getRank(0) != 1 || lse.getRank(1) != 2 || lse.getTotalRank() != 3) { std::printf("FAIL ranks\n"); return 1; }
        // get_data() returns the constraints as given, get_lexqr() the factorization: a one-row level needs no reflector,
        // so its R row is the row itself; level 2 (I x = 0 after eliminating x1) has |R(1,1)| = ||(-1, 1, 0)|| = sqrt(2)
        if (lse.get_data()(0, 3) != 3.0 || lse.get_data()(3, 2) != 1.0) { std::printf("FAIL get_data\n"); return 1; }
        if (lse.get_lexqr()(0, 0) != 1.0 || lse.get_lexqr()(0, 3) != 3.0 ||
            std::fabs(std::fabs(lse.get_lexqr()(1, 1)) - std::sqrt(2.0)) > 1e-12)
        {
            std::printf("FAIL get_lexqr %g %g %g\n", lse.get_lexqr()(0, 0), lse.get_lexqr()(0, 3), lse.get_lexqr()(1, 1));
            return 1;
        }
        // setCtr: replace the first equation by 2 x1 = 2 -> x = (1, 0, 0)
        {
            dRowVectorType row(3);
            row(0) = 2.0; row(1) = 0.0; row(2) = 0.0;
            lse.setCtr(0, row, 2.0);
            lse.setCtrType(0, 0, CTR_ACTIVE_EQ);
            lse.factorize();
            lse.solve();
            if (std::fabs(lse.get_x()(0) - 1.0) > 1e-12 || std::fabs(lse.get_x()(1)) > 1e-12) { std::printf("FAIL setCtr\n"); return 1; }
        }
        lse.setObjDim(d2);
        lse.setFixedVariablesCount(1);
        lse.fixVariable(2, 2.0);
        lse.setData(0, a);
        lse.setData(1, e);
        lse.factorize();
        lse.solve();
        if (std::fabs(lse.get_x()(0) - 0.5) > 1e-12 || std::fabs(lse.get_x()(2) - 2.0) > 1e-12) { std::printf("FAIL fixed\n"); return 1; }
        // regularization and the least-norm solve through the facade: x1 + x2 + x3 = 3 alone (rank 1, two free variables):
        // solve() gives the basic solution (3, 0, 0) up to the pivot choice, solveLeastNorm the minimum-norm one (1, 1, 1);
        // R_NO_Z damping with factor 1 halves the reduced right-hand side: the basic solution becomes (1.5, 0, 0)
        {
            Index d1[1] = {1};
            internal::LexLSE one(3, 1, d1);
            one.setData(0, a);
            one.factorize();
            one.solve();
            double nb = 0.0;
            for (Index k = 0; k < 3; k++) nb += one.get_x()(k) * one.get_x()(k);
            if (std::fabs(nb - 9.0) > 1e-9) { std::printf("FAIL basic solution norm %g\n", nb); return 1; }
            one.solveLeastNorm_2();
            for (Index k = 0; k < 3; k++)
                if (std::fabs(one.get_x()(k) - 1.0) > 1e-12) { std::printf("FAIL least norm x[%u] = %g\n", k, one.get_x()(k)); return 1; }
            internal::LexLSE reg(3, 1, d1);
            ParametersLexLSE pr;
            pr.regularization_type = REGULARIZATION_R_NO_Z;
            reg.setParameters(pr);
            reg.setRegularizationFactor(0, 1.0);
            reg.setData(0, a);
            reg.factorize();
            reg.solve();
            // R = |(1,1,1)| up to sign: the damped right-hand side is R (R'R + 1)^-1 R' b~ = b~ * R^2 / (R^2 + 1) with R^2 = 1 (one pivot)
            double s = 0.0;
            for (Index k = 0; k < 3; k++) s += reg.get_x()(k);
            if (std::fabs(s - 1.5) > 1e-9) { std::printf("FAIL regularized basic solution sum %g\n", s); return 1; }
            // the conjugate-gradient variant solves the same damped problem as TIKHONOV (lexlse.h:2256-2279 vs :317-331)
            double xs[2][3];
            for (int v = 0; v < 2; v++)
            {
                internal::LexLSE t(3, 1, d1);
                ParametersLexLSE pt;
                pt.regularization_type         = v ? REGULARIZATION_TIKHONOV_CG : REGULARIZATION_TIKHONOV;
                pt.max_number_of_CG_iterations = 50;
                t.setParameters(pt);
                t.setRegularizationFactor(0, 0.5);
                t.setData(0, a);
                t.factorize();
                t.solve();
                for (Index k = 0; k < 3; k++) xs[v][k] = t.get_x()(k);
            }
            for (Index k = 0; k < 3; k++)
                if (std::fabs(xs[0][k] - xs[1][k]) > 1e-9) { std::printf("FAIL TIKHONOV_CG vs TIKHONOV x[%u]: %g %g\n", k, xs[1][k], xs[0][k]); return 1; }
        }
        std::printf("OK\n");
        return 0;
    }
    catch (std::exception &e)
    {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
}
