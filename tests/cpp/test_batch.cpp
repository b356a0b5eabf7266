// The C++ batch entry point (include/lexls/batch.h): LexLS::solveBatch / LexLSIBatch / LexLSEBatch against the single-problem
// drop-in classes on the same inputs.  Both go through the same C-ABI and the same kernels, so the results must be identical
// bit for bit; the warm-started batch must need no more factorizations than the cold one.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <lexls/batch.h>

using namespace LexLS;

static double rnd()
{
    static unsigned long long s = 88172645463325252ull;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return double(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

int main()
{
    try
    {
        const uint64_t B = 37;
        // ---- LexLSI: 6 joint limits, 3 equalities, 4 inequalities, 5 equalities; n = 6
        const Index n = 6, P = 4;
        Index dims[P]          = {6, 3, 4, 5};
        ObjectiveType types[P] = {SIMPLE_BOUNDS_OBJECTIVE, GENERAL_OBJECTIVE, GENERAL_OBJECTIVE, GENERAL_OBJECTIVE};
        Index var_index[6]     = {0, 1, 2, 3, 4, 5};
        std::vector<std::vector<dMatrixType>> data(B, std::vector<dMatrixType>(P));
        for (uint64_t q = 0; q < B; q++)
            for (Index k = 0; k < P; k++)
            {
                const Index m = dims[k], cols = (k == 0) ? 2 : n + 2;
                dMatrixType M(m, cols);
                for (Index i = 0; i < m; i++)
                {
                    for (Index j = 0; j + 2 < cols; j++) M(i, j) = rnd();
                    const double ub = 0.6 + 0.4 * rnd(), w = (k == 2 || k == 0) ? 0.2 + 0.5 * (rnd() + 1.0) : 0.0;
                    M(i, cols - 1) = ub;
                    M(i, cols - 2) = ub - w;
                }
                data[q][k] = M;
            }
        ParametersLexLSI par;
        BatchResult r = solveBatch(n, P, dims, types, par, data, var_index);
        if (r.x.size() != B) { std::printf("FAIL batch size\n"); return 1; }
        unsigned fact_cold = 0;
        std::vector<std::vector<std::vector<ConstraintActivationType>>> act(P, std::vector<std::vector<ConstraintActivationType>>(B));
        for (uint64_t q = 0; q < B; q++)
        {
            internal::LexLSI one(n, P, dims, types);
            one.setParameters(par);
            one.setData(0, var_index, data[q][0]);
            for (Index k = 1; k < P; k++) one.setData(k, data[q][k]);
            const TerminationStatus st = one.solve();
            if (st != r.status[q]) { std::printf("FAIL status of problem %u\n", unsigned(q)); return 1; }
            for (Index j = 0; j < n; j++)
                if (one.get_x()(j) != r.x[q](j)) { std::printf("FAIL x of problem %u differs from the single-problem solve\n", unsigned(q)); return 1; }
            fact_cold += one.getFactorizationsCount();
            for (Index k = 0; k < P; k++) one.getActiveCtr(k, act[k][q]);
        }
        // warm start through the batch class: the final active sets and x of the cold solve as the guess
        {
            LexLSIBatch b(n, P, dims, types, B);
            b.setParameters(par);
            std::vector<dMatrixType> level(B);
            for (Index k = 0; k < P; k++)
            {
                for (uint64_t q = 0; q < B; q++) level[q] = data[q][k];
                if (k == 0) b.setData(0, var_index, level); else b.setData(k, level);
                b.setActiveGuess(k, act[k]);
            }
            b.set_x0(r.x);
            b.solve();
            unsigned fact_warm = 0;
            for (uint64_t q = 0; q < B; q++)
            {
                fact_warm += b.getFactorizationsCount(q);
                if (b.getStatus(q) != r.status[q]) { std::printf("FAIL warm status\n"); return 1; }
                std::vector<ConstraintActivationType> a;
                for (Index k = 0; k < P; k++)
                {
                    b.getActiveCtr(q, k, a);
                    if (a != act[k][q]) { std::printf("FAIL warm active set of problem %u level %u\n", unsigned(q), k); return 1; }
                }
                for (Index j = 0; j < n; j++)
                    if (std::fabs(b.get_x(q)(j) - r.x[q](j)) > 1e-9) { std::printf("FAIL warm x\n"); return 1; }
            }
            if (fact_warm > fact_cold || fact_warm != B) { std::printf("FAIL warm start needed %u factorizations (cold %u)\n", fact_warm, fact_cold); return 1; }
        }

        // ---- LexLSE: 3 levels of 2 / 3 / 4 rows, n = 7
        const Index ne = 7, Pe = 3;
        Index de[Pe] = {2, 3, 4};
        std::vector<std::vector<dMatrixType>> eq(B, std::vector<dMatrixType>(Pe));
        for (uint64_t q = 0; q < B; q++)
            for (Index k = 0; k < Pe; k++)
            {
                dMatrixType M(de[k], ne + 1);
                for (Index i = 0; i < de[k]; i++)
                    for (Index j = 0; j <= ne; j++) M(i, j) = rnd();
                eq[q][k] = M;
            }
        ParametersLexLSE pe;
        BatchResult re = solveBatch(ne, Pe, de, pe, eq);
        for (uint64_t q = 0; q < B; q++)
        {
            internal::LexLSE one(ne, Pe, de);
            for (Index k = 0; k < Pe; k++) one.setData(k, eq[q][k]);
            one.factorize();
            one.solve();
            for (Index j = 0; j < ne; j++)
                if (one.get_x()(j) != re.x[q](j)) { std::printf("FAIL LexLSE x of problem %u\n", unsigned(q)); return 1; }
        }
        // the same entry point over a device list (lexls_b200_create_batch_multi): one device, then every device of the box
        {
            int ndev = 0;
            lexls_b200_device_count(&ndev);
            for (int use = 1; use <= ndev; use = (use < ndev && 2 * use > ndev) ? ndev : 2 * use)
            {
                std::vector<int> devices;
                for (int d = 0; d < use; d++) devices.push_back(d);
                BatchResult rm = solveBatch(ne, Pe, de, pe, eq, devices);
                for (uint64_t q = 0; q < B; q++)
                    for (Index j = 0; j < ne; j++)
                        if (rm.x[q](j) != re.x[q](j)) { std::printf("FAIL multi-device LexLSE x (%d devices) of problem %u\n", use, unsigned(q)); return 1; }
                BatchResult ri = solveBatch(n, P, dims, types, par, data, var_index, devices);
                for (uint64_t q = 0; q < B; q++)
                {
                    if (ri.status[q] != r.status[q]) { std::printf("FAIL multi-device LexLSI status\n"); return 1; }
                    for (Index j = 0; j < n; j++)
                        if (ri.x[q](j) != r.x[q](j)) { std::printf("FAIL multi-device LexLSI x (%d devices) of problem %u\n", use, unsigned(q)); return 1; }
                }
                std::printf("multi-device solveBatch: %d device(s) ok\n", use);
                if (use == ndev) break;
            }
        }
        // host-resident inputs pipelined over chunk handles (asynchronous copies, deferred repacks): the same bits
        {
            std::vector<std::vector<double> > lvl(Pe);
            std::vector<uint64_t> stride(Pe);
            for (Index k = 0; k < Pe; k++)
            {
                stride[k] = uint64_t(de[k]) * (ne + 1);
                lvl[k].resize(B * stride[k]);
                for (uint64_t q = 0; q < B; q++) std::copy(eq[q][k].data(), eq[q][k].data() + stride[k], lvl[k].begin() + q * stride[k]);
            }
            PipelinedBatch<LexLSEBatch> pb(B, 3, [&](uint64_t count) { return new LexLSEBatch(ne, Pe, de, count); });
            std::vector<double> xp(B * ne);
            for (int round = 0; round < 3; round++) // the staging arena grows during the first rounds
            {
                std::fill(xp.begin(), xp.end(), 0.0);
                pb.solve([&](LexLSEBatch &c, uint64_t first, uint64_t) {
                    for (Index k = 0; k < Pe; k++) c.setDataPacked(k, lvl[k].data() + first * stride[k], de[k], stride[k]);
                }, xp.data(), ne);
                for (uint64_t q = 0; q < B; q++)
                    for (Index j = 0; j < ne; j++)
                        if (xp[q * ne + j] != re.x[q](j)) { std::printf("FAIL pipelined LexLSE x of problem %u (round %d)\n", unsigned(q), round); return 1; }
            }
            PipelinedBatch<LexLSIBatch> pi(B, 4, [&](uint64_t count) { return new LexLSIBatch(n, P, dims, types, count); });
            std::vector<double> xi(B * n);
            pi.solve([&](LexLSIBatch &c, uint64_t first, uint64_t count) {
                std::vector<dMatrixType> level(count);
                for (Index k = 0; k < P; k++)
                {
                    for (uint64_t q = 0; q < count; q++) level[q] = data[first + q][k];
                    if (k == 0) c.setData(0, var_index, level); else c.setData(k, level);
                }
            }, xi.data(), n);
            for (uint64_t q = 0; q < B; q++)
                for (Index j = 0; j < n; j++)
                    if (xi[q * n + j] != r.x[q](j)) { std::printf("FAIL pipelined LexLSI x of problem %u\n", unsigned(q)); return 1; }
            std::printf("PipelinedBatch ok\n");
        }
        // the handle classes: ranks, packed data, fixed variables, residuals
        {
            LexLSEBatch b(ne, Pe, de, B);
            std::vector<double> packed;
            for (Index k = 0; k < Pe; k++)
            {
                packed.assign(B * de[k] * (ne + 1), 0.0);
                for (uint64_t q = 0; q < B; q++)
                    for (Index j = 0; j <= ne; j++)
                        for (Index i = 0; i < de[k]; i++) packed[q * de[k] * (ne + 1) + j * de[k] + i] = eq[q][k](i, j);
                b.setDataPacked(k, packed.data(), de[k], uint64_t(de[k]) * (ne + 1));
            }
            std::vector<Index> fi(B, 3);         // x4 = 0.25 in every problem
            std::vector<double> fv(B, 0.25);
            b.fixVariables(1, fi.data(), fv.data());
            b.solve();
            for (uint64_t q = 0; q < B; q++)
            {
                internal::LexLSE one(ne, Pe, de);
                one.setFixedVariablesCount(1);
                one.fixVariable(3, 0.25);
                for (Index k = 0; k < Pe; k++) one.setData(k, eq[q][k]);
                one.factorize();
                one.solve();
                for (Index j = 0; j < ne; j++)
                    if (one.get_x()(j) != b.get_x(q)(j)) { std::printf("FAIL LexLSEBatch x (fixed variable) of problem %u\n", unsigned(q)); return 1; }
                for (Index k = 0; k < Pe; k++)
                    if (one.getRank(k) != b.getRank(q, k)) { std::printf("FAIL LexLSEBatch rank\n"); return 1; }
                if (b.get_x(q)(3) != 0.25) { std::printf("FAIL fixed value\n"); return 1; }
            }
        }
        {
            LexLSIBatch b(n, P, dims, types, B);
            std::vector<dMatrixType> level(B);
            for (Index k = 0; k < P; k++)
            {
                for (uint64_t q = 0; q < B; q++) level[q] = data[q][k];
                if (k == 0) b.setData(0, var_index, level); else b.setData(k, level);
            }
            b.solve();
            internal::LexLSI one(n, P, dims, types);
            one.setData(0, var_index, data[5][0]);
            for (Index k = 1; k < P; k++) one.setData(k, data[5][k]);
            one.solve();
            for (Index k = 0; k < P; k++)
            {
                const dVectorType vb = b.get_v(5, k);
                for (Index i = 0; i < dims[k]; i++)
                    if (vb(i) != one.get_v(k)(i)) { std::printf("FAIL LexLSIBatch get_v level %u\n", k); return 1; }
            }
            if (b.getTotalRank(5) != one.getTotalRank() || b.getActivationsCount(5) != one.getActivationsCount()) { std::printf("FAIL counters\n"); return 1; }
            b.reset();
            bool threw = false;
            try { b.get_x(0); } catch (Exception &) { threw = true; }
            if (!threw) { std::printf("FAIL get_x before solve must throw\n"); return 1; }
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
