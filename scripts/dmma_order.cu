// dmma_order.cu — which accumulation order does mma.sync f64 (DMMA) use on this part?
// The arithmetic contract of the l-QR kernels is "(S) sequential fma, one accumulator, ascending k".  Before the Gauss step
// may use DMMA for Left * Up we must know whether D = A*B + C is that chain (bit for bit), some other chain, or a fused
// multi-term sum.  Random operands with heavy cancellation make the candidates distinguishable.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -fmad=false -o scripts/bin/dmma_order scripts/dmma_order.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

template <int K> struct Cand
{
    static __device__ double asc(const double *a, const double *b, double c)
    {
        for (int k = 0; k < K; k++) c = fma(a[k], b[k], c);
        return c;
    }
    static __device__ double desc(const double *a, const double *b, double c)
    {
        for (int k = K; k--;) c = fma(a[k], b[k], c);
        return c;
    }
    // ascending chain inside blocks of 4, blocks ascending (what a k4 hardware atom chained K/4 times would give)
    static __device__ double pair(const double *a, const double *b, double c)
    {
        double s = 0.0;
        for (int k = 0; k < K; k += 2) s += fma(a[k], b[k], a[k + 1] * b[k + 1]);
        return s + c;
    }
    static __device__ double dot_then_c(const double *a, const double *b, double c)
    {
        double s = 0.0;
        for (int k = 0; k < K; k++) s = fma(a[k], b[k], s);
        return s + c;
    }
};

// one warp; trial t uses A[t] (M x K row-major), B[t] (K x 8, element (k,n) at k*8+n), C[t] (M x 8)
template <int M, int K> __global__ void probe(const double *A, const double *B, const double *C, int trials, unsigned long long *match)
{
    const int lane = threadIdx.x, g = lane >> 2, t4 = lane & 3;
    unsigned long long m_asc = 0, m_desc = 0, m_pair = 0, m_dot = 0, total = 0;
    for (int t = 0; t < trials; t++)
    {
        const double *a = A + (size_t)t * M * K, *b = B + (size_t)t * K * 8, *c = C + (size_t)t * M * 8;
        double d[4] = {0, 0, 0, 0};
        if constexpr (M == 8 && K == 4)
        {
            double a0 = a[g * K + t4], b0 = b[t4 * 8 + g];
            d[0] = c[g * 8 + 2 * t4];
            d[1] = c[g * 8 + 2 * t4 + 1];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a0), "d"(b0));
        }
        else if constexpr (M == 16 && K == 8)
        {
            double a0 = a[g * K + t4], a1 = a[(g + 8) * K + t4], a2 = a[g * K + t4 + 4], a3 = a[(g + 8) * K + t4 + 4];
            double b0 = b[t4 * 8 + g], b1 = b[(t4 + 4) * 8 + g];
            d[0] = c[g * 8 + 2 * t4];
            d[1] = c[g * 8 + 2 * t4 + 1];
            d[2] = c[(g + 8) * 8 + 2 * t4];
            d[3] = c[(g + 8) * 8 + 2 * t4 + 1];
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        }
        else if constexpr (M == 16 && K == 16)
        {
            double av[8], bv[4];
            for (int i = 0; i < 8; i++) av[i] = a[(g + 8 * (i & 1)) * K + t4 + 4 * (i >> 1)];
            for (int i = 0; i < 4; i++) bv[i] = b[(t4 + 4 * i) * 8 + g];
            d[0] = c[g * 8 + 2 * t4];
            d[1] = c[g * 8 + 2 * t4 + 1];
            d[2] = c[(g + 8) * 8 + 2 * t4];
            d[3] = c[(g + 8) * 8 + 2 * t4 + 1];
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                         : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(av[4]), "d"(av[5]), "d"(av[6]), "d"(av[7]), "d"(bv[0]),
                           "d"(bv[1]), "d"(bv[2]), "d"(bv[3]));
        }
        const int nout = (M == 8) ? 2 : 4;
        for (int i = 0; i < nout; i++)
        {
            const int row = g + 8 * (i >> 1), col = 2 * t4 + (i & 1);
            double ar[K], br[K];
            for (int k = 0; k < K; k++)
            {
                ar[k] = a[row * K + k];
                br[k] = b[k * 8 + col];
            }
            const double c0 = c[row * 8 + col];
            total++;
            m_asc += (__double_as_longlong(Cand<K>::asc(ar, br, c0)) == __double_as_longlong(d[i]));
            m_desc += (__double_as_longlong(Cand<K>::desc(ar, br, c0)) == __double_as_longlong(d[i]));
            m_pair += (__double_as_longlong(Cand<K>::pair(ar, br, c0)) == __double_as_longlong(d[i]));
            m_dot += (__double_as_longlong(Cand<K>::dot_then_c(ar, br, c0)) == __double_as_longlong(d[i]));
        }
    }
    atomicAdd(match + 0, total);
    atomicAdd(match + 1, m_asc);
    atomicAdd(match + 2, m_desc);
    atomicAdd(match + 3, m_pair);
    atomicAdd(match + 4, m_dot);
}

static double rnd()
{
    // wide dynamic range and mixed signs: cancellation makes the summation order visible
    double m = (double)rand() / RAND_MAX * 2.0 - 1.0;
    int e    = rand() % 9 - 4;
    return ldexp(m, e);
}

template <int M, int K> static void run(const char *name)
{
    const int trials = 4096;
    size_t na = (size_t)trials * M * K, nb = (size_t)trials * K * 8, nc = (size_t)trials * M * 8;
    double *hA = (double *)malloc(na * 8), *hB = (double *)malloc(nb * 8), *hC = (double *)malloc(nc * 8);
    for (size_t i = 0; i < na; i++) hA[i] = rnd();
    for (size_t i = 0; i < nb; i++) hB[i] = rnd();
    for (size_t i = 0; i < nc; i++) hC[i] = rnd();
    double *dA, *dB, *dC;
    unsigned long long *dm, hm[5];
    cudaMalloc(&dA, na * 8);
    cudaMalloc(&dB, nb * 8);
    cudaMalloc(&dC, nc * 8);
    cudaMalloc(&dm, 40);
    cudaMemset(dm, 0, 40);
    cudaMemcpy(dA, hA, na * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, nb * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, hC, nc * 8, cudaMemcpyHostToDevice);
    probe<M, K><<<1, 32>>>(dA, dB, dC, trials, dm);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hm, dm, 40, cudaMemcpyDeviceToHost);
    printf("{\"shape\": \"%s\", \"cuda\": \"%s\", \"outputs\": %llu, \"match_fma_chain_ascending_k\": %llu, \"match_fma_chain_descending_k\": %llu, "
           "\"match_pairwise\": %llu, \"match_dot_then_add_c\": %llu}\n",
           name, cudaGetErrorString(e), hm[0], hm[1], hm[2], hm[3], hm[4]);
}

int main()
{
    srand(12345);
    run<8, 4>("m8n8k4");
    run<16, 8>("m16n8k8");
    run<16, 16>("m16n8k16");
    return 0;
}
