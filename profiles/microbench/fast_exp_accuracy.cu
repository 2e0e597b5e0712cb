#include <cstdio>
#include <cmath>
#include "/root/repo/riemannhamiltonianmontecarlo_b200/csrc/metric_kernel.cuh"
using namespace rmhmc;
__global__ void k(const double* x, double* e, double* q, int n) {
    __shared__ double tab[256];
    tab[threadIdx.x] = exp_table_entry(threadIdx.x);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) { e[i] = fast_exp_nonpos(x[i], tab); q[i] = fast_rcp_1to2(1.0 + e[i]); }
}
int main() {
    const int n = 1 << 16;
    double *x, *e, *q; cudaMallocManaged(&x, n*8); cudaMallocManaged(&e, n*8); cudaMallocManaged(&q, n*8);
    for (int i = 0; i < n; ++i) x[i] = -760.0 * (double)i / n * ((i % 7) ? 0.05 : 1.0);
    x[1] = -1e-300; x[2] = -0.0; x[3] = -709.0; x[4] = -745.0; x[5] = -800.0; x[6] = NAN;
    k<<<1,256>>>(x, e, q, n); cudaDeviceSynchronize();
    double worst = 0, worstq = 0;
    for (int i = 0; i < n; ++i) {
        double ref = exp(x[i]);
        if (x[i] != x[i]) { printf("nan in -> %g\n", e[i]); continue; }
        if (ref < 1e-300) { if (e[i] != 0.0 && fabs(e[i]-ref) > 1e-300) printf("tiny mismatch x=%g got %g ref %g\n", x[i], e[i], ref); continue; }
        double rel = fabs(e[i] - ref) / ref; if (rel > worst) worst = rel;
        double rq = fabs(q[i] - 1.0/(1.0+e[i])) * (1.0 + e[i]); if (rq > worstq) worstq = rq;
    }
    printf("max rel err exp %.3e  rcp %.3e  (eps=2.2e-16)\n", worst, worstq);
    return 0;
}
