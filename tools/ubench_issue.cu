// tools/ubench_issue.cu -- micro-benchmarks behind the FP64 roofline of DESIGN.md:
//   (1) does an FP64 instruction (2 issue cycles on its 16-lane pipe) leave the other issue slot free for
//       integer work, or does an FP64:INT mix cost N_fp64*2 + N_int cycles?
//   (2) dependent-issue latency of DADD/DMUL/DFMA, MUFU.RSQ64H, MUFU.RCP64H
//   (3) FP64 throughput as a function of resident warps per SM sub-partition and independent chains per warp
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue tools/ubench_issue.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int NI>
__global__ void k_mix(double *out, int *iout, int iters)
{
    double a[CH]; unsigned v[CH * (NI > 0 ? NI : 1)];
#pragma unroll
    for (int j = 0; j < CH; j++) a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
#pragma unroll
    for (int j = 0; j < CH * (NI > 0 ? NI : 1); j++) v[j] = threadIdx.x * 7 + j;
    const double m = 1.0000000001, c = 1e-12;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < CH; j++) {
            a[j] = __dmul_rn(a[j], m);
#pragma unroll
            for (int q = 0; q < NI; q++) v[j * NI + q] = (v[j * NI + q] ^ (v[j * NI + q] >> 3)) + 0x9e37u;   // LOP3/SHF + IADD (alu pipe)
            a[j] = __dadd_rn(a[j], c);
#pragma unroll
            for (int q = 0; q < NI; q++) v[j * NI + q] = (v[j * NI + q] & 0xfffffu) | (v[j * NI + q] << 5);
        }
    }
    double r = 0; unsigned s = 0;
#pragma unroll
    for (int j = 0; j < CH; j++) r += a[j];
#pragma unroll
    for (int j = 0; j < CH * (NI > 0 ? NI : 1); j++) s += v[j];
    if (r == 123.456) out[0] = r;
    if (s == 0x12345u) iout[0] = s;
}

template <int OP>
__global__ void k_lat(double *out, int iters, long long *cycles)
{
    double a = 1.0 + 1e-9 * threadIdx.x;
    const double m = 1.0000000001, c = 1e-12;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (OP == 0) a = __dadd_rn(a, c);
            if (OP == 1) a = __dmul_rn(a, m);
            if (OP == 2) a = __fma_rn(a, m, c);
            if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y; }
            if (OP == 4) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (a == 123.456) out[0] = a;
}

template <int CH, int NI>
static void run_mix(int warps_per_sm, const char *name)
{
    int dev, sms; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d; int *di; cudaMalloc(&d, 8); cudaMalloc(&di, 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = warps_per_sm * 32;     // one CTA per SM
    k_mix<CH, NI><<<sms, threads>>>(d, di, iters / 10);
    cudaEventRecord(e0);
    k_mix<CH, NI><<<sms, threads>>>(d, di, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double fp64_warp_instr = (double)iters * CH * 2 * warps_per_sm;   // per SM
    const double int_warp_instr = (double)iters * CH * 2 * NI * 2 * warps_per_sm;   // ~2 alu instr per statement
    const double cyc = ms * 1e-3 * khz * 1e3;
    printf("%-28s warps/SM=%2d  fp64 lanes/clk/SM = %5.1f   cycles per (fp64 instr per SMSP) = %.2f   int:fp64 = %d:1\n", name, warps_per_sm,
           fp64_warp_instr * 32 / cyc, cyc / (fp64_warp_instr / 4), 2 * NI);
    (void)int_warp_instr;
    cudaFree(d); cudaFree(di);
}

int main()
{
    long long *dc; double *d; cudaMalloc(&dc, 8); cudaMalloc(&d, 8);
    const char *names[] = {"DADD", "DMUL", "DFMA", "MUFU.RSQ64H", "MUFU.RCP64H"};
    for (int op = 0; op < 5; op++) {
        long long c;
        const int iters = 2000;
        if (op == 0) k_lat<0><<<1, 32>>>(d, iters, dc);
        if (op == 1) k_lat<1><<<1, 32>>>(d, iters, dc);
        if (op == 2) k_lat<2><<<1, 32>>>(d, iters, dc);
        if (op == 3) k_lat<3><<<1, 32>>>(d, iters, dc);
        if (op == 4) k_lat<4><<<1, 32>>>(d, iters, dc);
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("dependent-issue latency %-12s %.1f cycles\n", names[op], (double)c / (iters * 16.0));
    }
    for (int w : {4, 8, 16, 32}) {
        if (w == 4) { run_mix<1, 0>(w, "1 chain, fp64 only"); run_mix<3, 0>(w, "3 chains, fp64 only"); run_mix<8, 0>(w, "8 chains, fp64 only"); run_mix<8, 1>(w, "8 chains, +int 2:1"); run_mix<8, 2>(w, "8 chains, +int 4:1"); }
        if (w == 8) { run_mix<1, 0>(w, "1 chain, fp64 only"); run_mix<3, 0>(w, "3 chains, fp64 only"); run_mix<8, 0>(w, "8 chains, fp64 only"); run_mix<8, 1>(w, "8 chains, +int 2:1"); run_mix<8, 2>(w, "8 chains, +int 4:1"); }
        if (w == 16) { run_mix<3, 0>(w, "3 chains, fp64 only"); run_mix<8, 0>(w, "8 chains, fp64 only"); run_mix<8, 1>(w, "8 chains, +int 2:1"); run_mix<8, 2>(w, "8 chains, +int 4:1"); }
        if (w == 32) { run_mix<8, 0>(w, "8 chains, fp64 only"); run_mix<8, 1>(w, "8 chains, +int 2:1"); run_mix<8, 2>(w, "8 chains, +int 4:1"); }
    }
    return 0;
}
