// Diagnostics: issue-rate microbenchmarks for the pipes K1 and K2 are bound by (include/ocb_probe.h).
// One CTA of 1024 threads per SM runs `iters` rounds of 8 independent dependent chains of one opcode
// (inline PTX so ptxas keeps them); rate = lane-ops / SM clock cycles / SM, cycles from clock64().
#include "ocb_internal.cuh"

#include "../../include/ocb_probe.h"

#include <vector>

namespace ocb
{
namespace
{
enum ProbeOp
{
    P_POPC = 0,
    P_LOP3,
    P_IMAD,
    P_IADD3,
    P_MIX_POPC_LOP3, // 1 POPC : 4 LOP3, independent chains -- do XU and ALU overlap?
    P_MIX_POPC_LOP3_IMAD, // 1 POPC : 3 LOP3 : 1 IMAD
    P_VIMNMX,
    P_ISETP,
    P_DADD,
    P_DMUL,
    P_DFMA,
    P_DDIV,
    P_DSQRT,
    P_COUNT
};

template <int OP> __global__ void __launch_bounds__(1024, 1) probe_kernel(uint32_t iters, uint32_t seed, uint32_t *sink, long long *cycles)
{
    uint32_t x[8], y = seed | 1u, z = seed * 2654435761u;
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        x[i] = seed + threadIdx.x * 8 + i;
        d[i] = 1.0 + 1e-9 * (double)(x[i] & 1023);
    }
    const double dy = 1.0 + 1e-12 * (double)(seed & 7), dz = 1e-13;
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; it++)
    {
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                if (OP == P_POPC)
                    asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                else if (OP == P_LOP3)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
                else if (OP == P_IMAD)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
                else if (OP == P_IADD3)
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
                else if (OP == P_VIMNMX)
                    asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
                else if (OP == P_ISETP)
                    asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %2, p; }" : "+r"(x[i]) : "r"(y), "r"(z));
                else if (OP == P_MIX_POPC_LOP3)
                {
                    if (i < 2) // 2 POPC chains : 6 LOP3 chains ... ratio 1:3 per round; see host for accounting
                        asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                    else
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
                }
                else if (OP == P_MIX_POPC_LOP3_IMAD)
                {
                    if (i < 2)
                        asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                    else if (i < 6)
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
                    else
                        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
                }
                else if (OP == P_DADD)
                    asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dz));
                else if (OP == P_DMUL)
                    asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dy));
                else if (OP == P_DFMA)
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dy), "d"(dz));
                else if (OP == P_DDIV)
                    d[i] = __ddiv_rn(d[i], dy);
                else if (OP == P_DSQRT)
                    d[i] = __dsqrt_rn(d[i] + dy);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
    double dacc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        acc ^= x[i];
        dacc += d[i];
    }
    if (acc == 0x12345678u && dacc == 3.25)
        sink[0] = acc; // never true in practice; keeps the chains alive
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = t1 - t0;
}

template <int OP> int run_probe(int sms, uint32_t iters, uint32_t *d_sink, long long *d_cycles, double *rate_out, double *mhz_out)
{
    cudaEvent_t e0, e1;
    OCB_CUDA(cudaEventCreate(&e0));
    OCB_CUDA(cudaEventCreate(&e1));
    probe_kernel<OP><<<sms, 1024>>>(iters / 8 + 1, 12345u, d_sink, d_cycles); // warm-up
    OCB_CUDA(cudaEventRecord(e0));
    probe_kernel<OP><<<sms, 1024>>>(iters, 12345u, d_sink, d_cycles);
    OCB_CUDA(cudaEventRecord(e1));
    OCB_CUDA(cudaEventSynchronize(e1));
    count_launch(2);
    float ms = 0;
    OCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(sms);
    OCB_CUDA(cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : cyc)
        mx = c > mx ? c : mx;
    const double lane_ops = 1024.0 * (double)iters * 32.0; // per SM: 8 chains x 4 rounds per iteration
    *rate_out = lane_ops / (double)mx;
    *mhz_out = (double)mx / (ms * 1e-3) / 1e6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
} // namespace
} // namespace ocb

using namespace ocb;

extern "C" int ocb_probe_pipes(double *out, int n_out)
{
    if (!out || n_out < OCB_PROBE_COUNT)
        return fail_invalid("out must hold OCB_PROBE_COUNT doubles");
    int dev = 0;
    OCB_CUDA(cudaGetDevice(&dev));
    const int sms = sm_count(dev);
    uint32_t *d_sink = nullptr;
    long long *d_cycles = nullptr;
    OCB_CUDA(cudaMalloc(&d_sink, 64));
    OCB_CUDA(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    double mhz = 0, rate = 0;
    int rc = 0;
    const uint32_t it = 4096;
#define RUN(OP, SLOT, ITERS)                                                                                           \
    if (!rc)                                                                                                           \
    {                                                                                                                  \
        rc = run_probe<OP>(sms, ITERS, d_sink, d_cycles, &rate, &mhz);                                                 \
        out[SLOT] = rate;                                                                                              \
    }
    RUN(P_POPC, OCB_PROBE_POPC, it)
    RUN(P_LOP3, OCB_PROBE_LOP3, it)
    out[OCB_PROBE_SM_MHZ] = mhz;
    RUN(P_IMAD, OCB_PROBE_IMAD, it)
    RUN(P_IADD3, OCB_PROBE_IADD, it)
    RUN(P_VIMNMX, OCB_PROBE_IMNMX, it)
    RUN(P_ISETP, OCB_PROBE_ISETP_SEL, it)
    RUN(P_MIX_POPC_LOP3, OCB_PROBE_MIX_POPC_LOP3, it)
    RUN(P_MIX_POPC_LOP3_IMAD, OCB_PROBE_MIX_POPC_LOP3_IMAD, it)
    RUN(P_DADD, OCB_PROBE_DADD, it / 4)
    RUN(P_DMUL, OCB_PROBE_DMUL, it / 4)
    RUN(P_DFMA, OCB_PROBE_DFMA, it / 4)
    RUN(P_DDIV, OCB_PROBE_DDIV, it / 32)
    RUN(P_DSQRT, OCB_PROBE_DSQRT, it / 32)
#undef RUN
    out[OCB_PROBE_SMS] = sms;
    cudaFree(d_sink);
    cudaFree(d_cycles);
    return rc;
}
