/* ocb_probe.h -- diagnostics exported by libocb.so next to the C ABI of ocb.h: issue-rate microbenchmarks
 * that calibrate the roofline denominators (lane-operations per clock per SM) on the device in use.
 * Not part of the drop-in boundary. */
#ifndef OCB_PROBE_H
#define OCB_PROBE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C"
{
#endif
    enum ocb_probe_slot
    {
        OCB_PROBE_SMS = 0,
        OCB_PROBE_SM_MHZ,  /* SM clock held during the LOP3 probe (cycles / event time) */
        OCB_PROBE_POPC,    /* everything below: lane-ops / clk / SM */
        OCB_PROBE_LOP3,
        OCB_PROBE_IMAD,
        OCB_PROBE_IADD,
        OCB_PROBE_IMNMX,
        OCB_PROBE_ISETP_SEL,          /* setp+selp pairs */
        OCB_PROBE_MIX_POPC_LOP3,      /* 2 POPC chains + 6 LOP3 chains: total ops/clk/SM */
        OCB_PROBE_MIX_POPC_LOP3_IMAD, /* 2 POPC + 4 LOP3 + 2 IMAD chains */
        OCB_PROBE_DADD,
        OCB_PROBE_DMUL,
        OCB_PROBE_DFMA,
        OCB_PROBE_DDIV,  /* IEEE divisions */
        OCB_PROBE_DSQRT, /* IEEE square roots (+1 add) */
        OCB_PROBE_DFMA3, /* DFMA with three distinct register operands (register-file bandwidth bound) */
        OCB_PROBE_COUNT
    };
    /* Runs the probes on the current device; out must hold OCB_PROBE_COUNT doubles. 0 or a negative code. */
    int ocb_probe_pipes(double *out, int n_out);

    /* Checks the shared-reciprocal division and the straight-line square root that K2 uses (csrc/exact_math.cuh)
     * against div.rn.f64 / sqrt.rn.f64 on the device, on n pseudo-random operand triples whose exponents spread over
     * 2^-exponent_spread .. 2^+exponent_spread. counts5 = {divisions compared, divisions that differ, square roots
     * compared, square roots that differ, divisions outside the guarded range (not compared)}. 0 or a negative code. */
    int ocb_probe_exact_math(uint64_t seed, uint64_t n, uint32_t exponent_spread, uint64_t *counts5);

    /* Tensor pipe: average cycles per tcgen05.mma.kind::i8 (M 128 x N n x K 32, one CTA per SM on `ctas` SMs) when one
     * thread issues rounds x chains of them back to back, rotating over `chains` independent accumulators of n columns,
     * with A read from shared memory (a_in_tmem = 0) or from tensor memory (1). The arithmetic alone takes n / 2 cycles.
     * chains * n <= 512 (384 with A in tensor memory). 0 or a negative code. */
    int ocb_probe_umma(int a_in_tmem, int n, int chains, int rounds, int ctas, double *cycles_per_mma);

#ifdef __cplusplus
}
#endif
#endif
