// scopa_b200/csrc/ms_div.cuh -- one helper shared by the SDCFR kernels (ms_sdcfr.cu, ms_sd_train*.cuh, ms_sd_avgpol.cuh).
// Compiles as CUDA and, behind tests/emu/cta_emu.h, as host C++.
#pragma once

namespace ms {

// a / b in IEEE arithmetic where a may be exactly 0 (the result is then a itself).  A zero numerator sends the compiler's
// IEEE division into its special-case subroutine, a call the whole warp waits for (profiles/README.md R2.5); the lanes
// that hold a zero divide a stand-in instead, hidden from the optimiser behind an empty asm (it would otherwise fold the
// stand-in away and divide `a` again).  Host builds (tests/emu) just divide: same bits.
__device__ __forceinline__ float ms_div_or_zero(float a, float b) {
    const bool nz = a != 0.f;
    float num = nz ? a : 1.f;
#ifdef __CUDA_ARCH__
    asm("" : "+f"(num));
#endif
    const float q = num / b;
    return nz ? q : a;
}

// the same for fp64 (regret matching divides positive parts that are often exactly 0; the reference-semantics MCCFR kernels
// are one dependent chain per run, where a trip through the division's special-case subroutine is pure latency)
__device__ __forceinline__ double ms_ddiv_or_zero(double a, double b) {
    const bool nz = a != 0.0;
    double num = nz ? a : 1.0;
#ifdef __CUDA_ARCH__
    asm("" : "+d"(num));
    const double q = __ddiv_rn(num, b);
#else
    const double q = num / b;
#endif
    return nz ? q : a;
}

}  // namespace ms
