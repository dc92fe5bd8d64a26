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
#ifndef MS_CTA_EMU
    asm("" : "+f"(num));
#endif
    const float q = num / b;
    return nz ? q : a;
}

}  // namespace ms
