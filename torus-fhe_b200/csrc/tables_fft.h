// tables_fft.h -- host-side generation of the twiddle tables of the FP64 FFT channel (fft64.cuh).
// Entries are complex doubles (re, im), computed in long double and rounded once.
#pragma once
#include <cmath>
#include <vector>
#include "fft64_layout.h"

namespace mkf {

inline unsigned brev_bits(unsigned x, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

struct HostTablesFFT {
    std::vector<double> tw;   // [T_ENTRIES][2]
    HostTablesFFT() : tw((size_t)T_ENTRIES * 2, 0.0) {
        const long double PI = 3.14159265358979323846264338327950288L;
        auto put = [&](int idx, long double ang) {
            tw[2 * (size_t)idx] = (double)cosl(ang);
            tw[2 * (size_t)idx + 1] = (double)sinl(ang);
        };
        // forward, Cooley-Tukey natural -> bit-reversed on C[X] / (X^512 - i): stage d (span 256 >> d), group g:
        //   s(d, g) = exp(i theta / 2), theta = pi / 2^(d+1) + 2 pi bitrev_d(g) / 2^d
        auto fwd_ang = [&](int d, int g) {
            const long double theta = PI / (long double)(1 << (d + 1)) + 2.0L * PI * (long double)brev_bits((unsigned)g, d) / (long double)(1 << d);
            return 0.5L * theta;
        };
        for (int d = 1; d <= 4; d++)
            for (int g = 0; g < (1 << d); g++) put(TF_A + (1 << d) - 2 + g, fwd_ang(d, g));
        for (int d = 5; d <= 8; d++)
            for (int lane = 0; lane < 32; lane++) put(TF_B + 32 * (d - 5) + lane, fwd_ang(d, lane * (1 << (d - 5))));
        // inverse, decimation in time, bit-reversed -> natural: span sp, position twiddle exp(-2 pi i (pos mod sp) / (2 sp));
        // pass A' covers sp = 16 rs, rs = 2^t, pos mod sp = (r mod rs) 16 + l16; the table holds the rows r mod rs = 0
        for (int t = 0; t < 4; t++)
            for (int e = 0; e < 16; e++) put(TI_A + 16 * t + e, -2.0L * PI * (long double)e / (long double)(32 << t));
        for (int j = 0; j < 256; j++) {
            put(T_WJ + j, -2.0L * PI * (long double)j / 512.0L);        // last inverse stage (span 256)
            put(T_UT + j, -PI * (long double)j / 1024.0L);              // untwist zeta^-j, zeta = exp(i pi / N)
        }
    }
};

}  // namespace mkf
