// tables.h -- host-side generation of the general-twiddle tables of the
// 1024-point negacyclic Goldilocks NTT (see ntt32.cuh / ntt1024.cuh).
#pragma once
#include <vector>
#include "ntt32.cuh"

namespace ntt {

// primitive 2048th root of unity psi with psi^32 = 2^3
inline u64 find_psi() {
    u64 rho = gl::pow(7, (gl::P - 1) / 2048);  // 7 generates F_p^*
    u64 c = gl::pow(rho, 32);                  // primitive 64th root = 8^m, m odd
    for (u64 e = 1; e < 64; e += 2)
        if (gl::pow(c, e) == 8) return gl::pow(rho, e);
    return 0;
}

struct Tables {
    u64 psi, psi_inv;
    std::vector<u64> tw_fwd;  // [32 r][32 i0]
    std::vector<u64> tw_inv;  // [32 j][32 lane]
    Tables() : tw_fwd(1024), tw_inv(1024) {
        psi = find_psi();
        psi_inv = gl::pow(psi, gl::P - 2);
        u64 n_inv = gl::pow(1024, gl::P - 2);
        for (int r = 0; r < 32; r++)
            for (int i0 = 0; i0 < 32; i0++)
                tw_fwd[r * 32 + i0] = gl::pow(psi, (u64)i0 * (2 * brev5(r) + 1));
        for (int j = 0; j < 32; j++)
            for (int lane = 0; lane < 32; lane++)
                tw_inv[j * 32 + lane] = gl::mul(gl::pow(psi_inv, (u64)j * (2 * brev5(lane) + 1)), n_inv);
    }
};

}  // namespace ntt
