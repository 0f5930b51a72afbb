// philox.cuh -- counter-based Philox-4x32-10 (Salmon et al., SC'11) element keys.
// The occupancy generator replaces the reference's sequential srand/rand shuffle
// (Fortran/permute.f:39-44) by one 64-bit key per lattice element, recomputed wherever it is
// needed instead of being stored: element e of realization `stream` is occupied iff its key is
// among the k smallest (ties by element id).
#pragma once
#include <stdint.h>
#include "geometry.cuh"

namespace perc {

PERC_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

// One Philox call yields 128 bits = TWO 64-bit element keys, A = (c0, c1) and B = (c2, c3):
//   sites : call (type 0, counter = i >> 1)        -> site i takes half i & 1
//   bonds : call (type 1, counter = owner site i)  -> E bond = A, N bond = B
//           call (type 2, counter = owner site i)  -> NW bond = A, NE bond = B   (triangular, x even)
// (half as many Philox evaluations per lattice site as one call per element).  Ties between equal keys
// are broken by the element id: site i, bond dir * t + i.
PERC_HD void elem_key_pair(unsigned long long seed, unsigned long long stream, int type, unsigned long long counter,
                           unsigned long long& A, unsigned long long& B)
{
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t k0 = (uint32_t)seed;
    uint32_t k1 = (uint32_t)(seed >> 32) ^ (type == 0 ? 0u : type == 1 ? 0x5bd1e995u : 0x2545f491u);
    philox4x32_10(c, k0, k1);
    A = ((unsigned long long)c[0] << 32) | (unsigned long long)c[1];
    B = ((unsigned long long)c[2] << 32) | (unsigned long long)c[3];
}

PERC_HD unsigned long long site_key(unsigned long long seed, unsigned long long stream, unsigned long long i)
{
    unsigned long long A, B;
    elem_key_pair(seed, stream, 0, i >> 1, A, B);
    return (i & 1) ? B : A;
}

PERC_HD unsigned long long bond_key(unsigned long long seed, unsigned long long stream, int dir, unsigned long long i)
{
    unsigned long long A, B;
    elem_key_pair(seed, stream, dir < 2 ? 1 : 2, i, A, B);
    return (dir & 1) ? B : A;
}

}  // namespace perc
