/*
 * oracle/ref_philox.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11), restated from the published algorithm.  It stands in for UnityEngine.Random.value
 * (FFTMesh.cs:171-172), which is a closed-source engine service that is not in /root/reference
 * and cannot be reproduced -- RNG parity with Unity is unpinned.
 * Known-answer vectors from the Random123 distribution are checked in tests/test_oracle.py.
 */
#ifndef REF_PHILOX_H
#define REF_PHILOX_H
#include <stdint.h>

static inline void ref_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                     uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 24-bit uniform in (0, 1]: never 0, so Log(z1) stays finite (Random.value can return 0). */
static inline float ref_u32_to_unit_open0(uint32_t u)
{
    return (float)((u >> 8) + 1u) * (1.0f / 16777216.0f);
}

#endif
