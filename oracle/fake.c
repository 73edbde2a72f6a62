/* oracle/fake.c -- CPU ORACLE (test infrastructure only): hash-derived "network" for search-parity tests.
 *
 * The reference has no fake backend (no test ever runs a search); this one exists so that the
 * search itself can be compared bit-for-bit between the oracle and the CUDA engine without any float produced by
 * a transcendental function.  Every output is (9-bit integer) * 2^k, hence exact in fp32 on both sides.  The
 * product implements the same definition in crazyara_b200/csrc/search_kernels.cuh (fake_eval).
 */
#include "mcts.h"

static uint64_t zmix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void ofake_eval(unsigned long long key, int n_labels, float* value, float* prob) {
    const uint64_t h0 = zmix(key ^ 0x9E3779B97F4A7C15ULL);
    *value = (float)((int)((h0 >> 11) & 0xFFFF) - 32768) * (1.0f / 65536.0f);
    for (int i = 0; i < n_labels; ++i) {
        const uint64_t h = zmix(key + (uint64_t)(i + 1) * 0xD6E8FEB86659FD93ULL);
        int t = __builtin_ctz((uint32_t)h | 0x80u);      /* 0..7, geometric */
        const int m = (int)((h >> 40) & 0xFF);
        prob[i] = (float)(256 + m) * (1.0f / 65536.0f) * (float)(1 << t); /* (256+m) * 2^(t-16) */
    }
}
