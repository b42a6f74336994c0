// host-side check of the integer model behind csrc/seqsum.cuh (no GPU needed): the map of a term and the composition of maps
// must reproduce IEEE float addition exactly while the sum stays in its binade.  Built and run by tests/test_seqsum.py.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include "../../lsd_slam_b200/csrc/seqsum.cuh"

static uint32_t bitsOf(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float floatOf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 16); }

int main()
{
    long checked = 0, ties = 0;
    // 1. single terms: every binade 0..23 of the sum, terms from far below the ulp to far above, forced ties included
    for (int k = -3; k < 28; k++)
        for (int it = 0; it < 200000; it++) {
            const uint32_t S = 0x800000u | (rnd() & 0x7fffffu);
            const float s = floatOf(((uint32_t)(k + 127) << 23) | (S & 0x7fffffu));
            int ex = k - 30 + (int)(rnd() % 34);                       // exponent of the term
            if (ex < -126) ex = -126;
            uint32_t xm = rnd() & 0x7fffffu;
            const int mode = rnd() % 4;
            if (mode == 0) xm &= ~((1u << (rnd() % 23)) - 1u);          // short mantissas: make ties likely
            if (mode == 1) xm = 0;
            const float x = floatOf(((uint32_t)(ex + 127) << 23) | xm);
            volatile float r = s + x;
            const SeqMap m = seqElement(bitsOf(x), k);
            const long Sout = (long)S + ((S & 1) ? m.d1 : m.d0);
            if (Sout < (1 << 24)) {
                const float mine = floatOf(((uint32_t)(k + 127) << 23) | ((uint32_t)Sout & 0x7fffffu));
                if (bitsOf(mine) != bitsOf(r)) { printf("FAIL term k=%d s=%a x=%a got %a want %a\n", k, s, x, mine, (float)r); return 1; }
                if (m.d0 != m.d1) ties++;
            } else if (!((float)r >= floatOf((uint32_t)(k + 128) << 23))) {
                printf("FAIL crossing k=%d s=%a x=%a: model left the binade, the float add did not (%a)\n", k, s, x, (float)r); return 1;
            }
            checked++;
        }
    // 2. runs: compose up to 2048 maps and compare with the sequential sum as long as it stays in the binade
    for (int it = 0; it < 20000; it++) {
        const int k = rnd() % 24;
        const uint32_t S0 = 0x800000u | (rnd() & 0x3fffffu);
        float s = floatOf(((uint32_t)(k + 127) << 23) | (S0 & 0x7fffffu));
        SeqMap acc = seqIdentity();
        const int n = 1 + rnd() % 2048;
        const int exBase = k - 23 - (int)(rnd() % 4) + (int)(rnd() % 12) - 6;
        for (int i = 0; i < n; i++) {
            int ex = exBase + (int)(rnd() % 3);
            if (ex < -126) ex = -126;
            uint32_t xm = rnd() & 0x7fffffu;
            if (rnd() & 1) xm &= 0x7f0000u;
            const float x = floatOf(((uint32_t)(ex + 127) << 23) | xm);
            volatile float r = s + x;
            if (!((float)r < floatOf((uint32_t)(k + 128) << 23))) break;
            s = r;
            acc = seqCompose(acc, seqElement(bitsOf(x), k));
            const long Sout = (long)S0 + ((S0 & 1) ? acc.d1 : acc.d0);
            const float mine = floatOf(((uint32_t)(k + 127) << 23) | ((uint32_t)Sout & 0x7fffffu));
            if (Sout >= (1 << 24) || bitsOf(mine) != bitsOf(s)) { printf("FAIL run k=%d i=%d got %a want %a\n", k, i, mine, s); return 1; }
            checked++;
        }
    }
    printf("OK %ld additions checked, %ld ties\n", checked, ties);
    return 0;
}
