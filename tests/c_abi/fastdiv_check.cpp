// Host-side check of om::make_fastdiv (csrc/common.cuh): the device code computes umulhi(n, mul) >> shift (or n when d == 1).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>

#include "../../orienmask_b200/csrc/common.cuh"

static int fdiv_host(int n, const FastDiv& f) {
    const int q = (int)((((uint64_t)(uint32_t)n * f.mul) >> 32) >> f.shift);
    return f.one ? n : q;
}

int main() {
    long bad = 0;
    for (int d = 1; d <= 70000; ++d) {
        const FastDiv f = om::make_fastdiv(d);
        for (int n : {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffff, 0x7ffffffe, 0x40000000, 123456789})
            if (n >= 0 && fdiv_host(n, f) != n / d) ++bad;
        for (int k = 0; k < 100; ++k) {
            const int n = (int)((((unsigned)rand() << 16) ^ (unsigned)rand()) & 0x7fffffff);
            if (fdiv_host(n, f) != n / d) ++bad;
            const long long m = (long long)(n % 30000) * d + ((k & 1) ? d - 1 : 0);       // multiples of d and the value just below
            if (m < 0x7fffffffLL && fdiv_host((int)m, f) != (int)(m / d)) ++bad;
        }
    }
    for (int d : {1 << 20, (1 << 20) + 1, 9437184, 2367488, 0x7fffffff, 0x40000001})
        for (int n : {0, d - 1, d, 0x7fffffff})
            if (fdiv_host(n, om::make_fastdiv(d)) != n / d) ++bad;
    printf("bad %ld\n", bad);
    return bad != 0;
}
