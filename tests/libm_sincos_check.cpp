// Host build of q1physrl_b200/csrc/q1_libm_sincos.cuh compared bit for bit with the installed libm.
// usage: libm_sincos_check <count> <threads>; prints "mismatches <m> of <n>" and exits 0 iff m == 0.
// Built by tests/test_libm_sincos.py with: g++ -O2 -std=c++17 -ffp-contract=off -mfma
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../q1physrl_b200/csrc/q1_libm_sincos.cuh"

static inline uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double unit(uint64_t &s) { return (double)(splitmix(s) >> 11) * 0x1p-53; }
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t to_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

struct Tally { uint64_t n = 0, bad = 0; double first_bad = 0; };

static void check(double x, Tally &t)
{
    double s, c;
    volatile double vx = x;   // keep the compiler from folding sin/cos of a constant
    if (!q1libm::sincos(x, s, c))
        return;
    const double rs = sin(vx), rc = cos(vx);
    t.n++;
    if (to_bits(s) != to_bits(rs) || to_bits(c) != to_bits(rc)) {
        if (!t.bad) t.first_bad = x;
        t.bad++;
    }
}

static void worker(uint64_t seed, uint64_t count, Tally *out)
{
    Tally t;
    uint64_t s = seed;
    const double thresholds[] = {0.126, 0.855469, 0.85546875, 2.426265, 2.4262650012969971,
                                 1.5707963267948966, 3.141592653589793, 105414350.0, 1.0 / 128,
                                 0.5 / 128, 109.5 / 128, 0x1p-26, 0x1p-27};
    for (uint64_t i = 0; i < count; i++) {
        double x;
        switch (i & 7) {
        case 0: x = (unit(s) * 2 - 1) * 3.0; break;                       // the direct / fold ranges
        case 1: x = (unit(s) * 2 - 1) * 130.0; break;                     // an episode's yaw range
        case 2: x = (unit(s) * 2 - 1) * 1.0e5; break;
        case 3: x = ldexp(unit(s) + 0.5, (int)(splitmix(s) % 1080) - 56) * ((splitmix(s) & 1) ? 1 : -1); break;
        case 4: {                                                          // near a threshold, +-2^20 ulps
            double th = thresholds[splitmix(s) % (sizeof thresholds / sizeof *thresholds)];
            x = from_bits(to_bits(th) + (int64_t)(splitmix(s) % 2097152) - 1048576);
            if (splitmix(s) & 1) x = -x;
            break; }
        case 5: {                                                          // near a multiple of pi/2
            double m = (double)(int64_t)(splitmix(s) % 200000) - 100000.0;
            x = m * 1.5707963267948966 + (unit(s) * 2 - 1) * ldexp(1.0, -(int)(splitmix(s) % 50));
            break; }
        case 6: {                                                          // near a table node k/128 + 1/256
            double m = (double)(splitmix(s) % 220) * (1.0 / 256);
            x = from_bits(to_bits(m + 1.0e-300) + (int64_t)(splitmix(s) % 4096) - 2048);
            break; }
        default: {                                                         // degrees -> radians as phys.py:58 does
            double yaw = (unit(s) * 2 - 1) * 8000.0;
            if (splitmix(s) & 1) yaw = (double)(float)yaw;
            x = (yaw * 3.14159265358979323846) / 180.0;
            break; }
        }
        check(x, t);
    }
    *out = t;
}

int main(int argc, char **argv)
{
    uint64_t count = argc > 1 ? strtoull(argv[1], 0, 10) : 10000000ull;
    int threads = argc > 2 ? atoi(argv[2]) : 4;
    std::vector<Tally> tallies(threads);
    std::vector<std::thread> pool;
    for (int i = 0; i < threads; i++)
        pool.emplace_back(worker, 0x1234567ull + 977ull * i, count / threads, &tallies[i]);
    for (auto &th : pool) th.join();
    Tally total;
    for (auto &t : tallies) {
        if (t.bad && !total.bad) total.first_bad = t.first_bad;
        total.n += t.n; total.bad += t.bad;
    }
    // every integer number of degrees an episode can start from, and the exact special values
    Tally sp;
    for (int d = -72000; d <= 72000; d++) check(((double)d * 3.14159265358979323846) / 180.0, sp);
    const double specials[] = {0.0, -0.0, 0x1p-1022, 0x1p-1074, 0x1p-27, 0x1p-26, 0.126, 0.855469, 1.5707963267948966,
                               105414336.0, 105414350.0, 1e9, 1e22, 0x1p1023, 1.7976931348623157e308};
    for (double v : specials) { check(v, sp); check(-v, sp); }
    total.n += sp.n; total.bad += sp.bad;
    if (sp.bad && total.bad == sp.bad) total.first_bad = sp.first_bad;
    printf("mismatches %llu of %llu", (unsigned long long)total.bad, (unsigned long long)total.n);
    if (total.bad) printf(" first %a", total.first_bad);
    printf("\n");
    return total.bad ? 1 : 0;
}
