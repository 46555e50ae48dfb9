// msm_plan.h — host-side planning of the multi-scalar multiplication behind combine / decrypt / lincomb (pure C++, no CUDA):
// how many partial sums ("groups") per item, and whether the G2 accumulation uses the spill layout.  Shared by tcb200.cu and the
// host emulation (tests/hostemu), so that the choices for the BASELINE shapes are pinned by a CPU test (tests/test_hostemu_parity.py).
#pragma once
#include <stddef.h>

namespace tcbk {

// G partial sums per item trade shared work (small G) against parallelism (large G): the G that minimises
//   waves(n G) * (fixed + per_share * ceil(m / G)).
inline size_t pick_groups(size_t n, size_t m, size_t units_per_wave, double fixed_cost, double share_cost) {
    size_t best = 1;
    double best_cost = 1e300;
    for (size_t G = 1; G <= m; G++) {
        size_t per = (m + G - 1) / G;
        if (G > 1 && (m + G - 2) / (G - 1) == per) continue;          // same depth as G - 1 with more units
        double waves = (double)((n * G + units_per_wave - 1) / units_per_wave);
        double cost = waves * (fixed_cost + share_cost * (double)per);
        if (cost < best_cost * 0.999) { best_cost = cost; best = G; }
    }
    return best;
}

// Spill layout (scheme.cuh: task_g2_msm_acc_spill) for a batch that runs one unit per item (G == 1) and leaves unit slots of the
// single wave idle: the last share of every item moves to the spare units, q items each.  Returns q, or 0 when the layout does not
// shorten the longest unit by at least 3 % (or does not apply).
inline size_t pick_spill(size_t n, size_t m, size_t units_per_wave, double fixed_cost, double share_cost) {
    if (m < 3 || n >= units_per_wave || n < units_per_wave / 2) return 0;
    size_t spare = units_per_wave - n;
    size_t q = (n + spare - 1) / spare;
    double now = fixed_cost + share_cost * (double)m;
    double main_unit = fixed_cost + share_cost * (double)(m - 1), spare_unit = (double)q * (fixed_cost + share_cost);
    return (main_unit > spare_unit ? main_unit : spare_unit) < 0.97 * now ? q : 0;
}

// relative costs in field multiplications of one digit position: a doubling (fixed) and one mixed addition per share
// (G1 reads two digit positions per look-up: two doublings per position); batch-affine variants in tcb200.cu
inline void straus_costs(bool g2, double &fixed_cost, double &share_cost) {
    fixed_cost = g2 ? 4.8 : 14.0;
    share_cost = g2 ? 8.6 : 11.0;
}

}  // namespace tcbk
