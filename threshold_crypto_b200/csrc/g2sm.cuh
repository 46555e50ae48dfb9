// g2sm.cuh — G2 point arithmetic on shared-memory cells for the lane-PAIR kernels (device only): the accumulation loop of the
// multi-scalar multiplication behind combine_signatures / the G2 linear combinations.
//
// Why: with one (item, group) unit per lane pair, config #3 (2^14 items, 11 shares) is 1024 warps on 592 schedulers; at the 206
// registers of the register engine (2 blocks/SM) pick_groups cannot split items into more units without a second wave, so the
// kernel ran with <= 2 warps per scheduler and the multiply pipe ~60 % busy.  Here the running point, the table entry being added
// and four scratch values live in cells (quadsm.cuh: one Fp per thread and slot, 128-bit LDS/STS, the partner's half is the
// neighbouring column), products are 2-term dots streamed from shared memory (q_mul2) and nothing bigger than an Fp crosses a
// function boundary: 9 slots = 54 KB per 128-thread block and < 128 registers -> 4 blocks/SM, so the same batch runs as twice as many
// units (G = 2) in one wave.  Same formulas as tower.cuh (dbl-2009-l, madd-2007-bl): the Jacobian partial sums are bit-identical.
//
// Measured (profiles/kbench_r2m_*, r2n_*): 23.0 ms against the register engine's 19.3 ms at 2^14 items — the multiply pipe is 70 % busy in
// both, and G = 2 with the uniform step count executes 17 % more multiply-accumulates — and 4.2 against 4.6 ms at 2048 items.  Hence
// selectable (tcb_set_msm_algo 4) but not the default.
#pragma once
#include "quadsm.cuh"

#if defined(__CUDACC__)
namespace tcb {

enum { G_AX = 0, G_AY, G_AZ, G_BX, G_BY, G_S0, G_S1, G_S2, G_S3, G_NSLOT };
constexpr size_t G_SMEM_BYTES = (size_t)G_NSLOT * 3 * QNT * 16;
TCB_D u32 g_me() { return threadIdx.x & ~1u; }          // the "re" column of my lane pair

// acc <- 2 acc   (dbl-2009-l, a = 0; infinity stays infinity: Z3 = 2 Y Z)
static __device__ __noinline__ void g_dbl() {
    const u32 t = q_tid(), me = g_me();
    Fp A = q_sqr(G_AX), B = q_sqr(G_AY);
    Fp yz = q_mul2(q_cell(G_AY, me), q_cell(G_AZ, me));
    Fp E = dbl(A) + A;
    q_st(G_S0, B); q_st(G_S1, q_ld(G_AX, t) + B); q_st(G_S2, E);
    __syncwarp();
    Fp C = q_sqr(G_S0), XB2 = q_sqr(G_S1), F = q_sqr(G_S2);
    Fp D = dbl(XB2 - A - C);
    Fp X3 = F - dbl(D);
    q_st(G_S3, D - X3);
    __syncwarp();
    Fp Y3 = q_mul2(q_cell(G_S2, me), q_cell(G_S3, me)) - dbl(dbl(dbl(C)));
    q_st(G_AX, X3); q_st(G_AY, Y3); q_st(G_AZ, dbl(yz));          // the inputs were last read before the first hand-off
    __syncwarp();
}
// acc <- acc + (BX, BY)  (madd-2007-bl with the exceptional cases of tower.cuh jac_add_mixed); binf: the addend is infinity
static __device__ __noinline__ void g_madd(bool binf) {
    const u32 t = q_tid(), me = g_me();
    const bool e = q_role();
    Fp X1 = q_ld(G_AX, t), Y1 = q_ld(G_AY, t), Z1 = q_ld(G_AZ, t);
    const bool ainf = pair_and(Z1.is_zero());
    Fp Z1Z1 = q_sqr(G_AZ);
    Fp tt = q_mul2(q_cell(G_BY, me), q_cell(G_AZ, me));
    q_st(G_S0, Z1Z1); q_st(G_S1, tt);
    __syncwarp();
    Fp U2 = q_mul2(q_cell(G_BX, me), q_cell(G_S0, me));
    Fp S2 = q_mul2(q_cell(G_S1, me), q_cell(G_S0, me));
    Fp H = U2 - X1, r = dbl(S2 - Y1);
    const bool hz = pair_and(H.is_zero()), rz = pair_and(r.is_zero());
    __syncwarp();
    q_st(G_S0, H); q_st(G_S1, r);
    __syncwarp();
    Fp HH = q_sqr(G_S0);
    Fp r2 = q_sqr(G_S1);
    Fp Z3 = dbl(q_mul2(q_cell(G_AZ, me), q_cell(G_S0, me)));      // (Z1 + H)^2 - Z1Z1 - HH = 2 Z1 H
    q_st(G_S2, dbl(dbl(HH)));                                      // I
    __syncwarp();
    Fp J = q_mul2(q_cell(G_S0, me), q_cell(G_S2, me));
    Fp V = q_mul2(q_cell(G_AX, me), q_cell(G_S2, me));
    Fp X3 = r2 - J - dbl(V);
    q_st(G_S3, J);
    __syncwarp();
    q_st(G_S2, V - X3);                                            // I has been read by both lanes
    __syncwarp();
    Fp Y3 = q_mul2(q_cell(G_S1, me), q_cell(G_S2, me)) - dbl(q_mul2(q_cell(G_AY, me), q_cell(G_S3, me)));
    const bool reg = !binf && !ainf;
    const bool dblcase = reg && hz && rz, infcase = reg && hz && !rz;
    Fp one = e ? Fp::zero() : fp_one();
    Fp nx = X3, ny = Y3, nz = Z3;
    if (infcase) { nx = Fp::zero(); ny = one; nz = Fp::zero(); }
    if (ainf) { nx = q_ld(G_BX, t); ny = q_ld(G_BY, t); nz = one; }
    if (binf) { nx = X1; ny = Y1; nz = Z1; }
    if (__any_sync(0xffffffffu, dblcase)) {      // some pair of the warp adds a point to itself: the whole warp doubles, the others keep their sum
        __syncwarp();
        g_dbl();
        if (dblcase) { nx = q_ld(G_AX, t); ny = q_ld(G_AY, t); nz = q_ld(G_AZ, t); }
    }
    __syncwarp();
    q_st(G_AX, nx); q_st(G_AY, ny); q_st(G_AZ, nz);
    __syncwarp();
}
// my half of table entry `idx` of unit u (negated y for a negative digit)
TCB_D void g_fetch(const AffStore<Fp2S> *tab, size_t u, u32 idx, bool neg, Fp &x, Fp &y) {
    const AffStore<Fp2S> &en = tab[8 * u + idx];
    const bool e = q_role();
    x = ldg_fp2(e ? &en.x.c1 : &en.x.c0);
    y = ldg_fp2(e ? &en.y.c1 : &en.y.c0);
    if (neg) y = -y;
}
// Straus accumulation of task_g2_msm_acc on cells: unit w = (item, group g) adds the shares g, g + G, ... of its item
TCB_D void g2_msm_acc_cells(size_t units, size_t m, size_t G, const AffStore<Fp2S> *tab, const Gls4Digits *dgs, JacStore<Fp2S> *out) {
    size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool live = w < units;
    if (!live) w = units - 1;                    // tail pairs recompute the last unit: the warp stays converged around the hand-offs
    const size_t item = w / G, g = w % G;
    const size_t cnt = (m - g + G - 1) / G, cnt_max = (m + G - 1) / G;       // cnt_max: every pair of the warp walks the same number of steps
    const size_t base = item * m + g;
    const bool e = q_role();
    const u32 t = q_tid();
    Fp one = e ? Fp::zero() : fp_one();
    q_st(G_AX, Fp::zero()); q_st(G_AY, one); q_st(G_AZ, Fp::zero());
    __syncwarp();
    // flattened (digit position, share) walk with the next table entry fetched before the current addition
    Fp nx, ny;
    bool ninf;
    {
        const size_t u = base;
        u32 d = dgs[u].digit(GLS4_L);
        g_fetch(tab, u, d >> 1, d & 1, nx, ny);
        ninf = (dgs[u].flags & 2u) != 0;
    }
    for (int j = GLS4_L; j >= 0; j--) {
        g_dbl();
        for (size_t si = 0; si < cnt_max; si++) {
            q_st(G_BX, nx); q_st(G_BY, ny);
            const bool binf = ninf;
            size_t si2 = si + 1;
            int j2 = j;
            if (si2 == cnt_max) { si2 = 0; j2 = j - 1; }
            if (j2 >= 0) {
                const bool has = si2 < cnt;
                const size_t u = base + (has ? si2 : 0) * G;
                u32 d = dgs[u].digit(j2);
                g_fetch(tab, u, d >> 1, d & 1, nx, ny);
                ninf = !has || (dgs[u].flags & 2u) != 0;
            }
            __syncwarp();
            g_madd(binf);
        }
    }
    for (size_t si = 0; si < cnt_max; si++) {    // first mini-scalar was even: subtract P (table entry 0)
        const bool has = si < cnt;
        const size_t u = base + (has ? si : 0) * G;
        const bool sub = has && (dgs[u].flags & 3u) == 1u;
        Fp x, y;
        g_fetch(tab, u, 0, true, x, y);
        q_st(G_BX, x); q_st(G_BY, y);
        __syncwarp();
        g_madd(!sub);
    }
    if (live) {
        JacStore<Fp2S> &o = out[w];
        stg_fp2(e ? &o.x.c1 : &o.x.c0, q_ld(G_AX, t));
        stg_fp2(e ? &o.y.c1 : &o.y.c0, q_ld(G_AY, t));
        stg_fp2(e ? &o.z.c1 : &o.z.c0, q_ld(G_AZ, t));
    }
}

}  // namespace tcb
#endif
