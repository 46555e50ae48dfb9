// k_g1.cu — G1 kernels (one item per thread), Lagrange coefficients, Commitment::evaluate, roofline probes.
// The Fp multiply is a real function in this translation unit (fp.cuh, TCB_FP_NOINLINE): with it inlined
// k_commit_eval spent 71 % of its stall samples on instruction fetch.
#define TCB_FP_NOINLINE 1
#include "kern.h"
#include "scheme.cuh"
using namespace tcb;

__global__ void __launch_bounds__(128) k_lagrange(size_t n, size_t m, const u8 *xs, u32 *lam, u8 *status) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    size_t item = t / m, i = t % m;
    u8 st = 0;
    lagrange_coeff(xs + item * m * 32, m, i, lam + 8 * t, st);
    if (st) status[item] = st;
}
__global__ void __launch_bounds__(128) k_lagrange_nd(size_t n, size_t m, const u8 *xs, LagrangeND *nd, u8 *status) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    size_t item = t / m, i = t % m;
    u8 st = 0;
    lagrange_num_den(xs + item * m * 32, m, i, nd[t], st);
    if (st) status[item] = st;
}
__global__ void __launch_bounds__(128) k_lagrange_finish(size_t n, size_t m, LagrangeND *nd, u32 *lam) {
    size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item < n) lagrange_finish_item(nd + item * m, m, lam + 8 * item * m);
}
__global__ void __launch_bounds__(128) k_g1_mul(size_t n, const u8 *sk, const u8 *pts, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g1_mul(i, sk, pts, out);
}
__global__ void __launch_bounds__(128) k_g1_mul_store(size_t units, const u32 *k, const u8 *pts, Jac1Store *out, u8 *status, size_t per_item) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < units) task_g1_mul_store(i, k, pts, out, status, per_item);
}
__global__ void __launch_bounds__(128) k_g1_msm_prep(size_t units, const u32 *k, const u8 *pts, Aff1Store *tab, Glv2Digits *dg, u8 *status, size_t per_item) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < units) task_g1_msm_prep(i, k, pts, tab, dg, status, per_item);
}
// pinned to 3 blocks/SM (168 registers, no spills): left to ptxas the kernel drifted to 180 registers = 2 blocks/SM when unrelated code of
// this translation unit changed (11.36 -> 11.75 ms per 2^12 x 65 shares, profiles/r2t_ vs r2z_other_kernels_raw_subset.json)
#ifndef TCB_G1_MSM_MINB
#define TCB_G1_MSM_MINB 3
#endif
__global__ void __launch_bounds__(128, TCB_G1_MSM_MINB) k_g1_msm_acc(size_t units, size_t m, size_t G, const Aff1Store *tab, const Glv2Digits *dg, Jac1Store *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < units) task_g1_msm_acc(i, m, G, tab, dg, out);
}
__global__ void __launch_bounds__(128) k_g1_msm_acc_ba(size_t units, size_t m, size_t G, const Aff1Store *tab, const Glv2Digits *dg,
                                                       Aff1Store *buf_a, Aff1Store *buf_b, Fp *prefix, size_t cnt_max, Jac1Store *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < units) task_msm_acc_ba<MsmG1>(i, m, G, tab, dg, buf_a, buf_b, prefix, cnt_max, out);
}
// experiment (tcb_set_msm_algo 3): the G2 accumulation with ONE item per thread (unsliced Fp2, multiplies as calls)
// instead of one per lane pair; same tables, digits and outputs as k_g2_msm_acc
__global__ void __launch_bounds__(128) k_g2_msm_acc_thread(size_t units, size_t m, size_t G, const AffStore<Fp2> *tab, const Gls4Digits *dg, JacStore<Fp2> *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < units) task_g2_msm_acc<Fp2>(i, m, G, tab, dg, out);
}
__global__ void __launch_bounds__(128) k_g1_sum(size_t n, size_t m, const Jac1Store *terms, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_g1(out + 96 * i, g1_sum(i, m, terms));
}
// PublicKeySet::decrypt tail: g = sum of terms (or the first share when t == 0), then xor_with_hash
__global__ void __launch_bounds__(128) k_decrypt_finish(size_t n, size_t m, const Jac1Store *terms, const u8 *first_shares,
                                                        const u8 *v, const u64 *voff, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Aff<Fp> g;
    if (terms) g = g1_sum(i, m, terms);
    else { bool ok = true; g = load_g1(first_shares + 96 * i, ok); }
    xor_with_hash(out + voff[i], g, v + voff[i], (size_t)(voff[i + 1] - voff[i]));
}
// first half of the two-kernel hash_g2 (scheme.cuh: g2_random_point): SHA3, ChaCha, candidate search and the square root, one
// THREAD per item; tail lanes recompute the last item so that whole warps stay alive for the __syncwarp() after the rejection loop
__global__ void __launch_bounds__(128, 4) k_hash_g2_point(size_t n, const u8 *msgs, const u64 *off, G2PointStore *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < n;
    task_hash_g2_point(live ? i : n - 1, msgs, off, out + (live ? i : n));      // record n absorbs the tail lanes
}
__global__ void __launch_bounds__(128, 4) k_hash_g1_g2_point(size_t n, const u8 *g1, const u8 *msgs, const u64 *off, G2PointStore *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < n;
    task_hash_g1_g2_point(live ? i : n - 1, g1, msgs, off, out + (live ? i : n));
}
// experiment (tcb_set_hash_algo 2): the cofactor clearing with one THREAD per item as well (unsliced Fp2, no lane exchanges)
#ifndef TCB_CLEAR_THREAD_MINB
#define TCB_CLEAR_THREAD_MINB 2
#endif
__global__ void __launch_bounds__(128, TCB_CLEAR_THREAD_MINB) k_g2_clear_thread(size_t n, const G2PointStore *pts, u8 *out, int exact, u8 *redo) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g2_clear<Fp2>(i, pts, out, exact != 0, redo);
}
__global__ void __launch_bounds__(128) k_g1_decode(size_t n, const u8 *pts, Aff1Store *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g1_decode(i, pts, out);
}
// 4 blocks/SM (<= 128 registers): 2^16 evaluation points fit one wave of 148 x 512 threads
__global__ void __launch_bounds__(128, 4) k_commit_eval(size_t n, size_t deg, const Aff1Store *coeff, const u8 *x, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_commit_eval(i, deg, coeff, x, out);
}

__global__ void __launch_bounds__(128, 4) k_commit_eval_part(size_t units, size_t B, size_t L, size_t deg, const Aff1Store *coeff, const u8 *x, Jac1Store *out) {
    size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < units) task_commit_eval_part(u, B, L, deg, coeff, x, out);
}

__global__ void __launch_bounds__(128) k_fr_to_mont(size_t n, const u8 *in, Fr *out, u8 *bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_fr_to_mont(i, in, out, bad);
}
__global__ void __launch_bounds__(128) k_poly_eval(size_t n, size_t deg, const Fr *cm, const u8 *x, u8 *out, u8 *bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_poly_eval(i, deg, cm, x, out, bad);
}
__global__ void __launch_bounds__(128) k_poly_mul(size_t units, size_t da, size_t db, const Fr *am, const Fr *bm, u8 *out) {
    size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < units) task_poly_mul(u, da, db, am, bm, out);
}

static __device__ __forceinline__ u64 splitmix(u64 &s) {
    u64 z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static __device__ Fp rand_fp(u64 &s, int mode) {
    Fp r;
    for (;;) {
        for (int i = 0; i < 12; i += 2) { u64 v = splitmix(s); r.l[i] = (u32)v; r.l[i + 1] = (u32)(v >> 32); }
        if (mode == 1) { for (int i = 0; i < 12; i++) r.l[i] = FpParams::mod(i); r.l[0] -= 1; return r; }
        if (mode == 2) { for (int i = 0; i < 12; i++) r.l[i] = 0; return r; }
        if (mode == 3) { for (int i = 0; i < 12; i++) r.l[i] = 0; r.l[0] = 1; return r; }
        r.l[11] &= 0x1fffffffu;
        if (limbs_lt_mod<FpParams>(r.l)) return r;
    }
}
// 16 independent IMAD.WIDE.U32 accumulators per thread and nothing else in the loop body: the
// integer-MAC ceiling (32 MAC/clk/SM on B200; the first version of this probe had 8 accumulators
// plus a dependent add per row and under-read the ceiling by ~5%, see profiles/r1g_microbench.txt)
__global__ void __launch_bounds__(256) k_probe_imad(u64 *out, int iters, u32 seed) {
    u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    u64 acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc[k] = (u64)(a + k) * (u32)(b + r) + acc[k];
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2 independent Montgomery-multiply chains per thread
__global__ void __launch_bounds__(256) k_probe_fpmul(Fp *out, int iters, u64 seed) {
    u64 s = seed + threadIdx.x + (u64)blockIdx.x * 1024;
    Fp a = rand_fp(s, 0), b = rand_fp(s, 0), c = rand_fp(s, 0);
    for (int it = 0; it < iters; it++) { a = mmul<FpParams>(a, c); b = mmul<FpParams>(b, c); }   // inlined on purpose
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}


__global__ void __launch_bounds__(128) k_encrypt_uv(size_t n, const u8 *pk, const u8 *r, const u8 *msgs, const u64 *off, u8 *u_out, u8 *v_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_encrypt_uv(i, pk, r, msgs, off, u_out, v_out);
}
__global__ void __launch_bounds__(128) k_g1_compress(size_t n, const u8 *unc, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g1_compress(i, unc, out);
}
__global__ void __launch_bounds__(128) k_g1_decompress(size_t n, const u8 *in, u8 *out, u8 *status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g1_decompress(i, in, out, status);
}
namespace tcbk {
static inline unsigned grid1(size_t n) { return (unsigned)((n + 127) / 128); }
cudaError_t upload_consts_g1(const Consts &c) { return cudaMemcpyToSymbol(d_consts, &c, sizeof c); }
size_t g1_term_bytes() { return sizeof(Jac1Store); }
size_t fp_bytes() { return sizeof(Fp); }
void run_lagrange(cudaStream_t st, size_t n, size_t m, const u8 *xs, u32 *lam, u8 *status) {
    if (n * m) k_lagrange<<<grid1(n * m), 128, 0, st>>>(n, m, xs, lam, status);
}
size_t lagrange_nd_bytes() { return sizeof(LagrangeND); }
void run_lagrange_two_pass(cudaStream_t st, size_t n, size_t m, const u8 *xs, void *nd, u32 *lam, u8 *status) {
    if (!(n * m)) return;
    k_lagrange_nd<<<grid1(n * m), 128, 0, st>>>(n, m, xs, (LagrangeND *)nd, status);
    k_lagrange_finish<<<grid1(n), 128, 0, st>>>(n, m, (LagrangeND *)nd, lam);
}
void run_g1_mul(cudaStream_t st, size_t n, const u8 *sk, const u8 *pts, u8 *out) {
    if (n) k_g1_mul<<<grid1(n), 128, 0, st>>>(n, sk, pts, out);
}
void run_g1_mul_store(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *terms, u8 *status, size_t per_item) {
    if (units) k_g1_mul_store<<<grid1(units), 128, 0, st>>>(units, k, pts, (Jac1Store *)terms, status, per_item);
}
size_t g1_msm_tab_bytes() { return 8 * sizeof(Aff1Store); }
size_t g1_msm_dg_bytes() { return sizeof(Glv2Digits); }
size_t g1_msm_units_per_sm() {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_g1_msm_acc, 128, 0) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * 128;
}
void run_g1_msm_prep(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *tab, void *dg, u8 *status, size_t per_item) {
    if (units) k_g1_msm_prep<<<grid1(units), 128, 0, st>>>(units, k, pts, (Aff1Store *)tab, (Glv2Digits *)dg, status, per_item);
}
void run_g1_msm_acc(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out) {
    if (units) k_g1_msm_acc<<<grid1(units), 128, 0, st>>>(units, m, G, (const Aff1Store *)tab, (const Glv2Digits *)dg, (Jac1Store *)out);
}
size_t g2_msm_thread_units_per_sm() {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_g2_msm_acc_thread, 128, 0) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * 128;
}
void run_g2_msm_acc_thread(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out) {
    if (units) k_g2_msm_acc_thread<<<grid1(units), 128, 0, st>>>(units, m, G, (const AffStore<Fp2> *)tab, (const Gls4Digits *)dg, (JacStore<Fp2> *)out);
}
size_t g1_msm_ba_units_per_sm() {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_g1_msm_acc_ba, 128, 0) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * 128;
}
size_t g1_msm_ba_point_bytes(size_t cnt_max) { return ba_points_per_unit<MsmG1>(cnt_max) * sizeof(Aff1Store); }
size_t g1_msm_ba_prefix_bytes(size_t cnt_max) { return ba_prefix_per_unit<MsmG1>(cnt_max) * sizeof(Fp); }
void run_g1_msm_acc_ba(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *buf_a, void *buf_b, void *prefix, size_t cnt_max, void *out) {
    if (units) k_g1_msm_acc_ba<<<grid1(units), 128, 0, st>>>(units, m, G, (const Aff1Store *)tab, (const Glv2Digits *)dg, (Aff1Store *)buf_a,
                                                            (Aff1Store *)buf_b, (Fp *)prefix, cnt_max, (Jac1Store *)out);
}
void run_g1_sum(cudaStream_t st, size_t n, size_t m, const void *terms, u8 *out) {
    if (n) k_g1_sum<<<grid1(n), 128, 0, st>>>(n, m, (const Jac1Store *)terms, out);
}
void run_decrypt_finish(cudaStream_t st, size_t n, size_t m, const void *terms, const u8 *first_shares, const u8 *v, const u64 *voff, u8 *out) {
    if (n) k_decrypt_finish<<<grid1(n), 128, 0, st>>>(n, m, (const Jac1Store *)terms, first_shares, v, voff, out);
}
void run_hash_g2_point(cudaStream_t st, size_t n, const u8 *msgs, const u64 *off, void *pts) {   // pts: n + 1 records (the last one absorbs the tail lanes)
    if (n) k_hash_g2_point<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, msgs, off, (G2PointStore *)pts);
}
void run_hash_g1_g2_point(cudaStream_t st, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, void *pts) {
    if (n) k_hash_g1_g2_point<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, g1, msgs, off, (G2PointStore *)pts);
}
void run_g2_clear_thread(cudaStream_t st, size_t n, const void *pts, u8 *out, bool exact, u8 *redo) {
    if (n) k_g2_clear_thread<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, (const G2PointStore *)pts, out, exact ? 1 : 0, redo);
}
void run_g1_decode(cudaStream_t st, size_t n, const u8 *pts, void *tab) {
    if (n) k_g1_decode<<<grid1(n), 128, 0, st>>>(n, pts, (Aff1Store *)tab);
}
void run_commit_eval(cudaStream_t st, size_t n, size_t deg, const void *tab, const u8 *x, u8 *out) {
    if (n) k_commit_eval<<<grid1(n), 128, 0, st>>>(n, deg, (const Aff1Store *)tab, x, out);
}
size_t commit_eval_units_per_sm() {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_commit_eval_part, 128, 0) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * 128;
}
void run_commit_eval_part(cudaStream_t st, size_t n, size_t B, size_t L, size_t deg, const void *tab, const u8 *x, void *terms) {
    if (n * B) k_commit_eval_part<<<grid1(n * B), 128, 0, st>>>(n * B, B, L, deg, (const Aff1Store *)tab, x, (Jac1Store *)terms);
}
size_t fr_bytes() { return sizeof(Fr); }
void run_fr_to_mont(cudaStream_t st, size_t n, const u8 *in, void *out, u8 *bad) { if (n) k_fr_to_mont<<<grid1(n), 128, 0, st>>>(n, in, (Fr *)out, bad); }
void run_poly_eval(cudaStream_t st, size_t n, size_t deg, const void *cm, const u8 *x, u8 *out, u8 *bad) {
    if (n) k_poly_eval<<<grid1(n), 128, 0, st>>>(n, deg, (const Fr *)cm, x, out, bad);
}
void run_poly_mul(cudaStream_t st, size_t n, size_t da, size_t db, const void *am, const void *bm, u8 *out) {
    size_t units = n * (da + db + 1);
    if (units) k_poly_mul<<<grid1(units), 128, 0, st>>>(units, da, db, (const Fr *)am, (const Fr *)bm, out);
}
void run_encrypt_uv(cudaStream_t st, size_t n, const u8 *pk, const u8 *r, const u8 *msgs, const u64 *off, u8 *u_out, u8 *v_out) {
    if (n) k_encrypt_uv<<<grid1(n), 128, 0, st>>>(n, pk, r, msgs, off, u_out, v_out);
}
void run_g1_compress(cudaStream_t st, size_t n, const u8 *unc, u8 *out) { if (n) k_g1_compress<<<grid1(n), 128, 0, st>>>(n, unc, out); }
void run_g1_decompress(cudaStream_t st, size_t n, const u8 *in, u8 *out, u8 *status) { if (n) k_g1_decompress<<<grid1(n), 128, 0, st>>>(n, in, out, status); }
void run_probe_imad(cudaStream_t st, int blocks, int threads, u64 *out, int iters) { k_probe_imad<<<blocks, threads, 0, st>>>(out, iters, 12345u); }
void run_probe_fpmul(cudaStream_t st, int blocks, int threads, void *out, int iters) { k_probe_fpmul<<<blocks, threads, 0, st>>>((Fp *)out, iters, 99ULL); }
}  // namespace tcbk
