// k_g2.cu — G2 kernels on the lane-pair engine (Fp2S): hash_g2, sign, per-share terms and sums.
// The sliced Fp2 multiply / square are real functions in this translation unit: with them inlined
// the hash kernel spent 61% of its stall samples on instruction fetch (profiles/r1d_*), as functions
// k_hash_g2 runs 2x faster.  (The quad pairing kernel in k_pairing.cu prefers them inlined.)
#define TCB_FP2S_NOINLINE 1
#define TCB_FP_POW_CALL 1      // fixed-exponent powers (square roots) call one shared Fp multiply (tower.cuh)
#define TCB_Q_CALLS 1          // cell products of g2sm.cuh as calls: inlined (the pairing kernels' policy, quadsm.cuh) k_g2_msm_acc_sm is slower (29.9 vs 27.7 ms per 2^14 combines)
#include "kern.h"
#include "scheme.cuh"
#include "g2sm.cuh"
using namespace tcb;
typedef Fp2S F2;
#ifndef TCB_G2_MINB
#define TCB_G2_MINB 2
#endif

static __device__ __forceinline__ size_t unit_index() { return ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1; }
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_hash_g2(size_t n, const u8 *msgs, const u64 *off, u8 *out, int exact, const u8 *only) {
    size_t i = unit_index();
    if (i < n) task_hash_g2<F2>(i, msgs, off, out, exact != 0, only);
}
// second half of the two-kernel hash_g2 (scheme.cuh: g2_random_point / task_g2_clear): cofactor clearing of the curve points
// The clearing kernel alone needs few registers (its Fp2 products are by-value calls): 128 registers = 4 blocks/SM without spills.
// Measured on one box (profiles/kbench_r2g2*.json): exact hash_g2 19.47 ms at 2 blocks/SM, 18.76 at 3, 18.73 at 4; the other lane-pair
// kernels stay at 2 blocks/SM (all of them at 3: combine 29.3 instead of 23.1 ms per 2^14, but 5.2 instead of 5.9 ms per 2048).
#ifndef TCB_CLEAR_MINB
#define TCB_CLEAR_MINB 4
#endif
__global__ void __launch_bounds__(128, TCB_CLEAR_MINB) k_g2_clear(size_t n, const G2PointStore *pts, u8 *out, int exact, u8 *redo) {
    size_t i = unit_index();
    if (i < n) task_g2_clear<F2>(i, pts, out, exact != 0, redo);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_hash_g1_g2(size_t n, const u8 *g1, const u8 *msgs, const u64 *off, u8 *out, const u8 *only) {
    size_t i = unit_index();
    if (i < n) task_hash_g1_g2<F2>(i, g1, msgs, off, out, only);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_sign(size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    size_t i = unit_index();
    if (i < n) task_sign<F2>(i, sk, msgs, off, h, out);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_mul_store(size_t units, const u32 *k, const u8 *pts, JacStore<F2> *out, u8 *status, size_t per_item) {
    size_t i = unit_index();
    if (i < units) task_g2_mul_store<F2>(i, k, pts, out, status, per_item);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_sum(size_t n, size_t m, const JacStore<F2> *terms, u8 *out) {
    size_t i = unit_index();
    if (i < n) task_g2_sum<F2>(i, m, terms, out);
}
// shared-doubling multi-scalar multiplication (scheme.cuh): per-(item, share) digits + affine table, then
// one lane pair per (item, group)
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_msm_prep(size_t units, const u32 *k, const u8 *pts, AffStore<F2> *tab, Gls4Digits *dg, u8 *status, size_t per_item) {
    size_t i = unit_index();
    if (i < units) task_g2_msm_prep<F2>(i, k, pts, tab, dg, status, per_item);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_msm_acc(size_t units, size_t m, size_t G, const AffStore<F2> *tab, const Gls4Digits *dg, JacStore<F2> *out) {
    size_t i = unit_index();
    if (i < units) task_g2_msm_acc<F2>(i, m, G, tab, dg, out);
}
// spill layout (scheme.cuh: task_g2_msm_acc_spill): n main units + ceil(n / q) units that take the last share of q items each
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_msm_acc_spill(size_t units, size_t n, size_t m, size_t q, const AffStore<F2> *tab, const Gls4Digits *dg, JacStore<F2> *out) {
    size_t i = unit_index();
    if (i < units) task_g2_msm_acc_spill<F2>(i, n, m, q, tab, dg, out);
}
// the same accumulation with the running point, the table entry and the temporaries in shared-memory cells (g2sm.cuh): 4 blocks/SM
#ifndef TCB_G2SM_MINB
#define TCB_G2SM_MINB 4
#endif
__global__ void __launch_bounds__(QNT, TCB_G2SM_MINB) k_g2_msm_acc_sm(size_t units, size_t m, size_t G, const AffStore<F2> *tab, const Gls4Digits *dg, JacStore<F2> *out) {
    g2_msm_acc_cells(units, m, G, tab, dg, out);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_msm_acc_ba(size_t units, size_t m, size_t G, const AffStore<F2> *tab, const Gls4Digits *dg,
                                                                   AffStore<F2> *buf_a, AffStore<F2> *buf_b, Fp2c *prefix, size_t cnt_max, JacStore<F2> *out) {
    size_t i = unit_index();
    if (i < units) task_msm_acc_ba<MsmG2<F2>>(i, m, G, tab, dg, buf_a, buf_b, prefix, cnt_max, out);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_compress(size_t n, const u8 *unc, u8 *out) {
    size_t i = unit_index();
    if (i < n) task_g2_compress<F2>(i, unc, out);
}
__global__ void __launch_bounds__(128, TCB_G2_MINB) k_g2_decompress(size_t n, const u8 *in, u8 *out, u8 *status) {
    size_t i = unit_index();
    if (i < n) task_g2_decompress<F2>(i, in, out, status);
}
namespace tcbk {
static inline unsigned grid2(size_t units) { return (unsigned)((units * 2 + 127) / 128); }
cudaError_t upload_consts_g2(const Consts &c) {
    cudaError_t e = cudaFuncSetAttribute(k_g2_msm_acc_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);    // per device
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(d_consts, &c, sizeof c);
}
void run_hash_g2(cudaStream_t st, size_t n, const u8 *msgs, const u64 *off, u8 *out, bool exact, const u8 *only) {
    if (n) k_hash_g2<<<grid2(n), 128, 0, st>>>(n, msgs, off, out, exact ? 1 : 0, only);
}
size_t g2_point_bytes() { return sizeof(G2PointStore); }
void run_g2_clear(cudaStream_t st, size_t n, const void *pts, u8 *out, bool exact, u8 *redo) {   // (32-thread blocks to smooth the partial last wave: measured, no change)
    if (n) k_g2_clear<<<grid2(n), 128, 0, st>>>(n, (const G2PointStore *)pts, out, exact ? 1 : 0, redo);
}
void run_hash_g1_g2(cudaStream_t st, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, u8 *out, const u8 *only) {
    if (n) k_hash_g1_g2<<<grid2(n), 128, 0, st>>>(n, g1, msgs, off, out, only);
}
void run_sign(cudaStream_t st, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    if (n) k_sign<<<grid2(n), 128, 0, st>>>(n, sk, msgs, off, h, out);
}
void run_g2_compress(cudaStream_t st, size_t n, const u8 *unc, u8 *out) { if (n) k_g2_compress<<<grid2(n), 128, 0, st>>>(n, unc, out); }
void run_g2_decompress(cudaStream_t st, size_t n, const u8 *in, u8 *out, u8 *status) { if (n) k_g2_decompress<<<grid2(n), 128, 0, st>>>(n, in, out, status); }
size_t g2_term_bytes() { return sizeof(JacStore<F2>); }
void run_g2_mul_store(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *terms, u8 *status, size_t per_item) {
    if (units) k_g2_mul_store<<<grid2(units), 128, 0, st>>>(units, k, pts, (JacStore<F2> *)terms, status, per_item);
}
void run_g2_msm_acc_spill(cudaStream_t st, size_t n, size_t m, size_t q, const void *tab, const void *dg, void *out) {
    size_t units = n + (n + q - 1) / q;
    if (n) k_g2_msm_acc_spill<<<grid2(units), 128, 0, st>>>(units, n, m, q, (const AffStore<F2> *)tab, (const Gls4Digits *)dg, (JacStore<F2> *)out);
}
size_t g2_msm_tab_bytes() { return 8 * sizeof(AffStore<F2>); }
size_t g2_msm_dg_bytes() { return sizeof(Gls4Digits); }
size_t g2_msm_units_per_sm() {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_g2_msm_acc, 128, 0) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * 64;
}
void run_g2_msm_prep(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *tab, void *dg, u8 *status, size_t per_item) {
    if (units) k_g2_msm_prep<<<grid2(units), 128, 0, st>>>(units, k, pts, (AffStore<F2> *)tab, (Gls4Digits *)dg, status, per_item);
}
void run_g2_msm_acc(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out) {
    if (units) k_g2_msm_acc<<<grid2(units), 128, 0, st>>>(units, m, G, (const AffStore<F2> *)tab, (const Gls4Digits *)dg, (JacStore<F2> *)out);
}
size_t g2_msm_sm_units_per_sm() {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_g2_msm_acc_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES); attr = true; }
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_g2_msm_acc_sm, QNT, G_SMEM_BYTES) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * (QNT / 2);
}
void run_g2_msm_acc_sm(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out) {
    if (units) k_g2_msm_acc_sm<<<(unsigned)((units * 2 + QNT - 1) / QNT), QNT, G_SMEM_BYTES, st>>>(units, m, G, (const AffStore<F2> *)tab, (const Gls4Digits *)dg, (JacStore<F2> *)out);
}
size_t g2_msm_ba_units_per_sm() {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_g2_msm_acc_ba, 128, 0) != cudaSuccess || blocks < 1) blocks = 1;
    return (size_t)blocks * 64;
}
size_t g2_msm_ba_point_bytes(size_t cnt_max) { return ba_points_per_unit<MsmG2<F2>>(cnt_max) * sizeof(AffStore<F2>); }
size_t g2_msm_ba_prefix_bytes(size_t cnt_max) { return ba_prefix_per_unit<MsmG2<F2>>(cnt_max) * sizeof(Fp2c); }
void run_g2_msm_acc_ba(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *buf_a, void *buf_b, void *prefix, size_t cnt_max, void *out) {
    if (units) k_g2_msm_acc_ba<<<grid2(units), 128, 0, st>>>(units, m, G, (const AffStore<F2> *)tab, (const Gls4Digits *)dg, (AffStore<F2> *)buf_a,
                                                            (AffStore<F2> *)buf_b, (Fp2c *)prefix, cnt_max, (JacStore<F2> *)out);
}
void run_g2_sum(cudaStream_t st, size_t n, size_t m, const void *terms, u8 *out) {
    if (n) k_g2_sum<<<grid2(n), 128, 0, st>>>(n, m, (const JacStore<F2> *)terms, out);
}
}  // namespace tcbk
