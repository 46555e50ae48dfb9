// kern.h — host-callable launchers of the kernels, one group per translation unit so the
// groups compile in parallel (k_pairing.cu, k_g2.cu, k_g1.cu).  tcb200.cu (the C ABI) only
// sees these prototypes.  Every launcher enqueues on `st` and returns; errors are picked up
// by the caller with cudaGetLastError().
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "tower.cuh"

namespace tcbk {
typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

// ---- k_pairing.cu
cudaError_t upload_consts_pairing(const tcb::Consts &c);
void run_verify_g2_quad(cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok);
void run_selftest(cudaStream_t st, size_t n, u64 seed, unsigned long long *bad);
void run_final_exp_quad(cudaStream_t st, size_t n, const void *fbuf, const u8 *enc_ok, u8 *ok, void *fe_out);   // register engine
void run_final_exp_sm(cudaStream_t st, size_t n, const void *fbuf, const u8 *enc_ok, u8 *ok, void *fe_out);     // products / squarings on cells
void run_miller_quad_reg(cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, void *fbuf, u8 *enc_ok);
void run_count_diff(cudaStream_t st, size_t bytes, const void *x, const void *y, unsigned long long *bad);
// ---- k_miller.cu  (shared-memory engine: operands staged in shared memory, dot-product form)
cudaError_t upload_consts_miller(const tcb::Consts &c);
size_t miller_f_bytes();         // bytes of one Miller-loop value in the scratch buffer between the two kernels
void run_miller_quad(cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, void *fbuf, u8 *enc_ok, bool gen_scaled = false);   // c == nullptr: the G1 generator, or [3 (x^2 - 1)] times it
// ---- k_g2.cu  (lane-pair engine)
cudaError_t upload_consts_g2(const tcb::Consts &c);
void run_hash_g2(cudaStream_t st, size_t n, const u8 *msgs, const u64 *off, u8 *out, bool exact = true, const u8 *only = nullptr);   // exact == false: [3 (x^2 - 1)] H(m) (verify only); only != nullptr: just the flagged items
size_t g2_point_bytes();         // two-kernel hash_g2: curve point per item between run_hash_g2_point (k_g1.cu, one thread per item) and run_g2_clear
void run_g2_clear(cudaStream_t st, size_t n, const void *pts, u8 *out, bool exact, u8 *redo);
void run_hash_g1_g2(cudaStream_t st, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, u8 *out, const u8 *only = nullptr);
void run_sign(cudaStream_t st, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out);
void run_g2_compress(cudaStream_t st, size_t n, const u8 *unc, u8 *out);
void run_g2_decompress(cudaStream_t st, size_t n, const u8 *in, u8 *out, u8 *status);
size_t g2_term_bytes();
void run_g2_mul_store(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *terms, u8 *status, size_t per_item);
void run_g2_sum(cudaStream_t st, size_t n, size_t m, const void *terms, u8 *out);
size_t g2_msm_tab_bytes();       // per (item, share): 8 affine table entries
size_t g2_msm_dg_bytes();        // per (item, share): recoded scalar
size_t g2_msm_units_per_sm();    // resident (item, group) units per SM of k_g2_msm_acc
void run_g2_msm_prep(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *tab, void *dg, u8 *status, size_t per_item);
void run_g2_msm_acc(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out);
void run_g2_msm_acc_spill(cudaStream_t st, size_t n, size_t m, size_t q, const void *tab, const void *dg, void *out);   // 2 partial sums per item
size_t g2_msm_sm_units_per_sm();  // the accumulation on shared-memory cells (g2sm.cuh): 4 blocks/SM
void run_g2_msm_acc_sm(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out);
// batch-affine accumulation: scratch per unit = 2 point buffers + prefix products for up to cnt_max shares per unit
size_t g2_msm_ba_units_per_sm();
size_t g2_msm_ba_point_bytes(size_t cnt_max);
size_t g2_msm_ba_prefix_bytes(size_t cnt_max);
void run_g2_msm_acc_ba(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *buf_a, void *buf_b, void *prefix, size_t cnt_max, void *out);
// ---- k_g1.cu
cudaError_t upload_consts_g1(const tcb::Consts &c);
size_t g1_term_bytes();
void run_lagrange(cudaStream_t st, size_t n, size_t m, const u8 *xs, u32 *lam, u8 *status);
size_t lagrange_nd_bytes();
void run_lagrange_two_pass(cudaStream_t st, size_t n, size_t m, const u8 *xs, void *nd, u32 *lam, u8 *status);   // 2 launches
void run_g1_mul(cudaStream_t st, size_t n, const u8 *sk, const u8 *pts, u8 *out);
void run_g1_mul_store(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *terms, u8 *status, size_t per_item);
void run_g1_sum(cudaStream_t st, size_t n, size_t m, const void *terms, u8 *out);
size_t g1_msm_tab_bytes();
size_t g1_msm_dg_bytes();
size_t g1_msm_units_per_sm();
void run_g1_msm_prep(cudaStream_t st, size_t units, const u32 *k, const u8 *pts, void *tab, void *dg, u8 *status, size_t per_item);
void run_g1_msm_acc(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out);
size_t g2_msm_thread_units_per_sm();
void run_g2_msm_acc_thread(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *out);
size_t g1_msm_ba_units_per_sm();
size_t g1_msm_ba_point_bytes(size_t cnt_max);
size_t g1_msm_ba_prefix_bytes(size_t cnt_max);
void run_g1_msm_acc_ba(cudaStream_t st, size_t units, size_t m, size_t G, const void *tab, const void *dg, void *buf_a, void *buf_b, void *prefix, size_t cnt_max, void *out);
void run_decrypt_finish(cudaStream_t st, size_t n, size_t m, const void *terms, const u8 *first_shares, const u8 *v, const u64 *voff, u8 *out);
void run_hash_g2_point(cudaStream_t st, size_t n, const u8 *msgs, const u64 *off, void *pts);
void run_hash_g1_g2_point(cudaStream_t st, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, void *pts);
void run_g2_clear_thread(cudaStream_t st, size_t n, const void *pts, u8 *out, bool exact, u8 *redo);
void run_g1_decode(cudaStream_t st, size_t n, const u8 *pts, void *tab);
void run_commit_eval(cudaStream_t st, size_t n, size_t deg, const void *tab, const u8 *x, u8 *out);
size_t commit_eval_units_per_sm();
void run_commit_eval_part(cudaStream_t st, size_t n, size_t B, size_t L, size_t deg, const void *tab, const u8 *x, void *terms);   // then run_g1_sum(n, B, ...)
size_t fr_bytes();
void run_fr_to_mont(cudaStream_t st, size_t n, const u8 *in, void *out, u8 *bad);
void run_poly_eval(cudaStream_t st, size_t n, size_t deg, const void *cm, const u8 *x, u8 *out, u8 *bad);
void run_poly_mul(cudaStream_t st, size_t n, size_t da, size_t db, const void *am, const void *bm, u8 *out);
void run_encrypt_uv(cudaStream_t st, size_t n, const u8 *pk, const u8 *r, const u8 *msgs, const u64 *off, u8 *u_out, u8 *v_out);
void run_g1_compress(cudaStream_t st, size_t n, const u8 *unc, u8 *out);
void run_g1_decompress(cudaStream_t st, size_t n, const u8 *in, u8 *out, u8 *status);
void run_probe_imad(cudaStream_t st, int blocks, int threads, u64 *out, int iters);
void run_probe_fpmul(cudaStream_t st, int blocks, int threads, void *out, int iters);
size_t fp_bytes();
}  // namespace tcbk
