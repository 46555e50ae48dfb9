// tcb200.cu — the extern "C" boundary (include/tcb200.h) of the B200-native threshold_crypto
// hot path: context, scratch arena, host<->device staging, device sharding, kernel sequencing.
// The kernels live in k_pairing.cu / k_g2.cu / k_g1.cu (launchers in kern.h).  No CPU fallback:
// every entry point enqueues CUDA kernels or fails with an error code.
//
// Kernel inventory:
//   k_miller_quad + k_final_exp_sm   a1/a3  pairing equality, one item per lane QUAD: Miller loop and final exponentiation with their
//                               operands staged in shared memory (quadsm.cuh), f through HBM between the two
//   k_final_exp_quad            the round-1 register-engine final exponentiation (TCB_ENGINE_QUAD_SMEM_REGFE: A/B measurement, self-test reference)
//   k_verify_g2_quad            the round-1 fused register-engine kernel (TCB_ENGINE_QUAD_REG: self-test reference, A/B measurement)
//   k_hash_g2                   a2     SHA3 -> ChaCha20 -> G2::random -> exact cofactor (lane pairs)
//   k_sign                      a4     sk * H(m)                                   (lane pairs)
//   k_lagrange | k_lagrange_nd + k_lagrange_finish   a5   lambda_i(0): one thread per (item, share), or two passes with one inversion per item
//   k_g2_msm_prep / k_g2_msm_acc / k_g2_sum   a6   multi-scalar multiplication with shared doublings (combine_signatures)
//   k_g1_mul / k_g1_msm_prep / k_g1_msm_acc / k_g1_sum / k_decrypt_finish   a7 (decrypt shares, decrypt)
//   k_g*_msm_acc_ba (batch-affine tree: fewer multiplications, slower: memory latency) and k_g*_mul_store (one
//   multiplication per share): alternatives kept for measurement (tcb_set_msm_algo)
//   k_g1_decode / k_commit_eval a8     Commitment::evaluate (Horner)
//   k_selftest_*, k_probe_*     measurement / self-test
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "../../include/tcb200.h"
#include "kern.h"
#include "scheme.cuh"
#include "msm_plan.h"

using namespace tcb;
using namespace tcbk;

// ----------------------------------------------------------------------------- context
struct Chunk { void *p; size_t sz, used; };
struct DevState {
    int dev;
    cudaStream_t stream, stream2;      // host-buffer calls alternate the chunks of a large batch between the two (copy k+1 overlaps compute k)
    cudaStream_t cur_stream = nullptr; // stream of the piece being enqueued (up / down)
    cudaEvent_t last_ev = nullptr;     // end of the last call that used the scratch arena: the next call's stream waits on it
    bool has_last = false;
    std::vector<Chunk> chunks;
    size_t high_water = 0, cur = 0;
};
struct Download { size_t g; cudaStream_t st; void *host; const void *dev; size_t bytes; };
struct tcb_ctx {
    std::vector<DevState> devs;
    std::string err;
    std::vector<Download> pending;  // device -> host copies of the running host-buffer call, enqueued after every device has its work
    uint64_t launches = 0;
    int engine = TCB_ENGINE_QUAD_SMEM;
    int sm_count = 148;
    size_t msm_groups = 0;          // 0 = auto (pick_groups)
    size_t eval_split = 0;          // units per point of Commitment::evaluate: 0 = auto (tcb_set_eval_split)
    bool eval_split_off = false;
    bool verify_exact_hash = false; // tcb_set_verify_hash
    int hash_algo = 0;              // tcb_set_hash_algo: 0 point + clearing kernels, 1 the one-kernel hash_g2 (round 1/2a), 2 clearing per thread (experiment)
    int msm_algo = 0;               // MSM_STRAUS (default) | MSM_BATCH_AFFINE | MSM_PER_SHARE (tcb_set_msm_algo; the others are measurement knobs)
};

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char buf_[256];                                                                             \
            snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = buf_;                                                                            \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

// bump arena: pointers stay valid for the whole API call; consolidated on the next reset
static int arena_reset(tcb_ctx *ctx, DevState &d) {
    if (d.chunks.size() > 1 || (d.chunks.size() == 1 && d.chunks[0].sz < d.high_water)) {
        CK(cudaSetDevice(d.dev));
        CK(cudaDeviceSynchronize());   // scratch of earlier calls may still be in use on the caller's stream
        for (auto &c : d.chunks) cudaFree(c.p);
        d.chunks.clear();
        void *p = nullptr;
        size_t sz = d.high_water + (d.high_water >> 2) + (1 << 20);
        CK(cudaMalloc(&p, sz));
        d.chunks.push_back({p, sz, 0});
    }
    for (auto &c : d.chunks) c.used = 0;
    d.cur = 0;
    return 0;
}
static void *arena_alloc(tcb_ctx *ctx, DevState &d, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    d.cur += bytes;
    if (d.cur > d.high_water) d.high_water = d.cur;
    for (auto &c : d.chunks)
        if (c.sz - c.used >= bytes) { void *p = (char *)c.p + c.used; c.used += bytes; return p; }
    void *p = nullptr;
    size_t sz = bytes > (size_t)(8 << 20) ? bytes : (size_t)(8 << 20);
    cudaSetDevice(d.dev);
    if (cudaMalloc(&p, sz) != cudaSuccess) { ctx->err = "cudaMalloc failed in arena_alloc"; return nullptr; }
    d.chunks.push_back({p, sz, bytes});
    return p;
}

// count a launch and pick up launch errors
#define RUN(call)                      \
    do {                               \
        call;                          \
        ctx->launches++;               \
        CK(cudaGetLastError());        \
    } while (0)

// ----------------------------------------------------------------------------- device-side implementations
static int impl_verify_g2(tcb_ctx *ctx, DevState &dv, cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok, bool gen_scaled = false) {
    if (!n) return 0;
    if (ctx->engine == TCB_ENGINE_QUAD_REG) { RUN(run_verify_g2_quad(st, n, a, b, c, d, ok)); return 0; }   // callers never ask this engine for gen_scaled
    // Miller loop (shared-memory engine) -> f, 576 B per item in scratch -> final exponentiation and "== 1"
    void *fbuf = arena_alloc(ctx, dv, n * miller_f_bytes());
    u8 *enc = (u8 *)arena_alloc(ctx, dv, n);
    if (!fbuf || !enc) return -1;
    RUN(run_miller_quad(st, n, a, b, c, d, fbuf, enc, gen_scaled));
    if (ctx->engine == TCB_ENGINE_QUAD_SMEM_REGFE) RUN(run_final_exp_quad(st, n, fbuf, enc, ok, nullptr));
    else RUN(run_final_exp_sm(st, n, fbuf, enc, ok, nullptr));
    return 0;
}
// hash_g2 (exact) or, for the verifier, [3 (x^2 - 1)] hash_g2.  Default: two kernels — the curve point with one thread per item
// (k_hash_g2_point), the cofactor clearing on lane pairs (k_g2_clear) — and a pass of the one-kernel version over the items whose
// cleared point was the identity (the reference draws further candidates then; probability ~2^-255, the launch is ~free).
static int impl_hash_g2(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, const u8 *msgs, const u64 *off, u8 *out, bool exact = true) {
    if (!n) return 0;
    if (ctx->hash_algo == 1) { RUN(run_hash_g2(st, n, msgs, off, out, exact)); return 0; }
    void *pts = arena_alloc(ctx, d, (n + 1) * g2_point_bytes());
    u8 *redo = (u8 *)arena_alloc(ctx, d, n);
    if (!pts || !redo) return -1;
    RUN(run_hash_g2_point(st, n, msgs, off, pts));
    if (ctx->hash_algo == 2) RUN(run_g2_clear_thread(st, n, pts, out, exact, redo));
    else RUN(run_g2_clear(st, n, pts, out, exact, redo));
    RUN(run_hash_g2(st, n, msgs, off, out, exact, redo));
    return 0;
}
static int impl_verify(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    if (!n) return 0;
    // hash_g2 on lane pairs into scratch, then the quad pairing check e(pk, H) == e(g1, sig) — by default as
    // e(pk, cH) == e(c g1, sig) with c = 3 (x^2 - 1): the hash kernel skips the last third of the cofactor clearing (tcb200.h)
    u8 *h = (u8 *)arena_alloc(ctx, d, n * 192);
    if (!h) return -1;
    const bool scaled = !ctx->verify_exact_hash && ctx->engine != TCB_ENGINE_QUAD_REG;
    if (impl_hash_g2(ctx, d, st, n, msgs, off, h, !scaled)) return -1;
    return impl_verify_g2(ctx, d, st, n, pk, h, nullptr, sig, ok, scaled);
}
// hash_g1_g2 (src/lib.rs:697-707) the same way: point kernel with the compressed g1 in the SHA3 input, exact clearing, fallback pass
static int impl_hash_g1_g2(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, u8 *out) {
    if (!n) return 0;
    if (ctx->hash_algo == 1) { RUN(run_hash_g1_g2(st, n, g1, msgs, off, out)); return 0; }
    void *pts = arena_alloc(ctx, d, (n + 1) * g2_point_bytes());
    u8 *redo = (u8 *)arena_alloc(ctx, d, n);
    if (!pts || !redo) return -1;
    RUN(run_hash_g1_g2_point(st, n, g1, msgs, off, pts));
    if (ctx->hash_algo == 2) RUN(run_g2_clear_thread(st, n, pts, out, true, redo));
    else RUN(run_g2_clear(st, n, pts, out, true, redo));
    RUN(run_hash_g1_g2(st, n, g1, msgs, off, out, redo));
    return 0;
}
// sign = sk * hash_g2(msg): the hash through the two-kernel path into scratch, then the multiplication kernel (h == nullptr); sign_g2
// takes the caller's points
static int impl_sign(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    if (!n) return 0;
    if (!h && ctx->hash_algo != 1) {
        u8 *hs = (u8 *)arena_alloc(ctx, d, n * 192);
        if (!hs || impl_hash_g2(ctx, d, st, n, msgs, off, hs)) return -1;
        h = hs;
    }
    RUN(run_sign(st, n, sk, msgs, off, h, out));
    return 0;
}
// ---- sum_s k_s P_s per item as a multi-scalar multiplication (scheme.cuh, *_msm_*).
// G partial sums ("groups") per item trade shared work (small G) against parallelism (large G); pick the G
// that minimises  waves(n G) * (fixed + per_share * ceil(m / G)).
// (the planning functions live in msm_plan.h so that the host emulation's tests can pin their choices)
enum { MSM_STRAUS = 0, MSM_BATCH_AFFINE = 1, MSM_PER_SHARE = 2, MSM_STRAUS_G2_THREAD = 3, MSM_STRAUS_G2_CELLS = 4, MSM_STRAUS_NO_SPILL = 5, MSM_STRAUS_FORCE_SPILL = 6 };
// out: G Jacobian partial sums per item in `part` (then run_g*_sum(n, G, part, ...))
static int impl_msm(tcb_ctx *ctx, DevState &d, cudaStream_t st, bool g2, size_t n, size_t m, const u32 *k, const u8 *pts, u8 *status, void *&part, size_t &G) {
    const size_t term = g2 ? g2_term_bytes() : g1_term_bytes();
    if (ctx->msm_algo == MSM_PER_SHARE) {
        G = m;
        part = arena_alloc(ctx, d, n * m * term);
        if (!part) return -1;
        if (g2) RUN(run_g2_mul_store(st, n * m, k, pts, part, status, m));
        else RUN(run_g1_mul_store(st, n * m, k, pts, part, status, m));
        return 0;
    }
    const bool ba = ctx->msm_algo == MSM_BATCH_AFFINE;
    const bool plain_straus = ctx->msm_algo == MSM_STRAUS || ctx->msm_algo == MSM_STRAUS_NO_SPILL || ctx->msm_algo == MSM_STRAUS_FORCE_SPILL;
    const bool g2_thread = g2 && ctx->msm_algo == MSM_STRAUS_G2_THREAD;
    // the G2 accumulation on shared-memory cells (g2sm.cuh, 4 blocks/SM): measured 23.0 vs 19.3 ms at 2^14 items (G = 2 costs 17 % more
    // multiply-accumulates and the multiply pipe is the limit either way), 4.2 vs 4.6 ms at 2048 items — a measurement knob, not the default
    const bool g2_cells = g2 && ctx->msm_algo == MSM_STRAUS_G2_CELLS;
    static size_t per_sm[6] = {0, 0, 0, 0, 0, 0};
    size_t &ps = per_sm[g2_cells ? 5 : g2_thread ? 4 : (g2 ? 2 : 0) + (ba ? 1 : 0)];
    if (!ps) ps = g2_cells ? g2_msm_sm_units_per_sm() : g2_thread ? g2_msm_thread_units_per_sm() : g2 ? (ba ? g2_msm_ba_units_per_sm() : g2_msm_units_per_sm()) : (ba ? g1_msm_ba_units_per_sm() : g1_msm_units_per_sm());
    // relative costs in field multiplications: doubling 4.8 / 7, mixed addition 8.6 / 11, affine addition ~5.7 (+ one inversion per tree level)
    // (G1 reads two digit positions per look-up: two doublings per position)
    double fixed = g2 ? (ba ? 4.8 + 8.6 + 6.0 : 4.8) : (ba ? 14.0 + 11.0 + 6.0 : 14.0);
    double share = g2 ? (ba ? 5.2 : 8.6) : (ba ? 5.7 : 11.0);
    G = ctx->msm_groups ? (ctx->msm_groups < m ? ctx->msm_groups : m) : pick_groups(n, m, ps * (size_t)ctx->sm_count, fixed, share);
    // G2, one unit per item, and the batch leaves unit slots of the single wave idle (2^14 items on 18 944 slots): move the last share of
    // every item to the spare units, q items each (scheme.cuh: task_g2_msm_acc_spill) — when that shortens the longest unit
    size_t spill_q = 0;
    if (g2 && ctx->msm_algo == MSM_STRAUS_FORCE_SPILL && m >= 2) { G = 1; spill_q = 2; }       // tests: the layout on any batch
    if (g2 && ctx->msm_algo == MSM_STRAUS && G == 1 && !ctx->msm_groups) spill_q = pick_spill(n, m, ps * (size_t)ctx->sm_count, fixed, share);
    if (spill_q) G = 2;
    void *tab = arena_alloc(ctx, d, n * m * (g2 ? g2_msm_tab_bytes() : g1_msm_tab_bytes()));
    void *dg = arena_alloc(ctx, d, n * m * (g2 ? g2_msm_dg_bytes() : g1_msm_dg_bytes()));
    part = arena_alloc(ctx, d, n * G * term);
    if (!tab || !dg || !part) return -1;
    if (g2) RUN(run_g2_msm_prep(st, n * m, k, pts, tab, dg, status, m));
    else RUN(run_g1_msm_prep(st, n * m, k, pts, tab, dg, status, m));
    if (!ba) {
        if (g2_thread) RUN(run_g2_msm_acc_thread(st, n * G, m, G, tab, dg, part));
        else if (g2_cells) RUN(run_g2_msm_acc_sm(st, n * G, m, G, tab, dg, part));
        else if (g2 && spill_q) RUN(run_g2_msm_acc_spill(st, n, m, spill_q, tab, dg, part));
        else if (g2) RUN(run_g2_msm_acc(st, n * G, m, G, tab, dg, part));
        else RUN(run_g1_msm_acc(st, n * G, m, G, tab, dg, part));
        return 0;
    }
    size_t cnt_max = (m + G - 1) / G;
    size_t pb = g2 ? g2_msm_ba_point_bytes(cnt_max) : g1_msm_ba_point_bytes(cnt_max);
    size_t fb = g2 ? g2_msm_ba_prefix_bytes(cnt_max) : g1_msm_ba_prefix_bytes(cnt_max);
    void *buf_a = arena_alloc(ctx, d, n * G * pb), *buf_b = arena_alloc(ctx, d, n * G * pb), *prefix = arena_alloc(ctx, d, n * G * fb);
    if (!buf_a || !buf_b || !prefix) return -1;
    if (g2) RUN(run_g2_msm_acc_ba(st, n * G, m, G, tab, dg, buf_a, buf_b, prefix, cnt_max, part));
    else RUN(run_g1_msm_acc_ba(st, n * G, m, G, tab, dg, buf_a, buf_b, prefix, cnt_max, part));
    return 0;
}
static int impl_msm_g2(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t m, const u32 *k, const u8 *pts, u8 *status, void *&part, size_t &G) {
    return impl_msm(ctx, d, st, true, n, m, k, pts, status, part, G);
}
static int impl_msm_g1(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t m, const u32 *k, const u8 *pts, u8 *status, void *&part, size_t &G) {
    return impl_msm(ctx, d, st, false, n, m, k, pts, status, part, G);
}
// Lagrange coefficients: with enough items to fill the GPU from one unit per ITEM, the two-pass form shares one inversion per item
static int impl_lagrange(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t m, const u8 *x, u32 *lam, u8 *status) {
    if (n >= (size_t)ctx->sm_count * 16) {
        void *nd = arena_alloc(ctx, d, n * m * lagrange_nd_bytes());
        if (!nd) return -1;
        run_lagrange_two_pass(st, n, m, x, nd, lam, status);
        ctx->launches += 2;
        CK(cudaGetLastError());
        return 0;
    }
    RUN(run_lagrange(st, n, m, x, lam, status));
    return 0;
}
static int impl_combine_g2(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    CK(cudaMemsetAsync(status, 0, n, st));
    if (n == 0) return 0;
    if (t == 0) { CK(cudaMemcpyAsync(out, shares, n * 192, cudaMemcpyDeviceToDevice, st)); return 0; }
    size_t m = t + 1;
    u32 *lam = (u32 *)arena_alloc(ctx, d, n * m * 32);
    if (!lam) return -1;
    if (impl_lagrange(ctx, d, st, n, m, x, lam, status)) return -1;
    void *part; size_t G;
    if (impl_msm_g2(ctx, d, st, n, m, lam, shares, status, part, G)) return -1;
    RUN(run_g2_sum(st, n, G, part, out));
    return 0;
}
// mode 0: write the combined G1 point; mode 1: xor_with_hash (decrypt)
static int impl_combine_g1(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t t, const u8 *x, const u8 *shares,
                           u8 *out, u8 *status, int mode, const u8 *v, const u64 *voff) {
    CK(cudaMemsetAsync(status, 0, n, st));
    if (n == 0) return 0;
    if (t == 0) {
        if (mode == 0) { CK(cudaMemcpyAsync(out, shares, n * 96, cudaMemcpyDeviceToDevice, st)); return 0; }
        RUN(run_decrypt_finish(st, n, 1, nullptr, shares, v, voff, out));
        return 0;
    }
    size_t m = t + 1;
    u32 *lam = (u32 *)arena_alloc(ctx, d, n * m * 32);
    if (!lam) return -1;
    if (impl_lagrange(ctx, d, st, n, m, x, lam, status)) return -1;
    void *part; size_t G;
    if (impl_msm_g1(ctx, d, st, n, m, lam, shares, status, part, G)) return -1;
    if (mode == 0) RUN(run_g1_sum(st, n, G, part, out));
    else RUN(run_decrypt_finish(st, n, G, part, shares, v, voff, out));
    return 0;
}
static int impl_commit_eval(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    void *tab = arena_alloc(ctx, d, (deg + 1) * g1_term_bytes());
    if (!tab) return -1;
    RUN(run_g1_decode(st, deg + 1, coeff, tab));
    if (!n) return 0;
    // Small batches: B units per point (coefficient blocks) so that the evaluation fills the GPU; each block keeps >= 32 coefficients
    // so that the 255-bit recombination multiply stays a small part of the unit's work.
    static size_t per_sm = 0;
    if (!per_sm) per_sm = commit_eval_units_per_sm();
    size_t resident = per_sm * (size_t)ctx->sm_count, B = 1;
    if (!ctx->eval_split_off)
        while (B < 16 && n * B * 2 <= resident && (deg + 1) / (B * 2) >= 32) B *= 2;
    if (ctx->eval_split) B = ctx->eval_split;
    if (B <= 1) { RUN(run_commit_eval(st, n, deg, tab, x, out)); return 0; }
    size_t L = (deg + B) / B;                                   // ceil((deg + 1) / B)
    void *terms = arena_alloc(ctx, d, n * B * g1_term_bytes());
    if (!terms) return -1;
    RUN(run_commit_eval_part(st, n, B, L, deg, tab, x, terms));
    RUN(run_g1_sum(st, n, B, terms, out));
    return 0;
}

// ----------------------------------------------------------------------------- init / free
extern "C" int tcb_init(tcb_ctx **out, const int *device_ids, int n_devices) {
    if (!out) return -2;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return -3;   // no CPU fallback
    tcb_ctx *ctx = new tcb_ctx();
    { const char *e = getenv("TCB200_MSM_ALGO"); if (e && e[0] >= '0' && e[0] <= '6') ctx->msm_algo = e[0] - '0'; }
    Consts C;
    build_consts(C);
    if (n_devices <= 0 || !device_ids) {
        int cur = 0;
        cudaGetDevice(&cur);
        n_devices = 1;
        ctx->devs.resize(1);
        ctx->devs[0].dev = cur;
    } else {
        ctx->devs.resize(n_devices);
        for (int i = 0; i < n_devices; i++) ctx->devs[i].dev = device_ids[i];
    }
    for (auto &d : ctx->devs) {
        if (d.dev < 0 || d.dev >= count) { delete ctx; return -4; }
        if (cudaSetDevice(d.dev) != cudaSuccess) { delete ctx; return -5; }
        if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.last_ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return -6; }
        if (upload_consts_pairing(C) != cudaSuccess || upload_consts_miller(C) != cudaSuccess || upload_consts_g2(C) != cudaSuccess ||
            upload_consts_g1(C) != cudaSuccess) { delete ctx; return -7; }
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->devs[0].dev) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    cudaSetDevice(ctx->devs[0].dev);
    *out = ctx;
    return 0;
}
extern "C" void tcb_free(tcb_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.dev);
        cudaStreamSynchronize(d.stream);
        cudaStreamSynchronize(d.stream2);
        if (d.has_last) cudaEventSynchronize(d.last_ev);
        for (auto &c : d.chunks) {
            cudaMemset(c.p, 0, c.sz);   // scratch may have held secret scalars (src/lib.rs:304-314 hygiene)
            cudaFree(c.p);
        }
        cudaStreamDestroy(d.stream);
        cudaStreamDestroy(d.stream2);
        cudaEventDestroy(d.last_ev);
    }
    delete ctx;
}
extern "C" const char *tcb_last_error(const tcb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int tcb_set_engine(tcb_ctx *ctx, int engine) {
    if (!ctx || (engine != TCB_ENGINE_QUAD_SMEM && engine != TCB_ENGINE_QUAD_REG && engine != TCB_ENGINE_QUAD_SMEM_REGFE)) return -2;
    ctx->engine = engine;
    return 0;
}
extern "C" int tcb_set_verify_hash(tcb_ctx *ctx, int exact) {
    if (!ctx) return -2;
    ctx->verify_exact_hash = exact != 0;
    return 0;
}
extern "C" int tcb_set_hash_algo(tcb_ctx *ctx, int algo) {
    if (!ctx || algo < 0 || algo > 2) return -2;
    ctx->hash_algo = algo;
    return 0;
}
extern "C" int tcb_set_msm_groups(tcb_ctx *ctx, size_t groups) {
    if (!ctx) return -2;
    ctx->msm_groups = groups;
    return 0;
}
extern "C" int tcb_set_eval_split(tcb_ctx *ctx, size_t units) {
    if (!ctx || units > 64) return -2;
    ctx->eval_split = units > 1 ? units : 0;
    ctx->eval_split_off = units == 1;
    return 0;
}
extern "C" int tcb_set_msm_algo(tcb_ctx *ctx, int algo) {
    if (!ctx || algo < 0 || algo > 6) return -2;
    ctx->msm_algo = algo;
    return 0;
}
extern "C" uint64_t tcb_launch_count(const tcb_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ----------------------------------------------------------------------------- _dev API (device pointers, caller's stream)
// The scratch arena is reused from call to call: a call on another stream first waits (on the device) for the end of the
// previous call, so consecutive _dev calls on one ctx may use different streams.
#define DEV_PROLOGUE                                                     \
    if (!ctx) return -2;                                                 \
    DevState &d = ctx->devs[0];                                          \
    cudaStream_t st = (cudaStream_t)stream;                              \
    CK(cudaSetDevice(d.dev));                                            \
    if (arena_reset(ctx, d)) return -1;                                  \
    if (d.has_last) CK(cudaStreamWaitEvent(st, d.last_ev, 0));
#define DEV_RETURN(expr)                                                 \
    do {                                                                 \
        int rc_ = (expr);                                                \
        if (rc_ == 0) { CK(cudaEventRecord(d.last_ev, st)); d.has_last = true; } \
        return rc_;                                                      \
    } while (0)

extern "C" int tcb_verify_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *dd, u8 *ok) {
    DEV_PROLOGUE
    DEV_RETURN(impl_verify_g2(ctx, d, st, n, a, b, c, dd, ok));
}
// the two halves of the pairing check on their own (bench.py times each kernel separately; also the batched form of the
// pairing engine's miller_loop / final_exponentiation pair): f = 576 B per item, Montgomery form, quad-sliced
extern "C" int tcb_miller_loop_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *dd, void *f_out, u8 *enc_ok) {
    DEV_PROLOGUE
    if (n) {
        if (ctx->engine == TCB_ENGINE_QUAD_REG) RUN(run_miller_quad_reg(st, n, a, b, c, dd, f_out, enc_ok));
        else RUN(run_miller_quad(st, n, a, b, c, dd, f_out, enc_ok));
    }
    DEV_RETURN(0);
}
extern "C" int tcb_final_exp_is_one_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const void *f_in, const u8 *enc_ok, u8 *ok) {
    DEV_PROLOGUE
    if (n) {
        if (ctx->engine == TCB_ENGINE_QUAD_SMEM) RUN(run_final_exp_sm(st, n, f_in, enc_ok, ok, nullptr));
        else RUN(run_final_exp_quad(st, n, f_in, enc_ok, ok, nullptr));
    }
    DEV_RETURN(0);
}
extern "C" size_t tcb_miller_value_bytes(void) { return miller_f_bytes(); }
extern "C" int tcb_hash_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    DEV_PROLOGUE
    DEV_RETURN(impl_hash_g2(ctx, d, st, n, msgs, off, out));
}
extern "C" int tcb_verifier_hash_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    DEV_PROLOGUE
    DEV_RETURN(impl_hash_g2(ctx, d, st, n, msgs, off, out, ctx->verify_exact_hash || ctx->engine == TCB_ENGINE_QUAD_REG));
}
extern "C" int tcb_verifier_generator(const tcb_ctx *ctx, u8 *out_g1) {
    if (!ctx || !out_g1) return -2;
    const bool scaled = !ctx->verify_exact_hash && ctx->engine != TCB_ENGINE_QUAD_REG;
    Aff<Fp> g;
    g.x = scaled ? h_consts.g1cx : h_consts.g1x; g.y = scaled ? h_consts.g1cy : h_consts.g1y; g.inf = false;
    store_g1(out_g1, g);
    return 0;
}
extern "C" int tcb_verify_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    DEV_PROLOGUE
    DEV_RETURN(impl_verify(ctx, d, st, n, pk, sig, msgs, off, ok));
}
extern "C" int tcb_sign_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    DEV_PROLOGUE
    DEV_RETURN(impl_sign(ctx, d, st, n, sk, msgs, off, h, out));
}
extern "C" int tcb_combine_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    DEV_PROLOGUE
    DEV_RETURN(impl_combine_g2(ctx, d, st, n, t, x, shares, out, status));
}
extern "C" int tcb_combine_g1_batch_dev(tcb_ctx *ctx, void *stream, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    DEV_PROLOGUE
    DEV_RETURN(impl_combine_g1(ctx, d, st, n, t, x, shares, out, status, 0, nullptr, nullptr));
}
extern "C" int tcb_decrypt_batch_dev(tcb_ctx *ctx, void *stream, size_t n, size_t t, const u8 *x, const u8 *shares, const u8 *v,
                                     const u64 *voff, u64 v_total, u8 *out, u8 *status) {
    DEV_PROLOGUE
    (void)v_total;
    DEV_RETURN(impl_combine_g1(ctx, d, st, n, t, x, shares, out, status, 1, v, voff));
}
extern "C" int tcb_g1_mul_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *sk, const u8 *pts, u8 *out) {
    DEV_PROLOGUE
    if (n) RUN(run_g1_mul(st, n, sk, pts, out));
    DEV_RETURN(0);
}
extern "C" int tcb_commitment_eval_batch_dev(tcb_ctx *ctx, void *stream, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    DEV_PROLOGUE
    DEV_RETURN(impl_commit_eval(ctx, d, st, deg, coeff, n, x, out));
}

// ----------------------------------------------------------------------------- host-buffer API: shard over ctx's devices
// Every item is independent: device g gets the contiguous slice [lo_g, hi_g) of
// each per-item array; H2D, kernels and D2H are queued per device, then all are awaited.
struct Slice { size_t lo, hi; };

template <class T>
static T *up(tcb_ctx *ctx, DevState &d, const T *host, size_t count) {
    T *p = (T *)arena_alloc(ctx, d, count * sizeof(T));
    if (!p) return nullptr;
    if (count && cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, d.cur_stream ? d.cur_stream : d.stream) != cudaSuccess) {
        ctx->err = "H2D copy failed";
        return nullptr;
    }
    return p;
}
// Results are copied back in a SECOND pass (sync_all), after every device and chunk has its uploads and kernels queued: a
// device-to-host copy into pageable memory blocks the calling thread until that stream has drained, which would otherwise run
// the devices one after the other.  With pinned caller buffers every copy is asynchronous.
static int down(tcb_ctx *ctx, DevState &d, void *host, const void *dev, size_t bytes) {
    if (bytes) ctx->pending.push_back({(size_t)(&d - ctx->devs.data()), d.cur_stream ? d.cur_stream : d.stream, host, dev, bytes});
    return 0;
}
static int sync_all(tcb_ctx *ctx) {
    int rc = 0;
    for (auto &dl : ctx->pending) {
        cudaSetDevice(ctx->devs[dl.g].dev);
        cudaError_t e = cudaMemcpyAsync(dl.host, dl.dev, dl.bytes, cudaMemcpyDeviceToHost, dl.st);
        if (e != cudaSuccess) { ctx->err = std::string("D2H copy: ") + cudaGetErrorString(e); rc = -1; }
    }
    ctx->pending.clear();
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.dev);
        for (cudaStream_t st : {d.stream, d.stream2}) {
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { ctx->err = std::string("stream sync: ") + cudaGetErrorString(e); rc = -1; }
        }
        d.cur_stream = nullptr;
    }
    cudaSetDevice(ctx->devs[0].dev);
    return rc;
}
// A host batch is cut into pieces: one contiguous slice per device, and a large slice into up to four chunks that
// alternate between the device's two streams.
struct Piece { size_t g, lo, hi; int sidx; };
static std::vector<Piece> make_pieces(tcb_ctx *ctx, size_t n, size_t grain) {
    std::vector<Piece> out;
    size_t G = ctx->devs.size();
    for (size_t g = 0; g < G; g++) {
        size_t lo = n * g / G, hi = n * (g + 1) / G, cnt = hi - lo;
        if (!cnt) continue;
        size_t chunks = grain ? cnt / grain : 1;
        chunks = chunks < 1 ? 1 : (chunks > 4 ? 4 : chunks);
        for (size_t k = 0; k < chunks; k++) out.push_back({g, lo + cnt * k / chunks, lo + cnt * (k + 1) / chunks, (int)(k & 1)});
    }
    return out;
}
// upload the message slice [lo,hi) with rebased offsets; keeps the rebased offsets alive in `keep`
static int up_msgs(tcb_ctx *ctx, DevState &d, const u8 *msgs, const u64 *off, size_t lo, size_t hi,
                   std::vector<std::vector<u64>> &keep, u8 *&d_msgs, u64 *&d_off) {
    keep.emplace_back(hi - lo + 1);
    std::vector<u64> &o = keep.back();
    for (size_t j = 0; j <= hi - lo; j++) o[j] = off[lo + j] - off[lo];
    size_t bytes = (size_t)o[hi - lo];
    d_msgs = up(ctx, d, msgs + off[lo], bytes ? bytes : 1);
    d_off = up(ctx, d, o.data(), o.size());
    return (d_msgs && d_off) ? 0 : -1;
}
// GRAIN: items per chunk below which a device's slice is not split (0 = never split)
#define HOST_PROLOGUE_G(GRAIN)                                                  \
    if (!ctx) return -2;                                                        \
    ctx->pending.clear();                                                       \
    for (auto &dv_ : ctx->devs) {                                               \
        CK(cudaSetDevice(dv_.dev));                                             \
        if (arena_reset(ctx, dv_)) return -1;                                   \
        if (dv_.has_last) {                                                     \
            CK(cudaStreamWaitEvent(dv_.stream, dv_.last_ev, 0));                \
            CK(cudaStreamWaitEvent(dv_.stream2, dv_.last_ev, 0));               \
        }                                                                       \
    }                                                                           \
    std::vector<Piece> pieces = make_pieces(ctx, n, GRAIN);
#define HOST_PROLOGUE HOST_PROLOGUE_G(0)
#define FOR_EACH_DEV                                        \
    for (const Piece &pc_ : pieces) {                       \
        DevState &d = ctx->devs[pc_.g];                     \
        Slice s{pc_.lo, pc_.hi};                            \
        size_t cnt = s.hi - s.lo;                           \
        CK(cudaSetDevice(d.dev));                           \
        cudaStream_t st = pc_.sidx ? d.stream2 : d.stream;  \
        d.cur_stream = st;
#define END_FOR_EACH_DEV }

extern "C" int tcb_verify_g2_batch(tcb_ctx *ctx, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *dd, u8 *ok) {
    HOST_PROLOGUE
    FOR_EACH_DEV
        u8 *da = up(ctx, d, a + 96 * s.lo, 96 * cnt), *db = up(ctx, d, b + 192 * s.lo, 192 * cnt);
        u8 *dc = c ? up(ctx, d, c + 96 * s.lo, 96 * cnt) : nullptr, *ddv = up(ctx, d, dd + 192 * s.lo, 192 * cnt);
        u8 *dok = (u8 *)arena_alloc(ctx, d, cnt);
        if (!da || !db || (c && !dc) || !ddv || !dok) return -1;
        if (impl_verify_g2(ctx, d, st, cnt, da, db, dc, ddv, dok)) return -1;
        if (down(ctx, d, ok + s.lo, dok, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_hash_g2_batch(tcb_ctx *ctx, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(pieces.size());
    FOR_EACH_DEV
        u8 *dm; u64 *doff;
        if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dout = (u8 *)arena_alloc(ctx, d, 192 * cnt);
        if (!dout) return -1;
        if (impl_hash_g2(ctx, d, st, cnt, dm, doff, dout)) return -1;
        if (down(ctx, d, out + 192 * s.lo, dout, 192 * cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_hash_g1_g2_batch(tcb_ctx *ctx, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, u8 *out) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(pieces.size());
    FOR_EACH_DEV
        u8 *dm; u64 *doff;
        if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dg = up(ctx, d, g1 + 96 * s.lo, 96 * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, 192 * cnt);
        if (!dg || !dout) return -1;
        if (impl_hash_g1_g2(ctx, d, st, cnt, dg, dm, doff, dout)) return -1;
        if (down(ctx, d, out + 192 * s.lo, dout, 192 * cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_verify_batch(tcb_ctx *ctx, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(pieces.size());
    FOR_EACH_DEV
        u8 *dm; u64 *doff;
        if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dpk = up(ctx, d, pk + 96 * s.lo, 96 * cnt), *dsig = up(ctx, d, sig + 192 * s.lo, 192 * cnt);
        u8 *dok = (u8 *)arena_alloc(ctx, d, cnt);
        if (!dpk || !dsig || !dok) return -1;
        if (impl_verify(ctx, d, st, cnt, dpk, dsig, dm, doff, dok)) return -1;
        if (down(ctx, d, ok + s.lo, dok, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
static int sign_common(tcb_ctx *ctx, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(pieces.size());
    FOR_EACH_DEV
        u8 *dm = nullptr, *dh = nullptr; u64 *doff = nullptr;
        if (h) { dh = up(ctx, d, h + 192 * s.lo, 192 * cnt); if (!dh) return -1; }
        else if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dsk = up(ctx, d, sk + 32 * s.lo, 32 * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, 192 * cnt);
        if (!dsk || !dout) return -1;
        if (impl_sign(ctx, d, st, cnt, dsk, dm, doff, dh, dout)) return -1;
        if (down(ctx, d, out + 192 * s.lo, dout, 192 * cnt)) return -1;
        CK(cudaMemsetAsync(dsk, 0, 32 * cnt, st));   // wipe the secret scalars from scratch
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_sign_batch(tcb_ctx *ctx, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, u8 *out) {
    return sign_common(ctx, n, sk, msgs, off, nullptr, out);
}
extern "C" int tcb_sign_g2_batch(tcb_ctx *ctx, size_t n, const u8 *sk, const u8 *h, u8 *out) {
    return sign_common(ctx, n, sk, nullptr, nullptr, h, out);
}
extern "C" int tcb_combine_g2_batch(tcb_ctx *ctx, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    HOST_PROLOGUE_G(16384)
    size_t m = t + 1;
    FOR_EACH_DEV
        u8 *dx = up(ctx, d, x + 32 * m * s.lo, 32 * m * cnt), *dsh = up(ctx, d, shares + 192 * m * s.lo, 192 * m * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, 192 * cnt), *dst = (u8 *)arena_alloc(ctx, d, cnt);
        if (!dx || !dsh || !dout || !dst) return -1;
        if (impl_combine_g2(ctx, d, st, cnt, t, dx, dsh, dout, dst)) return -1;
        if (down(ctx, d, out + 192 * s.lo, dout, 192 * cnt) || down(ctx, d, status + s.lo, dst, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_combine_g1_batch(tcb_ctx *ctx, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    HOST_PROLOGUE_G(8192)
    size_t m = t + 1;
    FOR_EACH_DEV
        u8 *dx = up(ctx, d, x + 32 * m * s.lo, 32 * m * cnt), *dsh = up(ctx, d, shares + 96 * m * s.lo, 96 * m * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, 96 * cnt), *dst = (u8 *)arena_alloc(ctx, d, cnt);
        if (!dx || !dsh || !dout || !dst) return -1;
        if (impl_combine_g1(ctx, d, st, cnt, t, dx, dsh, dout, dst, 0, nullptr, nullptr)) return -1;
        if (down(ctx, d, out + 96 * s.lo, dout, 96 * cnt) || down(ctx, d, status + s.lo, dst, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_decrypt_batch(tcb_ctx *ctx, size_t n, size_t t, const u8 *x, const u8 *shares, const u8 *v, const u64 *voff,
                                 u8 *out, u8 *status) {
    HOST_PROLOGUE_G(8192)
    size_t m = t + 1;
    std::vector<std::vector<u64>> keep;
    keep.reserve(pieces.size());
    FOR_EACH_DEV
        u8 *dv; u64 *dvoff;
        if (up_msgs(ctx, d, v, voff, s.lo, s.hi, keep, dv, dvoff)) return -1;
        size_t vbytes = (size_t)(voff[s.hi] - voff[s.lo]);
        u8 *dx = up(ctx, d, x + 32 * m * s.lo, 32 * m * cnt), *dsh = up(ctx, d, shares + 96 * m * s.lo, 96 * m * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, vbytes), *dst = (u8 *)arena_alloc(ctx, d, cnt);
        if (!dx || !dsh || !dout || !dst) return -1;
        if (impl_combine_g1(ctx, d, st, cnt, t, dx, dsh, dout, dst, 1, dv, dvoff)) return -1;
        if (down(ctx, d, out + voff[s.lo], dout, vbytes) || down(ctx, d, status + s.lo, dst, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
static int g1_mul_common(tcb_ctx *ctx, size_t n, const u8 *sk, const u8 *pts, u8 *out) {
    HOST_PROLOGUE
    FOR_EACH_DEV
        u8 *dsk = up(ctx, d, sk + 32 * s.lo, 32 * cnt);
        u8 *dp = pts ? up(ctx, d, pts + 96 * s.lo, 96 * cnt) : nullptr;
        u8 *dout = (u8 *)arena_alloc(ctx, d, 96 * cnt);
        if (!dsk || (pts && !dp) || !dout) return -1;
        RUN(run_g1_mul(st, cnt, dsk, dp, dout));
        if (down(ctx, d, out + 96 * s.lo, dout, 96 * cnt)) return -1;
        CK(cudaMemsetAsync(dsk, 0, 32 * cnt, st));
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_decrypt_share_batch(tcb_ctx *ctx, size_t n, const u8 *sk, const u8 *u, u8 *out) {
    if (!u) { if (ctx) ctx->err = "u_g1 is NULL"; return -2; }
    return g1_mul_common(ctx, n, sk, u, out);
}
extern "C" int tcb_g1_mul_gen_batch(tcb_ctx *ctx, size_t n, const u8 *sk, u8 *out) { return g1_mul_common(ctx, n, sk, nullptr, out); }
extern "C" int tcb_commitment_eval_batch(tcb_ctx *ctx, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    HOST_PROLOGUE
    FOR_EACH_DEV
        u8 *dc = up(ctx, d, coeff, 96 * (deg + 1));   // the commitment table is broadcast to every device
        u8 *dx = up(ctx, d, x + 32 * s.lo, 32 * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, 96 * cnt);
        if (!dc || !dx || !dout) return -1;
        if (impl_commit_eval(ctx, d, st, deg, dc, cnt, dx, dout)) return -1;
        if (down(ctx, d, out + 96 * s.lo, dout, 96 * cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}

// ---- out_i = sum_k s_{i,k} P_{i,k}  (BivarCommitment::row / evaluate, src/poly.rs:693-726:
// the caller supplies the Fr power products as scalars).  Same per-term + sum kernels as interpolate.
static int lincomb_common(tcb_ctx *ctx, size_t n, size_t m, const u8 *scalars, const u8 *pts, u8 *out, int g2) {
    HOST_PROLOGUE_G(8192)
    size_t pw = g2 ? 192 : 96;
    if (m == 0) { if (ctx) ctx->err = "m must be >= 1"; return -2; }
    std::vector<u8> hstatus(n, 0);
    FOR_EACH_DEV
        u8 *dsc = up(ctx, d, scalars + 32 * m * s.lo, 32 * m * cnt), *dp = up(ctx, d, pts + pw * m * s.lo, pw * m * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, pw * cnt), *dst = (u8 *)arena_alloc(ctx, d, cnt);
        if (!dsc || !dp || !dout || !dst) return -1;
        CK(cudaMemsetAsync(dst, 0, cnt, st));
        // canonical little-endian scalars are exactly the limb layout the recoding reads
        void *part; size_t Gp;
        if (g2) { if (impl_msm_g2(ctx, d, st, cnt, m, (const u32 *)dsc, dp, dst, part, Gp)) return -1; RUN(run_g2_sum(st, cnt, Gp, part, dout)); }
        else { if (impl_msm_g1(ctx, d, st, cnt, m, (const u32 *)dsc, dp, dst, part, Gp)) return -1; RUN(run_g1_sum(st, cnt, Gp, part, dout)); }
        if (down(ctx, d, out + pw * s.lo, dout, pw * cnt) || down(ctx, d, hstatus.data() + s.lo, dst, cnt)) return -1;
    END_FOR_EACH_DEV
    if (sync_all(ctx)) return -1;
    for (size_t i = 0; i < n; i++)
        if (hstatus[i]) { ctx->err = "lincomb: item " + std::to_string(i) + " holds a point whose field element is >= p (invalid encoding)"; return -10; }
    return 0;
}
extern "C" int tcb_g1_lincomb_batch(tcb_ctx *ctx, size_t n, size_t m, const u8 *scalars, const u8 *pts, u8 *out) { return lincomb_common(ctx, n, m, scalars, pts, out, 0); }
extern "C" int tcb_g2_lincomb_batch(tcb_ctx *ctx, size_t n, size_t m, const u8 *scalars, const u8 *pts, u8 *out) { return lincomb_common(ctx, n, m, scalars, pts, out, 1); }

// ---- Fr-side Poly algebra (src/poly.rs:173-194, 358-369)
extern "C" int tcb_poly_eval_batch(tcb_ctx *ctx, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    HOST_PROLOGUE
    std::vector<u8> bad(pieces.size() + 1, 0);
    size_t pi = 0;
    FOR_EACH_DEV
        u8 *dc = up(ctx, d, coeff, 32 * (deg + 1)), *dx = up(ctx, d, x + 32 * s.lo, 32 * cnt);
        void *cm = arena_alloc(ctx, d, (deg + 1) * fr_bytes());
        u8 *dout = (u8 *)arena_alloc(ctx, d, 32 * cnt), *dbad = (u8 *)arena_alloc(ctx, d, 4);
        if (!dc || !dx || !cm || !dout || !dbad) return -1;
        CK(cudaMemsetAsync(dbad, 0, 4, st));
        RUN(run_fr_to_mont(st, deg + 1, dc, cm, dbad));
        RUN(run_poly_eval(st, cnt, deg, cm, dx, dout, dbad));
        if (down(ctx, d, out + 32 * s.lo, dout, 32 * cnt) || down(ctx, d, &bad[pi++], dbad, 1)) return -1;
    END_FOR_EACH_DEV
    if (sync_all(ctx)) return -1;
    for (u8 b : bad) if (b) { ctx->err = "poly_eval: a scalar is not a canonical Fr (>= r)"; return -10; }
    return 0;
}
extern "C" int tcb_poly_mul_batch(tcb_ctx *ctx, size_t n, size_t da, const u8 *a, size_t db, const u8 *b, u8 *out) {
    HOST_PROLOGUE
    std::vector<u8> bad(pieces.size() + 1, 0);
    size_t pi = 0, w = da + db + 1;
    FOR_EACH_DEV
        u8 *da_ = up(ctx, d, a + 32 * (da + 1) * s.lo, 32 * (da + 1) * cnt), *db_ = up(ctx, d, b + 32 * (db + 1) * s.lo, 32 * (db + 1) * cnt);
        void *am = arena_alloc(ctx, d, (da + 1) * cnt * fr_bytes()), *bm = arena_alloc(ctx, d, (db + 1) * cnt * fr_bytes());
        u8 *dout = (u8 *)arena_alloc(ctx, d, 32 * w * cnt), *dbad = (u8 *)arena_alloc(ctx, d, 4);
        if (!da_ || !db_ || !am || !bm || !dout || !dbad) return -1;
        CK(cudaMemsetAsync(dbad, 0, 4, st));
        RUN(run_fr_to_mont(st, (da + 1) * cnt, da_, am, dbad));
        RUN(run_fr_to_mont(st, (db + 1) * cnt, db_, bm, dbad));
        RUN(run_poly_mul(st, cnt, da, db, am, bm, dout));
        if (down(ctx, d, out + 32 * w * s.lo, dout, 32 * w * cnt) || down(ctx, d, &bad[pi++], dbad, 1)) return -1;
    END_FOR_EACH_DEV
    if (sync_all(ctx)) return -1;
    for (u8 x : bad) if (x) { ctx->err = "poly_mul: a coefficient is not a canonical Fr (>= r)"; return -10; }
    return 0;
}

// ---- PublicKey::encrypt_with_rng with caller-supplied r (src/lib.rs:128-137)
extern "C" int tcb_encrypt_batch(tcb_ctx *ctx, size_t n, const u8 *pk, const u8 *r, const u8 *msgs, const u64 *off,
                                 u8 *u_out, u8 *v_out, u8 *w_out) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(pieces.size());
    FOR_EACH_DEV
        u8 *dm; u64 *doff;
        if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        size_t mbytes = (size_t)(off[s.hi] - off[s.lo]);
        u8 *dpk = up(ctx, d, pk + 96 * s.lo, 96 * cnt), *dr = up(ctx, d, r + 32 * s.lo, 32 * cnt);
        u8 *du = (u8 *)arena_alloc(ctx, d, 96 * cnt), *dv = (u8 *)arena_alloc(ctx, d, mbytes);
        u8 *dh = (u8 *)arena_alloc(ctx, d, 192 * cnt), *dw = (u8 *)arena_alloc(ctx, d, 192 * cnt);
        if (!dpk || !dr || !du || !dv || !dh || !dw) return -1;
        RUN(run_encrypt_uv(st, cnt, dpk, dr, dm, doff, du, dv));
        if (impl_hash_g1_g2(ctx, d, st, cnt, du, dv, doff, dh)) return -1;
        RUN(run_sign(st, cnt, dr, nullptr, nullptr, dh, dw));
        if (down(ctx, d, u_out + 96 * s.lo, du, 96 * cnt) || down(ctx, d, v_out + off[s.lo], dv, mbytes) ||
            down(ctx, d, w_out + 192 * s.lo, dw, 192 * cnt)) return -1;
        CK(cudaMemsetAsync(dr, 0, 32 * cnt, st));     // the encryption randomness is secret
    END_FOR_EACH_DEV
    return sync_all(ctx);
}

// ---- batched point (de)compression
static int codec_common(tcb_ctx *ctx, size_t n, const u8 *in, size_t in_w, u8 *out, size_t out_w, u8 *status, int which) {
    HOST_PROLOGUE
    FOR_EACH_DEV
        u8 *din = up(ctx, d, in + in_w * s.lo, in_w * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, out_w * cnt), *dst = (u8 *)arena_alloc(ctx, d, cnt);
        if (!din || !dout || !dst) return -1;
        switch (which) {
            case 0: RUN(run_g1_compress(st, cnt, din, dout)); break;
            case 1: RUN(run_g2_compress(st, cnt, din, dout)); break;
            case 2: RUN(run_g1_decompress(st, cnt, din, dout, dst)); break;
            default: RUN(run_g2_decompress(st, cnt, din, dout, dst)); break;
        }
        if (down(ctx, d, out + out_w * s.lo, dout, out_w * cnt)) return -1;
        if (status && down(ctx, d, status + s.lo, dst, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_g1_compress_batch(tcb_ctx *ctx, size_t n, const u8 *unc, u8 *out48) { return codec_common(ctx, n, unc, 96, out48, 48, nullptr, 0); }
extern "C" int tcb_g2_compress_batch(tcb_ctx *ctx, size_t n, const u8 *unc, u8 *out96) { return codec_common(ctx, n, unc, 192, out96, 96, nullptr, 1); }
extern "C" int tcb_g1_decompress_batch(tcb_ctx *ctx, size_t n, const u8 *in48, u8 *out96, u8 *status) { return codec_common(ctx, n, in48, 48, out96, 96, status, 2); }
extern "C" int tcb_g2_decompress_batch(tcb_ctx *ctx, size_t n, const u8 *in96, u8 *out192, u8 *status) { return codec_common(ctx, n, in96, 96, out192, 192, status, 3); }

// ----------------------------------------------------------------------------- self-test / probes
extern "C" int tcb_selftest_fp(tcb_ctx *ctx, size_t n, uint64_t seed) {
    if (!ctx) return -2;
    DevState &d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    unsigned long long *bad = nullptr, h = 0;
    CK(cudaMalloc(&bad, 8));
    CK(cudaMemsetAsync(bad, 0, 8, d.stream));
    n = (n + 127) & ~(size_t)127;
    run_selftest(d.stream, n, seed, bad);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h, bad, 8, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    cudaFree(bad);
    return (int)(h > 0x7fffffff ? 0x7fffffff : h);
}
// Miller loop of the shared-memory engine against the register engine's on the caller's items: number of differing 32-bit
// words of f and of the encoding flags (0 = bit-identical)
extern "C" int tcb_selftest_miller(tcb_ctx *ctx, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *dd) {
    if (!ctx) return -2;
    DevState &d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    if (arena_reset(ctx, d)) return -1;
    if (!n) return 0;
    cudaStream_t st = d.stream;
    u8 *da = up(ctx, d, a, 96 * n), *db = up(ctx, d, b, 192 * n), *dc = c ? up(ctx, d, c, 96 * n) : nullptr, *ddv = up(ctx, d, dd, 192 * n);
    size_t fb = n * miller_f_bytes(), eb = (n + 3) & ~(size_t)3;
    void *f1 = arena_alloc(ctx, d, fb), *f2 = arena_alloc(ctx, d, fb);
    u8 *e1 = (u8 *)arena_alloc(ctx, d, eb), *e2 = (u8 *)arena_alloc(ctx, d, eb);
    unsigned long long *bad = (unsigned long long *)arena_alloc(ctx, d, 8), h = 0;
    if (!da || !db || (c && !dc) || !ddv || !f1 || !f2 || !e1 || !e2 || !bad) return -1;
    CK(cudaMemsetAsync(bad, 0, 8, st));
    CK(cudaMemsetAsync(e1, 0, eb, st));
    CK(cudaMemsetAsync(e2, 0, eb, st));
    RUN(run_miller_quad(st, n, da, db, dc, ddv, f1, e1));
    RUN(run_miller_quad_reg(st, n, da, db, dc, ddv, f2, e2));
    RUN(run_count_diff(st, fb, f1, f2, bad));
    RUN(run_count_diff(st, eb, e1, e2, bad));
    // the two final-exponentiation kernels on the same Miller values: the full Fp12 results and the booleans must agree as well
    void *g1 = arena_alloc(ctx, d, fb), *g2 = arena_alloc(ctx, d, fb);
    u8 *o1 = (u8 *)arena_alloc(ctx, d, eb), *o2 = (u8 *)arena_alloc(ctx, d, eb);
    if (!g1 || !g2 || !o1 || !o2) return -1;
    CK(cudaMemsetAsync(o1, 0, eb, st));
    CK(cudaMemsetAsync(o2, 0, eb, st));
    RUN(run_final_exp_sm(st, n, f1, e1, o1, g1));
    RUN(run_final_exp_quad(st, n, f1, e1, o2, g2));
    RUN(run_count_diff(st, fb, g1, g2, bad));
    RUN(run_count_diff(st, eb, o1, o2, bad));
    CK(cudaMemcpyAsync(&h, bad, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (int)(h > 0x7fffffff ? 0x7fffffff : h);
}
template <class F>
static int timed(tcb_ctx *ctx, DevState &d, F &&enqueue, int reps, float &ms_best) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    ms_best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaEventRecord(e0, d.stream));
        enqueue();
        CK(cudaEventRecord(e1, d.stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < ms_best) ms_best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
extern "C" int tcb_probe_imad(tcb_ctx *ctx, double *macs_per_sec) {
    if (!ctx || !macs_per_sec) return -2;
    DevState &d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 2048;
    u64 *out = nullptr;
    CK(cudaMalloc(&out, (size_t)blocks * threads * 8));
    float ms = 0;
    if (timed(ctx, d, [&] { run_probe_imad(d.stream, blocks, threads, out, iters); ctx->launches++; }, 3, ms)) return -1;
    cudaFree(out);
    *macs_per_sec = (double)blocks * threads * iters * 64.0 / (ms * 1e-3);
    return 0;
}
extern "C" int tcb_probe_fpmul(tcb_ctx *ctx, double *muls_per_sec) {
    if (!ctx || !muls_per_sec) return -2;
    DevState &d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    const int blocks = ctx->sm_count * 4, threads = 256, iters = 4096;
    void *out = nullptr;
    CK(cudaMalloc(&out, (size_t)blocks * threads * fp_bytes()));
    float ms = 0;
    if (timed(ctx, d, [&] { run_probe_fpmul(d.stream, blocks, threads, out, iters); ctx->launches++; }, 3, ms)) return -1;
    cudaFree(out);
    *muls_per_sec = (double)blocks * threads * iters * 2.0 / (ms * 1e-3);
    return 0;
}
