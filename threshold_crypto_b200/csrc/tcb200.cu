// tcb200.cu — sm_100a kernels and the extern "C" boundary (include/tcb200.h) of the
// B200-native threshold_crypto hot path.  No CPU fallback: every entry point launches CUDA
// kernels or fails with an error code.
//
// Kernel inventory (one unit = one thread, or one lane pair for the Fp2S engine):
//   k_verify_g2 / k_verify      a1/a3  pairing equality (+ on-device hash_g2)
//   k_hash_g2                   a2     SHA3 -> ChaCha20 -> G2::random -> x h2
//   k_sign                      a4     sk * H(m)
//   k_lagrange                  a5     lambda_i(0), one thread per (item, share)
//   k_g2_mul_store / k_g2_sum   a6     per-share terms and their sum (combine_signatures)
//   k_g1_mul / k_g1_mul_store / k_g1_sum / k_decrypt_finish   a7 (decrypt shares, decrypt)
//   k_g1_decode / k_commit_eval a8     Commitment::evaluate (Horner)
//   k_selftest_fp, k_probe_*    measurement / self-test
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include "../../include/tcb200.h"
#include "scheme.cuh"

using namespace tcb;

// ----------------------------------------------------------------------------- kernels
template <class F2> __device__ __forceinline__ size_t unit_index() {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    return F2::SLICED ? (t >> 1) : t;
}
template <class F2>
__global__ void __launch_bounds__(128) k_verify_g2(size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok) {
    size_t i = unit_index<F2>();
    if (i < n) task_verify_g2<F2>(i, a, b, c, d, ok);
}
template <class F2>
__global__ void __launch_bounds__(128) k_hash_g2(size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    size_t i = unit_index<F2>();
    if (i < n) task_hash_g2<F2>(i, msgs, off, out);
}
template <class F2>
__global__ void __launch_bounds__(128) k_verify(size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    size_t i = unit_index<F2>();
    if (i < n) task_verify<F2>(i, pk, sig, msgs, off, ok);
}
template <class F2>
__global__ void __launch_bounds__(128) k_sign(size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    size_t i = unit_index<F2>();
    if (i < n) task_sign<F2>(i, sk, msgs, off, h, out);
}
__global__ void __launch_bounds__(128) k_lagrange(size_t n, size_t m, const u8 *xs, u32 *lam, u8 *status) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    size_t item = t / m, i = t % m;
    u8 st = 0;
    lagrange_coeff(xs + item * m * 32, m, i, lam + 8 * t, st);
    if (st) status[item] = st;
}
template <class F2>
__global__ void __launch_bounds__(128) k_g2_mul_store(size_t units, const u32 *k, const u8 *pts, JacStore<F2> *out, u8 *status, size_t per_item) {
    size_t i = unit_index<F2>();
    if (i < units) task_g2_mul_store<F2>(i, k, pts, out, status, per_item);
}
template <class F2>
__global__ void __launch_bounds__(128) k_g2_sum(size_t n, size_t m, const JacStore<F2> *terms, u8 *out) {
    size_t i = unit_index<F2>();
    if (i < n) task_g2_sum<F2>(i, m, terms, out);
}
__global__ void __launch_bounds__(128) k_g1_mul(size_t n, const u8 *sk, const u8 *pts, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g1_mul(i, sk, pts, out);
}
__global__ void __launch_bounds__(128) k_g1_mul_store(size_t units, const u32 *k, const u8 *pts, Jac1Store *out, u8 *status, size_t per_item) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < units) task_g1_mul_store(i, k, pts, out, status, per_item);
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
__global__ void __launch_bounds__(128) k_g1_decode(size_t n, const u8 *pts, Jac1Store *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_g1_decode(i, pts, out);
}
__global__ void __launch_bounds__(128) k_commit_eval(size_t n, size_t deg, const Jac1Store *coeff, const u8 *x, u8 *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) task_commit_eval(i, deg, coeff, x, out);
}

// ---- self-test and roofline probes
__device__ __forceinline__ u64 splitmix(u64 &s) {
    u64 z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ Fp rand_fp(u64 &s, int mode) {
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
__global__ void k_selftest_fp(size_t n, u64 seed, unsigned long long *bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s = seed + i * 0x632be59bd9b4e019ULL;
    int mode_a = i < 64 ? (int)(i & 3) : 0, mode_b = i < 64 ? (int)((i >> 2) & 3) : 0;
    Fp a = rand_fp(s, mode_a), b = rand_fp(s, mode_b), c = rand_fp(s, 0), d = rand_fp(s, mode_a);
    int errs = 0;
    Fp ref = mont_mul_portable<FpParams>(a, b);
    if ((a * b) != ref) errs++;
    if (sqr(a) != mont_mul_portable<FpParams>(a, a)) errs++;
    Fp d2 = dot2(a, b, c, d);
    if (d2 != (ref + mont_mul_portable<FpParams>(c, d))) errs++;
    if (((a + b) - b) != a) errs++;
    if (!(a + (-a)).is_zero()) errs++;
    if (!limbs_lt_mod<FpParams>((a + b).l) || !limbs_lt_mod<FpParams>((a - b).l) || !limbs_lt_mod<FpParams>(d2.l)) errs++;
    // sliced Fp2 against scalar Fp2: lanes of a pair share (a,b,c,d) of the even lane
    {
        u32 m = 3u << (threadIdx.x & 30u);
        Fp2 x, y;
        x.c0 = a; x.c1 = b; y.c0 = c; y.c1 = d;
        for (int k = 0; k < 12; k++) {
            x.c0.l[k] = __shfl_sync(m, x.c0.l[k], threadIdx.x & 30u); x.c1.l[k] = __shfl_sync(m, x.c1.l[k], threadIdx.x & 30u);
            y.c0.l[k] = __shfl_sync(m, y.c0.l[k], threadIdx.x & 30u); y.c1.l[k] = __shfl_sync(m, y.c1.l[k], threadIdx.x & 30u);
        }
        Fp2S xs = Fp2S::from_halves(x.c0, x.c1), ys = Fp2S::from_halves(y.c0, y.c1);
        Fp2 pm = x * y, ps = sqr(x), px = mul_xi(x);
        Fp2S qm = xs * ys, qs = sqr(xs), qx = mul_xi(xs);
        bool role = threadIdx.x & 1;
        if (qm.h != (role ? pm.c1 : pm.c0)) errs++;
        if (qs.h != (role ? ps.c1 : ps.c0)) errs++;
        if (qx.h != (role ? px.c1 : px.c0)) errs++;
        // Fp2 mul against the schoolbook with portable multiplies
        Fp t0 = mont_mul_portable<FpParams>(x.c0, y.c0) - mont_mul_portable<FpParams>(x.c1, y.c1);
        Fp t1 = mont_mul_portable<FpParams>(x.c0, y.c1) + mont_mul_portable<FpParams>(x.c1, y.c0);
        if (pm.c0 != t0 || pm.c1 != t1) errs++;
        // predicates and rarely-used ops of the sliced engine against the scalar one
        Fp2 z0 = x; z0.c1 = Fp::zero();          // only one half zero: exercises pair_and
        Fp2S z0s = Fp2S::from_halves(z0.c0, z0.c1);
        if (is_zero(z0s) != is_zero(z0)) errs++;
        if (is_zero(Fp2S::zero()) != true) errs++;
        if (eq(xs, ys) != eq(x, y) || !eq(xs, xs)) errs++;
        if (eq(z0s, xs) != eq(z0, x)) errs++;
        Fp2 pc = conj(x), pi = inv(x), pn = -x, pf = mul_fp(x, c);
        Fp2S qc = conj(xs), qi = inv(xs), qn = -xs, qf = mul_fp(xs, c);
        if (qc.h != (role ? pc.c1 : pc.c0)) errs++;
        if (qi.h != (role ? pi.c1 : pi.c0)) errs++;
        if (qn.h != (role ? pn.c1 : pn.c0)) errs++;
        if (qf.h != (role ? pf.c1 : pf.c0)) errs++;
        if (fp2_cmp(xs, ys) != fp2_cmp(x, y)) errs++;
        if (norm(xs) != norm(x)) errs++;
        Fp2 one_s; Fp2S::one().gather(one_s.c0, one_s.c1);
        if (!eq(one_s, Fp2::one())) errs++;
        if (((i >> 1) & 31) == 0) {   // a few square roots (expensive); the condition is pair-uniform
            Fp2 sq = sqr(x), r1;
            Fp2S r2;
            bool ok1 = fp2_sqrt(r1, sq), ok2 = fp2_sqrt(r2, Fp2S::from_halves(sq.c0, sq.c1));
            if (!ok1 || !ok2) errs++;
            if (r2.h != (role ? r1.c1 : r1.c0)) errs++;
            Fp2 nr;
            bool ok3 = fp2_sqrt(nr, x), ok4 = fp2_sqrt(r2, xs);
            if (ok3 != ok4) errs++;
        }
    }
    if (errs) atomicAdd(bad, (unsigned long long)errs);
}
// 8 independent IMAD.WIDE accumulators per thread, no carries: the integer-MAC ceiling
__global__ void __launch_bounds__(256) k_probe_imad(u64 *out, int iters, u32 seed) {
    u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    u64 acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = (u64)a * (u32)(b + k) + acc[k];
            a += (u32)acc[0];
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2 independent Montgomery-multiply chains per thread
__global__ void __launch_bounds__(256) k_probe_fpmul(Fp *out, int iters, u64 seed) {
    u64 s = seed + threadIdx.x + (u64)blockIdx.x * 1024;
    Fp a = rand_fp(s, 0), b = rand_fp(s, 0), c = rand_fp(s, 0);
    for (int it = 0; it < iters; it++) { a = a * c; b = b * c; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}

// ----------------------------------------------------------------------------- context
struct Chunk { void *p; size_t sz, used; };
struct DevState {
    int dev;
    cudaStream_t stream;
    std::vector<Chunk> chunks;
    size_t high_water = 0, cur = 0;
};
struct tcb_ctx {
    std::vector<DevState> devs;
    std::string err;
    uint64_t launches = 0;
    int engine = TCB_ENGINE_PAIR;
    int sm_count = 148;
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

template <class K, class... A>
static int launch(tcb_ctx *ctx, cudaStream_t st, K kern, size_t threads, A... args) {
    if (threads == 0) return 0;
    const int block = 128;
    size_t grid = (threads + block - 1) / block;
    kern<<<(unsigned)grid, block, 0, st>>>(args...);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
#define LAUNCH_G2(kern, units, ...)                                                                       \
    (ctx->engine == TCB_ENGINE_PAIR ? launch(ctx, st, kern<Fp2S>, (size_t)(units) * 2, __VA_ARGS__)       \
                                    : launch(ctx, st, kern<Fp2>, (size_t)(units), __VA_ARGS__))

// ----------------------------------------------------------------------------- device-side implementations
static int impl_verify_g2(tcb_ctx *ctx, cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok) {
    return LAUNCH_G2(k_verify_g2, n, n, a, b, c, d, ok);
}
static int impl_hash_g2(tcb_ctx *ctx, cudaStream_t st, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    return LAUNCH_G2(k_hash_g2, n, n, msgs, off, out);
}
static int impl_verify(tcb_ctx *ctx, cudaStream_t st, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    return LAUNCH_G2(k_verify, n, n, pk, sig, msgs, off, ok);
}
static int impl_sign(tcb_ctx *ctx, cudaStream_t st, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    return LAUNCH_G2(k_sign, n, n, sk, msgs, off, h, out);
}
static int impl_combine_g2(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    CK(cudaMemsetAsync(status, 0, n, st));
    if (n == 0) return 0;
    if (t == 0) { CK(cudaMemcpyAsync(out, shares, n * 192, cudaMemcpyDeviceToDevice, st)); return 0; }
    size_t m = t + 1;
    u32 *lam = (u32 *)arena_alloc(ctx, d, n * m * 32);
    if (!lam) return -1;
    if (launch(ctx, st, k_lagrange, n * m, n, m, x, lam, status)) return -1;
    if (ctx->engine == TCB_ENGINE_PAIR) {
        JacStore<Fp2S> *terms = (JacStore<Fp2S> *)arena_alloc(ctx, d, n * m * sizeof(JacStore<Fp2S>));
        if (!terms) return -1;
        if (launch(ctx, st, k_g2_mul_store<Fp2S>, n * m * 2, n * m, lam, shares, terms, status, m)) return -1;
        return launch(ctx, st, k_g2_sum<Fp2S>, n * 2, n, m, terms, out);
    }
    JacStore<Fp2> *terms = (JacStore<Fp2> *)arena_alloc(ctx, d, n * m * sizeof(JacStore<Fp2>));
    if (!terms) return -1;
    if (launch(ctx, st, k_g2_mul_store<Fp2>, n * m, n * m, lam, shares, terms, status, m)) return -1;
    return launch(ctx, st, k_g2_sum<Fp2>, n, n, m, terms, out);
}
// mode 0: write the combined G1 point; mode 1: xor_with_hash (decrypt)
static int impl_combine_g1(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t n, size_t t, const u8 *x, const u8 *shares,
                           u8 *out, u8 *status, int mode, const u8 *v, const u64 *voff) {
    CK(cudaMemsetAsync(status, 0, n, st));
    if (n == 0) return 0;
    if (t == 0) {
        if (mode == 0) { CK(cudaMemcpyAsync(out, shares, n * 96, cudaMemcpyDeviceToDevice, st)); return 0; }
        return launch(ctx, st, k_decrypt_finish, n, n, (size_t)1, (const Jac1Store *)nullptr, shares, v, voff, out);
    }
    size_t m = t + 1;
    u32 *lam = (u32 *)arena_alloc(ctx, d, n * m * 32);
    Jac1Store *terms = (Jac1Store *)arena_alloc(ctx, d, n * m * sizeof(Jac1Store));
    if (!lam || !terms) return -1;
    if (launch(ctx, st, k_lagrange, n * m, n, m, x, lam, status)) return -1;
    if (launch(ctx, st, k_g1_mul_store, n * m, n * m, lam, shares, terms, status, m)) return -1;
    if (mode == 0) return launch(ctx, st, k_g1_sum, n, n, m, terms, out);
    return launch(ctx, st, k_decrypt_finish, n, n, m, (const Jac1Store *)terms, shares, v, voff, out);
}
static int impl_commit_eval(tcb_ctx *ctx, DevState &d, cudaStream_t st, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    Jac1Store *tab = (Jac1Store *)arena_alloc(ctx, d, (deg + 1) * sizeof(Jac1Store));
    if (!tab) return -1;
    if (launch(ctx, st, k_g1_decode, deg + 1, deg + 1, coeff, tab)) return -1;
    return launch(ctx, st, k_commit_eval, n, n, deg, (const Jac1Store *)tab, x, out);
}

// ----------------------------------------------------------------------------- init / free
extern "C" int tcb_init(tcb_ctx **out, const int *device_ids, int n_devices) {
    if (!out) return -2;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return -3;   // no CPU fallback
    tcb_ctx *ctx = new tcb_ctx();
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
        if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return -6; }
        if (cudaMemcpyToSymbol(d_consts, &C, sizeof C) != cudaSuccess) { delete ctx; return -7; }
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
        for (auto &c : d.chunks) {
            cudaMemset(c.p, 0, c.sz);   // scratch may have held secret scalars (src/lib.rs:304-314 hygiene)
            cudaFree(c.p);
        }
        cudaStreamDestroy(d.stream);
    }
    delete ctx;
}
extern "C" const char *tcb_last_error(const tcb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int tcb_set_engine(tcb_ctx *ctx, int engine) {
    if (!ctx || (engine != TCB_ENGINE_PAIR && engine != TCB_ENGINE_THREAD)) return -2;
    ctx->engine = engine;
    return 0;
}
extern "C" uint64_t tcb_launch_count(const tcb_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ----------------------------------------------------------------------------- _dev API (device pointers, caller's stream)
#define DEV_PROLOGUE                         \
    if (!ctx) return -2;                     \
    DevState &d = ctx->devs[0];              \
    cudaStream_t st = (cudaStream_t)stream;  \
    CK(cudaSetDevice(d.dev));                \
    if (arena_reset(ctx, d)) return -1;

extern "C" int tcb_verify_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *dd, u8 *ok) {
    DEV_PROLOGUE
    return impl_verify_g2(ctx, st, n, a, b, c, dd, ok);
}
extern "C" int tcb_hash_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    DEV_PROLOGUE
    return impl_hash_g2(ctx, st, n, msgs, off, out);
}
extern "C" int tcb_verify_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    DEV_PROLOGUE
    return impl_verify(ctx, st, n, pk, sig, msgs, off, ok);
}
extern "C" int tcb_sign_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    DEV_PROLOGUE
    return impl_sign(ctx, st, n, sk, msgs, off, h, out);
}
extern "C" int tcb_combine_g2_batch_dev(tcb_ctx *ctx, void *stream, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    DEV_PROLOGUE
    return impl_combine_g2(ctx, d, st, n, t, x, shares, out, status);
}
extern "C" int tcb_combine_g1_batch_dev(tcb_ctx *ctx, void *stream, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    DEV_PROLOGUE
    return impl_combine_g1(ctx, d, st, n, t, x, shares, out, status, 0, nullptr, nullptr);
}
extern "C" int tcb_decrypt_batch_dev(tcb_ctx *ctx, void *stream, size_t n, size_t t, const u8 *x, const u8 *shares, const u8 *v,
                                     const u64 *voff, u64 v_total, u8 *out, u8 *status) {
    DEV_PROLOGUE
    (void)v_total;
    return impl_combine_g1(ctx, d, st, n, t, x, shares, out, status, 1, v, voff);
}
extern "C" int tcb_g1_mul_batch_dev(tcb_ctx *ctx, void *stream, size_t n, const u8 *sk, const u8 *pts, u8 *out) {
    DEV_PROLOGUE
    return launch(ctx, st, k_g1_mul, n, n, sk, pts, out);
}
extern "C" int tcb_commitment_eval_batch_dev(tcb_ctx *ctx, void *stream, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    DEV_PROLOGUE
    return impl_commit_eval(ctx, d, st, deg, coeff, n, x, out);
}

// ----------------------------------------------------------------------------- host-buffer API: shard over ctx's devices
// Every item is independent: device g gets the contiguous slice [lo_g, hi_g) of
// each per-item array; H2D, kernels and D2H are queued per device, then all are awaited.
struct Slice { size_t lo, hi; };
static Slice slice_of(size_t n, size_t g, size_t G) { return {n * g / G, n * (g + 1) / G}; }

template <class T>
static T *up(tcb_ctx *ctx, DevState &d, const T *host, size_t count) {
    T *p = (T *)arena_alloc(ctx, d, count * sizeof(T));
    if (!p) return nullptr;
    if (count && cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) {
        ctx->err = "H2D copy failed";
        return nullptr;
    }
    return p;
}
static int down(tcb_ctx *ctx, DevState &d, void *host, const void *dev, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, d.stream));
    return 0;
}
static int sync_all(tcb_ctx *ctx) {
    int rc = 0;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.dev);
        cudaError_t e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess) { ctx->err = std::string("stream sync: ") + cudaGetErrorString(e); rc = -1; }
    }
    cudaSetDevice(ctx->devs[0].dev);
    return rc;
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
#define HOST_PROLOGUE   \
    if (!ctx) return -2; \
    size_t G = ctx->devs.size();
#define FOR_EACH_DEV                                        \
    for (size_t g = 0; g < G; g++) {                        \
        DevState &d = ctx->devs[g];                         \
        Slice s = slice_of(n, g, G);                        \
        size_t cnt = s.hi - s.lo;                           \
        CK(cudaSetDevice(d.dev));                           \
        if (arena_reset(ctx, d)) return -1;                 \
        if (cnt == 0) continue;                             \
        cudaStream_t st = d.stream;                         \
        (void)st;
#define END_FOR_EACH_DEV }

extern "C" int tcb_verify_g2_batch(tcb_ctx *ctx, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *dd, u8 *ok) {
    HOST_PROLOGUE
    FOR_EACH_DEV
        u8 *da = up(ctx, d, a + 96 * s.lo, 96 * cnt), *db = up(ctx, d, b + 192 * s.lo, 192 * cnt);
        u8 *dc = c ? up(ctx, d, c + 96 * s.lo, 96 * cnt) : nullptr, *ddv = up(ctx, d, dd + 192 * s.lo, 192 * cnt);
        u8 *dok = (u8 *)arena_alloc(ctx, d, cnt);
        if (!da || !db || (c && !dc) || !ddv || !dok) return -1;
        if (impl_verify_g2(ctx, st, cnt, da, db, dc, ddv, dok)) return -1;
        if (down(ctx, d, ok + s.lo, dok, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_hash_g2_batch(tcb_ctx *ctx, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(G);
    FOR_EACH_DEV
        u8 *dm; u64 *doff;
        if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dout = (u8 *)arena_alloc(ctx, d, 192 * cnt);
        if (!dout) return -1;
        if (impl_hash_g2(ctx, st, cnt, dm, doff, dout)) return -1;
        if (down(ctx, d, out + 192 * s.lo, dout, 192 * cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
extern "C" int tcb_verify_batch(tcb_ctx *ctx, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(G);
    FOR_EACH_DEV
        u8 *dm; u64 *doff;
        if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dpk = up(ctx, d, pk + 96 * s.lo, 96 * cnt), *dsig = up(ctx, d, sig + 192 * s.lo, 192 * cnt);
        u8 *dok = (u8 *)arena_alloc(ctx, d, cnt);
        if (!dpk || !dsig || !dok) return -1;
        if (impl_verify(ctx, st, cnt, dpk, dsig, dm, doff, dok)) return -1;
        if (down(ctx, d, ok + s.lo, dok, cnt)) return -1;
    END_FOR_EACH_DEV
    return sync_all(ctx);
}
static int sign_common(tcb_ctx *ctx, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h, u8 *out) {
    HOST_PROLOGUE
    std::vector<std::vector<u64>> keep;
    keep.reserve(G);
    FOR_EACH_DEV
        u8 *dm = nullptr, *dh = nullptr; u64 *doff = nullptr;
        if (h) { dh = up(ctx, d, h + 192 * s.lo, 192 * cnt); if (!dh) return -1; }
        else if (up_msgs(ctx, d, msgs, off, s.lo, s.hi, keep, dm, doff)) return -1;
        u8 *dsk = up(ctx, d, sk + 32 * s.lo, 32 * cnt);
        u8 *dout = (u8 *)arena_alloc(ctx, d, 192 * cnt);
        if (!dsk || !dout) return -1;
        if (impl_sign(ctx, st, cnt, dsk, dm, doff, dh, dout)) return -1;
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
    HOST_PROLOGUE
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
    HOST_PROLOGUE
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
    HOST_PROLOGUE
    size_t m = t + 1;
    std::vector<std::vector<u64>> keep;
    keep.reserve(G);
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
        if (launch(ctx, st, k_g1_mul, cnt, cnt, (const u8 *)dsk, (const u8 *)dp, dout)) return -1;
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

// ----------------------------------------------------------------------------- self-test / probes
extern "C" int tcb_selftest_fp(tcb_ctx *ctx, size_t n, uint64_t seed) {
    if (!ctx) return -2;
    DevState &d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    unsigned long long *bad = nullptr, h = 0;
    CK(cudaMalloc(&bad, 8));
    CK(cudaMemsetAsync(bad, 0, 8, d.stream));
    n = (n + 127) & ~(size_t)127;
    k_selftest_fp<<<(unsigned)(n / 128), 128, 0, d.stream>>>(n, seed, bad);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h, bad, 8, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    cudaFree(bad);
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
    if (timed(ctx, d, [&] { k_probe_imad<<<blocks, threads, 0, d.stream>>>(out, iters, 12345u); ctx->launches++; }, 3, ms)) return -1;
    cudaFree(out);
    *macs_per_sec = (double)blocks * threads * iters * 64.0 / (ms * 1e-3);
    return 0;
}
extern "C" int tcb_probe_fpmul(tcb_ctx *ctx, double *muls_per_sec) {
    if (!ctx || !muls_per_sec) return -2;
    DevState &d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    const int blocks = ctx->sm_count * 4, threads = 256, iters = 4096;
    Fp *out = nullptr;
    CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(Fp)));
    float ms = 0;
    if (timed(ctx, d, [&] { k_probe_fpmul<<<blocks, threads, 0, d.stream>>>(out, iters, 99ULL); ctx->launches++; }, 3, ms)) return -1;
    cudaFree(out);
    *muls_per_sec = (double)blocks * threads * iters * 2.0 / (ms * 1e-3);
    return 0;
}
