// k_miller.cu — the Miller loop of the pairing-equality check on the shared-memory engine (quadsm.cuh): operands staged in
// shared memory, products as dot products with one reduction, multiplier core = one small loop.  The accumulator f goes through
// global memory to k_final_exp_quad (k_pairing.cu).  The single Fp multiplies of this unit are calls (by-value ABI).
#define TCB_FP_NOINLINE 1
#include "kern.h"
#include "quadsm.cuh"
using namespace tcb;

#ifndef TCB_MILLER_MINB
#define TCB_MILLER_MINB 2
#endif
__global__ void __launch_bounds__(QNT, TCB_MILLER_MINB) k_miller_quad(size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, Fp *fout, u8 *enc_ok, int gen_scaled) {
    q_miller_block(n, a, b, c, d, fout, enc_ok, gen_scaled != 0);
}

namespace tcbk {
cudaError_t upload_consts_miller(const Consts &c) {
    cudaError_t e = cudaFuncSetAttribute(k_miller_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(d_consts, &c, sizeof c);
}
size_t miller_f_bytes() { return 12 * sizeof(Fp); }
void run_miller_quad(cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, void *fbuf, u8 *enc_ok, bool gen_scaled) {
    if (!n) return;
    k_miller_quad<<<(unsigned)((n + QNT / 4 - 1) / (QNT / 4)), QNT, Q_SMEM_BYTES, st>>>(n, a, b, c, d, (Fp *)fbuf, enc_ok, gen_scaled ? 1 : 0);
}
}  // namespace tcbk
