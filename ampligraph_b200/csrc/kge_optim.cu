// kge_optim.cu -- dense optimizer update of one embedding table (sm_100a).
//
// Replaces OptimizerWrapper.minimize -> tf.keras.optimizers.legacy.*.apply_gradients
// (optimizers.py:136-168) and the LP regulariser (regularizers.py:14-37), fused:
// one streaming pass reads {table, grad, slots}, writes {table, slots} and zeroes
// the gradient accumulator for the next step.  Semantics are the reference's
// DENSE ones: legacy Keras sums duplicate IndexedSlices rows first and Adam then
// decays m, v and moves EVERY row each step, touched or not.
// Bound: HBM, 8 fp32 streams per element for Adam (4 read + 4 written).
#include <math.h>

#include "kge_internal.h"

namespace kge {

__device__ __forceinline__ float reg_grad(float x, int p, float lam, float *pow_out)
{
    // d/dx lam*|x|^p = lam*p*|x|^(p-1)*sign(x)
    float ax = fabsf(x), sg = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
    float pm1 = (p == 2) ? ax : (p == 3) ? ax * ax : (p == 1) ? 1.f : powf(ax, (float)(p - 1));
    *pow_out = pm1 * ax;
    return lam * (float)p * pm1 * sg;
}

template <int KIND, bool REG>
__global__ void __launch_bounds__(256) kge_optim_kernel(float4 *__restrict__ var, float4 *__restrict__ grad,
                                                        float4 *__restrict__ s0, float4 *__restrict__ s1,
                                                        long long n4, OptimParams o, double *reg_loss)
{
    float racc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 x4 = var[i], g4 = grad[i];
        float x[4] = {x4.x, x4.y, x4.z, x4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
        float a[4], b[4];
        if (KIND != KGE_OPT_SGD || o.momentum != 0.f) { float4 t = s0[i]; a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
        if (KIND == KGE_OPT_ADAM) { float4 t = s1[i]; b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float gg = g[e];
            if (REG) { float pw; gg += reg_grad(x[e], o.reg_p, o.reg_lambda, &pw); racc += pw; }
            if (KIND == KGE_OPT_ADAM) {
                a[e] = fmaf(gg - a[e], 1.f - o.beta1, a[e]);       // m += (g-m)(1-b1)
                b[e] = fmaf(gg * gg - b[e], 1.f - o.beta2, b[e]);  // v += (g^2-v)(1-b2)
                x[e] -= (a[e] * o.lr_t) / (sqrtf(b[e]) + o.eps);
            } else if (KIND == KGE_OPT_ADAGRAD) {
                a[e] = fmaf(gg, gg, a[e]);
                x[e] -= o.lr * gg / (sqrtf(a[e]) + o.eps);
            } else {  // SGD (momentum optional)
                if (o.momentum != 0.f) { a[e] = o.momentum * a[e] - o.lr * gg; x[e] += a[e]; }
                else x[e] -= o.lr * gg;
            }
        }
        var[i] = make_float4(x[0], x[1], x[2], x[3]);
        if (KIND != KGE_OPT_SGD || o.momentum != 0.f) s0[i] = make_float4(a[0], a[1], a[2], a[3]);
        if (KIND == KGE_OPT_ADAM) s1[i] = make_float4(b[0], b[1], b[2], b[3]);
        grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (REG && reg_loss) {
        racc = warp_sum(racc);
        __shared__ float part[8];
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = racc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)part[w];
            if (t != 0.0) atomicAdd(reg_loss, (double)o.reg_lambda * t);
        }
    }
}

// --------------------------------------------------------------------------
// Data-parallel optimizer step FUSED with the gradient exchange over NVLink peer memory.
// Tables are replicated, every rank holds a full gradient table produced by its own batch.
// Rank r owns the row shard [begin, end): it loads that shard of EVERY rank's gradient table
// (peer loads through NVSwitch), sums them in rank order (deterministic, identical on all
// ranks), applies the optimizer to its shard -- slots exist only for the shard -- and stores
// the updated parameters into EVERY rank's table (peer stores).  One kernel = reduce-scatter
// + sharded optimizer + all-gather; it replaces all-reduce(grad) + a full-table optimizer
// pass on every replica.  Per rank: (N-1)/N * table bytes in and out over NVLink, optimizer
// HBM traffic and slot memory divided by N.  The caller brackets it with two cross-rank
// barriers (all gradients complete / all parameters delivered) and zeroes its own gradient
// table afterwards.
// --------------------------------------------------------------------------
struct PeerPtrs {
    float4 *table[KGE_MAX_PEERS];
    const float4 *grad[KGE_MAX_PEERS];
};

template <int KIND, bool REG>
__global__ void __launch_bounds__(256) kge_optim_sharded_kernel(PeerPtrs pp, int world, int rank, float4 *__restrict__ s0,
                                                                float4 *__restrict__ s1, long long off4,
                                                                long long n4, OptimParams o, double *reg_loss)
{
    float racc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const long long gi = off4 + i;
        float4 gq[KGE_MAX_PEERS];
#pragma unroll
        for (int q = 0; q < KGE_MAX_PEERS; ++q)
            if (q < world) gq[q] = __ldcg(pp.grad[q] + gi);  // peer (or local) load, L2-coherent
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < KGE_MAX_PEERS; ++q)
            if (q < world) { g[0] += gq[q].x; g[1] += gq[q].y; g[2] += gq[q].z; g[3] += gq[q].w; }
        float4 x4 = pp.table[rank][gi];  // all replicas are identical; read the local one
        float x[4] = {x4.x, x4.y, x4.z, x4.w};
        float a[4], b[4];
        if (KIND != KGE_OPT_SGD || o.momentum != 0.f) { float4 t = s0[i]; a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
        if (KIND == KGE_OPT_ADAM) { float4 t = s1[i]; b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float gg = g[e];
            if (REG) { float pw; gg += reg_grad(x[e], o.reg_p, o.reg_lambda, &pw); racc += pw; }
            if (KIND == KGE_OPT_ADAM) {
                a[e] = fmaf(gg - a[e], 1.f - o.beta1, a[e]);
                b[e] = fmaf(gg * gg - b[e], 1.f - o.beta2, b[e]);
                x[e] -= (a[e] * o.lr_t) / (sqrtf(b[e]) + o.eps);
            } else if (KIND == KGE_OPT_ADAGRAD) {
                a[e] = fmaf(gg, gg, a[e]);
                x[e] -= o.lr * gg / (sqrtf(a[e]) + o.eps);
            } else {
                if (o.momentum != 0.f) { a[e] = o.momentum * a[e] - o.lr * gg; x[e] += a[e]; }
                else x[e] -= o.lr * gg;
            }
        }
        const float4 xn = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
        for (int q = 0; q < KGE_MAX_PEERS; ++q)
            if (q < world) __stcg(pp.table[q] + gi, xn);  // deliver the updated rows to every replica
        if (KIND != KGE_OPT_SGD || o.momentum != 0.f) s0[i] = make_float4(a[0], a[1], a[2], a[3]);
        if (KIND == KGE_OPT_ADAM) s1[i] = make_float4(b[0], b[1], b[2], b[3]);
    }
    if (REG && reg_loss) {
        racc = warp_sum(racc);
        __shared__ float part[8];
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = racc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)part[w];
            if (t != 0.0) atomicAdd(reg_loss, (double)o.reg_lambda * t);
        }
    }
}

cudaError_t launch_optimizer_sharded(const OptimParams &o, int world, int rank, float *const *tables, float *const *grads,
                                     float *slot0, float *slot1, long long off_floats, long long n_floats,
                                     double *reg_loss, int sm_count, cudaStream_t st)
{
    long long n4 = n_floats / 4;
    if (n4 == 0) return cudaSuccess;
    PeerPtrs pp;
    for (int q = 0; q < KGE_MAX_PEERS; ++q) {
        pp.table[q] = q < world ? (float4 *)tables[q] : nullptr;
        pp.grad[q] = q < world ? (const float4 *)grads[q] : nullptr;
    }
    long long want = (n4 + 255) / 256;
    int grid = (int)(want < (long long)sm_count * 8 ? want : (long long)sm_count * 8);
    float4 *a = (float4 *)slot0, *b = (float4 *)slot1;
    const bool reg = o.reg_p > 0;
    const long long off4 = off_floats / 4;
#define KGE_OPTS(K)                                                                                          \
    if (reg) kge_optim_sharded_kernel<K, true><<<grid, 256, 0, st>>>(pp, world, rank, a, b, off4, n4, o, reg_loss); \
    else kge_optim_sharded_kernel<K, false><<<grid, 256, 0, st>>>(pp, world, rank, a, b, off4, n4, o, reg_loss);
    switch (o.kind) {
    case KGE_OPT_SGD: KGE_OPTS(KGE_OPT_SGD) break;
    case KGE_OPT_ADAM: KGE_OPTS(KGE_OPT_ADAM) break;
    case KGE_OPT_ADAGRAD: KGE_OPTS(KGE_OPT_ADAGRAD) break;
    default: return cudaErrorInvalidValue;
    }
#undef KGE_OPTS
    return cudaGetLastError();
}

// --------------------------------------------------------------------------
// LAZY optimizer (extension for tables whose dense update is unaffordable, e.g. cfg5: 10 M entities x
// k=1000 -> 80 GB table, 560 GB of dense-Adam traffic per step).  Only rows stamped by
// this step's training kernel are read and updated: m, v of untouched rows do NOT decay and the rows do
// not move (the semantics of TensorFlow-Addons' LazyAdam), bias correction uses the global step.  This
// is NOT the reference's dense rule; it is opt-in ('lazy_adam' etc.).  A warp scans 32 row stamps with
// one coalesced read, ballots the touched ones and updates each touched row cooperatively.
// --------------------------------------------------------------------------
template <int KIND, bool REG>
__global__ void __launch_bounds__(256) kge_optim_lazy_kernel(float *__restrict__ var, float *__restrict__ grad,
                                                             float *__restrict__ s0, float *__restrict__ s1,
                                                             long long rows, int ld, const int *__restrict__ row_stamp,
                                                             int stamp, OptimParams o, double *reg_loss)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    float racc = 0.f;
    for (long long base = warp0 * 32; base < rows; base += n_warps * 32) {
        const long long r = base + lane;
        unsigned m = __ballot_sync(0xffffffffu, r < rows && row_stamp[r] == stamp);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const size_t off = (size_t)(base + b) * ld;
            for (int c = lane * 4; c < ld; c += 128) {
                float4 x4 = *reinterpret_cast<float4 *>(var + off + c), g4 = *reinterpret_cast<float4 *>(grad + off + c);
                float x[4] = {x4.x, x4.y, x4.z, x4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, a[4], bb[4];
                if (KIND != KGE_OPT_SGD || o.momentum != 0.f) { float4 t = *reinterpret_cast<float4 *>(s0 + off + c); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
                if (KIND == KGE_OPT_ADAM) { float4 t = *reinterpret_cast<float4 *>(s1 + off + c); bb[0] = t.x; bb[1] = t.y; bb[2] = t.z; bb[3] = t.w; }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float gg = g[e];
                    if (REG) { float pw; gg += reg_grad(x[e], o.reg_p, o.reg_lambda, &pw); racc += pw; }
                    if (KIND == KGE_OPT_ADAM) {
                        a[e] = fmaf(gg - a[e], 1.f - o.beta1, a[e]);
                        bb[e] = fmaf(gg * gg - bb[e], 1.f - o.beta2, bb[e]);
                        x[e] -= (a[e] * o.lr_t) / (sqrtf(bb[e]) + o.eps);
                    } else if (KIND == KGE_OPT_ADAGRAD) {
                        a[e] = fmaf(gg, gg, a[e]);
                        x[e] -= o.lr * gg / (sqrtf(a[e]) + o.eps);
                    } else {
                        if (o.momentum != 0.f) { a[e] = o.momentum * a[e] - o.lr * gg; x[e] += a[e]; }
                        else x[e] -= o.lr * gg;
                    }
                }
                *reinterpret_cast<float4 *>(var + off + c) = make_float4(x[0], x[1], x[2], x[3]);
                if (KIND != KGE_OPT_SGD || o.momentum != 0.f) *reinterpret_cast<float4 *>(s0 + off + c) = make_float4(a[0], a[1], a[2], a[3]);
                if (KIND == KGE_OPT_ADAM) *reinterpret_cast<float4 *>(s1 + off + c) = make_float4(bb[0], bb[1], bb[2], bb[3]);
                *reinterpret_cast<float4 *>(grad + off + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    if (REG && reg_loss) {
        racc = warp_sum(racc);
        if (lane == 0 && racc != 0.f) atomicAdd(reg_loss, (double)o.reg_lambda * (double)racc);
    }
}

cudaError_t launch_optimizer_lazy(const OptimParams &o, float *table, float *grad, float *slot0, float *slot1,
                                  long long rows, int ld, const int *row_stamp, int stamp, double *reg_loss,
                                  int sm_count, cudaStream_t st)
{
    if (rows == 0) return cudaSuccess;
    long long want = (rows + 255) / 256;  // 8 warps x 32 rows per block
    int grid = (int)(want < (long long)sm_count * 8 ? want : (long long)sm_count * 8);
    const bool reg = o.reg_p > 0;
#define KGE_OPTL(K)                                                                                                      \
    if (reg) kge_optim_lazy_kernel<K, true><<<grid, 256, 0, st>>>(table, grad, slot0, slot1, rows, ld, row_stamp, stamp, o, reg_loss); \
    else kge_optim_lazy_kernel<K, false><<<grid, 256, 0, st>>>(table, grad, slot0, slot1, rows, ld, row_stamp, stamp, o, reg_loss);
    switch (o.kind) {
    case KGE_OPT_SGD: KGE_OPTL(KGE_OPT_SGD) break;
    case KGE_OPT_ADAM: KGE_OPTL(KGE_OPT_ADAM) break;
    case KGE_OPT_ADAGRAD: KGE_OPTL(KGE_OPT_ADAGRAD) break;
    default: return cudaErrorInvalidValue;
    }
#undef KGE_OPTL
    return cudaGetLastError();
}

cudaError_t launch_optimizer(const OptimParams &o, float *table, float *grad, float *slot0, float *slot1,
                             long long n_floats, double *reg_loss, int sm_count, cudaStream_t st)
{
    long long n4 = n_floats / 4;  // ld is a multiple of 4
    if (n4 == 0) return cudaSuccess;
    long long want = (n4 + 255) / 256;
    int grid = (int)(want < (long long)sm_count * 8 ? want : (long long)sm_count * 8);
    float4 *v = (float4 *)table, *g = (float4 *)grad, *a = (float4 *)slot0, *b = (float4 *)slot1;
    const bool reg = o.reg_p > 0;
#define KGE_OPT(K)                                                                            \
    if (reg) kge_optim_kernel<K, true><<<grid, 256, 0, st>>>(v, g, a, b, n4, o, reg_loss);    \
    else kge_optim_kernel<K, false><<<grid, 256, 0, st>>>(v, g, a, b, n4, o, reg_loss);
    switch (o.kind) {
    case KGE_OPT_SGD: KGE_OPT(KGE_OPT_SGD) break;
    case KGE_OPT_ADAM: KGE_OPT(KGE_OPT_ADAM) break;
    case KGE_OPT_ADAGRAD: KGE_OPT(KGE_OPT_ADAGRAD) break;
    default: return cudaErrorInvalidValue;
    }
#undef KGE_OPT
    return cudaGetLastError();
}

}  // namespace kge
