// kge_train_common.cuh -- device helpers shared by the training kernels (kge_train.cu: the general kernel;
// kge_train_res.cu: the resident trilinear fast path): packed-fp32 float4 arithmetic, the gradient sink, warp
// reductions, and the per-positive loss + dL/dscore of the five reference losses.
#pragma once
#include <math.h>

#include "kge_internal.h"

namespace kge {

// --------------------------------------------------------------------------
// float4 helpers
// --------------------------------------------------------------------------
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4ld(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void f4st(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// Blackwell packed fp32: fma/mul/add.rn.f32x2 (SASS FFMA2/FMUL2/FADD2) do two IEEE fp32 operations per
// issue slot on a 64-bit register pair.  The kernel is issue-bound, so every float4 op is two f32x2 ops.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b)
{
    f32x2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b)
{
    f32x2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float4 f4from(f32x2_t lo, f32x2_t hi)
{
    float4 r;
    upk2(lo, r.x, r.y);
    upk2(hi, r.z, r.w);
    return r;
}
#define KGE_LO(v) pk2((v).x, (v).y)
#define KGE_HI(v) pk2((v).z, (v).w)
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return f4from(add2(KGE_LO(a), KGE_LO(b)), add2(KGE_HI(a), KGE_HI(b))); }
__device__ __forceinline__ float4 f4neg(float4 a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return a + f4neg(b); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return f4from(mul2(KGE_LO(a), KGE_LO(b)), mul2(KGE_HI(a), KGE_HI(b))); }
__device__ __forceinline__ float4 operator*(float a, float4 b)
{
    const f32x2_t aa = pk2(a, a);
    return f4from(mul2(aa, KGE_LO(b)), mul2(aa, KGE_HI(b)));
}
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c)
{
    return f4from(fma2(KGE_LO(a), KGE_LO(b), KGE_LO(c)), fma2(KGE_HI(a), KGE_HI(b), KGE_HI(c)));
}
__device__ __forceinline__ float4 f4fma(float a, float4 b, float4 c)
{
    const f32x2_t aa = pk2(a, a);
    return f4from(fma2(aa, KGE_LO(b), KGE_LO(c)), fma2(aa, KGE_HI(b), KGE_HI(c)));
}
__device__ __forceinline__ float f4hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }  // TF abs grad
__device__ __forceinline__ float4 f4sgn(float4 a) { return make_float4(sgnf(a.x), sgnf(a.y), sgnf(a.z), sgnf(a.w)); }
__device__ __forceinline__ float f4abssum(float4 a) { return (fabsf(a.x) + fabsf(a.y)) + (fabsf(a.z) + fabsf(a.w)); }

// --------------------------------------------------------------------------
// gradient sink: red.global.add.v4.f32 straight from registers into the gradient table
// (the scorers keep a Sink template parameter so that a profiling build can swap it).
// --------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float *g, float4 v)
{
#ifdef KGE_PROFILE_NOSCATTER
    if (v.x != 1.2345e38f) return;  // profiling build: the scatter is dropped, v stays live
#endif
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// -DKGE_PROFILE_NOSCATTER builds a profiling variant that drops the gradient scatter (how much of the kernel the atomics
// cost).  The product never tests a run-time flag here: a branch around every RED cost 3.5 % (cfg2) to 11 % (cfg3).
struct SinkRed {
    static __device__ __forceinline__ void put(float *, float *grow, int, int goff, float4 v) { red_add_v4(grow + goff, v); }
};

// RotatE's modulus and unit vector use the MUFU approximations (sqrt.approx: 2^-23 relative, rsqrt.approx: 2 ulp), far
// inside the 1e-4 training tolerance; the RANKING kernels keep the correctly rounded sqrt because their scores must be
// bit-identical to the oracle.
// single-instruction MUFU forms (flush-to-zero: a denormal x = re^2 + im^2 means a residual below 1e-19)
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}



__device__ __forceinline__ void warp_sum2(float &a, float &b)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
}

// Two warp sums for the price of five shuffles (instead of ten): after the xor-16 step the lower half-warp carries a and
// the upper one b; four more steps finish each sum inside its half.  On return lanes [0,16) hold sum(a), lanes [16,32) sum(b).
__device__ __forceinline__ float warp_sum2t(float a, float b, int lane)
{
    const bool hi16 = (lane & 16) != 0;
    float k = hi16 ? b : a;
    k += __shfl_xor_sync(0xffffffffu, hi16 ? a : b, 16);
    k += __shfl_xor_sync(0xffffffffu, k, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// Four warp sums for the price of six shuffles: a transposed butterfly.  After the xor-16 step the lower half-warp
// carries v0, v1 and the upper one v2, v3; after the xor-8 step each quarter-warp carries ONE value; three more
// steps finish the sum inside the quarter.  On return lanes [8q, 8q+8) hold the full sum of value q.
__device__ __forceinline__ float warp_sum4t(float v0, float v1, float v2, float v3, int lane)
{
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    float k0 = hi16 ? v2 : v0, k1 = hi16 ? v3 : v1;
    k0 += __shfl_xor_sync(0xffffffffu, hi16 ? v0 : v2, 16);
    k1 += __shfl_xor_sync(0xffffffffu, hi16 ? v1 : v3, 16);
    float k = hi8 ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// --------------------------------------------------------------------------
// explicit shared-memory accessors (32-bit shared addresses) and register-pinning helpers of the fast paths
// (kge_train_res.cu, kge_train_rot.cu)
// --------------------------------------------------------------------------
__device__ __forceinline__ float4 lds4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds_i(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_i(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_init_s(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RES_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RES_DONE;\n"
        "bra RES_WAIT_LOOP;\n"
        "RES_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load_s(uint32_t smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}
// The front end rematerialises cheap values (kernel parameters, %tid-derived addresses) inside the long per-positive
// loop instead of keeping them in registers -- a dozen ld.param / shift / mad per trip of the inner loops.  Passing a
// value through an empty asm makes it opaque: it is computed once and stays in its register.
#define KGE_KEEP32(x) asm volatile("" : "+r"(x))
#define KGE_KEEP64(x) asm volatile("" : "+l"(x))
#define KGE_KEEPF(x) asm volatile("" : "+f"(x))
__device__ __forceinline__ void red4(char *row, uint32_t byte_off, float4 v) { red_add_v4(reinterpret_cast<float *>(row + byte_off), v); }


// --------------------------------------------------------------------------
// Dynamic assignment of positives to warps.  A static stride (warp w takes w, w + n_warps, ...) quantises the launch to
// ceil(B / n_warps) rounds -- cfg2: 27,212 positives over 1,776 warps = 15.3, so a third of the warps run a 16th round
// while the rest idle (4 %); cfg4: 9.1 -> 10 rounds (9 %).  Instead every warp starts on its own index and draws the
// following ones from a counter: the draw for the NEXT positive is issued before the current one is processed, so its
// latency is never seen.  The counters reset themselves: the last warp to retire (all draws are complete by then)
// zeroes both words, so consecutive launches on one stream need no memset and the scheme survives graph replay.
// --------------------------------------------------------------------------
// (a FLOAT counter on purpose: the compiler -- nvcc for atomicAdd(), ptxas for atom.add / atom.inc in inline PTX, u32 and u64,
// even with a laundered addend -- aggregates a warp's integer atomic adds (VOTE + POPC, one ATOMG) and reads the result back
// with a shuffle right after it, which puts the whole L2 round trip, microseconds when HBM is saturated, at the top of every
// positive: measured +35 % on the HBM-resident `big` shape, profiles/r2n_kbench_dynamic.log.  atom.add.f32 is left alone;
// counts are exact below 2^24, the host falls back to the static stride for larger batches.)
#define KGE_SCHED_MAX_B (1ll << 24)
__device__ __forceinline__ float sched_draw(unsigned *sched)  // the caller converts when it consumes the value, not here
{
    float v;
    asm volatile("atom.relaxed.gpu.global.add.f32 %0, [%1], 0f3F800000;" : "=f"(v) : "l"(sched) : "memory");
    return v;
}
__device__ __forceinline__ void sched_retire(unsigned *sched, long long n_warps, int lane)
{
    if (sched && lane == 0) {
        const unsigned d = atomicAdd(sched + 1, 1u);
        if ((long long)d == n_warps - 1) { sched[0] = 0u; sched[1] = 0u; }
    }
}

// --------------------------------------------------------------------------
// per-positive loss and dL/dscore (warp-cooperative; lanes stride over j).
// in: P, sc[j] = N_j.  out: sc[j] = dL/dN_j, returns loss_i, *dP.
// --------------------------------------------------------------------------
// x / y for y in a range where the IEEE slow path (denormal operands) cannot matter: reciprocal + multiply, no FCHK branch
#ifdef KGE_FAST_LOSS_MATH
__device__ __forceinline__ float fdiv(float x, float y) { return __fdividef(x, y); }
#else
__device__ __forceinline__ float fdiv(float x, float y) { return x / y; }
#endif
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return fdiv(1.f, 1.f + expf(-x)); }
#define KGE_CLIP_LO (-75.0f)  // loss_functions.py:32
#define KGE_CLIP_HI (75.0f)   // loss_functions.py:35

static __device__ __forceinline__ float loss_and_dscores(const TrainParams &p, float P, float *sc, int lane, float *dP_out)
{
    const int eta = p.eta;
    const float w = (p.reduction == KGE_REDUCE_MEAN) ? 1.f / (float)eta : 1.f;
    float loss = 0.f, dP = 0.f;
    switch (p.loss) {
    case KGE_LOSS_PAIRWISE: {  // loss_functions.py:305-307
        float acc = 0.f, ds = 0.f;
        for (int j = lane; j < eta; j += 32) {
            float x = p.margin - P + sc[j];
            float d = (x >= 0.f) ? w : 0.f;  // TF maximum(): tie goes to the first argument
            acc += fmaxf(x, 0.f);
            ds += d;
            sc[j] = d;
        }
        loss = w * warp_sum(acc);
        dP = -warp_sum(ds);
        break;
    }
    case KGE_LOSS_NLL: {  // loss_functions.py:376-382 (positive term counted eta times)
        const float w2 = (p.reduction == KGE_REDUCE_MEAN) ? 1.f / (2.f * (float)eta) : 1.f;
        float Pc = fminf(fmaxf(P, KGE_CLIP_LO), KGE_CLIP_HI);
        float inP = (P >= KGE_CLIP_LO && P <= KGE_CLIP_HI) ? 1.f : 0.f;
        float acc = 0.f;
        for (int j = lane; j < eta; j += 32) {
            float N = sc[j];
            float Nc = fminf(fmaxf(N, KGE_CLIP_LO), KGE_CLIP_HI);
            float inN = (N >= KGE_CLIP_LO && N <= KGE_CLIP_HI) ? 1.f : 0.f;
            float e = expf(Nc);
            acc += logf(1.f + e);
            sc[j] = w2 * inN * (e / (1.f + e));
        }
        float ep = expf(-Pc);
        loss = w2 * ((float)eta * logf(1.f + ep) + warp_sum(acc));
        dP = -w2 * (float)eta * inP * (ep / (1.f + ep));
        break;
    }
    case KGE_LOSS_ABSOLUTE_MARGIN: {  // loss_functions.py:461-463
        float acc = 0.f;
        for (int j = lane; j < eta; j += 32) {
            float x = p.margin + sc[j];
            acc += fmaxf(x, 0.f);
            sc[j] = (x >= 0.f) ? w : 0.f;
        }
        loss = w * warp_sum(acc) - w * (float)eta * P;
        dP = -w * (float)eta;
        break;
    }
    case KGE_LOSS_SELF_ADVERSARIAL: {  // loss_functions.py:563-572, softmax NOT detached
        // with x = -N - margin, t = exp(-|x|):  log_sigmoid(x) = min(x,0) - log1p(t) and
        // sigmoid(N + margin) = sigmoid(-x) = (x < 0 ? 1 : t) / (1 + t): one exp + one log1p per corruption
        if (eta <= 32) {  // one corruption per lane: everything stays in registers
            const bool on = lane < eta;
            const float N = on ? sc[lane] : 0.f;
            const float mx = warp_max(on ? p.alpha * N : -INFINITY);
            const float e = on ? expf(p.alpha * N - mx) : 0.f;
            const float x = -N - p.margin, t = expf(-fabsf(x));
            const float lj = fminf(x, 0.f) - log1pf(t);
            const float sg = fdiv((x < 0.f) ? 1.f : t, 1.f + t);
            float z = e, sl = e * lj;
            warp_sum2(z, sl);
            const float S = fdiv(sl, z), pj = fdiv(e, z);
            if (on) sc[lane] = w * (pj * sg - p.alpha * pj * (lj - S));
            loss = -log_sigmoid(p.margin + P) - w * S;
            dP = -sigmoidf(-(p.margin + P));
            break;
        }
        float mx = -INFINITY;
        for (int j = lane; j < eta; j += 32) mx = fmaxf(mx, p.alpha * sc[j]);
        mx = warp_max(mx);
        float z = 0.f, sl = 0.f;
        for (int j = lane; j < eta; j += 32) {
            float e = expf(p.alpha * sc[j] - mx);
            z += e;
            sl += e * log_sigmoid(-sc[j] - p.margin);
        }
        z = warp_sum(z);
        float S = warp_sum(sl) / z;
        for (int j = lane; j < eta; j += 32) {
            float N = sc[j];
            float pj = expf(p.alpha * N - mx) / z;
            float lj = log_sigmoid(-N - p.margin);
            sc[j] = w * (pj * sigmoidf(N + p.margin) - p.alpha * pj * (lj - S));
        }
        loss = -log_sigmoid(p.margin + P) - w * S;
        dP = -sigmoidf(-(p.margin + P));
        break;
    }
    case KGE_LOSS_MULTICLASS_NLL: {  // loss_functions.py:647-653
        float Pc = fminf(fmaxf(P, KGE_CLIP_LO), KGE_CLIP_HI);
        float inP = (P >= KGE_CLIP_LO && P <= KGE_CLIP_HI) ? 1.f : 0.f;
        float acc = 0.f;
        for (int j = lane; j < eta; j += 32) acc += expf(fminf(fmaxf(sc[j], KGE_CLIP_LO), KGE_CLIP_HI));
        float pe = expf(Pc);
        float D = w * warp_sum(acc) + pe;
        for (int j = lane; j < eta; j += 32) {
            float N = sc[j];
            float inN = (N >= KGE_CLIP_LO && N <= KGE_CLIP_HI) ? 1.f : 0.f;
            sc[j] = w * inN * expf(fminf(fmaxf(N, KGE_CLIP_LO), KGE_CLIP_HI)) / D;
        }
        loss = -logf(pe / D);
        dP = inP * (pe / D - 1.f);
        break;
    }
    }
    *dP_out = dP;
    return loss;
}


}  // namespace kge
