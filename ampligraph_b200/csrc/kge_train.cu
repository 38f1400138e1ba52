// kge_train.cu -- fused forward+backward of one training batch (sm_100a).
//
// Replaces, in ONE kernel: EmbeddingLookupLayer.call (layers/encoding/
// EmbeddingLookupLayer.py:307-342) for positives and corruptions,
// CorruptionGenerationLayerTrain.call (layers/corruption_generation/
// CorruptionGenerationLayerTrain.py:35-94), the five _compute_scores
// (layers/scoring/{TransE.py:37,DistMult.py:34,ComplEx.py:39,HolE.py:31,RotatE.py:62}),
// the five losses (loss_functions.py:286-308,360-382,442-464,540-574,630-654 via
// Loss.__call__ :185-225) and tape.gradient (optimizers.py:166).
//
// Work decomposition: one warp owns one positive and its eta corruptions at a
// time.  The warp's (3+G) embedding rows are gathered from HBM/L2 into its
// shared-memory slot by the copy engine (cp.async.bulk, one 1-D bulk copy per
// row, completion on a per-warp mbarrier); scores, loss and dL/dscore are
// computed from shared memory; gradient rows are written IN PLACE over the
// gathered rows and pushed to the gradient tables by the copy engine again
// (cp.reduce.async.bulk ... .add.f32 -- an element-wise fp32 atomic add of a
// whole row).  The reference re-gathers s,p,o for every corruption
// (3(1+eta) rows per positive); here each needed row moves once: (3+eta) rows
// in, (3+eta) gradient rows out = 2(3+eta)*ld*4 bytes per positive.
#include <math.h>

#include "kge_internal.h"

namespace kge {

// --------------------------------------------------------------------------
// float4 helpers
// --------------------------------------------------------------------------
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4ld(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void f4st(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c)
{
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4fma(float a, float4 b, float4 c)
{
    return make_float4(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y), fmaf(a, b.z, c.z), fmaf(a, b.w, c.w));
}
__device__ __forceinline__ float f4hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float f4dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }  // TF abs grad
__device__ __forceinline__ float4 f4sgn(float4 a) { return make_float4(sgnf(a.x), sgnf(a.y), sgnf(a.z), sgnf(a.w)); }
__device__ __forceinline__ float f4abssum(float4 a) { return (fabsf(a.x) + fabsf(a.y)) + (fabsf(a.z) + fabsf(a.w)); }

// --------------------------------------------------------------------------
// Per-model scorers.  Each lane owns float4 chunks c = lane + 32*it (it < NIT)
// of every (padded) half-row; pad columns are zero in HBM and produce zero
// scores / gradients.  Methods (all warp-synchronous, no cross-lane traffic):
//   prep(s,p,o)          load per-positive state; returns this lane's partial of f(s,p,o)
//   neg_partial(r,side)  partial of f for a corruption whose replaced row is r
//                        side 0 = subject replaced (keep_subj = 0), 1 = object replaced
//   neg_grad(r,side,g)   overwrite r with g * df/dr and accumulate what the kept
//                        rows will need
//   finish(s,p,o,gP)     overwrite s,p,o with their total gradient rows
// Unscaled f; HolE's 2/k factor is folded into g by the caller.
// --------------------------------------------------------------------------
template <int MODEL, int NIT>
struct Scorer;

// ---- DistMult: f = sum s*p*o (DistMult.py:48) -----------------------------
template <int NIT>
struct Scorer<KGE_DISTMULT, NIT> {
    float4 A[NIT], C[NIT], Ws[NIT], Wo[NIT];
    int lane, nch;
    __device__ __forceinline__ float prep(const float *s, const float *p, const float *o, const TrainParams &, int ln)
    {
        lane = ln;
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            float4 vs = f4zero(), vp = f4zero(), vo = f4zero();
            if (c < nch) { vs = f4ld(s + 4 * c); vp = f4ld(p + 4 * c); vo = f4ld(o + 4 * c); }
            A[it] = vp * vo;
            C[it] = vs * vp;
            Ws[it] = f4zero();
            Wo[it] = f4zero();
            acc += f4dot(vs, A[it]);
        }
        return acc;
    }
    __device__ __forceinline__ float neg_partial(const float *r, int side) const
    {
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) acc += f4dot(f4ld(r + 4 * c), side ? C[it] : A[it]);
        }
        return acc;
    }
    __device__ __forceinline__ void neg_grad(float *r, int side, float g)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 v = f4ld(r + 4 * c);
                if (side) { Wo[it] = f4fma(g, v, Wo[it]); f4st(r + 4 * c, g * C[it]); }
                else { Ws[it] = f4fma(g, v, Ws[it]); f4st(r + 4 * c, g * A[it]); }
            }
        }
    }
    __device__ __forceinline__ void finish(float *s, float *p, float *o, float gP)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 vs = f4ld(s + 4 * c), vp = f4ld(p + 4 * c), vo = f4ld(o + 4 * c);
                float4 U = f4fma(gP, vs, Ws[it]);  // everything that sat in the subject slot
                float4 X = f4fma(gP, vo, Wo[it]);  // everything that sat in the object slot
                f4st(s + 4 * c, vp * X);
                f4st(o + 4 * c, U * vp);
                f4st(p + 4 * c, f4fma(U, vo, vs * Wo[it]));
            }
        }
    }
};

// ---- ComplEx / HolE (ComplEx.py:52-62; HolE.py:45 scales by 2/k) ----------
// f = sum s_re (p_re o_re + p_im o_im) + s_im (p_re o_im - p_im o_re); trilinear, so the
// kept-row gradients come from the weighted sums W of the replaced rows.
template <int NIT>
struct ComplexScorer {
    float4 A[NIT], Bv[NIT], C[NIT], D[NIT];          // subject-side / object-side query vectors
    float4 Wsr[NIT], Wsi[NIT], Wor[NIT], Woi[NIT];   // sum_j g_j r_j per side (re, im)
    int lane, nch, kp;
    __device__ __forceinline__ float prep(const float *s, const float *p, const float *o, const TrainParams &, int ln)
    {
        lane = ln;
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            float4 sr = f4zero(), si = f4zero(), pr = f4zero(), pi = f4zero(), orr = f4zero(), oi = f4zero();
            if (c < nch) {
                sr = f4ld(s + 4 * c); si = f4ld(s + kp + 4 * c);
                pr = f4ld(p + 4 * c); pi = f4ld(p + kp + 4 * c);
                orr = f4ld(o + 4 * c); oi = f4ld(o + kp + 4 * c);
            }
            A[it] = f4fma(pi, oi, pr * orr);
            Bv[it] = pr * oi - pi * orr;
            C[it] = sr * pr - si * pi;
            D[it] = f4fma(sr, pi, si * pr);
            Wsr[it] = Wsi[it] = Wor[it] = Woi[it] = f4zero();
            acc += f4dot(sr, A[it]) + f4dot(si, Bv[it]);
        }
        return acc;
    }
    __device__ __forceinline__ float neg_partial(const float *r, int side) const
    {
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 rr = f4ld(r + 4 * c), ri = f4ld(r + kp + 4 * c);
                acc += side ? (f4dot(rr, C[it]) + f4dot(ri, D[it])) : (f4dot(rr, A[it]) + f4dot(ri, Bv[it]));
            }
        }
        return acc;
    }
    __device__ __forceinline__ void neg_grad(float *r, int side, float g)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 rr = f4ld(r + 4 * c), ri = f4ld(r + kp + 4 * c);
                if (side) {
                    Wor[it] = f4fma(g, rr, Wor[it]); Woi[it] = f4fma(g, ri, Woi[it]);
                    f4st(r + 4 * c, g * C[it]); f4st(r + kp + 4 * c, g * D[it]);
                } else {
                    Wsr[it] = f4fma(g, rr, Wsr[it]); Wsi[it] = f4fma(g, ri, Wsi[it]);
                    f4st(r + 4 * c, g * A[it]); f4st(r + kp + 4 * c, g * Bv[it]);
                }
            }
        }
    }
    __device__ __forceinline__ void finish(float *s, float *p, float *o, float gP)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 sr = f4ld(s + 4 * c), si = f4ld(s + kp + 4 * c);
                float4 pr = f4ld(p + 4 * c), pi = f4ld(p + kp + 4 * c);
                float4 orr = f4ld(o + 4 * c), oi = f4ld(o + kp + 4 * c);
                float4 Ur = f4fma(gP, sr, Wsr[it]), Ui = f4fma(gP, si, Wsi[it]);    // subject-slot mass
                float4 Xr = f4fma(gP, orr, Wor[it]), Xi = f4fma(gP, oi, Woi[it]);   // object-slot mass
                // d/ds f(s,p,X)
                f4st(s + 4 * c, f4fma(pi, Xi, pr * Xr));
                f4st(s + kp + 4 * c, pr * Xi - pi * Xr);
                // d/do f(U,p,o)
                f4st(o + 4 * c, Ur * pr - Ui * pi);
                f4st(o + kp + 4 * c, f4fma(Ur, pi, Ui * pr));
                // d/dp [f(U,p,o) + f(s,p,Wo)]
                f4st(p + 4 * c, f4fma(Ur, orr, Ui * oi) + f4fma(sr, Wor[it], si * Woi[it]));
                f4st(p + kp + 4 * c, (Ur * oi - Ui * orr) + (sr * Woi[it] - si * Wor[it]));
            }
        }
    }
};
template <int NIT> struct Scorer<KGE_COMPLEX, NIT> : ComplexScorer<NIT> {};
template <int NIT> struct Scorer<KGE_HOLE, NIT> : ComplexScorer<NIT> {};

// ---- TransE: f = -sum |s+p-o| (TransE.py:51-53) ---------------------------
template <int NIT>
struct Scorer<KGE_TRANSE, NIT> {
    float4 Qs[NIT], Qo[NIT], Vs[NIT], Vo[NIT];  // p-o, s+p, sum g*sign per side
    int lane, nch;
    __device__ __forceinline__ float prep(const float *s, const float *p, const float *o, const TrainParams &, int ln)
    {
        lane = ln;
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            float4 vs = f4zero(), vp = f4zero(), vo = f4zero();
            if (c < nch) { vs = f4ld(s + 4 * c); vp = f4ld(p + 4 * c); vo = f4ld(o + 4 * c); }
            Qs[it] = vp - vo;
            Qo[it] = vs + vp;
            Vs[it] = Vo[it] = f4zero();
            acc -= f4abssum(Qo[it] - vo);
        }
        return acc;
    }
    __device__ __forceinline__ float neg_partial(const float *r, int side) const
    {
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 v = f4ld(r + 4 * c);
                acc -= f4abssum(side ? (Qo[it] - v) : (v + Qs[it]));
            }
        }
        return acc;
    }
    __device__ __forceinline__ void neg_grad(float *r, int side, float g)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 v = f4ld(r + 4 * c);
                if (side) {  // t = (s+p) - r ; f = -|t| ; df/dr = +sign(t)
                    float4 gs = g * f4sgn(Qo[it] - v);
                    Vo[it] = Vo[it] + gs;
                    f4st(r + 4 * c, gs);
                } else {  // t = r + (p-o) ; df/dr = -sign(t)
                    float4 gs = g * f4sgn(v + Qs[it]);
                    Vs[it] = Vs[it] + gs;
                    f4st(r + 4 * c, -1.f * gs);
                }
            }
        }
    }
    __device__ __forceinline__ void finish(float *s, float *p, float *o, float gP)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 vo = f4ld(o + 4 * c);
                float4 Vp = gP * f4sgn(Qo[it] - vo);
                f4st(s + 4 * c, -1.f * (Vp + Vo[it]));
                f4st(p + 4 * c, -1.f * (Vp + Vs[it] + Vo[it]));
                f4st(o + 4 * c, Vp + Vs[it]);
            }
        }
    }
};

// ---- RotatE (RotatE.py:76-104) --------------------------------------------
// The relation row handed to this kernel is the per-step rotation table row
// [cos(phi) | sin(phi)], phi = theta/div (kge_rotation_table_kernel).
// f = -sum_d |R(phi) s - o|.  With (a,b) = residual/|residual|:
//   df/do = (a,b); df/ds = -R(-phi)(a,b); df/dphi = a*y_im - b*y_re, y = R(phi) s.
// The reference's gradient is NaN at an exactly-zero residual; here it is 0.
template <int NIT>
struct Scorer<KGE_ROTATE, NIT> {
    float4 Cs[NIT], Sn[NIT], Yr[NIT], Yi[NIT], Or_[NIT], Oi[NIT];
    float4 Zor[NIT], Zoi[NIT], Zsr[NIT], Zsi[NIT], Aphi[NIT];
    int lane, nch, kp;
    static __device__ __forceinline__ void unit(float4 re, float4 im, float4 &a, float4 &b, float &msum)
    {
        float m0 = sqrtf(fmaf(im.x, im.x, re.x * re.x)), m1 = sqrtf(fmaf(im.y, im.y, re.y * re.y));
        float m2 = sqrtf(fmaf(im.z, im.z, re.z * re.z)), m3 = sqrtf(fmaf(im.w, im.w, re.w * re.w));
        msum = (m0 + m1) + (m2 + m3);
        float i0 = m0 > 0.f ? 1.f / m0 : 0.f, i1 = m1 > 0.f ? 1.f / m1 : 0.f;
        float i2 = m2 > 0.f ? 1.f / m2 : 0.f, i3 = m3 > 0.f ? 1.f / m3 : 0.f;
        a = make_float4(re.x * i0, re.y * i1, re.z * i2, re.w * i3);
        b = make_float4(im.x * i0, im.y * i1, im.z * i2, im.w * i3);
    }
    __device__ __forceinline__ float prep(const float *s, const float *p, const float *o, const TrainParams &, int ln)
    {
        lane = ln;
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            float4 sr = f4zero(), si = f4zero();
            Cs[it] = Sn[it] = Or_[it] = Oi[it] = f4zero();
            if (c < nch) {
                sr = f4ld(s + 4 * c); si = f4ld(s + kp + 4 * c);
                Cs[it] = f4ld(p + 4 * c); Sn[it] = f4ld(p + kp + 4 * c);
                Or_[it] = f4ld(o + 4 * c); Oi[it] = f4ld(o + kp + 4 * c);
            }
            Yr[it] = sr * Cs[it] - si * Sn[it];
            Yi[it] = f4fma(sr, Sn[it], si * Cs[it]);
            Zor[it] = Zoi[it] = Zsr[it] = Zsi[it] = Aphi[it] = f4zero();
            float4 re = Yr[it] - Or_[it], im = Yi[it] - Oi[it];
            acc -= (sqrtf(fmaf(im.x, im.x, re.x * re.x)) + sqrtf(fmaf(im.y, im.y, re.y * re.y))) +
                   (sqrtf(fmaf(im.z, im.z, re.z * re.z)) + sqrtf(fmaf(im.w, im.w, re.w * re.w)));
        }
        return acc;
    }
    __device__ __forceinline__ float neg_partial(const float *r, int side) const
    {
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 rr = f4ld(r + 4 * c), ri = f4ld(r + kp + 4 * c), re, im;
                if (side) { re = Yr[it] - rr; im = Yi[it] - ri; }
                else {
                    re = (rr * Cs[it] - ri * Sn[it]) - Or_[it];
                    im = f4fma(rr, Sn[it], ri * Cs[it]) - Oi[it];
                }
                acc -= (sqrtf(fmaf(im.x, im.x, re.x * re.x)) + sqrtf(fmaf(im.y, im.y, re.y * re.y))) +
                       (sqrtf(fmaf(im.z, im.z, re.z * re.z)) + sqrtf(fmaf(im.w, im.w, re.w * re.w)));
            }
        }
        return acc;
    }
    __device__ __forceinline__ void neg_grad(float *r, int side, float g)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 rr = f4ld(r + 4 * c), ri = f4ld(r + kp + 4 * c), a, b;
                float ms;
                if (side) {  // residual = y(s) - r ; df/dr = +(a,b)
                    unit(Yr[it] - rr, Yi[it] - ri, a, b, ms);
                    a = g * a; b = g * b;
                    Zor[it] = Zor[it] + a; Zoi[it] = Zoi[it] + b;
                    f4st(r + 4 * c, a); f4st(r + kp + 4 * c, b);
                } else {  // residual = R(phi) r - o ; df/dr = -R(-phi)(a,b)
                    float4 yr = rr * Cs[it] - ri * Sn[it], yi = f4fma(rr, Sn[it], ri * Cs[it]);
                    unit(yr - Or_[it], yi - Oi[it], a, b, ms);
                    a = g * a; b = g * b;
                    Zsr[it] = Zsr[it] + a; Zsi[it] = Zsi[it] + b;
                    Aphi[it] = Aphi[it] + (a * yi - b * yr);
                    f4st(r + 4 * c, -1.f * f4fma(a, Cs[it], b * Sn[it]));
                    f4st(r + kp + 4 * c, a * Sn[it] - b * Cs[it]);
                }
            }
        }
    }
    // p row receives d/dtheta = (1/div) d/dphi in its first half, zeros in the second
    // (the second half of a RotatE relation row is allocated but unused, RotatE.py:76).
    __device__ __forceinline__ void finish(float *s, float *p, float *o, float gP, float inv_div)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = lane + 32 * it;
            if (c < nch) {
                float4 a, b;
                float ms;
                unit(Yr[it] - Or_[it], Yi[it] - Oi[it], a, b, ms);
                a = gP * a; b = gP * b;
                float4 Zr = a + Zor[it], Zi = b + Zoi[it];  // everything with y = R(phi) s
                f4st(s + 4 * c, -1.f * f4fma(Zr, Cs[it], Zi * Sn[it]));
                f4st(s + kp + 4 * c, Zr * Sn[it] - Zi * Cs[it]);
                f4st(o + 4 * c, a + Zsr[it]);
                f4st(o + kp + 4 * c, b + Zsi[it]);
                f4st(p + 4 * c, inv_div * (Aphi[it] + (Zr * Yi[it] - Zi * Yr[it])));
                f4st(p + kp + 4 * c, f4zero());
            }
        }
    }
};

// --------------------------------------------------------------------------
// per-positive loss and dL/dscore (warp-cooperative; lanes stride over j).
// in: P, sc[j] = N_j.  out: sc[j] = dL/dN_j, returns loss_i, *dP.
// --------------------------------------------------------------------------
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }
#define KGE_CLIP_LO (-75.0f)  // loss_functions.py:32
#define KGE_CLIP_HI (75.0f)   // loss_functions.py:35

__device__ float loss_and_dscores(const TrainParams &p, float P, float *sc, int lane, float *dP_out)
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

// --------------------------------------------------------------------------
// the kernel
// --------------------------------------------------------------------------
template <int MODEL, int NIT>
__global__ void __launch_bounds__(KGE_TRAIN_THREADS_FOR_NIT(NIT)) kge_train_kernel(const TrainParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *region = smem_raw + (size_t)warp * p.region_bytes;
    float *rows = reinterpret_cast<float *>(region);  // (3+G) rows of ld floats
    float *sc = reinterpret_cast<float *>(region + p.rows_bytes);
    int *nid = reinterpret_cast<int *>(sc + p.eta_pad);
    int *nside = nid + p.eta_pad;
    uint64_t *bar = reinterpret_cast<uint64_t *>(nside + p.eta_pad);

    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    __syncthreads();

    const int ld = p.ld, eta = p.eta, G = p.G;
    const uint32_t row_bytes = (uint32_t)ld * 4u;
    const int n_groups = (eta + G - 1) / G;
    float *srow = rows, *prow = rows + ld, *orow = rows + 2 * ld, *nrows = rows + 3 * ld;
    uint32_t phase = 0;
    double loss_acc = 0.0;

    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + warp; i < p.B; i += n_warps) {
        const int s_id = p.triples[3 * i], p_id = p.triples[3 * i + 1], o_id = p.triples[3 * i + 2];
        // ---- corruptions of this positive (A3) ----
        for (int j = lane; j < eta; j += 32) {
            int keep, repl;
            const unsigned long long r = (unsigned long long)j * (unsigned long long)p.B + (unsigned long long)i;
            if (p.neg_ent) { repl = p.neg_ent[r]; keep = p.neg_keep[r] ? 1 : 0; }
            else draw_corruption(p.seed, p.step, r, p.n_ent, &keep, &repl);
            nid[j] = repl;
            nside[j] = keep;  // keep_subj = 1 -> object replaced -> side 1
        }
        __syncwarp();
        // ---- gather s, p, o and the first group of replaced rows (A2) ----
        const int g0 = min(G, eta);
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(3 + g0) * row_bytes);
        __syncwarp();
        for (int r = lane; r < 3 + g0; r += 32) {
            const float *src = (r == 0)   ? p.ent + (size_t)s_id * ld
                               : (r == 1) ? p.rel + (size_t)p_id * ld
                               : (r == 2) ? p.ent + (size_t)o_id * ld
                                          : p.ent + (size_t)nid[r - 3] * ld;
            bulk_load(rows + (size_t)r * ld, src, row_bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;

        Scorer<MODEL, NIT> S;
        S.nch = p.nch;
        if constexpr (MODEL != KGE_TRANSE && MODEL != KGE_DISTMULT) S.kp = p.kp;
        const float P = warp_sum(S.prep(srow, prow, orow, p, lane));

        // ---- pass A: scores of all corruptions (A4) ----
        for (int g = 0; g < n_groups; ++g) {
            const int j0 = g * G, gs = min(G, eta - j0);
            if (g > 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)gs * row_bytes);
                __syncwarp();
                for (int r = lane; r < gs; r += 32)
                    bulk_load(nrows + (size_t)r * ld, p.ent + (size_t)nid[j0 + r] * ld, row_bytes, bar);
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
            for (int jj = 0; jj < gs; ++jj) {
                float v = warp_sum(S.neg_partial(nrows + (size_t)jj * ld, nside[j0 + jj]));
                if (lane == 0) sc[j0 + jj] = v;
            }
        }
        __syncwarp();

        // ---- loss and dL/dscore (A5) ----
        const float scale = p.score_scale;  // HolE 2/k, else 1
        float dP;
        if (p.mode != KGE_STEP_BACKWARD_EXT) {
            if (scale != 1.f)
                for (int j = lane; j < eta; j += 32) sc[j] *= scale;
            __syncwarp();
            if (p.scores_neg)
                for (int j = lane; j < eta; j += 32) p.scores_neg[(size_t)j * p.B + i] = sc[j];
            if (p.scores_pos && lane == 0) p.scores_pos[i] = scale * P;
            if (p.mode == KGE_STEP_FORWARD_ONLY) { __syncwarp(); continue; }
            float li = loss_and_dscores(p, scale * P, sc, lane, &dP);
            if (lane == 0) loss_acc += (double)li;
        } else {
            for (int j = lane; j < eta; j += 32) sc[j] = p.dneg[(size_t)j * p.B + i];
            dP = p.dpos[i];
        }
        __syncwarp();

        // ---- pass B: gradient rows, last group first (it is still resident) ----
        for (int g = n_groups - 1; g >= 0; --g) {
            const int j0 = g * G, gs = min(G, eta - j0);
            if (g != n_groups - 1) {
                bulk_wait_read_all();  // the copy engine must be done reading the previous group's rows
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)gs * row_bytes);
                __syncwarp();
                for (int r = lane; r < gs; r += 32)
                    bulk_load(nrows + (size_t)r * ld, p.ent + (size_t)nid[j0 + r] * ld, row_bytes, bar);
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
            for (int jj = 0; jj < gs; ++jj) S.neg_grad(nrows + (size_t)jj * ld, nside[j0 + jj], scale * sc[j0 + jj]);
            fence_proxy_async_smem();
            __syncwarp();
            for (int r = lane; r < gs; r += 32)
                bulk_reduce_add_f32(p.grad_ent + (size_t)nid[j0 + r] * ld, nrows + (size_t)r * ld, row_bytes);
            bulk_commit();
        }
        if constexpr (MODEL == KGE_ROTATE) S.finish(srow, prow, orow, scale * dP, p.inv_div);
        else S.finish(srow, prow, orow, scale * dP);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) bulk_reduce_add_f32(p.grad_ent + (size_t)s_id * ld, srow, row_bytes);
        if (lane == 1) bulk_reduce_add_f32(p.grad_rel + (size_t)p_id * ld, prow, row_bytes);
        if (lane == 2) bulk_reduce_add_f32(p.grad_ent + (size_t)o_id * ld, orow, row_bytes);
        bulk_commit();
        bulk_wait_read_all();  // slot is reused by the next positive's gather
        __syncwarp();
    }
    bulk_wait_all();
    if (p.loss_out && p.mode == KGE_STEP_FUSED && lane == 0 && loss_acc != 0.0) atomicAdd(p.loss_out, loss_acc);
}

// --------------------------------------------------------------------------
// RotatE rotation table: rot[r] = [cos(theta/div) | sin(theta/div)] (RotatE.py:96-98),
// canonical sin/cos so that training, predict and ranking agree bit for bit.
// --------------------------------------------------------------------------
__global__ void kge_rotation_table_kernel(const float *__restrict__ rel, float *__restrict__ rot, long long n_rel,
                                          int kp, int ld, float div)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rel * kp) return;
    long long r = idx / kp;
    int d = (int)(idx - r * kp);
    float s, c;
    kge_sincosf(__fdiv_rn(rel[r * ld + d], div), &s, &c);
    rot[r * ld + d] = c;
    rot[r * ld + kp + d] = s;
}

// --------------------------------------------------------------------------
// materialise the Philox corruption stream (inspection / parity only)
// --------------------------------------------------------------------------
__global__ void kge_corruptions_kernel(const int32_t *__restrict__ triples, long long B, int eta,
                                       unsigned long long seed, unsigned long long step, unsigned n_ent,
                                       int32_t *__restrict__ out)
{
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B * eta) return;
    long long i = r % B;
    int keep, repl;
    draw_corruption(seed, step, (unsigned long long)r, n_ent, &keep, &repl);
    out[3 * r + 0] = keep ? triples[3 * i + 0] : repl;
    out[3 * r + 1] = triples[3 * i + 1];
    out[3 * r + 2] = keep ? repl : triples[3 * i + 2];
}

// --------------------------------------------------------------------------
// host launchers
// --------------------------------------------------------------------------
template <int MODEL>
static cudaError_t launch_train_model(const TrainParams &p, int nit, int sm_count, int threads, size_t smem,
                                      cudaStream_t st)
{
#define KGE_LAUNCH(N)                                                                                       \
    {                                                                                                       \
        auto kern = kge_train_kernel<MODEL, N>;                                                             \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                     \
        int occ = 0;                                                                                        \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);                       \
        if (e != cudaSuccess) return e;                                                                     \
        if (occ < 1) return cudaErrorLaunchOutOfResources;                                                  \
        long long want = (p.B + (threads / 32) - 1) / (threads / 32);                                       \
        long long cap = (long long)occ * sm_count;                                                          \
        int grid = (int)(want < cap ? want : cap);                                                          \
        kern<<<grid, threads, smem, st>>>(p);                                                               \
        return cudaGetLastError();                                                                          \
    }
    switch (nit) {
    case 1: KGE_LAUNCH(1)
    case 2: KGE_LAUNCH(2)
    case 4: KGE_LAUNCH(4)
    default: return cudaErrorInvalidValue;
    }
#undef KGE_LAUNCH
}

cudaError_t launch_train(const TrainParams &p, int nit, int sm_count, int threads, size_t smem, cudaStream_t st)
{
    if (p.B == 0) return cudaSuccess;
    switch (p.model) {
    case KGE_TRANSE: return launch_train_model<KGE_TRANSE>(p, nit, sm_count, threads, smem, st);
    case KGE_DISTMULT: return launch_train_model<KGE_DISTMULT>(p, nit, sm_count, threads, smem, st);
    case KGE_COMPLEX: return launch_train_model<KGE_COMPLEX>(p, nit, sm_count, threads, smem, st);
    case KGE_HOLE: return launch_train_model<KGE_HOLE>(p, nit, sm_count, threads, smem, st);
    case KGE_ROTATE: return launch_train_model<KGE_ROTATE>(p, nit, sm_count, threads, smem, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_rotation_table(const float *rel, float *rot, long long n_rel, int kp, int ld, float div,
                                  cudaStream_t st)
{
    long long n = n_rel * kp;
    if (n == 0) return cudaSuccess;
    kge_rotation_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rel, rot, n_rel, kp, ld, div);
    return cudaGetLastError();
}

cudaError_t launch_corruptions(const int32_t *triples, long long B, int eta, unsigned long long seed,
                               unsigned long long step, unsigned n_ent, int32_t *out, cudaStream_t st)
{
    long long n = B * eta;
    if (n == 0) return cudaSuccess;
    kge_corruptions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(triples, B, eta, seed, step, n_ent, out);
    return cudaGetLastError();
}

}  // namespace kge
