// kge_rank.cu -- score every test triple against every candidate entity and count
// ranks (sm_100a).
//
// Replaces AbstractScoringLayer.get_ranks (layers/scoring/AbstractScoringLayer.py:156-422)
// and the five _get_subject_corruption_scores / _get_object_corruption_scores
// (TransE.py:56-114, DistMult.py:51-99, ComplEx.py:65-151, HolE.py:47-89,
// RotatE.py:107-217).  The reference materialises a [b, E, K] broadcast; here the
// candidate table is streamed ONCE per batch of queries through shared memory and
// only int32 counters leave the SM.
//
// Bit-exactness.  Ranks are counts of int32(score*1000) comparisons
// (AbstractScoringLayer.py:11,:201), so every score must be bit-identical to the
// oracle's.  Every (query, candidate) accumulator therefore runs the canonical
// chain -- ascending column index, one explicitly-rounded operation at a time
// (__fmaf_rn/__fadd_rn/__fsqrt_rn) -- the same sequence oracle/kge_oracle.c states
// in C.  This rules out tensor cores (TF32/bf16 products, unspecified accumulation
// order): the kernel is an fp32-FMA register-tiled contraction.  Zero pad columns
// are exact no-ops in every chain (fma(0,0,a)=a, a+|0|=a, a+sqrt(0)=a).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "kge_rank_common.cuh"

namespace kge {

// --------------------------------------------------------------------------
// prepare: per-query vectors for both sides + quantised positive score
// --------------------------------------------------------------------------
__device__ __forceinline__ const float *shard_row(const ShardView &v, const float *ent, int id, int ld)
{
    if (v.world <= 1) return ent + (size_t)id * ld;
    const int q = id / v.rows_per_shard;
    return v.ent[q] + (size_t)(id - q * v.rows_per_shard) * ld;
}

__global__ void kge_rank_qvec_kernel(int model, const ShardView sv, const float *__restrict__ ent, const float *__restrict__ rel,
                                     const float *__restrict__ rot, const int32_t *__restrict__ triples, long long b,
                                     int kp, int ld, float *__restrict__ qs, float *__restrict__ qo,
                                     float *__restrict__ qaux)
{
    const int halves = model_halves(model);
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * kp) return;
    long long i = idx / kp;
    int d = (int)(idx - i * kp);
    const float *s = shard_row(sv, ent, triples[3 * i], ld);
    const float *o = shard_row(sv, ent, triples[3 * i + 2], ld);
    const size_t prow = (size_t)triples[3 * i + 1] * ld;
    float *os = qs + i * ld, *oo = qo + i * ld;
    if (halves == 1) {
        float p = rel[prow + d];
        if (model == KGE_TRANSE) { os[d] = __fsub_rn(p, o[d]); oo[d] = __fadd_rn(s[d], p); }
        else { os[d] = __fmul_rn(p, o[d]); oo[d] = __fmul_rn(s[d], p); }
    } else if (model == KGE_ROTATE) {
        float c = rot[prow + d], sn = rot[prow + kp + d];
        os[d] = c; os[kp + d] = sn;
        qaux[i * ld + d] = o[d]; qaux[i * ld + kp + d] = o[kp + d];
        oo[d] = __fmaf_rn(-s[kp + d], sn, __fmul_rn(s[d], c));
        oo[kp + d] = __fmaf_rn(s[kp + d], c, __fmul_rn(s[d], sn));
    } else {
        float pr = rel[prow + d], pi = rel[prow + kp + d];
        os[d] = __fmaf_rn(pi, o[kp + d], __fmul_rn(pr, o[d]));        // ComplEx.py:95-99
        os[kp + d] = __fmaf_rn(-pi, o[d], __fmul_rn(pr, o[kp + d]));   // ComplEx.py:101-106
        oo[d] = __fmaf_rn(-s[kp + d], pi, __fmul_rn(s[d], pr));        // ComplEx.py:139-143
        oo[kp + d] = __fmaf_rn(s[d], pi, __fmul_rn(s[kp + d], pr));    // ComplEx.py:145-149
    }
}

// Positive score.  Its canonical chain (oracle: kgeo_score_triple) is, operation for operation, the corruption-score chain
// of the triple's OWN entity on one side: TransE |(s+p) - o| and RotatE |R(s) - o| are the object-side chains with
// candidate o (query vector s+p / R(s)); DistMult fma(s*p, o, acc) is the object-side chain with candidate o (query
// vector s*p); ComplEx / HolE fma(s_re, p_re o_re + p_im o_im, acc), fma(s_im, ...) is the subject-side chain with candidate s.
// So the positive is scored by the same warp-cooperative pair scorer as the filter entries, from the query vectors
// kge_rank_qvec_kernel has just written.
template <int OP>
__global__ void kge_rank_qpos_kernel(const ShardView sv, const float *__restrict__ ent, const int32_t *__restrict__ triples,
                                     long long b, int col, const float *__restrict__ qvec, const float *__restrict__ qaux, int kp,
                                     int ld, float scale, int32_t *__restrict__ qpos)
{
    extern __shared__ __align__(16) float pair_sm[];
    constexpr int ROWS = OP == OP_ROT_S ? 3 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    float *sm = pair_sm + (size_t)warp * ROWS * ld;
    for (long long i = (long long)blockIdx.x * wpc + warp; i < b; i += (long long)gridDim.x * wpc) {
        const float *e = shard_row(sv, ent, triples[3 * i + col], ld);
        const float sc = pair_score_warp<OP>(e, qvec + (size_t)i * ld, qaux ? qaux + (size_t)i * ld : nullptr, ld, kp, scale, sm, lane);
        if (lane == 0) qpos[i] = quantise(sc);
    }
}

cudaError_t launch_rank_prepare(const Layout &L, const ShardView &sv, const float *ent, const float *rel, const float *rot,
                                const int32_t *triples, long long b, float scale, float *qvec_s, float *qvec_o,
                                float *qaux, int32_t *qpos, cudaStream_t st)
{
    if (b == 0) return cudaSuccess;
    long long n = b * L.kp;
    kge_rank_qvec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.model, sv, ent, rel, rot, triples, b, L.kp, L.ld,
                                                                       qvec_s, qvec_o, qaux);
    const int wpc = pair_score_warps(L.ld, 2);
    const size_t sm = (size_t)wpc * 2 * L.ld * sizeof(float);
    const unsigned grid = (unsigned)((b + wpc - 1) / wpc);
#define KGE_QPOS(OP, COL, QV)                                                                                        \
    {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(kge_rank_qpos_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        if (e != cudaSuccess) return e;                                                                              \
        kge_rank_qpos_kernel<OP><<<grid, wpc * 32, sm, st>>>(sv, ent, triples, b, COL, QV, nullptr, L.kp, L.ld, scale, qpos); \
    }
    switch (L.model) {
    case KGE_TRANSE: KGE_QPOS(OP_L1_SUB, 2, qvec_o) break;
    case KGE_ROTATE: KGE_QPOS(OP_ROT_O, 2, qvec_o) break;
    case KGE_DISTMULT: KGE_QPOS(OP_DOT, 2, qvec_o) break;
    default: KGE_QPOS(OP_DOT, 0, qvec_s) break;  // ComplEx, HolE
    }
#undef KGE_QPOS
    return cudaGetLastError();
}

// --------------------------------------------------------------------------
// tile kernel: CTA = 256 candidates x 32 queries, 8 warps = 2 candidate groups x 4
// query groups, thread tile = 4 candidates (its own rows) x 8 queries (broadcast).
// Operands are staged column-chunk by column-chunk with cp.async (16 B, zero-fill
// past the row end), double buffered.  Shared rows are padded to 36 floats so the
// per-thread 128-bit row reads are bank-conflict free.
// --------------------------------------------------------------------------
constexpr int RK_TC = 256, RK_TQ = 32, RK_DK = 32, RK_LDS = RK_DK + 4, RK_THREADS = 256;
constexpr int RK_E_FLOATS = RK_TC * RK_LDS, RK_Q_FLOATS = RK_TQ * RK_LDS;
constexpr int RK_STAGE_FLOATS = RK_E_FLOATS + 2 * RK_Q_FLOATS;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid)
{
    unsigned sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int OP>
__global__ void __launch_bounds__(RK_THREADS) kge_rank_tile_kernel(const RankParams p, int32_t *__restrict__ cnt)
{
    constexpr bool ROT = (OP == OP_ROT_S || OP == OP_ROT_O);
    extern __shared__ __align__(128) float smem[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cg = warp & 1, qg = warp >> 1;
    const long long c0 = (long long)blockIdx.x * RK_TC, q0 = (long long)blockIdx.y * RK_TQ;
    const int ld = p.L.ld, kp = p.L.kp;
    const int n_chunks = ROT ? (kp + 15) / 16 : (ld + RK_DK - 1) / RK_DK;
    const int col_end = ROT ? kp : ld;  // valid columns per half (ROT) / per row

    // rows this thread stages: E rows t/8 + 32n (n<8), column quad t%8; Q row t/8
    const int lrow = t >> 3, c4 = t & 7;
    const float *erow[8];
    bool evalid[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        long long c = c0 + lrow + 32 * n;
        evalid[n] = c < p.n_cand;
        long long id = evalid[n] ? (p.cand_ids ? (long long)p.cand_ids[c] : p.cand_begin + c) : 0;
        erow[n] = p.ent + (size_t)id * ld;
    }
    const bool qvalid = (q0 + lrow) < p.b;
    const float *qrow = p.qvec + (size_t)(qvalid ? q0 + lrow : 0) * ld;
    const float *arow = (OP == OP_ROT_S) ? p.qaux + (size_t)(qvalid ? q0 + lrow : 0) * ld : nullptr;

    auto stage_load = [&](int chunk, int buf) {
        float *Es = smem + buf * RK_STAGE_FLOATS, *Qs = Es + RK_E_FLOATS, *As = Qs + RK_Q_FLOATS;
        int col;  // global column of this thread's float4
        bool cvalid;
        if (ROT) {
            int d = chunk * 16 + 4 * (c4 & 3);
            cvalid = d < col_end;
            col = (c4 < 4 ? 0 : kp) + d;
        } else {
            col = chunk * RK_DK + 4 * c4;
            cvalid = col < col_end;
        }
        if (!cvalid) col = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            cp_async16(Es + (lrow + 32 * n) * RK_LDS + 4 * c4, erow[n] + col, cvalid && evalid[n]);
        cp_async16(Qs + lrow * RK_LDS + 4 * c4, qrow + col, cvalid && qvalid);
        if (OP == OP_ROT_S) cp_async16(As + lrow * RK_LDS + 4 * c4, arow + col, cvalid && qvalid);
        cp_async_commit();
    };

    float acc[4][8];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[m][i] = 0.f;

    stage_load(0, 0);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < n_chunks) { stage_load(ch + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float *Es = smem + buf * RK_STAGE_FLOATS + (cg * 128 + lane) * RK_LDS;
        const float *Qs = smem + buf * RK_STAGE_FLOATS + RK_E_FLOATS + (qg * 8) * RK_LDS;
        const float *As = Qs + RK_Q_FLOATS;
        if (!ROT) {
#pragma unroll
            for (int d4 = 0; d4 < RK_DK / 4; ++d4) {
                float4 e[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) e[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RK_LDS + 4 * d4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 q = *reinterpret_cast<const float4 *>(Qs + i * RK_LDS + 4 * d4);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        float a = acc[m][i];
                        a = rank_step<OP>(a, e[m].x, q.x);
                        a = rank_step<OP>(a, e[m].y, q.y);
                        a = rank_step<OP>(a, e[m].z, q.z);
                        a = rank_step<OP>(a, e[m].w, q.w);
                        acc[m][i] = a;
                    }
                }
            }
        } else {
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
                float4 er[4], ei[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    er[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RK_LDS + 4 * d4);
                    ei[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RK_LDS + 16 + 4 * d4);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 qa = *reinterpret_cast<const float4 *>(Qs + i * RK_LDS + 4 * d4);
                    float4 qb = *reinterpret_cast<const float4 *>(Qs + i * RK_LDS + 16 + 4 * d4);
                    float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
                    if (OP == OP_ROT_S) {
                        oa = *reinterpret_cast<const float4 *>(As + i * RK_LDS + 4 * d4);
                        ob = *reinterpret_cast<const float4 *>(As + i * RK_LDS + 16 + 4 * d4);
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        float a = acc[m][i];
                        a = rank_step_rot<OP>(a, er[m].x, ei[m].x, qa.x, qb.x, oa.x, ob.x);
                        a = rank_step_rot<OP>(a, er[m].y, ei[m].y, qa.y, qb.y, oa.y, ob.y);
                        a = rank_step_rot<OP>(a, er[m].z, ei[m].z, qa.z, qb.z, oa.z, ob.z);
                        a = rank_step_rot<OP>(a, er[m].w, ei[m].w, qa.w, qb.w, oa.w, ob.w);
                        acc[m][i] = a;
                    }
                }
            }
        }
        __syncthreads();
    }

    // epilogue: quantise, compare with the positive, count over this warp's 128 candidates
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long q = q0 + qg * 8 + i;
        const bool qok = q < p.b;
        const int qp = qok ? p.qpos[q] : 0;
        int gt = 0, eq = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const long long c = c0 + cg * 128 + lane + 32 * m;
            const bool cok = c < p.n_cand;
            const float sc = rank_finish<OP>(acc[m][i], p.scale);
            if (p.scores && qok && cok) p.scores[(size_t)q * p.n_cand + c] = sc;
            const int qc = quantise(sc);
            gt += __popc(__ballot_sync(0xffffffffu, cok && (qp < qc)));
            eq += __popc(__ballot_sync(0xffffffffu, cok && (qp == qc)));
        }
        if (lane == 0 && qok) {
            if (gt) atomicAdd(&cnt[3 * q + 0], gt);
            if (eq) atomicAdd(&cnt[3 * q + 1], eq);
        }
    }
}

// --------------------------------------------------------------------------
// DOT specialisation (DistMult / ComplEx / HolE): CTA = 256 candidates x 64 queries, thread tile =
// 4 candidates (its own rows) x 16 queries, accumulators as 32 f32x2 pairs of adjacent queries.
// The query tile is staged TRANSPOSED ([column][query]) so that one broadcast 128-bit read returns
// four queries of one column, and each Blackwell packed FFMA2 (fma.rn.f32x2 with the candidate value
// as the broadcast scalar operand) advances two canonical chains at once: same IEEE fma per chain,
// same ascending column order, half the FP issue slots.
// --------------------------------------------------------------------------
constexpr int RD_TC = 256, RD_TQ = 64, RD_DK = 32, RD_ELDS = RD_DK + 4, RD_QLDS = RD_TQ + 4, RD_THREADS = 256;
constexpr int RD_E_FLOATS = RD_TC * RD_ELDS, RD_Q_FLOATS = RD_DK * RD_QLDS;
constexpr int RD_STAGE_FLOATS = RD_E_FLOATS + RD_Q_FLOATS;

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src, bool valid)
{
    unsigned sz = valid ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ unsigned long long rk_pk2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long rk_fma2(float e, unsigned long long q, unsigned long long acc)
{
    unsigned long long d, ee = rk_pk2(e, e);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(ee), "l"(q), "l"(acc));
    return d;
}

__global__ void __launch_bounds__(RD_THREADS, 2) kge_rank_dot_kernel(const RankParams p, int32_t *__restrict__ cnt)
{
    if (p.gate && *p.gate <= p.gate_cap) return;  // armed only as the overflow fallback of the tensor-core filter
    extern __shared__ __align__(128) float smem[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cg = warp & 1, qg = warp >> 1;  // 2 candidate groups of 128, 4 query groups of 16
    const long long c0 = (long long)blockIdx.x * RD_TC, q0 = (long long)blockIdx.y * RD_TQ;
    const int ld = p.L.ld;
    const int n_chunks = (ld + RD_DK - 1) / RD_DK;

    // E staging: rows t/8 + 32n (n<8), column quad t%8 (16-byte copies)
    const int lrow = t >> 3, c4 = t & 7;
    const float *erow[8];
    bool evalid[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        long long c = c0 + lrow + 32 * n;
        evalid[n] = c < p.n_cand;
        long long id = evalid[n] ? (p.cand_ids ? (long long)p.cand_ids[c] : p.cand_begin + c) : 0;
        erow[n] = p.ent + (size_t)id * ld;
    }
    // Q staging (transposed): column t%32, queries t/32 + 8n (n<8), 4-byte copies
    const int qd = t & 31, qq = t >> 5;

    auto stage_load = [&](int chunk, int buf) {
        float *Es = smem + buf * RD_STAGE_FLOATS, *Qs = Es + RD_E_FLOATS;
        const int col = chunk * RD_DK + 4 * c4;
        const bool cvalid = col < ld;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            cp_async16(Es + (lrow + 32 * n) * RD_ELDS + 4 * c4, erow[n] + (cvalid ? col : 0), cvalid && evalid[n]);
        const int qcol = chunk * RD_DK + qd;
        const bool qcvalid = qcol < ld;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const long long q = q0 + qq + 8 * n;
            const bool ok = qcvalid && q < p.b;
            cp_async4(Qs + qd * RD_QLDS + qq + 8 * n, p.qvec + (size_t)(ok ? q : 0) * ld + (ok ? qcol : 0), ok);
        }
        cp_async_commit();
    };

    unsigned long long acc[4][8];  // [candidate][query pair]
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[m][i] = 0ull;

    stage_load(0, 0);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < n_chunks) { stage_load(ch + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float *Es = smem + buf * RD_STAGE_FLOATS + (cg * 128 + lane) * RD_ELDS;
        const float *Qs = smem + buf * RD_STAGE_FLOATS + RD_E_FLOATS + qg * 16;
#pragma unroll
        for (int d4 = 0; d4 < RD_DK / 4; ++d4) {
            float4 e[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) e[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RD_ELDS + 4 * d4);
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
                const float *qrow = Qs + (4 * d4 + dd) * RD_QLDS;
                unsigned long long q2[8];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(qrow + 4 * v);  // 4 queries
                    q2[2 * v] = w.x;
                    q2[2 * v + 1] = w.y;
                }
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const float ev = dd == 0 ? e[m].x : dd == 1 ? e[m].y : dd == 2 ? e[m].z : e[m].w;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[m][i] = rk_fma2(ev, q2[i], acc[m][i]);
                }
            }
        }
        __syncthreads();
    }

    // epilogue: quantise, compare with the positive, count over this warp's 128 candidates
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const long long q = q0 + qg * 16 + i;
        const bool qok = q < p.b;
        const int qp = qok ? p.qpos[q] : 0;
        int gt = 0, eq = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[m][i >> 1]));
            const float a = (i & 1) ? hi : lo;
            const long long c = c0 + cg * 128 + lane + 32 * m;
            const bool cok = c < p.n_cand;
            const float sc = rank_finish<OP_DOT>(a, p.scale);
            if (p.scores && qok && cok) p.scores[(size_t)q * p.n_cand + c] = sc;
            const int qc = quantise(sc);
            gt += __popc(__ballot_sync(0xffffffffu, cok && (qp < qc)));
            eq += __popc(__ballot_sync(0xffffffffu, cok && (qp == qc)));
        }
        if (lane == 0 && qok) {
            if (gt) atomicAdd(&cnt[3 * q + 0], gt);
            if (eq) atomicAdd(&cnt[3 * q + 1], eq);
        }
    }
}

// --------------------------------------------------------------------------
// TransE / RotatE specialisation ("pair" kernel): CTA = 256 candidates x 32 queries (the tile kernel's geometry), thread
// tile = 4 candidates x 8 queries, accumulators as 16 f32x2 pairs of ADJACENT QUERIES.  As in the DOT kernel the query
// tile (and, on RotatE's subject side, the tile of object rows) is staged transposed ([column][query]), so one broadcast
// 128-bit read returns four queries of one column and every packed instruction advances two canonical chains:
//   TransE: acc += |e + q| (or |q - e|) is FADD2 + FADD2 with the absolute value as an operand modifier: one issue slot
//     per chain step instead of two;
//   RotatE, per pair of chain steps: 2 FADD2 + FMUL2 + FFMA2 (residual, x = re^2 + im^2; the subject side adds the rotation:
//     2 FMUL2 + 2 FFMA2), 2 FMNMX + 2 MUFU.RSQ, 2 FMUL2 + 2 FFMA2 (the correctly rounded sqrt of kge_rank_common.cuh,
//     same five operations per chain), 2 FSET and one FFMA2 that adds [x >= 2^-101] * sqrt -- 15 issue slots instead of 26.
// Every operation is the IEEE operation of the scalar chain (rank_step / rank_step_rot) on the same operands in the same
// order: a - b, a * b and fma are sign-symmetric, negations and absolute values are operand modifiers, fma(r, 1, acc) ==
// acc + r and fma(r, 0, acc) == acc for finite r.
// --------------------------------------------------------------------------
constexpr int RR_QLDS = RK_TQ + 4;                 // row stride of the transposed query tile
constexpr int RR_Q_FLOATS = 32 * RR_QLDS;          // 32 tile columns (RotatE: 16 re + 16 im)
constexpr int RR_STAGE_FLOATS = RK_E_FLOATS + 2 * RR_Q_FLOATS;

__device__ __forceinline__ unsigned long long rk_add2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long rk_sub2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long rk_mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long rk_mul2_ftz(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long rk_fma2v(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void rk_upk2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// acc + |v| for two chains (the absolute value becomes an operand modifier of the FADD2)
__device__ __forceinline__ unsigned long long rk_add_abs2(unsigned long long acc, unsigned long long v)
{
    float lo, hi;
    rk_upk2(v, lo, hi);
    return rk_add2(acc, rk_pk2(fabsf(lo), fabsf(hi)));
}

// acc += sqrt_rn_nonneg(re^2 + im^2) for two chains at once (see sqrt_rn_nonneg for the five-operation sqrt)
__device__ __forceinline__ unsigned long long rk_mod_step2(unsigned long long acc, unsigned long long re, unsigned long long im)
{
    const float lo = 3.9443045e-31f;  // 2^-101
    const unsigned long long x2 = rk_fma2v(im, im, rk_mul2(re, re));
    float x0, x1, y0, y1, s0, s1;
    rk_upk2(x2, x0, x1);
    const float c0 = fmaxf(x0, lo), c1 = fmaxf(x1, lo);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(c0));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(c1));
    const unsigned long long xc = rk_pk2(c0, c1), y = rk_pk2(y0, y1);
    const unsigned long long s = rk_mul2_ftz(xc, y), h = rk_mul2_ftz(y, rk_pk2(0.5f, 0.5f));
    rk_upk2(s, s0, s1);
    const unsigned long long r = rk_fma2v(rk_pk2(-s0, -s1), s, xc);  // the negation becomes an operand modifier of the FFMA2
    const unsigned long long res = rk_fma2v(r, h, s);
    const unsigned long long flag = rk_pk2(x0 >= lo ? 1.f : 0.f, x1 >= lo ? 1.f : 0.f);
    return rk_fma2v(res, flag, acc);
}

template <int OP>
__global__ void __launch_bounds__(RK_THREADS, 2) kge_rank_pair_kernel(const RankParams p, int32_t *__restrict__ cnt)
{
    static_assert(OP != OP_DOT, "the bilinear models have kge_rank_dot_kernel");
    constexpr bool ROT = (OP == OP_ROT_S || OP == OP_ROT_O);
    extern __shared__ __align__(128) float smem[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cg = warp & 1, qg = warp >> 1;  // 2 candidate groups of 128, 4 query groups of 8
    const long long c0 = (long long)blockIdx.x * RK_TC, q0 = (long long)blockIdx.y * RK_TQ;
    const int ld = p.L.ld, kp = p.L.kp;
    const int n_chunks = ROT ? (kp + 15) / 16 : (ld + RK_DK - 1) / RK_DK;
    const int col_end = ROT ? kp : ld;  // valid columns per half (ROT) / per row

    // E staging as in the tile kernel: rows t/8 + 32n (n<8), column quad t%8 (ROT: quads 0-3 re columns, 4-7 im columns)
    const int lrow = t >> 3, c4 = t & 7;
    const float *erow[8];
    bool evalid[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        long long c = c0 + lrow + 32 * n;
        evalid[n] = c < p.n_cand;
        long long id = evalid[n] ? (p.cand_ids ? (long long)p.cand_ids[c] : p.cand_begin + c) : 0;
        erow[n] = p.ent + (size_t)id * ld;
    }
    // Q (and A) staging, transposed: tile column t%32 (ROT: 0-15 re, 16-31 im), queries t/32 + 8n (n<4), 4-byte copies
    const int qd = t & 31, qq = t >> 5;

    auto stage_load = [&](int chunk, int buf) {
        float *Es = smem + buf * RR_STAGE_FLOATS, *Qs = Es + RK_E_FLOATS, *As = Qs + RR_Q_FLOATS;
        {
            const int d = ROT ? chunk * 16 + 4 * (c4 & 3) : chunk * RK_DK + 4 * c4;
            const bool cvalid = d < col_end;
            const int col = cvalid ? (ROT && c4 >= 4 ? kp : 0) + d : 0;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                cp_async16(Es + (lrow + 32 * n) * RK_LDS + 4 * c4, erow[n] + col, cvalid && evalid[n]);
        }
        const int d = ROT ? chunk * 16 + (qd & 15) : chunk * RK_DK + qd;
        const bool qcvalid = d < col_end;
        const int qcol = (ROT && qd >= 16 ? kp : 0) + d;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const long long q = q0 + qq + 8 * n;
            const bool ok = qcvalid && q < p.b;
            const size_t src = (size_t)(ok ? q : 0) * ld + (ok ? qcol : 0);
            cp_async4(Qs + qd * RR_QLDS + qq + 8 * n, p.qvec + src, ok);
            if (OP == OP_ROT_S) cp_async4(As + qd * RR_QLDS + qq + 8 * n, p.qaux + src, ok);
        }
        cp_async_commit();
    };
    // four adjacent query pairs of one tile column
    auto ldq = [](const float *row, unsigned long long (&v)[4]) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(row), b = *reinterpret_cast<const ulonglong2 *>(row + 4);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    };
    auto comp = [](const float4 &v, int dd) { return dd == 0 ? v.x : dd == 1 ? v.y : dd == 2 ? v.z : v.w; };

    unsigned long long acc[4][4];  // [candidate][query pair]
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][i] = 0ull;

    stage_load(0, 0);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < n_chunks) { stage_load(ch + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float *Es = smem + buf * RR_STAGE_FLOATS + (cg * 128 + lane) * RK_LDS;
        const float *Qs = smem + buf * RR_STAGE_FLOATS + RK_E_FLOATS + qg * 8;
        const float *As = Qs + RR_Q_FLOATS;
        if constexpr (!ROT) {
#pragma unroll
            for (int d4 = 0; d4 < RK_DK / 4; ++d4) {
                float4 e[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) e[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RK_LDS + 4 * d4);
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) {
                    unsigned long long q[4];
                    ldq(Qs + (4 * d4 + dd) * RR_QLDS, q);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float ev = comp(e[m], dd);
                        const unsigned long long e2 = rk_pk2(ev, ev);
#pragma unroll
                        for (int i = 0; i < 4; ++i)  // TransE.py:78-84 (e + q) / :107-113 (q - e)
                            acc[m][i] = rk_add_abs2(acc[m][i], OP == OP_L1_ADD ? rk_add2(e2, q[i]) : rk_sub2(q[i], e2));
                    }
                }
            }
        } else {
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
                float4 er[4], ei[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    er[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RK_LDS + 4 * d4);
                    ei[m] = *reinterpret_cast<const float4 *>(Es + 32 * m * RK_LDS + 16 + 4 * d4);
                }
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) {
                    const int col = 4 * d4 + dd;
                    unsigned long long qa[4], qb[4], oa[4], ob[4];
                    ldq(Qs + col * RR_QLDS, qa);
                    ldq(Qs + (16 + col) * RR_QLDS, qb);
                    if (OP == OP_ROT_S) { ldq(As + col * RR_QLDS, oa); ldq(As + (16 + col) * RR_QLDS, ob); }
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float evr = comp(er[m], dd), evi = comp(ei[m], dd);
                        const unsigned long long er2 = rk_pk2(evr, evr), ei2 = rk_pk2(evi, evi);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            unsigned long long re, im;
                            if (OP == OP_ROT_S) {  // RotatE.py:151-163: qa = cos, qb = sin, (oa, ob) = the object row
                                re = rk_sub2(rk_fma2v(rk_pk2(-evi, -evi), qb[i], rk_mul2(er2, qa[i])), oa[i]);
                                im = rk_sub2(rk_fma2v(ei2, qa[i], rk_mul2(er2, qb[i])), ob[i]);
                            } else {  // RotatE.py:208-216: (qa, qb) = the rotated subject
                                re = rk_sub2(qa[i], er2);
                                im = rk_sub2(qb[i], ei2);
                            }
                            acc[m][i] = rk_mod_step2(acc[m][i], re, im);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // epilogue: quantise, compare with the positive, count over this warp's 128 candidates
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long q = q0 + qg * 8 + i;
        const bool qok = q < p.b;
        const int qp = qok ? p.qpos[q] : 0;
        int gt = 0, eq = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            float lo, hi;
            rk_upk2(acc[m][i >> 1], lo, hi);
            const long long c = c0 + cg * 128 + lane + 32 * m;
            const bool cok = c < p.n_cand;
            const float sc = rank_finish<OP>((i & 1) ? hi : lo, p.scale);
            if (p.scores && qok && cok) p.scores[(size_t)q * p.n_cand + c] = sc;
            const int qc = quantise(sc);
            gt += __popc(__ballot_sync(0xffffffffu, cok && (qp < qc)));
            eq += __popc(__ballot_sync(0xffffffffu, cok && (qp == qc)));
        }
        if (lane == 0 && qok) {
            if (gt) atomicAdd(&cnt[3 * q + 0], gt);
            if (eq) atomicAdd(&cnt[3 * q + 1], eq);
        }
    }
}

cudaError_t launch_rank_count(const RankParams &p, int32_t *cnt, cudaStream_t st)
{
    if (p.b == 0 || p.n_cand == 0) return cudaSuccess;
    const int op = rank_op(p.L.model, p.side);
    if (op == OP_DOT) {
        dim3 gd((unsigned)((p.n_cand + RD_TC - 1) / RD_TC), (unsigned)((p.b + RD_TQ - 1) / RD_TQ));
        const size_t sm = 2 * RD_STAGE_FLOATS * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(kge_rank_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        kge_rank_dot_kernel<<<gd, RD_THREADS, sm, st>>>(p, cnt);
        return cudaGetLastError();
    }
    dim3 grid((unsigned)((p.n_cand + RK_TC - 1) / RK_TC), (unsigned)((p.b + RK_TQ - 1) / RK_TQ));
    // KGE_B200_RANK_KERNEL=tile: the scalar tile kernel (one chain per accumulator register) instead of the packed pair kernel --
    // an A/B and cross-check aid, both run the same canonical chains
    const char *force = getenv("KGE_B200_RANK_KERNEL");
    const bool tile = force && strcmp(force, "tile") == 0;
    const size_t smem = (tile ? 2 * RK_STAGE_FLOATS : 2 * RR_STAGE_FLOATS) * sizeof(float);
#define KGE_RK(OP)                                                                                           \
    case OP: {                                                                                               \
        auto kern = tile ? kge_rank_tile_kernel<OP> : kge_rank_pair_kernel<OP>;                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e != cudaSuccess) return e;                                                                      \
        kern<<<grid, RK_THREADS, smem, st>>>(p, cnt);                                                        \
        break;                                                                                               \
    }
    switch (op) {
    KGE_RK(OP_L1_ADD)
    KGE_RK(OP_L1_SUB)
    KGE_RK(OP_ROT_S)
    KGE_RK(OP_ROT_O)
    }
#undef KGE_RK
    return cudaGetLastError();
}

// --------------------------------------------------------------------------
// filter kernel (AbstractScoringLayer.py:260-307): one thread per (query, known-true
// candidate) pair recomputes that candidate's score with the same canonical chain and
// counts qpos <= q_f (always '<=', whatever the tie strategy).
// --------------------------------------------------------------------------
template <int OP>
__global__ void kge_rank_filter_kernel(const RankParams p, const long long *__restrict__ off,
                                       const int32_t *__restrict__ idx, int32_t *__restrict__ cnt)
{
    extern __shared__ __align__(16) float pair_sm[];
    constexpr int ROWS = OP == OP_ROT_S ? 3 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    float *sm = pair_sm + (size_t)warp * ROWS * p.L.ld;
    const long long total = off[p.b];
    for (long long f = (long long)blockIdx.x * wpc + warp; f < total; f += (long long)gridDim.x * wpc) {
        long long lo = 0, hi = p.b;  // largest i with off[i] <= f
        while (hi - lo > 1) {
            long long mid = (lo + hi) >> 1;
            if (off[mid] <= f) lo = mid; else hi = mid;
        }
        const long long i = lo;
        const long long pos = (long long)idx[f] - p.filt_base;  // filter ids are global; the shard's table is local
        if (pos < p.cand_begin || pos >= p.cand_begin + p.n_cand) continue;  // :280-288 (warp-uniform)
        const long long id = p.cand_ids ? (long long)p.cand_ids[pos] : pos;
        const float sc = pair_score_warp<OP>(p.ent + (size_t)id * p.L.ld, p.qvec + (size_t)i * p.L.ld,
                                             OP == OP_ROT_S ? p.qaux + (size_t)i * p.L.ld : nullptr, p.L.ld, p.L.kp, p.scale, sm, lane);
        if (lane == 0 && p.qpos[i] <= quantise(sc)) atomicAdd(&cnt[3 * i + 2], 1);
    }
}

cudaError_t launch_rank_filter_n(const RankParams &p, const long long *filt_off, const int32_t *filt_idx,
                                 long long n_pairs, int32_t *cnt, cudaStream_t st)
{
    if (n_pairs == 0 || p.b == 0) return cudaSuccess;
    const int op = rank_op(p.L.model, p.side);
    const int wpc = pair_score_warps(p.L.ld, op == OP_ROT_S ? 3 : 2);
    const size_t sm = (size_t)wpc * (op == OP_ROT_S ? 3 : 2) * p.L.ld * sizeof(float);
    long long want = (n_pairs + wpc - 1) / wpc;
    const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
#define KGE_FILT(OP)                                                                                                  \
    {                                                                                                                 \
        cudaError_t e = cudaFuncSetAttribute(kge_rank_filter_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        if (e != cudaSuccess) return e;                                                                               \
        kge_rank_filter_kernel<OP><<<grid, wpc * 32, sm, st>>>(p, filt_off, filt_idx, cnt);                           \
        break;                                                                                                        \
    }
    switch (op) {
    case OP_DOT: KGE_FILT(OP_DOT)
    case OP_L1_ADD: KGE_FILT(OP_L1_ADD)
    case OP_L1_SUB: KGE_FILT(OP_L1_SUB)
    case OP_ROT_S: KGE_FILT(OP_ROT_S)
    case OP_ROT_O: KGE_FILT(OP_ROT_O)
    }
#undef KGE_FILT
    return cudaGetLastError();
}

// ranks[i] += f(strategy)(gt, eq) - filtered   (AbstractScoringLayer.py:218-258, :304-307)
__global__ void kge_rank_finalize_kernel(const int32_t *__restrict__ cnt, long long b, int strategy,
                                         int32_t *__restrict__ ranks)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    int gt = cnt[3 * i], eq = cnt[3 * i + 1], fl = cnt[3 * i + 2];
    int r = (strategy == KGE_RANK_BEST) ? gt : (strategy == KGE_RANK_MIDDLE) ? gt + (eq + 1) / 2 : gt + eq;
    ranks[i] += r - fl;
}

__global__ void kge_rank_accumulate_kernel(const int32_t *__restrict__ cnt, long long n, int32_t *__restrict__ counts)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] += cnt[i];
}

cudaError_t launch_rank_accumulate(const int32_t *cnt, long long b, int32_t *counts, cudaStream_t st)
{
    if (b == 0) return cudaSuccess;
    kge_rank_accumulate_kernel<<<(unsigned)((3 * b + 255) / 256), 256, 0, st>>>(cnt, 3 * b, counts);
    return cudaGetLastError();
}

cudaError_t launch_rank_finalize(const int32_t *cnt, long long b, int strategy, int32_t *ranks, cudaStream_t st)
{
    if (b == 0) return cudaSuccess;
    kge_rank_finalize_kernel<<<(unsigned)((b + 255) / 256), 256, 0, st>>>(cnt, b, strategy, ranks);
    return cudaGetLastError();
}

}  // namespace kge
