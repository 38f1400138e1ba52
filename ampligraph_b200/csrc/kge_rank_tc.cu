// kge_rank_tc.cu -- tensor-core FILTER pass of the all-entity ranking for the bilinear models (sm_100a: tcgen05 + TMEM).
//
// For DistMult / ComplEx / HolE the corruption scores of a batch of queries are a contraction
// S[b, E] = Q[b, ld] . T[E, ld]^T (DistMult.py:51-99, ComplEx.py:65-151, HolE.py:47-89 compute it by broadcasting to
// [b, E, K]).  Ranks are counts of int32(score*1000) comparisons (AbstractScoringLayer.py:11,:201) and must be
// bit-identical to the canonical FP32 chain, which no tensor-core accumulation order reproduces.  But a rank only needs,
// per (query, candidate) pair, on which side of the positive's quantisation bin the candidate's score falls:
//
//   1. split      every fp32 operand x into two bf16 numbers, x ~ hi + lo (|x - hi - lo| <= 2^-16 |x|), laid out in HBM
//                 as ready-made UMMA operand tiles (8x8 "core matrices", K-major, no swizzle);
//   2. filter     (this kernel) approximate scores a = Qhi.Thi + Qhi.Tlo + Qlo.Thi on the 5th-generation tensor cores:
//                 tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), accumulators in TMEM, operand tiles staged by the copy
//                 engine (cp.async.bulk + mbarrier), one issuing thread, double-buffered accumulators so that the epilogue
//                 of one tile overlaps the MMAs of the next.  The epilogue (tcgen05.ld) compares a with the two edges
//                 U, L of the positive's bin: |a - s_canonical| <= delta is proven below, so a >= U + delta means
//                 "greater", L + delta <= a <= U - delta means "equal", a <= L - delta means "smaller";
//   3. refine     the few pairs inside a +-delta band around an edge are appended to a list and re-scored with the exact
//                 canonical chain (kge_rank_refine_kernel) -- same arithmetic as the FP32 kernel, same counters.
//   If the list overflows (it is sized for b*n_cand/16 pairs) the exact FP32 kernel redoes the whole call, gated on the
//   device-side counter: correctness never depends on the filter being selective.
//
// Error bound.  With u = 2^-8 (bf16 round-to-nearest): |x - hi| <= u|x|, |x - hi - lo| <= u^2 |x|.  The three computed
// products differ from q_d*e_d by the dropped terms lo*lo and the two residuals: <= 3 u^2 |q_d e_d| (+ O(u^3)); the
// bf16 x bf16 products themselves are exact in fp32.  The tensor core sums 3*ld products into an fp32 accumulator in an
// unspecified order; any order with fp32 roundings (truncation included, unit 2^-23) errs by <= 3 ld 2^-23 sum|q_d e_d|,
// and the canonical chain itself by <= ld 2^-24 sum|q_d e_d|.  With sum|q_d e_d| <= |q| |e| (Cauchy-Schwarz):
//     |a - s_canonical| <= (2^-14 + 4 ld 2^-23) |q|_2 |e|_2  =: eps_rel |q| |e|.
// tests/test_gpu_parity.py::test_tc_filter_error_bound measures the actual error (>= 10x smaller) and the bit-exact rank
// tests run both modes.
#include <cuda_bf16.h>
#include <math.h>

#include "kge_rank_common.cuh"

namespace kge {

// ---------------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TC_MQ = 128;                         // queries per A tile    (UMMA M, one TMEM lane per query)
constexpr int TC_NC = 256;                         // candidates per B tile (UMMA N, one TMEM column per candidate)
constexpr int TC_KB = 32;                          // columns per k-block (2 MMA k-steps of 16)
constexpr int TC_KC = TC_KB / 8;                   // 8-column chunks (core-matrix columns) per k-block
constexpr int TC_KSTEPS = TC_KB / 16;              // MMA k-steps per k-block
constexpr int TC_A_ELEMS = TC_MQ * TC_KB;          // bf16 elements of one part (hi or lo) of an A tile
constexpr int TC_B_ELEMS = TC_NC * TC_KB;
constexpr int TC_A_BYTES = 2 * TC_A_ELEMS * 2;     // hi + lo, contiguous in HBM and in shared memory: 16 KB
constexpr int TC_B_BYTES = 2 * TC_B_ELEMS * 2;     // 32 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;  // 48 KB per pipeline stage
constexpr int TC_STAGES = 4;                       // operand ring depth: the loop is bound by the latency of the bulk copies (a 96 KB
                                                   // stage x 2 measured 15 us per 7-k-block tile against 4.9 us of MMAs), so more,
                                                   // smaller stages in flight beat fewer, larger ones
constexpr int TC_THREADS = 192;                    // warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue
constexpr int TC_TMEM_COLS = 512;                  // two accumulator stages of 256 fp32 columns

// ---------------------------------------------------------------------------------------------------------------------
// 1. split: fp32 rows -> blocked bf16 hi/lo operand tiles.
//    tile (rb, kb) of a matrix with RB rows per block: [hi: RB x TC_KB][lo: RB x TC_KB]; inside a part, element (r, c) lives
//    at ((r/8 * TC_KC + c/8) * 8 + r%8) * 8 + c%8: 8x8 core matrices of 128 contiguous bytes, TC_KC of them (one k-block)
//    per 8-row group -> the canonical K-major SWIZZLE_NONE layout with LBO = 128 B (next k-chunk), SBO = TC_KC*128 B
//    (next 8-row group), so a whole tile is ONE contiguous bulk copy and needs no tensor map.
// ---------------------------------------------------------------------------------------------------------------------
// One WARP per row: lanes stride over the row's 8-column chunks (coalesced 32-byte reads), write the hi / lo core-matrix
// rows (16 bytes each) and accumulate the row's sum of squares on the way, so the norms the error bound needs cost no
// second pass.  Candidates (RB = 256): max |row| of each tile via atomicMax (positive floats order like their bit
// patterns).  Queries (RB = 128): the warp's lane 0 also writes the query's bin edges {U, L} in SCORE units (unscaled
// accumulator for HolE) and its error-bound factors.  qc >= n  <=>  trunc(x) >= n with x = fl(score*1000): x >= n for
// n >= 1, x > n-1 for n <= 0.
template <int RB>
__global__ void __launch_bounds__(256) kge_rank_split_kernel(const float *__restrict__ src, const int32_t *__restrict__ ids,
                                                             long long row_begin, long long n_rows, long long rows_pad,
                                                             int ld, int nkb, __nv_bfloat16 *__restrict__ out,
                                                             unsigned *__restrict__ tile_max, const int32_t *__restrict__ qpos,
                                                             float scale, float eps_rel, float4 *__restrict__ thr)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = nkb * TC_KC;
    for (long long r = warp0; r < rows_pad; r += n_warps) {
        const bool live = r < n_rows;
        const float *row = live ? src + (size_t)(ids ? (long long)ids[r] : row_begin + r) * ld : nullptr;
        const long long rb = r / RB;
        const int rl = (int)(r - rb * RB);
        float ss = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            const int kb = c / TC_KC, kc = c - kb * TC_KC, col0 = c * 8;
            float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) {
                if (col0 + 4 <= ld) { const float4 v = *reinterpret_cast<const float4 *>(row + col0); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
                if (col0 + 8 <= ld) { const float4 v = *reinterpret_cast<const float4 *>(row + col0 + 4); x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w; }
            }
            __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                hi[e] = __float2bfloat16_rn(x[e]);
                lo[e] = __float2bfloat16_rn(x[e] - __bfloat162float(hi[e]));  // x - hi is exact in fp32
                ss = fmaf(x[e], x[e], ss);
            }
            const size_t tile = ((size_t)rb * nkb + kb) * (size_t)(2 * RB * TC_KB);
            const size_t off = (size_t)(((rl >> 3) * TC_KC + kc) * 8 + (rl & 7)) * 8;
            *reinterpret_cast<uint4 *>(out + tile + off) = *reinterpret_cast<const uint4 *>(hi);
            *reinterpret_cast<uint4 *>(out + tile + (size_t)RB * TC_KB + off) = *reinterpret_cast<const uint4 *>(lo);
        }
        ss = warp_sum(ss);
        if (lane != 0) continue;
        const float norm = __fmul_ru(__fsqrt_ru(ss), 1.00001f);  // rounded up: the sum's own rounding error is << 1e-5
        if (tile_max) {
            if (live) atomicMax(tile_max + rb, __float_as_uint(norm));
        } else if (!live) {
            thr[r] = make_float4(INFINITY, INFINITY, 0.f, 0.f);  // padding queries never count, never push
        } else {
            const int n = qpos[r];
            const double edge_hi = (n + 1 >= 1) ? (double)(n + 1) : (double)n;  // T(n+1)
            const double edge_lo = (n >= 1) ? (double)n : (double)n - 1.0;      // T(n)
            const float U = (float)(edge_hi / 1000.0 / (double)scale), L = (float)(edge_lo / 1000.0 / (double)scale);
            // slack: rounding of U, L themselves, of fl(scale*acc) and of fl(score*1000), each <= 2^-24 relative (64x margin)
            const float slack = fmaxf(fabsf(U), fabsf(L)) * 3.814697265625e-6f + 1e-37f;
            thr[r] = make_float4(U, L, __fmul_ru(eps_rel, norm), slack);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers: tcgen05 (MMA, TMEM alloc / load / fences, commit)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM lane quarter
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, mma_sm100_desc.hpp): K-major, SWIZZLE_NONE,
// leading byte offset (next 8-column chunk along K) 128 B, stride byte offset (next 8-row group) TC_KC*128 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)((TC_KC * 128u) >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, N = 256, M = 128, dense
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_NC >> 3) << 17) | ((uint32_t)(TC_MQ >> 4) << 24);

struct TcParams {
    const __nv_bfloat16 *a_split;  // [n_qb][nkb][hi|lo][128 x 64] query tiles
    const __nv_bfloat16 *b_split;  // [n_ct][nkb][hi|lo][256 x 64] candidate tiles
    const float4 *thr;             // [n_qb*128] {U, L, eps_rel*|q|, slack}
    const float *tile_norm;        // [n_ct] max |e| of the tile
    int32_t *cnt;                  // [b,3]
    int2 *pairs;                   // undecided (query, candidate position) pairs
    unsigned *pair_count;
    unsigned pair_cap;
    long long b, n_cand;
    int nkb, ksteps, n_qb, n_ct, ctas_per_qb;
    float *probe_a, *probe_d;      // diagnostic outputs (PROBE instantiation only)
    float scale;
};

// ---------------------------------------------------------------------------------------------------------------------
// 2. the filter kernel.  CTA (qb, j): query block qb, candidate tiles j, j + ctas_per_qb, ...; one CTA per SM.
//    Per tile: nkb trips of a TC_STAGES-deep ring {cp.async.bulk A tile + B tile -> 3 MMAs per k-step}, then 128x256 accumulators are
//    read back by four warps (one query per thread) while the next tile's MMAs fill the other accumulator stage.
// ---------------------------------------------------------------------------------------------------------------------
template <bool PROBE>
__global__ void __launch_bounds__(TC_THREADS, 1) kge_rank_tc_kernel(const TcParams p)
{
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);  // [TC_STAGES] operands landed
    uint64_t *empty = full + TC_STAGES;                                        // [TC_STAGES] operands consumed
    uint64_t *tfull = empty + TC_STAGES;                                       // [2] accumulator stage complete
    uint64_t *tempty = tfull + 2;                                              // [2] accumulator stage drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x % p.n_qb, first = blockIdx.x / p.n_qb;
    const int n_items = first < p.n_ct ? (p.n_ct - first + p.ctas_per_qb - 1) / p.ctas_per_qb : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- producer: one thread feeds the two-stage operand ring with bulk copies ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int i = 0; i < n_items; ++i) {
                const int ct = first + i * p.ctas_per_qb;
                for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(empty + s, ((it / TC_STAGES) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(full + s, (uint32_t)TC_STAGE_BYTES);
                    unsigned char *st = smem + (size_t)s * TC_STAGE_BYTES;
                    bulk_load(st, p.a_split + ((size_t)qb * p.nkb + kb) * (size_t)(2 * TC_A_ELEMS), TC_A_BYTES, full + s);
                    bulk_load(st + TC_A_BYTES, p.b_split + ((size_t)ct * p.nkb + kb) * (size_t)(2 * TC_B_ELEMS), TC_B_BYTES, full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: lane 0 issues every tcgen05.mma of the CTA ----------------
        uint32_t it = 0;
        for (int i = 0; i < n_items; ++i) {
            const int acc = i & 1;
            mbar_wait(tempty + acc, ((i >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_NC);
            for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                const int s = it % TC_STAGES;
                mbar_wait(full + s, (it / TC_STAGES) & 1u);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * TC_STAGE_BYTES), a_lo = a_hi + TC_A_ELEMS * 2;
                    const uint32_t b_hi = a_hi + TC_A_BYTES, b_lo = b_hi + TC_B_ELEMS * 2;
                    const int ks_n = min(TC_KSTEPS, p.ksteps - kb * TC_KSTEPS);
                    for (int ks = 0; ks < ks_n; ++ks) {
                        const uint32_t o = (uint32_t)ks * 256u;  // two 128-byte k-chunks per k-step of 16
                        const uint64_t da_hi = umma_smem_desc(a_hi + o), da_lo = umma_smem_desc(a_lo + o);
                        const uint64_t db_hi = umma_smem_desc(b_hi + o), db_lo = umma_smem_desc(b_lo + o);
                        umma_bf16(d_tmem, da_hi, db_hi, TC_IDESC, (kb | ks) != 0);
                        umma_bf16(d_tmem, da_hi, db_lo, TC_IDESC, 1u);
                        umma_bf16(d_tmem, da_lo, db_hi, TC_IDESC, 1u);
                    }
                    umma_commit(empty + s);                       // operands of this stage may be overwritten
                    if (kb == p.nkb - 1) umma_commit(tfull + acc);  // accumulators of this tile are final
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: thread = one query (TMEM lane), loop over the tile's 256 candidate columns ----------------
        const int lg = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
        const int ql = lg * 32 + lane;
        const long long q = (long long)qb * TC_MQ + ql;
        const float4 thr = p.thr[q];
        const float U = thr.x, L = thr.y;
        int gt = 0, ge = 0;
        for (int i = 0; i < n_items; ++i) {
            const int acc = i & 1;
            const int ct = first + i * p.ctas_per_qb;
            const long long c0 = (long long)ct * TC_NC;
            const int nvalid = (int)min((long long)TC_NC, p.n_cand - c0);
            const float delta = __fmaf_ru(__fmul_ru(thr.z, p.tile_norm[ct]), 1.0001f, thr.w);
            mbar_wait(tfull + acc, (i >> 1) & 1u);
            tc_fence_after();
            for (int ch = 0; ch < TC_NC / 32; ++ch) {
                if (ch * 32 >= nvalid) break;  // warp-uniform
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * TC_NC + ch * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float a = v[j];
                    const bool ok = ch * 32 + j < nvalid;
                    if (PROBE) {
                        if (ok && q < p.b) {
                            p.probe_a[(size_t)q * p.n_cand + c0 + ch * 32 + j] = a * p.scale;
                            p.probe_d[(size_t)q * p.n_cand + c0 + ch * 32 + j] = delta * p.scale;
                        }
                        continue;
                    }
                    const bool undecided = fminf(fabsf(a - U), fabsf(a - L)) < delta;
                    gt += (ok && !undecided && a >= U) ? 1 : 0;
                    ge += (ok && !undecided && a >= L) ? 1 : 0;
                    if (ok && undecided) {
                        const unsigned slot = atomicAdd(p.pair_count, 1u);
                        if (slot < p.pair_cap) p.pairs[slot] = make_int2((int)q, (int)(c0 + ch * 32 + j));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
        }
        if (q < p.b) {
            if (gt) atomicAdd(p.cnt + 3 * q + 0, gt);
            if (ge - gt) atomicAdd(p.cnt + 3 * q + 1, ge - gt);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// 3. refine: the exact canonical chain (same arithmetic as kge_rank_dot_kernel / the filter kernel of kge_rank.cu) for the
//    undecided pairs; one thread per pair.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void kge_rank_refine_kernel(const RankParams p, const int2 *__restrict__ pairs, const unsigned *__restrict__ pair_count,
                                       unsigned pair_cap, int32_t *__restrict__ cnt)
{
    extern __shared__ __align__(16) float pair_sm[];
    const unsigned n = *pair_count;
    if (n > pair_cap) {  // overflow: forget what the filter counted; the gated exact kernel (next launch) recounts everything
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.b; i += (long long)gridDim.x * blockDim.x) {
            cnt[3 * i] = 0;
            cnt[3 * i + 1] = 0;
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    float *sm = pair_sm + (size_t)warp * 2 * p.L.ld;
    for (unsigned f = blockIdx.x * wpc + warp; f < n; f += gridDim.x * wpc) {  // one warp per undecided pair
        const int2 pr = pairs[f];
        const long long pos = pr.y;
        const long long id = p.cand_ids ? (long long)p.cand_ids[pos] : p.cand_begin + pos;
        const float sc = pair_score_warp<OP_DOT>(p.ent + (size_t)id * p.L.ld, p.qvec + (size_t)pr.x * p.L.ld, nullptr, p.L.ld, p.L.kp,
                                                 p.scale, sm, lane);
        if (lane == 0) {
            const int qc = quantise(sc), qp = p.qpos[pr.x];
            if (qp < qc) atomicAdd(cnt + 3 * pr.x + 0, 1);
            else if (qp == qc) atomicAdd(cnt + 3 * pr.x + 1, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static size_t tc_align(size_t x) { return (x + 1023) & ~(size_t)1023; }

bool rank_tc_applicable(const Layout &L, int side, long long b, long long n_cand)
{
    (void)side;
    if (!(L.model == KGE_DISTMULT || L.model == KGE_COMPLEX || L.model == KGE_HOLE)) return false;
    // below this the FP32 kernel is already launch-bound; above, int positions / the per-SM query-block mapping must hold
    return b >= 1 && n_cand >= 512 && n_cand < (1ll << 31) && (b + TC_MQ - 1) / TC_MQ <= 128;
}

RankTcLayout rank_tc_layout(const Layout &L, long long b, long long n_cand, int pair_cap_override)
{
    RankTcLayout w;
    w.nkb = (L.ld + TC_KB - 1) / TC_KB;
    w.ksteps = (L.ld + 15) / 16;
    w.n_qb = (int)((b + TC_MQ - 1) / TC_MQ);
    w.n_ct = (int)((n_cand + TC_NC - 1) / TC_NC);
    long long cap = b * n_cand / 16;
    if (cap < (1ll << 20)) cap = 1ll << 20;
    if (cap > (1ll << 25)) cap = 1ll << 25;
    if (pair_cap_override > 0) cap = pair_cap_override;
    w.pair_cap = (unsigned)cap;
    size_t off = 0;
    w.off_a = off; off += tc_align((size_t)w.n_qb * w.nkb * TC_A_BYTES);
    w.off_b = off; off += tc_align((size_t)w.n_ct * w.nkb * TC_B_BYTES);
    w.off_thr = off; off += tc_align((size_t)w.n_qb * TC_MQ * sizeof(float4));
    w.off_count = off; off += 1024;                                       // pair counter, then the tile norms: one memset
    w.off_tnorm = off; off += tc_align((size_t)w.n_ct * sizeof(float));
    w.off_pairs = off; off += tc_align((size_t)w.pair_cap * sizeof(int2));
    w.bytes = off;
    return w;
}

cudaError_t launch_rank_count_tc(const RankParams &p, const RankTcLayout &w, void *ws, int32_t *cnt, int sm_count, cudaStream_t st,
                                 float *probe_a, float *probe_d)
{
    char *base = (char *)ws;
    __nv_bfloat16 *a_split = (__nv_bfloat16 *)(base + w.off_a), *b_split = (__nv_bfloat16 *)(base + w.off_b);
    float4 *thr = (float4 *)(base + w.off_thr);
    float *tnorm = (float *)(base + w.off_tnorm);
    unsigned *count = (unsigned *)(base + w.off_count);
    int2 *pairs = (int2 *)(base + w.off_pairs);
    const int ld = p.L.ld;
    cudaError_t e;
    if ((e = cudaMemsetAsync(count, 0, 1024 + (size_t)w.n_ct * sizeof(float), st)) != cudaSuccess) return e;
    // split (+ norms, + per-query thresholds): candidates in 256-row tiles, this side's query vectors in 128-row tiles
    const float eps_rel = (float)(ldexp(1.0, -14) + 4.0 * (double)ld * ldexp(1.0, -23));
    {
        const long long rows_pad = (long long)w.n_ct * TC_NC, want = (rows_pad + 7) / 8;
        kge_rank_split_kernel<TC_NC><<<(unsigned)(want < (long long)sm_count * 32 ? want : (long long)sm_count * 32), 256, 0, st>>>(
            p.ent, p.cand_ids, p.cand_begin, p.n_cand, rows_pad, ld, w.nkb, b_split, (unsigned *)tnorm, nullptr, 1.f, 0.f, nullptr);
    }
    {
        const long long rows_pad = (long long)w.n_qb * TC_MQ, want = (rows_pad + 7) / 8;
        kge_rank_split_kernel<TC_MQ><<<(unsigned)(want < (long long)sm_count * 32 ? want : (long long)sm_count * 32), 256, 0, st>>>(
            p.qvec, nullptr, 0, p.b, rows_pad, ld, w.nkb, a_split, nullptr, p.qpos, p.scale, eps_rel, thr);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    TcParams t;
    t.a_split = a_split; t.b_split = b_split; t.thr = thr; t.tile_norm = tnorm; t.cnt = cnt;
    t.pairs = pairs; t.pair_count = count; t.pair_cap = w.pair_cap;
    t.b = p.b; t.n_cand = p.n_cand; t.nkb = w.nkb; t.ksteps = w.ksteps; t.n_qb = w.n_qb; t.n_ct = w.n_ct;
    int per = sm_count / w.n_qb;
    if (per < 1) per = 1;
    if (per > w.n_ct) per = w.n_ct;
    t.ctas_per_qb = per;
    t.probe_a = probe_a; t.probe_d = probe_d; t.scale = p.scale;
    const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment*/ + 128 /*barriers + TMEM slot*/;
    if (probe_a) {
        if ((e = cudaFuncSetAttribute(kge_rank_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kge_rank_tc_kernel<true><<<(unsigned)(w.n_qb * per), TC_THREADS, smem, st>>>(t);
        return cudaGetLastError();
    }
    if ((e = cudaFuncSetAttribute(kge_rank_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    kge_rank_tc_kernel<false><<<(unsigned)(w.n_qb * per), TC_THREADS, smem, st>>>(t);
    {
        const int wpc = pair_score_warps(ld, 2);
        const size_t rsm = (size_t)wpc * 2 * ld * sizeof(float);
        if ((e = cudaFuncSetAttribute(kge_rank_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm)) != cudaSuccess) return e;
        kge_rank_refine_kernel<<<sm_count * 2, wpc * 32, rsm, st>>>(p, pairs, count, w.pair_cap, cnt);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // overflow fallback: the exact FP32 kernel, every CTA of which returns at once unless the pair list overflowed
    RankParams g = p;
    g.gate = count;
    g.gate_cap = w.pair_cap;
    return launch_rank_count(g, cnt, st);
}

}  // namespace kge
