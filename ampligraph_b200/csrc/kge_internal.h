// kge_internal.h -- internal launcher interfaces of libkge_b200 (not part of the C-ABI).
#pragma once
#include "kge_common.cuh"

// per-lane register state grows with NIT (float4 chunks per lane): trade threads for registers
// (RotatE carries 11 float4 vectors of per-positive state per chunk -- cos, sin, o, y = R s, four Z sums, d/dphi: give it more registers)
// (RotatE with 12 warps at 168 registers was measured against 8 warps at 229: 273 us vs 267 us on cfg4, profiles/r2m_kbench_cfg4_sweep.log)
#define KGE_TRAIN_THREADS(model, nit) ((nit) <= 1 ? 512 : (nit) == 2 ? ((model) == KGE_ROTATE ? 256 : 384) : 256)
#define KGE_MAX_PEERS 8
#define KGE_MIN_RESIDENT_WARPS 8  // tuned on B200: fewer warps/SM than this costs more than re-gathering
#define KGE_TARGET_WARPS 10
#define KGE_MIN_RESIDENT_WARPS_WIDE 6  // the fast path on DistMult rows of 257..512 floats (NIT = 4): cfg3 113.6 us with 6 resident warps against 133 us for the grouped general kernel with 8 (profiles/r2l_*)

namespace kge {

struct TrainParams {
    const float *ent;   // [n_ent, ld]
    const float *rel;   // [n_rel, ld]; for RotatE the per-step rotation table [cos|sin]
    float *grad_ent;
    float *grad_rel;
    const int32_t *triples;  // [B,3]
    long long B;
    const int32_t *neg_ent;  // [eta*B] or nullptr (Philox)
    const uint8_t *neg_keep; // [eta*B] or nullptr
    unsigned long long seed, step;
    unsigned n_ent;
    int model, eta, kp, ld, nch;  // nch = kp/4 float4 chunks per half
    int G;                        // replaced rows resident per pass
    int nbuf;                     // group buffers per slot (non-resident): 2 = next group prefetched, 1 = fetched on demand
    int wk, n_cb;                 // column window (floats per half) and number of windows per row
    int slot_floats;              // stride between row windows in the shared-memory slot
    int resident;                 // 1: one window, one group, rows stay in place between the passes
    int eta_pad;                  // round_up(eta,4)
    int rows_bytes, region_bytes; // per-warp shared-memory carve-up
    int loss, reduction, mode;
    float margin, alpha, score_scale, inv_div;
    double *loss_out;
    float *scores_pos, *scores_neg;
    const float *dpos, *dneg;
    // row-sharded entity table (shard_world <= 1: single table at ent / grad_ent)
    int shard_world, rows_per_shard;
    const float *ent_shard[KGE_MAX_PEERS];
    float *grad_ent_shard[KGE_MAX_PEERS];
    // lazy optimizer support: rows touched by this step get stamp written (nullptr: off)
    int stamp;
    int *stamp_ent, *stamp_rel;
    int *stamp_ent_shard[KGE_MAX_PEERS];
    int hot_ent[2];  // entities whose subject/object gradient rows are summed per warp before they are scattered (-1: none); fast path only
    float *stash;  // [B, eta, ld] local copy of the replaced rows gathered by the score pass (sharded runs) or nullptr
    unsigned *sched;  // {next positive, retired warps}: dynamic assignment of positives to warps (see next_positive), or nullptr = static stride
};

// grid = min(occupancy * sm_count, ceil(B / warps))
cudaError_t launch_train(const TrainParams &p, int nit, int sm_count, int threads, size_t smem, cudaStream_t st);
// kge_train_res.cu: the resident trilinear fast path (DistMult / ComplEx / HolE, all rows of a positive resident, eta <= 32, one
// table); p.rows_bytes / p.region_bytes carry ITS slot geometry (kge_create: res_*), threads = res_warps * 32
cudaError_t launch_train_res(const TrainParams &p, int nit, int sm_count, int threads, size_t smem, cudaStream_t st);
// kge_train_rot.cu: the RotatE fast path (one column window, single table, any eta in groups of p.G with two buffers -- or one
// group when everything fits); p.G / p.rows_bytes / p.region_bytes carry ITS slot geometry (kge_create: rot_*), threads = rot_warps * 32
cudaError_t launch_train_rot(const TrainParams &p, int nit, int sm_count, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_rotation_table(const float *rel, float *rot, long long n_rel, int kp, int ld, float div,
                                  cudaStream_t st);
cudaError_t launch_corruptions(const int32_t *triples, long long B, int eta, unsigned long long seed,
                               unsigned long long step, unsigned n_ent, int32_t *out, cudaStream_t st);

// kge_misc.cu
cudaError_t launch_score_triples(const Layout &L, int nit, const float *ent, const float *rel_or_rot,
                                 const int32_t *triples, long long n, float scale, float *out, int sm_count,
                                 cudaStream_t st);
cudaError_t launch_pack(const Layout &L, const float *src, float *dst, long long rows, bool unpack, cudaStream_t st);
cudaError_t launch_glorot(const Layout &L, float *table, long long rows, unsigned long long seed, cudaStream_t st);
cudaError_t launch_init_table(const Layout &L, float *table, long long rows, int kind, float a, float b,
                              unsigned long long seed, cudaStream_t st);

// kge_optim.cu
struct RegParams {  // lambda*sum|x|^p (+ optional second term: Keras 'l1_l2')
    int p, p2;
    float lambda, lambda2;
};
struct OptimParams {
    int kind;
    float lr, lr_t, beta1, beta2, eps, momentum;
    RegParams reg;
};
struct ExchangeParams {  // kge_optimizer_step_exchange
    int world, rank, phases;
    float *table[KGE_MAX_PEERS];        // [ent|rel] parameter block of every rank
    const float *grad[KGE_MAX_PEERS];   // this step's gradient block of every rank
    unsigned *flags[KGE_MAX_PEERS];     // flag pad of every rank: 2*world uint32
    float *table_mc;                    // NVLS: multicast mapping of the parameter block / this step's gradient block
    const float *grad_mc;               // (both non-null: reduce and broadcast inside the switch), else nullptr
    unsigned token;
    float *zero_grad;                   // local: next step's gradient block, zeroed here (or nullptr)
    long long total4;                   // float4s in one block
    long long off4, n4;                 // this rank's shard
    long long ent4;                     // float4s of the entity part (regulariser switch point)
    RegParams reg_ent, reg_rel;
    float *slot0, *slot1;
    double *reg_loss;
    unsigned *done_counter;             // handle-owned device counter (last-CTA detection), self-resetting
    unsigned long long *trace;          // nullptr or 8 x uint64 phase stamps (kge_set_exchange_trace)
};
cudaError_t launch_optimizer(const OptimParams &o, float *table, float *grad, float *slot0, float *slot1,
                             long long n_floats, double *reg_loss, int sm_count, cudaStream_t st);
cudaError_t launch_optimizer_lazy(const OptimParams &o, float *table, float *grad, float *slot0, float *slot1,
                                  long long rows, int ld, const int *row_stamp, int stamp, double *reg_loss,
                                  int sm_count, cudaStream_t st);
// tables/grads: HOST arrays of `world` device pointers (own rank first is NOT required; index = rank)
cudaError_t launch_optimizer_sharded(const OptimParams &o, int world, int rank, float *const *tables, float *const *grads,
                                     float *slot0, float *slot1, long long off_floats, long long n_floats,
                                     double *reg_loss, int sm_count, cudaStream_t st);
cudaError_t launch_optimizer_exchange(const OptimParams &o, const ExchangeParams &x, int sm_count, cudaStream_t st);
cudaError_t launch_peer_barrier(int world, int rank, unsigned *const *flags, int slot, unsigned token, cudaStream_t st);

// kge_rank.cu
struct ShardView {  // row-sharded entity table seen through peer pointers (world <= 1: not sharded)
    int world, rows_per_shard;
    const float *ent[KGE_MAX_PEERS];
};
struct RankParams {
    Layout L;
    int side, strategy;
    const float *ent;      // [n_ent, ld]
    const float *qvec;     // [b, ld] prepared query vectors for this side
    const float *qaux;     // [b, ld] object rows (RotatE subject side) or nullptr
    const int32_t *qpos;   // [b] quantised positive scores
    const int32_t *cand_ids;  // nullptr or [n_cand]
    long long cand_begin, n_cand, b;
    long long filt_base;   // subtracted from filter ids (global id of the local shard's first row)
    float scale;           // HolE
    float *scores;         // nullptr, or [b, n_cand] output of every candidate's score (kge_corruption_scores)
    const unsigned *gate;  // nullptr, or device counter: kge_rank_dot_kernel only runs when *gate > gate_cap (overflow fallback
    unsigned gate_cap;     // of the tensor-core filter, kge_rank_tc.cu)
};
cudaError_t launch_rank_prepare(const Layout &L, const ShardView &sv, const float *ent, const float *rel, const float *rot,
                                const int32_t *triples, long long b, float scale, float *qvec_s, float *qvec_o,
                                float *qaux, int32_t *qpos, cudaStream_t st);
// cnt = [b,3] int32 workspace: #(qpos < qc), #(qpos == qc), #filtered
cudaError_t launch_rank_count(const RankParams &p, int32_t *cnt, cudaStream_t st);
cudaError_t launch_rank_filter_n(const RankParams &p, const long long *filt_off, const int32_t *filt_idx,
                                 long long n_pairs, int32_t *cnt, cudaStream_t st);
cudaError_t launch_rank_finalize(const int32_t *cnt, long long b, int strategy, int32_t *ranks, cudaStream_t st);
cudaError_t launch_rank_accumulate(const int32_t *cnt, long long b, int32_t *counts, cudaStream_t st);  // counts += cnt

// kge_rank_tc.cu: tensor-core filter + exact refine for the bilinear models
struct RankTcLayout {  // carve-up of the tensor-core part of the caller's workspace
    int nkb, ksteps, n_qb, n_ct;
    unsigned pair_cap;
    size_t off_a, off_b, off_thr, off_tnorm, off_count, off_pairs, bytes;
};
bool rank_tc_applicable(const Layout &L, int side, long long b, long long n_cand);
RankTcLayout rank_tc_layout(const Layout &L, long long b, long long n_cand, int pair_cap_override);
// same contract as launch_rank_count (cnt[3q+0] += #greater, cnt[3q+1] += #equal), bit-identical results.
// probe_a/probe_d != nullptr: diagnostic run -- write approximate scores and assumed error bounds [b, n_cand], count nothing
cudaError_t launch_rank_count_tc(const RankParams &p, const RankTcLayout &w, void *ws, int32_t *cnt, int sm_count, cudaStream_t st,
                                 float *probe_a = nullptr, float *probe_d = nullptr);

}  // namespace kge
