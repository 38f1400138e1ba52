// kge_api.cu -- the extern "C" boundary of libkge_b200.so (see include/kge_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "kge_internal.h"

using namespace kge;

struct kge_handle {
    kge_config cfg;
    Layout L;
    int sm_count;
    int max_smem;
    float score_scale;  // HolE 2/k
    float rot_div;      // RotatE range/pi
    // the only device memory the library owns (allocated once by kge_create, freed by kge_destroy):
    float *rot;              // [n_rel, ld] rotation table (RotatE only)
    unsigned *done_counter;  // [0]: last-CTA counter of kge_optimizer_step_exchange; [1], [2]: the training kernels' positive counter and retired-warp counter (self-resetting)
    unsigned long long *exchange_trace;  // caller-owned 8 x uint64 phase stamps of the next exchange launches, or nullptr
    // training launch geometry
    int nit, G, nbuf, warps, eta_pad, rows_bytes, region_bytes, wk, n_cb, slot_floats, resident;
    int res_warps, res_rows_bytes, res_region_bytes;  // slot geometry of the resident trilinear fast path (kge_train_res.cu); res_warps = 0: not applicable
    int rot_warps, rot_G, rot_rows_bytes, rot_region_bytes;  // slot geometry of the RotatE fast path (kge_train_rot.cu); rot_warps = 0: not applicable
    int *stamp_ent, *stamp_rel;  // lazy-optimizer row stamps (caller-owned) or nullptr
    float *stash;                // caller-owned row stash for sharded runs (kge_set_row_stash) or nullptr
    long long stash_rows;
    int hot_ent[2];              // kge_set_hot_entities (-1: none)
    int l2_bytes;                // cudaDeviceProp::l2CacheSize
};

// Make the handle's device current for the duration of an entry point (two handles on different GPUs
// driven from one thread); restores the caller's device on return.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};
#define KGE_GUARD(h)                       \
    DeviceGuard guard_((h)->cfg.device);   \
    if (guard_.err != cudaSuccess) return cuda_fail(guard_.err, "cudaSetDevice")

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char *what)
{
    return fail(KGE_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define KGE_CUDA(call, what)                                 \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, what);   \
    } while (0)

extern "C" const char *kge_last_error(void) { return g_err; }
extern "C" int kge_abi_version(void) { return KGE_B200_ABI_VERSION; }

extern "C" int kge_create(const kge_config *cfg, kge_handle **out)
{
    if (!cfg || !out) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: null argument");
    if (cfg->struct_size != (int32_t)sizeof(kge_config))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: kge_config.struct_size %d != %d (ABI mismatch)",
                    cfg->struct_size, (int)sizeof(kge_config));
    if (cfg->scoring < KGE_TRANSE || cfg->scoring > KGE_ROTATE)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: unknown scoring_type id %d", cfg->scoring);
    if (cfg->loss < KGE_LOSS_PAIRWISE || cfg->loss > KGE_LOSS_MULTICLASS_NLL)
        return fail(KGE_ERR_INVALID_ARGUMENT, "Could not interpret loss identifier: %d", cfg->loss);
    if (cfg->reduction != KGE_REDUCE_SUM && cfg->reduction != KGE_REDUCE_MEAN)
        return fail(KGE_ERR_INVALID_ARGUMENT, "Invalid value for reduction!");
    if (cfg->rank_mode != KGE_RANK_MODE_AUTO && cfg->rank_mode != KGE_RANK_MODE_EXACT)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: unknown rank_mode %d", cfg->rank_mode);
    if (cfg->max_rel_size < 0 || cfg->rank_pair_cap < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: max_rel_size / rank_pair_cap < 0");
    if (cfg->k < 1 || cfg->eta < 1 || cfg->n_ent < 1 || cfg->n_rel < 1)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: k, eta, n_ent, n_rel must be >= 1");
    if (cfg->n_ent > 0x7fffffffLL || cfg->n_rel > 0x7fffffffLL)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: ids are int32 (n_ent, n_rel < 2^31)");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(KGE_ERR_CUDA, "kge_create: no CUDA device visible (%s) -- libkge_b200 has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_create: device %d out of range [0,%d)", cfg->device, ndev);
    DeviceGuard guard_(cfg->device);
    if (guard_.err != cudaSuccess) return cuda_fail(guard_.err, "cudaSetDevice");
    cudaDeviceProp prop;
    KGE_CUDA(cudaGetDeviceProperties(&prop, cfg->device), "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(KGE_ERR_UNSUPPORTED, "kge_create: device is sm_%d%d; libkge_b200 is built for sm_100a only",
                    prop.major, prop.minor);

    kge_handle *h = new (std::nothrow) kge_handle();
    if (!h) return fail(KGE_ERR_CUDA, "kge_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    Layout &L = h->L;
    L.model = cfg->scoring;
    L.k = cfg->k;
    L.halves = model_halves(cfg->scoring);
    L.kp = (cfg->k + 3) / 4 * 4;
    L.ld = L.halves * L.kp;
    L.K = L.halves * cfg->k;
    h->sm_count = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->l2_bytes = prop.l2CacheSize;
    h->score_scale = (cfg->scoring == KGE_HOLE) ? hole_scale(L.K) : 1.f;
    h->rot_div = (cfg->scoring == KGE_ROTATE) ? rotate_divisor(L.K, cfg->max_rel_size > 0 ? cfg->max_rel_size : cfg->n_rel) : 1.f;

    // ---- training geometry: one warp per positive; its slot holds (3+G) row windows ----
    // window = wk floats per half (<= 128*NIT so that NIT float4 per lane cover it); rows wider than
    // 512 floats per half are processed in n_cb column windows, eta > G negatives in groups of G.
    const int nch = L.kp / 4;
    const int nit_raw = (nch + 31) / 32;
    h->nit = nit_raw <= 1 ? 1 : nit_raw <= 2 ? 2 : 4;
    h->wk = L.kp <= 512 ? L.kp : 512;
    h->n_cb = (L.kp + h->wk - 1) / h->wk;
    h->eta_pad = (cfg->eta + 3) / 4 * 4;
    const int row_bytes = L.halves * h->wk * 4, aux = 4 * h->eta_pad * 4 + 16;  // sc | nid | nside | jorig, two mbarriers
    h->slot_floats = row_bytes / 4;
    int max_warps = KGE_TRAIN_THREADS(cfg->scoring, h->nit) / 32;
    int nbuf = 2;  // group buffers of a non-resident slot
    // tuning aids (scripts/kbench.py): cap the warps per CTA, choose the buffering of non-resident slots
    if (const char *ev = getenv("KGE_B200_TRAIN_WARPS")) { const int w = atoi(ev); if (w >= 1 && w < max_warps) max_warps = w; }
    if (const char *ev = getenv("KGE_B200_TRAIN_NBUF")) nbuf = atoi(ev) == 1 ? 1 : 2;
    // Residency policy, tuned on B200 (scripts/kbench.py sweep, profiles/): keep all eta corruptions resident
    // (one gather per row) while >= KGE_MIN_RESIDENT_WARPS warps still fit; otherwise shrink G until
    // min(max_warps, KGE_TARGET_WARPS) warps fit -- the gradient pass then re-gathers the other groups, which
    // costs less than running with few warps.
    // non-resident slots hold TWO group buffers (the next group is prefetched while the current one is processed)
    int G = cfg->eta;
    bool res = h->n_cb == 1;
    if (cfg->neg_group > 0) {
        G = cfg->neg_group < cfg->eta ? cfg->neg_group : cfg->eta;
        res = res && G >= cfg->eta;
    } else {
        const long long full = (3 + (long long)cfg->eta) * row_bytes + aux;
        const int min_res = max_warps < KGE_MIN_RESIDENT_WARPS ? max_warps : KGE_MIN_RESIDENT_WARPS;
        if (!res || full * min_res > h->max_smem) {
            res = false;
            const int want = max_warps < KGE_TARGET_WARPS ? max_warps : KGE_TARGET_WARPS;
            while (G > 1 && (long long)((3 + nbuf * G) * (long long)row_bytes + aux) * want > h->max_smem) --G;
        }
    }
    h->resident = res ? 1 : 0;
    h->G = G;
    h->nbuf = nbuf;
    h->rows_bytes = (res ? 3 + G : 3 + nbuf * G) * row_bytes;
    h->region_bytes = h->rows_bytes + aux;
    int warps = h->max_smem / h->region_bytes;
    if (warps > max_warps) warps = max_warps;
    h->warps = warps;  // 0 => even (3+1) windows do not fit (reported by kge_train_step)
    h->hot_ent[0] = h->hot_ent[1] = -1;

    // ---- the resident trilinear fast path (kge_train_res.cu): DistMult / ComplEx / HolE, one window, eta <= 32; its slot
    // holds s, o and the eta replaced rows (the relation row lives in registers) + sc | nid | jorig + one mbarrier.
    // Same residency rule as above, with its own (one row smaller) slot.
    h->res_warps = 0;
    // DistMult rows of up to 512 floats (NIT = 4) qualify too: its per-lane state is a third of ComplEx's.
    if ((cfg->scoring == KGE_DISTMULT || cfg->scoring == KGE_COMPLEX || cfg->scoring == KGE_HOLE) && h->n_cb == 1 &&
        (h->nit <= 2 || cfg->scoring == KGE_DISTMULT) && cfg->eta <= 32 && cfg->neg_group <= 0) {
        const int res_aux = (3 * h->eta_pad * 4 + 8 + 15) / 16 * 16;
        h->res_rows_bytes = (2 + cfg->eta) * row_bytes;
        h->res_region_bytes = h->res_rows_bytes + res_aux;
        int rw = h->max_smem / h->res_region_bytes;
        if (rw > max_warps) rw = max_warps;
        int min_res = max_warps < KGE_MIN_RESIDENT_WARPS ? max_warps : KGE_MIN_RESIDENT_WARPS;
        if (h->nit == 4) min_res = KGE_MIN_RESIDENT_WARPS_WIDE;
        if (const char *ev = getenv("KGE_B200_RES_MIN_WARPS")) min_res = atoi(ev);  // tuning aid
        if (rw >= min_res) h->res_warps = rw;
    }

    // ---- the RotatE fast path (kge_train_rot.cu): one window, rows of up to 256 floats per half; its slot holds ONLY replaced
    // rows (s, o and the rotation row live in registers): all eta of them when they fit beside 8 warps, else two buffers of G.
    h->rot_warps = 0;
    if (cfg->scoring == KGE_ROTATE && h->n_cb == 1 && h->nit <= 2 && cfg->neg_group <= 0) {
        const int rot_aux = (4 * h->eta_pad * 4 + 16 + 15) / 16 * 16;  // sc | nid | jorig | side, two mbarriers
        const int rw = 8;                                                // 256 threads (launch bound of the kernel)
        const int cap = (h->max_smem / rw - rot_aux) / row_bytes;        // rows per warp
        int rG = cap >= cfg->eta ? cfg->eta : cap / 2;
        if (const char *ev = getenv("KGE_B200_ROT_G")) {  // tuning aid: smaller groups
            const int g = atoi(ev);
            if (g >= 1 && g <= cap / 2 && g < cfg->eta) rG = g;
        }
        if (rG >= 1) {
            h->rot_warps = rw;
            h->rot_G = rG;
            h->rot_rows_bytes = (rG >= cfg->eta ? rG : 2 * rG) * row_bytes;
            h->rot_region_bytes = h->rot_rows_bytes + rot_aux;
        }
    }

    if (cfg->scoring == KGE_ROTATE) {
        cudaError_t e2 = cudaMalloc(&h->rot, (size_t)cfg->n_rel * L.ld * sizeof(float));
        if (e2 != cudaSuccess) { delete h; return cuda_fail(e2, "cudaMalloc(rotation table)"); }
    }
    {
        cudaError_t e2 = cudaMalloc(&h->done_counter, 4 * sizeof(unsigned));
        if (e2 == cudaSuccess) e2 = cudaMemset(h->done_counter, 0, 4 * sizeof(unsigned));
        if (e2 != cudaSuccess) { if (h->rot) cudaFree(h->rot); delete h; return cuda_fail(e2, "cudaMalloc(counter)"); }
    }
    *out = h;
    return KGE_OK;
}

extern "C" void kge_destroy(kge_handle *h)
{
    if (!h) return;
    DeviceGuard guard_(h->cfg.device);
    if (h->rot) cudaFree(h->rot);
    if (h->done_counter) cudaFree(h->done_counter);
    delete h;
}

extern "C" int32_t kge_internal_k(const kge_handle *h) { return h ? h->L.K : 0; }
extern "C" int32_t kge_half_stride(const kge_handle *h) { return h ? h->L.kp : 0; }
extern "C" int32_t kge_row_stride(const kge_handle *h) { return h ? h->L.ld : 0; }

// null check + device guard (every entry point below launches on the handle's device)
#define KGE_CHECK_HANDLE(h, fn)                                           \
    if (!(h)) return fail(KGE_ERR_INVALID_ARGUMENT, fn ": null handle"); \
    KGE_GUARD(h)

extern "C" int kge_pack_rows(kge_handle *h, const float *dense_dev, float *table_dev, int64_t rows, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_pack_rows");
    if (rows < 0 || (rows > 0 && (!dense_dev || !table_dev))) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_pack_rows: bad argument");
    KGE_CUDA(launch_pack(h->L, dense_dev, table_dev, rows, false, (cudaStream_t)stream), "kge_pack_rows");
    return KGE_OK;
}
extern "C" int kge_unpack_rows(kge_handle *h, const float *table_dev, float *dense_dev, int64_t rows, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_unpack_rows");
    if (rows < 0 || (rows > 0 && (!dense_dev || !table_dev))) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_unpack_rows: bad argument");
    KGE_CUDA(launch_pack(h->L, table_dev, dense_dev, rows, true, (cudaStream_t)stream), "kge_unpack_rows");
    return KGE_OK;
}

extern "C" int kge_init_glorot_uniform(kge_handle *h, float *table_dev, int64_t rows, uint64_t seed, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_init_glorot_uniform");
    if (rows < 0 || (rows > 0 && !table_dev)) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_init_glorot_uniform: bad argument");
    KGE_CUDA(launch_glorot(h->L, table_dev, rows, seed, (cudaStream_t)stream), "kge_init_glorot_uniform");
    return KGE_OK;
}

extern "C" int kge_init_table(kge_handle *h, float *table_dev, int64_t rows, int32_t kind, float a, float b, uint64_t seed,
                              void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_init_table");
    if (rows < 0 || (rows > 0 && !table_dev)) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_init_table: bad argument");
    if (kind < KGE_INIT_UNIFORM || kind > KGE_INIT_CONSTANT) return fail(KGE_ERR_INVALID_ARGUMENT, "Unknown initializer kind: %d", kind);
    if ((kind == KGE_INIT_NORMAL || kind == KGE_INIT_TRUNCATED_NORMAL) && b < 0.f)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_init_table: negative stddev");
    KGE_CUDA(launch_init_table(h->L, table_dev, rows, kind, a, b, seed, (cudaStream_t)stream), "kge_init_table");
    return KGE_OK;
}

static int refresh_rotation(kge_handle *h, const float *rel_dev, cudaStream_t st)
{
    if (h->cfg.scoring != KGE_ROTATE) return KGE_OK;
    KGE_CUDA(launch_rotation_table(rel_dev, h->rot, h->cfg.n_rel, h->L.kp, h->L.ld, h->rot_div, st),
             "rotation table");
    return KGE_OK;
}

extern "C" int kge_score_triples(kge_handle *h, const float *ent_dev, const float *rel_dev,
                                 const int32_t *triples_dev, int64_t n, float *scores_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_score_triples");
    if (n < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_score_triples: n < 0");
    if (n == 0) return KGE_OK;
    if (!ent_dev || !rel_dev || !triples_dev || !scores_dev)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_score_triples: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = refresh_rotation(h, rel_dev, st)) return rc;
    const float *rel = h->cfg.scoring == KGE_ROTATE ? h->rot : rel_dev;
    KGE_CUDA(launch_score_triples(h->L, h->nit, ent_dev, rel, triples_dev, n, h->score_scale, scores_dev,
                                  h->sm_count, st),
             "kge_score_triples");
    return KGE_OK;
}

extern "C" int kge_generate_corruptions(kge_handle *h, const int32_t *triples_dev, int64_t B, uint64_t seed,
                                        uint64_t step, int32_t *corruptions_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_generate_corruptions");
    if (B < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_generate_corruptions: B < 0");
    if (B == 0) return KGE_OK;
    if (!triples_dev || !corruptions_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_generate_corruptions: null pointer");
    KGE_CUDA(launch_corruptions(triples_dev, B, h->cfg.eta, seed, step, (unsigned)h->cfg.n_ent, corruptions_dev,
                                (cudaStream_t)stream),
             "kge_generate_corruptions");
    return KGE_OK;
}

// ---- host-side replay of the corruption stream (no GPU, no handle) -----------------------------------------
extern "C" void kge_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const u32x4 v = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}

extern "C" int kge_host_corruptions(const int32_t *triples_host, int64_t B, int32_t eta, int64_t n_ent, uint64_t seed,
                                    uint64_t step, int32_t *corruptions_host)
{
    if (B < 0 || eta < 1 || n_ent < 1 || n_ent > 0x7fffffffLL)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_host_corruptions: need B >= 0, eta >= 1, 1 <= n_ent < 2^31");
    if (B == 0) return KGE_OK;
    if (!triples_host || !corruptions_host) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_host_corruptions: null pointer");
    for (int64_t r = 0; r < B * (int64_t)eta; ++r) {  // same draw, same row order as kge_corruptions_kernel
        const int64_t i = r % B;
        int keep, repl;
        draw_corruption(seed, step, (unsigned long long)r, (uint32_t)n_ent, &keep, &repl);
        corruptions_host[3 * r + 0] = keep ? triples_host[3 * i + 0] : repl;
        corruptions_host[3 * r + 1] = triples_host[3 * i + 1];
        corruptions_host[3 * r + 2] = keep ? repl : triples_host[3 * i + 2];
    }
    return KGE_OK;
}

static int check_shard_map(const kge_handle *h, const kge_shard_map *map, const char *fn)
{
    if (!map || map->struct_size != (int32_t)sizeof(kge_shard_map))
        return fail(KGE_ERR_INVALID_ARGUMENT, "%s: bad kge_shard_map (ABI mismatch)", fn);
    if (map->world < 1 || map->world > KGE_MAX_PEERS || map->rows_per_shard < 1)
        return fail(KGE_ERR_INVALID_ARGUMENT, "%s: world must be 1..%d, rows_per_shard >= 1", fn, KGE_MAX_PEERS);
    if (map->rows_per_shard * (int64_t)map->world < h->cfg.n_ent)
        return fail(KGE_ERR_INVALID_ARGUMENT, "%s: world*rows_per_shard < n_ent", fn);
    for (int q = 0; q < map->world; ++q)
        if (!map->ent[q]) return fail(KGE_ERR_INVALID_ARGUMENT, "%s: null entity shard pointer for rank %d", fn, q);
    return KGE_OK;
}

static int train_step_impl(kge_handle *h, int32_t mode, const kge_shard_map *map, const float *ent_dev,
                           const float *rel_dev, float *grad_ent_dev, float *grad_rel_dev,
                           const int32_t *triples_dev, int64_t B, const int32_t *neg_ent_dev,
                           const uint8_t *neg_keep_subj_dev, uint64_t seed, uint64_t step, double *loss_dev,
                           float *scores_pos_dev, float *scores_neg_dev, const float *dpos_dev,
                           const float *dneg_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_train_step");
    if (mode < KGE_STEP_FUSED || mode > KGE_STEP_BACKWARD_EXT)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: unknown mode %d", mode);
    if (B < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: B < 0");
    if (B == 0) return KGE_OK;
    if (!ent_dev || !rel_dev || !triples_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: null table/triples");
    if (mode != KGE_STEP_FORWARD_ONLY && (!grad_ent_dev || !grad_rel_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: null gradient buffer");
    if ((neg_ent_dev == nullptr) != (neg_keep_subj_dev == nullptr))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: neg_ent_dev and neg_keep_subj_dev must be given together");
    if (mode == KGE_STEP_BACKWARD_EXT && (!dpos_dev || !dneg_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: BACKWARD_EXT needs dpos_dev and dneg_dev");
    if (mode == KGE_STEP_FORWARD_ONLY && (!scores_pos_dev || !scores_neg_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step: FORWARD_ONLY needs both score outputs");
    if (h->warps < 1)
        return fail(KGE_ERR_UNSUPPORTED, "kge_train_step: one positive's working set (%d B) exceeds shared memory (%d B)",
                    h->region_bytes, h->max_smem);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = refresh_rotation(h, rel_dev, st)) return rc;

    TrainParams p;
    memset(&p, 0, sizeof(p));
    p.ent = ent_dev;
    p.rel = h->cfg.scoring == KGE_ROTATE ? h->rot : rel_dev;
    p.grad_ent = grad_ent_dev;
    p.grad_rel = grad_rel_dev;
    p.triples = triples_dev;
    p.B = B;
    p.neg_ent = neg_ent_dev;
    p.neg_keep = neg_keep_subj_dev;
    p.seed = seed;
    p.step = step;
    p.n_ent = (unsigned)h->cfg.n_ent;
    p.model = h->cfg.scoring;
    p.eta = h->cfg.eta;
    p.kp = h->L.kp;
    p.ld = h->L.ld;
    p.nch = h->L.kp / 4;
    p.G = h->G;
    p.nbuf = h->nbuf;
    p.wk = h->wk;
    p.n_cb = h->n_cb;
    p.slot_floats = h->slot_floats;
    p.resident = h->resident;
    p.eta_pad = h->eta_pad;
    p.rows_bytes = h->rows_bytes;
    p.region_bytes = h->region_bytes;
    p.loss = h->cfg.loss;
    p.reduction = h->cfg.reduction;
    p.mode = mode;
    p.margin = h->cfg.margin;
    p.alpha = h->cfg.alpha;
    p.score_scale = h->score_scale;
    p.inv_div = 1.f / h->rot_div;
    p.loss_out = loss_dev;
    p.scores_pos = scores_pos_dev;
    p.scores_neg = scores_neg_dev;
    p.dpos = dpos_dev;
    p.dneg = dneg_dev;
    p.stamp = (int)(step & 0x3fffffffu) + 1;
    p.stamp_ent = h->stamp_ent;
    p.stamp_rel = h->stamp_rel;
    p.hot_ent[0] = h->hot_ent[0];
    p.hot_ent[1] = h->hot_ent[1];
    p.stash = (h->stash && h->stash_rows >= B * (int64_t)h->cfg.eta) ? h->stash : nullptr;
    if (map && map->world > 1) {
        p.shard_world = map->world;
        p.rows_per_shard = (int)map->rows_per_shard;
        for (int q = 0; q < map->world; ++q) { p.ent_shard[q] = map->ent[q]; p.grad_ent_shard[q] = map->grad_ent[q]; }
        p.stamp_ent = map->stamp_ent[0];
        for (int q = 0; q < map->world; ++q) {
            p.stamp_ent_shard[q] = map->stamp_ent[q];
            if (!map->stamp_ent[q]) p.stamp_ent = nullptr;  // all or nothing
        }
    }
    // KGE_B200_TRAIN_KERNEL=general keeps every shape on the general kernel (A/B runs, and the tests that cover its resident
    // trilinear instantiations); default: the specialised kernel of kge_train_res.cu where it applies
    const char *force = getenv("KGE_B200_TRAIN_KERNEL");
    const bool fast = h->res_warps > 0 && p.shard_world <= 1 && p.stash == nullptr && !(force && strcmp(force, "general") == 0);
    // Dynamic assignment of positives to warps (kge_train_common.cuh) where it was measured to pay: the resident fast path on
    // tables that live in L2 with enough positives per warp for the imbalance to matter -- cfg2 145 -> 137 us.  On an
    // HBM-resident table it LOSES (`big`, 3.2 GB: 668 -> 905 us, reproducibly; the general kernel on cfg4 / cfg5w: +-1 %),
    // so those keep the static stride (profiles/r2o_kbench_{static,dynamic}.log).  KGE_B200_TRAIN_SCHED=static|dynamic forces.
    {
        const char *ev = getenv("KGE_B200_TRAIN_SCHED");
        const long long n_warps = (long long)h->sm_count * (fast ? h->res_warps : h->warps);
        bool dyn = fast && B >= 8 * n_warps && 2ll * h->cfg.n_ent * h->L.ld * 4 <= (long long)h->l2_bytes;
        if (ev && strcmp(ev, "static") == 0) dyn = false;
        if (ev && strcmp(ev, "dynamic") == 0) dyn = true;
        p.sched = (dyn && B < (1ll << 24)) ? h->done_counter + 1 : nullptr;  // float counter: exact below 2^24
        // The hot-entity privatisation (kge_set_hot_entities) was a remedy for the same imbalance -- warps that drew many
        // positives of a hot entity fell behind on its serialised atomics -- and with the dynamic assignment it only costs:
        // bench.py's process 148.5 us with the hint, 138.3 us without (profiles/r2t_bench_probes*.log).  It stays in effect for
        // the static stride (153.8 vs 168 us there, profiles/r2h_*).
        if (p.sched) p.hot_ent[0] = p.hot_ent[1] = -1;
    }
    if (h->rot_warps > 0 && p.shard_world <= 1 && p.stash == nullptr && !(force && strcmp(force, "general") == 0)) {
        p.G = h->rot_G;
        p.rows_bytes = h->rot_rows_bytes;
        p.region_bytes = h->rot_region_bytes;
        KGE_CUDA(launch_train_rot(p, h->nit, h->sm_count, h->rot_warps * 32, (size_t)h->rot_warps * h->rot_region_bytes, st),
                 "kge_train_step");
        return KGE_OK;
    }
    if (fast) {
        p.rows_bytes = h->res_rows_bytes;
        p.region_bytes = h->res_region_bytes;
        KGE_CUDA(launch_train_res(p, h->nit, h->sm_count, h->res_warps * 32, (size_t)h->res_warps * h->res_region_bytes, st),
                 "kge_train_step");
        return KGE_OK;
    }
    KGE_CUDA(launch_train(p, h->nit, h->sm_count, h->warps * 32, (size_t)h->warps * h->region_bytes, st),
             "kge_train_step");
    return KGE_OK;
}

extern "C" int kge_train_step(kge_handle *h, int32_t mode, const float *ent_dev, const float *rel_dev,
                              float *grad_ent_dev, float *grad_rel_dev, const int32_t *triples_dev, int64_t B,
                              const int32_t *neg_ent_dev, const uint8_t *neg_keep_subj_dev, uint64_t seed,
                              uint64_t step, double *loss_dev, float *scores_pos_dev, float *scores_neg_dev,
                              const float *dpos_dev, const float *dneg_dev, void *stream)
{
    return train_step_impl(h, mode, nullptr, ent_dev, rel_dev, grad_ent_dev, grad_rel_dev, triples_dev, B, neg_ent_dev,
                           neg_keep_subj_dev, seed, step, loss_dev, scores_pos_dev, scores_neg_dev, dpos_dev, dneg_dev,
                           stream);
}

extern "C" int kge_train_step_sharded(kge_handle *h, int32_t mode, const kge_shard_map *map, const float *rel_dev,
                                      float *grad_rel_dev, const int32_t *triples_dev, int64_t B,
                                      const int32_t *neg_ent_dev, const uint8_t *neg_keep_subj_dev, uint64_t seed,
                                      uint64_t step, double *loss_dev, float *scores_pos_dev,
                                      float *scores_neg_dev, const float *dpos_dev, const float *dneg_dev,
                                      void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_train_step_sharded");
    if (int rc = check_shard_map(h, map, "kge_train_step_sharded")) return rc;
    if (mode != KGE_STEP_FORWARD_ONLY)
        for (int q = 0; q < map->world; ++q)
            if (!map->grad_ent[q]) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_train_step_sharded: null gradient shard pointer for rank %d", q);
    return train_step_impl(h, mode, map, map->ent[0], rel_dev, map->grad_ent[0], grad_rel_dev, triples_dev, B,
                           neg_ent_dev, neg_keep_subj_dev, seed, step, loss_dev, scores_pos_dev, scores_neg_dev,
                           dpos_dev, dneg_dev, stream);
}

static int check_optim(const kge_optimizer_config *opt, const char *fn)
{
    if (!opt || opt->struct_size != (int32_t)sizeof(kge_optimizer_config))
        return fail(KGE_ERR_INVALID_ARGUMENT, "%s: bad kge_optimizer_config (ABI mismatch)", fn);
    if (opt->kind < KGE_OPT_SGD || opt->kind > KGE_OPT_ADAGRAD)
        return fail(KGE_ERR_INVALID_ARGUMENT, "Could not interpret optimizer identifier: %d", opt->kind);
    if (opt->reg_p < 0 || opt->reg_p2 < 0 || (opt->reg_p == 0 && opt->reg_p2 != 0))
        return fail(KGE_ERR_INVALID_ARGUMENT, "%s: bad regulariser (reg_p >= 0; reg_p2 needs reg_p)", fn);
    return KGE_OK;
}
static RegParams reg_of(const kge_optimizer_config *opt)
{
    RegParams r;
    r.p = opt->reg_p; r.lambda = opt->reg_lambda; r.p2 = opt->reg_p2; r.lambda2 = opt->reg_lambda2;
    return r;
}
static void fill_optim(const kge_optimizer_config *opt, int64_t t, OptimParams &o)
{
    o.kind = opt->kind;
    o.lr = opt->learning_rate;
    o.beta1 = opt->beta_1;
    o.beta2 = opt->beta_2;
    o.eps = opt->epsilon;
    o.momentum = opt->momentum;
    o.reg = reg_of(opt);
    // legacy Adam: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), in fp64 on the host
    o.lr_t = (float)((double)opt->learning_rate * sqrt(1.0 - pow((double)opt->beta_2, (double)t)) /
                     (1.0 - pow((double)opt->beta_1, (double)t)));
}
static bool slots_ok(const kge_optimizer_config *opt, const float *s0, const float *s1)
{
    const bool need0 = opt->kind != KGE_OPT_SGD || opt->momentum != 0.f;
    const bool need1 = opt->kind == KGE_OPT_ADAM;
    return !((need0 && !s0) || (need1 && !s1));
}

extern "C" int kge_optimizer_step(kge_handle *h, const kge_optimizer_config *opt, int64_t t, float *table_dev,
                                  float *grad_dev, float *slot0_dev, float *slot1_dev, int64_t rows,
                                  double *reg_loss_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_optimizer_step");
    if (int rc = check_optim(opt, "kge_optimizer_step")) return rc;
    if (rows < 0 || t < 1) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step: rows >= 0 and t >= 1 required");
    if (rows == 0) return KGE_OK;
    if (!table_dev || !grad_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step: null table/grad");
    if (!slots_ok(opt, slot0_dev, slot1_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step: optimizer slot buffer missing");
    OptimParams o;
    fill_optim(opt, t, o);
    KGE_CUDA(launch_optimizer(o, table_dev, grad_dev, slot0_dev, slot1_dev, rows * (long long)h->L.ld, reg_loss_dev,
                              h->sm_count, (cudaStream_t)stream),
             "kge_optimizer_step");
    return KGE_OK;
}

extern "C" int kge_set_row_stamps(kge_handle *h, int32_t *ent_stamps_dev, int32_t *rel_stamps_dev)
{
    KGE_CHECK_HANDLE(h, "kge_set_row_stamps");
    h->stamp_ent = ent_stamps_dev;
    h->stamp_rel = rel_stamps_dev;
    return KGE_OK;
}

extern "C" int kge_rows_resident(const kge_handle *h) { return h ? (h->resident ? 1 : 0) : -1; }

extern "C" int kge_set_hot_entities(kge_handle *h, const int32_t *ids_host, int32_t n)
{
    KGE_CHECK_HANDLE(h, "kge_set_hot_entities");
    if (n < 0 || (n > 0 && !ids_host)) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_set_hot_entities: bad arguments");
    h->hot_ent[0] = h->hot_ent[1] = -1;
    for (int i = 0; i < n && i < 2; ++i) {
        if (ids_host[i] < 0 || ids_host[i] >= h->cfg.n_ent) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_set_hot_entities: id %d out of range", ids_host[i]);
        if (i == 1 && ids_host[1] == ids_host[0]) break;
        h->hot_ent[i] = ids_host[i];
    }
    return KGE_OK;
}

extern "C" int kge_set_row_stash(kge_handle *h, float *stash_dev, int64_t rows)
{
    KGE_CHECK_HANDLE(h, "kge_set_row_stash");
    if (rows < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_set_row_stash: rows < 0");
    h->stash = stash_dev;
    h->stash_rows = stash_dev ? rows : 0;
    return KGE_OK;
}

extern "C" int32_t kge_step_stamp(uint64_t step) { return (int32_t)(step & 0x3fffffffu) + 1; }

extern "C" int kge_optimizer_step_lazy(kge_handle *h, const kge_optimizer_config *opt, int64_t t, float *table_dev,
                                       float *grad_dev, float *slot0_dev, float *slot1_dev, int64_t rows,
                                       const int32_t *row_stamp_dev, int32_t stamp, double *reg_loss_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_optimizer_step_lazy");
    if (int rc = check_optim(opt, "kge_optimizer_step_lazy")) return rc;
    if (rows < 0 || t < 1) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_lazy: rows >= 0 and t >= 1 required");
    if (rows == 0) return KGE_OK;
    if (!table_dev || !grad_dev || !row_stamp_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_lazy: null pointer");
    if (!slots_ok(opt, slot0_dev, slot1_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_lazy: optimizer slot buffer missing");
    OptimParams o;
    fill_optim(opt, t, o);
    KGE_CUDA(launch_optimizer_lazy(o, table_dev, grad_dev, slot0_dev, slot1_dev, rows, h->L.ld, (const int *)row_stamp_dev,
                                   stamp, reg_loss_dev, h->sm_count, (cudaStream_t)stream),
             "kge_optimizer_step_lazy");
    return KGE_OK;
}

static int check_world(int32_t world, int32_t rank, const char *fn)
{
    if (world < 1 || world > KGE_MAX_PEERS || rank < 0 || rank >= world)
        return fail(KGE_ERR_INVALID_ARGUMENT, "%s: world must be 1..%d and 0 <= rank < world", fn, KGE_MAX_PEERS);
    return KGE_OK;
}

extern "C" int kge_optimizer_step_sharded(kge_handle *h, const kge_optimizer_config *opt, int64_t t, int32_t world,
                                          int32_t rank, float *const *peer_tables, float *const *peer_grads,
                                          float *slot0_shard_dev, float *slot1_shard_dev, int64_t row_begin,
                                          int64_t row_end, double *reg_loss_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_optimizer_step_sharded");
    if (int rc = check_optim(opt, "kge_optimizer_step_sharded")) return rc;
    if (int rc = check_world(world, rank, "kge_optimizer_step_sharded")) return rc;
    if (row_begin < 0 || row_end < row_begin || t < 1) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_sharded: bad row range / t");
    if (row_end == row_begin) return KGE_OK;
    if (!peer_tables || !peer_grads) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_sharded: null pointer arrays");
    for (int q = 0; q < world; ++q)
        if (!peer_tables[q] || !peer_grads[q]) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_sharded: null peer pointer for rank %d", q);
    if (!slots_ok(opt, slot0_shard_dev, slot1_shard_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_sharded: optimizer slot buffer missing");
    OptimParams o;
    fill_optim(opt, t, o);
    KGE_CUDA(launch_optimizer_sharded(o, world, rank, peer_tables, peer_grads, slot0_shard_dev, slot1_shard_dev,
                                      row_begin * (long long)h->L.ld, (row_end - row_begin) * (long long)h->L.ld,
                                      reg_loss_dev, h->sm_count, (cudaStream_t)stream),
             "kge_optimizer_step_sharded");
    return KGE_OK;
}

extern "C" int kge_optimizer_step_exchange(kge_handle *h, const kge_optimizer_config *opt_ent,
                                           const kge_optimizer_config *opt_rel, int64_t t, int32_t world, int32_t rank,
                                           float *const *peer_tables, float *const *peer_grads, float *table_mc_dev,
                                           const float *grads_mc_dev, float *zero_grads_dev,
                                           float *slot0_shard_dev, float *slot1_shard_dev, int64_t row_begin,
                                           int64_t row_end, uint32_t *const *peer_flags, uint32_t token, int32_t phases,
                                           double *reg_loss_dev, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_optimizer_step_exchange");
    if (int rc = check_optim(opt_ent, "kge_optimizer_step_exchange")) return rc;
    if (int rc = check_optim(opt_rel, "kge_optimizer_step_exchange")) return rc;
    if (int rc = check_world(world, rank, "kge_optimizer_step_exchange")) return rc;
    if (opt_ent->kind != opt_rel->kind || opt_ent->learning_rate != opt_rel->learning_rate || opt_ent->beta_1 != opt_rel->beta_1 ||
        opt_ent->beta_2 != opt_rel->beta_2 || opt_ent->epsilon != opt_rel->epsilon || opt_ent->momentum != opt_rel->momentum)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: the two tables may differ in their regulariser only");
    const int64_t total_rows = h->cfg.n_ent + h->cfg.n_rel;
    if (row_begin < 0 || row_end < row_begin || row_end > total_rows || t < 1)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: bad row range / t");
    if (phases < 0 || phases > 3) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: phases must be 0..3");
    if (!peer_tables || !peer_grads || (phases && !peer_flags))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: null pointer arrays");
    for (int q = 0; q < world; ++q)
        if (!peer_tables[q] || !peer_grads[q] || (phases && !peer_flags[q]))
            return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: null peer pointer for rank %d", q);
    if (row_end > row_begin && !slots_ok(opt_ent, slot0_shard_dev, slot1_shard_dev))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: optimizer slot buffer missing");
    OptimParams o;
    fill_optim(opt_ent, t, o);
    ExchangeParams x;
    memset(&x, 0, sizeof(x));
    x.world = world; x.rank = rank; x.phases = phases;
    for (int q = 0; q < world; ++q) { x.table[q] = peer_tables[q]; x.grad[q] = peer_grads[q]; x.flags[q] = phases ? peer_flags[q] : nullptr; }
    if ((table_mc_dev == nullptr) != (grads_mc_dev == nullptr))
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_optimizer_step_exchange: give both multicast mappings or neither");
    x.table_mc = table_mc_dev;
    x.grad_mc = grads_mc_dev;
    x.token = token;
    x.zero_grad = zero_grads_dev;
    const long long ld4 = h->L.ld / 4;
    x.total4 = total_rows * ld4;
    x.off4 = row_begin * ld4;
    x.n4 = (row_end - row_begin) * ld4;
    x.ent4 = h->cfg.n_ent * ld4;
    x.reg_ent = reg_of(opt_ent);
    x.reg_rel = reg_of(opt_rel);
    x.slot0 = slot0_shard_dev; x.slot1 = slot1_shard_dev;
    x.reg_loss = reg_loss_dev;
    x.done_counter = h->done_counter;
    x.trace = h->exchange_trace;
    KGE_CUDA(launch_optimizer_exchange(o, x, h->sm_count, (cudaStream_t)stream), "kge_optimizer_step_exchange");
    return KGE_OK;
}

extern "C" int kge_set_exchange_trace(kge_handle *h, uint64_t *stamps_dev)
{
    KGE_CHECK_HANDLE(h, "kge_set_exchange_trace");
    h->exchange_trace = reinterpret_cast<unsigned long long *>(stamps_dev);
    return KGE_OK;
}

extern "C" int kge_peer_barrier(kge_handle *h, int32_t world, int32_t rank, uint32_t *const *peer_flags, int32_t slot,
                                uint32_t token, void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_peer_barrier");
    if (int rc = check_world(world, rank, "kge_peer_barrier")) return rc;
    if (slot < 0 || slot > 1 || !peer_flags) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_peer_barrier: bad slot / null flags");
    for (int q = 0; q < world; ++q)
        if (!peer_flags[q]) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_peer_barrier: null flag pad for rank %d", q);
    KGE_CUDA(launch_peer_barrier(world, rank, peer_flags, slot, token, (cudaStream_t)stream), "kge_peer_barrier");
    return KGE_OK;
}

// ---- ranking ---------------------------------------------------------------------------------------------
// caller-owned workspace, carved into 256-byte aligned regions
struct RankWorkspace {
    float *qs, *qo, *qaux;   // [b, ld] query vectors (subject side / object side / RotatE object rows)
    int32_t *qpos, *cnt;     // [b], [b,3]
    bool tc;                 // tensor-core filter region present (KGE_RANK_MODE_AUTO, bilinear model, large enough call)
    RankTcLayout tcl;
    void *tc_base;
    size_t bytes;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static RankWorkspace carve_rank_workspace(const kge_handle *h, void *base, int64_t b, int64_t n_cand)
{
    RankWorkspace w;
    char *p = (char *)base;
    size_t off = 0;
    const size_t qbytes = align256((size_t)b * h->L.ld * sizeof(float));
    w.qs = (float *)(p + off); off += qbytes;
    w.qo = (float *)(p + off); off += qbytes;
    w.qaux = (float *)(p + off); off += qbytes;
    w.qpos = (int32_t *)(p + off); off += align256((size_t)b * sizeof(int32_t));
    w.cnt = (int32_t *)(p + off); off += align256((size_t)3 * b * sizeof(int32_t));
    w.tc = h->cfg.rank_mode == KGE_RANK_MODE_AUTO && rank_tc_applicable(h->L, 0, b, n_cand);
    w.tc_base = nullptr;
    if (w.tc) {
        off = (off + 1023) & ~(size_t)1023;
        w.tcl = rank_tc_layout(h->L, b, n_cand, h->cfg.rank_pair_cap);
        w.tc_base = p + off;
        off += w.tcl.bytes;
    }
    w.bytes = off;
    return w;
}

extern "C" int64_t kge_rank_workspace_bytes(const kge_handle *h, int64_t b, int64_t n_cand)
{
    if (!h || b < 0 || n_cand < 0) return 0;
    return (int64_t)carve_rank_workspace(h, nullptr, b, n_cand).bytes;
}

static int rank_impl(kge_handle *h, const kge_shard_map *map, int64_t filt_base, int32_t side, int32_t strategy,
                     const float *ent_dev, const float *rel_dev, const int32_t *triples_dev, int64_t b,
                     const int32_t *cand_ids_dev, int64_t cand_begin, int64_t n_cand, const int64_t *filt_off_dev,
                     const int32_t *filt_idx_dev, int64_t n_filt, int32_t *ranks_dev, int32_t *counts_dev,
                     float *scores_dev, void *workspace_dev, int64_t workspace_bytes, void *stream,
                     float *probe_delta_dev = nullptr)
{
    KGE_CHECK_HANDLE(h, "kge_rank");
    if (side != KGE_SIDE_S && side != KGE_SIDE_O) return fail(KGE_ERR_INVALID_ARGUMENT, "Invalid value for corrupt_side");
    if (strategy < KGE_RANK_WORST || strategy > KGE_RANK_MIDDLE)
        return fail(KGE_ERR_INVALID_ARGUMENT, "Invalid value for ranking_strategy");
    if (b < 0 || n_cand < 0 || cand_begin < 0 || n_filt < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: negative size");
    if (b == 0) return KGE_OK;
    if (!ent_dev || !rel_dev || !triples_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: null pointer");
    if (!ranks_dev && !counts_dev && !scores_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: no output buffer");
    if (cand_ids_dev && cand_begin != 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: cand_begin must be 0 with cand_ids_dev");
    if (!map && !cand_ids_dev && cand_begin + n_cand > h->cfg.n_ent)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: candidate range [%lld,%lld) exceeds n_ent %lld",
                    (long long)cand_begin, (long long)(cand_begin + n_cand), (long long)h->cfg.n_ent);
    if ((filt_off_dev == nullptr) != (filt_idx_dev == nullptr) && n_filt > 0)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: filter offsets and indices must be given together");
    const int64_t need = kge_rank_workspace_bytes(h, b, n_cand);
    if (!workspace_dev || workspace_bytes < need)
        return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: workspace of %lld bytes needed (kge_rank_workspace_bytes), got %lld",
                    (long long)need, (long long)(workspace_dev ? workspace_bytes : 0));
    if (((uintptr_t)workspace_dev & 1023u) != 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank: workspace must be 1024-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const RankWorkspace w = carve_rank_workspace(h, workspace_dev, b, n_cand);
    if (int rc = refresh_rotation(h, rel_dev, st)) return rc;
    ShardView sv;
    memset(&sv, 0, sizeof(sv));
    if (map && map->world > 1) {
        sv.world = map->world;
        sv.rows_per_shard = (int)map->rows_per_shard;
        for (int q = 0; q < map->world; ++q) sv.ent[q] = map->ent[q];
    }
    KGE_CUDA(launch_rank_prepare(h->L, sv, ent_dev, rel_dev, h->rot, triples_dev, b, h->score_scale, w.qs, w.qo, w.qaux,
                                 w.qpos, st),
             "kge_rank: prepare");
    // raw counters of THIS call in a zeroed scratch; added to the caller's accumulator or finalized into ranks_dev below
    int32_t *cnt = w.cnt;
    KGE_CUDA(cudaMemsetAsync(cnt, 0, (size_t)3 * b * sizeof(int32_t), st), "kge_rank: memset");
    RankParams p;
    memset(&p, 0, sizeof(p));
    p.L = h->L;
    p.side = side;
    p.strategy = strategy;
    p.ent = ent_dev;
    p.qvec = side == KGE_SIDE_S ? w.qs : w.qo;
    p.qaux = w.qaux;
    p.qpos = w.qpos;
    p.cand_ids = cand_ids_dev;
    p.cand_begin = cand_begin;
    p.n_cand = n_cand;
    p.b = b;
    p.scale = h->score_scale;
    p.filt_base = filt_base;
    p.scores = scores_dev;
    if (probe_delta_dev) {  // kge_rank_filter_probe: approximate scores + assumed bound, nothing is counted
        if (!w.tc || n_cand == 0) return fail(KGE_ERR_UNSUPPORTED, "kge_rank_filter_probe: the tensor-core filter does not apply to this call");
        p.scores = nullptr;
        KGE_CUDA(launch_rank_count_tc(p, w.tcl, w.tc_base, cnt, h->sm_count, st, scores_dev, probe_delta_dev), "kge_rank_filter_probe");
        return KGE_OK;
    }
    if (w.tc && !scores_dev && n_cand > 0)  // tensor-core filter + exact refine: bit-identical counters (kge_rank_tc.cu)
        KGE_CUDA(launch_rank_count_tc(p, w.tcl, w.tc_base, cnt, h->sm_count, st), "kge_rank: tensor-core count");
    else
        KGE_CUDA(launch_rank_count(p, cnt, st), "kge_rank: count");
    if (filt_off_dev && n_filt > 0)
        KGE_CUDA(launch_rank_filter_n(p, (const long long *)filt_off_dev, filt_idx_dev, n_filt, cnt, st), "kge_rank: filter");
    if (counts_dev) KGE_CUDA(launch_rank_accumulate(cnt, b, counts_dev, st), "kge_rank: accumulate");
    else if (ranks_dev) KGE_CUDA(launch_rank_finalize(cnt, b, strategy, ranks_dev, st), "kge_rank: finalize");
    return KGE_OK;
}

extern "C" int kge_rank(kge_handle *h, int32_t side, int32_t strategy, const float *ent_dev, const float *rel_dev,
                        const int32_t *triples_dev, int64_t b, const int32_t *cand_ids_dev, int64_t cand_begin,
                        int64_t n_cand, const int64_t *filt_off_dev, const int32_t *filt_idx_dev, int64_t n_filt,
                        int32_t *ranks_dev, int32_t *counts_dev, void *workspace_dev, int64_t workspace_bytes, void *stream)
{
    return rank_impl(h, nullptr, 0, side, strategy, ent_dev, rel_dev, triples_dev, b, cand_ids_dev, cand_begin, n_cand,
                     filt_off_dev, filt_idx_dev, n_filt, ranks_dev, counts_dev, nullptr, workspace_dev, workspace_bytes, stream);
}

extern "C" int kge_rank_sharded(kge_handle *h, const kge_shard_map *map, int32_t rank, int32_t side, int32_t strategy,
                                const float *rel_dev, const int32_t *triples_dev, int64_t b,
                                const int64_t *filt_off_dev, const int32_t *filt_idx_dev, int64_t n_filt,
                                int32_t *ranks_dev, int32_t *counts_dev, void *workspace_dev, int64_t workspace_bytes,
                                void *stream)
{
    if (!h) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank_sharded: null handle");
    if (int rc = check_shard_map(h, map, "kge_rank_sharded")) return rc;
    if (rank < 0 || rank >= map->world) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank_sharded: rank out of range");
    const int64_t first = (int64_t)rank * map->rows_per_shard;
    int64_t n_local = h->cfg.n_ent - first;
    if (n_local > map->rows_per_shard) n_local = map->rows_per_shard;
    if (n_local < 0) n_local = 0;
    return rank_impl(h, map, first, side, strategy, map->ent[rank], rel_dev, triples_dev, b, nullptr, 0, n_local,
                     filt_off_dev, filt_idx_dev, n_filt, ranks_dev, counts_dev, nullptr, workspace_dev, workspace_bytes, stream);
}

extern "C" int kge_rank_finalize(kge_handle *h, const int32_t *counts_dev, int64_t b, int32_t strategy, int32_t *ranks_dev,
                                 void *stream)
{
    KGE_CHECK_HANDLE(h, "kge_rank_finalize");
    if (strategy < KGE_RANK_WORST || strategy > KGE_RANK_MIDDLE)
        return fail(KGE_ERR_INVALID_ARGUMENT, "Invalid value for ranking_strategy");
    if (b < 0) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank_finalize: b < 0");
    if (b == 0) return KGE_OK;
    if (!counts_dev || !ranks_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank_finalize: null pointer");
    KGE_CUDA(launch_rank_finalize(counts_dev, b, strategy, ranks_dev, (cudaStream_t)stream), "kge_rank_finalize");
    return KGE_OK;
}

extern "C" int kge_corruption_scores(kge_handle *h, int32_t side, const float *ent_dev, const float *rel_dev,
                                     const int32_t *triples_dev, int64_t b, const int32_t *cand_ids_dev,
                                     int64_t cand_begin, int64_t n_cand, float *scores_dev, void *workspace_dev,
                                     int64_t workspace_bytes, void *stream)
{
    if (b > 0 && n_cand > 0 && !scores_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_corruption_scores: null output");
    return rank_impl(h, nullptr, 0, side, KGE_RANK_WORST, ent_dev, rel_dev, triples_dev, b, cand_ids_dev, cand_begin, n_cand,
                     nullptr, nullptr, 0, nullptr, nullptr, scores_dev, workspace_dev, workspace_bytes, stream);
}

extern "C" int kge_rank_filter_probe(kge_handle *h, int32_t side, const float *ent_dev, const float *rel_dev,
                                     const int32_t *triples_dev, int64_t b, const int32_t *cand_ids_dev, int64_t cand_begin,
                                     int64_t n_cand, float *approx_dev, float *delta_dev, void *workspace_dev,
                                     int64_t workspace_bytes, void *stream)
{
    if (!approx_dev || !delta_dev) return fail(KGE_ERR_INVALID_ARGUMENT, "kge_rank_filter_probe: null output");
    return rank_impl(h, nullptr, 0, side, KGE_RANK_WORST, ent_dev, rel_dev, triples_dev, b, cand_ids_dev, cand_begin, n_cand,
                     nullptr, nullptr, 0, nullptr, nullptr, approx_dev, workspace_dev, workspace_bytes, stream, delta_dev);
}
