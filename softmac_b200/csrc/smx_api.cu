// smx_api.cu -- host side of libsoftmac_b200.so: the C ABI declared in include/softmac_b200.h.
//
// Owns all device memory (per-substep particle checkpoints, grid, SDF tables, primitive time series,
// adjoint ping-pong buffers), the particle orderings produced by the periodic cell sort, and the
// stream-ordered launch sequences of the forward substep (mpm_simulator.py:320-337) and of its adjoint
// (mpm_simulator.py:339-378).  No CPU compute path exists here: without a CUDA device smx_create fails.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <set>
#include <string>
#include <algorithm>

#include "../../include/softmac_b200.h"
#include "smx_kernels.cuh"
#include "smx_sdf.cuh"
#include "smx_rigid.cuh"

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <thread>
#include <omp.h>

using namespace smx;

// Threads for the host-side f64 <-> f32 conversion loops.  Not taken from OMP_NUM_THREADS (torchrun exports
// OMP_NUM_THREADS=1 to every rank): SMX_HOST_THREADS, else hardware threads / LOCAL_WORLD_SIZE, capped at 16.
__attribute__((used)) static int smx_host_threads() {
    static int n = 0;
    if (n) return n;
    const char* e = getenv("SMX_HOST_THREADS");
    if (e && atoi(e) > 0) return n = atoi(e);
    int hw = (int)std::thread::hardware_concurrency();
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    int ranks = (lw && atoi(lw) > 0) ? atoi(lw) : 1;
    n = std::max(1, std::min(16, hw / ranks));
    return n;
}

// SMX_BACKTRACE=1: print a native backtrace on SIGSEGV (debug aid; resolve offsets with addr2line -e <lib>)
static void smx_segv_handler(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "\n[libsoftmac_b200] fatal signal, native backtrace:\n";
    ssize_t w = write(2, msg, sizeof msg - 1); (void)w;
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
__attribute__((constructor)) static void smx_install_handler() {
    const char* e = getenv("SMX_BACKTRACE");
    if (e && e[0] == '1') { signal(SIGSEGV, smx_segv_handler); signal(SIGABRT, smx_segv_handler); }
}

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(SMX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CKLN(sim, name) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(SMX_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); (sim)->launches++; if ((sim)->prof) { int r2_ = prof_mark((sim), (name)); if (r2_) return r2_; } } while (0)
#define CKL(sim) CKLN(sim, "other")
#define TRY(expr) do { int r_ = (expr); if (r_ != SMX_OK) return r_; } while (0)

namespace {

struct Order {                  // one physical ordering of the particles (an "epoch" of the cell sort)
    uint32_t* perm = nullptr;   // storage slot -> particle id (nullptr: identity)
    uint32_t* idx = nullptr;    // storage slot -> slot in the parent ordering (nullptr for roots)
    uint32_t* flags = nullptr;  // [nb^3] active-block flags
    uint32_t* blocks = nullptr; // active-block list
    int* nblocks = nullptr;     // device count
    int* ctrl_slot = nullptr;   // control_idx in storage order (lazily built)
    int ctrl_version = -1;
    bool live = true;
    long long uid = 0;          // never reused (ids are)
};

struct HostPrim { PrimDev d; float* sdf_dev = nullptr; float4* nrm_dev = nullptr; };

}  // namespace

#define SMX_IO_CHUNKS 16
struct smx_sim {
    smx_config cfg;
    Params P;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    bool pdl = true, pdl_grid = false;
    int pf_sc = 0, pf_g = 0, pf_g2p = 0, pf_g2pg = 0;   // L2 prefetch distances (particles): one wave of resident CTAs of the scatter / gather kernels
    int B = 1;                          // batched independent rollouts
    // spatial slab decomposition (one rank of several): owned x-block columns [slab_lo, slab_hi), neighbours present?
    bool slab = false, halo_lo = false, halo_hi = false;
    int slab_lo = 0, slab_hi = 0;
    bool dense = true;
    // particle frames
    float* pool = nullptr;
    long long frame_floats = 0;
    std::vector<int> slot_of, order_of, trans_from;
    std::vector<int> age_of;            // substeps run since the ordering of frame f was binned (re-sort when it reaches sort_every: also in copy mode, where frame indices never grow)
    int spare_slot = 0;
    std::vector<Order> orders;
    std::vector<Order> free_orders;     // recycled device buffers
    std::vector<int> dead_ids;
    long long next_uid = 1;
    // grid
    size_t G = 0;
    float4 *g_in = nullptr, *g_out = nullptr, *g_mix = nullptr, *gg_out = nullptr, *gg_out_b = nullptr, *gg_mix = nullptr, *g_lin = nullptr;
    uint8_t* mixflag = nullptr;     // [2][blocks]: blocks of gg_mix the contact adjoint of an (even / odd) substep scattered into (k_grid_grad sweeps only those)
    uint8_t* mixflag_of(int f, bool other = false) { return (mixflag && !slab) ? mixflag + (size_t)(((f & 1) != 0) != other) * (G / 64) : nullptr; }
    // adjoint grid of substep f: double-buffered by parity so that k_grid_grad(f) can already clear the one of substep f-1
    // halo exchange over peer memory (smx_slab_halo_*): receive slots + stamps + flags of both sides live in ONE allocation (`halo_mem`,
    // IPC-exportable) that the x-neighbours write into; peer_base[side] is the neighbour's allocation as mapped here
    // CUDA-graph replay of smx_step / smx_step_grad for launch-latency-bound scenes (smx_step_graph): one executable graph per direction,
    // re-targeted to the frames of every call by a whole-graph update of the freshly captured launch sequence
    std::map<size_t, cudaGraphExec_t> gexec[2];     // per direction, keyed by the node count of the captured sequence (its few recurring shapes)
    long long graph_launches = 0, graph_fallbacks = 0, graph_reinstantiations = 0;
    int slab_checked = -1;              // last frame whose positions went through k_check_slab
    unsigned char* halo_mem = nullptr; size_t halo_bytes = 0;
    unsigned char* peer_base[2] = {nullptr, nullptr}; bool peer_ipc[2] = {false, false};
    unsigned* halo_done = nullptr;      // [2] completion counters of the push kernels
    unsigned halo_seq[2] = {0u, 0u};    // exchanges done per side (both ends count in lockstep)
    long long halo_exchanges = 0;
    bool peers_on() const { return slab && (!halo_lo || peer_base[0]) && (!halo_hi || peer_base[1]) && (halo_lo || halo_hi) && halo_mem; }
    // slab mode driven phase by phase from the host (the caller exchanges the halos between the phases): no cross-substep fusion there
    bool slab_legacy() const { return slab && !peers_on(); }
    float4* gg_of(int f) { return (slab_legacy() || !(f & 1)) ? gg_out : gg_out_b; }
    int bwd_prepared = -1; long long bwd_prepared_uid = -1;   // substep whose g_out / g_mix / cleared gg are already in place (by k_grid_grad of the next substep)
    // grid checkpoints (g_in, g_out[, g_mix] on the active blocks of every substep) so that the adjoint does not
    // re-run P2G and the grid update (north_star: "per-substep state buffers resident in HBM rather than recomputed")
    float4* ckpt = nullptr;
    size_t ckpt_rec = 0;                // float4 entries per substep record
    int ckpt_cap = 0;                   // blocks reserved per array
    int ckpt_narr = 0;
    bool ckpt_enabled = true, ckpt_dirty = true;
    unsigned long long ckpt_overflow_seen = 0;
    int ckpt_cap_hint = 1;
    int* ckpt_need = nullptr;           // [max_steps] blocks the grid record of substep f needed (written by the saving kernel)
    int ckpt_need_max = 0;              // largest need seen by a backward pass: sizes the arena at the next reset
    int defer_save = -1;                // substep whose post-contact g_out / g_mix are saved by the NEXT substep's k_grid_op (smx_step)
    size_t ckpt_bytes = 0;
    std::vector<long long> ckpt_order;  // uid of the ordering the record of substep f was written in (-1: none)
    // SVD records (U, V, sigma - 1, J - 1 of the forward P2G of every substep) so that the adjoint does not repeat the SVD
    // per-substep "within reach of a primitive" bits of the forecast contact kernel (one word per warp of 32 slots)
    uint32_t* near_pool = nullptr;
    std::vector<long long> near_order;  // uid of the ordering the bits of substep f were written in (-1: none)
    size_t near_words() const { return ((size_t)std::max(P.n, 1) + 31) / 32; }
    size_t near_rec() const { return 32 + 2 * near_words(); }     // counter (padded) + reach bits + work list, per substep
    uint32_t* near_of(int f) { return near_pool + (size_t)f * near_rec(); }
    float4* svd_pool = nullptr;
    std::vector<long long> svd_order;   // uid of the ordering the SVD record of substep f was written in (-1: none)
    float4* svd_rec(int f) { return svd_pool + (long long)f * SMX_RPLANES * P.stride; }
    std::vector<char> ckpt_contact;
    // adjoint ping-pong
    float *adj_cur = nullptr, *adj_nxt = nullptr;
    int adj_frame = -1, adj_order = -1;
    bool adj_partial = false;           // inside smx_step_grad: adj_cur holds only the F planes of frame adj_frame (fused backward launches)
    bool bwd_fusion = true;             // P2G adjoint (f) + G2P adjoint (f-1) in one launch inside smx_step_grad (SMX_NO_BWD_FUSION=1: off)
    int grad_pending = -1, mid_done = -1, grad_mid_done = -1;
    float* ch_target = nullptr; int ch_m = 0; double* ch_loss = nullptr;   // Chamfer target cloud (m,3) and loss accumulator
    unsigned long long* cd_buf = nullptr; int cd_cap = 0;                 // contact-distance loss scratch: per (rollout, group) [min bits, id], winner slots, per-rollout loss
    int last_fwd = -1;
    long long g_in_clean_uid = -1;      // ordering whose active blocks of g_in are known to be zero (k_grid_op re-zeroes them)
    struct Seed { float* dev = nullptr; int ncols = 24; };
    std::map<int, Seed> seeds;          // frame -> device (n, 3 | 24) fp32 AoS in particle-id order
    std::multimap<size_t, float*> seed_pool;    // released seed buffers by size: clear_grads / add_*_grad never call cudaMalloc / cudaFree in the steady state
    // primitives
    std::vector<HostPrim> prims;
    PrimDev* prims_dev = nullptr;
    float* pstate = nullptr; double* pgrad = nullptr; double* ext_f = nullptr; float* ext_f_grad = nullptr;
    float* abuf = nullptr; double* gabuf = nullptr;     // velocity-control action buffers [np][T][6]
    // device-resident affine rigid coupling (smx_rigid_linear_*): one f64 arena holding the matrices and the per-env-step series
    bool rig_on = false; RigidLin rig = {}; double* rig_arena = nullptr; int* rig_enable = nullptr; unsigned char* rig_masks = nullptr;
    std::vector<double> rig_init;
    // control
    int* ctrl_id = nullptr; int ctrl_version = 0; float* action = nullptr; double* action_grad = nullptr;
    // staging / sort scratch
    float* stage_dev = nullptr; float* stage_host = nullptr;
    uint32_t *keys_a = nullptr, *keys_b = nullptr, *iota = nullptr;
    void* cub_tmp = nullptr; size_t cub_bytes = 0;
    unsigned long long* counters = nullptr;
    long long n_resorts = 0, last_ckpt_overflow = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_chunk[SMX_IO_CHUNKS] = {};  // pipelined host <-> device staging
    long long launches = 0;
    std::set<const void*> smem_optin;   // kernels whose dynamic shared-memory limit was raised on this handle's device
    bool prof = false;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;

    float* frame_ptr(int f) { return pool + (long long)slot_of[f] * frame_floats; }
    PrimSet primset() const {
        PrimSet ps; ps.prims = prims_dev; ps.pstate = pstate; ps.pgrad = pgrad; ps.ext_f = ext_f; ps.ext_f_grad = ext_f_grad; ps.T = cfg.max_steps;
        return ps;
    }
    bool has_contact() const {
        if (cfg.collision_type != 2) return false;
        for (auto& p : prims) if (p.d.enabled) return true;
        return false;
    }
};

struct smx_sim;
static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }
// Programmatic dependent launch for the kernels of the substep sequence: the next kernel's CTAs may become resident while the
// last CTAs of the previous one drain (every such kernel starts with griddepcontrol.wait, so it touches no data before its
// predecessor has completed and flushed).  SMX_NO_PDL=1 launches them with plain stream serialisation (ablation).
template <typename... KArgs, typename... Args>
static void launch_pdl(smx_sim* s, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args);
static int seed_alloc(smx_sim* s, size_t bytes, float** out) {
    auto it = s->seed_pool.find(bytes);
    if (it != s->seed_pool.end()) { *out = it->second; s->seed_pool.erase(it); return SMX_OK; }
    if (cudaMalloc(out, bytes) != cudaSuccess) { cudaGetLastError(); g_err[0] = 0; snprintf(g_err, sizeof g_err, "cannot allocate a %.1f MB adjoint seed buffer", bytes / 1e6); return SMX_ERR_NOMEM; }
    return SMX_OK;
}
static void seed_release(smx_sim* s, float* p, size_t bytes) { if (p) s->seed_pool.insert({bytes, p}); }
static inline size_t prim_slot_of(int b, int id) { return (size_t)b * SMX_MAXP + id; }
template <typename... KArgs, typename... Args>
static void launch_pdl(smx_sim* s, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = s->pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);       // errors are picked up by CKLN (cudaGetLastError)
}

// profiling: an event after every launch of a profiled substep (smx_profile_substep)
static int prof_mark(smx_sim* s, const char* name) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e, s->stream));
    s->marks.push_back({name, e});
    return SMX_OK;
}

static int sync_prims(smx_sim* s) {
    if (s->prims.empty()) return SMX_OK;
    std::vector<PrimDev> h(s->prims.size());
    for (size_t i = 0; i < h.size(); i++) h[i] = s->prims[i].d;
    CK(cudaMemcpyAsync(s->prims_dev, h.data(), h.size() * sizeof(PrimDev), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

static void free_order(Order& o) {
    cudaFree(o.perm); cudaFree(o.idx); cudaFree(o.flags); cudaFree(o.blocks); cudaFree(o.nblocks); cudaFree(o.ctrl_slot);
    o = Order(); o.live = false;
}

// Orderings no frame refers to any more are recycled: their device buffers go to a free list and are reused by
// the next re-sort on the same stream (no cudaMalloc / cudaFree, hence no implicit sync, in the steady state).
static void gc_orders(smx_sim* s) {
    std::vector<char> used(s->orders.size(), 0);
    for (size_t f = 0; f < s->order_of.size(); f++) {
        if (s->order_of[f] >= 0) used[s->order_of[f]] = 1;
        if (s->trans_from[f] >= 0) used[s->trans_from[f]] = 1;
    }
    if (s->adj_order >= 0) used[s->adj_order] = 1;
    for (size_t i = 0; i < s->orders.size(); i++)
        if (!used[i] && s->orders[i].live) {
            Order& o = s->orders[i];
            if (o.perm || o.idx || o.flags || o.ctrl_slot) { Order keep = o; keep.ctrl_version = -1; s->free_orders.push_back(keep); }
            o = Order(); o.live = false;
            s->dead_ids.push_back((int)i);
        }
}
static int new_order_id(smx_sim* s, const Order& o) {
    int id;
    if (!s->dead_ids.empty()) { id = s->dead_ids.back(); s->dead_ids.pop_back(); s->orders[id] = o; }
    else { s->orders.push_back(o); id = (int)s->orders.size() - 1; }
    s->orders[id].live = true;
    s->orders[id].uid = s->next_uid++;
    return id;
}
// an Order with perm / idx buffers allocated (recycled when possible)
static int alloc_order(smx_sim* s, Order& o, bool need_idx) {
    int n = std::max(s->P.n, 1);
    if (!s->free_orders.empty()) { o = s->free_orders.back(); s->free_orders.pop_back(); o.live = true; o.ctrl_version = -1; }
    if (!o.perm) CK(cudaMalloc(&o.perm, (size_t)n * sizeof(uint32_t)));
    if (need_idx && !o.idx) CK(cudaMalloc(&o.idx, (size_t)n * sizeof(uint32_t)));
    return SMX_OK;
}

static int build_blocks(smx_sim* s, Order& o, const float* frame, const uint32_t* keys_sorted = nullptr) {
    if (s->dense) return SMX_OK;
    int nb3 = s->B * s->P.nb3;
    if (!o.flags) {
        CK(cudaMalloc(&o.flags, nb3 * sizeof(uint32_t))); CK(cudaMalloc(&o.blocks, nb3 * sizeof(uint32_t))); CK(cudaMalloc(&o.nblocks, sizeof(int)));
    }
    CK(cudaMemsetAsync(o.flags, 0, nb3 * sizeof(uint32_t), s->stream));
    CK(cudaMemsetAsync(o.nblocks, 0, sizeof(int), s->stream));
    if (s->P.n > 0) {
        if (keys_sorted) k_mark_blocks_sorted<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P, keys_sorted, 2, o.flags);
        else k_mark_blocks<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P, frame, 2, o.flags);
        CKL(s);
    }
    if (s->slab) {
        int nb = s->P.nb;
        if (s->halo_lo) for (int c = s->slab_lo - 1; c <= s->slab_lo; c++) { k_mark_column<<<nblk(nb * nb, 256), 256, 0, s->stream>>>(o.flags, nb, c); CKL(s); }
        if (s->halo_hi) for (int c = s->slab_hi - 1; c <= s->slab_hi; c++) { k_mark_column<<<nblk(nb * nb, 256), 256, 0, s->stream>>>(o.flags, nb, c); CKL(s); }
    }
    k_compact_blocks<<<nblk(nb3, 256), 256, 0, s->stream>>>(nb3, o.flags, o.blocks, o.nblocks); CKL(s);
    return SMX_OK;
}

// (re)sort frame f by cell key: new ordering, permuted copy of the frame, active-block list
static int resort(smx_sim* s, int f, bool keep_transition) {
    int n = s->P.n;
    int old_id = s->order_of[f];
    gc_orders(s);
    Order no;
    float* src = s->frame_ptr(f);
    bool do_sort = !(s->cfg.flags & SMX_FLAG_NO_SORT) && n > 0;
    if (do_sort) {
        if (!s->dense && old_id >= 0 && s->orders[old_id].flags) { k_check_active<<<nblk(n, 256), 256, 0, s->stream>>>(s->P, src, s->orders[old_id].flags, s->counters); CKL(s); }
        k_keys<<<nblk(n, 256), 256, 0, s->stream>>>(s->P, src, s->keys_a, s->iota, s->counters); CKL(s);
        TRY(alloc_order(s, no, true));
        int bits = 0; while ((1ull << bits) < s->G) bits++;
        CK(cub::DeviceRadixSort::SortPairs(s->cub_tmp, s->cub_bytes, s->keys_a, s->keys_b, s->iota, no.idx, n, 0, bits, s->stream));
        s->launches += 3;
        const uint32_t* old_perm = old_id >= 0 ? s->orders[old_id].perm : nullptr;
        k_compose_perm<<<nblk(n, 256), 256, 0, s->stream>>>(n, old_perm, no.idx, no.perm); CKL(s);
        float* dst = s->pool + (long long)s->spare_slot * s->frame_floats;
        k_gather_frame<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, src, dst, no.idx); CKLN(s, "resort");
        std::swap(s->slot_of[f], s->spare_slot);
        s->n_resorts++;
    } else if (old_id >= 0 && s->orders[old_id].perm && n > 0) {
        TRY(alloc_order(s, no, false));
        CK(cudaMemcpyAsync(no.perm, s->orders[old_id].perm, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s->stream));
    }
    TRY(build_blocks(s, no, s->frame_ptr(f), do_sort ? s->keys_b : nullptr));
    s->order_of[f] = new_order_id(s, no);
    s->age_of[f] = 0;
    s->ckpt_order[f] = -1; s->svd_order[f] = -1; s->near_order[f] = -1;
    // keep_transition: frame f was produced by substep f-1 in the old ordering, so the adjoint has to be carried
    // back through idx; otherwise (user-written frame) the adjoint chain is cut here, as in the reference
    s->trans_from[f] = (keep_transition && do_sort) ? old_id : -1;
    return SMX_OK;
}

static int ctrl_slots(smx_sim* s, int order_id, const int** out) {
    *out = nullptr;
    if (s->cfg.n_control <= 0) return SMX_OK;
    Order& o = s->orders[order_id];
    if (!o.ctrl_slot) CK(cudaMalloc(&o.ctrl_slot, std::max(1, s->P.n) * sizeof(int)));
    if (o.ctrl_version != s->ctrl_version) {
        if (s->P.n > 0) { k_permute_i32<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P.n, s->ctrl_id, o.ctrl_slot, o.perm); CKL(s); }
        o.ctrl_version = s->ctrl_version;
    }
    *out = o.ctrl_slot;
    return SMX_OK;
}

template <typename F>
static int dispatch_mat(const Params& P, F&& fn) {
    const int material = P.material, ptype = P.ptype;
    int mat = material * 3 + ptype;
    switch (mat) {
        case 0:
#ifndef SMX_ONLY_MAT0
            if (P.vm) return fn(std::integral_constant<int, 6>());     // von Mises return mapping (smx_set_plasticity)
#endif
            return fn(std::integral_constant<int, 0>());
#ifndef SMX_ONLY_MAT0       // A/B variant builds (tools/gpu_variants.sh) instantiate the bench material only
        case 1: return fn(std::integral_constant<int, 1>());
        case 2: return fn(std::integral_constant<int, 2>());
        case 3: return fn(std::integral_constant<int, 4>());   // neo-Hookean "plastic" falls through to elastic (mpm_simulator.py:237-241)
        case 4: return fn(std::integral_constant<int, 4>());
        case 5: return fn(std::integral_constant<int, 5>());
#endif
    }
    return fail(SMX_ERR_ARG, "bad material_model/ptype %d/%d", material, ptype);
}

static int grid_blocks_launch(smx_sim* s) { return s->sm_count * 8; }

static int clear_grids(smx_sim* s, const Order& o, float4* a, float4* b, float4* c) {
    if (s->dense) {
        if (a) CK(cudaMemsetAsync(a, 0, s->G * sizeof(float4), s->stream));
        if (b) CK(cudaMemsetAsync(b, 0, s->G * sizeof(float4), s->stream));
        if (c) CK(cudaMemsetAsync(c, 0, s->G * sizeof(float4), s->stream));
    } else {
        k_clear_blocks<<<grid_blocks_launch(s), 256, 0, s->stream>>>(o.blocks, o.nblocks, a, b, c); CKLN(s, "clear");
        return SMX_OK;
    }
    if (s->prof) TRY(prof_mark(s, "clear"));
    return SMX_OK;
}

// Size the grid-checkpoint arena from the ordering of frame f: one record per substep holding g_in, g_out (+ g_mix
// with contact) of up to `cap` active blocks, cap = 1.5 x the current active-block count (the count only grows slowly as
// the material spreads; a record that does not fit is flagged on the device and the adjoint recomputes instead).
// One host sync, done once after reset / set_frame.
static int ensure_ckpt(smx_sim* s, int f) {
    s->ckpt_dirty = false;
    if (!s->ckpt_enabled) return SMX_OK;
    int total = s->B * s->P.nb3, nbk = total;
    Order& o = s->orders[s->order_of[f]];
    if (!s->dense && o.nblocks) {
        CK(cudaStreamSynchronize(s->stream));
        CK(cudaMemcpy(&nbk, o.nblocks, sizeof(int), cudaMemcpyDeviceToHost));
    }
    long long want = std::max<long long>((long long)s->ckpt_cap_hint * (nbk + nbk / 2) + 64, (long long)s->ckpt_need_max + s->ckpt_need_max / 8 + 64);
    int cap = (int)std::min<long long>(total, want);
    int narr = s->has_contact() ? 3 : 2;
    if (s->ckpt && cap <= s->ckpt_cap && narr <= s->ckpt_narr) return SMX_OK;
    if (s->ckpt) { CK(cudaStreamSynchronize(s->stream)); cudaFree(s->ckpt); s->ckpt = nullptr; }
    std::fill(s->ckpt_order.begin(), s->ckpt_order.end(), -1);
    size_t rec = (size_t)narr * cap * 64;
    size_t bytes = (size_t)s->cfg.max_steps * rec * sizeof(float4), free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (bytes > free_b / 2 || cudaMalloc(&s->ckpt, bytes) != cudaSuccess) { cudaGetLastError(); s->ckpt = nullptr; s->ckpt_rec = 0; s->ckpt_cap = 0; return SMX_OK; }
    s->ckpt_rec = rec; s->ckpt_cap = cap; s->ckpt_narr = narr; s->ckpt_bytes = bytes;
    return SMX_OK;
}

static int forward_p2g(smx_sim* s, int f, bool write_F, bool accumulate, bool fuse_prev_g2p = false) {
    const Params& P = s->P;
    Order& o = s->orders[s->order_of[f]];
    float* fin = s->frame_ptr(f);
    const float* fprev = fuse_prev_g2p ? s->frame_ptr(f - 1) : nullptr;
    const float4* gprev = fuse_prev_g2p ? s->g_out : nullptr;
    float* fout = write_F ? s->frame_ptr(f + 1) : nullptr;
    PrimSet ps = s->primset();
    const int* cslot = nullptr;
    TRY(ctrl_slots(s, s->order_of[f], &cslot));
    if (s->g_in_clean_uid != o.uid) TRY(clear_grids(s, o, s->g_in, nullptr, nullptr));
    s->g_in_clean_uid = -1;             // P2G is about to write it
    s->bwd_prepared = -1;
    if (s->slab && P.n > 0) {
        // particles that left slab + halo are counted; with the G2P of f-1 folded into this launch x[f] does not exist yet: look at x[f-1]
        const int cf = fprev ? f - 1 : f;
        if (s->slab_checked != cf) {
            k_check_slab<<<nblk(P.n, 256), 256, 0, s->stream>>>(P, fprev ? fprev : fin, s->halo_lo ? s->slab_lo - 1 : 0, s->halo_hi ? s->slab_hi : P.nb - 1, s->counters); CKL(s);
            s->slab_checked = cf;
        }
    }
    if (P.n > 0) {
        float4* rec = (write_F && s->svd_pool) ? s->svd_rec(f) : nullptr;
        const bool extra = P.ctype == 1 || P.n_control > 0;
        const int grid = nblk(P.n, SMX_TPB_SC), acc = accumulate ? 1 : 0;
        TRY(dispatch_mat(P, [&](auto mat) {
            constexpr int M = decltype(mat)::value;
            auto go = [&](auto staged_c, auto extra_c) {
                constexpr bool ST = decltype(staged_c)::value, EX = decltype(extra_c)::value;
                const size_t smem = ST ? (size_t)(SMX_TPB_SC / 32) * sizeof(WarpStage) : 0;
                // more than 48 KB of dynamic shared memory needs the per-device opt-in (remembered per handle)
                if (ST && s->smem_optin.insert((const void*)k_p2g<M, ST, EX>).second) {
                    cudaFuncSetAttribute(k_p2g<M, ST, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    cudaFuncSetAttribute(k_p2g<M, ST, EX>, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("SMX_CARVEOUT") ? atoi(getenv("SMX_CARVEOUT")) : (int)cudaSharedmemCarveoutMaxShared);   // SMX_CARVEOUT: A/B experiments (percent of shared memory)
                }
                launch_pdl(s, k_p2g<M, ST, EX>, grid, SMX_TPB_SC, smem, P, ps, f, fin, fout, s->g_in, cslot, s->action, acc, fprev, gprev, rec, s->pf_sc);
            };
            const bool staged = !(s->cfg.flags & SMX_FLAG_DIRECT_RED);
            if (staged) { if (extra) go(std::true_type(), std::true_type()); else go(std::true_type(), std::false_type()); }
            else { if (extra) go(std::false_type(), std::true_type()); else go(std::false_type(), std::false_type()); }
            CKLN(s, fprev ? "k_g2p2g" : "k_p2g"); return (int)SMX_OK;
        }));
        if (rec) s->svd_order[f] = o.uid;
    }
    if (write_F && s->cfg.rigid_velocity_control && !s->prims.empty()) {
        k_forward_kinematics<<<nblk((long long)s->prims.size() * s->B, 64), 64, 0, s->stream>>>(s->pstate, s->cfg.max_steps, (int)s->prims.size(), s->B, f, P.dt); CKL(s);
    }
    return SMX_OK;
}
static int forward_grid(smx_sim* s, int f, bool accumulate, bool checkpoint) {
    const Params& P = s->P;
    Order& o = s->orders[s->order_of[f]];
    PrimSet ps = s->primset();
    bool contact = s->has_contact();
    bool save = checkpoint && s->ckpt && (!contact || s->ckpt_narr == 3);
    float4* rec = save ? s->ckpt + (size_t)f * s->ckpt_rec : nullptr;
    // checkpoint == forward pass proper: also save the grid record and re-zero g_in for the next substep's P2G
    uint32_t* near = nullptr;
    if (contact && P.n > 0) {
        if (!s->near_pool && cudaMalloc(&s->near_pool, (size_t)s->cfg.max_steps * s->near_rec() * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); s->near_pool = nullptr; }
        near = s->near_pool ? s->near_of(f) : nullptr;
    }
    // the previous substep of a fused smx_step sequence left its post-contact record to this launch (same ordering, same block list)
    float4* rec_prev = nullptr;
    if (checkpoint && s->defer_save == f - 1 && f > 0 && s->ckpt && s->order_of[f - 1] == s->order_of[f]) rec_prev = s->ckpt + (size_t)(f - 1) * s->ckpt_rec;
    { const bool saved_pdl = s->pdl; s->pdl = s->pdl && s->pdl_grid;
    if (P.ctype == 0) launch_pdl(s, k_grid_op<true>, grid_blocks_launch(s), 256, 0, P, ps, f, s->dense ? nullptr : o.blocks, o.nblocks, s->g_in, s->g_out, contact ? s->g_mix : nullptr,
                                                           accumulate ? 1 : 0, rec, s->ckpt_cap, contact ? 0 : 1, checkpoint ? 1 : 0, s->counters, near, s->ckpt_need + f, rec_prev);
    else launch_pdl(s, k_grid_op<false>, grid_blocks_launch(s), 256, 0, P, ps, f, s->dense ? nullptr : o.blocks, o.nblocks, s->g_in, s->g_out, contact ? s->g_mix : nullptr,
                                                           accumulate ? 1 : 0, rec, s->ckpt_cap, contact ? 0 : 1, checkpoint ? 1 : 0, s->counters, near, s->ckpt_need + f, rec_prev);
    s->pdl = saved_pdl; }
    CKLN(s, "k_grid_op");
    if (rec_prev) { s->ckpt_order[f - 1] = o.uid; s->ckpt_contact[f - 1] = 1; }
    if (checkpoint) s->defer_save = -1;
    if (checkpoint) s->g_in_clean_uid = o.uid;
    if (contact && P.n > 0) {
        float life = 1.0f / (float)(P.substeps - f % P.substeps);      // mpm_simulator.py:425 (f32 in the reference too)
        launch_pdl(s, k_contact, nblk(P.n, SMX_TPB), SMX_TPB, 0, P, ps, f, life, s->frame_ptr(f), s->g_mix, s->g_out, accumulate ? 1 : 0, near, (int)s->near_words()); CKLN(s, "k_contact");
        s->near_order[f] = near ? o.uid : -1;
    }
    if (save && !contact) { s->ckpt_order[f] = o.uid; s->ckpt_contact[f] = 0; }
    return SMX_OK;
}
// with contact, g_out is final only after the contact scatter (and, in slab mode, after the halo exchange of that scatter)
static int forward_grid_save_contact(smx_sim* s, int f) {
    Order& o = s->orders[s->order_of[f]];
    if (!(s->has_contact() && s->ckpt && s->ckpt_narr == 3)) return SMX_OK;
    launch_pdl(s, k_ckpt_copy, grid_blocks_launch(s), 256, 0, s->dense ? nullptr : o.blocks, o.nblocks, s->B * s->P.nb3, s->ckpt_cap, s->ckpt + (size_t)f * s->ckpt_rec,
                                                             nullptr, s->g_out, s->g_mix, 0, s->counters, nullptr, nullptr, s->ckpt_need + f);
    CKLN(s, "ckpt_save");
    s->ckpt_order[f] = o.uid; s->ckpt_contact[f] = 1;
    return SMX_OK;
}
// P2G + grid update + forecast contact of substep f (everything before G2P); shared by forward and adjoint
static int forward_to_grid(smx_sim* s, int f, bool write_F, bool accumulate) {
    TRY(forward_p2g(s, f, write_F, accumulate));
    return forward_grid(s, f, accumulate, false);
}

static int check_frame(smx_sim* s, int f, const char* what) {
    if (!s) return fail(SMX_ERR_ARG, "%s: null simulator", what);
    if (f < 0 || f >= s->cfg.max_steps) return fail(SMX_ERR_RANGE, "%s: frame %d outside [0, %d)", what, f, s->cfg.max_steps);
    return SMX_OK;
}
static int check_prim(smx_sim* s, int id, const char* what) {
    if (!s) return fail(SMX_ERR_ARG, "%s: null simulator", what);
    if (id < 0 || id >= (int)s->prims.size()) return fail(SMX_ERR_RANGE, "%s: primitive %d outside [0, %d)", what, id, (int)s->prims.size());
    return SMX_OK;
}

// Pipelined staging: host f64 -> pinned fp32 -> device in SMX_IO_CHUNKS pieces, the conversion of piece k+1 (host threads)
// overlapping the H2D copy of piece k; and the reverse for read-back (the conversion of piece k overlaps the D2H copy of k+1).
// `fill(dst, i0, i1)` writes staging elements [i0, i1).  The caller has synchronised the stream (staging buffer reuse).
static size_t io_chunk(size_t cnt) {
    static const bool single = getenv("SMX_IO_SINGLE_CHUNK") != nullptr;      // ablation: one conversion, then one copy
    if (single) return std::max<size_t>(cnt, 1);
    size_t chunk = (cnt + SMX_IO_CHUNKS - 1) / SMX_IO_CHUNKS;
    return std::max<size_t>((chunk + 1023) / 1024 * 1024, 1 << 16);
}
template <typename Fill>
static int h2d_pipelined(smx_sim* s, size_t cnt, Fill&& fill) {
    size_t chunk = io_chunk(cnt);
    for (size_t i0 = 0; i0 < cnt; i0 += chunk) {
        size_t i1 = std::min(cnt, i0 + chunk);
        fill(s->stage_host, i0, i1);
        CK(cudaMemcpyAsync(s->stage_dev + i0, s->stage_host + i0, (i1 - i0) * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    }
    return SMX_OK;
}
static int d2h_pipelined(smx_sim* s, size_t cnt, double* host) {
    // measured on the B200 host (16 threads, 96 MB): overlapping the conversion with the copy is SLOWER for read-back (9.1 vs 7.9 ms:
    // the first-touch page faults of the caller's fresh f64 array compete with the DMA), so one piece unless SMX_IO_D2H_CHUNKS is set
    static const bool piecewise = getenv("SMX_IO_D2H_CHUNKS") != nullptr;
    size_t chunk = piecewise ? io_chunk(cnt) : std::max<size_t>(cnt, 1);
    int k = 0;
    for (size_t i0 = 0; i0 < cnt; i0 += chunk, k++) {
        size_t i1 = std::min(cnt, i0 + chunk);
        CK(cudaMemcpyAsync(s->stage_host + i0, s->stage_dev + i0, (i1 - i0) * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaEventRecord(s->ev_chunk[k], s->stream));
    }
    // one parallel region for all pieces (the worker threads stay hot): thread 0 waits for the copy of piece k, then everybody converts it
    const int nchunks = k;
    const float* src = s->stage_host;
    cudaError_t err = cudaSuccess;
    #pragma omp parallel num_threads(smx_host_threads())
    {
        for (int c = 0; c < nchunks; c++) {
            #pragma omp master
            { cudaError_t e = cudaEventSynchronize(s->ev_chunk[c]); if (e != cudaSuccess) err = e; }
            #pragma omp barrier
            const long long i0 = (long long)c * (long long)chunk, i1 = (long long)std::min(cnt, (size_t)i0 + chunk);
            #pragma omp for schedule(static) nowait
            for (long long i = i0; i < i1; i++) host[i] = (double)src[i];
        }
    }
    if (err != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(err));
    return SMX_OK;
}

// host (n, ncomp) f64 in id order -> frame components [c0, c0+ncomp) in storage order
static int upload_cols(smx_sim* s, int f, const double* host, int ncomp, int c0) {
    int n = s->P.n;
    if (n == 0) return SMX_OK;
    size_t cnt = (size_t)n * ncomp;
    CK(cudaStreamSynchronize(s->stream));       // staging buffer reuse
    auto fill = [&](float* dst, size_t i0, size_t i1) {
        #pragma omp parallel for schedule(static) num_threads(smx_host_threads())
        for (long long i = (long long)i0; i < (long long)i1; i++) dst[i] = (float)host[i];
    };
        TRY(h2d_pipelined(s, cnt, fill));
    const uint32_t* perm = s->orders[s->order_of[f]].perm;
    k_upload<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->stage_dev, ncomp, c0, s->frame_ptr(f), perm, 0); CKL(s);
    return SMX_OK;
}
static void launch_plain_download(smx_sim* s, const float* frame, const uint32_t* perm, int ncomp, int c0) {
    k_download<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P.n, s->P.stride, s->stage_dev, ncomp, c0, frame, perm);
}
static int download_cols(smx_sim* s, const float* frame, const uint32_t* perm, double* host, int ncomp, int c0) {
    int n = s->P.n;
    if (n == 0) return SMX_OK;
    size_t cnt = (size_t)n * ncomp;
    k_download<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->stage_dev, ncomp, c0, frame, perm); CKL(s);
    return d2h_pipelined(s, cnt, host);
}
static int ensure_order(smx_sim* s, int f) {
    if (s->order_of[f] >= 0) return SMX_OK;
    // a frame that was never written: give it an identity ordering
    Order o;
    int id = new_order_id(s, o);
    s->order_of[f] = id;
    s->trans_from[f] = -1;
    CK(cudaMemsetAsync(s->frame_ptr(f), 0, s->frame_floats * sizeof(float), s->stream));
    return build_blocks(s, s->orders[id], s->frame_ptr(f));
}
static int apply_seed(smx_sim* s, int f, float* adj, int order_id) {
    auto it = s->seeds.find(f);
    if (it == s->seeds.end() || s->P.n == 0) return SMX_OK;
    k_upload<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P.n, s->P.stride, it->second.dev, it->second.ncols, 0, adj, s->orders[order_id].perm, 1); CKL(s);
    return SMX_OK;
}

// ---- CUDA-graph replay helpers (smx_step_graph below) ----------------------------------------------------------------------------
static bool graph_ok_common(smx_sim* s) {
    return s->P.n > 0 && !s->slab && !s->prof && !s->ckpt_dirty && !s->cfg.rigid_velocity_control && s->cfg.n_control == 0 &&
           !(s->cfg.flags & (SMX_FLAG_NO_SORT | SMX_FLAG_DENSE_GRID | SMX_FLAG_NO_GRID_CKPT)) && (!s->has_contact() || s->near_pool) && s->ckpt;
}
template <typename Body>
static int graph_run(smx_sim* s, int dir, Body&& body) {
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    if (rc != SMX_OK || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (rc != SMX_OK) return rc;
        return fail(SMX_ERR_CUDA, "smx_step_graph: stream capture failed: %s", cudaGetErrorString(e));
    }
    size_t nodes = 0;
    cudaGraphGetNodes(g, nullptr, &nodes);
    cudaGraphExec_t& ex = s->gexec[dir][nodes];
    if (ex) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(ex, g, &info) != cudaSuccess) { cudaGetLastError(); cudaGraphExecDestroy(ex); ex = nullptr; s->graph_reinstantiations++; }
    }
    if (!ex && cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) {
        const cudaError_t e2 = cudaGetLastError();
        cudaGraphDestroy(g); ex = nullptr;
        return fail(SMX_ERR_CUDA, "smx_step_graph: cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
    }
    const cudaError_t e3 = cudaGraphLaunch(ex, s->stream);
    cudaGraphDestroy(g);
    if (e3 != cudaSuccess) return fail(SMX_ERR_CUDA, "smx_step_graph: cudaGraphLaunch failed: %s", cudaGetErrorString(e3));
    s->graph_launches++;
    return SMX_OK;
}

// =================================================================================================
extern "C" {

const char* smx_last_error(void) { return g_err; }

static int create_body(smx_sim* s, const smx_config* cfg);
int smx_create(const smx_config* cfg, smx_sim** out) {
    if (!cfg || !out) return fail(SMX_ERR_ARG, "smx_create: null argument");
    if (cfg->n_particles < 0 || cfg->n_grid < 8 || cfg->n_grid % 4 || cfg->max_steps < 2) return fail(SMX_ERR_ARG, "smx_create: need n_particles >= 0, n_grid >= 8 and a multiple of 4, max_steps >= 2");
    if (cfg->substeps < 1) return fail(SMX_ERR_ARG, "smx_create: substeps must be >= 1");
    if (cfg->material_model < 0 || cfg->material_model > 1 || cfg->ptype < 0 || cfg->ptype > 2 || cfg->collision_type < 0 || cfg->collision_type > 2)
        return fail(SMX_ERR_ARG, "smx_create: material_model in {0,1}, ptype in {0,1,2}, collision_type in {0,1,2}");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(SMX_ERR_CUDA, "smx_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SMX_ERR_ARG, "smx_create: device %d outside [0, %d)", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));
    smx_sim* s = new smx_sim();
    s->cfg = *cfg;
    const int rc = create_body(s, cfg);
    if (rc != SMX_OK) { smx_destroy(s); return rc; }     // frees whatever was allocated before the failure (g_err keeps the message)
    *out = s;
    return SMX_OK;
}
static int create_body(smx_sim* s, const smx_config* cfg) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device));
    s->sm_count = prop.multiProcessorCount;
    s->pf_sc = s->sm_count * SMX_SC_MINB * SMX_TPB_SC; s->pf_g = s->sm_count * SMX_P2GG_MINB * SMX_TPB; s->pf_g2p = s->sm_count * 8 * SMX_TPB; s->pf_g2pg = s->sm_count * SMX_G2PG_MINB * SMX_TPB_G2PG;
    s->pdl = getenv("SMX_NO_PDL") == nullptr;
    // the grid kernels are launched with plain stream serialisation: at 32 registers all their CTAs become resident next to the
    // draining particle kernel and PDL then costs 10 % (measured 3.47 vs 3.84 G/s); SMX_PDL_GRID=1 turns it on for experiments
    s->pdl_grid = getenv("SMX_PDL_GRID") != nullptr;
    s->bwd_fusion = getenv("SMX_NO_BWD_FUSION") == nullptr;
    if (getenv("SMX_NO_PREFETCH")) s->pf_sc = s->pf_g = s->pf_g2p = s->pf_g2pg = 1 << 30;
    if (cfg->stream || (cfg->flags & SMX_FLAG_EXTERNAL_STREAM)) s->stream = (cudaStream_t)cfg->stream;   // NULL + flag: the legacy default stream
    else { CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->own_stream = true; }
    Params& P = s->P;
    int n = cfg->n_particles;
    const int B = std::max(cfg->n_batch, 1);
    if (B > 255) return fail(SMX_ERR_ARG, "smx_create: at most 255 batched rollouts per handle");
    if (cfg->n_grid > 256) return fail(SMX_ERR_ARG, "smx_create: n_grid <= 256 (8-bit base-cell packing of the staged scatter)");
    if ((long long)B * cfg->n_particles > 2000000000LL) return fail(SMX_ERR_ARG, "smx_create: n_batch * n_particles too large");
    if ((long long)B * cfg->n_grid * cfg->n_grid * cfg->n_grid >= (1LL << 31)) return fail(SMX_ERR_ARG, "smx_create: n_batch * n_grid^3 must stay below 2^31 (32-bit node indices)");
    P.nbatch = B; P.npb = cfg->n_particles;
    n = B * cfg->n_particles;           // total particle slots of the handle
    P.n = n; P.stride = ((long long)std::max(n, 1) + 31) / 32 * 32;
    P.ng = cfg->n_grid; P.nb = cfg->n_grid / 4;
    double dx = 1.0 / cfg->n_grid, p_vol = (dx * 0.5) * (dx * 0.5), p_mass = p_vol;       // mpm_simulator.py:32-35
    double mu = cfg->E / (2 * (1 + cfg->nu)), lam = cfg->E * cfg->nu / ((1 + cfg->nu) * (1 - 2 * cfg->nu));
    if (cfg->ptype == 1) { mu *= 0.3; lam *= 0.3; } else if (cfg->ptype == 2) mu = 0.0;     // :42-45
    P.dt = (float)cfg->dt; P.dx = (float)dx; P.inv_dx = (float)cfg->n_grid; P.p_mass = (float)p_mass; P.mu = (float)mu; P.lam = (float)lam;
    P.cs = (float)(-cfg->dt * p_vol * 4 * (double)cfg->n_grid * (double)cfg->n_grid);
    P.gx = (float)cfg->gravity[0]; P.gy = (float)cfg->gravity[1]; P.gz = (float)cfg->gravity[2];
    P.sticky = cfg->ground_friction >= 10.0;
    P.vm = 0; P.vm_c = 0.f;
    P.material = cfg->material_model; P.ptype = cfg->ptype; P.ctype = cfg->collision_type; P.substeps = cfg->substeps; P.n_control = cfg->n_control; P.np = 0;
    s->dense = (cfg->flags & SMX_FLAG_DENSE_GRID) || (cfg->flags & SMX_FLAG_NO_SORT) || cfg->sort_every <= 0;
    P.Gb = P.ng * P.ng * P.ng; P.nb3 = P.nb * P.nb * P.nb;
    P.dbg = getenv("SMX_DBG") ? atoi(getenv("SMX_DBG")) : 0;
    s->G = (size_t)B * P.Gb;
    s->B = B;
    s->frame_floats = 24 * P.stride;
    int T = cfg->max_steps;
    s->slot_of.resize(T); s->order_of.assign(T, -1); s->trans_from.assign(T, -1); s->age_of.assign(T, 0);
    for (int f = 0; f < T; f++) s->slot_of[f] = f;
    s->spare_slot = T;
    size_t pool_bytes = (size_t)(T + 1) * s->frame_floats * sizeof(float);
    if (cudaMalloc(&s->pool, pool_bytes) != cudaSuccess) { cudaGetLastError(); return fail(SMX_ERR_NOMEM, "smx_create: cannot allocate %.1f MB of particle checkpoints", pool_bytes / 1e6); }
    s->ckpt_order.assign(T, -1); s->ckpt_contact.assign(T, 0); s->svd_order.assign(T, -1); s->near_order.assign(T, -1);
    if (cfg->material_model == 0 && cfg->ptype != 2 && !(cfg->flags & SMX_FLAG_NO_SVD_REC)) {
        // optional: without it (flag, or not enough memory) the adjoint recomputes the SVD
        size_t svd_bytes = (size_t)T * SMX_RPLANES * P.stride * sizeof(float4), free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (svd_bytes > free_b / 2 || cudaMalloc(&s->svd_pool, svd_bytes) != cudaSuccess) { cudaGetLastError(); s->svd_pool = nullptr; }
    }
    s->ckpt_enabled = !(cfg->flags & SMX_FLAG_NO_GRID_CKPT);
    CK(cudaMalloc(&s->g_in, s->G * sizeof(float4))); CK(cudaMalloc(&s->g_out, s->G * sizeof(float4))); CK(cudaMalloc(&s->g_mix, s->G * sizeof(float4)));
    CK(cudaMalloc(&s->gg_out, s->G * sizeof(float4))); CK(cudaMalloc(&s->gg_out_b, s->G * sizeof(float4))); CK(cudaMalloc(&s->gg_mix, s->G * sizeof(float4)));
    CK(cudaMemsetAsync(s->gg_out_b, 0, s->G * sizeof(float4), s->stream));
    CK(cudaMemsetAsync(s->g_in, 0, s->G * sizeof(float4), s->stream)); CK(cudaMemsetAsync(s->g_out, 0, s->G * sizeof(float4), s->stream));
    CK(cudaMemsetAsync(s->g_mix, 0, s->G * sizeof(float4), s->stream)); CK(cudaMemsetAsync(s->gg_out, 0, s->G * sizeof(float4), s->stream));
    CK(cudaMemsetAsync(s->gg_mix, 0, s->G * sizeof(float4), s->stream));
    CK(cudaMalloc(&s->mixflag, 2 * (s->G / 64))); CK(cudaMemsetAsync(s->mixflag, 0, 2 * (s->G / 64), s->stream));
    CK(cudaMalloc(&s->adj_cur, s->frame_floats * sizeof(float))); CK(cudaMalloc(&s->adj_nxt, s->frame_floats * sizeof(float)));
    size_t stage = (size_t)std::max(n, 1) * 24;
    CK(cudaMalloc(&s->stage_dev, stage * sizeof(float))); CK(cudaMallocHost(&s->stage_host, stage * sizeof(float)));
    CK(cudaMalloc(&s->keys_a, std::max(n, 1) * sizeof(uint32_t))); CK(cudaMalloc(&s->keys_b, std::max(n, 1) * sizeof(uint32_t))); CK(cudaMalloc(&s->iota, std::max(n, 1) * sizeof(uint32_t)));
    CK(cub::DeviceRadixSort::SortPairs(nullptr, s->cub_bytes, s->keys_a, s->keys_b, s->iota, s->iota, std::max(n, 1), 0, 32, s->stream));
    CK(cudaMalloc(&s->cub_tmp, s->cub_bytes));
    CK(cudaMalloc(&s->counters, 4 * sizeof(unsigned long long))); CK(cudaMemsetAsync(s->counters, 0, 4 * sizeof(unsigned long long), s->stream));
    CK(cudaMalloc(&s->ckpt_need, (size_t)s->cfg.max_steps * sizeof(int))); CK(cudaMemsetAsync(s->ckpt_need, 0, (size_t)s->cfg.max_steps * sizeof(int), s->stream));
    CK(cudaMalloc(&s->prims_dev, SMX_MAXP * sizeof(PrimDev)));
    CK(cudaMalloc(&s->pstate, (size_t)B * SMX_MAXP * T * 13 * sizeof(float))); CK(cudaMemsetAsync(s->pstate, 0, (size_t)B * SMX_MAXP * T * 13 * sizeof(float), s->stream));
    CK(cudaMalloc(&s->pgrad, (size_t)B * SMX_MAXP * T * 13 * sizeof(double))); CK(cudaMemsetAsync(s->pgrad, 0, (size_t)B * SMX_MAXP * T * 13 * sizeof(double), s->stream));
    CK(cudaMalloc(&s->ext_f, (size_t)B * SMX_MAXP * 6 * sizeof(double))); CK(cudaMemsetAsync(s->ext_f, 0, (size_t)B * SMX_MAXP * 6 * sizeof(double), s->stream));
    CK(cudaMalloc(&s->ext_f_grad, (size_t)B * SMX_MAXP * 6 * sizeof(float))); CK(cudaMemsetAsync(s->ext_f_grad, 0, (size_t)B * SMX_MAXP * 6 * sizeof(float), s->stream));
    CK(cudaMalloc(&s->abuf, (size_t)B * SMX_MAXP * T * 6 * sizeof(float))); CK(cudaMemsetAsync(s->abuf, 0, (size_t)B * SMX_MAXP * T * 6 * sizeof(float), s->stream));
    CK(cudaMalloc(&s->gabuf, (size_t)B * SMX_MAXP * T * 6 * sizeof(double))); CK(cudaMemsetAsync(s->gabuf, 0, (size_t)B * SMX_MAXP * T * 6 * sizeof(double), s->stream));
    int nc = std::max(cfg->n_control, 1) * B;
    CK(cudaMalloc(&s->ctrl_id, std::max(n, 1) * sizeof(int))); CK(cudaMemsetAsync(s->ctrl_id, 0xff, std::max(n, 1) * sizeof(int), s->stream));
    CK(cudaMalloc(&s->action, nc * 3 * sizeof(float))); CK(cudaMemsetAsync(s->action, 0, nc * 3 * sizeof(float), s->stream));
    CK(cudaMalloc(&s->action_grad, nc * 3 * sizeof(double))); CK(cudaMemsetAsync(s->action_grad, 0, nc * 3 * sizeof(double), s->stream));
    CK(cudaEventCreate(&s->ev0)); CK(cudaEventCreate(&s->ev1));
    for (int i = 0; i < SMX_IO_CHUNKS; i++) CK(cudaEventCreateWithFlags(&s->ev_chunk[i], cudaEventDisableTiming));
    // the staged G2P adjoint wants 7 CTAs x 31 KB of shared memory per SM: ask for the largest carve-out (k_p2g: at its first launch)
    cudaFuncSetAttribute(k_g2p_grad<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

int smx_destroy(smx_sim* s) {
    if (!s) return SMX_OK;
    cudaSetDevice(s->cfg.device);
    cudaStreamSynchronize(s->stream);
    for (auto& o : s->orders) if (o.live) free_order(o);
    for (auto& o : s->free_orders) free_order(o);
    for (auto& kv : s->seeds) cudaFree(kv.second.dev);
    for (auto& kv : s->seed_pool) cudaFree(kv.second);
    for (auto& p : s->prims) { cudaFree(p.sdf_dev); cudaFree(p.nrm_dev); }
    for (int side = 0; side < 2; side++) if (s->peer_base[side] && s->peer_ipc[side]) cudaIpcCloseMemHandle(s->peer_base[side]);
    for (int d = 0; d < 2; d++) for (auto& kv : s->gexec[d]) if (kv.second) cudaGraphExecDestroy(kv.second);
    void* ptrs[] = {s->halo_mem, s->halo_done, s->near_pool, s->svd_pool, s->ch_target, s->ch_loss, s->cd_buf, s->ckpt, s->pool, s->g_in, s->g_out, s->g_mix, s->gg_out, s->gg_out_b, s->gg_mix, s->mixflag, s->g_lin, s->adj_cur, s->adj_nxt, s->stage_dev, s->keys_a, s->keys_b, s->iota, s->cub_tmp,
                    s->counters, s->prims_dev, s->pstate, s->pgrad, s->ext_f, s->ext_f_grad, s->abuf, s->gabuf, s->ctrl_id, s->action, s->action_grad,
                    s->rig_arena, s->rig_enable, s->rig_masks, s->ckpt_need};
    for (void* p : ptrs) cudaFree(p);
    cudaFreeHost(s->stage_host);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    for (int i = 0; i < SMX_IO_CHUNKS; i++) if (s->ev_chunk[i]) cudaEventDestroy(s->ev_chunk[i]);
    if (s->own_stream) cudaStreamDestroy(s->stream);
    delete s;
    return SMX_OK;
}

int smx_synchronize(smx_sim* s) {
    if (!s) return fail(SMX_ERR_ARG, "smx_synchronize: null simulator");
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

int smx_add_primitive(smx_sim* s, const double* sdf, const double* normal, const int32_t res[3], const double lower[3], const double upper[3],
                      double sdf_dx, double friction, double softness, int32_t enabled) {
    if (!s) return fail(SMX_ERR_ARG, "smx_add_primitive: null simulator");
    if ((int)s->prims.size() >= SMX_MAXP) return fail(SMX_ERR_RANGE, "smx_add_primitive: at most %d primitives", SMX_MAXP);
    CK(cudaSetDevice(s->cfg.device));
    HostPrim hp; memset(&hp.d, 0, sizeof hp.d);
    hp.d.friction = (float)friction; hp.d.softness = (float)softness; hp.d.enabled = enabled ? 1 : 0;
    if (sdf) {
        if (!normal || !res || !lower || !upper || sdf_dx <= 0 || res[0] < 2 || res[1] < 2 || res[2] < 2) return fail(SMX_ERR_ARG, "smx_add_primitive: incomplete table description");
        size_t R = (size_t)res[0] * res[1] * res[2];
        std::vector<float> hs(R); std::vector<float4> hn(R);
        for (size_t i = 0; i < R; i++) { hs[i] = (float)sdf[i]; hn[i] = make_float4((float)normal[3 * i], (float)normal[3 * i + 1], (float)normal[3 * i + 2], 0.f); }
        CK(cudaMalloc(&hp.sdf_dev, R * sizeof(float))); CK(cudaMalloc(&hp.nrm_dev, R * sizeof(float4)));
        CK(cudaMemcpy(hp.sdf_dev, hs.data(), R * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(hp.nrm_dev, hn.data(), R * sizeof(float4), cudaMemcpyHostToDevice));
        hp.d.sdf = hp.sdf_dev; hp.d.nrm = hp.nrm_dev; hp.d.has_table = 1;
        hp.d.r0 = res[0]; hp.d.r1 = res[1]; hp.d.r2 = res[2];
        for (int i = 0; i < 3; i++) { hp.d.lo[i] = (float)lower[i]; hp.d.hi[i] = (float)upper[i]; }
        hp.d.inv_dx = (float)(1.0 / sdf_dx);
    }
    s->prims.push_back(hp);
    s->ckpt_dirty = true;
    s->P.np = (int)s->prims.size();
    TRY(sync_prims(s));
    return (int)s->prims.size() - 1;
}
int smx_set_primitive_params(smx_sim* s, int32_t id, double friction, double softness) {
    TRY(check_prim(s, id, "smx_set_primitive_params"));
    s->prims[id].d.friction = (float)friction; s->prims[id].d.softness = (float)softness;
    return sync_prims(s);
}
int smx_set_primitive_contact(smx_sim* s, int32_t id, int32_t enabled) {
    TRY(check_prim(s, id, "smx_set_primitive_contact"));
    s->prims[id].d.enabled = enabled ? 1 : 0;
    s->ckpt_dirty = true;
    std::fill(s->near_order.begin(), s->near_order.end(), -1);
    return sync_prims(s);
}

static void reset_bookkeeping(smx_sim* s) {
    std::fill(s->order_of.begin(), s->order_of.end(), -1);
    std::fill(s->trans_from.begin(), s->trans_from.end(), -1);
    std::fill(s->age_of.begin(), s->age_of.end(), 0);
    std::fill(s->ckpt_order.begin(), s->ckpt_order.end(), -1);
    std::fill(s->svd_order.begin(), s->svd_order.end(), -1);
    std::fill(s->near_order.begin(), s->near_order.end(), -1);
    s->adj_frame = -1; s->adj_order = -1;
    s->ckpt_dirty = true;
    s->defer_save = -1;
    s->slab_checked = -1;
    if (s->ckpt_need) cudaMemsetAsync(s->ckpt_need, 0, (size_t)s->cfg.max_steps * sizeof(int), s->stream);
    gc_orders(s);                       // every ordering is unreferenced now: recycle all of them
}
int smx_reset(smx_sim* s, const double* state, int32_t ncols) {
    if (!s || !state) return fail(SMX_ERR_ARG, "smx_reset: null argument");
    if (ncols != 3 && ncols != 24) return fail(SMX_ERR_ARG, "smx_reset: state must have 3 or 24 columns, got %d", ncols);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    reset_bookkeeping(s);
    int n = s->P.n;
    Order root; s->order_of[0] = new_order_id(s, root);
    if (n > 0) {
        auto fill = [&](float* stage, size_t i0, size_t i1) {     // element ranges (a chunk may end inside a row)
            if (ncols == 24) {
                #pragma omp parallel for schedule(static) num_threads(smx_host_threads())
                for (long long i = (long long)i0; i < (long long)i1; i++) stage[i] = (float)state[i];
            } else {
                #pragma omp parallel for schedule(static) num_threads(smx_host_threads())
                for (long long i = (long long)i0; i < (long long)i1; i++) {
                    long long p = i / 24; int c = (int)(i - 24 * p);
                    stage[i] = c < 3 ? (float)state[3 * p + c] : ((c == 6 || c == 10 || c == 14) ? 1.f : 0.f);
                }
            }
        };
        TRY(h2d_pipelined(s, (size_t)n * 24, fill));
        k_upload<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->stage_dev, 24, 0, s->frame_ptr(0), nullptr, 0); CKL(s);
    }
    return resort(s, 0, false);
}
// ---- fp32 host entry points: the caller's buffers travel as they are (no f64 <-> f32 conversion pass on the host; with buffers pinned
// once through smx_host_register the copies are plain DMA).  Same semantics as the f64 calls they mirror.
int smx_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || !bytes) return fail(SMX_ERR_ARG, "smx_host_register: null buffer");
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return SMX_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SMX_ERR_CUDA, "cudaHostRegister(%llu bytes) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    return SMX_OK;
}
int smx_host_unregister(void* ptr) {
    if (!ptr) return fail(SMX_ERR_ARG, "smx_host_unregister: null buffer");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); if (e != cudaErrorHostMemoryNotRegistered) return fail(SMX_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
    return SMX_OK;
}
int smx_reset_f32(smx_sim* s, const float* state, int32_t ncols) {
    if (!s || !state) return fail(SMX_ERR_ARG, "smx_reset_f32: null argument");
    if (ncols != 3 && ncols != 24) return fail(SMX_ERR_ARG, "smx_reset_f32: state must have 3 or 24 columns, got %d", ncols);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    reset_bookkeeping(s);
    const int n = s->P.n;
    Order root; s->order_of[0] = new_order_id(s, root);
    if (n > 0) {
        CK(cudaMemcpyAsync(s->stage_dev, state, (size_t)n * ncols * sizeof(float), cudaMemcpyHostToDevice, s->stream));
        if (ncols == 3) {       // x only: v = 0, F = I, C = 0 (MPMSimulator.reset, mpm_simulator.py:478-492)
            CK(cudaMemsetAsync(s->frame_ptr(0), 0, s->frame_floats * sizeof(float), s->stream));
            k_fill_comp<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->frame_ptr(0), 6, 10, 14, 1.f); CKL(s);
        }
        k_upload<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->stage_dev, ncols, 0, s->frame_ptr(0), nullptr, 0); CKL(s);
    }
    return resort(s, 0, false);
}
static int add_seed_dev(smx_sim* s, int f, const float* src_dev, int ncols);
static int add_seed_f32(smx_sim* s, int f, const float* g, int ncols, const char* what) {
    TRY(check_frame(s, f, what));
    if (!g) return fail(SMX_ERR_ARG, "%s: null input", what);
    const int n = s->P.n;
    if (n == 0) return SMX_OK;
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));       // staging buffer reuse
    CK(cudaMemcpyAsync(s->stage_dev, g, (size_t)n * ncols * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    return add_seed_dev(s, f, s->stage_dev, ncols);
}
int smx_add_x_grad_f32(smx_sim* s, int32_t f, const float* g3) { return add_seed_f32(s, f, g3, 3, "smx_add_x_grad_f32"); }
int smx_add_state_grad_f32(smx_sim* s, int32_t f, const float* g24) { return add_seed_f32(s, f, g24, 24, "smx_add_state_grad_f32"); }
int smx_get_state_f32(smx_sim* s, int32_t f, float* out24) {
    TRY(check_frame(s, f, "smx_get_state_f32"));
    if (!out24) return fail(SMX_ERR_ARG, "smx_get_state_f32: null output");
    TRY(smx_get_state_dev(s, f, s->stage_dev));
    if (s->P.n > 0) CK(cudaMemcpyAsync(out24, s->stage_dev, (size_t)s->P.n * 24 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_get_state_grad_f32(smx_sim* s, int32_t f, float* out24) {
    TRY(check_frame(s, f, "smx_get_state_grad_f32"));
    if (!out24) return fail(SMX_ERR_ARG, "smx_get_state_grad_f32: null output");
    TRY(smx_get_state_grad_dev(s, f, s->stage_dev));
    if (s->P.n > 0) CK(cudaMemcpyAsync(out24, s->stage_dev, (size_t)s->P.n * 24 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
// smx_reset with the (n, 24) fp32 rows already on the device (particle migration between slab ranks: no host round trip)
int smx_reset_dev(smx_sim* s, const float* rows_dev) {
    if (!s || (!rows_dev && s->P.n > 0)) return fail(SMX_ERR_ARG, "smx_reset_dev: null argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));       // orderings are recycled below
    reset_bookkeeping(s);
    const int n = s->P.n;
    Order root; s->order_of[0] = new_order_id(s, root);
    if (n > 0) { k_upload<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, rows_dev, 24, 0, s->frame_ptr(0), nullptr, 0); CKL(s); }
    return resort(s, 0, false);
}
// frame f / its adjoint as (n, 24) fp32 rows in particle-id order, written to DEVICE memory on the simulator's stream (no sync)
int smx_get_state_dev(smx_sim* s, int32_t f, float* out_dev) {
    TRY(check_frame(s, f, "smx_get_state_dev"));
    if (!out_dev && s->P.n > 0) return fail(SMX_ERR_ARG, "smx_get_state_dev: null output");
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_state_dev: frame %d has not been written", f);
    CK(cudaSetDevice(s->cfg.device));
    if (s->P.n > 0) { k_download<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P.n, s->P.stride, out_dev, 24, 0, s->frame_ptr(f), s->orders[s->order_of[f]].perm); CKL(s); }
    return SMX_OK;
}
int smx_get_state_grad_dev(smx_sim* s, int32_t f, float* out_dev) {
    TRY(check_frame(s, f, "smx_get_state_grad_dev"));
    if (!out_dev && s->P.n > 0) return fail(SMX_ERR_ARG, "smx_get_state_grad_dev: null output");
    CK(cudaSetDevice(s->cfg.device));
    if (s->P.n == 0) return SMX_OK;
    if (s->adj_frame != f) {            // no backward step has produced this frame's adjoint: it is just the loss seed (if any)
        if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_state_grad_dev: frame %d has not been written", f);
        CK(cudaMemsetAsync(s->adj_nxt, 0, s->frame_floats * sizeof(float), s->stream));
        TRY(apply_seed(s, f, s->adj_nxt, s->order_of[f]));
        k_download<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P.n, s->P.stride, out_dev, 24, 0, s->adj_nxt, s->orders[s->order_of[f]].perm); CKL(s);
    } else {
        k_download<<<nblk(s->P.n, 256), 256, 0, s->stream>>>(s->P.n, s->P.stride, out_dev, 24, 0, s->adj_cur, s->orders[s->adj_order].perm); CKL(s);
    }
    return SMX_OK;
}

int smx_set_frame(smx_sim* s, int32_t f, const double* x, const double* v, const double* F, const double* C) {
    TRY(check_frame(s, f, "smx_set_frame"));
    CK(cudaSetDevice(s->cfg.device));
    TRY(ensure_order(s, f));
    s->ckpt_order[f] = -1; s->svd_order[f] = -1; s->near_order[f] = -1;
    if (x) TRY(upload_cols(s, f, x, 3, 0));
    if (v) TRY(upload_cols(s, f, v, 3, 3));
    if (F) TRY(upload_cols(s, f, F, 9, 6));
    if (C) TRY(upload_cols(s, f, C, 9, 15));
    if (x) { TRY(resort(s, f, false)); s->ckpt_dirty = true; }      // positions changed: re-bin (cuts the adjoint chain at f, as in the reference)
    return SMX_OK;
}
int smx_get_state(smx_sim* s, int32_t f, double* out24) {
    TRY(check_frame(s, f, "smx_get_state"));
    if (!out24) return fail(SMX_ERR_ARG, "smx_get_state: null output");
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_state: frame %d has not been written", f);
    CK(cudaSetDevice(s->cfg.device));
    return download_cols(s, s->frame_ptr(f), s->orders[s->order_of[f]].perm, out24, 24, 0);
}
static int get_cols(smx_sim* s, int f, double* out, int ncomp, int c0, const char* what) {
    TRY(check_frame(s, f, what));
    if (!out) return fail(SMX_ERR_ARG, "%s: null output", what);
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "%s: frame %d has not been written", what, f);
    CK(cudaSetDevice(s->cfg.device));
    return download_cols(s, s->frame_ptr(f), s->orders[s->order_of[f]].perm, out, ncomp, c0);
}
int smx_get_x(smx_sim* s, int32_t f, double* x) { return get_cols(s, f, x, 3, 0, "smx_get_x"); }
int smx_get_v(smx_sim* s, int32_t f, double* v) { return get_cols(s, f, v, 3, 3, "smx_get_v"); }
int smx_set_x(smx_sim* s, int32_t f, const double* x) { if (!x) return fail(SMX_ERR_ARG, "smx_set_x: null input"); return smx_set_frame(s, f, x, nullptr, nullptr, nullptr); }
int smx_set_v(smx_sim* s, int32_t f, const double* v) { if (!v) return fail(SMX_ERR_ARG, "smx_set_v: null input"); return smx_set_frame(s, f, nullptr, v, nullptr, nullptr); }

int smx_copy_frame(smx_sim* s, int32_t src, int32_t dst) {
    if (s) s->slab_checked = -1;
    TRY(check_frame(s, src, "smx_copy_frame")); TRY(check_frame(s, dst, "smx_copy_frame"));
    if (s->order_of[src] < 0) return fail(SMX_ERR_STATE, "smx_copy_frame: source frame %d has not been written", src);
    CK(cudaSetDevice(s->cfg.device));
    if (src != dst) {
        CK(cudaMemcpyAsync(s->frame_ptr(dst), s->frame_ptr(src), s->frame_floats * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        s->order_of[dst] = s->order_of[src]; s->age_of[dst] = s->age_of[src]; s->trans_from[dst] = -1; s->ckpt_order[dst] = -1; s->svd_order[dst] = -1; s->near_order[dst] = -1;
        int T = s->cfg.max_steps;
        for (int b = 0; b < s->B; b++)
            for (size_t ii = 0; ii < s->prims.size(); ii++) {
                size_t i = prim_slot_of(b, (int)ii);
                for (int j = 0; j < s->cfg.substeps; j++) {
                    if (src + j >= T || dst + j >= T) break;
                    CK(cudaMemcpyAsync(s->pstate + (i * T + dst + j) * 13, s->pstate + (i * T + src + j) * 13, 13 * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
                    CK(cudaMemcpyAsync(s->abuf + (i * T + dst + j) * 6, s->abuf + (i * T + src + j) * 6, 6 * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
                }
            }
    }
    return SMX_OK;
}

// ---- primitives ---------------------------------------------------------------------------------
// Batch-addressed implementations; the un-suffixed entry points act on every batch (setters / clears) or on batch 0
// (getters), which is the whole handle when n_batch == 1.
static int check_batch(smx_sim* s, int b, const char* what) {
    if (b < 0 || b >= s->B) return fail(SMX_ERR_RANGE, "%s: batch %d outside [0, %d)", what, b, s->B);
    return SMX_OK;
}
static inline size_t prim_slot(smx_sim* s, int b, int id) { return (size_t)b * SMX_MAXP + id; }

static int set_prim_state(smx_sim* s, int b0, int b1, int id, int f0, int f1, const double* s13) {
    if (!s13) return fail(SMX_ERR_ARG, "smx_set_primitive_state: null state");
    if (f0 < 0 || f1 > s->cfg.max_steps || f0 >= f1) return fail(SMX_ERR_RANGE, "smx_set_primitive_state: frame range [%d, %d) outside [0, %d)", f0, f1, s->cfg.max_steps);
    CK(cudaSetDevice(s->cfg.device));
    std::vector<float> h((size_t)(f1 - f0) * 13);
    for (int f = 0; f < f1 - f0; f++) for (int c = 0; c < 13; c++) h[(size_t)f * 13 + c] = (float)s13[c];
    for (int f = f0; f < f1; f++) { s->near_order[f] = -1; s->ckpt_order[f] = -1; }   // reach bits and the grid record (post-contact g_out) of these substeps are stale
    for (int b = b0; b < b1; b++)
        CK(cudaMemcpyAsync(s->pstate + (prim_slot(s, b, id) * s->cfg.max_steps + f0) * 13, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    // small pageable sources are staged by the runtime before the call returns: the per-env-step pose hand-over of the rigid bridge
    // (rigid_simulator.py:200-201) does not drain the stream; a bulk fill (reset of all frames) still does
    if (h.size() * sizeof(float) > 32768) CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_set_primitive_state(smx_sim* s, int32_t id, int32_t f0, int32_t f1, const double* s13) {
    TRY(check_prim(s, id, "smx_set_primitive_state"));
    return set_prim_state(s, 0, s->B, id, f0, f1, s13);
}
int smx_set_primitive_state_b(smx_sim* s, int32_t batch, int32_t id, int32_t f0, int32_t f1, const double* s13) {
    TRY(check_prim(s, id, "smx_set_primitive_state_b")); TRY(check_batch(s, batch, "smx_set_primitive_state_b"));
    return set_prim_state(s, batch, batch + 1, id, f0, f1, s13);
}
int smx_get_primitive_state_b(smx_sim* s, int32_t batch, int32_t id, int32_t f, double* out13) {
    TRY(check_prim(s, id, "smx_get_primitive_state")); TRY(check_frame(s, f, "smx_get_primitive_state")); TRY(check_batch(s, batch, "smx_get_primitive_state"));
    if (!out13) return fail(SMX_ERR_ARG, "smx_get_primitive_state: null output");
    CK(cudaSetDevice(s->cfg.device));
    float h[13];
    CK(cudaMemcpyAsync(h, s->pstate + (prim_slot(s, batch, id) * s->cfg.max_steps + f) * 13, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int c = 0; c < 13; c++) out13[c] = h[c];
    return SMX_OK;
}
int smx_get_primitive_state(smx_sim* s, int32_t id, int32_t f, double* out13) { return smx_get_primitive_state_b(s, 0, id, f, out13); }
int smx_get_primitive_state_grad_b(smx_sim* s, int32_t batch, int32_t id, int32_t f0, int32_t f1, double* out13) {
    TRY(check_prim(s, id, "smx_get_primitive_state_grad")); TRY(check_batch(s, batch, "smx_get_primitive_state_grad"));
    if (!out13) return fail(SMX_ERR_ARG, "smx_get_primitive_state_grad: null output");
    if (f0 < 0 || f1 > s->cfg.max_steps || f0 >= f1) return fail(SMX_ERR_RANGE, "smx_get_primitive_state_grad: frame range [%d, %d) outside [0, %d)", f0, f1, s->cfg.max_steps);
    CK(cudaSetDevice(s->cfg.device));
    std::vector<double> h((size_t)(f1 - f0) * 13);
    CK(cudaMemcpyAsync(h.data(), s->pgrad + (prim_slot(s, batch, id) * s->cfg.max_steps + f0) * 13, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int c = 0; c < 13; c++) out13[c] = 0;
    for (int f = 0; f < f1 - f0; f++) for (int c = 0; c < 13; c++) out13[c] += h[(size_t)f * 13 + c];
    return SMX_OK;
}
int smx_get_primitive_state_grad(smx_sim* s, int32_t id, int32_t f0, int32_t f1, double* out13) { return smx_get_primitive_state_grad_b(s, 0, id, f0, f1, out13); }
int smx_add_primitive_state_grad_b(smx_sim* s, int32_t batch, int32_t id, int32_t f, const double* g13) {
    TRY(check_prim(s, id, "smx_add_primitive_state_grad")); TRY(check_frame(s, f, "smx_add_primitive_state_grad")); TRY(check_batch(s, batch, "smx_add_primitive_state_grad"));
    if (!g13) return fail(SMX_ERR_ARG, "smx_add_primitive_state_grad: null input");
    CK(cudaSetDevice(s->cfg.device));
    double h[13];
    double* d = s->pgrad + (prim_slot(s, batch, id) * s->cfg.max_steps + f) * 13;
    CK(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int c = 0; c < 13; c++) h[c] += g13[c];
    CK(cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_add_primitive_state_grad(smx_sim* s, int32_t id, int32_t f, const double* g13) { return smx_add_primitive_state_grad_b(s, 0, id, f, g13); }
int smx_get_ext_f_b(smx_sim* s, int32_t batch, int32_t id, double* out6) {
    TRY(check_prim(s, id, "smx_get_ext_f")); TRY(check_batch(s, batch, "smx_get_ext_f"));
    if (!out6) return fail(SMX_ERR_ARG, "smx_get_ext_f: null output");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out6, s->ext_f + 6 * prim_slot(s, batch, id), 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_get_ext_f(smx_sim* s, int32_t id, double* out6) { return smx_get_ext_f_b(s, 0, id, out6); }
static int clear_ext_f(smx_sim* s, int b0, int b1, int id) {
    CK(cudaSetDevice(s->cfg.device));
    for (int b = b0; b < b1; b++) {
        CK(cudaMemsetAsync(s->ext_f + 6 * prim_slot(s, b, id), 0, 6 * sizeof(double), s->stream));
        CK(cudaMemsetAsync(s->ext_f_grad + 6 * prim_slot(s, b, id), 0, 6 * sizeof(float), s->stream));
    }
    return SMX_OK;
}
int smx_clear_ext_f(smx_sim* s, int32_t id) { TRY(check_prim(s, id, "smx_clear_ext_f")); return clear_ext_f(s, 0, s->B, id); }
int smx_clear_ext_f_b(smx_sim* s, int32_t batch, int32_t id) {
    TRY(check_prim(s, id, "smx_clear_ext_f_b")); TRY(check_batch(s, batch, "smx_clear_ext_f_b"));
    return clear_ext_f(s, batch, batch + 1, id);
}
static int set_ext_f_grad(smx_sim* s, int b0, int b1, int id, const double* g6) {
    if (!g6) return fail(SMX_ERR_ARG, "smx_set_ext_f_grad: null input");
    CK(cudaSetDevice(s->cfg.device));
    float h[6]; for (int i = 0; i < 6; i++) h[i] = (float)g6[i];
    // pageable source: the runtime stages it before the call returns, so no stream synchronisation is needed (the reference calls
    // this once per primitive and substep in the backward pass, mpm_simulator.py:344-346)
    for (int b = b0; b < b1; b++) CK(cudaMemcpyAsync(s->ext_f_grad + 6 * prim_slot(s, b, id), h, sizeof h, cudaMemcpyHostToDevice, s->stream));
    return SMX_OK;
}
int smx_set_ext_f_grad(smx_sim* s, int32_t id, const double* g6) { TRY(check_prim(s, id, "smx_set_ext_f_grad")); return set_ext_f_grad(s, 0, s->B, id, g6); }
int smx_set_ext_f_grad_b(smx_sim* s, int32_t batch, int32_t id, const double* g6) {
    TRY(check_prim(s, id, "smx_set_ext_f_grad_b")); TRY(check_batch(s, batch, "smx_set_ext_f_grad_b"));
    return set_ext_f_grad(s, batch, batch + 1, id, g6);
}
// ---- bulk coupling (batched handles): one transfer for all rollouts and primitives ---------------------------------
int smx_get_ext_f_all(smx_sim* s, double* out) {
    if (!s || !out) return fail(SMX_ERR_ARG, "smx_get_ext_f_all: null argument");
    CK(cudaSetDevice(s->cfg.device));
    int np = (int)s->prims.size();
    std::vector<double> h((size_t)s->B * SMX_MAXP * 6);
    CK(cudaMemcpyAsync(h.data(), s->ext_f, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int b = 0; b < s->B; b++) for (int i = 0; i < np; i++) for (int c = 0; c < 6; c++) out[((size_t)b * np + i) * 6 + c] = h[((size_t)b * SMX_MAXP + i) * 6 + c];
    return SMX_OK;
}
int smx_clear_ext_f_all(smx_sim* s) {
    if (!s) return fail(SMX_ERR_ARG, "smx_clear_ext_f_all: null simulator");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemsetAsync(s->ext_f, 0, (size_t)s->B * SMX_MAXP * 6 * sizeof(double), s->stream));
    CK(cudaMemsetAsync(s->ext_f_grad, 0, (size_t)s->B * SMX_MAXP * 6 * sizeof(float), s->stream));
    return SMX_OK;
}
int smx_set_ext_f_grads_all(smx_sim* s, const double* g) {
    if (!s || !g) return fail(SMX_ERR_ARG, "smx_set_ext_f_grads_all: null argument");
    CK(cudaSetDevice(s->cfg.device));
    int np = (int)s->prims.size();
    std::vector<float> h((size_t)s->B * SMX_MAXP * 6, 0.f);
    for (int b = 0; b < s->B; b++) for (int i = 0; i < np; i++) for (int c = 0; c < 6; c++) h[((size_t)b * SMX_MAXP + i) * 6 + c] = (float)g[((size_t)b * np + i) * 6 + c];
    CK(cudaMemcpyAsync(s->ext_f_grad, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_set_primitive_states_all(smx_sim* s, int32_t f0, int32_t f1, const double* st) {
    if (!s || !st) return fail(SMX_ERR_ARG, "smx_set_primitive_states_all: null argument");
    if (f0 < 0 || f1 > s->cfg.max_steps || f0 >= f1) return fail(SMX_ERR_RANGE, "smx_set_primitive_states_all: frame range [%d, %d) outside [0, %d)", f0, f1, s->cfg.max_steps);
    CK(cudaSetDevice(s->cfg.device));
    int np = (int)s->prims.size();
    if (np == 0) return SMX_OK;
    size_t cnt = (size_t)s->B * np * 13;
    if (cnt > (size_t)std::max(s->P.n, 1) * 24) return fail(SMX_ERR_ARG, "smx_set_primitive_states_all: staging buffer too small");
    CK(cudaStreamSynchronize(s->stream));
    for (size_t i = 0; i < cnt; i++) s->stage_host[i] = (float)st[i];
    CK(cudaMemcpyAsync(s->stage_dev, s->stage_host, cnt * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    for (int f = f0; f < f1; f++) s->near_order[f] = -1;
    long long nt = (long long)s->B * np * (f1 - f0);
    k_fill_prim_states<<<nblk(nt, 128), 128, 0, s->stream>>>(s->pstate, s->stage_dev, s->cfg.max_steps, np, s->B, f0, f1); CKL(s);
    return SMX_OK;
}
int smx_get_primitive_state_grads_all(smx_sim* s, int32_t f0, int32_t f1, double* out) {
    if (!s || !out) return fail(SMX_ERR_ARG, "smx_get_primitive_state_grads_all: null argument");
    if (f0 < 0 || f1 > s->cfg.max_steps || f0 >= f1) return fail(SMX_ERR_RANGE, "smx_get_primitive_state_grads_all: frame range [%d, %d) outside [0, %d)", f0, f1, s->cfg.max_steps);
    CK(cudaSetDevice(s->cfg.device));
    int np = (int)s->prims.size();
    if (np == 0) return SMX_OK;
    size_t cnt = (size_t)s->B * np * 13;
    if (cnt * 2 > (size_t)std::max(s->P.n, 1) * 24) return fail(SMX_ERR_ARG, "smx_get_primitive_state_grads_all: staging buffer too small");
    double* dev = reinterpret_cast<double*>(s->stage_dev);
    k_sum_prim_grads<<<nblk((long long)cnt, 128), 128, 0, s->stream>>>(s->pgrad, dev, s->cfg.max_steps, np, s->B, f0, f1); CKL(s);
    CK(cudaMemcpyAsync(out, dev, cnt * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

int smx_set_primitive_action(smx_sim* s, int32_t id, int32_t st, int32_t n, const double* a6) {
    TRY(check_prim(s, id, "smx_set_primitive_action"));
    if (!a6) return fail(SMX_ERR_ARG, "smx_set_primitive_action: null input");
    int T = s->cfg.max_steps;
    if (st < 0 || n < 1 || (long long)(st + 1) * n > T) return fail(SMX_ERR_RANGE, "smx_set_primitive_action: frames [%d, %d) outside [0, %d)", st * n, (st + 1) * n, T);
    CK(cudaSetDevice(s->cfg.device));
    // action_buffer[s] = a ; v[j] = a[3:6], w[j] = a[0:3] for j in [s*n, (s+1)*n)   (primitive_base.py:285-304)
    float a[6]; for (int i = 0; i < 6; i++) a[i] = (float)a6[i];
    for (int b = 0; b < s->B; b++) {        // velocity-control actions are shared by all batches
        size_t ps = prim_slot(s, b, id);
        CK(cudaMemcpyAsync(s->abuf + (ps * T + st) * 6, a, sizeof a, cudaMemcpyHostToDevice, s->stream));
        std::vector<float> h((size_t)n * 13);
        CK(cudaMemcpyAsync(h.data(), s->pstate + (ps * T + (size_t)st * n) * 13, h.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        for (int j = 0; j < n; j++) for (int k = 0; k < 3; k++) { h[(size_t)j * 13 + 7 + k] = a[3 + k]; h[(size_t)j * 13 + 10 + k] = a[k]; }
        CK(cudaMemcpyAsync(s->pstate + (ps * T + (size_t)st * n) * 13, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return SMX_OK;
}
int smx_get_primitive_action_grad(smx_sim* s, int32_t id, int32_t st, int32_t n, double* out6) {
    TRY(check_prim(s, id, "smx_get_primitive_action_grad"));
    if (!out6) return fail(SMX_ERR_ARG, "smx_get_primitive_action_grad: null output");
    int T = s->cfg.max_steps;
    if (st < 0 || n < 1 || (long long)(st + 1) * n > T) return fail(SMX_ERR_RANGE, "smx_get_primitive_action_grad: frames outside [0, %d)", T);
    CK(cudaSetDevice(s->cfg.device));
    // set_velocity_from_action_kernel.grad accumulates into action_buffer.grad[s] (primitive_base.py:298-319); the action is shared
    // by every batched rollout, so its gradient is the sum over batches (kept in batch 0's action_buffer.grad)
    std::vector<double> h((size_t)n * 13); double g[6];
    CK(cudaMemcpyAsync(g, s->gabuf + ((size_t)id * T + st) * 6, sizeof g, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int b = 0; b < s->B; b++) {
        CK(cudaMemcpyAsync(h.data(), s->pgrad + ((size_t)prim_slot(s, b, id) * T + (size_t)st * n) * 13, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        for (int j = 0; j < n; j++) for (int k = 0; k < 3; k++) { g[3 + k] += h[(size_t)j * 13 + 7 + k]; g[k] += h[(size_t)j * 13 + 10 + k]; }
    }
    CK(cudaMemcpyAsync(s->gabuf + ((size_t)id * T + st) * 6, g, sizeof g, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < 6; i++) out6[i] = g[i];
    return SMX_OK;
}
// Primitive.reset of the reference (primitive_base.py:236-246, 267-275, 321-326): zero the state series and their adjoints, the
// velocity-control action buffer and its adjoint, the wrench and its adjoint -- for every batch of the handle
int smx_reset_primitive(smx_sim* s, int32_t id) {
    TRY(check_prim(s, id, "smx_reset_primitive"));
    CK(cudaSetDevice(s->cfg.device));
    const size_t T = (size_t)s->cfg.max_steps;
    for (int b = 0; b < s->B; b++) {
        const size_t ps = prim_slot(s, b, id);
        CK(cudaMemsetAsync(s->pstate + ps * T * 13, 0, T * 13 * sizeof(float), s->stream));
        CK(cudaMemsetAsync(s->pgrad + ps * T * 13, 0, T * 13 * sizeof(double), s->stream));
        CK(cudaMemsetAsync(s->abuf + ps * T * 6, 0, T * 6 * sizeof(float), s->stream));
        CK(cudaMemsetAsync(s->gabuf + ps * T * 6, 0, T * 6 * sizeof(double), s->stream));
    }
    std::fill(s->near_order.begin(), s->near_order.end(), -1);
    std::fill(s->ckpt_order.begin(), s->ckpt_order.end(), -1);
    return clear_ext_f(s, 0, s->B, id);
}

// ---- material variants ---------------------------------------------------------------------------
int smx_set_plasticity(smx_sim* s, int32_t mode, double yield_stress) {
    if (!s) return fail(SMX_ERR_ARG, "smx_set_plasticity: null simulator");
    if (mode != 0 && mode != 1) return fail(SMX_ERR_ARG, "smx_set_plasticity: mode %d (0 sigma clip, 1 von Mises)", mode);
    if (mode == 1 && !(s->cfg.material_model == 0 && s->cfg.ptype == 0))
        return fail(SMX_ERR_STATE, "smx_set_plasticity: the von Mises return mapping belongs to the co-rotated plastic material (material_model 0, ptype 0)");
    if (mode == 1 && !(yield_stress >= 0.0)) return fail(SMX_ERR_ARG, "smx_set_plasticity: yield_stress %g", yield_stress);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    s->P.vm = mode;
    s->P.vm_c = mode ? (float)(yield_stress / (2.0 * (s->cfg.E / (2.0 * (1.0 + s->cfg.nu))))) : 0.f;      // yield_stress / (2 mu), mu as in mpm_simulator.py:41
    return SMX_OK;
}

// ---- control ------------------------------------------------------------------------------------
int smx_set_action(smx_sim* s, const double* action) {
    if (!s || !action) return fail(SMX_ERR_ARG, "smx_set_action: null argument");
    if (s->cfg.n_control <= 0) return fail(SMX_ERR_STATE, "smx_set_action: simulator was created with n_control == 0");
    CK(cudaSetDevice(s->cfg.device));
    std::vector<float> h((size_t)s->B * s->cfg.n_control * 3);       // (n_batch * n_control, 3)
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)action[i];
    CK(cudaMemcpyAsync(s->action, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemsetAsync(s->action_grad, 0, h.size() * sizeof(double), s->stream));    // set_action_kernel zeroes action.grad (:584-586)
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_set_control_idx(smx_sim* s, const int32_t* idx) {
    if (!s || !idx) return fail(SMX_ERR_ARG, "smx_set_control_idx: null argument");
    CK(cudaSetDevice(s->cfg.device));
    if (s->P.n > 0) CK(cudaMemcpyAsync(s->ctrl_id, idx, (size_t)s->P.n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->ctrl_version++;
    return SMX_OK;
}
int smx_get_action_grad(smx_sim* s, double* out) {
    if (!s || !out) return fail(SMX_ERR_ARG, "smx_get_action_grad: null argument");
    if (s->cfg.n_control <= 0) return fail(SMX_ERR_STATE, "smx_get_action_grad: simulator was created with n_control == 0");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out, s->action_grad, (size_t)s->B * s->cfg.n_control * 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

// ---- device-resident rigid coupling (fixed / prismatic / revolute / free joints): RigidSimulator.step / step_grad on the simulator's stream ------------
int smx_rigid_linear_create(smx_sim* s, const smx_rigid_linear* d) {
    if (!s || !d) return fail(SMX_ERR_ARG, "smx_rigid_linear_create: null argument");
    const int np = (int)s->prims.size(), sd = d->state_dim, ad = d->action_dim, K = d->max_env_steps, B = s->B;
    if (np == 0) return fail(SMX_ERR_STATE, "smx_rigid_linear_create: the simulator has no primitives");
    if (sd < 1 || sd > SMX_RIG_MAXS || ad < 0 || ad > SMX_RIG_MAXA || K < 1) return fail(SMX_ERR_RANGE, "smx_rigid_linear_create: state_dim in [1, %d], action_dim in [0, %d], max_env_steps >= 1", SMX_RIG_MAXS, SMX_RIG_MAXA);
    if (!d->As || !d->Aw || !d->c || !d->body || !d->joint || !d->enable || !d->init_state || (ad > 0 && !d->Aa)) return fail(SMX_ERR_ARG, "smx_rigid_linear_create: null matrix");
    if (s->cfg.rigid_velocity_control) return fail(SMX_ERR_STATE, "smx_rigid_linear_create: not available with rigid velocity control");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(s->rig_arena); cudaFree(s->rig_enable); cudaFree(s->rig_masks); s->rig_arena = nullptr; s->rig_enable = nullptr; s->rig_masks = nullptr; s->rig_on = false;
    for (int i = 0; i < np; i++) {
        const int jt = d->joint[2 * i], o = d->joint[2 * i + 1], nd = rig_ndof(jt);
        if (jt < 0 || jt > 3 || o < 0 || 2 * (o + nd) > sd || (sd & 1)) return fail(SMX_ERR_RANGE, "smx_rigid_linear_create: joint %d of primitive %d (dof offset %d) does not fit state_dim %d", jt, i, o, sd);
    }
    const size_t nAs = (size_t)sd * sd, nAa = (size_t)ad * sd, nAw = (size_t)6 * np * sd, nc = sd, nbody = (size_t)np * 10;
    const size_t nst = (size_t)(K + 1) * B * sd, nact = (size_t)K * B * std::max(ad, 1), nsg = (size_t)B * sd;
    const size_t consts = nAs + nAa + nAw + nc + nbody, total = consts + nst + 2 * nact + nsg;
    CK(cudaMalloc(&s->rig_arena, total * sizeof(double)));
    CK(cudaMalloc(&s->rig_enable, 3 * np * sizeof(int)));
    CK(cudaMalloc(&s->rig_masks, (size_t)K * B * np));
    std::vector<double> h(consts);
    size_t o = 0;
    auto put = [&](const double* src, size_t n) { if (n) memcpy(h.data() + o, src, n * sizeof(double)); size_t at = o; o += n; return at; };
    const size_t oAs = put(d->As, nAs), oAa = put(d->Aa, nAa), oAw = put(d->Aw, nAw), oc = put(d->c, nc), obody = put(d->body, nbody);
    CK(cudaMemcpy(s->rig_arena, h.data(), consts * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemset(s->rig_arena + consts, 0, (total - consts) * sizeof(double)));
    std::vector<int> en(3 * np);          // enable (np) then joint (np, 2)
    for (int i = 0; i < np; i++) { en[i] = d->enable[i] ? 1 : 0; en[np + 2 * i] = d->joint[2 * i]; en[np + 2 * i + 1] = d->joint[2 * i + 1]; }
    CK(cudaMemcpy(s->rig_enable, en.data(), en.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(s->rig_masks, 0, (size_t)K * B * np));
    RigidLin& R = s->rig;
    R.sd = sd; R.ad = ad; R.np = np; R.B = B; R.S = std::max(s->cfg.substeps, 1); R.T = s->cfg.max_steps; R.K = K; R.fp32 = d->fp32_bridge ? 1 : 0;
    R.scale = d->ext_grad_scale;
    double* a = s->rig_arena;
    R.As = a + oAs; R.Aa = a + oAa; R.Aw = a + oAw; R.c = a + oc; R.body = a + obody; R.enable = s->rig_enable; R.joint = s->rig_enable + np;
    R.states = a + consts; R.actions = R.states + nst; R.action_grad = R.actions + nact; R.state_grad = R.action_grad + nact; R.masks = s->rig_masks;
    s->rig_init.assign(d->init_state, d->init_state + sd);
    s->rig_on = true;
    return SMX_OK;
}
static int rig_check(smx_sim* s, int k, const char* what) {
    if (!s) return fail(SMX_ERR_ARG, "%s: null simulator", what);
    if (!s->rig_on) return fail(SMX_ERR_STATE, "%s: smx_rigid_linear_create has not been called", what);
    if (k < 0 || k >= s->rig.K) return fail(SMX_ERR_RANGE, "%s: env step %d outside [0, %d)", what, k, s->rig.K);
    return SMX_OK;
}
// RigidSimulator.reset (rigid_simulator.py:360-369): state 0 = init_state for every rollout, poses of frames [0, substeps), wrench and adjoints cleared
int smx_rigid_linear_reset(smx_sim* s) {
    TRY(rig_check(s, 0, "smx_rigid_linear_reset"));
    CK(cudaSetDevice(s->cfg.device));
    RigidLin& R = s->rig;
    std::vector<double> h((size_t)R.B * R.sd);
    for (int b = 0; b < R.B; b++) for (int i = 0; i < R.sd; i++) h[(size_t)b * R.sd + i] = s->rig_init[i];
    CK(cudaMemcpyAsync(R.states, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemsetAsync(R.state_grad, 0, (size_t)R.B * R.sd * sizeof(double), s->stream));
    CK(cudaMemsetAsync(R.action_grad, 0, (size_t)R.K * R.B * std::max(R.ad, 1) * sizeof(double), s->stream));
    const int f1 = std::min(R.S, s->cfg.max_steps);
    for (int f = 0; f < f1; f++) s->near_order[f] = -1;
    k_rigid_linear_step<<<R.B, 128, 0, s->stream>>>(R, 0, 0, 0, f1, s->ext_f, s->ext_f_grad, s->pstate); CKLN(s, "rigid");
    return SMX_OK;
}
int smx_rigid_linear_set_actions(smx_sim* s, int32_t k, const double* actions) {
    TRY(rig_check(s, k, "smx_rigid_linear_set_actions"));
    if (s->rig.ad == 0) return SMX_OK;
    if (!actions) return fail(SMX_ERR_ARG, "smx_rigid_linear_set_actions: null input");
    CK(cudaSetDevice(s->cfg.device));
    // pageable source: staged by the runtime before the call returns, no stream synchronisation
    CK(cudaMemcpyAsync(s->rig.actions + (size_t)k * s->B * s->rig.ad, actions, (size_t)s->B * s->rig.ad * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return SMX_OK;
}
int smx_rigid_linear_step(smx_sim* s, int32_t k) {
    TRY(rig_check(s, k, "smx_rigid_linear_step"));
    CK(cudaSetDevice(s->cfg.device));
    RigidLin& R = s->rig;
    const int f0 = std::min((k + 1) * R.S, s->cfg.max_steps), f1 = std::min((k + 2) * R.S, s->cfg.max_steps);
    for (int f = f0; f < f1; f++) s->near_order[f] = -1;
    k_rigid_linear_step<<<R.B, 128, 0, s->stream>>>(R, k, 1, f0, f1, s->ext_f, s->ext_f_grad, s->pstate); CKLN(s, "rigid");
    return SMX_OK;
}
int smx_rigid_linear_step_grad(smx_sim* s, int32_t k) {
    TRY(rig_check(s, k, "smx_rigid_linear_step_grad"));
    CK(cudaSetDevice(s->cfg.device));
    RigidLin& R = s->rig;
    const int f0 = std::min((k + 1) * R.S, s->cfg.max_steps), f1 = std::min((k + 2) * R.S, s->cfg.max_steps);
    k_rigid_linear_step_grad<<<R.B, 128, 0, s->stream>>>(R, k, 0, f0, f1, s->pgrad, s->ext_f_grad); CKLN(s, "rigid_grad");
    return SMX_OK;
}
int smx_rigid_linear_finish(smx_sim* s) {
    TRY(rig_check(s, 0, "smx_rigid_linear_finish"));
    CK(cudaSetDevice(s->cfg.device));
    RigidLin& R = s->rig;
    k_rigid_linear_step_grad<<<R.B, 128, 0, s->stream>>>(R, 0, 1, 0, std::min(R.S, s->cfg.max_steps), s->pgrad, s->ext_f_grad); CKLN(s, "rigid_grad");
    return SMX_OK;
}
int smx_rigid_linear_get_states(smx_sim* s, int32_t k, double* out) {
    if (!s || !s->rig_on) return fail(SMX_ERR_STATE, "smx_rigid_linear_get_states: smx_rigid_linear_create has not been called");
    if (!out) return fail(SMX_ERR_ARG, "smx_rigid_linear_get_states: null output");
    if (k < 0 || k > s->rig.K) return fail(SMX_ERR_RANGE, "smx_rigid_linear_get_states: state index %d outside [0, %d]", k, s->rig.K);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out, s->rig.states + (size_t)k * s->B * s->rig.sd, (size_t)s->B * s->rig.sd * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_rigid_linear_get_action_grads(smx_sim* s, int32_t k0, int32_t k1, double* out) {
    if (!s || !s->rig_on) return fail(SMX_ERR_STATE, "smx_rigid_linear_get_action_grads: smx_rigid_linear_create has not been called");
    if (!out) return fail(SMX_ERR_ARG, "smx_rigid_linear_get_action_grads: null output");
    if (k0 < 0 || k1 > s->rig.K || k0 >= k1) return fail(SMX_ERR_RANGE, "smx_rigid_linear_get_action_grads: env steps [%d, %d) outside [0, %d)", k0, k1, s->rig.K);
    if (s->rig.ad == 0) return SMX_OK;
    CK(cudaSetDevice(s->cfg.device));
    const size_t row = (size_t)s->B * s->rig.ad;
    CK(cudaMemcpyAsync(out, s->rig.action_grad + k0 * row, (k1 - k0) * row * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_rigid_linear_get_state_grad(smx_sim* s, double* out) {
    if (!s || !s->rig_on) return fail(SMX_ERR_STATE, "smx_rigid_linear_get_state_grad: smx_rigid_linear_create has not been called");
    if (!out) return fail(SMX_ERR_ARG, "smx_rigid_linear_get_state_grad: null output");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out, s->rig.state_grad, (size_t)s->B * s->rig.sd * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

static int halo_exchange(smx_sim* s, int which, int f);     // slab halo sum over peer memory (defined with the slab API below)

// ---- the hot path -------------------------------------------------------------------------------
// substep = begin (clear + P2G) ; [slab mode: halo exchange of g_in by the caller] ; end (grid update, contact, G2P)
int smx_substep_begin(smx_sim* s, int32_t f) {
    TRY(check_frame(s, f, "smx_substep"));
    if (f + 1 >= s->cfg.max_steps) return fail(SMX_ERR_RANGE, "smx_substep: substep %d would write frame %d >= max_steps %d", f, f + 1, s->cfg.max_steps);
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_substep: frame %d has not been written (call smx_reset / smx_set_frame first)", f);
    CK(cudaSetDevice(s->cfg.device));
    s->order_of[f + 1] = s->order_of[f]; s->trans_from[f + 1] = -1; s->age_of[f + 1] = s->age_of[f] + 1;
    s->ckpt_order[f] = -1; s->ckpt_order[f + 1] = -1; s->svd_order[f] = -1; s->near_order[f] = -1; s->svd_order[f + 1] = -1;
    if (s->ckpt_dirty) TRY(ensure_ckpt(s, f));
    return forward_p2g(s, f, true, true);
}
// grid update + contact scatter; separate from `end` so that a slab caller can exchange the contact scatter in between
int smx_substep_mid(smx_sim* s, int32_t f) {
    TRY(check_frame(s, f, "smx_substep"));
    if (f + 1 >= s->cfg.max_steps || s->order_of[f] < 0 || s->order_of[f + 1] != s->order_of[f]) return fail(SMX_ERR_STATE, "smx_substep_mid: smx_substep_begin(%d) has not been called", f);
    CK(cudaSetDevice(s->cfg.device));
    TRY(forward_grid(s, f, true, true));
    s->mid_done = f;
    return SMX_OK;
}
int smx_substep_end(smx_sim* s, int32_t f) {
    TRY(check_frame(s, f, "smx_substep"));
    if (f + 1 >= s->cfg.max_steps || s->order_of[f] < 0 || s->order_of[f + 1] != s->order_of[f]) return fail(SMX_ERR_STATE, "smx_substep_end: smx_substep_begin(%d) has not been called", f);
    CK(cudaSetDevice(s->cfg.device));
    if (s->mid_done != f) TRY(smx_substep_mid(s, f));
    s->mid_done = -1;
    TRY(forward_grid_save_contact(s, f));
    s->last_fwd = f;
    if (s->P.n > 0) { launch_pdl(s, k_g2p, nblk(s->P.n, SMX_TPB), SMX_TPB, 0, s->P, s->frame_ptr(f), s->frame_ptr(f + 1), s->g_out, s->pf_g2p); CKLN(s, "k_g2p"); }
    if (s->cfg.sort_every > 0 && s->age_of[f + 1] >= s->cfg.sort_every && !(s->cfg.flags & SMX_FLAG_NO_SORT)) TRY(resort(s, f + 1, true));
    return SMX_OK;
}
int smx_substep(smx_sim* s, int32_t f) {
    TRY(smx_substep_begin(s, f));
    if (s->peers_on()) {                // slab rank with connected neighbours: the halo sums run here, on the stream
        TRY(halo_exchange(s, 0, f));
        if (s->has_contact()) { TRY(smx_substep_mid(s, f)); TRY(halo_exchange(s, 1, f)); }
    }
    return smx_substep_end(s, f);
}

// adjoint substep = begin (adjoint set-up, grid restore, G2P adjoint, contact adjoint) ; [slab mode: halo exchange of
// gg_out by the caller] ; end (grid adjoint, P2G adjoint)
int smx_substep_grad_begin(smx_sim* s, int32_t f) {
    TRY(check_frame(s, f, "smx_substep_grad"));
    if (f + 1 >= s->cfg.max_steps) return fail(SMX_ERR_RANGE, "smx_substep_grad: substep %d outside the stored range", f);
    if (s->order_of[f] < 0 || s->order_of[f + 1] < 0) return fail(SMX_ERR_STATE, "smx_substep_grad: substep %d has not been run forward", f);
    CK(cudaSetDevice(s->cfg.device));
    const Params& P = s->P;
    int o = s->order_of[f];
    if (s->adj_frame != f + 1 && s->ckpt) {     // start of a backward pass: did every grid record fit? (one sync)
        unsigned long long ov = 0;
        CK(cudaStreamSynchronize(s->stream));
        CK(cudaMemcpy(&ov, s->counters + 2, sizeof ov, cudaMemcpyDeviceToHost));
        if (ov != s->ckpt_overflow_seen) {
            // some record did not fit (the active region grew): the substeps whose record is incomplete recompute their grid in the
            // adjoint, the others keep using theirs; the arena is sized for the largest need at the next reset
            s->ckpt_overflow_seen = ov;
            std::vector<int> need(s->cfg.max_steps, 0);
            CK(cudaMemcpy(need.data(), s->ckpt_need, need.size() * sizeof(int), cudaMemcpyDeviceToHost));
            for (int g = 0; g < s->cfg.max_steps; g++) {
                if (need[g] > s->ckpt_cap) s->ckpt_order[g] = -1;
                s->ckpt_need_max = std::max(s->ckpt_need_max, need[g]);
            }
        }
    }
    if (s->adj_frame != f + 1) {        // start of a backward pass: the adjoint of frame f+1 is its loss seed
        CK(cudaMemsetAsync(s->adj_cur, 0, s->frame_floats * sizeof(float), s->stream));
        s->adj_frame = f + 1; s->adj_order = s->order_of[f + 1];
        TRY(apply_seed(s, f + 1, s->adj_cur, s->adj_order));
    }
    if (s->adj_order != o) {            // frame f+1 was re-sorted after substep f produced it: carry the adjoint back
        if (s->adj_order != s->order_of[f + 1] || s->trans_from[f + 1] != o || !s->orders[s->adj_order].idx)
            return fail(SMX_ERR_STATE, "smx_substep_grad: adjoint of frame %d is not connected to substep %d (frame was overwritten?)", f + 1, f);
        if (P.n > 0) { k_scatter_frame<<<nblk(P.n, 256), 256, 0, s->stream>>>(P.n, P.stride, s->adj_cur, s->adj_nxt, s->orders[s->adj_order].idx); CKLN(s, "resort_adjoint"); }
        std::swap(s->adj_cur, s->adj_nxt);
        s->adj_order = o;
    }
    Order& ord = s->orders[o];
    bool contact = s->has_contact();
    PrimSet ps = s->primset();
    float4* gg = s->gg_of(f);
    const bool have_rec = s->ckpt && s->ckpt_order[f] == ord.uid && (bool)s->ckpt_contact[f] == contact;
    if (have_rec && s->bwd_prepared == f && s->bwd_prepared_uid == ord.uid) {
        // nothing to do: k_grid_grad of substep f+1 restored g_out (/ g_mix) and cleared the adjoint grids; g_in is read from the record
    } else if (have_rec) {
        // restore g_in / g_out (/ g_mix) and zero the adjoint grids of the same blocks in one launch
        launch_pdl(s, k_ckpt_copy, grid_blocks_launch(s), 256, 0, s->dense ? nullptr : ord.blocks, ord.nblocks, s->B * s->P.nb3, s->ckpt_cap, s->ckpt + (size_t)f * s->ckpt_rec,
                                                                 s->g_in, s->g_out, contact ? s->g_mix : nullptr, 1, s->counters, gg, contact ? s->gg_mix : nullptr, nullptr);
        CKLN(s, "ckpt_restore");
    } else {
        if (s->slab) return fail(SMX_ERR_STATE, "smx_substep_grad: slab decomposition needs the grid checkpoint of substep %d (run it forward in this handle; do not set SMX_FLAG_NO_GRID_CKPT)", f);
        TRY(forward_to_grid(s, f, false, false));
        TRY(clear_grids(s, ord, gg, contact ? s->gg_mix : nullptr, nullptr));
    }
    s->g_in_clean_uid = -1;             // g_in now holds substep f's values (or stale ones when the record is used directly)
    s->bwd_prepared = -1;
    const float* fin = s->frame_ptr(f);
    if (P.n > 0) {
#ifdef SMX_G2PG_F4
        const size_t g2pg_smem = (size_t)(SMX_TPB_G2PG / 32) * sizeof(WarpStage);
        if (s->smem_optin.insert((const void*)k_g2p_grad<true>).second) cudaFuncSetAttribute(k_g2p_grad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g2pg_smem);
#else
        const size_t g2pg_smem = 0;
#endif
        if (s->cfg.flags & SMX_FLAG_DIRECT_RED) launch_pdl(s, k_g2p_grad<false>, nblk(P.n, SMX_TPB_G2PG), SMX_TPB_G2PG, 0, P, fin, s->adj_cur, s->adj_nxt, s->g_out, gg, s->pf_g2pg);
        else launch_pdl(s, k_g2p_grad<true>, nblk(P.n, SMX_TPB_G2PG), SMX_TPB_G2PG, g2pg_smem, P, fin, s->adj_cur, s->adj_nxt, s->g_out, gg, s->pf_g2pg);
        CKLN(s, "k_g2p_grad");
    }
    s->grad_pending = f;
    (void)ps;
    return SMX_OK;
}
// adjoint of the forecast contact: gathers gg_out (complete after the slab caller's halo sum), scatters into gg_mix
int smx_substep_grad_mid(smx_sim* s, int32_t f) {
    TRY(check_frame(s, f, "smx_substep_grad"));
    if (s->grad_pending != f) return fail(SMX_ERR_STATE, "smx_substep_grad_mid: smx_substep_grad_begin(%d) has not been called", f);
    CK(cudaSetDevice(s->cfg.device));
    const Params& P = s->P;
    if (s->has_contact() && P.n > 0) {
        PrimSet ps = s->primset();
        float life = 1.0f / (float)(P.substeps - f % P.substeps);
        const uint32_t* near = (s->near_pool && s->near_order[f] == s->orders[s->order_of[f]].uid) ? s->near_of(f) : nullptr;
        if (near) {
            const int nwords = (int)s->near_words();
            const int grid = s->sm_count * 2;
            launch_pdl(s, k_contact_grad_sparse, grid, SMX_TPB, 0, P, ps, f, life, s->frame_ptr(f), s->adj_nxt, s->g_mix, s->gg_of(f), s->gg_mix, near, nwords, s->mixflag_of(f));
        } else {
            launch_pdl(s, k_contact_grad, nblk(P.n, SMX_TPB), SMX_TPB, 0, P, ps, f, life, s->frame_ptr(f), s->adj_nxt, s->g_mix, s->gg_of(f), s->gg_mix, s->mixflag_of(f));
        }
        CKLN(s, "k_contact_grad");
    }
    s->grad_mid_done = f;
    return SMX_OK;
}
// P2G adjoint of substep f and G2P adjoint of substep f-1 may share a launch when f-1 runs in the same ordering, both have their grid
// record (so that k_grid_grad of f also restores g_out of f-1 and clears its adjoint grid) and no loss seed enters at frame f
static bool can_fuse_bwd(smx_sim* s, int f) {
    if (!s->bwd_fusion || f < 1 || s->slab_legacy() || (s->cfg.flags & (SMX_FLAG_NO_FUSION | SMX_FLAG_DIRECT_RED)) || s->P.n <= 0) return false;
    const int o = s->order_of[f];
    if (o < 0 || s->order_of[f - 1] != o || s->trans_from[f] >= 0) return false;
    const Order& ord = s->orders[o];
    const bool contact = s->has_contact();
    auto have = [&](int g) { return s->ckpt && s->ckpt_order[g] == ord.uid && (bool)s->ckpt_contact[g] == contact; };
    return have(f) && have(f - 1) && s->seeds.find(f) == s->seeds.end();
}
static int grad_end(smx_sim* s, int f, bool fuse);
int smx_substep_grad_end(smx_sim* s, int32_t f) { return grad_end(s, f, false); }
static int grad_end(smx_sim* s, int f, bool fuse) {
    TRY(check_frame(s, f, "smx_substep_grad"));
    if (s->grad_pending != f) return fail(SMX_ERR_STATE, "smx_substep_grad_end: smx_substep_grad_begin(%d) has not been called", f);
    if (s->grad_mid_done != f) TRY(smx_substep_grad_mid(s, f));
    s->grad_pending = -1; s->grad_mid_done = -1;
    CK(cudaSetDevice(s->cfg.device));
    const Params& P = s->P;
    int o = s->order_of[f];
    Order& ord = s->orders[o];
    bool contact = s->has_contact();
    PrimSet ps = s->primset();
    const float* fin = s->frame_ptr(f);
    float4* gg = s->gg_of(f);
    {
        // g_in comes straight from the grid checkpoint when there is one; and when substep f-1 shares the ordering and has a
        // record too, this launch also prepares its adjoint (restore of g_out / g_mix, clear of the other adjoint grid)
        const bool have_rec = s->ckpt && s->ckpt_order[f] == ord.uid && (bool)s->ckpt_contact[f] == contact;
        const bool prep = have_rec && !s->slab_legacy() && !(s->cfg.flags & SMX_FLAG_NO_FUSION) && f > 0 && s->order_of[f - 1] == o && s->ckpt_order[f - 1] == ord.uid &&
                          (bool)s->ckpt_contact[f - 1] == contact && s->trans_from[f] < 0;
        const float4* rec_in = have_rec ? s->ckpt + (size_t)f * s->ckpt_rec : nullptr;
        const float4* rec_prev = prep ? s->ckpt + (size_t)(f - 1) * s->ckpt_rec : nullptr;
        { const bool saved_pdl = s->pdl; s->pdl = s->pdl && s->pdl_grid;
        if (P.ctype == 0) launch_pdl(s, k_grid_grad<true>, grid_blocks_launch(s), 256, 0, P, ps, f, s->dense ? nullptr : ord.blocks, ord.nblocks, s->g_in, gg, contact ? s->gg_mix : nullptr,
                                                                 rec_in, rec_prev, s->ckpt_cap, contact ? 1 : 0, s->g_out, s->g_mix, s->gg_of(f - 1), contact ? s->mixflag_of(f) : nullptr, contact ? s->mixflag_of(f, true) : nullptr);
        else launch_pdl(s, k_grid_grad<false>, grid_blocks_launch(s), 256, 0, P, ps, f, s->dense ? nullptr : ord.blocks, ord.nblocks, s->g_in, gg, contact ? s->gg_mix : nullptr,
                                                                 rec_in, rec_prev, s->ckpt_cap, contact ? 1 : 0, s->g_out, s->g_mix, s->gg_of(f - 1), contact ? s->mixflag_of(f) : nullptr, contact ? s->mixflag_of(f, true) : nullptr);
        s->pdl = saved_pdl; }
        CKLN(s, "k_grid_grad");
        s->bwd_prepared = prep ? f - 1 : -1; s->bwd_prepared_uid = ord.uid;
        if (fuse && !prep) return fail(SMX_ERR_STATE, "internal: fused backward launch without a prepared substep %d", f - 1);
    }
    if (s->cfg.rigid_velocity_control && !s->prims.empty()) {
        k_forward_kinematics_grad<<<nblk((long long)s->prims.size() * s->B, 64), 64, 0, s->stream>>>(s->pstate, s->pgrad, s->cfg.max_steps, (int)s->prims.size(), s->B, f, P.dt); CKL(s);
    }
    const int* cslot = nullptr;
    TRY(ctrl_slots(s, o, &cslot));
    if (P.n > 0) {
        const bool use_rec = s->svd_pool && s->svd_order[f] == ord.uid;
        const float4* rec = use_rec ? s->svd_rec(f) : nullptr;
        const bool extra = P.ctype == 1 || P.n_control > 0;
        const bool tiled = !(s->cfg.flags & SMX_FLAG_NO_TMA);
        TRY(dispatch_mat(P, [&](auto mat) {
            constexpr int M = decltype(mat)::value;
            auto go = [&](auto rec_c, auto extra_c) {
                constexpr bool R = decltype(rec_c)::value, E = decltype(extra_c)::value;
                if (fuse) {
                    // + G2P adjoint of substep f-1: g_out of f-1 and its cleared adjoint grid were put in place by k_grid_grad above
                    if (getenv("SMX_CARVEOUT_FB") && s->smem_optin.insert((const void*)k_p2g_grad_g2p_grad<M, R, E>).second)      // A/B experiments: L1 / shared split
                        cudaFuncSetAttribute(k_p2g_grad_g2p_grad<M, R, E>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("SMX_CARVEOUT_FB")));
                    launch_pdl(s, k_p2g_grad_g2p_grad<M, R, E>, nblk(P.n, SMX_TPB_FB), SMX_TPB_FB, 0, P, ps, f, fin, s->adj_cur, s->adj_nxt, gg, cslot, s->action, s->action_grad, rec,
                               (const float*)s->frame_ptr(f - 1), (const float4*)s->g_out, s->gg_of(f - 1), s->pf_g);
                } else if (tiled) {
                    // persistent CTAs, double-buffered TMA staging of the streaming planes
                    const int ntiles = nblk(P.n, SMX_P2GG_TPB), grid = std::min(ntiles, s->sm_count * SMX_P2GG_TILED_MINB);
                    const size_t smem = (size_t)2 * SMX_P2GG_NPL(M, R) * SMX_P2GG_TPB * sizeof(float4);
                    // the opt-in above 48 KB is per device: remembered per handle (a process may hold handles on several GPUs)
                    if (s->smem_optin.insert((const void*)k_p2g_grad_tiled<M, R, E>).second) cudaFuncSetAttribute(k_p2g_grad_tiled<M, R, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    launch_pdl(s, k_p2g_grad_tiled<M, R, E>, grid, SMX_P2GG_TPB, smem, P, ps, f, fin, s->adj_cur, s->adj_nxt, gg, cslot, s->action, s->action_grad, rec, ntiles);
                } else {
                    launch_pdl(s, k_p2g_grad<M, R, E>, nblk(P.n, SMX_TPB), SMX_TPB, 0, P, ps, f, fin, s->adj_cur, s->adj_nxt, gg, cslot, s->action, s->action_grad, rec, s->pf_g);
                }
            };
            if (use_rec) { if (extra) go(std::true_type(), std::true_type()); else go(std::true_type(), std::false_type()); }
            else { if (extra) go(std::false_type(), std::true_type()); else go(std::false_type(), std::false_type()); }
            CKLN(s, fuse ? "k_p2g_grad+g2p_grad" : "k_p2g_grad"); return (int)SMX_OK;
        }));
    }
    std::swap(s->adj_cur, s->adj_nxt);
    s->adj_frame = f; s->adj_order = o;
    s->adj_partial = fuse;
    if (fuse) {         // the G2P adjoint of substep f-1 is done: what smx_substep_grad_begin(f - 1) would leave behind
        s->g_in_clean_uid = -1; s->bwd_prepared = -1;
        s->grad_pending = f - 1;
        return SMX_OK;
    }
    TRY(apply_seed(s, f, s->adj_cur, o));
    return SMX_OK;
}
int smx_substep_grad(smx_sim* s, int32_t f) {
    TRY(smx_substep_grad_begin(s, f));
    if (s->peers_on()) {
        TRY(halo_exchange(s, 3, f));
        if (s->has_contact()) { TRY(smx_substep_grad_mid(s, f)); TRY(halo_exchange(s, 4, f)); }
    }
    return smx_substep_grad_end(s, f);
}

// ---- spatial slab decomposition ----------------------------------------------------------------------------------------
int smx_set_slab(smx_sim* s, int32_t xb_lo, int32_t xb_hi, int32_t has_lo_neighbour, int32_t has_hi_neighbour) {
    if (!s) return fail(SMX_ERR_ARG, "smx_set_slab: null simulator");
    int nb = s->P.nb;
    if (xb_lo < 0 || xb_hi > nb || xb_lo >= xb_hi) return fail(SMX_ERR_RANGE, "smx_set_slab: block columns [%d, %d) outside [0, %d)", xb_lo, xb_hi, nb);
    if ((has_lo_neighbour && xb_lo < 1) || (has_hi_neighbour && xb_hi > nb - 1)) return fail(SMX_ERR_RANGE, "smx_set_slab: no room for a halo column");
    if (s->B != 1) return fail(SMX_ERR_STATE, "smx_set_slab: not available for batched handles");
    if (s->dense) return fail(SMX_ERR_STATE, "smx_set_slab: needs active-block lists (sort_every > 0, no SMX_FLAG_DENSE_GRID / SMX_FLAG_NO_SORT)");
    s->slab = true; s->slab_lo = xb_lo; s->slab_hi = xb_hi; s->halo_lo = has_lo_neighbour != 0; s->halo_hi = has_hi_neighbour != 0;
    return SMX_OK;
}
// ---- halo exchange over peer memory ---------------------------------------------------------------------------------------
// Layout of a rank's halo allocation (identical on every rank: it only depends on nb): [flag lo | flag hi] (128 B each), then per
// (side, slot) the block stamps (H u32, padded to 256 B), then per (side, slot) the receive slot (H x 64 float4).  H = 2 nb^2.
struct HaloLayout {
    size_t H, stamp_bytes, total;
    explicit HaloLayout(int nb) { H = (size_t)2 * nb * nb; stamp_bytes = (H * 4 + 255) / 256 * 256; total = 256 + 4 * stamp_bytes + 4 * H * 64 * sizeof(float4); }
    size_t flag(int side) const { return (size_t)side * 128; }
    size_t stamp(int side, int slot) const { return 256 + (size_t)(side * 2 + slot) * stamp_bytes; }
    size_t rx(int side, int slot) const { return 256 + 4 * stamp_bytes + (size_t)(side * 2 + slot) * H * 64 * sizeof(float4); }
};
static int halo_alloc(smx_sim* s) {
    if (s->halo_mem) return SMX_OK;
    if (!s->slab) return fail(SMX_ERR_STATE, "smx_slab_halo: call smx_set_slab first");
    HaloLayout L(s->P.nb);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMalloc(&s->halo_mem, L.total));
    CK(cudaMemset(s->halo_mem, 0, L.total));
    CK(cudaMalloc(&s->halo_done, 2 * sizeof(unsigned)));
    CK(cudaMemset(s->halo_done, 0, 2 * sizeof(unsigned)));
    s->halo_bytes = L.total;
    return SMX_OK;
}
// this rank's halo allocation: device pointer (ranks emulated in one process connect through it) and a 64-byte IPC handle (one process per GPU)
int smx_slab_halo_export(smx_sim* s, void** base, void* ipc_handle64) {
    if (!s) return fail(SMX_ERR_ARG, "smx_slab_halo_export: null simulator");
    TRY(halo_alloc(s));
    if (base) *base = s->halo_mem;
    if (ipc_handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, s->halo_mem));
        memcpy(ipc_handle64, &h, 64);
    }
    return SMX_OK;
}
// side 0: the neighbour that owns the columns below slab_lo, side 1: the one above slab_hi.  Exactly one of (base, ipc_handle64).
int smx_slab_halo_connect(smx_sim* s, int32_t side, void* base, const void* ipc_handle64) {
    if (!s || (side != 0 && side != 1)) return fail(SMX_ERR_ARG, "smx_slab_halo_connect: bad argument");
    if (!(side == 0 ? s->halo_lo : s->halo_hi)) return fail(SMX_ERR_STATE, "smx_slab_halo_connect: this slab has no neighbour on side %d", side);
    if ((base != nullptr) == (ipc_handle64 != nullptr)) return fail(SMX_ERR_ARG, "smx_slab_halo_connect: pass a device pointer or an IPC handle");
    TRY(halo_alloc(s));
    CK(cudaSetDevice(s->cfg.device));
    if (s->peer_base[side] && s->peer_ipc[side]) cudaIpcCloseMemHandle(s->peer_base[side]);
    if (ipc_handle64) {
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handle64, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_base[side] = (unsigned char*)p; s->peer_ipc[side] = true;
    } else { s->peer_base[side] = (unsigned char*)base; s->peer_ipc[side] = false; }
    return SMX_OK;
}
// halo arrays: 0 g_in (P2G sums), 1 g_out += neighbour's (g_out - g_mix) (contact scatter), 3 adjoint grid of substep f, 4 gg_mix
static float4* halo_array(smx_sim* s, int which, int f) { return which == 0 ? s->g_in : which == 1 ? s->g_out : which == 3 ? s->gg_of(f) : s->gg_mix; }
// push == true: the neighbours' slots / stamps / flags (what this rank writes); else this rank's own (what it waits for and adds)
static HaloSides halo_sides(smx_sim* s, bool push) {
    const int nb = s->P.nb;
    HaloLayout L(nb);
    HaloSides hs;
    for (int side = 0; side < 2; side++) {
        hs.on[side] = (side == 0 ? s->halo_lo : s->halo_hi) ? 1 : 0;
        const int b = side == 0 ? s->slab_lo : s->slab_hi;
        hs.node0[side] = hs.on[side] ? (size_t)(b - 1) * nb * nb * 64 : 0;
        hs.seq1[side] = s->halo_seq[side] + 1;
        const int slot = (int)(s->halo_seq[side] & 1u);
        unsigned char* base = push ? s->peer_base[side] : s->halo_mem;
        const int at = push ? 1 - side : side;          // my lo neighbour receives on ITS hi side
        hs.rx[side] = hs.on[side] ? (float4*)(base + L.rx(at, slot)) : nullptr;
        hs.stamp[side] = hs.on[side] ? (uint32_t*)(base + L.stamp(at, slot)) : nullptr;
        hs.flag[side] = hs.on[side] ? (unsigned*)(base + L.flag(at)) : nullptr;
        hs.done[side] = s->halo_done + side;
    }
    return hs;
}
static int halo_push(smx_sim* s, int which, int f) {
    HaloLayout L(s->P.nb);
    const HaloSides hs = halo_sides(s, true);
    const dim3 grid(std::min((int)((L.H + 7) / 8), s->sm_count * 2), 2);
    if (which == 1) k_halo_push<1><<<grid, 256, 0, s->stream>>>(halo_array(s, which, f), s->g_mix, (int)L.H, hs);
    else k_halo_push<0><<<grid, 256, 0, s->stream>>>(halo_array(s, which, f), nullptr, (int)L.H, hs);
    CKLN(s, "halo_push");
    return SMX_OK;
}
static int halo_add(smx_sim* s, int which, int f) {
    HaloLayout L(s->P.nb);
    static const long long timeout_ns = getenv("SMX_HALO_TIMEOUT_MS") ? atoll(getenv("SMX_HALO_TIMEOUT_MS")) * 1000000ll : 2000000000ll;
    const HaloSides hs = halo_sides(s, false);
    k_halo_wait<<<1, 32, 0, s->stream>>>(hs, s->counters, timeout_ns); CKLN(s, "halo_wait");
    const dim3 grid(std::min((int)((L.H + 7) / 8), s->sm_count * 2), 2);
    k_halo_add<<<grid, 256, 0, s->stream>>>(halo_array(s, which, f), (int)L.H, hs); CKLN(s, "halo_add");
    for (int side = 0; side < 2; side++) if (hs.on[side]) s->halo_seq[side]++;
    s->halo_exchanges++;
    return SMX_OK;
}
// the two halves of one exchange as separate calls (ranks emulated in one process on one stream must all push before anyone waits)
int smx_slab_halo_push(smx_sim* s, int32_t which, int32_t f) {
    if (!s || !s->peers_on()) return fail(SMX_ERR_STATE, "smx_slab_halo_push: neighbours are not connected (smx_slab_halo_connect)");
    if (which != 0 && which != 1 && which != 3 && which != 4) return fail(SMX_ERR_RANGE, "smx_slab_halo_push: array in {0, 1, 3, 4}");
    CK(cudaSetDevice(s->cfg.device));
    return halo_push(s, which, f);
}
int smx_slab_halo_add(smx_sim* s, int32_t which, int32_t f) {
    if (!s || !s->peers_on()) return fail(SMX_ERR_STATE, "smx_slab_halo_add: neighbours are not connected (smx_slab_halo_connect)");
    if (which != 0 && which != 1 && which != 3 && which != 4) return fail(SMX_ERR_RANGE, "smx_slab_halo_add: array in {0, 1, 3, 4}");
    CK(cudaSetDevice(s->cfg.device));
    return halo_add(s, which, f);
}
static int halo_exchange(smx_sim* s, int which, int f) {
    TRY(halo_push(s, which, f));
    return halo_add(s, which, f);
}
// out[0]: exchanges that gave up waiting for the neighbour (must be 0), out[1]: exchanges done, out[2]: bytes of the halo allocation
int smx_slab_halo_status(smx_sim* s, int64_t out[3]) {
    if (!s || !out) return fail(SMX_ERR_ARG, "smx_slab_halo_status: null argument");
    CK(cudaSetDevice(s->cfg.device));
    unsigned long long t = 0;
    CK(cudaMemcpyAsync(&t, s->counters + 3, sizeof t, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    out[0] = (int64_t)t; out[1] = s->halo_exchanges; out[2] = (int64_t)s->halo_bytes;
    return SMX_OK;
}

// device pointer and size of a grid array: 0 g_in, 1 g_out, 2 g_mix, 3 gg_out, 4 gg_mix (float4 per node, block-major:
// x-block column c is the contiguous element range [c * nb^2 * 64, (c+1) * nb^2 * 64))
int smx_grid_dev(smx_sim* s, int32_t which, void** ptr, int64_t* n_float4) {
    if (!s || !ptr || !n_float4) return fail(SMX_ERR_ARG, "smx_grid_dev: null argument");
    float4* t[5] = {s->g_in, s->g_out, s->g_mix, s->gg_out, s->gg_mix};
    if (which < 0 || which > 4) return fail(SMX_ERR_RANGE, "smx_grid_dev: which in [0, 5)");
    *ptr = t[which]; *n_float4 = (int64_t)s->G;
    return SMX_OK;
}
int smx_stream(smx_sim* s, void** stream) {
    if (!s || !stream) return fail(SMX_ERR_ARG, "smx_stream: null argument");
    *stream = (void*)s->stream;
    return SMX_OK;
}

// `count` consecutive substeps.  Between two substeps that share an ordering the G2P of the first is fused into the P2G
// of the second ("G2P2G": one launch, frame f's x, v, C go from registers straight into the next scatter); results are
// identical to calling smx_substep in a loop.  Disabled by SMX_FLAG_NO_FUSION, with velocity control and for a slab rank whose halos are
// exchanged by the caller; a slab rank with connected neighbours (smx_slab_halo_connect) keeps the fusion: its halo sums run in here.
int smx_step(smx_sim* s, int32_t s0, int32_t count) {
    if (!s) return fail(SMX_ERR_ARG, "smx_step: null simulator");
    bool fuse_ok = !(s->cfg.flags & SMX_FLAG_NO_FUSION) && !s->slab_legacy() && !s->cfg.rigid_velocity_control && s->P.n > 0;
    if (!fuse_ok || count < 2) {
        for (int i = 0; i < count; i++) TRY(smx_substep(s, s0 + i));
        return SMX_OK;
    }
    bool pending_g2p = false;           // G2P of substep f-1 still to be done (fused into the P2G of f)
    for (int i = 0; i < count; i++) {
        int f = s0 + i;
        if (!pending_g2p) TRY(smx_substep_begin(s, f));
        else {
            // the part of smx_substep_begin after the checks, with the previous G2P folded into the P2G launch
            TRY(check_frame(s, f, "smx_step"));
            if (f + 1 >= s->cfg.max_steps) return fail(SMX_ERR_RANGE, "smx_step: substep %d would write frame %d >= max_steps %d", f, f + 1, s->cfg.max_steps);
            CK(cudaSetDevice(s->cfg.device));
            s->order_of[f + 1] = s->order_of[f]; s->trans_from[f + 1] = -1; s->age_of[f + 1] = s->age_of[f] + 1;
            s->ckpt_order[f] = -1; s->ckpt_order[f + 1] = -1; s->svd_order[f] = -1; s->near_order[f] = -1; s->svd_order[f + 1] = -1;
            TRY(forward_p2g(s, f, true, true, true));
            pending_g2p = false;
        }
        if (s->peers_on()) TRY(halo_exchange(s, 0, f));
        TRY(smx_substep_mid(s, f));
        if (s->peers_on() && s->has_contact()) TRY(halo_exchange(s, 1, f));
        s->mid_done = -1;
        s->last_fwd = f;
        bool resort_next = s->cfg.sort_every > 0 && s->age_of[f + 1] >= s->cfg.sort_every && !(s->cfg.flags & SMX_FLAG_NO_SORT);
        bool last = (i == count - 1);
        const bool fuse_next = !last && !resort_next && f + 2 < s->cfg.max_steps && !s->ckpt_dirty;
        // with contact the record of g_out / g_mix can only be taken after the contact scatter: when the next substep follows in this
        // call with the same ordering, its k_grid_op saves them just before overwriting them (no copy launch); otherwise copy now
        if (fuse_next && s->has_contact() && s->ckpt && s->ckpt_narr == 3) s->defer_save = f;
        else TRY(forward_grid_save_contact(s, f));
        if (fuse_next) { pending_g2p = true; continue; }
        launch_pdl(s, k_g2p, nblk(s->P.n, SMX_TPB), SMX_TPB, 0, s->P, s->frame_ptr(f), s->frame_ptr(f + 1), s->g_out, s->pf_g2p); CKLN(s, "k_g2p");
        if (resort_next) TRY(resort(s, f + 1, true));
    }
    return SMX_OK;
}
// `count` consecutive adjoint substeps s1-1 ... s1-count.  Between two substeps that share an ordering the G2P adjoint of the earlier one
// is fused into the P2G adjoint launch of the later one (the mirror image of G2P2G in smx_step); the last substep of the call runs the
// plain kernels, so the adjoint of frame s1-count is complete when the call returns.  Results equal smx_substep_grad in a loop.
int smx_step_grad(smx_sim* s, int32_t s1, int32_t count) {
    if (!s) return fail(SMX_ERR_ARG, "smx_step_grad: null simulator");
    bool g2p_done = false;              // the G2P adjoint of the substep at hand was part of the previous launch
    for (int i = 1; i <= count; i++) {
        const int f = s1 - i;
        if (!g2p_done) TRY(smx_substep_grad_begin(s, f));
        if (s->peers_on()) TRY(halo_exchange(s, 3, f));
        TRY(smx_substep_grad_mid(s, f));
        if (s->peers_on() && s->has_contact()) TRY(halo_exchange(s, 4, f));
        const bool fuse = i < count && can_fuse_bwd(s, f);
        const int rc = grad_end(s, f, fuse);
        if (rc != SMX_OK) { s->adj_frame = -1; s->adj_partial = false; s->grad_pending = -1; return rc; }
        g2p_done = fuse;
    }
    return SMX_OK;
}

// ---- CUDA-graph replay for small scenes ------------------------------------------------------------------------------------------
// At ~10 k particles every kernel of a substep is a single partial wave and an env step is bound by the latency of ~16 dependent launches.
// smx_step_graph / smx_step_grad_graph capture the launch sequence of the call into a graph (stream capture: the host code below runs
// as always, its launches are recorded instead of submitted), update the handle's executable graph with it -- same topology from call to
// call, only the frame-dependent kernel arguments change, so cudaGraphExecUpdate patches the nodes in place -- and launch that ONE graph.
// Calls that would allocate, synchronise or re-sort inside (the first substeps after a reset, a re-sort falling into the call, the
// first adjoint call of a pass, slab ranks, profiling) go down the ordinary path: results are identical either way.
int smx_step_graph(smx_sim* s, int32_t s0, int32_t count) {
    if (!s) return fail(SMX_ERR_ARG, "smx_step_graph: null simulator");
    bool ok = graph_ok_common(s) && count >= 1 && s0 >= 0 && s0 + count + 1 < s->cfg.max_steps && s->order_of[s0] >= 0;
    // no re-sort may fall into the call (it recycles orderings and sizes CUB's workspace): the substeps stay in the ordering of frame s0
    ok = ok && s->cfg.sort_every > 0 && s->age_of[s0] + count < s->cfg.sort_every;
    if (ok && s->g_in_clean_uid != s->orders[s->order_of[s0]].uid) ok = true;   // the clear launch is captured like any other
    if (!ok) { s->graph_fallbacks++; return smx_step(s, s0, count); }
    return graph_run(s, 0, [&]() { return smx_step(s, s0, count); });
}
int smx_step_grad_graph(smx_sim* s, int32_t s1, int32_t count) {
    if (!s) return fail(SMX_ERR_ARG, "smx_step_grad_graph: null simulator");
    bool ok = graph_ok_common(s) && count >= 1 && s1 - count >= 0 && s1 < s->cfg.max_steps;
    // not the first call of a backward pass (it looks at the checkpoint counters on the host), one ordering throughout, every grid record there
    ok = ok && s->adj_frame == s1 && s->order_of[s1] >= 0 && s->adj_order == s->order_of[s1] && s->trans_from[s1] < 0;
    if (ok) {
        const int o = s->order_of[s1];
        const int uid = s->orders[o].uid;
        const bool contact = s->has_contact();
        for (int f = s1 - 1; f >= s1 - count && ok; f--)
            ok = s->order_of[f] == o && s->trans_from[f + 1] < 0 && s->ckpt_order[f] == uid && (bool)s->ckpt_contact[f] == contact;
    }
    if (!ok) { s->graph_fallbacks++; return smx_step_grad(s, s1, count); }
    return graph_run(s, 1, [&]() { return smx_step_grad(s, s1, count); });
}
// out[0] calls replayed as one graph launch, out[1] calls that took the ordinary path, out[2] re-instantiations (topology changed)
int smx_graph_status(smx_sim* s, int64_t out[3]) {
    if (!s || !out) return fail(SMX_ERR_ARG, "smx_graph_status: null argument");
    out[0] = s->graph_launches; out[1] = s->graph_fallbacks; out[2] = s->graph_reinstantiations;
    return SMX_OK;
}

// ---- adjoint seeds / read-out ---------------------------------------------------------------------
static int add_seed_dev(smx_sim* s, int f, const float* src_dev, int ncols);
static int add_seed(smx_sim* s, int f, const double* g, int ncols) {
    int n = s->P.n;
    if (n == 0) return SMX_OK;
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    size_t cnt = (size_t)n * ncols;
    auto fill = [&](float* dst, size_t i0, size_t i1) {
        #pragma omp parallel for schedule(static) num_threads(smx_host_threads())
        for (long long i = (long long)i0; i < (long long)i1; i++) dst[i] = (float)g[i];
    };
        TRY(h2d_pipelined(s, cnt, fill));
    return add_seed_dev(s, f, s->stage_dev, ncols);
}
// accumulate (n, ncols) fp32 rows that are already on the device into the loss seed of frame f
static int add_seed_dev(smx_sim* s, int f, const float* src_dev, int ncols) {
    int n = s->P.n;
    if (n == 0) return SMX_OK;
    size_t cnt = (size_t)n * ncols;
    auto it = s->seeds.find(f);
    if (it != s->seeds.end() && it->second.ncols < ncols) {
        // widen an x-only seed to the full 24 columns
        float* d = nullptr;
        TRY(seed_alloc(s, (size_t)n * 24 * sizeof(float), &d));
        CK(cudaMemsetAsync(d, 0, (size_t)n * 24 * sizeof(float), s->stream));
        CK(cudaMemcpy2DAsync(d, 24 * sizeof(float), it->second.dev, 3 * sizeof(float), 3 * sizeof(float), n, cudaMemcpyDeviceToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        seed_release(s, it->second.dev, (size_t)n * 3 * sizeof(float));
        it->second.dev = d; it->second.ncols = 24;
    }
    if (it == s->seeds.end()) {
        smx_sim::Seed sd; sd.ncols = ncols;
        TRY(seed_alloc(s, cnt * sizeof(float), &sd.dev));
        CK(cudaMemcpyAsync(sd.dev, src_dev, cnt * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        s->seeds[f] = sd;
    } else if (it->second.ncols == ncols) {
        k_axpy<<<nblk((long long)cnt, 256), 256, 0, s->stream>>>((long long)cnt, it->second.dev, src_dev); CKL(s);
    } else {    // stored 24 columns, adding 3
        k_add_cols<<<nblk(n, 256), 256, 0, s->stream>>>(n, it->second.dev, 24, src_dev, 3); CKL(s);
    }
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_add_state_grad(smx_sim* s, int32_t f, const double* g24) {
    TRY(check_frame(s, f, "smx_add_state_grad"));
    if (!g24) return fail(SMX_ERR_ARG, "smx_add_state_grad: null input");
    return add_seed(s, f, g24, 24);
}
int smx_add_state_grad_dev(smx_sim* s, int32_t f, const float* g24_dev) {
    TRY(check_frame(s, f, "smx_add_state_grad_dev"));
    if (!g24_dev && s->P.n > 0) return fail(SMX_ERR_ARG, "smx_add_state_grad_dev: null input");
    CK(cudaSetDevice(s->cfg.device));
    return add_seed_dev(s, f, g24_dev, 24);
}
int smx_add_x_grad(smx_sim* s, int32_t f, const double* g3) {
    TRY(check_frame(s, f, "smx_add_x_grad"));
    if (!g3) return fail(SMX_ERR_ARG, "smx_add_x_grad: null input");
    return add_seed(s, f, g3, 3);
}
// Chamfer loss between the particles of frame f and a target cloud, and its seed on x.grad[f] (loss_grip.py:45-68,117-140)
int smx_set_chamfer_target(smx_sim* s, const double* target, int32_t m) {
    if (!s || !target || m < 1) return fail(SMX_ERR_ARG, "smx_set_chamfer_target: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    std::vector<float> h((size_t)m * 3);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)target[i];
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(s->ch_target); s->ch_target = nullptr;
    CK(cudaMalloc(&s->ch_target, h.size() * sizeof(float)));
    CK(cudaMemcpy(s->ch_target, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    s->ch_m = m;
    if (!s->ch_loss) CK(cudaMalloc(&s->ch_loss, sizeof(double)));
    return SMX_OK;
}
int smx_chamfer_loss(smx_sim* s, int32_t f, double weight, double* loss_out) {
    TRY(check_frame(s, f, "smx_chamfer_loss"));
    if (!loss_out) return fail(SMX_ERR_ARG, "smx_chamfer_loss: null output");
    if (!s->ch_target) return fail(SMX_ERR_STATE, "smx_chamfer_loss: call smx_set_chamfer_target first");
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_chamfer_loss: frame %d has not been written", f);
    CK(cudaSetDevice(s->cfg.device));
    int n = s->P.n;
    *loss_out = 0.0;
    if (n == 0) return SMX_OK;
    auto it = s->seeds.find(f);
    if (it == s->seeds.end()) {         // create an x-only seed buffer for this frame
        smx_sim::Seed sd; sd.ncols = 3;
        TRY(seed_alloc(s, (size_t)n * 3 * sizeof(float), &sd.dev));
        CK(cudaMemsetAsync(sd.dev, 0, (size_t)n * 3 * sizeof(float), s->stream));
        s->seeds[f] = sd;
        it = s->seeds.find(f);
    }
    CK(cudaMemsetAsync(s->ch_loss, 0, sizeof(double), s->stream));
    const uint32_t* perm = s->orders[s->order_of[f]].perm;
    const int per_cta = 128 / SMX_CH_SPLIT;        // queries per CTA (SMX_CH_SPLIT lanes share one query)
    k_chamfer<<<dim3(nblk(s->P.npb, per_cta), s->B), 128, 0, s->stream>>>(s->P, s->frame_ptr(f), perm, s->ch_target, s->ch_m, (float)weight, it->second.dev, it->second.ncols, s->ch_loss, 0); CKL(s);
    k_chamfer<<<dim3(nblk(s->ch_m, per_cta), s->B), 128, 0, s->stream>>>(s->P, s->frame_ptr(f), perm, s->ch_target, s->ch_m, (float)weight, it->second.dev, it->second.ncols, s->ch_loss, 1); CKL(s);
    CK(cudaMemcpyAsync(loss_out, s->ch_loss, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

// Contact-distance term of DoorLoss / TransportLoss (loss_door.py:46-56, loss_transport.py:54-70) on the device: value per rollout
// into loss_out[n_batch]; gradient into the loss seed of frame f (minimising particle) and into the primitive's position adjoint.
int smx_contact_distance_loss(smx_sim* s, int32_t f, int32_t prim_id, int32_t n_groups, double weight, double* loss_out) {
    TRY(check_frame(s, f, "smx_contact_distance_loss"));
    TRY(check_prim(s, prim_id, "smx_contact_distance_loss"));
    if (!loss_out) return fail(SMX_ERR_ARG, "smx_contact_distance_loss: null output");
    if (n_groups < 1 || n_groups > 64) return fail(SMX_ERR_RANGE, "smx_contact_distance_loss: n_groups in [1, 64]");
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_contact_distance_loss: frame %d has not been written", f);
    CK(cudaSetDevice(s->cfg.device));
    const int n = s->P.n, B = s->B, npc = s->P.npb / n_groups, nt = B * n_groups;
    for (int b = 0; b < B; b++) loss_out[b] = 0.0;
    if (n == 0 || npc == 0) return SMX_OK;
    if (s->cd_cap < nt) {
        CK(cudaStreamSynchronize(s->stream));
        cudaFree(s->cd_buf); s->cd_buf = nullptr;
        CK(cudaMalloc(&s->cd_buf, (size_t)nt * 4 * sizeof(unsigned long long)));      // [2 nt] best, [nt] slots (u32 in u64 cells), [nt >= B] loss
        s->cd_cap = nt;
    }
    auto it = s->seeds.find(f);
    if (it == s->seeds.end()) {         // create an x-only seed buffer for this frame
        smx_sim::Seed sd; sd.ncols = 3;
        TRY(seed_alloc(s, (size_t)n * 3 * sizeof(float), &sd.dev));
        CK(cudaMemsetAsync(sd.dev, 0, (size_t)n * 3 * sizeof(float), s->stream));
        s->seeds[f] = sd;
        it = s->seeds.find(f);
    }
    unsigned long long* best = s->cd_buf;
    uint32_t* slots = reinterpret_cast<uint32_t*>(s->cd_buf + 2 * (size_t)nt);
    double* loss = reinterpret_cast<double*>(s->cd_buf + 3 * (size_t)nt);
    CK(cudaMemsetAsync(best, 0xff, (size_t)nt * 2 * sizeof(unsigned long long), s->stream));
    CK(cudaMemsetAsync(s->cd_buf + 2 * (size_t)nt, 0, (size_t)nt * 2 * sizeof(unsigned long long), s->stream));
    const uint32_t* perm = s->orders[s->order_of[f]].perm;
    const float* fr = s->frame_ptr(f);
    const int T = s->cfg.max_steps;
    k_contact_dist_min<<<nblk(n, 256), 256, 0, s->stream>>>(s->P, fr, perm, s->pstate, T, prim_id, f, n_groups, npc, best, 0); CKL(s);
    k_contact_dist_min<<<nblk(n, 256), 256, 0, s->stream>>>(s->P, fr, perm, s->pstate, T, prim_id, f, n_groups, npc, best, 1); CKL(s);
    k_contact_dist_slot<<<nblk(n, 256), 256, 0, s->stream>>>(s->P, perm, n_groups, npc, best, slots); CKL(s);
    k_contact_dist_finish<<<nblk(nt, 64), 64, 0, s->stream>>>(s->P, fr, s->pstate, s->pgrad, T, prim_id, f, n_groups, npc, best, weight, it->second.dev, it->second.ncols, loss, slots); CKL(s);
    CK(cudaMemcpyAsync(loss_out, loss, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}

int smx_get_state_grad(smx_sim* s, int32_t f, double* out24) {
    TRY(check_frame(s, f, "smx_get_state_grad"));
    if (!out24) return fail(SMX_ERR_ARG, "smx_get_state_grad: null output");
    CK(cudaSetDevice(s->cfg.device));
    if (s->adj_frame != f) {
        // no backward step has produced this frame's adjoint: it is just the loss seed (if any)
        if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_state_grad: frame %d has not been written", f);
        CK(cudaMemsetAsync(s->adj_nxt, 0, s->frame_floats * sizeof(float), s->stream));
        TRY(apply_seed(s, f, s->adj_nxt, s->order_of[f]));
        return download_cols(s, s->adj_nxt, s->orders[s->order_of[f]].perm, out24, 24, 0);
    }
    return download_cols(s, s->adj_cur, s->orders[s->adj_order].perm, out24, 24, 0);
}
int smx_get_grad(smx_sim* s, int32_t f, double* xg, double* vg) {
    TRY(check_frame(s, f, "smx_get_grad"));
    if (!xg || !vg) return fail(SMX_ERR_ARG, "smx_get_grad: null output");
    CK(cudaSetDevice(s->cfg.device));
    // MPMSimulator.get_grad (mpm_simulator.py:561-574): only the x and v columns travel (2 x n x 3 fp32 over PCIe)
    const float* adj = s->adj_cur;
    int order = s->adj_order;
    if (s->adj_frame != f) {
        if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_grad: frame %d has not been written", f);
        CK(cudaMemsetAsync(s->adj_nxt, 0, s->frame_floats * sizeof(float), s->stream));
        TRY(apply_seed(s, f, s->adj_nxt, s->order_of[f]));
        adj = s->adj_nxt; order = s->order_of[f];
    }
    // x and v are adjacent get_state columns (0..5): one gather kernel, one copy, then the split into the two host arrays
    const int n = s->P.n;
    if (n == 0) return SMX_OK;
    launch_plain_download(s, adj, s->orders[order].perm, 6, 0);
    CKL(s);
    CK(cudaMemcpyAsync(s->stage_host, s->stage_dev, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const float* src = s->stage_host;
    #pragma omp parallel for schedule(static) num_threads(smx_host_threads())
    for (long long p = 0; p < (long long)n; p++)
        for (int c = 0; c < 3; c++) { xg[3 * p + c] = (double)src[6 * p + c]; vg[3 * p + c] = (double)src[6 * p + 3 + c]; }
    return SMX_OK;
}
// MPMSimulator.get_grad(f) into two caller (n, 3) fp32 arrays: two gathers, two plain copies, no conversion pass
int smx_get_grad_f32(smx_sim* s, int32_t f, float* xg, float* vg) {
    TRY(check_frame(s, f, "smx_get_grad_f32"));
    if (!xg || !vg) return fail(SMX_ERR_ARG, "smx_get_grad_f32: null output");
    CK(cudaSetDevice(s->cfg.device));
    const float* adj = s->adj_cur;
    int order = s->adj_order;
    if (s->adj_frame != f) {
        if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_grad_f32: frame %d has not been written", f);
        CK(cudaMemsetAsync(s->adj_nxt, 0, s->frame_floats * sizeof(float), s->stream));
        TRY(apply_seed(s, f, s->adj_nxt, s->order_of[f]));
        adj = s->adj_nxt; order = s->order_of[f];
    }
    const int n = s->P.n;
    if (n == 0) return SMX_OK;
    const uint32_t* perm = s->orders[order].perm;
    k_download<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->stage_dev, 3, 0, adj, perm); CKL(s);
    k_download<<<nblk(n, 256), 256, 0, s->stream>>>(n, s->P.stride, s->stage_dev + (size_t)3 * n, 3, 3, adj, perm); CKL(s);
    CK(cudaMemcpyAsync(xg, s->stage_dev, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(vg, s->stage_dev + (size_t)3 * n, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_clear_grads(smx_sim* s) {
    if (!s) return fail(SMX_ERR_ARG, "smx_clear_grads: null simulator");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    for (auto& kv : s->seeds) seed_release(s, kv.second.dev, (size_t)s->P.n * kv.second.ncols * sizeof(float));
    s->seeds.clear();
    s->adj_frame = -1; s->adj_order = -1;
    int T = s->cfg.max_steps;
    const int B = s->B;
    CK(cudaMemsetAsync(s->pgrad, 0, (size_t)B * SMX_MAXP * T * 13 * sizeof(double), s->stream));
    CK(cudaMemsetAsync(s->gabuf, 0, (size_t)B * SMX_MAXP * T * 6 * sizeof(double), s->stream));
    CK(cudaMemsetAsync(s->ext_f_grad, 0, (size_t)B * SMX_MAXP * 6 * sizeof(float), s->stream));
    CK(cudaMemsetAsync(s->action_grad, 0, (size_t)B * std::max(s->cfg.n_control, 1) * 3 * sizeof(double), s->stream));
    return SMX_OK;
}

// ---- introspection --------------------------------------------------------------------------------
int smx_get_sort_keys(smx_sim* s, int32_t f, uint32_t* keys) {
    TRY(check_frame(s, f, "smx_get_sort_keys"));
    if (!keys) return fail(SMX_ERR_ARG, "smx_get_sort_keys: null output");
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_sort_keys: frame %d has not been written", f);
    CK(cudaSetDevice(s->cfg.device));
    int n = s->P.n;
    if (n == 0) return SMX_OK;
    k_keys<<<nblk(n, 256), 256, 0, s->stream>>>(s->P, s->frame_ptr(f), s->keys_a, nullptr, nullptr); CKL(s);
    CK(cudaMemcpyAsync(keys, s->keys_a, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_get_permutation(smx_sim* s, int32_t f, uint32_t* perm) {
    TRY(check_frame(s, f, "smx_get_permutation"));
    if (!perm) return fail(SMX_ERR_ARG, "smx_get_permutation: null output");
    if (s->order_of[f] < 0) return fail(SMX_ERR_STATE, "smx_get_permutation: frame %d has not been written", f);
    CK(cudaSetDevice(s->cfg.device));
    int n = s->P.n;
    const uint32_t* d = s->orders[s->order_of[f]].perm;
    if (!d) { for (int i = 0; i < n; i++) perm[i] = i; return SMX_OK; }
    CK(cudaMemcpyAsync(perm, d, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SMX_OK;
}
int smx_get_grid(smx_sim* s, float* g_in, float* g_out) {
    if (!s) return fail(SMX_ERR_ARG, "smx_get_grid: null simulator");
    CK(cudaSetDevice(s->cfg.device));
    if (g_in && s->last_fwd >= 0) {
        // k_grid_op re-zeroes g_in for the next substep; the values of the last substep live in its checkpoint record
        int f = s->last_fwd;
        if (!s->ckpt || s->order_of[f] < 0 || s->ckpt_order[f] != s->orders[s->order_of[f]].uid)
            return fail(SMX_ERR_STATE, "smx_get_grid: g_in of substep %d was not retained (grid checkpoints disabled)", f);
        Order& o = s->orders[s->order_of[f]];
        launch_pdl(s, k_ckpt_copy, grid_blocks_launch(s), 256, 0, s->dense ? nullptr : o.blocks, o.nblocks, s->B * s->P.nb3, s->ckpt_cap, s->ckpt + (size_t)f * s->ckpt_rec,
                                                                 s->g_in, nullptr, nullptr, 1, s->counters, nullptr, nullptr, nullptr);
        CKL(s);
        s->g_in_clean_uid = -1;
    }
    size_t Gb = (size_t)s->P.Gb;      // batch 0 only
    if (!s->g_lin) CK(cudaMalloc(&s->g_lin, Gb * sizeof(float4)));
    const float4* src[2] = {s->g_in, s->g_out}; float* dst[2] = {g_in, g_out};
    for (int a = 0; a < 2; a++) {
        if (!dst[a]) continue;
        k_grid_linear<<<nblk((long long)Gb, 256), 256, 0, s->stream>>>(s->P.ng, s->P.nb, src[a], s->g_lin); CKL(s);
        CK(cudaMemcpyAsync(dst[a], s->g_lin, Gb * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return SMX_OK;
}
int smx_get_counters(smx_sim* s, int64_t out[4]) {
    if (!s || !out) return fail(SMX_ERR_ARG, "smx_get_counters: null argument");
    CK(cudaSetDevice(s->cfg.device));
    unsigned long long h[4];
    CK(cudaMemcpyAsync(h, s->counters, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    int nb = -1;
    CK(cudaStreamSynchronize(s->stream));
    if (!s->dense)
        for (int f = s->cfg.max_steps - 1; f >= 0; f--)
            if (s->order_of[f] >= 0 && s->orders[s->order_of[f]].nblocks) { CK(cudaMemcpy(&nb, s->orders[s->order_of[f]].nblocks, sizeof(int), cudaMemcpyDeviceToHost)); break; }
    out[0] = (int64_t)h[0]; out[1] = (int64_t)h[1]; out[2] = s->n_resorts; out[3] = nb;
    s->last_ckpt_overflow = (long long)h[2];
    return SMX_OK;
}
int smx_frame_component_dev(smx_sim* s, int32_t f, int32_t c, void** ptr) {
    TRY(check_frame(s, f, "smx_frame_component_dev"));
    if (!ptr || c < 0 || c >= 24) return fail(SMX_ERR_ARG, "smx_frame_component_dev: bad component");
    *ptr = s->frame_ptr(f) + comp_pos(c) / 4 * 4 * s->P.stride + comp_pos(c) % 4;       // element j at ptr[4 * j]
    return SMX_OK;
}
int smx_timer_start(smx_sim* s) {
    if (!s) return fail(SMX_ERR_ARG, "smx_timer_start: null simulator");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaEventRecord(s->ev0, s->stream));
    return SMX_OK;
}
int smx_timer_stop(smx_sim* s, float* ms) {
    if (!s || !ms) return fail(SMX_ERR_ARG, "smx_timer_stop: null argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaEventRecord(s->ev1, s->stream));
    CK(cudaEventSynchronize(s->ev1));
    CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return SMX_OK;
}
int64_t smx_launch_count(smx_sim* s) { return s ? s->launches : 0; }

// Mesh.trimesh2sdf (mesh.py:178-241) on the GPU; stand-alone (no simulator handle).  Outputs are host arrays
// sdf[r0*r1*r2], normal[r0*r1*r2*3] sampled at lower + (i,j,k)*dx.
int smx_build_sdf_table(const double* vertices, int32_t nv, const int32_t* faces, int32_t nf, const int32_t res[3], const double lower[3], double dx,
                        double* sdf_out, double* normal_out, int32_t device) {
    if (!vertices || !faces || !res || !lower || !sdf_out || !normal_out || nv < 3 || nf < 1 || dx <= 0) return fail(SMX_ERR_ARG, "smx_build_sdf_table: bad argument");
    for (int i = 0; i < 3 * nf; i++) if (faces[i] < 0 || faces[i] >= nv) return fail(SMX_ERR_RANGE, "smx_build_sdf_table: face index %d outside [0, %d)", faces[i], nv);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SMX_ERR_CUDA, "smx_build_sdf_table: no CUDA device; this library has no CPU path"); }
    CK(cudaSetDevice(device));
    long long total = (long long)res[0] * res[1] * res[2];
    double *dv = nullptr, *ds = nullptr, *dn = nullptr; int* df = nullptr;
    CK(cudaMalloc(&dv, (size_t)nv * 3 * sizeof(double))); CK(cudaMalloc(&df, (size_t)nf * 3 * sizeof(int)));
    CK(cudaMalloc(&ds, (size_t)total * sizeof(double))); CK(cudaMalloc(&dn, (size_t)total * 3 * sizeof(double)));
    CK(cudaMemcpy(dv, vertices, (size_t)nv * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(df, faces, (size_t)nf * 3 * sizeof(int), cudaMemcpyHostToDevice));
    k_build_sdf<<<(unsigned)((total + 127) / 128), 128>>>(dv, df, nf, res[0], res[1], res[2], lower[0], lower[1], lower[2], dx, ds, dn);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(sdf_out, ds, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(normal_out, dn, (size_t)total * 3 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dv); cudaFree(df); cudaFree(ds); cudaFree(dn);
    if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "smx_build_sdf_table: %s", cudaGetErrorString(e));
    return SMX_OK;
}

}  // extern "C"
// shared by smx_profile_substep / smx_profile_step: run `body` with an event after every launch, sum the device time per kernel class
template <typename Body>
static int profile_run(smx_sim* s, const char** names, float* ms, int32_t* launches, int32_t* count, Body&& body) {
    if (!s || !names || !ms || !count) return fail(SMX_ERR_ARG, "smx_profile_*: null argument");
    CK(cudaSetDevice(s->cfg.device));
    s->marks.clear();
    s->prof = true;
    int r = prof_mark(s, "start");
    if (r == SMX_OK) r = body();
    s->prof = false;
    if (r != SMX_OK) { for (auto& m : s->marks) cudaEventDestroy(m.second); s->marks.clear(); return r; }
    CK(cudaStreamSynchronize(s->stream));
    int n = 0;
    for (size_t i = 1; i < s->marks.size(); i++) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, s->marks[i - 1].second, s->marks[i].second));
        int k = -1;
        for (int j = 0; j < n; j++) if (!strcmp(names[j], s->marks[i].first)) k = j;
        if (k < 0) { if (n >= 32) continue; k = n++; names[k] = s->marks[i].first; ms[k] = 0.f; if (launches) launches[k] = 0; }
        ms[k] += t;
        if (launches) launches[k]++;
    }
    for (auto& m : s->marks) cudaEventDestroy(m.second);
    s->marks.clear();
    *count = n;
    return SMX_OK;
}
extern "C" {
int smx_profile_substep(smx_sim* s, int32_t f, int32_t backward, const char** names, float* ms, int32_t* count) {
    return profile_run(s, names, ms, nullptr, count, [&]() { return backward ? smx_substep_grad(s, f) : smx_substep(s, f); });
}
int smx_profile_step(smx_sim* s, int32_t f, int32_t n, int32_t backward, const char** names, float* ms, int32_t* launches, int32_t* count) {
    return profile_run(s, names, ms, launches, count, [&]() { return backward ? smx_step_grad(s, f, n) : smx_step(s, f, n); });
}

// 16 floats on the device: [0..2] sum of the adjoint of x, [3..5] of v, [6] |x adjoint|^2, [7] |v adjoint|^2, [8] particle count of
// the adjoint of frame f -- what a rank all-reduces after a rollout (stream-ordered on the simulator's stream, no host sync)
int smx_grad_summary_dev(smx_sim* s, int32_t f, float* out16_dev) {
    TRY(check_frame(s, f, "smx_grad_summary_dev"));
    if (!out16_dev) return fail(SMX_ERR_ARG, "smx_grad_summary_dev: null output");
    if (s->adj_frame != f) return fail(SMX_ERR_STATE, "smx_grad_summary_dev: no backward pass has produced the adjoint of frame %d", f);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemsetAsync(out16_dev, 0, 16 * sizeof(float), s->stream));
    if (s->P.n > 0) { k_grad_summary<<<std::min(nblk(s->P.n, 256), s->sm_count * 8), 256, 0, s->stream>>>(s->P.n, s->P.stride, s->adj_cur, out16_dev); CKL(s); }
    return SMX_OK;
}

}  // extern "C"
