// smx_kernels.cuh -- hand-written sm_100a kernels of the MLS-MPM substep and of its adjoint.
//
// Data layout (all fp32 in HBM):
//   particle frame  : six float4 PLANES [6][stride]: (x0 x1 x2 v0) (v1 v2 C0 C1) (C2..C5) (C6 C7 C8 F0) (F1..F4) (F5..F8),
//                     i.e. storage position of get_state column c (x0..2 v3..5 F6..14 C15..23) = comp_pos(c); one 16-byte
//                     load/store per plane and lane (512 contiguous bytes per warp).  x, v, C (written by G2P) come first,
//                     F (written by P2G) last; particles are physically sorted by block-major cell key, so a warp's 32
//                     particles share ~4 cells
//   SVD record      : four float4 planes [4][stride] per substep: (u0 | u1.x) (u1.yz | v0.xy) (v0.z | v1) (e0 e1 e2 J-1):
//                     the first two columns of U and V (the third is their cross product), sigma - 1 and det(F_tmp) - 1 of
//                     the forward P2G, so that the adjoint does not repeat the Jacobi SVD
//   grid            : float4 per node, BLOCK-MAJOR: 4x4x4-node blocks are contiguous 1 KB chunks
//                     g_in  = (momentum xyz, mass)          target of P2G        (mpm_simulator.py:261-262)
//                     g_out = (velocity xyz, active flag)   source of G2P        (:296-297, :403-404, :443)
//                     g_mix = (velocity xyz, active flag)   grid_v_mixed         (:403)
//                     gg_out / gg_mix: adjoints; grid_grad overwrites gg_out with (d g_in xyz, d mass)
//   primitives      : pstate [B][P][T][13] fp32, pgrad [B][P][T][13] f64, ext_f [B][P][6] f64, ext_f_grad [B][P][6] fp32
//   batching        : B independent rollouts share one handle: particle slot j belongs to batch j / npb (the sort key
//                     carries the batch in its high bits), node indices are offset by batch * ng^3
//
// Stage map (reference kernel -> kernel here):
//   compute_F_tmp + svd + p2g (:125-133, :198-262)           -> k_p2g          (SVD and stress stay in registers)
//   grid_op / grid_op_mixed1 (:283-297, :396-404)            -> k_grid_op
//   grid_op_mixed2 + 3 + 4 (:406-443)                        -> k_contact      (gather, forecast contact, scatter)
//   g2p (:299-318)                                           -> k_g2p
//   g2p.grad                                                 -> k_g2p_grad
//   grid_op_mixed4.grad + 3.grad + 2.grad                    -> k_contact_grad
//   grid_op.grad / grid_op_mixed1.grad                       -> k_grid_grad
//   p2g.grad + svd_grad + compute_F_tmp.grad (:135-157)      -> k_p2g_grad
#pragma once
#include "smx_contact.cuh"

namespace smx {

struct Params {
    int n;                  // particles
    long long stride;       // floats between components of a frame
    int ng, nb;             // grid nodes per axis, blocks per axis (ng/4)
    float dt, dx, inv_dx, p_mass, mu, lam, cs;      // cs = -dt*p_vol*4*inv_dx^2 (mpm_simulator.py:247)
    float vm_c;             // yield_stress / (2 mu): von Mises return mapping (smx_set_plasticity)
    int vm;                 // 1: co-rotated plastic uses the von Mises return mapping instead of the sigma clip
    float gx, gy, gz;       // gravity
    int sticky;             // ground_friction >= 10
    int material, ptype, ctype, substeps, n_control, np;
    int dbg;                // timing experiments only (SMX_DBG): 1 skip the reductions of the flush, 2 skip the flush walk, 4 skip staging stores
    int nbatch, npb, Gb, nb3;   // independent rollouts batched in one handle: count, particles per batch, nodes and blocks per batch
};

struct PrimSet {            // device pointers shared by all kernels that touch primitives
    const PrimDev* prims;   // [np]
    const float* pstate;    // [np][T][13]
    double* pgrad;          // [np][T][13]
    double* ext_f;          // [np][6]
    const float* ext_f_grad;// [np][6]
    int T;
};
// batch-major primitive arrays: [batch][SMX_MAXP][...]
__device__ __forceinline__ const float* pstate_at(const PrimSet& ps, int b, int i, int f) { return ps.pstate + (((size_t)b * SMX_MAXP + i) * ps.T + f) * 13; }
__device__ __forceinline__ double* pgrad_at(const PrimSet& ps, int b, int i, int f) { return ps.pgrad + (((size_t)b * SMX_MAXP + i) * ps.T + f) * 13; }
__device__ __forceinline__ double* ext_f_at(const PrimSet& ps, int b, int i) { return ps.ext_f + ((size_t)b * SMX_MAXP + i) * 6; }
__device__ __forceinline__ const float* ext_f_grad_at(const PrimSet& ps, int b, int i) { return ps.ext_f_grad + ((size_t)b * SMX_MAXP + i) * 6; }
__device__ __forceinline__ int batch_of(const Params& P, int j) { return P.nbatch > 1 ? j / P.npb : 0; }

// ---- particle-frame layout -------------------------------------------------------------------------------------------
#define SMX_NPLANES 6       // float4 planes per particle frame
#define SMX_RPLANES 4       // float4 planes per SVD record
// get_state column (x0..2 v3..5 F6..14 C15..23) -> storage position (x v C F)
__host__ __device__ __forceinline__ int comp_pos(int c) { return c < 6 ? c : (c < 15 ? c + 9 : c - 9); }
__device__ __forceinline__ long long comp_index(long long stride, int j, int c) {
    int p = comp_pos(c);
    return (((long long)(p >> 2) * stride + j) << 2) + (p & 3);
}
// read-once load of a streaming particle plane: ld.global.nc.L1::no_allocate -- the planes pass through L1 without evicting the grid
// nodes that the 27-node gathers of the same kernel re-use (fused adjoint 142.6 -> 140.4 us; -DSMX_STREAM_ALLOC for the plain __ldg)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
#ifndef SMX_STREAM_ALLOC
    float4 r;       // not volatile: read-only data, the compiler may schedule the load freely
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ float4 ld_plane(const float* __restrict__ fr, long long stride, int j, int p) {
    return ldg_stream(reinterpret_cast<const float4*>(fr) + ((long long)p * stride + j));
}
__device__ __forceinline__ void st_plane(float* __restrict__ fr, long long stride, int j, int p, float4 v) {
    reinterpret_cast<float4*>(fr)[(long long)p * stride + j] = v;
}
// L2 prefetch of planes [p0, p1) of particle slot j (issued one wave of CTAs ahead of the loads: the streaming frame data
// then arrives from L2 instead of HBM when its CTA starts)
#ifndef SMX_PF_WAVE
#define SMX_PF_WAVE 1
#endif
__device__ __forceinline__ void prefetch_planes(const void* fr, long long stride, long long j, long long n, int p0, int p1) {
    if (!SMX_PF_WAVE || j >= n) return;
    const float4* b = reinterpret_cast<const float4*>(fr) + j;
    for (int p = p0; p < p1; p++) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + p * stride));
}
__device__ __forceinline__ V3 load_x(const float* __restrict__ fr, long long stride, int j) {
    float4 a = ld_plane(fr, stride, j, 0);
    return v3(a.x, a.y, a.z);
}

// Programmatic dependent launch: let the next kernel of the stream be scheduled early, then wait until the previous one has
// completed and flushed its memory.  Both are no-ops for a kernel that was not launched with the PDL attribute.
__device__ __forceinline__ void pdl_prologue() {
    // wait first, then allow the dependents: the look-ahead stays at one kernel.  (Triggering before the wait lets the CTAs of the
    // kernel after next become resident behind a small grid kernel that is itself still waiting: measured 3.86 -> 3.47 G/s once the
    // grid kernels were lean enough to be co-resident with the particle kernels.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
}

// Timing experiment (SMX_DBG bits 8..: step in ns): de-phase the warps that share an SM.  A PDL launch releases every resident CTA of
// the first wave at the same instant (griddepcontrol.wait), so all warps of an SM walk through the load / gather / SVD / scatter phases
// of the kernel together; slot s of the SM (blockIdx / #SMs) and warp w start (s * warps + w) * step later.  Later waves inherit the offsets.
__device__ __forceinline__ void stagger_first_wave(int dbg, int slots) {
#ifndef SMX_STAGGER
    return;
#endif
    const unsigned step = (unsigned)dbg >> 8;
    if (step) {
        unsigned nsm;
        asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        const unsigned slot = blockIdx.x / nsm;
        if (slot < (unsigned)slots) {
            const unsigned u = slot * (blockDim.x >> 5) + (threadIdx.x >> 5);
            if (u) __nanosleep(u * step);
        }
    }
}

#define SMX_TPB 128         // gather-type particle kernels
#ifndef SMX_TPB_SC
#define SMX_TPB_SC 64       // P2G: two warps x 13.6 KB of staging, eight CTAs per SM at 128 registers (16 warps).  Same residency as four
#endif                      // CTAs of four warps, finer-grained CTA turnover: G2P2G 94.9 -> 92.7 us (96- and 32-thread CTAs measured slower)
#ifndef SMX_SC_MINB
#define SMX_SC_MINB 8
#endif
#ifndef SMX_TPB_G2PG
#define SMX_TPB_G2PG 96     // G2P adjoint: three warps x 10.3 KB of staging, seven CTAs per SM at <= 96 registers (21 warps)
#endif

__device__ __forceinline__ void red_add_f4(float4* addr, float a, float b, float c, float d) {
    // one 16-byte reduction (SASS: REDG.E.ADD.F32x4) instead of four scalar atomics
    atomicAdd(addr, make_float4(a, b, c, d));
}

// ------------------------------------------------------------------------------------------------
// Warp-level aggregation of the 27-node scatter.  Particles are stored sorted by cell, so the 32
// particles of a warp fall into a handful of runs with the same base cell.  Every lane parks its 27
// float4 contributions and its 9 per-axis node offsets in shared memory; then lane o < 27 owns stencil
// node o and walks the warp's particles in order, run by run (the run boundaries come from one ballot,
// so the walk is warp-uniform: no divergence, conflict-free 128-bit shared loads), and issues ONE
// REDG.E.ADD.F32x4 per run and node.  With ~8 particles per cell this cuts the L2 reductions ~7x
// (27 per particle -> 27 per run) and makes the within-run summation order deterministic.
// ------------------------------------------------------------------------------------------------
struct WarpStage {
    float4 val[32 * 27];    // [lane][offset]; row stride 27 float4 -> conflict-free 128-bit stores
    uint32_t key[32];       // [lane] block-major index of the slot's base cell: the owning lane derives its node address per run (RunNode)
};
// block-major index of a stencil's base cell (batch offset included): the key of the scatter runs
__device__ __forceinline__ uint32_t base_node(const struct Stencil& s);

// packed fp32x2 add (sm_100a: one FADD2 instead of two FADD)
__device__ __forceinline__ void add_f4(float4& a, const float4& b) {
    float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Common head of the flush: park the packed base cells, find the runs.  ends bit p: slot p is the last of its run; alive bit p: slot p
// holds a particle.
__device__ __forceinline__ void flush_prologue(uint32_t* keys, uint32_t key, bool live, unsigned& ends, unsigned& alive) {
    const unsigned lane = threadIdx.x & 31;
    keys[lane] = key;
    uint32_t k = live ? key : 0xffffffffu;
    uint32_t prev = __shfl_up_sync(0xffffffffu, k, 1);
    bool head = (lane == 0) || (k != prev);
    ends = (__ballot_sync(0xffffffffu, head) >> 1) | 0x80000000u;
    alive = __ballot_sync(0xffffffffu, live);
    __syncwarp();
}
// Node (bx + a, by + b, bz + c) of the run whose first slot carries `key` = the block-major index of the run's base cell (its low six
// bits are the base's position inside its 4x4x4 block).  Stepping `a` nodes along x adds a * 16, plus (block stride - 64) when the step
// leaves the block, i.e. when (bx & 3) + a > 3: for a = 1 that is (key & 0x30) == 0x30, for a = 2 it is key & 0x20; likewise y, z.
// The owning lane keeps its three masks and the in-block offset: one AND + compare + predicated add per axis at a run end.
struct RunNode {
    uint32_t d0, mx, my, mz, wx, wy;
    __device__ __forceinline__ RunNode(int a, int b, int c, int nb) {
        d0 = (uint32_t)(a * 16 + b * 4 + c);
        mx = a == 0 ? 0xffffffffu : (a == 1 ? 0x30u : 0x20u);       // all ones never matches a live key
        my = b == 0 ? 0xffffffffu : (b == 1 ? 0x0cu : 0x08u);
        mz = c == 0 ? 0xffffffffu : (c == 1 ? 0x03u : 0x02u);
        wx = (uint32_t)(nb * nb * 64 - 64); wy = (uint32_t)(nb * 64 - 16);
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t key) const {
        uint32_t n = key + d0;
        if ((key & mx) == mx) n += wx;
        if ((key & my) == my) n += wy;
        if ((key & mz) == mz) n += 60u;
        return n;
    }
};
// The walk is unrolled over the 32 slots in batches of 8: the eight 128-bit loads of a batch are in flight together and the additions
// alternate between two accumulators (two dependent chains instead of one).  A batch without a run end (one warp-uniform test of
// the ballot) is eight plain additions; otherwise the run ends are tested slot by slot against immediate bits, and after the last
// slot of a run the lane issues its reduction.
__device__ __forceinline__ void warp_stage_flush(WarpStage& st, uint32_t key, bool live, float4* __restrict__ grid, int nb, int Gb, int dbg = 0) {
    unsigned ends, alive;
    flush_prologue(st.key, key, live, ends, alive);
    const unsigned lane = threadIdx.x & 31;
    if (lane >= 27 || (dbg & 2)) return;
    const RunNode rn(lane / 9, (lane / 3) % 3, lane % 3, nb);
    const float4* src = st.val + lane;
    const int nl = __popc(alive);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc0 = zero, acc1 = zero;
    int p0 = 0;
#pragma unroll 1     // four trips of eight unrolled slots: the fully unrolled walk (1000 instructions) stalled on instruction fetch
    for (int base = 0; base < 32; base += 8) {
        if (base >= nl) break;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = src[(base + i) * 27];
        const unsigned e8 = (ends >> base) & 0xffu, l8 = (alive >> base) & 0xffu;
        if (e8 == 0u) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) { add_f4(acc0, v[i]); add_f4(acc1, v[i + 1]); }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            add_f4((i & 1) ? acc1 : acc0, v[i]);
            if (e8 & (1u << i)) {                               // last slot of its run
                if ((l8 & (1u << i)) && !(dbg & 1)) {
                    add_f4(acc0, acc1);
                    atomicAdd(grid + rn(st.key[p0]), acc0);
                }
                acc0 = zero; acc1 = zero;
                p0 = base + i + 1;
            }
        }
    }
}

// Three-component variant (adjoint of G2P: nothing is scattered into the mass slot): xy and z are staged in separate arrays
// (8-byte and 4-byte accesses, both conflict-free), 25 % less shared-memory traffic and 25 % less shared memory per warp.
struct WarpStage3 {
    float2 xy[32 * 27];     // [lane][offset]
    float z[32 * 27];
    uint32_t key[32];
};
__device__ __forceinline__ void warp_stage_flush3(WarpStage3& st, uint32_t key, bool live, float4* __restrict__ grid, int nb, int Gb, int dbg = 0) {
    unsigned ends, alive;
    flush_prologue(st.key, key, live, ends, alive);
    const unsigned lane = threadIdx.x & 31;
    if (lane >= 27 || (dbg & 2)) return;
    const RunNode rn(lane / 9, (lane / 3) % 3, lane % 3, nb);
    const float2* sxy = st.xy + lane;
    const float* sz = st.z + lane;
    const int nl = __popc(alive);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
    float az0 = 0.f, az1 = 0.f;
    int p0 = 0;
#pragma unroll 1
    for (int base = 0; base < 32; base += 8) {
        if (base >= nl) break;
        float2 v[8]; float z[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { v[i] = sxy[(base + i) * 27]; z[i] = sz[(base + i) * 27]; }
        const unsigned e8 = (ends >> base) & 0xffu, l8 = (alive >> base) & 0xffu;
        if (e8 == 0u) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) { acc0 = __fadd2_rn(acc0, v[i]); az0 += z[i]; acc1 = __fadd2_rn(acc1, v[i + 1]); az1 += z[i + 1]; }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i & 1) { acc1 = __fadd2_rn(acc1, v[i]); az1 += z[i]; } else { acc0 = __fadd2_rn(acc0, v[i]); az0 += z[i]; }
            if (e8 & (1u << i)) {                               // last slot of its run
                if ((l8 & (1u << i)) && !(dbg & 1)) {
                    acc0 = __fadd2_rn(acc0, acc1);
                    atomicAdd(grid + rn(st.key[p0]), make_float4(acc0.x, acc0.y, az0 + az1, 0.f));
                }
                acc0 = make_float2(0.f, 0.f); acc1 = acc0; az0 = 0.f; az1 = 0.f;
                p0 = base + i + 1;
            }
        }
    }
}

// quadratic B-spline stencil of one particle (mpm_simulator.py:215-217)
struct Stencil {
    uint32_t ox[3], oy[3], oz[3];   // block-major address contributions per axis offset (unsigned: one IMAD.WIDE.U32 per node address)
    int bx, by, bz;
    float fx, fy, fz;
    float wx[3], wy[3], wz[3];
};
__device__ __forceinline__ uint32_t base_node(const Stencil& s) { return s.ox[0] + s.oy[0] + s.oz[0]; }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ uint32_t node_index(int i, int j, int k, int nb) {
    return (uint32_t)((((i >> 2) * nb + (j >> 2)) * nb + (k >> 2)) * 64 + (((i & 3) << 4) | ((j & 3) << 2) | (k & 3)));
}
// sort key of a particle: block-major index of its base cell (the documented variant of SURVEY.md 8a-0)
__device__ __forceinline__ uint32_t cell_key(float x, float y, float z, float inv_dx, int ng, int nb, int* clamped) {
    int bx = (int)(x * inv_dx - 0.5f), by = (int)(y * inv_dx - 0.5f), bz = (int)(z * inv_dx - 0.5f);
    int cx = clampi(bx, 0, ng - 3), cy = clampi(by, 0, ng - 3), cz = clampi(bz, 0, ng - 3);
    if (clamped) *clamped = (cx != bx) | (cy != by) | (cz != bz);
    return node_index(cx, cy, cz, nb);
}
__device__ __forceinline__ void axis_weights(float f, float* w) {
    w[0] = 0.5f * (1.5f - f) * (1.5f - f); w[1] = 0.75f - (f - 1.f) * (f - 1.f); w[2] = 0.5f * (f - 0.5f) * (f - 0.5f);
}
__device__ __forceinline__ void axis_dweights(float f, float* d) { d[0] = f - 1.5f; d[1] = -2.f * (f - 1.f); d[2] = f - 0.5f; }

__device__ __forceinline__ Stencil make_stencil(float x, float y, float z, const Params& P, int bt = 0) {
    Stencil s;
    // x*inv_dx (power-of-two scale), -0.5 and the subtraction of the integer base are exact in fp32, so base and
    // fx equal the reference's f64 values for the same fp32 x
    float sx = x * P.inv_dx, sy = y * P.inv_dx, sz = z * P.inv_dx;
    s.bx = clampi((int)(sx - 0.5f), 0, P.ng - 3); s.by = clampi((int)(sy - 0.5f), 0, P.ng - 3); s.bz = clampi((int)(sz - 0.5f), 0, P.ng - 3);
    s.fx = sx - (float)s.bx; s.fy = sy - (float)s.by; s.fz = sz - (float)s.bz;
    axis_weights(s.fx, s.wx); axis_weights(s.fy, s.wy); axis_weights(s.fz, s.wz);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int i = s.bx + a, j = s.by + a, k = s.bz + a;
        s.ox[a] = (uint32_t)(bt * P.Gb + ((i >> 2) * P.nb * P.nb) * 64 + ((i & 3) << 4));
        s.oy[a] = (uint32_t)(((j >> 2) * P.nb) * 64 + ((j & 3) << 2));
        s.oz[a] = (uint32_t)((k >> 2) * 64 + (k & 3));
    }
    return s;
}

// warp-reduce `v` and let lane 0 add it to a double accumulator
__device__ __forceinline__ void warp_sum_to(double* dst, float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dst, (double)v);
}
// lanes of a warp normally belong to one batch; a warp straddling two batches falls back to per-lane atomics
__device__ __forceinline__ bool warp_one_batch(int bt) { return __all_sync(0xffffffffu, bt == __shfl_sync(0xffffffffu, bt, 0)); }
__device__ __forceinline__ void commit_wrench(double* ext_f, V3 bf, V3 r, bool active, int bt) {
    if (!__any_sync(0xffffffffu, active)) return;
    V3 bt3 = cross(r, bf);
    if (!active) { bf = v3(0, 0, 0); bt3 = v3(0, 0, 0); }
    if (warp_one_batch(bt)) {
        warp_sum_to(ext_f + 0, bf.x); warp_sum_to(ext_f + 1, bf.y); warp_sum_to(ext_f + 2, bf.z);
        warp_sum_to(ext_f + 3, bt3.x); warp_sum_to(ext_f + 4, bt3.y); warp_sum_to(ext_f + 5, bt3.z);
    } else if (active) {
        atomicAdd(ext_f + 0, (double)bf.x); atomicAdd(ext_f + 1, (double)bf.y); atomicAdd(ext_f + 2, (double)bf.z);
        atomicAdd(ext_f + 3, (double)bt3.x); atomicAdd(ext_f + 4, (double)bt3.y); atomicAdd(ext_f + 5, (double)bt3.z);
    }
}
__device__ __forceinline__ void commit_prim_grad(double* g13, const PrimGrad& G, bool active, int bt) {
    if (!__any_sync(0xffffffffu, active)) return;
    float v[13] = {G.pos.x, G.pos.y, G.pos.z, G.rot.w, G.rot.x, G.rot.y, G.rot.z, G.v.x, G.v.y, G.v.z, G.w.x, G.w.y, G.w.z};
    if (warp_one_batch(bt)) {
#pragma unroll
        for (int i = 0; i < 13; i++) warp_sum_to(g13 + i, active ? v[i] : 0.f);
    } else if (active) {
#pragma unroll
        for (int i = 0; i < 13; i++) if (v[i] != 0.f) atomicAdd(g13 + i, (double)v[i]);
    }
}

// MAT = material_model * 3 + ptype (0..5); 6 = co-rotated plastic with the von Mises return mapping of the soft_cloth variant
// (soft_cloth/engine/mpm_simulator.py:172-189, :232) instead of the sigma clip (softmac mpm_simulator.py:226-229)
__host__ __device__ constexpr int mat_model(int MAT) { return MAT == 6 ? 0 : MAT / 3; }
__host__ __device__ constexpr int mat_ptype(int MAT) { return MAT == 6 ? 0 : MAT % 3; }
__host__ __device__ constexpr bool mat_vm(int MAT) { return MAT == 6; }

// compute_von_mises in deviation form: e = sigma - 1 in, g = sigma_new - 1 out; returns whether the particle yields (otherwise g = e and
// new_F stays F_tmp).  c = yield_stress / (2 mu).  sig = max(sig, 0.05); epsilon = log(sig); epsilon_hat = epsilon - mean;
// norm = sqrt(epsilon_hat . epsilon_hat + 1e-8); delta_gamma = norm - c; yields: epsilon -= (delta_gamma / norm) epsilon_hat, sig = exp(epsilon).
// D (optional): d g_k / d e_i with Taichi's conventions (max passes its gradient iff 0.05 < sig; the yield test carries none).
__device__ __forceinline__ bool von_mises_dev(const float* e, float c, float* g, float (*D)[3] = nullptr) {
    float eps[3], eh[3], sc[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { sc[d] = fmaxf(e[d], 0.05f - 1.f); eps[d] = log1pf(sc[d]); }
    float m = (eps[0] + eps[1] + eps[2]) * (1.f / 3.f);
    float q = 1e-8f;
#pragma unroll
    for (int d = 0; d < 3; d++) { eh[d] = eps[d] - m; q = fmaf(eh[d], eh[d], q); }
    float nrm = sqrtf(q);
    bool yields = nrm - c > 0.f;
    float k = 1.f - c / nrm;            // delta_gamma / norm
#pragma unroll
    for (int d = 0; d < 3; d++) g[d] = yields ? expm1f(fmaf(-k, eh[d], eps[d])) : e[d];
    if (D) {
        float c3 = c / (nrm * nrm * nrm);
#pragma unroll
        for (int kk = 0; kk < 3; kk++)
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float id = kk == i ? 1.f : 0.f;
                float de = id - k * (id - 1.f / 3.f) - c3 * eh[i] * eh[kk];
                D[kk][i] = (e[i] > 0.05f - 1.f) ? (1.f + g[kk]) * de / (1.f + sc[i]) : 0.f;
            }
    }
    return yields;
}

// ------------------------------------------------------------------------------------------------
// material point update (mpm_simulator.py:125-133, 219-248), deviation form.  Et = F_tmp - I.
// ------------------------------------------------------------------------------------------------
struct Material {
    M3 newF;        // F[f+1]
    M3 stress;      // before the cs prefactor
    float J, Jm1;
    Svd svd;        // valid for co-rotated plastic / elastic
};
__device__ __forceinline__ M3 compute_Et(const M3& C, const M3& F, float dt) {
    // F_tmp - I = (F - I) + dt*C + dt*C*(F - I);  F - I is exact in fp32 for F near I
    M3 EF = F; EF.m[0] -= 1.f; EF.m[4] -= 1.f; EF.m[8] -= 1.f;
    M3 CE = mul(C, EF), Et;
#pragma unroll
    for (int i = 0; i < 9; i++) Et.m[i] = fmaf(dt, C.m[i] + CE.m[i], EF.m[i]);
    return Et;
}
// REC: m.svd and m.Jm1 were loaded from the SVD record of the forward pass (co-rotated plastic / elastic only)
template <int MAT, bool REC = false>
__device__ __forceinline__ void material_update(const M3& Et, const Params& P, Material& m) {
    constexpr int model = mat_model(MAT), ptype = mat_ptype(MAT);
    constexpr bool rec = REC && model == 0 && ptype != 2;
    if (!rec) m.Jm1 = det_minus_one(Et);
    m.J = 1.f + m.Jm1;
    M3 Ftmp = Et; Ftmp.m[0] += 1.f; Ftmp.m[4] += 1.f; Ftmp.m[8] += 1.f;
    if (model == 0) {
        float iso = P.lam * m.J * m.Jm1;
        if (ptype == 2) {                 // liquid: mu == 0, R never contributes; skip the SVD
            float c = cbrtf(m.J);
            m.newF = scale(c, m3_identity());
            m.stress = m3_zero();
            m.stress.m[0] = iso; m.stress.m[4] = iso; m.stress.m[8] = iso;
        } else {
            if (!rec) m.svd = svd_dev(Et);
            float g0 = m.svd.e[0], g1 = m.svd.e[1], g2 = m.svd.e[2];
            m.newF = Ftmp;
            if (mat_vm(MAT)) {            // soft_cloth plastic: von Mises return mapping; new_F stays F_tmp unless the particle yields
                float g3[3];
                bool yields = von_mises_dev(m.svd.e, P.vm_c, g3);
                g0 = g3[0]; g1 = g3[1]; g2 = g3[2];
                if (__any_sync(__activemask(), yields))
                    m.newF = add(Ftmp, udvt(m.svd.U, g0 - m.svd.e[0], g1 - m.svd.e[1], g2 - m.svd.e[2], m.svd.V));
            } else if (ptype == 0) {      // plastic: clip sigma to [1-2e-3, 1+3e-3] (:226-229)
                g0 = fminf(fmaxf(g0, -2e-3f), 3e-3f); g1 = fminf(fmaxf(g1, -2e-3f), 3e-3f); g2 = fminf(fmaxf(g2, -2e-3f), 3e-3f);
                // new_F = F_tmp + U (Sc - S) V^T: exactly F_tmp when nothing is clipped (skipped warp-wide in that case)
                bool clipped = (g0 != m.svd.e[0]) | (g1 != m.svd.e[1]) | (g2 != m.svd.e[2]);
                if (__any_sync(__activemask(), clipped))
                    m.newF = add(Ftmp, udvt(m.svd.U, g0 - m.svd.e[0], g1 - m.svd.e[1], g2 - m.svd.e[2], m.svd.V));
            }
            // stress = 2 mu (new_F - R) new_F^T + lam J (J - 1) I with new_F = U Sc V^T, R = U V^T (:234-236)
            //        = U diag(2 mu g (1 + g) + lam J (J - 1)) U^T,  g = Sc - 1: symmetric, and accurate because g is carried, not Sc
            float mu2 = 2.f * P.mu;
            m.stress = udut(m.svd.U, fmaf(mu2 * g0, 1.f + g0, iso), fmaf(mu2 * g1, 1.f + g1, iso), fmaf(mu2 * g2, 1.f + g2, iso));
        }
    } else {                              // neo-Hookean (:237-245)
        if (ptype == 2) {
            float sq = sqrtf(m.J);
            m.newF = m3_zero(); m.newF.m[0] = sq; m.newF.m[4] = sq; m.newF.m[8] = 1.f;
        } else m.newF = Ftmp;
        m.stress = scale(P.mu, mulT(m.newF, m.newF));
        float iso = P.lam * log1pf(m.Jm1) - P.mu;
        m.stress.m[0] += iso; m.stress.m[4] += iso; m.stress.m[8] += iso;
    }
}

// unpack the six planes of one particle
__device__ __forceinline__ void unpack_state(const float4& p0, const float4& p1, const float4& p2, const float4& p3, const float4& p4, const float4& p5,
                                             V3& x, V3& v, M3& F, M3& C) {
    x = v3(p0.x, p0.y, p0.z); v = v3(p0.w, p1.x, p1.y);
    C.m[0] = p1.z; C.m[1] = p1.w; C.m[2] = p2.x; C.m[3] = p2.y; C.m[4] = p2.z; C.m[5] = p2.w; C.m[6] = p3.x; C.m[7] = p3.y; C.m[8] = p3.z;
    F.m[0] = p3.w; F.m[1] = p4.x; F.m[2] = p4.y; F.m[3] = p4.z; F.m[4] = p4.w; F.m[5] = p5.x; F.m[6] = p5.y; F.m[7] = p5.z; F.m[8] = p5.w;
}
__device__ __forceinline__ void load_state(const float* __restrict__ fr, long long stride, int j, V3& x, V3& v, M3& F, M3& C) {
    const float4* b = reinterpret_cast<const float4*>(fr) + j;
    float4 p0 = ldg_stream(b), p1 = ldg_stream(b + stride), p2 = ldg_stream(b + 2 * stride), p3 = ldg_stream(b + 3 * stride), p4 = ldg_stream(b + 4 * stride), p5 = ldg_stream(b + 5 * stride);
    unpack_state(p0, p1, p2, p3, p4, p5, x, v, F, C);
}
// loaders on a per-particle base pointer b (plane p at b[p * stride]); SM: the planes were staged in shared memory by a bulk copy
template <bool SM> __device__ __forceinline__ float4 ld4(const float4* p) { if (SM) return *p; else return ldg_stream(p); }
// re-read of a staged plane that the compiler must not merge with an earlier read (keeps the value out of registers in between)
__device__ __forceinline__ float4 lds4_again(const float4* p) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return r;
}
template <bool SM> __device__ __forceinline__ void load_state_b(const float4* b, long long stride, V3& x, V3& v, M3& F, M3& C) {
    float4 p0 = ld4<SM>(b), p1 = ld4<SM>(b + stride), p2 = ld4<SM>(b + 2 * stride), p3 = ld4<SM>(b + 3 * stride), p4 = ld4<SM>(b + 4 * stride), p5 = ld4<SM>(b + 5 * stride);
    unpack_state(p0, p1, p2, p3, p4, p5, x, v, F, C);
}
template <bool SM> __device__ __forceinline__ M3 load_F_b(const float4* b, long long stride) {
    float4 p3 = ld4<SM>(b + 3 * stride), p4 = ld4<SM>(b + 4 * stride), p5 = ld4<SM>(b + 5 * stride);
    M3 F; F.m[0] = p3.w; F.m[1] = p4.x; F.m[2] = p4.y; F.m[3] = p4.z; F.m[4] = p4.w; F.m[5] = p5.x; F.m[6] = p5.y; F.m[7] = p5.z; F.m[8] = p5.w;
    return F;
}
// x, v, C of a frame (planes 0..2 and the first three components of plane 3; F0 in plane 3.w belongs to P2G)
__device__ __forceinline__ void store_xvC(float* __restrict__ fr, long long stride, int j, V3 x, V3 v, const M3& C) {
    float4* b = reinterpret_cast<float4*>(fr) + j;
    b[0] = make_float4(x.x, x.y, x.z, v.x);
    b[stride] = make_float4(v.y, v.z, C.m[0], C.m[1]);
    b[2 * stride] = make_float4(C.m[2], C.m[3], C.m[4], C.m[5]);
    float* q = reinterpret_cast<float*>(b + 3 * stride);
    *reinterpret_cast<float2*>(q) = make_float2(C.m[6], C.m[7]);
    q[2] = C.m[8];
}
// F of a frame (plane 3.w and planes 4, 5)
__device__ __forceinline__ void store_F(float* __restrict__ fr, long long stride, int j, const M3& F) {
    float4* b = reinterpret_cast<float4*>(fr) + j;
    reinterpret_cast<float*>(b + 3 * stride)[3] = F.m[0];
    b[4 * stride] = make_float4(F.m[1], F.m[2], F.m[3], F.m[4]);
    b[5 * stride] = make_float4(F.m[5], F.m[6], F.m[7], F.m[8]);
}
__device__ __forceinline__ M3 load_F(const float* __restrict__ fr, long long stride, int j) {
    const float4* b = reinterpret_cast<const float4*>(fr) + j;
    float f0 = __ldg(reinterpret_cast<const float*>(b + 3 * stride) + 3);
    float4 p4 = ldg_stream(b + 4 * stride), p5 = ldg_stream(b + 5 * stride);
    M3 F; F.m[0] = f0; F.m[1] = p4.x; F.m[2] = p4.y; F.m[3] = p4.z; F.m[4] = p4.w; F.m[5] = p5.x; F.m[6] = p5.y; F.m[7] = p5.z; F.m[8] = p5.w;
    return F;
}
// SVD record of the forward P2G (see the layout comment at the top)
__device__ __forceinline__ void store_svd_rec(float4* __restrict__ rec, long long stride, int j, const Svd& sv, float Jm1) {
    float4* b = rec + j;
    b[0] = make_float4(sv.U.m[0], sv.U.m[3], sv.U.m[6], sv.U.m[1]);
    b[stride] = make_float4(sv.U.m[4], sv.U.m[7], sv.V.m[0], sv.V.m[3]);
    b[2 * stride] = make_float4(sv.V.m[6], sv.V.m[1], sv.V.m[4], sv.V.m[7]);
    b[3 * stride] = make_float4(sv.e[0], sv.e[1], sv.e[2], Jm1);
}
template <bool SM> __device__ __forceinline__ void load_svd_rec_b(const float4* b, long long stride, Svd& sv, float& Jm1) {
    float4 a = ld4<SM>(b), c = ld4<SM>(b + stride), d = ld4<SM>(b + 2 * stride), e = ld4<SM>(b + 3 * stride);
    V3 u0 = v3(a.x, a.y, a.z), u1 = v3(a.w, c.x, c.y), v0 = v3(c.z, c.w, d.x), v1 = v3(d.y, d.z, d.w);
    V3 u2 = cross(u0, u1), v2 = cross(v0, v1);       // U, V are proper rotations
    sv.U.m[0] = u0.x; sv.U.m[3] = u0.y; sv.U.m[6] = u0.z; sv.U.m[1] = u1.x; sv.U.m[4] = u1.y; sv.U.m[7] = u1.z; sv.U.m[2] = u2.x; sv.U.m[5] = u2.y; sv.U.m[8] = u2.z;
    sv.V.m[0] = v0.x; sv.V.m[3] = v0.y; sv.V.m[6] = v0.z; sv.V.m[1] = v1.x; sv.V.m[4] = v1.y; sv.V.m[7] = v1.z; sv.V.m[2] = v2.x; sv.V.m[5] = v2.y; sv.V.m[8] = v2.z;
    sv.e[0] = e.x; sv.e[1] = e.y; sv.e[2] = e.z; Jm1 = e.w;
}

// particle-contact impulses (collision_type == 1, :203-206) and the control impulse (:209-213)
__device__ __forceinline__ V3 particle_impulses(const Params& P, const PrimSet& ps, int f, int j, int bt, bool live, V3 x, V3 v,
                                                const int* __restrict__ ctrl_slot, const float* __restrict__ action, bool accumulate) {
    V3 imp = v3(0, 0, 0);
    if (P.ctype == 1) {
        for (int i = 0; i < P.np; i++) {
            if (!ps.prims[i].enabled) continue;
            PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
            bool act = false; V3 bf = v3(0, 0, 0), r = v3(0, 0, 0);
            if (live) imp += collide_particle_fwd(ps.prims[i], S, x, v, P.dt, act, bf, r);
            if (accumulate) commit_wrench(ext_f_at(ps, bt, i), bf, r, act, bt);
        }
    }
    if (P.n_control > 0 && live) {
        int ci = ctrl_slot[j];
        if (ci >= 0) { ci += bt * P.n_control; imp += (6e-4f * P.dt) * v3(action[3 * ci], action[3 * ci + 1], action[3 * ci + 2]); }
    }
    return imp;
}

// G2P gather of one particle (mpm_simulator.py:299-318): new velocity and APIC matrix from the 27 nodes of its stencil
__device__ __forceinline__ void g2p_gather(const Params& P, const Stencil& s, const float4* __restrict__ g_out, V3& nv, M3& Cn) {
    nv = v3(0, 0, 0);
    M3 B = m3_zero();       // sum w g (x) offset
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            // row sums over c first: R0 = sum wz g, Rz = sum c wz g
            const uint32_t oab = s.ox[a] + s.oy[b];
            float4 g0 = g_out[oab + s.oz[0]], g1 = g_out[oab + s.oz[1]], g2 = g_out[oab + s.oz[2]];
            float w1 = s.wz[1], w2 = s.wz[2], w0 = s.wz[0];
            V3 R0 = v3(fmaf(w2, g2.x, fmaf(w1, g1.x, w0 * g0.x)), fmaf(w2, g2.y, fmaf(w1, g1.y, w0 * g0.y)), fmaf(w2, g2.z, fmaf(w1, g1.z, w0 * g0.z)));
            V3 Rz = v3(fmaf(2.f * w2, g2.x, w1 * g1.x), fmaf(2.f * w2, g2.y, w1 * g1.y), fmaf(2.f * w2, g2.z, w1 * g1.z));
            float wab = s.wx[a] * s.wy[b];
            V3 r0 = wab * R0;
            nv += r0;
            if (a) { B.m[0] += a * r0.x; B.m[3] += a * r0.y; B.m[6] += a * r0.z; }
            if (b) { B.m[1] += b * r0.x; B.m[4] += b * r0.y; B.m[7] += b * r0.z; }
            B.m[2] = fmaf(wab, Rz.x, B.m[2]); B.m[5] = fmaf(wab, Rz.y, B.m[5]); B.m[8] = fmaf(wab, Rz.z, B.m[8]);
        }
    float k4 = 4.f * P.inv_dx;
    float f3[3] = {s.fx, s.fy, s.fz}, n3[3] = {nv.x, nv.y, nv.z};
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) Cn.m[3 * r + c] = k4 * (B.m[3 * r + c] - n3[r] * f3[c]);
}

// ------------------------------------------------------------------------------------------------
// P2G: F_tmp, SVD, plasticity, stress, APIC scatter.  One thread per particle slot.
// FUSED: the G2P of the PREVIOUS substep runs first in the same thread ("G2P2G"): x, v, C of frame f are produced from
// frame f-1 and g_out, written to the checkpoint, and consumed from registers (fprev / g_prev non-null).
// STAGED: warp-aggregated scatter through shared memory (default); otherwise one REDG per node.
// ------------------------------------------------------------------------------------------------
template <int MAT, bool STAGED, bool EXTRA>
__global__ void __launch_bounds__(SMX_TPB_SC, SMX_SC_MINB) k_p2g(Params P, PrimSet ps, int f, float* __restrict__ fin, float* __restrict__ fout,
                                                    float4* __restrict__ g_in, const int* __restrict__ ctrl_slot,
                                                    const float* __restrict__ action, int accumulate,
                                                    const float* __restrict__ fprev, const float4* __restrict__ g_prev, float4* __restrict__ rec, int pf_dist) {
    pdl_prologue();
    stagger_first_wave(P.dbg, SMX_SC_MINB);
    // EXTRA: particle-contact impulses (collision_type == 1) and / or particle control forces are present
    extern __shared__ __align__(128) float4 smx_dyn_smem[];     // STAGED: one WarpStage per warp (more than the 48 KB static limit)
    WarpStage* stage = reinterpret_cast<WarpStage*>(smx_dyn_smem);
    constexpr bool has_svd = (mat_model(MAT) == 0) && (mat_ptype(MAT) != 2);
    int j = blockIdx.x * SMX_TPB_SC + threadIdx.x;
    bool live = j < P.n;
    int jj = live ? j : P.n - 1;
    if (!fprev) prefetch_planes(fin, P.stride, (long long)j + pf_dist, P.n, 0, SMX_NPLANES);
    else { prefetch_planes(fprev, P.stride, (long long)j + pf_dist, P.n, 0, 1); prefetch_planes(fin, P.stride, (long long)j + pf_dist, P.n, 3, SMX_NPLANES); }
    V3 x, v; M3 F, C;
    int bt = batch_of(P, jj);
    if (fprev) {        // fused G2P of substep f-1: frame f-1 -> x, v, C of frame f
        V3 xo = load_x(fprev, P.stride, jj);
        {   // F of frame f was written by the previous P2G launch (plain loads: this kernel writes the same planes); issued before
            // the gather so that the HBM latency overlaps the 27 grid loads
            const float4* b = reinterpret_cast<const float4*>(fin) + jj;
            float f0 = reinterpret_cast<const float*>(b + 3 * P.stride)[3];
            float4 p4 = b[4 * P.stride], p5 = b[5 * P.stride];
            F.m[0] = f0; F.m[1] = p4.x; F.m[2] = p4.y; F.m[3] = p4.z; F.m[4] = p4.w; F.m[5] = p5.x; F.m[6] = p5.y; F.m[7] = p5.z; F.m[8] = p5.w;
        }
        Stencil so = make_stencil(xo.x, xo.y, xo.z, P, bt);
        g2p_gather(P, so, g_prev, v, C);
        x = xo + P.dt * v;
        if (live) store_xvC(fin, P.stride, j, x, v, C);
    } else load_state(fin, P.stride, jj, x, v, F, C);
    V3 imp = v3(0, 0, 0);
    if (EXTRA) imp = particle_impulses(P, ps, f, jj, bt, live, x, v, ctrl_slot, action, accumulate != 0);
    Stencil s = make_stencil(x.x, x.y, x.z, P, bt);
    if (live) {
        Material m;
        material_update<MAT>(compute_Et(C, F, P.dt), P, m);
        if (fout) store_F(fout, P.stride, j, m.newF);
        if (has_svd && rec) store_svd_rec(rec, P.stride, j, m.svd, m.Jm1);
        // affine' = (cs*stress + p_mass*C) * dx ; value(node) = w * (q0 + affine' * offset), q0 = p_mass*v + imp - affine' * fx
        M3 A;
#pragma unroll
        for (int i = 0; i < 9; i++) A.m[i] = (P.cs * m.stress.m[i] + P.p_mass * C.m[i]) * P.dx;
        V3 q0 = P.p_mass * v + imp - mulv(A, v3(s.fx, s.fy, s.fz));
        V3 c0 = v3(A.m[0], A.m[3], A.m[6]), c1 = v3(A.m[1], A.m[4], A.m[7]), c2 = v3(A.m[2], A.m[5], A.m[8]);
        float4* row = STAGED ? stage[threadIdx.x >> 5].val + (threadIdx.x & 31) * 27 : nullptr;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            V3 qa = q0 + (float)a * c0;
#pragma unroll
            for (int b = 0; b < 3; b++) {
                V3 qb = qa + (float)b * c1;
                float wab = s.wx[a] * s.wy[b];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    V3 q = qb + (float)c * c2;
                    float w = wab * s.wz[c];
                    if (STAGED) row[a * 9 + b * 3 + c] = make_float4(w * q.x, w * q.y, w * q.z, w * P.p_mass);
                    else red_add_f4(g_in + (s.ox[a] + s.oy[b] + s.oz[c]), w * q.x, w * q.y, w * q.z, w * P.p_mass);
                }
            }
        }
    }
    if (STAGED) warp_stage_flush(stage[threadIdx.x >> 5], base_node(s), live, g_in, P.nb, P.Gb, P.dbg);
}

// boundary_condition (mpm_simulator.py:268-281); mask bit d cleared where component d was zeroed
__device__ __forceinline__ V3 boundary_condition(int i, int j, int k, V3 v, const Params& P, int& mask) {
    const int bound = 3;
    mask = 7;
    if (i < bound && v.x < 0.f) { v.x = 0.f; mask &= ~1; }
    if (i > P.ng - bound && v.x > 0.f) { v.x = 0.f; mask &= ~1; }
    if (j < bound && v.y < 0.f) { v.y = 0.f; mask &= ~2; }
    if (j > P.ng - bound && v.y > 0.f) { v.y = 0.f; mask &= ~2; }
    if (j < bound && P.sticky) { v = v3(0, 0, 0); mask = 0; }
    if (k < bound && v.z < 0.f) { v.z = 0.f; mask &= ~4; }
    if (k > P.ng - bound && v.z > 0.f) { v.z = 0.f; mask &= ~4; }
    return v;
}
__device__ __forceinline__ void node_coords(uint32_t node, int nb, int& i, int& j, int& k) {
    uint32_t b = node >> 6, l = node & 63;
    int bk = b % nb, bj = (b / nb) % nb, bi = b / (nb * nb);
    i = bi * 4 + (l >> 4); j = bj * 4 + ((l >> 2) & 3); k = bk * 4 + (l & 3);
}

// ------------------------------------------------------------------------------------------------
// grid update: momentum -> velocity, gravity, (grid contact), boundary conditions.
// blocks == nullptr: dense sweep; else one 64-node grid block per 64 threads from the active list.
// Also re-zeroes nothing: inactive nodes are written as zero so g_out / g_mix never need a memset.
// ------------------------------------------------------------------------------------------------
// Fused with it (to keep the launch count down): the grid checkpoint of this substep (rec: g_in always, g_out when
// no contact kernel follows) and the zeroing of g_in for the next substep's P2G.
// GC: grid contact (collision_type == 0) compiled in; the common variants stay lean (few registers: every active block resident at once)
template <bool GC>
__global__ void __launch_bounds__(256) k_grid_op(Params P, PrimSet ps, int f, const uint32_t* __restrict__ blocks, const int* __restrict__ nblocks,
                                                 float4* __restrict__ g_in, float4* __restrict__ g_out, float4* __restrict__ g_mix,
                                                 int accumulate, float4* __restrict__ rec, int cap, int save_out, int zero_in,
                                                 unsigned long long* __restrict__ counters, uint32_t* __restrict__ near_count, int* __restrict__ need = nullptr,
                                                 float4* __restrict__ rec_prev = nullptr) {
    // rec_prev: grid record of the PREVIOUS substep (same ordering): its g_out / g_mix (final after that substep's contact scatter) are
    // saved here, node by node, just before this substep overwrites them -- instead of a separate copy launch
    pdl_prologue();
    if (near_count && blockIdx.x == 0 && threadIdx.x == 0) *near_count = 0u;        // work list of the contact kernel that follows
    int total = blocks ? *nblocks : P.nbatch * P.nb3;
    if (rec && total > cap && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counters + 2, 1ull);   // record does not fit
    if (rec && need && blockIdx.x == 0 && threadIdx.x == 0) *need = total;          // blocks this substep's record needs (host: which substeps fit)
    for (int bi = blockIdx.x * 4 + (threadIdx.x >> 6); bi < total; bi += gridDim.x * 4) {
        uint32_t node = (blocks ? blocks[bi] : (uint32_t)bi) * 64u + (threadIdx.x & 63);
        float4 g = g_in[node];
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        bool on = g.w > 1e-10f;
        int i, j, k;
        int bt = P.nbatch > 1 ? (int)(node / (uint32_t)P.Gb) : 0;
        node_coords(node - (uint32_t)(bt * P.Gb), P.nb, i, j, k);
        V3 v = v3(0, 0, 0);
        if (on) {
            float inv = 1.f / g.w;
            v = v3(inv * g.x + P.dt * P.gx, inv * g.y + P.dt * P.gy, inv * g.z + P.dt * P.gz);
        }
        if (GC && P.ctype == 0) {
            V3 gp = v3(i * P.dx, j * P.dx, k * P.dx);
            for (int q = 0; q < P.np; q++) {
                if (!ps.prims[q].enabled) continue;
                PrimState S = load_prim_state(pstate_at(ps, bt, q, f));
                bool act = false; V3 r = v3(0, 0, 0), vin = v;
                if (on) v = collide_grid_fwd(ps.prims[q], S, gp, v, act, r);
                if (accumulate) commit_wrench(ext_f_at(ps, bt, q), (g.w / P.dt) * (vin - v), r, act, bt);
            }
        }
        if (on) {
            int mask;
            v = boundary_condition(i, j, k, v, P, mask);
            o = make_float4(v.x, v.y, v.z, 1.f);
        }
        if (rec_prev && bi < cap) {
            size_t slot = (size_t)bi * 64 + (threadIdx.x & 63);
            rec_prev[(size_t)cap * 64 + slot] = g_out[node];
            if (g_mix) rec_prev[(size_t)2 * cap * 64 + slot] = g_mix[node];
        }
        g_out[node] = o;
        if (g_mix) g_mix[node] = o;
        if (rec && bi < cap) {
            size_t slot = (size_t)bi * 64 + (threadIdx.x & 63);
            rec[slot] = g;
            if (save_out) rec[(size_t)cap * 64 + slot] = o;
        }
        if (zero_in) g_in[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// forecast contact: gather v_tmp from g_mix, chain collide_mixed over the enabled primitives,
// scatter -2 w (v_tmp - v_tgt) into g_out where the node is active; reduce the wrench per body.
// ------------------------------------------------------------------------------------------------
// eight CTAs per SM (64 registers): nearly every warp leaves after the reach test, so the kernel is bound by the latency of the x load and
// occupancy pays (20.4 -> 18.6 us at 1M particles); the few warps that stay spill 96 bytes
#ifndef SMX_CONTACT_MINB
#define SMX_CONTACT_MINB 8
#endif
__global__ void __launch_bounds__(SMX_TPB, SMX_CONTACT_MINB) k_contact(Params P, PrimSet ps, int f, float life, const float* __restrict__ fin,
                                                     const float4* __restrict__ g_mix, float4* __restrict__ g_out, int accumulate,
                                                     uint32_t* __restrict__ near_mask, int nwords) {
    pdl_prologue();
    int j = blockIdx.x * SMX_TPB + threadIdx.x;
    bool live = j < P.n;
    int jj = live ? j : P.n - 1;
    V3 x = load_x(fin, P.stride, jj);
    int bt = batch_of(P, jj);
    // cheap reject: is the particle within reach of any enabled primitive?
    bool near = false;
    for (int i = 0; i < P.np; i++) {
        if (!ps.prims[i].enabled) continue;
        PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
        near |= live && (prim_sdf(ps.prims[i], S, x) <= 5e-3f);
    }
    const unsigned near_bits = __ballot_sync(0xffffffffu, near);
    // kept per substep for the adjoint kernel: [0] number of warps with a particle in reach, [32 + w] the reach bits of warp w,
    // [32 + nwords + i] the i-th such warp (work list; order irrelevant)
    if (near_mask && (threadIdx.x & 31) == 0) {
        near_mask[32 + (j >> 5)] = near_bits;
        if (near_bits) near_mask[32 + nwords + atomicAdd(near_mask, 1u)] = (uint32_t)(j >> 5);
    }
    if (!near_bits) return;
    Stencil s = make_stencil(x.x, x.y, x.z, P, bt);
    V3 vtmp = v3(0, 0, 0);
    uint32_t onmask = 0;
    if (near) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float4 g = g_mix[s.ox[a] + s.oy[b] + s.oz[c]];
                    float w = s.wx[a] * s.wy[b] * s.wz[c];
                    vtmp.x = fmaf(w, g.x, vtmp.x); vtmp.y = fmaf(w, g.y, vtmp.y); vtmp.z = fmaf(w, g.z, vtmp.z);
                    if (g.w > 0.f) onmask |= 1u << (a * 9 + b * 3 + c);
                }
    }
    V3 vt = vtmp;
    for (int i = 0; i < P.np; i++) {
        if (!ps.prims[i].enabled) continue;
        PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
        CmTape T; T.active = false; T.r = v3(0, 0, 0);
        V3 vin = vt;
        if (near) vt = collide_mixed_fwd(ps.prims[i], S, x, vin, P.dt, life, T);
        if (accumulate) commit_wrench(ext_f_at(ps, bt, i), (P.p_mass / P.dt) * (vin - vt), T.r, near && T.active, bt);
    }
    if (!near) return;
    V3 d = vtmp - vt;
    if (d.x == 0.f && d.y == 0.f && d.z == 0.f) return;
    d = -2.f * d;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (onmask & (1u << (a * 9 + b * 3 + c))) {
                    float w = s.wx[a] * s.wy[b] * s.wz[c];
                    red_add_f4(g_out + (s.ox[a] + s.oy[b] + s.oz[c]), w * d.x, w * d.y, w * d.z, 0.f);
                }
}

// ------------------------------------------------------------------------------------------------
// G2P: gather v, C (APIC) and advect.  Writes x, v, C of frame f+1.
// ------------------------------------------------------------------------------------------------
#ifndef SMX_G2P_MINB
#define SMX_G2P_MINB 8
#endif
__global__ void __launch_bounds__(SMX_TPB, SMX_G2P_MINB) k_g2p(Params P, const float* __restrict__ fin, float* __restrict__ fout, const float4* __restrict__ g_out, int pf_dist) {
    pdl_prologue();
    int j = blockIdx.x * SMX_TPB + threadIdx.x;
    if (j >= P.n) return;
    prefetch_planes(fin, P.stride, (long long)j + pf_dist, P.n, 0, 1);
    V3 x = load_x(fin, P.stride, j);
    Stencil s = make_stencil(x.x, x.y, x.z, P, batch_of(P, j));
    V3 nv; M3 Cn;
    g2p_gather(P, s, g_out, nv, Cn);
    store_xvC(fout, P.stride, j, x + P.dt * nv, nv, Cn);
}

// ------------------------------------------------------------------------------------------------
// adjoint of G2P: scatter d g_out, accumulate d x through the weights.
//   ain  = adjoint of frame f+1 (x 0..2, v 3..5, C 15..23), aout = adjoint of frame f (x written here)
// ------------------------------------------------------------------------------------------------
#ifndef SMX_G2PG_MINB
#define SMX_G2PG_MINB 6     // 18 warps; 7 CTAs (21 warps, 80 registers) measured slower: 80.7 vs 76.2 us
#endif
// one particle of the G2P adjoint: stages (or reduces) d g_out of its 27 nodes and returns the partial d x of frame f
//   gx1 = adjoint of x[f+1], gnv = adjoint of v[f+1] + dt * gx1, gC = adjoint of C[f+1]
template <bool STAGED, bool F4>
__device__ __forceinline__ V3 g2p_grad_particle(const Params& P, const Stencil& s, V3 gx1, V3 gnv, const M3& gC, const float4* __restrict__ g_out,
                                                float4* __restrict__ gg_out, float4* row, float2* row_xy, float* row_z) {
    float dwx[3], dwy[3], dwz[3];
    axis_dweights(s.fx, dwx); axis_dweights(s.fy, dwy); axis_dweights(s.fz, dwz);
    float k4 = 4.f * P.inv_dx;
    // d g_out(node) = w * (gnv + k4 * gC * (offset - fx)) = w * (q0 + k4*gC*offset)
    M3 K = scale(k4, gC);
    V3 q0 = gnv - mulv(K, v3(s.fx, s.fy, s.fz));
    V3 c0 = v3(K.m[0], K.m[3], K.m[6]), c1 = v3(K.m[1], K.m[4], K.m[7]), c2 = v3(K.m[2], K.m[5], K.m[8]);
    V3 gfx = v3(0, 0, 0), S0 = v3(0, 0, 0);     // S0 = sum w * g (for d dpos)
#pragma unroll
    for (int a = 0; a < 3; a++) {
        V3 qa = q0 + (float)a * c0;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            V3 qb = qa + (float)b * c1;
            float wab = s.wx[a] * s.wy[b];
            float G0 = 0.f, G1 = 0.f;       // sum_c gw wz[c], sum_c gw dwz[c]
            V3 R0 = v3(0, 0, 0);            // sum_c wz[c] g
#pragma unroll
            for (int c = 0; c < 3; c++) {
                V3 q = qb + (float)c * c2;
                uint32_t node = s.ox[a] + s.oy[b] + s.oz[c];
                float4 g = g_out[node];
                float w = wab * s.wz[c];
                if (STAGED) {
                    if (F4) row[a * 9 + b * 3 + c] = make_float4(w * q.x, w * q.y, w * q.z, 0.f);
                    else { row_xy[a * 9 + b * 3 + c] = make_float2(w * q.x, w * q.y); row_z[a * 9 + b * 3 + c] = w * q.z; }
                } else red_add_f4(gg_out + node, w * q.x, w * q.y, w * q.z, 0.f);
                float gw = fmaf(g.x, q.x, fmaf(g.y, q.y, g.z * q.z));       // d weight
                G0 = fmaf(gw, s.wz[c], G0); G1 = fmaf(gw, dwz[c], G1);
                R0.x = fmaf(s.wz[c], g.x, R0.x); R0.y = fmaf(s.wz[c], g.y, R0.y); R0.z = fmaf(s.wz[c], g.z, R0.z);
            }
            gfx.x = fmaf(dwx[a] * s.wy[b], G0, gfx.x);
            gfx.y = fmaf(s.wx[a] * dwy[b], G0, gfx.y);
            gfx.z = fmaf(wab, G1, gfx.z);
            S0 += wab * R0;
        }
    }
    // d dpos = k4 * w * gC^T g  ->  d fx -= sum = K^T S0
    gfx -= Tmulv(K, S0);
    return gx1 + P.inv_dx * gfx;
}

#ifdef SMX_G2PG_F4
#define SMX_G2PG_F4_ON true
#else
#define SMX_G2PG_F4_ON false
#endif
template <bool STAGED>
__global__ void __launch_bounds__(SMX_TPB_G2PG, SMX_G2PG_MINB) k_g2p_grad(Params P, const float* __restrict__ fin, const float* __restrict__ ain,
                                                         float* __restrict__ aout, const float4* __restrict__ g_out, float4* __restrict__ gg_out, int pf_dist) {
    pdl_prologue();
    constexpr bool F4 = SMX_G2PG_F4_ON;
#ifdef SMX_G2PG_F4
    extern __shared__ __align__(128) float4 smx_dyn_smem[];
    WarpStage* stage = reinterpret_cast<WarpStage*>(smx_dyn_smem);
#else
    __shared__ WarpStage3 stage[STAGED ? SMX_TPB_G2PG / 32 : 1];
#endif
    int j = blockIdx.x * SMX_TPB_G2PG + threadIdx.x;
    bool live = j < P.n;
    int jj = live ? j : P.n - 1;
    prefetch_planes(fin, P.stride, (long long)j + pf_dist, P.n, 0, 1);
    prefetch_planes(ain, P.stride, (long long)j + pf_dist, P.n, 0, 4);
    V3 x = load_x(fin, P.stride, jj);
    int bt = batch_of(P, jj);
    Stencil s = make_stencil(x.x, x.y, x.z, P, bt);
    if (live) {
        V3 gx1, gnv; M3 gC;
        {
            float4 a0 = ld_plane(ain, P.stride, j, 0), a1 = ld_plane(ain, P.stride, j, 1), a2 = ld_plane(ain, P.stride, j, 2), a3 = ld_plane(ain, P.stride, j, 3);
            gx1 = v3(a0.x, a0.y, a0.z);
            gnv = v3(a0.w, a1.x, a1.y) + P.dt * gx1;
            gC.m[0] = a1.z; gC.m[1] = a1.w; gC.m[2] = a2.x; gC.m[3] = a2.y; gC.m[4] = a2.z; gC.m[5] = a2.w; gC.m[6] = a3.x; gC.m[7] = a3.y; gC.m[8] = a3.z;
        }
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#ifdef SMX_G2PG_F4
        V3 gxp = g2p_grad_particle<STAGED, F4>(P, s, gx1, gnv, gC, g_out, gg_out, STAGED ? stage[w].val + l * 27 : nullptr, nullptr, nullptr);
#else
        V3 gxp = g2p_grad_particle<STAGED, F4>(P, s, gx1, gnv, gC, g_out, gg_out, nullptr, STAGED ? stage[w].xy + l * 27 : nullptr, STAGED ? stage[w].z + l * 27 : nullptr);
#endif
        // partial d x of frame f (the contact adjoint and P2G adjoint add theirs); the rest of the plane is written by P2G adjoint
        st_plane(aout, P.stride, j, 0, make_float4(gxp.x, gxp.y, gxp.z, 0.f));
    }
#ifdef SMX_G2PG_F4
    if (STAGED) warp_stage_flush(stage[threadIdx.x >> 5], base_node(s), live, gg_out, P.nb, P.Gb, P.dbg);
#else
    if (STAGED) warp_stage_flush3(stage[threadIdx.x >> 5], base_node(s), live, gg_out, P.nb, P.Gb, P.dbg);
#endif
}

// ------------------------------------------------------------------------------------------------
// adjoint of the forecast contact (mixed4.grad, mixed3.grad, mixed2.grad fused).
// Particles that are not within reach of a primitive contribute exactly zero and exit early.
// ------------------------------------------------------------------------------------------------
// one storage slot j of the contact adjoint; called by all 32 lanes of a warp together (warp-level reductions inside)
__device__ __forceinline__ void contact_grad_slot(const Params& P, const PrimSet& ps, int f, float life, const float* __restrict__ fin,
                                                  float* __restrict__ aout, const float4* __restrict__ g_mix, const float4* __restrict__ gg_out,
                                                  float4* __restrict__ gg_mix, int j, bool live, int jj, V3 x, int bt, bool near,
                                                  uint8_t* __restrict__ mixflag = nullptr) {
    // mixflag: one byte per 4x4x4 block, set for every block this substep's contact adjoint scatters into; k_grid_grad reads and re-zeroes
    // gg_mix only there (a full sweep of the otherwise untouched gg_mix cost 17 us per adjoint substep at 1M particles)
    Stencil s = make_stencil(x.x, x.y, x.z, P, bt);
    float dwx[3], dwy[3], dwz[3];
    axis_dweights(s.fx, dwx); axis_dweights(s.fy, dwy); axis_dweights(s.fz, dwz);
    V3 vtmp = v3(0, 0, 0);
    uint32_t onmask = 0;
    if (near) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float4 g = g_mix[s.ox[a] + s.oy[b] + s.oz[c]];
                    float w = s.wx[a] * s.wy[b] * s.wz[c];
                    vtmp.x = fmaf(w, g.x, vtmp.x); vtmp.y = fmaf(w, g.y, vtmp.y); vtmp.z = fmaf(w, g.z, vtmp.z);
                    if (g.w > 0.f) onmask |= 1u << (a * 9 + b * 3 + c);
                }
    }
    // forward chain with tape
    V3 vin[SMX_MAXP], vout[SMX_MAXP];
    CmTape tape[SMX_MAXP];
    int which[SMX_MAXP], nq = 0;
    V3 vt = vtmp;
    for (int i = 0; i < P.np; i++) {
        if (!ps.prims[i].enabled) continue;
        PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
        vin[nq] = vt; tape[nq].active = false;
        if (near) vt = collide_mixed_fwd(ps.prims[i], S, x, vt, P.dt, life, tape[nq]);
        vout[nq] = vt; which[nq] = i; nq++;
    }
    // mixed4.grad: d(v_tmp - v_tgt) = -2 sum_on w * d g_out ; d weight = -2 d g_out . (v_tmp - v_tgt)
    V3 d = vtmp - vt, gd = v3(0, 0, 0), gfx = v3(0, 0, 0);
    if (near) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    if (onmask & (1u << (a * 9 + b * 3 + c))) {
                        float4 go = gg_out[s.ox[a] + s.oy[b] + s.oz[c]];
                        float w = s.wx[a] * s.wy[b] * s.wz[c];
                        gd.x = fmaf(-2.f * w, go.x, gd.x); gd.y = fmaf(-2.f * w, go.y, gd.y); gd.z = fmaf(-2.f * w, go.z, gd.z);
                        float gw = -2.f * (go.x * d.x + go.y * d.y + go.z * d.z);
                        gfx.x = fmaf(gw, dwx[a] * s.wy[b] * s.wz[c], gfx.x);
                        gfx.y = fmaf(gw, s.wx[a] * dwy[b] * s.wz[c], gfx.y);
                        gfx.z = fmaf(gw, s.wx[a] * s.wy[b] * dwz[c], gfx.z);
                    }
    }
    // mixed3.grad: chain in reverse
    V3 g = -gd;             // d v_tgt
    V3 gxc = v3(0, 0, 0);   // d x from the contact model
    for (int a = nq - 1; a >= 0; a--) {
        int i = which[a];
        PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
        PrimGrad G = prim_grad_zero();
        V3 gin = v3(0, 0, 0);
        bool act = near && tape[a].active;
        if (near) collide_mixed_adj(ps.prims[i], S, x, vin[a], vout[a], P.p_mass, P.dt, life, tape[a], g, ext_f_grad_at(ps, bt, i), gxc, gin, G);
        commit_prim_grad(pgrad_at(ps, bt, i, f), G, act, bt);
        g = gin;
    }
    if (!near) return;
    V3 gvtmp = gd + g;      // d v_tmp
    // mixed2.grad: scatter w * d v_tmp into d g_mix ; d weight = g_mix . d v_tmp
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                uint32_t node = s.ox[a] + s.oy[b] + s.oz[c];
                float4 gm = g_mix[node];
                float w = s.wx[a] * s.wy[b] * s.wz[c];
                red_add_f4(gg_mix + node, w * gvtmp.x, w * gvtmp.y, w * gvtmp.z, 0.f);
                if (mixflag && a != 1 && b != 1 && c != 1) mixflag[node >> 6] = 1;      // every block the 3x3x3 stencil touches holds one of its 8 corners
                float gw = gm.x * gvtmp.x + gm.y * gvtmp.y + gm.z * gvtmp.z;
                gfx.x = fmaf(gw, dwx[a] * s.wy[b] * s.wz[c], gfx.x);
                gfx.y = fmaf(gw, s.wx[a] * dwy[b] * s.wz[c], gfx.y);
                gfx.z = fmaf(gw, s.wx[a] * s.wy[b] * dwz[c], gfx.z);
            }
    {
        float4* ap = reinterpret_cast<float4*>(aout) + j;
        float4 a = *ap;
        a.x += gxc.x + P.inv_dx * gfx.x; a.y += gxc.y + P.inv_dx * gfx.y; a.z += gxc.z + P.inv_dx * gfx.z;
        *ap = a;
    }
}

// dense form: one thread per slot, the reach test is repeated (used when no reach bits were recorded for the substep)
__global__ void __launch_bounds__(SMX_TPB) k_contact_grad(Params P, PrimSet ps, int f, float life, const float* __restrict__ fin,
                                                          float* __restrict__ aout, const float4* __restrict__ g_mix,
                                                          const float4* __restrict__ gg_out, float4* __restrict__ gg_mix, uint8_t* __restrict__ mixflag) {
    pdl_prologue();
    int j = blockIdx.x * SMX_TPB + threadIdx.x;
    bool live = j < P.n;
    int jj = live ? j : P.n - 1;
    V3 x = load_x(fin, P.stride, jj);
    int bt = batch_of(P, jj);
    bool near = false;
    for (int i = 0; i < P.np; i++) {
        if (!ps.prims[i].enabled) continue;
        PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
        near |= live && (prim_sdf(ps.prims[i], S, x) <= 5e-3f);
    }
    if (!__any_sync(0xffffffffu, near)) return;
    contact_grad_slot(P, ps, f, life, fin, aout, g_mix, gg_out, gg_mix, j, live, jj, x, bt, near, mixflag);
}
// sparse form: persistent warps walk the work list that k_contact recorded for this substep and only visit the warps' worth of slots
// that have a particle within reach of a primitive -- a handful in a typical scene, where the dense form pays 168 registers x 7813 CTAs
// of launch latency at 1M particles (30 us for 0.8 M warp instructions)
__global__ void __launch_bounds__(SMX_TPB) k_contact_grad_sparse(Params P, PrimSet ps, int f, float life, const float* __restrict__ fin,
                                                                 float* __restrict__ aout, const float4* __restrict__ g_mix,
                                                                 const float4* __restrict__ gg_out, float4* __restrict__ gg_mix,
                                                                 const uint32_t* __restrict__ near_mask, int nwords, uint8_t* __restrict__ mixflag) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * SMX_TPB + threadIdx.x) >> 5, nwarps = (gridDim.x * SMX_TPB) >> 5;
    const int count = (int)min(near_mask[0], (uint32_t)nwords);
    for (int i = warp; i < count; i += nwarps) {
        const uint32_t w = near_mask[32 + nwords + i];
        const unsigned bits = near_mask[32 + w];
        const int j = (int)w * 32 + lane;
        const bool live = j < P.n;
        const int jj = live ? j : P.n - 1;
        V3 x = load_x(fin, P.stride, jj);
        contact_grad_slot(P, ps, f, life, fin, aout, g_mix, gg_out, gg_mix, j, live, jj, x, batch_of(P, jj), live && ((bits >> lane) & 1u), mixflag);
    }
}

// ------------------------------------------------------------------------------------------------
// adjoint of the grid update (grid_op.grad / grid_op_mixed1.grad): gg_out <- (d g_in xyz, d mass) in place.
// ------------------------------------------------------------------------------------------------
// Fused with it (launch count): rec_in != nullptr reads g_in straight from the grid checkpoint of substep f, and rec_prev != nullptr
// prepares the adjoint of substep f-1 (same ordering): g_out (/ g_mix) <- its checkpoint, the OTHER adjoint grid gg_next <- 0
// (gg_out is double-buffered by substep parity because the P2G adjoint of substep f still reads this one), gg_mix <- 0 in place.
template <bool GC>
__global__ void __launch_bounds__(256) k_grid_grad(Params P, PrimSet ps, int f, const uint32_t* __restrict__ blocks, const int* __restrict__ nblocks,
                                                   const float4* __restrict__ g_in, float4* __restrict__ gg_out, float4* __restrict__ gg_mix,
                                                   const float4* __restrict__ rec_in, const float4* __restrict__ rec_prev, int cap, int prev_mix,
                                                   float4* __restrict__ g_out, float4* __restrict__ g_mix, float4* __restrict__ gg_next,
                                                   const uint8_t* __restrict__ mix_touched = nullptr, uint8_t* __restrict__ mix_other = nullptr) {
    // mix_touched: per-block flags written by this substep's contact adjoint (nullptr: every block may hold something); mix_other: the flags
    // of the other substep parity, consumed by the previous launch of this kernel and reset here for the next contact adjoint
    pdl_prologue();
    int total = blocks ? *nblocks : P.nbatch * P.nb3;
    for (int bi = blockIdx.x * 4 + (threadIdx.x >> 6); bi < total; bi += gridDim.x * 4) {
        uint32_t node = (blocks ? blocks[bi] : (uint32_t)bi) * 64u + (threadIdx.x & 63);
        const size_t slot = (size_t)bi * 64 + (threadIdx.x & 63);
        float4 g = rec_in ? rec_in[slot] : g_in[node];
        if (rec_prev) {         // records are only used when every active block fits (bi < cap)
            g_out[node] = rec_prev[(size_t)cap * 64 + slot];
            if (prev_mix) g_mix[node] = rec_prev[(size_t)2 * cap * 64 + slot];
            gg_next[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        bool on = g.w > 1e-10f;
        int i, j, k;
        int bt = P.nbatch > 1 ? (int)(node / (uint32_t)P.Gb) : 0;
        node_coords(node - (uint32_t)(bt * P.Gb), P.nb, i, j, k);
        float4 go = gg_out[node];
        V3 gv = v3(go.x, go.y, go.z);
        if (gg_mix) {
            if (!mix_touched || mix_touched[node >> 6]) {
                float4 gm = gg_mix[node]; gv += v3(gm.x, gm.y, gm.z);
                if (rec_prev) gg_mix[node] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (mix_other && (threadIdx.x & 63) == 0) mix_other[node >> 6] = 0;
        }
        float inv = on ? 1.f / g.w : 0.f;
        V3 v = v3(inv * g.x + P.dt * P.gx, inv * g.y + P.dt * P.gy, inv * g.z + P.dt * P.gz);
        float gmass = 0.f;
        if (GC && P.ctype == 0) {
            V3 gp = v3(i * P.dx, j * P.dx, k * P.dx);
            V3 vins[SMX_MAXP]; int which[SMX_MAXP], nq = 0;
            for (int q = 0; q < P.np; q++) {
                if (!ps.prims[q].enabled) continue;
                PrimState S = load_prim_state(pstate_at(ps, bt, q, f));
                vins[nq] = v; which[nq++] = q;
                bool act; V3 r;
                if (on) v = collide_grid_fwd(ps.prims[q], S, gp, v, act, r);
            }
            int mask = 7;
            if (on) boundary_condition(i, j, k, v, P, mask);
            gv = v3((mask & 1) ? gv.x : 0.f, (mask & 2) ? gv.y : 0.f, (mask & 4) ? gv.z : 0.f);
            for (int a = nq - 1; a >= 0; a--) {
                int q = which[a];
                PrimState S = load_prim_state(pstate_at(ps, bt, q, f));
                PrimGrad G = prim_grad_zero();
                V3 gin = v3(0, 0, 0);
                if (on) collide_grid_adj(ps.prims[q], S, gp, vins[a], P.dt, g.w, gv, ext_f_grad_at(ps, bt, q), gin, gmass, G);
                commit_prim_grad(pgrad_at(ps, bt, q, f), G, on, bt);
                gv = gin;
            }
        } else {
            int mask = 7;
            if (on) boundary_condition(i, j, k, v, P, mask);
            gv = v3((mask & 1) ? gv.x : 0.f, (mask & 2) ? gv.y : 0.f, (mask & 4) ? gv.z : 0.f);
        }
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) {
            float dotv = g.x * gv.x + g.y * gv.y + g.z * gv.z;
            o = make_float4(inv * gv.x, inv * gv.y, inv * gv.z, gmass - dotv * inv * inv);
        }
        gg_out[node] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// adjoint of P2G + svd_grad + compute_F_tmp.grad.  gg = (d g_in xyz, d mass) from k_grid_grad.
//   ain  = adjoint of frame f+1 (only F, comps 6..14, is read here)
//   aout = adjoint of frame f: x holds the partial from g2p/contact grads; v, F, C are written
// ------------------------------------------------------------------------------------------------
#ifndef SMX_P2GG_MINB
#define SMX_P2GG_MINB 4
#endif
// ---- TMA bulk copies (cp.async.bulk, 1-D) signalled on an mbarrier: staging of the streaming particle planes ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }

// REC:   U, V, sigma - 1 and J - 1 come from the SVD record written by the forward P2G (no Jacobi sweeps here)
// EXTRA: particle-contact impulses (collision_type == 1) and / or particle control forces are present
// SM:    the streaming planes of this particle were staged in shared memory (k_p2g_grad_tiled); otherwise they are read from HBM
// fb / rb / ab: per-particle base pointers of the frame, the SVD record and the adjoint of frame f+1 (plane p at base[p * stride])
// FUSE:  the adjoints of x, v, C of frame f are returned in registers (gxo, gvo, gCo) for the G2P adjoint of substep f-1 that follows in
//        the same thread; only the planes that carry the adjoint of F (3, 4, 5) are stored
template <int MAT, bool REC, bool EXTRA, bool SM, bool FUSE = false>
__device__ __forceinline__ void p2g_grad_particle(const Params& P, const PrimSet& ps, int f, int j, int jj, bool live, const float4* fb, long long fs,
                                                  const float4* rb, long long rs, const float4* ab, long long astr, float4 gx_part, float* __restrict__ aout,
                                                  const float4* __restrict__ gg, const int* __restrict__ ctrl_slot, const float* __restrict__ action,
                                                  double* __restrict__ action_grad, V3* gxo = nullptr, V3* gvo = nullptr, M3* gCo = nullptr) {
    constexpr int model = mat_model(MAT), ptype = mat_ptype(MAT);
    constexpr bool corot = model == 0 && ptype != 2;
    V3 x, v; M3 F, C;
    load_state_b<SM>(fb, fs, x, v, F, C);
    Material m;
    if (REC && corot) load_svd_rec_b<SM>(rb, rs, m.svd, m.Jm1);
    // every streaming load is issued up front (they are all in flight together); the adjoint inputs are consumed last
    M3 gnewF;
    if (!SM) gnewF = load_F_b<false>(ab, astr);
    int bt = batch_of(P, jj);
    V3 imp = v3(0, 0, 0);
    if (EXTRA) imp = particle_impulses(P, ps, f, jj, bt, live, x, v, ctrl_slot, action, false);
    M3 Et = compute_Et(C, F, P.dt);
    material_update<MAT, REC>(Et, P, m);
    M3 A;
#pragma unroll
    for (int i = 0; i < 9; i++) A.m[i] = (P.cs * m.stress.m[i] + P.p_mass * C.m[i]) * P.dx;
    Stencil s = make_stencil(x.x, x.y, x.z, P, bt);
    float dwx[3], dwy[3], dwz[3];
    axis_dweights(s.fx, dwx); axis_dweights(s.fy, dwy); axis_dweights(s.fz, dwz);
    V3 q0 = P.p_mass * v + imp - mulv(A, v3(s.fx, s.fy, s.fz));
    V3 c0 = v3(A.m[0], A.m[3], A.m[6]), c1 = v3(A.m[1], A.m[4], A.m[7]), c2 = v3(A.m[2], A.m[5], A.m[8]);
    V3 S0 = v3(0, 0, 0), gfx = v3(0, 0, 0);
    M3 S1 = m3_zero();      // sum w * g (x) offset
#pragma unroll
    for (int a = 0; a < 3; a++) {
        V3 qa = q0 + (float)a * c0;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            V3 qb = qa + (float)b * c1;
            float wab = s.wx[a] * s.wy[b];
            const uint32_t oab = s.ox[a] + s.oy[b];
            float G0 = 0.f, G1 = 0.f;       // sum_c gw wz[c], sum_c gw dwz[c]
            V3 R0 = v3(0, 0, 0), Rz = v3(0, 0, 0);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                V3 q = qb + (float)c * c2;      // value scattered to this node, per unit weight
                float4 g = gg[oab + s.oz[c]];
                float gw = fmaf(g.x, q.x, fmaf(g.y, q.y, fmaf(g.z, q.z, g.w * P.p_mass)));
                G0 = fmaf(gw, s.wz[c], G0); G1 = fmaf(gw, dwz[c], G1);
                float wz = s.wz[c];
                R0.x = fmaf(wz, g.x, R0.x); R0.y = fmaf(wz, g.y, R0.y); R0.z = fmaf(wz, g.z, R0.z);
                if (c) { float cw = (float)c * wz; Rz.x = fmaf(cw, g.x, Rz.x); Rz.y = fmaf(cw, g.y, Rz.y); Rz.z = fmaf(cw, g.z, Rz.z); }
            }
            gfx.x = fmaf(dwx[a] * s.wy[b], G0, gfx.x);
            gfx.y = fmaf(s.wx[a] * dwy[b], G0, gfx.y);
            gfx.z = fmaf(wab, G1, gfx.z);
            V3 r0 = wab * R0;
            S0 += r0;
            if (a) { S1.m[0] += a * r0.x; S1.m[3] += a * r0.y; S1.m[6] += a * r0.z; }
            if (b) { S1.m[1] += b * r0.x; S1.m[4] += b * r0.y; S1.m[7] += b * r0.z; }
            S1.m[2] = fmaf(wab, Rz.x, S1.m[2]); S1.m[5] = fmaf(wab, Rz.y, S1.m[5]); S1.m[8] = fmaf(wab, Rz.z, S1.m[8]);
        }
    }
    // d affine = sum w g (x) dpos = dx * (S1 - S0 (x) fx);  d fx -= dx * affine^T S0 = A^T S0
    gfx -= Tmulv(A, S0);
    M3 gaff;
    {
        float f3[3] = {s.fx, s.fy, s.fz}, s3[3] = {S0.x, S0.y, S0.z};
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) gaff.m[3 * r + c] = P.dx * (S1.m[3 * r + c] - s3[r] * f3[c]);
    }
    V3 gx = P.inv_dx * gfx, gv = P.p_mass * S0, gimp = S0;
    M3 gC = scale(P.p_mass, gaff);
    M3 gstress = scale(P.cs, gaff);
    if (SM) {       // staged in shared memory: read the late inputs only now, and F, C again (not carried through the gather loop)
        float4 p1 = lds4_again(fb + fs), p2 = lds4_again(fb + 2 * fs), p3 = lds4_again(fb + 3 * fs), p4 = lds4_again(fb + 4 * fs), p5 = lds4_again(fb + 5 * fs);
        C.m[0] = p1.z; C.m[1] = p1.w; C.m[2] = p2.x; C.m[3] = p2.y; C.m[4] = p2.z; C.m[5] = p2.w; C.m[6] = p3.x; C.m[7] = p3.y; C.m[8] = p3.z;
        F.m[0] = p3.w; F.m[1] = p4.x; F.m[2] = p4.y; F.m[3] = p4.z; F.m[4] = p4.w; F.m[5] = p5.x; F.m[6] = p5.y; F.m[7] = p5.z; F.m[8] = p5.w;
        gnewF = load_F_b<true>(ab, astr);
    }
    float tr = trace(gstress), gJ = 0.f;
    M3 gFtmp = m3_zero();
    if (model == 0) {
        gJ = P.lam * (2.f * m.J - 1.f) * tr;
        if (ptype == 2) {
            gJ += (1.f / 3.f) * cbrtf(m.J) / m.J * trace(gnewF);       // d/dJ J^(1/3); mu == 0 so R carries nothing
        } else {
            // stress = 2 mu D newF^T, D = newF - R, newF = U diag(1 + g) V^T, D = U diag(g) V^T (g = clipped sigma - 1).
            // Everything is carried in the singular frame:  Gh = U^T gstress U,  N = U^T gnewF V, and with s = 1 + g
            //   M1 = U^T (gnewF + 2 mu gstress^T D + 2 mu gstress newF) V = N + 2 mu (Gh^T diag(g) + Gh diag(s))
            //   M2 = U^T (d R) V = -2 mu Gh diag(s)
            // followed by the folded svd_grad (mpm_simulator.py:140-157): see DESIGN.md "SVD adjoint in divided-difference form"
            const M3 &U = m.svd.U, &V = m.svd.V;
            const float* e = m.svd.e;
            float g3[3], Dvm[3][3];
            bool yields = false;        // von Mises only: a particle that does not yield keeps new_F = F_tmp (the elastic form below)
            if (mat_vm(MAT)) yields = von_mises_dev(e, P.vm_c, g3, Dvm);
            else {
#pragma unroll
                for (int i = 0; i < 3; i++) g3[i] = ptype == 0 ? fminf(fmaxf(e[i], -2e-3f), 3e-3f) : e[i];
            }
            M3 Gh = mul(Tmul(U, gstress), U);
            M3 M1 = mul(Tmul(U, gnewF), V), M2;
            float mu2 = 2.f * P.mu;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int jx = 0; jx < 3; jx++) {
                    float gs = mu2 * Gh.m[3 * i + jx] * (1.f + g3[jx]);
                    M2.m[3 * i + jx] = -gs;
                    M1.m[3 * i + jx] += fmaf(mu2 * Gh.m[3 * jx + i], g3[jx], gs);
                }
            M3 inner;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                // plastic: min(max(sig, lo), hi): gradient reaches sig iff lo < sig and max(sig, lo) < hi;  elastic: new_F = F_tmp
                if (mat_vm(MAT)) inner.m[4 * i] = yields ? M1.m[0] * Dvm[0][i] + M1.m[4] * Dvm[1][i] + M1.m[8] * Dvm[2][i] : M1.m[4 * i];
                else inner.m[4 * i] = (ptype == 1 || (e[i] > -2e-3f && e[i] < 3e-3f)) ? M1.m[4 * i] : 0.f;
#pragma unroll
                for (int jx = 0; jx < 3; jx++) {
                    if (jx == i) continue;
                    float de = e[jx] - e[i];
                    float K = __fdividef(1.f, clamp_ref(de * (2.f + e[i] + e[jx])));
                    if (ptype == 0 && (!mat_vm(MAT) || yields)) {
                        float dg = g3[jx] - g3[i];
                        float P1 = dg + de + (g3[jx] * e[jx] - g3[i] * e[i]);
                        float Q1 = dg - de + (g3[jx] * e[i] - g3[i] * e[jx]);
                        inner.m[3 * i + jx] = K * (M1.m[3 * i + jx] * P1 + M1.m[3 * jx + i] * Q1 + (M2.m[3 * i + jx] - M2.m[3 * jx + i]) * de);
                    } else {
                        inner.m[3 * i + jx] = M1.m[3 * i + jx] + K * (M2.m[3 * i + jx] - M2.m[3 * jx + i]) * de;
                    }
                }
            }
            gFtmp = mulT(mul(U, inner), V);
        }
    } else {
        M3 sym;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) sym.m[3 * r + c] = gstress.m[3 * r + c] + gstress.m[3 * c + r];
        gnewF = add(gnewF, scale(P.mu, mul(sym, m.newF)));
        gJ = P.lam / m.J * tr;
        if (ptype == 2) gJ += (gnewF.m[0] + gnewF.m[4]) / (2.f * sqrtf(m.J));
        else gFtmp = gnewF;
    }
    {
        if (SM) Et = compute_Et(C, F, P.dt);       // recomputed from the re-read C, F
        M3 Ftmp = Et; Ftmp.m[0] += 1.f; Ftmp.m[4] += 1.f; Ftmp.m[8] += 1.f;
        gFtmp = add(gFtmp, scale(gJ, cofactor(Ftmp)));
    }
    // compute_F_tmp.grad: dC += dt * gFtmp F^T ; dF = (I + dt C)^T gFtmp
    gC = add(gC, scale(P.dt, mulT(gFtmp, F)));
    M3 gF = add(gFtmp, scale(P.dt, Tmul(C, gFtmp)));
    // control and particle-contact adjoints
    if (EXTRA) {
        if (P.n_control > 0 && live) {
            int ci = ctrl_slot[jj];
            if (ci >= 0) {
                ci += bt * P.n_control;
                float k = 6e-4f * P.dt;
                atomicAdd(action_grad + 3 * ci, (double)(k * gimp.x)); atomicAdd(action_grad + 3 * ci + 1, (double)(k * gimp.y));
                atomicAdd(action_grad + 3 * ci + 2, (double)(k * gimp.z));
            }
        }
        if (P.ctype == 1) {
            for (int i = P.np - 1; i >= 0; i--) {
                if (!ps.prims[i].enabled) continue;
                PrimState S = load_prim_state(pstate_at(ps, bt, i, f));
                PrimGrad G = prim_grad_zero();
                if (live) collide_particle_adj(ps.prims[i], S, x, v, P.dt, gimp, ext_f_grad_at(ps, bt, i), gx, gv, G);
                commit_prim_grad(pgrad_at(ps, bt, i, f), G, live, bt);
            }
        }
    }
    if (FUSE) { *gxo = v3(gx_part.x + gx.x, gx_part.y + gx.y, gx_part.z + gx.z); *gvo = gv; *gCo = gC; }
    if (live) {
        float4* b = reinterpret_cast<float4*>(aout) + j;
        if (!FUSE) {
            b[0] = make_float4(gx_part.x + gx.x, gx_part.y + gx.y, gx_part.z + gx.z, gv.x);
            b[P.stride] = make_float4(gv.y, gv.z, gC.m[0], gC.m[1]);
            b[2 * P.stride] = make_float4(gC.m[2], gC.m[3], gC.m[4], gC.m[5]);
        }
        b[3 * P.stride] = make_float4(gC.m[6], gC.m[7], gC.m[8], gF.m[0]);
        b[4 * P.stride] = make_float4(gF.m[1], gF.m[2], gF.m[3], gF.m[4]);
        b[5 * P.stride] = make_float4(gF.m[5], gF.m[6], gF.m[7], gF.m[8]);
    }
}

// one thread per particle slot, planes read straight from HBM (ablation: SMX_FLAG_NO_TMA)
template <int MAT, bool REC, bool EXTRA>
__global__ void __launch_bounds__(SMX_TPB, SMX_P2GG_MINB) k_p2g_grad(Params P, PrimSet ps, int f, const float* __restrict__ fin, const float* __restrict__ ain,
                                                      float* __restrict__ aout, const float4* __restrict__ gg, const int* __restrict__ ctrl_slot,
                                                      const float* __restrict__ action, double* __restrict__ action_grad, const float4* __restrict__ rec, int pf_dist) {
    pdl_prologue();
    constexpr bool corot = (mat_model(MAT) == 0) && (mat_ptype(MAT) != 2);
    int j = blockIdx.x * SMX_TPB + threadIdx.x;
    bool live = j < P.n;
    int jj = live ? j : P.n - 1;
    {
        long long jp = (long long)j + pf_dist;
        prefetch_planes(fin, P.stride, jp, P.n, 0, SMX_NPLANES);
        if (REC && corot) prefetch_planes(rec, P.stride, jp, P.n, 0, SMX_RPLANES);
        prefetch_planes(ain, P.stride, jp, P.n, 3, SMX_NPLANES);
        prefetch_planes(aout, P.stride, jp, P.n, 0, 1);
    }
    const float4 gx_part = reinterpret_cast<const float4*>(aout)[jj];      // partial d x from the G2P / contact adjoints
    p2g_grad_particle<MAT, REC, EXTRA, false>(P, ps, f, j, jj, live, reinterpret_cast<const float4*>(fin) + jj, P.stride, rec + jj, P.stride,
                                              reinterpret_cast<const float4*>(ain) + jj, P.stride, gx_part, aout, gg, ctrl_slot, action, action_grad);
}

// ------------------------------------------------------------------------------------------------
// Backward counterpart of G2P2G: the P2G adjoint of substep f and the G2P adjoint of substep f-1 in ONE thread per particle slot
// (same ordering, no loss seed on frame f).  The adjoints of x, v, C of frame f go from registers straight into the scatter of
// d g_out of substep f-1: they are neither written to nor re-read from HBM (-112 B per particle), and the pair costs one launch.
//   ain  : adjoint buffer of frame f+1 (planes 3.w, 4, 5 = adjoint of F are read); its plane 0 RECEIVES the partial d x of frame
//          f-1 (it is the `aout` of the next adjoint substep)
//   aout : adjoint buffer of frame f: plane 0 holds the partial d x of frame f (read), planes 3, 4, 5 are written
//   g_prev / gg_prev : g_out of substep f-1 (restored by k_grid_grad of substep f) and its cleared adjoint grid
// ------------------------------------------------------------------------------------------------
#ifndef SMX_FUSEB_MINB
#define SMX_FUSEB_MINB 4
#endif
#ifndef SMX_TPB_FB
#define SMX_TPB_FB 128
#endif
template <int MAT, bool REC, bool EXTRA>
__global__ void __launch_bounds__(SMX_TPB_FB, SMX_FUSEB_MINB) k_p2g_grad_g2p_grad(Params P, PrimSet ps, int f, const float* __restrict__ fin, float* ain, float* __restrict__ aout,
                                                                  const float4* __restrict__ gg, const int* __restrict__ ctrl_slot, const float* __restrict__ action,
                                                                  double* __restrict__ action_grad, const float4* __restrict__ rec, const float* __restrict__ fprev,
                                                                  const float4* __restrict__ g_prev, float4* __restrict__ gg_prev, int pf_dist) {
    pdl_prologue();
    stagger_first_wave(P.dbg, SMX_FUSEB_MINB);
    constexpr bool corot = (mat_model(MAT) == 0) && (mat_ptype(MAT) != 2);
    __shared__ WarpStage3 stage[SMX_TPB_FB / 32];
    int j = blockIdx.x * SMX_TPB_FB + threadIdx.x;
    bool live = j < P.n;
    int jj = live ? j : P.n - 1;
    {
        long long jp = (long long)j + pf_dist;
        prefetch_planes(fin, P.stride, jp, P.n, 0, SMX_NPLANES);
        if (REC && corot) prefetch_planes(rec, P.stride, jp, P.n, 0, SMX_RPLANES);
        prefetch_planes(ain, P.stride, jp, P.n, 3, SMX_NPLANES);
        prefetch_planes(aout, P.stride, jp, P.n, 0, 1);
        prefetch_planes(fprev, P.stride, jp, P.n, 0, 1);
    }
    const float4 gx_part = reinterpret_cast<const float4*>(aout)[jj];      // partial d x of frame f from the G2P / contact adjoints of substep f
    const V3 xp = load_x(fprev, P.stride, jj);                             // x of frame f-1 (needed last; issued with the other streaming loads)
    V3 gx, gv; M3 gC;
    p2g_grad_particle<MAT, REC, EXTRA, false, true>(P, ps, f, j, jj, live, reinterpret_cast<const float4*>(fin) + jj, P.stride, rec + jj, P.stride,
                                                    reinterpret_cast<const float4*>(ain) + jj, P.stride, gx_part, aout, gg, ctrl_slot, action, action_grad, &gx, &gv, &gC);
    const int bt = batch_of(P, jj);
    Stencil s = make_stencil(xp.x, xp.y, xp.z, P, bt);
    if (live) {
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        V3 gxp = g2p_grad_particle<true, false>(P, s, gx, gv + P.dt * gx, gC, g_prev, gg_prev, nullptr, stage[w].xy + l * 27, stage[w].z + l * 27);
        st_plane(ain, P.stride, j, 0, make_float4(gxp.x, gxp.y, gxp.z, 0.f));
    }
    warp_stage_flush3(stage[threadIdx.x >> 5], base_node(s), live, gg_prev, P.nb, P.Gb, P.dbg);
}

#ifndef SMX_P2GG_TPB
#define SMX_P2GG_TPB 128            // tile of the persistent P2G adjoint (particle slots per CTA trip)
#endif
#define SMX_P2GG_TILED_MINB (512 / SMX_P2GG_TPB)
// Persistent, software-pipelined variant: every CTA walks tiles of SMX_P2GG_TPB particle slots; while tile i is being processed the
// streaming planes of tile i+1 (frame f: 6, adjoint of F[f+1]: 3, SVD record: 4 -- each a contiguous 2 KB row) are brought into
// the other half of a double buffer by TMA bulk copies (cp.async.bulk) that complete on an mbarrier, so no warp ever waits for HBM.
#define SMX_P2GG_NPL(MAT, REC) ((mat_model(MAT) == 0 && mat_ptype(MAT) != 2 && (REC)) ? 13 : 9)
template <int MAT, bool REC, bool EXTRA>
__global__ void __launch_bounds__(SMX_P2GG_TPB, SMX_P2GG_TILED_MINB) k_p2g_grad_tiled(Params P, PrimSet ps, int f, const float* __restrict__ fin, const float* __restrict__ ain,
                                                            float* __restrict__ aout, const float4* __restrict__ gg, const int* __restrict__ ctrl_slot,
                                                            const float* __restrict__ action, double* __restrict__ action_grad, const float4* __restrict__ rec, int ntiles) {
    pdl_prologue();
    constexpr int NPL = SMX_P2GG_NPL(MAT, REC);
    extern __shared__ __align__(128) float4 smx_dyn_smem[];
    __shared__ uint64_t bar[2];
    float4* buf = smx_dyn_smem;         // [2][NPL][SMX_P2GG_TPB]
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
    __syncthreads();
    auto issue = [&](int tile, int stage) {     // one thread: arm the barrier with the byte count, then one bulk copy per plane row
        const long long base = (long long)tile * SMX_P2GG_TPB;
        const uint32_t bytes = (uint32_t)min((long long)SMX_P2GG_TPB, (long long)P.n - base) * 16u;
        float4* dst = buf + stage * NPL * SMX_P2GG_TPB;
        mbar_expect_tx(&bar[stage], bytes * NPL);
        const float4* f4 = reinterpret_cast<const float4*>(fin) + base;
        const float4* a4 = reinterpret_cast<const float4*>(ain) + base;
#pragma unroll
        for (int p = 0; p < 6; p++) bulk_g2s(dst + p * SMX_P2GG_TPB, f4 + p * P.stride, bytes, &bar[stage]);
#pragma unroll
        for (int p = 0; p < 3; p++) bulk_g2s(dst + (6 + p) * SMX_P2GG_TPB, a4 + (3 + p) * P.stride, bytes, &bar[stage]);
        if (NPL == 13) {
#pragma unroll
            for (int p = 0; p < 4; p++) bulk_g2s(dst + (9 + p) * SMX_P2GG_TPB, rec + base + p * P.stride, bytes, &bar[stage]);
        }
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue(tile, 0);
    for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
        const int stage = it & 1;
        // the other stage was released by the __syncthreads that ended the previous trip
        if (tid == 0 && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, stage ^ 1);
        int j = tile * SMX_P2GG_TPB + tid;
        bool live = j < P.n;
        int jj = live ? j : P.n - 1;
        const float4 gx_part = reinterpret_cast<const float4*>(aout)[jj];  // partial d x from the G2P / contact adjoints (needed last)
        mbar_wait(&bar[stage], (it >> 1) & 1);
        const float4* sb = buf + stage * NPL * SMX_P2GG_TPB + (jj - tile * SMX_P2GG_TPB);
        p2g_grad_particle<MAT, REC, EXTRA, true>(P, ps, f, j, jj, live, sb, SMX_P2GG_TPB, sb + 9 * SMX_P2GG_TPB, SMX_P2GG_TPB, sb + 3 * SMX_P2GG_TPB, SMX_P2GG_TPB, gx_part, aout, gg,
                                                 ctrl_slot, action, action_grad);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// small kernels: keys, permutation, IO, seeds, kinematics
// ------------------------------------------------------------------------------------------------
__global__ void k_keys(Params P, const float* __restrict__ fr, uint32_t* __restrict__ keys, uint32_t* __restrict__ iota, unsigned long long* counters) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    int clamped;
    V3 xq = load_x(fr, P.stride, j);
    keys[j] = (uint32_t)(batch_of(P, j) * P.Gb) + cell_key(xq.x, xq.y, xq.z, P.inv_dx, P.ng, P.nb, &clamped);
    if (iota) iota[j] = j;
    if (clamped && counters) atomicAdd(counters, 1ull);
}
// dst[plane][j] = src[plane][idx[j]] for all six planes
__global__ void k_gather_frame(int n, long long stride, const float* __restrict__ src, float* __restrict__ dst, const uint32_t* __restrict__ idx) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t i = idx[j];
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int c = 0; c < SMX_NPLANES; c++) d4[c * stride + j] = s4[c * stride + i];
}
// dst[c][idx[j]] = src[c][j]  (inverse of the above; used to carry an adjoint back across a re-sort)
__global__ void k_scatter_frame(int n, long long stride, const float* __restrict__ src, float* __restrict__ dst, const uint32_t* __restrict__ idx) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t i = idx[j];
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int c = 0; c < SMX_NPLANES; c++) d4[c * stride + i] = s4[c * stride + j];
}
__global__ void k_compose_perm(int n, const uint32_t* __restrict__ old_perm, const uint32_t* __restrict__ idx, uint32_t* __restrict__ new_perm) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) new_perm[j] = old_perm ? old_perm[idx[j]] : idx[j];
}
// staging (n, ncomp) AoS in particle-id order, get_state columns c0.. <->  frame planes in storage order
__global__ void k_upload(int n, long long stride, const float* __restrict__ aos, int ncomp, int c0, float* __restrict__ fr, const uint32_t* __restrict__ perm, int accumulate) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t p = perm ? perm[j] : j;
    for (int c = 0; c < ncomp; c++) {
        float v = aos[(size_t)p * ncomp + c];
        long long k = comp_index(stride, j, c0 + c);
        if (accumulate) fr[k] += v; else fr[k] = v;
    }
}
// frame components c0, c1, c2 (get_state columns) <- value (F = I of a reset from positions only)
__global__ void k_fill_comp(int n, long long stride, float* __restrict__ fr, int c0, int c1, int c2, float value) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    fr[comp_index(stride, j, c0)] = value; fr[comp_index(stride, j, c1)] = value; fr[comp_index(stride, j, c2)] = value;
}
__global__ void k_download(int n, long long stride, float* __restrict__ aos, int ncomp, int c0, const float* __restrict__ fr, const uint32_t* __restrict__ perm) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t p = perm ? perm[j] : j;
    for (int c = 0; c < ncomp; c++) aos[(size_t)p * ncomp + c] = fr[comp_index(stride, j, c0 + c)];
}
__global__ void k_permute_i32(int n, const int* __restrict__ src, int* __restrict__ dst, const uint32_t* __restrict__ perm) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[perm ? perm[j] : j];
}
__global__ void k_axpy(long long n, float* __restrict__ y, const float* __restrict__ x) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
// dst (n, nd) += src (n, ns) on the first ns columns
__global__ void k_add_cols(int n, float* __restrict__ dst, int nd, const float* __restrict__ src, int ns) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int c = 0; c < ns; c++) dst[(size_t)j * nd + c] += src[(size_t)j * ns + c];
}
// sums over the particles of the adjoint of x and v (planes 0 and 1 of an adjoint frame): the per-rollout gradient summary a rank
// all-reduces (bench.py); out[0..2] x, [3..5] v, [6] |x|^2, [7] |v|^2, [8] count
__global__ void __launch_bounds__(256) k_grad_summary(int n, long long stride, const float* __restrict__ adj, float* __restrict__ out) {
    float a[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const float4 p0 = ld_plane(adj, stride, j, 0), p1 = ld_plane(adj, stride, j, 1);
        a[0] += p0.x; a[1] += p0.y; a[2] += p0.z; a[3] += p0.w; a[4] += p1.x; a[5] += p1.y;
        a[6] += p0.x * p0.x + p0.y * p0.y + p0.z * p0.z; a[7] += p0.w * p0.w + p1.x * p1.x + p1.y * p1.y; a[8] += 1.f;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        if ((threadIdx.x & 31) == 0) atomicAdd(out + i, a[i]);
    }
}
// block-major grid -> (i,j,k) linear order, for tests
__global__ void k_grid_linear(int ng, int nb, const float4* __restrict__ g, float4* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ng * ng * ng) return;
    int k = t % ng, j = (t / ng) % ng, i = t / (ng * ng);
    out[t] = g[node_index(i, j, k, nb)];
}
// active-block bookkeeping: flag every 4^3 block a particle's stencil (+- margin cells) can touch
__global__ void k_mark_blocks(Params P, const float* __restrict__ fr, int margin, uint32_t* __restrict__ flags) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    int b[3];
    V3 xq = load_x(fr, P.stride, j);
    b[0] = clampi((int)(xq.x * P.inv_dx - 0.5f), 0, P.ng - 3);
    b[1] = clampi((int)(xq.y * P.inv_dx - 0.5f), 0, P.ng - 3);
    b[2] = clampi((int)(xq.z * P.inv_dx - 0.5f), 0, P.ng - 3);
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = max(b[d] - margin, 0) >> 2; hi[d] = min(b[d] + 2 + margin, P.ng - 1) >> 2; }
    for (int i = lo[0]; i <= hi[0]; i++)
        for (int jj = lo[1]; jj <= hi[1]; jj++)
            for (int k = lo[2]; k <= hi[2]; k++) {
                uint32_t id = (uint32_t)(batch_of(P, j) * P.nb3 + (i * P.nb + jj) * P.nb + k);
                if (!flags[id]) flags[id] = 1u;
            }
}
// same, driven by the SORTED key array: only the first particle of every distinct cell does the marking
__global__ void k_mark_blocks_sorted(Params P, const uint32_t* __restrict__ keys_sorted, int margin, uint32_t* __restrict__ flags) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    uint32_t key = keys_sorted[j];
    if (j > 0 && keys_sorted[j - 1] == key) return;
    int b[3];
    {   // invert node_index: key = block * 64 + local
        uint32_t blk = (key >> 6) % (uint32_t)P.nb3, l = key & 63;
        int bk = blk % P.nb, bj = (blk / P.nb) % P.nb, bi = blk / (P.nb * P.nb);
        b[0] = bi * 4 + (l >> 4); b[1] = bj * 4 + ((l >> 2) & 3); b[2] = bk * 4 + (l & 3);
    }
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = max(b[d] - margin, 0) >> 2; hi[d] = min(b[d] + 2 + margin, P.ng - 1) >> 2; }
    for (int i = lo[0]; i <= hi[0]; i++)
        for (int jj = lo[1]; jj <= hi[1]; jj++)
            for (int k = lo[2]; k <= hi[2]; k++) flags[(key >> 6) / (uint32_t)P.nb3 * P.nb3 + (i * P.nb + jj) * P.nb + k] = 1u;
}
// grid checkpoints: the active blocks of up to three float4 grids <-> a compact per-substep record
// record layout: [array][active-block slot][64 nodes]; `cap` = blocks reserved per array
__global__ void __launch_bounds__(256) k_ckpt_copy(const uint32_t* __restrict__ blocks, const int* __restrict__ nblocks, int nb3, int cap,
                                                    float4* __restrict__ rec, float4* __restrict__ a, float4* __restrict__ b, float4* __restrict__ c, int restore,
                                                    unsigned long long* __restrict__ counters, float4* __restrict__ zero_a = nullptr, float4* __restrict__ zero_b = nullptr,
                                                    int* __restrict__ need = nullptr) {
    pdl_prologue();
    int total = blocks ? *nblocks : nb3;
    if (need && !restore && blockIdx.x == 0 && threadIdx.x == 0) *need = total;
    if (total > cap) {      // record does not fit: flagged, the host falls back to recomputation (counters[2])
        if (blockIdx.x == 0 && threadIdx.x == 0 && !restore) atomicAdd(counters + 2, 1ull);
        total = cap;
    }
    for (int bi = blockIdx.x * 4 + (threadIdx.x >> 6); bi < total; bi += gridDim.x * 4) {
        uint32_t node = (blocks ? blocks[bi] : (uint32_t)bi) * 64u + (threadIdx.x & 63);
        size_t slot = (size_t)bi * 64 + (threadIdx.x & 63);
        if (zero_a) zero_a[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (zero_b) zero_b[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (restore) {
            if (a) a[node] = rec[slot];
            if (b) b[node] = rec[(size_t)cap * 64 + slot];
            if (c) c[node] = rec[(size_t)2 * cap * 64 + slot];
        } else {
            if (a) rec[slot] = a[node];
            if (b) rec[(size_t)cap * 64 + slot] = b[node];
            if (c) rec[(size_t)2 * cap * 64 + slot] = c[node];
        }
    }
}
// slab decomposition: force every block of x-block column `col` active (the halo columns exchanged with a neighbour
// must be cleared / updated every substep even where no local particle reaches them)
// ------------------------------------------------------------------------------------------------
// Slab halo exchange over peer memory (NVLink / NVSwitch; the same kernels serve several ranks emulated on one device).
// The 2-column halo range of a grid array is H = 2 nb^2 blocks of 64 nodes.  k_halo_push: one warp per halo block; a block that holds
// anything is written straight into the NEIGHBOUR's receive slot (128-bit peer stores) and stamped with the sequence number of this
// exchange -- empty blocks (~85 % of a halo) never travel and are never read -- then the last CTA releases the neighbour's flag
// (system-scope store after a system fence).  k_halo_wait: ONE thread polls the own flag (so that a waiting rank occupies no SM);
// k_halo_add: adds the stamped blocks of the own receive slot into the grid.  Slots alternate with the sequence number: the push of
// exchange q + 2 into slot q & 1 is ordered after the neighbour's add of exchange q through the flags of exchange q + 1.
// MODE 0: push A; MODE 1: push A - B (own contact scatter g_out - g_mix of the forecast contact model).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// both sides of a slab in ONE launch (blockIdx.y = side; a side without a neighbour has n == 0)
struct HaloSides {
    size_t node0[2];            // first node of the 2-column halo range of the side
    float4* rx[2];              // push: the NEIGHBOUR's receive slot;  add: the own receive slot
    uint32_t* stamp[2];         // block stamps of that slot
    unsigned* flag[2];          // push: the neighbour's flag;  wait: the own flag
    unsigned seq1[2];
    unsigned* done[2];          // push: own completion counter of the side
    int on[2];
};
template <int MODE>
__global__ void __launch_bounds__(256) k_halo_push(const float4* __restrict__ A, const float4* __restrict__ B, int H, HaloSides hs) {
    const int side = blockIdx.y;
    if (!hs.on[side]) return;
    const size_t node0 = hs.node0[side];
    float4* __restrict__ peer_rx = hs.rx[side];
    uint32_t* __restrict__ peer_stamp = hs.stamp[side];
    const unsigned seq1 = hs.seq1[side];
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int hb = warp; hb < H; hb += nwarps) {
        const size_t base = node0 + (size_t)hb * 64;
        float4 v0 = A[base + lane], v1 = A[base + 32 + lane];
        if (MODE == 1) {
            const float4 b0 = B[base + lane], b1 = B[base + 32 + lane];
            v0 = make_float4(v0.x - b0.x, v0.y - b0.y, v0.z - b0.z, 0.f); v1 = make_float4(v1.x - b1.x, v1.y - b1.y, v1.z - b1.z, 0.f);
        }
        const bool nz = v0.x != 0.f || v0.y != 0.f || v0.z != 0.f || v0.w != 0.f || v1.x != 0.f || v1.y != 0.f || v1.z != 0.f || v1.w != 0.f;
        if (__any_sync(0xffffffffu, nz)) {
            peer_rx[(size_t)hb * 64 + lane] = v0; peer_rx[(size_t)hb * 64 + 32 + lane] = v1;
            if (lane == 0) peer_stamp[hb] = seq1;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(hs.done[side], 1u) == gridDim.x - 1) {    // last CTA of the side: everything every CTA wrote is fenced
            *hs.done[side] = 0u;
            __threadfence_system();
            st_release_sys(hs.flag[side], seq1);
        }
    }
}
// counters[3] counts exchanges that gave up waiting (neighbour gone): the run is invalid then, but the GPU is not left spinning.
// One thread per side.
__global__ void k_halo_wait(HaloSides hs, unsigned long long* __restrict__ counters, long long timeout_ns) {
    const int side = threadIdx.x;
    if (side > 1 || !hs.on[side]) return;
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_acquire_sys(hs.flag[side]) - hs.seq1[side]) < 0) {
        __nanosleep(200);
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { atomicAdd(counters + 3, 1ull); break; }
    }
}
__global__ void __launch_bounds__(256) k_halo_add(float4* __restrict__ A, int H, HaloSides hs) {
    const int side = blockIdx.y;
    if (!hs.on[side]) return;
    const size_t node0 = hs.node0[side];
    const float4* __restrict__ rx = hs.rx[side];
    const uint32_t* __restrict__ stamp = hs.stamp[side];
    const unsigned seq1 = hs.seq1[side];
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int hb = warp; hb < H; hb += nwarps) {
        if (__ldcg(stamp + hb) != seq1) continue;
        const size_t base = node0 + (size_t)hb * 64;
        const float4* r = rx + (size_t)hb * 64;                 // written by the neighbour: read past L1
        float4 a0 = A[base + lane], a1 = A[base + 32 + lane];
        const float4 r0 = __ldcg(r + lane), r1 = __ldcg(r + 32 + lane);
        A[base + lane] = make_float4(a0.x + r0.x, a0.y + r0.y, a0.z + r0.z, a0.w + r0.w);
        A[base + 32 + lane] = make_float4(a1.x + r1.x, a1.y + r1.y, a1.z + r1.z, a1.w + r1.w);
    }
}
__global__ void k_mark_column(uint32_t* __restrict__ flags, int nb, int col) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nb * nb) flags[(size_t)col * nb * nb + t] = 1u;
}
// slab decomposition: a local particle whose stencil leaves [lo_col, hi_col] (own slab + one halo column per side) would
// scatter into nodes that are not exchanged: counted in counters[1]
__global__ void k_check_slab(Params P, const float* __restrict__ fr, int lo_col, int hi_col, unsigned long long* counters) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    int b0 = clampi((int)(load_x(fr, P.stride, j).x * P.inv_dx - 0.5f), 0, P.ng - 3);
    if ((b0 >> 2) < lo_col || ((b0 + 2) >> 2) > hi_col) atomicAdd(counters + 1, 1ull);
}
__global__ void k_compact_blocks(int nblk, const uint32_t* __restrict__ flags, uint32_t* __restrict__ list, int* __restrict__ count) {
    // order of the list does not matter (each entry is processed independently); one atomic per set flag
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk && flags[b]) list[atomicAdd(count, 1)] = (uint32_t)b;
}
// zero the float4 grid arrays on the active blocks only
__global__ void __launch_bounds__(256) k_clear_blocks(const uint32_t* __restrict__ blocks, const int* __restrict__ nblocks, float4* __restrict__ a,
                                                      float4* __restrict__ b, float4* __restrict__ c) {
    int total = *nblocks;
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int bi = blockIdx.x * 4 + (threadIdx.x >> 6); bi < total; bi += gridDim.x * 4) {
        uint32_t node = blocks[bi] * 64u + (threadIdx.x & 63);
        if (a) a[node] = z;
        if (b) b[node] = z;
        if (c) c[node] = z;
    }
}
// a particle whose stencil touches an inactive block would scatter into stale memory: count it
__global__ void k_check_active(Params P, const float* __restrict__ fr, const uint32_t* __restrict__ flags, unsigned long long* counters) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    V3 xq = load_x(fr, P.stride, j);
    int b0 = clampi((int)(xq.x * P.inv_dx - 0.5f), 0, P.ng - 3), b1 = clampi((int)(xq.y * P.inv_dx - 0.5f), 0, P.ng - 3),
        b2 = clampi((int)(xq.z * P.inv_dx - 0.5f), 0, P.ng - 3);
    bool ok = true;
    for (int i = b0 >> 2; i <= (b0 + 2) >> 2; i++)
        for (int jj = b1 >> 2; jj <= (b1 + 2) >> 2; jj++)
            for (int k = b2 >> 2; k <= (b2 + 2) >> 2; k++) ok &= flags[batch_of(P, j) * P.nb3 + (i * P.nb + jj) * P.nb + k] != 0;
    if (!ok) atomicAdd(counters + 1, 1ull);
}
// bulk coupling for batched handles: one launch writes / reads the state series of every (batch, primitive)
__global__ void k_fill_prim_states(float* __restrict__ pstate, const float* __restrict__ staged, int T, int np, int nbatch, int f0, int f1) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (batch, primitive, frame)
    int nf = f1 - f0;
    if (t >= nbatch * np * nf) return;
    int f = f0 + t % nf, bi = t / nf, i = bi % np, b = bi / np;
    const float* src = staged + (size_t)(b * np + i) * 13;
    float* dst = pstate + (((size_t)b * SMX_MAXP + i) * T + f) * 13;
#pragma unroll
    for (int c = 0; c < 13; c++) dst[c] = src[c];
}
__global__ void k_sum_prim_grads(const double* __restrict__ pgrad, double* __restrict__ out, int T, int np, int nbatch, int f0, int f1) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (batch, primitive, component)
    if (t >= nbatch * np * 13) return;
    int c = t % 13, bi = t / 13, i = bi % np, b = bi / np;
    const double* src = pgrad + (((size_t)b * SMX_MAXP + i) * T) * 13 + c;
    double acc = 0;
    for (int f = f0; f < f1; f++) acc += src[(size_t)f * 13];
    out[t] = acc;
}
// ------------------------------------------------------------------------------------------------
// Chamfer loss seeds on the device (softmac/engine/losses/loss_grip.py:45-68, GripLoss with weight (1,0,0)):
//   loss = sum_i |x_i - t_nn(i)|^2 + sum_j |x_nn(j) - t_j|^2, nearest neighbours found by brute force with the reference's
//   tie rule (strict <, scanning in index order => smallest index wins), held fixed in the gradient.
// pass 0: one thread per particle slot, targets tiled through shared memory; pass 1: one thread per (batch, target).
// seed: (n, ncols) AoS in particle-id order (the library's loss-seed buffer of the frame), accumulated with atomics.
// ------------------------------------------------------------------------------------------------
#define SMX_CH_TILE 512
#define SMX_CH_SPLIT 8      // lanes that share one query: each scans every 8th candidate, then a (distance, index) min over the 8 lanes
// lexicographic (distance, index) minimum over the SMX_CH_SPLIT lanes of a query group: equal to a serial scan in index order with
// a strict `<` (the smallest index among equal distances wins)
__device__ __forceinline__ void chamfer_group_min(float& best, uint32_t& bi) {
#pragma unroll
    for (int o = SMX_CH_SPLIT / 2; o > 0; o >>= 1) {
        float d = __shfl_xor_sync(0xffffffffu, best, o);
        uint32_t i = __shfl_xor_sync(0xffffffffu, bi, o);
        if (d < best || (d == best && i < bi)) { best = d; bi = i; }
    }
}
// grid: (ceil(queries / 16), nbatch).  pass 0: queries = the particles of rollout blockIdx.y, candidates = the m targets;
// pass 1: queries = the m targets, candidates = the particles of rollout blockIdx.y (index = particle id: ties go to the smallest id).
__global__ void __launch_bounds__(128) k_chamfer(Params P, const float* __restrict__ fr, const uint32_t* __restrict__ perm, const float* __restrict__ tgt, int m,
                                                 float weight, float* __restrict__ seed, int ncols, double* __restrict__ loss, int pass) {
    __shared__ float4 tile[SMX_CH_TILE];
    const int b = blockIdx.y, part = threadIdx.x % SMX_CH_SPLIT;
    const int q = blockIdx.x * (128 / SMX_CH_SPLIT) + threadIdx.x / SMX_CH_SPLIT;     // query handled by this group of lanes
    const int j0 = b * P.npb;
    const int nq = pass == 0 ? P.npb : m, nc = pass == 0 ? m : P.npb;
    const bool live = q < nq;
    const int qq = live ? q : nq - 1;
    float x, y, z;
    if (pass == 0) { V3 xq = load_x(fr, P.stride, j0 + qq); x = xq.x; y = xq.y; z = xq.z; }
    else { x = tgt[3 * qq]; y = tgt[3 * qq + 1]; z = tgt[3 * qq + 2]; }
    float best = 3.0e38f; uint32_t bi = 0xffffffffu; int bslot = 0;
    for (int base = 0; base < nc; base += SMX_CH_TILE) {
        const int cnt = min(SMX_CH_TILE, nc - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            if (pass == 0) tile[e] = make_float4(tgt[3 * (base + e)], tgt[3 * (base + e) + 1], tgt[3 * (base + e) + 2], __uint_as_float((uint32_t)(base + e)));
            else {
                V3 c = load_x(fr, P.stride, j0 + base + e);
                tile[e] = make_float4(c.x, c.y, c.z, __uint_as_float(perm ? perm[j0 + base + e] : (uint32_t)(j0 + base + e)));
            }
        }
        __syncthreads();
        for (int e = part; e < cnt; e += SMX_CH_SPLIT) {
            const float4 c = tile[e];
            const float dx = x - c.x, dy = y - c.y, dz = z - c.z;
            const float d = dx * dx + dy * dy + dz * dz;
            const uint32_t id = __float_as_uint(c.w);
            if (d < best || (d == best && id < bi)) { best = d; bi = id; bslot = base + e; }
        }
    }
    // which lane holds the winner: after the group minimum every lane knows (best, bi); the owner contributes the seed
    const float my_best = best; const uint32_t my_bi = bi;
    chamfer_group_min(best, bi);
    float partial = 0.f;
    if (live && nc > 0 && my_best == best && my_bi == bi) {
        // pass 0: d loss / d x_q = 2 w (x_q - t_nn);  pass 1: d loss / d x_nn = 2 w (x_nn - t_q)
        float cx, cy, cz; uint32_t row;
        if (pass == 0) { cx = tgt[3 * bslot]; cy = tgt[3 * bslot + 1]; cz = tgt[3 * bslot + 2]; row = perm ? perm[j0 + q] : (uint32_t)(j0 + q); }
        else { V3 c = load_x(fr, P.stride, j0 + bslot); cx = c.x; cy = c.y; cz = c.z; row = bi; }
        const float sgn = pass == 0 ? 1.f : -1.f;
        const float dx = sgn * (x - cx), dy = sgn * (y - cy), dz = sgn * (z - cz);
        atomicAdd(seed + (size_t)row * ncols, 2.f * weight * dx); atomicAdd(seed + (size_t)row * ncols + 1, 2.f * weight * dy);
        atomicAdd(seed + (size_t)row * ncols + 2, 2.f * weight * dz);
        partial = dx * dx + dy * dy + dz * dz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, o);
    if ((threadIdx.x & 31) == 0 && partial != 0.f) atomicAdd(loss, (double)(weight * partial));
}
// ------------------------------------------------------------------------------------------------
// Contact-distance term of DoorLoss / TransportLoss on the device (softmac/engine/losses/loss_door.py:46-56,
// loss_transport.py:54-70): per controller group g (particle ids [g npc, (g+1) npc) of a rollout),
//   m_g = min(min_i max(|x_i - p|^2 - 0.01, 0), 1e6),   loss += weight * sum_g m_g^2,
// p = position of one primitive at frame f; the gradient 2 m_g * 2 (x_i - p) goes to the minimising particle (smallest id among
// equal minima, as numpy.argmin in the host mirror) and, negated, to the primitive position.  Distances in f64 from the fp32 frame,
// exactly as the host mirror computes them.  pass 0: minimum of the distance bits; pass 1: smallest id that attains it; then
// k_contact_dist_finish (one thread per (rollout, group)).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double contact_dist_of(const Params& P, const float* __restrict__ fr, int j, const float* __restrict__ ps13) {
    V3 x = load_x(fr, P.stride, j);
    const double dx = (double)x.x - (double)ps13[0], dy = (double)x.y - (double)ps13[1], dz = (double)x.z - (double)ps13[2];
    return fmax(dx * dx + dy * dy + dz * dz - 0.01, 0.0);
}
__global__ void __launch_bounds__(256) k_contact_dist_min(Params P, const float* __restrict__ fr, const uint32_t* __restrict__ perm, const float* __restrict__ pstate,
                                                          int T, int prim, int f, int ngroups, int npc, unsigned long long* __restrict__ best, int pass) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    const int bt = batch_of(P, j);
    const uint32_t id = (perm ? perm[j] : (uint32_t)j) - (uint32_t)(bt * P.npb);       // particle id inside its rollout
    const int g = (int)(id / (uint32_t)npc);
    if (g >= ngroups) return;                                                           // the reference loops over n_groups * npc particles only
    const double d = contact_dist_of(P, fr, j, pstate + (((size_t)bt * SMX_MAXP + prim) * T + f) * 13);
    unsigned long long* slot = best + 2 * ((size_t)bt * ngroups + g);
    // non-negative doubles order like their bit patterns
    if (pass == 0) atomicMin(slot, (unsigned long long)__double_as_longlong(d));
    else if ((unsigned long long)__double_as_longlong(d) == slot[0]) atomicMin(slot + 1, (unsigned long long)id);
}
__global__ void k_contact_dist_finish(Params P, const float* __restrict__ fr, const float* __restrict__ pstate, double* __restrict__ pgrad, int T, int prim, int f, int ngroups, int npc,
                                      const unsigned long long* __restrict__ best, double weight, float* __restrict__ seed, int ncols, double* __restrict__ loss,
                                      uint32_t* __restrict__ slot_of_winner) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nbatch * ngroups) return;
    const int bt = t / ngroups, g = t % ngroups;
    const double dmin = __longlong_as_double((long long)best[2 * t]);
    const double m = fmin(dmin, 1e6);
    atomicAdd(loss + bt, weight * m * m);
    if (!(m > 0.0 && m < 1e6)) return;
    const uint32_t id = (uint32_t)best[2 * t + 1];
    const int j = (int)slot_of_winner[t];                                               // storage slot of the winner (found by k_contact_dist_slot)
    const float* ps13 = pstate + (((size_t)bt * SMX_MAXP + prim) * T + f) * 13;
    V3 x = load_x(fr, P.stride, j);
    const double k = weight * 2.0 * m * 2.0;
    const double gx = k * ((double)x.x - (double)ps13[0]), gy = k * ((double)x.y - (double)ps13[1]), gz = k * ((double)x.z - (double)ps13[2]);
    float* row = seed + ((size_t)bt * P.npb + id) * ncols;
    atomicAdd(row, (float)gx); atomicAdd(row + 1, (float)gy); atomicAdd(row + 2, (float)gz);
    double* pg = pgrad + (((size_t)bt * SMX_MAXP + prim) * T + f) * 13;
    atomicAdd(pg, -gx); atomicAdd(pg + 1, -gy); atomicAdd(pg + 2, -gz);
}
// storage slot of every (rollout, group) winner: the slot whose particle id equals the recorded one
__global__ void __launch_bounds__(256) k_contact_dist_slot(Params P, const uint32_t* __restrict__ perm, int ngroups, int npc, const unsigned long long* __restrict__ best,
                                                           uint32_t* __restrict__ slot_of_winner) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    const int bt = batch_of(P, j);
    const uint32_t id = (perm ? perm[j] : (uint32_t)j) - (uint32_t)(bt * P.npb);
    const int g = (int)(id / (uint32_t)npc);
    if (g >= ngroups) return;
    if ((unsigned long long)id == best[2 * ((size_t)bt * ngroups + g) + 1]) slot_of_winner[bt * ngroups + g] = (uint32_t)j;
}
// Primitive.forward_kinematics and its adjoint (primitive_base.py:280-283, primitive_utils.py:20-40); one thread
__global__ void k_forward_kinematics(float* __restrict__ pstate, int T, int np, int nbatch, int f, float dt) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np * nbatch) return;
    int i = (t / np) * SMX_MAXP + (t % np);
    float* s0 = pstate + ((size_t)i * T + f) * 13;
    float* s1 = s0 + 13;
    for (int d = 0; d < 3; d++) s1[d] = s0[d] + s0[7 + d] * dt;
    double aa[3] = {(double)s0[10] * dt, (double)s0[11] * dt, (double)s0[12] * dt};
    double w = sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2] + 1e-12);
    double sn = sin(w / 2), q[4] = {cos(w / 2), aa[0] / w * sn, aa[1] / w * sn, aa[2] / w * sn};
    double r[4] = {s0[3], s0[4], s0[5], s0[6]}, o[4];
    o[0] = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3];
    o[1] = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2];
    o[2] = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1];
    o[3] = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0];
    double nn = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
    for (int d = 0; d < 4; d++) s1[3 + d] = (float)(o[d] / nn);
}
__global__ void k_forward_kinematics_grad(const float* __restrict__ pstate, double* __restrict__ pgrad, int T, int np, int nbatch, int f, float dtf) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np * nbatch) return;
    int i = (t / np) * SMX_MAXP + (t % np);
    const float* s0 = pstate + ((size_t)i * T + f) * 13;
    double* g0 = pgrad + ((size_t)i * T + f) * 13;
    const double* g1 = g0 + 13;
    double dt = dtf;
    for (int d = 0; d < 3; d++) { g0[d] += g1[d]; g0[7 + d] += dt * g1[d]; }
    double aa[3] = {(double)s0[10] * dt, (double)s0[11] * dt, (double)s0[12] * dt};
    double w = sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2] + 1e-12);
    double sn = sin(w / 2), cs = cos(w / 2), q[4] = {cs, aa[0] / w * sn, aa[1] / w * sn, aa[2] / w * sn};
    double r[4] = {s0[3], s0[4], s0[5], s0[6]}, o[4];
    o[0] = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3];
    o[1] = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2];
    o[2] = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1];
    o[3] = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0];
    double nn = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
    double y[4], dd = 0, g[4];
    for (int d = 0; d < 4; d++) { y[d] = o[d] / nn; dd += y[d] * g1[3 + d]; }
    for (int d = 0; d < 4; d++) g[d] = (g1[3 + d] - y[d] * dd) / nn;
    double gr[4], gq[4];
    gr[0] = g[0] * q[0] + g[1] * q[1] + g[2] * q[2] + g[3] * q[3];
    gr[1] = -g[0] * q[1] + g[1] * q[0] + g[2] * q[3] - g[3] * q[2];
    gr[2] = -g[0] * q[2] - g[1] * q[3] + g[2] * q[0] + g[3] * q[1];
    gr[3] = -g[0] * q[3] + g[1] * q[2] - g[2] * q[1] + g[3] * q[0];
    gq[0] = g[0] * r[0] + g[1] * r[1] + g[2] * r[2] + g[3] * r[3];
    gq[1] = -g[0] * r[1] + g[1] * r[0] - g[2] * r[3] + g[3] * r[2];
    gq[2] = -g[0] * r[2] + g[1] * r[3] + g[2] * r[0] - g[3] * r[1];
    gq[3] = -g[0] * r[3] - g[1] * r[2] + g[2] * r[1] + g[3] * r[0];
    for (int d = 0; d < 4; d++) g0[3 + d] += gr[d];
    double gw = -0.5 * sn * gq[0], gaa[3];
    for (int d = 0; d < 3; d++) { gaa[d] = gq[d + 1] * sn / w; gw += gq[d + 1] * aa[d] * (0.5 * cs / w - sn / (w * w)); }
    for (int d = 0; d < 3; d++) { gaa[d] += gw * aa[d] / w; g0[10 + d] += dt * gaa[d]; }
}

}  // namespace smx
