/*
 * softmac_b200.h -- C ABI of libsoftmac_b200.so: a B200-native (sm_100a) implementation of SoftMAC's
 * differentiable MLS-MPM substep loop, forward and adjoint, and of the rigid-primitive coupling
 * interface (contact wrench out, primitive pose/twist in, and the adjoints of both).
 *
 * The reference (damianliumin/SoftMAC) has no FFI: its boundary is the duck-typed Python surface of
 * `MPMSimulator` (softmac/engine/mpm_simulator.py:16-618) and `Primitive`
 * (softmac/engine/primitive/primitive_base.py:8-335) as consumed by `TaichiEnv`
 * (softmac/engine/taichi_env.py:93-151) and `RigidSimulator` (softmac/engine/rigid_simulator.py:85-220).
 * Each entry point below names the reference method it replaces.  The Python mirror of that surface
 * (softmac_b200/engine/*.py) binds exactly these symbols through ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative smx_status; the message of the last failure
 *     on the calling thread is available from smx_last_error().
 *   - pointers are HOST pointers unless the name ends in _dev.  Host arrays are borrowed for the
 *     duration of the call.  Particle arrays are in PARTICLE-ID order (the order of the array passed
 *     to smx_reset), float64, row-major: x,v (n,3); F,C (n,3,3); state (n,24) = [x v F C] -- the
 *     layout of MPMSimulator.get_state (mpm_simulator.py:481-489).  Device storage is fp32, SoA,
 *     physically sorted by grid cell; the library maps between the two.
 *   - one smx_sim <-> one CUDA device and one stream.  Calls on one handle are not thread-safe;
 *     different handles may be driven from different threads/processes.  Kernels are launched
 *     asynchronously; only the smx_get_* calls, smx_synchronize and smx_timer_stop block.
 *   - frame f of the particle state is the input of substep f, frame f+1 its output
 *     (taichi_env.py:94-95).  Primitive frame f is what substep f collides against.
 */
#ifndef SOFTMAC_B200_H
#define SOFTMAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smx_sim smx_sim;

typedef enum {
    SMX_OK = 0,
    SMX_ERR_ARG = -1,      /* bad argument / null pointer */
    SMX_ERR_RANGE = -2,    /* frame / primitive / controller index out of range */
    SMX_ERR_CUDA = -3,     /* CUDA runtime error (message carries cudaGetErrorString) */
    SMX_ERR_STATE = -4,    /* call sequence not valid (e.g. adjoint of a frame that was never computed) */
    SMX_ERR_NOMEM = -5,
    SMX_ERR_DOMAIN = -6    /* particles left the active region / the grid (see smx_get_counters) */
} smx_status;

/* MPMSimulator.__init__(cfg, primitives, env_dt, rigid_velocity_control) -- mpm_simulator.py:17-84.
 * Field names follow cfg.SIMULATOR (softmac/config/default_config.py:14-29). */
typedef struct {
    int32_t n_particles;
    int32_t n_grid;               /* int(128 * quality * 0.5); must be a multiple of 4 */
    int32_t max_steps;            /* number of particle / primitive frames kept in HBM */
    double dt;
    double E, nu;                 /* Lame parameters derived as in mpm_simulator.py:41-45 */
    double gravity[3];
    double ground_friction;       /* >= 10 -> sticky floor (mpm_simulator.py:278) */
    int32_t material_model;       /* 0 co-rotated, 1 neo-Hookean */
    int32_t ptype;                /* 0 plastic, 1 elastic, 2 liquid */
    int32_t collision_type;       /* 0 grid, 1 particle, 2 mixed (forecast) */
    int32_t substeps;             /* int(env_dt / dt): period of `life` in grid_op_mixed3 (:425) */
    int32_t n_control;            /* cfg.n_controllers */
    int32_t rigid_velocity_control;
    int32_t sort_every;           /* re-bin + re-sort particles every this many substeps (0: only at reset) */
    int32_t device;               /* CUDA device ordinal */
    int32_t flags;                /* SMX_FLAG_* */
    void* stream;                 /* cudaStream_t to run on; NULL -> the library creates its own */
    int32_t n_batch;              /* independent rollouts batched in this handle (0/1: one).  n_particles, primitives' states,
                                     wrenches and actions are PER BATCH; particle arrays at the API have n_batch * n_particles
                                     rows, batch-major (BASELINE config 4: several rollouts per GPU in one handle) */
} smx_config;

#define SMX_FLAG_DENSE_GRID 1     /* sweep the whole grid every substep instead of the active-block list */
#define SMX_FLAG_NO_SORT 2        /* keep particles in id order (debug / ablation) */
#define SMX_FLAG_NO_GRID_CKPT 8    /* adjoint re-runs P2G + grid update instead of restoring the per-substep grid checkpoint */
#define SMX_FLAG_EXTERNAL_STREAM 16 /* run on smx_config.stream even when it is NULL (the legacy default stream): lets torch ops on the
                                     same stream interleave with the kernels without extra synchronisation (slab halo exchange) */
#define SMX_FLAG_NO_FUSION 32       /* smx_step launches G2P and the next P2G separately instead of the fused G2P2G kernel */
#define SMX_FLAG_DIRECT_RED 4     /* one L2 reduction per particle and node instead of the warp-aggregated scatter (ablation) */
#define SMX_FLAG_NO_TMA 128       /* particle kernels read their streaming planes straight from HBM instead of the persistent,
                                     TMA-staged (cp.async.bulk + mbarrier, double-buffered) variants (ablation) */
#define SMX_FLAG_NO_SVD_REC 64    /* do not keep the per-substep SVD records (64 B per particle and substep); the adjoint repeats the SVD */

/* lifetime ------------------------------------------------------------------------------------- */
int smx_create(const smx_config* cfg, smx_sim** out);
int smx_destroy(smx_sim* sim);
const char* smx_last_error(void);
int smx_synchronize(smx_sim* sim);

/* primitives: Mesh(...) construction -- softmac/engine/primitive/mesh.py:26-43.  sdf is the
 * (r0,r1,r2) table, normal the (r0,r1,r2,3) table, lower/upper = sdf["position"], sdf_dx = sdf["dx"][0].
 * sdf == NULL registers a primitive without geometry (never in contact).  Returns the primitive id. */
int smx_add_primitive(smx_sim* sim, const double* sdf, const double* normal, const int32_t res[3], const double lower[3],
                      const double upper[3], double sdf_dx, double friction, double softness, int32_t contact_enabled);
/* Primitive.friction / softness fields (primitive_base.py:26-27, primitives.py:43-45) */
int smx_set_primitive_params(smx_sim* sim, int32_t id, double friction, double softness);
/* MPMSimulator.primitives_contact[id] = flag (mpm_simulator.py:70, demo_grip.py:117) */
int smx_set_primitive_contact(smx_sim* sim, int32_t id, int32_t enabled);

/* particle state IO ---------------------------------------------------------------------------- */
/* MPMSimulator.reset(x): ncols == 3 -> v=0, F=I, C=0; ncols == 24 -> full state (mpm_simulator.py:494-519) */
int smx_reset(smx_sim* sim, const double* state, int32_t ncols);
/* MPMSimulator.set_state(f, [x, v, F, C]) / setframe (mpm_simulator.py:458-466, 491-492) */
int smx_set_frame(smx_sim* sim, int32_t f, const double* x, const double* v, const double* F, const double* C);
/* MPMSimulator.get_state(f) -> (n,24) (mpm_simulator.py:481-489) */
int smx_get_state(smx_sim* sim, int32_t f, double* out24);
/* get_x / set_x / get_v / set_v (mpm_simulator.py:521-559) */
int smx_get_x(smx_sim* sim, int32_t f, double* x);
int smx_set_x(smx_sim* sim, int32_t f, const double* x);
int smx_get_v(smx_sim* sim, int32_t f, double* v);
int smx_set_v(smx_sim* sim, int32_t f, const double* v);
/* MPMSimulator.copyframe(source, target) incl. the primitives' `substeps` frames (mpm_simulator.py:468-479) */
int smx_copy_frame(smx_sim* sim, int32_t src, int32_t dst);

/* primitive state / coupling ------------------------------------------------------------------- */
/* Primitive.set_all_states(f, state13) for f in [f0, f1) (primitive_base.py:258-260; the rigid bridge loops
 * `substeps` such calls, rigid_simulator.py:200-201).  state13 = [x(3) q(4, w first) v(3) w(3)]. */
int smx_set_primitive_state(smx_sim* sim, int32_t id, int32_t f0, int32_t f1, const double* s13);
/* Primitive.get_state(f) (first 7) + v, w */
int smx_get_primitive_state(smx_sim* sim, int32_t id, int32_t f, double* out13);
/* sum over f in [f0, f1) of Primitive.get_all_states_grad(f) (primitive_base.py:262-265, rigid_simulator.py:207-208) */
int smx_get_primitive_state_grad(smx_sim* sim, int32_t id, int32_t f0, int32_t f1, double* out13);
/* loss seed on position/rotation/v/w.grad[f] (what Taichi losses write directly, e.g. loss_grip.py:70-87) */
int smx_add_primitive_state_grad(smx_sim* sim, int32_t id, int32_t f, const double* g13);
/* Primitive.ext_f.to_numpy() (rigid_simulator.py:92): [force(3), torque about the primitive origin(3)] */
int smx_get_ext_f(smx_sim* sim, int32_t id, double* out6);
/* Primitive.clear_ext_f(): zeroes value and adjoint (primitive_base.py:183-187) */
int smx_clear_ext_f(smx_sim* sim, int32_t id);
/* Primitive.set_ext_f_grad(g6) (primitive_base.py:189-192) */
int smx_set_ext_f_grad(smx_sim* sim, int32_t id, const double* g6);
/* batch-addressed variants of the six calls above (n_batch > 1: every rollout has its own rigid bodies).  The
 * un-suffixed setters / clears act on every batch, the un-suffixed getters on batch 0. */
int smx_set_primitive_state_b(smx_sim* sim, int32_t batch, int32_t id, int32_t f0, int32_t f1, const double* s13);
int smx_get_primitive_state_b(smx_sim* sim, int32_t batch, int32_t id, int32_t f, double* out13);
int smx_get_primitive_state_grad_b(smx_sim* sim, int32_t batch, int32_t id, int32_t f0, int32_t f1, double* out13);
int smx_add_primitive_state_grad_b(smx_sim* sim, int32_t batch, int32_t id, int32_t f, const double* g13);
int smx_get_ext_f_b(smx_sim* sim, int32_t batch, int32_t id, double* out6);
int smx_clear_ext_f_b(smx_sim* sim, int32_t batch, int32_t id);
int smx_set_ext_f_grad_b(smx_sim* sim, int32_t batch, int32_t id, const double* g6);
/* bulk coupling for batched handles: ONE transfer for all rollouts and primitives per env step.  Arrays are
 * [n_batch][n_primitives][6 | 13], batch-major.  Same semantics as looping the calls above over (batch, id). */
int smx_get_ext_f_all(smx_sim* sim, double* out6);
int smx_clear_ext_f_all(smx_sim* sim);
int smx_set_ext_f_grads_all(smx_sim* sim, const double* g6);
int smx_set_primitive_states_all(smx_sim* sim, int32_t f0, int32_t f1, const double* s13);
int smx_get_primitive_state_grads_all(smx_sim* sim, int32_t f0, int32_t f1, double* out13);
/* velocity-control mode: Primitive.set_action(s, n, a6) / get_action_grad(s, n) (primitive_base.py:285-319) */
int smx_set_primitive_action(smx_sim* sim, int32_t id, int32_t s, int32_t n, const double* a6);
int smx_get_primitive_action_grad(smx_sim* sim, int32_t id, int32_t s, int32_t n, double* out6);
/* Primitive.reset(): clear_all_states (state series AND their adjoints, primitive_base.py:236-246), clear_ext_f (:183-187) and
 * clear_action_buffer (:321-326) in one call, for every batched rollout */
int smx_reset_primitive(smx_sim* sim, int32_t id);

/* Device-resident rigid coupling: the stand-in rigid integrator (bodies on fixed, prismatic, revolute or free joints -- the gripper
 * of demo_grip, the glass and bowl of demo_pour, the hinged door of demo_door) behind the RigidSimulator interface, on the GPU.
 * Replaces the per-env-step host round trip of RigidSimulator.step / set_ext_state / step_grad / get_ext_state_grad
 * (softmac/engine/rigid_simulator.py:85-220): the bridge reads primitive.ext_f / substeps in float32 (:92-93), ignores
 * wrenches below 1e-10 or of primitives with enable_external_force == False (:96), advances the bodies, and writes pose +
 * twist into the next `substeps` primitive frames in float32 (:185, :200-201); backwards it sums get_all_states_grad over
 * those frames (:207-216), emits the action gradient and set_ext_f_grad(. / substeps) (:166-168).
 * The integrator is affine,  s' = s As + a Aa + w Aw + c,  and the pose map of a body is closed form (its Jacobian in closed
 * form for prismatic / revolute joints, by central differences, eps 1e-6, for the free joint -- as the host bridge), so all of that is one small kernel per env step on the
 * simulator's stream: no synchronisation inside an episode.  State layout: [q (state_dim / 2), q_dot (state_dim / 2)],
 * the dofs of primitive i's body at dof_offset_i (prismatic: 1; revolute: 1 hinge angle about `axis` through the link origin;
 * free: 3 exponential coordinates + 3 translations).
 * Row-major f64: As (state_dim, state_dim), Aa (action_dim, state_dim), Aw (6 * n_primitives, state_dim), c (state_dim),
 * body (n_primitives, 10) = origin(3) quat0(4, w first) axis(3); int32: joint (n_primitives, 2) = (type: 0 fixed,
 * 1 prismatic, 2 free, 3 revolute; dof_offset), enable (n_primitives); init_state (state_dim). */
typedef struct {
    int32_t state_dim, action_dim;
    int32_t max_env_steps;        /* env steps kept (states, actions, action gradients, wrench masks) */
    int32_t fp32_bridge;          /* truncate the wrench to float32 as the Jade bridge does (rigid_simulator.py:92) */
    double ext_grad_scale;        /* RigidSimulator.ext_grad_scale (rigid_simulator.py:148) */
    const double *As, *Aa, *Aw, *c, *body, *init_state;
    const int32_t *joint, *enable;
} smx_rigid_linear;
int smx_rigid_linear_create(smx_sim* sim, const smx_rigid_linear* desc);
/* RigidSimulator.reset: every rollout back to init_state, poses of frames [0, substeps) written, wrench cleared */
int smx_rigid_linear_reset(smx_sim* sim);
/* actions of env step k for every rollout, [n_batch][action_dim] */
int smx_rigid_linear_set_actions(smx_sim* sim, int32_t k, const double* actions);
/* RigidSimulator.step(k) after the substeps of env step k; RigidSimulator.step_grad(k) before their adjoints */
int smx_rigid_linear_step(smx_sim* sim, int32_t k);
int smx_rigid_linear_step_grad(smx_sim* sim, int32_t k);
/* state_grad += get_ext_state_grad(0) at the end of TaichiEnv.backward (taichi_env.py:149-150) */
int smx_rigid_linear_finish(smx_sim* sim);
/* read-out (blocking): rigid state after k env steps [n_batch][state_dim]; action gradients of env steps [k0, k1)
 * [k1 - k0][n_batch][action_dim]; adjoint of the initial rigid state [n_batch][state_dim] */
int smx_rigid_linear_get_states(smx_sim* sim, int32_t k, double* out);
int smx_rigid_linear_get_action_grads(smx_sim* sim, int32_t k0, int32_t k1, double* out);
int smx_rigid_linear_get_state_grad(smx_sim* sim, double* out);

/* Plastic flow rule of the co-rotated plastic material (material_model 0, ptype 0).  mode 0 (default): sigma clip to [1 - 2e-3, 1 + 3e-3]
 * (softmac/engine/mpm_simulator.py:226-229).  mode 1: von Mises return mapping in log strain with cfg.SIMULATOR.yield_stress, the rule the
 * soft_cloth variant runs (soft_cloth/engine/mpm_simulator.py:172-189 compute_von_mises, call site :232; yield_stress field :20,49,92) and
 * that softmac keeps commented out at mpm_simulator.py:225.  Forward and adjoint (Taichi sub-gradients: max(sig, 0.05) passes its gradient
 * iff 0.05 < sig, the yield test carries none).  Call between rollouts, not inside one. */
int smx_set_plasticity(smx_sim* sim, int32_t mode, double yield_stress);

/* particle-force control ("mpm" control mode) --------------------------------------------------- */
/* MPMSimulator.set_action(action (n_control,3)); also zeroes action.grad (mpm_simulator.py:579-592).
 * With n_batch > 1 the array is (n_batch * n_control, 3), batch-major (likewise smx_get_action_grad). */
int smx_set_action(smx_sim* sim, const double* action);
/* MPMSimulator.set_control_idx(idx (n,) int, -1 = uncontrolled) (mpm_simulator.py:594-602) */
int smx_set_control_idx(smx_sim* sim, const int32_t* idx);
/* action.grad.to_numpy() after substep_grad (mpm_simulator.py:378) */
int smx_get_action_grad(smx_sim* sim, double* out);

/* the hot path ---------------------------------------------------------------------------------- */
/* MPMSimulator.substep(s) (mpm_simulator.py:320-337): frame s -> frame s+1, accumulates ext_f */
int smx_substep(smx_sim* sim, int32_t s);
/* MPMSimulator.substep_grad(s) (mpm_simulator.py:339-378): consumes the adjoint of frame s+1 (+ its seeds),
 * produces the adjoint of frame s (+ its seeds), accumulates primitive-state and action adjoints */
int smx_substep_grad(smx_sim* sim, int32_t s);
/* The two halves of a substep, for callers that exchange grid halos between them (spatial slab decomposition):
 * begin = clear + P2G;  end = grid update, contact, G2P.  smx_substep == begin followed by end.  Likewise for the
 * adjoint: begin = adjoint set-up, grid restore, G2P adjoint (scatter into gg_out);  end = grid adjoint + P2G adjoint. */
int smx_substep_begin(smx_sim* sim, int32_t s);
int smx_substep_end(smx_sim* sim, int32_t s);
int smx_substep_grad_begin(smx_sim* sim, int32_t s);
int smx_substep_grad_end(smx_sim* sim, int32_t s);
/* With the forecast contact model a slab caller needs one more cut per direction: forward mid = grid update + contact
 * scatter into g_out (then exchange g_out - g_mix on the halo, add the neighbour's part, call end); adjoint mid = contact
 * adjoint (gathers the halo-complete gg_out, scatters gg_mix; then sum the gg_mix halo, call end).  Optional: end runs mid
 * itself when the caller did not. */
int smx_substep_mid(smx_sim* sim, int32_t s);
int smx_substep_grad_mid(smx_sim* sim, int32_t s);
/* Declares this handle one rank of an x-slab decomposition: it owns the x-block columns [xb_lo, xb_hi) (4 nodes per
 * column) and has neighbours on the flagged sides.  The halo columns {xb_lo-1, xb_lo} / {xb_hi-1, xb_hi} are kept active
 * every substep; the caller sums them with the neighbour's copies between begin and end (g_in forward, gg_out backward).
 * Call before smx_reset.  Wrenches, primitive-state adjoints and action adjoints are per-rank partial sums (the caller
 * all-reduces them).  Particle migration between ranks is not implemented: particles whose stencil leaves the own
 * slab + halo are counted in counters[1]. */
int smx_set_slab(smx_sim* sim, int32_t xb_lo, int32_t xb_hi, int32_t has_lo_neighbour, int32_t has_hi_neighbour);
/* Halo exchange over peer memory (NVLink / NVSwitch P2P; SURVEY.md 8e "halo exchange of P2G ghost cells"): every slab rank owns ONE
 * allocation of receive slots that its x-neighbours write into directly.  smx_slab_halo_export returns its device pointer (ranks
 * emulated in one process) and a 64-byte CUDA IPC handle (one process per GPU: exchange the handles once, e.g. with an all-gather);
 * smx_slab_halo_connect maps a neighbour's allocation (side 0: below xb_lo, side 1: above xb_hi; pass the pointer OR the handle).
 * Once every neighbour is connected, smx_substep / smx_substep_grad / smx_step / smx_step_grad run the halo sums themselves, on the
 * simulator's stream -- push of the non-empty halo blocks into the neighbour's slot, flag, wait, add -- with no host round trip, and
 * smx_step keeps its cross-substep fusion.  smx_slab_halo_push / _add are the two halves of one exchange for ranks emulated on ONE
 * stream (all ranks must push before any rank adds).  array: 0 g_in, 1 contact scatter into g_out, 3 adjoint grid of substep f,
 * 4 gg_mix.  smx_slab_halo_status: out[0] exchanges that gave up waiting for a neighbour (must be 0), out[1] exchanges done,
 * out[2] bytes of the halo allocation. */
int smx_slab_halo_export(smx_sim* sim, void** base, void* ipc_handle64);
int smx_slab_halo_connect(smx_sim* sim, int32_t side, void* base, const void* ipc_handle64);
int smx_slab_halo_push(smx_sim* sim, int32_t array, int32_t f);
int smx_slab_halo_add(smx_sim* sim, int32_t array, int32_t f);
int smx_slab_halo_status(smx_sim* sim, int64_t out[3]);
/* device pointer / element count of a grid array: 0 g_in, 1 g_out, 2 g_mix, 3 gg_out, 4 gg_mix (float4 per node,
 * block-major; x-block column c is the contiguous range [c*nb^2*64, (c+1)*nb^2*64), nb = n_grid/4) */
int smx_grid_dev(smx_sim* sim, int32_t which, void** ptr_dev, int64_t* n_float4);
int smx_stream(smx_sim* sim, void** stream);
/* `count` substeps in one call: s0, s0+1, ... (TaichiEnv.step inner loop, taichi_env.py:101-102) */
int smx_step(smx_sim* sim, int32_t s0, int32_t count);
/* adjoint of substeps s1-1, s1-2, ..., s1-count (TaichiEnv.step_grad inner loop, taichi_env.py:128-131) */
int smx_step_grad(smx_sim* sim, int32_t s1, int32_t count);

/* smx_step / smx_step_grad replayed as ONE CUDA-graph launch (for scenes of ~10 k particles, where an env step is bound by the latency of its
 * dependent launches): the call's launch sequence is captured, the handle's executable graph is updated in place with it (same topology,
 * new frame pointers) and launched.  Calls that would allocate, synchronise or re-sort inside take the ordinary path; results are identical.
 * smx_graph_status: out[0] calls replayed as a graph, out[1] calls on the ordinary path, out[2] re-instantiations. */
int smx_step_graph(smx_sim* sim, int32_t s0, int32_t count);
int smx_step_grad_graph(smx_sim* sim, int32_t s1, int32_t count);
int smx_graph_status(smx_sim* sim, int64_t out[3]);

/* fp32 host entry points (the optimiser's float32 tensors never pass through a host-side f64 conversion): same semantics as the f64
 * calls they mirror -- MPMSimulator.reset / get_state / get_grad (mpm_simulator.py:448-574) and the loss seeds -- with float rows that
 * travel as they are.  smx_host_register pins a caller buffer once (cudaHostRegister) so that these copies are plain DMA. */
int smx_host_register(void* ptr, uint64_t bytes);
int smx_host_unregister(void* ptr);
int smx_reset_f32(smx_sim* sim, const float* state, int32_t ncols);
int smx_get_state_f32(smx_sim* sim, int32_t f, float* out24);
int smx_add_x_grad_f32(smx_sim* sim, int32_t f, const float* g3);
int smx_add_state_grad_f32(smx_sim* sim, int32_t f, const float* g24);
int smx_get_state_grad_f32(smx_sim* sim, int32_t f, float* out24);
int smx_get_grad_f32(smx_sim* sim, int32_t f, float* xg, float* vg);

/* adjoint seeds and read-out -------------------------------------------------------------------- */
/* loss -> x.grad[f] (+ v, F, C .grad[f]): g24 is (n,24) in get_state layout; g3 is (n,3) */
int smx_add_state_grad(smx_sim* sim, int32_t f, const double* g24);
int smx_add_x_grad(smx_sim* sim, int32_t f, const double* g3);
/* GripLoss with weight (1,0,0) on the device (softmac/engine/losses/loss_grip.py:45-68, 117-140): Chamfer distance between
 * the particles of frame f (of every batched rollout) and a target cloud (m,3); the value is returned and its gradient is
 * accumulated into the loss seed of frame f (nearest neighbours by brute force, reference tie rule, fixed in the gradient). */
int smx_set_chamfer_target(smx_sim* sim, const double* target, int32_t m);
int smx_chamfer_loss(smx_sim* sim, int32_t f, double weight, double* loss_out);
/* Contact-distance term of DoorLoss / TransportLoss on the device (softmac/engine/losses/loss_door.py:46-56, loss_transport.py:54-70):
 * per controller group g (particle ids [g npc, (g+1) npc), npc = n_particles / n_groups) m_g = min(min_i max(|x_i - p|^2 - 0.01, 0), 1e6)
 * with p the position of primitive prim_id at frame f; loss_out[b] = weight * sum_g m_g^2 of rollout b (n_batch values); the gradient
 * goes into the loss seed of frame f (minimising particle, smallest id among ties) and into the primitive's position adjoint. */
int smx_contact_distance_loss(smx_sim* sim, int32_t f, int32_t prim_id, int32_t n_groups, double weight, double* loss_out);
/* adjoint of frame f; only the frame most recently produced by smx_substep_grad is resident
 * (no per-frame gradient arrays are kept) */
int smx_get_state_grad(smx_sim* sim, int32_t f, double* out24);
/* MPMSimulator.get_grad(f) -> (x.grad[f], v.grad[f]) (mpm_simulator.py:561-574) */
int smx_get_grad(smx_sim* sim, int32_t f, double* xg, double* vg);
/* ti.ad.clear_all_gradients() restricted to this path (demo_grip.py:135) */
int smx_clear_grads(smx_sim* sim);

/* device-resident variants for particle migration between slab ranks (softmac_b200/slabs.py): (n, 24) fp32 rows in
 * particle-id order, get_state layout, in DEVICE memory; ordered on the simulator's stream, no host round trip.
 * smx_reset_dev == smx_reset, smx_get_state_dev == smx_get_state, smx_get_state_grad_dev == smx_get_state_grad,
 * smx_add_state_grad_dev == smx_add_state_grad. */
int smx_reset_dev(smx_sim* sim, const float* rows_dev);
int smx_get_state_dev(smx_sim* sim, int32_t f, float* out_dev);
int smx_get_state_grad_dev(smx_sim* sim, int32_t f, float* out_dev);
int smx_add_state_grad_dev(smx_sim* sim, int32_t f, const float* g24_dev);

/* introspection for tests, benches and zero-copy consumers -------------------------------------- */
/* sort key of every particle of frame f in STORAGE order, and the storage permutation (slot -> particle id).
 * Contract: perm == numpy.argsort(key_in_previous_order, kind="stable") composed over re-sorts. */
int smx_get_sort_keys(smx_sim* sim, int32_t f, uint32_t* keys);
int smx_get_permutation(smx_sim* sim, int32_t f, uint32_t* perm);
/* dense (n_grid^3, 4) fp32 copies of the grid after the last substep: g_in = (momentum xyz, mass),
 * g_out = (velocity xyz, active flag), in (i*n_grid + j)*n_grid + k order */
int smx_get_grid(smx_sim* sim, float* g_in, float* g_out);
/* counters[0] = particles clamped into the grid, [1] = particles that left the active-block region,
 * [2] = number of re-sorts, [3] = active blocks of the current ordering */
int smx_get_counters(smx_sim* sim, int64_t out[4]);
/* device pointer of component c (0..23, get_state column order) of frame f in storage order: the frame is six float4
 * planes, so the value of storage slot j is at ptr[4 * j] (fp32) */
int smx_frame_component_dev(smx_sim* sim, int32_t f, int32_t c, void** ptr_dev);
/* CUDA-event timer on the simulator's stream */
int smx_timer_start(smx_sim* sim);
int smx_timer_stop(smx_sim* sim, float* ms);
/* number of kernels this handle has launched since creation */
int64_t smx_launch_count(smx_sim* sim);
/* Mesh.task / Mesh.trimesh2sdf (softmac/engine/primitive/mesh.py:167-241) on the GPU: signed distance (negative inside)
 * and nearest-face unit normal / (1 + 1e-8) at lower + (i,j,k)*dx for a triangle mesh (vertices (nv,3) f64, faces (nf,3)
 * int32).  Stand-alone: no simulator handle.  Outputs: sdf[r0*r1*r2], normal[r0*r1*r2*3] (host). */
int smx_build_sdf_table(const double* vertices, int32_t nv, const int32_t* faces, int32_t nf, const int32_t res[3], const double lower[3],
                        double dx, double* sdf_out, double* normal_out, int32_t device);
/* runs smx_substep (backward == 0) or smx_substep_grad (backward != 0) of substep f with a CUDA event after every
 * launch and returns the per-kernel-class device time: names[i] (static strings), ms[i], i < *count (<= 32) */
int smx_profile_substep(smx_sim* sim, int32_t f, int32_t backward, const char** names, float* ms, int32_t* count);
/* the same around smx_step(sim, f, n) (backward == 0) or smx_step_grad(sim, f, n): the fused launches of the timed hot path;
 * launches[i] = number of marks (launches) summed into ms[i] */
int smx_profile_step(smx_sim* sim, int32_t f, int32_t n, int32_t backward, const char** names, float* ms, int32_t* launches, int32_t* count);
/* gradient summary of the adjoint of frame f (after a backward pass) written to 16 floats of DEVICE memory on the simulator's
 * stream: [0..2] sum of x.grad, [3..5] sum of v.grad, [6] |x.grad|^2, [7] |v.grad|^2, [8] particles -- what a rank hands to the
 * gradient all-reduce of independent rollouts (BASELINE config 4) without a host round trip */
int smx_grad_summary_dev(smx_sim* sim, int32_t f, float* out16_dev);

#ifdef __cplusplus
}
#endif
#endif /* SOFTMAC_B200_H */
