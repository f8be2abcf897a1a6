// smx_rigid.cuh -- device-resident rigid coupling for bodies on fixed, prismatic, revolute or free joints (the gripper of demo_grip,
// the glass and bowl of demo_pour, the hinged door of demo_door: config/demo_door_config.py:31-56, assets/door/door.urdf).  SURVEY.md 8f row 3: removes the per-env-step host round trip of the rigid bridge.
//
// What the reference does once per env step on the host (softmac/engine/rigid_simulator.py):
//   step          :85-137   read primitive.ext_f / substeps (float32, :92-93), ignore wrenches below 1e-10 or of primitives with
//                           enable_external_force == False (:96), advance the bodies, record the Jacobians
//   set_ext_state :176-203  write pose + twist of every primitive into the next `substeps` frames (float32, :185, :200-201)
//   step_grad     :139-174  state_grad += sum over those frames of get_all_states_grad . d pose / d state (:207-216), action
//                           gradient, wrench adjoint / substeps -> set_ext_f_grad (:166-168), state_grad <- state_grad . ds'/ds
// The stand-in integrator (semi-implicit Euler, softmac_b200/engine/rigid_simulator.py) is affine in (state, action, wrench) for
// all three joint types:
//   s' = s As + a Aa + w Aw + c
// and the pose map state -> [x(3) q(4) v(3) w(3)] of a body is closed form (rigid_pose below; nonlinear for free joints: exponential
// coordinates -> quaternion, twist rotated into the body frame).  Its Jacobian is taken by central differences in f64 exactly as the
// host bridge does (eps 1e-6).  So the whole bridge is a few small constant matrices plus one closed-form map and runs here as one
// tiny kernel per env step on the simulator's stream: no device->host read of the wrench, no host->device write of the poses, no
// stream synchronisation inside an episode.  One CTA per batched rollout; all arithmetic in f64 like the host bridge, which stays
// as the checker of this path (tests/test_cuda_batch.py).
#pragma once
#include "smx_contact.cuh"

namespace smx {

#define SMX_RIG_MAXS 32     // rigid state dimension (positions + velocities of all dofs)
#define SMX_RIG_MAXA 16     // action dimension

struct RigidLin {
    int sd, ad, np, B, S, T, K, fp32;
    double scale;                                   // ext_grad_scale (rigid_simulator.py:83, :148)
    const double *As, *Aa, *Aw, *c;                 // (sd,sd) (ad,sd) (6 np,sd) (sd), row-major
    const double* body;                             // (np,10): origin(3) quat0(4, w first) axis(3) of the body behind primitive i
    const int* joint;                               // (np,2): joint type (0 fixed, 1 prismatic, 2 free, 3 revolute), offset of its dofs in the state
    const int* enable;                              // (np) enable_external_force
    double *states, *actions, *action_grad, *state_grad;    // [K+1][B][sd] [K][B][ad] [K][B][ad] [B][sd]
    unsigned char* masks;                           // [K][B][np]: wrench of env step k was fed to the bodies
};

__device__ __forceinline__ void rig_qmul(const double* a, const double* b, double* o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
__device__ __forceinline__ void rig_qrot(const double* q, const double* v, double* o) {       // v + 2 (q_w (q_v x v) + q_v x (q_v x v))
    const double ux = q[2] * v[2] - q[3] * v[1], uy = q[3] * v[0] - q[1] * v[2], uz = q[1] * v[1] - q[2] * v[0];
    o[0] = v[0] + 2 * (q[0] * ux + q[2] * uz - q[3] * uy);
    o[1] = v[1] + 2 * (q[0] * uy + q[3] * ux - q[1] * uz);
    o[2] = v[2] + 2 * (q[0] * uz + q[1] * uy - q[2] * ux);
}
// pose + twist of one body from the rigid state: [x(3) q(4) v(3) w(3)], position / quaternion in the world, body-frame twist
// (what the Jade bridge reads back per body, rigid_simulator.py:176-186); h = sd / 2 separates positions from velocities.
// `bump` / `delta` perturb state entry `bump` for the central differences of the adjoint (free joints; the other joints use the
// closed-form Jacobian, rigid_pose_vjp).
__host__ __device__ __forceinline__ int rig_ndof(int joint) { return joint == 0 ? 0 : (joint == 2 ? 6 : 1); }
__device__ void rigid_pose(const double* __restrict__ body, int joint, int o, int h, const double* __restrict__ s, int bump, double delta, double* out) {
    auto S = [&](int i) { return s[i] + (i == bump ? delta : 0.0); };
    const double* org = body; const double* q0 = body + 3; const double* ax = body + 7;
    if (joint == 0) {
        for (int i = 0; i < 3; i++) out[i] = org[i];
        for (int i = 0; i < 4; i++) out[3 + i] = q0[i];
        for (int i = 7; i < 13; i++) out[i] = 0.0;
    } else if (joint == 1) {
        double aw[3]; rig_qrot(q0, ax, aw);
        const double q = S(o), qd = S(h + o);
        for (int i = 0; i < 3; i++) out[i] = org[i] + aw[i] * q;
        for (int i = 0; i < 4; i++) out[3 + i] = q0[i];
        for (int i = 0; i < 3; i++) { out[7 + i] = ax[i] * qd; out[10 + i] = 0.0; }
    } else if (joint == 3) {        // hinge through the link origin: R = R0 Rot(axis, theta); body-frame twist (0, axis * omega)
        const double th = S(o), om = S(h + o);
        const double sn = sin(0.5 * th), qt[4] = {cos(0.5 * th), sn * ax[0], sn * ax[1], sn * ax[2]};
        for (int i = 0; i < 3; i++) out[i] = org[i];
        rig_qmul(q0, qt, out + 3);
        for (int i = 0; i < 3; i++) { out[7 + i] = 0.0; out[10 + i] = ax[i] * om; }
    } else {
        const double e[3] = {S(o), S(o + 1), S(o + 2)};
        const double th = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        double qe[4];
        if (th < 1e-12) { qe[0] = 1.0; qe[1] = 0.5 * e[0]; qe[2] = 0.5 * e[1]; qe[3] = 0.5 * e[2]; }
        else { const double sn = sin(0.5 * th) / th; qe[0] = cos(0.5 * th); qe[1] = sn * e[0]; qe[2] = sn * e[1]; qe[3] = sn * e[2]; }
        double q[4]; rig_qmul(qe, q0, q);
        const double inv[4] = {q[0], -q[1], -q[2], -q[3]};
        const double wv[3] = {S(h + o), S(h + o + 1), S(h + o + 2)}, lv[3] = {S(h + o + 3), S(h + o + 4), S(h + o + 5)};
        for (int i = 0; i < 3; i++) out[i] = org[i] + S(o + 3 + i);
        for (int i = 0; i < 4; i++) out[3 + i] = q[i];
        rig_qrot(inv, lv, out + 7);
        rig_qrot(inv, wv, out + 10);
    }
}

// pg (13) . d pose / d state entry for prismatic (joint 1) and revolute (joint 3) joints, closed form; vel: the entry is the joint velocity
__device__ __forceinline__ double rigid_pose_vjp(const double* __restrict__ body, int joint, const double* __restrict__ s, int o, bool vel, const double* __restrict__ pg) {
    const double* q0 = body + 3; const double* ax = body + 7;
    if (joint == 1) {
        if (vel) return pg[7] * ax[0] + pg[8] * ax[1] + pg[9] * ax[2];
        double aw[3]; rig_qrot(q0, ax, aw);
        return pg[0] * aw[0] + pg[1] * aw[1] + pg[2] * aw[2];
    }
    if (vel) return pg[10] * ax[0] + pg[11] * ax[1] + pg[12] * ax[2];
    const double th = s[o], c = 0.5 * cos(0.5 * th), dqt[4] = {-0.5 * sin(0.5 * th), c * ax[0], c * ax[1], c * ax[2]};
    double dq[4]; rig_qmul(q0, dqt, dq);
    return pg[3] * dq[0] + pg[4] * dq[1] + pg[5] * dq[2] + pg[6] * dq[3];
}

// advance == 1: env step k (after its substeps ran): wrench -> s[k+1], poses of frames [f0, f1), wrench cleared.
// advance == 0: poses of state k only (reset: k = 0, frames [0, substeps)), wrench cleared.
__global__ void __launch_bounds__(128) k_rigid_linear_step(RigidLin R, int k, int advance, int f0, int f1, double* __restrict__ ext_f,
                                                          float* __restrict__ ext_f_grad, float* __restrict__ pstate) {
    __shared__ double s[SMX_RIG_MAXS], sn[SMX_RIG_MAXS], w[6 * SMX_MAXP], a[SMX_RIG_MAXA];
    __shared__ int msk[SMX_MAXP];
    const int b = blockIdx.x, t = threadIdx.x;
    const int nw = 6 * R.np;
    if (t < R.sd) s[t] = R.states[((size_t)k * R.B + b) * R.sd + t];
    if (t < nw) {
        double v = ext_f[((size_t)b * SMX_MAXP + t / 6) * 6 + t % 6];
        if (R.fp32) v = (double)(float)v;
        w[t] = v / R.S;
        ext_f[((size_t)b * SMX_MAXP + t / 6) * 6 + t % 6] = 0.0;            // clear_ext_f: value and adjoint (primitive_base.py:183-187)
        ext_f_grad[((size_t)b * SMX_MAXP + t / 6) * 6 + t % 6] = 0.f;
    }
    if (advance && t < R.ad) a[t] = R.actions[((size_t)k * R.B + b) * R.ad + t];
    __syncthreads();
    if (advance) {
        if (t < R.np) {
            bool any = false;
            for (int c = 0; c < 6; c++) any |= fabs(w[6 * t + c]) > 1e-10;
            msk[t] = (any && R.enable[t]) ? 1 : 0;
            R.masks[((size_t)k * R.B + b) * R.np + t] = (unsigned char)msk[t];
        }
        __syncthreads();
        if (t < R.sd) {
            double acc = R.c[t];
            for (int i = 0; i < R.sd; i++) acc += s[i] * R.As[i * R.sd + t];
            for (int i = 0; i < R.ad; i++) acc += a[i] * R.Aa[i * R.sd + t];
            for (int i = 0; i < nw; i++) if (msk[i / 6]) acc += w[i] * R.Aw[i * R.sd + t];
            sn[t] = acc;
            R.states[((size_t)(k + 1) * R.B + b) * R.sd + t] = acc;
        }
    } else if (t < R.sd) sn[t] = s[t];
    __syncthreads();
    __shared__ double pose[SMX_MAXP * 13];
    if (t < R.np) rigid_pose(R.body + 10 * t, R.joint[2 * t], R.joint[2 * t + 1], R.sd / 2, sn, -1, 0.0, pose + 13 * t);
    __syncthreads();
    for (int e = t; e < R.np * 13; e += blockDim.x) {
        const int p = e / 13, q = e % 13;
        const float v = (float)pose[e];             // the bridge hands poses over in float32 (rigid_simulator.py:185)
        float* dst = pstate + (((size_t)b * SMX_MAXP + p) * R.T) * 13 + q;
        for (int f = f0; f < f1; f++) dst[(size_t)f * 13] = v;
    }
}

// finish == 0: adjoint of env step k (before its adjoint substeps run): pulls the primitive-state adjoints of frames [f0, f1)
//              (the frames env step k wrote), emits the action gradient of step k and the wrench adjoint of its substeps.
// finish == 1: frames [0, substeps) of the initial pose into state_grad (taichi_env.py:149-150); nothing else.
__global__ void __launch_bounds__(128) k_rigid_linear_step_grad(RigidLin R, int k, int finish, int f0, int f1, const double* __restrict__ pgrad,
                                                               float* __restrict__ ext_f_grad) {
    __shared__ double pg[SMX_MAXP * 13], g[SMX_RIG_MAXS], g2[SMX_RIG_MAXS];
    const int b = blockIdx.x, t = threadIdx.x;
    for (int e = t; e < R.np * 13; e += blockDim.x) {
        const int p = e / 13, q = e % 13;
        const double* src = pgrad + (((size_t)b * SMX_MAXP + p) * R.T) * 13 + q;
        double acc = 0;
        for (int f = f0; f < f1; f++) acc += src[(size_t)f * 13];
        pg[e] = acc;
    }
    __shared__ double sk[SMX_RIG_MAXS];
    // the poses of frames [f0, f1) were produced from the state AFTER env step k (state k + 1); the initial ones from state 0
    if (t < R.sd) { g[t] = R.state_grad[(size_t)b * R.sd + t]; g2[t] = g[t]; sk[t] = R.states[((size_t)(finish ? 0 : k + 1) * R.B + b) * R.sd + t]; }
    __syncthreads();
    {   // one thread per (primitive, own state entry): d pose / d state in closed form (prismatic, revolute) or by central differences
        // (free joint; eps 1e-6, as the host bridge)
        const int p = t / 12, l = t % 12;
        if (p < R.np) {
            const int joint = R.joint[2 * p], o = R.joint[2 * p + 1], h = R.sd / 2;
            const int ndof = rig_ndof(joint);
            if (l < 2 * ndof) {
                const int i = l < ndof ? o + l : h + o + (l - ndof);
                double acc = 0;
                if (joint == 2) {
                    double hi[13], lo[13];
                    rigid_pose(R.body + 10 * p, joint, o, h, sk, i, 1e-6, hi);
                    rigid_pose(R.body + 10 * p, joint, o, h, sk, i, -1e-6, lo);
                    for (int q = 0; q < 13; q++) acc += pg[p * 13 + q] * (hi[q] - lo[q]) / 2e-6;
                } else acc = rigid_pose_vjp(R.body + 10 * p, joint, sk, o, l >= ndof, pg + p * 13);
                g2[i] = g[i] + (finish ? 1.0 : R.scale) * acc;      // every state entry belongs to exactly one body: no race
            }
        }
    }
    __syncthreads();
    if (finish) {
        if (t < R.sd) R.state_grad[(size_t)b * R.sd + t] = g2[t];
        return;
    }
    if (t < R.ad) {
        double acc = 0;
        for (int i = 0; i < R.sd; i++) acc += g2[i] * R.Aa[t * R.sd + i];
        R.action_grad[((size_t)k * R.B + b) * R.ad + t] = acc;
    }
    if (t < 6 * R.np) {
        double acc = 0;
        for (int i = 0; i < R.sd; i++) acc += g2[i] * R.Aw[t * R.sd + i];
        const int m = R.masks[((size_t)k * R.B + b) * R.np + t / 6];
        ext_f_grad[((size_t)b * SMX_MAXP + t / 6) * 6 + t % 6] = m ? (float)(acc / R.S) : 0.f;
    }
    if (t < R.sd) {
        double acc = 0;
        for (int j = 0; j < R.sd; j++) acc += g2[j] * R.As[t * R.sd + j];
        R.state_grad[(size_t)b * R.sd + t] = acc;
    }
}

}  // namespace smx
