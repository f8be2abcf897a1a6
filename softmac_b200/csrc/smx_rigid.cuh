// smx_rigid.cuh -- device-resident rigid coupling for articulated bodies whose joints are all fixed or prismatic
// (the gripper of demo_grip).  SURVEY.md 8f row 3: removes the per-env-step host round trip of the rigid bridge.
//
// What the reference does once per env step on the host (softmac/engine/rigid_simulator.py):
//   step          :85-137   read primitive.ext_f / substeps (float32, :92-93), ignore wrenches below 1e-10 or of primitives with
//                           enable_external_force == False (:96), advance the bodies, record the Jacobians
//   set_ext_state :176-203  write pose + twist of every primitive into the next `substeps` frames (float32, :185, :200-201)
//   step_grad     :139-174  state_grad += sum over those frames of get_all_states_grad . d pose / d state (:207-216), action
//                           gradient, wrench adjoint / substeps -> set_ext_f_grad (:166-168), state_grad <- state_grad . ds'/ds
// For fixed / prismatic joints the stand-in integrator is affine in (state, action, wrench):
//   s' = s As + a Aa + w Aw + c          pose_i = pose0_i + s' M_i
// so the whole bridge is a handful of small constant matrices and runs here as one tiny kernel per env step on the simulator's
// stream: no device->host read of the wrench, no host->device write of the poses, no stream synchronisation inside an episode.
// One CTA per batched rollout; all arithmetic in f64 like the host bridge (softmac_b200/engine/batched_env.py:LinearBatchedRigid,
// which stays as the checker of this path in tests/test_cuda_batch.py).
#pragma once
#include "smx_contact.cuh"

namespace smx {

#define SMX_RIG_MAXS 32     // rigid state dimension (positions + velocities of all dofs)
#define SMX_RIG_MAXA 16     // action dimension

struct RigidLin {
    int sd, ad, np, B, S, T, K, fp32;
    double scale;                                   // ext_grad_scale (rigid_simulator.py:83, :148)
    const double *As, *Aa, *Aw, *c, *M, *pose0;     // (sd,sd) (ad,sd) (6 np,sd) (sd) (np,sd,13) (np,13), row-major
    const int* enable;                              // (np) enable_external_force
    double *states, *actions, *action_grad, *state_grad;    // [K+1][B][sd] [K][B][ad] [K][B][ad] [B][sd]
    unsigned char* masks;                           // [K][B][np]: wrench of env step k was fed to the bodies
};

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
    for (int e = t; e < R.np * 13; e += blockDim.x) {
        const int p = e / 13, q = e % 13;
        double acc = R.pose0[e];
        for (int i = 0; i < R.sd; i++) acc += sn[i] * R.M[((size_t)p * R.sd + i) * 13 + q];
        const float v = (float)acc;                 // the bridge hands poses over in float32 (rigid_simulator.py:185)
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
    if (t < R.sd) g[t] = R.state_grad[(size_t)b * R.sd + t];
    __syncthreads();
    if (t < R.sd) {
        double acc = 0;
        for (int p = 0; p < R.np; p++)
            for (int q = 0; q < 13; q++) acc += pg[p * 13 + q] * R.M[((size_t)p * R.sd + t) * 13 + q];
        g2[t] = g[t] + (finish ? 1.0 : R.scale) * acc;
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
