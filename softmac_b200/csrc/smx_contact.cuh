// smx_contact.cuh -- rigid-primitive SDF lookup and the three contact models, forward and reverse mode.
//
// Device-side counterpart of softmac/engine/primitive/primitive_base.py:53-181 (sdf, normal,
// collider_v, collide, collide_particle, collide_mixed) and softmac/engine/primitive/mesh.py:45-108
// (trilinear SDF / normal table lookup).  Reverse mode replays the forward branches and follows the
// Taichi autodiff conventions of SURVEY.md Appendix B.
#pragma once
#include "smx_math.cuh"

namespace smx {

#define SMX_MAXP 8
#define SMX_INF_SDF 1e10f

// one rigid primitive: tables + scalar parameters (device-resident array of these)
struct PrimDev {
    const float* sdf;       // [r0][r1][r2]
    const float4* nrm;      // [r0][r1][r2] (xyz, pad) -- one 16-byte load per corner
    int r0, r1, r2;
    float lo[3], hi[3];
    float inv_dx;
    float friction, softness;
    int enabled;            // primitives_contact[i] (mpm_simulator.py:70)
    int has_table;
};

// pose and twist of a primitive at one frame: [x(3) q(4, w first) v(3) w(3)]
struct PrimState { V3 pos; Q4 rot; V3 v, w; };
struct PrimGrad { V3 pos; Q4 rot; V3 v, w; };

__device__ __forceinline__ PrimState load_prim_state(const float* s13) {
    PrimState s;
    s.pos = v3(s13[0], s13[1], s13[2]);
    s.rot.w = s13[3]; s.rot.x = s13[4]; s.rot.y = s13[5]; s.rot.z = s13[6];
    s.v = v3(s13[7], s13[8], s13[9]); s.w = v3(s13[10], s13[11], s13[12]);
    return s;
}
__device__ __forceinline__ PrimGrad prim_grad_zero() {
    PrimGrad g; g.pos = v3(0, 0, 0); g.rot = q4_zero(); g.v = v3(0, 0, 0); g.w = v3(0, 0, 0); return g;
}

// inv_trans, primitive_utils.py:43-46
__device__ __forceinline__ V3 inv_trans(V3 pos, const PrimState& s) { return qrot(qnormalize(qconj(s.rot)), pos - s.pos); }
__device__ __forceinline__ void inv_trans_adj(V3 pos, const PrimState& s, V3 go, V3* gpos, PrimGrad& G) {
    Q4 c = qconj(s.rot), iq = qnormalize(c);
    V3 d = pos - s.pos, gd = v3(0, 0, 0);
    Q4 giq = q4_zero(), gc = q4_zero();
    qrot_adj(iq, d, go, giq, gd);
    if (gpos) *gpos += gd;
    G.pos -= gd;
    qnormalize_adj(c, giq, gc);
    G.rot.w += gc.w; G.rot.x -= gc.x; G.rot.y -= gc.y; G.rot.z -= gc.z;
}

struct Tri { int b0, b1, b2; float fx, fy, fz; bool in; };
__device__ __forceinline__ Tri tri_setup(const PrimDev& P, V3 pl) {
    Tri t;
    // written so that a NaN coordinate is outside the box (mesh.py:49-51)
    t.in = P.has_table && (pl.x >= P.lo[0] && pl.x < P.hi[0]) && (pl.y >= P.lo[1] && pl.y < P.hi[1]) && (pl.z >= P.lo[2] && pl.z < P.hi[2]);
    if (t.in) {
        float px = (pl.x - P.lo[0]) * P.inv_dx, py = (pl.y - P.lo[1]) * P.inv_dx, pz = (pl.z - P.lo[2]) * P.inv_dx;
        t.b0 = min((int)px, P.r0 - 2); t.b1 = min((int)py, P.r1 - 2); t.b2 = min((int)pz, P.r2 - 2);   // clamp: fp32 rounding at the upper face
        t.fx = px - (float)t.b0; t.fy = py - (float)t.b1; t.fz = pz - (float)t.b2;
    }
    return t;
}
__device__ __forceinline__ int tidx(const PrimDev& P, const Tri& t, int i, int j, int k) { return ((t.b0 + i) * P.r1 + (t.b1 + j)) * P.r2 + (t.b2 + k); }

// Mesh._sdf, mesh.py:45-64 (detail = False)
__device__ __forceinline__ float sdf_local(const PrimDev& P, V3 pl) {
    Tri t = tri_setup(P, pl);
    if (!t.in) return SMX_INF_SDF;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float w = (i ? t.fx : 1.f - t.fx) * (j ? t.fy : 1.f - t.fy) * (k ? t.fz : 1.f - t.fz);
                s = fmaf(w, __ldg(P.sdf + tidx(P, t, i, j, k)), s);
            }
    return s;
}
__device__ __forceinline__ void sdf_local_adj(const PrimDev& P, V3 pl, float gs, V3& gpl) {
    Tri t = tri_setup(P, pl);
    if (!t.in) return;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float wx = (i ? t.fx : 1.f - t.fx), wy = (j ? t.fy : 1.f - t.fy), wz = (k ? t.fz : 1.f - t.fz);
                float v = __ldg(P.sdf + tidx(P, t, i, j, k));
                g0 += (i ? 1.f : -1.f) * wy * wz * v; g1 += wx * (j ? 1.f : -1.f) * wz * v; g2 += wx * wy * (k ? 1.f : -1.f) * v;
            }
    float c = gs * P.inv_dx;
    gpl.x += c * g0; gpl.y += c * g1; gpl.z += c * g2;
}
// Mesh._normal, mesh.py:88-108
__device__ __forceinline__ V3 normal_raw(const PrimDev& P, const Tri& t) {
    V3 r = v3(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float w = (i ? t.fx : 1.f - t.fx) * (j ? t.fy : 1.f - t.fy) * (k ? t.fz : 1.f - t.fz);
                float4 n = __ldg(P.nrm + tidx(P, t, i, j, k));
                r.x = fmaf(w, n.x, r.x); r.y = fmaf(w, n.y, r.y); r.z = fmaf(w, n.z, r.z);
            }
    return r;
}
__device__ __forceinline__ V3 normal_local(const PrimDev& P, V3 pl) {
    Tri t = tri_setup(P, pl);
    if (!t.in) return v3(0.f, 1.f, 0.f);
    return normalize3(normal_raw(P, t));
}
__device__ __forceinline__ void normal_local_adj(const PrimDev& P, V3 pl, V3 gn, V3& gpl) {
    Tri t = tri_setup(P, pl);
    if (!t.in) return;
    V3 r = normal_raw(P, t), gr = v3(0, 0, 0);
    normalize3_adj(r, gn, gr);
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float wx = (i ? t.fx : 1.f - t.fx), wy = (j ? t.fy : 1.f - t.fy), wz = (k ? t.fz : 1.f - t.fz);
                float4 n = __ldg(P.nrm + tidx(P, t, i, j, k));
                float v = n.x * gr.x + n.y * gr.y + n.z * gr.z;
                g0 += (i ? 1.f : -1.f) * wy * wz * v; g1 += wx * (j ? 1.f : -1.f) * wz * v; g2 += wx * wy * (k ? 1.f : -1.f) * v;
            }
    gpl.x += P.inv_dx * g0; gpl.y += P.inv_dx * g1; gpl.z += P.inv_dx * g2;
}

// Primitive.sdf / normal / collider_v, primitive_base.py:53-70
__device__ __forceinline__ float prim_sdf(const PrimDev& P, const PrimState& s, V3 pos) { return sdf_local(P, inv_trans(pos, s)); }
__device__ __forceinline__ void prim_sdf_adj(const PrimDev& P, const PrimState& s, V3 pos, float gs, V3* gpos, PrimGrad& G) {
    V3 pl = inv_trans(pos, s), gpl = v3(0, 0, 0);
    sdf_local_adj(P, pl, gs, gpl);
    inv_trans_adj(pos, s, gpl, gpos, G);
}
__device__ __forceinline__ V3 prim_normal(const PrimDev& P, const PrimState& s, V3 pos) { return qrot(s.rot, normal_local(P, inv_trans(pos, s))); }
__device__ __forceinline__ void prim_normal_adj(const PrimDev& P, const PrimState& s, V3 pos, V3 gn, V3* gpos, PrimGrad& G) {
    V3 pl = inv_trans(pos, s), nl = normal_local(P, pl), gnl = v3(0, 0, 0), gpl = v3(0, 0, 0);
    qrot_adj(s.rot, nl, gn, G.rot, gnl);
    normal_local_adj(P, pl, gnl, gpl);
    inv_trans_adj(pos, s, gpl, gpos, G);
}
__device__ __forceinline__ V3 prim_collider_v(const PrimState& s, V3 r) {
    Q4 qn = qnormalize(s.rot);
    V3 rl = qrot(qconj(qn), r);
    return qrot(qn, s.v + cross(s.w, rl));
}
__device__ __forceinline__ void prim_collider_v_adj(const PrimState& s, V3 r, V3 go, V3& gr, PrimGrad& G) {
    Q4 qn = qnormalize(s.rot), iq = qconj(qn);
    V3 rl = qrot(iq, r), cl = s.v + cross(s.w, rl);
    Q4 gqn = q4_zero(), giq = q4_zero();
    V3 gcl = v3(0, 0, 0), grl;
    qrot_adj(qn, cl, go, gqn, gcl);
    G.v += gcl;
    G.w += cross(rl, gcl);
    grl = cross(gcl, s.w);
    qrot_adj(iq, r, grl, giq, gr);
    gqn.w += giq.w; gqn.x -= giq.x; gqn.y -= giq.y; gqn.z -= giq.z;
    qnormalize_adj(s.rot, gqn, G.rot);
}

// friction projection shared by collide and collide_mixed (primitive_base.py:86-89, 153-156)
__device__ __forceinline__ V3 friction_proj(V3 t, float nc, float fric) {
    float tt = dot(t, t), tn = sqrtf(tt + 1e-8f);
    bool flag = (nc < 0.f) && (sqrtf(tt) > 1e-30f);
    return flag ? (fmaxf(0.f, tn + nc * fric) / tn) * t : t;
}
__device__ __forceinline__ void friction_proj_adj(V3 t, float nc, float fric, V3 go, V3& gt, float& gnc) {
    float tt = dot(t, t), tn = sqrtf(tt + 1e-8f), b = tn + nc * fric, mx = fmaxf(0.f, b);
    bool flag = (nc < 0.f) && (sqrtf(tt) > 1e-30f);
    if (!flag) { gt += go; return; }
    float tg = dot(t, go), gmx = tg / tn, gtn = -tg * mx / (tn * tn);
    gt += (mx / tn) * go;
    if (!(b < 0.f)) { gtn += gmx; gnc += gmx * fric; }    // max(0, b): gradient to b unless b < 0
    gt += (gtn / tn) * t;
}

// ---------------------------------------------------------------------------------------------
// collide_mixed -- the forecast-based contact model, primitive_base.py:139-181
// ---------------------------------------------------------------------------------------------
struct CmTape {
    bool active, moving_in, outside, pen;
    float nc, infl, e, s;
    V3 D, r, cv, iv, vt0, vt, x_new, n;
};

// returns the new particle velocity; bf = p_mass (v_in - v_out)/dt is the reaction on the body
__device__ __forceinline__ V3 collide_mixed_fwd(const PrimDev& P, const PrimState& S, V3 x, V3 pv_in, float dt, float life, CmTape& T) {
    float dist = prim_sdf(P, S, x);
    T.active = dist <= 5e-3f;
    if (!T.active) return pv_in;
    V3 pv = pv_in;
    T.D = prim_normal(P, S, x);
    T.r = x - S.pos;
    T.cv = prim_collider_v(S, T.r);
    T.iv = pv - T.cv;
    T.nc = dot(T.iv, T.D);
    T.moving_in = T.nc < 0.f;
    T.outside = false;
    if (T.moving_in) {
        T.vt0 = T.iv - T.nc * T.D;
        T.vt = friction_proj(T.vt0, T.nc, P.friction);
        pv = T.cv + T.vt;
        if (dist > 0.f) {
            T.outside = true;
            T.e = expf(-dist * P.softness);
            T.infl = fminf(T.e, 1.f);
            pv = T.cv + (1.f - T.infl) * T.iv + T.infl * T.vt;
        }
    }
    T.x_new = x + dt * pv;
    T.s = prim_sdf(P, S, T.x_new);
    T.pen = T.s < 0.f;
    if (T.pen) {
        T.n = prim_normal(P, S, T.x_new);
        pv = pv - ((T.s / dt) * life) * T.n;
    }
    return pv;
}
// gout: adjoint of the returned velocity; gext: ext_f.grad (6). Accumulates gx, gpv_in, G.
__device__ __forceinline__ void collide_mixed_adj(const PrimDev& P, const PrimState& S, V3 x, V3 pv_in, V3 pv_out, float p_mass, float dt,
                                                  float life, const CmTape& T, V3 gout, const float* gext, V3& gx, V3& gpv_in, PrimGrad& G) {
    if (!T.active) { gpv_in += gout; return; }
    float dist = prim_sdf(P, S, x);
    float c = p_mass / dt;
    V3 ge_f = v3(gext[0], gext[1], gext[2]), ge_t = v3(gext[3], gext[4], gext[5]);
    V3 bf = c * (pv_in - pv_out);
    V3 g_r = cross(bf, ge_t);                 // b_t = r x b_f
    V3 g_bf = ge_f + cross(ge_t, T.r);
    V3 g_vin = c * g_bf, g_pv = gout - c * g_bf;
    V3 g_xnew = v3(0, 0, 0);
    float g_s = 0.f;
    if (T.pen) {
        g_s = -(life / dt) * dot(g_pv, T.n);
        V3 g_n = (-(T.s / dt) * life) * g_pv;
        prim_normal_adj(P, S, T.x_new, g_n, &g_xnew, G);
    }
    prim_sdf_adj(P, S, T.x_new, g_s, &g_xnew, G);
    V3 g_mid = g_pv + dt * g_xnew;
    gx += g_xnew;
    V3 g_cv = v3(0, 0, 0), g_iv = v3(0, 0, 0), g_D = v3(0, 0, 0);
    float g_nc = 0.f, g_dist = 0.f;
    if (T.moving_in) {
        V3 g_vt;
        if (T.outside) {
            g_cv += g_mid; g_iv += (1.f - T.infl) * g_mid; g_vt = T.infl * g_mid;
            float g_infl = dot(g_mid, T.vt - T.iv);
            if (T.e < 1.f) g_dist += g_infl * (-P.softness) * T.e;
        } else { g_cv += g_mid; g_vt = g_mid; }
        V3 g_vt0 = v3(0, 0, 0);
        friction_proj_adj(T.vt0, T.nc, P.friction, g_vt, g_vt0, g_nc);
        g_iv += g_vt0; g_nc -= dot(g_vt0, T.D); g_D -= T.nc * g_vt0;
    } else g_vin += g_mid;
    g_iv += g_nc * T.D; g_D += g_nc * T.iv;
    g_vin += g_iv; g_cv -= g_iv;
    prim_collider_v_adj(S, T.r, g_cv, g_r, G);
    gx += g_r; G.pos -= g_r;
    prim_normal_adj(P, S, x, g_D, &gx, G);
    prim_sdf_adj(P, S, x, g_dist, &gx, G);
    (void)dist;
    gpv_in += g_vin;
}

// ---------------------------------------------------------------------------------------------
// collide -- grid contact, primitive_base.py:72-103
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 collide_grid_fwd(const PrimDev& P, const PrimState& S, V3 gp, V3 v_in, bool& active, V3& r_out) {
    float dist = prim_sdf(P, S, gp);
    float infl = fminf(expf(-dist * P.softness), 1.f);
    active = (P.softness > 0.f && infl > 0.1f) || dist <= 0.f;
    if (!active) return v_in;
    V3 D = prim_normal(P, S, gp), r = gp - S.pos, cv = prim_collider_v(S, r), iv = v_in - cv;
    float nc = dot(iv, D);
    V3 vt = friction_proj(iv - fminf(nc, 0.f) * D, nc, P.friction);
    r_out = r;
    return cv + (1.f - infl) * iv + infl * vt;
}
__device__ __forceinline__ void collide_grid_adj(const PrimDev& P, const PrimState& S, V3 gp, V3 v_in, float dt, float gm, V3 gvout,
                                                 const float* gext, V3& gvin, float& ggm, PrimGrad& G) {
    float dist = prim_sdf(P, S, gp);
    float e = expf(-dist * P.softness), infl = fminf(e, 1.f);
    if (!((P.softness > 0.f && infl > 0.1f) || dist <= 0.f)) { gvin += gvout; return; }
    V3 D = prim_normal(P, S, gp), r = gp - S.pos, cv = prim_collider_v(S, r), iv = v_in - cv;
    float nc = dot(iv, D), mn = fminf(nc, 0.f);
    V3 vt0 = iv - mn * D, vtf = friction_proj(vt0, nc, P.friction);
    V3 v_out = cv + (1.f - infl) * iv + infl * vtf;
    float c = gm / dt;
    V3 ge_f = v3(gext[0], gext[1], gext[2]), ge_t = v3(gext[3], gext[4], gext[5]);
    V3 bf = c * (v_in - v_out);
    V3 g_r = cross(bf, ge_t), g_bf = ge_f + cross(ge_t, r);
    ggm += dot(g_bf, v_in - v_out) / dt;
    V3 g_vi = c * g_bf, g_vo = gvout - c * g_bf;
    V3 g_cv = g_vo, g_iv = (1.f - infl) * g_vo, g_vtf = infl * g_vo;
    float g_infl = dot(g_vo, vtf - iv), g_nc = 0.f;
    V3 g_vt = v3(0, 0, 0), g_D = v3(0, 0, 0);
    friction_proj_adj(vt0, nc, P.friction, g_vtf, g_vt, g_nc);
    g_iv += g_vt;
    float g_mn = -dot(g_vt, D);
    g_D -= mn * g_vt;
    if (nc < 0.f) g_nc += g_mn;                 // min(nc, 0): gradient to nc iff nc < 0
    g_iv += g_nc * D; g_D += g_nc * iv;
    g_vi += g_iv; g_cv -= g_iv;
    prim_collider_v_adj(S, r, g_cv, g_r, G);
    G.pos -= g_r;                               // grid_pos carries no gradient
    prim_normal_adj(P, S, gp, g_D, nullptr, G);
    float g_dist = (e < 1.f) ? g_infl * (-P.softness) * e : 0.f;
    prim_sdf_adj(P, S, gp, g_dist, nullptr, G);
    gvin += g_vi;
}

// ---------------------------------------------------------------------------------------------
// collide_particle -- penalty contact, primitive_base.py:105-137.  Returns the impulse p_f * dt and
// the reaction force b_f on the body (zero when inactive).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 collide_particle_fwd(const PrimDev& P, const PrimState& S, V3 x, V3 pv, float dt, bool& active, V3& bf, V3& r_out) {
    float c = prim_sdf(P, S, x) - 5e-3f;
    active = c < 0.f;
    bf = v3(0, 0, 0);
    if (!active) return v3(0, 0, 0);
    V3 D = prim_normal(P, S, x), r = x - S.pos, cv = prim_collider_v(S, r), iv = pv - cv;
    float nc = dot(iv, D);
    V3 vt = iv - nc * D;
    float vtn = sqrtf(dot(vt, vt) + 1e-8f);
    V3 f = (-c * 50.f) * D - (fabsf(nc) * P.friction / vtn) * vt;
    bf = -f; r_out = r;
    return dt * f;
}
__device__ __forceinline__ void collide_particle_adj(const PrimDev& P, const PrimState& S, V3 x, V3 pv, float dt, V3 gimp, const float* gext,
                                                     V3& gx, V3& gpv, PrimGrad& G) {
    float c = prim_sdf(P, S, x) - 5e-3f;
    if (!(c < 0.f)) return;
    const float k1 = 50.f, kf = P.friction;
    V3 D = prim_normal(P, S, x), r = x - S.pos, cv = prim_collider_v(S, r), iv = pv - cv;
    float nc = dot(iv, D);
    V3 vt = iv - nc * D;
    float vtn = sqrtf(dot(vt, vt) + 1e-8f);
    V3 f = (-c * k1) * D - (fabsf(nc) * kf / vtn) * vt;
    V3 bf = -f;
    V3 ge_f = v3(gext[0], gext[1], gext[2]), ge_t = v3(gext[3], gext[4], gext[5]);
    V3 g_r = cross(bf, ge_t), g_bf = ge_f + cross(ge_t, r);
    V3 g_f = dt * gimp - g_bf;
    V3 g_D = (-c * k1) * g_f;
    float g_c = -k1 * dot(g_f, D);
    V3 g_vt = (-fabsf(nc) * kf / vtn) * g_f;
    float g_vtn = dot(g_f, vt) * fabsf(nc) * kf / (vtn * vtn);
    float g_abs = -dot(g_f, vt) * kf / vtn;
    g_vt += (g_vtn / vtn) * vt;
    float g_nc = g_abs * (nc > 0.f ? 1.f : (nc < 0.f ? -1.f : 0.f));
    V3 g_iv = g_vt;
    g_nc -= dot(g_vt, D); g_D -= nc * g_vt;
    g_iv += g_nc * D; g_D += g_nc * iv;
    gpv += g_iv;
    V3 g_cv = -g_iv;
    prim_collider_v_adj(S, r, g_cv, g_r, G);
    gx += g_r; G.pos -= g_r;
    prim_normal_adj(P, S, x, g_D, &gx, G);
    prim_sdf_adj(P, S, x, g_c, &gx, G);
}

}  // namespace smx
