/*
 * mpm_oracle.c -- float64 CPU restatement of SoftMAC's differentiable MLS-MPM substep
 * (forward + adjoint) and of the primitive contact model it calls.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product (softmac_b200/) never
 * links, imports or falls back to anything in oracle/.
 *
 * PARITY UNPINNED: the reference is Python + Taichi 1.4.1 (+ Jade).  None of them can be installed
 * in this image and the reference ships no tests, golden vectors or known-answer fixtures for this
 * path, so this restatement is pinned only by (i) line-by-line correspondence with the cited
 * reference source, (ii) central finite differences of its own forward for every adjoint, (iii)
 * conservation / reconstruction properties.  See oracle/README.md and DESIGN.md.
 *
 * Structure mirrors the reference one-to-one: one function per @ti.kernel, one parallel-for per
 * kernel, `omp atomic` wherever Taichi uses an atomic add, per-frame value *and* gradient fields.
 * Citations are relative to /root/reference/.
 *
 *   fields / constants         softmac/engine/mpm_simulator.py:17-90
 *   clear_grid                 softmac/engine/mpm_simulator.py:93-114
 *   clear_SVD_grad             :116-123
 *   compute_F_tmp              :125-128
 *   svd                        :130-133   (ti.svd, taichi==1.4.1 -- third party, restated: see svd3)
 *   svd_grad / backward_svd    :135-157, clamp :184-192
 *   p2g                        :198-262
 *   boundary_condition         :268-281
 *   grid_op                    :283-297
 *   g2p                        :299-318
 *   substep / substep_grad     :320-378
 *   grid_op_mixed1..4          :396-443
 *   Primitive.sdf/normal       softmac/engine/primitive/primitive_base.py:53-61
 *   Primitive.collider_v       :63-70
 *   Primitive.collide          :72-103
 *   Primitive.collide_particle :105-137
 *   Primitive.collide_mixed    :139-181
 *   forward_kinematics         :280-283
 *   Mesh._sdf/_normal          softmac/engine/primitive/mesh.py:45-108
 *   qrot/qmul/w2quat/inv_trans softmac/engine/primitive/primitive_utils.py:4-46
 *
 * Adjoints are hand-written reverse mode of exactly those statements, following the Taichi autodiff
 * conventions listed in SURVEY.md Appendix B (zero gradient through int casts and comparisons,
 * branch replay, max/min sub-gradient choice, abs -> sign, eps-guarded sqrt differentiated as
 * written).  svd_grad is the reference's explicit formula.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXP 8
#define INF_SDF 1e10

typedef struct {
    double friction, softness;
    int enabled;            /* primitives_contact[i], mpm_simulator.py:70 */
    int has_table;
    int res[3];
    double lower[3], upper[3], sdf_dx, inv_sdf_dx;
    double *sdf;            /* [res0][res1][res2]       mesh.py:36 */
    double *nrm;            /* [res0][res1][res2][3]    mesh.py:37 */
    double *pos, *rot, *v, *w;          /* [T][3],[T][4],[T][3],[T][3]  primitive_base.py:28-36 */
    double *gpos, *grot, *gv, *gw;
    double ext_f[6], ext_f_grad[6];     /* primitive_base.py:39 */
    double *abuf, *gabuf;               /* action_buffer [T][6], primitive_base.py:43 */
} orc_prim;

typedef struct {
    int n, ng, T;
    double dt, dx, inv_dx, p_vol, p_mass, mu, lam;
    double gravity[3];
    int sticky_ground;      /* ground_friction >= 10, mpm_simulator.py:278 */
    int material_model, ptype, collision_type, substeps, n_control;
    int rigid_velocity_control;
    int plasticity;         /* 0: sigma clip (softmac mpm_simulator.py:226-229); 1: von Mises return mapping (soft_cloth/engine/mpm_simulator.py:232) */
    double yield_stress;    /* cfg.yield_stress (soft_cloth/engine/mpm_simulator.py:20,92) */
    int np;
    orc_prim prim[ORC_MAXP];
    double *x, *v, *C, *F, *gx, *gv, *gC, *gF;              /* [T][n][3|9] */
    double *Ftmp, *U, *S, *V, *gFtmp, *gU, *gS, *gV;        /* [n][9] */
    double *gvin, *gvout, *gm, *gvmix;                      /* grid_v_in, grid_v_out, grid_m, grid_v_mixed */
    double *ggvin, *ggvout, *ggm, *ggvmix;                  /* their .grad */
    double *vtmp, *vtgt, *gvtmp, *gvtgt;                    /* [n][3] */
    int *control_idx;
    double *action, *gaction;                               /* [n_control][3] */
} orc_sim;

/* ------------------------------------------------------------------------------------------ */
/* small linear algebra                                                                        */
/* ------------------------------------------------------------------------------------------ */
static inline double dot3(const double *a, const double *b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static inline void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1]*b[2] - a[2]*b[1]; c[1] = a[2]*b[0] - a[0]*b[2]; c[2] = a[0]*b[1] - a[1]*b[0];
}
/* c += a x b */
static inline void cross3_acc(const double *a, const double *b, double *c) {
    c[0] += a[1]*b[2] - a[2]*b[1]; c[1] += a[2]*b[0] - a[0]*b[2]; c[2] += a[0]*b[1] - a[1]*b[0];
}
static inline void mm(const double *A, const double *B, double *C) {        /* C = A B */
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        t[3*i+j] = A[3*i]*B[j] + A[3*i+1]*B[3+j] + A[3*i+2]*B[6+j];
    memcpy(C, t, sizeof t);
}
static inline void mmT(const double *A, const double *B, double *C) {       /* C = A B^T */
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        t[3*i+j] = A[3*i]*B[3*j] + A[3*i+1]*B[3*j+1] + A[3*i+2]*B[3*j+2];
    memcpy(C, t, sizeof t);
}
static inline void mTm(const double *A, const double *B, double *C) {       /* C = A^T B */
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        t[3*i+j] = A[i]*B[j] + A[3+i]*B[3+j] + A[6+i]*B[6+j];
    memcpy(C, t, sizeof t);
}
static inline double det3(const double *A) {
    return A[0]*(A[4]*A[8]-A[5]*A[7]) - A[1]*(A[3]*A[8]-A[5]*A[6]) + A[2]*(A[3]*A[7]-A[4]*A[6]);
}
static inline void cof3(const double *A, double *K) {   /* d det / dA */
    K[0] =  (A[4]*A[8]-A[5]*A[7]); K[1] = -(A[3]*A[8]-A[5]*A[6]); K[2] =  (A[3]*A[7]-A[4]*A[6]);
    K[3] = -(A[1]*A[8]-A[2]*A[7]); K[4] =  (A[0]*A[8]-A[2]*A[6]); K[5] = -(A[0]*A[7]-A[1]*A[6]);
    K[6] =  (A[1]*A[5]-A[2]*A[4]); K[7] = -(A[0]*A[5]-A[2]*A[3]); K[8] =  (A[0]*A[4]-A[1]*A[3]);
}

/* ------------------------------------------------------------------------------------------ */
/* ti.svd restated (third party: taichi==1.4.1, requirements.txt:2; call site                   */
/* mpm_simulator.py:133).  Published contract of the McAdams/Sifakis 3x3 SVD that Taichi uses:  */
/* F = U diag(s) V^T, U and V proper rotations, s sorted by decreasing magnitude, the sign of   */
/* det F carried by s[2].  Implemented here as cyclic Jacobi on F^T F to convergence in f64     */
/* (every quantity the path derives from it -- U S V^T, U V^T and the svd_grad formula -- is    */
/* invariant to the remaining basis freedom except at exactly repeated singular values).        */
/* ------------------------------------------------------------------------------------------ */
static void jacobi_eig3(double A[9], double V[9]) {
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = A[1]*A[1] + A[2]*A[2] + A[5]*A[5];
        double dg = A[0]*A[0] + A[4]*A[4] + A[8]*A[8];
        if (off <= 1e-34 * dg || off == 0.0) break;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            double apq = A[3*p+q];
            if (apq == 0.0) continue;
            double app = A[3*p+p], aqq = A[3*q+q];
            double theta = (aqq - app) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta*theta + 1.0));
            double c = 1.0 / sqrt(t*t + 1.0), s = t * c;
            /* A <- J^T A J with J = [[c, s], [-s, c]] on (p,q) */
            for (int k = 0; k < 3; k++) {
                double akp = A[3*k+p], akq = A[3*k+q];
                A[3*k+p] = c*akp - s*akq; A[3*k+q] = s*akp + c*akq;
            }
            for (int k = 0; k < 3; k++) {
                double apk = A[3*p+k], aqk = A[3*q+k];
                A[3*p+k] = c*apk - s*aqk; A[3*q+k] = s*apk + c*aqk;
            }
            for (int k = 0; k < 3; k++) {
                double vkp = V[3*k+p], vkq = V[3*k+q];
                V[3*k+p] = c*vkp - s*vkq; V[3*k+q] = s*vkp + c*vkq;
            }
        }
    }
}

static void svd3(const double *F, double *U, double *S /* 3x3 diag */, double *V) {
    double A[9], Vv[9];
    mTm(F, F, A);
    jacobi_eig3(A, Vv);
    double lam[3] = {A[0], A[4], A[8]};
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2 - i; j++)
        if (lam[idx[j]] < lam[idx[j+1]]) { int t = idx[j]; idx[j] = idx[j+1]; idx[j+1] = t; }
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) V[3*r+c] = Vv[3*r+idx[c]];
    if (det3(V) < 0) for (int r = 0; r < 3; r++) V[3*r+2] = -V[3*r+2];
    double B[9]; mm(F, V, B);
    double b1[3] = {B[0], B[3], B[6]}, b2[3] = {B[1], B[4], B[7]}, b3[3] = {B[2], B[5], B[8]};
    double u1[3], u2[3], u3[3];
    double s1 = sqrt(dot3(b1, b1));
    if (s1 > 1e-300) { for (int i = 0; i < 3; i++) u1[i] = b1[i]/s1; } else { u1[0] = 1; u1[1] = u1[2] = 0; }
    double d = dot3(u1, b2);
    for (int i = 0; i < 3; i++) u2[i] = b2[i] - d*u1[i];
    double s2 = sqrt(dot3(u2, u2));
    if (s2 > 1e-300 && s2 > 1e-14 * s1) { for (int i = 0; i < 3; i++) u2[i] /= s2; }
    else {  /* rank <= 1: any unit vector orthogonal to u1 */
        int k = fabs(u1[0]) < fabs(u1[1]) ? (fabs(u1[0]) < fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) < fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[k] = 1; double dd = dot3(u1, e);
        for (int i = 0; i < 3; i++) u2[i] = e[i] - dd*u1[i];
        double nn = sqrt(dot3(u2, u2)); for (int i = 0; i < 3; i++) u2[i] /= nn;
    }
    cross3(u1, u2, u3);
    double s3 = dot3(u3, b3);
    s2 = dot3(u2, b2);
    for (int r = 0; r < 3; r++) { U[3*r] = u1[r]; U[3*r+1] = u2[r]; U[3*r+2] = u3[r]; }
    memset(S, 0, 9 * sizeof(double));
    S[0] = s1; S[4] = s2; S[8] = s3;
}

/* mpm_simulator.py:184-192 */
static inline double clamp_ref(double a) { return a >= 0 ? fmax(a, 1e-6) : fmin(a, -1e-6); }

/* mpm_simulator.py:140-157, literal */
static void backward_svd(const double *gu, const double *gsig, const double *gv,
                         const double *u, const double *sig, const double *v, double *out) {
    double s[3] = {sig[0]*sig[0], sig[4]*sig[4], sig[8]*sig[8]};
    double Fm[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        Fm[3*i+j] = (i == j) ? 0.0 : 1.0 / clamp_ref(s[j] - s[i]);
    double utgu[9], gutu[9], vtgv[9], gvtv[9], t1[9], t2[9], sigma_term[9], u_term[9], v_term[9];
    mmT(gsig, v, t1); mm(u, t1, sigma_term);                  /* u @ gsigma @ vt */
    mTm(u, gu, utgu); mTm(gu, u, gutu);
    for (int i = 0; i < 9; i++) t1[i] = Fm[i] * (utgu[i] - gutu[i]);
    mm(t1, sig, t2); mmT(t2, v, t1); mm(u, t1, u_term);       /* u @ ((F*(ut@gu - gut@u)) @ sig) @ vt */
    mTm(v, gv, vtgv); mTm(gv, v, gvtv);
    for (int i = 0; i < 9; i++) t1[i] = Fm[i] * (vtgv[i] - gvtv[i]);
    mmT(t1, v, t2); mm(sig, t2, t1); mm(u, t1, v_term);       /* u @ (sig @ ((F*(vt@gv - gvt@v)) @ vt)) */
    for (int i = 0; i < 9; i++) out[i] = u_term[i] + v_term[i] + sigma_term[i];
}

/* ------------------------------------------------------------------------------------------ */
/* quaternion helpers, primitive_utils.py:4-46, and their reverse mode                         */
/* ------------------------------------------------------------------------------------------ */
static inline double length_eps(const double *x) { return sqrt(dot3(x, x) + 1e-8); }   /* :4-5 */

static void qrot(const double *q, const double *v, double *out) {                       /* :8-13 */
    double uv[3], uuv[3];
    cross3(q + 1, v, uv); cross3(q + 1, uv, uuv);
    for (int i = 0; i < 3; i++) out[i] = v[i] + 2.0 * (q[0]*uv[i] + uuv[i]);
}
/* accumulates into gq[4], gv[3] */
static void qrot_adj(const double *q, const double *v, const double *go, double *gq, double *gv) {
    double uv[3], guv[3], guuv[3];
    cross3(q + 1, v, uv);
    for (int i = 0; i < 3; i++) { gv[i] += go[i]; guv[i] = 2.0*q[0]*go[i]; guuv[i] = 2.0*go[i]; }
    gq[0] += 2.0 * dot3(uv, go);
    /* uuv = qv x uv */
    cross3_acc(uv, guuv, gq + 1);
    cross3_acc(guuv, q + 1, guv);
    /* uv = qv x v */
    cross3_acc(v, guv, gq + 1);
    cross3_acc(guv, q + 1, gv);
}
static inline void normalize_n(const double *x, int n, double *y) {
    double s = 0; for (int i = 0; i < n; i++) s += x[i]*x[i];
    s = sqrt(s); for (int i = 0; i < n; i++) y[i] = x[i]/s;
}
/* y = x/|x| ; gx += (gy - y (y.gy))/|x| */
static inline void normalize_adj(const double *x, int n, const double *gy, double *gx) {
    double s = 0, y[4], d = 0;
    for (int i = 0; i < n; i++) s += x[i]*x[i];
    s = sqrt(s);
    for (int i = 0; i < n; i++) { y[i] = x[i]/s; d += y[i]*gy[i]; }
    for (int i = 0; i < n; i++) gx[i] += (gy[i] - y[i]*d)/s;
}
static void inv_trans(const double *pos, const double *position, const double *rot, double *out) {  /* :43-46 */
    double c[4] = {rot[0], -rot[1], -rot[2], -rot[3]}, iq[4], d[3];
    normalize_n(c, 4, iq);
    for (int i = 0; i < 3; i++) d[i] = pos[i] - position[i];
    qrot(iq, d, out);
}
static void inv_trans_adj(const double *pos, const double *position, const double *rot, const double *go,
                          double *gpos /* may be NULL */, double *gposition, double *grot) {
    double c[4] = {rot[0], -rot[1], -rot[2], -rot[3]}, iq[4], d[3];
    normalize_n(c, 4, iq);
    for (int i = 0; i < 3; i++) d[i] = pos[i] - position[i];
    double giq[4] = {0, 0, 0, 0}, gd[3] = {0, 0, 0}, gc[4] = {0, 0, 0, 0};
    qrot_adj(iq, d, go, giq, gd);
    for (int i = 0; i < 3; i++) { if (gpos) gpos[i] += gd[i]; gposition[i] -= gd[i]; }
    normalize_adj(c, 4, giq, gc);
    grot[0] += gc[0]; grot[1] -= gc[1]; grot[2] -= gc[2]; grot[3] -= gc[3];
}
static void qmul(const double *q, const double *r, double *out) {                         /* :20-27 */
    /* terms = r.outer_product(q): terms[i][j] = r[i]*q[j] */
    double w = r[0]*q[0] - r[1]*q[1] - r[2]*q[2] - r[3]*q[3];
    double x = r[0]*q[1] + r[1]*q[0] - r[2]*q[3] + r[3]*q[2];
    double y = r[0]*q[2] + r[1]*q[3] + r[2]*q[0] - r[3]*q[1];
    double z = r[0]*q[3] - r[1]*q[2] + r[2]*q[1] + r[3]*q[0];
    double o[4] = {w, x, y, z};
    normalize_n(o, 4, out);
}
static void qmul_adj(const double *q, const double *r, const double *go, double *gq, double *gr) {
    double o[4];
    o[0] = r[0]*q[0] - r[1]*q[1] - r[2]*q[2] - r[3]*q[3];
    o[1] = r[0]*q[1] + r[1]*q[0] - r[2]*q[3] + r[3]*q[2];
    o[2] = r[0]*q[2] + r[1]*q[3] + r[2]*q[0] - r[3]*q[1];
    o[3] = r[0]*q[3] - r[1]*q[2] + r[2]*q[1] + r[3]*q[0];
    double g[4] = {0, 0, 0, 0};
    normalize_adj(o, 4, go, g);
    gr[0] += g[0]*q[0] + g[1]*q[1] + g[2]*q[2] + g[3]*q[3];
    gr[1] += -g[0]*q[1] + g[1]*q[0] + g[2]*q[3] - g[3]*q[2];
    gr[2] += -g[0]*q[2] - g[1]*q[3] + g[2]*q[0] + g[3]*q[1];
    gr[3] += -g[0]*q[3] + g[1]*q[2] - g[2]*q[1] + g[3]*q[0];
    gq[0] += g[0]*r[0] + g[1]*r[1] + g[2]*r[2] + g[3]*r[3];
    gq[1] += -g[0]*r[1] + g[1]*r[0] - g[2]*r[3] + g[3]*r[2];
    gq[2] += -g[0]*r[2] + g[1]*r[3] + g[2]*r[0] - g[3]*r[1];
    gq[3] += -g[0]*r[3] - g[1]*r[2] + g[2]*r[1] + g[3]*r[0];
}
static void w2quat(const double *aa, double *out) {                                        /* :30-40 */
    double w = sqrt(dot3(aa, aa) + 1e-12);      /* axis_angle.norm(1e-12) = sqrt(sum sq + eps) */
    double s = sin(w/2);
    out[0] = cos(w/2);
    for (int i = 0; i < 3; i++) out[i+1] = aa[i]/w*s;
}
static void w2quat_adj(const double *aa, const double *go, double *gaa) {
    double w = sqrt(dot3(aa, aa) + 1e-12);
    double s = sin(w/2), c = cos(w/2);
    double gw = -0.5*s*go[0];
    for (int i = 0; i < 3; i++) {
        gaa[i] += go[i+1]*s/w;
        gw += go[i+1]*aa[i]*(0.5*c/w - s/(w*w));
    }
    for (int i = 0; i < 3; i++) gaa[i] += gw*aa[i]/w;
}

/* ------------------------------------------------------------------------------------------ */
/* Mesh SDF / normal table lookup, mesh.py:45-108 (detail=False path), and reverse mode        */
/* ------------------------------------------------------------------------------------------ */
static inline int in_box(const orc_prim *P, const double *pl) {
    /* written so that a NaN coordinate (frame never filled: all-zero quaternion) is outside the box */
    for (int i = 0; i < 3; i++) if (!(pl[i] >= P->lower[i] && pl[i] < P->upper[i])) return 0;
    return 1;
}
static inline void tri_setup(const orc_prim *P, const double *pl, int *base, double *fx) {
    for (int i = 0; i < 3; i++) {
        double pos = (pl[i] - P->lower[i]) * P->inv_sdf_dx;
        base[i] = (int)pos;
        fx[i] = pos - base[i];
    }
}
static double sdf_local(const orc_prim *P, const double *pl) {
    if (!P->has_table || !in_box(P, pl)) return INF_SDF;
    int b[3]; double fx[3]; tri_setup(P, pl, b, fx);
    double s = 0;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        double wgt = (i ? fx[0] : 1.-fx[0]) * (j ? fx[1] : 1.-fx[1]) * (k ? fx[2] : 1.-fx[2]);
        s += wgt * P->sdf[((size_t)(b[0]+i)*P->res[1] + (b[1]+j))*P->res[2] + (b[2]+k)];
    }
    return s;
}
static void sdf_local_adj(const orc_prim *P, const double *pl, double gs, double *gpl) {
    if (!P->has_table || !in_box(P, pl)) return;
    int b[3]; double fx[3]; tri_setup(P, pl, b, fx);
    double g[3] = {0, 0, 0};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        double wx = (i ? fx[0] : 1.-fx[0]), wy = (j ? fx[1] : 1.-fx[1]), wz = (k ? fx[2] : 1.-fx[2]);
        double t = P->sdf[((size_t)(b[0]+i)*P->res[1] + (b[1]+j))*P->res[2] + (b[2]+k)];
        g[0] += (i ? 1. : -1.) * wy * wz * t;
        g[1] += wx * (j ? 1. : -1.) * wz * t;
        g[2] += wx * wy * (k ? 1. : -1.) * t;
    }
    for (int a = 0; a < 3; a++) gpl[a] += gs * g[a] * P->inv_sdf_dx;
}
static void normal_local(const orc_prim *P, const double *pl, double *n) {
    if (!P->has_table || !in_box(P, pl)) { n[0] = 0; n[1] = 1; n[2] = 0; return; }
    int b[3]; double fx[3]; tri_setup(P, pl, b, fx);
    double r[3] = {0, 0, 0};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        double wgt = (i ? fx[0] : 1.-fx[0]) * (j ? fx[1] : 1.-fx[1]) * (k ? fx[2] : 1.-fx[2]);
        const double *t = P->nrm + 3*(((size_t)(b[0]+i)*P->res[1] + (b[1]+j))*P->res[2] + (b[2]+k));
        r[0] += wgt*t[0]; r[1] += wgt*t[1]; r[2] += wgt*t[2];
    }
    normalize_n(r, 3, n);
}
static void normal_local_adj(const orc_prim *P, const double *pl, const double *gn, double *gpl) {
    if (!P->has_table || !in_box(P, pl)) return;
    int b[3]; double fx[3]; tri_setup(P, pl, b, fx);
    double r[3] = {0, 0, 0};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        double wgt = (i ? fx[0] : 1.-fx[0]) * (j ? fx[1] : 1.-fx[1]) * (k ? fx[2] : 1.-fx[2]);
        const double *t = P->nrm + 3*(((size_t)(b[0]+i)*P->res[1] + (b[1]+j))*P->res[2] + (b[2]+k));
        r[0] += wgt*t[0]; r[1] += wgt*t[1]; r[2] += wgt*t[2];
    }
    double gr[3] = {0, 0, 0};
    normalize_adj(r, 3, gn, gr);
    double g[3] = {0, 0, 0};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        double wx = (i ? fx[0] : 1.-fx[0]), wy = (j ? fx[1] : 1.-fx[1]), wz = (k ? fx[2] : 1.-fx[2]);
        const double *t = P->nrm + 3*(((size_t)(b[0]+i)*P->res[1] + (b[1]+j))*P->res[2] + (b[2]+k));
        double tv = dot3(t, gr);
        g[0] += (i ? 1. : -1.) * wy * wz * tv;
        g[1] += wx * (j ? 1. : -1.) * wz * tv;
        g[2] += wx * wy * (k ? 1. : -1.) * tv;
    }
    for (int a = 0; a < 3; a++) gpl[a] += g[a] * P->inv_sdf_dx;
}

/* per-call gradient sink for the primitive state at frame f (private per thread, reduced later) */
typedef struct { double pos[3], rot[4], v[3], w[3], m; } prim_grad;

/* Primitive.sdf, primitive_base.py:53-56 */
static double prim_sdf(const orc_prim *P, int f, const double *pos) {
    double pl[3]; inv_trans(pos, P->pos + 3*f, P->rot + 4*f, pl);
    return sdf_local(P, pl);
}
static void prim_sdf_adj(const orc_prim *P, int f, const double *pos, double gs, double *gpos, prim_grad *G) {
    double pl[3], gpl[3] = {0, 0, 0};
    inv_trans(pos, P->pos + 3*f, P->rot + 4*f, pl);
    sdf_local_adj(P, pl, gs, gpl);
    inv_trans_adj(pos, P->pos + 3*f, P->rot + 4*f, gpl, gpos, G->pos, G->rot);
}
/* Primitive.normal, primitive_base.py:58-61 */
static void prim_normal(const orc_prim *P, int f, const double *pos, double *n) {
    double pl[3], nl[3]; inv_trans(pos, P->pos + 3*f, P->rot + 4*f, pl);
    normal_local(P, pl, nl);
    qrot(P->rot + 4*f, nl, n);
}
static void prim_normal_adj(const orc_prim *P, int f, const double *pos, const double *gn, double *gpos, prim_grad *G) {
    double pl[3], nl[3], gnl[3] = {0, 0, 0}, gpl[3] = {0, 0, 0};
    inv_trans(pos, P->pos + 3*f, P->rot + 4*f, pl);
    normal_local(P, pl, nl);
    qrot_adj(P->rot + 4*f, nl, gn, G->rot, gnl);
    normal_local_adj(P, pl, gnl, gpl);
    inv_trans_adj(pos, P->pos + 3*f, P->rot + 4*f, gpl, gpos, G->pos, G->rot);
}
/* Primitive.collider_v, primitive_base.py:63-70 */
static void prim_collider_v(const orc_prim *P, int f, const double *r, double *out) {
    double qn[4], iq[4], rl[3], cl[3];
    normalize_n(P->rot + 4*f, 4, qn);
    iq[0] = qn[0]; iq[1] = -qn[1]; iq[2] = -qn[2]; iq[3] = -qn[3];
    qrot(iq, r, rl);
    cross3(P->w + 3*f, rl, cl);
    for (int i = 0; i < 3; i++) cl[i] += P->v[3*f+i];
    qrot(qn, cl, out);
}
static void prim_collider_v_adj(const orc_prim *P, int f, const double *r, const double *go, double *gr, prim_grad *G) {
    double qn[4], iq[4], rl[3], cl[3];
    normalize_n(P->rot + 4*f, 4, qn);
    iq[0] = qn[0]; iq[1] = -qn[1]; iq[2] = -qn[2]; iq[3] = -qn[3];
    qrot(iq, r, rl);
    cross3(P->w + 3*f, rl, cl);
    for (int i = 0; i < 3; i++) cl[i] += P->v[3*f+i];
    double gqn[4] = {0, 0, 0, 0}, gcl[3] = {0, 0, 0}, grl[3] = {0, 0, 0}, giq[4] = {0, 0, 0, 0};
    qrot_adj(qn, cl, go, gqn, gcl);
    for (int i = 0; i < 3; i++) G->v[i] += gcl[i];
    cross3_acc(rl, gcl, G->w);              /* cl = w x rl : gw += rl x gcl */
    cross3_acc(gcl, P->w + 3*f, grl);       /*              grl += gcl x w  */
    qrot_adj(iq, r, grl, giq, gr);          /* rl = qrot(iq, r) */
    gqn[0] += giq[0]; gqn[1] -= giq[1]; gqn[2] -= giq[2]; gqn[3] -= giq[3];
    normalize_adj(P->rot + 4*f, 4, gqn, G->rot);
}

static inline void pg_zero(prim_grad *G) { memset(G, 0, sizeof *G); }
static inline void wrench_add(orc_prim *P, const double *bf, const double *bt) {
    for (int i = 0; i < 3; i++) {
        #pragma omp atomic
        P->ext_f[i] += bf[i];
    }
    for (int i = 0; i < 3; i++) {
        #pragma omp atomic
        P->ext_f[i+3] += bt[i];
    }
}
static inline void prim_grad_commit(orc_prim *P, int f, const prim_grad *G) {
    for (int i = 0; i < 3; i++) {
        #pragma omp atomic
        P->gpos[3*f+i] += G->pos[i];
        #pragma omp atomic
        P->gv[3*f+i] += G->v[i];
        #pragma omp atomic
        P->gw[3*f+i] += G->w[i];
    }
    for (int i = 0; i < 4; i++) {
        #pragma omp atomic
        P->grot[4*f+i] += G->rot[i];
    }
}

/* shared friction projection used by collide and collide_mixed (primitive_base.py:86-89, 153-156):
 *   t_norm = length(t); t_fr = t / t_norm * max(0, t_norm + nc * friction)
 *   flag = (nc < 0 and sqrt(t.t) > 1e-30); t = t_fr * flag + t * (1 - flag)                      */
static void friction_proj(const double *t, double nc, double fric, double *out) {
    double tn = length_eps(t);
    double mx = fmax(0.0, tn + nc*fric);
    int flag = (nc < 0 && sqrt(dot3(t, t)) > 1e-30);
    for (int i = 0; i < 3; i++) out[i] = flag ? t[i]/tn*mx : t[i];
}
static void friction_proj_adj(const double *t, double nc, double fric, const double *go, double *gt, double *gnc) {
    double tn = length_eps(t);
    double b = tn + nc*fric;
    double mx = fmax(0.0, b);
    int flag = (nc < 0 && sqrt(dot3(t, t)) > 1e-30);
    if (!flag) { for (int i = 0; i < 3; i++) gt[i] += go[i]; return; }
    /* out = t/tn*mx ; ti.max(0, b): gradient to 0 iff b < 0 (Appendix B.2), else to b */
    double tg = dot3(t, go);
    double gmx = tg/tn;
    double gtn = -tg*mx/(tn*tn);
    for (int i = 0; i < 3; i++) gt[i] += go[i]*mx/tn;
    if (!(b < 0.0)) { gtn += gmx; *gnc += gmx*fric; }
    for (int i = 0; i < 3; i++) gt[i] += gtn*t[i]/tn;
}

/* ------------------------------------------------------------------------------------------ */
/* Primitive.collide  (grid contact)  primitive_base.py:72-103                                  */
/* ------------------------------------------------------------------------------------------ */
static void prim_collide(orc_prim *P, int f, const double *gp, const double *v_in, double dt, double gm,
                         double *v_out, int accumulate) {
    double dist = prim_sdf(P, f, gp);
    double infl = fmin(exp(-dist*P->softness), 1.0);
    for (int i = 0; i < 3; i++) v_out[i] = v_in[i];
    if ((P->softness > 0 && infl > 0.1) || dist <= 0) {
        double D[3], r[3], cv[3], iv[3], vt[3], vtf[3];
        prim_normal(P, f, gp, D);
        for (int i = 0; i < 3; i++) r[i] = gp[i] - P->pos[3*f+i];
        prim_collider_v(P, f, r, cv);
        for (int i = 0; i < 3; i++) iv[i] = v_in[i] - cv[i];
        double nc = dot3(iv, D);
        double mn = fmin(nc, 0.0);
        for (int i = 0; i < 3; i++) vt[i] = iv[i] - mn*D[i];
        friction_proj(vt, nc, P->friction, vtf);
        for (int i = 0; i < 3; i++) v_out[i] = cv[i] + iv[i]*(1-infl) + vtf[i]*infl;
        if (accumulate) {
            double bf[3], bt[3];
            for (int i = 0; i < 3; i++) bf[i] = gm*(v_in[i] - v_out[i])*(1.0/dt);
            cross3(r, bf, bt);
            wrench_add(P, bf, bt);
        }
    }
}
/* returns adjoint of v_in in gvin (accumulated), of grid_m in *ggm (accumulated) */
static void prim_collide_adj(const orc_prim *P, int f, const double *gp, const double *v_in, double dt, double gm,
                             const double *gvout, double *gvin, double *ggm, prim_grad *G) {
    double dist = prim_sdf(P, f, gp);
    double e = exp(-dist*P->softness);
    double infl = fmin(e, 1.0);
    if (!((P->softness > 0 && infl > 0.1) || dist <= 0)) { for (int i = 0; i < 3; i++) gvin[i] += gvout[i]; return; }
    double D[3], r[3], cv[3], iv[3], vt[3], vtf[3], v_out[3];
    prim_normal(P, f, gp, D);
    for (int i = 0; i < 3; i++) r[i] = gp[i] - P->pos[3*f+i];
    prim_collider_v(P, f, r, cv);
    for (int i = 0; i < 3; i++) iv[i] = v_in[i] - cv[i];
    double nc = dot3(iv, D);
    double mn = fmin(nc, 0.0);
    for (int i = 0; i < 3; i++) vt[i] = iv[i] - mn*D[i];
    friction_proj(vt, nc, P->friction, vtf);
    for (int i = 0; i < 3; i++) v_out[i] = cv[i] + iv[i]*(1-infl) + vtf[i]*infl;
    double bf[3];
    for (int i = 0; i < 3; i++) bf[i] = gm*(v_in[i] - v_out[i])*(1.0/dt);

    double g_vo[3], g_r[3] = {0,0,0}, g_bf[3], g_vi[3] = {0,0,0};
    for (int i = 0; i < 3; i++) { g_vo[i] = gvout[i]; g_bf[i] = P->ext_f_grad[i]; }
    /* b_t = r x b_f */
    cross3_acc(bf, P->ext_f_grad + 3, g_r);
    cross3_acc(P->ext_f_grad + 3, r, g_bf);
    for (int i = 0; i < 3; i++) {
        *ggm += g_bf[i]*(v_in[i] - v_out[i])*(1.0/dt);
        g_vi[i] += g_bf[i]*gm/dt;
        g_vo[i] -= g_bf[i]*gm/dt;
    }
    /* v_out = cv + iv (1-infl) + vtf infl */
    double g_cv[3], g_iv[3], g_vtf[3], g_infl = 0;
    for (int i = 0; i < 3; i++) {
        g_cv[i] = g_vo[i]; g_iv[i] = g_vo[i]*(1-infl); g_vtf[i] = g_vo[i]*infl;
        g_infl += g_vo[i]*(vtf[i] - iv[i]);
    }
    double g_vt[3] = {0,0,0}, g_nc = 0, g_D[3] = {0,0,0};
    friction_proj_adj(vt, nc, P->friction, g_vtf, g_vt, &g_nc);
    /* vt = iv - min(nc,0) D ; min(a,b): gradient to a iff a < b */
    double g_mn = 0;
    for (int i = 0; i < 3; i++) { g_iv[i] += g_vt[i]; g_mn -= g_vt[i]*D[i]; g_D[i] -= mn*g_vt[i]; }
    if (nc < 0.0) g_nc += g_mn;
    for (int i = 0; i < 3; i++) { g_iv[i] += g_nc*D[i]; g_D[i] += g_nc*iv[i]; }
    for (int i = 0; i < 3; i++) { g_vi[i] += g_iv[i]; g_cv[i] -= g_iv[i]; }
    prim_collider_v_adj(P, f, r, g_cv, g_r, G);
    for (int i = 0; i < 3; i++) G->pos[i] -= g_r[i];    /* grid_pos carries no gradient */
    prim_normal_adj(P, f, gp, g_D, NULL, G);
    double g_dist = 0;
    if (e < 1.0) g_dist += g_infl * (-P->softness) * e;
    prim_sdf_adj(P, f, gp, g_dist, NULL, G);
    for (int i = 0; i < 3; i++) gvin[i] += g_vi[i];
}

/* ------------------------------------------------------------------------------------------ */
/* Primitive.collide_particle  (penalty contact)  primitive_base.py:105-137                     */
/* ------------------------------------------------------------------------------------------ */
static void prim_collide_particle(orc_prim *P, int f, const double *x, const double *pv, double dt,
                                  double *impulse, int accumulate) {
    double dist = prim_sdf(P, f, x);
    double c = dist - 5e-3;
    impulse[0] = impulse[1] = impulse[2] = 0;
    if (c < 0.0) {
        double D[3], r[3], cv[3], iv[3], vt[3];
        prim_normal(P, f, x, D);
        for (int i = 0; i < 3; i++) r[i] = x[i] - P->pos[3*f+i];
        prim_collider_v(P, f, r, cv);
        for (int i = 0; i < 3; i++) iv[i] = pv[i] - cv[i];
        double nc = dot3(iv, D);
        for (int i = 0; i < 3; i++) vt[i] = iv[i] - nc*D[i];
        double k1 = 50.0, kf = P->friction;
        double vtn = sqrt(dot3(vt, vt) + 1e-8);
        double pf[3], bf[3], bt[3];
        for (int i = 0; i < 3; i++) {
            double f1 = -D[i]*c*k1, f2 = -vt[i]/vtn*fabs(nc)*kf;
            pf[i] = (f1 + f2)*1.0; bf[i] = -(f1 + f2)*1.0;
        }
        if (accumulate) { cross3(r, bf, bt); wrench_add(P, bf, bt); }
        for (int i = 0; i < 3; i++) impulse[i] = pf[i]*dt;
    }
}
static void prim_collide_particle_adj(const orc_prim *P, int f, const double *x, const double *pv, double dt,
                                      const double *gimp, double *gx, double *gpv, prim_grad *G) {
    double dist = prim_sdf(P, f, x);
    double c = dist - 5e-3;
    if (!(c < 0.0)) return;
    double D[3], r[3], cv[3], iv[3], vt[3];
    prim_normal(P, f, x, D);
    for (int i = 0; i < 3; i++) r[i] = x[i] - P->pos[3*f+i];
    prim_collider_v(P, f, r, cv);
    for (int i = 0; i < 3; i++) iv[i] = pv[i] - cv[i];
    double nc = dot3(iv, D);
    for (int i = 0; i < 3; i++) vt[i] = iv[i] - nc*D[i];
    double k1 = 50.0, kf = P->friction;
    double vtn = sqrt(dot3(vt, vt) + 1e-8);
    double bf[3];
    for (int i = 0; i < 3; i++) bf[i] = -(-D[i]*c*k1 - vt[i]/vtn*fabs(nc)*kf);
    /* adjoint of p_f (via impulse) and of b_f, b_t (via ext_f.grad) */
    double g_pf[3], g_bf[3], g_r[3] = {0,0,0};
    for (int i = 0; i < 3; i++) { g_pf[i] = gimp[i]*dt; g_bf[i] = P->ext_f_grad[i]; }
    cross3_acc(bf, P->ext_f_grad + 3, g_r);
    cross3_acc(P->ext_f_grad + 3, r, g_bf);
    double g_f[3];                                  /* adjoint of (f1+f2) */
    for (int i = 0; i < 3; i++) g_f[i] = g_pf[i] - g_bf[i];
    /* f1 = -D c k1 ; f2 = -vt/vtn*|nc|*kf */
    double g_D[3], g_c = 0, g_vt[3], g_vtn = 0, g_abs = 0;
    for (int i = 0; i < 3; i++) {
        g_D[i] = -g_f[i]*c*k1; g_c -= g_f[i]*D[i]*k1;
        g_vt[i] = -g_f[i]/vtn*fabs(nc)*kf;
        g_vtn += g_f[i]*vt[i]/(vtn*vtn)*fabs(nc)*kf;
        g_abs -= g_f[i]*vt[i]/vtn*kf;
    }
    for (int i = 0; i < 3; i++) g_vt[i] += g_vtn*vt[i]/vtn;
    double g_nc = g_abs * (nc > 0 ? 1.0 : (nc < 0 ? -1.0 : 0.0));
    double g_iv[3];
    for (int i = 0; i < 3; i++) { g_iv[i] = g_vt[i]; g_nc -= g_vt[i]*D[i]; g_D[i] -= nc*g_vt[i]; }
    for (int i = 0; i < 3; i++) { g_iv[i] += g_nc*D[i]; g_D[i] += g_nc*iv[i]; }
    double g_cv[3];
    for (int i = 0; i < 3; i++) { gpv[i] += g_iv[i]; g_cv[i] = -g_iv[i]; }
    prim_collider_v_adj(P, f, r, g_cv, g_r, G);
    for (int i = 0; i < 3; i++) { gx[i] += g_r[i]; G->pos[i] -= g_r[i]; }
    prim_normal_adj(P, f, x, g_D, gx, G);
    prim_sdf_adj(P, f, x, g_c, gx, G);
}

/* ------------------------------------------------------------------------------------------ */
/* Primitive.collide_mixed  (forecast-based contact)  primitive_base.py:139-181                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {    /* forward intermediates kept for the adjoint */
    int active, moving_in, outside, pen;
    double dist, D[3], r[3], cv[3], iv[3], nc, vt0[3], vt[3], infl, e, pv_mid[3], x_new[3], s, n[3];
} cm_tape;

static void prim_collide_mixed_fwd(orc_prim *P, int f, const double *x, const double *pv_in, double p_mass,
                                   double dt, double life, double *pv_out, cm_tape *T, int accumulate) {
    T->dist = prim_sdf(P, f, x);
    T->active = (T->dist <= 5e-3);
    for (int i = 0; i < 3; i++) pv_out[i] = pv_in[i];
    if (!T->active) return;
    double pv[3] = {pv_in[0], pv_in[1], pv_in[2]};
    prim_normal(P, f, x, T->D);
    for (int i = 0; i < 3; i++) T->r[i] = x[i] - P->pos[3*f+i];
    prim_collider_v(P, f, T->r, T->cv);
    for (int i = 0; i < 3; i++) T->iv[i] = pv[i] - T->cv[i];
    T->nc = dot3(T->iv, T->D);
    T->moving_in = (T->nc < 0);
    T->outside = 0;
    if (T->moving_in) {
        for (int i = 0; i < 3; i++) T->vt0[i] = T->iv[i] - T->nc*T->D[i];
        friction_proj(T->vt0, T->nc, P->friction, T->vt);
        for (int i = 0; i < 3; i++) pv[i] = T->cv[i] + T->vt[i];
        if (T->dist > 0) {
            T->outside = 1;
            T->e = exp(-T->dist*P->softness);
            T->infl = fmin(T->e, 1.0);
            for (int i = 0; i < 3; i++) pv[i] = T->cv[i] + T->iv[i]*(1-T->infl) + T->vt[i]*T->infl;
        }
    }
    for (int i = 0; i < 3; i++) { T->pv_mid[i] = pv[i]; T->x_new[i] = pv[i]*dt + x[i]; }
    T->s = prim_sdf(P, f, T->x_new);
    T->pen = (T->s < 0);
    if (T->pen) {
        prim_normal(P, f, T->x_new, T->n);
        for (int i = 0; i < 3; i++) pv[i] = pv[i] - (T->s/dt)*T->n[i]*life;
    }
    if (accumulate) {
        double bf[3], bt[3];
        for (int i = 0; i < 3; i++) bf[i] = p_mass*(pv_in[i] - pv[i])*(1.0/dt);
        cross3(T->r, bf, bt);
        wrench_add(P, bf, bt);
    }
    for (int i = 0; i < 3; i++) pv_out[i] = pv[i];
}
static void prim_collide_mixed_adj(const orc_prim *P, int f, const double *x, const double *pv_in, const double *pv_out,
                                   double p_mass, double dt, double life, const cm_tape *T,
                                   const double *gout, double *gx, double *gpv_in, prim_grad *G) {
    if (!T->active) { for (int i = 0; i < 3; i++) gpv_in[i] += gout[i]; return; }
    double g_pv[3], g_bf[3], g_r[3] = {0,0,0}, g_vin[3], bf[3];
    for (int i = 0; i < 3; i++) { g_pv[i] = gout[i]; g_bf[i] = P->ext_f_grad[i]; bf[i] = p_mass*(pv_in[i] - pv_out[i])*(1.0/dt); }
    cross3_acc(bf, P->ext_f_grad + 3, g_r);
    cross3_acc(P->ext_f_grad + 3, T->r, g_bf);
    for (int i = 0; i < 3; i++) { g_vin[i] = g_bf[i]*p_mass/dt; g_pv[i] -= g_bf[i]*p_mass/dt; }
    double g_xnew[3] = {0,0,0}, g_s = 0;
    if (T->pen) {
        double g_n[3];
        for (int i = 0; i < 3; i++) { g_s -= g_pv[i]*T->n[i]*life/dt; g_n[i] = -(T->s/dt)*life*g_pv[i]; }
        prim_normal_adj(P, f, T->x_new, g_n, g_xnew, G);
    }
    prim_sdf_adj(P, f, T->x_new, g_s, g_xnew, G);
    /* x_new = pv_mid*dt + x */
    double g_mid[3];
    for (int i = 0; i < 3; i++) { g_mid[i] = g_pv[i] + dt*g_xnew[i]; gx[i] += g_xnew[i]; }
    double g_cv[3] = {0,0,0}, g_iv[3] = {0,0,0}, g_D[3] = {0,0,0}, g_nc = 0, g_dist = 0;
    if (T->moving_in) {
        double g_vt[3];
        if (T->outside) {
            double g_infl = 0;
            for (int i = 0; i < 3; i++) {
                g_cv[i] += g_mid[i]; g_iv[i] += g_mid[i]*(1-T->infl); g_vt[i] = g_mid[i]*T->infl;
                g_infl += g_mid[i]*(T->vt[i] - T->iv[i]);
            }
            if (T->e < 1.0) g_dist += g_infl*(-P->softness)*T->e;
        } else {
            for (int i = 0; i < 3; i++) { g_cv[i] += g_mid[i]; g_vt[i] = g_mid[i]; }
        }
        double g_vt0[3] = {0,0,0};
        friction_proj_adj(T->vt0, T->nc, P->friction, g_vt, g_vt0, &g_nc);
        for (int i = 0; i < 3; i++) { g_iv[i] += g_vt0[i]; g_nc -= g_vt0[i]*T->D[i]; g_D[i] -= T->nc*g_vt0[i]; }
    } else {
        for (int i = 0; i < 3; i++) g_vin[i] += g_mid[i];       /* p_v unchanged = p_v_in */
    }
    for (int i = 0; i < 3; i++) { g_iv[i] += g_nc*T->D[i]; g_D[i] += g_nc*T->iv[i]; }
    for (int i = 0; i < 3; i++) { g_vin[i] += g_iv[i]; g_cv[i] -= g_iv[i]; }
    prim_collider_v_adj(P, f, T->r, g_cv, g_r, G);
    for (int i = 0; i < 3; i++) { gx[i] += g_r[i]; G->pos[i] -= g_r[i]; }
    prim_normal_adj(P, f, x, g_D, gx, G);
    prim_sdf_adj(P, f, x, g_dist, gx, G);
    for (int i = 0; i < 3; i++) gpv_in[i] += g_vin[i];
}

/* ------------------------------------------------------------------------------------------ */
/* simulator object                                                                            */
/* ------------------------------------------------------------------------------------------ */
static double *zalloc(size_t n) { double *p = (double *)calloc(n ? n : 1, sizeof(double)); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }

/* cfg -> constants, mpm_simulator.py:17-52 */
orc_sim *orc_create(int n_particles, int n_grid, int max_steps, double dt, double E, double nu,
                    const double *gravity, double ground_friction, int material_model, int ptype,
                    int collision_type, int substeps, int n_control, int rigid_velocity_control) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    s->n = n_particles; s->ng = n_grid; s->T = max_steps;
    s->dt = dt; s->dx = 1.0 / n_grid; s->inv_dx = (double)n_grid;
    s->p_vol = (s->dx * 0.5) * (s->dx * 0.5);       /* squared even in 3-D, :34 */
    s->p_mass = s->p_vol * 1;
    s->mu = E / (2 * (1 + nu)); s->lam = E * nu / ((1 + nu) * (1 - 2 * nu));
    if (ptype == 1) { s->mu *= 0.3; s->lam *= 0.3; } else if (ptype == 2) { s->mu = 0.0; }
    for (int i = 0; i < 3; i++) s->gravity[i] = gravity[i];
    s->sticky_ground = ground_friction >= 10.0;
    s->material_model = material_model; s->ptype = ptype; s->collision_type = collision_type;
    s->substeps = substeps; s->n_control = n_control; s->rigid_velocity_control = rigid_velocity_control;
    size_t n = n_particles, T = max_steps, G = (size_t)n_grid*n_grid*n_grid;
    s->x = zalloc(T*n*3); s->v = zalloc(T*n*3); s->C = zalloc(T*n*9); s->F = zalloc(T*n*9);
    s->gx = zalloc(T*n*3); s->gv = zalloc(T*n*3); s->gC = zalloc(T*n*9); s->gF = zalloc(T*n*9);
    s->Ftmp = zalloc(n*9); s->U = zalloc(n*9); s->S = zalloc(n*9); s->V = zalloc(n*9);
    s->gFtmp = zalloc(n*9); s->gU = zalloc(n*9); s->gS = zalloc(n*9); s->gV = zalloc(n*9);
    s->gvin = zalloc(G*3); s->gvout = zalloc(G*3); s->gm = zalloc(G); s->gvmix = zalloc(G*3);
    s->ggvin = zalloc(G*3); s->ggvout = zalloc(G*3); s->ggm = zalloc(G); s->ggvmix = zalloc(G*3);
    s->vtmp = zalloc(n*3); s->vtgt = zalloc(n*3); s->gvtmp = zalloc(n*3); s->gvtgt = zalloc(n*3);
    s->control_idx = (int *)calloc(n ? n : 1, sizeof(int));
    s->action = zalloc((size_t)(n_control > 0 ? n_control : 1)*3);
    s->gaction = zalloc((size_t)(n_control > 0 ? n_control : 1)*3);
    return s;
}
/* soft_cloth variant of the plastic material: von Mises return mapping with cfg.yield_stress instead of the sigma clip */
void orc_set_plasticity(orc_sim *s, int mode, double yield_stress) { s->plasticity = mode; s->yield_stress = yield_stress; }
void orc_destroy(orc_sim *s) {
    if (!s) return;
    double *ps[] = {s->x, s->v, s->C, s->F, s->gx, s->gv, s->gC, s->gF, s->Ftmp, s->U, s->S, s->V, s->gFtmp, s->gU, s->gS, s->gV,
                    s->gvin, s->gvout, s->gm, s->gvmix, s->ggvin, s->ggvout, s->ggm, s->ggvmix, s->vtmp, s->vtgt, s->gvtmp, s->gvtgt,
                    s->action, s->gaction};
    for (size_t i = 0; i < sizeof ps / sizeof *ps; i++) free(ps[i]);
    free(s->control_idx);
    for (int i = 0; i < s->np; i++) {
        orc_prim *P = &s->prim[i];
        double *pp[] = {P->sdf, P->nrm, P->pos, P->rot, P->v, P->w, P->gpos, P->grot, P->gv, P->gw, P->abuf, P->gabuf};
        for (size_t k = 0; k < sizeof pp / sizeof *pp; k++) free(pp[k]);
    }
    free(s);
}
/* Mesh primitive: tables + per-frame state series (primitive_base.py:26-43, mesh.py:26-43) */
int orc_add_primitive(orc_sim *s, const double *sdf, const double *nrm, const int *res, const double *lower,
                      const double *upper, double sdf_dx, double friction, double softness, int enabled) {
    if (s->np >= ORC_MAXP) return -1;
    orc_prim *P = &s->prim[s->np];
    memset(P, 0, sizeof *P);
    P->friction = friction; P->softness = softness; P->enabled = enabled;
    size_t T = s->T;
    if (sdf) {
        P->has_table = 1;
        size_t R = (size_t)res[0]*res[1]*res[2];
        P->sdf = zalloc(R); P->nrm = zalloc(R*3);
        memcpy(P->sdf, sdf, R*sizeof(double)); memcpy(P->nrm, nrm, R*3*sizeof(double));
        for (int i = 0; i < 3; i++) { P->res[i] = res[i]; P->lower[i] = lower[i]; P->upper[i] = upper[i]; }
        P->sdf_dx = sdf_dx; P->inv_sdf_dx = 1.0 / sdf_dx;
    }
    P->pos = zalloc(T*3); P->rot = zalloc(T*4); P->v = zalloc(T*3); P->w = zalloc(T*3);
    P->gpos = zalloc(T*3); P->grot = zalloc(T*4); P->gv = zalloc(T*3); P->gw = zalloc(T*3);
    P->abuf = zalloc(T*6); P->gabuf = zalloc(T*6);
    return s->np++;
}
void orc_set_primitive_enabled(orc_sim *s, int i, int enabled) { s->prim[i].enabled = enabled; }
void orc_set_primitive_params(orc_sim *s, int i, double friction, double softness) { s->prim[i].friction = friction; s->prim[i].softness = softness; }

/* state IO: (n,24) row = [x(3) v(3) F(9 row-major) C(9 row-major)], mpm_simulator.py:481-512 */
void orc_set_frame(orc_sim *s, int f, const double *st24) {
    size_t n = s->n;
    for (size_t p = 0; p < n; p++) {
        const double *r = st24 + 24*p;
        memcpy(s->x + (f*n + p)*3, r, 3*sizeof(double));
        memcpy(s->v + (f*n + p)*3, r + 3, 3*sizeof(double));
        memcpy(s->F + (f*n + p)*9, r + 6, 9*sizeof(double));
        memcpy(s->C + (f*n + p)*9, r + 15, 9*sizeof(double));
    }
}
void orc_get_frame(const orc_sim *s, int f, double *st24) {
    size_t n = s->n;
    for (size_t p = 0; p < n; p++) {
        double *r = st24 + 24*p;
        memcpy(r, s->x + (f*n + p)*3, 3*sizeof(double));
        memcpy(r + 3, s->v + (f*n + p)*3, 3*sizeof(double));
        memcpy(r + 6, s->F + (f*n + p)*9, 9*sizeof(double));
        memcpy(r + 15, s->C + (f*n + p)*9, 9*sizeof(double));
    }
}
void orc_add_frame_grad(orc_sim *s, int f, const double *g24) {
    size_t n = s->n;
    for (size_t p = 0; p < n; p++) {
        const double *r = g24 + 24*p;
        for (int i = 0; i < 3; i++) { s->gx[(f*n + p)*3 + i] += r[i]; s->gv[(f*n + p)*3 + i] += r[3+i]; }
        for (int i = 0; i < 9; i++) { s->gF[(f*n + p)*9 + i] += r[6+i]; s->gC[(f*n + p)*9 + i] += r[15+i]; }
    }
}
void orc_get_frame_grad(const orc_sim *s, int f, double *g24) {
    size_t n = s->n;
    for (size_t p = 0; p < n; p++) {
        double *r = g24 + 24*p;
        memcpy(r, s->gx + (f*n + p)*3, 3*sizeof(double));
        memcpy(r + 3, s->gv + (f*n + p)*3, 3*sizeof(double));
        memcpy(r + 6, s->gF + (f*n + p)*9, 9*sizeof(double));
        memcpy(r + 15, s->gC + (f*n + p)*9, 9*sizeof(double));
    }
}
/* ti.ad.clear_all_gradients() (demo_grip.py:135) restricted to the fields of this path */
void orc_clear_grads(orc_sim *s) {
    size_t n = s->n, T = s->T;
    memset(s->gx, 0, T*n*3*sizeof(double)); memset(s->gv, 0, T*n*3*sizeof(double));
    memset(s->gC, 0, T*n*9*sizeof(double)); memset(s->gF, 0, T*n*9*sizeof(double));
    memset(s->gaction, 0, (size_t)(s->n_control > 0 ? s->n_control : 1)*3*sizeof(double));
    for (int i = 0; i < s->np; i++) {
        orc_prim *P = &s->prim[i];
        memset(P->gpos, 0, T*3*sizeof(double)); memset(P->grot, 0, T*4*sizeof(double));
        memset(P->gv, 0, T*3*sizeof(double)); memset(P->gw, 0, T*3*sizeof(double));
        memset(P->gabuf, 0, T*6*sizeof(double));
        memset(P->ext_f_grad, 0, sizeof P->ext_f_grad);
    }
}
/* Primitive.set_all_states / get_all_states_grad, primitive_base.py:258-265: [x(3) q(4) v(3) w(3)] */
void orc_set_primitive_state(orc_sim *s, int i, int f, const double *s13) {
    orc_prim *P = &s->prim[i];
    memcpy(P->pos + 3*f, s13, 3*sizeof(double)); memcpy(P->rot + 4*f, s13 + 3, 4*sizeof(double));
    memcpy(P->v + 3*f, s13 + 7, 3*sizeof(double)); memcpy(P->w + 3*f, s13 + 10, 3*sizeof(double));
}
void orc_get_primitive_state(const orc_sim *s, int i, int f, double *s13) {
    const orc_prim *P = &s->prim[i];
    memcpy(s13, P->pos + 3*f, 3*sizeof(double)); memcpy(s13 + 3, P->rot + 4*f, 4*sizeof(double));
    memcpy(s13 + 7, P->v + 3*f, 3*sizeof(double)); memcpy(s13 + 10, P->w + 3*f, 3*sizeof(double));
}
void orc_get_primitive_state_grad(const orc_sim *s, int i, int f, double *g13) {
    const orc_prim *P = &s->prim[i];
    memcpy(g13, P->gpos + 3*f, 3*sizeof(double)); memcpy(g13 + 3, P->grot + 4*f, 4*sizeof(double));
    memcpy(g13 + 7, P->gv + 3*f, 3*sizeof(double)); memcpy(g13 + 10, P->gw + 3*f, 3*sizeof(double));
}
void orc_add_primitive_state_grad(orc_sim *s, int i, int f, const double *g13) {
    orc_prim *P = &s->prim[i];
    for (int k = 0; k < 3; k++) { P->gpos[3*f+k] += g13[k]; P->gv[3*f+k] += g13[7+k]; P->gw[3*f+k] += g13[10+k]; }
    for (int k = 0; k < 4; k++) P->grot[4*f+k] += g13[3+k];
}
void orc_get_ext_f(const orc_sim *s, int i, double *o6) { memcpy(o6, s->prim[i].ext_f, 6*sizeof(double)); }
/* clear_ext_f zeroes value and grad, primitive_base.py:183-187 */
void orc_clear_ext_f(orc_sim *s, int i) { memset(s->prim[i].ext_f, 0, 48); memset(s->prim[i].ext_f_grad, 0, 48); }
void orc_set_ext_f_grad(orc_sim *s, int i, const double *g6) { memcpy(s->prim[i].ext_f_grad, g6, 48); }
/* set_action_kernel also zeroes action.grad, mpm_simulator.py:579-586 */
void orc_set_action(orc_sim *s, const double *a) {
    memcpy(s->action, a, (size_t)s->n_control*3*sizeof(double));
    memset(s->gaction, 0, (size_t)s->n_control*3*sizeof(double));
}
void orc_get_action_grad(const orc_sim *s, double *g) { memcpy(g, s->gaction, (size_t)s->n_control*3*sizeof(double)); }
void orc_set_control_idx(orc_sim *s, const int *idx) { memcpy(s->control_idx, idx, (size_t)s->n*sizeof(int)); }
/* velocity control: set_action / set_velocity_from_action_kernel (+.grad), primitive_base.py:285-319 */
void orc_set_primitive_action(orc_sim *s, int i, int st, int n, const double *a6) {
    orc_prim *P = &s->prim[i];
    memcpy(P->abuf + 6*st, a6, 48);
    for (int j = st*n; j < (st+1)*n; j++)
        for (int k = 0; k < 3; k++) { P->v[3*j+k] = P->abuf[6*st+k+3]; P->w[3*j+k] = P->abuf[6*st+k]; }
}
void orc_get_primitive_action_grad(orc_sim *s, int i, int st, int n, double *g6) {
    orc_prim *P = &s->prim[i];
    for (int j = st*n; j < (st+1)*n; j++)
        for (int k = 0; k < 3; k++) { P->gabuf[6*st+k+3] += P->gv[3*j+k]; P->gabuf[6*st+k] += P->gw[3*j+k]; }
    memcpy(g6, P->gabuf + 6*st, 48);
}

/* ------------------------------------------------------------------------------------------ */
/* kernels                                                                                     */
/* ------------------------------------------------------------------------------------------ */
#define GIDX(s, i, j, k) (((size_t)(i)*(s)->ng + (j))*(s)->ng + (k))

static void clear_grid(orc_sim *s) {                                            /* :93-114 */
    size_t G = (size_t)s->ng*s->ng*s->ng, n = s->n;
    #pragma omp parallel for
    for (size_t I = 0; I < G; I++) {
        for (int d = 0; d < 3; d++) {
            s->gvin[3*I+d] = s->gvout[3*I+d] = s->gvmix[3*I+d] = 0;
            s->ggvin[3*I+d] = s->ggvout[3*I+d] = s->ggvmix[3*I+d] = 0;
        }
        s->gm[I] = 0; s->ggm[I] = 0;
    }
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++)
        for (int d = 0; d < 3; d++) s->vtmp[3*p+d] = s->vtgt[3*p+d] = s->gvtmp[3*p+d] = s->gvtgt[3*p+d] = 0;
}
static void clear_SVD_grad(orc_sim *s) {                                        /* :116-123 */
    size_t n = s->n;
    memset(s->gU, 0, n*72); memset(s->gS, 0, n*72); memset(s->gV, 0, n*72); memset(s->gFtmp, 0, n*72);
}
static void compute_F_tmp(orc_sim *s, int f) {                                  /* :125-128 */
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        double A[9];
        const double *C = s->C + (f*n + p)*9;
        for (int i = 0; i < 9; i++) A[i] = (i % 4 == 0) + s->dt*C[i];
        mm(A, s->F + (f*n + p)*9, s->Ftmp + 9*p);
    }
}
static void compute_F_tmp_grad(orc_sim *s, int f) {
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        double A[9], t[9];
        const double *C = s->C + (f*n + p)*9, *g = s->gFtmp + 9*p;
        for (int i = 0; i < 9; i++) A[i] = (i % 4 == 0) + s->dt*C[i];
        mmT(g, s->F + (f*n + p)*9, t);                  /* dC += dt * g F^T */
        for (int i = 0; i < 9; i++) s->gC[(f*n + p)*9 + i] += s->dt*t[i];
        mTm(A, g, t);                                   /* dF += A^T g */
        for (int i = 0; i < 9; i++) s->gF[(f*n + p)*9 + i] += t[i];
    }
}
static void svd_kernel(orc_sim *s) {                                            /* :130-133 */
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) svd3(s->Ftmp + 9*p, s->U + 9*p, s->S + 9*p, s->V + 9*p);
}
static void svd_grad(orc_sim *s) {                                              /* :135-138 */
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        double o[9];
        backward_svd(s->gU + 9*p, s->gS + 9*p, s->gV + 9*p, s->U + 9*p, s->S + 9*p, s->V + 9*p, o);
        for (int i = 0; i < 9; i++) s->gFtmp[9*p+i] += o[i];
    }
}

/* quadratic B-spline stencil, :215-217 */
static inline void stencil(const orc_sim *s, const double *x, int *base, double *fx, double w[3][3]) {
    for (int d = 0; d < 3; d++) {
        base[d] = (int)(x[d]*s->inv_dx - 0.5);
        /* The reference has no bounds check here (out-of-domain particles are undefined behaviour, SURVEY.md
           section 5); the restatement clamps the base cell exactly like the CUDA path does, so that a scene
           that blows up cannot corrupt the test process. In-domain particles are unaffected. */
        if (!(base[d] >= 0)) base[d] = 0;
        if (base[d] > s->ng - 3) base[d] = s->ng - 3;
        fx[d] = x[d]*s->inv_dx - (double)base[d];
        w[0][d] = 0.5*(1.5 - fx[d])*(1.5 - fx[d]);
        w[1][d] = 0.75 - (fx[d] - 1.0)*(fx[d] - 1.0);
        w[2][d] = 0.5*(fx[d] - 0.5)*(fx[d] - 0.5);
    }
}
static inline void stencil_dw(const double *fx, double dw[3][3]) {
    for (int d = 0; d < 3; d++) { dw[0][d] = fx[d] - 1.5; dw[1][d] = -2.0*(fx[d] - 1.0); dw[2][d] = fx[d] - 0.5; }
}

/* compute_von_mises, soft_cloth/engine/mpm_simulator.py:172-189 (the softmac copy, mpm_simulator.py:166-182, lacks the 0.05 floor and is
 * commented out at its call site :225).  S: diagonal of sig.  Returns 1 when the particle yields; then sn = exp(epsilon') and, if D is not
 * NULL, D[k][i] = d sn_k / d S_i with Taichi's conventions: ti.max(sig, 0.05) passes its gradient to sig iff 0.05 < sig, the condition
 * delta_gamma > 0 carries none, norm(x) = sqrt(x.x + 1e-8) (:201-202). */
static int von_mises_sig(const orc_sim *s, const double *S, double sn[3], double D[3][3]) {
    double sc[3], eps[3], eh[3], m = 0, q = 1e-8;
    for (int d = 0; d < 3; d++) { sc[d] = fmax(S[4*d], 0.05); eps[d] = log(sc[d]); m += eps[d]/3; }
    for (int d = 0; d < 3; d++) { eh[d] = eps[d] - m; q += eh[d]*eh[d]; }
    double nrm = sqrt(q), c = s->yield_stress/(2*s->mu);
    if (!(nrm - c > 0)) return 0;
    for (int d = 0; d < 3; d++) sn[d] = exp(eps[d] - ((nrm - c)/nrm)*eh[d]);
    if (D) for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) {
        /* eps'_k = eps_k - (1 - c/n) eh_k;  d eh_k/d eps_i = [k==i] - 1/3;  d n/d eps_i = eh_i/n (sum eh = 0) */
        double de = (k == i) - (1 - c/nrm)*((k == i) - 1.0/3.0) - (c/(nrm*nrm*nrm))*eh[i]*eh[k];
        D[k][i] = (0.05 < S[4*i]) ? sn[k]*de/sc[i] : 0.0;
    }
    return 1;
}

/* material update: F_tmp,U,S,V -> new_F, stress (before the -dt*p_vol*4*inv_dx^2 prefactor), :219-245 */
static void material_fwd(const orc_sim *s, const double *Ftmp, const double *U, const double *S, const double *V,
                         double *newF, double *stress, double *J_out) {
    double J = det3(Ftmp);
    memcpy(newF, Ftmp, 72);
    if (s->material_model == 0) {
        if (s->ptype == 0 && s->plasticity == 1) {
            double sn[3];
            if (von_mises_sig(s, S, sn, NULL)) {                /* otherwise new_F stays F_tmp */
                double Sn[9] = {0}, t[9];
                for (int d = 0; d < 3; d++) Sn[4*d] = sn[d];
                mm(U, Sn, t); mmT(t, V, newF);
            }
        } else if (s->ptype == 0) {
            double Sn[9] = {0}, t[9];
            for (int d = 0; d < 3; d++) Sn[4*d] = fmin(fmax(S[4*d], 1 - 2e-3), 1 + 3e-3);
            mm(U, Sn, t); mmT(t, V, newF);
        } else if (s->ptype == 2) {
            double c = pow(J, 1.0/3.0);
            for (int i = 0; i < 9; i++) newF[i] = (i % 4 == 0) ? c : 0.0;
        }
        double r[9], A[9];
        mmT(U, V, r);
        for (int i = 0; i < 9; i++) A[i] = newF[i] - r[i];
        mmT(A, newF, stress);
        for (int i = 0; i < 9; i++) stress[i] = 2*s->mu*stress[i] + ((i % 4 == 0) ? s->lam*J*(J - 1) : 0.0);
    } else {
        if (s->ptype == 2) {
            double sq = sqrt(J);
            memset(newF, 0, 72); newF[0] = sq; newF[4] = sq; newF[8] = 1;
        }
        mmT(newF, newF, stress);
        for (int i = 0; i < 9; i++) stress[i] = s->mu*stress[i] + ((i % 4 == 0) ? (s->lam*log(J) - s->mu) : 0.0);
    }
    *J_out = J;
}

static void p2g(orc_sim *s, int f, int accumulate_wrench) {                     /* :198-262 */
    size_t n = s->n;
    const double cs = -s->dt*s->p_vol*4*s->inv_dx*s->inv_dx;
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; p++) {
        const double *x = s->x + (f*n + p)*3, *v = s->v + (f*n + p)*3, *C = s->C + (f*n + p)*9;
        double imp[3] = {0, 0, 0};
        if (s->collision_type == 1)
            for (int i = 0; i < s->np; i++) if (s->prim[i].enabled) {
                double t[3]; prim_collide_particle(&s->prim[i], f, x, v, s->dt, t, accumulate_wrench);
                for (int d = 0; d < 3; d++) imp[d] += t[d];
            }
        if (s->n_control > 0) {
            int ci = s->control_idx[p];
            if (ci >= 0) for (int d = 0; d < 3; d++) imp[d] += 6e-4*s->action[3*ci+d]*s->dt;
        }
        int base[3]; double fx[3], w[3][3];
        stencil(s, x, base, fx, w);
        double newF[9], stress[9], J, affine[9];
        material_fwd(s, s->Ftmp + 9*p, s->U + 9*p, s->S + 9*p, s->V + 9*p, newF, stress, &J);
        for (int i = 0; i < 9; i++) affine[i] = cs*stress[i] + s->p_mass*C[i];
        memcpy(s->F + ((f+1)*n + p)*9, newF, 72);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double dpos[3] = {(i - fx[0])*s->dx, (j - fx[1])*s->dx, (k - fx[2])*s->dx};
            double wt = w[i][0]*w[j][1]*w[k][2];
            size_t I = GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            for (int d = 0; d < 3; d++) {
                double val = wt*(s->p_mass*v[d] + affine[3*d]*dpos[0] + affine[3*d+1]*dpos[1] + affine[3*d+2]*dpos[2] + imp[d]);
                #pragma omp atomic
                s->gvin[3*I+d] += val;
            }
            #pragma omp atomic
            s->gm[I] += wt*s->p_mass;
        }
    }
}

static void p2g_grad(orc_sim *s, int f) {
    size_t n = s->n;
    const double cs = -s->dt*s->p_vol*4*s->inv_dx*s->inv_dx;
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; p++) {
        const double *x = s->x + (f*n + p)*3, *v = s->v + (f*n + p)*3, *C = s->C + (f*n + p)*9;
        const double *Ftmp = s->Ftmp + 9*p, *U = s->U + 9*p, *S = s->S + 9*p, *V = s->V + 9*p;
        double imp[3] = {0, 0, 0};
        if (s->collision_type == 1)
            for (int i = 0; i < s->np; i++) if (s->prim[i].enabled) {
                double t[3]; prim_collide_particle(&s->prim[i], f, x, v, s->dt, t, 0);
                for (int d = 0; d < 3; d++) imp[d] += t[d];
            }
        int ci = -1;
        if (s->n_control > 0) {
            ci = s->control_idx[p];
            if (ci >= 0) for (int d = 0; d < 3; d++) imp[d] += 6e-4*s->action[3*ci+d]*s->dt;
        }
        int base[3]; double fx[3], w[3][3], dw[3][3];
        stencil(s, x, base, fx, w); stencil_dw(fx, dw);
        double newF[9], stress[9], J, affine[9];
        material_fwd(s, Ftmp, U, S, V, newF, stress, &J);
        for (int i = 0; i < 9; i++) affine[i] = cs*stress[i] + s->p_mass*C[i];

        double g_aff[9] = {0}, g_v[3] = {0,0,0}, g_imp[3] = {0,0,0}, g_fx[3] = {0,0,0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double dpos[3] = {(i - fx[0])*s->dx, (j - fx[1])*s->dx, (k - fx[2])*s->dx};
            double wt = w[i][0]*w[j][1]*w[k][2];
            size_t I = GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            const double *gg = s->ggvin + 3*I;
            double val[3], g_wt = s->ggm[I]*s->p_mass, g_dpos[3] = {0,0,0};
            for (int d = 0; d < 3; d++) {
                val[d] = s->p_mass*v[d] + affine[3*d]*dpos[0] + affine[3*d+1]*dpos[1] + affine[3*d+2]*dpos[2] + imp[d];
                g_wt += gg[d]*val[d];
                double gval = wt*gg[d];
                g_v[d] += s->p_mass*gval; g_imp[d] += gval;
                for (int e = 0; e < 3; e++) { g_aff[3*d+e] += gval*dpos[e]; g_dpos[e] += affine[3*d+e]*gval; }
            }
            for (int e = 0; e < 3; e++) g_fx[e] -= s->dx*g_dpos[e];
            g_fx[0] += g_wt*dw[i][0]*w[j][1]*w[k][2];
            g_fx[1] += g_wt*w[i][0]*dw[j][1]*w[k][2];
            g_fx[2] += g_wt*w[i][0]*w[j][1]*dw[k][2];
        }
        double *gx = s->gx + (f*n + p)*3, *gv = s->gv + (f*n + p)*3, *gC = s->gC + (f*n + p)*9;
        for (int d = 0; d < 3; d++) { gx[d] += s->inv_dx*g_fx[d]; gv[d] += g_v[d]; }
        for (int i = 0; i < 9; i++) gC[i] += s->p_mass*g_aff[i];
        double g_stress[9], g_newF[9], g_J = 0;
        for (int i = 0; i < 9; i++) { g_stress[i] = cs*g_aff[i]; g_newF[i] = s->gF[((f+1)*n + p)*9 + i]; }
        double tr = g_stress[0] + g_stress[4] + g_stress[8];
        double *gFtmp = s->gFtmp + 9*p, *gU = s->gU + 9*p, *gS = s->gS + 9*p, *gV = s->gV + 9*p;
        if (s->material_model == 0) {
            double r[9], A[9], t[9], g_A[9];
            mmT(U, V, r);
            for (int i = 0; i < 9; i++) A[i] = newF[i] - r[i];
            mm(g_stress, newF, g_A);                                    /* stress = 2mu A newF^T */
            mTm(g_stress, A, t);
            for (int i = 0; i < 9; i++) { g_A[i] *= 2*s->mu; g_newF[i] += 2*s->mu*t[i] + g_A[i]; }
            g_J += s->lam*(2*J - 1)*tr;
            /* r = U V^T */
            mm(g_A, V, t); for (int i = 0; i < 9; i++) gU[i] -= t[i];
            mTm(g_A, U, t); for (int i = 0; i < 9; i++) gV[i] -= t[i];
            if (s->ptype == 0 && s->plasticity == 1) {
                double sn[3], D[3][3];
                if (von_mises_sig(s, S, sn, D)) {
                    double Sn[9] = {0}, t2[9];
                    for (int d = 0; d < 3; d++) Sn[4*d] = sn[d];
                    mm(g_newF, V, t); mm(t, Sn, t2); for (int i = 0; i < 9; i++) gU[i] += t2[i];
                    mTm(g_newF, U, t); mm(t, Sn, t2); for (int i = 0; i < 9; i++) gV[i] += t2[i];
                    mTm(U, g_newF, t); mm(t, V, t2);
                    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) gS[4*i] += t2[4*k]*D[k][i];
                } else {
                    for (int i = 0; i < 9; i++) gFtmp[i] += g_newF[i];
                }
            } else if (s->ptype == 0) {
                double Sn[9] = {0}, t2[9];
                for (int d = 0; d < 3; d++) Sn[4*d] = fmin(fmax(S[4*d], 1 - 2e-3), 1 + 3e-3);
                mm(g_newF, V, t); mm(t, Sn, t2); for (int i = 0; i < 9; i++) gU[i] += t2[i];      /* gU += g V Sn */
                mTm(g_newF, U, t); mm(t, Sn, t2); for (int i = 0; i < 9; i++) gV[i] += t2[i];     /* gV += g^T U Sn */
                mTm(U, g_newF, t); mm(t, V, t2);                                                    /* U^T g V */
                for (int d = 0; d < 3; d++) {
                    /* min(max(sig, lo), hi): max -> grad to sig iff lo < sig; min -> grad to inner iff inner < hi */
                    double sg = S[4*d], lo = 1 - 2e-3, hi = 1 + 3e-3;
                    double inner = fmax(sg, lo);
                    if (lo < sg && inner < hi) gS[4*d] += t2[4*d];
                }
            } else if (s->ptype == 1) {
                for (int i = 0; i < 9; i++) gFtmp[i] += g_newF[i];
            } else {
                g_J += (1.0/3.0)*pow(J, 1.0/3.0 - 1.0)*(g_newF[0] + g_newF[4] + g_newF[8]);
            }
        } else {
            double t[9], sym[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sym[3*i+j] = g_stress[3*i+j] + g_stress[3*j+i];
            mm(sym, newF, t);
            for (int i = 0; i < 9; i++) g_newF[i] += s->mu*t[i];
            g_J += s->lam/J*tr;
            if (s->ptype == 2) {
                double sq = sqrt(J);
                g_J += (g_newF[0] + g_newF[4])/(2*sq);
            } else {
                for (int i = 0; i < 9; i++) gFtmp[i] += g_newF[i];
            }
        }
        double K[9]; cof3(Ftmp, K);
        for (int i = 0; i < 9; i++) gFtmp[i] += g_J*K[i];
        if (ci >= 0) for (int d = 0; d < 3; d++) {
            #pragma omp atomic
            s->gaction[3*ci+d] += 6e-4*s->dt*g_imp[d];
        }
        if (s->collision_type == 1)
            for (int i = s->np - 1; i >= 0; i--) if (s->prim[i].enabled) {
                prim_grad G; pg_zero(&G);
                prim_collide_particle_adj(&s->prim[i], f, x, v, s->dt, g_imp, gx, gv, &G);
                prim_grad_commit(&s->prim[i], f, &G);
            }
    }
}

/* boundary_condition, :268-281; mask[d]=0 where the component was zeroed (for the adjoint) */
static void boundary_condition(const orc_sim *s, const int *I, double *v, int *mask) {
    const int bound = 3;
    mask[0] = mask[1] = mask[2] = 1;
    for (int d = 0; d < 3; d++) {
        if (I[d] < bound && v[d] < 0) { v[d] = 0; mask[d] = 0; }
        if (I[d] > s->ng - bound && v[d] > 0) { v[d] = 0; mask[d] = 0; }
        if (d == 1 && I[d] < bound && s->sticky_ground) { v[0] = v[1] = v[2] = 0; mask[0] = mask[1] = mask[2] = 0; }
    }
}

static void grid_op(orc_sim *s, int f, int accumulate_wrench) {                 /* :283-297 */
    int ng = s->ng;
    #pragma omp parallel for collapse(2)
    for (int i = 0; i < ng; i++) for (int j = 0; j < ng; j++) for (int k = 0; k < ng; k++) {
        size_t I = GIDX(s, i, j, k);
        if (s->gm[I] > 1e-10) {
            double v[3], inv = 1 / s->gm[I];
            for (int d = 0; d < 3; d++) v[d] = inv*s->gvin[3*I+d] + s->dt*s->gravity[d];
            if (s->collision_type == 0) {
                double gp[3] = {i*s->dx, j*s->dx, k*s->dx};
                for (int q = 0; q < s->np; q++) if (s->prim[q].enabled) {
                    double o[3]; prim_collide(&s->prim[q], f, gp, v, s->dt, s->gm[I], o, accumulate_wrench);
                    v[0] = o[0]; v[1] = o[1]; v[2] = o[2];
                }
            }
            int II[3] = {i, j, k}, mask[3];
            boundary_condition(s, II, v, mask);
            for (int d = 0; d < 3; d++) s->gvout[3*I+d] = v[d];
        }
    }
}
static void grid_op_grad(orc_sim *s, int f) {
    int ng = s->ng;
    #pragma omp parallel for collapse(2)
    for (int i = 0; i < ng; i++) for (int j = 0; j < ng; j++) for (int k = 0; k < ng; k++) {
        size_t I = GIDX(s, i, j, k);
        if (s->gm[I] > 1e-10) {
            double v[3], inv = 1 / s->gm[I], vin[ORC_MAXP + 1][3];
            for (int d = 0; d < 3; d++) v[d] = inv*s->gvin[3*I+d] + s->dt*s->gravity[d];
            double gp[3] = {i*s->dx, j*s->dx, k*s->dx};
            int nq = 0, which[ORC_MAXP];
            if (s->collision_type == 0)
                for (int q = 0; q < s->np; q++) if (s->prim[q].enabled) {
                    memcpy(vin[nq], v, 24); which[nq++] = q;
                    double o[3]; prim_collide(&s->prim[q], f, gp, v, s->dt, s->gm[I], o, 0);
                    v[0] = o[0]; v[1] = o[1]; v[2] = o[2];
                }
            int II[3] = {i, j, k}, mask[3];
            boundary_condition(s, II, v, mask);
            double g[3], gmass = 0;
            for (int d = 0; d < 3; d++) g[d] = mask[d] ? s->ggvout[3*I+d] : 0.0;
            for (int a = nq - 1; a >= 0; a--) {
                double gin[3] = {0, 0, 0}; prim_grad G; pg_zero(&G);
                prim_collide_adj(&s->prim[which[a]], f, gp, vin[a], s->dt, s->gm[I], g, gin, &gmass, &G);
                prim_grad_commit(&s->prim[which[a]], f, &G);
                g[0] = gin[0]; g[1] = gin[1]; g[2] = gin[2];
            }
            double dotv = 0;
            for (int d = 0; d < 3; d++) { s->ggvin[3*I+d] += inv*g[d]; dotv += s->gvin[3*I+d]*g[d]; }
            s->ggm[I] += gmass - dotv*inv*inv;
        }
    }
}

static void grid_op_mixed1(orc_sim *s, int f) {                                 /* :396-404 */
    (void)f; int ng = s->ng;
    #pragma omp parallel for collapse(2)
    for (int i = 0; i < ng; i++) for (int j = 0; j < ng; j++) for (int k = 0; k < ng; k++) {
        size_t I = GIDX(s, i, j, k);
        if (s->gm[I] > 1e-10) {
            double v[3], inv = 1 / s->gm[I];
            for (int d = 0; d < 3; d++) v[d] = inv*s->gvin[3*I+d] + s->dt*s->gravity[d];
            int II[3] = {i, j, k}, mask[3];
            boundary_condition(s, II, v, mask);
            for (int d = 0; d < 3; d++) { s->gvmix[3*I+d] = v[d]; s->gvout[3*I+d] += s->gvmix[3*I+d]; }
        }
    }
}
static void grid_op_mixed1_grad(orc_sim *s, int f) {
    (void)f; int ng = s->ng;
    #pragma omp parallel for collapse(2)
    for (int i = 0; i < ng; i++) for (int j = 0; j < ng; j++) for (int k = 0; k < ng; k++) {
        size_t I = GIDX(s, i, j, k);
        if (s->gm[I] > 1e-10) {
            double v[3], inv = 1 / s->gm[I];
            for (int d = 0; d < 3; d++) v[d] = inv*s->gvin[3*I+d] + s->dt*s->gravity[d];
            int II[3] = {i, j, k}, mask[3];
            boundary_condition(s, II, v, mask);
            double dotv = 0;
            for (int d = 0; d < 3; d++) {
                /* grid_v_out += grid_v_mixed ; grid_v_mixed = v_out */
                double g = mask[d] ? (s->ggvmix[3*I+d] + s->ggvout[3*I+d]) : 0.0;
                s->ggvin[3*I+d] += inv*g; dotv += s->gvin[3*I+d]*g;
            }
            s->ggm[I] -= dotv*inv*inv;
        }
    }
}
static void grid_op_mixed2(orc_sim *s, int f) {                                 /* :406-419 */
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        int base[3]; double fx[3], w[3][3];
        stencil(s, s->x + (f*n + p)*3, base, fx, w);
        double nv[3] = {0, 0, 0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2];
            const double *g = s->gvmix + 3*GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            for (int d = 0; d < 3; d++) nv[d] += wt*g[d];
        }
        for (int d = 0; d < 3; d++) s->vtmp[3*p+d] = nv[d];
    }
}
static void grid_op_mixed2_grad(orc_sim *s, int f) {
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        int base[3]; double fx[3], w[3][3], dw[3][3];
        stencil(s, s->x + (f*n + p)*3, base, fx, w); stencil_dw(fx, dw);
        const double *gnv = s->gvtmp + 3*p;
        double g_fx[3] = {0, 0, 0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2];
            size_t I = GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            double g_wt = 0;
            for (int d = 0; d < 3; d++) {
                #pragma omp atomic
                s->ggvmix[3*I+d] += wt*gnv[d];
                g_wt += s->gvmix[3*I+d]*gnv[d];
            }
            g_fx[0] += g_wt*dw[i][0]*w[j][1]*w[k][2];
            g_fx[1] += g_wt*w[i][0]*dw[j][1]*w[k][2];
            g_fx[2] += g_wt*w[i][0]*w[j][1]*dw[k][2];
        }
        for (int d = 0; d < 3; d++) s->gx[(f*n + p)*3 + d] += s->inv_dx*g_fx[d];
    }
}
static inline double life_of(const orc_sim *s, int f) {
    /* :425 `1 / (self.substeps - f % self.substeps)`: int/int true division, evaluated in f32 under
       Taichi's default_fp (Appendix B.5) and then promoted */
    return (double)(1.0f / (float)(s->substeps - f % s->substeps));
}
static void grid_op_mixed3(orc_sim *s, int f, int accumulate_wrench) {          /* :421-429 */
    size_t n = s->n; double life = life_of(s, f);
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        double vt[3] = {s->vtmp[3*p], s->vtmp[3*p+1], s->vtmp[3*p+2]};
        for (int i = 0; i < s->np; i++) if (s->prim[i].enabled) {
            double o[3]; cm_tape T;
            prim_collide_mixed_fwd(&s->prim[i], f, s->x + (f*n + p)*3, vt, s->p_mass, s->dt, life, o, &T, accumulate_wrench);
            vt[0] = o[0]; vt[1] = o[1]; vt[2] = o[2];
        }
        for (int d = 0; d < 3; d++) s->vtgt[3*p+d] = vt[d];
    }
}
static void grid_op_mixed3_grad(orc_sim *s, int f) {
    size_t n = s->n; double life = life_of(s, f);
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        const double *x = s->x + (f*n + p)*3;
        double vin[ORC_MAXP][3], vout[ORC_MAXP][3]; cm_tape T[ORC_MAXP]; int which[ORC_MAXP], nq = 0;
        double vt[3] = {s->vtmp[3*p], s->vtmp[3*p+1], s->vtmp[3*p+2]};
        for (int i = 0; i < s->np; i++) if (s->prim[i].enabled) {
            memcpy(vin[nq], vt, 24);
            prim_collide_mixed_fwd(&s->prim[i], f, x, vt, s->p_mass, s->dt, life, vout[nq], &T[nq], 0);
            memcpy(vt, vout[nq], 24); which[nq++] = i;
        }
        double g[3] = {s->gvtgt[3*p], s->gvtgt[3*p+1], s->gvtgt[3*p+2]};
        double *gx = s->gx + (f*n + p)*3;
        for (int a = nq - 1; a >= 0; a--) {
            double gin[3] = {0, 0, 0}; prim_grad G; pg_zero(&G);
            prim_collide_mixed_adj(&s->prim[which[a]], f, x, vin[a], vout[a], s->p_mass, s->dt, life, &T[a], g, gx, gin, &G);
            prim_grad_commit(&s->prim[which[a]], f, &G);
            g[0] = gin[0]; g[1] = gin[1]; g[2] = gin[2];
        }
        for (int d = 0; d < 3; d++) s->gvtmp[3*p+d] += g[d];
    }
}
static void grid_op_mixed4(orc_sim *s, int f) {                                 /* :431-443 */
    size_t n = s->n; const double alpha = 2.0;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        int base[3]; double fx[3], w[3][3];
        stencil(s, s->x + (f*n + p)*3, base, fx, w);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2];
            size_t I = GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            if (s->gm[I] > 1e-10)
                for (int d = 0; d < 3; d++) {
                    double val = alpha*wt*(s->vtmp[3*p+d] - s->vtgt[3*p+d]);
                    #pragma omp atomic
                    s->gvout[3*I+d] -= val;
                }
        }
    }
}
static void grid_op_mixed4_grad(orc_sim *s, int f) {
    size_t n = s->n; const double alpha = 2.0;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        int base[3]; double fx[3], w[3][3], dw[3][3];
        stencil(s, s->x + (f*n + p)*3, base, fx, w); stencil_dw(fx, dw);
        double g_fx[3] = {0, 0, 0}, g_d[3] = {0, 0, 0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2];
            size_t I = GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            if (s->gm[I] > 1e-10) {
                double g_wt = 0;
                for (int d = 0; d < 3; d++) {
                    double go = s->ggvout[3*I+d];
                    g_wt -= alpha*go*(s->vtmp[3*p+d] - s->vtgt[3*p+d]);
                    g_d[d] -= alpha*wt*go;
                }
                g_fx[0] += g_wt*dw[i][0]*w[j][1]*w[k][2];
                g_fx[1] += g_wt*w[i][0]*dw[j][1]*w[k][2];
                g_fx[2] += g_wt*w[i][0]*w[j][1]*dw[k][2];
            }
        }
        for (int d = 0; d < 3; d++) {
            s->gvtmp[3*p+d] += g_d[d]; s->gvtgt[3*p+d] -= g_d[d];
            s->gx[(f*n + p)*3 + d] += s->inv_dx*g_fx[d];
        }
    }
}

static void g2p(orc_sim *s, int f) {                                            /* :299-318 */
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        const double *x = s->x + (f*n + p)*3;
        int base[3]; double fx[3], w[3][3];
        stencil(s, x, base, fx, w);
        double nv[3] = {0, 0, 0}, nC[9] = {0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double dpos[3] = {i - fx[0], j - fx[1], k - fx[2]};
            double wt = w[i][0]*w[j][1]*w[k][2];
            const double *g = s->gvout + 3*GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            for (int d = 0; d < 3; d++) {
                nv[d] += wt*g[d];
                for (int e = 0; e < 3; e++) nC[3*d+e] += 4*s->inv_dx*wt*g[d]*dpos[e];
            }
        }
        double *v1 = s->v + ((f+1)*n + p)*3, *x1 = s->x + ((f+1)*n + p)*3;
        memcpy(v1, nv, 24); memcpy(s->C + ((f+1)*n + p)*9, nC, 72);
        for (int d = 0; d < 3; d++) x1[d] = x[d] + s->dt*v1[d];
    }
}
static void g2p_grad(orc_sim *s, int f) {
    size_t n = s->n;
    #pragma omp parallel for
    for (size_t p = 0; p < n; p++) {
        const double *x = s->x + (f*n + p)*3;
        int base[3]; double fx[3], w[3][3], dw[3][3];
        stencil(s, x, base, fx, w); stencil_dw(fx, dw);
        const double *gx1 = s->gx + ((f+1)*n + p)*3, *gv1 = s->gv + ((f+1)*n + p)*3, *gC1 = s->gC + ((f+1)*n + p)*9;
        double gnv[3], g_fx[3] = {0, 0, 0};
        for (int d = 0; d < 3; d++) gnv[d] = gv1[d] + s->dt*gx1[d];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double dpos[3] = {i - fx[0], j - fx[1], k - fx[2]};
            double wt = w[i][0]*w[j][1]*w[k][2];
            size_t I = GIDX(s, base[0]+i, base[1]+j, base[2]+k);
            const double *g = s->gvout + 3*I;
            double g_wt = 0, g_dpos[3] = {0, 0, 0};
            for (int d = 0; d < 3; d++) {
                double cd = gC1[3*d]*dpos[0] + gC1[3*d+1]*dpos[1] + gC1[3*d+2]*dpos[2];
                double gg = wt*gnv[d] + 4*s->inv_dx*wt*cd;
                #pragma omp atomic
                s->ggvout[3*I+d] += gg;
                g_wt += g[d]*gnv[d] + 4*s->inv_dx*g[d]*cd;
                for (int e = 0; e < 3; e++) g_dpos[e] += 4*s->inv_dx*wt*g[d]*gC1[3*d+e];
            }
            for (int e = 0; e < 3; e++) g_fx[e] -= g_dpos[e];
            g_fx[0] += g_wt*dw[i][0]*w[j][1]*w[k][2];
            g_fx[1] += g_wt*w[i][0]*dw[j][1]*w[k][2];
            g_fx[2] += g_wt*w[i][0]*w[j][1]*dw[k][2];
        }
        for (int d = 0; d < 3; d++) s->gx[(f*n + p)*3 + d] += gx1[d] + s->inv_dx*g_fx[d];
    }
}

/* forward_kinematics, primitive_base.py:280-283 */
static void forward_kinematics(orc_prim *P, int f, double dt) {
    double aa[3], dq[4];
    for (int i = 0; i < 3; i++) { P->pos[3*(f+1)+i] = P->pos[3*f+i] + P->v[3*f+i]*dt; aa[i] = P->w[3*f+i]*dt; }
    w2quat(aa, dq);
    qmul(dq, P->rot + 4*f, P->rot + 4*(f+1));
}
static void forward_kinematics_grad(orc_prim *P, int f, double dt) {
    double aa[3], dq[4], gdq[4] = {0,0,0,0}, gaa[3] = {0,0,0};
    for (int i = 0; i < 3; i++) {
        P->gpos[3*f+i] += P->gpos[3*(f+1)+i]; P->gv[3*f+i] += dt*P->gpos[3*(f+1)+i];
        aa[i] = P->w[3*f+i]*dt;
    }
    w2quat(aa, dq);
    qmul_adj(dq, P->rot + 4*f, P->grot + 4*(f+1), gdq, P->grot + 4*f);
    w2quat_adj(aa, gdq, gaa);
    for (int i = 0; i < 3; i++) P->gw[3*f+i] += dt*gaa[i];
}

/* ------------------------------------------------------------------------------------------ */
/* drivers                                                                                     */
/* ------------------------------------------------------------------------------------------ */
static void forward_to_grid(orc_sim *s, int f, int accumulate_wrench, int with_fk) {
    clear_grid(s);
    compute_F_tmp(s, f);
    if (s->material_model == 0) svd_kernel(s);
    p2g(s, f, accumulate_wrench);
    if (with_fk && s->rigid_velocity_control)
        for (int i = 0; i < s->np; i++) forward_kinematics(&s->prim[i], f, s->dt);
    if (s->collision_type == 2) {
        grid_op_mixed1(s, f); grid_op_mixed2(s, f); grid_op_mixed3(s, f, accumulate_wrench); grid_op_mixed4(s, f);
    } else {
        grid_op(s, f, accumulate_wrench);
    }
}
void orc_substep(orc_sim *s, int f) {                                           /* :320-337 */
    forward_to_grid(s, f, 1, 1);
    g2p(s, f);
}
void orc_substep_grad(orc_sim *s, int f) {                                      /* :339-378 */
    clear_grid(s);
    if (s->material_model == 0) clear_SVD_grad(s); else memset(s->gFtmp, 0, (size_t)s->n*72);
    compute_F_tmp(s, f);
    if (s->material_model == 0) svd_kernel(s);
    /* the reference re-accumulates ext_f here (harmless: cleared by the rigid bridge, :339-359) */
    p2g(s, f, 1);
    if (s->collision_type == 2) {
        grid_op_mixed1(s, f); grid_op_mixed2(s, f); grid_op_mixed3(s, f, 1); grid_op_mixed4(s, f);
    } else grid_op(s, f, 1);
    g2p_grad(s, f);
    if (s->collision_type == 2) {
        grid_op_mixed4_grad(s, f); grid_op_mixed3_grad(s, f); grid_op_mixed2_grad(s, f); grid_op_mixed1_grad(s, f);
    } else grid_op_grad(s, f);
    if (s->rigid_velocity_control)
        for (int i = s->np - 1; i >= 0; i--) forward_kinematics_grad(&s->prim[i], f, s->dt);
    p2g_grad(s, f);
    if (s->material_model == 0) svd_grad(s);
    compute_F_tmp_grad(s, f);
}

/* exposed pieces for unit tests */
void orc_svd3(const double *F, double *U, double *S, double *V) { svd3(F, U, S, V); }
void orc_backward_svd(const double *gu, const double *gs, const double *gv, const double *u, const double *s, const double *v, double *o) { backward_svd(gu, gs, gv, u, s, v, o); }
double orc_prim_sdf(orc_sim *s, int i, int f, const double *pos) { return prim_sdf(&s->prim[i], f, pos); }
void orc_prim_normal(orc_sim *s, int i, int f, const double *pos, double *n) { prim_normal(&s->prim[i], f, pos, n); }
void orc_get_grid(const orc_sim *s, double *gvin, double *gm, double *gvout) {
    size_t G = (size_t)s->ng*s->ng*s->ng;
    if (gvin) memcpy(gvin, s->gvin, G*24);
    if (gm) memcpy(gm, s->gm, G*8);
    if (gvout) memcpy(gvout, s->gvout, G*24);
}
/* torchrun exports OMP_NUM_THREADS=1; the timed CPU baseline asks for its thread count explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
