// smx_math.cuh -- register-resident 3x3 / quaternion math for the MLS-MPM particle kernels (sm_100a).
//
// Everything here is fp32.  The reference computes in f64 (softmac/engine/mpm_simulator.py:19); the two
// places where fp32 would lose the answer -- the co-rotated stress of a near-isotropic F
// (mpm_simulator.py:228-236) and the 1/(s_j^2 - s_i^2) factors of the SVD adjoint (mpm_simulator.py:140-157)
// -- are evaluated in *deviation form*: all quantities are carried as differences from the identity
// (E = F - I, e_i = sigma_i - 1), which are exactly representable / accurately computable in fp32, and
// the SVD adjoint is folded analytically into divided differences so no large factor ever multiplies
// a rounded cancellation.  See DESIGN.md "Numerics".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smx {

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };
struct M3 { float m[9]; };   // row-major

__device__ __forceinline__ V3 v3(float a, float b, float c) { V3 r; r.x = a; r.y = b; r.z = c; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ void operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ void operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

__device__ __forceinline__ M3 m3_zero() {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.m[i] = 0.f;
    return r;
}
__device__ __forceinline__ M3 m3_identity() { M3 r = m3_zero(); r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
__device__ __forceinline__ M3 mul(const M3& A, const M3& B) {          // A B
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = fmaf(A.m[3 * i], B.m[j], fmaf(A.m[3 * i + 1], B.m[3 + j], A.m[3 * i + 2] * B.m[6 + j]));
    return C;
}
__device__ __forceinline__ M3 mulT(const M3& A, const M3& B) {         // A B^T
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = fmaf(A.m[3 * i], B.m[3 * j], fmaf(A.m[3 * i + 1], B.m[3 * j + 1], A.m[3 * i + 2] * B.m[3 * j + 2]));
    return C;
}
__device__ __forceinline__ M3 Tmul(const M3& A, const M3& B) {         // A^T B
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = fmaf(A.m[i], B.m[j], fmaf(A.m[3 + i], B.m[3 + j], A.m[6 + i] * B.m[6 + j]));
    return C;
}
__device__ __forceinline__ M3 add(const M3& A, const M3& B) { M3 C;
#pragma unroll
    for (int i = 0; i < 9; i++) C.m[i] = A.m[i] + B.m[i]; return C; }
__device__ __forceinline__ M3 sub(const M3& A, const M3& B) { M3 C;
#pragma unroll
    for (int i = 0; i < 9; i++) C.m[i] = A.m[i] - B.m[i]; return C; }
__device__ __forceinline__ M3 scale(float s, const M3& A) { M3 C;
#pragma unroll
    for (int i = 0; i < 9; i++) C.m[i] = s * A.m[i]; return C; }
__device__ __forceinline__ V3 mulv(const M3& A, V3 v) {
    return v3(fmaf(A.m[0], v.x, fmaf(A.m[1], v.y, A.m[2] * v.z)), fmaf(A.m[3], v.x, fmaf(A.m[4], v.y, A.m[5] * v.z)),
              fmaf(A.m[6], v.x, fmaf(A.m[7], v.y, A.m[8] * v.z)));
}
__device__ __forceinline__ V3 Tmulv(const M3& A, V3 v) {               // A^T v
    return v3(fmaf(A.m[0], v.x, fmaf(A.m[3], v.y, A.m[6] * v.z)), fmaf(A.m[1], v.x, fmaf(A.m[4], v.y, A.m[7] * v.z)),
              fmaf(A.m[2], v.x, fmaf(A.m[5], v.y, A.m[8] * v.z)));
}
__device__ __forceinline__ float trace(const M3& A) { return A.m[0] + A.m[4] + A.m[8]; }
__device__ __forceinline__ float det(const M3& A) {
    return A.m[0] * (A.m[4] * A.m[8] - A.m[5] * A.m[7]) - A.m[1] * (A.m[3] * A.m[8] - A.m[5] * A.m[6]) +
           A.m[2] * (A.m[3] * A.m[7] - A.m[4] * A.m[6]);
}
__device__ __forceinline__ M3 cofactor(const M3& A) {                  // d det / dA
    M3 K;
    K.m[0] = A.m[4] * A.m[8] - A.m[5] * A.m[7]; K.m[1] = -(A.m[3] * A.m[8] - A.m[5] * A.m[6]); K.m[2] = A.m[3] * A.m[7] - A.m[4] * A.m[6];
    K.m[3] = -(A.m[1] * A.m[8] - A.m[2] * A.m[7]); K.m[4] = A.m[0] * A.m[8] - A.m[2] * A.m[6]; K.m[5] = -(A.m[0] * A.m[7] - A.m[1] * A.m[6]);
    K.m[6] = A.m[1] * A.m[5] - A.m[2] * A.m[4]; K.m[7] = -(A.m[0] * A.m[5] - A.m[2] * A.m[3]); K.m[8] = A.m[0] * A.m[4] - A.m[1] * A.m[3];
    return K;
}
// det(I + E) - 1 without cancellation: tr E + (sum of principal 2x2 minors) + det E
__device__ __forceinline__ float det_minus_one(const M3& E) {
    float t1 = E.m[0] + E.m[4] + E.m[8];
    float t2 = (E.m[0] * E.m[4] - E.m[1] * E.m[3]) + (E.m[0] * E.m[8] - E.m[2] * E.m[6]) + (E.m[4] * E.m[8] - E.m[5] * E.m[7]);
    return t1 + t2 + det(E);
}

// ---------------------------------------------------------------------------------------------
// 3x3 SVD in deviation form.  Input E = F - I.  Output rotations U, V and e[i] = sigma_i - 1 with
// sigma_0 >= sigma_1 >= |sigma_2| and sign(det F) carried by sigma_2 -- the contract of ti.svd
// (taichi==1.4.1; call site softmac/engine/mpm_simulator.py:133).  Cyclic Jacobi on the symmetric
// S = F^T F - I = E + E^T + E^T E: rotation angles depend only on off-diagonals and diagonal
// *differences*, so eigenvalues lambda_i = sigma_i^2 - 1 come out with error ~1e-7 * |S|, i.e.
// relative to the deviation, not to 1.  sigma_i - 1 = lambda_i / (1 + sqrt(1 + lambda_i)).
// ---------------------------------------------------------------------------------------------
struct Svd { M3 U, V; float e[3]; };

__device__ __forceinline__ void jacobi_rot(float& app, float& aqq, float& apq, float& arp, float& arq,
                                           float& v0p, float& v0q, float& v1p, float& v1q, float& v2p, float& v2q) {
    // annihilate apq; r is the third index.  A <- J^T A J, V <- V J with J = [[c, s], [-s, c]] on (p, q).
    // t = tan(phi) = h / (d + sign(d) sqrt(d^2 + h^2)), d = aqq - app, h = 2 apq: branch-free, two MUFU.RSQ and
    // one MUFU.RCP; c gets one Newton step so that V stays orthonormal to fp32 rounding.
    float d = aqq - app, h = 2.f * apq;
    float x = fmaf(d, d, h * h);
    float r = x * rsqrtf(fmaxf(x, 1e-37f));
    float den = d + copysignf(r, d);
    float t = (x > 1e-37f) ? __fdividef(h, den) : 0.f;
    float y = fmaf(t, t, 1.f);
    float c = rsqrtf(y);
    c = c * fmaf(-0.5f * y, c * c, 1.5f);
    float s = t * c;
    app = fmaf(-t, apq, app); aqq = fmaf(t, apq, aqq); apq = 0.f;
    float nrp = c * arp - s * arq, nrq = s * arp + c * arq; arp = nrp; arq = nrq;
    float a;
    a = c * v0p - s * v0q; v0q = s * v0p + c * v0q; v0p = a;
    a = c * v1p - s * v1q; v1q = s * v1p + c * v1q; v1p = a;
    a = c * v2p - s * v2q; v2q = s * v2p + c * v2q; v2p = a;
}

#ifndef SMX_JACOBI_SWEEPS
#define SMX_JACOBI_SWEEPS 4
#endif

__device__ __forceinline__ Svd svd_dev(const M3& E) {
    // S = E + E^T + E^T E
    float s00 = 2.f * E.m[0] + (E.m[0] * E.m[0] + E.m[3] * E.m[3] + E.m[6] * E.m[6]);
    float s11 = 2.f * E.m[4] + (E.m[1] * E.m[1] + E.m[4] * E.m[4] + E.m[7] * E.m[7]);
    float s22 = 2.f * E.m[8] + (E.m[2] * E.m[2] + E.m[5] * E.m[5] + E.m[8] * E.m[8]);
    float s01 = E.m[1] + E.m[3] + (E.m[0] * E.m[1] + E.m[3] * E.m[4] + E.m[6] * E.m[7]);
    float s02 = E.m[2] + E.m[6] + (E.m[0] * E.m[2] + E.m[3] * E.m[5] + E.m[6] * E.m[8]);
    float s12 = E.m[5] + E.m[7] + (E.m[1] * E.m[2] + E.m[4] * E.m[5] + E.m[7] * E.m[8]);
    float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
#pragma unroll 1
    for (int sweep = 0; sweep < SMX_JACOBI_SWEEPS; sweep++) {
        // warp-uniform convergence test: off-diagonal mass below fp32 resolution of the diagonal, or below 1e-9 absolute
        // (an eigenvalue error of 1e-9 is 5e-7 of the plastic clip range; quantities this small do not move x/v/F/C)
        float off2 = s01 * s01 + s02 * s02 + s12 * s12, dg2 = s00 * s00 + s11 * s11 + s22 * s22;
        if (__all_sync(__activemask(), off2 <= fmaxf(1e-15f * dg2, 1e-18f))) break;
        jacobi_rot(s00, s11, s01, s02, s12, v00, v01, v10, v11, v20, v21);   // (0,1), r = 2
        jacobi_rot(s00, s22, s02, s01, s12, v00, v02, v10, v12, v20, v22);   // (0,2), r = 1
        jacobi_rot(s11, s22, s12, s01, s02, v01, v02, v11, v12, v21, v22);   // (1,2), r = 0
    }
    float l0 = s00, l1 = s11, l2 = s22;
    // sort descending, swapping columns of V (a swap flips det(V); fixed below)
    float t;
    bool flip = false;
#define SMX_SWAPCOL(la, lb, a0, b0, a1, b1, a2, b2) if (la < lb) { t = la; la = lb; lb = t; t = a0; a0 = b0; b0 = t; t = a1; a1 = b1; b1 = t; t = a2; a2 = b2; b2 = t; flip = !flip; }
    SMX_SWAPCOL(l0, l1, v00, v01, v10, v11, v20, v21)
    SMX_SWAPCOL(l1, l2, v01, v02, v11, v12, v21, v22)
    SMX_SWAPCOL(l0, l1, v00, v01, v10, v11, v20, v21)
#undef SMX_SWAPCOL
    if (flip) { v02 = -v02; v12 = -v12; v22 = -v22; }
    Svd r;
    r.V.m[0] = v00; r.V.m[1] = v01; r.V.m[2] = v02; r.V.m[3] = v10; r.V.m[4] = v11; r.V.m[5] = v12; r.V.m[6] = v20; r.V.m[7] = v21; r.V.m[8] = v22;
    // sigma = sqrt(1 + lambda) as x * rsqrt(x) (2 ulp) and sigma - 1 = lambda / (1 + sigma) with a fast reciprocal:
    // both errors are relative, i.e. ~1e-7 of the *deviation*
    float y0 = fmaxf(1.f + l0, 0.f), y1 = fmaxf(1.f + l1, 0.f), y2 = fmaxf(1.f + l2, 0.f);
    float sg0 = y0 * rsqrtf(fmaxf(y0, 1e-37f)), sg1 = y1 * rsqrtf(fmaxf(y1, 1e-37f)), sg2 = y2 * rsqrtf(fmaxf(y2, 1e-37f));
    r.e[0] = __fdividef(l0, 1.f + sg0); r.e[1] = __fdividef(l1, 1.f + sg1); r.e[2] = __fdividef(l2, 1.f + sg2);
    // B = F V = V + E V ; U from Gram-Schmidt on its columns (they are orthogonal up to rounding)
    M3 B = add(r.V, mul(E, r.V));
    V3 b0 = v3(B.m[0], B.m[3], B.m[6]), b1 = v3(B.m[1], B.m[4], B.m[7]), b2 = v3(B.m[2], B.m[5], B.m[8]);
    V3 u0, u1, u2;
    float n0 = dot(b0, b0);
    u0 = n0 > 1e-30f ? rsqrtf(n0) * b0 : v3(1.f, 0.f, 0.f);
    u1 = b1 - dot(u0, b1) * u0;
    float n1 = dot(u1, u1);
    if (n1 > 1e-30f && n1 > 1e-12f * n0) u1 = rsqrtf(n1) * u1;
    else {      // rank <= 1: any unit vector orthogonal to u0
        V3 a = fabsf(u0.x) < 0.6f ? v3(1.f, 0.f, 0.f) : v3(0.f, 1.f, 0.f);
        u1 = a - dot(u0, a) * u0; u1 = rsqrtf(dot(u1, u1)) * u1;
    }
    u2 = cross(u0, u1);
    if (dot(u2, b2) < 0.f) r.e[2] = -sg2 - 1.f;     // det F < 0: sigma_2 carries the sign
    r.U.m[0] = u0.x; r.U.m[3] = u0.y; r.U.m[6] = u0.z; r.U.m[1] = u1.x; r.U.m[4] = u1.y; r.U.m[7] = u1.z; r.U.m[2] = u2.x; r.U.m[5] = u2.y; r.U.m[8] = u2.z;
    return r;
}

// U diag(d) V^T
__device__ __forceinline__ M3 udvt(const M3& U, float d0, float d1, float d2, const M3& V) {
    M3 T = U;
    T.m[0] *= d0; T.m[3] *= d0; T.m[6] *= d0; T.m[1] *= d1; T.m[4] *= d1; T.m[7] *= d1; T.m[2] *= d2; T.m[5] *= d2; T.m[8] *= d2;
    return mulT(T, V);
}

// U diag(d) U^T (symmetric)
__device__ __forceinline__ M3 udut(const M3& U, float d0, float d1, float d2) {
    float t0 = U.m[0] * d0, t1 = U.m[1] * d1, t2 = U.m[2] * d2, t3 = U.m[3] * d0, t4 = U.m[4] * d1, t5 = U.m[5] * d2, t6 = U.m[6] * d0, t7 = U.m[7] * d1,
          t8 = U.m[8] * d2;
    M3 S;
    S.m[0] = fmaf(t0, U.m[0], fmaf(t1, U.m[1], t2 * U.m[2]));
    S.m[4] = fmaf(t3, U.m[3], fmaf(t4, U.m[4], t5 * U.m[5]));
    S.m[8] = fmaf(t6, U.m[6], fmaf(t7, U.m[7], t8 * U.m[8]));
    S.m[1] = S.m[3] = fmaf(t0, U.m[3], fmaf(t1, U.m[4], t2 * U.m[5]));
    S.m[2] = S.m[6] = fmaf(t0, U.m[6], fmaf(t1, U.m[7], t2 * U.m[8]));
    S.m[5] = S.m[7] = fmaf(t3, U.m[6], fmaf(t4, U.m[7], t5 * U.m[8]));
    return S;
}

// mpm_simulator.py:184-192
__device__ __forceinline__ float clamp_ref(float a) { return a >= 0.f ? fmaxf(a, 1e-6f) : fminf(a, -1e-6f); }

// ---------------------------------------------------------------------------------------------
// quaternion helpers (softmac/engine/primitive/primitive_utils.py:4-46) and their reverse mode
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 qvec(Q4 q) { return v3(q.x, q.y, q.z); }
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    V3 qv = qvec(q), uv = cross(qv, v), uuv = cross(qv, uv);
    return v + 2.f * (q.w * uv + uuv);
}
__device__ __forceinline__ void qrot_adj(Q4 q, V3 v, V3 go, Q4& gq, V3& gv) {
    V3 qv = qvec(q), uv = cross(qv, v);
    gv += go;
    V3 guv = (2.f * q.w) * go, guuv = 2.f * go;
    gq.w += 2.f * dot(uv, go);
    V3 gqv = cross(uv, guuv);           // uuv = qv x uv
    guv += cross(guuv, qv);
    gqv += cross(v, guv);               // uv = qv x v
    gv += cross(guv, qv);
    gq.x += gqv.x; gq.y += gqv.y; gq.z += gqv.z;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
    float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    Q4 r; r.w = q.w * s; r.x = q.x * s; r.y = q.y * s; r.z = q.z * s; return r;
}
// y = q/|q| ; gq += (gy - y (y.gy))/|q|
__device__ __forceinline__ void qnormalize_adj(Q4 q, Q4 gy, Q4& gq) {
    float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    Q4 y; y.w = q.w * s; y.x = q.x * s; y.y = q.y * s; y.z = q.z * s;
    float d = y.w * gy.w + y.x * gy.x + y.y * gy.y + y.z * gy.z;
    gq.w += (gy.w - y.w * d) * s; gq.x += (gy.x - y.x * d) * s; gq.y += (gy.y - y.y * d) * s; gq.z += (gy.z - y.z * d) * s;
}
__device__ __forceinline__ V3 normalize3(V3 v) { return rsqrtf(dot(v, v)) * v; }
__device__ __forceinline__ void normalize3_adj(V3 x, V3 gy, V3& gx) {
    float s = rsqrtf(dot(x, x));
    V3 y = s * x;
    gx += s * (gy - dot(y, gy) * y);
}
__device__ __forceinline__ Q4 qconj(Q4 q) { Q4 r; r.w = q.w; r.x = -q.x; r.y = -q.y; r.z = -q.z; return r; }
__device__ __forceinline__ Q4 q4_zero() { Q4 r; r.w = r.x = r.y = r.z = 0.f; return r; }

}  // namespace smx
