// smx_sdf.cuh -- SDF / normal table construction from a triangle mesh on the GPU (set-up time).
//
// Replaces Mesh.task + Mesh.trimesh2sdf (softmac/engine/primitive/mesh.py:167-241), which call
// trimesh.proximity.ProximityQuery.signed_distance / on_surface (trimesh==3.21.5, requirements.txt:8):
//   sdf[i,j,k]    = signed distance (negative inside) at lower + (i,j,k)*dx
//   normal[i,j,k] = unit normal of the nearest triangle / (1 + 1e-8)
// One thread per sample, brute force over the triangles staged in shared memory (meshes have 12..2556 triangles,
// tables <= 360k samples), f64 so that the distances agree with the reference's cached tables to ~1e-16.
// Sign: generalised winding number (Van Oosterom-Strackee solid angles).  Ties between equidistant triangles (samples
// whose closest point lies on an edge / vertex) go to the lowest triangle index; trimesh resolves them by rounding
// noise, so on those samples any adjacent face normal is a valid answer.
#pragma once
#include <cuda_runtime.h>

namespace smx {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double a, double b, double c) { D3 r; r.x = a; r.y = b; r.z = c; return r; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 dcross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// Ericson, Real-Time Collision Detection 5.1.5
__device__ __forceinline__ D3 closest_on_triangle(D3 p, D3 a, D3 b, D3 c) {
    D3 ab = b - a, ac = c - a, ap = p - a;
    double d1 = ddot(ab, ap), d2 = ddot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return a;
    D3 bp = p - b;
    double d3_ = ddot(ab, bp), d4 = ddot(ac, bp);
    if (d3_ >= 0 && d4 <= d3_) return b;
    double vc = d1 * d4 - d3_ * d2;
    if (vc <= 0 && d1 >= 0 && d3_ <= 0) return a + (d1 / (d1 - d3_)) * ab;
    D3 cp = p - c;
    double d5 = ddot(ab, cp), d6 = ddot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return c;
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) return a + (d2 / (d2 - d6)) * ac;
    double va = d3_ * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3_) >= 0 && (d5 - d6) >= 0) return b + ((d4 - d3_) / ((d4 - d3_) + (d5 - d6))) * (c - b);
    double den = 1.0 / (va + vb + vc);
    return a + (vb * den) * ab + (vc * den) * ac;
}

#define SMX_SDF_TILE 256
__global__ void __launch_bounds__(128) k_build_sdf(const double* __restrict__ verts, const int* __restrict__ faces, int nf, int r0, int r1, int r2,
                                                  double lx, double ly, double lz, double dx, double* __restrict__ sdf, double* __restrict__ nrm) {
    __shared__ double tri[SMX_SDF_TILE][9];
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)r0 * r1 * r2;
    bool live = t < total;
    long long tt = live ? t : 0;
    int k = (int)(tt % r2), j = (int)((tt / r2) % r1), i = (int)(tt / ((long long)r1 * r2));
    D3 p = d3(lx + i * dx, ly + j * dx, lz + k * dx);
    double best = 1e300, wind = 0.0;
    int best_tri = 0;
    for (int base = 0; base < nf; base += SMX_SDF_TILE) {
        int cnt = min(SMX_SDF_TILE, nf - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 9; e += blockDim.x) {
            int f = e / 9, c = e % 9;
            tri[f][c] = verts[3 * faces[3 * (base + f) + c / 3] + c % 3];
        }
        __syncthreads();
        for (int f = 0; f < cnt; f++) {
            D3 a = d3(tri[f][0], tri[f][1], tri[f][2]), b = d3(tri[f][3], tri[f][4], tri[f][5]), c = d3(tri[f][6], tri[f][7], tri[f][8]);
            D3 q = closest_on_triangle(p, a, b, c) - p;
            double dd = ddot(q, q);
            if (dd < best * (1.0 - 1e-9) - 1e-30) { best = dd; best_tri = base + f; }       // strict: ties keep the lowest index
            D3 ua = a - p, ub = b - p, uc = c - p;
            double la = sqrt(ddot(ua, ua)), lb = sqrt(ddot(ub, ub)), lc = sqrt(ddot(uc, uc));
            double num = ddot(ua, dcross(ub, uc));
            double den = la * lb * lc + ddot(ua, ub) * lc + ddot(ub, uc) * la + ddot(uc, ua) * lb;
            wind += 2.0 * atan2(num, den);
        }
    }
    if (!live) return;
    bool inside = wind / (4.0 * 3.14159265358979323846) > 0.5;
    double dist = sqrt(best);
    sdf[t] = inside ? -dist : dist;
    D3 a = d3(verts[3 * faces[3 * best_tri]], verts[3 * faces[3 * best_tri] + 1], verts[3 * faces[3 * best_tri] + 2]);
    D3 b = d3(verts[3 * faces[3 * best_tri + 1]], verts[3 * faces[3 * best_tri + 1] + 1], verts[3 * faces[3 * best_tri + 1] + 2]);
    D3 c = d3(verts[3 * faces[3 * best_tri + 2]], verts[3 * faces[3 * best_tri + 2] + 1], verts[3 * faces[3 * best_tri + 2] + 2]);
    D3 n = dcross(b - a, c - a);
    double s = 1.0 / (sqrt(ddot(n, n)) * (1.0 + 1e-8));
    nrm[3 * t] = n.x * s; nrm[3 * t + 1] = n.y * s; nrm[3 * t + 2] = n.z * s;
}

}  // namespace smx
