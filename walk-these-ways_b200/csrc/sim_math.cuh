// Device-side small linear algebra for the rigid-body step: 3-vectors, symmetric 3x3, spatial
// (6-D) vectors and symmetric 6x6 articulated inertias in [angular; linear] body coordinates,
// plus Philox4x32-10 for device randomisation.  fp32 throughout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct V3 { float x, y, z; };
struct Sym3 { float xx, xy, xz, yy, yz, zz; };
struct M3 { float m00, m01, m02, m10, m11, m12, m20, m21, m22; };
struct SV { V3 a, l; };                 // spatial motion or force vector
struct SI { Sym3 A; M3 B; Sym3 C; };    // symmetric 6x6 [[A, B],[B^T, C]]

#define DI __device__ __forceinline__

DI V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DI V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DI V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DI V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
DI V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
DI float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DI V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DI float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
DI void addcomp(V3& a, int k, float v) { if (k == 0) a.x += v; else if (k == 1) a.y += v; else a.z += v; }

DI V3 mul(const M3& M, V3 v) {
    return v3(M.m00 * v.x + M.m01 * v.y + M.m02 * v.z, M.m10 * v.x + M.m11 * v.y + M.m12 * v.z,
              M.m20 * v.x + M.m21 * v.y + M.m22 * v.z);
}
DI V3 mulT(const M3& M, V3 v) {
    return v3(M.m00 * v.x + M.m10 * v.y + M.m20 * v.z, M.m01 * v.x + M.m11 * v.y + M.m21 * v.z,
              M.m02 * v.x + M.m12 * v.y + M.m22 * v.z);
}
DI V3 mul(const Sym3& S, V3 v) {
    return v3(S.xx * v.x + S.xy * v.y + S.xz * v.z, S.xy * v.x + S.yy * v.y + S.yz * v.z,
              S.xz * v.x + S.yz * v.y + S.zz * v.z);
}
DI M3 matmul(const M3& A, const M3& B) {
    M3 C;
    C.m00 = A.m00 * B.m00 + A.m01 * B.m10 + A.m02 * B.m20; C.m01 = A.m00 * B.m01 + A.m01 * B.m11 + A.m02 * B.m21; C.m02 = A.m00 * B.m02 + A.m01 * B.m12 + A.m02 * B.m22;
    C.m10 = A.m10 * B.m00 + A.m11 * B.m10 + A.m12 * B.m20; C.m11 = A.m10 * B.m01 + A.m11 * B.m11 + A.m12 * B.m21; C.m12 = A.m10 * B.m02 + A.m11 * B.m12 + A.m12 * B.m22;
    C.m20 = A.m20 * B.m00 + A.m21 * B.m10 + A.m22 * B.m20; C.m21 = A.m20 * B.m01 + A.m21 * B.m11 + A.m22 * B.m21; C.m22 = A.m20 * B.m02 + A.m21 * B.m12 + A.m22 * B.m22;
    return C;
}
DI M3 transpose(const M3& A) { M3 T; T.m00 = A.m00; T.m01 = A.m10; T.m02 = A.m20; T.m10 = A.m01; T.m11 = A.m11; T.m12 = A.m21; T.m20 = A.m02; T.m21 = A.m12; T.m22 = A.m22; return T; }
DI M3 tofull(const Sym3& S) { M3 M; M.m00 = S.xx; M.m01 = S.xy; M.m02 = S.xz; M.m10 = S.xy; M.m11 = S.yy; M.m12 = S.yz; M.m20 = S.xz; M.m21 = S.yz; M.m22 = S.zz; return M; }
DI Sym3 symof(const M3& M) { Sym3 S; S.xx = M.m00; S.xy = 0.5f * (M.m01 + M.m10); S.xz = 0.5f * (M.m02 + M.m20); S.yy = M.m11; S.yz = 0.5f * (M.m12 + M.m21); S.zz = M.m22; return S; }
DI V3 col(const M3& M, int k) { return k == 0 ? v3(M.m00, M.m10, M.m20) : (k == 1 ? v3(M.m01, M.m11, M.m21) : v3(M.m02, M.m12, M.m22)); }
DI V3 row(const M3& M, int k) { return k == 0 ? v3(M.m00, M.m01, M.m02) : (k == 1 ? v3(M.m10, M.m11, M.m12) : v3(M.m20, M.m21, M.m22)); }
DI V3 col(const Sym3& S, int k) { return k == 0 ? v3(S.xx, S.xy, S.xz) : (k == 1 ? v3(S.xy, S.yy, S.yz) : v3(S.xz, S.yz, S.zz)); }

// rotation matrix of a child frame rotated by angle (c = cos, s = sin) about AXIS: child -> parent coords
template <int AXIS> DI M3 axis_rot(float c, float s) {
    M3 R;
    if (AXIS == 0) { R.m00 = 1; R.m01 = 0; R.m02 = 0; R.m10 = 0; R.m11 = c; R.m12 = -s; R.m20 = 0; R.m21 = s; R.m22 = c; }
    else           { R.m00 = c; R.m01 = 0; R.m02 = s; R.m10 = 0; R.m11 = 1; R.m12 = 0; R.m20 = -s; R.m21 = 0; R.m22 = c; }
    return R;
}
// v_parent = R v_child, specialised (exploits the zeros)
template <int AXIS> DI V3 rot_c2p(float c, float s, V3 v) {
    if (AXIS == 0) return v3(v.x, c * v.y - s * v.z, s * v.y + c * v.z);
    return v3(c * v.x + s * v.z, v.y, -s * v.x + c * v.z);
}
template <int AXIS> DI V3 rot_p2c(float c, float s, V3 v) {   // R^T v
    if (AXIS == 0) return v3(v.x, c * v.y + s * v.z, -s * v.y + c * v.z);
    return v3(c * v.x - s * v.z, v.y, s * v.x + c * v.z);
}

// M * R_axis(c, s): only two columns mix
template <int AXIS> DI M3 mul_axis(const M3& M, float c, float s) {
    M3 O = M;
    if (AXIS == 0) {
        O.m01 = c * M.m01 + s * M.m02; O.m11 = c * M.m11 + s * M.m12; O.m21 = c * M.m21 + s * M.m22;
        O.m02 = c * M.m02 - s * M.m01; O.m12 = c * M.m12 - s * M.m11; O.m22 = c * M.m22 - s * M.m21;
    } else {
        O.m00 = c * M.m00 - s * M.m02; O.m10 = c * M.m10 - s * M.m12; O.m20 = c * M.m20 - s * M.m22;
        O.m02 = s * M.m00 + c * M.m02; O.m12 = s * M.m10 + c * M.m12; O.m22 = s * M.m20 + c * M.m22;
    }
    return O;
}
// R_axis(c, s) * M * R_axis(c, s)^T: rotate the columns, then the rows (6 two-term rotations instead of two 3x3 products)
template <int AXIS> DI M3 rot_mat_axis(float c, float s, const M3& M) {
    const V3 c0 = rot_c2p<AXIS>(c, s, v3(M.m00, M.m10, M.m20)), c1 = rot_c2p<AXIS>(c, s, v3(M.m01, M.m11, M.m21)),
             c2 = rot_c2p<AXIS>(c, s, v3(M.m02, M.m12, M.m22));                     // X = R M (column k of X = R * column k of M)
    const V3 r0 = rot_c2p<AXIS>(c, s, v3(c0.x, c1.x, c2.x)), r1 = rot_c2p<AXIS>(c, s, v3(c0.y, c1.y, c2.y)),
             r2 = rot_c2p<AXIS>(c, s, v3(c0.z, c1.z, c2.z));                        // row i of X R^T = R * (row i of X)
    M3 O; O.m00 = r0.x; O.m01 = r0.y; O.m02 = r0.z; O.m10 = r1.x; O.m11 = r1.y; O.m12 = r1.z; O.m20 = r2.x; O.m21 = r2.y; O.m22 = r2.z;
    return O;
}

DI M3 quat_to_R(float x, float y, float z, float w) {   // xyzw, body -> world
    M3 R;
    R.m00 = 1 - 2 * (y * y + z * z); R.m01 = 2 * (x * y - z * w);     R.m02 = 2 * (x * z + y * w);
    R.m10 = 2 * (x * y + z * w);     R.m11 = 1 - 2 * (x * x + z * z); R.m12 = 2 * (y * z - x * w);
    R.m20 = 2 * (x * z - y * w);     R.m21 = 2 * (y * z + x * w);     R.m22 = 1 - 2 * (x * x + y * y);
    return R;
}
// isaacgym.torch_utils.quat_rotate_inverse restated: a - b + c  (public formula, xyzw)
DI V3 quat_rotate_inverse(float qx, float qy, float qz, float qw, V3 v) {
    V3 q = v3(qx, qy, qz);
    float s = 2.0f * qw * qw - 1.0f;
    V3 a = s * v;
    V3 b = (qw * 2.0f) * cross(q, v);
    V3 c = (2.0f * dot(q, v)) * q;
    return a - b + c;
}

// ---- spatial algebra ----
DI SV sv(V3 a, V3 l) { SV r; r.a = a; r.l = l; return r; }
DI SV operator+(SV p, SV q) { return sv(p.a + q.a, p.l + q.l); }
DI SV operator-(SV p, SV q) { return sv(p.a - q.a, p.l - q.l); }
DI SV operator*(float s, SV p) { return sv(s * p.a, s * p.l); }
DI float dot(SV p, SV q) { return dot(p.a, q.a) + dot(p.l, q.l); }
DI SV crf(SV v, SV f) { return sv(cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)); }   // v x* f
DI SV mul(const SI& I, SV v) { return sv(mul(I.A, v.a) + mul(I.B, v.l), mulT(I.B, v.a) + mul(I.C, v.l)); }

// rigid-body inertia about the link origin from the table: [Ixx Ixy Ixz Iyy Iyz Izz hx hy hz m], h = m*c
DI SI rigid_inertia(const float* t) {
    SI I;
    I.A.xx = t[0]; I.A.xy = t[1]; I.A.xz = t[2]; I.A.yy = t[3]; I.A.yz = t[4]; I.A.zz = t[5];
    float hx = t[6], hy = t[7], hz = t[8], m = t[9];
    I.B.m00 = 0; I.B.m01 = -hz; I.B.m02 = hy; I.B.m10 = hz; I.B.m11 = 0; I.B.m12 = -hx; I.B.m20 = -hy; I.B.m21 = hx; I.B.m22 = 0;
    I.C.xx = m; I.C.xy = 0; I.C.xz = 0; I.C.yy = m; I.C.yz = 0; I.C.zz = m;
    return I;
}
DI void add_inplace(SI& P, const SI& Q) {
    P.A.xx += Q.A.xx; P.A.xy += Q.A.xy; P.A.xz += Q.A.xz; P.A.yy += Q.A.yy; P.A.yz += Q.A.yz; P.A.zz += Q.A.zz;
    P.B.m00 += Q.B.m00; P.B.m01 += Q.B.m01; P.B.m02 += Q.B.m02; P.B.m10 += Q.B.m10; P.B.m11 += Q.B.m11; P.B.m12 += Q.B.m12; P.B.m20 += Q.B.m20; P.B.m21 += Q.B.m21; P.B.m22 += Q.B.m22;
    P.C.xx += Q.C.xx; P.C.xy += Q.C.xy; P.C.xz += Q.C.xz; P.C.yy += Q.C.yy; P.C.yz += Q.C.yz; P.C.zz += Q.C.zz;
}
// U = I e_k for an angular unit axis k
DI SV inertia_col_ang(const SI& I, int k) { return sv(col(I.A, k), row(I.B, k)); }
// I - U U^T * dinv
DI SI downdate(const SI& I, SV U, float dinv) {
    SI R = I;
    V3 ua = dinv * U.a, ul = dinv * U.l;
    R.A.xx -= U.a.x * ua.x; R.A.xy -= U.a.x * ua.y; R.A.xz -= U.a.x * ua.z; R.A.yy -= U.a.y * ua.y; R.A.yz -= U.a.y * ua.z; R.A.zz -= U.a.z * ua.z;
    R.B.m00 -= U.a.x * ul.x; R.B.m01 -= U.a.x * ul.y; R.B.m02 -= U.a.x * ul.z;
    R.B.m10 -= U.a.y * ul.x; R.B.m11 -= U.a.y * ul.y; R.B.m12 -= U.a.y * ul.z;
    R.B.m20 -= U.a.z * ul.x; R.B.m21 -= U.a.z * ul.y; R.B.m22 -= U.a.z * ul.z;
    R.C.xx -= U.l.x * ul.x; R.C.xy -= U.l.x * ul.y; R.C.xz -= U.l.x * ul.z; R.C.yy -= U.l.y * ul.y; R.C.yz -= U.l.y * ul.z; R.C.zz -= U.l.z * ul.z;
    return R;
}
DI Sym3 rot_sym(const M3& R, const Sym3& S) { return symof(matmul(matmul(R, tofull(S)), transpose(R))); }
DI M3 skew_mul(V3 r, const M3& M) {   // [r]x M : cross r with every column
    V3 c0 = cross(r, col(M, 0)), c1 = cross(r, col(M, 1)), c2 = cross(r, col(M, 2));
    M3 O; O.m00 = c0.x; O.m10 = c0.y; O.m20 = c0.z; O.m01 = c1.x; O.m11 = c1.y; O.m21 = c1.z; O.m02 = c2.x; O.m12 = c2.y; O.m22 = c2.z;
    return O;
}
// X^T Ia X: express the child's articulated inertia in the parent frame (child frame rotated by (c, s) about AXIS, origin r in parent)
template <int AXIS> DI SI transform_to_parent(const SI& Ia, float c, float s, V3 r) {
    Sym3 Ar = symof(rot_mat_axis<AXIS>(c, s, tofull(Ia.A))), Cr = symof(rot_mat_axis<AXIS>(c, s, tofull(Ia.C)));
    M3 Br = rot_mat_axis<AXIS>(c, s, Ia.B);
    SI P;
    P.C = Cr;
    M3 KC = skew_mul(r, tofull(Cr));
    M3 Bp;   // Br + [r]x Cr
    Bp.m00 = Br.m00 + KC.m00; Bp.m01 = Br.m01 + KC.m01; Bp.m02 = Br.m02 + KC.m02;
    Bp.m10 = Br.m10 + KC.m10; Bp.m11 = Br.m11 + KC.m11; Bp.m12 = Br.m12 + KC.m12;
    Bp.m20 = Br.m20 + KC.m20; Bp.m21 = Br.m21 + KC.m21; Bp.m22 = Br.m22 + KC.m22;
    P.B = Bp;
    // A_p = Ar + K Bp^T + Br K^T  (second + third terms sum to a symmetric matrix)
    M3 KBpT = skew_mul(r, transpose(Bp));
    M3 KBrT = skew_mul(r, transpose(Br));    // (Br K^T) = (K Br^T)^T
    M3 S;
    S.m00 = KBpT.m00 + KBrT.m00; S.m01 = KBpT.m01 + KBrT.m10; S.m02 = KBpT.m02 + KBrT.m20;
    S.m10 = KBpT.m10 + KBrT.m01; S.m11 = KBpT.m11 + KBrT.m11; S.m12 = KBpT.m12 + KBrT.m21;
    S.m20 = KBpT.m20 + KBrT.m02; S.m21 = KBpT.m21 + KBrT.m12; S.m22 = KBpT.m22 + KBrT.m22;
    Sym3 Ss = symof(S);
    P.A.xx = Ar.xx + Ss.xx; P.A.xy = Ar.xy + Ss.xy; P.A.xz = Ar.xz + Ss.xz; P.A.yy = Ar.yy + Ss.yy; P.A.yz = Ar.yz + Ss.yz; P.A.zz = Ar.zz + Ss.zz;
    return P;
}

// ---- 4-lane (one env) shuffles ----
DI float shfl4(float v, int src_leg) { return __shfl_sync(0xffffffffu, v, src_leg, 4); }
DI float allsum4(float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); return v; }
DI V3 allsum4(V3 v) { return v3(allsum4(v.x), allsum4(v.y), allsum4(v.z)); }
DI SV allsum4(SV v) { return sv(allsum4(v.a), allsum4(v.l)); }
DI V3 shfl4(V3 v, int s) { return v3(shfl4(v.x, s), shfl4(v.y, s), shfl4(v.z, s)); }
DI SV shfl4(SV v, int s) { return sv(shfl4(v.a, s), shfl4(v.l, s)); }

// ---- 6x6 SPD LDL^T (unit lower L stored row-wise below the diagonal, D inverted) ----
struct LDL6 { float L[15]; float Dinv[6]; };
DI void si_to_array(const SI& I, float A[6][6]) {
    M3 a = tofull(I.A), c = tofull(I.C); const M3& b = I.B; M3 bt = transpose(I.B);
    const M3* blk[2][2] = {{&a, &b}, {&bt, &c}};
#pragma unroll
    for (int bi = 0; bi < 2; bi++)
#pragma unroll
        for (int bj = 0; bj < 2; bj++) {
            const M3& m = *blk[bi][bj];
            A[3 * bi + 0][3 * bj + 0] = m.m00; A[3 * bi + 0][3 * bj + 1] = m.m01; A[3 * bi + 0][3 * bj + 2] = m.m02;
            A[3 * bi + 1][3 * bj + 0] = m.m10; A[3 * bi + 1][3 * bj + 1] = m.m11; A[3 * bi + 1][3 * bj + 2] = m.m12;
            A[3 * bi + 2][3 * bj + 0] = m.m20; A[3 * bi + 2][3 * bj + 1] = m.m21; A[3 * bi + 2][3 * bj + 2] = m.m22;
        }
}
DI LDL6 ldl_factor(const SI& I) {
    float A[6][6];
    si_to_array(I, A);
    float L[6][6], D[6];
    LDL6 F;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        float d = A[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * D[k];
        D[j] = d;
        float di = 1.0f / d;
        F.Dinv[j] = di;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            float s = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k] * D[k];
            L[i][j] = s * di;
        }
    }
    int n = 0;
#pragma unroll
    for (int i = 1; i < 6; i++)
#pragma unroll
        for (int j = 0; j < i; j++) F.L[n++] = L[i][j];
    return F;
}
DI SV ldl_solve(const LDL6& F, SV b) {   // solves I x = b
    float y[6] = {b.a.x, b.a.y, b.a.z, b.l.x, b.l.y, b.l.z};
    int n = 0;
#pragma unroll
    for (int i = 1; i < 6; i++)
#pragma unroll
        for (int j = 0; j < i; j++) y[i] -= F.L[n++] * y[j];
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] *= F.Dinv[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
#pragma unroll
        for (int k = i + 1; k < 6; k++) y[i] -= F.L[(k * (k - 1)) / 2 + i] * y[k];
    }
    return sv(v3(y[0], y[1], y[2]), v3(y[3], y[4], y[5]));
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG ----
struct Philox { uint32_t k0, k1; };
DI uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
DI float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }   // [0,1)
// uniform draw for (seed, env, step counter, slot)
static __device__ __noinline__ float philox_uniform(uint64_t seed, uint32_t env, uint64_t step, uint32_t slot) {
    uint4 c = make_uint4(slot >> 2, env, (uint32_t)step, (uint32_t)(step >> 32));
    uint4 r = philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t s = slot & 3;
    return u01(s == 0 ? r.x : (s == 1 ? r.y : (s == 2 ? r.z : r.w)));
}
