// smalfit_math.cuh -- closed-form maths of the SMAL fitting path, shared by every
// kernel.  Everything here is __host__ __device__ so tests/cpu_check can compile
// the same functions with g++ and compare them with the oracle; the product only
// ever runs them inside the CUDA kernels of smalfit_kernels.cu.
//
// Reference behaviour each block restates (paths in the SMALify checkout):
//   rodrigues_*      smal_model/batch_lbs.py:33-52 (eps added per component inside the norm)
//   chain_*          smal_model/batch_lbs.py:75-170 in the telescoped form
//                    G_i = Rw_i diag(s_i),  t_i = t_p + Rw_p diag(s_p)(J_i - J_p),
//                    A_i = [G_i | t_i - G_i J_i]   (S_parent^-1 cancels)
//   camera_*         PyTorch3D 0.2.5 look_at_view_transform(2.7,0,0) + OpenGLPerspectiveCameras
//                    as used at smal_fitter/p3d_renderer.py:22-23,67-68
//   face_setup/eval  PyTorch3D 0.2.5 CheckPixelInsideFace / PointTriangleDistance{Forward,Backward}
//                    (csrc/rasterize_meshes, csrc/utils/geometry_utils) as configured at
//                    smal_fitter/p3d_renderer.py:26-39
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SMF_HD __host__ __device__ __forceinline__
#else
#define SMF_HD inline
#endif

namespace smf {

constexpr int NJ = 35;            // joints incl. root
constexpr int NMJ = 41;           // model joints (35 regressed + 6 picked vertices)
constexpr int NKP = 25;           // annotated keypoints
constexpr int NBETA = 20;
constexpr int NLS = 6;
constexpr int MAXINF = 8;         // skinning influences per vertex

constexpr float CAM_DIST = 2.7f;
constexpr float CAM_F = 1.7320508075688772f;         // 1/tan(30 deg)
constexpr float RAST_SIGMA = 1e-4f;
constexpr float RAST_BLUR = 9.210240366975849e-4f;   // log(1/1e-4 - 1) * 1e-4   (NDC^2)
constexpr float RAST_BLUR_SQRT = 0.030348377826f;    // sqrt(RAST_BLUR)
constexpr float RAST_EPS = 1e-8f;                    // PyTorch3D kEpsilon
constexpr int RAST_K = 100;                          // faces_per_pixel
constexpr float RODRIGUES_EPS = 1e-8f;

// exact-rounding helpers: on the device these pin the operation order (no
// re-contraction), on the host they are the plain operators.
#if defined(__CUDA_ARCH__)
SMF_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
SMF_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
SMF_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
SMF_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SMF_HD float fsat(float a) { return __saturatef(a); }
#else
SMF_HD float fmul(float a, float b) { return a * b; }
SMF_HD float fsub(float a, float b) { return a - b; }
SMF_HD float fadd(float a, float b) { return a + b; }
SMF_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
SMF_HD float fsat(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
#endif

// a x b (2-D) with one rounding-pinned FMA; shared by every place that forms edge functions so
// that forward and backward see bit-identical barycentric numerators and depths
SMF_HD float cross2(float ax, float ay, float bx, float by) { return ffma(ax, by, -fmul(ay, bx)); }
SMF_HD float dot2(float ax, float ay, float bx, float by) { return ffma(ax, bx, fmul(ay, by)); }

// ---------------------------------------------------------------------------
// 3x3 helpers (row-major float[9])
// ---------------------------------------------------------------------------
SMF_HD void mat3_mul(const float* A, const float* B, float* C) {      // C = A B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
SMF_HD void mat3_mul_bt(const float* A, const float* B, float* C) {   // C = A B^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
SMF_HD void mat3_mul_at(const float* A, const float* B, float* C) {   // C = A^T B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}

// ---------------------------------------------------------------------------
// Rodrigues (batch_lbs.py:33-52)
// ---------------------------------------------------------------------------
SMF_HD void rodrigues_fwd(const float* th, float* R) {
    const float t0 = th[0], t1 = th[1], t2 = th[2];
    const float ux = t0 + RODRIGUES_EPS, uy = t1 + RODRIGUES_EPS, uz = t2 + RODRIGUES_EPS;
    const float a = sqrtf(ux * ux + uy * uy + uz * uz);
    const float rx = t0 / a, ry = t1 / a, rz = t2 / a;
    const float c = cosf(a), s = sinf(a), k = 1.f - c;
    R[0] = c + k * rx * rx; R[1] = k * rx * ry - s * rz; R[2] = k * rx * rz + s * ry;
    R[3] = k * ry * rx + s * rz; R[4] = c + k * ry * ry; R[5] = k * ry * rz - s * rx;
    R[6] = k * rz * rx - s * ry; R[7] = k * rz * ry + s * rx; R[8] = c + k * rz * rz;
}

// thb += dL/dtheta given Rb = dL/dR, differentiating the expression above.
SMF_HD void rodrigues_bwd(const float* th, const float* Rb, float* thb) {
    const float u[3] = {th[0] + RODRIGUES_EPS, th[1] + RODRIGUES_EPS, th[2] + RODRIGUES_EPS};
    const float a = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const float ia = 1.f / a;
    const float r[3] = {th[0] * ia, th[1] * ia, th[2] * ia};
    const float c = cosf(a), s = sinf(a), k = 1.f - c;
    // dR/da = -s I + s r r^T + c [r]x
    float ab = 0.f;
    ab += Rb[0] * (-s + s * r[0] * r[0]) + Rb[4] * (-s + s * r[1] * r[1]) + Rb[8] * (-s + s * r[2] * r[2]);
    ab += s * (r[0] * r[1] * (Rb[1] + Rb[3]) + r[0] * r[2] * (Rb[2] + Rb[6]) + r[1] * r[2] * (Rb[5] + Rb[7]));
    ab += c * (r[0] * (Rb[7] - Rb[5]) + r[1] * (Rb[2] - Rb[6]) + r[2] * (Rb[3] - Rb[1]));
    // dR/dr_k = k (e_k r^T + r e_k^T) + s [e_k]x
    float rb[3];
    rb[0] = k * (Rb[0] * r[0] + Rb[1] * r[1] + Rb[2] * r[2] + Rb[0] * r[0] + Rb[3] * r[1] + Rb[6] * r[2]) + s * (Rb[7] - Rb[5]);
    rb[1] = k * (Rb[3] * r[0] + Rb[4] * r[1] + Rb[5] * r[2] + Rb[1] * r[0] + Rb[4] * r[1] + Rb[7] * r[2]) + s * (Rb[2] - Rb[6]);
    rb[2] = k * (Rb[6] * r[0] + Rb[7] * r[1] + Rb[8] * r[2] + Rb[2] * r[0] + Rb[5] * r[1] + Rb[8] * r[2]) + s * (Rb[3] - Rb[1]);
    // r = th / a ,  a = |th + eps|
    ab -= (rb[0] * th[0] + rb[1] * th[1] + rb[2] * th[2]) * ia * ia;
    for (int i = 0; i < 3; ++i) thb[i] += rb[i] * ia + ab * u[i] * ia;
}

// ---------------------------------------------------------------------------
// Kinematic chain.  Arrays are [NJ][..] floats, anywhere (shared / host memory).
// ---------------------------------------------------------------------------
struct ChainFwd {          // pointers into caller storage
    float* R;    // [NJ*9] local rotations
    float* Rw;   // [NJ*9] accumulated rotations
    float* s;    // [NJ*3] per-axis scales
    float* t;    // [NJ*3] posed joint positions
    float* J;    // [NJ*3] rest joints
    float* G;    // [NJ*9] Rw diag(s)
    float* off;  // [NJ*3] t - G J
};

// scales of joint j from the 6 log-scales; scale_axis[j*3+a] in {-1,0..5}
SMF_HD void chain_scale(int j, const float* ls, const int* scale_axis, float* s) {
    for (int a = 0; a < 3; ++a) {
        const int k = scale_axis[j * 3 + a];
        s[j * 3 + a] = (k >= 0) ? expf(ls[k]) : 1.f;
    }
}

// joint i given its parent p is done (root: p < 0).  Inputs are read into locals first and results stored at the end:
// the arrays live in shared memory, and stores interleaved with loads through pointers the compiler cannot prove
// distinct would serialise every access (the chain is latency-bound: ~10 dependent levels per frame).
SMF_HD void chain_fwd_joint(const ChainFwd& c, int i, int p) {
    float Rw[9], G[9], t[3], off[3], J[3], si[3];
    for (int k = 0; k < 3; ++k) { J[k] = c.J[i * 3 + k]; si[k] = c.s[i * 3 + k]; }
    if (p < 0) {
        for (int k = 0; k < 9; ++k) Rw[k] = c.R[k];
        for (int k = 0; k < 3; ++k) t[k] = J[k];
    } else {
        float Rp[9], Ri[9], tp[3], d[3];
        for (int k = 0; k < 9; ++k) { Rp[k] = c.Rw[p * 9 + k]; Ri[k] = c.R[i * 9 + k]; }
        for (int k = 0; k < 3; ++k) { tp[k] = c.t[p * 3 + k]; d[k] = c.s[p * 3 + k] * (J[k] - c.J[p * 3 + k]); }
        mat3_mul(Rp, Ri, Rw);
        for (int r = 0; r < 3; ++r) t[r] = tp[r] + Rp[r * 3 + 0] * d[0] + Rp[r * 3 + 1] * d[1] + Rp[r * 3 + 2] * d[2];
    }
    for (int r = 0; r < 3; ++r)
        for (int a = 0; a < 3; ++a) G[r * 3 + a] = Rw[r * 3 + a] * si[a];
    for (int r = 0; r < 3; ++r) off[r] = t[r] - (G[r * 3 + 0] * J[0] + G[r * 3 + 1] * J[1] + G[r * 3 + 2] * J[2]);
    for (int k = 0; k < 9; ++k) { c.Rw[i * 9 + k] = Rw[k]; c.G[i * 9 + k] = G[k]; }
    for (int k = 0; k < 3; ++k) { c.t[i * 3 + k] = t[k]; c.off[i * 3 + k] = off[k]; }
}

struct ChainBwd {
    float* Gb;    // [NJ*9] in: dL/dG from skinning
    float* offb;  // [NJ*3] in: dL/doff from skinning
    float* tb;    // [NJ*3] work: dL/dt
    float* Rwb;   // [NJ*9] work: dL/dRw
    float* sb;    // [NJ*3] out: dL/ds
    float* Jb;    // [NJ*3] out: dL/dJ (rest joints)
    float* Rb;    // [NJ*9] out: dL/dR (local)
};

// Step 1 (any order, per joint): fold the A_i = [G_i | t_i - G_i J_i] layer.
// Initialises tb, Rwb, sb, Jb of joint i.  (locals first, stores last: see chain_fwd_joint)
SMF_HD void chain_bwd_local(const ChainFwd& c, const ChainBwd& b, int i) {
    float G[9], Rw[9], Gb[9], si[3], ob[3], Ji[3];
    for (int k = 0; k < 9; ++k) { G[k] = c.G[i * 9 + k]; Rw[k] = c.Rw[i * 9 + k]; Gb[k] = b.Gb[i * 9 + k]; }
    for (int k = 0; k < 3; ++k) { si[k] = c.s[i * 3 + k]; ob[k] = b.offb[i * 3 + k]; Ji[k] = c.J[i * 3 + k]; }
    // off = t - G J :  Gb_eff = Gb - offb J^T ,  Jb = -G^T offb
    float Ge[9], Jb[3], sb[3], Rwb[9];
    for (int r = 0; r < 3; ++r)
        for (int a = 0; a < 3; ++a) Ge[r * 3 + a] = Gb[r * 3 + a] - ob[r] * Ji[a];
    for (int a = 0; a < 3; ++a) Jb[a] = -(G[0 * 3 + a] * ob[0] + G[1 * 3 + a] * ob[1] + G[2 * 3 + a] * ob[2]);
    // G = Rw diag(s)
    for (int a = 0; a < 3; ++a) {
        sb[a] = Rw[0 * 3 + a] * Ge[0 * 3 + a] + Rw[1 * 3 + a] * Ge[1 * 3 + a] + Rw[2 * 3 + a] * Ge[2 * 3 + a];
        for (int r = 0; r < 3; ++r) Rwb[r * 3 + a] = Ge[r * 3 + a] * si[a];
    }
    for (int k = 0; k < 3; ++k) { b.tb[i * 3 + k] = ob[k]; b.Jb[i * 3 + k] = Jb[k]; b.sb[i * 3 + k] = sb[k]; }
    for (int k = 0; k < 9; ++k) b.Rwb[i * 9 + k] = Rwb[k];
}

// Step 2 (children before parents): push joint i's (complete) tb / Rwb to its parent p
// and emit Rb_i.  Must be serialised per parent (the kernels let the parent's thread run its children in turn).
SMF_HD void chain_bwd_push(const ChainFwd& c, const ChainBwd& b, int i, int p) {
    if (p < 0) {
        for (int k = 0; k < 9; ++k) b.Rb[k] = b.Rwb[k];
        for (int k = 0; k < 3; ++k) b.Jb[k] += b.tb[k];     // t_0 = J_0
        return;
    }
    float Rp[9], Ri[9], Rwbi[9], sp[3], tbi[3], delta[3];
    for (int k = 0; k < 9; ++k) { Rp[k] = c.Rw[p * 9 + k]; Ri[k] = c.R[i * 9 + k]; Rwbi[k] = b.Rwb[i * 9 + k]; }
    for (int k = 0; k < 3; ++k) { sp[k] = c.s[p * 3 + k]; tbi[k] = b.tb[i * 3 + k]; delta[k] = c.J[i * 3 + k] - c.J[p * 3 + k]; }
    float tbp[3], Rwbp[9], sbp[3], Jbi[3], Jbp[3];
    for (int k = 0; k < 3; ++k) { tbp[k] = b.tb[p * 3 + k]; sbp[k] = b.sb[p * 3 + k]; Jbi[k] = b.Jb[i * 3 + k]; Jbp[k] = b.Jb[p * 3 + k]; }
    for (int k = 0; k < 9; ++k) Rwbp[k] = b.Rwb[p * 9 + k];
    float RtT[3];
    for (int a = 0; a < 3; ++a) RtT[a] = Rp[0 * 3 + a] * tbi[0] + Rp[1 * 3 + a] * tbi[1] + Rp[2 * 3 + a] * tbi[2];   // (Rw_p^T tb_i)_a
    for (int r = 0; r < 3; ++r) {
        tbp[r] += tbi[r];
        for (int a = 0; a < 3; ++a) Rwbp[r * 3 + a] += tbi[r] * sp[a] * delta[a];
    }
    for (int a = 0; a < 3; ++a) {
        sbp[a] += RtT[a] * delta[a];
        const float db = sp[a] * RtT[a];
        Jbi[a] += db;
        Jbp[a] -= db;
    }
    // Rw_i = Rw_p R_i
    float tmp[9], Rbi[9];
    mat3_mul_bt(Rwbi, Ri, tmp);             // Rwb_i R_i^T
    for (int k = 0; k < 9; ++k) Rwbp[k] += tmp[k];
    mat3_mul_at(Rp, Rwbi, Rbi);             // Rw_p^T Rwb_i
    for (int k = 0; k < 3; ++k) { b.tb[p * 3 + k] = tbp[k]; b.sb[p * 3 + k] = sbp[k]; b.Jb[i * 3 + k] = Jbi[k]; b.Jb[p * 3 + k] = Jbp[k]; }
    for (int k = 0; k < 9; ++k) { b.Rwb[p * 9 + k] = Rwbp[k]; b.Rb[i * 9 + k] = Rbi[k]; }
}

// ---------------------------------------------------------------------------
// Camera
// ---------------------------------------------------------------------------
// f: focal factor of the NDC projection (1/tan(fov/2); the reference's fixed 60 degree camera = CAM_F).
SMF_HD void camera_fwd(float X, float Y, float Z, float& xn, float& yn, float& zv, float f = CAM_F) {
    zv = CAM_DIST - Z;
    const float iz = 1.f / zv;
    xn = -f * X * iz;
    yn = f * Y * iz;
}
// (gx, gy) = dL/d(x_ndc, y_ndc) -> dL/d(X,Y,Z); z_view carries no gradient.
SMF_HD void camera_bwd(float xn, float yn, float zv, float gx, float gy, float& gX, float& gY, float& gZ, float f = CAM_F) {
    const float iz = 1.f / zv;
    gX = -f * iz * gx;
    gY = f * iz * gy;
    gZ = (xn * gx + yn * gy) * iz;
}
// dL/df of one projected point: x_ndc and y_ndc are linear in f
SMF_HD float camera_bwd_focal(float xn, float yn, float gx, float gy, float f) { return (xn * gx + yn * gy) / f; }
// keypoint pixel (row, col) from NDC, transform_points_screen with (S-1)/2
SMF_HD void screen_fwd(float xn, float yn, float half, float& row, float& col) {
    col = half * (1.f - xn);
    row = half * (1.f - yn);
}

// pixel centre of the rasteriser (x is flipped: +X left, +Y up)
// (one rounding-pinned FMA, so that the packed two-pixel steps form bit-identical coordinates)
SMF_HD float pix_to_ndc(int i, float inv_s) { return ffma(-ffma(2.f, (float)i, 1.f), inv_s, 1.f); }

// ---------------------------------------------------------------------------
// Soft rasteriser: one face against one pixel
// ---------------------------------------------------------------------------
struct FaceSetup {         // 24 floats
    float x0, y0, x1, y1, x2, y2;
    float z0, z1, z2;
    float rden;            // 1 / (area + eps)
    float e01x, e01y, e02x, e02y, e12x, e12y;
    float rl01, rl02, rl12;          // 1 / |edge|^2 (0 for a degenerate edge)
    float bx0, bx1, by0, by1;        // blur-expanded bbox
    float valid;                     // 0/1
};

SMF_HD FaceSetup face_setup(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2) {
    FaceSetup f;
    f.x0 = x0; f.y0 = y0; f.x1 = x1; f.y1 = y1; f.x2 = x2; f.y2 = y2;
    f.z0 = z0; f.z1 = z1; f.z2 = z2;
    f.e01x = fsub(x1, x0); f.e01y = fsub(y1, y0);
    f.e02x = fsub(x2, x0); f.e02y = fsub(y2, y0);
    f.e12x = fsub(x2, x1); f.e12y = fsub(y2, y1);
    // area = EdgeFunction(v2; v0, v1) = (x2-x0)(y1-y0) - (y2-y0)(x1-x0)
    const float area = cross2(f.e02x, f.e02y, f.e01x, f.e01y);
    const float zmax = fmaxf(z0, fmaxf(z1, z2));
    const bool degenerate = (area <= RAST_EPS) && (area >= -RAST_EPS);
    f.valid = (zmax >= 0.f && !degenerate) ? 1.f : 0.f;
    f.rden = 1.f / fadd(area, RAST_EPS);
    const float l01 = dot2(f.e01x, f.e01y, f.e01x, f.e01y);
    const float l02 = dot2(f.e02x, f.e02y, f.e02x, f.e02y);
    const float l12 = dot2(f.e12x, f.e12y, f.e12x, f.e12y);
    f.rl01 = (l01 <= RAST_EPS) ? 0.f : 1.f / l01;
    f.rl02 = (l02 <= RAST_EPS) ? 0.f : 1.f / l02;
    f.rl12 = (l12 <= RAST_EPS) ? 0.f : 1.f / l12;
    f.bx0 = fsub(fminf(x0, fminf(x1, x2)), RAST_BLUR_SQRT);
    f.bx1 = fadd(fmaxf(x0, fmaxf(x1, x2)), RAST_BLUR_SQRT);
    f.by0 = fsub(fminf(y0, fminf(y1, y2)), RAST_BLUR_SQRT);
    f.by1 = fadd(fmaxf(y0, fmaxf(y1, y2)), RAST_BLUR_SQRT);
    return f;
}

// Conservative pixel rectangle [c0,c1]x[r0,r1] (inclusive, clipped) containing every
// pixel whose centre passes the bbox test; returns false if empty / face invalid.
SMF_HD bool face_pixel_rect(const FaceSetup& f, int S, int& c0, int& c1, int& r0, int& r1) {
    if (f.valid == 0.f) return false;
    const float hs = 0.5f * (float)S;
    // x(c) = 1 - (2c+1)/S in [bx0, bx1]  <=>  c in [ (1-bx1) S/2 - 1/2 , (1-bx0) S/2 - 1/2 ]
    const float cl = (1.f - f.bx1) * hs - 0.5f, ch = (1.f - f.bx0) * hs - 0.5f;
    const float rl = (1.f - f.by1) * hs - 0.5f, rh = (1.f - f.by0) * hs - 0.5f;
    if (!(ch >= -0.02f) || !(cl <= (float)S - 0.98f) || !(rh >= -0.02f) || !(rl <= (float)S - 0.98f)) return false;
    const float fS = (float)(S - 1);
    c0 = (int)fminf(fmaxf(ceilf(cl - 0.01f), 0.f), fS);
    c1 = (int)fminf(fmaxf(floorf(ch + 0.01f), 0.f), fS);
    r0 = (int)fminf(fmaxf(ceilf(rl - 0.01f), 0.f), fS);
    r1 = (int)fminf(fmaxf(floorf(rh + 0.01f), 0.f), fS);
    return (c0 <= c1) && (r0 <= r1);
}

struct Fragment {
    float pz;        // interpolated (unclipped barycentric) depth, >= 0
    float sd;        // signed squared distance: <0 inside
    int edge;        // closest edge: 0 = v0v1, 1 = v0v2, 2 = v1v2
    float t;         // clamped parameter on that edge
    float qx, qy;    // p_proj - p on that edge
};

// CheckPixelInsideFace.  Returns true when (face, pixel) yields a fragment.
// REGULAR: the caller knows that no edge of the face is degenerate (rl* != 0): the "distance to the end point" selects
// are compiled out; the value computed is the same.
template <bool REGULAR = false> SMF_HD bool face_eval_core(const FaceSetup& f, float px, float py, Fragment& fr);
SMF_HD bool face_eval(const FaceSetup& f, float px, float py, Fragment& fr) {
    if (f.valid == 0.f) return false;
    if (px > f.bx1 || px < f.bx0 || py > f.by1 || py < f.by0) return false;
    return face_eval_core<false>(f, px, py, fr);
}
// face_eval without the validity / bounding-box tests (implied for the pixels of a valid face's
// rectangle by the distance test): needs only the vertices, edges, rden and rl* of the set-up.
template <bool REGULAR> SMF_HD bool face_eval_core(const FaceSetup& f, float px, float py, Fragment& fr) {
    const float ax = fsub(px, f.x0), ay = fsub(py, f.y0);     // p - v0
    const float bx = fsub(px, f.x1), by = fsub(py, f.y1);     // p - v1
    const float cx = fsub(px, f.x2), cy = fsub(py, f.y2);     // p - v2
    // barycentric numerators (edge functions)
    const float n0 = cross2(bx, by, f.e12x, f.e12y);                  // E(p; v1, v2)
    const float n1 = cross2(f.e02x, f.e02y, cx, cy);                  // E(p; v2, v0) = (p-v2) x (v0-v2)
    const float n2 = cross2(ax, ay, f.e01x, f.e01y);                  // E(p; v0, v1)
    const float w0 = fmul(n0, f.rden), w1 = fmul(n1, f.rden), w2 = fmul(n2, f.rden);
    const float pz = ffma(w2, f.z2, ffma(w1, f.z1, fmul(w0, f.z0)));
    if (pz < 0.f) return false;
    // squared distances to the three segments
    const float t01 = (!REGULAR && f.rl01 == 0.f) ? 1.f : fsat(fmul(dot2(f.e01x, f.e01y, ax, ay), f.rl01));
    const float q01x = ffma(t01, f.e01x, -ax), q01y = ffma(t01, f.e01y, -ay);
    const float d01 = dot2(q01x, q01y, q01x, q01y);
    const float t02 = (!REGULAR && f.rl02 == 0.f) ? 1.f : fsat(fmul(dot2(f.e02x, f.e02y, ax, ay), f.rl02));
    const float q02x = ffma(t02, f.e02x, -ax), q02y = ffma(t02, f.e02y, -ay);
    const float d02 = dot2(q02x, q02y, q02x, q02y);
    const float t12 = (!REGULAR && f.rl12 == 0.f) ? 1.f : fsat(fmul(dot2(f.e12x, f.e12y, bx, by), f.rl12));
    const float q12x = ffma(t12, f.e12x, -bx), q12y = ffma(t12, f.e12y, -by);
    const float d12 = dot2(q12x, q12y, q12x, q12y);
    // closest edge, ties 01 -> 02 -> 12 (PointTriangleDistanceBackward order)
    int e; float d, t, qx, qy;
    if (d01 <= d02 && d01 <= d12) { e = 0; d = d01; t = t01; qx = q01x; qy = q01y; }
    else if (d02 <= d01 && d02 <= d12) { e = 1; d = d02; t = t02; qx = q02x; qy = q02y; }
    else { e = 2; d = d12; t = t12; qx = q12x; qy = q12y; }
    const bool inside = (w0 > 0.f) && (w1 > 0.f) && (w2 > 0.f);
    if (!inside && d >= RAST_BLUR) return false;
    fr.pz = pz; fr.sd = inside ? -d : d; fr.edge = e; fr.t = t; fr.qx = qx; fr.qy = qy;
    return true;
}

// Forward-side fragment test against a prepared face (edges, 1/(area+eps) and 1/|e|^2 from
// face_setup): face_eval without the bounding-box test (implied by the distance test) and without
// the closest-edge bookkeeping.  Depth and signed distance are formed with exactly face_eval's
// operations, so the forward's K-nearest thresholds and acceptance decisions agree bit for bit
// with what the backward recomputes.  Branch-free after the depth test.
template <bool REGULAR = false> SMF_HD bool frag_setup_forward(const FaceSetup& f, float px, float py, float& sd, float& pz) {
    const float ax = fsub(px, f.x0), ay = fsub(py, f.y0);
    const float bx = fsub(px, f.x1), by = fsub(py, f.y1);
    const float cx = fsub(px, f.x2), cy = fsub(py, f.y2);
    const float n0 = cross2(bx, by, f.e12x, f.e12y);
    const float n1 = cross2(f.e02x, f.e02y, cx, cy);
    const float n2 = cross2(ax, ay, f.e01x, f.e01y);
    const float w0 = fmul(n0, f.rden), w1 = fmul(n1, f.rden), w2 = fmul(n2, f.rden);
    pz = ffma(w2, f.z2, ffma(w1, f.z1, fmul(w0, f.z0)));
    if (pz < 0.f) return false;
    const float t01 = (!REGULAR && f.rl01 == 0.f) ? 1.f : fsat(fmul(dot2(f.e01x, f.e01y, ax, ay), f.rl01));
    const float q01x = ffma(t01, f.e01x, -ax), q01y = ffma(t01, f.e01y, -ay);
    const float d01 = dot2(q01x, q01y, q01x, q01y);
    const float t02 = (!REGULAR && f.rl02 == 0.f) ? 1.f : fsat(fmul(dot2(f.e02x, f.e02y, ax, ay), f.rl02));
    const float q02x = ffma(t02, f.e02x, -ax), q02y = ffma(t02, f.e02y, -ay);
    const float d02 = dot2(q02x, q02y, q02x, q02y);
    const float t12 = (!REGULAR && f.rl12 == 0.f) ? 1.f : fsat(fmul(dot2(f.e12x, f.e12y, bx, by), f.rl12));
    const float q12x = ffma(t12, f.e12x, -bx), q12y = ffma(t12, f.e12y, -by);
    const float d12 = dot2(q12x, q12y, q12x, q12y);
    const float d = fminf(d01, fminf(d02, d12));
    const bool inside = (w0 > 0.f) && (w1 > 0.f) && (w2 > 0.f);
    if (!inside && d >= RAST_BLUR) return false;
    sd = inside ? -d : d;
    return true;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// Packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2, PTX *.f32x2): one instruction works on two floats held in a 64-bit
// register pair.  The rasteriser sweeps a face's rectangle two horizontally adjacent pixels per lane: the x terms
// are packed, the y terms (same row) are scalars that the instructions broadcast.  Every packed operation is the IEEE
// round-to-nearest operation of its scalar twin, applied in the same order, so depths and signed distances are
// bit-identical to frag_setup_forward / face_eval_core (which the single-pixel steps and the host checks use).
// ---------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 f2_bc(float s) { return f2_pack(s, s); }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct Fragment2 {         // two pixels of one row against one face
    float pz[2], sd[2];
    // backward only: squared distances to the three segments, clamped parameters and -(p_proj - p) per edge
    f32x2 d01, d02, d12, t01, t02, t12, nq01x, nq01y, nq02x, nq02y, nq12x, nq12y;
};

// Two pixels (px.lo, py) and (px.hi, py) against a prepared face WITHOUT degenerate edges (rl* != 0; callers send the
// others through the scalar functions).  ok[k]: the pixel yields a fragment.  Same acceptance, depth and signed
// distance as frag_setup_forward / face_eval_core, operation for operation:
//   cross2(a, b) = fma(ax, by, -(ay * bx)),  dot2 = fma(ax, bx, ay * by)   (the rounded product is the y term where a
//   scalar is available; negations are moved onto scalars: (-a) * b == -(a * b), fma(t, -e, a) == -fma(t, e, -a))
template <bool BACKWARD>
__device__ __forceinline__ void face_eval2(const FaceSetup& f, f32x2 px, float py, bool ok[2], Fragment2& fr) {
    const f32x2 ax = f2_sub(px, f2_bc(f.x0)), bx = f2_sub(px, f2_bc(f.x1)), cx = f2_sub(px, f2_bc(f.x2));
    const float ay = fsub(py, f.y0), by = fsub(py, f.y1), cy = fsub(py, f.y2);
    // barycentric numerators
    const f32x2 n0 = f2_fma(bx, f2_bc(f.e12y), f2_bc(-fmul(by, f.e12x)));            // cross2(b, e12)
    const f32x2 n1 = f2_fma(f2_bc(f.e02x), f2_bc(cy), f2_mul(cx, f2_bc(-f.e02y)));   // cross2(e02, c)
    const f32x2 n2 = f2_fma(ax, f2_bc(f.e01y), f2_bc(-fmul(ay, f.e01x)));            // cross2(a, e01)
    const f32x2 rden = f2_bc(f.rden);
    const f32x2 w0 = f2_mul(n0, rden), w1 = f2_mul(n1, rden), w2 = f2_mul(n2, rden);
    const f32x2 pz = f2_fma(w2, f2_bc(f.z2), f2_fma(w1, f2_bc(f.z1), f2_mul(w0, f2_bc(f.z0))));
    // clamped projections on the three segments
    float u0, u1;
    f2_unpack(f2_mul(f2_fma(f2_bc(f.e01x), ax, f2_bc(fmul(f.e01y, ay))), f2_bc(f.rl01)), u0, u1);
    const f32x2 t01 = f2_pack(fsat(u0), fsat(u1));
    f2_unpack(f2_mul(f2_fma(f2_bc(f.e02x), ax, f2_bc(fmul(f.e02y, ay))), f2_bc(f.rl02)), u0, u1);
    const f32x2 t02 = f2_pack(fsat(u0), fsat(u1));
    f2_unpack(f2_mul(f2_fma(f2_bc(f.e12x), bx, f2_bc(fmul(f.e12y, by))), f2_bc(f.rl12)), u0, u1);
    const f32x2 t12 = f2_pack(fsat(u0), fsat(u1));
    // -(t e - a): the same magnitude bits as q = t e - a
    const f32x2 q01x = f2_fma(t01, f2_bc(-f.e01x), ax), q01y = f2_fma(t01, f2_bc(-f.e01y), f2_bc(ay));
    const f32x2 q02x = f2_fma(t02, f2_bc(-f.e02x), ax), q02y = f2_fma(t02, f2_bc(-f.e02y), f2_bc(ay));
    const f32x2 q12x = f2_fma(t12, f2_bc(-f.e12x), bx), q12y = f2_fma(t12, f2_bc(-f.e12y), f2_bc(by));
    const f32x2 d01 = f2_fma(q01x, q01x, f2_mul(q01y, q01y));
    const f32x2 d02 = f2_fma(q02x, q02x, f2_mul(q02y, q02y));
    const f32x2 d12 = f2_fma(q12x, q12x, f2_mul(q12y, q12y));
    float pz0, pz1, a0, a1, b0, b1, c0, c1, wa0, wa1, wb0, wb1, wc0, wc1;
    f2_unpack(pz, pz0, pz1);
    f2_unpack(d01, a0, a1); f2_unpack(d02, b0, b1); f2_unpack(d12, c0, c1);
    f2_unpack(w0, wa0, wa1); f2_unpack(w1, wb0, wb1); f2_unpack(w2, wc0, wc1);
    const float d0 = fminf(a0, fminf(b0, c0)), d1 = fminf(a1, fminf(b1, c1));
    const bool in0 = (wa0 > 0.f) && (wb0 > 0.f) && (wc0 > 0.f), in1 = (wa1 > 0.f) && (wb1 > 0.f) && (wc1 > 0.f);
    ok[0] = !(pz0 < 0.f) && (in0 || !(d0 >= RAST_BLUR));
    ok[1] = !(pz1 < 0.f) && (in1 || !(d1 >= RAST_BLUR));
    fr.pz[0] = pz0; fr.pz[1] = pz1;
    fr.sd[0] = in0 ? -d0 : d0; fr.sd[1] = in1 ? -d1 : d1;
    if (BACKWARD) {
        fr.d01 = d01; fr.d02 = d02; fr.d12 = d12; fr.t01 = t01; fr.t02 = t02; fr.t12 = t12;
        fr.nq01x = q01x; fr.nq01y = q01y; fr.nq02x = q02x; fr.nq02y = q02y; fr.nq12x = q12x; fr.nq12y = q12y;
    }
}

// frag_prob for two pixels: m = 1 - sigmoid(-sd / sigma)
__device__ __forceinline__ void frag_prob2(const float sd[2], float p[2], float m[2]) {
    float x0, x1;
    f2_unpack(f2_mul(f2_pack(sd[0], sd[1]), f2_bc(1.4426950408889634f / RAST_SIGMA)), x0, x1);
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
    f2_unpack(f2_add(f2_bc(1.f), f2_pack(e0, e1)), x0, x1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(p[0]) : "f"(x0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(p[1]) : "f"(x1));
    f2_unpack(f2_sub(f2_bc(1.f), f2_pack(p[0], p[1])), m[0], m[1]);
}
#endif  // __CUDACC__

// 1 - sigmoid(-sd/sigma) the way the reference forms it in fp32: p = sigmoid(x), m = 1 - p.
SMF_HD void frag_prob(float sd, float& p, float& m) {
#if defined(__CUDA_ARCH__)
    float e;      // exp(sd / sigma) as one scaled MUFU.EX2 (flushes to 0 / overflows to inf at the ends: p = 1 / p = 0)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(sd * (1.4426950408889634f / RAST_SIGMA)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(1.f + e));
#else
    p = 1.f / (1.f + expf(sd * (1.f / RAST_SIGMA)));
#endif
    m = 1.f - p;
}

// gradient of the signed distance w.r.t. the face's xy: g[0..5] += gs * d(sd)/d(x0,y0,x1,y1,x2,y2)
SMF_HD void frag_grad(const Fragment& fr, float gs, float* g) {
    const float gd = (fr.sd < 0.f) ? -gs : gs;            // d(sd)/d(d2) = inside ? -1 : 1
    const float ga = gd * (1.f - fr.t) * 2.f, gb = gd * fr.t * 2.f;
    int ia, ib;
    if (fr.edge == 0) { ia = 0; ib = 2; } else if (fr.edge == 1) { ia = 0; ib = 4; } else { ia = 2; ib = 4; }
    g[ia] += ga * fr.qx; g[ia + 1] += ga * fr.qy;
    g[ib] += gb * fr.qx; g[ib + 1] += gb * fr.qy;
}

}  // namespace smf
