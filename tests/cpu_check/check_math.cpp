// Host build of smalfit_math.cuh for tests/test_math_host.py (TEST INFRASTRUCTURE:
// lets the closed forms the CUDA kernels use be compared with the oracle's autograd
// on a box without a GPU; the product never loads this library).
#include "../../smalify_b200/csrc/smalfit_math.cuh"
#include <cstring>
using namespace smf;

extern "C" {

void chk_rodrigues(const float* th, float* R) { rodrigues_fwd(th, R); }
void chk_rodrigues_bwd(const float* th, const float* Rb, float* thb) { thb[0] = thb[1] = thb[2] = 0.f; rodrigues_bwd(th, Rb, thb); }

// forward chain from theta [35*3], J [35*3], ls[6]; outputs G [35*9], off [35*3]
void chk_chain(const float* theta, const float* J, const float* ls, const int* parents, const int* scale_axis,
               float* G, float* off) {
    float R[NJ * 9], Rw[NJ * 9], s[NJ * 3], t[NJ * 3], Jc[NJ * 3];
    memcpy(Jc, J, sizeof(Jc));
    ChainFwd c{R, Rw, s, t, Jc, G, off};
    for (int j = 0; j < NJ; ++j) { rodrigues_fwd(theta + 3 * j, R + 9 * j); chain_scale(j, ls, scale_axis, s); }
    for (int j = 0; j < NJ; ++j) chain_fwd_joint(c, j, parents[j]);
}

// backward: given Gb, offb -> dtheta [35*3], dJ [35*3], dls [6]
void chk_chain_bwd(const float* theta, const float* J, const float* ls, const int* parents, const int* scale_axis,
                   const float* Gb_in, const float* offb_in, float* dtheta, float* dJ, float* dls) {
    float R[NJ * 9], Rw[NJ * 9], s[NJ * 3], t[NJ * 3], Jc[NJ * 3], G[NJ * 9], off[NJ * 3];
    memcpy(Jc, J, sizeof(Jc));
    ChainFwd c{R, Rw, s, t, Jc, G, off};
    for (int j = 0; j < NJ; ++j) { rodrigues_fwd(theta + 3 * j, R + 9 * j); chain_scale(j, ls, scale_axis, s); }
    for (int j = 0; j < NJ; ++j) chain_fwd_joint(c, j, parents[j]);
    float Gb[NJ * 9], offb[NJ * 3], tb[NJ * 3], Rwb[NJ * 9], sb[NJ * 3], Rb[NJ * 9];
    memcpy(Gb, Gb_in, sizeof(Gb)); memcpy(offb, offb_in, sizeof(offb));
    ChainBwd b{Gb, offb, tb, Rwb, sb, dJ, Rb};
    for (int j = 0; j < NJ; ++j) chain_bwd_local(c, b, j);
    for (int j = NJ - 1; j >= 0; --j) chain_bwd_push(c, b, j, parents[j]);
    for (int j = 0; j < NJ; ++j) { dtheta[3 * j] = dtheta[3 * j + 1] = dtheta[3 * j + 2] = 0.f; rodrigues_bwd(theta + 3 * j, Rb + 9 * j, dtheta + 3 * j); }
    for (int k = 0; k < NLS; ++k) dls[k] = 0.f;
    for (int i = 0; i < NJ * 3; ++i) if (scale_axis[i] >= 0) dls[scale_axis[i]] += sb[i] * s[i];
}

// one face against one pixel: returns 1 if a fragment exists; out = (pz, sd, p, m), grad[6] = d(sd)/dxy
int chk_face_eval(const float* tri /* x0,y0,z0,x1,y1,z1,x2,y2,z2 */, float px, float py, float* out, float* grad) {
    const FaceSetup fs = face_setup(tri[0], tri[1], tri[2], tri[3], tri[4], tri[5], tri[6], tri[7], tri[8]);
    Fragment fr;
    for (int k = 0; k < 6; ++k) grad[k] = 0.f;
    if (!face_eval(fs, px, py, fr)) return 0;
    float p, m;
    frag_prob(fr.sd, p, m);
    out[0] = fr.pz; out[1] = fr.sd; out[2] = p; out[3] = m;
    frag_grad(fr, 1.0f, grad);
    return 1;
}

// tile-rasteriser forward variant (prepared face): out = (pz, sd)
int chk_frag_setup_forward(const float* tri, float px, float py, float* out) {
    const FaceSetup fs = face_setup(tri[0], tri[1], tri[2], tri[3], tri[4], tri[5], tri[6], tri[7], tri[8]);
    if (fs.valid == 0.f) return 0;
    float sd = 0.f, pz = 0.f;
    const bool ok = frag_setup_forward(fs, px, py, sd, pz);
    out[0] = pz; out[1] = sd;
    return ok ? 1 : 0;
}

int chk_face_rect(const float* tri, int S, int* rect) {
    const FaceSetup fs = face_setup(tri[0], tri[1], tri[2], tri[3], tri[4], tri[5], tri[6], tri[7], tri[8]);
    return face_pixel_rect(fs, S, rect[0], rect[1], rect[2], rect[3]) ? 1 : 0;
}

void chk_camera(const float* X, float* ndc) { camera_fwd(X[0], X[1], X[2], ndc[0], ndc[1], ndc[2]); }
void chk_camera_bwd(const float* ndc, const float* g2, float* g3) { camera_bwd(ndc[0], ndc[1], ndc[2], g2[0], g2[1], g3[0], g3[1], g3[2]); }
}
