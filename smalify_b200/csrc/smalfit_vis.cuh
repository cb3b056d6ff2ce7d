// smalfit_vis.cuh -- visualisation pass (row 8f-3; included by smalfit_kernels.cu): the reference's
// `color_renderer` (smal_fitter/p3d_renderer.py:41-59,70-72): MeshRasterizer(blur_radius = 0,
// faces_per_pixel = 1) + HardPhongShader with one PointLights at (0, 0, 3), constant vertex colour
// (config.MESH_COLOR), PyTorch3D 0.2.5 defaults otherwise: light ambient / diffuse / specular colours
// 0.5 / 0.3 / 0.2, material colours 1, shininess 64, white background, barycentrics not perspective
// corrected, camera centre (0, 0, 2.7).  Runs every 100 epochs in the reference (generate_visualization),
// so it is written for clarity, not speed:
//   vis_prepare   per vertex: camera (NDC + view depth) and the area-weighted vertex normal
//                 (Meshes.verts_normals_packed: sum of the incident faces' cross products, normalised)
//   vis_zbuffer   warp per face: lanes sweep the face's pixel box, inside test (all barycentrics > 0,
//                 CheckPixelInsideFace with blur 0), 64-bit atomicMin of (depth bits, face id): the
//                 nearest face, lower id first on equal depth -- independent of the order of arrival
//   vis_shade     per pixel: barycentric interpolation of the world position and normal, Phong terms,
//                 hard_rgb_blend.
#pragma once

namespace smf {

struct VisArgs {
    const float* verts;          // [n][V][3] world vertices (caller)
    float4* ndc;                 // [n][Vp]
    float* normals;              // [n][V][3]
    unsigned long long* zbuf;    // [n][S*S]
    float* rgb;                  // [n][3][S][S] (caller)
    int S;
    float color[3];
    float focal;
};

__global__ void __launch_bounds__(256) vis_prepare_kernel(ModelDev m, VisArgs a) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, fr = blockIdx.y;
    if (v >= m.V) return;
    const float* vb = a.verts + (size_t)fr * m.V * 3;
    float xn, yn, zv;
    camera_fwd(vb[v * 3], vb[v * 3 + 1], vb[v * 3 + 2], xn, yn, zv, a.focal);
    a.ndc[(size_t)fr * m.Vp + v] = make_float4(xn, yn, zv, 0.f);
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    for (int e = m.v2f_ptr[v]; e < m.v2f_ptr[v + 1]; ++e) {
        const ushort4 f4 = m.faces4[m.v2f_fc[e] >> 2];
        const float* p0 = vb + f4.x * 3; const float* p1 = vb + f4.y * 3; const float* p2 = vb + f4.z * 3;
        const float ux = p1[0] - p0[0], uy = p1[1] - p0[1], uz = p1[2] - p0[2];
        const float wx = p2[0] - p0[0], wy = p2[1] - p0[1], wz = p2[2] - p0[2];
        n0 += uy * wz - uz * wy; n1 += uz * wx - ux * wz; n2 += ux * wy - uy * wx;      // (v1 - v0) x (v2 - v0)
    }
    const float len = fmaxf(sqrtf(n0 * n0 + n1 * n1 + n2 * n2), 1e-6f);                 // F.normalize(eps = 1e-6)
    float* o = a.normals + ((size_t)fr * m.V + v) * 3;
    o[0] = n0 / len; o[1] = n1 / len; o[2] = n2 / len;
}

__global__ void __launch_bounds__(256) vis_zbuffer_kernel(ModelDev m, VisArgs a) {
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), fr = blockIdx.y;
    if (f >= m.F) return;
    const ushort4 f4 = m.faces4[f];
    const float4* ndc = a.ndc + (size_t)fr * m.Vp;
    const float4 A = ndc[f4.x], B = ndc[f4.y], C = ndc[f4.z];
    const FaceSetup fs = face_setup(A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z);
    if (fs.valid == 0.f) return;
    // pixel box of the triangle itself (blur 0): undo the blur margin of the set-up's box
    const int S = a.S;
    const float hs = 0.5f * (float)S;
    const float bx0 = fs.bx0 + RAST_BLUR_SQRT, bx1 = fs.bx1 - RAST_BLUR_SQRT, by0 = fs.by0 + RAST_BLUR_SQRT, by1 = fs.by1 - RAST_BLUR_SQRT;
    const int c0 = max((int)floorf((1.f - bx1) * hs - 0.5f) - 1, 0), c1 = min((int)ceilf((1.f - bx0) * hs - 0.5f) + 1, S - 1);
    const int r0 = max((int)floorf((1.f - by1) * hs - 0.5f) - 1, 0), r1 = min((int)ceilf((1.f - by0) * hs - 0.5f) + 1, S - 1);
    if (c0 > c1 || r0 > r1) return;
    const float inv_s = 1.f / (float)S;
    const int wd = c1 - c0 + 1, npx = wd * (r1 - r0 + 1);
    for (int i = lane; i < npx; i += 32) {
        const int rr = i / wd, x = c0 + (i - rr * wd), y = r0 + rr;
        Fragment frag;
        if (!face_eval_core(fs, pix_to_ndc(x, inv_s), pix_to_ndc(y, inv_s), frag)) continue;
        if (!(frag.sd < 0.f)) continue;                      // blur 0: only pixels inside the triangle
        const unsigned long long key = ((unsigned long long)__float_as_uint(frag.pz + 0.f) << 32) | (unsigned)f;
        atomicMin(a.zbuf + ((size_t)fr * S + y) * S + x, key);
    }
}

__global__ void __launch_bounds__(256) vis_shade_kernel(ModelDev m, VisArgs a) {
    const int S = a.S, fr = blockIdx.y;
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= S * S) return;
    const unsigned long long key = a.zbuf[(size_t)fr * S * S + pi];
    float r = 1.f, g = 1.f, b = 1.f;                         // BlendParams().background_color
    if (key != ~0ull) {
        const int f = (int)(key & 0xffffffffu);
        const ushort4 f4 = m.faces4[f];
        const float4* ndc = a.ndc + (size_t)fr * m.Vp;
        const float4 A = ndc[f4.x], B = ndc[f4.y], C = ndc[f4.z];
        const FaceSetup fs = face_setup(A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z);
        const float inv_s = 1.f / (float)S;
        const float px = pix_to_ndc(pi % S, inv_s), py = pix_to_ndc(pi / S, inv_s);
        // barycentrics exactly as the fragment test forms them
        const float w0 = fmul(cross2(fsub(px, fs.x1), fsub(py, fs.y1), fs.e12x, fs.e12y), fs.rden);
        const float w1 = fmul(cross2(fs.e02x, fs.e02y, fsub(px, fs.x2), fsub(py, fs.y2)), fs.rden);
        const float w2 = fmul(cross2(fsub(px, fs.x0), fsub(py, fs.y0), fs.e01x, fs.e01y), fs.rden);
        const float* vb = a.verts + (size_t)fr * m.V * 3;
        const float* nb = a.normals + (size_t)fr * m.V * 3;
        float P[3], N[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            P[k] = w0 * vb[f4.x * 3 + k] + w1 * vb[f4.y * 3 + k] + w2 * vb[f4.z * 3 + k];
            N[k] = w0 * nb[f4.x * 3 + k] + w1 * nb[f4.y * 3 + k] + w2 * nb[f4.z * 3 + k];
        }
        // lighting (pytorch3d/renderer/lighting.py diffuse / specular, F.normalize eps 1e-6)
        const float nl = fmaxf(sqrtf(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]), 1e-6f);
        const float n0 = N[0] / nl, n1 = N[1] / nl, n2 = N[2] / nl;
        float d0 = 0.f - P[0], d1 = 0.f - P[1], d2 = 3.f - P[2];                   // light at (0, 0, 3)
        const float dl = fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-6f);
        d0 /= dl; d1 /= dl; d2 /= dl;
        const float cosang = n0 * d0 + n1 * d1 + n2 * d2;
        const float diffuse = 0.3f * fmaxf(cosang, 0.f);
        float v0 = 0.f - P[0], v1 = 0.f - P[1], v2 = CAM_DIST - P[2];               // camera centre (0, 0, 2.7)
        const float vl = fmaxf(sqrtf(v0 * v0 + v1 * v1 + v2 * v2), 1e-6f);
        v0 /= vl; v1 /= vl; v2 /= vl;
        const float q0 = -d0 + 2.f * cosang * n0, q1 = -d1 + 2.f * cosang * n1, q2 = -d2 + 2.f * cosang * n2;
        const float al = fmaxf(v0 * q0 + v1 * q1 + v2 * q2, 0.f) * (cosang > 0.f ? 1.f : 0.f);
        const float specular = 0.2f * powf(al, 64.f);
        const float amb = 0.5f;
        r = (amb + diffuse) * a.color[0] + specular;
        g = (amb + diffuse) * a.color[1] + specular;
        b = (amb + diffuse) * a.color[2] + specular;
    }
    float* o = a.rgb + (size_t)fr * 3 * S * S;
    o[pi] = r; o[(size_t)S * S + pi] = g; o[2 * (size_t)S * S + pi] = b;
}

void launch_vis(const ModelDev& m, const VisArgs& a, int n, cudaStream_t st) {
    cudaMemsetAsync(a.zbuf, 0xff, sizeof(unsigned long long) * (size_t)n * a.S * a.S, st);
    vis_prepare_kernel<<<dim3((m.V + 255) / 256, n), 256, 0, st>>>(m, a);
    vis_zbuffer_kernel<<<dim3((m.F + 7) / 8, n), 256, 0, st>>>(m, a);
    vis_shade_kernel<<<dim3((a.S * a.S + 255) / 256, n), 256, 0, st>>>(m, a);
}

// Scratch comes from the fit's workspace (overwritten by the next step anyway): NDC vertices -> w.ndc,
// normals -> w.dvs, the 64-bit z-buffer -> w.pix.
void launch_vis_color(const ModelDev& m, const Workspace& w, const float* verts, int n, const float color[3], float focal,
                      float* rgb, cudaStream_t st) {
    VisArgs a;
    a.verts = verts; a.ndc = w.ndc; a.normals = w.dvs; a.zbuf = reinterpret_cast<unsigned long long*>(w.pix);
    a.rgb = rgb; a.S = w.S; a.color[0] = color[0]; a.color[1] = color[1]; a.color[2] = color[2]; a.focal = focal;
    launch_vis(m, a, n, st);
}

}  // namespace smf
