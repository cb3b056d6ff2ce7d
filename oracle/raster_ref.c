/* raster_ref.c -- CPU ORACLE (test / baseline infrastructure, never shipped in the product).
 *
 * Plain-C float32 restatement of the PyTorch3D 0.2.5 soft-silhouette path that
 * smal_fitter/p3d_renderer.py:26-39,66 configures (PyTorch3D is not vendored in the
 * reference and not installable here -- parity of this file is UNPINNED, see
 * oracle/smal_oracle.py):
 *
 *   RasterizeMeshesNaiveCpu      (csrc/rasterize_meshes/rasterize_meshes_cpu.cpp): for every
 *                                pixel, every face; keep the K smallest pz
 *   geometry                     (csrc/utils/geometry_utils.h): EdgeFunction, barycentrics over
 *                                (area + eps), PointLineDistance (segment), PointTriangleDistance
 *   sigmoid_alpha_blend          (renderer/blending.py): alpha = 1 - prod(1 - sigmoid(-d/sigma))
 *   RasterizeMeshesBackwardCpu   closest-edge segment-distance gradient
 *
 * mode 0 ("faithful"): single thread, every face against every pixel, exactly the
 *                      reference's O(S^2 F) loop.
 * mode 1 ("culled")  : OpenMP over pixel rows, faces pre-filtered per row by their
 *                      blur-expanded y-range (same results, what a fair CPU port would do).
 *
 * Build: see oracle/Makefile.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* -DRASTER_DOUBLE builds the same restatement in double precision (libraster_ref64.so, entry points suffixed
 * _f64): the float64 arm of whole-fit comparisons, where the float64 torch oracle would take hours. */
#ifdef RASTER_DOUBLE
#define float double
#define fminf fmin
#define fmaxf fmax
#define sqrtf sqrt
#define expf exp
#define raster_soft_silhouette raster_soft_silhouette_f64
#define raster_set_threads raster_set_threads_f64
#define raster_num_threads raster_num_threads_f64
#endif

#define K_EPS 1e-8f

typedef struct { float pz, sd; int f; } frag_t;

static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* squared distance from p to segment a-b; also returns t and (proj - p) */
static inline float seg_dist2(float px, float py, float ax, float ay, float bx, float by, float* t_out, float* qx, float* qy) {
    const float bax = bx - ax, bay = by - ay;
    const float l2 = bax * bax + bay * bay;
    if (l2 <= K_EPS) {
        *t_out = 1.f; *qx = bx - px; *qy = by - py;
        return (*qx) * (*qx) + (*qy) * (*qy);
    }
    float t = (bax * (px - ax) + bay * (py - ay)) / l2;
    t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    *t_out = t;
    *qx = ax + t * bax - px; *qy = ay + t * bay - py;
    return (*qx) * (*qx) + (*qy) * (*qy);
}

/* CheckPixelInsideFace.  verts: (V,3) = (x_ndc, y_ndc, z_view). */
static inline int eval_pair(const float* verts, const int* face, float px, float py, float blur, float rad,
                            float* pz_out, float* sd_out, int* edge_out, float* t_out, float* qx_out, float* qy_out) {
    const float* v0 = verts + 3 * face[0];
    const float* v1 = verts + 3 * face[1];
    const float* v2 = verts + 3 * face[2];
    const float xmin = fminf(v0[0], fminf(v1[0], v2[0])), xmax = fmaxf(v0[0], fmaxf(v1[0], v2[0]));
    const float ymin = fminf(v0[1], fminf(v1[1], v2[1])), ymax = fmaxf(v0[1], fmaxf(v1[1], v2[1]));
    const float zmax = fmaxf(v0[2], fmaxf(v1[2], v2[2]));
    if (px > xmax + rad || px < xmin - rad || py > ymax + rad || py < ymin - rad) return 0;
    const float area = edge_fn(v2[0], v2[1], v0[0], v0[1], v1[0], v1[1]);
    if (zmax < 0.f || (area <= K_EPS && area >= -K_EPS)) return 0;
    const float den = area + K_EPS;
    const float w0 = edge_fn(px, py, v1[0], v1[1], v2[0], v2[1]) / den;
    const float w1 = edge_fn(px, py, v2[0], v2[1], v0[0], v0[1]) / den;
    const float w2 = edge_fn(px, py, v0[0], v0[1], v1[0], v1[1]) / den;
    const float pz = w0 * v0[2] + w1 * v1[2] + w2 * v2[2];
    if (pz < 0.f) return 0;
    float t01, q01x, q01y, t02, q02x, q02y, t12, q12x, q12y;
    const float d01 = seg_dist2(px, py, v0[0], v0[1], v1[0], v1[1], &t01, &q01x, &q01y);
    const float d02 = seg_dist2(px, py, v0[0], v0[1], v2[0], v2[1], &t02, &q02x, &q02y);
    const float d12 = seg_dist2(px, py, v1[0], v1[1], v2[0], v2[1], &t12, &q12x, &q12y);
    float d; int e;
    if (d01 <= d02 && d01 <= d12) { d = d01; e = 0; *t_out = t01; *qx_out = q01x; *qy_out = q01y; }
    else if (d02 <= d01 && d02 <= d12) { d = d02; e = 1; *t_out = t02; *qx_out = q02x; *qy_out = q02y; }
    else { d = d12; e = 2; *t_out = t12; *qx_out = q12x; *qy_out = q12y; }
    const int inside = (w0 > 0.f) && (w1 > 0.f) && (w2 > 0.f);
    if (!inside && d >= blur) return 0;
    *pz_out = pz; *sd_out = inside ? -d : d; *edge_out = e;
    return 1;
}

/* keep the K smallest (pz, f): insertion into a sorted array */
static inline void push_frag(frag_t* q, int* n, int K, float pz, float sd, int f) {
    int i = *n;
    if (i == K) {
        const frag_t last = q[K - 1];
        if (!(pz < last.pz || (pz == last.pz && f < last.f))) return;
        i = K - 1;
    } else {
        (*n)++;
    }
    while (i > 0 && (q[i - 1].pz > pz || (q[i - 1].pz == pz && q[i - 1].f > f))) { q[i] = q[i - 1]; --i; }
    q[i].pz = pz; q[i].sd = sd; q[i].f = f;
}

/* One mesh.  Outputs alpha (S*S).  If grad_alpha != NULL also accumulates
 * grad_verts (V*3; only x,y receive gradient) for L = sum grad_alpha * alpha.
 * stats (may be NULL): [0] bbox-passing pairs, [1] fragments before the cap, [2] touched pixels,
 * [3] capped pixels. */
void raster_soft_silhouette(const float* verts, int V, const int* faces, int F, int S, int K, float sigma,
                            float blur, int mode, float* alpha, const float* grad_alpha, float* grad_verts,
                            long long* stats) {
    const float rad = sqrtf(blur);
    long long s_pair = 0, s_frag = 0, s_touch = 0, s_cap = 0;
    float* ylo = NULL; float* yhi = NULL;
    if (mode == 1) {
        ylo = (float*)malloc(sizeof(float) * F); yhi = (float*)malloc(sizeof(float) * F);
        for (int f = 0; f < F; ++f) {
            const float a = verts[3 * faces[3 * f] + 1], b = verts[3 * faces[3 * f + 1] + 1], c = verts[3 * faces[3 * f + 2] + 1];
            ylo[f] = fminf(a, fminf(b, c)) - rad; yhi[f] = fmaxf(a, fmaxf(b, c)) + rad;
        }
    }
    (void)V;
#pragma omp parallel if (mode == 1) reduction(+ : s_pair, s_frag, s_touch, s_cap)
    {
        frag_t* q = (frag_t*)malloc(sizeof(frag_t) * (size_t)K);
        int* rowf = (int*)malloc(sizeof(int) * (size_t)F);
        float* gloc = NULL;
        if (grad_alpha && grad_verts) gloc = (float*)calloc((size_t)V * 3, sizeof(float));
#pragma omp for schedule(dynamic, 4)
        for (int r = 0; r < S; ++r) {
            const float py = 1.f - (2.f * (float)r + 1.f) / (float)S;
            int nrow = F;
            if (mode == 1) {
                nrow = 0;
                for (int f = 0; f < F; ++f) if (!(py > yhi[f] || py < ylo[f])) rowf[nrow++] = f;
            }
            for (int c = 0; c < S; ++c) {
                const float px = 1.f - (2.f * (float)c + 1.f) / (float)S;
                int n = 0; long long nf = 0;
                for (int i = 0; i < nrow; ++i) {
                    const int f = (mode == 1) ? rowf[i] : i;
                    float pz, sd, t, qx, qy; int e;
                    const int* face = faces + 3 * f;
                    /* pair statistics: the bbox test alone */
                    {
                        const float* v0 = verts + 3 * face[0]; const float* v1 = verts + 3 * face[1]; const float* v2 = verts + 3 * face[2];
                        const float xmin = fminf(v0[0], fminf(v1[0], v2[0])), xmax = fmaxf(v0[0], fmaxf(v1[0], v2[0]));
                        const float ymin = fminf(v0[1], fminf(v1[1], v2[1])), ymax = fmaxf(v0[1], fmaxf(v1[1], v2[1]));
                        if (!(px > xmax + rad || px < xmin - rad || py > ymax + rad || py < ymin - rad)) s_pair++;
                        else continue;
                    }
                    if (!eval_pair(verts, face, px, py, blur, rad, &pz, &sd, &e, &t, &qx, &qy)) continue;
                    nf++;
                    push_frag(q, &n, K, pz, sd, f);
                }
                s_frag += nf;
                if (nf > 0) s_touch++;
                if (nf > K) s_cap++;
                float prod = 1.f;
                for (int k = 0; k < n; ++k) {
                    const float p = 1.f / (1.f + expf(q[k].sd / sigma));
                    prod *= (1.f - p);
                }
                alpha[(size_t)r * S + c] = 1.f - prod;
                if (gloc && n > 0) {
                    const float ga = grad_alpha[(size_t)r * S + c];
                    if (ga != 0.f && prod != 0.f) {
                        for (int k = 0; k < n; ++k) {
                            const int f = q[k].f;
                            float pz = 0.f, sd = 0.f, t = 0.f, qx = 0.f, qy = 0.f; int e = 0;
                            eval_pair(verts, faces + 3 * f, px, py, blur, rad, &pz, &sd, &e, &t, &qx, &qy);
                            const float p = 1.f / (1.f + expf(sd / sigma));
                            /* d alpha / d sd = -(1/sigma) p (1-p) prod_{j!=k}(1-p_j) = -(p/sigma) prod */
                            const float gs = ga * (-(p / sigma) * prod);
                            const float gd = (sd < 0.f) ? -gs : gs;
                            const int ia = (e == 2) ? 1 : 0, ib = (e == 0) ? 1 : 2;
                            const int va = faces[3 * f + ia], vb = faces[3 * f + ib];
                            gloc[3 * va + 0] += gd * (1.f - t) * 2.f * qx; gloc[3 * va + 1] += gd * (1.f - t) * 2.f * qy;
                            gloc[3 * vb + 0] += gd * t * 2.f * qx; gloc[3 * vb + 1] += gd * t * 2.f * qy;
                        }
                    }
                }
            }
        }
        if (gloc) {
#pragma omp critical
            for (int i = 0; i < V * 3; ++i) grad_verts[i] += gloc[i];
            free(gloc);
        }
        free(q); free(rowf);
    }
    if (ylo) { free(ylo); free(yhi); }
    if (stats) { stats[0] = s_pair; stats[1] = s_frag; stats[2] = s_touch; stats[3] = s_cap; }
}

void raster_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int raster_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
