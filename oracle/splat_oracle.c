/*
 * splat_oracle.c - CPU restatement of the reference's differentiable Gaussian
 * rasterizer.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle of the repo (tier rule 3): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker.  The product (libs3g_b200.so) never
 * links, loads or calls anything under oracle/.
 *
 * What it restates (paths relative to /root/reference,
 * DGR = submodules/depth-diff-gaussian-rasterization/cuda_rasterizer):
 *   preprocess      DGR/forward.cu:155-256, :74-152, :20-71; DGR/auxiliary.h:41-77,139-164
 *   binning         DGR/rasterizer_impl.cu:70-138 (keys, stable sort, tile ranges), :35-50
 *   composite fwd   DGR/forward.cu:261-379
 *   composite bwd   DGR/backward.cu:415-590
 *   preprocess bwd  DGR/backward.cu:144-274 (cov2D), :346-412 (means/depth), :20-139 (SH),
 *                   :278-341 (scale/rotation)
 * The algorithm is written from the arithmetic description (SURVEY.md appendix
 * A), one plain scalar loop per stage, single precision like the reference.
 * Compiled with -ffp-contract=off: the GPU builds (reference and ours) contract
 * a*b+c into FMAs, so float results agree to rounding (tests use 1e-4 relative)
 * and integer outputs (radii, tile lists) agree except where a value sits within
 * an ulp of a rounding boundary.
 *
 * Pinning: the reference ships no golden vectors for this path;
 * this oracle is pinned against outputs of the REAL reference extension
 * (oracle/_ref, built by oracle/build_ref.sh) captured on a B200 and committed
 * under tests/golden/ (see tests/golden/README.md and tools/make_golden_raster.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define NCH 3

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    int P, D, M, W, H, gx, gy;
    int64_t R;
    /* per Gaussian */
    float* depth;       /* [P]   view-space z                         */
    float* xy;          /* [P,2] pixel-space mean                     */
    float* conic_o;     /* [P,4] conic a,b,c + opacity                */
    float* rgb;         /* [P,3]                                      */
    float* cov3d;       /* [P,6]                                      */
    uint8_t* clamped;   /* [P,3]                                      */
    int* radii;         /* [P]                                        */
    uint32_t* tiles;    /* [P]   tiles touched                        */
    /* binning */
    uint64_t* keys;     /* [R] sorted                                 */
    uint32_t* plist;    /* [R] sorted Gaussian ids                    */
    uint32_t* ranges;   /* [tiles,2]                                  */
    /* image */
    float* final_T;     /* [H*W] */
    uint32_t* n_contrib;
} Oracle;

static void o_free_arrays(Oracle* o) {
    free(o->depth); free(o->xy); free(o->conic_o); free(o->rgb); free(o->cov3d); free(o->clamped);
    free(o->radii); free(o->tiles); free(o->keys); free(o->plist); free(o->ranges);
    free(o->final_T); free(o->n_contrib);
    memset(o, 0, sizeof(*o));
}

void* oracle_create(void) { return calloc(1, sizeof(Oracle)); }
void oracle_destroy(void* h) {
    if (!h) return;
    o_free_arrays((Oracle*)h);
    free(h);
}

/* column-major 4x4 applied to (p,1): rows 0..2 (auxiliary.h:58-66) and 0..3 (:68-77) */
static void xf3(const float* m, const float* p, float* r) {
    for (int i = 0; i < 3; ++i) r[i] = m[i] * p[0] + m[4 + i] * p[1] + m[8 + i] * p[2] + m[12 + i];
}
static void xf4(const float* m, const float* p, float* r) {
    for (int i = 0; i < 4; ++i) r[i] = m[i] * p[0] + m[4 + i] * p[1] + m[8 + i] * p[2] + m[12 + i];
}

/* Sigma = R S^2 R^T from scale and (r,x,y,z) quaternion; forward.cu:118-152 builds
 * M = S * Rq^T (row i of M = s_i * column i of the rotation) and Sigma = M^T M. */
static void rot_from_quat(const float* q, float Rm[3][3]) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    Rm[0][0] = 1.f - 2.f * (y * y + z * z); Rm[0][1] = 2.f * (x * y - r * z); Rm[0][2] = 2.f * (x * z + r * y);
    Rm[1][0] = 2.f * (x * y + r * z); Rm[1][1] = 1.f - 2.f * (x * x + z * z); Rm[1][2] = 2.f * (y * z - r * x);
    Rm[2][0] = 2.f * (x * z - r * y); Rm[2][1] = 2.f * (y * z + r * x); Rm[2][2] = 1.f - 2.f * (x * x + y * y);
}
static void cov3d_of(const float* scale, float mod, const float* q, float* c6) {
    float Rm[3][3];
    rot_from_quat(q, Rm);
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    /* Sigma_ij = sum_k Rm[i][k] s_k^2 Rm[j][k], evaluated as (s_k Rm[i][k]) (s_k Rm[j][k]) */
    float L[3][3];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) L[i][k] = s[k] * Rm[i][k];
    float S[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[i][j] = L[i][0] * L[j][0] + L[i][1] * L[j][1] + L[i][2] * L[j][2];
    c6[0] = S[0][0]; c6[1] = S[0][1]; c6[2] = S[0][2]; c6[3] = S[1][1]; c6[4] = S[1][2]; c6[5] = S[2][2];
}

/* EWA projection: cov2D = (J Wv) Sigma (J Wv)^T + 0.3 I  (forward.cu:74-113).
 * Also returns the intermediate 2x3 matrix A = J Wv and the clamped t. */
static void cov2d_of(const float* mean, float fx, float fy, float tfx, float tfy, const float* c6,
                     const float* view, float* cov /*a,b,c*/, float A[2][3], float* t_out,
                     float* txtz_out, float* tytz_out) {
    float t[3];
    xf3(view, mean, t);
    float limx = 1.3f * tfx, limy = 1.3f * tfy;
    float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
    float J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* Wv = rotation part of the view matrix, Wv[r][c] = view[c*4 + r] */
    for (int c = 0; c < 3; ++c) {
        float w0 = view[c * 4 + 0], w1 = view[c * 4 + 1], w2 = view[c * 4 + 2];
        A[0][c] = J00 * w0 + J02 * w2;
        A[1][c] = J11 * w1 + J12 * w2;
    }
    float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float AS[2][3];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) AS[i][j] = A[i][0] * S[0][j] + A[i][1] * S[1][j] + A[i][2] * S[2][j];
    cov[0] = AS[0][0] * A[0][0] + AS[0][1] * A[0][1] + AS[0][2] * A[0][2] + 0.3f;
    cov[1] = AS[0][0] * A[1][0] + AS[0][1] * A[1][1] + AS[0][2] * A[1][2];
    cov[2] = AS[1][0] * A[1][0] + AS[1][1] * A[1][1] + AS[1][2] * A[1][2] + 0.3f;
    if (t_out) { t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2]; }
    if (txtz_out) *txtz_out = txtz;
    if (tytz_out) *tytz_out = tytz;
}

/* SH basis up to degree 3 at unit direction d (forward.cu:30-59) */
static int sh_basis(int deg, const float* d, float* b) {
    float x = d[0], y = d[1], z = d[2];
    b[0] = SH_C0;
    if (deg < 1) return 1;
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (deg < 2) return 4;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (2.f * zz - xx - yy);
    b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
    if (deg < 3) return 9;
    b[9] = SH_C3[0] * y * (3.f * xx - yy); b[10] = SH_C3[1] * xy * z;
    b[11] = SH_C3[2] * y * (4.f * zz - xx - yy); b[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    b[13] = SH_C3[4] * x * (4.f * zz - xx - yy); b[14] = SH_C3[5] * z * (xx - yy);
    b[15] = SH_C3[6] * x * (xx - 3.f * yy);
    return 16;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static void rect_of(float px, float py, int radius, int gx, int gy, int* r) {   /* auxiliary.h:46-56 */
    r[0] = clampi((int)((px - radius) / TILE), 0, gx);
    r[1] = clampi((int)((py - radius) / TILE), 0, gy);
    r[2] = clampi((int)((px + radius + TILE - 1) / TILE), 0, gx);
    r[3] = clampi((int)((py + radius + TILE - 1) / TILE), 0, gy);
}

typedef struct { uint64_t key; uint32_t val; uint32_t seq; } KV;
static int kv_cmp(const void* a, const void* b) {
    const KV* x = (const KV*)a; const KV* y = (const KV*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);   /* stable */
}

/* Forward.  Returns num_rendered (>=0) or -1 on allocation failure. */
int64_t oracle_forward(void* h, int P, int D, int M, const float* bg, int W, int H,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* opacities, const float* scales, float scale_modifier,
                       const float* rotations, const float* cov3D_precomp, const float* view,
                       const float* proj, const float* campos, float tan_fovx, float tan_fovy,
                       float* out_color, float* out_depth, int* out_radii) {
    Oracle* o = (Oracle*)h;
    o_free_arrays(o);
    o->P = P; o->D = D; o->M = M; o->W = W; o->H = H;
    o->gx = (W + TILE - 1) / TILE; o->gy = (H + TILE - 1) / TILE;
    const int ntiles = o->gx * o->gy;
    const size_t HW = (size_t)W * H;
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    size_t Pn = P > 0 ? (size_t)P : 1;
    o->depth = calloc(Pn, 4); o->xy = calloc(Pn * 2, 4); o->conic_o = calloc(Pn * 4, 4);
    o->rgb = calloc(Pn * 3, 4); o->cov3d = calloc(Pn * 6, 4); o->clamped = calloc(Pn * 3, 1);
    o->radii = calloc(Pn, 4); o->tiles = calloc(Pn, 4);
    o->ranges = calloc((size_t)ntiles * 2, 4);
    o->final_T = calloc(HW, 4); o->n_contrib = calloc(HW, 4);

    /* ---- stage 1: per-Gaussian projection ------------------------------- */
    int64_t R = 0;
    for (int i = 0; i < P; ++i) {
        const float* p = means3D + 3 * (size_t)i;
        float pv[3];
        xf3(view, p, pv);
        if (pv[2] <= 0.2f) continue;                                  /* near cull, auxiliary.h:154 */
        float ph[4];
        xf4(proj, p, ph);
        float pw = 1.0f / (ph[3] + 0.0000001f);
        float ndc_x = ph[0] * pw, ndc_y = ph[1] * pw;
        float* c6 = o->cov3d + 6 * (size_t)i;
        if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * (size_t)i, 24);
        else cov3d_of(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i, c6);
        float cov[3], A[2][3];
        cov2d_of(p, focal_x, focal_y, tan_fovx, tan_fovy, c6, view, cov, A, NULL, NULL, NULL);
        float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
        float l1 = mid + sq, l2 = mid - sq;
        float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
        /* pixel coordinate in double, then rounded (auxiliary.h:43) */
        float px = (float)((((double)ndc_x + 1.0) * W - 1.0) * 0.5);
        float py = (float)((((double)ndc_y + 1.0) * H - 1.0) * 0.5);
        int r[4];
        rect_of(px, py, (int)my_radius, o->gx, o->gy, r);
        if ((r[2] - r[0]) * (r[3] - r[1]) == 0) continue;
        if (!colors_precomp) {
            float d[3] = {p[0] - campos[0], p[1] - campos[1], p[2] - campos[2]};
            float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            d[0] /= len; d[1] /= len; d[2] /= len;
            float b[16];
            int nb = sh_basis(D, d, b);
            const float* sh = shs + (size_t)i * M * 3;
            for (int c = 0; c < 3; ++c) {
                float acc = 0.f;
                for (int k = 0; k < nb; ++k) acc += b[k] * sh[k * 3 + c];
                acc += 0.5f;
                o->clamped[3 * (size_t)i + c] = acc < 0;
                o->rgb[3 * (size_t)i + c] = fmaxf(acc, 0.f);
            }
        } else {
            for (int c = 0; c < 3; ++c) o->rgb[3 * (size_t)i + c] = colors_precomp[3 * (size_t)i + c];
        }
        o->depth[i] = pv[2];
        o->radii[i] = (int)my_radius;
        o->xy[2 * (size_t)i] = px; o->xy[2 * (size_t)i + 1] = py;
        o->conic_o[4 * (size_t)i + 0] = conic[0]; o->conic_o[4 * (size_t)i + 1] = conic[1];
        o->conic_o[4 * (size_t)i + 2] = conic[2]; o->conic_o[4 * (size_t)i + 3] = opacities[i];
        o->tiles[i] = (uint32_t)((r[3] - r[1]) * (r[2] - r[0]));
        R += o->tiles[i];
    }
    if (out_radii) memcpy(out_radii, o->radii, (size_t)P * 4);
    o->R = R;

    /* ---- stage 2: duplicate, sort by (tile, depth bits), ranges ---------- */
    size_t Rn = R > 0 ? (size_t)R : 1;
    KV* kv = malloc(Rn * sizeof(KV));
    o->keys = malloc(Rn * 8); o->plist = malloc(Rn * 4);
    if (!kv || !o->keys || !o->plist) { free(kv); return -1; }
    {
        size_t off = 0;
        for (int i = 0; i < P; ++i) {
            if (o->radii[i] <= 0) continue;
            int r[4];
            rect_of(o->xy[2 * (size_t)i], o->xy[2 * (size_t)i + 1], o->radii[i], o->gx, o->gy, r);
            uint32_t dbits;
            memcpy(&dbits, &o->depth[i], 4);
            for (int y = r[1]; y < r[3]; ++y)
                for (int x = r[0]; x < r[2]; ++x) {
                    kv[off].key = ((uint64_t)(y * o->gx + x) << 32) | dbits;
                    kv[off].val = (uint32_t)i;
                    kv[off].seq = (uint32_t)off;
                    ++off;
                }
        }
    }
    qsort(kv, (size_t)R, sizeof(KV), kv_cmp);
    for (int64_t k = 0; k < R; ++k) { o->keys[k] = kv[k].key; o->plist[k] = kv[k].val; }
    free(kv);
    for (int64_t k = 0; k < R; ++k) {                                 /* rasterizer_impl.cu:116-138 */
        uint32_t cur = (uint32_t)(o->keys[k] >> 32);
        if (k == 0) o->ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(o->keys[k - 1] >> 32);
            if (cur != prev) { o->ranges[2 * prev + 1] = (uint32_t)k; o->ranges[2 * cur] = (uint32_t)k; }
        }
        if (k == R - 1) o->ranges[2 * cur + 1] = (uint32_t)R;
    }

    /* ---- stage 3: per-pixel front-to-back blend (forward.cu:308-378) ------ */
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            int tile = (py / TILE) * o->gx + (px / TILE);
            uint32_t s = o->ranges[2 * tile], e = o->ranges[2 * tile + 1];
            float T = 1.0f, C[NCH] = {0, 0, 0}, Dp = 0.f;
            uint32_t contributor = 0, last = 0;
            for (uint32_t k = s; k < e; ++k) {
                ++contributor;
                uint32_t g = o->plist[k];
                float dx = o->xy[2 * (size_t)g] - (float)px, dy = o->xy[2 * (size_t)g + 1] - (float)py;
                const float* co = o->conic_o + 4 * (size_t)g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float alpha = fminf(0.99f, co[3] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) break;
                for (int c = 0; c < NCH; ++c) C[c] += o->rgb[3 * (size_t)g + c] * alpha * T;
                Dp += o->depth[g] * alpha * T;
                T = test_T;
                last = contributor;
            }
            size_t pid = (size_t)py * W + px;
            o->final_T[pid] = T;
            o->n_contrib[pid] = last;
            for (int c = 0; c < NCH; ++c) out_color[c * HW + pid] = C[c] + T * bg[c];
            out_depth[pid] = Dp;                                      /* no background term, :377 */
        }
    return R;
}

/* accessors: copy state out (sizes are the caller's responsibility) */
int64_t oracle_num_rendered(void* h) { return ((Oracle*)h)->R; }
void oracle_get_geometry(void* h, float* depth, float* xy, float* conic_o, float* rgb, float* cov3d,
                         uint8_t* clamped, uint32_t* tiles) {
    Oracle* o = (Oracle*)h;
    size_t P = (size_t)o->P;
    if (depth) memcpy(depth, o->depth, P * 4);
    if (xy) memcpy(xy, o->xy, P * 8);
    if (conic_o) memcpy(conic_o, o->conic_o, P * 16);
    if (rgb) memcpy(rgb, o->rgb, P * 12);
    if (cov3d) memcpy(cov3d, o->cov3d, P * 24);
    if (clamped) memcpy(clamped, o->clamped, P * 3);
    if (tiles) memcpy(tiles, o->tiles, P * 4);
}
void oracle_get_binning(void* h, uint64_t* keys, uint32_t* plist, uint32_t* ranges) {
    Oracle* o = (Oracle*)h;
    if (keys) memcpy(keys, o->keys, (size_t)o->R * 8);
    if (plist) memcpy(plist, o->plist, (size_t)o->R * 4);
    if (ranges) memcpy(ranges, o->ranges, (size_t)o->gx * o->gy * 8);
}
void oracle_get_image(void* h, float* final_T, uint32_t* n_contrib) {
    Oracle* o = (Oracle*)h;
    size_t HW = (size_t)o->W * o->H;
    if (final_T) memcpy(final_T, o->final_T, HW * 4);
    if (n_contrib) memcpy(n_contrib, o->n_contrib, HW * 4);
}

/* Backward.  Accumulates in DOUBLE (the reference sums with unordered fp32
 * atomics, backward.cu:550-587, so its own result is only defined up to
 * summation order; the per-pair terms below are single precision like the
 * reference's).  All outputs are [P,*] and fully written. */
int oracle_backward(void* h, const float* bg, const float* means3D, const float* shs,
                    const float* colors_precomp, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp, const float* view,
                    const float* proj, const float* campos, float tan_fovx, float tan_fovy,
                    const float* dL_dpix, const float* dL_dpix_depth, float* dL_dmean2D /*[P,3]*/,
                    float* dL_dconic /*[P,4]*/, float* dL_dopacity /*[P]*/, float* dL_dcolor /*[P,3]*/,
                    float* dL_ddepth /*[P]*/, float* dL_dmean3D /*[P,3]*/, float* dL_dcov3D /*[P,6]*/,
                    float* dL_dsh /*[P,M,3]*/, float* dL_dscale /*[P,3]*/, float* dL_drot /*[P,4]*/) {
    Oracle* o = (Oracle*)h;
    const int P = o->P, W = o->W, H = o->H, D = o->D, M = o->M;
    const size_t HW = (size_t)W * H;
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    double* acc = calloc((size_t)(P > 0 ? P : 1) * 10, sizeof(double));
    if (!acc) return -1;
    /* ---- composite backward (backward.cu:488-589) ------------------------ */
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t pid = (size_t)py * W + px;
            int tile = (py / TILE) * o->gx + (px / TILE);
            uint32_t s = o->ranges[2 * tile];
            const float T_final = o->final_T[pid];
            float T = T_final;
            int last = (int)o->n_contrib[pid];
            float g[NCH], gd = dL_dpix_depth[pid];
            for (int c = 0; c < NCH; ++c) g[c] = dL_dpix[c * HW + pid];
            float bgdot = 0.f;
            for (int c = 0; c < NCH; ++c) bgdot += bg[c] * g[c];
            float accum[NCH] = {0, 0, 0}, accum_d = 0.f, last_alpha = 0.f, last_c[NCH] = {0, 0, 0}, last_d = 0.f;
            for (int k = last - 1; k >= 0; --k) {
                uint32_t gi = o->plist[s + (uint32_t)k];
                float dx = o->xy[2 * (size_t)gi] - (float)px, dy = o->xy[2 * (size_t)gi + 1] - (float)py;
                const float* co = o->conic_o + 4 * (size_t)gi;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float G = expf(power);
                float alpha = fminf(0.99f, co[3] * G);
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.f - alpha);
                float w = alpha * T;
                float dL_dalpha = 0.f;
                double* a = acc + (size_t)gi * 10;
                for (int c = 0; c < NCH; ++c) {
                    float col = o->rgb[3 * (size_t)gi + c];
                    accum[c] = last_alpha * last_c[c] + (1.f - last_alpha) * accum[c];
                    last_c[c] = col;
                    dL_dalpha += (col - accum[c]) * g[c];
                    a[6 + c] += w * g[c];
                }
                float dep = o->depth[gi];
                accum_d = last_alpha * last_d + (1.f - last_alpha) * accum_d;
                last_d = dep;
                dL_dalpha += (dep - accum_d) * gd;
                a[9] += w * gd;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
                float dL_dG = co[3] * dL_dalpha;
                float gdx = G * dx, gdy = G * dy;
                float dG_ddelx = -gdx * co[0] - gdy * co[1];
                float dG_ddely = -gdy * co[2] - gdx * co[1];
                a[0] += dL_dG * dG_ddelx * ddelx_dx;
                a[1] += dL_dG * dG_ddely * ddely_dy;
                a[2] += -0.5f * gdx * dx * dL_dG;
                a[3] += -0.5f * gdx * dy * dL_dG;
                a[4] += -0.5f * gdy * dy * dL_dG;
                a[5] += G * dL_dalpha;
            }
        }
    /* ---- per-Gaussian chain rule ----------------------------------------- */
    for (int i = 0; i < P; ++i) {
        const double* a = acc + (size_t)i * 10;
        float gm2[2] = {(float)a[0], (float)a[1]};
        float gcon[3] = {(float)a[2], (float)a[3], (float)a[4]};
        float gcol[3] = {(float)a[6], (float)a[7], (float)a[8]};
        float gdep = (float)a[9];
        dL_dmean2D[3 * (size_t)i] = gm2[0]; dL_dmean2D[3 * (size_t)i + 1] = gm2[1]; dL_dmean2D[3 * (size_t)i + 2] = 0.f;
        if (dL_dconic) { dL_dconic[4 * (size_t)i] = gcon[0]; dL_dconic[4 * (size_t)i + 1] = gcon[1];
                         dL_dconic[4 * (size_t)i + 2] = 0.f; dL_dconic[4 * (size_t)i + 3] = gcon[2]; }
        dL_dopacity[i] = (float)a[5];
        for (int c = 0; c < 3; ++c) dL_dcolor[3 * (size_t)i + c] = gcol[c];
        if (dL_ddepth) dL_ddepth[i] = gdep;
        float dmean[3] = {0, 0, 0}, dcov[6] = {0, 0, 0, 0, 0, 0}, dsc[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0};
        if (dL_dsh) memset(dL_dsh + (size_t)i * M * 3, 0, (size_t)M * 12);
        if (o->radii[i] > 0) {
            const float* p = means3D + 3 * (size_t)i;
            const float* c6 = cov3D_precomp ? cov3D_precomp + 6 * (size_t)i : o->cov3d + 6 * (size_t)i;
            /* conic -> cov2D (backward.cu:196-212) */
            float cov[3], A[2][3], t[3], txtz, tytz;
            cov2d_of(p, focal_x, focal_y, tan_fovx, tan_fovy, c6, view, cov, A, t, &txtz, &tytz);
            float ca = cov[0], cb = cov[1], cc = cov[2];
            float denom = ca * cc - cb * cb;
            float d2i = 1.0f / ((denom * denom) + 0.0000001f);
            float da = 0, db = 0, dc = 0;
            if (d2i != 0) {
                da = d2i * (-cc * cc * gcon[0] + 2 * cb * cc * gcon[1] + (denom - ca * cc) * gcon[2]);
                dc = d2i * (-ca * ca * gcon[2] + 2 * ca * cb * gcon[1] + (denom - ca * cc) * gcon[0]);
                db = d2i * 2 * (cb * cc * gcon[0] - (denom + 2 * cb * cb) * gcon[1] + ca * cb * gcon[2]);
                /* cov2D = A Sigma A^T  =>  dL/dSigma_jk (off-diagonals stored once, doubled; :217-227) */
                const int jj[6] = {0, 0, 0, 1, 1, 2}, kk[6] = {0, 1, 2, 1, 2, 2};
                for (int e = 0; e < 6; ++e) {
                    int j = jj[e], k = kk[e];
                    float v = A[0][j] * A[0][k] * da + A[1][j] * A[1][k] * dc;
                    float x = (A[0][j] * A[1][k] + A[0][k] * A[1][j]) * db;
                    dcov[e] = (j == k) ? (v + 0.5f * x) : (2 * v + x);
                }
            }
            /* dL/dA (backward.cu:237-248): rows of A */
            float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
            float dA[2][3];
            for (int j = 0; j < 3; ++j) {
                float sa0 = A[0][0] * S[0][j] + A[0][1] * S[1][j] + A[0][2] * S[2][j];
                float sa1 = A[1][0] * S[0][j] + A[1][1] * S[1][j] + A[1][2] * S[2][j];
                dA[0][j] = 2 * sa0 * da + sa1 * db;
                dA[1][j] = 2 * sa1 * dc + sa0 * db;
            }
            /* A = J Wv => dL/dJ (only the 4 non-zero entries, :252-255) */
            float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
            for (int c = 0; c < 3; ++c) {
                float w0 = view[c * 4 + 0], w1 = view[c * 4 + 1], w2 = view[c * 4 + 2];
                dJ00 += w0 * dA[0][c]; dJ02 += w2 * dA[0][c];
                dJ11 += w1 * dA[1][c]; dJ12 += w2 * dA[1][c];
            }
            float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
            float xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
            float ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
            float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
            float dtx = xmul * -focal_x * tz2 * dJ02;
            float dty = ymul * -focal_y * tz2 * dJ12;
            float dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 + (2 * focal_x * t[0]) * tz3 * dJ02 +
                        (2 * focal_y * t[1]) * tz3 * dJ12;
            for (int c = 0; c < 3; ++c)
                dmean[c] = view[c * 4 + 0] * dtx + view[c * 4 + 1] * dty + view[c * 4 + 2] * dtz;
            /* screen-space mean (backward.cu:375-389) */
            float mh[4];
            xf4(proj, p, mh);
            float mw = 1.0f / (mh[3] + 0.0000001f);
            float mul1 = mh[0] * mw * mw, mul2 = mh[1] * mw * mw;
            for (int c = 0; c < 3; ++c)
                dmean[c] += (proj[c * 4 + 0] * mw - proj[c * 4 + 3] * mul1) * gm2[0] +
                            (proj[c * 4 + 1] * mw - proj[c * 4 + 3] * mul2) * gm2[1];
            /* depth (backward.cu:397-403) */
            float mul3 = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
            for (int c = 0; c < 3; ++c) dmean[c] += (view[c * 4 + 2] - view[c * 4 + 3] * mul3) * gdep;
            /* SH (backward.cu:20-139) */
            if (!colors_precomp) {
                const float* sh = shs + (size_t)i * M * 3;
                float* dsh = dL_dsh + (size_t)i * M * 3;
                float dv[3] = {p[0] - campos[0], p[1] - campos[1], p[2] - campos[2]};
                float len = sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
                float d[3] = {dv[0] / len, dv[1] / len, dv[2] / len};
                float b[16];
                int nb = sh_basis(D, d, b);
                float gr[3];
                for (int c = 0; c < 3; ++c) gr[c] = o->clamped[3 * (size_t)i + c] ? 0.f : gcol[c];
                for (int k = 0; k < nb; ++k)
                    for (int c = 0; c < 3; ++c) dsh[k * 3 + c] = b[k] * gr[c];
                /* d(colour)/d(dir) by differentiating the basis polynomials */
                float ddir[3] = {0, 0, 0};
                float x = d[0], y = d[1], z = d[2];
                for (int c = 0; c < 3; ++c) {
                    float dbx[16] = {0}, dby[16] = {0}, dbz[16] = {0};
                    if (D > 0) { dby[1] = -SH_C1; dbz[2] = SH_C1; dbx[3] = -SH_C1; }
                    if (D > 1) {
                        dbx[4] = SH_C2[0] * y; dby[4] = SH_C2[0] * x;
                        dby[5] = SH_C2[1] * z; dbz[5] = SH_C2[1] * y;
                        dbx[6] = SH_C2[2] * -2.f * x; dby[6] = SH_C2[2] * -2.f * y; dbz[6] = SH_C2[2] * 4.f * z;
                        dbx[7] = SH_C2[3] * z; dbz[7] = SH_C2[3] * x;
                        dbx[8] = SH_C2[4] * 2.f * x; dby[8] = SH_C2[4] * -2.f * y;
                    }
                    if (D > 2) {
                        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        dbx[9] = SH_C3[0] * 6.f * xy;  dby[9] = SH_C3[0] * 3.f * (xx - yy);
                        dbx[10] = SH_C3[1] * yz; dby[10] = SH_C3[1] * xz; dbz[10] = SH_C3[1] * xy;
                        dbx[11] = SH_C3[2] * -2.f * xy; dby[11] = SH_C3[2] * (-3.f * yy + 4.f * zz - xx); dbz[11] = SH_C3[2] * 8.f * yz;
                        dbx[12] = SH_C3[3] * -6.f * xz; dby[12] = SH_C3[3] * -6.f * yz; dbz[12] = SH_C3[3] * 3.f * (2.f * zz - xx - yy);
                        dbx[13] = SH_C3[4] * (-3.f * xx + 4.f * zz - yy); dby[13] = SH_C3[4] * -2.f * xy; dbz[13] = SH_C3[4] * 8.f * xz;
                        dbx[14] = SH_C3[5] * 2.f * xz; dby[14] = SH_C3[5] * -2.f * yz; dbz[14] = SH_C3[5] * (xx - yy);
                        dbx[15] = SH_C3[6] * 3.f * (xx - yy); dby[15] = SH_C3[6] * -6.f * xy;
                    }
                    float sx = 0, sy = 0, sz = 0;
                    for (int k = 1; k < nb; ++k) {
                        sx += dbx[k] * sh[k * 3 + c]; sy += dby[k] * sh[k * 3 + c]; sz += dbz[k] * sh[k * 3 + c];
                    }
                    ddir[0] += sx * gr[c]; ddir[1] += sy * gr[c]; ddir[2] += sz * gr[c];
                }
                /* through the normalisation (auxiliary.h:107-117) */
                float s2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
                float inv32 = 1.0f / sqrtf(s2 * s2 * s2);
                float dot = dv[0] * ddir[0] + dv[1] * ddir[1] + dv[2] * ddir[2];
                for (int c = 0; c < 3; ++c) dmean[c] += (s2 * ddir[c] - dv[c] * dot) * inv32;
            }
            /* Sigma -> scale, quaternion (backward.cu:278-341) */
            if (!cov3D_precomp) {
                const float* q = rotations + 4 * (size_t)i;
                const float* sc = scales + 3 * (size_t)i;
                float Rm[3][3];
                rot_from_quat(q, Rm);
                float s[3] = {scale_modifier * sc[0], scale_modifier * sc[1], scale_modifier * sc[2]};
                /* symmetric dL/dSigma with halved off-diagonals */
                float G[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                 {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                 {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
                /* Sigma = L L^T with L[i][k] = Rm[i][k] s_k ; dL/dL = 2 G L */
                float L[3][3], dLm[3][3];
                for (int a_ = 0; a_ < 3; ++a_)
                    for (int k = 0; k < 3; ++k) L[a_][k] = Rm[a_][k] * s[k];
                for (int a_ = 0; a_ < 3; ++a_)
                    for (int k = 0; k < 3; ++k)
                        dLm[a_][k] = 2.f * (G[a_][0] * L[0][k] + G[a_][1] * L[1][k] + G[a_][2] * L[2][k]);
                float dR[3][3];
                for (int k = 0; k < 3; ++k) {
                    dsc[k] = Rm[0][k] * dLm[0][k] + Rm[1][k] * dLm[1][k] + Rm[2][k] * dLm[2][k];
                    for (int a_ = 0; a_ < 3; ++a_) dR[a_][k] = dLm[a_][k] * s[k];
                }
                float r = q[0], x = q[1], y = q[2], z = q[3];
                /* derivative of rot_from_quat, un-normalised quaternion (backward.cu:333-336) */
                dq[0] = 2 * z * (dR[1][0] - dR[0][1]) + 2 * y * (dR[0][2] - dR[2][0]) + 2 * x * (dR[2][1] - dR[1][2]);
                dq[1] = 2 * y * (dR[0][1] + dR[1][0]) + 2 * z * (dR[0][2] + dR[2][0]) + 2 * r * (dR[2][1] - dR[1][2]) - 4 * x * (dR[2][2] + dR[1][1]);
                dq[2] = 2 * x * (dR[0][1] + dR[1][0]) + 2 * r * (dR[0][2] - dR[2][0]) + 2 * z * (dR[2][1] + dR[1][2]) - 4 * y * (dR[2][2] + dR[0][0]);
                dq[3] = 2 * r * (dR[1][0] - dR[0][1]) + 2 * x * (dR[0][2] + dR[2][0]) + 2 * y * (dR[2][1] + dR[1][2]) - 4 * z * (dR[1][1] + dR[0][0]);
            }
        }
        for (int c = 0; c < 3; ++c) dL_dmean3D[3 * (size_t)i + c] = dmean[c];
        for (int e = 0; e < 6; ++e) dL_dcov3D[6 * (size_t)i + e] = dcov[e];
        for (int c = 0; c < 3; ++c) dL_dscale[3 * (size_t)i + c] = dsc[c];
        for (int c = 0; c < 4; ++c) dL_drot[4 * (size_t)i + c] = dq[c];
    }
    free(acc);
    return 0;
}

/* near-plane visibility (rasterizer_impl.cu:54-66) */
void oracle_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int i = 0; i < P; ++i) {
        float pv[3];
        xf3(view, means3D + 3 * (size_t)i, pv);
        present[i] = pv[2] > 0.2f;
    }
}
