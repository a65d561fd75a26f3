/*
 * oracle.c -- CPU restatement of the MojoSplat forward path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (mojosplat_b200/) never imports, links or executes anything under oracle/.
 *
 * Each function restates one reference routine in plain fp32 C and cites the
 * reference lines it follows (paths relative to the reference checkout):
 *
 *   oracle_project        mojosplat/projection.py:51-346   (torch backend, "torch" semantics)
 *                         mojosplat/kernels/projection.mojo:13-257 ("gsplat" semantics)
 *   oracle_bin_*          mojosplat/binning.py:108-262     (torch backend, bit-exact target)
 *   oracle_rasterize      mojosplat/kernels/rasterization.mojo:75-162
 *
 * Pinning status:
 *   - projection + binning: PINNED against outputs of the unmodified reference torch
 *     backend run in the build container (oracle/make_golden.py -> tests/golden/*.npz).
 *   - rasterization: the reference has no CPU rasterizer and gsplat is not installable
 *     offline, so gsplat parity is UNPINNED; the restatement is pinned only to the
 *     backend-independent known-answer checks of the reference's tests
 *     (tests/test_rasterization.py:154-248, tests/test_render.py:60-119) and to an
 *     independent numpy restatement (oracle/oracle_np.py).
 *
 * Build: oracle/build.sh  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off keeps every multiply and add separately rounded, like the eager
 * torch ops the reference issues.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_SEM_TORCH 0
#define ORACLE_SEM_GSPLAT 1

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------
 * Projection.  projection.py:285-346 (wrapper), :72-102 (covariance), :51-69 (quat->R),
 * :163-196 (world->cam), :105-160 (pinhole), :199-283 (fused tail).
 *
 * semantics == ORACLE_SEM_TORCH : exactly the torch backend (opacity ignored, radius
 *   3.33 sigma, culled rows keep their computed means2d/conics/depths, det clamp 1e-10).
 * semantics == ORACLE_SEM_GSPLAT: the Mojo kernel's / gsplat's rules
 *   (projection.mojo:59-87 near + opacity cull with zeroed rows, :213-226 opacity-aware
 *   extent, :240-244 viewport cull, :246-251 conic without det clamp).  `near` comes from
 *   the caller (the Mojo kernel hard-codes 0.1, the gsplat call passes camera.near).
 * ---------------------------------------------------------------------------------- */
void oracle_project(
    int64_t N,
    const float* means3d,   /* [N,3] */
    const float* log_scales,/* [N,3] */
    const float* quats,     /* [N,4] wxyz */
    const float* opacities, /* [N] or NULL (only read in gsplat semantics) */
    const float* viewmat,   /* [4,4] row-major world->cam */
    float fx, float fy, float cx, float cy,
    int W, int H, float near_plane, float far_plane, float eps2d,
    int semantics,
    float* means2d,  /* [N,2] */
    float* conics,   /* [N,3] */
    float* depths,   /* [N] */
    int32_t* radii)  /* [N,2] */
{
    const float Rv[3][3] = {
        {viewmat[0], viewmat[1], viewmat[2]},
        {viewmat[4], viewmat[5], viewmat[6]},
        {viewmat[8], viewmat[9], viewmat[10]}};
    const float tv[3] = {viewmat[3], viewmat[7], viewmat[11]};

    /* projection.py:137-146 */
    const float tan_fovx = 0.5f * (float)W / fx;
    const float tan_fovy = 0.5f * (float)H / fy;
    const float lim_x_pos = ((float)W - cx) / fx + 0.3f * tan_fovx;
    const float lim_x_neg = cx / fx + 0.3f * tan_fovx;
    const float lim_y_pos = ((float)H - cy) / fy + 0.3f * tan_fovy;
    const float lim_y_neg = cy / fy + 0.3f * tan_fovy;

#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float* mu = means3d + 3 * i;
        const float* ls = log_scales + 3 * i;
        const float* q = quats + 4 * i;

        /* world -> camera mean, projection.py:190-192 */
        float mc[3];
        for (int r = 0; r < 3; ++r)
            mc[r] = (Rv[r][0] * mu[0] + Rv[r][1] * mu[1] + Rv[r][2] * mu[2]) + tv[r];

        if (semantics == ORACLE_SEM_GSPLAT) {
            /* projection.mojo:59-87 */
            float op = opacities ? opacities[i] : 1.0f;
            if (mc[2] <= near_plane || mc[2] >= far_plane || op < (1.0f / 255.0f)) {
                means2d[2 * i] = means2d[2 * i + 1] = 0.0f;
                conics[3 * i] = conics[3 * i + 1] = conics[3 * i + 2] = 0.0f;
                depths[i] = 0.0f;
                radii[2 * i] = radii[2 * i + 1] = 0;
                continue;
            }
        }

        /* quat -> rotation, projection.py:51-69 (F.normalize: x / max(||x||, 1e-12)) */
        float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (nrm < 1e-12f) nrm = 1e-12f;
        const float w = q[0] / nrm, x = q[1] / nrm, y = q[2] / nrm, z = q[3] / nrm;
        float R[3][3] = {
            {1.0f - 2.0f * (y * y + z * z), 2.0f * (x * y - w * z), 2.0f * (x * z + w * y)},
            {2.0f * (x * y + w * z), 1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z - w * x)},
            {2.0f * (x * z - w * y), 2.0f * (y * z + w * x), 1.0f - 2.0f * (x * x + y * y)}};

        /* M = R * s ; Sigma = M M^T, projection.py:86-87 */
        const float s[3] = {expf(ls[0]), expf(ls[1]), expf(ls[2])};
        float M[3][3], S[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[r][c] = R[r][c] * s[c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                S[r][c] = M[r][0] * M[c][0] + M[r][1] * M[c][1] + M[r][2] * M[c][2];

        /* Sigma_c = Rv Sigma Rv^T, projection.py:193-195 */
        float A[3][3], Sc[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                A[r][c] = Rv[r][0] * S[0][c] + Rv[r][1] * S[1][c] + Rv[r][2] * S[2][c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                Sc[r][c] = A[r][0] * Rv[c][0] + A[r][1] * Rv[c][1] + A[r][2] * Rv[c][2];

        /* pinhole Jacobian, projection.py:134-159 */
        const float tz = mc[2];
        const float tz2 = tz * tz;
        float rx = mc[0] / tz, ry = mc[1] / tz;
        rx = fminf(fmaxf(rx, -lim_x_neg), lim_x_pos);
        ry = fminf(fmaxf(ry, -lim_y_neg), lim_y_pos);
        const float tx = tz * rx, ty = tz * ry;
        const float J[2][3] = {{fx / tz, 0.0f, -fx * tx / tz2}, {0.0f, fy / tz, -fy * ty / tz2}};
        float JS[2][3];
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                JS[r][c] = J[r][0] * Sc[0][c] + J[r][1] * Sc[1][c] + J[r][2] * Sc[2][c];
        float c00 = JS[0][0] * J[0][0] + JS[0][1] * J[0][1] + JS[0][2] * J[0][2];
        float c01 = JS[0][0] * J[1][0] + JS[0][1] * J[1][1] + JS[0][2] * J[1][2];
        float c10 = JS[1][0] * J[0][0] + JS[1][1] * J[0][1] + JS[1][2] * J[0][2];
        float c11 = JS[1][0] * J[1][0] + JS[1][1] * J[1][1] + JS[1][2] * J[1][2];

        /* means2d = (K[:2,:3] . mu_c) / z, projection.py:156-159 */
        const float m2x = (fx * mc[0] + 0.0f * mc[1] + cx * mc[2]) / tz;
        const float m2y = (0.0f * mc[0] + fy * mc[1] + cy * mc[2]) / tz;

        c00 += eps2d; /* projection.py:242 */
        c11 += eps2d;
        float det = c00 * c11 - c01 * c10;

        if (semantics == ORACLE_SEM_TORCH) {
            if (!(det >= 1e-10f)) det = (det != det) ? det : 1e-10f; /* clamp(min=1e-10), :248 */
            const float k0 = c11 / det;
            const float k1 = -(c01 + c10) / 2.0f / det;
            const float k2 = c00 / det;
            float r_x = ceilf(3.33f * sqrtf(c00)); /* :266-267 */
            float r_y = ceilf(3.33f * sqrtf(c11));
            const int valid = (det > 0.0f) && (tz > near_plane) && (tz < far_plane); /* :271 */
            if (!valid) { r_x = 0.0f; r_y = 0.0f; }
            const int inside = (m2x + r_x > 0.0f) && (m2x - r_x < (float)W) &&
                               (m2y + r_y > 0.0f) && (m2y - r_y < (float)H); /* :274-279 */
            if (!inside) { r_x = 0.0f; r_y = 0.0f; }
            means2d[2 * i] = m2x; means2d[2 * i + 1] = m2y;
            conics[3 * i] = k0; conics[3 * i + 1] = k1; conics[3 * i + 2] = k2;
            depths[i] = tz;
            radii[2 * i] = (int32_t)r_x; radii[2 * i + 1] = (int32_t)r_y;
        } else {
            /* projection.mojo:213-257 */
            float extend = 3.33f;
            const float op = opacities ? opacities[i] : 1.0f;
            const float oe = sqrtf(2.0f * logf(op / (1.0f / 255.0f)));
            if (oe < extend) extend = oe;
            const float r_x = ceilf(extend * sqrtf(c00));
            const float r_y = ceilf(extend * sqrtf(c11));
            if ((r_x <= 0.0f && r_y <= 0.0f) ||
                m2x + r_x <= 0.0f || m2x - r_x >= (float)W ||
                m2y + r_y <= 0.0f || m2y - r_y >= (float)H) {
                means2d[2 * i] = means2d[2 * i + 1] = 0.0f;
                conics[3 * i] = conics[3 * i + 1] = conics[3 * i + 2] = 0.0f;
                depths[i] = 0.0f;
                radii[2 * i] = radii[2 * i + 1] = 0;
                continue;
            }
            const float inv_det = 1.0f / det;
            means2d[2 * i] = m2x; means2d[2 * i + 1] = m2y;
            conics[3 * i] = c11 * inv_det;
            conics[3 * i + 1] = -(c01 + c10) / 2.0f * inv_det;
            conics[3 * i + 2] = c00 * inv_det;
            depths[i] = tz;
            radii[2 * i] = (int32_t)r_x; radii[2 * i + 1] = (int32_t)r_y;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Binning.  binning.py:108-262.
 * ---------------------------------------------------------------------------------- */

/* Monotone float -> uint32 map used as the canonical depth order:
 * ascending float order, -0.0 == +0.0 (torch.argsort ties them), NaN last. */
static inline uint32_t depth_key(float d) {
    uint32_t b;
    memcpy(&b, &d, 4);
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu; /* NaN sorts last */
    if (b == 0x80000000u) b = 0u;                             /* -0.0 -> +0.0 */
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

static inline float clampf(float v, float lo, float hi) {
    /* torch.clamp: min(max(v, lo), hi); NaN propagates */
    if (v != v) return v;
    v = v < lo ? lo : v;
    v = v > hi ? hi : v;
    return v;
}

static inline void tile_rect(
    const float* means2d, const float* radii_f, int64_t i, int W, int H, int tile_size,
    int tiles_w, int tiles_h, int semantics, int* x0, int* y0, int* x1, int* y1)
{
    const float mx = means2d[2 * i], my = means2d[2 * i + 1];
    const float rx = radii_f[2 * i], ry = radii_f[2 * i + 1];
    if (semantics == ORACLE_SEM_TORCH) {
        /* binning.py:139-155: pixel-space clamp, trunc-to-int32, INCLUSIVE max */
        const float ts = (float)tile_size;
        float ax = clampf(mx - rx, 0.0f, (float)(W - 1));
        float bx = clampf(mx + rx, 0.0f, (float)(W - 1));
        float ay = clampf(my - ry, 0.0f, (float)(H - 1));
        float by = clampf(my + ry, 0.0f, (float)(H - 1));
        int tx0 = (int)(ax / ts), tx1 = (int)(bx / ts);
        int ty0 = (int)(ay / ts), ty1 = (int)(by / ts);
        /* binning.py:181-184 second clamp in tile space (no-op for finite inputs) */
        tx0 = tx0 < 0 ? 0 : (tx0 > tiles_w - 1 ? tiles_w - 1 : tx0);
        tx1 = tx1 < 0 ? 0 : (tx1 > tiles_w - 1 ? tiles_w - 1 : tx1);
        ty0 = ty0 < 0 ? 0 : (ty0 > tiles_h - 1 ? tiles_h - 1 : ty0);
        ty1 = ty1 < 0 ? 0 : (ty1 > tiles_h - 1 ? tiles_h - 1 : ty1);
        *x0 = tx0; *x1 = tx1 + 1; *y0 = ty0; *y1 = ty1 + 1; /* exclusive ends */
    } else {
        /* gsplat isect_tiles [upstream-knowledge, SURVEY 8a-2]: tile-space floor/ceil,
         * exclusive max, nothing for radii <= 0 */
        if (rx <= 0.0f || ry <= 0.0f) { *x0 = *x1 = *y0 = *y1 = 0; return; }
        const float ts = (float)tile_size;
        float fx0 = floorf((mx - rx) / ts), fx1 = ceilf((mx + rx) / ts);
        float fy0 = floorf((my - ry) / ts), fy1 = ceilf((my + ry) / ts);
        fx0 = fminf(fmaxf(fx0, 0.0f), (float)tiles_w); fx1 = fminf(fmaxf(fx1, 0.0f), (float)tiles_w);
        fy0 = fminf(fmaxf(fy0, 0.0f), (float)tiles_h); fy1 = fminf(fmaxf(fy1, 0.0f), (float)tiles_h);
        *x0 = (int)fx0; *x1 = (int)fx1; *y0 = (int)fy0; *y1 = (int)fy1;
    }
}

/* Number of (gaussian, tile) intersections; also per-gaussian counts if `counts` != NULL.
 * binning.py:162-163.  radii are passed as float (the reference promotes int32 radii to
 * float in `means2d - radii`, binning.py:139). */
int64_t oracle_bin_count(
    int64_t N, const float* means2d, const float* radii_f, int W, int H, int tile_size,
    int semantics, int32_t* counts)
{
    const int tiles_w = (W + tile_size - 1) / tile_size, tiles_h = (H + tile_size - 1) / tile_size;
    int64_t total = 0;
    for (int64_t i = 0; i < N; ++i) {
        int x0, y0, x1, y1;
        tile_rect(means2d, radii_f, i, W, H, tile_size, tiles_w, tiles_h, semantics, &x0, &y0, &x1, &y1);
        int c = (x1 - x0) * (y1 - y0);
        if (c < 0) c = 0;
        if (counts) counts[i] = c;
        total += c;
    }
    return total;
}

static void radix_sort_u64_pairs(uint64_t* keys, int32_t* vals, int64_t n, int key_bits) {
    /* stable LSD radix sort, 11-bit digits */
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    int32_t* v2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    const int RB = 11, NB = 1 << RB;
    int64_t* hist = (int64_t*)malloc(sizeof(int64_t) * NB);
    uint64_t* src = keys; uint64_t* dst = k2; int32_t* vs = vals; int32_t* vd = v2;
    for (int shift = 0; shift < key_bits; shift += RB) {
        memset(hist, 0, sizeof(int64_t) * NB);
        for (int64_t i = 0; i < n; ++i) hist[(src[i] >> shift) & (NB - 1)]++;
        int64_t acc = 0;
        for (int b = 0; b < NB; ++b) { int64_t c = hist[b]; hist[b] = acc; acc += c; }
        for (int64_t i = 0; i < n; ++i) {
            int64_t p = hist[(src[i] >> shift) & (NB - 1)]++;
            dst[p] = src[i]; vd[p] = vs[i];
        }
        uint64_t* tk = src; src = dst; dst = tk;
        int32_t* tv = vs; vs = vd; vd = tv;
    }
    if (src != keys) { memcpy(keys, src, sizeof(uint64_t) * (size_t)n); memcpy(vals, vs, sizeof(int32_t) * (size_t)n); }
    free(k2); free(v2); free(hist);
}

/* Full binning: emits (tile, depth, gaussian) in the reference's emission order
 * (binning.py:172-200: gaussian ascending, ty ascending, tx ascending), sorts by depth and
 * then STABLY by tile (binning.py:223-231; the depth argsort is made stable here, the
 * canonical tie order of SURVEY H2), and derives tile_ranges with searchsorted-left
 * (binning.py:252-260).
 *   sorted_ids  [M] int32, isect_keys [M] uint64 (tile<<32 | depth_key) or NULL,
 *   tile_ranges [tiles_h*tiles_w*2] int32.
 * Returns M (must equal oracle_bin_count). */
int64_t oracle_bin(
    int64_t N, const float* means2d, const float* radii_f, const float* depths,
    int W, int H, int tile_size, int semantics,
    int64_t M_capacity, int32_t* sorted_ids, uint64_t* isect_keys, int32_t* tile_ranges)
{
    const int tiles_w = (W + tile_size - 1) / tile_size, tiles_h = (H + tile_size - 1) / tile_size;
    const int64_t n_tiles = (int64_t)tiles_w * tiles_h;
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(M_capacity > 0 ? M_capacity : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < N; ++i) {
        int x0, y0, x1, y1;
        tile_rect(means2d, radii_f, i, W, H, tile_size, tiles_w, tiles_h, semantics, &x0, &y0, &x1, &y1);
        const uint64_t dk = depth_key(depths[i]);
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                if (m >= M_capacity) { free(keys); return -1; }
                keys[m] = ((uint64_t)(ty * tiles_w + tx) << 32) | dk;
                sorted_ids[m] = (int32_t)i;
                ++m;
            }
    }
    int tile_bits = 1;
    while (((int64_t)1 << tile_bits) < n_tiles) ++tile_bits;
    radix_sort_u64_pairs(keys, sorted_ids, m, 32 + tile_bits);

    /* searchsorted(sorted_tile_ids, arange(n_tiles+1), side='left') */
    int64_t p = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
        while (p < m && (int64_t)(keys[p] >> 32) < t) ++p;
        int64_t e = p;
        while (e < m && (int64_t)(keys[e] >> 32) < t + 1) ++e;
        tile_ranges[2 * t] = (int32_t)p;
        tile_ranges[2 * t + 1] = (int32_t)e;
        p = e;
    }
    if (isect_keys) memcpy(isect_keys, keys, sizeof(uint64_t) * (size_t)m);
    free(keys);
    return m;
}

/* ------------------------------------------------------------------------------------
 * Rasterization.  kernels/rasterization.mojo:75-162.
 *   pixel centre +0.5 (:78-79); sigma (:138-142); alpha clamp and thresholds (:143-150);
 *   accumulate then update T (:152-157); background blend (:160-162).
 * stats[0] += evaluated (pixel, gaussian) pairs (every loop iteration reached),
 * stats[1] += contributing pairs (passed both tests).  stats may be NULL.
 * Out-of-range ids are skipped like the Mojo staging guard (:109).
 * ---------------------------------------------------------------------------------- */
void oracle_rasterize(
    int64_t N, int CDIM,
    const float* means2d, const float* conics, const float* colors, const float* opacities,
    const float* background,          /* [CDIM] */
    const int32_t* tile_ranges,       /* [tiles_h, tiles_w, 2] */
    const int32_t* sorted_ids,        /* [M] */
    int W, int H, int tile_size,
    float* image,                     /* [H, W, CDIM] */
    int64_t* stats)
{
    const int tiles_w = (W + tile_size - 1) / tile_size, tiles_h = (H + tile_size - 1) / tile_size;
    int64_t e_all = 0, e_pass = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : e_all, e_pass)
    for (int t = 0; t < tiles_w * tiles_h; ++t) {
        const int tr = t / tiles_w, tc = t % tiles_w;
        const int32_t r0 = tile_ranges[2 * t], r1 = tile_ranges[2 * t + 1];
        float pix[CDIM > 0 ? CDIM : 1];
        for (int ii = 0; ii < tile_size; ++ii) {
            const int i = tr * tile_size + ii;
            if (i >= H) break;
            for (int jj = 0; jj < tile_size; ++jj) {
                const int j = tc * tile_size + jj;
                if (j >= W) break;
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.0f;
                for (int c = 0; c < CDIM; ++c) pix[c] = 0.0f;
                for (int32_t k = r0; k < r1; ++k) {
                    const int32_t g = sorted_ids[k];
                    if (g < 0 || g >= N) continue;
                    ++e_all;
                    const float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    const float a = conics[3 * g], b = conics[3 * g + 1], cc = conics[3 * g + 2];
                    const float sigma = 0.5f * (a * dx * dx + cc * dy * dy) + b * dx * dy;
                    float alpha = opacities[g] * expf(-sigma);
                    if (alpha > 0.999f) alpha = 0.999f;
                    if (sigma < 0.0f || alpha < (1.0f / 255.0f)) continue;
                    const float next_T = T * (1.0f - alpha);
                    if (next_T <= 1e-4f) break;
                    const float vis = alpha * T;
                    for (int c = 0; c < CDIM; ++c) pix[c] += colors[(int64_t)g * CDIM + c] * vis;
                    T = next_T;
                    ++e_pass;
                }
                float* out = image + ((int64_t)i * W + j) * CDIM;
                for (int c = 0; c < CDIM; ++c) out[c] = pix[c] + T * background[c];
            }
        }
    }
    if (stats) { stats[0] += e_all; stats[1] += e_pass; }
}
