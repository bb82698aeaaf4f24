/*
 * n3tree_oracle.c — CPU restatement of the reference's per-ray hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product (mega-nerf-viewer_b200/csrc) never links, loads or calls it.
 *
 * The reference has no CPU implementation of this path (all of it is
 * __device__ code), so this file restates, function by function, the CUDA
 * sources it cites, using the arithmetic the reference's sm_100 build actually
 * executes (read from its PTX/SASS: where nvcc fuses a multiply-add this file
 * calls fmaf(), where it does not the operations stay separate; the file must
 * be compiled with -ffp-contract=off so the compiler adds no fusion of its
 * own).  Double-precision promotions caused by unsuffixed literals in the
 * reference (1e-9, 1e-6, SH constants) are reproduced.
 *
 * Parity status: PINNED against outputs of the reference's own CUDA kernel
 * (oracle/_ref, built from the unmodified sources) through tests/golden/ —
 * see tests/golden/README.md and oracle/make_golden.py.  The one operation
 * that cannot be reproduced bit-for-bit on a CPU is CUDA's expf (its
 * MUFU.EX2 table); visit sequences are therefore compared as "equal, or
 * prefix-equal where |T - stop_thresh| is within a few ulp" and pixels to
 * <= 1/255.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "../include/mnv_b200.h"

/* ------------------------------------------------------------------ helpers */

static inline float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal half -> normal float */
            int e = -1;
            do {
                ++e;
                man <<= 1;
            } while ((man & 0x400u) == 0);
            man &= 0x3ffu;
            bits = sign | ((uint32_t) (127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* CUDA fminf/fmaxf (FMNMX): if one operand is NaN the other is returned. */
static inline float fmin_c(float a, float b) { return fminf(a, b); }
static inline float fmax_c(float a, float b) { return fmaxf(a, b); }

/* include/cuda/common.cuh:10-14 — sqrtf(d0*d0 + d1*d1 + d2*d2); the sm_100
 * build evaluates FMUL(d1,d1), FFMA(d0,d0,.), FFMA(d2,d2,.). */
static inline float norm3(const float *d) {
    float s = d[1] * d[1];
    s = fmaf(d[0], d[0], s);
    s = fmaf(d[2], d[2], s);
    return sqrtf(s);
}

/* ------------------------------------------------------------------- query */

/* include/cuda/rt_core.cuh:117-159 query_single_from_root.
 * xyz is modified in place (becomes the position relative to the leaf, in
 * leaf units), exactly like the reference. Returns the depth. */
static inline int query_single_from_root(const mnv_tree_desc *tree, int32_t *visited, float *xyz,
                                         int32_t *chunk_idx, int32_t *child_idx,
                                         int track_visit) {
    const float hi = 1.f - 1e-6f; /* 0x3F7FFFEF */
    xyz[0] = fmax_c(fmin_c(xyz[0], hi), 0.f);
    xyz[1] = fmax_c(fmin_c(xyz[1], hi), 0.f);
    xyz[2] = fmax_c(fmin_c(xyz[2], hi), 0.f);
    const int N = tree->N;
    const int N3 = N * N * N;
    int32_t cur = 0;
    int depth = 1;
    for (;;) {
        if (track_visit) {
            if (visited[cur] == 0) visited[cur] = 1; /* atomicCAS(&visited[cur], 0, 1) */
        }
        int cidx = 0;
        for (int i = 0; i < 3; ++i) {
            xyz[i] *= (float) N;
            const float f = floorf(xyz[i]);
            cidx = (int) ((float) (cidx * N) + f); /* int*int + float, as written */
            xyz[i] -= f;
        }
        const int32_t skip = tree->child[(int64_t) cur * N3 + cidx];
        if (skip == 0) {
            *chunk_idx = cur;
            *child_idx = cidx;
            return depth;
        }
        depth += 1;
        cur += skip;
    }
}

/* ---- tiny pthread parallel-for (dynamic chunk claiming) ------------------- */
typedef void (*pfor_body)(void *ctx, int64_t begin, int64_t end);
typedef struct {
    pfor_body body;
    void *ctx;
    int64_t n, grain;
    atomic_llong next;
} pfor_job;

static void *pfor_worker(void *arg) {
    pfor_job *j = (pfor_job *) arg;
    for (;;) {
        const int64_t b = atomic_fetch_add(&j->next, j->grain);
        if (b >= j->n) break;
        const int64_t e = b + j->grain < j->n ? b + j->grain : j->n;
        j->body(j->ctx, b, e);
    }
    return NULL;
}

int oracle_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int) n : 1;
}

static void pfor(pfor_body body, void *ctx, int64_t n, int64_t grain, int nthreads) {
    if (nthreads <= 0) nthreads = oracle_num_threads();
    if (nthreads > 256) nthreads = 256;
    pfor_job job = {body, ctx, n, grain < 1 ? 1 : grain, 0};
    if (nthreads == 1 || n <= job.grain) {
        pfor_worker(&job);
        return;
    }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
        if (pthread_create(&th[started], NULL, pfor_worker, &job) == 0) ++started;
    pfor_worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

typedef struct {
    const mnv_tree_desc *tree;
    const float *xyz;
    int32_t *out;
} query_ctx;

static void query_body(void *vctx, int64_t b, int64_t e) {
    query_ctx *c = (query_ctx *) vctx;
    for (int64_t i = b; i < e; ++i) {
        float p[3] = {c->xyz[3 * i], c->xyz[3 * i + 1], c->xyz[3 * i + 2]};
        int32_t chunk, child;
        const int depth = query_single_from_root(c->tree, NULL, p, &chunk, &child, 0);
        c->out[3 * i] = chunk;
        c->out[3 * i + 1] = child;
        c->out[3 * i + 2] = depth;
    }
}

int oracle_query_points(const mnv_tree_desc *tree, const float *xyz, int64_t n, int32_t *out,
                        int nthreads) {
    if (!tree || !xyz || !out) return MNV_ERR_INVALID;
    query_ctx c = {tree, xyz, out};
    pfor(query_body, &c, n, 4096, nthreads);
    return MNV_OK;
}

/* -------------------------------------------------------------- SH basis */

/* include/cuda/rt_core.cuh:12-68 maybe_precalc_basis.  The SH constants are
 * double literals, so each product is formed in double and rounded once to
 * float; inner float sub-expressions follow the sm_100 build (PTX lines cited
 * in DESIGN.md): e.g. (xx - yy) is a float subtract, (2.0*zz - xx - yy) is all
 * double. */
static void precalc_basis(int format, int basis_dim, const float *dir, float *out) {
    if (format != MNV_FORMAT_SH) return;
    out[0] = (float) 0.28209479177387814;
    const float x = dir[0], y = dir[1], z = dir[2];
    const float xx = x * x, yy = y * y, zz = z * z;
    const float xy = x * y, yz = y * z, xz = x * z;
    switch (basis_dim) {
        case 25: {
            const float xx_m_yy = xx - yy;
            const float t3xx_yy = fmaf(xx, 3.f, -yy);
            const float z7m1 = fmaf(zz, 7.f, -1.f);
            const float z7m3 = fmaf(zz, 7.f, -3.f);
            const float xx_3yy = fmaf(yy, -3.f, xx);
            out[16] = (float) (((double) xy * 2.5033429417967046) * (double) xx_m_yy);
            out[17] = (float) (((double) yz * -1.7701307697799304) * (double) t3xx_yy);
            out[18] = (float) (((double) xy * 0.9461746957575601) * (double) z7m1);
            out[19] = (float) (((double) yz * -0.6690465435572892) * (double) z7m3);
            out[20] = (float) ((double) fmaf(zz, fmaf(zz, 35.f, -30.f), 3.f) * 0.10578554691520431);
            out[21] = (float) (((double) xz * -0.6690465435572892) * (double) z7m3);
            out[22] = (float) (((double) xx_m_yy * 0.47308734787878004) * (double) z7m1);
            out[23] = (float) (((double) xz * -1.7701307697799304) * (double) xx_3yy);
            out[24] = (float) ((double) fmaf(xx, xx_3yy, -(yy * t3xx_yy)) * 0.6258357354491761);
        } /* fallthrough */
        case 16: {
            const float a = fmaf(xx, 3.f, -yy);
            const float b = fmaf(zz, 4.f, -xx) - yy;
            const float c = fmaf(yy, -3.f, fmaf(xx, -3.f, zz + zz));
            out[9] = (float) (((double) y * -0.5900435899266435) * (double) a);
            out[10] = (float) (((double) xy * 2.890611442640554) * (double) z);
            out[11] = (float) (((double) y * -0.4570457994644658) * (double) b);
            out[12] = (float) (((double) z * 0.3731763325901154) * (double) c);
            out[13] = (float) (((double) x * -0.4570457994644658) * (double) b);
            out[14] = (float) (((double) z * 1.445305721320277) * (double) (xx - yy));
            out[15] = (float) (((double) x * -0.5900435899266435) * (double) fmaf(yy, -3.f, xx));
        } /* fallthrough */
        case 9:
            out[4] = (float) ((double) xy * 1.0925484305920792);
            out[5] = (float) ((double) yz * -1.0925484305920792);
            out[6] = (float) ((((double) zz + (double) zz) - (double) xx - (double) yy) *
                              0.31539156525252005);
            out[7] = (float) ((double) xz * -1.0925484305920792);
            out[8] = (float) ((double) (xx - yy) * 0.5462742152960396);
            /* fallthrough */
        case 4:
            out[1] = (float) ((double) y * -0.4886025119029199);
            out[2] = (float) ((double) z * 0.4886025119029199);
            out[3] = (float) ((double) x * -0.4886025119029199);
        default:
            break;
    }
}

/* ----------------------------------------------------------------- ray gen */

/* src/cuda/renderer_kernel.cu:30-38 screen2worlddir (+ common.cuh:16-23,25-31). */
static void screen2worlddir(int ix, int iy, const mnv_camera *cam, float *out, float *cen) {
    const float *m = cam->c2w;
    const float vx = ((float) ix + 0.5f - cam->cx) / cam->fx;
    const float vy = -((float) iy + 0.5f - cam->cy) / cam->fy;
    /* _mv3 with v[2] = -1: FMUL(vy, m[3+r]); FFMA(vx, m[r], .); FADD(., -m[6+r]) */
    out[0] = fmaf(vx, m[0], vy * m[3]) - m[6];
    out[1] = fmaf(vx, m[1], vy * m[4]) - m[7];
    out[2] = fmaf(vx, m[2], vy * m[5]) - m[8];
    const float invnorm = 1.f / norm3(out);
    out[0] *= invnorm;
    out[1] *= invnorm;
    out[2] *= invnorm;
    cen[0] = m[9];
    cen[1] = m[10];
    cen[2] = m[11];
}

/* src/cuda/renderer_kernel.cu:40-61 rodrigues. `angle < 1e-6` compares in
 * double; (1.0 - cos_angle) is double, so the last term is a double FMA. */
static void rodrigues(const float *aa, float *dir) {
    const float angle = norm3(aa);
    if ((double) angle < 1e-6) return;
    float k[3];
    for (int i = 0; i < 3; ++i) k[i] = aa[i] / angle;
    const float cos_angle = cosf(angle), sin_angle = sinf(angle);
    float cross[3];
    cross[0] = fmaf(k[1], dir[2], -(k[2] * dir[1]));
    cross[1] = fmaf(k[2], dir[0], -(k[0] * dir[2]));
    cross[2] = fmaf(k[0], dir[1], -(k[1] * dir[0]));
    float dot = k[1] * dir[1];
    dot = fmaf(k[0], dir[0], dot);
    dot = fmaf(k[2], dir[2], dot);
    const double omc = 1.0 - (double) cos_angle;
    for (int i = 0; i < 3; ++i) {
        const float a = fmaf(cos_angle, dir[i], sin_angle * cross[i]);
        const float b = dot * k[i];
        dir[i] = (float) fma(omc, (double) b, (double) a);
    }
}

/* ------------------------------------------------------------------ march */

typedef struct {
    uint64_t *hash;    /* per ray FNV-1a over chunk*8+child */
    int32_t *count;    /* per ray visits */
    int32_t *shaded;   /* per ray visits with sigma > sigma_thresh */
    int32_t *log;      /* per ray first log_cap packed leaf indices */
    int log_cap;
} visit_sink;

/* include/cuda/rt_core.cuh:162-332 render_voxels_trace_ray. */
static void trace_ray(const mnv_tree_desc *tree, int32_t *visited, float *dir, const float *vdir,
                      const float *cen, const mnv_render_options *opt, float tmax_bg, float *out,
                      float *split /* [3] = priority, chunk, child */,
                      float *sample /* [3] = priority, chunk, child */, int track_visit,
                      const visit_sink *sink, int64_t ray) {
    const int N3 = tree->N * tree->N * tree->N;
    const int D = tree->data_dim;
    split[0] = (float) (opt->max_depth + 1);
    sample[0] = (float) (opt->max_sample_count + 1);

    /* _get_delta_scale, rt_core.cuh:102-115 */
    dir[0] *= tree->scale[0];
    dir[1] *= tree->scale[1];
    dir[2] *= tree->scale[2];
    const float delta_scale = 1.f / norm3(dir);
    dir[0] *= delta_scale;
    dir[1] *= delta_scale;
    dir[2] *= delta_scale;
    tmax_bg = tmax_bg / delta_scale;

    float invdir[3];
    for (int i = 0; i < 3; ++i) invdir[i] = (float) (1.0 / ((double) dir[i] + 1e-9));

    /* _dda_world, rt_core.cuh:70-86 (double arithmetic, rounded to float) */
    float tmin = 0.f, tmax = 1e4f;
    for (int i = 0; i < 3; ++i) {
        const float t1 = (float) ((((double) opt->render_bbox[i] + 1e-6) - (double) cen[i]) *
                                  (double) invdir[i]);
        const float t2 = (float) ((((double) opt->render_bbox[i + 3] - 1e-6) - (double) cen[i]) *
                                  (double) invdir[i]);
        tmin = fmax_c(tmin, fmin_c(t1, t2));
        tmax = fmin_c(tmax, fmax_c(t1, t2));
    }
    tmax = fmin_c(tmax, tmax_bg);

    if (tmax < 0 || tmin > tmax) {
        if (opt->render_depth) out[3] = 1.f;
        if (sink) { /* ray misses the box: empty visit sequence */
            if (sink->hash) sink->hash[ray] = 0xcbf29ce484222325ULL;
            if (sink->count) sink->count[ray] = 0;
            if (sink->shaded) sink->shaded[ray] = 0;
        }
        return;
    }

    float basis_fn[MNV_GLOBAL_BASIS_MAX];
    memset(basis_fn, 0, sizeof(basis_fn)); /* the reference leaves unused slots undefined */
    precalc_basis(tree->format, tree->basis_dim, vdir, basis_fn);
    for (int i = 0; i < opt->basis_minmax[0] && i < MNV_GLOBAL_BASIS_MAX; ++i) basis_fn[i] = 0.f;
    for (int i = opt->basis_minmax[1] + 1; i < MNV_GLOBAL_BASIS_MAX; ++i)
        if (i >= 0) basis_fn[i] = 0.f;

    float light_intensity = 1.f;
    float t = tmin;
    float max_weight = -1.f, max_sample_weight = -1.f;
    uint64_t h = 0xcbf29ce484222325ULL;
    int nvis = 0, nshaded = 0;
    float pos[3];
    int32_t chunk_idx, child_idx;
    const int bd = tree->basis_dim;

    while (t < tmax) {
        pos[0] = fmaf(t, dir[0], cen[0]);
        pos[1] = fmaf(t, dir[1], cen[1]);
        pos[2] = fmaf(t, dir[2], cen[2]);
        const int depth =
                query_single_from_root(tree, visited, pos, &chunk_idx, &child_idx, track_visit);
        if (sink) {
            const int32_t packed = chunk_idx * 8 + child_idx;
            h = (h ^ (uint64_t) (int64_t) ((int64_t) chunk_idx * 8 + child_idx)) * 0x100000001b3ULL;
            if (sink->log && nvis < sink->log_cap) sink->log[ray * sink->log_cap + nvis] = packed;
        }
        ++nvis;
        const float cube_size = powf((float) tree->N, (float) depth); /* exact for N = 2 */

        /* _dda_unit, rt_core.cuh:88-100: FMUL then FADD (not fused in the sm_100 build) */
        float tm = 1e4f;
        for (int i = 0; i < 3; ++i) {
            const float t1 = -pos[i] * invdir[i];
            const float t2 = t1 + invdir[i];
            tm = fmin_c(tm, fmax_c(t1, t2));
        }
        const float t_subcube = tm / cube_size;
        const float delta_t = t_subcube + opt->step_size;
        const int64_t leaf = ((int64_t) chunk_idx * N3 + child_idx);
        const uint16_t *rec = tree->data + leaf * D;
        const float sigma = half_to_float(rec[D - 1]);

        if (sigma > opt->sigma_thresh) {
            ++nshaded;
            const float att = expf((-delta_t * delta_scale) * sigma);
            const float weight = light_intensity * (1.f - att);

            if (weight > max_weight && depth < opt->max_depth) {
                split[1] = (float) chunk_idx;
                split[2] = (float) child_idx;
                split[0] = (float) depth;
                max_weight = weight;
            }
            const int16_t sc = tree->sample_counts ? tree->sample_counts[leaf] : 8;
            if (weight > max_sample_weight && sc < opt->max_sample_count) {
                sample[1] = (float) chunk_idx;
                sample[2] = (float) child_idx;
                sample[0] = (float) sc;
                max_sample_weight = weight;
            }

            if (opt->render_depth) {
                out[0] = fmaf(t, weight, out[0]);
            } else if (bd >= 0) {
                int off = 0;
                for (int ch = 0; ch < 3; ++ch) {
                    const uint16_t *c = rec + off;
#define C(k) half_to_float(c[k])
#define B(k) basis_fn[k]
                    float tmp = B(0) * C(0);
                    /* each case group: FMUL of the 2nd product, FFMA chain, one FADD */
                    switch (bd) {
                        case 25: {
                            float s = B(17) * C(17);
                            s = fmaf(B(16), C(16), s);
                            for (int k = 18; k <= 24; ++k) s = fmaf(B(k), C(k), s);
                            tmp = tmp + s;
                        } /* fallthrough */
                        case 16: {
                            float s = B(10) * C(10);
                            s = fmaf(B(9), C(9), s);
                            for (int k = 11; k <= 15; ++k) s = fmaf(B(k), C(k), s);
                            tmp = tmp + s;
                        } /* fallthrough */
                        case 9: {
                            float s = B(5) * C(5);
                            s = fmaf(B(4), C(4), s);
                            for (int k = 6; k <= 8; ++k) s = fmaf(B(k), C(k), s);
                            tmp = tmp + s;
                        } /* fallthrough */
                        case 4: {
                            float s = B(2) * C(2);
                            s = fmaf(B(1), C(1), s);
                            s = fmaf(B(3), C(3), s);
                            tmp = tmp + s;
                        }
                        default:
                            break;
                    }
#undef C
#undef B
                    out[ch] += weight / (1.f + expf(-tmp));
                    off += bd;
                }
            } else {
                for (int j = 0; j < 3; ++j) out[j] = fmaf(weight, half_to_float(rec[j]), out[j]);
            }

            light_intensity *= att;

            if (light_intensity < opt->stop_thresh) {
                if (opt->render_depth) out[0] = out[1] = out[2] = fmin_c(out[0] * 0.3f, 1.0f);
                const float scale = 1.f / (1.f - light_intensity);
                out[0] *= scale;
                out[1] *= scale;
                out[2] *= scale;
                out[3] = 1.f;
                goto done;
            }
        } else {
            if (max_weight == -1 && depth < opt->max_depth) {
                split[1] = (float) chunk_idx;
                split[2] = (float) child_idx;
                split[0] = (float) depth;
            }
            const int16_t sc = tree->sample_counts ? tree->sample_counts[leaf] : 8;
            if (max_sample_weight == -1 && sc < opt->max_sample_count) {
                sample[1] = (float) chunk_idx;
                sample[2] = (float) child_idx;
                sample[0] = (float) sc;
            }
        }
        t += delta_t;
    }
    if (opt->render_depth) {
        out[0] = out[1] = out[2] = fmin_c(out[0] * 0.3f, 1.0f);
        out[3] = 1.f;
    } else {
        out[3] = 1.f - light_intensity;
    }
done:
    if (sink) {
        if (sink->hash) sink->hash[ray] = h;
        if (sink->count) sink->count[ray] = nvis;
        if (sink->shaded) sink->shaded[ray] = nshaded;
    }
}

static inline uint8_t to_u8(float v) {
    /* uint8_t(out * 255): cvt.rzi.u32.f32 (saturating, NaN -> 0) then low byte */
    const float s = v * 255.f;
    uint32_t u;
    if (!(s > 0.f)) u = 0;
    else if (s >= 4294967296.f) u = 0xffffffffu;
    else u = (uint32_t) s;
    return (uint8_t) (u & 0xffu);
}

typedef struct {
    const mnv_tree_desc *tree;
    const mnv_camera *cam;
    const mnv_render_options *opt;
    uint8_t *rgba;
    const float *depth_in;
    float *to_split, *to_sample;
    int32_t *visited;
    int track_visit, offscreen;
    const visit_sink *sink;
    int y0, row_step;
} render_ctx;

static void render_rows(void *vctx, int64_t rb, int64_t re) {
    render_ctx *c = (render_ctx *) vctx;
    const mnv_tree_desc *tree = c->tree;
    const mnv_render_options *opt = c->opt;
    const int W = c->cam->width;
    for (int64_t r = rb; r < re; ++r) {
        const int y = c->y0 + (int) r * c->row_step;
        for (int x = 0; x < W; ++x) {
            const int64_t idx = (int64_t) y * W + x;
            float dir[3], cen[3], out[4] = {0.f, 0.f, 0.f, 0.f};
            uint8_t *px = c->rgba + idx * 4;
            const uint8_t init[3] = {px[0], px[1], px[2]};
            float dummy_split[3], dummy_sample[3];
            float *split = c->to_split ? c->to_split + idx * 3 : dummy_split;
            float *sample = c->to_sample ? c->to_sample + idx * 3 : dummy_sample;
            if (tree->N > 0) {
                screen2worlddir(x, y, c->cam, dir, cen);
                for (int i = 0; i < 3; ++i) cen[i] = fmaf(tree->scale[i], cen[i], tree->offset[i]);
                float t_max = 1e9f;
                if (!c->offscreen && c->depth_in) t_max = c->depth_in[idx];
                float vdir[3] = {dir[0], dir[1], dir[2]};
                rodrigues(opt->rot_dirs, vdir);
                trace_ray(tree, c->visited, dir, vdir, cen, opt, t_max, out, split, sample,
                          c->track_visit, c->sink, idx);
            }
            /* composite_and_write, renderer_kernel.cu:215-241 */
            const float nalpha = 1.f - out[3];
            if (c->offscreen) {
                const float remain = opt->background_brightness * nalpha;
                out[0] += remain;
                out[1] += remain;
                out[2] += remain;
            } else {
                out[0] = fmaf(init[0] / 255.f, nalpha, out[0]);
                out[1] = (init[1] / 255.f) * nalpha + out[1];
                out[2] = (init[2] / 255.f) * nalpha + out[2];
            }
            px[0] = to_u8(out[0]);
            px[1] = to_u8(out[1]);
            px[2] = to_u8(out[2]);
            px[3] = 255;
        }
    }
}

/*
 * src/cuda/renderer_kernel.cu:243-292 render_voxels_kernel + :215-241
 * composite_and_write, for pixels of rows y0, y0+row_step, ... < y1.
 *   rgba       u8 [H][W][4]; when !offscreen it must hold the prior colour on entry
 *   depth_in   f32 [H][W] t_max surface (only read when !offscreen), may be NULL
 *   to_split / to_sample  f32 [P][3] pre-filled with -1 by the caller
 *              (cuda_renderer.cpp:97-98), may be NULL
 *   visited    i32 [capacity] (only touched when track_visit; racy-but-benign
 *              like the reference's atomicCAS marks)
 * Every other pointer may be NULL.  nthreads <= 0 -> all online cores.
 */
int oracle_render_voxels(const mnv_tree_desc *tree, const mnv_camera *cam,
                         const mnv_render_options *opt, uint8_t *rgba, const float *depth_in,
                         float *to_split, float *to_sample, int32_t *visited, int track_visit,
                         int offscreen, uint64_t *visit_hash, int32_t *visit_count,
                         int32_t *shaded_count, int32_t *visit_log, int log_cap, int y0, int y1,
                         int row_step, int nthreads) {
    if (!tree || !cam || !opt || !rgba) return MNV_ERR_INVALID;
    if (tree->N != 2) return MNV_ERR_INVALID;
    if (row_step < 1) row_step = 1;
    if (y1 > cam->height) y1 = cam->height;
    if (y0 < 0 || y0 >= y1) return MNV_ERR_INVALID;
    visit_sink sink_v = {visit_hash, visit_count, shaded_count, visit_log, log_cap};
    const visit_sink *sink =
            (visit_hash || visit_count || shaded_count || visit_log) ? &sink_v : NULL;
    const int nrows = (y1 - y0 + row_step - 1) / row_step;
    render_ctx c = {tree,    cam,         opt,       rgba, depth_in, to_split, to_sample,
                    visited, track_visit, offscreen, sink, y0,       row_step};
    pfor(render_rows, &c, nrows, 1, nthreads);
    return MNV_OK;
}

/* ===================================================================== A7 ==
 * include/cuda/rt_core.cuh:418-576 get_samples_trace_ray +
 * src/cuda/renderer_kernel.cu:329-363 get_samples_from_voxels_kernel, dense
 * outputs exactly like the reference (only sensible at test sizes):
 *   num_samples i16 [P]; samples f32 [P][S][sd]; cluster_indices i16 [P][S]
 * sd = 4 + 3*need_viewdir + (appearance_embedding != -1).
 */
typedef struct {
    const mnv_tree_desc *tree;
    const mnv_camera *cam;
    const mnv_render_options *opt;
    const int32_t *grid_dim;
    const float *min_position, *range;
    int16_t *num_samples;
    float *samples;
    int16_t *cluster;
    float *to_split, *to_sample;
    int S, sd;
} samples_ctx;

static void samples_rows(void *vctx, int64_t rb, int64_t re) {
    samples_ctx *c = (samples_ctx *) vctx;
    const mnv_tree_desc *tree = c->tree;
    const mnv_render_options *opt = c->opt;
    const int W = c->cam->width, N3 = 8, D = tree->data_dim;
    for (int64_t y = rb; y < re; ++y) {
        for (int x = 0; x < W; ++x) {
            const int64_t idx = y * W + x;
            float true_dir[3], true_cen[3];
            screen2worlddir(x, (int) y, c->cam, true_dir, true_cen);
            float vdir[3] = {true_dir[0], true_dir[1], true_dir[2]};
            rodrigues(opt->rot_dirs, vdir);
            float dummy_a[3], dummy_b[3];
            float *split = c->to_split ? c->to_split + idx * 3 : dummy_a;
            float *sample = c->to_sample ? c->to_sample + idx * 3 : dummy_b;
            split[0] = (float) (opt->max_depth + 1);
            sample[0] = (float) (opt->max_sample_count + 1);
            float cen[3], dir[3];
            for (int i = 0; i < 3; ++i) cen[i] = fmaf(tree->scale[i], true_cen[i], tree->offset[i]);
            for (int i = 0; i < 3; ++i) dir[i] = true_dir[i] * tree->scale[i];
            const float delta_scale = 1.f / norm3(dir);
            for (int i = 0; i < 3; ++i) dir[i] *= delta_scale;
            const float tmax_bg = 1e9f / delta_scale;
            float invdir[3];
            for (int i = 0; i < 3; ++i) invdir[i] = (float) (1.0 / ((double) dir[i] + 1e-9));
            float tmin = 0.f, tmax = 1e4f;
            for (int i = 0; i < 3; ++i) {
                const float t1 = (float) ((((double) opt->render_bbox[i] + 1e-6) - (double) cen[i]) *
                                          (double) invdir[i]);
                const float t2 = (float) ((((double) opt->render_bbox[i + 3] - 1e-6) - (double) cen[i]) *
                                          (double) invdir[i]);
                tmin = fmax_c(tmin, fmin_c(t1, t2));
                tmax = fmin_c(tmax, fmax_c(t1, t2));
            }
            tmax = fmin_c(tmax, tmax_bg);
            int16_t *ns = c->num_samples + idx;
            if (tmax < 0 || tmin > tmax) continue;
            float light_intensity = 1.f, t = tmin, max_weight = -1.f, max_sample_weight = -1.f;
            float pos[3];
            int32_t chunk_idx, child_idx;
            while (t < tmax) {
                /* NOT fused here: the reference's get_samples kernel executes FMUL + FADD
                 * (t*dir is reused for true_z), unlike its render kernel's FFMA */
                for (int i = 0; i < 3; ++i) pos[i] = cen[i] + t * dir[i];
                const int depth = query_single_from_root(tree, NULL, pos, &chunk_idx, &child_idx, 0);
                const float cube_size = powf((float) tree->N, (float) depth);
                float tm = 1e4f;
                for (int i = 0; i < 3; ++i) {
                    const float t1 = -pos[i] * invdir[i];
                    const float t2 = t1 + invdir[i];
                    tm = fmin_c(tm, fmax_c(t1, t2));
                }
                const float delta_t = tm / cube_size + opt->step_size;
                const int64_t leaf = (int64_t) chunk_idx * N3 + child_idx;
                const float sigma = half_to_float(tree->data[leaf * D + D - 1]);
                const int16_t sc = tree->sample_counts ? tree->sample_counts[leaf] : 8;
                if (sigma > opt->sigma_thresh) {
                    const float att = expf((-delta_t * delta_scale) * sigma);
                    const float weight = light_intensity * (1.f - att);
                    if (weight > max_weight && depth < opt->max_depth) {
                        split[1] = (float) chunk_idx;
                        split[2] = (float) child_idx;
                        split[0] = (float) depth;
                        max_weight = weight;
                    }
                    if (weight > max_sample_weight && sc < opt->max_sample_count) {
                        sample[1] = (float) chunk_idx;
                        sample[2] = (float) child_idx;
                        sample[0] = (float) sc;
                        max_sample_weight = weight;
                    }
                    if (*ns < opt->max_guided_samples) {
                        float *row = c->samples + ((int64_t) idx * c->S + *ns) * c->sd;
                        float tz[3];
                        for (int i = 0; i < 3; ++i) tz[i] = (t * dir[i]) / tree->scale[i];
                        row[0] = norm3(tz);
                        for (int i = 0; i < 3; ++i) row[1 + i] = fmaf(true_dir[i], row[0], true_cen[i]);
                        if (opt->need_viewdir) {
                            row[4] = vdir[0];
                            row[5] = vdir[1];
                            row[6] = vdir[2];
                            if (opt->appearance_embedding != -1) row[7] = (float) opt->appearance_embedding;
                        } else if (opt->appearance_embedding != -1) {
                            row[4] = (float) opt->appearance_embedding;
                        }
                        const float g0 = (float) c->grid_dim[0], g1 = (float) c->grid_dim[1];
                        const int a = (int) fmax_c(
                                fmin_c((row[2] - c->min_position[1]) / c->range[1] * g0, g0 - 1.0f), 0.0f);
                        const int b = (int) fmax_c(
                                fmin_c((row[3] - c->min_position[2]) / c->range[2] * g1, g1 - 1.0f), 0.0f);
                        c->cluster[(int64_t) idx * c->S + *ns] = (int16_t) (a * c->grid_dim[1] + b);
                        *ns += 1;
                    }
                    light_intensity *= att;
                    if (light_intensity < opt->stop_thresh) break;
                } else {
                    if (max_weight == -1 && depth < opt->max_depth) {
                        split[1] = (float) chunk_idx;
                        split[2] = (float) child_idx;
                        split[0] = (float) depth;
                    }
                    if (max_sample_weight == -1 && sc < opt->max_sample_count) {
                        sample[1] = (float) chunk_idx;
                        sample[2] = (float) child_idx;
                        sample[0] = (float) sc;
                    }
                }
                t += delta_t;
            }
        }
    }
}

int oracle_get_samples(const mnv_tree_desc *tree, const mnv_camera *cam, const mnv_render_options *opt,
                       const int32_t *grid_dim, const float *min_position, const float *range,
                       int16_t *num_samples, float *samples, int16_t *cluster, int S, int sd,
                       float *to_split, float *to_sample, int nthreads) {
    if (!tree || !cam || !opt || !num_samples || !samples || !cluster) return MNV_ERR_INVALID;
    samples_ctx c = {tree, cam, opt, grid_dim, min_position, range, num_samples,
                     samples, cluster, to_split, to_sample, S, sd};
    pfor(samples_rows, &c, cam->height, 1, nthreads);
    return MNV_OK;
}

/* ===================================================================== A8 ==
 * include/cuda/rt_core.cuh:334-416 composite_nerf_results +
 * src/cuda/renderer_kernel.cu:294-327 render_nerf_results_kernel (offscreen).
 * sample_values f32 [V][stride]; sigma is read from column `sigma_col`
 * (the reference always uses 3, rt_core.cuh:365).
 */
int oracle_composite_nerf(int format, int basis_dim, const mnv_camera *cam,
                          const mnv_render_options *opt, const float *sample_values, int stride,
                          int sigma_col, const float *z_vals, const int64_t *offsets, uint8_t *rgba) {
    const int W = cam->width, H = cam->height;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int64_t idx = (int64_t) y * W + x;
            float out[3] = {0.f, 0.f, 0.f};
            const int64_t start = idx == 0 ? 0 : offsets[idx - 1], end = offsets[idx];
            if (start != end) {
                float dir[3], cen[3];
                screen2worlddir(x, y, cam, dir, cen);
                rodrigues(opt->rot_dirs, dir);
                float basis_fn[MNV_GLOBAL_BASIS_MAX];
                memset(basis_fn, 0, sizeof(basis_fn));
                precalc_basis(format, basis_dim, dir, basis_fn);
                for (int i = 0; i < opt->basis_minmax[0] && i < MNV_GLOBAL_BASIS_MAX; ++i) basis_fn[i] = 0.f;
                for (int i = opt->basis_minmax[1] + 1; i < MNV_GLOBAL_BASIS_MAX; ++i)
                    if (i >= 0) basis_fn[i] = 0.f;
                float ti = 1.f, wc = 0.f;
                for (int64_t i = start; i < end; ++i) {
                    const float *sv = sample_values + i * stride;
                    float weight;
                    if (i < end - 1) {
                        const float delta = z_vals[i + 1] - z_vals[i];
                        wc = expf(-sv[sigma_col] * delta);
                        weight = ti * (1.0f - wc);
                    } else {
                        weight = ti;
                    }
                    if (opt->render_depth) {
                        out[0] = fmaf(ti, weight, out[0]);
                    } else if (basis_dim >= 0) {
                        for (int ch = 0; ch < 3; ++ch) {
                            const float *c = sv + ch * basis_dim;
                            float tmp = basis_fn[0] * c[0];
                            if (basis_dim == 25) {
                                float s = basis_fn[17] * c[17];
                                s = fmaf(basis_fn[16], c[16], s);
                                for (int k = 18; k <= 24; ++k) s = fmaf(basis_fn[k], c[k], s);
                                tmp = tmp + s;
                            }
                            if (basis_dim == 25 || basis_dim == 16) {
                                float s = basis_fn[10] * c[10];
                                s = fmaf(basis_fn[9], c[9], s);
                                for (int k = 11; k <= 15; ++k) s = fmaf(basis_fn[k], c[k], s);
                                tmp = tmp + s;
                            }
                            if (basis_dim == 25 || basis_dim == 16 || basis_dim == 9) {
                                float s = basis_fn[5] * c[5];
                                s = fmaf(basis_fn[4], c[4], s);
                                for (int k = 6; k <= 8; ++k) s = fmaf(basis_fn[k], c[k], s);
                                tmp = tmp + s;
                            }
                            if (basis_dim == 25 || basis_dim == 16 || basis_dim == 9 || basis_dim == 4) {
                                float s = basis_fn[2] * c[2];
                                s = fmaf(basis_fn[1], c[1], s);
                                s = fmaf(basis_fn[3], c[3], s);
                                tmp = tmp + s;
                            }
                            out[ch] += weight / (1.f + expf(-tmp));
                        }
                    } else {
                        for (int j = 0; j < 3; ++j) out[j] = fmaf(weight, sv[j], out[j]);
                    }
                    ti *= wc;
                }
                if (opt->render_depth) out[0] = out[1] = out[2] = fmin_c(out[0] * 0.3f, 1.0f);
            }
            uint8_t *px = rgba + idx * 4;
            px[0] = to_u8(out[0]);
            px[1] = to_u8(out[1]);
            px[2] = to_u8(out[2]);
            px[3] = 255;
        }
    return MNV_OK;
}
