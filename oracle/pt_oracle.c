/* oracle/pt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's path-tracing loop (DeltaPavonis/cpp_raytracer), used
 * ONLY as a checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The
 * product (cpp_raytracer_b200/) never includes, links or calls anything in this directory.
 *
 * PARITY PINNING: the reference has no tests or golden vectors of its own,
 * so this restatement is pinned against the reference ITSELF, compiled here from its own
 * sources as oracle/_ref/ref_bridge (oracle/Makefile):
 *   - closest hit: bit-exact (prim index and t) vs the reference's Scene::hit_by on the golden
 *     ray sets of all seven small scenes            (tests/test_oracle_cpu.py)
 *   - render: BIT-EXACT double pixels vs the reference's single-threaded Camera::render on
 *     golden renders (same LCG state in, same pixels out), which pins ray generation, every
 *     material, the recursion, the RNG draw order and the accumulation
 *   - rand_double / reflect / refract / Schlick / tone map: known answers in tests/golden/kat.json
 *
 * Every function cites the reference lines it follows.  Arithmetic is written operation by
 * operation in the reference's order and compiled with -ffp-contract=off (the reference is
 * built for baseline x86-64: no FMA).  Note the reference's idiom `v / d  ==  v * (1 / d)`
 * (vec3d.h:31), which matters at the last bit.
 */
#include "pt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y, z; } V3;

static V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
static V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }       /* vec3d.h:21,93 */
static V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }       /* vec3d.h:23,95 */
static V3 vmul(V3 a, double d) { return v3(a.x * d, a.y * d, a.z * d); }         /* vec3d.h:25,97 */
static V3 vdiv(V3 a, double d) { return vmul(a, 1 / d); }                        /* vec3d.h:31: multiply by 1/d */
static V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }     /* vec3d.h:109 */
static V3 vcross(V3 a, V3 b) {                                                   /* vec3d.h:112-114 */
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static double vmag2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }          /* vec3d.h:37 */
static double vmag(V3 a) { return sqrt(vmag2(a)); }                              /* vec3d.h:34 */
static V3 vunit(V3 a) { return vdiv(a, vmag(a)); }                               /* vec3d.h:127-130 */
static V3 from(const double *p) { return v3(p[0], p[1], p[2]); }
static void to(double *p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* ---- RNG: rand_util.h ------------------------------------------------------------------- */
uint32_t oracle_seed_sequence_next(uint32_t *s) {                                /* rand_util.h:66-68 */
    *s = 2483477u * (*s) + 2987434823u;
    return *s;
}
double oracle_rand_double(uint32_t *state, double min, double max) {             /* rand_util.h:85-117 */
    *state = 1664525u * (*state) + 1013904223u;
    const double scale = 1 / (double)(4294967295u - 1);
    return min + (max - min) * (double)(*state) * scale;
}

/* ---- camera: camera.h:87-157 ------------------------------------------------------------ */
void oracle_camera_init(OCamera *c) {
    double aspect = (double)c->image_w / (double)c->image_h;
    V3 dir = from(c->dir), center = from(c->center);
    if (c->focus_dist < 0) c->focus_dist = vmag(dir);                            /* :101-103 */
    double focal = c->focus_dist, vw, vh;
    if (c->vfov >= 0) { vh = 2 * focal * tan(c->vfov / 2); vw = vh * aspect; }   /* :113-119 */
    else { vw = 2 * focal * tan(c->hfov / 2); vh = vw / aspect; }
    V3 bz = vneg(vunit(dir));                                                    /* :122 */
    V3 bx = vunit(vcross(from(c->up), bz));                                      /* :123 */
    V3 by = vcross(bz, bx);                                                      /* :124 */
    V3 xv = vmul(bx, vw), yv = vmul(by, -vh);                                    /* :130 */
    V3 dx = vdiv(xv, (double)c->image_w), dy = vdiv(yv, (double)c->image_h);     /* :131-132 */
    V3 ulc = vsub(vsub(vsub(center, vmul(bz, focal)), vdiv(xv, 2)), vdiv(yv, 2));/* :139 */
    V3 p00 = vadd(vadd(ulc, vdiv(dx, 2)), vdiv(dy, 2));                          /* :144 */
    double rad = focal * tan(c->defocus_angle / 2);                              /* :150 */
    to(c->delta_x, dx); to(c->delta_y, dy); to(c->pixel00, p00);
    to(c->disk_x, vmul(bx, rad)); to(c->disk_y, vmul(by, rad));                  /* :151-152 */
}

/* ---- primitives ------------------------------------------------------------------------- */
typedef struct { int hit; double t; V3 p, n; int front; uint32_t mat; } HitInfo;

/* hittable.h:46-71 */
static void set_face(HitInfo *h, V3 dir, V3 outward) {
    if (vdot(dir, outward) > 0) { h->n = vneg(outward); h->front = 0; }
    else { h->n = outward; h->front = 1; }
}
static int in_open(double tmin, double tmax, double t) { return tmin < t && t < tmax; }   /* interval.h:38 */

/* sphere.h:45-96 */
static HitInfo sphere_hit(const OSphere *s, V3 o, V3 d, double tmin, double tmax) {
    HitInfo h; h.hit = 0;
    V3 c = from(s->c);
    V3 oc = vsub(o, c);
    double a = vdot(d, d), b_half = vdot(d, oc), cc = vdot(oc, oc) - s->r * s->r;
    double disc = b_half * b_half - a * cc;
    if (disc < 0) return h;
    double sq = sqrt(disc);
    double root = (-b_half - sq) / a;
    if (!in_open(tmin, tmax, root)) {
        root = (-b_half + sq) / a;
        if (!in_open(tmin, tmax, root)) return h;
    }
    h.hit = 1; h.t = root;
    h.p = vadd(o, vmul(d, root));                                                /* ray3d.h:16 */
    set_face(&h, d, vdiv(vsub(h.p, c), s->r));                                   /* sphere.h:94 */
    h.mat = s->mat;
    return h;
}

/* parallelogram.h:177-240 with the constructor's precompute (parallelogram.h:269-279) */
static HitInfo quad_hit(const OQuad *q, V3 o, V3 d, double tmin, double tmax) {
    HitInfo h; h.hit = 0;
    V3 v = from(q->v), s1 = from(q->s1), s2 = from(q->s2);
    V3 n = vcross(s1, s2);
    V3 un = vunit(n);
    V3 w = vdiv(n, vmag2(n));
    double den = vdot(un, d);
    if (fabs(den) < 1e-9) return h;
    double t = vdot(un, vsub(v, o)) / den;
    if (!in_open(tmin, tmax, t)) return h;
    V3 p = vadd(o, vmul(d, t));
    V3 pl = vsub(p, v);
    double alpha = vdot(w, vcross(pl, s2)), beta = vdot(w, vcross(s1, pl));
    if (!(0 <= alpha && alpha <= 1 && 0 <= beta && beta <= 1)) return h;         /* interval.h:35 inclusive */
    h.hit = 1; h.t = t; h.p = p;
    set_face(&h, d, un);
    h.mat = q->mat;
    return h;
}

/* Canonical primitive order = Scene::get_primitive_components() (scene.h:85-105); the flat
 * arrays carry each primitive's canonical index, so merge the two arrays by `prim`. */
typedef struct { uint32_t n; uint32_t *type_index; } PrimOrder;   /* bit31 = quad */
static PrimOrder prim_order(const OScene *s) {
    PrimOrder po; po.n = (uint32_t)(s->n_spheres + s->n_quads);
    po.type_index = (uint32_t *)malloc(sizeof(uint32_t) * (po.n ? po.n : 1));
    for (uint64_t i = 0; i < s->n_spheres; ++i) po.type_index[s->spheres[i].prim] = (uint32_t)i;
    for (uint64_t i = 0; i < s->n_quads; ++i) po.type_index[s->quads[i].prim] = 0x80000000u | (uint32_t)i;
    return po;
}
static HitInfo prim_hit(const OScene *s, uint32_t ti, V3 o, V3 d, double tmin, double tmax) {
    return (ti & 0x80000000u) ? quad_hit(&s->quads[ti & 0x7fffffffu], o, d, tmin, tmax)
                              : sphere_hit(&s->spheres[ti], o, d, tmin, tmax);
}
/* scene.h:59-75: first primitive wins ties because later ones must be STRICTLY closer */
static HitInfo closest(const OScene *s, const PrimOrder *po, V3 o, V3 d, double tmin, double tmax, int32_t *prim) {
    HitInfo best; best.hit = 0;
    *prim = -1;
    for (uint32_t i = 0; i < po->n; ++i) {
        HitInfo h = prim_hit(s, po->type_index[i], o, d, tmin, tmax);
        if (h.hit) { best = h; tmax = h.t; *prim = (int32_t)i; }
    }
    return best;
}

void oracle_raycast_brute(const OScene *scene, const double *rays, int64_t n, double tmin, double tmax,
                          int32_t *prim_out, double *t_out) {
    PrimOrder po = prim_order(scene);
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; ++k) {
        const double *r = rays + 6 * k;
        int32_t prim;
        HitInfo h = closest(scene, &po, from(r), from(r + 3), tmin, tmax, &prim);
        prim_out[k] = prim;
        t_out[k] = h.hit ? h.t : 0.0;
    }
    free(po.type_index);
}

double oracle_prim_hit(const OScene *scene, uint32_t prim, const double *ray, double tmin, double tmax) {
    PrimOrder po = prim_order(scene);
    HitInfo h = prim_hit(scene, po.type_index[prim], from(ray), from(ray + 3), tmin, tmax);
    free(po.type_index);
    return h.hit ? h.t : -1.0;
}

/* ---- materials --------------------------------------------------------------------------- */
void oracle_reflected(const double d[3], const double n[3], double out[3]) {     /* vec3d.h:144-155 */
    V3 dd = from(d), nn = from(n);
    to(out, vsub(dd, vmul(nn, 2 * vdot(dd, nn))));
}
int oracle_refracted(const double ud[3], const double n[3], double ratio, double out[3]) {   /* vec3d.h:168-200 */
    V3 u = from(ud), nn = from(n);
    double cos_theta = fmin(vdot(vneg(u), nn), 1.);
    double sin_theta = sqrt(1 - cos_theta * cos_theta);
    if (ratio * sin_theta > 1) return 0;
    V3 perp = vmul(vadd(u, vmul(nn, cos_theta)), ratio);
    V3 para = vmul(nn, -sqrt(fabs(1 - vmag2(perp))));
    to(out, vadd(perp, para));
    return 1;
}
double oracle_reflectance(double cos_theta, double ratio) {                      /* material.h:175-181 */
    double r0 = (1 - ratio) / (1 + ratio);
    r0 *= r0;
    return r0 + (1 - r0) * pow(1 - cos_theta, 5);
}

static V3 random_unit_vector(uint32_t *rng) {                                    /* vec3d.h:64-75 */
    V3 r;
    do {
        r.x = oracle_rand_double(rng, -1, 1);   /* braced init list: x, y, z in order (vec3d.h:60) */
        r.y = oracle_rand_double(rng, -1, 1);
        r.z = oracle_rand_double(rng, -1, 1);
    } while (!(vmag2(r) < 1));
    return vunit(r);
}

static V3 random_vector_in_unit_disk(uint32_t *rng) {                            /* vec3d.h:79-85 */
    V3 v;
    do {
        v.x = oracle_rand_double(rng, -1, 1);
        v.y = oracle_rand_double(rng, -1, 1);
        v.z = 0;
    } while (!(vmag2(v) < 1));
    return v;
}

/* n draws of the two rejection samplers from one LCG stream (test hooks for the distribution tests of the
 * device's direct samplers; the same two functions feed oracle_render, whose bit-exact match with the reference
 * pins them). */
void oracle_random_unit_vectors(uint32_t *lcg_state, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) to(out + 3 * i, random_unit_vector(lcg_state));
}
void oracle_random_vectors_in_unit_disk(uint32_t *lcg_state, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        V3 v = random_vector_in_unit_disk(lcg_state);
        out[2 * i] = v.x; out[2 * i + 1] = v.y;
    }
}

typedef struct { double r, g, b; } RGBd;
static RGBd rgb(double r, double g, double b) { RGBd c = {r, g, b}; return c; }

typedef struct { const OScene *s; PrimOrder po; RGBd background; uint32_t *rng; uint64_t rays; } RenderCtx;

/* camera.h:205-258 (recursive, as the reference) */
static RGBd ray_color(RenderCtx *cx, V3 o, V3 d, uint64_t depth_left) {
    if (depth_left == 0) return rgb(0, 0, 0);
    int32_t prim;
    cx->rays++;
    HitInfo h = closest(cx->s, &cx->po, o, d, 0.00001, INFINITY, &prim);         /* :217 */
    if (!h.hit) return cx->background;                                           /* :248 */
    const OMaterial *m = &cx->s->materials[h.mat];
    RGBd emitted = rgb(0, 0, 0);                                                 /* material.h:38-40 */
    V3 sd;
    RGBd att;
    switch (m->kind) {
    case 3:                                                                      /* material.h:248-263 */
        return rgb(m->rgb[0] * m->param, m->rgb[1] * m->param, m->rgb[2] * m->param);
    case 0: {                                                                    /* material.h:64-86 */
        sd = vadd(h.n, random_unit_vector(cx->rng));
        if (fabs(sd.x) < 1e-8 && fabs(sd.y) < 1e-8 && fabs(sd.z) < 1e-8) sd = h.n;
        att = rgb(m->rgb[0], m->rgb[1], m->rgb[2]);
        break;
    }
    case 1: {                                                                    /* material.h:116-139 */
        V3 u = vunit(d);
        V3 refl = vsub(u, vmul(h.n, 2 * vdot(u, h.n)));
        double fuzz = fmin(m->param, 1.);                                        /* :150-151 */
        sd = vadd(refl, vmul(random_unit_vector(cx->rng), fuzz));
        if (vdot(h.n, sd) < 0) return emitted;
        att = rgb(m->rgb[0], m->rgb[1], m->rgb[2]);
        break;
    }
    default: {                                                                   /* material.h:185-218 */
        double ratio = h.front ? 1. / m->param : m->param / 1.;
        V3 u = vunit(d);
        double ud[3], nn[3], out[3];
        to(ud, u); to(nn, h.n);
        if (!oracle_refracted(ud, nn, ratio, out)) {
            oracle_reflected(ud, nn, out);
        } else {
            double cos_theta = fmin(vdot(vneg(u), h.n), 1.);
            if (oracle_rand_double(cx->rng, 0, 1) < oracle_reflectance(cos_theta, ratio)) oracle_reflected(ud, nn, out);
        }
        sd = from(out);
        att = rgb(1, 1, 1);
        break;
    }
    }
    RGBd in = ray_color(cx, h.p, sd, depth_left - 1);
    return rgb(emitted.r + att.r * in.r, emitted.g + att.g * in.g, emitted.b + att.b * in.b);   /* :233-234 */
}

/* camera.h:184-200 with the random draws supplied: (vx, vy) = the accepted point in the unit disk (vec3d.h:79-85;
 * ignored for a pinhole camera), r1 / r2 = the U(-0.5, 0.5) factors of delta_x / delta_y.  Used by oracle_render
 * below, so the bit-exact render test pins it against the reference. */
void oracle_camera_ray(const OCamera *c, uint64_t row, uint64_t col, double vx, double vy, double r1, double r2, double out[6]) {
    V3 center = from(c->center), p00 = from(c->pixel00), dx = from(c->delta_x), dy = from(c->delta_y);
    V3 kx = from(c->disk_x), ky = from(c->disk_y);
    V3 o = center;
    if (!(c->defocus_angle <= 0)) o = vadd(vadd(center, vmul(kx, vx)), vmul(ky, vy));        /* camera.h:167 */
    V3 pc = vadd(vadd(p00, vmul(dy, (double)row)), vmul(dx, (double)col));
    V3 ps = vadd(vadd(pc, vmul(dx, r1)), vmul(dy, r2));
    V3 d = vsub(ps, o);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
}

uint64_t oracle_render(const OScene *scene, const OCamera *c, uint32_t *lcg_state, double *out) {
    RenderCtx cx;
    cx.s = scene; cx.po = prim_order(scene); cx.rng = lcg_state; cx.rays = 0;
    cx.background = rgb(c->background[0], c->background[1], c->background[2]);
    for (uint64_t row = 0; row < c->image_h; ++row) {                            /* camera.h:280-293 */
        for (uint64_t col = 0; col < c->image_w; ++col) {
            RGBd px = rgb(0, 0, 0);
            for (uint64_t s = 0; s < c->spp; ++s) {
                V3 v = {0, 0, 0};                                                /* camera.h:184-200 */
                if (!(c->defocus_angle <= 0)) v = random_vector_in_unit_disk(cx.rng);
                /* camera.h:197-198: `pc + rand*dx + rand*dy` -- the two draws are unsequenced in C++;
                 * g++ 13 (the compiler the reference is built with here) evaluates the right-hand
                 * operand first, i.e. the delta_y factor is drawn BEFORE the delta_x factor.  Pinned by
                 * the bit-exact render test. */
                double r2 = oracle_rand_double(cx.rng, -0.5, 0.5);
                double r1 = oracle_rand_double(cx.rng, -0.5, 0.5);
                double ray[6];
                oracle_camera_ray(c, row, col, v.x, v.y, r1, r2, ray);
                V3 o = {ray[0], ray[1], ray[2]}, ps_minus_o = {ray[3], ray[4], ray[5]};
                RGBd col_s = ray_color(&cx, o, ps_minus_o, c->max_depth);
                px.r += col_s.r; px.g += col_s.g; px.b += col_s.b;
            }
            double inv = 1 / (double)c->spp;                                     /* rgb.h:76: /= multiplies by 1/d */
            double *dst = out + (row * c->image_w + col) * 3;
            dst[0] = px.r * inv; dst[1] = px.g * inv; dst[2] = px.b * inv;
        }
    }
    free(cx.po.type_index);
    return cx.rays;
}

/* rgb.h:90-113 (defaults: tone mapping on, gamma 2, max magnitude 255) */
void oracle_tonemap(const double *c, int64_t n, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) {
        double r = c[3 * i], g = c[3 * i + 1], b = c[3 * i + 2];
        double L = 0.2126 * r + 0.7152 * g + 0.0722 * b;
        r /= 1 + L; g /= 1 + L; b /= 1 + L;
        double scale = 255 + 0.999999;
        out[3 * i] = (int)(scale * pow(r, 1 / 2.));
        out[3 * i + 1] = (int)(scale * pow(g, 1 / 2.));
        out[3 * i + 2] = (int)(scale * pow(b, 1 / 2.));
    }
}
