/* oracle/pt_oracle.h -- TEST INFRASTRUCTURE ONLY (see pt_oracle.c). */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stdint.h>

#pragma pack(push, 1)
typedef struct { uint32_t kind, pad; double rgb[3]; double param; } OMaterial;   /* 0 lambertian 1 metal 2 dielectric 3 light */
typedef struct { double c[3]; double r; uint32_t mat, prim; } OSphere;
typedef struct { double v[3], s1[3], s2[3]; uint32_t mat, prim; } OQuad;
typedef struct {
    uint64_t image_w, image_h, spp, max_depth;
    double center[3], dir[3], up[3];
    double focus_dist, defocus_angle, vfov, hfov;
    double background[3];
    double pixel00[3], delta_x[3], delta_y[3], disk_x[3], disk_y[3];
} OCamera;
#pragma pack(pop)

typedef struct {
    uint64_t n_materials, n_spheres, n_quads;
    const OMaterial *materials;
    const OSphere *spheres;
    const OQuad *quads;
} OScene;

/* rand_util.h:51-117 */
uint32_t oracle_seed_sequence_next(uint32_t *seed_state);
double oracle_rand_double(uint32_t *lcg_state, double min, double max);

/* camera.h:87-157 */
void oracle_camera_init(OCamera *cam);

/* scene.h:59-75 over the canonical primitive order (sphere.h:45-96, parallelogram.h:177-240) */
void oracle_raycast_brute(const OScene *scene, const double *rays, int64_t n, double tmin, double tmax,
                          int32_t *prim_out, double *t_out);
/* hit time of ONE primitive (canonical index) or -1 */
double oracle_prim_hit(const OScene *scene, uint32_t prim, const double *ray, double tmin, double tmax);

/* vec3d.h:144-200, material.h:175-181 */
void oracle_reflected(const double d[3], const double n[3], double out[3]);
int oracle_refracted(const double unit_d[3], const double n[3], double ratio, double out[3]);
double oracle_reflectance(double cos_theta, double ratio);

/* camera.h:184-297 single-threaded, starting from the given per-thread LCG state; returns the
 * number of hit_by calls; writes image_h*image_w*3 doubles (linear HDR) and the final state. */
void oracle_camera_ray(const OCamera *cam, uint64_t row, uint64_t col, double vx, double vy, double r1, double r2, double out[6]);
uint64_t oracle_render(const OScene *scene, const OCamera *cam, uint32_t *lcg_state, double *out_rgb);

/* vec3d.h:64-85: n draws of random_unit_vector() (n x 3) / random_vector_in_unit_disk() (n x 2) */
void oracle_random_unit_vectors(uint32_t *lcg_state, int64_t n, double *out);
void oracle_random_vectors_in_unit_disk(uint32_t *lcg_state, int64_t n, double *out);

/* rgb.h:90-113 */
void oracle_tonemap(const double *rgb, int64_t n_pixels, int32_t *out);

#endif
