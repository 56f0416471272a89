/* float32-faithful RK4 advection step -- CPU oracle / timed CPU baseline, TEST INFRASTRUCTURE ONLY.
 *
 * C restatement of what ``pset.execute(parcels.AdvectionRK4, runtime=dt, dt=dt)``
 * (/root/reference/particle_advecter.py:222-223) runs per particle in parcels 2.0.0beta2's
 * JIT-generated C (AdvectionRK4 + parcels.h sampling + spherical unit converters).  parcels is
 * not available here: PARITY UNPINNED -- see oracle/rk4.py for the full statement of the
 * semantics this follows.  Build: gcc -O2 -ffp-contract=off -fopenmp (oracle/rps.py::build_c);
 * no FMA contraction, so every operation is a separately rounded IEEE op like the x86-64
 * baseline build of the JIT code.
 *
 * The per-stage time decisions (cached time index, interpolate-or-hold, float fraction) are
 * identical for all particles (they share one clock) and are computed by the caller
 * (oracle/rk4.py::stage_times).
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const float *u, *v, *lon, *lat;
    int T, Y, X;
} field_t;

/* index i with vals[i] <= x <= vals[i+1]; returns -1 when out of range */
static inline int search_axis(const float *vals, int n, float x, int guess)
{
    if (!(x >= vals[0]) || !(x <= vals[n - 1])) return -1;
    int i = guess;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i < n - 2 && x > vals[i + 1]) ++i;          /* parcels.h: local linear search */
    while (i > 0 && x < vals[i]) --i;
    return i;
}

static inline float bilinear(const float *d, int X, int yi, int xi, double xsi, double eta)
{
    const float *r0 = d + (int64_t)yi * X + xi, *r1 = r0 + X;
    return (float)((1 - xsi) * (1 - eta) * r0[0] + xsi * (1 - eta) * r0[1] + xsi * eta * r1[1] + (1 - xsi) * eta * r1[0]);
}

/* sample converted (u, v) [deg/s] at float position (x, y); returns 0 on success */
static inline int sample_uv(const field_t *f, float x, float y, int ti, int interp, float frac,
                            int *xi, int *yi, float *u, float *v)
{
    *xi = search_axis(f->lon, f->X, x, *xi);
    *yi = search_axis(f->lat, f->Y, y, *yi);
    if (*xi < 0 || *yi < 0) return 1;
    const double xsi = (x - f->lon[*xi]) / (f->lon[*xi + 1] - f->lon[*xi]);     /* float ops, widened */
    const double eta = (y - f->lat[*yi]) / (f->lat[*yi + 1] - f->lat[*yi]);
    const int64_t slab = (int64_t)f->Y * f->X;
    float uu, vv;
    {
        const float f0 = bilinear(f->u + ti * slab, f->X, *yi, *xi, xsi, eta);
        if (interp) {
            const float f1 = bilinear(f->u + (ti + 1) * slab, f->X, *yi, *xi, xsi, eta);
            uu = f0 + (f1 - f0) * frac;
        } else uu = f0;
    }
    {
        const float f0 = bilinear(f->v + ti * slab, f->X, *yi, *xi, xsi, eta);
        if (interp) {
            const float f1 = bilinear(f->v + (ti + 1) * slab, f->X, *yi, *xi, xsi, eta);
            vv = f0 + (f1 - f0) * frac;
        } else vv = f0;
    }
    *u = (float)(uu * (1.0 / (1852. * 60. * cos(y * M_PI / 180))));
    *v = (float)(vv * (1.0 / (1852. * 60.)));
    return 0;
}

int64_t rk4_step_f32(float *lon, float *lat, int64_t n,
                     const float *U, const float *V, const float *glon, const float *glat,
                     int T, int Y, int X,
                     const int *ti, const int *interp, const float *frac, float dt, int threads)
{
    const field_t f = {U, V, glon, glat, T, Y, X};
    int64_t n_oob = 0;
    const float x0 = glon[0], y0 = glat[0];
    const float inv_dx = (float)(X - 1) / (glon[X - 1] - glon[0]);
    const float inv_dy = (float)(Y - 1) / (glat[Y - 1] - glat[0]);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : n_oob)
    for (int64_t p = 0; p < n; ++p) {
        const float plon = lon[p], plat = lat[p];
        /* a fresh index guess per step stands in for the particle's cached xi/yi (the value
           of a sample does not depend on which of two abutting cells a grid-line point lands in) */
        int xi = (int)((plon - x0) * inv_dx), yi = (int)((plat - y0) * inv_dy);
        float u1, v1, u2, v2, u3, v3, u4, v4;
        if (sample_uv(&f, plon, plat, ti[0], interp[0], frac[0], &xi, &yi, &u1, &v1)) { ++n_oob; continue; }
        const float lon1 = plon + u1 * .5 * dt, lat1 = plat + v1 * .5 * dt;
        if (sample_uv(&f, lon1, lat1, ti[1], interp[1], frac[1], &xi, &yi, &u2, &v2)) { ++n_oob; continue; }
        const float lon2 = plon + u2 * .5 * dt, lat2 = plat + v2 * .5 * dt;
        if (sample_uv(&f, lon2, lat2, ti[2], interp[2], frac[2], &xi, &yi, &u3, &v3)) { ++n_oob; continue; }
        const float lon3 = plon + u3 * dt, lat3 = plat + v3 * dt;
        if (sample_uv(&f, lon3, lat3, ti[3], interp[3], frac[3], &xi, &yi, &u4, &v4)) { ++n_oob; continue; }
        lon[p] = plon + (u1 + 2 * u2 + 2 * u3 + u4) / 6. * dt;
        lat[p] = plat + (v1 + 2 * v2 + 2 * v3 + v4) / 6. * dt;
    }
    return n_oob;
}
