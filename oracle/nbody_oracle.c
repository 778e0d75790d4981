/* nbody_oracle.c — see nbody_oracle.h.  TEST INFRASTRUCTURE ONLY; never linked into the product.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (no -ffast-math: the restatement keeps IEEE semantics;
 * the reference itself is built with -O3 -ffast-math, CMakeLists.txt:128-131, so parity with it is by tolerance). */
#include "nbody_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ ICs */
static float unit_centred(void) /* (rand() - RAND_MAX/2) / (T)(RAND_MAX/2), Bodies.cpp:231-237 */
{
    int k = rand() - RAND_MAX / 2;
    return (float)k / (float)(RAND_MAX / 2);
}
static float unit_flipped(void) /* (RAND_MAX - rand()) / (T)RAND_MAX, Bodies.cpp:185-187 */
{
    int k = RAND_MAX - rand();
    return (float)k / (float)RAND_MAX;
}

int oracle_init_bodies(int scheme, uint64_t n, unsigned seed, float *qx, float *qy, float *qz, float *vx, float *vy,
                       float *vz, float *m, float *r)
{
    if (n == 0 || (scheme != 0 && scheme != 1)) return 1;
    srand(seed);
    for (uint64_t i = 0; i < n; i++) {
        float mi, ri, x, y, z, u, v, w;
        if (scheme == 0) {
            if (i == 0) { /* Bodies.cpp:171-180 */
                mi = 2.0e24f; ri = 0.f; x = y = z = 0.f; u = v = w = 0.f;
            } else { /* Bodies.cpp:181-197; double literals promote, assignment rounds to float */
                mi = (float)((double)((float)rand() / (float)RAND_MAX) * 5e20);
                ri = (float)((double)mi * 2.5e-15);
                float ha = (float)((double)unit_flipped() * 2.0 * M_PI);
                float va = (float)((double)unit_flipped() * 2.0 * M_PI);
                float dc = (float)((double)unit_flipped() * 1.0e8 + 1.0e8);
                x = cosf(va) * sinf(ha) * dc;
                y = sinf(va) * dc;
                z = cosf(va) * cosf(ha) * dc;
                u = (float)((double)y * 4.0e-6);
                v = (float)((double)(-x) * 4.0e-6);
                w = 0.f;
            }
        } else { /* Bodies.cpp:225-241 */
            mi = (float)((double)((float)rand() / (float)RAND_MAX) * 5.0e21);
            ri = (float)((double)mi * 0.5e-14);
            x = (float)((double)unit_centred() * (5.0e8 * 1.33));
            y = (float)((double)unit_centred() * 5.0e8);
            z = (float)((double)unit_centred() * 5.0e8 - 10.0e8);
            u = (float)((double)unit_centred() * 1.0e2);
            v = (float)((double)unit_centred() * 1.0e2);
            w = (float)((double)unit_centred() * 1.0e2);
        }
        if (m) m[i] = mi;
        if (r) r[i] = ri;
        if (qx) qx[i] = x;
        if (qy) qy[i] = y;
        if (qz) qz[i] = z;
        if (vx) vx[i] = u;
        if (vy) vy[i] = v;
        if (vz) vz[i] = w;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ forces */
void oracle_accel_naive_f32(uint64_t n, const float *qx, const float *qy, const float *qz, const float *m, float G,
                            float soft, float *ax, float *ay, float *az)
{
    /* SimulationNBodyNaive.cpp:38-52 */
    for (uint64_t i = 0; i < n; i++) {
        float sx = 0.f, sy = 0.f, sz = 0.f; /* initIteration, :20-27 */
        for (uint64_t j = 0; j < n; j++) {
            const float rx = qx[j] - qx[i];
            const float ry = qy[j] - qy[i];
            const float rz = qz[j] - qz[i];
            const float r2 = rx * rx + ry * ry + rz * rz; /* pow(.,2) x3 */
            const float s2 = soft * soft;
            const float ai = G * m[j] / powf(r2 + s2, 3.f / 2.f);
            sx += ai * rx;
            sy += ai * ry;
            sz += ai * rz;
        }
        ax[i] = sx; ay[i] = sy; az[i] = sz;
    }
}

void oracle_accel_f64(uint64_t n, const float *qx, const float *qy, const float *qz, const float *m, float G, float soft,
                      const uint64_t *idx, uint64_t n_idx, double *ax, double *ay, double *az)
{
    const double Gd = (double)G, s2 = (double)soft * (double)soft;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)n_idx; t++) {
        const uint64_t i = idx ? idx[t] : (uint64_t)t;
        const double xi = qx[i], yi = qy[i], zi = qz[i];
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (uint64_t j = 0; j < n; j++) {
            const double rx = (double)qx[j] - xi, ry = (double)qy[j] - yi, rz = (double)qz[j] - zi;
            const double d = rx * rx + ry * ry + rz * rz + s2;
            const double ai = Gd * (double)m[j] / (d * sqrt(d));
            sx += ai * rx;
            sy += ai * ry;
            sz += ai * rz;
        }
        ax[t] = sx; ay[t] = sy; az[t] = sz;
    }
}

/* ------------------------------------------------------------------------------------------------ integrators */
void oracle_integrate_murb(uint64_t n, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz, const float *ax,
                           const float *ay, const float *az, float dt)
{
    /* Bodies.cpp:264-276 with T = float: `aixDt * 0.5` is a double expression */
    for (uint64_t i = 0; i < n; i++) {
        const float axdt = ax[i] * dt, aydt = ay[i] * dt, azdt = az[i] * dt;
        const float nx = (float)((double)qx[i] + ((double)vx[i] + (double)axdt * 0.5) * (double)dt);
        const float ny = (float)((double)qy[i] + ((double)vy[i] + (double)aydt * 0.5) * (double)dt);
        const float nz = (float)((double)qz[i] + ((double)vz[i] + (double)azdt * 0.5) * (double)dt);
        vx[i] = vx[i] + axdt; vy[i] = vy[i] + aydt; vz[i] = vz[i] + azdt;
        qx[i] = nx; qy[i] = ny; qz[i] = nz;
    }
}

void oracle_run_naive(uint64_t n, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz, const float *m,
                      float G, float soft, float dt, int n_iter, float *ax, float *ay, float *az)
{
    float *bx = ax ? ax : (float *)malloc(n * sizeof(float));
    float *by = ay ? ay : (float *)malloc(n * sizeof(float));
    float *bz = az ? az : (float *)malloc(n * sizeof(float));
    for (int it = 0; it < n_iter; it++) { /* SimulationNBodyNaive.cpp:56-61 */
        oracle_accel_naive_f32(n, qx, qy, qz, m, G, soft, bx, by, bz);
        oracle_integrate_murb(n, qx, qy, qz, vx, vy, vz, bx, by, bz, dt);
    }
    if (!ax) free(bx);
    if (!ay) free(by);
    if (!az) free(bz);
}

static void accel_f64_to_f32(uint64_t n, const float *qx, const float *qy, const float *qz, const float *m, float G,
                             float soft, double *tx, double *ty, double *tz, float *ax, float *ay, float *az)
{
    oracle_accel_f64(n, qx, qy, qz, m, G, soft, NULL, n, tx, ty, tz);
    for (uint64_t i = 0; i < n; i++) { ax[i] = (float)tx[i]; ay[i] = (float)ty[i]; az[i] = (float)tz[i]; }
}

void oracle_run_f64force(uint64_t n, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz, const float *m,
                         float G, float soft, float dt, int integrator, int n_iter)
{
    double *tx = (double *)malloc(3 * n * sizeof(double)), *ty = tx + n, *tz = ty + n;
    float *ax = (float *)malloc(3 * n * sizeof(float)), *ay = ax + n, *az = ay + n;
    if (integrator == 0) {
        for (int it = 0; it < n_iter; it++) {
            accel_f64_to_f32(n, qx, qy, qz, m, G, soft, tx, ty, tz, ax, ay, az);
            oracle_integrate_murb(n, qx, qy, qz, vx, vy, vz, ax, ay, az, dt);
        }
    } else { /* KDK, CUDABodies.cu:172-177 */
        const float hdt = dt * 0.5f;
        accel_f64_to_f32(n, qx, qy, qz, m, G, soft, tx, ty, tz, ax, ay, az);
        for (int it = 0; it < n_iter; it++) {
            for (uint64_t i = 0; i < n; i++) {
                vx[i] = fmaf(ax[i], hdt, vx[i]); vy[i] = fmaf(ay[i], hdt, vy[i]); vz[i] = fmaf(az[i], hdt, vz[i]);
                qx[i] = (float)fma((double)vx[i], (double)dt, (double)qx[i]);
                qy[i] = (float)fma((double)vy[i], (double)dt, (double)qy[i]);
                qz[i] = (float)fma((double)vz[i], (double)dt, (double)qz[i]);
            }
            accel_f64_to_f32(n, qx, qy, qz, m, G, soft, tx, ty, tz, ax, ay, az);
            for (uint64_t i = 0; i < n; i++) {
                vx[i] = fmaf(ax[i], hdt, vx[i]); vy[i] = fmaf(ay[i], hdt, vy[i]); vz[i] = fmaf(az[i], hdt, vz[i]);
            }
        }
    }
    free(tx);
    free(ax);
}

/* ------------------------------------------------------------------------------------------------ energy */
double oracle_energy_f64(uint64_t n, const float *qx, const float *qy, const float *qz, const float *vx, const float *vy,
                         const float *vz, const float *m, float G, float soft)
{
    const double Gd = (double)G, s2 = (double)soft * (double)soft;
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const double xi = qx[i], yi = qy[i], zi = qz[i], mi = m[i];
        double pot = 0.0;
        for (uint64_t j = 0; j < n; j++) {
            if ((uint64_t)i == j) continue; /* the reference adds the self term back, PropertyTracking.cu:298 */
            const double rx = (double)qx[j] - xi, ry = (double)qy[j] - yi, rz = (double)qz[j] - zi;
            pot += Gd * (double)m[j] / sqrt(rx * rx + ry * ry + rz * rz + s2);
        }
        const double v2 = (double)vx[i] * vx[i] + (double)vy[i] * vy[i] + (double)vz[i] * vz[i];
        total += 0.5 * mi * v2 - 0.5 * mi * pot;
    }
    return total;
}

void oracle_metrics_f64(uint64_t n, const float *qx, const float *qy, const float *qz, const float *vx, const float *vy,
                        const float *vz, const float *m, float G, float soft, double *out)
{
    const double Gd = (double)G, s2 = (double)soft * (double)soft;
    double e = 0, lx = 0, ly = 0, lz = 0, M = 0, mx = 0, my = 0, mz = 0, W = 0, wx = 0, wy = 0, wz = 0;
#pragma omp parallel for schedule(static) reduction(+ : e, lx, ly, lz, M, mx, my, mz, W, wx, wy, wz)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const double xi = qx[i], yi = qy[i], zi = qz[i], mi = m[i], ui = vx[i], vi = vy[i], wi = vz[i];
        double pot = 0.0;
        for (uint64_t j = 0; j < n; j++) {
            if ((uint64_t)i == j) continue;
            const double rx = (double)qx[j] - xi, ry = (double)qy[j] - yi, rz = (double)qz[j] - zi;
            pot += Gd * (double)m[j] / sqrt(rx * rx + ry * ry + rz * rz + s2);
        }
        const double w = mi * pot;
        e += 0.5 * mi * (ui * ui + vi * vi + wi * wi) - 0.5 * w;
        lx += mi * (yi * wi - zi * vi);
        ly += mi * (zi * ui - xi * wi);
        lz += mi * (xi * vi - yi * ui);
        M += mi; mx += mi * xi; my += mi * yi; mz += mi * zi;
        W += w; wx += w * xi; wy += w * yi; wz += w * zi;
    }
    out[0] = e; out[1] = lx; out[2] = ly; out[3] = lz; out[4] = M;
    out[5] = M != 0 ? mx / M : 0; out[6] = M != 0 ? my / M : 0; out[7] = M != 0 ? mz / M : 0;
    out[8] = W != 0 ? wx / W : out[5]; out[9] = W != 0 ? wy / W : out[6]; out[10] = W != 0 ? wz / W : out[7];
}
