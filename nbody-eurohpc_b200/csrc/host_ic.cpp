// Host-side initial conditions: a restatement of Bodies<float>::initGalaxy / initRandomly
// (reference: src/common/core/Bodies.cpp:158-214 and :217-257) for drivers that do not link the reference
// (bench.py, the ctypes tests).  The MUrB glue does NOT use this: there the reference's own Bodies<float> builds
// the bodies and hands them to b200nb_upload().
//
// Bit-identity with the reference matters (murb-test compares iteration 0 with eps = 0,
// src/test/implem/test_SimulationNBody.cpp:63), so every expression keeps the reference's evaluation types for
// T = float: int -> float conversion of rand(), float division by (float)RAND_MAX, promotion to double wherever a
// double literal takes part, one rounding back to float on assignment, and the float overloads of sin/cos.
// tests/test_ic_parity.py pins it against the reference's generator (golden fixtures + live when available).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/b200nb.h"

namespace {

struct BodyOut {
    float *qx, *qy, *qz, *vx, *vy, *vz, *m, *r;
    void set(uint64_t i, float mi, float ri, float x, float y, float z, float u, float v, float w) const
    {
        if (m) m[i] = mi;
        if (r) r[i] = ri;
        if (qx) qx[i] = x;
        if (qy) qy[i] = y;
        if (qz) qz[i] = z;
        if (vx) vx[i] = u;
        if (vy) vy[i] = v;
        if (vz) vz[i] = w;
    }
};

// uniform in [-1, 1): (rand() - RAND_MAX/2) / (float)(RAND_MAX/2), integer subtraction first
inline float centred_unit()
{
    const int k = rand() - RAND_MAX / 2;
    return (float)k / (float)(RAND_MAX / 2);
}
// uniform in (0, 1]: (RAND_MAX - rand()) / (float)RAND_MAX
inline float flipped_unit()
{
    const int k = RAND_MAX - rand();
    return (float)k / (float)RAND_MAX;
}

// a body of the "random" box; also what the reference puts in its padding zone (6 rand() calls)
inline void random_box_body(float &x, float &y, float &z, float &u, float &v, float &w)
{
    x = (float)((double)centred_unit() * (5.0e8 * 1.33));
    y = (float)((double)centred_unit() * 5.0e8);
    z = (float)((double)centred_unit() * 5.0e8 - 10.0e8);
    u = (float)((double)centred_unit() * 1.0e2);
    v = (float)((double)centred_unit() * 1.0e2);
    w = (float)((double)centred_unit() * 1.0e2);
}

void galaxy(uint64_t n, unsigned seed, const BodyOut &o)
{
    srand(seed);
    o.set(0, 2.0e24f, 0.0f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); // the central mass, no rand() consumed
    for (uint64_t i = 1; i < n; ++i) {
        const float mi = (float)((double)((float)rand() / (float)RAND_MAX) * 5e20);
        const float ri = (float)((double)mi * 2.5e-15);
        const float h = (float)((double)flipped_unit() * 2.0 * M_PI); // horizontal angle
        const float v = (float)((double)flipped_unit() * 2.0 * M_PI); // vertical angle
        const float d = (float)((double)flipped_unit() * 1.0e8 + 1.0e8);
        const float x = std::cos(v) * std::sin(h) * d;
        const float y = std::sin(v) * d;
        const float z = std::cos(v) * std::cos(h) * d;
        const float u = (float)((double)y * 4.0e-6);
        const float w = (float)((double)(-x) * 4.0e-6);
        o.set(i, mi, ri, x, y, z, u, w, 0.0f);
    }
}

void random_box(uint64_t n, unsigned seed, const BodyOut &o)
{
    srand(seed);
    for (uint64_t i = 0; i < n; ++i) {
        const float mi = (float)((double)((float)rand() / (float)RAND_MAX) * 5.0e21);
        const float ri = (float)((double)mi * 0.5e-14);
        float x, y, z, u, v, w;
        random_box_body(x, y, z, u, v, w);
        o.set(i, mi, ri, x, y, z, u, v, w);
    }
}

} // namespace

extern "C" int b200nb_init_bodies(int scheme, uint64_t n, unsigned seed, float *qx, float *qy, float *qz, float *vx,
                                  float *vy, float *vz, float *m, float *r)
{
    if (n == 0) return B200NB_EINVAL;
    const BodyOut o{qx, qy, qz, vx, vy, vz, m, r};
    if (scheme == 0) galaxy(n, seed, o);
    else if (scheme == 1) random_box(n, seed, o);
    else return B200NB_EINVAL;
    return B200NB_OK;
}

// ---------------------------------------------------------------------------------------------- .tab loader
// Bodies<float>::initMilkyWayAndromeda (src/common/core/Bodies.cpp:82-153): one body per non-empty line,
// "mass x y z vx vy vz" in galaxy units, rescaled per component: the disk (16384), bulge (8192) and halo (16384)
// bodies of the Milky Way come first inside each component pair and use (4.5e10, 4.0, 220), Andromeda's use
// (9.4e10, 6.0, 260); radius is 1e5 for every body.  Values are parsed as float with operator>>, like the reference.
extern "C" int b200nb_tab_count(const char *path, uint64_t *n_bodies)
{
    if (!path || !n_bodies) return B200NB_EINVAL;
    std::ifstream file(path);
    if (!file.is_open()) return B200NB_EINVAL;
    uint64_t n = 0;
    std::string line;
    while (std::getline(file, line))
        if (!line.empty()) ++n;
    *n_bodies = n;
    return B200NB_OK;
}

extern "C" int b200nb_tab_load(const char *path, uint64_t n, float *qx, float *qy, float *qz, float *vx, float *vy,
                               float *vz, float *m, float *r)
{
    if (!path) return B200NB_EINVAL;
    std::ifstream file(path);
    if (!file.is_open()) return B200NB_EINVAL;
    const BodyOut o{qx, qy, qz, vx, vy, vz, m, r};
    const uint64_t disk = 16384, bulge = 8192, halo = 16384;
    uint64_t i = 0;
    std::string line;
    while (i < n && std::getline(file, line)) {
        if (line.empty()) continue;
        std::istringstream iss(line);
        float mi, x, y, z, u, v, w;
        iss >> mi >> x >> y >> z >> u >> v >> w;
        if (iss.fail()) return B200NB_EINVAL;
        const bool milky_way = i < disk || (i >= 2 * disk && i < 2 * disk + bulge) ||
                               (i >= 2 * (disk + bulge) && i < 2 * (disk + bulge) + halo);
        const double sm = milky_way ? 4.5e10 : 9.4e10, sq = milky_way ? 4.0 : 6.0;
        const int sv = milky_way ? 220 : 260;
        mi = (float)((double)mi * sm);
        x = (float)((double)x * sq); y = (float)((double)y * sq); z = (float)((double)z * sq);
        u = u * (float)sv; v = v * (float)sv; w = w * (float)sv;
        o.set(i, mi, 1e5f, x, y, z, u, v, w);
        ++i;
    }
    return i == n ? B200NB_OK : B200NB_EINVAL;
}
