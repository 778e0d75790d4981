// ref_wrap.cpp — a thin extern "C" driver around the UNMODIFIED reference classes, compiled together with the
// reference's own sources (by path, from /root/reference) into oracle/_ref/libmurbref*.so by oracle/build_ref.sh.
// TEST INFRASTRUCTURE ONLY: used to pin the oracle, to generate tests/golden/, and as the CPU baseline
// (bench.py `cpu_baseline` / `--impl reference`).  It contains no reference code, only calls into it:
//   Bodies<float>                (src/common/core/Bodies.hpp)
//   BodiesAllocator<float>       (src/common/core/BodiesAllocator.hpp)
//   SimulationNBody{Naive,Optim,SIMD,OpenMP}<float>  (src/murb/implem/)
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "core/Bodies.hpp"
#include "core/BodiesAllocator.hpp"
#include "core/SimulationNBodyInterface.hpp"
#include "implem/SimulationNBodyNaive.hpp"
#include "implem/SimulationNBodyOpenMP.hpp"
#include "implem/SimulationNBodyOptim.hpp"
#include "implem/SimulationNBodySIMD.hpp"

namespace {

template <class Sim> struct Exposed : public Sim {
    using Sim::Sim;
    void forceOnly()
    {
        this->initIteration();
        this->computeBodiesAcceleration();
    }
};

struct Runner {
    virtual ~Runner() = default;
    virtual SimulationNBodyInterface<float> &sim() = 0;
    virtual void forceOnly() = 0;
    virtual const std::vector<accAoS_t<float>> &acc() = 0;
};
template <class Sim> struct RunnerT : Runner {
    Exposed<Sim> s;
    RunnerT(const BodiesAllocatorInterface<float> &a, float soft) : s(a, soft) {}
    SimulationNBodyInterface<float> &sim() override { return s; }
    void forceOnly() override { s.forceOnly(); }
    const std::vector<accAoS_t<float>> &acc() override { return s.getAccAoS(); }
};

std::unique_ptr<Runner> make(const std::string &tag, const BodiesAllocatorInterface<float> &a, float soft)
{
    if (tag == "cpu+naive") return std::make_unique<RunnerT<SimulationNBodyNaive<float>>>(a, soft);
    if (tag == "cpu+optim") return std::make_unique<RunnerT<SimulationNBodyOptim<float>>>(a, soft);
    if (tag == "cpu+simd") return std::make_unique<RunnerT<SimulationNBodySIMD<float>>>(a, soft);
    if (tag == "cpu+omp") return std::make_unique<RunnerT<SimulationNBodyOpenMP<float>>>(a, soft);
    return nullptr;
}

void copy_state(const Bodies<float> &b, uint64_t n, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz,
                float *m, float *r)
{
    const dataSoA_t<float> &d = b.getDataSoA();
    auto cp = [&](float *dst, const std::vector<float> &src) { if (dst) std::memcpy(dst, src.data(), n * sizeof(float)); };
    cp(qx, d.qx); cp(qy, d.qy); cp(qz, d.qz); cp(vx, d.vx); cp(vy, d.vy); cp(vz, d.vz); cp(m, d.m); cp(r, d.r);
}

} // namespace

extern "C" {

// Bodies<float>(n, scheme, 0): the first n bodies (padding excluded).  Returns the padding count.
int ref_init_bodies(uint64_t n, const char *scheme, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz,
                    float *m, float *r)
{
    const std::string sch(scheme);
    Bodies<float> b(n, sch, 0);
    copy_state(b, n, qx, qy, qz, vx, vy, vz, m, r);
    return (int)b.getPadding();
}

// `iters` x computeOneIteration() of implementation `tag`; state out, accelerations of the last iteration out.
// Returns the wall-clock milliseconds of the iteration loop only (construction excluded), < 0 on a bad tag.
double ref_run(const char *tag, uint64_t n, const char *scheme, float soft, float dt, int iters, float *qx, float *qy,
               float *qz, float *vx, float *vy, float *vz, float *ax, float *ay, float *az)
{
    const std::string sch(scheme);
    BodiesAllocator<float> alloc(n, sch);
    std::unique_ptr<Runner> r = make(tag, alloc, soft);
    if (!r) return -1.0;
    r->sim().setDt(dt);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) r->sim().computeOneIteration();
    const auto t1 = std::chrono::steady_clock::now();
    copy_state(*r->sim().getBodies(), n, qx, qy, qz, vx, vy, vz, nullptr, nullptr);
    if (iters > 0) {
        const auto &a = r->acc();
        for (uint64_t i = 0; i < n; ++i) {
            if (ax) ax[i] = a[i].ax;
            if (ay) ay[i] = a[i].ay;
            if (az) az[i] = a[i].az;
        }
    }
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// computeBodiesAcceleration() only, `reps` times on the initial positions; returns ms per force pass.
double ref_accel(const char *tag, uint64_t n, const char *scheme, float soft, int reps, float *ax, float *ay, float *az)
{
    const std::string sch(scheme);
    BodiesAllocator<float> alloc(n, sch);
    std::unique_ptr<Runner> r = make(tag, alloc, soft);
    if (!r || reps < 1) return -1.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) r->forceOnly();
    const auto t1 = std::chrono::steady_clock::now();
    const auto &a = r->acc();
    for (uint64_t i = 0; i < n; ++i) {
        if (ax) ax[i] = a[i].ax;
        if (ay) ay[i] = a[i].ay;
        if (az) az[i] = a[i].az;
    }
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
}

// Bodies<float>::updatePositionsAndVelocities(accSoA, dt) `iters` times with fixed accelerations
// (the shape of test_CUDABodies.cpp:42-75).
int ref_integrate(uint64_t n, const char *scheme, const float *ax, const float *ay, const float *az, float dt, int iters,
                  float *qx, float *qy, float *qz, float *vx, float *vy, float *vz)
{
    const std::string sch(scheme);
    Bodies<float> b(n, sch, 0);
    accSoA_t<float> acc;
    acc.ax.assign(ax, ax + n); acc.ay.assign(ay, ay + n); acc.az.assign(az, az + n);
    for (int i = 0; i < iters; ++i) b.updatePositionsAndVelocities(acc, dt);
    copy_state(b, n, qx, qy, qz, vx, vy, vz, nullptr, nullptr);
    return 0;
}

} // extern "C"
