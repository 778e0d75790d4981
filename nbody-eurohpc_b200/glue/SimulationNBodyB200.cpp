#include "SimulationNBodyB200.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <stdexcept>

#include "b200nb.h"

namespace {

// The reference's device variants print and exit on a CUDA error (CUDA_CHECK,
// SimulationNBodyCUDATileFullDevice.cu:10-17); the library only returns codes, the glue keeps the CLI behaviour.
void check(int rc, const b200nb_ctx *ctx, const char *what)
{
    if (rc == B200NB_OK) return;
    std::fprintf(stderr, "gpu+b200: %s failed (%d): %s\n", what, rc, b200nb_last_error(ctx));
    std::exit(rc);
}

int envInt(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

} // namespace

// ================================================================================================ B200Bodies
B200Bodies::B200Bodies(const unsigned long n, const std::string &scheme, const unsigned long randInit)
    : Bodies<float>(n, scheme, randInit)
{
}

B200Bodies::~B200Bodies() { b200nb_destroy(ctx); }

void B200Bodies::bind(float G, float soft, int nGpus)
{
    if (ctx && G == boundG && soft == boundSoft) return;
    this->getDataSoA(); // pull the state back before the old context goes away
    b200nb_destroy(ctx);
    ctx = nullptr;
    check(b200nb_create(&ctx, this->n, nGpus, G, soft), nullptr, "b200nb_create");
    const dataSoA_t<float> &d = this->dataSoA;
    check(b200nb_upload(ctx, d.qx.data(), d.qy.data(), d.qz.data(), d.m.data(), d.vx.data(), d.vy.data(), d.vz.data()),
          ctx, "b200nb_upload");
    boundG = G;
    boundSoft = soft;
    hostCurrent = true;
    this->allocatedBytes += (float)b200nb_allocated_bytes(ctx);
}

b200nb_ctx *B200Bodies::context()
{
    if (!ctx) bind(6.67384e-11f, 1.0f, envInt("MURB_B200_NGPUS", 1)); // integrator-only use: softening is irrelevant
    return ctx;
}

void B200Bodies::invalidateDataSoA() { hostCurrent = false; }

const dataSoA_t<float> &B200Bodies::getDataSoA() const
{
    if (!hostCurrent && ctx) {
        // mass, radius and the padding tail never change on the device: only n positions / velocities come back
        dataSoA_t<float> &d = this->dataSoA;
        check(b200nb_download_state(ctx, d.qx.data(), d.qy.data(), d.qz.data(), d.vx.data(), d.vy.data(), d.vz.data()),
              ctx, "b200nb_download_state");
        hostCurrent = true;
    }
    return this->dataSoA;
}

const std::vector<dataAoS_t<float>> &B200Bodies::getDataAoS() const
{
    const dataSoA_t<float> &d = this->getDataSoA();
    for (unsigned long i = 0; i < this->n; i++) {
        dataAoS_t<float> &b = this->dataAoS[i];
        b.qx = d.qx[i]; b.qy = d.qy[i]; b.qz = d.qz[i];
        b.vx = d.vx[i]; b.vy = d.vy[i]; b.vz = d.vz[i];
    }
    return this->dataAoS;
}

void B200Bodies::updatePositionsAndVelocities(const accSoA_t<float> &accelerations, float &dt)
{
    b200nb_ctx *c = this->context();
    check(b200nb_integrate_host_accel(c, accelerations.ax.data(), accelerations.ay.data(), accelerations.az.data(), dt),
          c, "b200nb_integrate_host_accel");
    this->invalidateDataSoA();
}

void B200Bodies::updatePositionsAndVelocities(const std::vector<accAoS_t<float>> &accelerations, float &dt)
{
    accSoA_t<float> soa;
    soa.ax.resize(this->n); soa.ay.resize(this->n); soa.az.resize(this->n);
    for (unsigned long i = 0; i < this->n; i++) {
        soa.ax[i] = accelerations[i].ax; soa.ay[i] = accelerations[i].ay; soa.az[i] = accelerations[i].az;
    }
    this->updatePositionsAndVelocities(soa, dt);
}

// ================================================================================================ allocator
B200BodiesAllocator::B200BodiesAllocator(const unsigned long n, const std::string &scheme, const unsigned long randInit)
    : n{n}, scheme{scheme}, randInit{randInit}
{
}

std::unique_ptr<Bodies<float>> B200BodiesAllocator::allocate_unique() const
{
    return std::make_unique<B200Bodies>(n, scheme, randInit);
}

std::shared_ptr<Bodies<float>> B200BodiesAllocator::allocate_shared() const
{
    return std::make_shared<B200Bodies>(n, scheme, randInit);
}

// ================================================================================================ simulation
SimulationNBodyB200::SimulationNBodyB200(const BodiesAllocatorInterface<float> &allocator, const float soft,
                                         const bool leapfrog)
    : SimulationNBodyInterface<float>(allocator, soft),
      HistoryTrackingInterface<double>(std::make_shared<SimulationHistory<double>>(0))
{
    this->init(leapfrog);
}

SimulationNBodyB200::SimulationNBodyB200(const BodiesAllocatorInterface<float> &allocator,
                                         std::shared_ptr<SimulationHistory<double>> history, const float soft,
                                         const bool leapfrog)
    : SimulationNBodyInterface<float>(allocator, soft), HistoryTrackingInterface<double>(history)
{
    if (!this->history) {
        std::fprintf(stderr, "gpu+b200: the history must not be null\n");
        std::exit(-1);
    }
    this->tracking = true;
    this->init(leapfrog);
}

void SimulationNBodyB200::init(const bool leapfrog)
{
    this->integrator = leapfrog ? B200NB_INTEGRATOR_LEAPFROG : B200NB_INTEGRATOR_MURB;
    // never touch this->allocator after the base constructor: the CLI passes a stack-local one (main.cpp:210,238)
    const float n = (float)this->getBodies()->getN();
    this->flopsPerIte = 20.f * n * n; // SimulationNBodyNaive.cpp:15
    this->b200Bodies = std::dynamic_pointer_cast<B200Bodies>(this->bodies);
    if (!this->b200Bodies) {
        std::fprintf(stderr, "gpu+b200: the allocator must be a B200BodiesAllocator\n");
        std::exit(-1);
    }
    const char *integ = std::getenv("MURB_B200_INTEGRATOR");
    if (integ && !std::strcmp(integ, "leapfrog")) this->integrator = B200NB_INTEGRATOR_LEAPFROG;
    this->nGpus = envInt("MURB_B200_NGPUS", 1);
    this->b200Bodies->bind(this->G, this->soft, this->nGpus);
    this->allocatedBytes = this->bodies->getAllocatedBytes();
    const char *csv = std::getenv("MURB_B200_METRICS_CSV");
    if (csv && *csv) {
        this->metricsPath = csv;
        this->tracking = true;
    }
    this->hostMirror = envInt("MURB_B200_HOST_MIRROR", 0) != 0;
}

SimulationNBodyB200::~SimulationNBodyB200()
{
    if (!this->metricsPath.empty()) this->saveMetricsToCSV(this->metricsPath);
}

void SimulationNBodyB200::saveMetricsToCSV(const std::string &filePath) const
{
    // the reference's own writer (SimulationHistory.cpp:103-122), on exactly the rows recorded so far.  Upstream only
    // ever fills in the energy; |L| and the density centre carry the definitions of include/b200nb.h
    SimulationHistory<double> rows((int)this->recorded);
    for (size_t i = 0; i < this->recorded; i++) {
        rows.setEnergyAt((int)i, this->history->getEnergyAt((int)i));
        rows.setAngMomentumAt((int)i, this->history->getAngMomentumAt((int)i));
        rows.setDensityCenterAt((int)i, this->history->getDensityCenterAt((int)i));
    }
    try {
        rows.saveMetricsToCSV(filePath);
    } catch (const std::exception &e) { // the reference throws when the file cannot be opened
        std::fprintf(stderr, "gpu+b200: %s\n", e.what());
    }
}

std::vector<double> SimulationNBodyB200::getEnergies() const
{
    const std::vector<double> &all = this->history->getAllEnergy();
    return std::vector<double>(all.begin(), all.begin() + (long)this->recorded);
}

void SimulationNBodyB200::syncHistoryToDevice()
{
#ifdef USE_CUDA
    if (auto gpu = std::dynamic_pointer_cast<GPUSimulationHistory<double>>(this->history)) gpu->copyToDevice();
#endif
}

void SimulationNBodyB200::computeOneIteration()
{
    b200nb_ctx *c = this->b200Bodies->context();
    // like gpu+tracking, the metrics belong to the state the iteration starts from (...PropertyTracking.cu:121-133)
    if (this->tracking) this->recordMetrics();
    check(b200nb_step(c, this->dt, this->integrator, 1), c, "b200nb_step");
    // main.cpp:353-371 joins the current device only; with several GPUs the others are joined here
    if (b200nb_n_local_gpus(c) > 1) check(b200nb_sync(c), c, "b200nb_sync");
    this->b200Bodies->invalidateDataSoA();
    if (this->hostMirror) (void)this->b200Bodies->getDataSoA(); // in place: the vectors never move, captured pointers stay valid
}

void SimulationNBodyB200::recordMetrics()
{
    const std::array<double, B200NB_N_METRICS> m = this->computeMetrics();
    // a caller-sized history (NIterations rows, main.cpp:247-248) is filled in place; otherwise it grows geometrically
    // (GPUSimulationHistory reallocates its device mirror on every resize)
    if ((int)this->recorded >= this->history->getNumIterations())
        this->history->setNumIterations(std::max<int>(16, 2 * this->history->getNumIterations()));
    const int k = (int)this->recorded++;
    this->history->setEnergyAt(k, m[B200NB_METRIC_ENERGY]);
    this->history->setAngMomentumAt(k, std::sqrt(m[B200NB_METRIC_ANG_X] * m[B200NB_METRIC_ANG_X] +
                                              m[B200NB_METRIC_ANG_Y] * m[B200NB_METRIC_ANG_Y] +
                                              m[B200NB_METRIC_ANG_Z] * m[B200NB_METRIC_ANG_Z]));
    this->history->setDensityCenterAt(k, {m[B200NB_METRIC_DENSITY_X], m[B200NB_METRIC_DENSITY_Y], m[B200NB_METRIC_DENSITY_Z]});
}

std::array<double, B200NB_N_METRICS> SimulationNBodyB200::computeMetrics()
{
    b200nb_ctx *c = this->b200Bodies->context();
    std::array<double, B200NB_N_METRICS> m{};
    check(b200nb_metrics(c, m.data()), c, "b200nb_metrics");
    return m;
}

void SimulationNBodyB200::computeAccelerationsOnly()
{
    b200nb_ctx *c = this->b200Bodies->context();
    check(b200nb_accel(c), c, "b200nb_accel");
}

const accSoA_t<float> &SimulationNBodyB200::getAccSoA()
{
    b200nb_ctx *c = this->b200Bodies->context();
    const unsigned long n = this->getBodies()->getN();
    accSoA.ax.resize(n); accSoA.ay.resize(n); accSoA.az.resize(n);
    check(b200nb_download_accel(c, accSoA.ax.data(), accSoA.ay.data(), accSoA.az.data()), c, "b200nb_download_accel");
    return accSoA;
}

double SimulationNBodyB200::computeEnergy()
{
    b200nb_ctx *c = this->b200Bodies->context();
    double e = 0.0;
    check(b200nb_energy(c, &e), c, "b200nb_energy");
    return e;
}

const char *SimulationNBodyB200::kernelName() const { return b200nb_kernel_name(this->b200Bodies->context()); }
