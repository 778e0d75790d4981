// MUrB glue for `--im gpu+b200`: the host side of the drop-in, compiled against the reference's own headers.
//
//   B200Bodies           : Bodies<float>                     (pattern: CUDABodies,  src/common/core/CUDABodies.hpp:24-65)
//   B200BodiesAllocator  : BodiesAllocatorInterface<float>   (pattern: CUDABodiesAllocator, BodiesAllocator.hpp:33-47)
//   SimulationNBodyB200  : SimulationNBodyInterface<float>   (pattern: SimulationNBodyCUDATileFullDevice.hpp:11-35)
//
// All device work goes through the C ABI in include/b200nb.h; this file holds no CUDA.  The reference CLI only
// instantiates createImplem<float>() (main.cpp:316), so the glue is fp32 only, like gpu+tile+full200k (…200k.cu:210).
//
// Environment (precedent: MURB_HETERO_GPU_FRACTION / MURB_HETERO_MIN_N, SimulationNBodyHetero.cu:217-227):
//   MURB_B200_NGPUS       number of GPUs to shard the targets over (default 1; 0 = all visible)
//   MURB_B200_INTEGRATOR  "murb" (default) or "leapfrog"; the tag gpu+b200+leapfrog selects leapfrog too
//   MURB_B200_METRICS_CSV path: track the total energy (what gpu+tracking does,
//                         SimulationNBodyCUDAPropertyTracking.cu:217-369), |angular momentum| and the density centre
//                         of the state every iteration starts from (row k = state before iteration k, the reference's
//                         convention: computeMetrics() precedes the update, ...PropertyTracking.cu:121-133) into a
//                         SimulationHistory<double> and write it on destruction with the reference's own
//                         SimulationHistory::saveMetricsToCSV (src/common/core/SimulationHistory.cpp:103-122)
//   MURB_B200_HOST_MIRROR 1: copy positions and velocities back into the host SoA after every iteration.  The OpenGL
//                         visualisers keep the raw host pointers they were given once (main.cpp:279-296) and read them
//                         every frame, so this is what makes the tag usable without --nv (24 B/body per iteration)
#ifndef SIMULATION_N_BODY_B200_HPP_
#define SIMULATION_N_BODY_B200_HPP_

#include <array>
#include <memory>
#include <string>
#include <vector>

#include "core/Bodies.hpp"
#include "core/BodiesAllocator.hpp"
#include "core/HistoryTrackingInterface.hpp"
#include "core/SimulationNBodyInterface.hpp"

#include "b200nb.h"

class B200Bodies : public Bodies<float> {
  protected:
    b200nb_ctx *ctx = nullptr;
    float boundG = 0.f, boundSoft = 0.f;
    mutable bool hostCurrent = true; // host SoA mirrors the device state

  public:
    B200Bodies(const unsigned long n, const std::string &scheme = "galaxy", const unsigned long randInit = 0);
    virtual ~B200Bodies();
    B200Bodies(const B200Bodies &) = delete;
    B200Bodies &operator=(const B200Bodies &) = delete;

    // (re)creates the device context for this G / softening and uploads the current host state
    void bind(float G, float soft, int nGpus);
    b200nb_ctx *context();        // binds with defaults if needed (standalone integrator use)
    void invalidateDataSoA();     // device state moved on: next getDataSoA() copies back (CUDABodies.cu:58-61)

    virtual const dataSoA_t<float> &getDataSoA() const;               // lazy D2H (CUDABodies.cu:63-93)
    virtual const std::vector<dataAoS_t<float>> &getDataAoS() const;  // rebuilt from the SoA mirror
    virtual void updatePositionsAndVelocities(const accSoA_t<float> &accelerations, float &dt);
    virtual void updatePositionsAndVelocities(const std::vector<accAoS_t<float>> &accelerations, float &dt);
};

class B200BodiesAllocator : public BodiesAllocatorInterface<float> {
  public:
    B200BodiesAllocator(const unsigned long n, const std::string &scheme = "galaxy", const unsigned long randInit = 0);
    virtual std::unique_ptr<Bodies<float>> allocate_unique() const;
    virtual std::shared_ptr<Bodies<float>> allocate_shared() const;
    virtual ~B200BodiesAllocator() = default;

  private:
    const unsigned long n;
    const std::string scheme; // by value: the reference keeps a reference to the caller's string (BodiesAllocator.hpp:28)
    const unsigned long randInit;
};

// Like gpu+tracking (SimulationNBodyCUDAPropertyTracking.hpp:13) the class is also a history tracker: getHistory()
// returns the reference's own container.  The base is the host-side HistoryTrackingInterface<double> so that the glue
// stays free of CUDA; a caller may hand in any SimulationHistory<double>, in particular the GPUSimulationHistory<double>
// the CLI makes for gpu+tracking (main.cpp:247-251) - syncHistoryToDevice() then refreshes its device mirror.
class SimulationNBodyB200 : public SimulationNBodyInterface<float>, public HistoryTrackingInterface<double> {
  protected:
    std::shared_ptr<B200Bodies> b200Bodies;
    int integrator; // B200NB_INTEGRATOR_*
    int nGpus;
    accSoA_t<float> accSoA;
    bool hostMirror = false;       // refresh the host SoA after every iteration (visualiser hand-off)
    bool tracking = false;         // record the metrics of the state every iteration starts from
    std::string metricsPath;       // empty: no CSV on destruction
    size_t recorded = 0;           // rows of the history filled so far
    void init(const bool leapfrog);
    void recordMetrics();

  public:
    SimulationNBodyB200(const BodiesAllocatorInterface<float> &allocator, const float soft = 0.035f,
                        const bool leapfrog = false);
    // the shape of SimulationNBodyCUDAPropertyTracking's constructor (...PropertyTracking.hpp:26-28): the caller owns
    // the history and tracking is on
    SimulationNBodyB200(const BodiesAllocatorInterface<float> &allocator, std::shared_ptr<SimulationHistory<double>> history,
                        const float soft = 0.035f, const bool leapfrog = false);
    virtual ~SimulationNBodyB200();
    virtual void computeOneIteration();
    const accSoA_t<float> &getAccSoA(); // accelerations of the last force pass (...PropertyTracking.cu:308-319)
    void computeAccelerationsOnly();    // force pass without integration (accuracy tests)
    double computeEnergy();             // fp64 total energy (...PropertyTracking.cu:217-304 definition)
    const char *kernelName() const;
    std::array<double, B200NB_N_METRICS> computeMetrics(); // indices: B200NB_METRIC_* (include/b200nb.h)
    size_t getNumRecorded() const { return recorded; }
    // rows [0, getNumRecorded()) of getHistory(): energy, |L|, density centre
    std::vector<double> getEnergies() const;
    void syncHistoryToDevice();         // GPUSimulationHistory only: copyToDevice(); otherwise a no-op
    void saveMetricsToCSV(const std::string &filePath) const;
};

#endif /* SIMULATION_N_BODY_B200_HPP_ */
