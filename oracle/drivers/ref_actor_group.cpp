// The reference's own ActorGroup (actor/actor_group.{h,cpp}), unmodified, with the network
// device made selectable (stock createNeuralNetworks() insists on >= 1 CUDA device,
// actor_group.cpp:168-177; it is a protected virtual, actor_group.h:55, so this subclass is the
// only change). TEST / BASELINE INFRASTRUCTURE ONLY.
//
//   ref_actor_group sp    <conf_str> [device=-1]                -> speaks the zero-server wire protocol on stdin/stdout
//   ref_actor_group bench <conf_str> <warmup_cycles> <cycles> [device=-1]
//        -> runs whole ActorGroup cycles (one simulation for every game) and prints
//           "REFBENCH evals=<n> seconds=<t> threads=<zero_num_threads> device=<d>"
#include "actor_group.h"
#include "configuration.h"
#include "configure_loader.h"
#include "create_network.h"
#include "environment.h"
#include "random.h"
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>

using namespace minizero;
using namespace minizero::actor;

class SelectableDeviceActorGroup : public ActorGroup {
public:
    explicit SelectableDeviceActorGroup(int device) : device_(device) {}

    double runCycles(long cycles)
    {
        auto t0 = std::chrono::steady_clock::now();
        for (long i = 0; i < 2 * cycles; ++i) { // CPU phase + GPU phase per cycle (actor_group.cpp:139-147)
            getSharedData()->actor_index_ = 0;
            for (auto& t : slave_threads_) { t->start(); }
            for (auto& t : slave_threads_) { t->finish(); }
            getSharedData()->do_cpu_job_ = !getSharedData()->do_cpu_job_;
        }
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

protected:
    void createNeuralNetworks() override
    {
        getSharedData()->networks_.resize(1);
        getSharedData()->network_outputs_.resize(1);
        getSharedData()->networks_[0] = network::createNetwork(config::nn_file_name, device_);
    }
    int device_;
};

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::cerr << "usage: ref_actor_group sp|bench <conf_str> ..." << std::endl;
        return 2;
    }
    const std::string mode = argv[1];
    env::setUpEnv();
    config::ConfigureLoader cl;
    config::setConfiguration(cl);
    if (!cl.loadFromString(argv[2])) { return 1; }
    utils::Random::seed(config::program_seed);

    if (mode == "sp") {
        SelectableDeviceActorGroup ag(argc > 3 ? atoi(argv[3]) : -1);
        ag.run();
    } else if (mode == "bench") {
        if (argc < 5) { return 2; }
        const long warm = atol(argv[3]), cycles = atol(argv[4]);
        const int device = (argc > 5 ? atoi(argv[5]) : -1);
        SelectableDeviceActorGroup ag(device);
        ag.initialize();
        ag.runCycles(warm);
        double t = ag.runCycles(cycles);
        std::cout << "REFBENCH evals=" << cycles * config::zero_num_parallel_games << " seconds=" << t
                  << " threads=" << config::zero_num_threads << " device=" << device << std::endl;
        std::cout.flush();
        _exit(0); // slave threads never terminate (paralleler.h:24-32)
    }
    return 0;
}
