// worker_main.hpp -- shared main() body of the worker executables (OTHWorker, C4Worker,
// GoWorker).  Same command line and file contract as the reference's workers
// (cpp/src/OTHWorker.cpp:31-69): `<exe> <task_id> <num_tasks>`, models read from
// data/models/<run>/traced_<run>_iteration_<i>.pt, samples written to
// data/games/<run>/<group>/<task>/<run>_iteration_<i>_{states,distributions,outcomes}.npy.
// The reference fixes its parameters as compile-time constants per executable; the same
// defaults are used here and may be overridden through SPRL_* environment variables, which
// is how a deployment sizes one GPU worker to replace many CPU worker tasks.
#ifndef SPRL_B200_WORKER_MAIN_HPP
#define SPRL_B200_WORKER_MAIN_HPP

#include "networks/GridNetwork.hpp"
#include "selfplay/GridWorker.hpp"

#include <cstdlib>
#include <iostream>
#include <string>

namespace SPRL {

struct WorkerDefaults {
    const char* runName;
    int numGroups, numWorkerTasks, numIters;
    int initGames, initTraversals, initBatch, initQueue;
    int games, traversals, batch, queue;
    float dirEps, dirAlpha;
};

inline int envInt(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : fallback;
}
inline float envFloat(const char* name, float fallback) {
    const char* v = std::getenv(name);
    return v ? (float)std::atof(v) : fallback;
}

template <typename ImplNode, typename Symmetrizer, int ROWS, int COLS, int HISTORY, int ACTIONS>
int workerMain(int argc, char* argv[], const WorkerDefaults& d, const char* exeName) {
    if (argc != 3) {
        std::cerr << "Usage: ./" << exeName << " <task_id> <num_tasks>" << std::endl;
        return 1;
    }
    const char* rn = std::getenv("SPRL_RUN_NAME");
    std::string runName = rn ? rn : d.runName;
    int myTaskId = std::stoi(argv[1]);
    int numTasks = std::stoi(argv[2]);
    int numGroups = envInt("SPRL_NUM_GROUPS", d.numGroups);
    if (numTasks < numGroups || numTasks <= 0 || myTaskId < 0 || myTaskId >= numTasks) {
        std::cerr << "task id / task count out of range" << std::endl;
        return 1;
    }
    int myGroup = myTaskId / (numTasks / numGroups);
    std::cout << "Task " << myTaskId << " of " << numTasks << ", in group " << myGroup << "." << std::endl;
    std::string saveDir = "data/games/" + runName + "/" + std::to_string(myGroup) + "/" + std::to_string(myTaskId);

    DeviceOptions& opt = deviceOptions();
    opt.device = envInt("SPRL_DEVICE", 0);
    // opt.seed: SPRL_SEED, else drawn from std::random_device and logged (deviceOptions())
    opt.numSlots = envInt("SPRL_NUM_SLOTS", 0);
    opt.fixSymmetryMask = envInt("SPRL_FIX_SYMMETRY_MASK", 0) != 0;      // default: the reference's behaviour (quirk Q3)
    // task t plays stream ids t, t + numTasks, ...: the tasks of a run never share a game stream
    opt.firstGame = (uint64_t)myTaskId;
    opt.gameStride = (uint64_t)numTasks;

    using State = GridState<ROWS * COLS, HISTORY>;
    RandomNetwork<State, ACTIONS> randomNetwork {};
    Symmetrizer symmetrizer {};
    try {
        runWorker<GridNetwork<ROWS, COLS, HISTORY, ACTIONS>, ImplNode, ROWS, COLS, HISTORY, ACTIONS>(
            runName, saveDir, &randomNetwork, &symmetrizer,
            envInt("SPRL_NUM_ITERS", d.numIters),
            envInt("SPRL_INIT_NUM_GAMES", d.initGames), envInt("SPRL_INIT_UCT_TRAVERSALS", d.initTraversals),
            envInt("SPRL_INIT_MAX_BATCH_SIZE", d.initBatch), envInt("SPRL_INIT_MAX_QUEUE_SIZE", d.initQueue),
            envInt("SPRL_NUM_GAMES", d.games), envInt("SPRL_UCT_TRAVERSALS", d.traversals),
            envInt("SPRL_MAX_BATCH_SIZE", d.batch), envInt("SPRL_MAX_QUEUE_SIZE", d.queue),
            envFloat("SPRL_DIRICHLET_EPSILON", d.dirEps), envFloat("SPRL_DIRICHLET_ALPHA", d.dirAlpha));
    } catch (const std::exception& e) {
        std::cerr << exeName << ": " << e.what() << std::endl;
        return 2;
    }
    return 0;
}

}  // namespace SPRL
#endif
