// ref_worker.cpp -- TEST / BASELINE INFRASTRUCTURE.  The reference's CPU self-play
// worker path, unmodified sources: SPRL::UCTTree (uct/UCTTree.hpp) searching with
// SPRL::GridNetwork (networks/GridNetwork.hpp: the traced LibTorch module on the
// CPU) and SPRL::D4GridSymmetrizer, in the move loop of selfplay/SelfPlay.hpp:82-146,
// timed.  This is what `bench.py --impl reference` and the cpu_baseline leg run,
// one single-threaded process per host core like the reference's deployment
// (README.md:124-128).  The stock OTHWorker main is not used because it hard-codes
// its parameters and blocks on the controller's model files (OTHWorker.cpp:12-28,
// selfplay/GridWorker.hpp:35-55).
//
//   ref_worker othello <model.pt|uniform> <seed> <first_game> <ngames> <sims> <batch> <queue>
//              <eps> <alpha> <max_moves_per_game (0 = full games)>
//
// Prints one JSON line: moves, sims (descents), evals, seconds.
#include "games/OthelloNode.hpp"
#include "networks/GridNetwork.hpp"
#include "networks/RandomNetwork.hpp"
#include "selfplay/SelfPlay.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"

#include <chrono>
#include <cstdlib>
#include <iostream>

extern "C" void sprl_shim_set_stream(uint64_t seed, uint64_t game);

using namespace SPRL;

int main(int argc, char** argv) {
    if (argc != 12 || std::string(argv[1]) != "othello") {
        std::cerr << "usage: ref_worker othello <model.pt|uniform> <seed> <first_game> <ngames> <sims> <batch> <queue> <eps> <alpha> <max_moves>\n";
        return 2;
    }
    using State = GridState<OTH_BOARD_SIZE, OTH_HISTORY_SIZE>;
    constexpr int A = OTH_ACTION_SIZE;
    std::string modelPath = argv[2];
    uint64_t seed = std::strtoull(argv[3], 0, 10), firstGame = std::strtoull(argv[4], 0, 10);
    int nGames = std::atoi(argv[5]), sims = std::atoi(argv[6]), maxBatch = std::atoi(argv[7]), maxQueue = std::atoi(argv[8]);
    float eps = (float)std::atof(argv[9]), alpha = (float)std::atof(argv[10]);
    int maxMoves = std::atoi(argv[11]);

    torch::set_num_threads(1);      // one core per worker process, as deployed
    RandomNetwork<State, A> uniformNet;
    GridNetwork<OTH_BOARD_WIDTH, OTH_BOARD_WIDTH, OTH_HISTORY_SIZE, A> gridNet(modelPath == "uniform" ? "random" : modelPath);
    INetwork<State, A>* net = (modelPath == "uniform") ? (INetwork<State, A>*)&uniformNet : (INetwork<State, A>*)&gridNet;
    D4GridSymmetrizer<OTH_BOARD_WIDTH, OTH_HISTORY_SIZE> sym;

    long long moves = 0, descents = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < nGames; ++g) {
        sprl_shim_set_stream(seed, firstGame + g);
        UCTTree<OthelloNode, State, A> tree { std::make_unique<OthelloNode>(), eps, alpha, InitQ::PARENT, &sym, true };
        int moveCount = 0;
        while (!tree.getDecisionNode()->isTerminal() && (maxMoves == 0 || moveCount < maxMoves)) {
            int trav = 0;
            while (trav < sims) {
                auto [leaves, t] = tree.searchAndGetLeaves(maxBatch, maxQueue, net, U_WEIGHT);
                if (!leaves.empty()) tree.evaluateAndBackpropLeaves(leaves, net);
                trav += t;
            }
            descents += trav;
            GameActionDist<A> visits = tree.getDecisionNode()->getEdgeStatistics()->m_numVisits;
            GameActionDist<A> pdf = visits / visits.sum();
            pdf = (moveCount < EARLY_GAME_CUTOFF) ? pdf.pow(EARLY_GAME_EXP) : pdf.pow(REST_GAME_EXP);
            pdf = pdf / pdf.sum();
            GameActionDist<A> cdf = pdf.cumsum();
            cdf = cdf / cdf[A - 1];
            int action = GetRandom().SampleCDF(std::vector<float>(cdf.begin(), cdf.end()));
            tree.advanceDecision((ActionIdx)action);
            ++moveCount;
            ++moves;
        }
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "{\"moves\": " << moves << ", \"sims\": " << descents << ", \"evals\": " << net->getNumEvals()
              << ", \"seconds\": " << secs << "}" << std::endl;
    return 0;
}
