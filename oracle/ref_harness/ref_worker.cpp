// ref_worker.cpp -- TEST / BASELINE INFRASTRUCTURE.  The reference's CPU self-play worker path, unmodified sources,
// timed: SPRL::runIteration (selfplay/SelfPlay.hpp:203-248) with SPRL::GridNetwork (networks/GridNetwork.hpp: the traced
// LibTorch module on the CPU) or SPRL::RandomNetwork, and the game's symmetrizer.  This is what `bench.py --impl
// reference` and the cpu_baseline leg run, one single-threaded process per host core like the reference's deployment
// (README.md:124-128).  The stock OTHWorker / C4Worker mains are not used because they hard-code their parameters and
// block on the controller's model files (OTHWorker.cpp:12-28, selfplay/GridWorker.hpp:35-55).
//
//   ref_worker <othello|c4> <model.pt|uniform> <seed> <first_game> <sims> <batch> <queue> <eps> <alpha> <seconds> [<max_games>]
//
// Protocol: one warm-up game through runIteration, discarded; then full games, one
// runIteration(numGames = 1) call each, until `seconds` of wall time have passed (the game in progress is finished and
// counted; at most max_games).  runIteration does not report how many descents it ran (the move loop adds whole batches,
// so a move takes numTraversals .. numTraversals + maxBatchSize - 1 of them): the warm-up game is therefore played a
// second time, untimed, from the same random streams through an instrumented copy of the move loop that counts them;
// the copy must reproduce the warm-up game's sample count, and its descents per move scale the timed moves.
//
// Prints one JSON line: games, moves, sims, evals, seconds (timed games only), sims_per_move.
#include "games/ConnectFourNode.hpp"
#include "games/OthelloNode.hpp"
#include "networks/GridNetwork.hpp"
#include "networks/RandomNetwork.hpp"
#include "selfplay/SelfPlay.hpp"
#include "symmetry/ConnectFourSymmetrizer.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>

extern "C" void sprl_shim_set_stream(uint64_t seed, uint64_t game);

using namespace SPRL;

struct OthelloW {
    using Node = OthelloNode;
    static constexpr int R = OTH_BOARD_WIDTH, C = OTH_BOARD_WIDTH, H = OTH_HISTORY_SIZE, A = OTH_ACTION_SIZE;
    using Sym = D4GridSymmetrizer<OTH_BOARD_WIDTH, OTH_HISTORY_SIZE>;
};
struct C4W {
    using Node = ConnectFourNode;
    static constexpr int R = C4_NUM_ROWS, C = C4_NUM_COLS, H = C4_HISTORY_SIZE, A = C4_ACTION_SIZE;
    using Sym = ConnectFourSymmetrizer;
};

template <class D>
int run(int argc, char** argv) {
    using State = GridState<D::R * D::C, D::H>;
    constexpr int A = D::A;
    const std::string modelPath = argv[2];
    const uint64_t seed = std::strtoull(argv[3], 0, 10), firstGame = std::strtoull(argv[4], 0, 10);
    const int sims = std::atoi(argv[5]), maxBatch = std::atoi(argv[6]), maxQueue = std::atoi(argv[7]);
    const float eps = (float)std::atof(argv[8]), alpha = (float)std::atof(argv[9]);
    const double seconds = std::atof(argv[10]);
    const int maxGames = argc > 11 ? std::atoi(argv[11]) : 1 << 30;

    torch::set_num_threads(1);      // one core per worker process, as deployed
    RandomNetwork<State, A> uniformNet;
    GridNetwork<D::R, D::C, D::H, A> gridNet(modelPath == "uniform" ? "random" : modelPath);
    INetwork<State, A>* net = (modelPath == "uniform") ? (INetwork<State, A>*)&uniformNet : (INetwork<State, A>*)&gridNet;
    typename D::Sym sym;
    const int S = sym.numSymmetries();

    std::ostringstream sink;                              // runIteration reports progress on std::cout
    std::streambuf* coutBuf = std::cout.rdbuf(sink.rdbuf());
    auto playGame = [&](uint64_t game) {                  // one full game through the reference's runIteration
        sprl_shim_set_stream(seed, game);
        auto [states, dists, outcomes] = runIteration<typename D::Node, State, A>(net, 1, sims, maxBatch, maxQueue, eps, alpha,
                                                                                  InitQ::PARENT, &sym, true);
        return (long long)states.size() / S;              // moves
    };

    // ---- warm-up game, then the same game through the instrumented move loop (selfplay/SelfPlay.hpp:82-146)
    const long long warmMoves = playGame(firstGame);
    long long descents = 0, replayMoves = 0;
    {
        sprl_shim_set_stream(seed, firstGame);
        UCTTree<typename D::Node, State, A> tree { std::make_unique<typename D::Node>(), eps, alpha, InitQ::PARENT, &sym, true };
        while (!tree.getDecisionNode()->isTerminal()) {
            int trav = 0;
            while (trav < sims) {
                auto [leaves, t] = tree.searchAndGetLeaves(maxBatch, maxQueue, net, U_WEIGHT);
                if (!leaves.empty()) tree.evaluateAndBackpropLeaves(leaves, net);
                trav += t;
            }
            descents += trav;
            GameActionDist<A> visits = tree.getDecisionNode()->getEdgeStatistics()->m_numVisits;
            GameActionDist<A> pdf = visits / visits.sum();
            pdf = ((int)replayMoves < EARLY_GAME_CUTOFF) ? pdf.pow(EARLY_GAME_EXP) : pdf.pow(REST_GAME_EXP);
            pdf = pdf / pdf.sum();
            GameActionDist<A> cdf = pdf.cumsum();
            cdf = cdf / cdf[A - 1];
            int action = GetRandom().SampleCDF(std::vector<float>(cdf.begin(), cdf.end()));
            tree.advanceDecision((ActionIdx)action);
            ++replayMoves;
        }
    }
    if (replayMoves != warmMoves) {
        std::cout.rdbuf(coutBuf);
        std::cerr << "the instrumented move loop (" << replayMoves << " moves) diverged from runIteration (" << warmMoves << " moves)\n";
        return 3;
    }
    const double simsPerMove = (double)descents / (double)replayMoves;

    // ---- timed games
    const int evals0 = net->getNumEvals();
    long long games = 0, moves = 0;
    const auto t0 = std::chrono::steady_clock::now();
    double secs = 0.0;
    while (games < maxGames) {
        moves += playGame(firstGame + 1 + (uint64_t)games);
        ++games;
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (secs >= seconds) break;
    }
    std::cout.rdbuf(coutBuf);
    std::cout << "{\"games\": " << games << ", \"moves\": " << moves << ", \"sims\": " << (long long)((double)moves * simsPerMove + 0.5)
              << ", \"evals\": " << net->getNumEvals() - evals0 << ", \"seconds\": " << secs << ", \"sims_per_move\": " << simsPerMove
              << ", \"warmup_moves\": " << warmMoves << "}" << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 11) {
        std::cerr << "usage: ref_worker <othello|c4> <model.pt|uniform> <seed> <first_game> <sims> <batch> <queue> <eps> <alpha> <seconds> [<max_games>]\n";
        return 2;
    }
    const std::string game = argv[1];
    if (game == "othello") return run<OthelloW>(argc, argv);
    if (game == "c4") return run<C4W>(argc, argv);
    std::cerr << "unknown game " << game << "\n";
    return 2;
}
