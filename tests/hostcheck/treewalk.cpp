// treewalk.cpp -- TEST PROGRAM for the C++ veneer's UCTTree (sprl_b200/host/include/sprl/veneer.hpp), written the way a
// user of the reference's uct/UCTTree.hpp:62-210 writes a move loop: searchAndGetLeaves / evaluateAndBackpropLeaves until
// the budget is spent, read getDecisionNode()'s statistics, advanceDecision(first most-visited action).  The same
// protocol as oracle/ref_harness/ref_trace.cpp `treewalk`, so its output is compared with the golden fixtures
// tests/golden/treewalk_*.npz (tests/test_host_gpu.py).
//
//   treewalk <othello|c4|go> <seed> <first_game> <ngames> <sims> <batch> <queue> <eps> <alpha> <noise> <sym> <parent|zero|drop> <out.bin>
//
// out.bin: int32 ngames, int32 A, then per game: int32 moves, int32 winner, per move: N[A], W[A], P[A] floats,
// root_N, root_W floats, int32 action, int32 traversals, int32 player.
#include "games/ConnectFourNode.hpp"
#include "games/GoNode.hpp"
#include "games/OthelloNode.hpp"
#include "symmetry/ConnectFourSymmetrizer.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"
#include "uct/UCTTree.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace SPRL;

template <class Node, class Sym, int B, int H, int A>
int run(char** argv) {
    using State = GridState<B, H>;
    deviceOptions().seed = std::strtoull(argv[2], 0, 10);
    deviceOptions().firstGame = std::strtoull(argv[3], 0, 10);
    const int nGames = std::atoi(argv[4]), sims = std::atoi(argv[5]), maxBatch = std::atoi(argv[6]), maxQueue = std::atoi(argv[7]);
    const float eps = (float)std::atof(argv[8]), alpha = (float)std::atof(argv[9]);
    const bool noise = std::atoi(argv[10]) != 0, useSym = std::atoi(argv[11]) != 0;
    const std::string q = argv[12];
    const InitQ initQ = q == "parent" ? InitQ::PARENT : (q == "drop" ? InitQ::DROP_PARENT : InitQ::ZERO);
    FILE* f = std::fopen(argv[13], "wb");
    if (!f) return 2;
    HashNetwork<State, A> net;
    Sym sym;
    int32_t hdr[2] = { nGames, A };
    std::fwrite(hdr, 4, 2, f);
    for (int g = 0; g < nGames; ++g) {
        UCTTree<Node, State, A> tree { std::make_unique<Node>(), eps, alpha, initQ, useSym ? &sym : nullptr, noise };
        std::vector<float> rows;
        std::vector<int32_t> ints;
        int moves = 0;
        while (!tree.getDecisionNode()->isTerminal()) {
            int trav = 0;
            while (trav < sims) {
                auto [leaves, n] = tree.searchAndGetLeaves(maxBatch, maxQueue, &net, 1.1f);
                if (!leaves.empty()) tree.evaluateAndBackpropLeaves(leaves, &net);
                trav += n;
            }
            const auto* root = tree.getDecisionNode();
            const auto* es = root->getEdgeStatistics();
            for (int a = 0; a < A; ++a) rows.push_back(es->m_numVisits[a]);
            for (int a = 0; a < A; ++a) rows.push_back(es->m_totalValues[a]);
            for (int a = 0; a < A; ++a) rows.push_back(es->m_childPriors[a]);
            rows.push_back(root->N());
            rows.push_back(root->W());
            const auto& visits = es->m_numVisits;
            const int action = (int)std::distance(visits.begin(), std::max_element(visits.begin(), visits.end()));
            ints.push_back(action); ints.push_back(trav); ints.push_back((int)root->getPlayer());
            tree.advanceDecision((ActionIdx)action);
            ++moves;
        }
        int32_t gh[2] = { moves, (int32_t)tree.getDecisionNode()->getWinner() };
        std::fwrite(gh, 4, 2, f);
        for (int m = 0; m < moves; ++m) {
            std::fwrite(rows.data() + (size_t)m * (3 * A + 2), 4, 3 * A + 2, f);
            std::fwrite(ints.data() + (size_t)m * 3, 4, 3, f);
        }
    }
    std::fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 14) { std::fprintf(stderr, "usage: see the header of treewalk.cpp\n"); return 2; }
    try {
        if (!std::strcmp(argv[1], "othello")) return run<OthelloNode, D4GridSymmetrizer<OTH_BOARD_WIDTH, OTH_HISTORY_SIZE>, OTH_BOARD_SIZE, OTH_HISTORY_SIZE, OTH_ACTION_SIZE>(argv);
        if (!std::strcmp(argv[1], "c4")) return run<ConnectFourNode, ConnectFourSymmetrizer, C4_BOARD_SIZE, C4_HISTORY_SIZE, C4_ACTION_SIZE>(argv);
        if (!std::strcmp(argv[1], "go")) return run<GoNode, D4GridSymmetrizer<GO_BOARD_WIDTH, GO_HISTORY_SIZE>, GO_BOARD_SIZE, GO_HISTORY_SIZE, GO_ACTION_SIZE>(argv);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "treewalk: %s\n", e.what());
        return 3;
    }
    return 2;
}
