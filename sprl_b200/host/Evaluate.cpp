// Evaluate -- GPU drop-in for the reference's match executable (cpp/src/Evaluate.cpp): same command line,
//   ./Evaluate <modelPath0> <modelPath1> <numGames> <numTraversals> <maxBatchSize> <maxQueueSize>
//              <model0UseSymmetrize> <model0UseParentQ> <model1UseSymmetrize> <model1UseParentQ>
// a model path of "random" selects the uniform evaluator (Evaluate.cpp:70-84), "heuristic" the Othello heuristic
// the reference keeps commented out (:62-66).  The reference hard-wires Connect Four (:30-35); SPRL_GAME =
// c4 | othello | go selects the game here.  All games are played concurrently (SPRL_NUM_SLOTS pairs of trees at a
// time); the final tally line is the reference's progress-bar text (:155).
#include "games/ConnectFourNode.hpp"
#include "games/GoNode.hpp"
#include "games/OthelloNode.hpp"
#include "networks/GridNetwork.hpp"
#include "networks/OthelloHeuristic.hpp"
#include "networks/RandomNetwork.hpp"
#include "symmetry/ConnectFourSymmetrizer.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>

namespace {

template <typename ImplNode, typename Symmetrizer, int ROWS, int COLS, int HISTORY, int ACTIONS>
int evaluateMain(char* argv[]) {
    using State = SPRL::GridState<ROWS * COLS, HISTORY>;
    using Net = SPRL::INetwork<State, ACTIONS>;
    const std::string modelPath[2] = { argv[1], argv[2] };
    const int numGames = std::stoi(argv[3]), numTraversals = std::stoi(argv[4]);
    const int maxBatchSize = std::stoi(argv[5]), maxQueueSize = std::stoi(argv[6]);
    const bool useSym[2] = { std::stoi(argv[7]) > 0, std::stoi(argv[9]) > 0 };
    const bool useParentQ[2] = { std::stoi(argv[8]) > 0, std::stoi(argv[10]) > 0 };

    SPRL::RandomNetwork<State, ACTIONS> randomNetwork {};
    Symmetrizer symmetrizer {};
    std::unique_ptr<Net> owned[2];
    Net* network[2];
    for (int k = 0; k < 2; ++k) {
        if (modelPath[k] == "random") {
            std::cout << "Using random network..." << std::endl;
            network[k] = &randomNetwork;
        } else if constexpr (std::is_same_v<ImplNode, SPRL::OthelloNode>) {
            if (modelPath[k] == "heuristic") {
                std::cout << "Using the Othello heuristic..." << std::endl;
                owned[k] = std::make_unique<SPRL::OthelloHeuristic>();
                network[k] = owned[k].get();
                continue;
            }
        }
        if (modelPath[k] != "random") {
            std::cout << "Using traced PyTorch network..." << std::endl;
            owned[k] = std::make_unique<SPRL::GridNetwork<ROWS, COLS, HISTORY, ACTIONS>>(modelPath[k]);
            network[k] = owned[k].get();
        }
    }
    SPRL::DeviceOptions& opt = SPRL::deviceOptions();
    const char* v;
    if ((v = std::getenv("SPRL_DEVICE"))) opt.device = std::atoi(v);
    if ((v = std::getenv("SPRL_SEED"))) opt.seed = (uint64_t)std::atoll(v);
    if ((v = std::getenv("SPRL_NUM_SLOTS"))) opt.numSlots = std::atoi(v);
    try {
        SPRL::MatchResult r = SPRL::playMatch<ImplNode, State, ACTIONS>(
            network[0], network[1], numGames, numTraversals, maxBatchSize, maxQueueSize,
            useSym[0] ? &symmetrizer : nullptr, useParentQ[0] ? SPRL::InitQ::PARENT : SPRL::InitQ::ZERO,
            useSym[1] ? &symmetrizer : nullptr, useParentQ[1] ? SPRL::InitQ::PARENT : SPRL::InitQ::ZERO);
        std::cout << "Player 0 wins: " << r.wins0 << ", Player 1 wins: " << r.wins1 << ", Draws: " << r.draws << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "Evaluate: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc != 11) {
        std::cerr << "Usage: ./Evaluate.exe <modelPath0> <modelPath1> <numGames> <numTraversals> <maxBatchSize> <maxQueueSize> "
                     "<model0UseSymmetrize> <model0UseParentQ> <model1UseSymmetrize> <model1UseParentQ>" << std::endl;
        return 1;
    }
    const char* g = std::getenv("SPRL_GAME");
    const std::string game = g ? g : "c4";
    if (game == "othello")
        return evaluateMain<SPRL::OthelloNode, SPRL::D4GridSymmetrizer<SPRL::OTH_BOARD_WIDTH, SPRL::OTH_HISTORY_SIZE>,
                            SPRL::OTH_BOARD_WIDTH, SPRL::OTH_BOARD_WIDTH, SPRL::OTH_HISTORY_SIZE, SPRL::OTH_ACTION_SIZE>(argv);
    if (game == "go")
        return evaluateMain<SPRL::GoNode, SPRL::D4GridSymmetrizer<SPRL::GO_BOARD_WIDTH, SPRL::GO_HISTORY_SIZE>,
                            SPRL::GO_BOARD_WIDTH, SPRL::GO_BOARD_WIDTH, SPRL::GO_HISTORY_SIZE, SPRL::GO_ACTION_SIZE>(argv);
    return evaluateMain<SPRL::ConnectFourNode, SPRL::ConnectFourSymmetrizer, SPRL::C4_NUM_ROWS, SPRL::C4_NUM_COLS,
                        SPRL::C4_HISTORY_SIZE, SPRL::C4_ACTION_SIZE>(argv);
}
