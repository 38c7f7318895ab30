// GoWorker -- GPU drop-in for the reference's Go worker (cpp/src/GoWorker.cpp).
// Defaults are the reference's constants (GoWorker.cpp:11-27); build with
// -DSPRL_GO_BOARD_WIDTH=9 for the 9x9 / komi 7.5 rules.
#include "games/GoNode.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"
#include "worker_main.hpp"

int main(int argc, char* argv[]) {
    const SPRL::WorkerDefaults d = { "panda_alpha", 4, 384, 100, 3, 262144, 1, 1, 3, 32768, 16, 8, 0.25f, 0.2f };
    return SPRL::workerMain<SPRL::GoNode, SPRL::D4GridSymmetrizer<SPRL::GO_BOARD_WIDTH, SPRL::GO_HISTORY_SIZE>,
                            SPRL::GO_BOARD_WIDTH, SPRL::GO_BOARD_WIDTH, SPRL::GO_HISTORY_SIZE, SPRL::GO_ACTION_SIZE>(
        argc, argv, d, "GoWorker");
}
