// OTHWorker -- GPU drop-in for the reference's Othello worker (cpp/src/OTHWorker.cpp).
// Defaults are the reference's constants (OTHWorker.cpp:12-28).
#include "games/OthelloNode.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"
#include "worker_main.hpp"

int main(int argc, char* argv[]) {
    const SPRL::WorkerDefaults d = { "orangutan_alpha", 4, 384, 50, 3, 131072, 1, 1, 3, 8192, 8, 4, 0.25f, 0.3f };
    return SPRL::workerMain<SPRL::OthelloNode, SPRL::D4GridSymmetrizer<SPRL::OTH_BOARD_WIDTH, SPRL::OTH_HISTORY_SIZE>,
                            SPRL::OTH_BOARD_WIDTH, SPRL::OTH_BOARD_WIDTH, SPRL::OTH_HISTORY_SIZE, SPRL::OTH_ACTION_SIZE>(
        argc, argv, d, "OTHWorker");
}
