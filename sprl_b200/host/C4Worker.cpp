// C4Worker -- GPU drop-in for the reference's Connect Four worker (cpp/src/C4Worker.cpp).
// Defaults are the reference's constants (C4Worker.cpp:11-27).
#include "games/ConnectFourNode.hpp"
#include "symmetry/ConnectFourSymmetrizer.hpp"
#include "worker_main.hpp"

int main(int argc, char* argv[]) {
    const SPRL::WorkerDefaults d = { "c4_test", 1, 1, 25, 10, 2048, 1, 1, 5, 512, 8, 4, 0.25f, 0.5f };
    return SPRL::workerMain<SPRL::ConnectFourNode, SPRL::ConnectFourSymmetrizer, SPRL::C4_NUM_ROWS, SPRL::C4_NUM_COLS,
                            SPRL::C4_HISTORY_SIZE, SPRL::C4_ACTION_SIZE>(argc, argv, d, "C4Worker");
}
