// networks/OthelloHeuristic.hpp -- drop-in for the reference's SPRL::OthelloHeuristic
// (/root/reference/cpp/src/networks/OthelloHeuristic.hpp:16-39, OthelloHeuristic.cpp:5-53): uniform policy over
// the legal actions, value = (own legal actions - opponent's placements) / empty squares.  Evaluated on the device
// inside the search kernel (csrc/search.cu: evaluate_leaf, SPRL_EVAL_OTHELLO_HEURISTIC); this class only names it.
#ifndef SPRL_B200_OTHELLO_HEURISTIC_HPP
#define SPRL_B200_OTHELLO_HEURISTIC_HPP
#include "../sprl/veneer.hpp"

namespace SPRL {

class OthelloHeuristic : public INetwork<GridState<OTH_BOARD_SIZE, OTH_HISTORY_SIZE>, OTH_ACTION_SIZE> {
public:
    using State = GridState<OTH_BOARD_SIZE, OTH_HISTORY_SIZE>;
    OthelloHeuristic() = default;
    int evaluatorKind() const override { return SPRL_EVAL_OTHELLO_HEURISTIC; }
};

}  // namespace SPRL
#endif
