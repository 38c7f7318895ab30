// networks/OthelloHeuristic.hpp -- include-path compatibility only.  The reference's
// OTHWorker.cpp includes this header but its use is commented out (OTHWorker.cpp:50);
// the heuristic evaluator is interactive-play tooling and is out of the hot path's scope.
#ifndef SPRL_B200_COMPAT_OTHELLO_HEURISTIC_HPP
#define SPRL_B200_COMPAT_OTHELLO_HEURISTIC_HPP
#include "../sprl/veneer.hpp"
#endif
